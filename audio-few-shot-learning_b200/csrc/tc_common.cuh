// tcgen05 / TMEM / TMA building blocks shared by the tensor-core kernels of libafsl (sm_100a only).
//
// Conventions (checked on B200 by tools/micro/tf32x3_probe.cu):
//  * operand tiles are K-major, 32 fp32 (128 bytes) per row, 8-row atoms of 1 KB, SWIZZLE_128B: the 16-byte chunk c of
//    row r sits at chunk c ^ (r & 7) - the layout TMA writes with CU_TENSOR_MAP_SWIZZLE_128B and UMMA reads with layout
//    type 2; tile bases are 1 KB aligned; a K step of 8 fp32 advances the descriptor's start address by 32 bytes;
//  * kind::tf32 reads fp32 containers and ignores the low 13 mantissa bits, fp32 accumulation in TMEM;
//  * accumulator of an M = 128 MMA: TMEM lane = row, column = n; warp w may only touch lanes 32 (w % 4) .. + 31.
#pragma once

#include <cuda.h>

#include "afsl_common.cuh"

namespace afsl {
namespace tc {

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrives on the barrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// whole warp: allocate / free kCols TMEM columns (power of two >= 32); the base address lands in *dst (shared memory)
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_free(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(kCols) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address / 16, stride between
// 8-row atoms 1 KB, descriptor version 1 (sm_100), layout type 2
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, TF32 x TF32, both K-major, M x N
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same MMA with the descriptors given as (low word, shared high word): the issue loop of a stage then only adds
// constants to two 32-bit values per MMA (the high word - stride, version, swizzle - is the same for every tile).
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return (addr & 0x3FFFFu) >> 4; }
__device__ __forceinline__ void mma_tf32_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.b32 p, %5, 0;\n mov.b64 da, {%1, %3};\n mov.b64 db, {%2, %3};\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (the others skip the guarded statement): keeps the issuing warp's control flow uniform
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same loads without their wait: several can be in flight before one tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- split precision: x = hi + lo with hi the TF32 the tensor core itself sees (low 13 mantissa bits dropped) or the
//      nearest TF32, lo rounded to nearest TF32; hi.hi + hi.lo + lo.hi (+ lo.lo) then carries ~21 bits per product
__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ float trunc_tf32(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
// lo for an operand the TENSOR CORE reads: x - trunc(x) is exact; adding half a TF32 ulp to its bit pattern and letting the
// MMA drop the low 13 bits is cvt.rna.tf32 (which ptxas expands to compare + add + mask: six instructions per element
// with the subtraction and the mask, against three here), bit for bit the same operand
__device__ __forceinline__ float lo_bits(float x) { return __uint_as_float(__float_as_uint(x - trunc_tf32(x)) + 0x1000u); }
__device__ __forceinline__ float4 lo_of_raw(const float4& v) {      // hi = the raw value as the MMA truncates it
  return make_float4(lo_bits(v.x), lo_bits(v.y), lo_bits(v.z), lo_bits(v.w));
}
__device__ __forceinline__ void split_rna(const float4& v, float4& hi, float4& lo) {
  hi = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
  lo = make_float4(rna_tf32(v.x - hi.x), rna_tf32(v.y - hi.y), rna_tf32(v.z - hi.z), rna_tf32(v.w - hi.w));
}

__device__ __forceinline__ void sts4(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a swizzled [rows x 32 fp32] tile
__device__ __forceinline__ uint32_t sw128(int row, int chunk) {
  return ((uint32_t)row << 7) | ((uint32_t)((chunk ^ row) & 7) << 4);      // (row / 8) KB + (row % 8) x 128 B = 128 row
}

// ---- TMA: 2-D tiled bulk tensor copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}
// the same box into L2 only: a later tma_load_2d of it then pays the L2 latency instead of the HBM round trip
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y)
               : "memory");
}
// `bytes` contiguous bytes (a multiple of 16, 16-byte aligned) into L2
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// host: tensor map of a dense row-major fp32 matrix [rows x cols], box = box_rows x 32 columns, 128-byte swizzle.
// cuTensorMapEncodeTiled is fetched through the runtime (no link-time dependency on libcuda).
inline int make_tensor_map_f32(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                               const char* name) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) fn = nullptr;
    return reinterpret_cast<EncodeFn>(fn);
  }();
  if (!encode) {
    set_error("%s: cuTensorMapEncodeTiled is not available from this driver", name);
    return AFSL_ECUDA;
  }
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {cols * sizeof(float)};
  const cuuint32_t box[2] = {32u, box_rows};
  const cuuint32_t estride[2] = {1u, 1u};
  const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estride,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed (code %d) for a [%llu x %llu] matrix, box %u rows", name, (int)rc,
              (unsigned long long)rows, (unsigned long long)cols, box_rows);
    return AFSL_ECUDA;
  }
  return AFSL_OK;
}

}  // namespace tc
}  // namespace afsl
