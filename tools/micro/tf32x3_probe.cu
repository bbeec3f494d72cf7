// Probe: one CTA, D[128 x 32] = A[128 x K] * B[32 x K]^T on tcgen05 (kind::tf32), operands written to shared memory by
// ordinary stores in the K-major SWIZZLE_128B layout, single pass and 3-pass split precision (hi*hi + hi*lo + lo*hi),
// accumulator read back from TMEM with tcgen05.ld.  Checks every layout assumption the libafsl tensor-core kernels make
// (smem descriptor fields, the 128-byte swizzle applied by hand, K advance inside the swizzle atom, TMEM lane = row).
// Build ON THE GPU BOX (shared runtime, outside the tree):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/tf32x3_probe tools/micro/tf32x3_probe.cu && /tmp/tf32x3_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int kM = 128, kN = 32, kK = 64;        // two k-blocks of 32 fp32 (128 bytes)
constexpr int kBlockK = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row, col) inside one k-block tile [rows x 32 fp32], K-major, 128-byte swizzle, 8-row atoms of 1 KB
__host__ __device__ inline uint32_t sw128_offset(int row, int col) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 2) ^ (row & 7)) & 7) << 4) + (col & 3) * 4);
}

// round-to-nearest TF32 (10 explicit mantissa bits) kept in an fp32 container: v = hi + lo exactly, |lo| <= 2^-11 |v|
__device__ __forceinline__ float split_hi(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, 16-byte units
  d |= (uint64_t)(0) << 16;                             // leading byte offset (unused: one swizzle atom along K)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset: 8-row atoms 1 KB apart
  d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d1,
                                              float* __restrict__ d3) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // tiles: [k-block][hi/lo] for A (16 KB each) then B (4 KB each)
  uint8_t* a_tiles = smem;                                  // 2 k-blocks x 2 (hi, lo) x 16 KB
  uint8_t* b_tiles = smem + 2 * 2 * 16384;                  // 2 k-blocks x 2 x 4 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kM * kK; i += 128) {
    const int r = i / kK, c = i % kK, kb = c / kBlockK, cc = c % kBlockK;
    const float v = a[i];
    const float hi = split_hi(v);
    *reinterpret_cast<float*>(a_tiles + (kb * 2 + 0) * 16384 + sw128_offset(r, cc)) = hi;
    *reinterpret_cast<float*>(a_tiles + (kb * 2 + 1) * 16384 + sw128_offset(r, cc)) = split_hi(v - hi);
  }
  for (int i = tid; i < kN * kK; i += 128) {
    const int r = i / kK, c = i % kK, kb = c / kBlockK, cc = c % kBlockK;
    const float v = b[i];
    const float hi = split_hi(v);
    *reinterpret_cast<float*>(b_tiles + (kb * 2 + 0) * 4096 + sw128_offset(r, cc)) = hi;
    *reinterpret_cast<float*>(b_tiles + (kb * 2 + 1) * 4096 + sw128_offset(r, cc)) = split_hi(v - hi);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
  if (tid == 0) {
    // accumulator 0 (columns 0..31): single pass hi*hi; accumulator 1 (columns 32..63): three passes, small terms first
    for (int acc = 0; acc < 2; ++acc) {
      uint32_t first = 1;
      for (int pass = 0; pass < (acc ? 3 : 1); ++pass) {
        const int sa = acc && pass == 0 ? 1 : 0, sb = acc && pass == 1 ? 1 : 0;     // (lo,hi), (hi,lo), (hi,hi)
        for (int kb = 0; kb < kK / kBlockK; ++kb)
          for (int k = 0; k < kBlockK / 8; ++k) {
            const uint64_t ad = make_desc(smem_u32(a_tiles + (kb * 2 + sa) * 16384) + k * 32);
            const uint64_t bd = make_desc(smem_u32(b_tiles + (kb * 2 + sb) * 4096) + k * 32);
            const uint32_t accum = first ? 0u : 1u;
            first = 0;
            asm volatile(
                "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem + acc * 32),
                "l"(ad), "l"(bd), "r"(idesc), "r"(accum)
                : "memory");
          }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMAs
  asm volatile(
      "{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}" ::"r"(
          smem_u32(&bar))
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int acc = 0; acc < 2; ++acc) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + acc * 32;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float* out = (acc ? d3 : d1) + (size_t)tid * kN;       // thread t of warp w holds row 32 w + t
    for (int j = 0; j < 32; ++j) out[j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
  const int na = kM * kK, nb = kN * kK, nd = kM * kN;
  float *ha = (float*)malloc(na * 4), *hb = (float*)malloc(nb * 4), *h1 = (float*)malloc(nd * 4), *h3 = (float*)malloc(nd * 4);
  srand(1);
  for (int i = 0; i < na; ++i) ha[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (int i = 0; i < nb; ++i) hb[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *da, *db, *d1, *d3;
  cudaMalloc(&da, na * 4); cudaMalloc(&db, nb * 4); cudaMalloc(&d1, nd * 4); cudaMalloc(&d3, nd * 4);
  cudaMemcpy(da, ha, na * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb, nb * 4, cudaMemcpyHostToDevice);
  const int smem = 2 * 2 * 16384 + 2 * 2 * 4096 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(da, db, d1, d3);
  cudaError_t err = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(err));
  if (err != cudaSuccess) return 1;
  cudaMemcpy(h1, d1, nd * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h3, d3, nd * 4, cudaMemcpyDeviceToHost);
  double e1 = 0, e3 = 0, ef = 0;
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      double ref = 0;
      float f = 0.f;
      for (int k = 0; k < kK; ++k) { ref += (double)ha[m * kK + k] * hb[n * kK + k]; f = fmaf(ha[m * kK + k], hb[n * kK + k], f); }
      e1 = fmax(e1, fabs(h1[m * kN + n] - ref));
      e3 = fmax(e3, fabs(h3[m * kN + n] - ref));
      ef = fmax(ef, fabs((double)f - ref));
    }
  printf("max |err| vs fp64: single tf32 pass %.3e, three passes %.3e, fp32 fma chain %.3e  (|dot| ~ %.1f)\n", e1, e3, ef, sqrt(kK / 9.0));
  printf("sample d3[5][7] = %.7f\n", h3[5 * kN + 7]);
  printf(e3 < 5e-6 && e1 < 5e-2 ? "PROBE OK\n" : "PROBE MISMATCH\n");
  return 0;
}
