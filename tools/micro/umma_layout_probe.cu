// Probe: which shared-memory operand layouts tcgen05.mma.kind::tf32 accepts beyond the K-major tiles of tf32x3_probe.cu.
//   test 0  B operand MN-major, SWIZZLE_128B: B[n][k] is read from a row-major [K rows][N fp32] block stored as two
//           column halves of 32 fp32 (128-byte rows, 8-row swizzle atoms, halves 16 KB apart) - i.e. the SAME bytes a
//           K-major A operand X[row][d] occupies, consumed as X^T without a transposed copy.  D[128 x 64] = A[128 x 32] * B.
//           Two descriptor readings are tried: (LBO, SBO) = (16 KB, 1 KB) and (1 KB, 16 KB).
//   test 1  M = 64: where the 64 accumulator rows land in TMEM (all 128 lanes are dumped).
// Build ON THE GPU BOX: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/umma_probe tools/micro/umma_layout_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline uint32_t sw128_offset(int row, int col) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 2) ^ (row & 7)) & 7) << 4) + (col & 3) * 4);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma(uint32_t tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
               "l"(ad), "l"(bd), "r"(idesc), "r"(accum)
               : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, float* out) {
  uint32_t v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
      "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 32; ++j) out[j] = __uint_as_float(v[j]);
}

constexpr int kM = 128, kK = 32, kN = 64;

// a: [128][32] row-major (A, K-major); x: [32][64] row-major (B^T: x[k][n]); out: [3][128 lanes][64]
__global__ void __launch_bounds__(128) probe(const float* __restrict__ a, const float* __restrict__ x, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_tile = smem;                      // [128 rows x 128 B] K-major
  uint8_t* x_tiles = smem + 16384;             // two halves [128 rows x 128 B] 16 KB apart; rows 0..31 hold x
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 3 * 16384 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < kM * kK; i += 128) *reinterpret_cast<float*>(a_tile + sw128_offset(i / kK, i % kK)) = a[i];
  for (int i = tid; i < kK * kN; i += 128) {
    const int k = i / kN, n = i % kN;
    *reinterpret_cast<float*>(x_tiles + (n / 32) * 16384 + sw128_offset(k, n % 32)) = x[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // b_major = MN: bit 16
  const uint32_t idesc_mn = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
  const uint32_t idesc_m64 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
  if (tid == 0) {
    for (int variant = 0; variant < 2; ++variant)
      for (int k = 0; k < kK / 8; ++k) {       // a K step of 8: 32 bytes along A's rows, 8 rows (one 1 KB atom) down B
        const uint64_t ad = make_desc(smem_u32(a_tile) + k * 32, 0, 1024);
        const uint64_t bd = variant == 0 ? make_desc(smem_u32(x_tiles) + k * 1024, 16384, 1024)
                                         : make_desc(smem_u32(x_tiles) + k * 1024, 1024, 16384);
        mma(tmem + variant * 64, ad, bd, idesc_mn, k != 0);
      }
    // M = 64: D[64 x 32] = A[0:64] * (A[64:96])^T, both K-major
    for (int k = 0; k < kK / 8; ++k)
      mma(tmem + 128, make_desc(smem_u32(a_tile) + k * 32, 0, 1024), make_desc(smem_u32(a_tile) + 64 * 128 + k * 32, 0, 1024),
          idesc_m64, k != 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile(
      "{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}" ::"r"(
          smem_u32(&bar))
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int t = 0; t < 3; ++t)
    for (int h = 0; h < 2; ++h) {
      if (t == 2 && h == 1) continue;
      ld32(tmem + ((uint32_t)(warp * 32) << 16) + t * 64 + h * 32, out + ((size_t)t * 128 + tid) * 64 + h * 32);
    }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main() {
  const int na = kM * kK, nx = kK * kN, no = 3 * 128 * 64;
  float *ha = (float*)malloc(na * 4), *hx = (float*)malloc(nx * 4), *ho = (float*)calloc(no, 4);
  srand(3);
  // values exactly representable in TF32 so that a correct layout reproduces the fp32 result to rounding of the sum only
  for (int i = 0; i < na; ++i) ha[i] = (float)(rand() % 65 - 32) / 32.f;
  for (int i = 0; i < nx; ++i) hx[i] = (float)(rand() % 65 - 32) / 32.f;
  float *da, *dx, *dout;
  cudaMalloc(&da, na * 4); cudaMalloc(&dx, nx * 4); cudaMalloc(&dout, no * 4);
  cudaMemcpy(da, ha, na * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dx, hx, nx * 4, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0, no * 4);
  const int smem = 3 * 16384 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(da, dx, dout);
  cudaError_t err = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(err));
  if (err != cudaSuccess) return 1;
  cudaMemcpy(ho, dout, no * 4, cudaMemcpyDeviceToHost);
  for (int variant = 0; variant < 2; ++variant) {
    double emax = 0;
    for (int m = 0; m < kM; ++m)
      for (int n = 0; n < kN; ++n) {
        double ref = 0;
        for (int k = 0; k < kK; ++k) ref += (double)ha[m * kK + k] * hx[k * kN + n];
        emax = fmax(emax, fabs(ho[((size_t)variant * 128 + m) * 64 + n] - ref));
      }
    printf("MN-major B, (LBO, SBO) = %s: max |err| %.3e -> %s\n", variant == 0 ? "(16 KB, 1 KB)" : "(1 KB, 16 KB)", emax,
           emax < 1e-4 ? "LAYOUT OK" : "mismatch");
  }
  // M = 64: find, for each accumulator row r (0..63), the TMEM lane that holds it
  int lane_of[64];
  for (int r = 0; r < 64; ++r) {
    lane_of[r] = -1;
    for (int lane = 0; lane < 128 && lane_of[r] < 0; ++lane) {
      double emax = 0;
      for (int n = 0; n < 32; ++n) {
        double ref = 0;
        for (int k = 0; k < kK; ++k) ref += (double)ha[r * kK + k] * ha[(64 + n) * kK + k];
        emax = fmax(emax, fabs(ho[((size_t)2 * 128 + lane) * 64 + n] - ref));
      }
      if (emax < 1e-4) lane_of[r] = lane;
    }
  }
  printf("M = 64 accumulator: row -> TMEM lane:");
  for (int r = 0; r < 64; ++r) printf(" %d", lane_of[r]);
  printf("\n");
  return 0;
}
