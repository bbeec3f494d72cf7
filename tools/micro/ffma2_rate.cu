// Microbenchmark: fp32 FMA issue rates on sm_100a - scalar FFMA, packed FFMA2 (reg pairs), FFMA2 with a scalar (.F32)
// multiplicand - in FMA lanes per clock per SM.  build ON THE GPU BOX into /tmp with the shared runtime (a statically linked
// cudart in the shipped tree trips the driver's closed-call scan): nvcc -arch=sm_100a -O3 -cudart shared -o /tmp/ffma2_rate ffma2_rate.cu && /tmp/ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
constexpr int kIters = 4096, kAcc = 12;
template <int kMode>
__global__ void __launch_bounds__(256) k(float* out, float s, long long* cyc) {
  float a[kAcc]; f32x2 a2[kAcc];
  for (int i = 0; i < kAcc; ++i) { a[i] = threadIdx.x + i; a2[i] = pack2(a[i], a[i] + 1.f); }
  const float m = s * 0.999f, n = s * 1e-3f;
  const f32x2 m2 = pack2(m, m * 1.0001f), n2 = pack2(n, n);
  const f32x2 ms = pack2(m, m);
  long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kAcc; ++i) {
      if (kMode == 0) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(a[i]) : "f"(a[i]), "f"(m), "f"(n));
      else if (kMode == 1) a2[i] = fma2(a2[i], m2, n2);
      else a2[i] = fma2(a2[i], ms, n2);
    }
  }
  long long t1 = clock64();
  float r = 0.f;
  for (int i = 0; i < kAcc; ++i) { r += a[i]; r += (float)(a2[i] & 0xffff); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&cyc, 148 * 8 * 8);
  const char* names[3] = {"FFMA  (3 regs)", "FFMA2 (pairs)", "FFMA2 (lo==hi multiplicand)"};
  for (int warps_per_sm : {8, 16, 32}) {
    for (int mode = 0; mode < 3; ++mode) {
      const int ctas = 148 * warps_per_sm / 8;
      if (mode == 0) k<0><<<ctas, 256>>>(out, 1.f, cyc); else if (mode == 1) k<1><<<ctas, 256>>>(out, 1.f, cyc); else k<2><<<ctas, 256>>>(out, 1.f, cyc);
      cudaDeviceSynchronize();
      long long h[148 * 8]; cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
      const double fma_per_thread = (double)kIters * kAcc * (mode == 0 ? 1 : 2);
      printf("%-30s warps/SM %2d: %.1f FMA lanes/clk/SM (%.0f cycles)\n", names[mode], warps_per_sm, fma_per_thread * warps_per_sm * 32 / avg, avg);
    }
  }
  return 0;
}
