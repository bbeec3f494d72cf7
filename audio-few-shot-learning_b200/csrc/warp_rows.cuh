// Warp-per-episode building blocks shared by proto_head_warp.cu and cpl_warp.cu: a lane owns D/32 columns of
// every row (packed fp32 pairs), rows move with coalesced 128-bit accesses, and per-lane partial sums of a
// batch of values are combined with one transposing butterfly.
#pragma once

#include "afsl_common.cuh"

namespace afsl {
namespace warp_rows {

constexpr int kWarpsPerCta = 4;
constexpr int kCtaThreads = kWarpsPerCta * kWarp;
constexpr unsigned kFull = 0xffffffffu;

// ---- 64-bit (two packed fp32) global accesses, streaming
__device__ __forceinline__ void ldg2(const float* p, f32x2& a, f32x2& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}
__device__ __forceinline__ f32x2 ldg1(const float* p) {
  f32x2 a;
  asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(a) : "l"(p));
  return a;
}
__device__ __forceinline__ void stg2(float* p, f32x2 a, f32x2 b) {
  asm volatile("st.global.L1::no_allocate.v2.b64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void stg1(float* p, f32x2 a) {
  asm volatile("st.global.L1::no_allocate.b64 [%0], %1;" ::"l"(p), "l"(a) : "memory");
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// a lane's kV = D/32 floats of one row as kV/2 packed pairs: float4 chunk c of the lane sits at float4 index c*32+lane
template <int kV>
__device__ __forceinline__ void load_row(const float* row, int lane, f32x2 (&v)[kV / 2]) {
  if constexpr (kV == 2) {
    v[0] = ldg1(row + 2 * lane);
  } else {
#pragma unroll
    for (int c = 0; c < kV / 4; ++c) ldg2(row + 4 * (c * 32 + lane), v[2 * c], v[2 * c + 1]);
  }
}
template <int kV>
__device__ __forceinline__ void store_row(float* row, int lane, const f32x2 (&v)[kV / 2]) {
  if constexpr (kV == 2) {
    stg1(row + 2 * lane, v[0]);
  } else {
#pragma unroll
    for (int c = 0; c < kV / 4; ++c) stg2(row + 4 * (c * 32 + lane), v[2 * c], v[2 * c + 1]);
  }
}

// the same lane layout in shared memory (conflict-free 128-bit accesses)
template <int kV>
__device__ __forceinline__ void lds_row(const float* row, int lane, f32x2 (&v)[kV / 2]) {
  if constexpr (kV == 2) {
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v[0]) : "r"(smem_u32(row + 2 * lane)));
  } else {
#pragma unroll
    for (int c = 0; c < kV / 4; ++c)
      asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(v[2 * c]), "=l"(v[2 * c + 1]) : "r"(smem_u32(row + 4 * (c * 32 + lane))));
  }
}
template <int kV>
__device__ __forceinline__ void sts_row(float* row, int lane, const f32x2 (&v)[kV / 2]) {
  if constexpr (kV == 2) {
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(smem_u32(row + 2 * lane)), "l"(v[0]) : "memory");
  } else {
#pragma unroll
    for (int c = 0; c < kV / 4; ++c)
      asm volatile("st.shared.v2.b64 [%0], {%1,%2};" ::"r"(smem_u32(row + 4 * (c * 32 + lane))), "l"(v[2 * c]), "l"(v[2 * c + 1]) : "memory");
  }
}

__host__ __device__ constexpr int pow2_ceil(int n) { int p = 1; while (p < n) p <<= 1; return p; }

// Transposing butterfly: every lane enters with N partial values; lane L leaves with the warp-wide sum of value
// L % N.  Each halving step keeps the half of the values whose index bit matches the lane's bit and ships the
// other half to the partner lane: N-1 shuffles for N <= 32 values, plus plain butterflies when N < 32.
template <int N>
__device__ __forceinline__ float transpose_reduce(float (&v)[N], int lane) {
  if constexpr (N == 1) {
    float t = v[0];
    return t;
  } else {
    constexpr int H = N / 2;
    const bool up = (lane & H) != 0;
    float k[H];
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const float keep = up ? v[i + H] : v[i];
      const float send = up ? v[i] : v[i + H];
      k[i] = keep + __shfl_xor_sync(kFull, send, H);
    }
    return transpose_reduce<H>(k, lane);
  }
}
template <int N>
__device__ __forceinline__ float reduce_values(float (&v)[N], int lane) {
  float t = transpose_reduce<N>(v, lane);
#pragma unroll
  for (int o = N; o < 32; o <<= 1) t += __shfl_xor_sync(kFull, t, o);
  return t;
}

}  // namespace warp_rows
}  // namespace afsl
