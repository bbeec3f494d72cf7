// Shared helpers for the libafsl kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "afsl.h"

namespace afsl {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

// thread-local error string + launch counter (afsl_abi.cu)
void set_error(const char* fmt, ...);
void count_launch();

#define AFSL_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::afsl::set_error(__VA_ARGS__);      \
      return AFSL_EINVAL;                  \
    }                                      \
  } while (0)

#define AFSL_CHECK_LAUNCH(name)                                                   \
  do {                                                                            \
    cudaError_t err__ = cudaGetLastError();                                       \
    if (err__ != cudaSuccess) {                                                   \
      ::afsl::set_error("%s: launch failed: %s", name, cudaGetErrorString(err__)); \
      return AFSL_ECUDA;                                                          \
    }                                                                             \
    ::afsl::count_launch();                                                       \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------- device side
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sum over aligned groups of `width` lanes (width a power of two <= 32); all lanes get the result
template <int kWidth>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = kWidth / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// packed fp32x2 arithmetic (Blackwell FADD2 / FFMA2): two IEEE fp32 operations per issued instruction,
// bit-identical to the scalar sub / fma they replace
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float sum2(f32x2 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return lo + hi;
}

// streaming 128-bit global accesses: data touched once, keep it out of L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}


// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier: global -> shared without register staging
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make barrier initialisation / prior generic-proxy accesses visible to the async proxy
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)       // suspend-time hint: the waiting thread sleeps in hardware until the phase flips
      : "memory");                      // instead of re-issuing try_wait every ~9 clocks (the spinning issuer warp of the
                                        // tensor-core head was 10 % of the kernel's instructions, all on one scheduler)
}

// Stable bucket of rows by label, whole CTA cooperating: on return (after the trailing barrier)
// row[start[w] .. start[w]+cnt[w]) lists the rows whose label is w in ascending row order - the order
// torch.nonzero / torch.where yield in the reference.  Labels outside 0..W-1 are left out.
__device__ inline void bucket_by_label(const int32_t* __restrict__ labels, int n, int W, int* lab, int* row, int* cnt,
                                       int* start) {
  for (int k = threadIdx.x; k < n; k += blockDim.x) lab[k] = labels[k];
  __syncthreads();
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    int c = 0;
    for (int k = 0; k < n; ++k) c += (lab[k] == w);
    cnt[w] = c;
  }
  __syncthreads();
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    int st = 0;
    for (int v = 0; v < w; ++v) st += cnt[v];
    start[w] = st;
    int j = st;
    for (int k = 0; k < n; ++k)
      if (lab[k] == w) row[j++] = k;
  }
  __syncthreads();
}

// persistent grid: one CTA per episode up to (SMs x resident CTAs per SM)
template <typename Kernel>
inline int persistent_grid(Kernel fn, int threads, size_t smem_bytes, int work_items) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem_bytes);
  if (per_sm < 1) per_sm = 1;
  int sms = kNumSMs, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int cap = sms * per_sm;
  return work_items < cap ? work_items : cap;
}

template <typename Kernel>
inline int opt_in_smem(Kernel fn, size_t bytes, const char* name) {
  if (bytes > 226 * 1024) {
    set_error("%s: needs %zu B of shared memory per CTA (limit 226 KB)", name, bytes);
    return AFSL_EINVAL;
  }
  if (bytes > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (err != cudaSuccess) {
      set_error("%s: cannot opt in to %zu B shared memory: %s", name, bytes, cudaGetErrorString(err));
      return AFSL_ECUDA;
    }
  }
  return AFSL_OK;
}

}  // namespace afsl
