// Row L2 normalisation y = x / max(||x||_2, eps) and its backward.
// Reference: F.normalize on the prototypes, loops/loops.py:47-48 (eps 1e-12); the same formula
// closes ProjectionHead (models/main_modules.py:253).  One warp per row, 128-bit accesses.
#include "afsl_common.cuh"

namespace afsl {
namespace {
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               int rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int D4 = D >> 2;
  for (int r = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); r < rows; r += gridDim.x * (kThreads / 32)) {
    const float4* x4 = reinterpret_cast<const float4*>(x) + (size_t)r * D4;
    float4* y4 = reinterpret_cast<float4*>(y) + (size_t)r * D4;
    float ss = 0.f;
    for (int c = lane; c < D4; c += 32) {
      const float4 v = __ldg(x4 + c);
      ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), eps);
    for (int c = lane; c < D4; c += 32) {
      float4 v = __ldg(x4 + c);
      v.x = __fdiv_rn(v.x, denom); v.y = __fdiv_rn(v.y, denom); v.z = __fdiv_rn(v.z, denom); v.w = __fdiv_rn(v.w, denom);
      y4[c] = v;
    }
  }
}

// d_x = (d_y - y (y . d_y)) / ||x||  when ||x|| > eps, else d_y / eps
__global__ void __launch_bounds__(kThreads) l2norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               float* __restrict__ dx, int rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int D4 = D >> 2;
  for (int r = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); r < rows; r += gridDim.x * (kThreads / 32)) {
    const float4* x4 = reinterpret_cast<const float4*>(x) + (size_t)r * D4;
    const float4* g4 = reinterpret_cast<const float4*>(dy) + (size_t)r * D4;
    float4* o4 = reinterpret_cast<float4*>(dx) + (size_t)r * D4;
    float ss = 0.f, dot = 0.f;
    for (int c = lane; c < D4; c += 32) {
      const float4 v = __ldg(x4 + c), g = __ldg(g4 + c);
      ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
      dot = fmaf(v.x, g.x, dot); dot = fmaf(v.y, g.y, dot); dot = fmaf(v.z, g.z, dot); dot = fmaf(v.w, g.w, dot);
    }
    ss = warp_sum(ss);
    dot = warp_sum(dot);
    const float nrm = sqrtf(ss);
    const bool clamped = !(nrm > eps);
    const float inv = 1.f / (clamped ? eps : nrm);
    const float k = clamped ? 0.f : dot * inv * inv * inv;   // (x . dy) / ||x||^3
    for (int c = lane; c < D4; c += 32) {
      const float4 v = __ldg(x4 + c), g = __ldg(g4 + c);
      float4 o;
      o.x = fmaf(-k, v.x, g.x * inv); o.y = fmaf(-k, v.y, g.y * inv);
      o.z = fmaf(-k, v.z, g.z * inv); o.w = fmaf(-k, v.w, g.w * inv);
      o4[c] = o;
    }
  }
}

int grid_for(int rows) {
  const int blocks = (rows + kThreads / 32 - 1) / (kThreads / 32);
  const int cap = kNumSMs * 8;
  return blocks < cap ? blocks : cap;
}
}  // namespace
}  // namespace afsl

extern "C" int afsl_l2_normalize_fwd_f32(const float* x, float* y, int rows, int D, float eps, void* stream) {
  AFSL_REQUIRE(x && y, "afsl_l2_normalize_fwd_f32: null pointer");
  AFSL_REQUIRE(rows >= 0 && D > 0 && D % 4 == 0, "afsl_l2_normalize_fwd_f32: rows=%d D=%d (D must be a multiple of 4)", rows, D);
  if (rows == 0) return AFSL_OK;
  afsl::l2norm_fwd_kernel<<<afsl::grid_for(rows), afsl::kThreads, 0, (cudaStream_t)stream>>>(x, y, rows, D, eps);
  AFSL_CHECK_LAUNCH("afsl_l2_normalize_fwd_f32");
  return AFSL_OK;
}

extern "C" int afsl_l2_normalize_bwd_f32(const float* x, const float* d_y, float* d_x, int rows, int D, float eps,
                                          void* stream) {
  AFSL_REQUIRE(x && d_y && d_x, "afsl_l2_normalize_bwd_f32: null pointer");
  AFSL_REQUIRE(rows >= 0 && D > 0 && D % 4 == 0, "afsl_l2_normalize_bwd_f32: rows=%d D=%d (D must be a multiple of 4)", rows, D);
  if (rows == 0) return AFSL_OK;
  afsl::l2norm_bwd_kernel<<<afsl::grid_for(rows), afsl::kThreads, 0, (cudaStream_t)stream>>>(x, d_y, d_x, rows, D, eps);
  AFSL_CHECK_LAUNCH("afsl_l2_normalize_bwd_f32");
  return AFSL_OK;
}
