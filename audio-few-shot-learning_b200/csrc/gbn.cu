// Grouped BatchNorm + ReLU + MaxPool(3, stride 3) for the encoder stages, forward and backward.
//
// Reference: conv_block = Conv3x3 -> BatchNorm2d -> ReLU -> MaxPool2d(3,3), models/main_modules.py:43-60,
// applied by EncoderModule once per view per set per episode (models/main_modules.py:18-23), i.e. in
// training mode every 25-sample set has its own batch statistics.  Batching E episodes into one
// cuDNN convolution call therefore needs BatchNorm statistics per GROUP of `group` consecutive
// samples.  The convolution stays cuDNN; these kernels replace the eager chain
// batch_norm -> relu -> max_pool2d (and its three backward kernels), which would otherwise move the
// full-resolution activation through HBM ~10 times per stage:
//   stats : one read of x                       -> mean, rstd, biased var per (group, channel)
//   fwd   : one read of x, pooled write         (ReLU and pooling fused; nothing saved but x)
//   bwd 1 : one read of x + dy                  -> sum(dz), sum(dz*xhat) per (group, channel)
//   bwd 2 : one read of x + dy, one write of dx (max-pool routing recomputed from x)
// x is [G*group, C, H, W] fp32 NCHW (cuDNN's default layout here); one CTA owns one (group, channel)
// slab at a time, so reductions need no atomics and are bit-reproducible.
#include "afsl_common.cuh"

namespace afsl {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / kWarp;

struct GbnParams {
  const float* x;       // [G*group, C, H, W]
  const float* mean;    // [G,C] (or [C] when stats_per_group == 0)
  const float* rstd;    // idem
  const float* gamma;   // [C]
  const float* beta;    // [C]
  const float* dy;      // [G*group, C, PH, PW]
  const float* sums;    // [G,C,2] from bwd 1
  float* y;             // [G*group, C, PH, PW]
  float* dx;            // like x
  float* out_mean;      // [G,C]
  float* out_rstd;      // [G,C]
  float* out_var;       // [G,C] biased
  float* out_sums;      // [G,C,2]
  float eps;
  int G, group, C, H, W, PH, PW, stats_per_group;
};

__device__ inline double block_sum_d(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int k = 0; k < kWarps; ++k) t += red[k];
  return t;
}

// ---------------------------------------------------------------- statistics
template <int kVec>
__device__ inline void slab_moments(const float* __restrict__ plane0, size_t plane_stride, int n_planes, int hw,
                                    float shift, float& s1, float& s2) {
  // shifted sums: sum(x - k), sum((x - k)^2) with k = first element of the slab (kills cancellation)
  const int nvec = hw / kVec;
  for (int i = 0; i < n_planes; ++i) {
    const float* pl = plane0 + (size_t)i * plane_stride;
    if (kVec == 4) {
      const float4* p4 = reinterpret_cast<const float4*>(pl);
      for (int v = threadIdx.x; v < nvec; v += kThreads) {
        const float4 a = ldg_stream(p4 + v);
        const float d0 = a.x - shift, d1 = a.y - shift, d2 = a.z - shift, d3 = a.w - shift;
        s1 += (d0 + d1) + (d2 + d3);
        s2 = fmaf(d0, d0, s2); s2 = fmaf(d1, d1, s2); s2 = fmaf(d2, d2, s2); s2 = fmaf(d3, d3, s2);
      }
    } else if (kVec == 2) {
      const float2* p2 = reinterpret_cast<const float2*>(pl);
      for (int v = threadIdx.x; v < nvec; v += kThreads) {
        const float2 a = __ldg(p2 + v);
        const float d0 = a.x - shift, d1 = a.y - shift;
        s1 += d0 + d1;
        s2 = fmaf(d0, d0, s2); s2 = fmaf(d1, d1, s2);
      }
    } else {
      for (int v = threadIdx.x; v < hw; v += kThreads) {
        const float d0 = __ldg(pl + v) - shift;
        s1 += d0;
        s2 = fmaf(d0, d0, s2);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads) gbn_stats_kernel(const GbnParams p) {
  __shared__ double red[kWarps];
  const int hw = p.H * p.W;
  const size_t plane_stride = (size_t)p.C * hw;
  for (int slab = blockIdx.x; slab < p.G * p.C; slab += gridDim.x) {
    const int g = slab / p.C, c = slab - g * p.C;
    const float* plane0 = p.x + ((size_t)g * p.group * p.C + c) * hw;
    const float shift = __ldg(plane0);
    float s1 = 0.f, s2 = 0.f;
    if ((hw & 3) == 0) slab_moments<4>(plane0, plane_stride, p.group, hw, shift, s1, s2);
    else if ((hw & 1) == 0) slab_moments<2>(plane0, plane_stride, p.group, hw, shift, s1, s2);
    else slab_moments<1>(plane0, plane_stride, p.group, hw, shift, s1, s2);
    const double t1 = block_sum_d((double)s1, red);
    const double t2 = block_sum_d((double)s2, red);
    if (threadIdx.x == 0) {
      const double m = (double)p.group * hw;
      const double dmean = t1 / m;
      double var = t2 / m - dmean * dmean;
      if (var < 0.0) var = 0.0;
      p.out_mean[slab] = (float)((double)shift + dmean);
      p.out_var[slab] = (float)var;
      p.out_rstd[slab] = (float)(1.0 / sqrt(var + (double)p.eps));
    }
  }
}

// ---------------------------------------------------------------- helpers shared by fwd / bwd
struct Affine {
  float a, b, mean, rstd;   // z = a*x + b ;  xhat = (x - mean)*rstd
};

__device__ __forceinline__ Affine load_affine(const GbnParams& p, int g, int c) {
  const int idx = p.stats_per_group ? g * p.C + c : c;
  Affine f;
  f.mean = __ldg(p.mean + idx);
  f.rstd = __ldg(p.rstd + idx);
  f.a = __ldg(p.gamma + c) * f.rstd;
  f.b = __ldg(p.beta + c) - f.mean * f.a;
  return f;
}

// max of z over the 3x3 window and its first (row-major) argmax, like at::max_pool2d_with_indices
__device__ __forceinline__ void window_max(const float* __restrict__ pl, int W, int ph, int pw, const Affine& f,
                                           float& zmax, int& arg) {
  const float* base = pl + (size_t)(3 * ph) * W + 3 * pw;
  zmax = -INFINITY;
  arg = 0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float z = fmaf(f.a, __ldg(base + r * W + q), f.b);
      if (z > zmax || z != z) {  // NaN propagates like PyTorch's pooling
        if (!(zmax != zmax)) { zmax = z; arg = r * 3 + q; }
      }
    }
  }
}

// same, on 9 values already in registers (row-major); also returns x at the argmax
__device__ __forceinline__ void window_load(const float* __restrict__ pl, int W, int ph, int pw, float (&xv)[9]) {
  const float* base = pl + (size_t)(3 * ph) * W + 3 * pw;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) xv[r * 3 + q] = __ldg(base + r * W + q);
}
__device__ __forceinline__ void window_pick(const float (&xv)[9], const Affine& f, float& zmax, int& arg, float& xmax) {
  zmax = -INFINITY;
  arg = 0;
  xmax = xv[0];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float z = fmaf(f.a, xv[k], f.b);
    if (z > zmax || z != z) {
      if (!(zmax != zmax)) { zmax = z; arg = k; xmax = xv[k]; }
    }
  }
}

// ---------------------------------------------------------------- forward
__global__ void __launch_bounds__(kThreads) gbn_fwd_kernel(const GbnParams p) {
  const int hw = p.H * p.W, phw = p.PH * p.PW;
  const size_t plane_stride = (size_t)p.C * hw, pplane_stride = (size_t)p.C * phw;
  for (int slab = blockIdx.x; slab < p.G * p.C; slab += gridDim.x) {
    const int g = slab / p.C, c = slab - g * p.C;
    const Affine f = load_affine(p, g, c);
    const float* x0 = p.x + ((size_t)g * p.group * p.C + c) * hw;
    float* y0 = p.y + ((size_t)g * p.group * p.C + c) * phw;
    const int total = p.group * phw;
    for (int o = threadIdx.x; o < total; o += kThreads) {
      const int i = o / phw, rem = o - i * phw;
      const int ph = rem / p.PW, pw = rem - ph * p.PW;
      float zmax;
      int arg;
      window_max(x0 + (size_t)i * plane_stride, p.W, ph, pw, f, zmax, arg);
      y0[(size_t)i * pplane_stride + rem] = zmax != zmax ? zmax : fmaxf(zmax, 0.f);
    }
  }
}

// ---------------------------------------------------------------- backward 1: reductions
__global__ void __launch_bounds__(kThreads) gbn_bwd_reduce_kernel(const GbnParams p) {
  __shared__ double red[kWarps];
  const int hw = p.H * p.W, phw = p.PH * p.PW;
  const size_t plane_stride = (size_t)p.C * hw, pplane_stride = (size_t)p.C * phw;
  for (int slab = blockIdx.x; slab < p.G * p.C; slab += gridDim.x) {
    const int g = slab / p.C, c = slab - g * p.C;
    const Affine f = load_affine(p, g, c);
    const float* x0 = p.x + ((size_t)g * p.group * p.C + c) * hw;
    const float* dy0 = p.dy + ((size_t)g * p.group * p.C + c) * phw;
    const int total = p.group * phw;
    float s1 = 0.f, s2 = 0.f;
    for (int o = threadIdx.x; o < total; o += kThreads) {
      const int i = o / phw, rem = o - i * phw;
      const int ph = rem / p.PW, pw = rem - ph * p.PW;
      const float* pl = x0 + (size_t)i * plane_stride;
      // all ten loads of the window are independent: issue them together, then decide
      const float dyv = __ldg(dy0 + (size_t)i * pplane_stride + rem);
      float xv[9];
      window_load(pl, p.W, ph, pw, xv);
      float zmax, xmax;
      int arg;
      window_pick(xv, f, zmax, arg, xmax);
      if (zmax > 0.f) {  // ReLU gate
        s1 += dyv;
        s2 = fmaf(dyv, (xmax - f.mean) * f.rstd, s2);
      }
    }
    const double t1 = block_sum_d((double)s1, red);
    const double t2 = block_sum_d((double)s2, red);
    if (threadIdx.x == 0) {
      p.out_sums[2 * slab] = (float)t1;
      p.out_sums[2 * slab + 1] = (float)t2;
    }
  }
}

// ---------------------------------------------------------------- backward 2: dx
// training (batch statistics): dx = a * (dz - mean(dz) - xhat * mean(dz * xhat)) over the slab
// eval (running statistics):   dx = a * dz
__global__ void __launch_bounds__(kThreads) gbn_bwd_dx_kernel(const GbnParams p) {
  const int hw = p.H * p.W, phw = p.PH * p.PW;
  const size_t plane_stride = (size_t)p.C * hw, pplane_stride = (size_t)p.C * phw;
  const int H3 = 3 * p.PH, W3 = 3 * p.PW;
  for (int slab = blockIdx.x; slab < p.G * p.C; slab += gridDim.x) {
    const int g = slab / p.C, c = slab - g * p.C;
    const Affine f = load_affine(p, g, c);
    float m1 = 0.f, m2 = 0.f;
    if (p.stats_per_group) {
      const float inv_m = 1.f / ((float)p.group * (float)hw);
      m1 = __ldg(p.sums + 2 * slab) * inv_m;
      m2 = __ldg(p.sums + 2 * slab + 1) * inv_m;
    }
    const float* x0 = p.x + ((size_t)g * p.group * p.C + c) * hw;
    const float* dy0 = p.dy + ((size_t)g * p.group * p.C + c) * phw;
    float* dx0 = p.dx + ((size_t)g * p.group * p.C + c) * hw;
    // pooled windows: 9 outputs each
    const int total = p.group * phw;
    for (int o = threadIdx.x; o < total; o += kThreads) {
      const int i = o / phw, rem = o - i * phw;
      const int ph = rem / p.PW, pw = rem - ph * p.PW;
      const float* pl = x0 + (size_t)i * plane_stride;
      float* dpl = dx0 + (size_t)i * plane_stride;
      const float dyv = __ldg(dy0 + (size_t)i * pplane_stride + rem);
      float xv[9];
      window_load(pl, p.W, ph, pw, xv);
      float zmax, xmax;
      int arg;
      window_pick(xv, f, zmax, arg, xmax);
      const float dz = zmax > 0.f ? dyv : 0.f;
      const size_t base = (size_t)(3 * ph) * p.W + 3 * pw;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const float xhat = (xv[r * 3 + q] - f.mean) * f.rstd;
          const float d = (r * 3 + q == arg) ? dz : 0.f;
          dpl[base + r * p.W + q] = f.a * (d - m1 - xhat * m2);
        }
      }
    }
    // elements dropped by the floor-mode pooling (right strip, bottom strip): dz = 0
    const int right = p.W - W3, bottom = p.H - H3;
    const int strip = p.H * right + bottom * W3;
    for (int o = threadIdx.x; o < p.group * strip; o += kThreads) {
      const int i = o / strip;
      int rem = o - i * strip, row, col;
      if (rem < p.H * right) { row = rem / right; col = W3 + (rem - row * right); }
      else { rem -= p.H * right; row = H3 + rem / W3; col = rem - (rem / W3) * W3; }
      const size_t off = (size_t)i * plane_stride + (size_t)row * p.W + col;
      const float xhat = (__ldg(x0 + off) - f.mean) * f.rstd;
      dx0[off] = f.a * (-m1 - xhat * m2);
    }
  }
}

int check(const GbnParams& p, const char* name) {
  AFSL_REQUIRE(p.G > 0 && p.group > 0 && p.C > 0 && p.H > 0 && p.W > 0, "%s: bad sizes G=%d group=%d C=%d H=%d W=%d", name,
               p.G, p.group, p.C, p.H, p.W);
  AFSL_REQUIRE((long long)p.G * p.C < (1LL << 31), "%s: too many (group, channel) slabs", name);
  return AFSL_OK;
}

int grid_for(const GbnParams& p) {
  const long long slabs = (long long)p.G * p.C;
  const long long cap = (long long)kNumSMs * 8;
  return (int)(slabs < cap ? slabs : cap);
}

}  // namespace
}  // namespace afsl

extern "C" int afsl_gbn_stats_f32(const float* x, float* mean, float* rstd, float* var_biased, int G, int group, int C,
                                   int H, int W, float eps, void* stream) {
  using namespace afsl;
  AFSL_REQUIRE(x && mean && rstd && var_biased, "afsl_gbn_stats_f32: null pointer");
  GbnParams p{};
  p.x = x; p.out_mean = mean; p.out_rstd = rstd; p.out_var = var_biased; p.eps = eps;
  p.G = G; p.group = group; p.C = C; p.H = H; p.W = W;
  if (int rc = check(p, "afsl_gbn_stats_f32")) return rc;
  gbn_stats_kernel<<<grid_for(p), kThreads, 0, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_gbn_stats_f32");
  return AFSL_OK;
}

extern "C" int afsl_gbn_relu_pool_fwd_f32(const float* x, const float* mean, const float* rstd, const float* gamma,
                                           const float* beta, float* y, int G, int group, int C, int H, int W,
                                           int stats_per_group, void* stream) {
  using namespace afsl;
  AFSL_REQUIRE(x && mean && rstd && gamma && beta && y, "afsl_gbn_relu_pool_fwd_f32: null pointer");
  AFSL_REQUIRE(H >= 3 && W >= 3, "afsl_gbn_relu_pool_fwd_f32: plane %dx%d smaller than the 3x3 pooling window", H, W);
  GbnParams p{};
  p.x = x; p.mean = mean; p.rstd = rstd; p.gamma = gamma; p.beta = beta; p.y = y;
  p.G = G; p.group = group; p.C = C; p.H = H; p.W = W; p.PH = H / 3; p.PW = W / 3; p.stats_per_group = stats_per_group;
  if (int rc = check(p, "afsl_gbn_relu_pool_fwd_f32")) return rc;
  gbn_fwd_kernel<<<grid_for(p), kThreads, 0, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_gbn_relu_pool_fwd_f32");
  return AFSL_OK;
}

extern "C" int afsl_gbn_relu_pool_bwd_f32(const float* x, const float* mean, const float* rstd, const float* gamma,
                                           const float* beta, const float* d_y, float* d_x, float* sums, int G,
                                           int group, int C, int H, int W, int stats_per_group, void* stream) {
  using namespace afsl;
  AFSL_REQUIRE(x && mean && rstd && gamma && beta && d_y && d_x && sums, "afsl_gbn_relu_pool_bwd_f32: null pointer");
  AFSL_REQUIRE(H >= 3 && W >= 3, "afsl_gbn_relu_pool_bwd_f32: plane %dx%d smaller than the 3x3 pooling window", H, W);
  GbnParams p{};
  p.x = x; p.mean = mean; p.rstd = rstd; p.gamma = gamma; p.beta = beta; p.dy = d_y; p.dx = d_x;
  p.out_sums = sums; p.sums = sums;
  p.G = G; p.group = group; p.C = C; p.H = H; p.W = W; p.PH = H / 3; p.PW = W / 3; p.stats_per_group = stats_per_group;
  if (int rc = check(p, "afsl_gbn_relu_pool_bwd_f32")) return rc;
  gbn_bwd_reduce_kernel<<<grid_for(p), kThreads, 0, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_gbn_relu_pool_bwd_f32 (reduce)");
  gbn_bwd_dx_kernel<<<grid_for(p), kThreads, 0, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_gbn_relu_pool_bwd_f32 (dx)");
  return AFSL_OK;
}
