// Grouped BatchNorm + ReLU + MaxPool(3, stride 3), channels-last (NHWC) variant of gbn.cu.
//
// Same reference and arithmetic as gbn.cu (conv_block, models/main_modules.py:43-60, one set of batch statistics
// per group of `group` consecutive samples).  cuDNN's sm_100 convolution kernels are NHWC-native: with NCHW
// activations every fprop / dgrad / wgrad call is bracketed by layout-conversion kernels (22 % of the training
// step, profiles/).  Keeping the encoder activations channels-last removes them, so the per-stage fused kernels
// exist in this layout too.  x is physically [G*group, H, W, C], C a multiple of 4:
//   a thread owns one channel quad (float4 = 16 B; the C/4 quads of a pixel are one contiguous C*4-byte run, so a
//   warp reads whole pixels) and walks pixels; per-(group, channel) reductions are two-stage: per-CTA partials
//   in double -> a tiny finalize kernel that adds the parts in order (no atomics, bit-reproducible).
#include "afsl_common.cuh"

namespace afsl {
namespace {

constexpr int kThreads = 256;

struct NhwcParams {
  const float* x;        // [G*group, H, W, C]
  const float* mean;     // [G,C] (or [C] when stats_per_group == 0)
  const float* rstd;
  const float* gamma;    // [C]
  const float* beta;     // [C]
  const float* dy;       // [G*group, PH, PW, C]
  float* y;              // [G*group, PH, PW, C]
  float* dx;             // like x
  double* partial;       // [G, parts, C, 2]
  float* out_mean;       // [G,C]
  float* out_rstd;
  float* out_var;        // biased
  float* out_sums;       // [G,C,2]
  float eps;
  int G, group, C, H, W, PH, PW, stats_per_group, parts;
};

// block-wide sum over the threads that share a channel quad: threads t, t + Q, t + 2Q, ... (Q = C/4 quads)
// red: [kThreads][8] doubles.  On return threads 0..Q-1 hold the totals of their quad's 8 values.
__device__ inline void quad_reduce(double (&v)[8], double* red, int quads) {
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = v[k];
  __syncthreads();
  if (threadIdx.x < quads) {
    for (int t = threadIdx.x + quads; t < kThreads; t += quads)
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += red[t * 8 + k];
  }
  __syncthreads();
}

// ---------------------------------------------------------------- statistics
// partial[g][part][c] = (sum(x - k_c), sum((x - k_c)^2)) over the part's pixels, k_c = the group's first pixel
__global__ void __launch_bounds__(kThreads) nhwc_stats_partial_kernel(const NhwcParams p) {
  extern __shared__ __align__(16) double red[];
  const int quads = p.C >> 2, lanes = kThreads / quads;       // pixel lanes per CTA
  const int cq = threadIdx.x % quads, pl = threadIdx.x / quads;
  const int g = blockIdx.x / p.parts, part = blockIdx.x - g * p.parts;
  const long long npix = (long long)p.group * p.H * p.W;
  const float4* x4 = reinterpret_cast<const float4*>(p.x) + (size_t)g * npix * quads;
  const float4 shift = __ldg(x4 + cq);
  const long long per = (npix + p.parts - 1) / p.parts;
  const long long lo = part * per, hi = min(npix, lo + per);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (pl < lanes) {
    long long i = lo + pl;
    for (; i + 3LL * lanes < hi; i += 4LL * lanes) {           // four independent 128-bit loads in flight
      const float4 a = ldg_stream(x4 + (size_t)i * quads + cq);
      const float4 b = ldg_stream(x4 + (size_t)(i + lanes) * quads + cq);
      const float4 c = ldg_stream(x4 + (size_t)(i + 2LL * lanes) * quads + cq);
      const float4 d = ldg_stream(x4 + (size_t)(i + 3LL * lanes) * quads + cq);
      const float4 vs[4] = {a, b, c, d};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float d0 = vs[u].x - shift.x, d1 = vs[u].y - shift.y, d2 = vs[u].z - shift.z, d3 = vs[u].w - shift.w;
        s1[0] += d0; s1[1] += d1; s1[2] += d2; s1[3] += d3;
        s2[0] = fmaf(d0, d0, s2[0]); s2[1] = fmaf(d1, d1, s2[1]); s2[2] = fmaf(d2, d2, s2[2]); s2[3] = fmaf(d3, d3, s2[3]);
      }
    }
    for (; i < hi; i += lanes) {
      const float4 a = ldg_stream(x4 + (size_t)i * quads + cq);
      const float d0 = a.x - shift.x, d1 = a.y - shift.y, d2 = a.z - shift.z, d3 = a.w - shift.w;
      s1[0] += d0; s1[1] += d1; s1[2] += d2; s1[3] += d3;
      s2[0] = fmaf(d0, d0, s2[0]); s2[1] = fmaf(d1, d1, s2[1]); s2[2] = fmaf(d2, d2, s2[2]); s2[3] = fmaf(d3, d3, s2[3]);
    }
  }
  double v[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) { v[2 * k] = (double)s1[k]; v[2 * k + 1] = (double)s2[k]; }
  quad_reduce(v, red, quads);
  if (threadIdx.x < quads) {
    double* out = p.partial + (((size_t)g * p.parts + part) * p.C + 4 * cq) * 2;
#pragma unroll
    for (int k = 0; k < 8; ++k) out[k] = v[k];
  }
}

__global__ void nhwc_stats_finalize_kernel(const NhwcParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // (g, c)
  if (idx >= p.G * p.C) return;
  const int g = idx / p.C, c = idx - g * p.C;
  double t1 = 0.0, t2 = 0.0;
  for (int part = 0; part < p.parts; ++part) {
    const double* src = p.partial + (((size_t)g * p.parts + part) * p.C + c) * 2;
    t1 += src[0];
    t2 += src[1];
  }
  const double m = (double)p.group * p.H * p.W;
  const float shift = __ldg(p.x + (size_t)g * p.group * p.H * p.W * p.C + c);
  const double dmean = t1 / m;
  double var = t2 / m - dmean * dmean;
  if (var < 0.0) var = 0.0;
  p.out_mean[idx] = (float)((double)shift + dmean);
  p.out_var[idx] = (float)var;
  p.out_rstd[idx] = (float)(1.0 / sqrt(var + (double)p.eps));
}

// ---------------------------------------------------------------- helpers shared by fwd / bwd
struct Affine4 {
  float a[4], b[4], mean[4], rstd[4];   // z = a*x + b ;  xhat = (x - mean)*rstd
};

__device__ __forceinline__ Affine4 load_affine4(const NhwcParams& p, int g, int cq) {
  Affine4 f;
  const int base = (p.stats_per_group ? g * p.C : 0) + 4 * cq;
  const float4 m = __ldg(reinterpret_cast<const float4*>(p.mean + base));
  const float4 r = __ldg(reinterpret_cast<const float4*>(p.rstd + base));
  const float4 ga = __ldg(reinterpret_cast<const float4*>(p.gamma + 4 * cq));
  const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + 4 * cq));
  const float mm[4] = {m.x, m.y, m.z, m.w}, rr[4] = {r.x, r.y, r.z, r.w}, gg[4] = {ga.x, ga.y, ga.z, ga.w},
              bb[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f.mean[k] = mm[k];
    f.rstd[k] = rr[k];
    f.a[k] = gg[k] * rr[k];
    f.b[k] = bb[k] - mm[k] * f.a[k];
  }
  return f;
}

// the 3x3 window of one pooled pixel for this thread's channel quad: 9 independent 128-bit loads
__device__ __forceinline__ void window_load4(const float4* __restrict__ img, int W, int quads, int ph, int pw, int cq,
                                             float4 (&xv)[9]) {
  const float4* base = img + ((size_t)(3 * ph) * W + 3 * pw) * quads + cq;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) xv[r * 3 + q] = ldg_stream(base + ((size_t)r * W + q) * quads);
}

__device__ __forceinline__ float comp(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

// max of z over the window and its first (row-major) argmax, like at::max_pool2d_with_indices; NaN sticks
__device__ __forceinline__ void window_pick4(const float4 (&xv)[9], const Affine4& f, int k, float& zmax, int& arg,
                                             float& xmax) {
  zmax = -INFINITY;
  arg = 0;
  xmax = comp(xv[0], k);
#pragma unroll
  for (int e = 0; e < 9; ++e) {
    const float xe = comp(xv[e], k);
    const float z = fmaf(f.a[k], xe, f.b[k]);
    if (z > zmax || z != z) {
      if (!(zmax != zmax)) { zmax = z; arg = e; xmax = xe; }
    }
  }
}

// work items = (sample, pooled row, pooled column, quad), quad fastest: a warp covers whole pooled pixels
__device__ __forceinline__ void decode(long long item, int quads, int PW, int PH, int& n, int& ph, int& pw, int& cq) {
  cq = (int)(item % quads);
  long long r = item / quads;
  pw = (int)(r % PW);
  r /= PW;
  ph = (int)(r % PH);
  n = (int)(r / PH);
}

// ---------------------------------------------------------------- forward
__global__ void __launch_bounds__(kThreads) nhwc_fwd_kernel(const NhwcParams p) {
  const int quads = p.C >> 2;
  const long long total = (long long)p.G * p.group * p.PH * p.PW * quads;
  const float4* x4 = reinterpret_cast<const float4*>(p.x);
  float4* y4 = reinterpret_cast<float4*>(p.y);
  int cur_g = -1, cur_cq = -1;
  Affine4 f;
  for (long long item = (long long)blockIdx.x * kThreads + threadIdx.x; item < total; item += (long long)gridDim.x * kThreads) {
    int n, ph, pw, cq;
    decode(item, quads, p.PW, p.PH, n, ph, pw, cq);
    const int g = n / p.group;
    if (g != cur_g || cq != cur_cq) { f = load_affine4(p, g, cq); cur_g = g; cur_cq = cq; }
    float4 xv[9];
    window_load4(x4 + (size_t)n * p.H * p.W * quads, p.W, quads, ph, pw, cq, xv);
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float zmax, xmax;
      int arg;
      window_pick4(xv, f, k, zmax, arg, xmax);
      o[k] = zmax != zmax ? zmax : fmaxf(zmax, 0.f);
    }
    stg_stream(y4 + item, make_float4(o[0], o[1], o[2], o[3]));
  }
}

// ---------------------------------------------------------------- backward 1: sum(dz), sum(dz * xhat) per (group, channel)
__global__ void __launch_bounds__(kThreads) nhwc_bwd_reduce_kernel(const NhwcParams p) {
  extern __shared__ __align__(16) double red[];
  const int quads = p.C >> 2, lanes = kThreads / quads;
  const int cq = threadIdx.x % quads, pl = threadIdx.x / quads;
  const int g = blockIdx.x / p.parts, part = blockIdx.x - g * p.parts;
  const long long npool = (long long)p.group * p.PH * p.PW;    // pooled pixels of the group
  const long long per = (npool + p.parts - 1) / p.parts;
  const long long lo = part * per, hi = min(npool, lo + per);
  const float4* x4 = reinterpret_cast<const float4*>(p.x) + (size_t)g * p.group * p.H * p.W * quads;
  const float4* dy4 = reinterpret_cast<const float4*>(p.dy) + (size_t)g * npool * quads;
  const Affine4 f = load_affine4(p, g, cq);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (pl < lanes) {
    for (long long i = lo + pl; i < hi; i += lanes) {
      const int pw = (int)(i % p.PW);
      const long long r = i / p.PW;
      const int ph = (int)(r % p.PH), n = (int)(r / p.PH);
      const float4 dyv = ldg_stream(dy4 + (size_t)i * quads + cq);
      float4 xv[9];
      window_load4(x4 + (size_t)n * p.H * p.W * quads, p.W, quads, ph, pw, cq, xv);
      const float dys[4] = {dyv.x, dyv.y, dyv.z, dyv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float zmax, xmax;
        int arg;
        window_pick4(xv, f, k, zmax, arg, xmax);
        if (zmax > 0.f) {   // ReLU gate
          s1[k] += dys[k];
          s2[k] = fmaf(dys[k], (xmax - f.mean[k]) * f.rstd[k], s2[k]);
        }
      }
    }
  }
  double v[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) { v[2 * k] = (double)s1[k]; v[2 * k + 1] = (double)s2[k]; }
  quad_reduce(v, red, quads);
  if (threadIdx.x < quads) {
    double* out = p.partial + (((size_t)g * p.parts + part) * p.C + 4 * cq) * 2;
#pragma unroll
    for (int k = 0; k < 8; ++k) out[k] = v[k];
  }
}

__global__ void nhwc_sums_finalize_kernel(const NhwcParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // (g, c)
  if (idx >= p.G * p.C) return;
  const int g = idx / p.C, c = idx - g * p.C;
  double t1 = 0.0, t2 = 0.0;
  for (int part = 0; part < p.parts; ++part) {
    const double* src = p.partial + (((size_t)g * p.parts + part) * p.C + c) * 2;
    t1 += src[0];
    t2 += src[1];
  }
  p.out_sums[2 * idx] = (float)t1;
  p.out_sums[2 * idx + 1] = (float)t2;
}

// ---------------------------------------------------------------- backward 2: dx
// training (batch statistics): dx = a * (dz - mean(dz) - xhat * mean(dz * xhat)) over the group's channel
// eval (running statistics):   dx = a * dz
__global__ void __launch_bounds__(kThreads) nhwc_bwd_dx_kernel(const NhwcParams p) {
  const int quads = p.C >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(p.x);
  const float4* dy4 = reinterpret_cast<const float4*>(p.dy);
  float4* dx4 = reinterpret_cast<float4*>(p.dx);
  const float inv_m = 1.f / ((float)p.group * (float)(p.H * p.W));
  int cur_g = -1, cur_cq = -1;
  Affine4 f;
  float m1[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f};
  auto refresh = [&](int g, int cq) {
    if (g == cur_g && cq == cur_cq) return;
    f = load_affine4(p, g, cq);
    if (p.stats_per_group) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        m1[k] = __ldg(p.out_sums + 2 * ((size_t)g * p.C + 4 * cq + k)) * inv_m;
        m2[k] = __ldg(p.out_sums + 2 * ((size_t)g * p.C + 4 * cq + k) + 1) * inv_m;
      }
    }
    cur_g = g;
    cur_cq = cq;
  };
  // pooled windows: 9 outputs each
  const long long total = (long long)p.G * p.group * p.PH * p.PW * quads;
  for (long long item = (long long)blockIdx.x * kThreads + threadIdx.x; item < total; item += (long long)gridDim.x * kThreads) {
    int n, ph, pw, cq;
    decode(item, quads, p.PW, p.PH, n, ph, pw, cq);
    refresh(n / p.group, cq);
    const float4 dyv = ldg_stream(dy4 + item);
    float4 xv[9];
    const float4* img = x4 + (size_t)n * p.H * p.W * quads;
    window_load4(img, p.W, quads, ph, pw, cq, xv);
    const float dys[4] = {dyv.x, dyv.y, dyv.z, dyv.w};
    float o[9][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float zmax, xmax;
      int arg;
      window_pick4(xv, f, k, zmax, arg, xmax);
      const float dz = zmax > 0.f ? dys[k] : 0.f;
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        const float xhat = (comp(xv[e], k) - f.mean[k]) * f.rstd[k];
        const float d = e == arg ? dz : 0.f;
        o[e][k] = f.a[k] * (d - m1[k] - xhat * m2[k]);
      }
    }
    float4* dimg = dx4 + (size_t)n * p.H * p.W * quads + ((size_t)(3 * ph) * p.W + 3 * pw) * quads + cq;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 3; ++q)
        stg_stream(dimg + ((size_t)r * p.W + q) * quads, make_float4(o[r * 3 + q][0], o[r * 3 + q][1], o[r * 3 + q][2], o[r * 3 + q][3]));
  }
  // pixels dropped by the floor-mode pooling (right strip, bottom strip): dz = 0
  const int H3 = 3 * p.PH, W3 = 3 * p.PW;
  const int right = p.W - W3, bottom = p.H - H3;
  const int strip = p.H * right + bottom * W3;
  const long long stotal = (long long)p.G * p.group * strip * quads;
  for (long long item = (long long)blockIdx.x * kThreads + threadIdx.x; item < stotal; item += (long long)gridDim.x * kThreads) {
    const int cq = (int)(item % quads);
    long long r = item / quads;
    int rem = (int)(r % strip);
    const int n = (int)(r / strip);
    int row, col;
    if (rem < p.H * right) { row = rem / right; col = W3 + (rem - row * right); }
    else { rem -= p.H * right; row = H3 + rem / W3; col = rem - (rem / W3) * W3; }
    refresh(n / p.group, cq);
    const size_t off = ((size_t)n * p.H * p.W + (size_t)row * p.W + col) * quads + cq;
    const float4 xe = ldg_stream(x4 + off);
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = f.a[k] * (-m1[k] - (comp(xe, k) - f.mean[k]) * f.rstd[k] * m2[k]);
    stg_stream(dx4 + off, make_float4(o[0], o[1], o[2], o[3]));
  }
}

int check(const NhwcParams& p, const char* name) {
  AFSL_REQUIRE(p.G > 0 && p.group > 0 && p.C > 0 && p.H > 0 && p.W > 0, "%s: bad sizes G=%d group=%d C=%d H=%d W=%d", name,
               p.G, p.group, p.C, p.H, p.W);
  AFSL_REQUIRE(p.C % 4 == 0 && p.C / 4 <= kThreads, "%s: channels-last kernels need C %% 4 == 0 and C <= %d (C=%d)", name,
               4 * kThreads, p.C);
  return AFSL_OK;
}

int stream_grid(long long items) {
  const long long ctas = (items + kThreads - 1) / kThreads;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(ctas < cap ? (ctas > 0 ? ctas : 1) : cap);
}

}  // namespace
}  // namespace afsl

extern "C" int afsl_gbn_nhwc_parts(int G) {
  // CTAs per group of the two-stage reductions: fill the machine about four times over
  const int parts = (4 * afsl::kNumSMs + G - 1) / (G > 0 ? G : 1);
  return parts < 1 ? 1 : (parts > 64 ? 64 : parts);
}

extern "C" int afsl_gbn_stats_nhwc_f32(const float* x, double* partial, int parts, float* mean, float* rstd,
                                        float* var_biased, int G, int group, int C, int H, int W, float eps, void* stream) {
  using namespace afsl;
  AFSL_REQUIRE(x && partial && mean && rstd && var_biased && parts > 0, "afsl_gbn_stats_nhwc_f32: null pointer / parts");
  NhwcParams p{};
  p.x = x; p.partial = partial; p.parts = parts; p.out_mean = mean; p.out_rstd = rstd; p.out_var = var_biased; p.eps = eps;
  p.G = G; p.group = group; p.C = C; p.H = H; p.W = W;
  if (int rc = check(p, "afsl_gbn_stats_nhwc_f32")) return rc;
  const size_t smem = (size_t)kThreads * 8 * sizeof(double);
  nhwc_stats_partial_kernel<<<G * parts, kThreads, smem, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_gbn_stats_nhwc_f32 (partial)");
  nhwc_stats_finalize_kernel<<<(G * C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_gbn_stats_nhwc_f32 (finalize)");
  return AFSL_OK;
}

extern "C" int afsl_gbn_relu_pool_nhwc_fwd_f32(const float* x, const float* mean, const float* rstd, const float* gamma,
                                                const float* beta, float* y, int G, int group, int C, int H, int W,
                                                int stats_per_group, void* stream) {
  using namespace afsl;
  AFSL_REQUIRE(x && mean && rstd && gamma && beta && y, "afsl_gbn_relu_pool_nhwc_fwd_f32: null pointer");
  AFSL_REQUIRE(H >= 3 && W >= 3, "afsl_gbn_relu_pool_nhwc_fwd_f32: plane %dx%d smaller than the 3x3 pooling window", H, W);
  NhwcParams p{};
  p.x = x; p.mean = mean; p.rstd = rstd; p.gamma = gamma; p.beta = beta; p.y = y;
  p.G = G; p.group = group; p.C = C; p.H = H; p.W = W; p.PH = H / 3; p.PW = W / 3; p.stats_per_group = stats_per_group;
  if (int rc = check(p, "afsl_gbn_relu_pool_nhwc_fwd_f32")) return rc;
  nhwc_fwd_kernel<<<stream_grid((long long)G * group * p.PH * p.PW * (C / 4)), kThreads, 0, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_gbn_relu_pool_nhwc_fwd_f32");
  return AFSL_OK;
}

extern "C" int afsl_gbn_relu_pool_nhwc_bwd_f32(const float* x, const float* mean, const float* rstd, const float* gamma,
                                                const float* beta, const float* d_y, float* d_x, double* partial,
                                                int parts, float* sums, int G, int group, int C, int H, int W,
                                                int stats_per_group, void* stream) {
  using namespace afsl;
  AFSL_REQUIRE(x && mean && rstd && gamma && beta && d_y && d_x && partial && sums && parts > 0,
               "afsl_gbn_relu_pool_nhwc_bwd_f32: null pointer / parts");
  AFSL_REQUIRE(H >= 3 && W >= 3, "afsl_gbn_relu_pool_nhwc_bwd_f32: plane %dx%d smaller than the 3x3 pooling window", H, W);
  NhwcParams p{};
  p.x = x; p.mean = mean; p.rstd = rstd; p.gamma = gamma; p.beta = beta; p.dy = d_y; p.dx = d_x;
  p.partial = partial; p.parts = parts; p.out_sums = sums;
  p.G = G; p.group = group; p.C = C; p.H = H; p.W = W; p.PH = H / 3; p.PW = W / 3; p.stats_per_group = stats_per_group;
  if (int rc = check(p, "afsl_gbn_relu_pool_nhwc_bwd_f32")) return rc;
  const size_t smem = (size_t)kThreads * 8 * sizeof(double);
  nhwc_bwd_reduce_kernel<<<G * parts, kThreads, smem, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_gbn_relu_pool_nhwc_bwd_f32 (reduce)");
  nhwc_sums_finalize_kernel<<<(G * C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_gbn_relu_pool_nhwc_bwd_f32 (finalize)");
  const long long windows = (long long)G * group * p.PH * p.PW * (C / 4);
  nhwc_bwd_dx_kernel<<<stream_grid(windows), kThreads, 0, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_gbn_relu_pool_nhwc_bwd_f32 (dx)");
  return AFSL_OK;
}
