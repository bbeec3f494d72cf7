// SpecAugment view generation: one read of the spectrogram, four views written.
//
// Reference: SpecAugment.apply_augmentations and its three transforms,
// utils/augmentations.py:33-157.  Views: 0 copy, 1 time-warp (cubic Hermite spline through the
// control points (0,-1), (p,(p-d)*2/(T-1)-1), (T-1,1) -> grid_sample bilinear, zeros padding,
// align_corners=True), 2 time mask, 3 frequency mask; each applied to the original.  Mask
// positions and warp control points are drawn on the host in the reference's RNG order, so the
// mask views are bit-exact and the warp view differs only by fp32 rounding of the spline.
//
// HBM-bound streaming kernel (algorithmic traffic 4 B x (1 read + up to 4 writes) per element).
// One CTA handles kRows consecutive mel rows of one sample: the tile is contiguous in memory, so
// loads/stores are flat 128-bit accesses even though T (157, 126) is not a multiple of 4.  The tile
// (+1 halo row each side, needed because grid_sample's y coordinate is f +- 1ulp) is staged in
// shared memory for the time-warp gather.
#include "afsl_common.cuh"

namespace afsl {
namespace {

constexpr int kThreads = 256;
constexpr int kRows = 16;

struct SpecParams {
  const float* x;           // [N,F,T]
  float* views;             // [4,N,F,T]
  const int32_t* warp_p;    // [N]
  const int32_t* warp_d;    // [N]
  const float* src_x;       // [N,T] or null
  const int32_t* row_lo;    // [F]
  const float* row_w;       // [F]
  const int32_t* time_masks;  // [sets,num_mask,2] (start,len)
  const int32_t* freq_masks;  // [sets,num_mask,2]
  int num_mask;
  float mask_value;
  int N, set_size, F, T, views_mask;
};

// Normalised source x of output column t, following hspline_interpolate_1D
// (utils/augmentations.py:96-108) one rounded fp32 operation at a time.
__device__ __forceinline__ float spline_source_x(int t, int p, int d, int T) {
  const float y1 = __fsub_rn(__fdiv_rn((float)(2 * (p - d)), (float)(T - 1)), 1.f);
  const float m01 = __fdiv_rn(__fsub_rn(y1, -1.f), (float)p);
  const float m12 = __fdiv_rn(__fsub_rn(1.f, y1), (float)(T - 1 - p));
  const float mid = __fdiv_rn(__fadd_rn(m12, m01), 2.f);
  const bool upper = t > p;                       // searchsorted(left) over [p, T-1]
  const float x_lo = upper ? (float)p : 0.f;
  const float dx = upper ? (float)(T - 1 - p) : (float)p;
  const float y_lo = upper ? y1 : -1.f, y_hi = upper ? 1.f : y1;
  const float s_lo = upper ? mid : m01, s_hi = upper ? m12 : mid;
  const float u = __fdiv_rn(__fsub_rn((float)t, x_lo), dx);
  const float u2 = __fmul_rn(u, u);
  const float u3 = (float)((double)u * (double)u * (double)u);
  // rows of the Hermite matrix times (1, u, u^2, u^3): torch's matmul accumulates k = 0..3 with FMAs
  // (verified against torch CPU; the only residual difference is torch's 1-ulp vectorised powf)
  const float h00 = __fmaf_rn(2.f, u3, __fmaf_rn(-3.f, u2, 1.f));
  const float h10 = __fmaf_rn(1.f, u3, __fmaf_rn(-2.f, u2, u));
  const float h01 = __fmaf_rn(-2.f, u3, __fmul_rn(3.f, u2));
  const float h11 = __fmaf_rn(1.f, u3, -u2);
  float r = __fmul_rn(h00, y_lo);
  r = __fadd_rn(r, __fmul_rn(__fmul_rn(h10, s_lo), dx));
  r = __fadd_rn(r, __fmul_rn(h01, y_hi));
  r = __fadd_rn(r, __fmul_rn(__fmul_rn(h11, s_hi), dx));
  return r;
}

__global__ void __launch_bounds__(kThreads) specaug_kernel(const SpecParams p) {
  extern __shared__ __align__(16) float smem[];
  const int T = p.T, F = p.F;
  float* tile = smem;                                   // [(kRows+2) * T], row 0 = halo above
  int* col_lo = reinterpret_cast<int*>(tile + (kRows + 2) * T);   // [T]
  float* col_w = reinterpret_cast<float*>(col_lo + T);            // [T]

  const int tiles_per_sample = (F + kRows - 1) / kRows;
  const long long total_tiles = (long long)p.N * tiles_per_sample;
  const size_t plane = (size_t)p.N * F * T;
  for (long long tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x) {
    const int n = (int)(tile_id / tiles_per_sample);
    const int f0 = (int)(tile_id - (long long)n * tiles_per_sample) * kRows;
    const int rows = min(kRows, F - f0);
    const int set = n / p.set_size;
    const float* xs = p.x + (size_t)n * F * T;
    const bool want_warp = (p.views_mask & 2) != 0;

    if (want_warp) {
      for (int t = threadIdx.x; t < T; t += kThreads) {
        const float gx = p.src_x ? p.src_x[(size_t)n * T + t] : spline_source_x(t, p.warp_p[n], p.warp_d[n], T);
        // grid_sampler_unnormalize, align_corners=True: ((g + 1) / 2) * (size - 1)
        const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(T - 1));
        const float fl = floorf(ix);
        col_lo[t] = (int)fl;
        col_w[t] = __fsub_rn(ix, fl);
      }
      // halo rows f0-1 and f0+rows (zero when outside the image: zeros padding)
      for (int t = threadIdx.x; t < 2 * T; t += kThreads) {
        const bool below = t >= T;
        const int tt = below ? t - T : t;
        const int f = below ? f0 + rows : f0 - 1;
        tile[(below ? rows + 1 : 0) * T + tt] = (f >= 0 && f < F) ? __ldg(xs + (size_t)f * T + tt) : 0.f;
      }
    }
    // body: flat 128-bit loads; keep a copy in registers for the copy / mask views
    const int n4 = (rows * T) >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(xs + (size_t)f0 * T);
    for (int q = threadIdx.x; q < n4; q += kThreads) {
      const float4 v = ldg_stream(src4 + q);
      if (want_warp) {
        // tile + T is 16-byte aligned only when T % 4 == 0, so the shared copy is written as scalars
        float* dst = tile + T + 4 * q;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
      }
      const int e0 = 4 * q;
      const size_t out_off = (size_t)n * F * T + (size_t)f0 * T + e0;
      if (p.views_mask & 1) stg_stream(reinterpret_cast<float4*>(p.views + out_off), v);
      if (p.views_mask & (4 | 8)) {
        float tm[4] = {v.x, v.y, v.z, v.w}, fm[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int e = e0 + k;
          const int r = e / T, t = e - r * T;
          const int f = f0 + r;
          for (int m = 0; m < p.num_mask; ++m) {
            const int* tmk = p.time_masks + ((size_t)set * p.num_mask + m) * 2;
            const int* fmk = p.freq_masks + ((size_t)set * p.num_mask + m) * 2;
            if (t >= tmk[0] && t < tmk[0] + tmk[1]) tm[k] = p.mask_value;
            if (f >= fmk[0] && f < fmk[0] + fmk[1]) fm[k] = p.mask_value;
          }
        }
        if (p.views_mask & 4)
          stg_stream(reinterpret_cast<float4*>(p.views + 2 * plane + out_off), make_float4(tm[0], tm[1], tm[2], tm[3]));
        if (p.views_mask & 8)
          stg_stream(reinterpret_cast<float4*>(p.views + 3 * plane + out_off), make_float4(fm[0], fm[1], fm[2], fm[3]));
      }
    }
    if (want_warp) {
      __syncthreads();
      for (int q = threadIdx.x; q < n4; q += kThreads) {
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int e = 4 * q + k;
          const int r = e / T, t = e - r * T;
          const int f = f0 + r;
          const int yn = p.row_lo[f];
          const float wn = p.row_w[f];           // n = iy - floor(iy)
          const float ws = __fsub_rn(1.f, wn);   // s
          const int xw = col_lo[t];
          const float ww = col_w[t];             // w = ix - floor(ix)
          const float we = __fsub_rn(1.f, ww);   // e
          // corner fetch: shared tile when the row is inside [f0-1, f0+rows], else global; zero outside the image
          auto fetch = [&](int yy, int xx) -> float {
            if (xx < 0 || xx >= T || yy < 0 || yy >= F) return 0.f;
            const int rr = yy - (f0 - 1);
            if (rr >= 0 && rr < rows + 2) return tile[rr * T + xx];
            return __ldg(xs + (size_t)yy * T + xx);
          };
          const float v_nw = fetch(yn, xw), v_ne = fetch(yn, xw + 1);
          const float v_sw = fetch(yn + 1, xw), v_se = fetch(yn + 1, xw + 1);
          float acc = __fmul_rn(v_nw, __fmul_rn(ws, we));
          acc = __fadd_rn(acc, __fmul_rn(v_ne, __fmul_rn(ws, ww)));
          acc = __fadd_rn(acc, __fmul_rn(v_sw, __fmul_rn(wn, we)));
          acc = __fadd_rn(acc, __fmul_rn(v_se, __fmul_rn(wn, ww)));
          o[k] = acc;
        }
        const size_t out_off = (size_t)n * F * T + (size_t)f0 * T + 4 * q;
        stg_stream(reinterpret_cast<float4*>(p.views + plane + out_off), make_float4(o[0], o[1], o[2], o[3]));
      }
      __syncthreads();
    }
  }
}

}  // namespace
}  // namespace afsl

extern "C" int afsl_specaug_views_f32(const float* x, float* views, const int32_t* warp_p, const int32_t* warp_d,
                                       const float* src_x, const int32_t* row_lo, const float* row_w,
                                       const int32_t* time_masks, const int32_t* freq_masks, int num_mask,
                                       float mask_value, int N, int set_size, int F, int T, int views_mask,
                                       void* stream) {
  using namespace afsl;
  AFSL_REQUIRE(x && views, "afsl_specaug_views_f32: null pointer");
  AFSL_REQUIRE(N >= 0 && F > 0 && T > 1 && set_size > 0, "afsl_specaug_views_f32: bad sizes N=%d F=%d T=%d set=%d", N, F, T,
               set_size);
  AFSL_REQUIRE(F % 4 == 0, "afsl_specaug_views_f32: F=%d must be a multiple of 4 (128-bit tiles)", F);
  AFSL_REQUIRE(views_mask > 0 && views_mask < 16, "afsl_specaug_views_f32: views_mask=%d", views_mask);
  if (views_mask & 2) {
    AFSL_REQUIRE(row_lo && row_w, "afsl_specaug_views_f32: time-warp view needs row_lo/row_w");
    AFSL_REQUIRE(src_x || (warp_p && warp_d), "afsl_specaug_views_f32: time-warp view needs src_x or warp_p/warp_d");
  }
  if (views_mask & (4 | 8)) {
    AFSL_REQUIRE(num_mask >= 0 && (num_mask == 0 || (time_masks && freq_masks)),
                 "afsl_specaug_views_f32: mask views need time_masks/freq_masks");
  }
  if (N == 0) return AFSL_OK;
  SpecParams p{};
  p.x = x; p.views = views; p.warp_p = warp_p; p.warp_d = warp_d; p.src_x = src_x; p.row_lo = row_lo; p.row_w = row_w;
  p.time_masks = time_masks; p.freq_masks = freq_masks; p.num_mask = num_mask; p.mask_value = mask_value;
  p.N = N; p.set_size = set_size; p.F = F; p.T = T; p.views_mask = views_mask;
  const size_t bytes = ((size_t)(kRows + 2) * T + 2 * (size_t)T) * sizeof(float);
  if (int rc = opt_in_smem(specaug_kernel, bytes, "afsl_specaug_views_f32")) return rc;
  const long long tiles = (long long)N * ((F + kRows - 1) / kRows);
  const int cap = persistent_grid(specaug_kernel, kThreads, bytes, 1 << 30);
  const int grid = (int)(tiles < cap ? tiles : cap);
  specaug_kernel<<<grid, kThreads, bytes, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_specaug_views_f32");
  return AFSL_OK;
}
