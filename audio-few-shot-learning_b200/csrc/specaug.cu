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
// One CTA handles kRows consecutive mel rows of one sample, one warp per row.  T (157, 126) is not a
// multiple of 4, so rows are not 16-byte aligned: accesses are 4 B per lane / 128 B per warp
// instruction, which keeps every 32 B sector fully used.  The tile (+1 halo row each side, needed
// because grid_sample's y coordinate is f +- 1 ulp for 36 of the 128 rows) is staged in shared
// memory for the time-warp gather.
#include <cstdlib>

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
  const int32_t* set_ids;     // [N] set of each sample, or null: sample n belongs to set n / set_size
  const int32_t* time_masks;  // [sets,num_mask,2] (start,len)
  const int32_t* freq_masks;  // [sets,num_mask,2]
  int num_mask;
  float mask_value;
  int N, set_size, F, T, views_mask, split;
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

// One warp per mel row: lane l owns columns l, l+32, ... (kTJ of them), so everything that depends
// only on the column (warp source column / weight, time-mask flag) lives in registers for the whole
// tile and everything that depends only on the row (frequency-mask flag, y blend) is warp-uniform.
// Global accesses are 4-byte per lane, 128 B per warp instruction, fully coalesced for any T.
template <int kTJ>
__global__ void __launch_bounds__(kThreads) specaug_kernel(const SpecParams p) {
  extern __shared__ __align__(16) float tile[];          // [(kRows+2) * T], row 0 = halo above
  const int T = p.T, F = p.F;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarpsPerCta = kThreads / 32;
  const size_t plane = (size_t)p.N * F * T;
  const bool want_copy = p.views_mask & 1, want_warp = p.views_mask & 2, want_tm = p.views_mask & 4,
             want_fm = p.views_mask & 8;

  // work item = kSplit-th part of one sample (a run of row tiles): the per-column tables below are
  // computed once per item and reused for all its rows
  const int tiles_per_sample = (F + kRows - 1) / kRows;
  const int tiles_per_item = (tiles_per_sample + p.split - 1) / p.split;
  const long long total_items = (long long)p.N * p.split;
  for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {
    const int n = (int)(item / p.split);
    const int part = (int)(item - (long long)n * p.split);
    const int set = p.set_ids ? __ldg(p.set_ids + n) : n / p.set_size;
    const float* xs = p.x + (size_t)n * F * T;
    const int32_t* tmk = p.time_masks + (size_t)set * p.num_mask * 2;
    const int32_t* fmk = p.freq_masks + (size_t)set * p.num_mask * 2;

    // per-column registers
    int col_lo[kTJ];
    float col_w[kTJ];
    unsigned tmask_bits = 0;
#pragma unroll
    for (int j = 0; j < kTJ; ++j) {
      const int t = lane + 32 * j;
      col_lo[j] = 0;
      col_w[j] = 0.f;
      if (t < T) {
        if (want_warp) {
          const float gx = p.src_x ? __ldg(p.src_x + (size_t)n * T + t)
                                   : spline_source_x(t, __ldg(p.warp_p + n), __ldg(p.warp_d + n), T);
          // grid_sampler_unnormalize, align_corners=True: ((g + 1) / 2) * (size - 1)
          const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(T - 1));
          const float fl = floorf(ix);
          col_lo[j] = (int)fl;
          col_w[j] = __fsub_rn(ix, fl);
        }
        if (want_tm)
          for (int m = 0; m < p.num_mask; ++m) {
            const int t0 = __ldg(tmk + 2 * m), len = __ldg(tmk + 2 * m + 1);
            if (t >= t0 && t < t0 + len) tmask_bits |= 1u << j;
          }
      }
    }
    for (int tl = part * tiles_per_item; tl < min((part + 1) * tiles_per_item, tiles_per_sample); ++tl) {
    const int f0 = tl * kRows;
    const int rows = min(kRows, F - f0);
    // pass 1: rows (with halo when warping) -> shared tile; copy / mask views straight from registers
    const int r_begin = want_warp ? 0 : 1, r_end = want_warp ? rows + 2 : rows + 1;
    for (int r = r_begin + warp; r < r_end; r += kWarpsPerCta) {
      const int f = f0 - 1 + r;
      const bool inside = f >= 0 && f < F;
      const bool body = r >= 1 && r <= rows;
      bool frow = false;
      if (body && want_fm)
        for (int m = 0; m < p.num_mask; ++m) {
          const int a = __ldg(fmk + 2 * m), len = __ldg(fmk + 2 * m + 1);
          frow |= (f >= a && f < a + len);
        }
      const float* src = xs + (size_t)f * T;
      const size_t out_off = (size_t)n * F * T + (size_t)f * T;
#pragma unroll
      for (int j = 0; j < kTJ; ++j) {
        const int t = lane + 32 * j;
        if (t < T) {
          const float v = inside ? __ldg(src + t) : 0.f;
          if (want_warp) tile[r * T + t] = v;
          if (body) {
            if (want_copy) p.views[out_off + t] = v;
            if (want_tm) p.views[2 * plane + out_off + t] = ((tmask_bits >> j) & 1u) ? p.mask_value : v;
            if (want_fm) p.views[3 * plane + out_off + t] = frow ? p.mask_value : v;
          }
        }
      }
    }
    if (want_warp) {
      __syncthreads();
      for (int r = 1 + warp; r <= rows; r += kWarpsPerCta) {
        const int f = f0 - 1 + r;
        const int yn = __ldg(p.row_lo + f);
        const float wn = __ldg(p.row_w + f);          // n = iy - floor(iy); s = 1 - n
        const float ws = __fsub_rn(1.f, wn);
        const int rr = yn - (f0 - 1);
        const bool north_in = yn >= 0 && yn < F, south_in = yn + 1 >= 0 && yn + 1 < F;
        const bool staged = rr >= 0 && rr + 1 < rows + 2;
        const float* north = tile + rr * T;
        const float* south = north + T;
        const size_t out_off = plane + (size_t)n * F * T + (size_t)f * T;
#pragma unroll
        for (int j = 0; j < kTJ; ++j) {
          const int t = lane + 32 * j;
          if (t < T) {
            const int xw = col_lo[j];
            const float ww = col_w[j], we = __fsub_rn(1.f, ww);
            const bool west_in = xw >= 0 && xw < T, east_in = xw + 1 >= 0 && xw + 1 < T;
            float v_nw = 0.f, v_ne = 0.f, v_sw = 0.f, v_se = 0.f;
            if (staged) {
              if (north_in) { if (west_in) v_nw = north[xw]; if (east_in) v_ne = north[xw + 1]; }
              if (south_in && wn != 0.f) { if (west_in) v_sw = south[xw]; if (east_in) v_se = south[xw + 1]; }
            } else {  // generic row tables: fall back to global memory
              if (north_in) { if (west_in) v_nw = __ldg(xs + (size_t)yn * T + xw); if (east_in) v_ne = __ldg(xs + (size_t)yn * T + xw + 1); }
              if (south_in) { if (west_in) v_sw = __ldg(xs + (size_t)(yn + 1) * T + xw); if (east_in) v_se = __ldg(xs + (size_t)(yn + 1) * T + xw + 1); }
            }
            // out = nw*(s*e) + ne*(s*w) + sw*(n*e) + se*(n*w), each product and sum rounded (grid_sampler order)
            float acc = __fmul_rn(v_nw, __fmul_rn(ws, we));
            acc = __fadd_rn(acc, __fmul_rn(v_ne, __fmul_rn(ws, ww)));
            acc = __fadd_rn(acc, __fmul_rn(v_sw, __fmul_rn(wn, we)));
            acc = __fadd_rn(acc, __fmul_rn(v_se, __fmul_rn(wn, ww)));
            p.views[out_off + t] = acc;
          }
        }
      }
      __syncthreads();
    }
    }  // row tiles of the item
  }
}


// ------------------------------------------------------------------------------------------------------
// Vectorised variant (the one that normally runs): a CTA owns kTileRows consecutive mel rows of one sample.
// A sample is F*T contiguous floats and kTileRows*T is a multiple of 4, so a tile is a 16-byte aligned run
// of float4s whatever T is: the tile is read ONCE with 128-bit loads (plus two scalar halo rows), parked in
// shared memory for the time-warp gather, and all four views leave as 128-bit stores (the op is 80 % writes).
// Per-column warp tables (source column, blend weight) are computed once per CTA into shared memory.
// Element arithmetic is identical to specaug_kernel above (same rounded operations, same order).
constexpr int kTileRows = 32;
constexpr int kMaxMasksSm = 8;

// masks of one set: up to two (start, len) pairs in registers (len 0 = unused), more through shared memory
struct MaskSet {
  int a0, l0, a1, l1;
  const int* more;   // pairs 2.. in shared memory
  int n_more;
  __device__ __forceinline__ bool hit(int v) const {
    bool h = ((unsigned)(v - a0) < (unsigned)l0) | ((unsigned)(v - a1) < (unsigned)l1);
    for (int k = 0; k < n_more; ++k) h |= (unsigned)(v - more[2 * k]) < (unsigned)more[2 * k + 1];
    return h;
  }
};

__device__ __forceinline__ MaskSet load_masks(const int* m, int num) {
  MaskSet s;
  s.a0 = num > 0 ? m[0] : 0; s.l0 = num > 0 ? max(m[1], 0) : 0;
  s.a1 = num > 1 ? m[2] : 0; s.l1 = num > 1 ? max(m[3], 0) : 0;
  s.more = m + 4;
  s.n_more = num > 2 ? num - 2 : 0;
  return s;
}

__global__ void __launch_bounds__(kThreads, 6) specaug_tile_kernel(const SpecParams p) {
  extern __shared__ __align__(16) float sm[];
  const int T = p.T, F = p.F, ldT = T + 4;              // tile rows carry two zero columns on either side
  // a CTA walks p.split consecutive tiles of ONE sample: the column tables and masks are set up once
  const int tiles_per_sample = (F + kTileRows - 1) / kTileRows;
  const int ctas_per_sample = (tiles_per_sample + p.split - 1) / p.split;
  const int n = blockIdx.x / ctas_per_sample, part = blockIdx.x - n * ctas_per_sample;
  float* tile = sm;                                     // [(rows + 2) * ldT]: row 0 = halo above
  float* col_w = sm + (kTileRows + 2) * ldT;            // [T] east weight
  int* col_lo = reinterpret_cast<int*>(col_w + T);      // [T] west column inside a padded tile row (0 .. T+2)
  float* row_wn = reinterpret_cast<float*>(col_lo + T); // [kTileRows + 1] south weight of the row
  int* row_off = reinterpret_cast<int*>(row_wn + kTileRows + 1);   // [kTileRows + 1] tile offset of the north source row
  int* masks = row_off + kTileRows + 1;                 // [2][kMaxMasksSm][2]: time masks then frequency masks
  const size_t plane = (size_t)p.N * F * T;
  const bool want_copy = p.views_mask & 1, want_warp = p.views_mask & 2, want_tm = p.views_mask & 4,
             want_fm = p.views_mask & 8;
  const int set = p.set_ids ? __ldg(p.set_ids + n) : n / p.set_size;
  const float* xs = p.x + (size_t)n * F * T;

  // per-column tables and masks (once per CTA)
  if (want_warp) {
    for (int t = threadIdx.x; t < T; t += kThreads) {
      const float gx = p.src_x ? __ldg(p.src_x + (size_t)n * T + t)
                               : spline_source_x(t, __ldg(p.warp_p + n), __ldg(p.warp_d + n), T);
      // grid_sampler_unnormalize, align_corners=True: ((g + 1) / 2) * (size - 1)
      const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(T - 1));
      const float fl = floorf(ix);
      // west neighbour xw, east xw + 1; columns outside [0, T) read the zero padding (zeros padding mode):
      // padded index = xw + 2, clamped so that fully outside pairs land on (0,1) or (T+2,T+3)
      const float cl = fminf(fmaxf(fl, -2.f), (float)T);
      col_lo[t] = (int)cl + 2;
      col_w[t] = __fsub_rn(ix, fl);
    }
  }
  if (threadIdx.x < 2 * p.num_mask) {
    masks[threadIdx.x] = __ldg(p.time_masks + (size_t)set * p.num_mask * 2 + threadIdx.x);
    masks[2 * kMaxMasksSm + threadIdx.x] = __ldg(p.freq_masks + (size_t)set * p.num_mask * 2 + threadIdx.x);
  }
  __syncthreads();
  const MaskSet tms = load_masks(masks, want_tm ? p.num_mask : 0);
  const MaskSet fms = load_masks(masks + 2 * kMaxMasksSm, want_fm ? p.num_mask : 0);
  const unsigned magic = (unsigned)((0x100000000ull + T - 1) / T);    // i / T for i < 2^16 via mulhi, one fix-up
  const size_t plane4 = plane / 4;

  for (int tix = part * p.split; tix < min((part + 1) * p.split, tiles_per_sample); ++tix) {
  const int f0 = tix * kTileRows, rows = min(kTileRows, F - f0);
  // per-row tables, zero padding columns and halo rows of this tile
  int staged_all = 1;
  if (want_warp) {
    for (int r = threadIdx.x; r <= rows; r += kThreads) {   // entry `rows` is a dummy for the wrap of the last float4
      const int f = min(f0 + r, F - 1);
      const int yn = __ldg(p.row_lo + f);
      const int rr = yn - (f0 - 1);                    // tile row of the north source row; south = rr + 1
      const bool ok = rr >= 0 && rr + 1 < rows + 2;
      row_off[r] = (ok ? rr : 0) * ldT;
      row_wn[r] = __ldg(p.row_w + f);
      if (!ok && r < rows) staged_all = 0;
    }
    for (int r = threadIdx.x; r < rows + 2; r += kThreads) {
      float* tr = tile + r * ldT;
      tr[0] = tr[1] = tr[T + 2] = tr[T + 3] = 0.f;
    }
    // halo rows (scalar: rows are not 16-byte aligned); rows outside the image are zeros
    for (int t = threadIdx.x; t < 2 * T; t += kThreads) {
      const bool top = t < T;
      const int f = top ? f0 - 1 : f0 + rows, tt = top ? t : t - T;
      tile[(top ? 0 : rows + 1) * ldT + 2 + tt] = (f >= 0 && f < F) ? __ldg(xs + (size_t)f * T + tt) : 0.f;
    }
  }
  staged_all = __syncthreads_and(staged_all);
  const int n4 = rows * T / 4;
  const float4* src4 = reinterpret_cast<const float4*>(xs + (size_t)f0 * T);
  float4* out4 = reinterpret_cast<float4*>(p.views + (size_t)n * F * T + (size_t)f0 * T);
  // pass 1: one 128-bit read of the tile; copy / mask views straight from registers.
  // Element k of a float4 sits at (r + wrap_k, t + k - wrap_k * T) with wrap_k = (t + k >= T): branch-free.
  for (int i4 = threadIdx.x; i4 < n4; i4 += kThreads) {
    const float4 v = ldg_stream(src4 + i4);
    if (want_copy) stg_stream(out4 + i4, v);
    const int i = 4 * i4;
    int r = (int)__umulhi((unsigned)i, magic);
    r -= (r * T > i);
    const int t = i - r * T;
    const float e[4] = {v.x, v.y, v.z, v.w};
    float tmv[4], fmv[4];
    const bool f_lo = fms.hit(f0 + r), f_hi = fms.hit(f0 + r + 1);
    float* trow = tile + (r + 1) * ldT + 2 + t;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool wrap = t + k >= T;
      const int tk = wrap ? t + k - T : t + k;
      if (want_warp) trow[wrap ? k + 4 : k] = e[k];         // next tile row starts ldT = T + 4 further on
      tmv[k] = tms.hit(tk) ? p.mask_value : e[k];
      fmv[k] = (wrap ? f_hi : f_lo) ? p.mask_value : e[k];
    }
    if (want_tm) stg_stream(out4 + 2 * plane4 + i4, make_float4(tmv[0], tmv[1], tmv[2], tmv[3]));
    if (want_fm) stg_stream(out4 + 3 * plane4 + i4, make_float4(fmv[0], fmv[1], fmv[2], fmv[3]));
  }
  if (!want_warp) continue;
  __syncthreads();
  // pass 2: time-warp view, bilinear gather out of the shared tile
  //   out = nw*(s*e) + ne*(s*w) + sw*(n*e) + se*(n*w), each product and sum rounded (grid_sampler order)
  for (int i4 = threadIdx.x; i4 < n4; i4 += kThreads) {
    const int i = 4 * i4;
    int r = (int)__umulhi((unsigned)i, magic);
    r -= (r * T > i);
    const int t = i - r * T;
    float o[4];
    if (staged_all) {
      const float wn_lo = row_wn[r], wn_hi = row_wn[r + 1];
      const int off_lo = row_off[r], off_hi = row_off[r + 1];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool wrap = t + k >= T;
        const int tk = wrap ? t + k - T : t + k;
        const float wn = wrap ? wn_hi : wn_lo, ws = __fsub_rn(1.f, wn);
        const float ww = col_w[tk], we = __fsub_rn(1.f, ww);
        const float* nw = tile + (wrap ? off_hi : off_lo) + col_lo[tk];
        float acc = __fmul_rn(nw[0], __fmul_rn(ws, we));
        acc = __fadd_rn(acc, __fmul_rn(nw[1], __fmul_rn(ws, ww)));
        acc = __fadd_rn(acc, __fmul_rn(nw[ldT], __fmul_rn(wn, we)));
        acc = __fadd_rn(acc, __fmul_rn(nw[ldT + 1], __fmul_rn(wn, ww)));
        o[k] = acc;
      }
    } else {
      // generic row tables (source rows outside the tile + halo): gather from global memory
      int rk = r, tk = t;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int f = f0 + rk;
        const int yn = __ldg(p.row_lo + f);
        const float wn = __ldg(p.row_w + f), ws = __fsub_rn(1.f, wn);
        const bool north_in = yn >= 0 && yn < F, south_in = yn + 1 >= 0 && yn + 1 < F;
        const int xw = col_lo[tk] - 2;
        const float ww = col_w[tk], we = __fsub_rn(1.f, ww);
        const bool west_in = xw >= 0 && xw < T, east_in = xw + 1 >= 0 && xw + 1 < T;
        float v_nw = 0.f, v_ne = 0.f, v_sw = 0.f, v_se = 0.f;
        if (north_in) { if (west_in) v_nw = __ldg(xs + (size_t)yn * T + xw); if (east_in) v_ne = __ldg(xs + (size_t)yn * T + xw + 1); }
        if (south_in) { if (west_in) v_sw = __ldg(xs + (size_t)(yn + 1) * T + xw); if (east_in) v_se = __ldg(xs + (size_t)(yn + 1) * T + xw + 1); }
        float acc = __fmul_rn(v_nw, __fmul_rn(ws, we));
        acc = __fadd_rn(acc, __fmul_rn(v_ne, __fmul_rn(ws, ww)));
        acc = __fadd_rn(acc, __fmul_rn(v_sw, __fmul_rn(wn, we)));
        acc = __fadd_rn(acc, __fmul_rn(v_se, __fmul_rn(wn, ww)));
        o[k] = acc;
        if (++tk == T) { tk = 0; ++rk; }
      }
    }
    stg_stream(out4 + plane4 + i4, make_float4(o[0], o[1], o[2], o[3]));
  }
  __syncthreads();          // the next tile overwrites the shared tile and row tables
  }  // tiles of this CTA
}

}  // namespace
}  // namespace afsl

extern "C" int afsl_specaug_views_f32(const float* x, float* views, const int32_t* warp_p, const int32_t* warp_d,
                                       const float* src_x, const int32_t* row_lo, const float* row_w,
                                       const int32_t* set_ids, const int32_t* time_masks, const int32_t* freq_masks,
                                       int num_mask, float mask_value, int N, int set_size, int F, int T, int views_mask,
                                       void* stream) {
  using namespace afsl;
  AFSL_REQUIRE(x && views, "afsl_specaug_views_f32: null pointer");
  AFSL_REQUIRE(N >= 0 && F > 0 && T > 1 && set_size > 0, "afsl_specaug_views_f32: bad sizes N=%d F=%d T=%d set=%d", N, F, T,
               set_size);
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
  p.set_ids = set_ids; p.time_masks = time_masks; p.freq_masks = freq_masks; p.num_mask = num_mask; p.mask_value = mask_value;
  p.N = N; p.set_size = set_size; p.F = F; p.T = T; p.views_mask = views_mask;
  // vectorised tile kernel whenever samples are whole float4 runs (F*T % 4 == 0: always for the 128-mel inputs)
  static_assert(kTileRows % 4 == 0, "tiles must start on a float4 boundary for every T");
  const bool vec_ok = ((size_t)F * T) % 4 == 0 && T >= 4 && num_mask <= kMaxMasksSm && aligned16(x) && aligned16(views) &&
                      ((size_t)N * F * T) % 4 == 0 && (size_t)kTileRows * T < 65536;
  const char* vec_env = getenv("AFSL_SPECAUG_TILE");     // read per launch so the tests can exercise both kernels
  if (vec_ok && (!vec_env || atoi(vec_env) != 0)) {
    const size_t tb = ((size_t)(kTileRows + 2) * (T + 4) + 2 * (size_t)T + 2 * (kTileRows + 1) + 4 * kMaxMasksSm) * sizeof(float);
    if (int rc = opt_in_smem(specaug_tile_kernel, tb, "afsl_specaug_views_f32")) return rc;
    // whole samples per CTA (column tables computed once per sample) when that still fills the machine several
    // times over, otherwise one tile per CTA
    const int tps = (F + kTileRows - 1) / kTileRows;
    const int cap = persistent_grid(specaug_tile_kernel, kThreads, tb, 1 << 30);
    p.split = (long long)N >= 4LL * cap ? tps : 1;
    const long long ctas = (long long)N * ((tps + p.split - 1) / p.split);
    AFSL_REQUIRE(ctas < (1ll << 31), "afsl_specaug_views_f32: too many tiles (%lld)", ctas);
    specaug_tile_kernel<<<(unsigned)ctas, kThreads, tb, (cudaStream_t)stream>>>(p);
    AFSL_CHECK_LAUNCH("afsl_specaug_views_f32");
    return AFSL_OK;
  }
  const size_t bytes = (size_t)(kRows + 2) * T * sizeof(float);
  const int tj = (T + 31) / 32;
  AFSL_REQUIRE(tj <= 16, "afsl_specaug_views_f32: T=%d too long (max 512 columns)", T);
  void (*fn)(const SpecParams) = tj <= 4 ? specaug_kernel<4> : tj <= 5 ? specaug_kernel<5> : tj <= 8 ? specaug_kernel<8>
                                                                                          : specaug_kernel<16>;
  if (int rc = opt_in_smem(fn, bytes, "afsl_specaug_views_f32")) return rc;
  // whole samples per CTA when there are enough of them to fill the machine twice, else split samples
  const int cap = persistent_grid(fn, kThreads, bytes, 1 << 30);
  const int tiles_per_sample = (F + kRows - 1) / kRows;
  int split = 1;
  while ((long long)N * split < 2LL * cap && split < tiles_per_sample) split *= 2;
  p.split = split;
  const long long items = (long long)N * split;
  const int grid = (int)(items < cap ? items : cap);
  fn<<<grid, kThreads, bytes, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_specaug_views_f32");
  return AFSL_OK;
}
