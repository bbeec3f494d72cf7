// Encoder stage 1 fused: Conv3x3(1 -> 64, pad 1) + grouped BatchNorm + ReLU + MaxPool(3,3), forward and
// backward, without ever materialising the full-resolution 64-channel activation.
//
// Reference: conv_block(in_channels=1, 64, pool 3) = the first block of conv_encoder,
// models/main_modules.py:43-81, applied per view per set per episode (:18-23) -> batch statistics per
// group of `group` consecutive samples.  With one input channel the convolution is 9 multiply-adds per
// output, while its output is 64x larger than its input (5.1 MB vs 80 KB per spectrogram): the cuDNN path
// (conv -> stats -> BN/ReLU/pool, and wgrad over the full-resolution gradient) is bound by ~10 passes over
// that tensor.  Here:
//   moments : per group the 9 shifted sums S_k = sum x(p+k) and the 45 products R_kl = sum x(p+k) x(p+l)
//             over all positions p (zero padding).  The conv output u_c = sum_k w_ck x(p+k) has
//             sum u_c = sum_k w_ck S_k and sum u_c^2 = sum_kl w_ck w_cl R_kl, so mean / variance of all 64
//             channels follow from 54 numbers per group (host-side, tiny) - exact in real arithmetic.
//   forward : per pooled position recompute the 3x3 window of conv outputs from a shared-memory input
//             tile, apply z = a u + b (BatchNorm folded, conv bias folded into the statistics), ReLU, max.
//   backward: per pooled position recompute the window, route dy to the first argmax (as
//             at::max_pool2d_with_indices), and accumulate per (group, channel)
//             s1 = sum dz, s2 = sum dz*xhat, T_k = sum dz * x(p_argmax + k); the weight gradient is
//             dW_c[k] = sum_g a_gc [ T_gck - m1 S_gk - m2 rstd (sum_l w_cl R_glk - mean S_gk) ] (host, tiny).
// Compute-bound on the fp32 pipe (~90 FFMA per pooled output per channel), traffic = input + pooled output.
#include "afsl_common.cuh"

namespace afsl {
namespace {

constexpr int kThreads = 512;        // 16 warps x 4 channels: keeps the 44 backward accumulators per thread in registers
constexpr int kWarps = kThreads / kWarp;
constexpr int kC = 64;              // output channels of stage 1
constexpr int kCPW = kC / kWarps;   // channels per warp
constexpr int kBands = 8;           // pooled rows per tile
constexpr int kAcc = 11;            // s1, s2, T_0..T_8
constexpr unsigned char kInactive = 15;   // argmax code of an output zeroed by the ReLU (or NaN): no gradient

struct S1Params {
  const float* x;        // [G*group, H, W]
  const float* w;        // [kC, 9]
  const float* a;        // [G,kC] or [kC]  (z = a*u + b)
  const float* b;
  const float* mean;     // [G,kC] or [kC]  mean of u   (backward)
  const float* rstd;
  const float* dy;       // [G*group, kC, PH, PW]
  float* y;              // [G*group, kC, PH, PW]
  unsigned char* arg_out;      // [G*group, kC, PH, PW] or null: window argmax 0..8 of each pooled output, kInactive if ReLU cut it
  const unsigned char* arg_in; // backward: the forward's codes (null = recompute the windows)
  double* moments;       // [G, 54]
  float* partial;        // [G, parts, kC, kAcc]
  int G, group, H, W, PH, PW, per_group, parts;
  int nhwc;              // y / argmax / dy are channels-last [G*group, PH, PW, kC] instead of [G*group, kC, PH, PW]
};

// ---------------------------------------------------------------- input moments
// The 54 moments are not accumulated tap pair by tap pair (45 multiply-adds per pixel): with zero padding
//   R_kl = sum_p x(p+k) x(p+l) = A(l-k) - [edge row left out by k] - [edge column left out by k] + [corner],
// where A(d) = sum_q x(q) x(q+d) is the image autocorrelation at displacement d = l - k (13 distinct values for
// d in [-2,2]^2 up to sign), the edge terms are 1-D autocorrelations (lags 0..2) of the first / last row and
// column, and likewise S_k = sum x - edge row sum - edge column sum + corner (derivation: substitute q = p + k;
// the pixels q whose p falls outside the image are one edge row and/or one edge column).  Exact in real arithmetic,
// 14 multiply-adds per pixel.  A warp walks a 32-column strip down a row band with a sliding 3-row window in
// registers (5 shared-memory loads + 14 FMA per pixel); bands are staged with cp.async, double buffered.
constexpr int kMomRows = 32;              // rows per band
constexpr int kMomThreads = 320;          // 10 warps: 5 column strips x 2 half bands at W = 157
constexpr int kMomWarps = kMomThreads / kWarp;
constexpr int kMomAcc = 14;               // sum x, A(0,0..2), A(1,-2..2), A(2,-2..2)

__host__ __device__ inline int mom_ld(int W) { return ((W + 31) / 32) * 32 + 4; }   // cols -2 .. 32*strips+1, zero outside

// 4-byte async copy global -> shared; `valid` = false writes a zero (padding) without touching global memory
__device__ __forceinline__ void cp_async_f32_zfill(float* dst, const float* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

// one window step: A = row y (cols x-2..x+2, centre A[2]), B = row y+1, C = row y+2
__device__ __forceinline__ void mom_step(const float (&A)[5], const float (&B)[5], const float (&C)[5], float (&acc)[kMomAcc]) {
  const float c = A[2];
  acc[0] += c;
  acc[1] = fmaf(c, c, acc[1]);
  acc[2] = fmaf(c, A[3], acc[2]);
  acc[3] = fmaf(c, A[4], acc[3]);
#pragma unroll
  for (int j = 0; j < 5; ++j) acc[4 + j] = fmaf(c, B[j], acc[4 + j]);
#pragma unroll
  for (int j = 0; j < 5; ++j) acc[9 + j] = fmaf(c, C[j], acc[9 + j]);
}
__device__ __forceinline__ void mom_load(const float* row, float (&A)[5]) {
#pragma unroll
  for (int j = 0; j < 5; ++j) A[j] = row[j];
}

__global__ void __launch_bounds__(kMomThreads) stage1_moments_kernel(const S1Params p) {
  extern __shared__ __align__(16) float mtile[];         // two buffers of [(kMomRows + 2) * ld]
  const int H = p.H, W = p.W, hw = H * W, ld = mom_ld(W);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // grid = G * parts CTAs; CTA (g, part) strides over the group's (sample, band) units
  const int parts = gridDim.x / p.G;
  const int g = blockIdx.x / parts, part = blockIdx.x - g * parts;
  const float* x0 = p.x + (size_t)g * p.group * hw;
  const int bands_per_sample = (H + kMomRows - 1) / kMomRows;
  const int total_bands = p.group * bands_per_sample;
  const int strips = (W + 31) / 32;
  const int tile_floats = (kMomRows + 2) * ld;
  float acc[kMomAcc];
#pragma unroll
  for (int k = 0; k < kMomAcc; ++k) acc[k] = 0.f;
  // edge accumulators, role by warp: 0 first row, 1 last row, 2 first column, 3 last column: sum, lag 0, lag 1, lag 2;
  // lane 0 of warps 4..7: the four corners (value, square)
  float edge[4] = {0.f, 0.f, 0.f, 0.f};

  auto prefetch = [&](int bd, float* dst) {               // rows [i0, i0 + kMomRows + 2) x cols [-2, ld - 2), zero outside
    const int sm = bd / bands_per_sample, i0 = (bd - sm * bands_per_sample) * kMomRows;
    const float* pl = x0 + (size_t)sm * hw;
    for (int r = warp; r < kMomRows + 2; r += kMomWarps) {
      const int i = i0 + r;
      const bool row_in = i < H;
      const float* src = pl + (size_t)(row_in ? i : 0) * W - 2;
      for (int c = lane; c < ld; c += 32) {
        const bool in = row_in && c >= 2 && c < W + 2;
        cp_async_f32_zfill(dst + r * ld + c, in ? src + c : pl, in);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int buf = 0;
  if (part < total_bands) prefetch(part, mtile);
  for (int bd = part; bd < total_bands; bd += parts) {
    const int sm = bd / bands_per_sample, i0 = (bd - sm * bands_per_sample) * kMomRows;
    const int rows = min(kMomRows, H - i0);
    const float* tile = mtile + buf * tile_floats;
    if (bd + parts < total_bands) {
      prefetch(bd + parts, mtile + (buf ^ 1) * tile_floats);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    // ---- autocorrelations: items = column strips x two half bands
    const int half = (rows + 1) / 2;
    for (int item = warp; item < 2 * strips; item += kMomWarps) {
      const int strip = item >> 1, ya = (item & 1) ? half : 0, yb = (item & 1) ? rows : half;
      if (ya >= yb) continue;
      const float* col = tile + 32 * strip + lane;        // window column x-2 of pixel x = 32*strip + lane
      float ra[5], rb[5], rc[5];
      mom_load(col + ya * ld, ra);
      mom_load(col + (ya + 1) * ld, rb);
      mom_load(col + (ya + 2) * ld, rc);
      int y = ya;
      for (; y + 3 <= yb; y += 3) {                        // window rows rotate through the three register sets
        mom_step(ra, rb, rc, acc);
        mom_load(col + (y + 3) * ld, ra);
        mom_step(rb, rc, ra, acc);
        mom_load(col + (y + 4) * ld, rb);
        mom_step(rc, ra, rb, acc);
        mom_load(col + min(y + 5, kMomRows + 1) * ld, rc);      // past the halo only when never used
      }
      if (y < yb) { mom_step(ra, rb, rc, acc); ++y; }
      if (y < yb) {                                        // rows past the band's halo are never used: the loads above
        mom_load(col + min(y + 2, kMomRows + 1) * ld, ra); // stay inside the tile by clamping
        mom_step(rb, rc, ra, acc);
      }
    }
    // ---- edge rows / columns / corners of the sample inside this band
    const bool has_top = i0 == 0, has_bot = i0 + rows == H;
    if (warp == 0 && has_top) {
      const float* r = tile + 2;
      for (int xx = lane; xx < W; xx += 32) {
        const float v = r[xx];
        edge[0] += v; edge[1] = fmaf(v, v, edge[1]); edge[2] = fmaf(v, r[xx + 1], edge[2]); edge[3] = fmaf(v, r[xx + 2], edge[3]);
      }
    } else if (warp == 1 && has_bot) {
      const float* r = tile + (rows - 1) * ld + 2;
      for (int xx = lane; xx < W; xx += 32) {
        const float v = r[xx];
        edge[0] += v; edge[1] = fmaf(v, v, edge[1]); edge[2] = fmaf(v, r[xx + 1], edge[2]); edge[3] = fmaf(v, r[xx + 2], edge[3]);
      }
    } else if (warp == 2 || warp == 3) {
      const float* c = tile + (warp == 2 ? 2 : W + 1);
      for (int yy = lane; yy < rows; yy += 32) {
        const float v = c[yy * ld];
        edge[0] += v; edge[1] = fmaf(v, v, edge[1]); edge[2] = fmaf(v, c[(yy + 1) * ld], edge[2]);
        edge[3] = fmaf(v, c[(yy + 2) * ld], edge[3]);
      }
    } else if (warp >= 4 && warp < 8 && lane == 0) {
      const bool bottom = warp >= 6, right = (warp & 1) != 0;
      if (bottom ? has_bot : has_top) {
        const float v = tile[(bottom ? rows - 1 : 0) * ld + (right ? W + 1 : 2)];
        edge[0] += v; edge[1] = fmaf(v, v, edge[1]);
      }
    }
    __syncthreads();                                       // this buffer is refilled during the next iteration
    buf ^= 1;
  }
  // ---- CTA reduction (fp32 partials of one band set -> double, warps added in order), then the 54 moments
  __shared__ float wsum[kMomWarps][kMomAcc];
  __shared__ float esum[8][4];
  __shared__ double tot[kMomAcc + 32];
#pragma unroll
  for (int k = 0; k < kMomAcc; ++k) {
    const float t = warp_sum(acc[k]);
    if (lane == 0) wsum[warp][k] = t;
  }
  if (warp < 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float t = warp_sum(edge[k]);
      if (lane == 0) esum[warp][k] = t;
    }
  } else if (warp < 8 && lane == 0) {
    esum[warp][0] = edge[0]; esum[warp][1] = edge[1]; esum[warp][2] = 0.f; esum[warp][3] = 0.f;
  }
  __syncthreads();
  if (threadIdx.x < kMomAcc) {
    double t = 0.0;
    for (int q = 0; q < kMomWarps; ++q) t += (double)wsum[q][threadIdx.x];
    tot[threadIdx.x] = t;
  } else if (threadIdx.x >= 32 && threadIdx.x < 64) {
    const int i = threadIdx.x - 32;
    tot[kMomAcc + i] = (double)esum[i >> 2][i & 3];
  }
  __syncthreads();
  if (threadIdx.x < 54) {
    // tap k = (ky, kx) in {-1,0,1}^2, row-major; edge row left out by k: first row if ky = +1, last row if ky = -1
    const double* A = tot + 1;                             // A[0..2]: (0,0..2); A[3..7]: (1,-2..2); A[8..12]: (2,-2..2)
    const double* E = tot + kMomAcc;                       // E[4*e + j]: e = 0 top, 1 bottom, 2 left, 3 right; corners 4..7 (tl,tr,bl,br)
    int k, l;
    if (threadIdx.x < 9) { k = l = threadIdx.x; }
    else {                                                 // upper triangle row by row after the 9 sums
      int o = threadIdx.x - 9; k = 0;
      while (o >= 9 - k) { o -= 9 - k; ++k; }
      l = k + o;
    }
    const int ky = k / 3 - 1, kx = k % 3 - 1, ly = l / 3 - 1, lx = l % 3 - 1;
    const int erow = ky == 1 ? 0 : 1, ecol = kx == 1 ? 2 : 3;
    const int corner = 4 + (ky == 1 ? 0 : 2) + (kx == 1 ? 0 : 1);
    double r;
    if (threadIdx.x < 9) {
      r = tot[0];
      if (ky) r -= E[4 * erow];
      if (kx) r -= E[4 * ecol];
      if (ky && kx) r += E[4 * corner];
    } else {
      int dy = ly - ky, dx = lx - kx;
      if (dy < 0 || (dy == 0 && dx < 0)) { dy = -dy; dx = -dx; }
      r = dy == 0 ? A[dx] : A[3 + 5 * (dy - 1) + dx + 2];
      if (ky && ly == ky) r -= E[4 * erow + 1 + abs(lx - kx)];
      if (kx && lx == kx) r -= E[4 * ecol + 1 + abs(ly - ky)];
      if (ky && kx && l == k) r += E[4 * corner + 1];
    }
    p.moments[((size_t)g * parts + part) * 54 + threadIdx.x] = r;
  }
}

// ---------------------------------------------------------------- tile staging shared by fwd / bwd
// tile = rows [3*ph0 - 1, 3*ph0 + 3*bands + 1) x cols [-1, W + 1) of one sample, zero padded; stride W + 2
__device__ inline void stage_tile(const float* __restrict__ pl, float* tile, int H, int W, int ph0, int bands) {
  const int rows = 3 * bands + 2, ld = W + 2;
  for (int o = threadIdx.x; o < rows * ld; o += kThreads) {
    const int r = o / ld, c = o - r * ld;
    const int i = 3 * ph0 - 1 + r, j = c - 1;
    tile[o] = (i >= 0 && i < H && j >= 0 && j < W) ? __ldg(pl + (size_t)i * W + j) : 0.f;
  }
}

// 5x5 input patch of the pooling window at (band bl, column pw) inside the tile
__device__ __forceinline__ void load_patch(const float* tile, int ld, int bl, int pw, float (&v)[25]) {
  const float* base = tile + (size_t)(3 * bl) * ld + 3 * pw;
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int c = 0; c < 5; ++c) v[r * 5 + c] = base[r * ld + c];
}

// conv output at window element (r, q) from the patch and the channel's 9 weights
__device__ __forceinline__ float conv_at(const float (&v)[25], const float (&w)[9], int r, int q) {
  float u = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) u = fmaf(w[a * 3 + b], v[(r + a) * 5 + (q + b)], u);
  return u;
}

__global__ void __launch_bounds__(kThreads) stage1_fwd_kernel(const S1Params p) {
  extern __shared__ __align__(16) float smem[];
  float* tile = smem;                                    // [(3*kBands+2) * (W+2)]
  float* sw = tile + (3 * kBands + 2) * (p.W + 2);       // [kC*9] weights, [kC] a, [kC] b
  const int H = p.H, W = p.W, PH = p.PH, PW = p.PW, ld = W + 2, hw = H * W, phw = PH * PW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_per_sample = (PH + kBands - 1) / kBands;
  const long long total_tiles = (long long)p.G * p.group * tiles_per_sample;
  for (int i = threadIdx.x; i < kC * 9; i += kThreads) sw[i] = __ldg(p.w + i);
  __syncthreads();
  // this warp's kCPW channels: weights live in registers for the whole kernel
  float w[kCPW][9];
#pragma unroll
  for (int cc = 0; cc < kCPW; ++cc)
#pragma unroll
    for (int k = 0; k < 9; ++k) w[cc][k] = sw[(warp * kCPW + cc) * 9 + k];
  // per group: z = fma(a, u, b) on the bias-free convolution output u.  The affine is NOT folded into the taps:
  // when |a| is small the window's z values collide in fp32 and the winner is the first of the collided set,
  // which only matches at::max_pool2d_with_indices if z is a monotone function of u as it is in the eager chain
  float ca[kCPW], cb[kCPW];
  f32x2 w2[kCPW / 2][9], ca2[kCPW / 2], cb2[kCPW / 2];     // channel pairs packed for FFMA2
#pragma unroll
  for (int cp = 0; cp < kCPW / 2; ++cp)
#pragma unroll
    for (int k = 0; k < 9; ++k) w2[cp][k] = pack2(w[2 * cp][k], w[2 * cp + 1][k]);
  int cur_g = -1;
  for (long long tl = blockIdx.x; tl < total_tiles; tl += gridDim.x) {
    const int s = (int)(tl / tiles_per_sample), tix = (int)(tl - (long long)s * tiles_per_sample);
    const int g = s / p.group;
    const int ph0 = tix * kBands, bands = min(kBands, PH - ph0);
    __syncthreads();
    if (g != cur_g) {                                    // per-(group, channel) affine z = a u + b
#pragma unroll
      for (int cc = 0; cc < kCPW; ++cc) {
        const int idx = (p.per_group ? g * kC : 0) + warp * kCPW + cc;
        ca[cc] = __ldg(p.a + idx);
        cb[cc] = __ldg(p.b + idx);
      }
#pragma unroll
      for (int cp = 0; cp < kCPW / 2; ++cp) {
        ca2[cp] = pack2(ca[2 * cp], ca[2 * cp + 1]);
        cb2[cp] = pack2(cb[2 * cp], cb[2 * cp + 1]);
      }
      cur_g = g;
    }
    stage_tile(p.x + (size_t)s * hw, tile, H, W, ph0, bands);
    __syncthreads();
    const int npos = bands * PW;
    for (int pos = lane; pos < npos; pos += 32) {
      const int bl = pos / PW, pw = pos - bl * PW;
      float v[25];
      load_patch(tile, ld, bl, pw, v);
      const size_t o0 = ((size_t)s * kC) * phw + (size_t)(ph0 + bl) * PW + pw;
      float outv[kCPW];
      unsigned char outc[kCPW];
      // packed fp32x2: two channels per FFMA2 (same fma as the scalar chain, tap order a, b ascending)
      f32x2 v2[25];
#pragma unroll
      for (int k = 0; k < 25; ++k) v2[k] = pack2(v[k], v[k]);
      float zz[kCPW][9];
#pragma unroll
      for (int cp = 0; cp < kCPW / 2; ++cp) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            f32x2 u = 0ull;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int b = 0; b < 3; ++b) u = fma2(w2[cp][a * 3 + b], v2[(r + a) * 5 + (q + b)], u);
            const f32x2 z2 = fma2(ca2[cp], u, cb2[cp]);
            unpack2(z2, zz[2 * cp][r * 3 + q], zz[2 * cp + 1][r * 3 + q]);
          }
      }
#pragma unroll
      for (int cc = 0; cc < kCPW; ++cc) {
        const float (&z)[9] = zz[cc];
        // NaN / inf inputs poison the batch statistics, hence a and b, hence every z of the group: a window is
        // either NaN-free or all NaN, and fmaxf of all-NaN operands is NaN - same output as a NaN-sticky scan
        const float zmax = fmaxf(fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3])), fmaxf(fmaxf(z[4], z[5]), fmaxf(fmaxf(z[6], z[7]), z[8])));
        outv[cc] = zmax != zmax ? zmax : fmaxf(zmax, 0.f);
        int arg = 8;                                      // first maximum in window order (at::max_pool2d_with_indices)
        if (p.arg_out) {
#pragma unroll
          for (int k = 7; k >= 0; --k) arg = z[k] == zmax ? k : arg;
        }
        outc[cc] = zmax > 0.f ? (unsigned char)arg : kInactive;
      }
      if (p.nhwc) {      // the warp's kCPW = 4 channels of this pooled pixel are one float4 / one 32-bit word of codes
        const size_t px = ((size_t)s * phw + (size_t)(ph0 + bl) * PW + pw) * kC + warp * kCPW;
        *reinterpret_cast<float4*>(p.y + px) = make_float4(outv[0], outv[1], outv[2], outv[3]);
        if (p.arg_out) *reinterpret_cast<uchar4*>(p.arg_out + px) = make_uchar4(outc[0], outc[1], outc[2], outc[3]);
      } else {
#pragma unroll
        for (int cc = 0; cc < kCPW; ++cc) {
          const size_t o = o0 + (size_t)(warp * kCPW + cc) * phw;
          p.y[o] = outv[cc];
          if (p.arg_out) p.arg_out[o] = outc[cc];
        }
      }
    }
  }
}

// ---------------------------------------------------------------- forward, channels-last: lane = channel pair
// One warp per pooled pixel, lane l owns channels (2l, 2l+1): the pair's 9 taps, a and b are per-lane constants that
// stay in registers for the whole kernel (18 + 4 registers instead of 36 weights + 25 patch values + 36 window
// values per lane in the lane-per-pixel kernel above), the 5x5 input patch is read as warp-uniform shared-memory
// broadcasts and enters FFMA2 as its scalar operand, and a pooled pixel is stored as one 256-byte row of y plus
// 64 code bytes.  The instruction count per pooled output is the same as in the lane-per-pixel kernel (~210 issued
// per pixel x 64 channels, 90 of them FFMA2); what changes is 72 instead of 122 registers -> 24 warps per SM (was 16),
// conflict-free broadcast patch reads, no index division, and three independent CTAs per SM with cp.async
// double-buffered tiles that hide each other's staging: 1.58 -> 1.09 ms per 64 groups (ncu: fma pipe 44 -> 52 % active).
constexpr int kFwdThreads = 256;
constexpr int kFwdBands = 14;       // pooled rows per tile: fewer, longer tiles = fewer CTA barriers per pixel
constexpr int kFwdWarps = kFwdThreads / kWarp;

// kCodes = false (no gradient wanted: evaluation sweeps): no argmax codes, and since z = fma(a, u, b) is monotone in u the
// affine is applied once per pooled pixel to the window's extreme u instead of to its nine elements: the taps are
// negated where a < 0 (u' = sign(a) u exactly, fma(a, u, b) = fma(|a|, u', b) bit for bit), so max_k z_k =
// fma(|a|, max_k u'_k, b) - 82 instead of 90 FFMA2 and no compare / select chain per pixel.
template <bool kCodes>
__global__ void __launch_bounds__(kFwdThreads, 3) stage1_fwd_nhwc_kernel(const S1Params p) {
  extern __shared__ __align__(16) float tiles[];         // two buffers of [(3*kFwdBands+2) * (W+2)]
  const int H = p.H, W = p.W, PH = p.PH, PW = p.PW, ld = W + 2, hw = H * W, phw = PH * PW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_per_sample = (PH + kFwdBands - 1) / kFwdBands;
  const int bands_per_tile = (PH + tiles_per_sample - 1) / tiles_per_sample;      // even split: 42 rows -> 6 x 7
  const long long total_tiles = (long long)p.G * p.group * tiles_per_sample;
  const int tile_floats = (3 * kFwdBands + 2) * ld;
  f32x2 w2[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w2[k] = pack2(__ldg(p.w + (2 * lane) * 9 + k), __ldg(p.w + (2 * lane + 1) * 9 + k));
  f32x2 a2 = 0ull, b2 = 0ull;
  int cur_g = -1, buf = 0;
  // asynchronous staging of tile `t` (LDGSTS, zero fill outside the image): a warp per tile row, lanes across columns
  auto prefetch = [&](long long t, float* dst) {
    const int s = (int)(t / tiles_per_sample), tix = (int)(t - (long long)s * tiles_per_sample);
    const int ph0 = tix * bands_per_tile, rows = 3 * min(bands_per_tile, PH - ph0) + 2;
    const float* pl = p.x + (size_t)s * hw;
    for (int r = warp; r < rows; r += kFwdWarps) {
      const int i = 3 * ph0 - 1 + r;
      const bool row_in = i >= 0 && i < H;
      const float* src = pl + (size_t)(row_in ? i : 0) * W - 1;
      for (int c = lane; c < ld; c += 32) {
        const bool in = row_in && c >= 1 && c <= W;
        cp_async_f32_zfill(dst + r * ld + c, in ? src + c : pl, in);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (blockIdx.x < total_tiles) prefetch(blockIdx.x, tiles);
  for (long long tl = blockIdx.x; tl < total_tiles; tl += gridDim.x) {
    const int s = (int)(tl / tiles_per_sample), tix = (int)(tl - (long long)s * tiles_per_sample);
    const int g = s / p.group;
    const int ph0 = tix * bands_per_tile, bands = min(bands_per_tile, PH - ph0);
    const float* tile = tiles + buf * tile_floats;
    if (g != cur_g) {                                    // z = fma(a, u, b) on the bias-free convolution output u
      const int idx = (p.per_group ? g * kC : 0) + 2 * lane;
      float a_lo = __ldg(p.a + idx), a_hi = __ldg(p.a + idx + 1);
      if (!kCodes) {                                     // taps carry the sign of a, a its magnitude
        const float s_lo = a_lo < 0.f ? -1.f : 1.f, s_hi = a_hi < 0.f ? -1.f : 1.f;
#pragma unroll
        for (int k = 0; k < 9; ++k)
          w2[k] = pack2(s_lo * __ldg(p.w + (2 * lane) * 9 + k), s_hi * __ldg(p.w + (2 * lane + 1) * 9 + k));
        a_lo = fabsf(a_lo); a_hi = fabsf(a_hi);
      }
      a2 = pack2(a_lo, a_hi);
      b2 = pack2(__ldg(p.b + idx), __ldg(p.b + idx + 1));
      cur_g = g;
    }
    // the other buffer was released by the barrier that ended the previous iteration: refill it while this tile is computed
    if (tl + gridDim.x < total_tiles) {
      prefetch(tl + gridDim.x, tiles + (buf ^ 1) * tile_floats);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    int bl = 0, pw = warp;
    while (pw >= PW) { pw -= PW; ++bl; }
    while (bl < bands) {
      const float* base = tile + (3 * bl) * ld + 3 * pw;
      float v[25];
#pragma unroll
      for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) v[r * 5 + c] = base[r * ld + c];
      float zl[9], zh[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          f32x2 u = 0ull;
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
              const float t = v[(r + a) * 5 + (q + b)];
              u = fma2(w2[a * 3 + b], pack2(t, t), u);       // same fma chain as the scalar kernel, taps ascending
            }
          if (kCodes) unpack2(fma2(a2, u, b2), zl[r * 3 + q], zh[r * 3 + q]);
          else unpack2(u, zl[r * 3 + q], zh[r * 3 + q]);
        }
      float outv[2];
      unsigned int outc[2] = {0u, 0u};
      if (!kCodes) {
        float um[2];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const float (&z)[9] = cc ? zh : zl;
          um[cc] = fmaxf(fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3])), fmaxf(fmaxf(z[4], z[5]), fmaxf(fmaxf(z[6], z[7]), z[8])));
        }
        float z0, z1;
        unpack2(fma2(a2, pack2(um[0], um[1]), b2), z0, z1);
        outv[0] = z0 != z0 ? z0 : fmaxf(z0, 0.f);
        outv[1] = z1 != z1 ? z1 : fmaxf(z1, 0.f);
      }
#pragma unroll
      for (int cc = 0; cc < 2 && kCodes; ++cc) {
        const float (&z)[9] = cc ? zh : zl;
        // NaN handling as in stage1_fwd_kernel: a window is NaN-free or all NaN
        const float zmax = fmaxf(fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3])), fmaxf(fmaxf(z[4], z[5]), fmaxf(fmaxf(z[6], z[7]), z[8])));
        outv[cc] = zmax != zmax ? zmax : fmaxf(zmax, 0.f);
        int arg = 8;                                      // first maximum in window order (at::max_pool2d_with_indices)
#pragma unroll
        for (int k = 7; k >= 0; --k) arg = z[k] == zmax ? k : arg;
        outc[cc] = zmax > 0.f ? (unsigned int)arg : (unsigned int)kInactive;
      }
      const size_t px = ((size_t)s * phw + (size_t)(ph0 + bl) * PW + pw) * kC + 2 * lane;
      *reinterpret_cast<float2*>(p.y + px) = make_float2(outv[0], outv[1]);
      if (kCodes) *reinterpret_cast<unsigned short*>(p.arg_out + px) = (unsigned short)(outc[0] | (outc[1] << 8));
      pw += kFwdWarps;
      while (pw >= PW) { pw -= PW; ++bl; }
    }
    __syncthreads();
    buf ^= 1;
  }
}

__global__ void __launch_bounds__(kThreads) stage1_bwd_kernel(const S1Params p) {
  extern __shared__ __align__(16) float smem[];
  float* tile = smem;
  float* sw = tile + (3 * kBands + 2) * (p.W + 2);       // [kC*9] w, [kC] a, [kC] b, [kC] mean, [kC] rstd
  const int H = p.H, W = p.W, PH = p.PH, PW = p.PW, ld = W + 2, hw = H * W, phw = PH * PW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_per_sample = (PH + kBands - 1) / kBands;
  const int tiles_per_group = p.group * tiles_per_sample;
  // CTA (g, part) walks the tiles part, part + parts, ... of group g and owns partial[g][part]
  const int g = blockIdx.x / p.parts, part = blockIdx.x - g * p.parts;
  for (int i = threadIdx.x; i < kC * 9; i += kThreads) sw[i] = __ldg(p.w + i);
  for (int c = threadIdx.x; c < kC; c += kThreads) {
    const int idx = p.per_group ? g * kC + c : c;
    sw[kC * 9 + c] = __ldg(p.a + idx);
    sw[kC * 10 + c] = __ldg(p.b + idx);
    sw[kC * 11 + c] = __ldg(p.mean + idx);
    sw[kC * 12 + c] = __ldg(p.rstd + idx);
  }
  float acc[kCPW][kAcc];
#pragma unroll
  for (int cc = 0; cc < kCPW; ++cc)
#pragma unroll
    for (int k = 0; k < kAcc; ++k) acc[cc][k] = 0.f;

  for (int tl = part; tl < tiles_per_group; tl += p.parts) {
    const int sl = tl / tiles_per_sample, tix = tl - sl * tiles_per_sample;
    const int s = g * p.group + sl;
    const int ph0 = tix * kBands, bands = min(kBands, PH - ph0);
    __syncthreads();
    stage_tile(p.x + (size_t)s * hw, tile, H, W, ph0, bands);
    __syncthreads();
    const int npos = bands * PW;
    if (p.arg_in) {
      // the forward recorded which window element won (and whether the ReLU let it through): only that
      // element's convolution output (for xhat) and its 3x3 input neighbourhood are needed
      for (int pos = lane; pos < npos; pos += 32) {
        const int bl = pos / PW, pw = pos - bl * PW;
        const size_t o0 = ((size_t)s * kC + warp * kCPW) * phw + (size_t)(ph0 + bl) * PW + pw;
        int codes[kCPW];
        float dys[kCPW];
        if (p.nhwc) {
          const size_t px = ((size_t)s * phw + (size_t)(ph0 + bl) * PW + pw) * kC + warp * kCPW;
          const uchar4 c4 = __ldg(reinterpret_cast<const uchar4*>(p.arg_in + px));
          const float4 d4 = __ldg(reinterpret_cast<const float4*>(p.dy + px));
          codes[0] = c4.x; codes[1] = c4.y; codes[2] = c4.z; codes[3] = c4.w;
          dys[0] = d4.x; dys[1] = d4.y; dys[2] = d4.z; dys[3] = d4.w;
        } else {
#pragma unroll
          for (int cc = 0; cc < kCPW; ++cc) {            // all loads of the position first: kCPW x 2 in flight
            codes[cc] = __ldg(p.arg_in + o0 + (size_t)cc * phw);
            dys[cc] = __ldg(p.dy + o0 + (size_t)cc * phw);
          }
        }
#pragma unroll
        for (int cc = 0; cc < kCPW; ++cc) {
          const int c = warp * kCPW + cc;
          const int code = codes[cc];
          if (code >= 9) continue;
          const float dyv = dys[cc];
          const int ar = code / 3, aq = code - ar * 3;
          const float* nb = tile + (size_t)(3 * bl + ar) * ld + 3 * pw + aq;
          float xs[9], u = 0.f;
#pragma unroll
          for (int ka = 0; ka < 3; ++ka)
#pragma unroll
            for (int kb = 0; kb < 3; ++kb) {
              xs[ka * 3 + kb] = nb[ka * ld + kb];
              u = fmaf(sw[c * 9 + ka * 3 + kb], xs[ka * 3 + kb], u);
            }
          const float xhat = (u - sw[kC * 11 + c]) * sw[kC * 12 + c];
          acc[cc][0] += dyv;
          acc[cc][1] = fmaf(dyv, xhat, acc[cc][1]);
#pragma unroll
          for (int k = 0; k < 9; ++k) acc[cc][2 + k] = fmaf(dyv, xs[k], acc[cc][2 + k]);
        }
      }
      continue;
    }
    for (int pos = lane; pos < npos; pos += 32) {
      const int bl = pos / PW, pw = pos - bl * PW;
      float v[25];
      load_patch(tile, ld, bl, pw, v);
      const float* dyp = p.dy + ((size_t)s * kC) * phw + (size_t)(ph0 + bl) * PW + pw;
#pragma unroll
      for (int cc = 0; cc < kCPW; ++cc) {
        const int c = warp * kCPW + cc;
        const float dyv = __ldg(dyp + (size_t)c * phw);
        float w[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k] = sw[c * 9 + k];
        const float a = sw[kC * 9 + c], b = sw[kC * 10 + c];
        float zmax = -INFINITY, umax = 0.f;
        int arg = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const float u = conv_at(v, w, r, q);
            const float z = fmaf(a, u, b);
            if ((z > zmax || z != z) && !(zmax != zmax)) { zmax = z; umax = u; arg = r * 3 + q; }
          }
        if (zmax > 0.f) {                               // ReLU gate; dz = dy routed to the window argmax
          const float xhat = (umax - sw[kC * 11 + c]) * sw[kC * 12 + c];
          acc[cc][0] += dyv;
          acc[cc][1] = fmaf(dyv, xhat, acc[cc][1]);
          // x(p_argmax + tap): the argmax is a runtime index, so read the 3x3 neighbourhood from the shared tile
          const int ar = arg / 3, aq = arg - ar * 3;
          const float* nb = tile + (size_t)(3 * bl + ar) * ld + 3 * pw + aq;
#pragma unroll
          for (int ka = 0; ka < 3; ++ka)
#pragma unroll
            for (int kb = 0; kb < 3; ++kb)
              acc[cc][2 + ka * 3 + kb] = fmaf(dyv, nb[ka * ld + kb], acc[cc][2 + ka * 3 + kb]);
        }
      }
    }
  }
  // warp reduction of the register accumulators, one (channel, slot) at a time, then one store per value
  float* out = p.partial + (((size_t)g * p.parts + part) * kC) * kAcc;
#pragma unroll
  for (int cc = 0; cc < kCPW; ++cc)
#pragma unroll
    for (int k = 0; k < kAcc; ++k) {
      const float t = warp_sum(acc[cc][k]);
      if (lane == 0) out[(size_t)(warp * kCPW + cc) * kAcc + k] = t;
    }
}

// ---------------------------------------------------------------- backward, channels-last: lanes = channels
// With y / argmax / dy channels-last, a WARP takes one pooled pixel at a time and each lane two of its 64
// channels: dy (256 B) and the codes (64 B) of a pixel are one coalesced load per warp, the per-channel
// accumulators (s1, s2, T_0..T_8) are 22 registers per lane, and the 3x3 neighbourhoods of all lanes fall inside
// the pixel's 5x5 input patch (tile row stride = 8 mod 32 banks: the 25 words sit in 25 different banks).
// The 16 warps' accumulators are added in warp order through shared memory (deterministic).
constexpr int kLanesCh = 2;   // channels per lane

__device__ __forceinline__ int bwd_tile_ld(int W) { return ((W + 2 + 23) / 32) * 32 + 8; }   // >= W + 2, = 8 (mod 32)

__global__ void __launch_bounds__(kThreads, 2) stage1_bwd_nhwc_kernel(const S1Params p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, W = p.W, PH = p.PH, PW = p.PW, hw = H * W, phw = PH * PW;
  const int ld = bwd_tile_ld(W);
  float* tile = smem;                                   // [(3*kBands+2) * ld]
  float* red = tile + 2 * (3 * kBands + 2) * ld;        // [kWarps][kC][kAcc] cross-warp reduction (after the two tile buffers)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_per_sample = (PH + kBands - 1) / kBands;
  const int tiles_per_group = p.group * tiles_per_sample;
  const int g = blockIdx.x / p.parts, part = blockIdx.x - g * p.parts;
  const int c0 = kLanesCh * lane;
  float w[kLanesCh][9], mean[kLanesCh], rstd[kLanesCh];
#pragma unroll
  for (int cc = 0; cc < kLanesCh; ++cc) {
    const int idx = (p.per_group ? g * kC : 0) + c0 + cc;
#pragma unroll
    for (int k = 0; k < 9; ++k) w[cc][k] = __ldg(p.w + (c0 + cc) * 9 + k);
    mean[cc] = __ldg(p.mean + idx);
    rstd[cc] = __ldg(p.rstd + idx);
  }
  float acc[kLanesCh][kAcc];
#pragma unroll
  for (int cc = 0; cc < kLanesCh; ++cc)
#pragma unroll
    for (int k = 0; k < kAcc; ++k) acc[cc][k] = 0.f;

  // asynchronous staging (LDGSTS, zero fill outside the image) of tile `t` of this group: rows [3*ph0 - 1, 3*ph0 + 3*bands + 1)
  // x cols [-1, W + 1), row stride ld; a warp per tile row.  Two buffers: the next tile lands while this one is processed.
  const int tile_floats = (3 * kBands + 2) * ld;
  auto prefetch = [&](int t, float* dst) {
    const int sl = t / tiles_per_sample, tix = t - sl * tiles_per_sample;
    const int ph0 = tix * kBands, rows = 3 * min(kBands, PH - ph0) + 2;
    const float* pl = p.x + (size_t)(g * p.group + sl) * hw;
    for (int r = warp; r < rows; r += kWarps) {
      const int i = 3 * ph0 - 1 + r;
      const bool row_in = i >= 0 && i < H;
      const float* src = pl + (size_t)(row_in ? i : 0) * W - 1;
      for (int c = lane; c < W + 2; c += 32) {
        const bool in = row_in && c >= 1 && c <= W;
        cp_async_f32_zfill(dst + r * ld + c, in ? src + c : pl, in);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int buf = 0;
  if (part < tiles_per_group) prefetch(part, tile);
  for (int tl = part; tl < tiles_per_group; tl += p.parts) {
    const int sl = tl / tiles_per_sample, tix = tl - sl * tiles_per_sample;
    const int s = g * p.group + sl;
    const int ph0 = tix * kBands, bands = min(kBands, PH - ph0);
    const float* cur_tile = tile + buf * tile_floats;
    if (tl + p.parts < tiles_per_group) {
      prefetch(tl + p.parts, tile + (buf ^ 1) * tile_floats);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int npos = bands * PW;
    const size_t px0 = ((size_t)s * phw + (size_t)ph0 * PW) * kC + c0;    // this lane's channels of the tile's first pixel
    // software pipeline: the dy / codes of the next TWO pixels of this warp are in flight while one is processed
    int pos = warp;
    float2 dyv = make_float2(0.f, 0.f), dyv2 = dyv;
    // the two code bytes travel as ONE 16-bit register until they are used: a uchar2 is split into its bytes right after the
    // load, which made every prefetch wait for its own data (44 % of the kernel's stall samples, long scoreboard)
    unsigned int cd = kInactive | (kInactive << 8), cd2 = cd;                // (32-bit variables: no half-register packing)
    if (pos < npos) {
      dyv = __ldg(reinterpret_cast<const float2*>(p.dy + px0 + (size_t)pos * kC));
      cd = __ldg(reinterpret_cast<const unsigned short*>(p.arg_in + px0 + (size_t)pos * kC));
    }
    if (pos + kWarps < npos) {
      dyv2 = __ldg(reinterpret_cast<const float2*>(p.dy + px0 + (size_t)(pos + kWarps) * kC));
      cd2 = __ldg(reinterpret_cast<const unsigned short*>(p.arg_in + px0 + (size_t)(pos + kWarps) * kC));
    }
    int bl = 0, pw = warp;                                 // pos = bl * PW + pw, kept incrementally (no division)
    while (pw >= PW) { pw -= PW; ++bl; }
    for (; pos < npos; pos += kWarps) {
      const float2 dy_cur = dyv;
      const unsigned int cd_cur = cd;
      dyv = dyv2;
      cd = cd2;
      const int nxt = pos + 2 * kWarps;
      if (nxt < npos) {
        dyv2 = __ldg(reinterpret_cast<const float2*>(p.dy + px0 + (size_t)nxt * kC));
        cd2 = __ldg(reinterpret_cast<const unsigned short*>(p.arg_in + px0 + (size_t)nxt * kC));
      }
      const float* patch = cur_tile + (size_t)(3 * bl) * ld + 3 * pw;
      pw += kWarps;
      while (pw >= PW) { pw -= PW; ++bl; }
      const float dys[kLanesCh] = {dy_cur.x, dy_cur.y};
      const int codes[kLanesCh] = {cd_cur & 0xff, cd_cur >> 8};
#pragma unroll
      for (int cc = 0; cc < kLanesCh; ++cc) {
        const int code = codes[cc];
        if (code >= 9) continue;                          // ReLU cut this output: no gradient
        const int ar = code / 3, aq = code - ar * 3;
        const float* nb = patch + ar * ld + aq;
        float xs[9], u = 0.f;
#pragma unroll
        for (int ka = 0; ka < 3; ++ka)
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) {
            xs[ka * 3 + kb] = nb[ka * ld + kb];
            u = fmaf(w[cc][ka * 3 + kb], xs[ka * 3 + kb], u);
          }
        const float xhat = (u - mean[cc]) * rstd[cc];
        acc[cc][0] += dys[cc];
        acc[cc][1] = fmaf(dys[cc], xhat, acc[cc][1]);
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[cc][2 + k] = fmaf(dys[cc], xs[k], acc[cc][2 + k]);
      }
    }
    __syncthreads();                                       // this buffer is refilled during the next iteration
    buf ^= 1;
  }
  // add the warps' accumulators in warp order
#pragma unroll
  for (int cc = 0; cc < kLanesCh; ++cc)
#pragma unroll
    for (int k = 0; k < kAcc; ++k) red[(warp * kC + c0 + cc) * kAcc + k] = acc[cc][k];
  __syncthreads();
  float* out = p.partial + (((size_t)g * p.parts + part) * kC) * kAcc;
  for (int o = threadIdx.x; o < kC * kAcc; o += kThreads) {
    float t = 0.f;
    for (int wq = 0; wq < kWarps; ++wq) t += red[wq * kC * kAcc + o];
    out[o] = t;
  }
}

size_t smem_bytes(int W, bool bwd) {
  return ((size_t)(3 * kBands + 2) * (W + 2) + kC * 9 + kC * (bwd ? 4 : 2)) * sizeof(float);
}

int check(const S1Params& p, const char* name) {
  AFSL_REQUIRE(p.G > 0 && p.group > 0 && p.H >= 3 && p.W >= 3, "%s: bad sizes G=%d group=%d H=%d W=%d", name, p.G, p.group, p.H,
               p.W);
  AFSL_REQUIRE(p.W <= 1024, "%s: W=%d too wide for the shared-memory tile", name, p.W);
  return AFSL_OK;
}

}  // namespace
}  // namespace afsl

extern "C" int afsl_stage1_channels(void) { return afsl::kC; }
extern "C" int afsl_stage1_acc_slots(void) { return afsl::kAcc; }

extern "C" int afsl_stage1_moments_f64(const float* x, double* moments, int parts, int G, int group, int H, int W,
                                        void* stream) {
  using namespace afsl;
  AFSL_REQUIRE(x && moments && parts > 0, "afsl_stage1_moments_f64: null pointer / parts");
  S1Params p{};
  p.x = x; p.moments = moments; p.G = G; p.group = group; p.H = H; p.W = W;
  if (int rc = check(p, "afsl_stage1_moments_f64")) return rc;
  const size_t mb = 2 * (size_t)(kMomRows + 2) * mom_ld(W) * sizeof(float);
  if (int rc = opt_in_smem(stage1_moments_kernel, mb, "afsl_stage1_moments_f64")) return rc;
  stage1_moments_kernel<<<G * parts, kMomThreads, mb, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_stage1_moments_f64");
  return AFSL_OK;
}

extern "C" int afsl_stage1_fwd_f32(const float* x, const float* weight, const float* a, const float* b, float* y,
                                    unsigned char* argmax, int G, int group, int H, int W, int per_group,
                                    int channels_last, void* stream) {
  using namespace afsl;
  AFSL_REQUIRE(x && weight && a && b && y, "afsl_stage1_fwd_f32: null pointer");
  static_assert(kCPW == 4, "the channels-last stores move one float4 / uchar4 per lane");
  S1Params p{};
  p.x = x; p.w = weight; p.a = a; p.b = b; p.y = y; p.arg_out = argmax; p.nhwc = channels_last;
  p.G = G; p.group = group; p.H = H; p.W = W; p.PH = H / 3; p.PW = W / 3; p.per_group = per_group;
  if (int rc = check(p, "afsl_stage1_fwd_f32")) return rc;
  if (channels_last) {        // lane = channel pair, warp = pooled pixel
    const size_t nb = 2 * (size_t)(3 * kFwdBands + 2) * (W + 2) * sizeof(float);   // double-buffered tile
    void (*fn)(const S1Params) = argmax ? stage1_fwd_nhwc_kernel<true> : stage1_fwd_nhwc_kernel<false>;
    if (int rc = opt_in_smem(fn, nb, "afsl_stage1_fwd_f32")) return rc;
    const long long tiles = (long long)G * group * ((p.PH + kFwdBands - 1) / kFwdBands);
    const int cap = persistent_grid(fn, kFwdThreads, nb, 1 << 30);
    fn<<<(int)(tiles < cap ? tiles : cap), kFwdThreads, nb, (cudaStream_t)stream>>>(p);
    AFSL_CHECK_LAUNCH("afsl_stage1_fwd_f32");
    return AFSL_OK;
  }
  const size_t bytes = smem_bytes(W, false);
  if (int rc = opt_in_smem(stage1_fwd_kernel, bytes, "afsl_stage1_fwd_f32")) return rc;
  const long long tiles = (long long)G * group * ((p.PH + kBands - 1) / kBands);
  const int cap = persistent_grid(stage1_fwd_kernel, kThreads, bytes, 1 << 30);
  stage1_fwd_kernel<<<(int)(tiles < cap ? tiles : cap), kThreads, bytes, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_stage1_fwd_f32");
  return AFSL_OK;
}

extern "C" int afsl_stage1_bwd_f32(const float* x, const float* weight, const float* a, const float* b, const float* mean,
                                    const float* rstd, const float* d_y, const unsigned char* argmax, float* partial, int parts,
                                    int G, int group, int H, int W, int per_group, int channels_last, void* stream) {
  using namespace afsl;
  AFSL_REQUIRE(x && weight && a && b && mean && rstd && d_y && partial && parts > 0, "afsl_stage1_bwd_f32: null pointer / parts");
  AFSL_REQUIRE(!channels_last || argmax, "afsl_stage1_bwd_f32: the channels-last backward needs the forward's argmax codes");
  S1Params p{};
  p.nhwc = channels_last;
  p.x = x; p.w = weight; p.a = a; p.b = b; p.mean = mean; p.rstd = rstd; p.dy = d_y; p.partial = partial; p.arg_in = argmax;
  p.G = G; p.group = group; p.H = H; p.W = W; p.PH = H / 3; p.PW = W / 3; p.per_group = per_group; p.parts = parts;
  if (int rc = check(p, "afsl_stage1_bwd_f32")) return rc;
  if (channels_last) {
    const int ld = ((W + 2 + 23) / 32) * 32 + 8;
    const size_t nb = (2 * (size_t)(3 * kBands + 2) * ld + (size_t)kWarps * kC * kAcc) * sizeof(float);
    if (int rc = opt_in_smem(stage1_bwd_nhwc_kernel, nb, "afsl_stage1_bwd_f32")) return rc;
    stage1_bwd_nhwc_kernel<<<G * parts, kThreads, nb, (cudaStream_t)stream>>>(p);
    AFSL_CHECK_LAUNCH("afsl_stage1_bwd_f32");
    return AFSL_OK;
  }
  const size_t bytes = smem_bytes(W, true);
  if (int rc = opt_in_smem(stage1_bwd_kernel, bytes, "afsl_stage1_bwd_f32")) return rc;
  stage1_bwd_kernel<<<G * parts, kThreads, bytes, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_stage1_bwd_f32");
  return AFSL_OK;
}
