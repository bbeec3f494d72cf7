// Prototype head forward for MANY-WAY episodes (20-way evaluation sweeps): one 128-thread CTA
// per episode, prototypes in shared memory, query rows in register batches.
//
// Same arithmetic and reference as proto_head.cu (models/util_functions.py:6-19, few_shot_classifier.py:108-116,
// loops/loss.py:24-37, loops/loops.py:79,271-272).  Why a third family: with W = 20 the warp-per-episode kernel
// (proto_head_warp.cu) is down to one row per butterfly (kB*W <= 32) and the lane-group kernel (proto_head.cu)
// re-reads all W*D prototype floats from shared memory for every query row - 100 rows x 20 KB = 2 MB of shared
// traffic per 205 KB of HBM traffic, i.e. bound by the 128 B/clk shared-memory pipe (0.20-0.30 of the HBM roofline).
// Here a warp keeps kB = 8 query rows in registers (lane l owns D/32 columns of every row, rows are coalesced
// 128-bit loads, warp_rows.cuh) and walks the prototypes kWB = 4 at a time: one shared-memory read of a prototype
// feeds 8 rows, and the 8 x 4 per-lane partial squared distances go through ONE transposing butterfly, after
// which lane u*4+b owns the distance of (row u, prototype w0+b).  Shared traffic per row drops 8x.
#include <cstdlib>

#include "proto_head.cuh"
#include "warp_rows.cuh"

namespace afsl {
namespace {

using namespace warp_rows;

constexpr int kWideThreads = 128;
constexpr int kWideWarps = kWideThreads / kWarp;
constexpr int kRB = 8;    // query rows per register batch
constexpr int kPB = 4;    // prototypes per butterfly (kRB * kPB = 32 values, one per lane)

struct WideSmem {
  float* protos;   // [W4][D]   rows W..W4-1 are zero
  float* score;    // [kWideWarps][kRB][W4]
  float* pp;       // [W4]      |p_w|^2 (matmul-form distances)
  float* part;     // [kWideWarps]
  int* hits;       // [kWideWarps]
  int* lab;        // [Ns]
};

__host__ __device__ inline int round4(int w) { return (w + 3) & ~3; }

__host__ __device__ inline size_t wide_words(int Ns, int W, int D) {
  const int W4 = round4(W);
  return (size_t)W4 * D + (size_t)kWideWarps * kRB * W4 + W4 + 2 * kWideWarps + (size_t)Ns + 4;
}

__device__ inline WideSmem wide_carve(float* base, int Ns, int W, int D) {
  const int W4 = round4(W);
  WideSmem s;
  s.protos = base;  base += (size_t)W4 * D;
  s.score = base;   base += (size_t)kWideWarps * kRB * W4;
  s.pp = base;      base += W4;
  s.part = base;    base += kWideWarps;
  s.hits = reinterpret_cast<int*>(base);  base += kWideWarps;
  s.lab = reinterpret_cast<int*>(base);
  return s;
}

template <int kV>
__global__ void __launch_bounds__(kWideThreads, 3) head_wide_fwd_kernel(const HeadParams p) {
  extern __shared__ __align__(16) float smem_raw[];
  constexpr int kH = kV / 2, kD = kV * 32;
  const int W = p.W, W4 = round4(W);
  const WideSmem s = wide_carve(smem_raw, p.Ns, W, kD);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int u_l = lane / kPB, b_l = lane - u_l * kPB;     // (row in batch, prototype in block) owned after the butterfly
  float* sc = s.score + (size_t)warp * kRB * W4;

  // zero the padding prototypes once (never overwritten)
  for (int i = threadIdx.x; i < (W4 - W) * kD; i += kWideThreads) s.protos[(size_t)W * kD + i] = 0.f;
  if (threadIdx.x < W4 - W) s.pp[W + threadIdx.x] = 0.f;

  for (int e = blockIdx.x; e < p.E; e += gridDim.x) {
    // ------------------------------------------------------------------ prototypes -> shared (and global)
    if (p.support) {
      for (int k = threadIdx.x; k < p.Ns; k += kWideThreads) s.lab[k] = p.s_labels[(size_t)e * p.Ns + k];
      __syncthreads();
      const float* sup = p.support + (size_t)e * p.Ns * kD;
      for (int w = warp; w < W; w += kWideWarps) {
        // rows of class w in ascending order (the order torch.nonzero yields in the reference), found with ballots:
        // up to kRB row loads in flight, added in row order
        int n = 0;
        f32x2 acc[kH];
#pragma unroll
        for (int j = 0; j < kH; ++j) acc[j] = 0ull;
        for (int k0 = 0; k0 < p.Ns; k0 += 32) {
          const int l = k0 + lane < p.Ns ? s.lab[k0 + lane] : -1;
          unsigned m = __ballot_sync(kFull, l == w);
          while (m) {
            f32x2 buf[kRB][kH];
            int got = 0;
#pragma unroll
            for (int u = 0; u < kRB; ++u)
              if (m) {
                load_row<kV>(sup + (size_t)(k0 + __ffs(m) - 1) * kD, lane, buf[u]);
                m &= m - 1;
                ++got;
              }
#pragma unroll
            for (int u = 0; u < kRB; ++u)
              if (u < got) {
#pragma unroll
                for (int j = 0; j < kH; ++j) acc[j] = add2(acc[j], buf[u][j]);
              }
            n += got;
          }
        }
        const float fn = (float)n;                          // n == 0 -> NaN, as the reference's empty mean
        f32x2 sq = 0ull;
#pragma unroll
        for (int j = 0; j < kH; ++j) {
          float a, b;
          unpack2(acc[j], a, b);
          acc[j] = pack2(__fdiv_rn(a, fn), __fdiv_rn(b, fn));
          sq = fma2(acc[j], acc[j], sq);
        }
        sts_row<kV>(s.protos + (size_t)w * kD, lane, acc);
        if (p.protos_out) store_row<kV>(p.protos_out + ((size_t)e * W + w) * kD, lane, acc);
        const float pw2 = warp_sum(sum2(sq));
        if (lane == 0) s.pp[w] = pw2;
      }
    } else {
      for (int w = warp; w < W; w += kWideWarps) {
        f32x2 x[kH], sq = 0ull;
        load_row<kV>(p.protos_in + ((size_t)e * W + w) * kD, lane, x);
        sts_row<kV>(s.protos + (size_t)w * kD, lane, x);
#pragma unroll
        for (int j = 0; j < kH; ++j) sq = fma2(x[j], x[j], sq);
        const float pw2 = warp_sum(sum2(sq));
        if (lane == 0) s.pp[w] = pw2;
      }
    }
    __syncthreads();

    // ------------------------------------------------------------------ query rows
    if (p.queries) {
      const int r0 = p.q_offsets ? p.q_offsets[e] : e * p.Nq;
      const int nrows = p.q_offsets ? p.q_offsets[e + 1] - r0 : p.Nq;
      const float* qry = p.queries + (size_t)r0 * kD;
      float nll_acc = 0.f;
      int hit = 0;
      // torch.cdist (few_shot_classifier.py:110) computes (q - p)^2 directly up to 25 rows on both sides and switches to
      // the matmul form |q|^2 + |p|^2 - 2 q.p (clamped at 0) beyond: same switch here, half the arithmetic per element
      const bool expanded = nrows > 25 || W > 25;
      for (int i0 = warp * kRB; i0 < nrows; i0 += kWideWarps * kRB) {
        f32x2 buf[kRB][kH];
#pragma unroll
        for (int u = 0; u < kRB; ++u) load_row<kV>(qry + (size_t)min(i0 + u, nrows - 1) * kD, lane, buf[u]);   // tail replays the last row
        if (!expanded) {
          for (int w0 = 0; w0 < W4; w0 += kPB) {
            float part[kRB * kPB];
#pragma unroll
            for (int b = 0; b < kPB; ++b) {
              f32x2 pr[kH];
              lds_row<kV>(s.protos + (size_t)(w0 + b) * kD, lane, pr);
#pragma unroll
              for (int u = 0; u < kRB; ++u) {
                f32x2 a0 = 0ull;
#pragma unroll
                for (int j = 0; j < kH; ++j) {
                  const f32x2 d0 = sub2(buf[u][j], pr[j]);
                  a0 = fma2(d0, d0, a0);
                }
                part[u * kPB + b] = sum2(a0);
              }
            }
            const float d2 = transpose_reduce<kRB * kPB>(part, lane);     // lane u*kPB+b: (row u, prototype w0+b)
            sc[u_l * W4 + w0 + b_l] = -sqrtf(d2);
          }
        } else {
          // |q_u|^2 of the batch: lane u*kPB+b receives row u's
          float qpart[kRB * kPB];
#pragma unroll
          for (int u = 0; u < kRB; ++u) {
            f32x2 a0 = 0ull;
#pragma unroll
            for (int j = 0; j < kH; ++j) a0 = fma2(buf[u][j], buf[u][j], a0);
            const float t = sum2(a0);
#pragma unroll
            for (int b = 0; b < kPB; ++b) qpart[u * kPB + b] = t;
          }
          const float qq = transpose_reduce<kRB * kPB>(qpart, lane);
          for (int w0 = 0; w0 < W4; w0 += kPB) {
            float part[kRB * kPB];
#pragma unroll
            for (int b = 0; b < kPB; ++b) {
              f32x2 pr[kH];
              lds_row<kV>(s.protos + (size_t)(w0 + b) * kD, lane, pr);
#pragma unroll
              for (int u = 0; u < kRB; ++u) {
                f32x2 a0 = 0ull;
#pragma unroll
                for (int j = 0; j < kH; ++j) a0 = fma2(buf[u][j], pr[j], a0);
                part[u * kPB + b] = sum2(a0);
              }
            }
            const float dot = transpose_reduce<kRB * kPB>(part, lane);
            const float d2 = fmaxf(fmaf(-2.f, dot, qq) + s.pp[w0 + b_l], 0.f);   // clamp_min(0) as at::_euclidean_dist
            sc[u_l * W4 + w0 + b_l] = -sqrtf(d2);
          }
        }
        __syncwarp();
        // per row: max / first argmax / sum of exponentials, kPB lanes per row
        const int i = i0 + u_l;
        const bool live = i < nrows;
        const float* srow = sc + u_l * W4;
        float m = -INFINITY;
        int am = 0x7fffffff;
        for (int w = b_l; w < W; w += kPB) {
          const float v = srow[w];
          if (v > m || (v == m && w < am)) { m = v; am = w; }
          if (v != v && am == 0x7fffffff) am = w;             // NaN row: keep something defined
        }
#pragma unroll
        for (int o = kPB / 2; o > 0; o >>= 1) {
          const float om = __shfl_xor_sync(kFull, m, o);
          const int oa = __shfl_xor_sync(kFull, am, o);
          if (om > m || (om == m && oa < am)) { m = om; am = oa; }
        }
        float se = 0.f;
        for (int w = b_l; w < W; w += kPB) se += expf(srow[w] - m);
#pragma unroll
        for (int o = kPB / 2; o > 0; o >>= 1) se += __shfl_xor_sync(kFull, se, o);
        if (live && b_l == 0) {
          if (p.pred) p.pred[r0 + i] = am;
          if (p.posterior) p.posterior[r0 + i] = m;
          if (p.q_labels) {
            const int y = p.q_labels[r0 + i];
            if (y >= 0 && y < W) nll_acc += -((srow[y] - m) - logf(se));   // log_softmax then NLL
            hit += (am == y);
          }
        }
        if (p.scores) {                                     // the batch's scores are one contiguous run of rows x W floats
          const int live_rows = min(kRB, nrows - i0);
          float* dst = p.scores + (size_t)(r0 + i0) * W;
          for (int o = lane; o < live_rows * W; o += 32) {
            const int u = o / W, w = o - u * W;
            dst[o] = sc[u * W4 + w];
          }
        }
        __syncwarp();
      }
      nll_acc = warp_sum(nll_acc);
      for (int o = 16; o > 0; o >>= 1) hit += __shfl_xor_sync(kFull, hit, o);
      if (lane == 0) { s.part[warp] = nll_acc; s.hits[warp] = hit; }
      __syncthreads();
      if (threadIdx.x == 0) {
        if (p.loss) {
          float tot = 0.f;
          for (int k = 0; k < kWideWarps; ++k) tot += s.part[k];
          p.loss[e] = tot / (float)nrows;
        }
        if (p.correct) {
          int c = 0;
          for (int k = 0; k < kWideWarps; ++k) c += s.hits[k];
          p.correct[e] = c;
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace

// Forward launches with W >= 8 prototypes and D in {128, 256}; AFSL_HEAD_WIDE=0 disables it (the parity tests run
// both this and the lane-group kernel).
int launch_head_wide(const HeadParams& p, bool bwd, cudaStream_t stream, const char* name, bool* handled) {
  *handled = false;
  if (bwd || p.W < 8 || !(p.support || p.protos_in)) return AFSL_OK;
  if (p.D != 128 && p.D != 256) return AFSL_OK;       // D = 64 rows are too short for 32 lanes: butterfly-bound, see proto_head.cu
  const char* env = getenv("AFSL_HEAD_WIDE");
  if (env && atoi(env) == 0) return AFSL_OK;
  const size_t bytes = wide_words(p.Ns, p.W, p.D) * sizeof(float);
  if (bytes > 200 * 1024) return AFSL_OK;
  void (*fn)(const HeadParams) = p.D == 128 ? head_wide_fwd_kernel<4> : head_wide_fwd_kernel<8>;
  *handled = true;
  if (int rc = opt_in_smem(fn, bytes, name)) return rc;
  const int grid = persistent_grid(fn, kWideThreads, bytes, p.E);
  fn<<<grid, kWideThreads, bytes, stream>>>(p);
  AFSL_CHECK_LAUNCH(name);
  return AFSL_OK;
}

}  // namespace afsl
