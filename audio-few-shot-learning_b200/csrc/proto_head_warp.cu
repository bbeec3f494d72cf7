// Prototype head, one WARP per episode, everything in registers.
//
// Same arithmetic as proto_head.cu (reference: models/util_functions.py:6-19, few_shot_classifier.py:108-116,
// loops/loss.py:24-37, loops/loops.py:79,271-272) for the shapes where the episode's prototypes fit the
// register file: W*D <= 1280 floats (5-way at D <= 256, 20-way at D = 64).  Lane l owns D/32 columns of
// every row, so
//   * prototypes P[w] and the prototype gradient dP[w] are register arrays (no shared memory, no CTA barrier);
//   * a support / query row is one or two fully coalesced 128-bit loads per lane (512 B per warp instruction),
//     kB rows in flight per warp, read from HBM exactly once;
//   * rows are processed kB at a time: the kB*W per-lane partial squared distances go through ONE transposing
//     butterfly (31 shuffles for 32 values instead of 5 per value) that leaves lane u*W+w with the distance of
//     (row u, prototype w), so sqrt / exp / log / the coefficient division run once per (row, prototype) and not
//     once per lane, and scores are stored coalesced;
//   * the backward is ONE pass over the query rows: the softmax coefficients of a batch are known while its rows
//     are still in registers, so dQ is stored and dP accumulated (ascending rows, deterministic) without a
//     second read; with the forward's prototypes passed back in, the support block is not re-read at all.
// The grid is one warp per episode (4 per CTA): the hardware CTA scheduler balances the tail.
#include "proto_head.cuh"
#include "warp_rows.cuh"

namespace afsl {
namespace {

using namespace warp_rows;

constexpr int kMaxSupport = 128;   // support labels of one episode staged in shared memory

// support labels -> shared (one slice per warp), per-class counts, and whether the labels are the
// block-sorted pattern 0..0 1..1 ... with kB rows per class (what datasets/batch_creation.py:35-60 produces)
template <int kW, int kB>
__device__ __forceinline__ bool stage_labels(const int32_t* labels, int Ns, int lane, int* slab, int (&cnt)[kW]) {
#pragma unroll
  for (int w = 0; w < kW; ++w) cnt[w] = 0;
  bool sorted = Ns == kW * kB;
  for (int k0 = 0; k0 < Ns; k0 += 32) {
    const int k = k0 + lane;
    const int l = k < Ns ? labels[k] : -1;
    if (k < Ns) slab[k] = l;
    sorted = sorted && __all_sync(kFull, k >= Ns || l == k / kB);
#pragma unroll
    for (int w = 0; w < kW; ++w) cnt[w] += __popc(__ballot_sync(kFull, l == w));
  }
  __syncwarp();
  return sorted;
}

__device__ __forceinline__ void query_span(const HeadParams& p, int e, int& r0, int& nrows) {
  r0 = p.q_offsets ? p.q_offsets[e] : e * p.Nq;
  nrows = p.q_offsets ? p.q_offsets[e + 1] - r0 : p.Nq;
}

template <int kW, int kV, int kB, bool kBwd>
__global__ void __launch_bounds__(kCtaThreads) head_warp_kernel(const HeadParams p) {
  __shared__ int slab_all[kWarpsPerCta][kMaxSupport];
  constexpr int kH = kV / 2, kD = kV * 32, kVals = kB * kW, kN = pow2_ceil(kVals);
  // the episode's prototypes: built in registers, then parked in this warp's shared-memory slice so the
  // row loop keeps its registers for the rows in flight (and dP in the backward)
  __shared__ __align__(16) float sp_all[kWarpsPerCta][kW * kD];
  static_assert(kVals <= 32, "a batch of rows must fit one value per lane");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * kWarpsPerCta + warp;
  if (e >= p.E) return;
  int* slab = slab_all[warp];
  float* sp = sp_all[warp];
  // this lane's (row-in-batch, prototype) after the transposing reduction
  const int u_l = lane / kW, w_l = lane - u_l * kW;
  const bool lane_valid = lane < kVals;

  int cnt[kW];
  // the forward builds prototypes from the support block; the backward only when they were not passed back in
  const bool build = kBwd ? (p.queries && !p.protos_in && p.support) : (p.support != nullptr);
  bool sorted = false;
  if (p.s_labels && (build || kBwd)) sorted = stage_labels<kW, kB>(p.s_labels + (size_t)e * p.Ns, p.Ns, lane, slab, cnt);

  // one padded row sequence: support rows in [0, ns_pad), query rows in [ns_pad, total), both in whole batches
  int r0 = 0, nrows = 0;
  if (p.queries) query_span(p, e, r0, nrows);
  const int ns = build ? p.Ns : 0;
  const int ns_pad = (ns + kB - 1) / kB * kB;
  const int total = ns_pad + (nrows + kB - 1) / kB * kB;
  const float* sup = build ? p.support + (size_t)e * p.Ns * kD : nullptr;
  const float* qry = p.queries ? p.queries + (size_t)r0 * kD : nullptr;
  auto row_ptr = [&](int t) -> const float* {       // slots past the end of a block replay its last row
    if (t < ns_pad) return sup + (size_t)min(t, ns - 1) * kD;
    return qry + (size_t)min(t - ns_pad, nrows - 1) * kD;
  };

  if (!build && p.protos_in && p.queries) {
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      f32x2 x[kH];
      load_row<kV>(p.protos_in + ((size_t)e * kW + w) * kD, lane, x);
      sts_row<kV>(sp + w * kD, lane, x);
    }
  }
  f32x2 buf[kB][kH];
#pragma unroll
  for (int u = 0; u < kB; ++u)
    if (u < total) load_row<kV>(row_ptr(u), lane, buf[u]);

  // ------------------------------------------------------------------ prototypes
  if (build) {
    f32x2 P[kW][kH];
#pragma unroll
    for (int w = 0; w < kW; ++w)
#pragma unroll
      for (int j = 0; j < kH; ++j) P[w][j] = 0ull;
    if (sorted) {                    // class w = batch w: no label tests, rows added in ascending order
#pragma unroll
      for (int w = 0; w < kW; ++w)
#pragma unroll
        for (int u = 0; u < kB; ++u) {
#pragma unroll
          for (int j = 0; j < kH; ++j) P[w][j] = add2(P[w][j], buf[u][j]);
          if ((w + 1) * kB + u < total) load_row<kV>(row_ptr((w + 1) * kB + u), lane, buf[u]);
        }
    } else {
      for (int t0 = 0; t0 < ns_pad; t0 += kB) {
#pragma unroll
        for (int u = 0; u < kB; ++u) {
          const int lk = t0 + u < ns ? slab[t0 + u] : -1;
#pragma unroll
          for (int w = 0; w < kW; ++w)
            if (lk == w) {
#pragma unroll
              for (int j = 0; j < kH; ++j) P[w][j] = add2(P[w][j], buf[u][j]);
            }
          if (t0 + u + kB < total) load_row<kV>(row_ptr(t0 + u + kB), lane, buf[u]);
        }
      }
    }
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      const float fn = (float)cnt[w];  // empty class -> NaN, as the reference's empty mean
#pragma unroll
      for (int j = 0; j < kH; ++j) {
        float a, b;
        unpack2(P[w][j], a, b);
        P[w][j] = pack2(__fdiv_rn(a, fn), __fdiv_rn(b, fn));
      }
      sts_row<kV>(sp + w * kD, lane, P[w]);
      if (!kBwd && p.protos_out) store_row<kV>(p.protos_out + ((size_t)e * kW + w) * kD, lane, P[w]);
    }
  }
  __syncwarp();

  // ------------------------------------------------------------------ query rows, kB at a time
  f32x2 dP[kBwd ? kW : 1][kH];     // holds +sum_i coef[i,w] (q_i - p_w); negated at the end
#pragma unroll
  for (int w = 0; w < (kBwd ? kW : 1); ++w)
#pragma unroll
    for (int j = 0; j < kH; ++j) dP[w][j] = 0ull;
  float nll = 0.f;
  int hit = 0;
  const float dl = (kBwd && p.d_loss && nrows > 0) ? p.d_loss[e] / (float)nrows : 0.f;
  for (int t0 = ns_pad; t0 < total; t0 += kB) {
    const int i0 = t0 - ns_pad;
    const int i_l = i0 + u_l;
    const bool live_l = lane_valid && i_l < nrows;
    const int y_l = (live_l && p.q_labels) ? p.q_labels[r0 + i_l] : -1;
    float dsc_l = 0.f;
    if (kBwd && p.d_scores && live_l) dsc_l = p.d_scores[(size_t)(r0 + i_l) * kW + w_l];
    float part[kN];
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      f32x2 pw[kH];
      lds_row<kV>(sp + w * kD, lane, pw);
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        f32x2 a0 = 0ull, a1 = 0ull;
#pragma unroll
        for (int j = 0; j < kH; ++j) {
          const f32x2 d = sub2(buf[u][j], pw[j]);
          if (j & 1) a1 = fma2(d, d, a1); else a0 = fma2(d, d, a0);
        }
        part[u * kW + w] = kH > 1 ? sum2(add2(a0, a1)) : sum2(a0);
      }
    }
    if (!kBwd) {                      // forward: the batch's rows are done with, fetch the next batch
#pragma unroll
      for (int u = 0; u < kB; ++u)
        if (t0 + u + kB < total) load_row<kV>(row_ptr(t0 + u + kB), lane, buf[u]);
    }
#pragma unroll
    for (int k = kVals; k < kN; ++k) part[k] = 0.f;
    // lane u*W+w: squared distance of (row u, prototype w); score = -distance
    const float sc = -sqrtf(reduce_values<kN>(part, lane));
    // max / first argmax / sum exp over the row's W lanes, prototypes in ascending order
    float m = -INFINITY, se = 0.f;
    int am = 0x7fffffff;
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      const float v = __shfl_sync(kFull, sc, u_l * kW + w);
      if (v > m || (v == m && w < am)) { m = v; am = w; }
      if (v != v && am == 0x7fffffff) am = w;  // NaN row: keep something defined
    }
    const float ex = expf(sc - m);
#pragma unroll
    for (int w = 0; w < kW; ++w) se += __shfl_sync(kFull, ex, u_l * kW + w);
    if (!kBwd) {
      if (live_l) {
        if (p.scores) p.scores[(size_t)(r0 + i0) * kW + lane] = sc;      // (r0+i0)*W + u*W + w: coalesced
        if (w_l == 0) {
          if (p.pred) p.pred[r0 + i_l] = am;
          if (p.posterior) p.posterior[r0 + i_l] = m;
          hit += (am == y_l);
        }
        if (p.loss && w_l == y_l) nll += -((sc - m) - logf(se));         // log_softmax then NLL, on the label's lane
      }
    } else {
      // dL/dscore = dl*(softmax - onehot) + d_scores ; score = -dist => dL/ddist = -dL/dscore
      // coef = dL/ddist / dist, zero where dist == 0 (cdist backward convention) and in replayed slots
      const float g = dl * (ex / se - (w_l == y_l ? 1.f : 0.f)) + dsc_l;
      const float dist = -sc;
      const float cf_l = (live_l && dist > 0.f) ? -g / dist : 0.f;
      float* dq = p.d_queries + (size_t)r0 * kD;
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        f32x2 out[kH];
#pragma unroll
        for (int j = 0; j < kH; ++j) out[j] = 0ull;
#pragma unroll
        for (int w = 0; w < kW; ++w) {
          const float cf = __shfl_sync(kFull, cf_l, u * kW + w);
          const f32x2 c2 = pack2(cf, cf);
          f32x2 pw[kH];
          lds_row<kV>(sp + w * kD, lane, pw);
#pragma unroll
          for (int j = 0; j < kH; ++j) {
            const f32x2 d = sub2(buf[u][j], pw[j]);
            out[j] = fma2(c2, d, out[j]);
            dP[w][j] = fma2(c2, d, dP[w][j]);   // rows in ascending order
          }
        }
        if (i0 + u < nrows) store_row<kV>(dq + (size_t)(i0 + u) * kD, lane, out);
        if (t0 + u + kB < total) load_row<kV>(row_ptr(t0 + u + kB), lane, buf[u]);
      }
    }
  }

  if (!kBwd) {
    if (p.queries && (p.loss || p.correct)) {
      nll = warp_sum(nll);
      hit = __reduce_add_sync(kFull, hit);
      if (lane == 0) {
        if (p.loss) p.loss[e] = nll / (float)nrows;
        if (p.correct) p.correct[e] = hit;
      }
    }
    return;
  }

  // ------------------------------------------------------------------ backward epilogue
  // dP[w] = -sum_i coef[i,w] (q_i - p_w) (+ the gradient arriving from the other consumer of the prototypes)
#pragma unroll
  for (int w = 0; w < (kBwd ? kW : 1); ++w) {
    f32x2 x[kH];
    if (p.d_protos_extra) load_row<kV>(p.d_protos_extra + ((size_t)e * kW + w) * kD, lane, x);
#pragma unroll
    for (int j = 0; j < kH; ++j) {
      float a, b;
      unpack2(dP[w][j], a, b);
      dP[w][j] = pack2(-a, -b);
      if (p.d_protos_extra) dP[w][j] = add2(dP[w][j], x[j]);
    }
    if (p.d_protos) store_row<kV>(p.d_protos + ((size_t)e * kW + w) * kD, lane, dP[w]);
  }
  // dS[k] = dP[label_k] / count[label_k]  (mean backward); rows without a prototype get zero
  if (p.d_support) {
    float* ds = p.d_support + (size_t)e * p.Ns * kD;
#pragma unroll
    for (int w = 0; w < (kBwd ? kW : 1); ++w) {
      const float fn = (float)cnt[w];
      f32x2 g[kH];
#pragma unroll
      for (int j = 0; j < kH; ++j) {
        float a, b;
        unpack2(dP[w][j], a, b);
        g[j] = pack2(__fdiv_rn(a, fn), __fdiv_rn(b, fn));
      }
      if (sorted) {
#pragma unroll
        for (int u = 0; u < kB; ++u) store_row<kV>(ds + (size_t)(w * kB + u) * kD, lane, g);
      } else {
        for (int k = 0; k < p.Ns; ++k)
          if (slab[k] == w) store_row<kV>(ds + (size_t)k * kD, lane, g);
      }
    }
    if (!sorted) {
      f32x2 zero[kH];
#pragma unroll
      for (int j = 0; j < kH; ++j) zero[j] = 0ull;
      for (int k = 0; k < p.Ns; ++k) {
        const int l = slab[k];
        if (l < 0 || l >= kW) store_row<kV>(ds + (size_t)k * kD, lane, zero);
      }
    }
  }
}

using KernelFn = void (*)(const HeadParams);

template <int kW, int kV>
void variant(bool bwd, KernelFn& fn) {
  constexpr int kB = 32 / kW < 6 ? 32 / kW : 6;          // rows per batch: kB*W values fit the 32 lanes
  constexpr int kBB = (kW == 5) ? 5 : kB;                // 5-way: batches of 5 divide the 25-row blocks evenly
  fn = bwd ? head_warp_kernel<kW, kV, kBB, true> : head_warp_kernel<kW, kV, kBB, false>;
}

bool pick_variant(int W, int D, bool bwd, KernelFn& fn) {
#define AFSL_WV(W_, D_)                  \
  if (W == W_ && D == D_) {              \
    variant<W_, D_ / 32>(bwd, fn);       \
    return true;                         \
  }
  AFSL_WV(5, 256) AFSL_WV(5, 128) AFSL_WV(5, 64)
  AFSL_WV(2, 256) AFSL_WV(3, 256) AFSL_WV(4, 256)
  AFSL_WV(10, 128) AFSL_WV(10, 64)
#undef AFSL_WV
  return false;
}

}  // namespace

int launch_head_warp(const HeadParams& p, bool bwd, cudaStream_t stream, const char* name, bool* handled) {
  *handled = false;
  KernelFn fn = nullptr;
  if (p.Ns > kMaxSupport || !pick_variant(p.W, p.D, bwd, fn)) return AFSL_OK;
  *handled = true;
  const int grid = (p.E + kWarpsPerCta - 1) / kWarpsPerCta;
  fn<<<grid, kCtaThreads, 0, stream>>>(p);
  AFSL_CHECK_LAUNCH(name);
  return AFSL_OK;
}

}  // namespace afsl
