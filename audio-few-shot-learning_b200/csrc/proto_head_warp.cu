// Prototype head, one WARP per episode, everything in registers.
//
// Same arithmetic as proto_head.cu (reference: models/util_functions.py:6-19, few_shot_classifier.py:108-116,
// loops/loss.py:24-37, loops/loops.py:79,271-272) for the shapes where the episode's prototypes fit the
// register file: W*D <= 1280 floats (5-way at D <= 256, 20-way at D = 64).  Lane l owns D/32 columns of
// every row, so
//   * prototypes P[w] and the prototype gradient dP[w] are register arrays (no shared memory, no CTA barrier);
//   * a support / query row is one or two fully coalesced 128-bit loads per lane (512 B per warp instruction),
//     software-prefetched kPF rows ahead, read from HBM exactly once;
//   * the backward is ONE pass over the query rows: the softmax coefficients of a row are known while the row
//     is still in registers, so dQ is stored and dP accumulated (ascending rows, deterministic) without a
//     second read; with the forward's prototypes passed back in, the support block is not re-read at all.
// The grid is one warp per episode (4 per CTA): the hardware CTA scheduler balances the tail.
#include "proto_head.cuh"

namespace afsl {
namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kCtaThreads = kWarpsPerCta * kWarp;
constexpr int kMaxSupport = 128;   // support labels of one episode staged in shared memory
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

// Walk rows 0..nrows-1 of a [nrows, D] block with kPF rows of loads in flight; body(i, row registers).
template <int kV, int kPF, typename Body>
__device__ __forceinline__ void stream_rows(const float* base, int nrows, int lane, Body&& body) {
  constexpr int kH = kV / 2, kD = kV * 32;
  if (nrows <= 0) return;
  f32x2 buf[kPF][kH];
#pragma unroll
  for (int u = 0; u < kPF; ++u) load_row<kV>(base + (size_t)min(u, nrows - 1) * kD, lane, buf[u]);
  for (int i0 = 0; i0 < nrows; i0 += kPF) {
#pragma unroll
    for (int u = 0; u < kPF; ++u) {
      const int i = i0 + u;
      f32x2 cur[kH];
#pragma unroll
      for (int j = 0; j < kH; ++j) cur[j] = buf[u][j];
      if (i + kPF < nrows) load_row<kV>(base + (size_t)(i + kPF) * kD, lane, buf[u]);
      if (i < nrows) body(i, cur);
    }
  }
}

template <int kW>
__device__ __forceinline__ void warp_sum_all(float (&v)[kW]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int w = 0; w < kW; ++w) v[w] += __shfl_xor_sync(kFull, v[w], o);
}

// support labels -> shared (one slice per warp) and per-class counts
template <int kW>
__device__ __forceinline__ void stage_labels(const int32_t* labels, int Ns, int lane, int* slab, int (&cnt)[kW]) {
#pragma unroll
  for (int w = 0; w < kW; ++w) cnt[w] = 0;
  for (int k0 = 0; k0 < Ns; k0 += 32) {
    const int k = k0 + lane;
    const int l = k < Ns ? labels[k] : -1;
    if (k < Ns) slab[k] = l;
#pragma unroll
    for (int w = 0; w < kW; ++w) cnt[w] += __popc(__ballot_sync(kFull, l == w));
  }
  __syncwarp();
}

// prototypes = per-label mean of the support rows (rows added in ascending order, then one IEEE division)
template <int kW, int kV, int kPF>
__device__ __forceinline__ void build_prototypes(const float* support, int Ns, int lane, const int* slab, const int (&cnt)[kW],
                                                 f32x2 (&P)[kW][kV / 2]) {
  constexpr int kH = kV / 2;
#pragma unroll
  for (int w = 0; w < kW; ++w)
#pragma unroll
    for (int j = 0; j < kH; ++j) P[w][j] = 0ull;
  stream_rows<kV, kPF>(support, Ns, lane, [&](int k, const f32x2(&x)[kH]) {
    const int lk = slab[k];
#pragma unroll
    for (int w = 0; w < kW; ++w)
      if (lk == w) {
#pragma unroll
        for (int j = 0; j < kH; ++j) P[w][j] = add2(P[w][j], x[j]);
      }
  });
#pragma unroll
  for (int w = 0; w < kW; ++w) {
    const float fn = (float)cnt[w];  // empty class -> NaN, as the reference's empty mean
#pragma unroll
    for (int j = 0; j < kH; ++j) {
      float a, b;
      unpack2(P[w][j], a, b);
      P[w][j] = pack2(__fdiv_rn(a, fn), __fdiv_rn(b, fn));
    }
  }
}

template <int kW, int kV>
__device__ __forceinline__ void load_prototypes(const float* protos, int lane, f32x2 (&P)[kW][kV / 2]) {
#pragma unroll
  for (int w = 0; w < kW; ++w) load_row<kV>(protos + (size_t)w * kV * 32, lane, P[w]);
}

// scores of one row: sc[w] = -||q - P[w]||, then (max, first argmax, sum exp) - every lane ends with all of them
template <int kW, int kV>
__device__ __forceinline__ void row_scores(const f32x2 (&q)[kV / 2], const f32x2 (&P)[kW][kV / 2], float (&sc)[kW], float& mx,
                                           int& amx, float& se) {
  constexpr int kH = kV / 2;
#pragma unroll
  for (int w = 0; w < kW; ++w) {
    f32x2 a0 = 0ull, a1 = 0ull;
#pragma unroll
    for (int j = 0; j < kH; ++j) {
      const f32x2 d = sub2(q[j], P[w][j]);
      if (j & 1) a1 = fma2(d, d, a1); else a0 = fma2(d, d, a0);
    }
    sc[w] = kH > 1 ? sum2(a0) + sum2(a1) : sum2(a0);
  }
  warp_sum_all<kW>(sc);
  float m = -INFINITY;
  int am = 0x7fffffff;
#pragma unroll
  for (int w = 0; w < kW; ++w) {
    const float v = -sqrtf(sc[w]);
    sc[w] = v;
    if (v > m || (v == m && w < am)) { m = v; am = w; }
    if (v != v && am == 0x7fffffff) am = w;  // NaN row: keep something defined
  }
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < kW; ++w) s += expf(sc[w] - m);
  mx = m; amx = am; se = s;
}

template <int kW>
__device__ __forceinline__ float pick(const float (&v)[kW], int idx) {
  float r = v[0];
#pragma unroll
  for (int w = 1; w < kW; ++w) r = idx == w ? v[w] : r;
  return r;
}

__device__ __forceinline__ void query_span(const HeadParams& p, int e, int& r0, int& nrows) {
  r0 = p.q_offsets ? p.q_offsets[e] : e * p.Nq;
  nrows = p.q_offsets ? p.q_offsets[e + 1] - r0 : p.Nq;
}

template <int kW, int kV, int kPF>
__global__ void __launch_bounds__(kCtaThreads) head_warp_fwd_kernel(const HeadParams p) {
  __shared__ int slab_all[kWarpsPerCta][kMaxSupport];
  constexpr int kH = kV / 2, kD = kV * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * kWarpsPerCta + warp;
  if (e >= p.E) return;
  int* slab = slab_all[warp];
  f32x2 P[kW][kH];
  if (p.support) {
    int cnt[kW];
    stage_labels<kW>(p.s_labels + (size_t)e * p.Ns, p.Ns, lane, slab, cnt);
    build_prototypes<kW, kV, kPF>(p.support + (size_t)e * p.Ns * kD, p.Ns, lane, slab, cnt, P);
    if (p.protos_out) {
#pragma unroll
      for (int w = 0; w < kW; ++w) store_row<kV>(p.protos_out + ((size_t)e * kW + w) * kD, lane, P[w]);
    }
  } else {
    load_prototypes<kW, kV>(p.protos_in + (size_t)e * kW * kD, lane, P);
  }
  if (!p.queries) return;
  int r0, nrows;
  query_span(p, e, r0, nrows);
  float nll = 0.f;
  int hit = 0, ql = -1;
  stream_rows<kV, kPF>(p.queries + (size_t)r0 * kD, nrows, lane, [&](int i, const f32x2(&q)[kH]) {
    if (p.q_labels && (i & 31) == 0) ql = i + lane < nrows ? p.q_labels[r0 + i + lane] : -1;
    float sc[kW], mx, se;
    int am;
    row_scores<kW, kV>(q, P, sc, mx, am, se);
    if (p.scores && lane < kW) p.scores[(size_t)(r0 + i) * kW + lane] = pick<kW>(sc, lane);
    if (p.pred && lane == 0) p.pred[r0 + i] = am;
    if (p.posterior && lane == 0) p.posterior[r0 + i] = mx;
    if (p.q_labels) {
      const int y = __shfl_sync(kFull, ql, i & 31);
      if (y >= 0 && y < kW) nll += -((pick<kW>(sc, y) - mx) - logf(se));  // log_softmax then NLL
      hit += (am == y);
    }
  });
  if (lane == 0) {
    if (p.loss) p.loss[e] = nll / (float)nrows;
    if (p.correct) p.correct[e] = hit;
  }
}

template <int kW, int kV, int kPF>
__global__ void __launch_bounds__(kCtaThreads) head_warp_bwd_kernel(const HeadParams p) {
  __shared__ int slab_all[kWarpsPerCta][kMaxSupport];
  constexpr int kH = kV / 2, kD = kV * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * kWarpsPerCta + warp;
  if (e >= p.E) return;
  int* slab = slab_all[warp];
  f32x2 P[kW][kH], dP[kW][kH];
  int cnt[kW];
  if (p.s_labels) stage_labels<kW>(p.s_labels + (size_t)e * p.Ns, p.Ns, lane, slab, cnt);
  if (p.queries) {
    if (p.protos_in) load_prototypes<kW, kV>(p.protos_in + (size_t)e * kW * kD, lane, P);
    else build_prototypes<kW, kV, kPF>(p.support + (size_t)e * p.Ns * kD, p.Ns, lane, slab, cnt, P);
  }
#pragma unroll
  for (int w = 0; w < kW; ++w)
#pragma unroll
    for (int j = 0; j < kH; ++j) dP[w][j] = 0ull;
  if (p.queries) {
    int r0, nrows;
    query_span(p, e, r0, nrows);
    const float dl = p.d_loss ? p.d_loss[e] / (float)nrows : 0.f;
    float* dq = p.d_queries + (size_t)r0 * kD;
    int ql = -1;
    stream_rows<kV, kPF>(p.queries + (size_t)r0 * kD, nrows, lane, [&](int i, const f32x2(&q)[kH]) {
      if (p.q_labels && (i & 31) == 0) ql = i + lane < nrows ? p.q_labels[r0 + i + lane] : -1;
      float sc[kW], mx, se;
      int am;
      row_scores<kW, kV>(q, P, sc, mx, am, se);
      const int y = p.q_labels ? __shfl_sync(kFull, ql, i & 31) : -1;
      float dsc = 0.f;
      if (p.d_scores && lane < kW) dsc = p.d_scores[(size_t)(r0 + i) * kW + lane];
      // dL/dscore = dl*(softmax - onehot) + d_scores ; score = -dist => dL/ddist = -dL/dscore
      // coef = dL/ddist / dist, zero where dist == 0 (cdist backward convention)
      f32x2 out[kH];
#pragma unroll
      for (int j = 0; j < kH; ++j) out[j] = 0ull;
#pragma unroll
      for (int w = 0; w < kW; ++w) {
        float g = dl * (expf(sc[w] - mx) / se - (w == y ? 1.f : 0.f));
        if (p.d_scores) g += __shfl_sync(kFull, dsc, w);
        const float dist = -sc[w];
        const float cf = dist > 0.f ? -g / dist : 0.f;
        const f32x2 c2 = pack2(cf, cf), n2 = pack2(-cf, -cf);
#pragma unroll
        for (int j = 0; j < kH; ++j) {
          const f32x2 d = sub2(q[j], P[w][j]);
          out[j] = fma2(c2, d, out[j]);
          dP[w][j] = fma2(n2, d, dP[w][j]);   // dP[w] = -sum_i coef[i,w] (q_i - p_w), rows ascending
        }
      }
      store_row<kV>(dq + (size_t)i * kD, lane, out);
    });
  }
  if (p.d_protos_extra) {
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      f32x2 x[kH];
      load_row<kV>(p.d_protos_extra + ((size_t)e * kW + w) * kD, lane, x);
#pragma unroll
      for (int j = 0; j < kH; ++j) dP[w][j] = add2(dP[w][j], x[j]);
    }
  }
  if (p.d_protos) {
#pragma unroll
    for (int w = 0; w < kW; ++w) store_row<kV>(p.d_protos + ((size_t)e * kW + w) * kD, lane, dP[w]);
  }
  // dS[k] = dP[label_k] / count[label_k]  (mean backward); rows without a prototype get zero
  if (p.d_support) {
    float* ds = p.d_support + (size_t)e * p.Ns * kD;
    f32x2 zero[kH];
#pragma unroll
    for (int j = 0; j < kH; ++j) zero[j] = 0ull;
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      const float fn = (float)cnt[w];
      f32x2 g[kH];
#pragma unroll
      for (int j = 0; j < kH; ++j) {
        float a, b;
        unpack2(dP[w][j], a, b);
        g[j] = pack2(__fdiv_rn(a, fn), __fdiv_rn(b, fn));
      }
      for (int k = 0; k < p.Ns; ++k)
        if (slab[k] == w) store_row<kV>(ds + (size_t)k * kD, lane, g);
    }
    for (int k = 0; k < p.Ns; ++k) {
      const int l = slab[k];
      if (l < 0 || l >= kW) store_row<kV>(ds + (size_t)k * kD, lane, zero);
    }
  }
}

using KernelFn = void (*)(const HeadParams);

template <int kW, int kV>
void variant(bool bwd, KernelFn& fn) {
  constexpr int kPF = kV >= 8 ? 4 : (kV == 4 ? 6 : 8);
  fn = bwd ? head_warp_bwd_kernel<kW, kV, kPF> : head_warp_fwd_kernel<kW, kV, kPF>;
}

bool pick_variant(int W, int D, bool bwd, KernelFn& fn) {
#define AFSL_WV(W_, D_)                  \
  if (W == W_ && D == D_) {              \
    variant<W_, D_ / 32>(bwd, fn);       \
    return true;                         \
  }
  AFSL_WV(5, 256) AFSL_WV(5, 128) AFSL_WV(5, 64)
  AFSL_WV(2, 256) AFSL_WV(3, 256) AFSL_WV(4, 256)
  AFSL_WV(10, 128) AFSL_WV(10, 64) AFSL_WV(20, 64)
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
