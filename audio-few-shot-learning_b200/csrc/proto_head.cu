// Prototype head kernels: per-label mean of support embeddings, query->prototype Euclidean
// distances, log-softmax / NLL, argmax / accuracy, and the fused backward.
//
// Reference semantics: models/util_functions.py:6-19 (compute_prototypes),
// models/few_shot_classifier.py:108-116 (-cdist), loops/loss.py:24-37 (FSL_Loss),
// loops/loops.py:79,271-272 (argmax, posterior, #correct).
//
// One CTA owns one episode at a time (persistent grid-stride loop over E).  Every support /
// query element is read from HBM exactly once per pass with 128-bit loads; prototypes live in
// shared memory; each query row is handled by a group of kLPR lanes holding the row in
// registers, with warp-shuffle reductions over the embedding dimension.  fp32 throughout
// (1e-5 parity bar; the contractions are ~1 flop/B, i.e. HBM-bound, so no tensor cores).
#include <cstdlib>

#include "afsl_common.cuh"
#include "proto_head.cuh"

namespace afsl {
namespace {

constexpr int kThreads = 64;
constexpr int kWarps = kThreads / kWarp;
constexpr bool kStagedByDefault = false;  // measured on B200: occupancy (8 CTAs/SM) beats staging (2 CTAs/SM) at 5w5s5q


// shared memory carve-up (floats/ints are both 4 bytes)
struct Smem {
  float* protos;   // [W*D]
  float* dprotos;  // [W*D]      (backward only)
  float* coef;     // [Nq*W]     (backward only)
  float* score;    // [slots*W]
  float* part;     // [slots]
  int* lab;        // [Ns]
  int* row;        // [Ns]
  int* cnt;        // [W]
  int* start;      // [W]
  int* correct;    // [1]
};

__host__ __device__ inline size_t smem_words(int Ns, int Nq, int W, int D, int slots, bool bwd) {
  size_t n = (size_t)W * D + (size_t)slots * W + slots + 2 * (size_t)Ns + 2 * (size_t)W + 4;
  if (bwd) n += (size_t)W * D + (size_t)Nq * W;
  return n;
}

__device__ inline Smem carve(float* base, int Ns, int Nq, int W, int D, int slots, bool bwd) {
  Smem s;
  s.protos = base;
  base += (size_t)W * D;
  s.dprotos = base;
  if (bwd) base += (size_t)W * D;
  s.coef = base;
  if (bwd) base += (size_t)Nq * W;
  s.score = base;
  base += (size_t)slots * W;
  s.part = base;
  base += slots;
  s.lab = reinterpret_cast<int*>(base);
  base += Ns;
  s.row = reinterpret_cast<int*>(base);
  base += Ns;
  s.cnt = reinterpret_cast<int*>(base);
  base += W;
  s.start = reinterpret_cast<int*>(base);
  base += W;
  s.correct = reinterpret_cast<int*>(base);
  return s;
}

__device__ inline void bucket_rows(const Smem& s, const int32_t* labels, int Ns, int W) {
  bucket_by_label(labels, Ns, W, s.lab, s.row, s.cnt, s.start);
}

// kStaged: the episode's support / query blocks were brought into shared memory by a TMA bulk copy, so
// they are read with plain shared loads; otherwise they are streamed from global memory.
template <bool kStaged>
__device__ __forceinline__ float4 ld_row(const float4* p) {
  if (kStaged) return *p;
  return ldg_stream(p);
}
template <bool kStaged>
__device__ __forceinline__ float4 ld_row_cached(const float4* p) {
  if (kStaged) return *p;
  return __ldg(p);
}

// prototypes of one episode -> shared memory (and global when requested)
template <bool kStaged>
__device__ inline void build_prototypes(const Smem& s, const float* support, float* protos_out, int W, int D) {
  const int D4 = D >> 2;
  const float4* sup4 = reinterpret_cast<const float4*>(support);
  float4* sp4 = reinterpret_cast<float4*>(s.protos);
  float4* out4 = reinterpret_cast<float4*>(protos_out);
  for (int item = threadIdx.x; item < W * D4; item += kThreads) {
    const int w = item / D4, c = item - w * D4;
    const int n = s.cnt[w];
    const int* rows = s.row + s.start[w];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int j = 0;
    for (; j + 4 <= n; j += 4) {  // 4 independent 128-bit loads in flight
      const float4 a = ld_row<kStaged>(sup4 + (size_t)rows[j] * D4 + c);
      const float4 b = ld_row<kStaged>(sup4 + (size_t)rows[j + 1] * D4 + c);
      const float4 d = ld_row<kStaged>(sup4 + (size_t)rows[j + 2] * D4 + c);
      const float4 g = ld_row<kStaged>(sup4 + (size_t)rows[j + 3] * D4 + c);
      acc.x = (((acc.x + a.x) + b.x) + d.x) + g.x;
      acc.y = (((acc.y + a.y) + b.y) + d.y) + g.y;
      acc.z = (((acc.z + a.z) + b.z) + d.z) + g.z;
      acc.w = (((acc.w + a.w) + b.w) + d.w) + g.w;
    }
    for (; j < n; ++j) {
      const float4 a = ld_row<kStaged>(sup4 + (size_t)rows[j] * D4 + c);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    const float fn = (float)n;  // n == 0 -> NaN, as the reference's empty mean
    acc.x = __fdiv_rn(acc.x, fn); acc.y = __fdiv_rn(acc.y, fn);
    acc.z = __fdiv_rn(acc.z, fn); acc.w = __fdiv_rn(acc.w, fn);
    sp4[item] = acc;
    if (protos_out) out4[item] = acc;
  }
}

__device__ inline void load_prototypes(const Smem& s, const float* protos_in, int W, int D) {
  const int n4 = (W * D) >> 2;
  const float4* in4 = reinterpret_cast<const float4*>(protos_in);
  float4* sp4 = reinterpret_cast<float4*>(s.protos);
  for (int i = threadIdx.x; i < n4; i += kThreads) sp4[i] = __ldg(in4 + i);
}

// Squared distances of the register-resident row q[] to kWB consecutive prototypes, reduced over the
// lane group.  The prototype block is unrolled so the row is read once per block, the arithmetic is
// packed (FADD2 + FFMA2: the same IEEE sub / fma as the scalar form, two per instruction).
template <int kLPR, int kCPL, int kWB>
__device__ __forceinline__ void row_dist2_block(const float4 (&q)[kCPL], const float4* sp4, int w0, int D4, int sub,
                                                float (&d2)[kWB]) {
  f32x2 acc[kWB][2];
#pragma unroll
  for (int b = 0; b < kWB; ++b) acc[b][0] = acc[b][1] = 0ull;
#pragma unroll
  for (int u = 0; u < kCPL; ++u) {
    const f32x2 qa = pack2(q[u].x, q[u].y), qb = pack2(q[u].z, q[u].w);
#pragma unroll
    for (int b = 0; b < kWB; ++b) {
      const float4 pr = sp4[(w0 + b) * D4 + sub + u * kLPR];
      const f32x2 da = sub2(qa, pack2(pr.x, pr.y)), db = sub2(qb, pack2(pr.z, pr.w));
      acc[b][0] = fma2(da, da, acc[b][0]);
      acc[b][1] = fma2(db, db, acc[b][1]);
    }
  }
#pragma unroll
  for (int b = 0; b < kWB; ++b) d2[b] = group_sum<kLPR>(sum2(acc[b][0]) + sum2(acc[b][1]));
}

// Scores of one row against all prototypes -> s.score[slot*W ..]; returns (max, argmax, sumexp)
// computed cooperatively by the lane group.  W must be a multiple of kWB.
template <int kLPR, int kCPL, int kWB>
__device__ __forceinline__ void row_scores(const float4 (&q)[kCPL], const Smem& s, int slot, int W, int D4, int sub,
                                           float& mx, int& amx, float& sumexp) {
  const float4* sp4 = reinterpret_cast<const float4*>(s.protos);
  float* sc = s.score + slot * W;
  for (int w0 = 0; w0 < W; w0 += kWB) {
    float d2[kWB];
    row_dist2_block<kLPR, kCPL, kWB>(q, sp4, w0, D4, sub, d2);
    if (sub == 0) {
#pragma unroll
      for (int b = 0; b < kWB; ++b) sc[w0 + b] = -sqrtf(d2[b]);
    }
  }
  __syncwarp();
  // max / first argmax over W
  float m = -INFINITY;
  int am = 0x7fffffff;
  for (int w = sub; w < W; w += kLPR) {
    const float v = sc[w];
    if (v > m || (v == m && w < am)) { m = v; am = w; }
    if (v != v && am == 0x7fffffff) am = w;  // NaN row: keep something defined
  }
#pragma unroll
  for (int o = kLPR / 2; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oa < am)) { m = om; am = oa; }
  }
  float se = 0.f;
  for (int w = sub; w < W; w += kLPR) se += expf(sc[w] - m);
  se = group_sum<kLPR>(se);
  mx = m; amx = am; sumexp = se;
}

// ---- TMA staging: double-buffered bulk copies of the episode's support / query blocks
struct Staging {
  float* buf;       // [2][(Ns + Nq) * D]
  uint64_t* bars;   // [2]
  int sup_floats, qry_floats;
};

__host__ __device__ inline size_t align4(size_t words) { return (words + 3) & ~(size_t)3; }

__device__ inline Staging carve_staging(float* smem_raw, size_t base_words, int Ns, int Nq, int D) {
  Staging g;
  g.buf = smem_raw + align4(base_words);
  g.sup_floats = Ns * D;
  g.qry_floats = Nq * D;
  g.bars = reinterpret_cast<uint64_t*>(g.buf + 2 * (size_t)(g.sup_floats + g.qry_floats));
  return g;
}

__device__ inline void query_rows(const HeadParams& p, int e, int& r0, int& nrows) {
  r0 = p.q_offsets ? p.q_offsets[e] : e * p.Nq;
  nrows = p.q_offsets ? p.q_offsets[e + 1] - r0 : p.Nq;
}

// one thread: start the bulk copies of episode e into stage st
__device__ inline void stage_issue(const HeadParams& p, const Staging& g, int e, int st) {
  float* dst = g.buf + (size_t)st * (g.sup_floats + g.qry_floats);
  uint32_t sup_bytes = p.support ? (uint32_t)g.sup_floats * 4u : 0u;
  uint32_t qry_bytes = 0;
  int r0 = 0, nrows = 0;
  if (p.queries) {
    query_rows(p, e, r0, nrows);
    qry_bytes = (uint32_t)nrows * (uint32_t)p.D * 4u;
  }
  mbar_arrive_expect_tx(&g.bars[st], sup_bytes + qry_bytes);
  if (sup_bytes) bulk_copy_g2s(dst, p.support + (size_t)e * g.sup_floats, sup_bytes, &g.bars[st]);
  if (qry_bytes) bulk_copy_g2s(dst + g.sup_floats, p.queries + (size_t)r0 * p.D, qry_bytes, &g.bars[st]);
}

template <int kLPR, int kCPL, int kWB, bool kStaged>
__global__ void __launch_bounds__(kThreads) head_fwd_kernel(const HeadParams p) {
  extern __shared__ __align__(16) float smem_raw[];
  constexpr int kGroups = kWarp / kLPR;       // rows handled concurrently by one warp
  constexpr int kSlots = kWarps * kGroups;
  const Smem s = carve(smem_raw, p.Ns, p.Nq, p.W, p.D, kSlots, false);
  const Staging g = carve_staging(smem_raw, smem_words(p.Ns, p.Nq, p.W, p.D, kSlots, false), p.Ns, p.Nq, p.D);
  const int D4 = p.D >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (kLPR - 1), grp = lane / kLPR;
  const int slot = warp * kGroups + grp;
  if (kStaged) {
    if (threadIdx.x == 0) {
      mbar_init(&g.bars[0], 1);
      mbar_init(&g.bars[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < p.E) stage_issue(p, g, blockIdx.x, 0);
  }

  int it = 0;
  for (int e = blockIdx.x; e < p.E; e += gridDim.x, ++it) {
    const int st = it & 1;
    const float* stage = g.buf + (size_t)st * (g.sup_floats + g.qry_floats);
    if (kStaged && threadIdx.x == 0 && e + (int)gridDim.x < p.E) {
      fence_async_proxy();               // the other stage was last read before the barrier ending the previous episode
      stage_issue(p, g, e + gridDim.x, st ^ 1);
    }
    if (threadIdx.x == 0) *s.correct = 0;
    if (p.support) {
      bucket_rows(s, p.s_labels + (size_t)e * p.Ns, p.Ns, p.W);
      if (kStaged) mbar_wait(&g.bars[st], (it >> 1) & 1);
      build_prototypes<kStaged>(s, kStaged ? stage : p.support + (size_t)e * p.Ns * p.D,
                                p.protos_out ? p.protos_out + (size_t)e * p.W * p.D : nullptr, p.W, p.D);
    } else {
      load_prototypes(s, p.protos_in + (size_t)e * p.W * p.D, p.W, p.D);
      if (kStaged) mbar_wait(&g.bars[st], (it >> 1) & 1);
    }
    __syncthreads();
    if (p.queries) {
      int r0, nrows;
      query_rows(p, e, r0, nrows);
      const float4* q4 = kStaged ? reinterpret_cast<const float4*>(stage + g.sup_floats)
                                 : reinterpret_cast<const float4*>(p.queries) + (size_t)r0 * D4;
      float nll_acc = 0.f;
      int hit = 0;
      // every lane group walks rows slot, slot+kSlots, ...; the trip count is warp-uniform
      const int trips = (nrows + kSlots - 1) / kSlots;
      for (int it = 0; it < trips; ++it) {
        const int i = it * kSlots + slot;
        const bool live = i < nrows;
        const int ii = live ? i : nrows - 1;  // clamp: dead groups redo the last row, results discarded
        float4 q[kCPL];
#pragma unroll
        for (int u = 0; u < kCPL; ++u) q[u] = ld_row<kStaged>(q4 + (size_t)ii * D4 + sub + u * kLPR);
        float mx, se;
        int am;
        row_scores<kLPR, kCPL, kWB>(q, s, slot, p.W, D4, sub, mx, am, se);
        if (live) {
          const float* sc = s.score + slot * p.W;
          if (p.scores)
            for (int w = sub; w < p.W; w += kLPR) p.scores[(size_t)(r0 + i) * p.W + w] = sc[w];
          if (sub == 0) {
            if (p.pred) p.pred[r0 + i] = am;
            if (p.posterior) p.posterior[r0 + i] = mx;
            if (p.q_labels) {
              const int y = p.q_labels[r0 + i];
              if (y >= 0 && y < p.W) nll_acc += -((sc[y] - mx) - logf(se));  // log_softmax then NLL
              hit += (am == y);
            }
          }
        }
        __syncwarp();
      }
      if (sub == 0) {
        s.part[slot] = nll_acc;
        if (hit) atomicAdd(s.correct, hit);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        if (p.loss) {
          float tot = 0.f;
          for (int k = 0; k < kSlots; ++k) tot += s.part[k];
          p.loss[e] = tot / (float)nrows;
        }
        if (p.correct) p.correct[e] = *s.correct;
      }
    }
    __syncthreads();
  }
}

template <int kLPR, int kCPL, int kWB, bool kStaged>
__global__ void __launch_bounds__(kThreads) head_bwd_kernel(const HeadParams p) {
  extern __shared__ __align__(16) float smem_raw[];
  constexpr int kGroups = kWarp / kLPR;
  constexpr int kSlots = kWarps * kGroups;
  const Smem s = carve(smem_raw, p.Ns, p.Nq, p.W, p.D, kSlots, true);
  const Staging g = carve_staging(smem_raw, smem_words(p.Ns, p.Nq, p.W, p.D, kSlots, true), p.Ns, p.Nq, p.D);
  const int D4 = p.D >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (kLPR - 1), grp = lane / kLPR;
  const int slot = warp * kGroups + grp;
  float4* sdp4 = reinterpret_cast<float4*>(s.dprotos);
  const float4* sp4 = reinterpret_cast<const float4*>(s.protos);
  if (kStaged) {
    if (threadIdx.x == 0) {
      mbar_init(&g.bars[0], 1);
      mbar_init(&g.bars[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < p.E) stage_issue(p, g, blockIdx.x, 0);
  }

  int it = 0;
  for (int e = blockIdx.x; e < p.E; e += gridDim.x, ++it) {
    const int st = it & 1;
    const float* stage = g.buf + (size_t)st * (g.sup_floats + g.qry_floats);
    if (kStaged && threadIdx.x == 0 && e + (int)gridDim.x < p.E) {
      fence_async_proxy();
      stage_issue(p, g, e + gridDim.x, st ^ 1);
    }
    if (p.support) {
      bucket_rows(s, p.s_labels + (size_t)e * p.Ns, p.Ns, p.W);
      if (kStaged) mbar_wait(&g.bars[st], (it >> 1) & 1);
      if (p.queries)
        build_prototypes<kStaged>(s, kStaged ? stage : p.support + (size_t)e * p.Ns * p.D, nullptr, p.W, p.D);
    } else if (p.protos_in) {
      if (p.s_labels) bucket_rows(s, p.s_labels + (size_t)e * p.Ns, p.Ns, p.W);   // saved prototypes + dS scatter
      load_prototypes(s, p.protos_in + (size_t)e * p.W * p.D, p.W, p.D);
      if (kStaged) mbar_wait(&g.bars[st], (it >> 1) & 1);
    } else {
      bucket_rows(s, p.s_labels + (size_t)e * p.Ns, p.Ns, p.W);  // prototype-only backward
    }
    __syncthreads();
    int nrows = 0, r0 = 0;
    const float4* q4 = nullptr;
    if (p.queries) {
      query_rows(p, e, r0, nrows);
      q4 = kStaged ? reinterpret_cast<const float4*>(stage + g.sup_floats)
                   : reinterpret_cast<const float4*>(p.queries) + (size_t)r0 * D4;
      float4* dq4 = reinterpret_cast<float4*>(p.d_queries) + (size_t)r0 * D4;
      const float dl = p.d_loss ? p.d_loss[e] / (float)nrows : 0.f;
      const int trips = (nrows + kSlots - 1) / kSlots;
      for (int it = 0; it < trips; ++it) {
        const int i = it * kSlots + slot;
        const bool live = i < nrows;
        const int ii = live ? i : nrows - 1;
        float4 q[kCPL];
#pragma unroll
        for (int u = 0; u < kCPL; ++u) q[u] = ld_row_cached<kStaged>(q4 + (size_t)ii * D4 + sub + u * kLPR);
        float mx, se;
        int am;
        row_scores<kLPR, kCPL, kWB>(q, s, slot, p.W, D4, sub, mx, am, se);
        const float* sc = s.score + slot * p.W;
        float* cf = s.coef + (size_t)ii * p.W;
        const int y = p.q_labels ? p.q_labels[r0 + ii] : -1;
        if (live) {
          // dL/dscore = dl*(softmax - onehot) + d_scores ;  score = -dist  =>  dL/ddist = -dL/dscore
          // coef = dL/ddist / dist, zero where dist == 0 (cdist backward convention)
          for (int w = sub; w < p.W; w += kLPR) {
            float g = dl * (expf(sc[w] - mx) / se - (w == y ? 1.f : 0.f));
            if (p.d_scores) g += p.d_scores[(size_t)(r0 + i) * p.W + w];
            const float dist = -sc[w];
            cf[w] = dist > 0.f ? -g / dist : 0.f;
          }
        }
        __syncwarp();
        if (live) {
          f32x2 acc[kCPL][2];
#pragma unroll
          for (int u = 0; u < kCPL; ++u) acc[u][0] = acc[u][1] = 0ull;
          for (int w0 = 0; w0 < p.W; w0 += kWB) {
#pragma unroll
            for (int b = 0; b < kWB; ++b) {
              const float c = cf[w0 + b];
              const f32x2 c2 = pack2(c, c);
#pragma unroll
              for (int u = 0; u < kCPL; ++u) {
                const float4 pr = sp4[(w0 + b) * D4 + sub + u * kLPR];
                acc[u][0] = fma2(c2, sub2(pack2(q[u].x, q[u].y), pack2(pr.x, pr.y)), acc[u][0]);
                acc[u][1] = fma2(c2, sub2(pack2(q[u].z, q[u].w), pack2(pr.z, pr.w)), acc[u][1]);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < kCPL; ++u) {
            float4 o;
            unpack2(acc[u][0], o.x, o.y);
            unpack2(acc[u][1], o.z, o.w);
            stg_stream(dq4 + (size_t)i * D4 + sub + u * kLPR, o);
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // dP[w] = -sum_i coef[i,w] (q_i - p_w)  (+ extra), rows in ascending order (deterministic)
    if (p.queries || p.d_protos_extra) {
      const float4* ex4 = p.d_protos_extra ? reinterpret_cast<const float4*>(p.d_protos_extra) + (size_t)e * p.W * D4 : nullptr;
      float4* dpo4 = p.d_protos ? reinterpret_cast<float4*>(p.d_protos) + (size_t)e * p.W * D4 : nullptr;
      for (int item = threadIdx.x; item < p.W * D4; item += kThreads) {
        const int w = item / D4, c = item - w * D4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q4) {
          const float4 pr = sp4[item];
          const f32x2 pa = pack2(pr.x, pr.y), pb = pack2(pr.z, pr.w);
          f32x2 a0 = 0ull, a1 = 0ull;
#pragma unroll 5
          for (int i = 0; i < nrows; ++i) {
            const float cf = -s.coef[(size_t)i * p.W + w];
            const f32x2 c2 = pack2(cf, cf);
            const float4 qv = ld_row_cached<kStaged>(q4 + (size_t)i * D4 + c);
            a0 = fma2(c2, sub2(pack2(qv.x, qv.y), pa), a0);
            a1 = fma2(c2, sub2(pack2(qv.z, qv.w), pb), a1);
          }
          unpack2(a0, acc.x, acc.y);
          unpack2(a1, acc.z, acc.w);
        }
        if (ex4) {
          const float4 x = __ldg(ex4 + item);
          acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
        }
        sdp4[item] = acc;
        if (dpo4) dpo4[item] = acc;
      }
    } else if (p.d_protos_extra == nullptr && p.protos_in == nullptr && !p.queries) {
      // unreachable: prototype-only backward always passes d_protos_extra
    }
    __syncthreads();
    // dS[k] = dP[label_k] / count[label_k]   (mean backward)
    if (p.d_support) {
      float4* ds4 = reinterpret_cast<float4*>(p.d_support) + (size_t)e * p.Ns * D4;
      for (int item = threadIdx.x; item < p.Ns * D4; item += kThreads) {
        const int k = item / D4, c = item - k * D4;
        const int w = s.lab[k];
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (w >= 0 && w < p.W) {
          g = sdp4[w * D4 + c];
          const float fn = (float)s.cnt[w];
          g.x = __fdiv_rn(g.x, fn); g.y = __fdiv_rn(g.y, fn); g.z = __fdiv_rn(g.z, fn); g.w = __fdiv_rn(g.w, fn);
        }
        stg_stream(ds4 + item, g);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ host dispatch
using KernelFn = void (*)(const HeadParams);

struct Variant {
  KernelFn fwd, bwd, fwd_staged, bwd_staged;
  int slots;
};

template <int kLPR, int kCPL>
void fill_variant(int W, Variant& v) {
  v.slots = kWarps * (kWarp / kLPR);
#define AFSL_WB(WB)                                          \
  {                                                          \
    v.fwd = head_fwd_kernel<kLPR, kCPL, WB, false>;          \
    v.bwd = head_bwd_kernel<kLPR, kCPL, WB, false>;          \
    v.fwd_staged = head_fwd_kernel<kLPR, kCPL, WB, true>;    \
    v.bwd_staged = head_bwd_kernel<kLPR, kCPL, WB, true>;    \
  }
  if (W % 5 == 0) AFSL_WB(5) else if (W % 4 == 0) AFSL_WB(4) else AFSL_WB(1)
#undef AFSL_WB
}

// lane group per query row: 8 lanes up to D = 256 (4 rows per warp), wider groups beyond
bool pick_variant(int D, int W, Variant& v) {
  switch (D) {
    case 16: fill_variant<4, 1>(W, v); return true;
    case 32: fill_variant<8, 1>(W, v); return true;
    case 64:      // many-way: 4 lanes per row (8 rows per warp, two shuffle steps per distance instead of three)
      if (W >= 8) fill_variant<4, 4>(W, v); else fill_variant<8, 2>(W, v);
      return true;
    case 128: fill_variant<8, 4>(W, v); return true;
    case 256: fill_variant<8, 8>(W, v); return true;
    case 512: fill_variant<16, 8>(W, v); return true;
    case 1024: fill_variant<32, 8>(W, v); return true;
    default: return false;
  }
}

int launch(const HeadParams& p, bool bwd, cudaStream_t stream, const char* name) {
  AFSL_REQUIRE(p.E >= 0 && p.W > 0 && p.D > 0, "%s: bad sizes E=%d W=%d D=%d", name, p.E, p.W, p.D);
  if (p.E == 0) return AFSL_OK;
  // small W*D: one warp per episode, prototypes in registers (proto_head_warp.cu); AFSL_HEAD_WARP=0 disables it
  const char* warp_env = getenv("AFSL_HEAD_WARP");     // read per launch so the tests can exercise both paths
  if (!warp_env || atoi(warp_env) != 0) {
    bool handled = false;
    const int rc = launch_head_warp(p, bwd, stream, name, &handled);
    if (rc != AFSL_OK || handled) return rc;
  }
  {   // many-way forward of fixed-size tasks: TMA-fed ring, q.p on the tcgen05 tensor cores (proto_head_tma.cu)
    bool handled = false;
    int rc = launch_head_tma(p, bwd, stream, name, &handled);
    if (rc != AFSL_OK || handled) return rc;
    rc = launch_head_mma(p, bwd, stream, name, &handled);      // its LDG-fed predecessor, AFSL_HEAD_MMA=2 only
    if (rc != AFSL_OK || handled) return rc;
  }
  {   // many-way forward: register batches of query rows against shared-memory prototypes (proto_head_wide.cu)
    bool handled = false;
    const int rc = launch_head_wide(p, bwd, stream, name, &handled);
    if (rc != AFSL_OK || handled) return rc;
  }
  Variant v;
  AFSL_REQUIRE(pick_variant(p.D, p.W, v), "%s: unsupported embedding dim D=%d (supported: 16,32,64,128,256,512,1024)", name, p.D);
  size_t bytes = smem_words(p.Ns, p.Nq, p.W, p.D, v.slots, bwd) * sizeof(float);
  AFSL_REQUIRE(bytes <= 220 * 1024, "%s: episode does not fit shared memory (W=%d D=%d Ns=%d Nq=%d -> %zu B)", name,
               p.W, p.D, p.Ns, p.Nq, bytes);
  KernelFn fn = bwd ? v.bwd : v.fwd;
  // TMA-staged variant: double-buffered bulk copies of the support / query blocks, when two stages fit
  // next to the working set with room for at least two CTAs per SM
  const size_t staged_bytes = align4(bytes / sizeof(float)) * sizeof(float) +
                              2 * (size_t)(p.Ns + p.Nq) * p.D * sizeof(float) + 2 * sizeof(uint64_t);
  const bool has_blocks = (p.support != nullptr || p.queries != nullptr) && (p.Ns + p.Nq) > 0;
  static const int staged_mode = [] {   // AFSL_HEAD_STAGED=0 / 1 overrides the default choice (measurement switch)
    const char* e = getenv("AFSL_HEAD_STAGED");
    return e ? atoi(e) : -1;
  }();
  const bool want_staged = staged_mode < 0 ? kStagedByDefault : staged_mode != 0;
  if (want_staged && has_blocks && staged_bytes <= 110 * 1024 && (size_t)(p.Ns + p.Nq) * p.D * 4 < (1u << 20)) {
    fn = bwd ? v.bwd_staged : v.fwd_staged;
    bytes = staged_bytes;
  }
  if (bytes > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (err != cudaSuccess) {
      set_error("%s: cannot opt in to %zu B shared memory: %s", name, bytes, cudaGetErrorString(err));
      return AFSL_ECUDA;
    }
  }
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, bytes);
  if (per_sm < 1) per_sm = 1;
  int sms = kNumSMs;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.E < sms * per_sm ? p.E : sms * per_sm;
  fn<<<grid, kThreads, bytes, stream>>>(p);
  AFSL_CHECK_LAUNCH(name);
  return AFSL_OK;
}

}  // namespace
}  // namespace afsl

using afsl::HeadParams;

extern "C" int afsl_prototypes_fwd_f32(const float* support, const int32_t* labels, float* protos, int E, int Ns,
                                        int W, int D, void* stream) {
  AFSL_REQUIRE(support && labels && protos, "afsl_prototypes_fwd_f32: null pointer");
  AFSL_REQUIRE(Ns > 0, "afsl_prototypes_fwd_f32: Ns=%d", Ns);
  HeadParams p{};
  p.support = support; p.s_labels = labels; p.protos_out = protos;
  p.E = E; p.Ns = Ns; p.Nq = 0; p.W = W; p.D = D;
  return afsl::launch(p, false, (cudaStream_t)stream, "afsl_prototypes_fwd_f32");
}

extern "C" int afsl_prototypes_bwd_f32(const float* d_protos, const int32_t* labels, float* d_support, int E, int Ns,
                                        int W, int D, void* stream) {
  AFSL_REQUIRE(d_protos && labels && d_support, "afsl_prototypes_bwd_f32: null pointer");
  AFSL_REQUIRE(Ns > 0, "afsl_prototypes_bwd_f32: Ns=%d", Ns);
  HeadParams p{};
  p.s_labels = labels; p.d_protos_extra = d_protos; p.d_support = d_support;
  p.E = E; p.Ns = Ns; p.Nq = 0; p.W = W; p.D = D;
  return afsl::launch(p, true, (cudaStream_t)stream, "afsl_prototypes_bwd_f32");
}

extern "C" int afsl_proto_scores_fwd_f32(const float* protos, const float* queries, const int32_t* q_labels,
                                          const int32_t* q_offsets, float* scores, float* loss, int32_t* pred,
                                          float* posterior, int32_t* correct, int E, int Nq, int W, int D,
                                          void* stream) {
  AFSL_REQUIRE(protos && queries, "afsl_proto_scores_fwd_f32: null pointer");
  AFSL_REQUIRE(q_labels || (!loss && !correct), "afsl_proto_scores_fwd_f32: loss/correct need q_labels");
  AFSL_REQUIRE(Nq > 0, "afsl_proto_scores_fwd_f32: Nq=%d", Nq);
  HeadParams p{};
  p.protos_in = protos; p.queries = queries; p.q_labels = q_labels; p.q_offsets = q_offsets;
  p.scores = scores; p.loss = loss; p.pred = pred; p.posterior = posterior; p.correct = correct;
  p.E = E; p.Ns = 0; p.Nq = Nq; p.W = W; p.D = D;
  return afsl::launch(p, false, (cudaStream_t)stream, "afsl_proto_scores_fwd_f32");
}

extern "C" int afsl_proto_scores_bwd_f32(const float* protos, const float* queries, const int32_t* q_labels,
                                          const int32_t* q_offsets, const float* d_loss, const float* d_scores,
                                          float* d_protos, float* d_queries, int E, int Nq, int W, int D,
                                          void* stream) {
  AFSL_REQUIRE(protos && queries && d_protos && d_queries, "afsl_proto_scores_bwd_f32: null pointer");
  AFSL_REQUIRE(d_loss || d_scores, "afsl_proto_scores_bwd_f32: no incoming gradient");
  AFSL_REQUIRE(!d_loss || q_labels, "afsl_proto_scores_bwd_f32: d_loss needs q_labels");
  AFSL_REQUIRE(Nq > 0, "afsl_proto_scores_bwd_f32: Nq=%d", Nq);
  HeadParams p{};
  p.protos_in = protos; p.queries = queries; p.q_labels = q_labels; p.q_offsets = q_offsets;
  p.d_loss = d_loss; p.d_scores = d_scores; p.d_protos = d_protos; p.d_queries = d_queries;
  p.E = E; p.Ns = 0; p.Nq = Nq; p.W = W; p.D = D;
  return afsl::launch(p, true, (cudaStream_t)stream, "afsl_proto_scores_bwd_f32");
}

extern "C" int afsl_proto_head_fwd_f32(const float* support, const int32_t* s_labels, const float* queries,
                                        const int32_t* q_labels, const int32_t* q_offsets, float* protos, float* scores,
                                        float* loss, int32_t* pred, float* posterior, int32_t* correct, int E, int Ns,
                                        int Nq, int W, int D, void* stream) {
  AFSL_REQUIRE(support && s_labels && queries, "afsl_proto_head_fwd_f32: null pointer");
  AFSL_REQUIRE(q_labels || (!loss && !correct), "afsl_proto_head_fwd_f32: loss/correct need q_labels");
  AFSL_REQUIRE(Ns > 0 && Nq > 0, "afsl_proto_head_fwd_f32: Ns=%d Nq=%d", Ns, Nq);
  HeadParams p{};
  p.support = support; p.s_labels = s_labels; p.queries = queries; p.q_labels = q_labels; p.q_offsets = q_offsets;
  p.protos_out = protos; p.scores = scores; p.loss = loss; p.pred = pred; p.posterior = posterior; p.correct = correct;
  p.E = E; p.Ns = Ns; p.Nq = Nq; p.W = W; p.D = D;
  return afsl::launch(p, false, (cudaStream_t)stream, "afsl_proto_head_fwd_f32");
}

extern "C" int afsl_proto_head_bwd_f32(const float* support, const float* protos, const int32_t* s_labels,
                                        const float* queries, const int32_t* q_labels, const int32_t* q_offsets,
                                        const float* d_loss, const float* d_protos_extra, float* d_support,
                                        float* d_queries, int E, int Ns, int Nq, int W, int D, void* stream) {
  AFSL_REQUIRE((support || protos) && s_labels && queries && q_labels && d_loss && d_support && d_queries,
               "afsl_proto_head_bwd_f32: null pointer");
  AFSL_REQUIRE(Ns > 0 && Nq > 0, "afsl_proto_head_bwd_f32: Ns=%d Nq=%d", Ns, Nq);
  HeadParams p{};
  // the forward's prototypes, when given, replace the re-read of the support block
  p.support = protos ? nullptr : support; p.protos_in = protos;
  p.s_labels = s_labels; p.queries = queries; p.q_labels = q_labels; p.q_offsets = q_offsets;
  p.d_loss = d_loss; p.d_protos_extra = d_protos_extra; p.d_support = d_support; p.d_queries = d_queries;
  p.E = E; p.Ns = Ns; p.Nq = Nq; p.W = W; p.D = D;
  return afsl::launch(p, true, (cudaStream_t)stream, "afsl_proto_head_bwd_f32");
}
