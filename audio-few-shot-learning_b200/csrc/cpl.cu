// Contrastive prototype learning (CPL) loss, forward and backward, batched over episodes.
//
// Reference: CPL_Loss.forward / similarity_sampling, loops/loss.py:118-165.  For query i the
// reference gathers up to M sampled queries of every other class plus the query itself, takes
// cosine similarity of each against the prototype of the query's own class (F.cosine_similarity:
// each vector divided by max(norm, 1e-8)), divides by T and applies log-softmax/NLL with the
// query itself as target, then divides the batch mean by Nq again (loss.py:131).  In closed form
//     C[w,j] = <p_w/|p_w|, q_j/|q_j|> / T
//     loss   = 1/Nq^2 * sum_i ( LSE_{j in keep_i} C[y_i,j] - C[y_i,i] )
// where keep_i = {i} U sampled negatives of i.  The sampling happens on the host (torch CPU
// generator, reference draw order) and arrives as a bit mask; keep == NULL means "every query of
// every other class" (M >= per-class count), which needs no randomness.
//
// One CTA per episode; prototypes (normalised) and the W x Nq similarity matrix live in shared
// memory; P and Q are read from HBM once in the forward and twice in the backward (second read of
// Q hits L2).  fp32, warp-shuffle reductions over the embedding dimension.
#include <cstdlib>

#include "afsl_common.cuh"
#include "cpl.cuh"

namespace afsl {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / kWarp;
constexpr float kCosEps = 1e-8f;


struct Smem {
  float* phat;   // [W*D]  normalised prototypes
  float* sim;    // [W*Nq] C
  float* grad;   // [W*Nq] dL/dC           (backward)
  float* rmax;   // [Nq]   row max over keep
  float* rsum;   // [Nq]   row sum of exp
  float* qinv;   // [Nq]   1/max(|q_j|,eps)  (0 flags a clamped norm in backward)
  float* pinv;   // [W]
  float* part;   // [kWarps]
  int* lab;      // [Nq]
  int* row;      // [Nq]
  int* cnt;      // [W]
  int* start;    // [W]
};

inline size_t smem_words(int Nq, int W, int D) {
  return (size_t)W * D + 2 * (size_t)W * Nq + 5 * (size_t)Nq + 3 * (size_t)W + kWarps + 8;
}

__device__ inline Smem carve(float* b, int Nq, int W, int D) {
  Smem s;
  s.phat = b; b += (size_t)W * D;
  s.sim = b; b += (size_t)W * Nq;
  s.grad = b; b += (size_t)W * Nq;
  s.rmax = b; b += Nq;
  s.rsum = b; b += Nq;
  s.qinv = b; b += Nq;
  s.pinv = b; b += W;
  s.part = b; b += kWarps;
  s.lab = reinterpret_cast<int*>(b); b += Nq;
  s.row = reinterpret_cast<int*>(b); b += Nq;
  s.cnt = reinterpret_cast<int*>(b); b += W;
  s.start = reinterpret_cast<int*>(b);
  return s;
}

__device__ __forceinline__ bool kept(const uint32_t* keep_e, const int* lab, int words, int i, int j) {
  if (keep_e) return (keep_e[(size_t)i * words + (j >> 5)] >> (j & 31)) & 1u;
  return j == i || lab[j] != lab[i];
}

// normalise the prototypes into shared memory: one lane group per row
template <int kLPR, int kCPL>
__device__ inline void stage_prototypes(const Smem& s, const float* protos, int W, int D4, int slot, int nslots, int sub) {
  const float4* p4 = reinterpret_cast<const float4*>(protos);
  float4* ph4 = reinterpret_cast<float4*>(s.phat);
  const int trips = (W + nslots - 1) / nslots;
  for (int it = 0; it < trips; ++it) {
    const int w = it * nslots + slot;
    const int ww = w < W ? w : W - 1;
    float4 v[kCPL];
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < kCPL; ++u) {
      v[u] = __ldg(p4 + (size_t)ww * D4 + sub + u * kLPR);
      ss = fmaf(v[u].x, v[u].x, ss); ss = fmaf(v[u].y, v[u].y, ss);
      ss = fmaf(v[u].z, v[u].z, ss); ss = fmaf(v[u].w, v[u].w, ss);
    }
    ss = group_sum<kLPR>(ss);
    const float nrm = sqrtf(ss);
    const float den = fmaxf(nrm, kCosEps);
    if (w < W) {
#pragma unroll
      for (int u = 0; u < kCPL; ++u) {
        float4 o;
        o.x = __fdiv_rn(v[u].x, den); o.y = __fdiv_rn(v[u].y, den);
        o.z = __fdiv_rn(v[u].z, den); o.w = __fdiv_rn(v[u].w, den);
        ph4[w * D4 + sub + u * kLPR] = o;
      }
      if (sub == 0) s.pinv[w] = nrm > kCosEps ? 1.f / nrm : -1.f / kCosEps;  // negative flags the clamp
    }
  }
}

// similarity matrix C[w,j] for all queries of the episode
template <int kLPR, int kCPL>
__device__ inline void similarities(const Smem& s, const float* queries, int Nq, int W, int D4, float temperature,
                                    int slot, int nslots, int sub) {
  const float4* q4 = reinterpret_cast<const float4*>(queries);
  const float4* ph4 = reinterpret_cast<const float4*>(s.phat);
  const int trips = (Nq + nslots - 1) / nslots;
  for (int it = 0; it < trips; ++it) {
    const int j = it * nslots + slot;
    const int jj = j < Nq ? j : Nq - 1;
    float4 v[kCPL];
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < kCPL; ++u) {
      v[u] = __ldg(q4 + (size_t)jj * D4 + sub + u * kLPR);
      ss = fmaf(v[u].x, v[u].x, ss); ss = fmaf(v[u].y, v[u].y, ss);
      ss = fmaf(v[u].z, v[u].z, ss); ss = fmaf(v[u].w, v[u].w, ss);
    }
    ss = group_sum<kLPR>(ss);
    const float nrm = sqrtf(ss);
    const float den = fmaxf(nrm, kCosEps);
#pragma unroll
    for (int u = 0; u < kCPL; ++u) {
      v[u].x = __fdiv_rn(v[u].x, den); v[u].y = __fdiv_rn(v[u].y, den);
      v[u].z = __fdiv_rn(v[u].z, den); v[u].w = __fdiv_rn(v[u].w, den);
    }
    for (int w = 0; w < W; ++w) {
      float dot = 0.f;
#pragma unroll
      for (int u = 0; u < kCPL; ++u) {
        const float4 p = ph4[w * D4 + sub + u * kLPR];
        dot = fmaf(p.x, v[u].x, dot); dot = fmaf(p.y, v[u].y, dot);
        dot = fmaf(p.z, v[u].z, dot); dot = fmaf(p.w, v[u].w, dot);
      }
      dot = group_sum<kLPR>(dot);
      if (j < Nq && sub == (w & (kLPR - 1))) s.sim[w * Nq + j] = __fdiv_rn(dot, temperature);
    }
    if (j < Nq && sub == 0) s.qinv[j] = nrm > kCosEps ? 1.f / nrm : -1.f / kCosEps;
  }
}

// per query i: max and sum-exp of C[y_i, keep_i]; returns this warp's partial sum of the row losses
__device__ inline float row_stats(const Smem& s, const uint32_t* keep_e, int Nq, int W, int words) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc = 0.f;
  for (int i = warp; i < Nq; i += kWarps) {
    const int y = s.lab[i];
    if (y < 0 || y >= W) {  // label without a prototype: row contributes nothing
      if (lane == 0) { s.rmax[i] = 0.f; s.rsum[i] = 1.f; }
      continue;
    }
    const float* row = s.sim + y * Nq;
    float m = -INFINITY;
    for (int j = lane; j < Nq; j += 32)
      if (kept(keep_e, s.lab, words, i, j)) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float se = 0.f;
    for (int j = lane; j < Nq; j += 32)
      if (kept(keep_e, s.lab, words, i, j)) se += expf(row[j] - m);
    se = warp_sum(se);
    if (lane == 0) {
      s.rmax[i] = m;
      s.rsum[i] = se;
      acc += -((row[i] - m) - logf(se));
    }
  }
  return acc;
}

template <int kLPR, int kCPL>
__global__ void __launch_bounds__(kThreads) cpl_fwd_kernel(const CplParams p) {
  extern __shared__ __align__(16) float smem_raw[];
  constexpr int kGroups = kWarp / kLPR, kSlots = kWarps * kGroups;
  const Smem s = carve(smem_raw, p.Nq, p.W, p.D);
  const int D4 = p.D >> 2, words = (p.Nq + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (kLPR - 1), slot = warp * kGroups + lane / kLPR;
  for (int e = blockIdx.x; e < p.E; e += gridDim.x) {
    for (int k = threadIdx.x; k < p.Nq; k += kThreads) s.lab[k] = p.labels[(size_t)e * p.Nq + k];
    stage_prototypes<kLPR, kCPL>(s, p.protos + (size_t)e * p.W * p.D, p.W, D4, slot, kSlots, sub);
    __syncthreads();
    similarities<kLPR, kCPL>(s, p.queries + (size_t)e * p.Nq * p.D, p.Nq, p.W, D4, p.temperature, slot, kSlots, sub);
    __syncthreads();
    const uint32_t* keep_e = p.keep ? p.keep + (size_t)e * p.Nq * words : nullptr;
    const float part = row_stats(s, keep_e, p.Nq, p.W, words);
    if (lane == 0) s.part[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int k = 0; k < kWarps; ++k) tot += s.part[k];
      // (1/Nq) * NLLLoss(mean): loops/loss.py:131
      p.loss[e] = (float)(1.0 / (double)p.Nq) * (tot / (float)p.Nq);
    }
    __syncthreads();
  }
}

template <int kLPR, int kCPL>
__global__ void __launch_bounds__(kThreads) cpl_bwd_kernel(const CplParams p) {
  extern __shared__ __align__(16) float smem_raw[];
  constexpr int kGroups = kWarp / kLPR, kSlots = kWarps * kGroups;
  const Smem s = carve(smem_raw, p.Nq, p.W, p.D);
  const int D4 = p.D >> 2, words = (p.Nq + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (kLPR - 1), slot = warp * kGroups + lane / kLPR;
  const float4* ph4 = reinterpret_cast<const float4*>(s.phat);
  const float inv_t = 1.f / p.temperature;
  for (int e = blockIdx.x; e < p.E; e += gridDim.x) {
    bucket_by_label(p.labels + (size_t)e * p.Nq, p.Nq, p.W, s.lab, s.row, s.cnt, s.start);
    stage_prototypes<kLPR, kCPL>(s, p.protos + (size_t)e * p.W * p.D, p.W, D4, slot, kSlots, sub);
    __syncthreads();
    const float* q_e = p.queries + (size_t)e * p.Nq * p.D;
    similarities<kLPR, kCPL>(s, q_e, p.Nq, p.W, D4, p.temperature, slot, kSlots, sub);
    __syncthreads();
    const uint32_t* keep_e = p.keep ? p.keep + (size_t)e * p.Nq * words : nullptr;
    row_stats(s, keep_e, p.Nq, p.W, words);
    __syncthreads();
    // dL/dC[w,j] = g * sum_{i: y_i = w} ( keep_ij * softmax_i[j] - [j == i] ),  g = d_loss / Nq^2
    const float g = p.d_loss[e] * (float)(1.0 / (double)p.Nq) / (float)p.Nq;
    for (int item = threadIdx.x; item < p.W * p.Nq; item += kThreads) {
      const int w = item / p.Nq, j = item - w * p.Nq;
      const float c = s.sim[item];
      float acc = 0.f;
      const int* members = s.row + s.start[w];
      for (int k = 0; k < s.cnt[w]; ++k) {
        const int i = members[k];
        float t = kept(keep_e, s.lab, words, i, j) ? expf(c - s.rmax[i]) / s.rsum[i] : 0.f;
        if (j == i) t -= 1.f;
        acc += t;
      }
      s.grad[item] = g * acc;
    }
    __syncthreads();
    // queries: dq^ = 1/T sum_w G[w,j] p^_w ; dq = (dq^ - q^ <q^,dq^>) / |q|
    {
      const float4* q4 = reinterpret_cast<const float4*>(q_e);
      float4* dq4 = reinterpret_cast<float4*>(p.d_queries + (size_t)e * p.Nq * p.D);
      const int trips = (p.Nq + kSlots - 1) / kSlots;
      for (int it = 0; it < trips; ++it) {
        const int j = it * kSlots + slot;
        const int jj = j < p.Nq ? j : p.Nq - 1;
        const float qi = s.qinv[jj];
        const bool clamped = qi < 0.f;
        const float inv = fabsf(qi);
        float4 qh[kCPL], acc[kCPL];
#pragma unroll
        for (int u = 0; u < kCPL; ++u) {
          const float4 v = __ldg(q4 + (size_t)jj * D4 + sub + u * kLPR);
          qh[u] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
          acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int w = 0; w < p.W; ++w) {
          const float gw = s.grad[w * p.Nq + jj] * inv_t;
#pragma unroll
          for (int u = 0; u < kCPL; ++u) {
            const float4 pr = ph4[w * D4 + sub + u * kLPR];
            acc[u].x = fmaf(gw, pr.x, acc[u].x); acc[u].y = fmaf(gw, pr.y, acc[u].y);
            acc[u].z = fmaf(gw, pr.z, acc[u].z); acc[u].w = fmaf(gw, pr.w, acc[u].w);
          }
        }
        float dot = 0.f;
#pragma unroll
        for (int u = 0; u < kCPL; ++u) {
          dot = fmaf(acc[u].x, qh[u].x, dot); dot = fmaf(acc[u].y, qh[u].y, dot);
          dot = fmaf(acc[u].z, qh[u].z, dot); dot = fmaf(acc[u].w, qh[u].w, dot);
        }
        dot = group_sum<kLPR>(dot);
        if (clamped) dot = 0.f;  // norm clamped to eps: the denominator is a constant
        if (j < p.Nq) {
#pragma unroll
          for (int u = 0; u < kCPL; ++u) {
            float4 o;
            o.x = (acc[u].x - qh[u].x * dot) * inv; o.y = (acc[u].y - qh[u].y * dot) * inv;
            o.z = (acc[u].z - qh[u].z * dot) * inv; o.w = (acc[u].w - qh[u].w * dot) * inv;
            stg_stream(dq4 + (size_t)j * D4 + sub + u * kLPR, o);
          }
        }
      }
    }
    // prototypes: dp^ = 1/T sum_j G[w,j] q^_j (ascending j) ; dp = (dp^ - p^ <p^,dp^>) / |p|
    {
      const float4* q4 = reinterpret_cast<const float4*>(q_e);
      float4* dp4 = reinterpret_cast<float4*>(p.d_protos + (size_t)e * p.W * p.D);
      const int trips = (p.W + kSlots - 1) / kSlots;
      for (int it = 0; it < trips; ++it) {
        const int w = it * kSlots + slot;
        const int ww = w < p.W ? w : p.W - 1;
        float4 acc[kCPL];
#pragma unroll
        for (int u = 0; u < kCPL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < p.Nq; ++j) {
          const float gj = s.grad[ww * p.Nq + j] * inv_t * fabsf(s.qinv[j]);
#pragma unroll
          for (int u = 0; u < kCPL; ++u) {
            const float4 v = __ldg(q4 + (size_t)j * D4 + sub + u * kLPR);
            acc[u].x = fmaf(gj, v.x, acc[u].x); acc[u].y = fmaf(gj, v.y, acc[u].y);
            acc[u].z = fmaf(gj, v.z, acc[u].z); acc[u].w = fmaf(gj, v.w, acc[u].w);
          }
        }
        const float pi = s.pinv[ww];
        const bool clamped = pi < 0.f;
        const float inv = fabsf(pi);
        float dot = 0.f;
        float4 ph[kCPL];
#pragma unroll
        for (int u = 0; u < kCPL; ++u) {
          ph[u] = ph4[ww * D4 + sub + u * kLPR];
          dot = fmaf(acc[u].x, ph[u].x, dot); dot = fmaf(acc[u].y, ph[u].y, dot);
          dot = fmaf(acc[u].z, ph[u].z, dot); dot = fmaf(acc[u].w, ph[u].w, dot);
        }
        dot = group_sum<kLPR>(dot);
        if (clamped) dot = 0.f;
        if (w < p.W) {
#pragma unroll
          for (int u = 0; u < kCPL; ++u) {
            float4 o;
            o.x = (acc[u].x - ph[u].x * dot) * inv; o.y = (acc[u].y - ph[u].y * dot) * inv;
            o.z = (acc[u].z - ph[u].z * dot) * inv; o.w = (acc[u].w - ph[u].w * dot) * inv;
            dp4[(size_t)w * D4 + sub + u * kLPR] = o;
          }
        }
      }
    }
    __syncthreads();
  }
}

using KernelFn = void (*)(const CplParams);

bool pick(int D, KernelFn& fwd, KernelFn& bwd) {
#define AFSL_VARIANT(LPR, CPL)            \
  if (D == 4 * LPR * CPL) {               \
    fwd = cpl_fwd_kernel<LPR, CPL>;       \
    bwd = cpl_bwd_kernel<LPR, CPL>;       \
    return true;                          \
  }
  AFSL_VARIANT(4, 1) AFSL_VARIANT(8, 1) AFSL_VARIANT(16, 1) AFSL_VARIANT(32, 1)
  AFSL_VARIANT(32, 2) AFSL_VARIANT(32, 4) AFSL_VARIANT(32, 8)
#undef AFSL_VARIANT
  return false;
}

int launch(const CplParams& p, bool bwd, cudaStream_t stream, const char* name) {
  AFSL_REQUIRE(p.protos && p.queries && p.labels, "%s: null pointer", name);
  AFSL_REQUIRE(p.E >= 0 && p.Nq > 0 && p.W > 0, "%s: bad sizes E=%d Nq=%d W=%d", name, p.E, p.Nq, p.W);
  AFSL_REQUIRE(p.temperature != 0.f, "%s: temperature must be non-zero", name);
  if (p.E == 0) return AFSL_OK;
  const char* warp_env = getenv("AFSL_CPL_WARP");      // read per launch so the tests can exercise both paths
  if (!warp_env || atoi(warp_env) != 0) {
    bool handled = false;
    const int rc = launch_cpl_warp(p, bwd, stream, name, &handled);
    if (rc != AFSL_OK || handled) return rc;
  }
  KernelFn f, b;
  AFSL_REQUIRE(pick(p.D, f, b), "%s: unsupported embedding dim D=%d (supported: 16,32,64,128,256,512,1024)", name, p.D);
  KernelFn fn = bwd ? b : f;
  const size_t bytes = smem_words(p.Nq, p.W, p.D) * sizeof(float);
  if (int rc = opt_in_smem(fn, bytes, name)) return rc;
  const int grid = persistent_grid(fn, kThreads, bytes, p.E);
  fn<<<grid, kThreads, bytes, stream>>>(p);
  AFSL_CHECK_LAUNCH(name);
  return AFSL_OK;
}

}  // namespace
}  // namespace afsl

extern "C" int afsl_cpl_fwd_f32(const float* protos, const float* queries, const int32_t* labels, const uint32_t* keep,
                                 float temperature, float* loss, int E, int Nq, int W, int D, void* stream) {
  AFSL_REQUIRE(loss, "afsl_cpl_fwd_f32: null loss");
  afsl::CplParams p{};
  p.protos = protos; p.queries = queries; p.labels = labels; p.keep = keep; p.temperature = temperature;
  p.loss = loss; p.E = E; p.Nq = Nq; p.W = W; p.D = D;
  return afsl::launch(p, false, (cudaStream_t)stream, "afsl_cpl_fwd_f32");
}

extern "C" int afsl_cpl_bwd_f32(const float* protos, const float* queries, const int32_t* labels, const uint32_t* keep,
                                 float temperature, const float* d_loss, float* d_protos, float* d_queries, int E, int Nq,
                                 int W, int D, void* stream) {
  AFSL_REQUIRE(d_loss && d_protos && d_queries, "afsl_cpl_bwd_f32: null pointer");
  afsl::CplParams p{};
  p.protos = protos; p.queries = queries; p.labels = labels; p.keep = keep; p.temperature = temperature;
  p.d_loss = d_loss; p.d_protos = d_protos; p.d_queries = d_queries; p.E = E; p.Nq = Nq; p.W = W; p.D = D;
  return afsl::launch(p, true, (cudaStream_t)stream, "afsl_cpl_bwd_f32");
}

// ---- the same pair with the forward's similarity matrix handed to the backward (warp family only: 5-way, D in {64, 128, 256})
extern "C" int afsl_cpl_saved_supported(int Nq, int W, int D) {
  const char* warp_env = getenv("AFSL_CPL_WARP");
  if (warp_env && atoi(warp_env) == 0) return 0;
  return afsl::cpl_warp_supported(Nq, W, D) ? 1 : 0;
}

extern "C" int afsl_cpl_fwd_save_f32(const float* protos, const float* queries, const int32_t* labels, const uint32_t* keep,
                                      float temperature, float* loss, float* sim, float* qinv, int E, int Nq, int W, int D,
                                      void* stream) {
  AFSL_REQUIRE(loss && sim && qinv, "afsl_cpl_fwd_save_f32: null pointer");
  AFSL_REQUIRE(afsl_cpl_saved_supported(Nq, W, D), "afsl_cpl_fwd_save_f32: shape Nq=%d W=%d D=%d is not taken by the warp kernels "
               "(ask afsl_cpl_saved_supported first)", Nq, W, D);
  afsl::CplParams p{};
  p.protos = protos; p.queries = queries; p.labels = labels; p.keep = keep; p.temperature = temperature;
  p.loss = loss; p.sim_out = sim; p.qinv_out = qinv; p.E = E; p.Nq = Nq; p.W = W; p.D = D;
  return afsl::launch(p, false, (cudaStream_t)stream, "afsl_cpl_fwd_save_f32");
}

extern "C" int afsl_cpl_bwd_saved_f32(const float* protos, const float* queries, const int32_t* labels, const uint32_t* keep,
                                       float temperature, const float* sim, const float* qinv, const float* d_loss,
                                       float* d_protos, float* d_queries, int E, int Nq, int W, int D, void* stream) {
  AFSL_REQUIRE(sim && qinv && d_loss && d_protos && d_queries, "afsl_cpl_bwd_saved_f32: null pointer");
  AFSL_REQUIRE(afsl_cpl_saved_supported(Nq, W, D), "afsl_cpl_bwd_saved_f32: shape Nq=%d W=%d D=%d is not taken by the warp kernels",
               Nq, W, D);
  afsl::CplParams p{};
  p.protos = protos; p.queries = queries; p.labels = labels; p.keep = keep; p.temperature = temperature;
  p.sim_in = sim; p.qinv_in = qinv; p.d_loss = d_loss; p.d_protos = d_protos; p.d_queries = d_queries;
  p.E = E; p.Nq = Nq; p.W = W; p.D = D;
  return afsl::launch(p, true, (cudaStream_t)stream, "afsl_cpl_bwd_saved_f32");
}
