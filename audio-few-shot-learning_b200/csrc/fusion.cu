// Multi-view self-attention fusion: one post-norm transformer encoder layer (1 head, ReLU FFN) over
// the V views of every sample, forward and backward.
//
// Reference: SelfAttention.forward, models/main_modules.py:222-228 =
// nn.TransformerEncoderLayer(d_model=64, nhead=1, dim_feedforward=256, dropout=0.1, batch_first=True):
//   qkv = x Win^T + bin ; a = softmax(q k^T / sqrt(d)) ; ctx = drop(a) v ; sa = ctx Wo^T + bo
//   x1 = LN1(x + drop1(sa)) ; h = relu(x1 W1^T + b1) ; ff = drop(h) W2^T + b2 ; y = LN2(x1 + drop2(ff))
// followed by laying the V tokens side by side ([N, V*d] is the same memory as [N, V, d]).
// Dropout masks (already scaled by 1/(1-p)) are inputs, so the kernel is deterministic; NULL = eval mode.
//
// The layer is a chain of small GEMMs that share 49 984 weights across all tokens (M = N*V rows): fp32 FFMA
// work at ~200 flop/B, i.e. compute-bound on the fp32 pipe (1e-5 parity rules out TF32 tensor cores).
// One CTA processes tiles of kTM = 32 tokens entirely in shared memory; weights stream through L1/L2
// (200 KB, shared by every CTA).  The backward recomputes the forward for its tile (cheaper than saving
// ~2.5 KB of activations per token to HBM) and accumulates weight gradients into a per-CTA partial buffer
// that the host sums - no atomics, bit-reproducible.
#include "afsl_common.cuh"

namespace afsl {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / kWarp;
constexpr int kTM = 32;          // tokens per tile
constexpr int kD = 64;           // d_model
constexpr int kF = 256;          // dim_feedforward
constexpr int kQ = 3 * kD;       // q | k | v
constexpr float kLnEps = 1e-5f;

// packed parameter block (floats); *_t = transposed ([in][out]) copy used by the forward GEMMs
struct WeightOffsets {
  static constexpr int win = 0;                       // [kQ][kD]   in_proj_weight
  static constexpr int bin = win + kQ * kD;           // [kQ]
  static constexpr int wo = bin + kQ;                 // [kD][kD]   out_proj.weight
  static constexpr int bo = wo + kD * kD;             // [kD]
  static constexpr int w1 = bo + kD;                  // [kF][kD]   linear1.weight
  static constexpr int b1 = w1 + kF * kD;             // [kF]
  static constexpr int w2 = b1 + kF;                  // [kD][kF]   linear2.weight
  static constexpr int b2 = w2 + kD * kF;             // [kD]
  static constexpr int g1 = b2 + kD;                  // norm1.weight
  static constexpr int be1 = g1 + kD;
  static constexpr int g2 = be1 + kD;                 // norm2.weight
  static constexpr int be2 = g2 + kD;
  static constexpr int count = be2 + kD;              // 49 984 parameters
  // transposed copies appended after the parameters
  static constexpr int win_t = count;                 // [kD][kQ]
  static constexpr int wo_t = win_t + kQ * kD;        // [kD][kD]
  static constexpr int w1_t = wo_t + kD * kD;         // [kD][kF]
  static constexpr int w2_t = w1_t + kF * kD;         // [kF][kD]
  static constexpr int total = w2_t + kD * kF;
};

struct FusionParams {
  const float* x;        // [N, V, kD]
  const float* w;        // packed block, WeightOffsets::total floats
  const float* drop_attn;  // [N, V, V] or null
  const float* drop1;      // [N, V, kD] or null
  const float* dropf;      // [N, V, kF] or null
  const float* drop2;      // [N, V, kD] or null
  float* y;              // [N, V, kD]                       (forward)
  const float* dy;       // [N, V, kD]                       (backward)
  float* dx;             // [N, V, kD]
  float* dw_part;        // [gridDim.x, WeightOffsets::count] per-CTA partial weight gradients (zero-initialised)
  int N, V;
};

// shared-memory tile buffers (row stride = width + 1 for conflict-free column walks in the dW loops)
struct Tile {
  float *x, *qkv, *p, *ctx, *r1h, *rs1, *x1, *h, *r2h, *rs2;   // forward state
  float *g, *g1, *dqkv, *dctx;                               // backward scratch
};
constexpr int ldD = kD + 4, ldQ = kQ + 4, ldF = kF + 4;          // +4 keeps rows 16-byte aligned

inline size_t tile_floats(bool bwd) {
  size_t n = (size_t)kTM * (ldD * 5 + ldQ + ldF + 8) + 2 * kTM;  // x, ctx, r1h, x1, r2h | qkv | h | p
  if (bwd) n += (size_t)kTM * (ldD * 3 + ldQ);                    // g, g1, dctx | dqkv
  return n;
}

__device__ inline Tile carve(float* b, bool bwd) {
  Tile t;
  t.x = b; b += kTM * ldD;
  t.ctx = b; b += kTM * ldD;
  t.r1h = b; b += kTM * ldD;
  t.x1 = b; b += kTM * ldD;
  t.r2h = b; b += kTM * ldD;
  t.qkv = b; b += kTM * ldQ;
  t.h = b; b += kTM * ldF;
  t.p = b; b += kTM * 8;
  t.rs1 = b; b += kTM;
  t.rs2 = b; b += kTM;
  t.g = t.g1 = t.dctx = t.dqkv = nullptr;
  if (bwd) {
    t.g = b; b += kTM * ldD;
    t.g1 = b; b += kTM * ldD;
    t.dctx = b; b += kTM * ldD;
    t.dqkv = b; b += kTM * ldQ;
  }
  return t;
}

// C[m][n] = (bias[n] | C[m][n]) + sum_k A[m][k] * B[k*ldb + n],  m < rows, n < N (N % 4 == 0)
// A, C in shared memory; B in global memory with contiguous n (coalesced 128-bit reads through L1).
template <int K, bool kAccum, int kR = 4>
__device__ inline void gemm_tile(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                 const float* __restrict__ bias, float* __restrict__ C, int ldc, int rows, int N) {
  // kR rows x 4 columns per thread: 4 rows for the wide outputs (N = 192 / 256), 2 rows for N = 64, where 4 would leave
  // half of the CTA's threads without rows of the 32-token tile
  const int ncg = N >> 2, nrg = kThreads / ncg;
  const int cg = threadIdx.x % ncg, rg = threadIdx.x / ncg;
  if (rg >= nrg) return;
  for (int r0 = rg * kR; r0 < rows; r0 += nrg * kR) {
    float4 acc[kR];
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      if (kAccum) acc[i] = *reinterpret_cast<const float4*>(C + (size_t)min(r0 + i, rows - 1) * ldc + cg * 4);
      else acc[i] = bias ? __ldg(reinterpret_cast<const float4*>(bias) + cg) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // four k per step: one 128-bit shared-memory read per row (a broadcast: the lanes of a warp share the row group) and
    // four 128-bit weight reads feed 16 kR FMAs - the first version read A element by element (5 loads per 16 FMAs: bound
    // by the load / store unit, not the FMA pipe).  Same fmaf sequence per output, so the results are bit-identical.
    const float* ar[kR];
#pragma unroll
    for (int i = 0; i < kR; ++i) ar[i] = A + (size_t)(r0 + min(i, rows - 1 - r0)) * lda;
    static_assert(K % 4 == 0, "K must be a multiple of 4");
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
      float4 a[kR], b[4];
#pragma unroll
      for (int i = 0; i < kR; ++i) a[i] = *reinterpret_cast<const float4*>(ar[i] + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) b[kk] = __ldg(reinterpret_cast<const float4*>(B + (size_t)(k + kk) * ldb) + cg);
#pragma unroll
      for (int i = 0; i < kR; ++i) {
        const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          acc[i].x = fmaf(av[kk], b[kk].x, acc[i].x); acc[i].y = fmaf(av[kk], b[kk].y, acc[i].y);
          acc[i].z = fmaf(av[kk], b[kk].z, acc[i].z); acc[i].w = fmaf(av[kk], b[kk].w, acc[i].w);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kR; ++i)
      if (r0 + i < rows) *reinterpret_cast<float4*>(C + (size_t)(r0 + i) * ldc + cg * 4) = acc[i];
  }
}

// per-CTA partial weight gradient: part[n*K + k] += sum_m G[m][n] * A[m][k] ; bias_part[n] += sum_m G[m][n]
template <int K>
__device__ inline void grad_weights(const float* __restrict__ G, int ldg, const float* __restrict__ A, int lda, int rows, int N,
                                    float* __restrict__ part, float* __restrict__ bias_part) {
  // register tile of 4 outputs n x 4 inputs k per thread: two 128-bit shared-memory reads feed 16 FMAs per row (the first
  // version had one output per thread: a scalar and a 128-bit read per 4 FMAs).  Rows are added in ascending order per
  // element as before: bit-identical.
  constexpr int K4 = K / 4;
  const int N4 = N >> 2;
  for (int item = threadIdx.x; item < N4 * K4; item += kThreads) {
    const int ng = item / K4, kg = item - ng * K4;
    float4 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int m = 0; m < rows; ++m) {
      const float4 g = *reinterpret_cast<const float4*>(G + (size_t)m * ldg + ng * 4);
      const float4 a = *reinterpret_cast<const float4*>(A + (size_t)m * lda + kg * 4);
      const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j].x = fmaf(gv[j], a.x, acc[j].x); acc[j].y = fmaf(gv[j], a.y, acc[j].y);
        acc[j].z = fmaf(gv[j], a.z, acc[j].z); acc[j].w = fmaf(gv[j], a.w, acc[j].w);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4* dst = reinterpret_cast<float4*>(part + (size_t)(ng * 4 + j) * K) + kg;
      float4 cur = *dst;
      cur.x += acc[j].x; cur.y += acc[j].y; cur.z += acc[j].z; cur.w += acc[j].w;
      *dst = cur;
    }
  }
  for (int n = threadIdx.x; n < N; n += kThreads) {
    float acc = 0.f;
    for (int m = 0; m < rows; ++m) acc += G[(size_t)m * ldg + n];
    bias_part[n] += acc;
  }
}

// out = LayerNorm(res + drop * add): keeps rhat (normalised, pre-affine) and rstd for the backward
__device__ inline void residual_layernorm(const float* res, const float* add, const float* __restrict__ drop, size_t drop_off,
                                          const float* __restrict__ gamma, const float* __restrict__ beta, float* out,
                                          float* rhat, float* rstd, int rows) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int m = warp; m < rows; m += kWarps) {
    float v[kD / 32];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < kD / 32; ++u) {
      const int j = lane + 32 * u;
      float a = add[(size_t)m * ldD + j];
      if (drop) a *= __ldg(drop + drop_off + (size_t)m * kD + j);
      v[u] = res[(size_t)m * ldD + j] + a;
      s += v[u];
    }
    const float mean = warp_sum(s) * (1.f / kD);
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < kD / 32; ++u) { const float d = v[u] - mean; q = fmaf(d, d, q); }
    const float rs = rsqrtf(warp_sum(q) * (1.f / kD) + kLnEps);
#pragma unroll
    for (int u = 0; u < kD / 32; ++u) {
      const int j = lane + 32 * u;
      const float n = (v[u] - mean) * rs;
      rhat[(size_t)m * ldD + j] = n;
      out[(size_t)m * ldD + j] = fmaf(n, __ldg(gamma + j), __ldg(beta + j));
    }
    if (lane == 0) rstd[m] = rs;
  }
}

// dr = rstd * (g - mean(g) - rhat * mean(g*rhat)), g = dy*gamma ; also per-CTA dgamma / dbeta partials
__device__ inline void layernorm_backward(float* dy_to_dr, const float* rhat, const float* rstd, const float* __restrict__ gamma,
                                          int rows, float* __restrict__ dgamma_part, float* __restrict__ dbeta_part) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  for (int j = threadIdx.x; j < kD; j += kThreads) {      // column sums first (dy still intact)
    float a = 0.f, b = 0.f;
    for (int m = 0; m < rows; ++m) {
      const float d = dy_to_dr[(size_t)m * ldD + j];
      a = fmaf(d, rhat[(size_t)m * ldD + j], a);
      b += d;
    }
    dgamma_part[j] += a;
    dbeta_part[j] += b;
  }
  __syncthreads();
  for (int m = warp; m < rows; m += kWarps) {
    float g[kD / 32], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int u = 0; u < kD / 32; ++u) {
      const int j = lane + 32 * u;
      g[u] = dy_to_dr[(size_t)m * ldD + j] * __ldg(gamma + j);
      s1 += g[u];
      s2 = fmaf(g[u], rhat[(size_t)m * ldD + j], s2);
    }
    s1 = warp_sum(s1) * (1.f / kD);
    s2 = warp_sum(s2) * (1.f / kD);
    const float rs = rstd[m];
#pragma unroll
    for (int u = 0; u < kD / 32; ++u) {
      const int j = lane + 32 * u;
      dy_to_dr[(size_t)m * ldD + j] = rs * (g[u] - s1 - rhat[(size_t)m * ldD + j] * s2);
    }
  }
  __syncthreads();
}

// forward of one tile; every intermediate the backward needs stays in the Tile buffers
__device__ inline void tile_forward(const FusionParams& p, const Tile& t, int tok0, int rows) {
  using O = WeightOffsets;
  const float* w = p.w;
  const int V = p.V;
  // x tile
  for (int i = threadIdx.x; i < rows * (kD / 4); i += kThreads) {
    const int m = i / (kD / 4), c = i - m * (kD / 4);
    *reinterpret_cast<float4*>(t.x + (size_t)m * ldD + c * 4) =
        __ldg(reinterpret_cast<const float4*>(p.x + (size_t)(tok0 + m) * kD) + c);
  }
  __syncthreads();
  gemm_tile<kD, false>(t.x, ldD, w + O::win_t, kQ, w + O::bin, t.qkv, ldQ, rows, kQ);
  __syncthreads();
  // attention probabilities: token m attends to the V tokens of its sample
  for (int i = threadIdx.x; i < rows * V; i += kThreads) {
    const int m = i / V, u = i - m * V;
    const int base = m - (m % V);                 // tiles start on a sample boundary (kTM % V == 0)
    const float* q = t.qkv + (size_t)m * ldQ;
    const float* k = t.qkv + (size_t)(base + u) * ldQ + kD;
    float dot = 0.f;
#pragma unroll 8
    for (int j = 0; j < kD; ++j) dot = fmaf(q[j], k[j], dot);
    t.p[m * 8 + u] = dot * 0.125f;                // 1/sqrt(64)
  }
  __syncthreads();
  for (int m = threadIdx.x; m < rows; m += kThreads) {
    float mx = -INFINITY, se = 0.f;
    for (int u = 0; u < V; ++u) mx = fmaxf(mx, t.p[m * 8 + u]);
    for (int u = 0; u < V; ++u) { const float e = expf(t.p[m * 8 + u] - mx); t.p[m * 8 + u] = e; se += e; }
    for (int u = 0; u < V; ++u) t.p[m * 8 + u] /= se;      // softmax (kept un-dropped for the backward)
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows * kD; i += kThreads) {
    const int m = i / kD, j = i - m * kD;
    const int base = m - (m % V);
    float acc = 0.f;
    for (int u = 0; u < V; ++u) {
      float a = t.p[m * 8 + u];
      if (p.drop_attn) a *= __ldg(p.drop_attn + ((size_t)(tok0 + m) * V + u));
      acc = fmaf(a, t.qkv[(size_t)(base + u) * ldQ + 2 * kD + j], acc);
    }
    t.ctx[(size_t)m * ldD + j] = acc;
  }
  __syncthreads();
  gemm_tile<kD, false, 2>(t.ctx, ldD, w + O::wo_t, kD, w + O::bo, t.x1, ldD, rows, kD);     // sa -> x1 buffer
  __syncthreads();
  residual_layernorm(t.x, t.x1, p.drop1, (size_t)tok0 * kD, w + O::g1, w + O::be1, t.x1, t.r1h, t.rs1, rows);
  __syncthreads();
  gemm_tile<kD, false>(t.x1, ldD, w + O::w1_t, kF, w + O::b1, t.h, ldF, rows, kF);
  __syncthreads();
  for (int i = threadIdx.x; i < rows * kF; i += kThreads) {   // relu, then dropout (kept: h' = relu(h) * drop)
    const int m = i / kF, j = i - m * kF;
    float v = fmaxf(t.h[(size_t)m * ldF + j], 0.f);
    if (p.dropf) v *= __ldg(p.dropf + (size_t)(tok0 + m) * kF + j);
    t.h[(size_t)m * ldF + j] = v;
  }
  __syncthreads();
  gemm_tile<kF, false, 2>(t.h, ldF, w + O::w2_t, kD, w + O::b2, t.r2h, ldD, rows, kD);     // ff -> r2h buffer
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads) fusion_fwd_kernel(const FusionParams p) {
  extern __shared__ __align__(16) float smem[];
  const Tile t = carve(smem, false);
  using O = WeightOffsets;
  const int tokens = p.N * p.V;
  for (int tok0 = blockIdx.x * kTM; tok0 < tokens; tok0 += gridDim.x * kTM) {
    const int rows = min(kTM, tokens - tok0);
    tile_forward(p, t, tok0, rows);
    // y = LN2(x1 + drop2(ff)); r2h currently holds ff
    residual_layernorm(t.x1, t.r2h, p.drop2, (size_t)tok0 * kD, p.w + O::g2, p.w + O::be2, t.x, t.r2h, t.rs2, rows);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * (kD / 4); i += kThreads) {
      const int m = i / (kD / 4), c = i - m * (kD / 4);
      reinterpret_cast<float4*>(p.y + (size_t)(tok0 + m) * kD)[c] = *reinterpret_cast<const float4*>(t.x + (size_t)m * ldD + c * 4);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads) fusion_bwd_kernel(const FusionParams p) {
  extern __shared__ __align__(16) float smem[];
  const Tile t = carve(smem, true);
  using O = WeightOffsets;
  const float* w = p.w;
  float* part = p.dw_part + (size_t)blockIdx.x * O::count;
  const int V = p.V, tokens = p.N * V;
  for (int tok0 = blockIdx.x * kTM; tok0 < tokens; tok0 += gridDim.x * kTM) {
    const int rows = min(kTM, tokens - tok0);
    tile_forward(p, t, tok0, rows);
    // ---- LN2: keep x1 intact, y is not needed; rhat2/rstd2 into r2h/rs2, scratch output into g1
    residual_layernorm(t.x1, t.r2h, p.drop2, (size_t)tok0 * kD, w + O::g2, w + O::be2, t.g1, t.r2h, t.rs2, rows);
    for (int i = threadIdx.x; i < rows * (kD / 4); i += kThreads) {           // g = dy tile
      const int m = i / (kD / 4), c = i - m * (kD / 4);
      *reinterpret_cast<float4*>(t.g + (size_t)m * ldD + c * 4) = __ldg(reinterpret_cast<const float4*>(p.dy + (size_t)(tok0 + m) * kD) + c);
    }
    layernorm_backward(t.g, t.r2h, t.rs2, w + O::g2, rows, part + O::g2, part + O::be2);   // g = d r2
    // d x1 (residual branch) = d r2 ; d ff = d r2 * drop2  -> g1
    for (int i = threadIdx.x; i < rows * kD; i += kThreads) {
      const int m = i / kD, j = i - m * kD;
      float v = t.g[(size_t)m * ldD + j];
      if (p.drop2) v *= __ldg(p.drop2 + (size_t)(tok0 + m) * kD + j);
      t.g1[(size_t)m * ldD + j] = v;
    }
    __syncthreads();
    grad_weights<kF>(t.g1, ldD, t.h, ldF, rows, kD, part + O::w2, part + O::b2);          // dW2 [kD][kF], db2
    __syncthreads();
    // d h = (d ff W2) * drop * [h' > 0], in place over h' (h' = relu(h)*drop is positive iff the unit is active and kept)
    {
      const int ncg = kF >> 2, nrg = kThreads / ncg;
      const int cg = threadIdx.x % ncg, rg = threadIdx.x / ncg;
      for (int r0 = rg; r0 < rows; r0 += nrg) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int k = 0; k < kD; ++k) {
          const float a = t.g1[(size_t)r0 * ldD + k];
          const float4 b = __ldg(reinterpret_cast<const float4*>(w + O::w2 + (size_t)k * kF) + cg);
          acc.x = fmaf(a, b.x, acc.x); acc.y = fmaf(a, b.y, acc.y); acc.z = fmaf(a, b.z, acc.z); acc.w = fmaf(a, b.w, acc.w);
        }
        float4* hp = reinterpret_cast<float4*>(t.h + (size_t)r0 * ldF + cg * 4);
        const float4 hv = *hp;
        float4 d = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p.dropf) d = __ldg(reinterpret_cast<const float4*>(p.dropf + (size_t)(tok0 + r0) * kF) + cg);
        acc.x = hv.x > 0.f ? acc.x * d.x : 0.f;
        acc.y = hv.y > 0.f ? acc.y * d.y : 0.f;
        acc.z = hv.z > 0.f ? acc.z * d.z : 0.f;
        acc.w = hv.w > 0.f ? acc.w * d.w : 0.f;
        *hp = acc;
      }
    }
    __syncthreads();
    grad_weights<kD>(t.h, ldF, t.x1, ldD, rows, kF, part + O::w1, part + O::b1);           // dW1 [kF][kD], db1
    __syncthreads();
    gemm_tile<kF, true, 2>(t.h, ldF, w + O::w1, kD, nullptr, t.g, ldD, rows, kD);             // g = d x1 = d r2 + d h W1
    layernorm_backward(t.g, t.r1h, t.rs1, w + O::g1, rows, part + O::g1, part + O::be1);   // g = d r1
    for (int i = threadIdx.x; i < rows * kD; i += kThreads) {                              // g1 = d sa = d r1 * drop1
      const int m = i / kD, j = i - m * kD;
      float v = t.g[(size_t)m * ldD + j];
      if (p.drop1) v *= __ldg(p.drop1 + (size_t)(tok0 + m) * kD + j);
      t.g1[(size_t)m * ldD + j] = v;
    }
    __syncthreads();
    grad_weights<kD>(t.g1, ldD, t.ctx, ldD, rows, kD, part + O::wo, part + O::bo);         // dWo, dbo
    gemm_tile<kD, false, 2>(t.g1, ldD, w + O::wo, kD, nullptr, t.dctx, ldD, rows, kD);        // d ctx = d sa Wo
    __syncthreads();
    // ---- attention backward, per token m (query role) and per key token
    // d a'[m][u] = d ctx[m] . v[u]; d a = d a' * drop; d s = a * (d a - sum a d a) / 8
    for (int i = threadIdx.x; i < rows * V; i += kThreads) {
      const int m = i / V, u = i - m * V, base = m - (m % V);
      const float* dc = t.dctx + (size_t)m * ldD;
      const float* v = t.qkv + (size_t)(base + u) * ldQ + 2 * kD;
      float dot = 0.f;
#pragma unroll 8
      for (int j = 0; j < kD; ++j) dot = fmaf(dc[j], v[j], dot);
      if (p.drop_attn) dot *= __ldg(p.drop_attn + ((size_t)(tok0 + m) * V + u));
      t.dqkv[(size_t)m * ldQ + u] = dot;            // scratch: first V columns of the dq slot hold d a
    }
    __syncthreads();
    for (int m = threadIdx.x; m < rows; m += kThreads) {
      float s = 0.f;
      for (int u = 0; u < V; ++u) s = fmaf(t.p[m * 8 + u], t.dqkv[(size_t)m * ldQ + u], s);
      float ds[8];
      for (int u = 0; u < V; ++u) ds[u] = t.p[m * 8 + u] * (t.dqkv[(size_t)m * ldQ + u] - s) * 0.125f;
      for (int u = 0; u < V; ++u) t.dqkv[(size_t)m * ldQ + kD + kD + kD - 8 + u] = ds[u];   // park d s at the tail of the v slot
    }
    __syncthreads();
    // d v[u] = sum_t a'[t][u] d ctx[t]  (t over the sample's tokens) -- must be computed before the tail is overwritten:
    // the parked d s occupies columns [3kD-8, 3kD) of dqkv, so d v is written to g (free now) and copied afterwards
    for (int i = threadIdx.x; i < rows * kD; i += kThreads) {
      const int m = i / kD, j = i - m * kD, base = m - (m % V), u = m - base;
      float acc = 0.f;
      for (int tt = 0; tt < V; ++tt) {
        float a = t.p[(base + tt) * 8 + u];
        if (p.drop_attn) a *= __ldg(p.drop_attn + ((size_t)(tok0 + base + tt) * V + u));
        acc = fmaf(a, t.dctx[(size_t)(base + tt) * ldD + j], acc);
      }
      t.g1[(size_t)m * ldD + j] = acc;              // d v (g1 is free: d sa already consumed)
    }
    __syncthreads();
    // d q[m] = sum_u d s[m][u] k[u] ; d k[m] = sum_t d s[t][m_local] q[t]
    for (int i = threadIdx.x; i < rows * kD; i += kThreads) {
      const int m = i / kD, j = i - m * kD, base = m - (m % V), ul = m - base;
      float dq = 0.f, dk = 0.f;
      for (int u = 0; u < V; ++u) {
        dq = fmaf(t.dqkv[(size_t)m * ldQ + 3 * kD - 8 + u], t.qkv[(size_t)(base + u) * ldQ + kD + j], dq);
        dk = fmaf(t.dqkv[(size_t)(base + u) * ldQ + 3 * kD - 8 + ul], t.qkv[(size_t)(base + u) * ldQ + j], dk);
      }
      t.dctx[(size_t)m * ldD + j] = dq;             // stage d q in dctx (consumed), d k in x1? no: x1 still needed -> r2h
      t.r2h[(size_t)m * ldD + j] = dk;              // r2h is free after the LN2 backward
    }
    __syncthreads();
    for (int i = threadIdx.x; i < rows * kD; i += kThreads) {     // assemble d qkv = [d q | d k | d v]
      const int m = i / kD, j = i - m * kD;
      t.dqkv[(size_t)m * ldQ + j] = t.dctx[(size_t)m * ldD + j];
      t.dqkv[(size_t)m * ldQ + kD + j] = t.r2h[(size_t)m * ldD + j];
      t.dqkv[(size_t)m * ldQ + 2 * kD + j] = t.g1[(size_t)m * ldD + j];
    }
    __syncthreads();
    grad_weights<kD>(t.dqkv, ldQ, t.x, ldD, rows, kQ, part + O::win, part + O::bin);       // dWin, dbin
    gemm_tile<kQ, true, 2>(t.dqkv, ldQ, w + O::win, kD, nullptr, t.g, ldD, rows, kD);         // g = d x = d r1 + d qkv Win
    __syncthreads();
    for (int i = threadIdx.x; i < rows * (kD / 4); i += kThreads) {
      const int m = i / (kD / 4), c = i - m * (kD / 4);
      reinterpret_cast<float4*>(p.dx + (size_t)(tok0 + m) * kD)[c] = *reinterpret_cast<const float4*>(t.g + (size_t)m * ldD + c * 4);
    }
    __syncthreads();
  }
}

int check(const FusionParams& p, int d, int ffn, const char* name) {
  AFSL_REQUIRE(p.x && p.w, "%s: null pointer", name);
  AFSL_REQUIRE(d == kD && ffn == kF, "%s: only d_model=%d, dim_feedforward=%d, nhead=1 are supported (got %d, %d)", name, kD,
               kF, d, ffn);
  AFSL_REQUIRE(p.V >= 1 && p.V <= 8 && kTM % p.V == 0, "%s: V=%d views (supported: 1, 2, 4, 8)", name, p.V);
  AFSL_REQUIRE(p.N >= 0, "%s: N=%d", name, p.N);
  return AFSL_OK;
}

}  // namespace
}  // namespace afsl

extern "C" int afsl_view_fusion_weight_floats(void) { return afsl::WeightOffsets::total; }
extern "C" int afsl_view_fusion_param_floats(void) { return afsl::WeightOffsets::count; }

extern "C" int afsl_view_fusion_grid(int N, int V) {
  const long long tiles = ((long long)N * V + afsl::kTM - 1) / afsl::kTM;
  const long long cap = afsl::kNumSMs;
  return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

extern "C" int afsl_view_fusion_fwd_f32(const float* x, const float* weights, float* y, const float* drop_attn,
                                         const float* drop1, const float* drop_ffn, const float* drop2, int N, int V, int d,
                                         int ffn, void* stream) {
  using namespace afsl;
  FusionParams p{};
  p.x = x; p.w = weights; p.y = y; p.drop_attn = drop_attn; p.drop1 = drop1; p.dropf = drop_ffn; p.drop2 = drop2;
  p.N = N; p.V = V;
  if (int rc = check(p, d, ffn, "afsl_view_fusion_fwd_f32")) return rc;
  AFSL_REQUIRE(y, "afsl_view_fusion_fwd_f32: null output");
  if (N == 0) return AFSL_OK;
  const size_t bytes = tile_floats(false) * sizeof(float);
  if (int rc = opt_in_smem(fusion_fwd_kernel, bytes, "afsl_view_fusion_fwd_f32")) return rc;
  const long long tiles = ((long long)N * V + kTM - 1) / kTM;
  const int cap = persistent_grid(fusion_fwd_kernel, kThreads, bytes, 1 << 30);
  fusion_fwd_kernel<<<(int)(tiles < cap ? tiles : cap), kThreads, bytes, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_view_fusion_fwd_f32");
  return AFSL_OK;
}

extern "C" int afsl_view_fusion_bwd_f32(const float* x, const float* weights, const float* d_y, const float* drop_attn,
                                         const float* drop1, const float* drop_ffn, const float* drop2, float* d_x,
                                         float* d_weights_partial, int N, int V, int d, int ffn, void* stream) {
  using namespace afsl;
  FusionParams p{};
  p.x = x; p.w = weights; p.dy = d_y; p.dx = d_x; p.dw_part = d_weights_partial;
  p.drop_attn = drop_attn; p.drop1 = drop1; p.dropf = drop_ffn; p.drop2 = drop2;
  p.N = N; p.V = V;
  if (int rc = check(p, d, ffn, "afsl_view_fusion_bwd_f32")) return rc;
  AFSL_REQUIRE(d_y && d_x && d_weights_partial, "afsl_view_fusion_bwd_f32: null pointer");
  if (N == 0) return AFSL_OK;
  const size_t bytes = tile_floats(true) * sizeof(float);
  if (int rc = opt_in_smem(fusion_bwd_kernel, bytes, "afsl_view_fusion_bwd_f32")) return rc;
  fusion_bwd_kernel<<<afsl_view_fusion_grid(N, V), kThreads, bytes, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_view_fusion_bwd_f32");
  return AFSL_OK;
}
