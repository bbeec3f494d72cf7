// Angular loss with angular mining, one WARP per episode (same arithmetic and reference as angular.cu:
// AngularLossClass.forward, loops/loss.py:48-97 + the pytorch_metric_learning miner / loss restated in
// oracle/angular.py), for the NSynth-shaped head: Dp = 64 and W + Nq <= 32 pooled rows.
//
// Lane i owns pooled row i (prototypes first, then queries): its normalised row lives in 64 registers, the
// other rows are broadcast out of the warp's shared-memory slice, so the N x N Gram matrix, the mined-triplet
// counts, the per-pair log-sum-exp and dL/dGram are all produced without a cross-lane reduction over the
// embedding dimension and without a CTA barrier.  Pairs are enumerated from per-lane bit masks (one bit per
// partner row), negatives from a per-lane mask, so all lanes walk their t-th pair together.
// dL/dGram rows are owned by one lane per pass (pass 1: the positive's row, pass 0: the anchor's row), which
// keeps the accumulation order fixed - the result is deterministic.
#include "angular.cuh"
#include "warp_rows.cuh"

namespace afsl {
namespace {

using namespace warp_rows;

constexpr int kD = 64;             // embedding dimension handled here
constexpr int kLd = kD + 4;        // row stride of the staged rows: conflict-free 128-bit own-row accesses
constexpr int kNMax = 32;          // pooled rows per episode (one per lane)
constexpr int kLg = kNMax + 1;     // Gram row stride: conflict-free row and column walks
constexpr float kNormEps = 1e-12f; // F.normalize
constexpr float kPairEps = 1e-6f;  // F.pairwise_distance

// per-warp shared-memory slice (floats): rows | gram | dgram | nu | norm | csum | gsum | manc (int) | cnt (bytes)
constexpr int kSliceWords = kNMax * kLd + 2 * kNMax * kLg + 5 * kNMax + kNMax * kNMax / 4;

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// masked, weighted log-sum-exp with the appended zero of one (anchor, positive) pair: max and normalised sum
__device__ __forceinline__ void pair_lse(const AngParams& p, const float* gram, const float* nu, const float* norm, int N,
                                         unsigned negmask, int a, int q, float base, float& m, float& tot) {
  float mx = 0.f;
  for (int k = 0; k < N; ++k) {
    if (!((negmask >> k) & 1u) || nu[k] <= 0.f) continue;
    const float rho = p.normalize_ref ? 1.f : norm[k];
    mx = fmaxf(mx, fmaf(4.f * p.t2 * rho, gram[a * kLg + k] + gram[q * kLg + k], base));
  }
  float acc = expf(-mx);
  for (int k = 0; k < N; ++k) {
    if (!((negmask >> k) & 1u) || nu[k] <= 0.f) continue;
    const float rho = p.normalize_ref ? 1.f : norm[k];
    const float f = fmaf(4.f * p.t2 * rho, gram[a * kLg + k] + gram[q * kLg + k], base);
    acc = fmaf(nu[k], expf(f - mx), acc);
  }
  m = mx;
  tot = acc;
}

template <bool kBwd>
__global__ void __launch_bounds__(kCtaThreads) angular_warp_kernel(const AngParams p) {
  extern __shared__ __align__(16) float smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * kWarpsPerCta + warp;
  if (e >= p.E) return;
  const int W = p.W, Nq = p.Nq, N = W + Nq;
  float* xh = smem_raw + (size_t)warp * kSliceWords;    // [kNMax][kLd] rows (raw, then normalised)
  float* gram = xh + kNMax * kLd;                       // [kNMax][kLg]
  float* dgram = gram + kNMax * kLg;                    // [kNMax][kLg]
  float* nu = dgram + kNMax * kLg;                      // [kNMax] negative weight
  float* norm = nu + kNMax;                             // [kNMax] raw norms
  float* csum = norm + kNMax;                           // [kNMax] component sum of the normalised row
  float* gsq = csum + kNMax;                            // [kNMax] sum_k g of the pair whose positive is row q (anchors branch)
  int* manc = reinterpret_cast<int*>(gsq + kNMax);      // [kNMax] mined triplets per anchor
  unsigned char* cnt = reinterpret_cast<unsigned char*>(manc + kNMax);   // [kNMax][kNMax] mined negatives per (a, q)
  const bool row = lane < N;

  // ---- stage the episode's rows with coalesced 128-bit loads (two 256 B rows per warp instruction)
  {
    const float4* p4 = reinterpret_cast<const float4*>(p.protos + (size_t)e * W * kD);
    const float4* q4 = reinterpret_cast<const float4*>(p.queries + (size_t)e * Nq * kD);
    for (int i = lane; i < N * (kD / 4); i += 32) {
      const int r = i / (kD / 4), c = i - r * (kD / 4);
      const float4 v = r < W ? __ldg(p4 + i) : __ldg(q4 + (i - W * (kD / 4)));
      *reinterpret_cast<float4*>(xh + r * kLd + 4 * c) = v;
    }
    for (int i = lane; i < kNMax * kNMax / 4; i += 32) reinterpret_cast<int*>(cnt)[i] = 0;
    manc[lane] = 0;
  }
  const int lab = row ? (lane < W ? lane : p.labels[(size_t)e * Nq + (lane - W)]) : -1;
  __syncwarp();

  // ---- own row: norm, normalise, component sum; normalised row back to shared for the other lanes
  f32x2 x[kD / 2];
  float nrm = 0.f, cs = 0.f;
  {
#pragma unroll
    for (int c = 0; c < kD / 4; ++c) {
      const float4 v = lds4(xh + (row ? lane : 0) * kLd + 4 * c);
      x[2 * c] = pack2(v.x, v.y);
      x[2 * c + 1] = pack2(v.z, v.w);
    }
    f32x2 a0 = 0ull, a1 = 0ull;
#pragma unroll
    for (int c = 0; c < kD / 2; c += 2) { a0 = fma2(x[c], x[c], a0); a1 = fma2(x[c + 1], x[c + 1], a1); }
    nrm = sqrtf(sum2(a0) + sum2(a1));
    const float r = 1.f / fmaxf(nrm, kNormEps);
    const f32x2 r2 = pack2(r, r);
    f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
    for (int c = 0; c < kD / 2; c += 2) {
      x[c] = mul2(x[c], r2); x[c + 1] = mul2(x[c + 1], r2);
      s0 = add2(s0, x[c]); s1 = add2(s1, x[c + 1]);
    }
    cs = sum2(s0) + sum2(s1);
    if (row) {
#pragma unroll
      for (int c = 0; c < kD / 4; ++c) {
        float4 v;
        unpack2(x[2 * c], v.x, v.y);
        unpack2(x[2 * c + 1], v.z, v.w);
        *reinterpret_cast<float4*>(xh + lane * kLd + 4 * c) = v;
      }
      norm[lane] = nrm;
      csum[lane] = cs;
    }
  }
  __syncwarp();

  // ---- Gram matrix: lane i computes row i against every broadcast row j
  for (int j = 0; j < N; ++j) {
    f32x2 a0 = 0ull, a1 = 0ull;
#pragma unroll
    for (int c = 0; c < kD / 4; ++c) {
      const float4 v = lds4(xh + j * kLd + 4 * c);
      a0 = fma2(x[2 * c], pack2(v.x, v.y), a0);
      a1 = fma2(x[2 * c + 1], pack2(v.z, v.w), a1);
    }
    if (row) gram[lane * kLg + j] = sum2(a0) + sum2(a1);
  }
  __syncwarp();

  // ---- pair / negative masks.  mask1: anchors a of the pairs (a, q = lane); mask0: positives q of the pairs
  //      (a = lane, q); negmask: rows usable as negatives for any pair of this lane's label
  unsigned mask1 = 0, mask0 = 0, negmask = 0;
  for (int j = 0; j < N; ++j) {
    const int lj = __shfl_sync(kFull, lab, j);
    if (!row) continue;
    if (p.anchors) {
      if (lane < W && j >= W && lj == lane) mask0 |= 1u << j;
      if (j >= W && lj != lab) negmask |= 1u << j;
    } else {
      if (j != lane && lj == lab) { mask0 |= 1u << j; mask1 |= 1u << j; }
      if (lj != lab) negmask |= 1u << j;
    }
  }
  if (p.anchors && row && lane >= W && lab >= 0 && lab < W) mask1 = 1u << lab;
  const int q = row ? lane : 0;      // this lane's row index as positive (pass 1) or anchor (pass 0)

  // ---- mining: count the negatives passing the angle test per pair; times each row was mined as a negative
  const float deps = (float)kD * kPairEps * kPairEps;
  const float gdiag = gram[q * kLg + q];
  int wneg = 0;
  for (unsigned m1 = mask1; __any_sync(kFull, m1 != 0);) {
    const bool act = m1 != 0;
    const int a = act ? __ffs(m1) - 1 : 0;
    m1 &= m1 - 1;
    const float gaq = gram[a * kLg + q];
    const float ap2 = gram[a * kLg + a] + gdiag - 2.f * gaq + 2.f * kPairEps * (csum[a] - cs) + deps;
    const float ap = sqrtf(fmaxf(ap2, 0.f));
    const float ra = norm[a], rq = nrm;
    const float sum_norm = sqrtf(fmaxf(ra * ra + rq * rq + 2.f * ra * rq * gaq, 0.f));
    const float inv = 1.f / fmaxf(sum_norm, kNormEps);
    const float cc = sum_norm > kNormEps ? 1.f : (sum_norm * inv) * (sum_norm * inv);
    const float csum_c = (ra * csum[a] + rq * cs) * inv;
    int count = 0;
    if (p.miner_tan == 0.f) {
      // angle 0: atan(ap / (2 nc)) > 0 iff ap > 0 for any finite nc (ap > 2 nc * 0): the distance to the pair centre is
      // not needed, every negative of a pair with distinct anchor / positive passes
      const bool on = act && !p.miner_never && ap > 0.f;
      for (int k = 0; k < N; ++k) {
        const bool pass = on && ((negmask >> k) & 1u);
        count += pass;
        const unsigned b = __ballot_sync(kFull, pass);
        if (lane == k) wneg += __popc(b);
      }
    } else
    for (int k = 0; k < N; ++k) {
      const float gkk = __shfl_sync(kFull, gdiag, k), csk = __shfl_sync(kFull, cs, k);
      const float dot = (ra * gram[k * kLg + a] + rq * gram[k * kLg + q]) * inv;
      const float nc2 = gkk + cc - 2.f * dot + 2.f * kPairEps * (csk - csum_c) + deps;
      const float nc = sqrtf(fmaxf(nc2, 0.f));
      // atan(ap / (2 nc)) > angle without the arctangent and the division: ap > 2 nc tan(angle)
      const bool pass = act && ((negmask >> k) & 1u) && !p.miner_never && ap > 2.f * nc * p.miner_tan;
      count += pass;
      const unsigned b = __ballot_sync(kFull, pass);
      if (lane == k) wneg += __popc(b);
    }
    if (act) {
      cnt[a * kNMax + q] = (unsigned char)count;
      if (count) atomicAdd(&manc[a], count);
    }
  }
  __syncwarp();
  // ---- negative weights
  {
    float w = row ? 1.f : 0.f;
    if (p.anchors) w = (row && lane >= W && lab >= 0 && lab < W) ? (float)((int)cnt[lab * kNMax + lane] + wneg) : 0.f;
    nu[lane] = w;
  }
  __syncwarp();
  const float nu_q = nu[q];

  // ---- forward: sum of pair weights and weighted pair losses (lane = positive)
  float num = 0.f, den = 0.f;
  for (unsigned m1 = mask1; __any_sync(kFull, m1 != 0);) {
    const bool act = m1 != 0;
    const int a = act ? __ffs(m1) - 1 : 0;
    m1 &= m1 - 1;
    const float omega = !act ? 0.f : p.anchors ? (float)manc[a] * nu_q : (float)cnt[a * kNMax + q];
    if (omega > 0.f) {
      float m, tot;
      pair_lse(p, gram, nu, norm, N, negmask, a, q, -2.f * (1.f + p.t2) * gram[a * kLg + q], m, tot);
      num = fmaf(omega, m + logf(tot), num);
      den += omega;
    }
  }
  num = warp_sum(num);
  den = warp_sum(den);
  if (!kBwd) {
    if (lane == 0) p.loss[e] = den > 0.f ? num / den : 0.f;
    return;
  }

  // ---- backward: dL/dGram, row `lane` owned by this lane in both passes
  for (int k = 0; k < kLg; ++k) dgram[lane * kLg + k] = 0.f;
  __syncwarp();
  const float scale = den > 0.f ? p.d_loss[e] / den : 0.f;
  gsq[lane] = 0.f;
  for (int pass = 1; pass >= 0; --pass) {
    if (pass == 0 && p.anchors) {
      // every query has exactly one pair, so after pass 1 row q of dL/dGram holds 4 t2 rho g of that pair:
      // the anchor's row is the sum of its positives' rows (ascending q), plus the -2(1+t2) sum_k g terms
      for (unsigned mm = mask0; mm != 0; mm &= mm - 1) {
        const int o = __ffs(mm) - 1;
        for (int k = 0; k < N; ++k) dgram[q * kLg + k] += dgram[o * kLg + k];
      }
      for (unsigned mm = mask0; mm != 0; mm &= mm - 1) {
        const int o = __ffs(mm) - 1;
        dgram[q * kLg + o] += -2.f * (1.f + p.t2) * gsq[o];
      }
      __syncwarp();
      break;
    }
    for (unsigned mm = pass ? mask1 : mask0; mm != 0; mm &= mm - 1) {     // no warp-level primitive inside: lanes run free
      const int o = __ffs(mm) - 1;
      const int a = pass ? o : q, qq = pass ? q : o;
      const float omega = p.anchors ? (float)manc[a] * nu[qq] : (float)cnt[a * kNMax + qq];
      if (omega <= 0.f) continue;
      float m, tot;
      const float base = -2.f * (1.f + p.t2) * gram[a * kLg + qq];
      pair_lse(p, gram, nu, norm, N, negmask, a, qq, base, m, tot);
      const float coef = scale * omega / tot;
      float gsum = 0.f;
      for (int k = 0; k < N; ++k) {
        if (!((negmask >> k) & 1u) || nu[k] <= 0.f) continue;
        const float rho = p.normalize_ref ? 1.f : norm[k];
        const float f = fmaf(4.f * p.t2 * rho, gram[a * kLg + k] + gram[qq * kLg + k], base);
        const float g = coef * nu[k] * expf(f - m);
        gsum += g;
        dgram[q * kLg + k] += 4.f * p.t2 * rho * g;
      }
      if (pass == 0) dgram[a * kLg + qq] += -2.f * (1.f + p.t2) * gsum;
      else gsq[q] = gsum;
    }
    __syncwarp();
  }
  // d rho_k = sum_i G[i,k] dG[i,k] / rho_k over rows i of another label (those entries hold only 4 t2 rho g sums)
  float drho = 0.f;
  for (int i = 0; i < N; ++i) {
    const int li = __shfl_sync(kFull, lab, i);
    if (row && li != lab) drho = fmaf(gram[i * kLg + q], dgram[i * kLg + q], drho);
  }
  drho = (!p.normalize_ref && row && nrm > 0.f) ? drho / nrm : 0.f;

  // d x^_i = sum_j (dG[i,j] + dG[j,i]) x^_j ;  d x_i = (d x^_i - x^_i <x^_i, d x^_i>) / |x_i| + d rho_i x^_i
  f32x2 acc[kD / 2];
#pragma unroll
  for (int c = 0; c < kD / 2; ++c) acc[c] = 0ull;
  for (int j = 0; j < N; ++j) {
    const float w = dgram[q * kLg + j] + dgram[j * kLg + q];
    const f32x2 w2 = pack2(w, w);
#pragma unroll
    for (int c = 0; c < kD / 4; ++c) {
      const float4 v = lds4(xh + j * kLd + 4 * c);
      acc[2 * c] = fma2(w2, pack2(v.x, v.y), acc[2 * c]);
      acc[2 * c + 1] = fma2(w2, pack2(v.z, v.w), acc[2 * c + 1]);
    }
  }
  if (row) {
    const float inv_i = 1.f / fmaxf(nrm, kNormEps);
    const bool clamped = !(nrm > kNormEps);
    f32x2 d0 = 0ull, d1 = 0ull;
#pragma unroll
    for (int c = 0; c < kD / 4; ++c) {
      const float4 v = lds4(xh + lane * kLd + 4 * c);
      x[2 * c] = pack2(v.x, v.y);
      x[2 * c + 1] = pack2(v.z, v.w);
      d0 = fma2(acc[2 * c], x[2 * c], d0);
      d1 = fma2(acc[2 * c + 1], x[2 * c + 1], d1);
    }
    const float dot = clamped ? 0.f : sum2(d0) + sum2(d1);
    const f32x2 nd2 = pack2(-dot, -dot), i2 = pack2(inv_i, inv_i), r2 = pack2(drho, drho);
    float* out = lane < W ? p.d_protos + ((size_t)e * W + lane) * kD : p.d_queries + ((size_t)e * Nq + (lane - W)) * kD;
#pragma unroll
    for (int c = 0; c < kD / 4; ++c) {
      const f32x2 o0 = fma2(r2, x[2 * c], mul2(fma2(nd2, x[2 * c], acc[2 * c]), i2));
      const f32x2 o1 = fma2(r2, x[2 * c + 1], mul2(fma2(nd2, x[2 * c + 1], acc[2 * c + 1]), i2));
      float4 v;
      unpack2(o0, v.x, v.y);
      unpack2(o1, v.z, v.w);
      *reinterpret_cast<float4*>(out + 4 * c) = v;
    }
  }
}

}  // namespace

int launch_angular_warp(const AngParams& p, bool bwd, cudaStream_t stream, const char* name, bool* handled) {
  *handled = false;
  if (p.D != kD || p.W + p.Nq > kNMax) return AFSL_OK;
  *handled = true;
  auto fn = bwd ? angular_warp_kernel<true> : angular_warp_kernel<false>;
  const size_t bytes = (size_t)kWarpsPerCta * kSliceWords * sizeof(float);
  if (int rc = opt_in_smem(fn, bytes, name)) return rc;
  const int grid = (p.E + kWarpsPerCta - 1) / kWarpsPerCta;
  fn<<<grid, kCtaThreads, bytes, stream>>>(p);
  AFSL_CHECK_LAUNCH(name);
  return AFSL_OK;
}

}  // namespace afsl
