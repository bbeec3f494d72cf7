// Angular loss, prototypes-as-anchors branch (AngularLossClass.forward, loops/loss.py:68-83 + the pytorch_metric_learning
// miner / loss restated in oracle/angular.py), Dp = 64 and W + Nq <= 31 pooled rows, on the tensor cores.
//
// What the loss needs per episode is the Gram matrix of its 30 pooled rows (forward) and, in the backward, the product
// dL/dGram . X: three 30 x 30 x 64 contractions that kept the one-warp-per-episode kernel (angular_warp.cu) at 0.07-0.11 of
// the HBM roofline on the fp32 pipe.  Here FOUR episodes share one 128-row tcgen05 tile in SPLIT FP16: every operand is
// x = h0 + h1 (two fp16, 22 bits), three passes h1.h0 + h0.h1 + h0.h0 with fp32 accumulation in TMEM.  fp16 has TF32's
// mantissa at four times its rate and twice its K per instruction (half the accumulator traffic per MAC: the first, TF32,
// version was bound by TMEM contention between its K = 8 MMAs and the epilogue's tcgen05.ld, profiles/r2q_*timeline*); its
// narrow exponent is harmless because the producers NORMALISE the rows first (|x^| <= 1: absolute representation error
// <= 2^-25 per element, ~6e-8 on a cosine) and the backward operand is scaled per row by a power of two.
//
//   rows 32 e + i of the tile: x^ of the prototypes (i < W) and queries (W <= i < N), zero rows, and row 31 = ones, so that
//   column 31 of the Gram block is the component sum the miner's pairwise_distance epsilon needs;
//   MMA 1   D1[128 x 128] = X^ X^^T: the 32 x 32 diagonal blocks are the episodes' cosine matrices;
//   E1      one warp per episode, lane = row: ONE tcgen05.ld of its 32 cosines into the warp's scratch (the accumulator is
//           free again at once), mining counts from ballots, the weighted log-sum-exp of its (prototype, query) pair, loss;
//           in the backward dL/dGram rows (transposed through the scratch, per-class sums in ascending row order:
//           deterministic) with the row normalisation's backward folded into the diagonal:
//           B[i][k] = r_i G'[i][k] + delta_ik (drho_i - r_i <x^_i, dx^_i>),  so that  dx_i = sum_k B[i][k] x^_k  exactly;
//   MMA 2   (dX)^T: D2[2 x 64 dims][2 x 32 rows] = X^^T-tile . B-rows^T per K half, every operand K-major (the producers
//           keep a transposed copy of the tile; MN-major operands: tools/micro/umma_layout_probe.cu);
//   E2      lane = embedding dimension, columns = rows: un-scaling and coalesced 128-byte stores of dP / dQ.
//
// One persistent CTA per SM, 17 warps, mbarriers only: 2 x 4 producer / E2 warps (warp = episode slot, the two sets take
// alternate tiles: 16 LDG.128 per lane straight from HBM, issued TWO tiles ahead and held in registers - 61 KB in flight
// per SM; norms, normalisation, fp16 split into the set's X buffer; transposed copy; E2 of the set's previous tile),
// 2 x 4 E1 warps (alternate tiles, two D1 accumulators), 1 issuer warp.  Every role is a compact loop over 4-column chunks
// of rows kept in shared memory (the first version, unrolled over 32-wide register arrays, was 8000 SASS instructions
// executed once per tile and bound by instruction fetch).
#include <cuda_fp16.h>

#include "angular.cuh"
#include "tc_common.cuh"

namespace afsl {
namespace {

using namespace tc;

constexpr int kD = 64;
constexpr int kBlk = 32;                      // tile rows (and TMEM lanes) per episode
constexpr int kEp = 4;                        // episodes per tile
constexpr int kMaxW = 8;
constexpr int kOnesRow = 31;
constexpr int kTile = 128 * 128;              // bytes of a [128 rows x 64 fp16] swizzled tile
constexpr int kLs = 36;                       // row stride (floats) of the per-warp scratch matrices: conflict-free
constexpr float kNormEps = 1e-12f;            // F.normalize
constexpr float kPairEps = 1e-6f;             // F.pairwise_distance
constexpr unsigned kFull = 0xffffffffu;

constexpr int kPeWarps = 8, kE1Warps = 8;                      // two sets of four each, alternating tiles
constexpr int kFirstE1 = kPeWarps, kIssuer = kFirstE1 + kE1Warps;
constexpr int kThreads = (kIssuer + 1) * 32;                    // 544

// per-E1-warp scratch (floats)
struct Scratch {
  float gn[kBlk * kLs];        // cosine rows of the block; in the backward the lane's own row becomes its B row
  float fx[kBlk * kLs];        // per-lane rows: exponents -> exponentials -> T (dL/dGram of the pairs, row = positive)
  float sp[kMaxW * kLs];       // per-class column sums of T = prototype rows of dL/dGram
  float nrm[kBlk], csn[kBlk], gdi[kBlk], nu[kBlk], gs[kBlk], coef[kBlk];
  int manc[kMaxW];
  int lab[kBlk];               // class of every row (-1: not a query of a known class)
};

struct Shared {
  float nrm[2][2][128];        // row norms of the tile in X buffer `set`, tile (iteration / 2) & 1 (producers -> E1)
  float rs[2][128];            // power-of-two row scales of the B operand of tile parity b (E1 -> E2)
  // mbarrier waits only tell adjacent phases apart, so a barrier that a role observes every OTHER tile (the two producer
  // sets, the two E1 sets) exists once per tile parity: index = tile iteration & 1, phase = iteration >> 1
  uint64_t xt_full, gp_full, gp_free;
  uint64_t x_full[2], x_free[2], xt_free[2], d2_full[2], d2_free[2], d1_full[2], d1_free[2];
  uint32_t tmem_base;
};

// operand tiles (bytes from the 1 KB aligned base): X[buffer][h0, h1], XT[h0, h1], GP[h0, h1] (64 rows each)
constexpr uint32_t kX = 0, kXT = 4 * kTile, kGP = 6 * kTile;
constexpr uint32_t kOpBytesBwd = 7 * kTile, kOpBytesFwd = 4 * kTile;

__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    if (!done) __nanosleep(32);
  } while (!done);
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float pick(const float4& t, int u) { return u == 0 ? t.x : u == 1 ? t.y : u == 2 ? t.z : t.w; }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// x = h0 + h1 in fp16 (round to nearest twice), two values per conversion instruction
__device__ __forceinline__ void split_h2(float x, float y, uint32_t& h0, uint32_t& h1) {
  const __half2 a = __floats2half2_rn(x, y);
  const float2 f = __half22float2(a);
  const __half2 b = __floats2half2_rn(x - f.x, y - f.y);
  h0 = *reinterpret_cast<const uint32_t*>(&a);
  h1 = *reinterpret_cast<const uint32_t*>(&b);
}
// four values -> 8 bytes of the h0 tile and 8 bytes of the h1 tile
__device__ __forceinline__ void split_h4(const float4& v, uint32_t (&h0)[2], uint32_t (&h1)[2]) {
  split_h2(v.x, v.y, h0[0], h1[0]);
  split_h2(v.z, v.w, h0[1], h1[1]);
}
// fp16 x fp16 -> fp32 instruction descriptor, both operands K-major
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.b32 p, %5, 0;\n mov.b64 da, {%1, %3};\n mov.b64 db, {%2, %3};\n"
      " tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// timeline of CTA 0 (AFSL_ANGULAR_DBG=1): SM clock at event `ev` of tile iteration `it`, one warp per role
#define DBG(ev)                                                                                              \
  do {                                                                                                       \
    if (p.dbg && blockIdx.x == 0 && (threadIdx.x & 127) == 0 && it < 24) p.dbg[it * 32 + (ev)] = clock64(); \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------- E1
// One warp, lane = row of the episode's block.  `tm` = TMEM address of the block's 32 cosine columns in this warp's lanes,
// `nrm_blk` = the block's 32 row norms (producers' buffer), `gp_row` = this lane's row of the B operand (h0 tile; the h1 tile
// is gp_h1 bytes further; gp_xor = the row's swizzle combined with its 64-byte half), `rs_out` = where the row's
// power-of-two scale goes.
template <bool kBwd>
__device__ __forceinline__ void e1_episode(const AngParams& p, int ep, int lab_in, uint32_t tm, Scratch* sc, const float* nrm_blk,
                                           uint64_t* d1_free, uint64_t* gp_free, uint32_t gp_parity, uint32_t gp_row,
                                           uint32_t gp_h1, int gp_xor, float* rs_out, long long* dbg) {
#define DBGE(ev)                                             \
  do {                                                       \
    if (dbg && (threadIdx.x & 31) == 0) dbg[ev] = clock64(); \
  } while (0)
  const int lane = threadIdx.x & 31;
  const int W = p.W, N = p.W + p.Nq;
  const bool qv = lane >= W && lane < N && lab_in >= 0 && lab_in < W;
  const int lab = lane < W ? lane : (qv ? lab_in : -1);
  const int a = qv ? lab : 0;
  const float c4 = 4.f * p.t2, cneg = -2.f * (1.f + p.t2);
  float* gn_row = sc->gn + lane * kLs;
  float* fx_row = sc->fx + lane * kLs;
  const float* pa_row = sc->gn + a * kLs;                       // cosines of my prototype

  // ---- the block's cosine rows and norms: accumulator / producers' buffer -> scratch, then both are free again
  float cs;
  const float nrm = nrm_blk[lane];
  {
    uint32_t v[32];
    tmem_ld32(tm, v);
    sc->nrm[lane] = nrm;
    fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(d1_free);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<float4*>(gn_row + 4 * c) = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]),
                                                               __uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3]));
    cs = __uint_as_float(v[kOnesRow]);                          // component sum of the normalised row
  }
  const float ri = 1.f / fmaxf(nrm, kNormEps);
  if (lane < kMaxW) sc->manc[lane] = 0;
  sc->csn[lane] = cs;
  sc->lab[lane] = qv ? lab : -1;
  __syncwarp();
  const float gdi = gn_row[lane];
  sc->gdi[lane] = gdi;
  const float gaq = pa_row[lane];                               // cos(prototype a, this query)
  __syncwarp();
  DBGE(13);

  // ---- masks: queries of the episode, lanes of my class, my pair's negatives
  const unsigned allq = __ballot_sync(kFull, qv);
  const unsigned sameq = __match_any_sync(kFull, lab) & allq;
  const unsigned negmask = qv ? (allq & ~sameq) : 0u;

  // ---- mining (AngularMiner): negatives of my pair that pass the angle test; times my row was mined as a negative
  const float deps = (float)kD * kPairEps * kPairEps;
  const float ap2 = sc->gdi[a] + gdi - 2.f * gaq + 2.f * kPairEps * (sc->csn[a] - cs) + deps;
  const float ap = sqrtf(fmaxf(ap2, 0.f));
  int count = 0, wneg = 0;
  if (p.miner_tan == 0.f) {
    // angle 0: atan(ap / (2 nc)) > 0 iff ap > 0: every negative of a pair with distinct anchor / positive passes
    const bool on = qv && !p.miner_never && ap > 0.f;
    const unsigned onmask = __ballot_sync(kFull, on);
    count = on ? __popc(negmask) : 0;
    wneg = qv ? __popc(onmask & ~sameq) : 0;
  } else {
    const float ra = sc->nrm[a], rq = nrm;
    const float sum_norm = sqrtf(fmaxf(ra * ra + rq * rq + 2.f * ra * rq * gaq, 0.f));
    const float inv = 1.f / fmaxf(sum_norm, kNormEps);
    const float cc = sum_norm > kNormEps ? 1.f : (sum_norm * inv) * (sum_norm * inv);
    const float csum_c = (ra * sc->csn[a] + rq * cs) * inv;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      const float4 g4 = ld4(gn_row + 4 * c), s4 = ld4(pa_row + 4 * c), gd4 = ld4(sc->gdi + 4 * c), cs4 = ld4(sc->csn + 4 * c);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = 4 * c + u;
        const float dot = (ra * pick(s4, u) + rq * pick(g4, u)) * inv;
        const float nc2 = pick(gd4, u) + cc - 2.f * dot + 2.f * kPairEps * (pick(cs4, u) - csum_c) + deps;
        const float nc = sqrtf(fmaxf(nc2, 0.f));
        // atan(ap / (2 nc)) > angle without the arctangent and the division: ap > 2 nc tan(angle)
        const bool pass = ((negmask >> k) & 1u) && !p.miner_never && ap > 2.f * nc * p.miner_tan;
        count += pass;
        const unsigned b = __ballot_sync(kFull, pass);
        if (lane == k) wneg = __popc(b);
      }
    }
  }
  if (count) atomicAdd(&sc->manc[a], count);
  const float nu_i = qv ? (float)(count + wneg) : 0.f;
  sc->nu[lane] = nu_i;
  __syncwarp();
  DBGE(14);

  // ---- my pair: weighted log-sum-exp over the negatives, with the appended zero
  const float omega = qv ? (float)sc->manc[a] * nu_i : 0.f;
  const float base = cneg * gaq;
  float mx = 0.f, tot = 1.f, gsum_un = 0.f;
  if (!kBwd) {
    // forward only: ONE pass, running maximum with rescaling of the running sum (the appended zero is the initial term)
#pragma unroll 4
    for (int c = 0; c < 8; ++c) {
      const float4 g4 = ld4(gn_row + 4 * c), s4 = ld4(pa_row + 4 * c), nu4 = ld4(sc->nu + 4 * c);
      float4 rho = make_float4(1.f, 1.f, 1.f, 1.f);
      if (!p.normalize_ref) rho = ld4(sc->nrm + 4 * c);
      float f[4], m_new = mx;
      bool use[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        f[u] = fmaf(c4 * pick(rho, u), pick(s4, u) + pick(g4, u), base);
        use[u] = ((negmask >> (4 * c + u)) & 1u) && pick(nu4, u) > 0.f;
        m_new = use[u] ? fmaxf(m_new, f[u]) : m_new;
      }
      float part = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) part += use[u] ? pick(nu4, u) * __expf(f[u] - m_new) : 0.f;   // a select: see below
      tot = fmaf(tot, __expf(mx - m_new), part);
      mx = m_new;
    }
  } else {
  #pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      const float4 g4 = ld4(gn_row + 4 * c), s4 = ld4(pa_row + 4 * c), nu4 = ld4(sc->nu + 4 * c);
      float4 rho = make_float4(1.f, 1.f, 1.f, 1.f);
      if (!p.normalize_ref) rho = ld4(sc->nrm + 4 * c);
      float f[4];
  #pragma unroll
      for (int u = 0; u < 4; ++u) {
        f[u] = fmaf(c4 * pick(rho, u), pick(s4, u) + pick(g4, u), base);
        const bool use = ((negmask >> (4 * c + u)) & 1u) && pick(nu4, u) > 0.f;
        mx = use ? fmaxf(mx, f[u]) : mx;
      }
      *reinterpret_cast<float4*>(fx_row + 4 * c) = make_float4(f[0], f[1], f[2], f[3]);
    }
    DBGE(15);
    tot = __expf(-mx);
    float sumex = 0.f;
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      const float4 f4 = ld4(fx_row + 4 * c), nu4 = ld4(sc->nu + 4 * c);
      float4 rho = make_float4(1.f, 1.f, 1.f, 1.f);
      if (!p.normalize_ref) rho = ld4(sc->nrm + 4 * c);
      float tu[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        // a select, not a product with a zero weight: the exponent of a masked column may overflow
        const bool use = ((negmask >> (4 * c + u)) & 1u) && pick(nu4, u) > 0.f;
        const float ex = use ? pick(nu4, u) * __expf(pick(f4, u) - mx) : 0.f;
        sumex += ex;
        tu[u] = c4 * pick(rho, u) * ex;     // dL/dGram[i][k] of my pair up to the row's coefficient (applied where it is read)
      }
      *reinterpret_cast<float4*>(fx_row + 4 * c) = make_float4(tu[0], tu[1], tu[2], tu[3]);
    }
    tot += sumex;
    gsum_un = sumex;
  }
  const float term = omega > 0.f ? omega * (mx + logf(tot)) : 0.f;
  const float num = warp_sum(term), den = warp_sum(omega);
  DBGE(16);
  if (!kBwd) {
    if (lane == 0) p.loss[ep] = den > 0.f ? num / den : 0.f;
    return;
  }

  // ---- backward.  T_i[k] = coef_i 4 t2 rho_k nu_k exp(f_k - m): dL/dGram[i][k] of my pair (and of its anchor's row); the
  //      scratch rows hold it without coef_i, which is known only now; gsum = sum_k g_k
  const float scale = den > 0.f ? p.d_loss[ep] / den : 0.f;
  const float coef = omega > 0.f ? scale * omega / tot : 0.f;
  const float gsum = coef * gsum_un;
  sc->coef[lane] = coef;
  sc->gs[lane] = gsum;
  __syncwarp();                                                 // T rows complete; every lane is done with the prototype rows
  DBGE(17);
  // per-class column sums (rows in ascending order): pr[w] = dL/dGram[prototype w][this lane's row].  Branch-free: the
  // row loads do not depend on the class masks, so they stream; the class of a row (the same for every lane) only selects
  // the accumulator (the first version walked per-class bit masks: find-first-set -> address -> load -> add, one
  // dependent chain per row behind uniform branches, 4800 clocks of the 15000 of this epilogue)
  {
    float pr[kMaxW];
#pragma unroll
    for (int w = 0; w < kMaxW; ++w) pr[w] = 0.f;
#pragma unroll 5
    for (int q = W; q < N; ++q) {
      const float t = sc->coef[q] * sc->fx[q * kLs + lane];
      const int wq = sc->lab[q];
#pragma unroll
      for (int w = 0; w < kMaxW; ++w) pr[w] += wq == w ? t : 0.f;
    }
#pragma unroll
    for (int w = 0; w < kMaxW; ++w)
      if (w < W) sc->sp[w * kLs + lane] = pr[w];
  }
  __syncwarp();
  DBGE(18);
  // B row of this lane: r_i (dL/dGram[i][k] + dL/dGram[k][i]) off the diagonal, written over its own cosine row (nobody
  // reads another lane's query row, and the prototype rows were last read before the barriers above)
  const bool isp = lane < W;
  const float* sp_row = sc->sp + (isp ? lane : 0) * kLs;
  float drho = 0.f, dot = 0.f, amax = 0.f;
#pragma unroll 2
  for (int c = 0; c < 8; ++c) {
    const float4 g4 = ld4(gn_row + 4 * c), t4 = ld4(fx_row + 4 * c), sp4 = ld4(sp_row + 4 * c), gs4 = ld4(sc->gs + 4 * c);
    const float4 cf4 = ld4(sc->coef + 4 * c);
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = 4 * c + u;
      const float gk = pick(g4, u);                             // cos(row i, row k)
      const float col = pick(cf4, u) * sc->fx[k * kLs + lane];  // T_k[i]
      drho = fmaf(col, gk, drho);
      const float vq = fmaf(coef, pick(t4, u), col);
      const float vp = ((sameq >> k) & 1u) ? cneg * pick(gs4, u) : pick(sp4, u);
      float val = qv ? vq : (isp ? vp : 0.f);
      if (c < 2 && k < W) {                                     // prototype columns of a query's row
        const float prk = sc->sp[k * kLs + lane];               // 0 for my own class: it never uses me as a negative
        if (qv) val = (k == lab) ? cneg * gsum : prk;
        drho = fmaf(prk, gk, drho);
      }
      dot = fmaf(val, gk, dot);
      o[u] = ri * val;
      amax = fmaxf(amax, fabsf(o[u]));
    }
    *reinterpret_cast<float4*>(gn_row + 4 * c) = make_float4(o[0], o[1], o[2], o[3]);
  }
  // the diagonal carries the backward of the row normalisation: drho_i - r_i <x^_i, dx^_i>  (G'[i][i] itself is 0)
  drho = (!p.normalize_ref && qv && nrm > 0.f) ? drho / nrm : 0.f;
  if (!(nrm > kNormEps)) dot = 0.f;                             // F.normalize clamps the norm: no projection term there
  const float cdiag = drho - ri * dot;
  amax = fmaxf(amax, fabsf(cdiag));
  gn_row[lane] = cdiag;
  // power-of-two row scale: the largest entry lands in [1, 2), so both fp16 halves keep their bits; E2 undoes it exactly
  const uint32_t ebits = (__float_as_uint(amax) >> 23) & 0xffu;
  const bool tiny = ebits < 16u || ebits > 240u;                // an all-zero (or non-finite) row stays as it is
  const float up = tiny ? 1.f : __uint_as_float((254u - ebits) << 23);
  DBGE(19);
  mbar_wait_sleep(gp_free, gp_parity);                          // MMA 2 and E2 of the previous tiles are done with GP / rs
  *rs_out = tiny ? 1.f : __uint_as_float(ebits << 23);
#pragma unroll 2
  for (int c = 0; c < 8; ++c) {
    float4 v = ld4(gn_row + 4 * c);
    v.x *= up; v.y *= up; v.z *= up; v.w *= up;
    uint32_t h0[2], h1[2];
    split_h4(v, h0, h1);
    const uint32_t off = ((uint32_t)(((c >> 1) ^ gp_xor) & 7) << 4) + (c & 1) * 8;
    sts_v2(gp_row + off, h0[0], h0[1]);
    sts_v2(gp_row + gp_h1 + off, h1[0], h1[1]);
  }
}

// ------------------------------------------------------------------------------------------------------------ kernel
template <bool kBwd>
__global__ void __launch_bounds__(kThreads, 1) angular_tc_kernel(const AngParams p) {
  extern __shared__ __align__(1024) uint8_t smem_ang_raw[];
  uint8_t* smem = smem_ang_raw + ((1024u - (smem_u32(smem_ang_raw) & 1023u)) & 1023u);
  const uint32_t base = smem_u32(smem);
  constexpr uint32_t kOps = kBwd ? kOpBytesBwd : kOpBytesFwd;
  Scratch* scratch = reinterpret_cast<Scratch*>(smem + kOps);
  Shared* sh = reinterpret_cast<Shared*>(scratch + kE1Warps);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = p.W, Nq = p.Nq, N = W + Nq;
  const int tiles = (p.E + kEp - 1) / kEp;

  if (tid == 0) {
    mbar_init(&sh->xt_full, 4);
    mbar_init(&sh->gp_full, 4);
    mbar_init(&sh->gp_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sh->x_full[b], 4);
      mbar_init(&sh->x_free[b], 1);
      mbar_init(&sh->xt_free[b], 1);
      mbar_init(&sh->d2_full[b], 1);
      mbar_init(&sh->d2_free[b], 4);
      mbar_init(&sh->d1_full[b], 1);
      mbar_init(&sh->d1_free[b], 4);
    }
    fence_barrier_init();
  }
  // operand tiles start as zeros (pad rows stay zero for the whole launch); row 31 of every block of X (h0) is ones
  for (uint32_t i = tid; i < kOps / 16; i += blockDim.x) sts4(base + i * 16, make_float4(0.f, 0.f, 0.f, 0.f));
  for (int i = tid; i < 2 * 2 * 128; i += blockDim.x) (&sh->nrm[0][0][0])[i] = 0.f;
  for (int i = tid; i < 2 * 128; i += blockDim.x) (&sh->rs[0][0])[i] = 1.f;
  __syncthreads();
  for (int i = tid; i < 2 * kEp * 8; i += blockDim.x) {
    const int b = i >> 5, e = (i >> 3) & 3, c = i & 7;
    const uint32_t one2 = 0x3C003C00u;                                  // two fp16 ones
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + kX + (b * 2) * kTile + sw128(e * kBlk + kOnesRow, c)), "r"(one2)
                 : "memory");
  }
  if (warp == kIssuer) tmem_alloc<kBwd ? 512 : 256>(&sh->tmem_base);       // D1 x 2 (+ D2 x 2 K halves in the backward)
  fence_async_proxy();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = sh->tmem_base;

  if (warp < kPeWarps) {
    // =========================================================== producers + E2: warp & 3 = episode slot, warp / 4 = set
    // 17 warps leave 96 registers per thread; this role holds a 64-register prefetch and takes the 16 registers per thread
    // that the E1 warps hand back (setmaxnreg moves registers inside the CTA's own allocation only)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int set = warp >> 2, e = warp & 3;
    const int n16 = N * 16, w16 = W * 16;                               // 16-byte chunks of the block / of its prototypes
    const uint32_t xb = base + kX + (uint32_t)set * 2 * kTile;          // this set's X buffer: h0 tile, h1 at + kTile
    float4 v[16];
    auto load_tile = [&](int tile) {
      const int ep = tile * kEp + e;
      const bool valid = tile < tiles && ep < p.E;
      // chunk ci = lane + 32 j of the block (row ci / 16, 16-byte chunk ci % 16): the first 16 W chunks are the episode's
      // prototype rows, the rest its query rows; W <= 8, so only j < 4 can still be prototypes
      const float4* p4 = reinterpret_cast<const float4*>(p.protos) + (size_t)(valid ? ep : 0) * w16 + lane;
      const float4* q4 = reinterpret_cast<const float4*>(p.queries) + (size_t)(valid ? ep : 0) * (n16 - w16) + lane - w16;
      const int lim = valid ? n16 - lane : 0;                           // chunk j exists when 32 j < lim
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (32 * j < lim) v[j] = ldg_stream((j < 4 && lane + 32 * j < w16 ? p4 : q4) + 32 * j);
      }
    };
    auto e2_tile = [&](int tile, int it) {                              // lane = embedding dimension, columns = rows
      const int d = (e & 1) * 32 + lane;
      mbar_wait(&sh->d2_full[it & 1], (it >> 1) & 1);
      fence_after();
#pragma unroll 1
      for (int hh = 0; hh < 4; ++hh) {                                  // K half h (the episode), 16 rows at a time
        const int h = hh >> 1, half = hh & 1;
        const int slot = 2 * (e >> 1) + h, ep = tile * kEp + slot;
        const bool valid = ep < p.E;
        const float* rs = sh->rs[it & 1] + slot * kBlk + half * 16;
        // row i of the block lives at d_protos[dp + 64 i] (prototypes) or d_queries[dq + 64 i] (queries)
        const ptrdiff_t dp = (ptrdiff_t)(valid ? ep : 0) * W * kD + d;
        const ptrdiff_t dq = (ptrdiff_t)(valid ? ep : 0) * Nq * kD + d - (ptrdiff_t)W * kD;
        uint32_t o[16];
        tmem_ld16_nowait(tmem + ((uint32_t)(e * 32) << 16) + 256 + h * 64 + (e >> 1) * 32 + half * 16, o);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int i = half * 16 + u;
            const float val = __uint_as_float(o[u]) * rs[u];
            if (i < W) p.d_protos[dp + i * kD] = val;
            else if (i < N) p.d_queries[dq + i * kD] = val;
          }
        }
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->d2_free[it & 1]);
    };
    int it = set, prev_tile = -1, prev_it = -1;
    int tile = blockIdx.x + set * gridDim.x;
    load_tile(tile);
    for (; tile < tiles; tile += 2 * gridDim.x, it += 2) {
      DBG(0);
      mbar_wait(&sh->x_free[set], ((it >> 1) & 1) ^ 1);                 // MMA 1 of my previous tile is done with the buffer
      DBG(1);
      {
        // row r = lane / 16 + 2 j, elements 4 (lane % 16) ..+3.  Squared norms: the 16 per-lane partial sums go through
        // one transposing butterfly over the row's 16 lanes (15 shuffles; lane cc ends up with the total of row j = cc),
        // one square root / reciprocal per lane, and the 16 reciprocals come back with one shuffle each
        const int cc = lane & 15;
        float* nrm_out = sh->nrm[set][(it >> 1) & 1] + e * kBlk;
        float rinv_mine;
        {
          float t8[8], t4[4], t2[2];
          const bool b8 = cc & 8, b4 = cc & 4, b2 = cc & 2, b1 = cc & 1;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float lo = fmaf(v[t].x, v[t].x, fmaf(v[t].y, v[t].y, fmaf(v[t].z, v[t].z, v[t].w * v[t].w)));
            const float hi = fmaf(v[t + 8].x, v[t + 8].x, fmaf(v[t + 8].y, v[t + 8].y, fmaf(v[t + 8].z, v[t + 8].z, v[t + 8].w * v[t + 8].w)));
            t8[t] = (b8 ? hi : lo) + __shfl_xor_sync(kFull, b8 ? lo : hi, 8);
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) t4[t] = (b4 ? t8[t + 4] : t8[t]) + __shfl_xor_sync(kFull, b4 ? t8[t] : t8[t + 4], 4);
#pragma unroll
          for (int t = 0; t < 2; ++t) t2[t] = (b2 ? t4[t + 2] : t4[t]) + __shfl_xor_sync(kFull, b2 ? t4[t] : t4[t + 2], 2);
          const float ss = (b1 ? t2[1] : t2[0]) + __shfl_xor_sync(kFull, b1 ? t2[0] : t2[1], 1);
          const float nrm = sqrtf(ss);
          rinv_mine = 1.f / fmaxf(nrm, kNormEps);
          const int r = (lane >> 4) + 2 * cc;
          if (r < N) nrm_out[r] = nrm;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int r = (lane >> 4) + 2 * j;
          const float rinv = __shfl_sync(kFull, rinv_mine, (lane & 16) | j);
          if (r < N) {
            uint32_t h0[2], h1[2];
            split_h4(make_float4(v[j].x * rinv, v[j].y * rinv, v[j].z * rinv, v[j].w * rinv), h0, h1);
            const uint32_t off = xb + sw128(e * kBlk + r, cc >> 1) + (cc & 1) * 8;
            sts_v2(off, h0[0], h0[1]);
            sts_v2(off + kTile, h1[0], h1[1]);
          }
        }
      }
      fence_async_proxy();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->x_full[set]);
      DBG(2);
      if (!kBwd) load_tile(tile + 2 * gridDim.x);                       // the set's next tile flies during the other set's turn
      if (kBwd) {
        if (prev_tile >= 0) e2_tile(prev_tile, prev_it);                // its MMA 2 completed long ago
        DBG(5);
        if (it > 0) mbar_wait(&sh->xt_free[(it - 1) & 1], ((it - 1) >> 1) & 1);   // MMA 2 of the previous tile is done with XT
        DBG(6);
        // transposed copy of my block: XT[64 (e / 2) + d][32 (e % 2) + j] = x^_j[d].  ldmatrix.trans hands every lane the
        // transposed 8 x 8 fp16 tiles of four row groups at once: lane i gets dim 8 c + i / 4 of rows 8 t + 2 (i % 4), + 1
        // as one 32-bit word per t - four conflict-free 4-byte stores per 8 dims (the first version moved single halves:
        // 16 st.shared.u16 per lane and 8 dims, 4400 instructions per tile)
        {
          const uint32_t src = xb + (uint32_t)(e * kBlk + lane) * 128;
          const uint32_t dst = base + kXT + (uint32_t)(64 * (e >> 1) + (lane >> 2)) * 128 + (lane & 3) * 4;
          uint32_t ct[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) ct[t] = (uint32_t)(((4 * (e & 1) + t) ^ (lane >> 2)) & 7) << 4;
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {
            const uint32_t sa = src + ((uint32_t)((c ^ lane) & 7) << 4);
#pragma unroll
            for (int comp = 0; comp < 2; ++comp) {
              uint32_t r0, r1, r2, r3;
              asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                           : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                           : "r"(sa + comp * kTile));
              const uint32_t da = dst + c * 1024 + comp * kTile;
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(da + ct[0]), "r"(r0) : "memory");
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(da + ct[1]), "r"(r1) : "memory");
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(da + ct[2]), "r"(r2) : "memory");
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(da + ct[3]), "r"(r3) : "memory");
            }
          }
        }
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->xt_full);
        DBG(7);
        prev_tile = tile;
        prev_it = it;
        // the 64 prefetch registers are live only from here to the split of the set's next tile (held across E2 and the
        // transposed copy they spilled, and every spill store waited for its load: 13000 clocks per tile)
        load_tile(tile + 2 * gridDim.x);
        DBG(3);
      }
    }
    if (kBwd && prev_tile >= 0) e2_tile(prev_tile, prev_it);
  } else if (warp == kIssuer) {
    // =========================================================== MMA issuer (warp-uniform control flow, one elected lane)
    constexpr uint32_t kIdesc1 = idesc_f16(128, 128), kIdesc2 = idesc_f16(128, 64);
    const int my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto mma1 = [&](int it) {
      const int b = it & 1;
      mbar_wait(&sh->x_full[b], (it >> 1) & 1);
      mbar_wait(&sh->d1_free[b], ((it >> 1) & 1) ^ 1);
      fence_after();
      DBG(8);
      if (elect_one()) {
        const uint32_t acc = tmem + b * 128;
        const uint32_t h0 = desc_lo(base + kX + b * 2 * kTile), h1 = desc_lo(base + kX + (b * 2 + 1) * kTile);
        uint32_t first = 0;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {                          // small terms first: h1.h0, h0.h1, h0.h0
          const uint32_t da = pass == 0 ? h1 : h0, db = pass == 1 ? h1 : h0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {                                 // K = 16 fp16 = 32 bytes per instruction
            mma_f16_lo(acc, da + 2 * k, db + 2 * k, kIdesc1, first);
            first = 1;
          }
        }
        commit(&sh->x_free[b]);
        commit(&sh->d1_full[b]);
      }
      __syncwarp();
    };
    if (my_tiles > 0) mma1(0);
    for (int it = 0; it < my_tiles; ++it) {
      if (it + 1 < my_tiles) mma1(it + 1);
      if (kBwd) {
        mbar_wait(&sh->xt_full, it & 1);
        mbar_wait(&sh->gp_full, it & 1);
        if (it > 0) mbar_wait(&sh->d2_free[(it - 1) & 1], ((it - 1) >> 1) & 1);
        fence_after();
        DBG(10);
        if (elect_one()) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {                                 // K half h: episodes h (rows 0-63) and 2 + h (64-127)
            const uint32_t acc = tmem + 256 + h * 64;
            uint32_t first = 0;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              const uint32_t da = desc_lo(base + kXT + (pass == 0 ? kTile : 0)) + 4 * h;
              const uint32_t db = desc_lo(base + kGP + (pass == 1 ? kTile / 2 : 0)) + 4 * h;
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                mma_f16_lo(acc, da + 2 * k, db + 2 * k, kIdesc2, first);
                first = 1;
              }
            }
          }
          commit(&sh->xt_free[it & 1]);
          commit(&sh->gp_free);
          commit(&sh->d2_full[it & 1]);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================================================== E1: set = (warp - 8) / 4 takes the tiles of its parity
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    const int set = (warp - kFirstE1) >> 2, quad = warp & 3;           // quad = episode slot = TMEM lane quarter
    Scratch* sc = scratch + (warp - kFirstE1);
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != set) continue;
      const int ep = tile * kEp + quad;
      const bool valid = ep < p.E;
      int lab = -1;
      if (valid && lane >= W && lane < N) lab = p.labels[(size_t)ep * Nq + (lane - W)];
      mbar_wait_sleep(&sh->d1_full[set], (it >> 1) & 1);
      fence_after();
      DBG(12);
      // B operand: row 32 (quad / 2) + lane of the 64-row tile, bytes 64 (quad % 2) ..+63; h1 tile 8 KB further
      const int grow = 32 * (quad >> 1) + lane;
      if (valid) {
        e1_episode<kBwd>(p, ep, lab, tmem + ((uint32_t)(quad * 32) << 16) + set * 128 + quad * 32, sc,
                         sh->nrm[set][(it >> 1) & 1] + quad * kBlk, &sh->d1_free[set], &sh->gp_free, (it & 1) ^ 1,
                         base + kGP + grow * 128, kTile / 2, (4 * (quad & 1)) ^ (grow & 7), sh->rs[it & 1] + quad * kBlk + lane,
                         (p.dbg && blockIdx.x == 0 && quad == 0 && it < 24) ? p.dbg + it * 32 : nullptr);
      } else {
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->d1_free[set]);
        if (kBwd) mbar_wait_sleep(&sh->gp_free, (it & 1) ^ 1);
      }
      if (kBwd) {
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->gp_full);
      }
      DBG(20);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == kIssuer) tmem_free<kBwd ? 512 : 256>(tmem);
}

}  // namespace

// Anchors branch, Dp = 64, W <= 8, W + Nq <= 31.  AFSL_ANGULAR_TC=0 leaves the launch to the fp32-pipe kernels (the parity
// tests run all of them on the same cases).
int launch_angular_tc(const AngParams& p, bool bwd, cudaStream_t stream, const char* name, bool* handled) {
  *handled = false;
  if (!p.anchors || p.D != kD || p.W > kMaxW || p.W + p.Nq > kOnesRow) return AFSL_OK;
  const char* env = getenv("AFSL_ANGULAR_TC");
  if (env && atoi(env) == 0) return AFSL_OK;
  *handled = true;
  const size_t bytes = (bwd ? kOpBytesBwd : kOpBytesFwd) + kE1Warps * sizeof(Scratch) + sizeof(Shared) + 1024;
  auto fn = bwd ? angular_tc_kernel<true> : angular_tc_kernel<false>;
  if (int rc = opt_in_smem(fn, bytes, name)) return rc;
  int sms = kNumSMs, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = (p.E + kEp - 1) / kEp;
  if (getenv("AFSL_ANGULAR_DBG")) {
    // timeline of CTA 0 (tools/angular_bench.py --timeline): SM clocks of every pipeline event of its first 24 tiles
    AngParams q = p;
    long long host[24 * 32];
    if (cudaMalloc(&q.dbg, sizeof(host)) != cudaSuccess) return AFSL_ECUDA;
    cudaMemsetAsync(q.dbg, 0, sizeof(host), stream);
    fn<<<tiles < sms ? tiles : sms, kThreads, bytes, stream>>>(q);
    cudaStreamSynchronize(stream);
    cudaMemcpy(host, q.dbg, sizeof(host), cudaMemcpyDeviceToHost);
    cudaFree(q.dbg);
    long long t0 = 0;
    for (long long t : host) if (t && (!t0 || t < t0)) t0 = t;
    fprintf(stderr, "angular_tc %s timeline (clocks since the first event; columns = events 0..20)\n", bwd ? "bwd" : "fwd");
    for (int it = 0; it < 24; ++it) {
      fprintf(stderr, "it %2d:", it);
      for (int ev = 0; ev <= 20; ++ev) fprintf(stderr, " %6lld", host[it * 32 + ev] ? host[it * 32 + ev] - t0 : -1);
      fprintf(stderr, "\n");
    }
    AFSL_CHECK_LAUNCH(name);
    return AFSL_OK;
  }
  fn<<<tiles < sms ? tiles : sms, kThreads, bytes, stream>>>(p);
  AFSL_CHECK_LAUNCH(name);
  return AFSL_OK;
}

}  // namespace afsl
