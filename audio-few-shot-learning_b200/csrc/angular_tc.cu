// Angular loss, prototypes-as-anchors branch (AngularLossClass.forward, loops/loss.py:68-83 + the pytorch_metric_learning
// miner / loss restated in oracle/angular.py), Dp = 64 and W + Nq <= 31 pooled rows, on the tensor cores.
//
// What the loss needs per episode is the Gram matrix of its 30 pooled rows (forward) and, in the backward, the product
// dL/dGram . X: three 30 x 30 x 64 contractions that kept the one-warp-per-episode kernel (angular_warp.cu) at 0.07-0.11 of
// the HBM roofline on the fp32 pipe.  Here FOUR episodes share one 128-row tcgen05 tile, split TF32 (x = hi + lo, three
// passes lo.hi + hi.lo + hi.hi with fp32 accumulation in TMEM, ~2^-22 per product):
//
//   rows 32 e + i of the tile: prototypes (i < W), queries (W <= i < N), zero rows, and row 31 = ones, so that column 31 of
//   the Gram block is the component sum the miner's pairwise_distance epsilon needs;
//   MMA 1   D1[128 x 128] = X X^T (raw rows; the 32 x 32 diagonal blocks are the episodes' Gram matrices, norms = sqrt of
//           the diagonal);
//   E1      one warp per episode, lane = row: tcgen05.ld of its 32 Gram entries, normalisation, mining counts from ballots,
//           the weighted log-sum-exp of its (prototype, query) pair with the negatives unrolled over registers, loss; in the
//           backward dL/dGram rows (transposed through the warp's scratch, per-class sums in ascending row order:
//           deterministic), with the row normalisation's backward folded in:
//           G3[i][j] = r_i r_j G'[i][j] + delta_ij r_i (drho_i - r_i <x^_i, dx^_i>),  so that  dX = G3 X  exactly;
//   MMA 2   (dX)^T of two episodes per instruction: D2[2 x 64 dims][2 x 32 rows] = X^T-tiles . G3-rows^T, every operand
//           K-major (the producers keep a transposed copy of the tile: tools/micro/umma_layout_probe.cu);
//   E2      lane = embedding dimension, columns = rows: coalesced 128-byte stores of dP / dQ.
//
// One persistent CTA per SM, 17 warps, mbarriers only: 2 x 4 producer / E2 warps (warp = episode slot, the two sets take
// alternate tiles: 16 LDG.128 per lane straight from HBM, issued TWO tiles ahead and held in registers - two tiles = 61 KB
// in flight per SM; hi / lo split; transposed copy; E2 of the set's previous tile), 2 x 4 E1 warps (alternate tiles, two D1
// accumulators), 1 issuer warp.  Every role is a compact LOOP over 4-column
// chunks (tcgen05.ld.x4 re-reads the accumulator instead of holding 32-wide register arrays): the first version, fully
// unrolled over 32 columns, was 8000 SASS instructions executed once per tile and bound by instruction fetch (no_inst stalls,
// profiles/r2q_*).
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
constexpr int kTile = 128 * 128;              // bytes of a [128 x 32 fp32] swizzled tile
constexpr int kLs = 36;                       // row stride (floats) of the per-warp scratch matrices: conflict-free
constexpr float kNormEps = 1e-12f;            // F.normalize
constexpr float kPairEps = 1e-6f;             // F.pairwise_distance
constexpr unsigned kFull = 0xffffffffu;

constexpr int kPeWarps = 8, kE1Warps = 8;                      // two sets of four each, alternating tiles
constexpr int kFirstE1 = kPeWarps, kIssuer = kFirstE1 + kE1Warps;
constexpr int kThreads = (kIssuer + 1) * 32;                    // 544

// per-E1-warp scratch (floats)
struct Scratch {
  float fx[kBlk * kLs];        // per-lane rows: exponents -> exponentials -> T (dL/dGram of the pairs, row = positive)
  float sp[kMaxW * kLs];       // raw Gram rows of the prototypes, later the per-class column sums of T
  float rinv[kBlk], nrm[kBlk], csn[kBlk], gdi[kBlk], nu[kBlk], gs[kBlk];
  int manc[kMaxW];
};

struct Bars {
  // mbarrier waits only tell adjacent phases apart, so a barrier that a role observes every OTHER tile (the two producer
  // sets, the two E1 sets) exists once per tile parity: index = tile iteration & 1, phase = iteration >> 1
  uint64_t x_full, xt_full, gp_full, gp_free;
  uint64_t x_free[2], xt_free[2], xt_done[2], d2_full[2], d2_free[2], d1_full[2], d1_free[2];
  uint32_t tmem_base;
};

constexpr uint32_t kXH = 0, kXL = 2 * kTile, kXTH = 4 * kTile, kXTL = 6 * kTile, kGPH = 8 * kTile, kGPL = 9 * kTile;
constexpr uint32_t kOpBytesBwd = 10 * kTile, kOpBytesFwd = 4 * kTile;

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
__device__ __forceinline__ void sts1(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
// four consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ float4 tmem_ld4(uint32_t taddr) {
  uint32_t a, b, c, d;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c), __uint_as_float(d));
}
__device__ __forceinline__ float pick(const float4& t, int u) { return u == 0 ? t.x : u == 1 ? t.y : u == 2 ? t.z : t.w; }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// timeline of CTA 0 (AFSL_ANGULAR_DBG=1): SM clock at event `ev` of tile iteration `it`, one warp per role
#define DBG(ev)                                                                                              \
  do {                                                                                                       \
    if (p.dbg && blockIdx.x == 0 && (threadIdx.x & 127) == 0 && it < 24) p.dbg[it * 32 + (ev)] = clock64(); \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------- E1
// One warp, lane = row of the episode's block.  `tm` = TMEM address of the block's 32 Gram columns in this warp's lanes.
template <bool kBwd>
__device__ __forceinline__ void e1_episode(const AngParams& p, int ep, int lab_in, uint32_t tm, Scratch* sc, uint64_t* d1_free,
                                           uint64_t* gp_free, uint32_t gp_parity, uint32_t gph_row, uint32_t gpl_row,
                                           long long* dbg) {
#define DBGE(ev)                                         \
  do {                                                   \
    if (dbg && (threadIdx.x & 31) == 0) dbg[ev] = clock64(); \
  } while (0)
  const int lane = threadIdx.x & 31;
  const int W = p.W, N = p.W + p.Nq;
  const bool qv = lane >= W && lane < N && lab_in >= 0 && lab_in < W;
  const int lab = lane < W ? lane : (qv ? lab_in : -1);
  const int a = qv ? lab : 0;
  const float c4 = 4.f * p.t2, cneg = -2.f * (1.f + p.t2);
  float* fx_row = sc->fx + lane * kLs;

  // ---- raw Gram row: diagonal (norm), column 31 (component sum); the prototypes' rows go to the scratch for every lane
  float gii = 0.f, srow = 0.f;
#pragma unroll 2
  for (int c = 0; c < 8; ++c) {
    const float4 t = tmem_ld4(tm + 4 * c);
    if (c == (lane >> 2)) gii = pick(t, lane & 3);
    if (c == 7) srow = t.w;
    if (lane < W) *reinterpret_cast<float4*>(sc->sp + lane * kLs + 4 * c) = t;
  }
  const float nrm = sqrtf(fmaxf(gii, 0.f));
  const float ri = 1.f / fmaxf(nrm, kNormEps);
  const float cs = srow * ri;                                   // component sum of the normalised row
  const float gdi = gii * ri * ri;
  sc->rinv[lane] = ri; sc->nrm[lane] = nrm; sc->csn[lane] = cs; sc->gdi[lane] = gdi;
  if (lane < kMaxW) sc->manc[lane] = 0;
  __syncwarp();
  DBGE(13);
  const float ra_inv = sc->rinv[a];
  const float* pa_row = sc->sp + a * kLs;                       // raw Gram row of my prototype
  const float gaq = pa_row[lane] * ra_inv * ri;                 // cos(prototype a, this query)

  // ---- masks: queries of the episode, lanes of my class, my pair's negatives
  const unsigned allq = __ballot_sync(kFull, qv);
  const unsigned sameq = __match_any_sync(kFull, lab) & allq;
  const unsigned negmask = qv ? (allq & ~sameq) : 0u;
  unsigned cm[kMaxW];
#pragma unroll
  for (int w = 0; w < kMaxW; ++w) cm[w] = __ballot_sync(kFull, qv && lab == w);

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
      const float4 t = tmem_ld4(tm + 4 * c), r4 = ld4(sc->rinv + 4 * c), s4 = ld4(pa_row + 4 * c);
      const float4 gd4 = ld4(sc->gdi + 4 * c), cs4 = ld4(sc->csn + 4 * c);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = 4 * c + u;
        const float gk = pick(t, u) * ri * pick(r4, u), pk = pick(s4, u) * ra_inv * pick(r4, u);
        const float dot = (ra * pk + rq * gk) * inv;
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
  float mx = 0.f;
#pragma unroll 2
  for (int c = 0; c < 8; ++c) {
    const float4 t = tmem_ld4(tm + 4 * c), r4 = ld4(sc->rinv + 4 * c), s4 = ld4(pa_row + 4 * c), nu4 = ld4(sc->nu + 4 * c);
    float4 rho = make_float4(1.f, 1.f, 1.f, 1.f);
    if (!p.normalize_ref) rho = ld4(sc->nrm + 4 * c);
    float f[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float rk = pick(r4, u);
      f[u] = fmaf(c4 * pick(rho, u), (pick(s4, u) * ra_inv + pick(t, u) * ri) * rk, base);
      const bool use = ((negmask >> (4 * c + u)) & 1u) && pick(nu4, u) > 0.f;
      mx = use ? fmaxf(mx, f[u]) : mx;
    }
    *reinterpret_cast<float4*>(fx_row + 4 * c) = make_float4(f[0], f[1], f[2], f[3]);
  }
  if (!kBwd) {                                                  // the accumulator is not needed any more
    fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(d1_free);
  }
  DBGE(15);
  float tot = __expf(-mx);
#pragma unroll 2
  for (int c = 0; c < 8; ++c) {
    const float4 f4 = ld4(fx_row + 4 * c), nu4 = ld4(sc->nu + 4 * c);
    float ex[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float w = ((negmask >> (4 * c + u)) & 1u) ? pick(nu4, u) : 0.f;
      ex[u] = w * __expf(pick(f4, u) - mx);
      tot += ex[u];
    }
    if (kBwd) *reinterpret_cast<float4*>(fx_row + 4 * c) = make_float4(ex[0], ex[1], ex[2], ex[3]);
  }
  const float term = omega > 0.f ? omega * (mx + logf(tot)) : 0.f;
  const float num = warp_sum(term), den = warp_sum(omega);
  DBGE(16);
  if (!kBwd) {
    if (lane == 0) p.loss[ep] = den > 0.f ? num / den : 0.f;
    return;
  }

  // ---- backward.  T_i[k] = 4 t2 rho_k g_k: dL/dGram[i][k] of my pair (and of its anchor's row), gsum = sum_k g_k
  const float scale = den > 0.f ? p.d_loss[ep] / den : 0.f;
  const float coef = omega > 0.f ? scale * omega / tot : 0.f;
  float gsum = 0.f;
#pragma unroll 2
  for (int c = 0; c < 8; ++c) {
    const float4 e4 = ld4(fx_row + 4 * c);
    float4 rho = make_float4(1.f, 1.f, 1.f, 1.f);
    if (!p.normalize_ref) rho = ld4(sc->nrm + 4 * c);
    float tk[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float gk = coef * pick(e4, u);
      gsum += gk;
      tk[u] = c4 * pick(rho, u) * gk;
    }
    *reinterpret_cast<float4*>(fx_row + 4 * c) = make_float4(tk[0], tk[1], tk[2], tk[3]);
  }
  sc->gs[lane] = gsum;
  __syncwarp();                                                 // T rows complete; every lane is done with the prototype rows
  DBGE(17);
  // per-class column sums (rows of a class in ascending order): pr[w] = dL/dGram[prototype w][this lane's row]
  {
    float pr[kMaxW];
#pragma unroll
    for (int w = 0; w < kMaxW; ++w) pr[w] = 0.f;
    unsigned any = 0;
#pragma unroll
    for (int w = 0; w < kMaxW; ++w) any |= cm[w];
    while (any) {                                               // one row of every class per round: W independent chains
      any = 0;
#pragma unroll
      for (int w = 0; w < kMaxW; ++w)
        if (cm[w]) {
          pr[w] += sc->fx[(__ffs(cm[w]) - 1) * kLs + lane];
          cm[w] &= cm[w] - 1;
          any |= cm[w];
        }
    }
#pragma unroll
    for (int w = 0; w < kMaxW; ++w)
      if (w < W) sc->sp[w * kLs + lane] = pr[w];
  }
  __syncwarp();
  // symmetric G'[i][k] = dL/dGram[i][k] + dL/dGram[k][i] of row i = this lane, scaled into the B operand of MMA 2:
  // G3[i][k] = r_i r_k G'[i][k] (hi = the raw value, the tensor core truncates it; lo = the rounded remainder)
  const bool isp = lane < W;
  const float* sp_row = sc->sp + (isp ? lane : 0) * kLs;
  float drho = 0.f, dot = 0.f;
  DBGE(18);
  mbar_wait_sleep(gp_free, gp_parity);                          // MMA 2 of the previous tile is done with the operand
  DBGE(19);
#pragma unroll 2
  for (int c = 0; c < 8; ++c) {
    const float4 t = tmem_ld4(tm + 4 * c), r4 = ld4(sc->rinv + 4 * c), t4 = ld4(fx_row + 4 * c), sp4 = ld4(sp_row + 4 * c);
    const float4 gs4 = ld4(sc->gs + 4 * c);
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = 4 * c + u;
      const float gk = pick(t, u) * ri * pick(r4, u);           // cos(row i, row k)
      const float col = sc->fx[k * kLs + lane];                 // T_k[i]
      drho = fmaf(col, gk, drho);
      const float vq = pick(t4, u) + col;
      const float vp = ((sameq >> k) & 1u) ? cneg * pick(gs4, u) : pick(sp4, u);
      float val = qv ? vq : (isp ? vp : 0.f);
      if (c < 2 && k < W) {                                     // prototype columns of a query's row
        const float prk = sc->sp[k * kLs + lane];               // 0 for my own class: it never uses me as a negative
        if (qv) val = (k == lab) ? cneg * gsum : prk;
        drho = fmaf(prk, gk, drho);
      }
      dot = fmaf(val, gk, dot);
      o[u] = ri * pick(r4, u) * val;
    }
    const float4 v = make_float4(o[0], o[1], o[2], o[3]);
    const uint32_t off = ((uint32_t)((c ^ lane) & 7) << 4);     // sw128 inside this lane's row
    sts4(gph_row + off, v);
    sts4(gpl_row + off, lo_of_raw(v));
  }
  fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(d1_free);
  // the diagonal carries the backward of the row normalisation: r_i (drho_i - r_i <x^_i, dx^_i>); G'[i][i] itself is 0
  drho = (!p.normalize_ref && qv && nrm > 0.f) ? drho / nrm : 0.f;
  if (!(nrm > kNormEps)) dot = 0.f;                             // F.normalize clamps the norm: no projection term there
  const float cdiag = ri * (drho - ri * dot);
  const uint32_t doff = ((uint32_t)(((lane >> 2) ^ lane) & 7) << 4) + (lane & 3) * 4;
  sts1(gph_row + doff, cdiag);
  sts1(gpl_row + doff, rna_tf32(cdiag - trunc_tf32(cdiag)));
}

// ------------------------------------------------------------------------------------------------------------ kernel
template <bool kBwd>
__global__ void __launch_bounds__(kThreads, 1) angular_tc_kernel(const AngParams p) {
  extern __shared__ __align__(1024) uint8_t smem_ang_raw[];
  uint8_t* smem = smem_ang_raw + ((1024u - (smem_u32(smem_ang_raw) & 1023u)) & 1023u);
  const uint32_t base = smem_u32(smem);
  constexpr uint32_t kOps = kBwd ? kOpBytesBwd : kOpBytesFwd;
  Scratch* scratch = reinterpret_cast<Scratch*>(smem + kOps);
  Bars* bars = reinterpret_cast<Bars*>(scratch + kE1Warps);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = p.W, Nq = p.Nq, N = W + Nq;
  const int tiles = (p.E + kEp - 1) / kEp;

  if (tid == 0) {
    mbar_init(&bars->x_full, 4);
    mbar_init(&bars->xt_full, 4);
    mbar_init(&bars->gp_full, 4);
    mbar_init(&bars->gp_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->x_free[b], 1);
      mbar_init(&bars->xt_free[b], 1);
      mbar_init(&bars->xt_done[b], 4);
      mbar_init(&bars->d2_full[b], 1);
      mbar_init(&bars->d2_free[b], 4);
      mbar_init(&bars->d1_full[b], 1);
      mbar_init(&bars->d1_free[b], 4);
    }
    fence_barrier_init();
  }
  // operand tiles start as zeros (pad rows stay zero for the whole launch), row 31 of every block of X is ones
  for (uint32_t i = tid; i < kOps / 16; i += blockDim.x) sts4(base + i * 16, make_float4(0.f, 0.f, 0.f, 0.f));
  __syncthreads();
  for (int i = tid; i < kEp * 2 * 8; i += blockDim.x) {
    const int e = i >> 4, kb = (i >> 3) & 1, c = i & 7;
    sts4(base + kXH + kb * kTile + sw128(e * kBlk + kOnesRow, c), make_float4(1.f, 1.f, 1.f, 1.f));
  }
  if (warp == kIssuer) tmem_alloc<512>(&bars->tmem_base);
  fence_async_proxy();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < kPeWarps) {
    // =========================================================== producers + E2: warp & 3 = episode slot, warp / 4 = set
    const int set = warp >> 2, e = warp & 3;
    const int n16 = N * 16, w16 = W * 16;                               // 16-byte chunks of the block / of its prototypes
    float4 v[16];
    auto load_tile = [&](int tile) {
      const int ep = tile * kEp + e;
      const bool valid = tile < tiles && ep < p.E;
      const float4* p4 = reinterpret_cast<const float4*>(p.protos) + (size_t)(valid ? ep : 0) * w16;
      const float4* q4 = reinterpret_cast<const float4*>(p.queries) + (size_t)(valid ? ep : 0) * (n16 - w16);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int ci = lane + 32 * j;                                   // row ci / 16, 16-byte chunk ci % 16 of the block
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && ci < n16) v[j] = ldg_stream(ci < w16 ? p4 + ci : q4 + (ci - w16));
      }
    };
    auto e2_tile = [&](int tile, int it) {                              // lane = embedding dimension, columns = rows
      const int d = (e & 1) * 32 + lane;
      mbar_wait(&bars->d2_full[it & 1], (it >> 1) & 1);
      fence_after();
#pragma unroll 1
      for (int s2 = 0; s2 < 2; ++s2) {
        const int ep = tile * kEp + 2 * s2 + (e >> 1);
        const bool valid = ep < p.E;
        const uint32_t ta = tmem + ((uint32_t)(e * 32) << 16) + 256 + s2 * 64 + (e >> 1) * 32;
        const size_t dp = (size_t)(valid ? ep : 0) * W * kD + d, dq = (size_t)(valid ? ep : 0) * Nq * kD + d;
#pragma unroll 2
        for (int c = 0; c < 8; ++c) {
          const float4 t = tmem_ld4(ta + 4 * c);
          if (valid) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int i = 4 * c + u;
              if (i < W) p.d_protos[dp + i * kD] = pick(t, u);
              else if (i < N) p.d_queries[dq + (i - W) * kD] = pick(t, u);
            }
          }
        }
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->d2_free[it & 1]);
    };
    int it = set, prev_tile = -1, prev_it = -1;
    int tile = blockIdx.x + set * gridDim.x;
    load_tile(tile);
    for (; tile < tiles; tile += 2 * gridDim.x, it += 2) {
      DBG(0);
      if (it > 0) mbar_wait(&bars->x_free[(it - 1) & 1], ((it - 1) >> 1) & 1);   // MMA 1 of the previous tile is done with X
      // ... and so is the other set's transposed copy, which reads X
      if (kBwd && it > 0) mbar_wait(&bars->xt_done[(it - 1) & 1], ((it - 1) >> 1) & 1);
      DBG(1);
      {
        const int cc = lane & 15;
        const uint32_t half = base + (cc >> 3) * kTile;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int r = (lane >> 4) + 2 * j;
          if (r < N) {
            const uint32_t off = half + sw128(e * kBlk + r, cc & 7);
            sts4(off + kXH, v[j]);
            sts4(off + kXL, lo_of_raw(v[j]));
          }
        }
      }
      fence_async_proxy();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->x_full);
      DBG(2);
      load_tile(tile + 2 * gridDim.x);                                  // the set's next tile flies while this one is finished
      DBG(3);
      if (kBwd) {
        if (prev_tile >= 0) e2_tile(prev_tile, prev_it);                // its MMA 2 completed long ago
        DBG(5);
        // transposed copy of my block: XT[e][d][j] = X[j][d] (lane = row j reads its own row, conflict-free both ways)
        if (it > 0) mbar_wait(&bars->xt_free[(it - 1) & 1], ((it - 1) >> 1) & 1);   // MMA 2 of the previous tile is done with XT
        DBG(6);
#pragma unroll 1
        for (int hc = 0; hc < 16; ++hc) {
          const int h = hc >> 3, c = hc & 7;
          const float4 hv = lds4(base + kXH + h * kTile + sw128(e * kBlk + lane, c));
          const float4 lv = lo_of_raw(hv);
          const uint32_t col = (uint32_t)e * (kD * 128) + (lane & 3) * 4;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int d = 4 * hc + u;
            const uint32_t off = col + sw128(d, lane >> 2);
            sts1(base + kXTH + off, pick(hv, u));
            sts1(base + kXTL + off, pick(lv, u));
          }
        }
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars->xt_full);
          mbar_arrive(&bars->xt_done[it & 1]);
        }
        DBG(7);
        prev_tile = tile;
        prev_it = it;
      }
    }
    if (kBwd && prev_tile >= 0) e2_tile(prev_tile, prev_it);
  } else if (warp == kIssuer) {
    // =========================================================== MMA issuer (warp-uniform control flow, one elected lane)
    constexpr uint32_t kIdesc1 = idesc_tf32(128, 128), kIdesc2 = idesc_tf32(128, 64);
    const int my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto mma1 = [&](int it) {
      const int b = it & 1;
      mbar_wait(&bars->x_full, it & 1);
      mbar_wait(&bars->d1_free[b], ((it >> 1) & 1) ^ 1);
      fence_after();
      DBG(8);
      if (elect_one()) {
        const uint32_t acc = tmem + b * 128;
        uint32_t first = 0;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {                          // small terms first: lo.hi, hi.lo, hi.hi
          const uint32_t pa_ = pass == 0 ? kXL : kXH, pb_ = pass == 1 ? kXL : kXH;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint32_t da = desc_lo(base + pa_ + kb * kTile), db = desc_lo(base + pb_ + kb * kTile);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              mma_tf32_lo(acc, da + 2 * k, db + 2 * k, kIdesc1, first);
              first = 1;
            }
          }
        }
        commit(&bars->x_free[b]);
        commit(&bars->d1_full[b]);
      }
      __syncwarp();
    };
    if (my_tiles > 0) mma1(0);
    for (int it = 0; it < my_tiles; ++it) {
      if (it + 1 < my_tiles) mma1(it + 1);
      if (kBwd) {
        mbar_wait(&bars->xt_full, it & 1);
        mbar_wait(&bars->gp_full, it & 1);
        if (it > 0) mbar_wait(&bars->d2_free[(it - 1) & 1], ((it - 1) >> 1) & 1);
        fence_after();
        DBG(10);
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {                                 // two episodes per instruction
            const uint32_t acc = tmem + 256 + s * 64;
            uint32_t first = 0;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              const uint32_t pa_ = pass == 0 ? kXTL : kXTH, pb_ = pass == 1 ? kGPL : kGPH;
              const uint32_t da = desc_lo(base + pa_ + s * kTile), db = desc_lo(base + pb_ + s * (kTile / 2));
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                mma_tf32_lo(acc, da + 2 * k, db + 2 * k, kIdesc2, first);
                first = 1;
              }
            }
          }
          commit(&bars->xt_free[it & 1]);
          commit(&bars->gp_free);
          commit(&bars->d2_full[it & 1]);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================================================== E1: set = (warp - 4) / 4 takes the tiles of its parity
    const int set = (warp - kFirstE1) >> 2, quad = warp & 3;           // quad = episode slot = TMEM lane quarter
    Scratch* sc = scratch + (warp - kFirstE1);
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != set) continue;
      const int ep = tile * kEp + quad;
      const bool valid = ep < p.E;
      int lab = -1;
      if (valid && lane >= W && lane < N) lab = p.labels[(size_t)ep * Nq + (lane - W)];
      mbar_wait_sleep(&bars->d1_full[set], (it >> 1) & 1);
      fence_after();
      DBG(12);
      const uint32_t row = (uint32_t)(quad * kBlk + lane) * 128u;
      if (valid) {
        e1_episode<kBwd>(p, ep, lab, tmem + ((uint32_t)(quad * 32) << 16) + set * 128 + quad * 32, sc, &bars->d1_free[set],
                         &bars->gp_free, (it & 1) ^ 1, base + kGPH + row, base + kGPL + row,
                         (p.dbg && blockIdx.x == 0 && quad == 0 && it < 24) ? p.dbg + it * 32 : nullptr);
      } else {
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->d1_free[set]);
        if (kBwd) mbar_wait_sleep(&bars->gp_free, (it & 1) ^ 1);
      }
      if (kBwd) {
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->gp_full);
      }
      DBG(20);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == kIssuer) tmem_free<512>(tmem);
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
  const size_t bytes = (bwd ? kOpBytesBwd : kOpBytesFwd) + kE1Warps * sizeof(Scratch) + sizeof(Bars) + 1024;
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
