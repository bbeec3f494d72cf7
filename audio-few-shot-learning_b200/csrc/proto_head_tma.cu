// Prototype head forward for MANY-WAY tasks: TMA-fed, tcgen05 tensor cores, TMEM accumulators (sm_100a).
//
// Same arithmetic and reference as proto_head_wide.cu / proto_head_mma.cu (models/util_functions.py:6-19: per-class mean of
// the support rows in ascending row order; few_shot_classifier.py:108-116: -torch.cdist, in the matmul form
// |q|^2 + |p|^2 - 2 q.p clamped at 0 that cdist itself uses beyond 25 rows; loops/loss.py:24-37: log-softmax + NLL;
// loops/loops.py:79: first-index argmax, #correct).  This is the shape of the path that is a real dense contraction
// (Nq = 100 rows x W = 20 prototypes x D = 64..256 per task, thousands of tasks: SURVEY 8d config 5), so q.p runs on
// `tcgen05.mma.kind::tf32` in split precision: the tensor core reads the RAW fp32 rows TMA delivered (it truncates them to
// TF32 itself: hi), the CUDA cores only derive lo = rna_tf32(x - trunc_tf32(x)), and
//     q.p = lo.lo + lo.hi + hi.lo + hi.hi          (four passes, fp32 accumulation in TMEM, ~22 bits per product).
//
// One persistent CTA per SM; every byte comes in through ONE ring of shared-memory stages filled by TMA
// (cp.async.bulk.tensor.2d, 128-byte swizzle = the UMMA operand layout), ~80-100 KB ahead of the consumers and across task
// boundaries - Little's law at ~1.2 us loaded latency needs ~50 KB in flight per SM, and no load latency is ever exposed.
// A stage holds kPair k-blocks (32 columns each, one TMA box per k-block) of either the task's support rows or its query
// rows; kPair = 2 halves the barrier round trips, proxy fences and reductions per task.  Roles (mbarriers only, no CTA
// barrier in the task loop):
//   warp 9     loader    one lane: arms the stage's mbarrier with the byte count and issues the TMA box
//   warps 0-7  producers two groups of four warps take alternate stages.  Support stage: per-class sums straight out of
//                        the swizzled rows (thread = class x 16-byte chunk, rows of a class in ascending order from a
//                        per-task bucket list) -> prototype chunk, |p|^2, the split prototype tiles (B operand);
//                        query stage: lo tile (ring of four) for the raw tile, |q|^2
//   warp 8     issuer    one lane: 16 MMAs (128 x 32 x 8) per query stage, tcgen05.commit hands the stage back to the loader
//   warps 10-13 epilogue tcgen05.ld of the 128 x 32 accumulator (thread = query row): distances, first-index argmax,
//                        log-softmax / NLL, #correct; two accumulators, so task t's epilogue overlaps task t+1's stages
#include <cstdlib>

#include "proto_head.cuh"
#include "tc_common.cuh"

namespace afsl {
namespace {

using namespace tc;

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kProducerWarps = 8;
constexpr int kGroups = 2;                                      // producer groups taking alternate stages
constexpr int kGroupWarps = kProducerWarps / kGroups;
constexpr int kGroupThreads = kGroupWarps * 32;                 // 128
constexpr int kIssuerWarp = 8;
constexpr int kLoaderWarp = 9;
constexpr int kEpiWarp0 = 10;
constexpr int kEpiWarps = 4;
constexpr int kBucketWarp = kEpiWarp0 + kEpiWarps;               // 14
constexpr int kTmaThreads = (kBucketWarp + 1) * 32;             // 480
constexpr int kTileM = 128;                                     // query rows per accumulator (TMEM lanes)
constexpr int kTileN = 32;                                      // prototype slots per accumulator (TMEM columns)
constexpr int kMaxWays = 24;                                    // B tiles hold 24 rows (3 KB); the MMA's slots 24..31 read on into
                                                                // the next tile / the lo ring: garbage columns nobody looks at
constexpr int kBlockK = 32;                                     // fp32 per 128-byte swizzle row = one stage's columns
constexpr int kTile = kTileM * 128;                             // bytes of a full [128 x 32] fp32 tile: what one MMA reads
constexpr int kBTile = kMaxWays * 128;
constexpr int kMaxLoRing = 4;
constexpr int kMaxRing = 9;
constexpr int kMaxSupportRows = kTileM;
constexpr int kAccN = 2 * kMaxWays;                             // accumulator columns: q.p_hi in 0..23, q.p_lo in 24..47
constexpr int kAccStride = 64;                                  // TMEM columns between the two accumulators
constexpr uint32_t kIdesc = idesc_tf32(kTileM, kAccN);

// bucket lists of kBk tasks: the bucket warp runs up to kBk - 1 tasks ahead.  With two sets, the lists of task t could only be
// built once EVERY producer warp was done with those of task t - 2 - and the first warp of a group, which alone has items in
// the second round of a support stage (classes 16-19 of 20), finishes ~2500 clocks after the others: the chain
// slowest warp (t - 2) -> bucket warp (+ an exposed label fetch) -> start of support (t) set the period at D = 64
constexpr int kBk = 4;

struct TmaBars {
  uint64_t tma_full[kMaxRing];    // stage filled by TMA (transaction bytes)
  uint64_t ready[kMaxRing];       // stage processed by its producer group (4 warps)
  uint64_t empty[kMaxRing];       // stage free again (issuer: plain arrive for support stages, tcgen05.commit for query stages)
  uint64_t lo_free[kMaxLoRing];   // lo tiles of a query stage free again (tcgen05.commit)
  uint64_t b_empty[2];            // the task's MMAs are done with the prototype tiles (of task parity b when double-buffered)
  uint64_t acc_full[2], epi_done[2], meta_full[2];
  uint64_t bk_full[kBk], bk_free[kBk];  // the task's bucket lists are built (bucket warp) / read for the last time (producers)
};

struct TmaMeta {
  TmaBars bars;
  uint32_t tmem_base;
  alignas(16) float qq[2][kGroups][kTileM];   // |q|^2 partial sums: the k-blocks each producer group handled
  alignas(16) float pp[2][kGroups][kTileN];   // |p|^2 partial sums, likewise (read 128 bits at a time by the epilogue)
  float part[kEpiWarps];
  int hits[kEpiWarps];
  int cnt[kBk][kTileN];                  // rows per class
};

// waiting roles that are not on the critical path (loader, epilogue) sleep between polls instead of competing with the
// producers for issue slots
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done) __nanosleep(64);
  } while (!done);
}

__device__ __forceinline__ void epilogue_bar() { asm volatile("bar.sync 2, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

// timeline of CTA 0 (AFSL_HEAD_DBG=1, tools/head_many_way_bench.py): SM clock at event `ev` of ring stage `c` (all tasks)
constexpr int kDbgStages = 160, kDbgEvents = 22;
#define HDBG(c, ev)                                                                                        \
  do {                                                                                                     \
    if (p.dbg && blockIdx.x == 0 && (c) < (uint32_t)kDbgStages) p.dbg[(c) * kDbgEvents + (ev)] = clock64(); \
  } while (0)

template <int kD, int kPair, int kLo, bool kOneSup>
__global__ void __launch_bounds__(kTmaThreads, 1)
head_tma_fwd_kernel(const HeadParams p, const __grid_constant__ CUtensorMap map_s, const __grid_constant__ CUtensorMap map_q) {
  constexpr int kKB = kD / kBlockK;
  constexpr int kSt = kKB / kPair;                  // ring stages per phase (support / query) of a task: kPair k-blocks each
  // A k-block tile of a ring stage holds only the rows a task has (its query rows / support rows rounded up to the 8-row
  // swizzle atom: 13 KB instead of 16 KB at 100 rows), so that one more stage fits the ring: the MMA's M = 128 rows read
  // on into the following tile, which only reaches accumulator rows nobody looks at.  tile_b / ring_n come from the host.
  const uint32_t tile_b = (uint32_t)p.tile_rows * 128u;      // bytes of one k-block tile of a stage (a multiple of 1 KB)
  const uint32_t kStageB = kPair * tile_b;                   // bytes of a ring stage (and of a lo-ring slot)
  const uint32_t kRing = (uint32_t)p.ring_stages;
  // support blocks of at most 32 rows (20-way 1-shot, given prototypes): ALL their k-blocks share ONE ring stage, 4 KB apart
  // - the support phase of a task is one barrier round trip instead of kSt (taken when the k-blocks fit a stage of 32-row
  // tiles: D = 64 with two k-blocks per stage)
  constexpr bool kOne = kOneSup && kKB <= kPair;
  constexpr int kSupSt = kOne ? 1 : kSt, kSupPair = kOne ? kKB : kPair;
  const uint32_t kSupStride = kOne ? 4096u : tile_b;
  static_assert(kKB % kPair == 0 && kLo <= kMaxLoRing, "bad stage configuration");
  // the producer group that owns a stage is (global stage index & 1)
  extern __shared__ __align__(1024) uint8_t smem_tma_raw[];
  uint8_t* smem = smem_tma_raw + ((1024u - (smem_u32(smem_tma_raw) & 1023u)) & 1023u);     // 1 KB: swizzle atoms
  const uint32_t ring = smem_u32(smem);                                        // [ring_n][kPair][tile_b] raw rows as TMA wrote them
  // split prototypes [kBBuf][kKB][hi, lo][3 KB]: two sets at D <= 128, so that the support phase of task t + 1 does not
  // wait for the MMAs of task t (at D = 64 that wait serialised the whole task: support 2200 + MMA 1200 clocks of a 3600
  // clock period); at D = 256 a second set (48 KB) would cost two ring stages
  constexpr int kBBuf = kD <= 128 ? 2 : 1;
  constexpr uint32_t kBSet = kKB * 2 * kBTile;
  const uint32_t b_base0 = ring + kRing * kStageB;
  const uint32_t lo_base = b_base0 + kBBuf * kBSet;                            // [kLo][kPair][tile_b] lo tiles of query stages
  // (the last lo tile's MMA reads up to kTile - tile_b bytes past its slot: that much padding before the metadata)
  TmaMeta* meta = reinterpret_cast<TmaMeta*>(smem + kRing * kStageB + kBBuf * kBSet + kLo * kStageB + (kTile - tile_b));
  uint8_t* rows_base = reinterpret_cast<uint8_t*>(meta + 1);                   // [kBk][W][row_stride]: bucketed support rows
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = p.W, Nq = p.Nq;
  const int sup_rows = p.support ? p.Ns : W;                                   // given prototypes: one "support row" per class
  const int row_stride = (sup_rows + 7) & ~7;                                  // a class's row list is read 8 ids at a time

  if (tid == 0) {
    for (int s = 0; s < kMaxRing; ++s) {
      mbar_init(&meta->bars.tma_full[s], 1);
      mbar_init(&meta->bars.ready[s], kGroupWarps);
      mbar_init(&meta->bars.empty[s], 1);
    }
    for (int s = 0; s < kLo; ++s) mbar_init(&meta->bars.lo_free[s], 1);
    mbar_init(&meta->bars.b_empty[0], 1);
    mbar_init(&meta->bars.b_empty[1], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&meta->bars.acc_full[i], 1);
      mbar_init(&meta->bars.epi_done[i], kEpiWarps);
      mbar_init(&meta->bars.meta_full[i], kProducerWarps);
    }
    for (int i = 0; i < kBk; ++i) {
      mbar_init(&meta->bars.bk_full[i], 1);
      mbar_init(&meta->bars.bk_free[i], kProducerWarps);
    }
    fence_barrier_init();
    prefetch_tensormap(&map_s);
    prefetch_tensormap(&map_q);
  }
  // prototype slots W..23 of every B tile stay zero for the whole launch
  for (int i = tid; i < kBBuf * kBSet / 16; i += kTmaThreads) sts4(b_base0 + i * 16, make_float4(0.f, 0.f, 0.f, 0.f));
  // a group that handles no support (query) stage of a task never writes its |p|^2 (|q|^2) partials: they stay zero
  for (int i = tid; i < 2 * kGroups * kTileM; i += kTmaThreads) (&meta->qq[0][0][0])[i] = 0.f;
  for (int i = tid; i < 2 * kGroups * kTileN; i += kTmaThreads) (&meta->pp[0][0][0])[i] = 0.f;
  if (warp == kIssuerWarp) tmem_alloc<2 * kAccStride>(&meta->tmem_base);
  fence_async_proxy();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = meta->tmem_base;

  if (warp == kLoaderWarp) {
    // =============================================================== loader
    if (lane == 0) {
      uint32_t c = 0;
      // The ring holds ~3 stages (shared memory is full) and a stage's slot is busy from the TMA issue to the end of its
      // MMAs; issue -> arrival is 1500-2400 clocks in the timeline.  Optionally the next task's blocks are PREFETCHED INTO
      // L2 one task ahead (<= 200 KB per SM, 30 MB of the 126 MB L2 over the GPU).  Measured: arrival stays at 1000-2200
      // clocks even from L2 (the stream runs at ~0.7 of the HBM rate: queueing, not DRAM latency), so the gain is small.
      const bool l2_prefetch = p.l2_prefetch;
      auto prefetch_task = [&](int e) {
        if (!l2_prefetch || e >= p.E) return;
        // the task's support block and query block are contiguous: two bulk prefetches, so that HBM sees long sequential
        // reads instead of the boxes' 128-byte pieces at a 4 D-byte stride
        const float* sup = p.support ? p.support : p.protos_in;
        bulk_prefetch_l2(sup + (size_t)e * sup_rows * kD, (uint32_t)(sup_rows * kD * 4));
        bulk_prefetch_l2(p.queries + (size_t)e * Nq * kD, (uint32_t)(Nq * kD * 4));
      };
      prefetch_task(blockIdx.x);
      for (int e = blockIdx.x; e < p.E; e += gridDim.x) {
        prefetch_task(e + gridDim.x);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const CUtensorMap* map = half ? &map_q : &map_s;
          const int rows = half ? Nq : sup_rows;
#pragma unroll 1
          const int stages = half ? kSt : kSupSt, per = half ? kPair : kSupPair;
          const uint32_t stride = half ? tile_b : kSupStride;
          for (int st = 0; st < stages; ++st, ++c) {
            const uint32_t s = c % kRing;
            mbar_wait_relaxed(&meta->bars.empty[s], ((c / kRing) & 1) ^ 1);
            HDBG(c, 0);
            mbar_arrive_expect_tx(&meta->bars.tma_full[s], (uint32_t)rows * 128u * per);
            for (int i = 0; i < per; ++i)
              tma_load_2d(ring + s * kStageB + i * stride, map, (st * per + i) * kBlockK, e * rows, &meta->bars.tma_full[s]);
          }
        }
      }
    }
  } else if (warp < kProducerWarps) {
    // =============================================================== producers
    const int grp = warp / kGroupWarps, gtid = tid & (kGroupThreads - 1);
    uint32_t c = 0;                        // stages issued so far (all tasks); this group handles those with (kb & 1) == grp
    uint32_t qc = 0;                       // query stages so far -> lo ring slot
    int it = 0;
    for (int e = blockIdx.x; e < p.E; e += gridDim.x, ++it) {
      const int par = it & 1;
      const int bq = it & (kBk - 1);
      mbar_wait(&meta->bars.bk_full[bq], (it / kBk) & 1);            // the bucket warp built this task's class lists
      const uint8_t* rows = rows_base + (size_t)bq * W * row_stride;
      const uint32_t b_base = b_base0 + (kBBuf == 2 ? par : 0) * kBSet;
      constexpr bool kFixedRoles = ((kSupSt + kSt) & 1) == 0;       // the same group takes the same stages of every task
      bool did_sup = false, did_qry = false, epi_ok = false;
      // ---------------------------------------------------------- support stages: prototypes, |p|^2, split prototype tiles
      {
        // item = (class w, 16-byte chunk j): 8 W items over the group's 128 threads, two rounds
        float pp_acc[2] = {0.f, 0.f};
        int n_it[2];
        float rcp_it[2], fn_it[2];
        const uint8_t* rows_it[2];
#pragma unroll
        for (int rnd = 0; rnd < 2; ++rnd) {
          const int w = (gtid + rnd * kGroupThreads) >> 3;
          n_it[rnd] = w < W ? meta->cnt[bq][w] : 0;
          rows_it[rnd] = rows + (w < W ? w : 0) * row_stride;
          // mean = sum / n through a refined reciprocal and one residual correction (the division's own fast path, without
          // its per-element reciprocal): n == 0 gives 0 * inf = NaN, as the reference's empty mean
          fn_it[rnd] = (float)n_it[rnd];
          const float r0 = __frcp_rn(fn_it[rnd]);
          rcp_it[rnd] = r0;
        }
        bool tiles_free = false;
#pragma unroll 1
        for (int st = 0; st < kSupSt; ++st, ++c) {
          if ((int)(c & 1) != grp) continue;
          did_sup = true;
          const uint32_t s = c % kRing;
          mbar_wait(&meta->bars.tma_full[s], (c / kRing) & 1);
          if (gtid == 0) HDBG(c, 1);
          if (gtid == 32) HDBG(c, 18);
          if (gtid == 96) HDBG(c, 20);
          if (gtid == 0) HDBG(c, 2);
          // A thread's two items (item = gtid and gtid + 128: classes 16.. belong to the first warp only) and the stage's
          // k-blocks two at a time are reduced TOGETHER: two rows per step, one 16-bit read of the row ids per item, up to
          // eight independent 128-bit reads in flight, then the adds in ascending row order.  One item after the other made
          // the first warp of a group - the only one with second items at 20 ways - finish every support stage ~800
          // clocks after the others (profiles/r2u_head_timeline_16.txt), and a stage is ready when its last warp is.
          float sq[2] = {0.f, 0.f};
          const int nmax = n_it[0] > n_it[1] ? n_it[0] : n_it[1];
#pragma unroll
          for (int ib = 0; ib < kSupPair; ib += 2) {
            constexpr int kI = kSupPair >= 2 ? 2 : 1;
            float4 acc[2][kI];
#pragma unroll
            for (int rnd = 0; rnd < 2; ++rnd)
#pragma unroll
              for (int i = 0; i < kI; ++i) acc[rnd][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            const uint32_t stage0 = ring + s * kStageB + ib * kSupStride;
            for (int i0 = 0; i0 < nmax; i0 += 2) {
              float4 v[2][kI][2];
#pragma unroll
              for (int rnd = 0; rnd < 2; ++rnd) {
                const uint32_t ids = *reinterpret_cast<const uint16_t*>(rows_it[rnd] + i0);
                const int j = (gtid + rnd * kGroupThreads) & 7;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const uint32_t off = sw128((int)((ids >> (8 * u)) & 0xffu), j);
#pragma unroll
                  for (int i = 0; i < kI; ++i)
                    if (i0 + u < n_it[rnd]) v[rnd][i][u] = lds4(stage0 + i * kSupStride + off);   // predicated: no traffic for missing rows
                }
              }
#pragma unroll
              for (int rnd = 0; rnd < 2; ++rnd)
#pragma unroll
                for (int u = 0; u < 2; ++u)
                  if (i0 + u < n_it[rnd]) {
#pragma unroll
                    for (int i = 0; i < kI; ++i) {
                      acc[rnd][i].x += v[rnd][i][u].x; acc[rnd][i].y += v[rnd][i][u].y;
                      acc[rnd][i].z += v[rnd][i][u].z; acc[rnd][i].w += v[rnd][i][u].w;
                    }
                  }
            }
            if (!tiles_free) {
              // the MMAs that last read this set of prototype tiles are done (waited for only here, with the stage's
              // rows already summed: at D = 256, one set, this hides most of the previous task's last MMAs)
              if (kBBuf == 2) mbar_wait(&meta->bars.b_empty[par], ((it >> 1) & 1) ^ 1);
              else mbar_wait(&meta->bars.b_empty[0], (it & 1) ^ 1);
              tiles_free = true;
            }
#pragma unroll
            for (int rnd = 0; rnd < 2; ++rnd) {
              const int item = gtid + rnd * kGroupThreads;
              const int w = item >> 3, j = item & 7;
              if (w < W) {
                const float rc = rcp_it[rnd], fn = fn_it[rnd];
                auto mean = [&](float sum) { const float q0 = sum * rc; return fmaf(fmaf(-q0, fn, sum), rc, q0); };
#pragma unroll
                for (int i = 0; i < kI; ++i) {
                  const int kb = st * kSupPair + ib + i;
                  const float4 m = make_float4(mean(acc[rnd][i].x), mean(acc[rnd][i].y), mean(acc[rnd][i].z), mean(acc[rnd][i].w));
                  if (p.protos_out && p.support)
                    *reinterpret_cast<float4*>(p.protos_out + ((size_t)e * W + w) * kD + kb * kBlockK + 4 * j) = m;
                  sq[rnd] = fmaf(m.x, m.x, fmaf(m.y, m.y, fmaf(m.z, m.z, fmaf(m.w, m.w, sq[rnd]))));
                  const uint32_t off = sw128(w, j);
                  sts4(b_base + (kb * 2 + 0) * kBTile + off, m);               // hi: the tensor core truncates it itself
                  sts4(b_base + (kb * 2 + 1) * kBTile + off, lo_of_raw(m));
                }
              }
            }
          }
#pragma unroll
          for (int rnd = 0; rnd < 2; ++rnd) {
            float x = sq[rnd];
            x += __shfl_xor_sync(kFullMask, x, 1);
            x += __shfl_xor_sync(kFullMask, x, 2);
            x += __shfl_xor_sync(kFullMask, x, 4);
            pp_acc[rnd] += x;
          }
          if (gtid == 0) HDBG(c, 8);
          fence_async_proxy();
          if (gtid == 0) HDBG(c, 9);
          __syncwarp();
          if (lane == 0) mbar_arrive(&meta->bars.ready[s]);
          if (gtid == 0) HDBG(c, 3);
          if (gtid == 32) HDBG(c, 19);
          if (gtid == 96) HDBG(c, 21);
        }
        // |p|^2, |q|^2 of parity `par` are free once the epilogue of task it - 2 is done: waited for only here, before the
        // first write (at the top of the task this wait chained support(t) behind epilogue(t - 2): with one support and
        // one query stage per task at D = 64 the period was (support + MMAs + epilogue) / 2)
        // (a group that took no support stage of this task has nothing to write - with an even number of stages per task
        // the groups' roles never change and its partials stay zero - and does not wait here: at D = 64 that wait held the
        // QUERY group's stage of task t behind the epilogue of task t - 2)
        if (did_sup || !kFixedRoles) {
          mbar_wait(&meta->bars.epi_done[par], ((it >> 1) & 1) ^ 1);
          epi_ok = true;
#pragma unroll
          for (int rnd = 0; rnd < 2; ++rnd) {
            const int item = gtid + rnd * kGroupThreads;
            if ((item & 7) == 0 && (item >> 3) < kTileN) meta->pp[par][grp][item >> 3] = pp_acc[rnd];
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&meta->bars.bk_free[bq]);                   // done with rows / cnt of this task
      }
      // ---------------------------------------------------------- query stages: lo tile for the raw tile, |q|^2
      {
        const int r0 = gtid >> 3, j = gtid & 7;                              // rows r0 + 16 t: 2 KB apart, same swizzle phase
        const uint32_t off0 = sw128(r0, j);
        float qacc[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) qacc[t] = 0.f;
#pragma unroll 1
        for (int st = 0; st < kSt; ++st, ++c, ++qc) {
          if ((int)(c & 1) != grp) continue;
          did_qry = true;
          const uint32_t s = c % kRing, ls = qc % kLo;
          mbar_wait(&meta->bars.tma_full[s], (c / kRing) & 1);
          if (gtid == 0) HDBG(c, 1);
          if (gtid == 32) HDBG(c, 18);
          if (gtid == 96) HDBG(c, 20);
#pragma unroll
          for (int i = 0; i < kPair; ++i) {
            const uint32_t src = ring + s * kStageB + i * tile_b + off0, dst = lo_base + ls * kStageB + i * tile_b + off0;
            // rows >= Nq of the tile were never written by TMA: whatever they hold only reaches accumulator rows nobody reads
            // (so rows >= Nq are neither read nor split: at Nq = 100 that is a fifth of the tile's shared-memory traffic)
            float4 v[8];
#pragma unroll
            for (int t = 0; t < 8; ++t)
              if (r0 + 16 * t < Nq) v[t] = lds4(src + t * 2048);
            if (i == 0) mbar_wait(&meta->bars.lo_free[ls], ((qc / kLo) & 1) ^ 1);   // MMAs of query stage qc - kLo are done with the slot
            if (i == 0 && gtid == 0) HDBG(c, 2);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              if (r0 + 16 * t < Nq) {
                sts4(dst + t * 2048, lo_of_raw(v[t]));
                qacc[t] = fmaf(v[t].x, v[t].x, fmaf(v[t].y, v[t].y, fmaf(v[t].z, v[t].z, fmaf(v[t].w, v[t].w, qacc[t]))));
              }
              if (t == 0 && gtid == 0) HDBG(c, 10 + 2 * (i & 1));
            }
            if (gtid == 0) HDBG(c, 11 + 2 * (i & 1));
          }
          if (gtid == 0) HDBG(c, 8);
          fence_async_proxy();
          if (gtid == 0) HDBG(c, 9);
          __syncwarp();
          if (lane == 0) mbar_arrive(&meta->bars.ready[s]);
          if (gtid == 0) HDBG(c, 3);
          if (gtid == 32) HDBG(c, 19);
          if (gtid == 96) HDBG(c, 21);
        }
        if (did_qry || !kFixedRoles) {
          if (!epi_ok) mbar_wait(&meta->bars.epi_done[par], ((it >> 1) & 1) ^ 1);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            float x = qacc[t];
            x += __shfl_xor_sync(kFullMask, x, 1);
            x += __shfl_xor_sync(kFullMask, x, 2);
            x += __shfl_xor_sync(kFullMask, x, 4);
            if (j == 0) meta->qq[par][grp][r0 + 16 * t] = x;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&meta->bars.meta_full[par]);
    }
  } else if (warp == kBucketWarp) {
    // =============================================================== bucket warp: class lists of the support rows, one task
    // ahead of the producers (it used to be a phase of the producers themselves, two 256-thread barriers and ~2000 clocks
    // per task on their critical path).  rows[w][0 .. cnt[w]) = the rows of class w in ascending order, 32 rows per step:
    // position inside the step from match.any, the running count of the class from shared memory (one writer per class)
    uint8_t* rows_w = rows_base;
    const int chunks = (sup_rows + 31) >> 5;
    int labs[kMaxSupportRows / 32];
    auto fetch = [&](int e) {
#pragma unroll
      for (int ch = 0; ch < kMaxSupportRows / 32; ++ch) {
        const int r = ch * 32 + lane;
        labs[ch] = -1;
        if (ch < chunks && r < sup_rows && e < p.E) labs[ch] = p.support ? p.s_labels[(size_t)e * p.Ns + r] : r;
      }
    };
    fetch(blockIdx.x);
    int it = 0;
    for (int e = blockIdx.x; e < p.E; e += gridDim.x, ++it) {
      const int bq = it & (kBk - 1);
      mbar_wait_relaxed(&meta->bars.bk_free[bq], ((it / kBk) & 1) ^ 1);
      int* cnt = meta->cnt[bq];
      uint8_t* rows = rows_w + (size_t)bq * W * row_stride;
      cnt[lane] = 0;
      __syncwarp();
#pragma unroll
      for (int ch = 0; ch < kMaxSupportRows / 32; ++ch) {
        if (ch < chunks) {
          const int lab = labs[ch];
          const bool valid = lab >= 0 && lab < W;                  // other labels are left out, as in the fp32-pipe kernels
          const unsigned same = __match_any_sync(kFullMask, lab);
          const int pos = __popc(same & ((1u << lane) - 1u));
          const int base = valid ? cnt[lab] : 0;
          __syncwarp();
          if (valid) {
            rows[lab * row_stride + base + pos] = (uint8_t)(ch * 32 + lane);
            if (pos == 0) cnt[lab] = base + __popc(same);
          }
          __syncwarp();
        }
      }
      fetch(e + gridDim.x);
      __syncwarp();
      if (lane == 0) mbar_arrive(&meta->bars.bk_full[bq]);
    }
  } else if (warp == kIssuerWarp) {
    // =============================================================== MMA issuer
    // The whole warp runs the (warp-uniform) control flow and the barrier waits; one elected lane issues.  The descriptors
    // of a stage are four 32-bit words computed once, each MMA adds a constant: ~4 instructions per MMA on the issuing
    // thread (a 128 x 32 x 8 MMA occupies the tensor pipe for 16 cycles, so the issue loop must not cost more than that).
    uint32_t c = 0, qc = 0;
    int it = 0;
    for (int e = blockIdx.x; e < p.E; e += gridDim.x, ++it) {
      const int par = it & 1;
#pragma unroll 1
      for (int st = 0; st < kSupSt; ++st, ++c) {                            // support stages: consumed by the producers only
        const uint32_t s = c % kRing;
        mbar_wait(&meta->bars.ready[s], (c / kRing) & 1);
        if (elect_one()) mbar_arrive(&meta->bars.empty[s]);
      }
      mbar_wait(&meta->bars.epi_done[par], ((it >> 1) & 1) ^ 1);            // accumulator `par` drained (task it-2)
      const uint32_t acc = tmem + par * kAccStride;
      const uint32_t b_base = b_base0 + (kBBuf == 2 ? par : 0) * kBSet;
#pragma unroll 1
      for (int st = 0; st < kSt; ++st, ++c, ++qc) {
        const uint32_t s = c % kRing, ls = qc % kLo;
        mbar_wait(&meta->bars.ready[s], (c / kRing) & 1);
        fence_after();
        if (lane == 0) HDBG(c, 4);
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < kPair; ++i) {
            const int kb = st * kPair + i;
            const uint32_t a_hi = desc_lo(ring + s * kStageB + i * tile_b), a_lo = desc_lo(lo_base + ls * kStageB + i * tile_b);
            // the hi and the lo tile of a k-block are adjacent (24 rows each): ONE B operand of 48 rows, so that a pass
            // over the A tile yields q.p_hi (columns 0..23) and q.p_lo (columns 24..47) at once - two reads of A per
            // k-block (lo, then hi: small terms first) instead of four.  The shared-memory data pipe was the limiter:
            // LSU + tensor-core wavefronts ~75 % of its cycles (profiles/r2r_ncu_kernels_summary.txt), 5 KB per
            // 128 x 32 x 8 MMA; a K step of 8 fp32 = 32 bytes = 2 descriptor units
            const uint32_t b_hl = desc_lo(b_base + kb * 2 * kBTile);
            mma_tf32_lo(acc, a_lo, b_hl, kIdesc, kb != 0);
#pragma unroll
            for (int k = 1; k < kBlockK / 8; ++k) mma_tf32_lo(acc, a_lo + 2 * k, b_hl + 2 * k, kIdesc, 1u);
#pragma unroll
            for (int k = 0; k < kBlockK / 8; ++k) mma_tf32_lo(acc, a_hi + 2 * k, b_hl + 2 * k, kIdesc, 1u);
          }
          commit(&meta->bars.empty[s]);
          commit(&meta->bars.lo_free[ls]);
          if (st == kSt - 1) {
            commit(&meta->bars.acc_full[par]);
            commit(&meta->bars.b_empty[kBBuf == 2 ? par : 0]);
          }
        }
        __syncwarp();
        if (lane == 0) HDBG(c, 5);
      }
    }
  } else {
    // =============================================================== epilogue: thread = query row
    const int quad = warp & 3;                                              // TMEM lanes 32 quad .. 32 quad + 31
    const int row = quad * 32 + lane;
    int it = 0;
    // the row's label is fetched ONE TASK AHEAD: read after the barrier waits, its HBM round trip (~1000 clocks under this
    // load) sat on the epilogue's critical path, a third of its ~3000 clocks per task (the bound of the D = 64 shapes)
    const bool live = row < Nq;
    int y_next = -1;
    if (live && p.q_labels && blockIdx.x < p.E) y_next = p.q_labels[(size_t)blockIdx.x * Nq + row];
    for (int e = blockIdx.x; e < p.E; e += gridDim.x, ++it) {
      const int par = it & 1, ph = (it >> 1) & 1;
      const int y = y_next;
      y_next = -1;
      if (live && p.q_labels && e + (int)gridDim.x < p.E) y_next = p.q_labels[(size_t)(e + gridDim.x) * Nq + row];
      mbar_wait_relaxed(&meta->bars.meta_full[par], ph);
      mbar_wait_relaxed(&meta->bars.acc_full[par], ph);
      fence_after();
      if (quad == 0 && lane == 0) HDBG((uint32_t)((it + 1) * (kSupSt + kSt) - 1), 6);
      // q.p of slot w = column w (q.p_hi) + column 24 + w (q.p_lo)
      float dot[kMaxWays];
      {
        uint32_t v[32], u[16];
        const uint32_t ta = tmem + ((uint32_t)(quad * 32) << 16) + par * kAccStride;
        tmem_ld32_nowait(ta, v);
        tmem_ld16_nowait(ta + 32, u);
        tmem_ld_wait();
        if (quad == 0 && lane == 0) HDBG((uint32_t)((it + 1) * (kSupSt + kSt) - 1), 14);
#pragma unroll
        for (int w = 0; w < kMaxWays; ++w)
          dot[w] = __uint_as_float(v[w]) + __uint_as_float(w + kMaxWays < 32 ? v[w + kMaxWays] : u[w + kMaxWays - 32]);
      }
      const float qq = meta->qq[par][0][live ? row : 0] + meta->qq[par][1][live ? row : 0];
      // |p|^2 of every slot, twelve 128-bit reads issued together (48 scalar reads inside the loops below, each waiting for
      // its own shared-memory round trip behind the producers' traffic, made the score loop 1900 of the epilogue's 2900
      // clocks per task: the bound of the D = 64 shapes, whose tasks have one stage per phase)
      float pw2s[kMaxWays];
      {
        const float4* pa = reinterpret_cast<const float4*>(meta->pp[par][0]);
        const float4* pb = reinterpret_cast<const float4*>(meta->pp[par][1]);
        float4 xa[kMaxWays / 4], xb[kMaxWays / 4];
#pragma unroll
        for (int k = 0; k < kMaxWays / 4; ++k) { xa[k] = pa[k]; xb[k] = pb[k]; }
#pragma unroll
        for (int k = 0; k < kMaxWays / 4; ++k) {
          pw2s[4 * k] = xa[k].x + xb[k].x; pw2s[4 * k + 1] = xa[k].y + xb[k].y;
          pw2s[4 * k + 2] = xa[k].z + xb[k].z; pw2s[4 * k + 3] = xa[k].w + xb[k].w;
        }
      }
      const bool need_scores = p.scores != nullptr || p.loss != nullptr;
      float m = -INFINITY, vy = 0.f, se = 1.f;
      int am = 0x7fffffff;
      if (need_scores) {
        float sc[kMaxWays];
#pragma unroll
        for (int w = 0; w < kMaxWays; ++w) {
          // |q|^2 + |p|^2 - 2 q.p clamped at 0, as at::_euclidean_dist; -sqrt = the score
          const float d2 = fmaxf(fmaf(-2.f, dot[w], qq) + pw2s[w], 0.f);
          const float s = -sqrtf(d2);
          sc[w] = s;
          if (w < W) {
            if (s > m) { m = s; am = w; }                                   // first index among equal maxima, like torch.max
            if (s != s && am == 0x7fffffff) am = w;                         // NaN row: keep something defined
            if (w == y) vy = s;
          }
        }
        if (p.loss) {
          se = 0.f;
#pragma unroll
          for (int w = 0; w < kMaxWays; ++w)
            if (w < W) se += expf(sc[w] - m);
        }
        if (live && p.scores) {
          float* dst = p.scores + ((size_t)e * Nq + row) * W;
#pragma unroll
          for (int w = 0; w < kMaxWays; ++w)
            if (w < W) dst[w] = sc[w];
        }
      } else {
        // evaluation only (labels, posterior, #correct): the smallest squared distance wins, one square root per row
        float best = INFINITY;
#pragma unroll
        for (int w = 0; w < kMaxWays; ++w) {
          const float d2 = fmaxf(fmaf(-2.f, dot[w], qq) + pw2s[w], 0.f);
          if (w < W) {
            if (d2 < best) { best = d2; am = w; }
            if (d2 != d2 && am == 0x7fffffff) am = w;
          }
        }
        m = -sqrtf(best);
      }
      if (quad == 0 && lane == 0) HDBG((uint32_t)((it + 1) * (kSupSt + kSt) - 1), 15);
      float nll = 0.f;
      int hit = 0;
      if (live) {
        const size_t r = (size_t)e * Nq + row;
        if (p.pred) p.pred[r] = am;
        if (p.posterior) p.posterior[r] = m;
        if (p.loss && y >= 0 && y < W) nll = -((vy - m) - logf(se));        // log_softmax then NLL
        hit = (am == y);
      }
      if (quad == 0 && lane == 0) HDBG((uint32_t)((it + 1) * (kSupSt + kSt) - 1), 16);
      if (p.loss || p.correct) {
        nll = warp_sum(nll);
        for (int o = 16; o > 0; o >>= 1) hit += __shfl_xor_sync(kFullMask, hit, o);
        if (lane == 0) { meta->part[quad] = nll; meta->hits[quad] = hit; }
        epilogue_bar();
        if (quad == 0 && lane == 0) {
          if (p.loss) p.loss[e] = (meta->part[0] + meta->part[1] + meta->part[2] + meta->part[3]) / (float)Nq;
          if (p.correct) p.correct[e] = meta->hits[0] + meta->hits[1] + meta->hits[2] + meta->hits[3];
        }
        epilogue_bar();
      }
      if (quad == 0 && lane == 0) HDBG((uint32_t)((it + 1) * (kSupSt + kSt) - 1), 17);
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&meta->bars.epi_done[par]);
      if (quad == 0 && lane == 0) HDBG((uint32_t)((it + 1) * (kSupSt + kSt) - 1), 7);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == kIssuerWarp) tmem_free<2 * kAccStride>(tmem);
}

template <int kD, int kPair, int kLo, bool kOneSup>
int launch_variant(const HeadParams& p_in, const CUtensorMap& ms, const CUtensorMap& mq, int sup_rows, cudaStream_t stream,
                   const char* name, bool* handled) {
  auto fn = head_tma_fwd_kernel<kD, kPair, kLo, kOneSup>;
  HeadParams p = p_in;
  // rows per k-block tile of a stage: the longer of the task's two blocks, rounded up to the swizzle atom
  const int longest = p.Nq > sup_rows ? p.Nq : sup_rows;
  p.tile_rows = (longest + 7) & ~7;
  const size_t tile_b = (size_t)p.tile_rows * 128, stage_b = kPair * tile_b;
  const size_t fixed = (size_t)kLo * stage_b + (size_t)(kD <= 128 ? 2 : 1) * (kD / kBlockK) * 2 * kBTile + (kTile - tile_b) + sizeof(TmaMeta) +
                       kBk * (size_t)p.W * ((sup_rows + 7) & ~7) + 1024 + 16;
  const size_t budget = 226 * 1024;
  if (fixed + 3 * stage_b > budget) return AFSL_OK;   // very long support blocks: the fp32-pipe kernels take the launch
  size_t ring_n = (budget - fixed) / stage_b;
  if (ring_n > (size_t)kMaxRing) ring_n = kMaxRing;
  if (const char* env = getenv("AFSL_HEAD_RING")) {   // fewer stages, for measurements
    const size_t want = (size_t)atoi(env);
    if (want >= 2 && want < ring_n) ring_n = want;
  }
  p.ring_stages = (int)ring_n;
  const size_t bytes = fixed + ring_n * stage_b;
  *handled = true;
  if (int rc = opt_in_smem(fn, bytes, name)) return rc;
  int sms = kNumSMs, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.E < sms ? p.E : sms;
  if (getenv("AFSL_HEAD_DBG")) {
    // columns: 0 TMA issued (stage free), 1 producers saw the bytes, 2 producers' tiles free, 3 producers done, 4 issuer saw
    // the stage, 5 MMAs issued, 6 epilogue saw the accumulator (last stage of a task), 7 epilogue done
    HeadParams q = p;
    static long long host[kDbgStages * kDbgEvents];
    if (cudaMalloc(&q.dbg, sizeof(host)) != cudaSuccess) return AFSL_ECUDA;
    cudaMemsetAsync(q.dbg, 0, sizeof(host), stream);
    fn<<<grid, kTmaThreads, bytes, stream>>>(q, ms, mq);
    cudaStreamSynchronize(stream);
    cudaMemcpy(host, q.dbg, sizeof(host), cudaMemcpyDeviceToHost);
    cudaFree(q.dbg);
    long long t0 = 0;
    for (long long t : host) if (t && (!t0 || t < t0)) t0 = t;
    fprintf(stderr, "head_tma timeline D=%d pair=%d (clocks since the first event; rows = ring stages, %d support + %d query per task)\n",
            kD, kPair, (kOneSup && kD / kBlockK <= kPair) ? 1 : kD / kBlockK / kPair, kD / kBlockK / kPair);
    for (int c = 0; c < kDbgStages; ++c) {
      fprintf(stderr, "st %3d:", c);
      for (int ev = 0; ev < kDbgEvents; ++ev) fprintf(stderr, " %7lld", host[c * kDbgEvents + ev] ? host[c * kDbgEvents + ev] - t0 : -1);
      fprintf(stderr, "\n");
    }
    AFSL_CHECK_LAUNCH(name);
    return AFSL_OK;
  }
  fn<<<grid, kTmaThreads, bytes, stream>>>(p, ms, mq);
  AFSL_CHECK_LAUNCH(name);
  return AFSL_OK;
}

}  // namespace

// Forward launches of fixed-size many-way tasks: 8 <= W <= 24, 25 < Nq <= 128 (the reference's cdist is in its matmul form
// there too), D in {64, 128, 256}, support block (<= 128 rows) or given prototypes.  AFSL_HEAD_MMA=0 disables it (the parity
// tests run this kernel and the fp32-pipe kernels on the same cases), AFSL_HEAD_MMA=2 selects the LDG-fed variant
// (proto_head_mma.cu).
int launch_head_tma(const HeadParams& p, bool bwd, cudaStream_t stream, const char* name, bool* handled) {
  *handled = false;
  if (bwd || !p.queries || p.q_offsets || !(p.support || p.protos_in)) return AFSL_OK;
  if (p.W < 8 || p.W > kMaxWays || p.Nq <= 25 || p.Nq > kTileM) return AFSL_OK;
  if (p.D != 64 && p.D != 128 && p.D != 256) return AFSL_OK;
  const int sup_rows = p.support ? p.Ns : p.W;
  if (sup_rows > kMaxSupportRows) return AFSL_OK;
  const char* env = getenv("AFSL_HEAD_MMA");
  if (env && atoi(env) != 1) return AFSL_OK;
  CUtensorMap ms, mq;
  const float* sup = p.support ? p.support : p.protos_in;
  if (int rc = make_tensor_map_f32(&ms, sup, (uint64_t)p.E * sup_rows, (uint64_t)p.D, (uint32_t)sup_rows, name)) return rc;
  if (int rc = make_tensor_map_f32(&mq, p.queries, (uint64_t)p.E * p.Nq, (uint64_t)p.D, (uint32_t)p.Nq, name)) return rc;
  // ring stages hold two k-blocks (half as many barrier round trips per task: 20w5s D = 256 0.59 -> 0.71 of HBM, 20w5s
  // D = 64 0.49 -> 0.59, gpurun_out/r2t_hb5_pair*.txt); AFSL_HEAD_PAIR=1 / 2 forces one / two (the parity tests run both)
  const char* pair_env = getenv("AFSL_HEAD_PAIR");
  const bool pair = pair_env ? atoi(pair_env) == 2 : true;
  HeadParams hp = p;
  const char* pf_env = getenv("AFSL_HEAD_L2PF");
  // measured (gpurun_out/r2t_hb6_*.txt, r2t_hb7_bulkpf.txt): +0.02 of HBM on the small-support shapes (20w1s D = 256: 0.497 ->
  // 0.518), within the noise elsewhere (20w5s D = 256: 0.707 / 0.697 / 0.687): on for support blocks of <= 32 rows
  hp.l2_prefetch = pf_env ? atoi(pf_env) != 0 : sup_rows <= 32;
  // support blocks of at most 32 rows in ONE ring stage: one barrier round trip for the support phase, but one producer
  // group then reduces all k-blocks while the other idles - a small gain at D = 64 (20w1s: 0.975 -> 0.948 ms per 48828
  // tasks), a loss at D = 256 (0.614 -> 0.802 ms per 12207 tasks), so only D = 64 takes it by default; AFSL_HEAD_ONESUP=1 / 0
  // forces it on / off (the parity test runs both)
  const char* one_env = getenv("AFSL_HEAD_ONESUP");
  const bool one = sup_rows <= 32 && (one_env ? atoi(one_env) != 0 : p.D == 64);
#define AFSL_TMA_VARIANT(D_, P_, L_) \
  (one ? launch_variant<D_, P_, L_, true>(hp, ms, mq, sup_rows, stream, name, handled) \
       : launch_variant<D_, P_, L_, false>(hp, ms, mq, sup_rows, stream, name, handled))
  // (the ring takes every stage that fits 226 KB of shared memory: 4 stages of 26 KB at D = 256 with 100-row blocks)
  if (p.D == 256) return pair ? AFSL_TMA_VARIANT(256, 2, 2) : AFSL_TMA_VARIANT(256, 1, 4);
  if (p.D == 128) return pair ? AFSL_TMA_VARIANT(128, 2, 2) : AFSL_TMA_VARIANT(128, 1, 4);
  return pair ? AFSL_TMA_VARIANT(64, 2, 2) : AFSL_TMA_VARIANT(64, 1, 4);
#undef AFSL_TMA_VARIANT
}

}  // namespace afsl
