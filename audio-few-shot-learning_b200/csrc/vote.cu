// Multi-segment majority vote with the reference's three tie strategies, batched over tasks.
//
// Reference: calculate_majority_vote_accuracy, loops/loops.py:169-247.  Per clip (distinct id, in
// any order - the counts do not depend on it): tally the predicted labels of its segments in
// segment order (collections.Counter keeps first-seen order); a unique top count wins; ties go to
// the smallest tied label ("min_label"), to the label of the first segment with the strictly
// greatest posterior among segments voting for a tied label ("max_posterior"; no segment beating
// -inf leaves no winner), or to the first-seen tied label (anything else).  The clip is correct
// when the winner equals the label of the clip's first segment.
//
// Integer work, bit-exact.  One WARP per task, 8 tasks per CTA, no CTA barrier: the task's four row arrays (16 bytes per
// segment) are staged in the warp's shared-memory slice with coalesced loads - every byte is read from HBM once - and the
// lowest-index segment of each clip ("leader") evaluates its clip out of shared memory.  Clip runs are short (<= 36
// segments), so the O(L^2) tally is a handful of shared-memory reads.  Tasks longer than the slice (kSliceRows segments)
// run the same code straight on the global arrays.
#include "afsl_common.cuh"

namespace afsl {
namespace {

constexpr int kVoteWarps = 8;
constexpr int kThreads = kVoteWarps * 32;
constexpr int kSliceRows = 320;                   // segments per staged task: 25 clips x up to ~12 segments on average
constexpr unsigned kAll = 0xffffffffu;

__device__ __forceinline__ void vote_task(const int32_t* pr, const int32_t* id, const int32_t* lb, const float* po, int n,
                                          int strategy, int lane, int& out_correct, int& out_clips) {
  bool ok = true;
  for (int k = lane + 1; k < n; k += 32) ok = ok && !(id[k] < id[k - 1]);
  const bool sorted = __all_sync(kAll, ok);
  int my_correct = 0, my_clips = 0;
  for (int k = lane; k < n; k += 32) {
    const int cid = id[k];
    bool leader;
    if (sorted) {
      leader = (k == 0) || (id[k - 1] != cid);
    } else {
      leader = true;
      for (int j = 0; j < k; ++j)
        if (id[j] == cid) { leader = false; break; }
    }
    if (!leader) continue;
    int end = n;
    if (sorted) {
      end = k + 1;
      while (end < n && id[end] == cid) ++end;
    }
    // pass 1: distinct labels in first-seen order, their counts, the top count
    int top = 0, n_tied = 0, win_first = -1, win_min = 0x7fffffff;
    for (int j = k; j < end; ++j) {
      if (id[j] != cid) continue;
      const int v = pr[j];
      bool seen = false;
      for (int i = k; i < j; ++i)
        if (id[i] == cid && pr[i] == v) { seen = true; break; }
      if (seen) continue;
      int c = 0;
      for (int i = j; i < end; ++i) c += (id[i] == cid && pr[i] == v);
      if (c > top) { top = c; n_tied = 1; win_first = v; win_min = v; }
      else if (c == top) { ++n_tied; if (v < win_min) win_min = v; }
    }
    int winner = win_first;
    if (n_tied > 1) {
      if (strategy == AFSL_TIE_MIN_LABEL) {
        winner = win_min;
      } else if (strategy == AFSL_TIE_MAX_POSTERIOR) {
        float best = -INFINITY;
        winner = -1;
        for (int j = k; j < end; ++j) {
          if (id[j] != cid) continue;
          const int v = pr[j];
          int c = 0;
          for (int i = k; i < end; ++i) c += (id[i] == cid && pr[i] == v);
          if (c == top && po[j] > best) { best = po[j]; winner = v; }
        }
      }
    }
    ++my_clips;
    my_correct += (winner == lb[k]);
  }
  out_correct = __reduce_add_sync(kAll, my_correct);
  out_clips = __reduce_add_sync(kAll, my_clips);
}

__global__ void __launch_bounds__(kThreads) vote_kernel(const int32_t* __restrict__ pred, const int32_t* __restrict__ ids,
                                                         const int32_t* __restrict__ labels,
                                                         const float* __restrict__ post,
                                                         const int32_t* __restrict__ offsets, int strategy,
                                                         int32_t* __restrict__ correct_clips,
                                                         int32_t* __restrict__ n_clips, int E) {
  __shared__ int32_t s_rows[kVoteWarps][4][kSliceRows];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t* s_pr = s_rows[warp][0];
  int32_t* s_id = s_rows[warp][1];
  int32_t* s_lb = s_rows[warp][2];
  float* s_po = reinterpret_cast<float*>(s_rows[warp][3]);
  for (int e = blockIdx.x * kVoteWarps + warp; e < E; e += gridDim.x * kVoteWarps) {
    const int s0 = offsets[e], n = offsets[e + 1] - s0;
    int c = 0, m = 0;
    if (n <= kSliceRows) {
      __syncwarp();                                       // the previous task's readers are done with the slice
      for (int k = lane; k < n; k += 32) {
        s_pr[k] = pred[s0 + k];
        s_id[k] = ids[s0 + k];
        s_lb[k] = labels[s0 + k];
        s_po[k] = post[s0 + k];
      }
      __syncwarp();
      vote_task(s_pr, s_id, s_lb, s_po, n, strategy, lane, c, m);
    } else {
      vote_task(pred + s0, ids + s0, labels + s0, post + s0, n, strategy, lane, c, m);
    }
    if (lane == 0) { correct_clips[e] = c; n_clips[e] = m; }
  }
}

}  // namespace
}  // namespace afsl

extern "C" int afsl_eval_vote_i32(const int32_t* pred, const int32_t* clip_ids, const int32_t* labels,
                                   const float* posterior, const int32_t* seg_offsets, int tie_strategy,
                                   int32_t* correct_clips, int32_t* n_clips, int E, void* stream) {
  AFSL_REQUIRE(pred && clip_ids && labels && posterior && seg_offsets && correct_clips && n_clips,
               "afsl_eval_vote_i32: null pointer");
  AFSL_REQUIRE(tie_strategy >= 0 && tie_strategy <= 2, "afsl_eval_vote_i32: unknown tie strategy %d", tie_strategy);
  AFSL_REQUIRE(E >= 0, "afsl_eval_vote_i32: E=%d", E);
  if (E == 0) return AFSL_OK;
  const int cap = afsl::kNumSMs * 8;                        // 8 CTAs of 8 warps per SM (40 KB of slices each)
  const int want = (E + afsl::kVoteWarps - 1) / afsl::kVoteWarps;
  const int grid = want < cap ? want : cap;
  afsl::vote_kernel<<<grid, afsl::kThreads, 0, (cudaStream_t)stream>>>(pred, clip_ids, labels, posterior, seg_offsets,
                                                                        tie_strategy, correct_clips, n_clips, E);
  AFSL_CHECK_LAUNCH("afsl_eval_vote_i32");
  return AFSL_OK;
}
