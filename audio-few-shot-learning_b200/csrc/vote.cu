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
// Integer work, bit-exact.  One CTA per task; the lowest-index segment of each clip ("leader")
// evaluates its clip.  Clip runs are short (<= 36 segments), so the O(L^2) tally stays in L1.
#include "afsl_common.cuh"

namespace afsl {
namespace {

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads) vote_kernel(const int32_t* __restrict__ pred, const int32_t* __restrict__ ids,
                                                         const int32_t* __restrict__ labels,
                                                         const float* __restrict__ post,
                                                         const int32_t* __restrict__ offsets, int strategy,
                                                         int32_t* __restrict__ correct_clips,
                                                         int32_t* __restrict__ n_clips, int E) {
  __shared__ int s_sorted, s_correct, s_clips;
  for (int e = blockIdx.x; e < E; e += gridDim.x) {
    const int s0 = offsets[e], n = offsets[e + 1] - s0;
    const int32_t* pr = pred + s0;
    const int32_t* id = ids + s0;
    const int32_t* lb = labels + s0;
    const float* po = post + s0;
    if (threadIdx.x == 0) { s_sorted = 1; s_correct = 0; s_clips = 0; }
    __syncthreads();
    for (int k = threadIdx.x + 1; k < n; k += kThreads)
      if (id[k] < id[k - 1]) s_sorted = 0;  // benign race: every writer stores 0
    __syncthreads();
    const bool sorted = s_sorted != 0;
    int my_correct = 0, my_clips = 0;
    for (int k = threadIdx.x; k < n; k += kThreads) {
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
    if (my_clips) { atomicAdd(&s_clips, my_clips); atomicAdd(&s_correct, my_correct); }
    __syncthreads();
    if (threadIdx.x == 0) { correct_clips[e] = s_correct; n_clips[e] = s_clips; }
    __syncthreads();
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
  const int cap = afsl::kNumSMs * 16;
  const int grid = E < cap ? E : cap;
  afsl::vote_kernel<<<grid, afsl::kThreads, 0, (cudaStream_t)stream>>>(pred, clip_ids, labels, posterior, seg_offsets,
                                                                        tie_strategy, correct_clips, n_clips, E);
  AFSL_CHECK_LAUNCH("afsl_eval_vote_i32");
  return AFSL_OK;
}
