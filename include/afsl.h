/*
 * libafsl - C ABI of the B200 (sm_100a) episodic prototypical-network head.
 *
 * The reference (magcil/audio-few-shot-learning) is pure Python/PyTorch and has
 * no FFI layer; its boundary for this path is the Python object API used by
 * loops/loops.py:40-49,76-79,268-277.  Each entry point below replaces the eager
 * ATen op chain of one reference function (cited per function) and is bound from
 * Python with ctypes (audio-few-shot-learning_b200/_lib.py); INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer is DEVICE memory of the current CUDA device, owned by the
 *    caller, contiguous, 16-byte aligned; nullable arguments are marked [opt];
 *  - floating point is IEEE fp32, labels/indices are int32;
 *  - a leading episode dimension E batches independent episodes / tasks
 *    (the reference always has E = 1);
 *  - `stream` is a cudaStream_t passed as void*; the library never allocates,
 *    frees, synchronises or touches any RNG; all calls are re-entrant;
 *  - host-drawn randomness (mask offsets, warp control points, CPL keep masks)
 *    enters as explicit arrays so results are bit-reproducible;
 *  - return value: 0 on success, otherwise an AFSL_E* code, with a thread-local
 *    message available from afsl_last_error().
 */
#ifndef AFSL_H
#define AFSL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFSL_OK 0
#define AFSL_EINVAL 1   /* bad shape / null pointer / unsupported size */
#define AFSL_ECUDA 2    /* a CUDA runtime call or launch failed */

#define AFSL_ABI_VERSION 1

/* tie strategies of calculate_majority_vote_accuracy (loops/loops.py:216-234) */
#define AFSL_TIE_FIRST_SEEN 0     /* any other string, e.g. the config default "" */
#define AFSL_TIE_MIN_LABEL 1      /* "min_label" */
#define AFSL_TIE_MAX_POSTERIOR 2  /* "max_posterior" */

int afsl_version(void);
const char* afsl_last_error(void);
/* number of kernel launches issued by this library in this process (bench bookkeeping) */
long long afsl_launch_count(void);

/* ---------------------------------------------------------------------------
 * Prototypes: per-label mean of the support rows.
 * Replaces compute_prototypes, models/util_functions.py:6-19.
 *   support [E,Ns,D], labels [E,Ns] (values 0..W-1)  ->  protos [E,W,D]
 * bwd: d_support[e,k,:] = d_protos[e,label[e,k],:] / count[e,label[e,k]]
 * ------------------------------------------------------------------------- */
int afsl_prototypes_fwd_f32(const float* support, const int32_t* labels, float* protos,
                            int E, int Ns, int W, int D, void* stream);
int afsl_prototypes_bwd_f32(const float* d_protos, const int32_t* labels, float* d_support,
                            int E, int Ns, int W, int D, void* stream);

/* ---------------------------------------------------------------------------
 * Scores / prototypical loss given prototypes.
 * Replaces l2_distance_to_prototypes (models/few_shot_classifier.py:108-116),
 * FSL_Loss.forward (loops/loss.py:24-37) and the argmax/accuracy of
 * evaluate_on_one_task (loops/loops.py:79) and loops/loops.py:271-272.
 *   protos [E,W,D], queries [rows,D]; episode e owns query rows
 *   q_offsets[e]..q_offsets[e+1] when q_offsets != NULL (ragged multi-segment
 *   tasks), else rows e*Nq..(e+1)*Nq.
 * outputs, each [opt]:
 *   scores [rows,W] = -||q - p||_2 ; loss [E] = mean_i NLL(log_softmax(scores_i), y_i)
 *   pred [rows] first-index argmax ; posterior [rows] = max score ; correct [E]
 * q_labels [opt] is required for loss / correct.
 * ------------------------------------------------------------------------- */
int afsl_proto_scores_fwd_f32(const float* protos, const float* queries, const int32_t* q_labels,
                              const int32_t* q_offsets, float* scores, float* loss, int32_t* pred,
                              float* posterior, int32_t* correct, int E, int Nq, int W, int D,
                              void* stream);
/* gradient of sum_e d_loss[e]*loss[e] (+ sum d_scores*scores when d_scores != NULL)
 * w.r.t. protos and queries.  d_protos / d_queries are overwritten. */
int afsl_proto_scores_bwd_f32(const float* protos, const float* queries, const int32_t* q_labels,
                              const int32_t* q_offsets, const float* d_loss, const float* d_scores,
                              float* d_protos, float* d_queries, int E, int Nq, int W, int D,
                              void* stream);

/* ---------------------------------------------------------------------------
 * Fused head: prototypes + distances + log-softmax + NLL (+ argmax/accuracy)
 * in one pass over the support and query embeddings.  Replaces the chain
 * process_support_set -> forward -> FSL_Loss of loops/loops.py:40-42 and the
 * evaluation chain of loops/loops.py:76-79.
 * Outputs as above plus protos [E,W,D] [opt].
 * bwd: d_support, d_queries of sum_e d_loss[e]*loss[e]; d_protos_extra [E,W,D]
 * [opt] is an additional gradient arriving at the prototypes from another
 * consumer (the CPL / angular branch, loops/loops.py:44-50).  protos [E,W,D]
 * [opt]: the prototypes the forward wrote; when given, the support block is
 * not re-read (support may then be NULL), otherwise they are recomputed.
 * ------------------------------------------------------------------------- */
int afsl_proto_head_fwd_f32(const float* support, const int32_t* s_labels, const float* queries,
                            const int32_t* q_labels, const int32_t* q_offsets, float* protos,
                            float* scores, float* loss, int32_t* pred, float* posterior,
                            int32_t* correct, int E, int Ns, int Nq, int W, int D, void* stream);
int afsl_proto_head_bwd_f32(const float* support, const float* protos, const int32_t* s_labels,
                            const float* queries, const int32_t* q_labels, const int32_t* q_offsets,
                            const float* d_loss, const float* d_protos_extra, float* d_support,
                            float* d_queries, int E, int Ns, int Nq, int W, int D, void* stream);

/* ---------------------------------------------------------------------------
 * Row L2 normalisation x / max(||x||, eps).  Replaces F.normalize on the
 * prototypes, loops/loops.py:47-48 (eps 1e-12).  rows = E*W.
 * ------------------------------------------------------------------------- */
int afsl_l2_normalize_fwd_f32(const float* x, float* y, int rows, int D, float eps, void* stream);
int afsl_l2_normalize_bwd_f32(const float* x, const float* d_y, float* d_x, int rows, int D, float eps,
                              void* stream);

/* ---------------------------------------------------------------------------
 * CPL loss.  Replaces CPL_Loss.forward / similarity_sampling, loops/loss.py:118-165,
 * in its closed form: C = cos(P,Q)/T (each vector / max(norm,1e-8)),
 *   loss[e] = 1/Nq^2 * sum_i [ LSE_{j in keep_i} C[y_i,j] - C[y_i,i] ].
 * keep [E,Nq,ceil(Nq/32)] is the host-drawn bit mask (bit j of row i set when
 * query j is a sampled negative of query i, or j == i); keep == NULL selects
 * the deterministic case M >= per-class count: keep_ij = (j == i) | (y_j != y_i).
 *   protos [E,W,D], queries [E,Nq,D], labels [E,Nq]
 * bwd overwrites d_protos [E,W,D], d_queries [E,Nq,D] with the gradient of
 * sum_e d_loss[e]*loss[e].
 * ------------------------------------------------------------------------- */
int afsl_cpl_fwd_f32(const float* protos, const float* queries, const int32_t* labels,
                     const uint32_t* keep, float temperature, float* loss, int E, int Nq, int W,
                     int D, void* stream);
int afsl_cpl_bwd_f32(const float* protos, const float* queries, const int32_t* labels,
                     const uint32_t* keep, float temperature, const float* d_loss, float* d_protos,
                     float* d_queries, int E, int Nq, int W, int D, void* stream);
/* The same pair with the forward's similarity matrix handed to the backward: afsl_cpl_fwd_save_f32 also writes
 * sim [E,W,Nq] = C and qinv [E,Nq] = 1/|q_j| (negative where the norm was clamped), afsl_cpl_bwd_saved_f32 takes
 * them back and walks the query rows ONCE instead of twice (what torch.autograd does for the reference: loss.py:118-165
 * keeps its similarity matrix alive for the backward).  Only for the shapes afsl_cpl_saved_supported(Nq, W, D) accepts
 * (5-way, D in {64,128,256}: the warp-per-episode kernels); elsewhere use the pair above. */
int afsl_cpl_saved_supported(int Nq, int W, int D);
int afsl_cpl_fwd_save_f32(const float* protos, const float* queries, const int32_t* labels,
                          const uint32_t* keep, float temperature, float* loss, float* sim, float* qinv,
                          int E, int Nq, int W, int D, void* stream);
int afsl_cpl_bwd_saved_f32(const float* protos, const float* queries, const int32_t* labels,
                           const uint32_t* keep, float temperature, const float* sim, const float* qinv,
                           const float* d_loss, float* d_protos, float* d_queries, int E, int Nq, int W,
                           int D, void* stream);

/* ---------------------------------------------------------------------------
 * Angular loss.  Replaces AngularLossClass.forward, loops/loss.py:48-97
 * (pytorch_metric_learning AngularMiner(angle) -> AngularLoss(alpha=40 deg)),
 * mining included, in the multiplicity-weighted closed form.
 * anchors != 0: prototypes_as_anchors=True branch (loss.py:68-83);
 * anchors == 0: prototypes and queries pooled (loss.py:84-96).
 * normalize_ref selects whether the reference rows inside the loss are
 * L2-normalised (identical for unit-norm inputs).
 * ------------------------------------------------------------------------- */
int afsl_angular_fwd_f32(const float* protos, const float* queries, const int32_t* labels,
                         float miner_angle_deg, float alpha_deg, int anchors, int normalize_ref,
                         float* loss, int E, int Nq, int W, int D, void* stream);
int afsl_angular_bwd_f32(const float* protos, const float* queries, const int32_t* labels,
                         float miner_angle_deg, float alpha_deg, int anchors, int normalize_ref,
                         const float* d_loss, float* d_protos, float* d_queries, int E, int Nq,
                         int W, int D, void* stream);

/* ---------------------------------------------------------------------------
 * SpecAugment views.  Replaces SpecAugment.apply_augmentations and the three
 * transforms, utils/augmentations.py:33-157.
 *   x [N,1,F,T]  ->  views [4,N,1,F,T] = {copy, time-warp, time-mask, freq-mask}
 * Samples are grouped in sets of `set_size` consecutive samples (one
 * apply_augmentations call of the reference = one set); masks are shared by a
 * set: time_masks/freq_masks [n_sets,num_mask,2] = (start, length) int32.
 * set_ids [N] int32 [opt] names the set of every sample instead (ragged sets: the
 * packed query segments of multi-segment tasks, one set per task).
 * warp_p, warp_d [N] int32 are the per-sample control point and displacement;
 * src_x [N,T] [opt] overrides the in-kernel Hermite spline with host-computed
 * normalised source coordinates.  row_lo [F] int32 / row_w [F] fp32 are the
 * source row and blend weight of grid_sample's y axis (host-computed once from
 * linspace(-1,1,F)).  views_mask selects which of the 4 views are written
 * (bit v), so callers can skip the plain copy.
 * ------------------------------------------------------------------------- */
int afsl_specaug_views_f32(const float* x, float* views, const int32_t* warp_p, const int32_t* warp_d,
                           const float* src_x, const int32_t* row_lo, const float* row_w,
                           const int32_t* set_ids, const int32_t* time_masks, const int32_t* freq_masks,
                           int num_mask, float mask_value, int N, int set_size, int F, int T, int views_mask,
                           void* stream);

/* ---------------------------------------------------------------------------
 * Multi-segment majority vote.  Replaces calculate_majority_vote_accuracy,
 * loops/loops.py:169-247.  Task e owns segments seg_offsets[e]..seg_offsets[e+1]
 * of pred / clip_ids / labels / posterior.  Outputs per task the number of
 * clips whose voted label equals the label of the clip's first segment, and the
 * number of distinct clips (accuracy = correct/clips, divided on the host in
 * float64 like the reference).
 * ------------------------------------------------------------------------- */
int afsl_eval_vote_i32(const int32_t* pred, const int32_t* clip_ids, const int32_t* labels,
                       const float* posterior, const int32_t* seg_offsets, int tie_strategy,
                       int32_t* correct_clips, int32_t* n_clips, int E, void* stream);

/* ---------------------------------------------------------------------------
 * View fusion: one post-norm transformer encoder layer (1 head, ReLU FFN, d_model 64,
 * dim_feedforward 256) over the V views of every sample; the output [N, V*64] is the
 * V tokens laid side by side.  Replaces SelfAttention.forward,
 * models/main_modules.py:222-228 (nn.TransformerEncoderLayer).
 *   x [N,V,64] -> y [N,V,64]
 * weights: packed fp32 block of afsl_view_fusion_weight_floats() floats =
 *   in_proj_weight[192,64] | in_proj_bias | out_proj.weight[64,64] | bias |
 *   linear1.weight[256,64] | bias | linear2.weight[64,256] | bias | norm1.w | norm1.b |
 *   norm2.w | norm2.b   (= afsl_view_fusion_param_floats() parameters)
 *   followed by the transposed copies in_proj^T | out_proj^T | linear1^T | linear2^T.
 * drop_* [opt]: keep masks already scaled by 1/(1-p) for the four dropout sites
 *   (attention probabilities [N,V,V], after out_proj [N,V,64], after ReLU [N,V,256],
 *   after linear2 [N,V,64]); NULL = no dropout (eval mode).
 * bwd recomputes the forward per tile; d_weights_partial is a ZERO-INITIALISED buffer
 * [afsl_view_fusion_grid(N,V), param_floats] of per-CTA partial sums (the caller adds
 * them up: no atomics, reproducible).
 * ------------------------------------------------------------------------- */
int afsl_view_fusion_weight_floats(void);
int afsl_view_fusion_param_floats(void);
int afsl_view_fusion_grid(int N, int V);
int afsl_view_fusion_fwd_f32(const float* x, const float* weights, float* y, const float* drop_attn,
                             const float* drop1, const float* drop_ffn, const float* drop2, int N, int V,
                             int d, int ffn, void* stream);
int afsl_view_fusion_bwd_f32(const float* x, const float* weights, const float* d_y, const float* drop_attn,
                             const float* drop1, const float* drop_ffn, const float* drop2, float* d_x,
                             float* d_weights_partial, int N, int V, int d, int ffn, void* stream);

/* ---------------------------------------------------------------------------
 * Grouped BatchNorm + ReLU + MaxPool(3, stride 3) over convolution outputs whose
 * batch axis holds G groups of `group` consecutive samples with independent batch
 * statistics - one group = one encoder call of the reference (conv_block,
 * models/main_modules.py:43-60, applied per view per set per episode, :18-23), so
 * batching E episodes keeps the per-(episode, set, view) statistics.
 *   x [G*group,C,H,W] NCHW -> y [G*group,C,H/3,W/3]
 * stats: mean / rstd = 1/sqrt(var_biased + eps) / var_biased, each [G,C].
 * fwd / bwd take mean, rstd as [G,C] (stats_per_group = 1, training) or [C]
 * (stats_per_group = 0, running statistics in eval mode).
 * bwd writes d_x (like x) and sums [G,C,2] = (sum dz, sum dz*xhat) per slab, from
 * which d_beta[c] = sum_g sums[g,c,0], d_gamma[c] = sum_g sums[g,c,1].
 * ------------------------------------------------------------------------- */
int afsl_gbn_stats_f32(const float* x, float* mean, float* rstd, float* var_biased, int G, int group,
                       int C, int H, int W, float eps, void* stream);
int afsl_gbn_relu_pool_fwd_f32(const float* x, const float* mean, const float* rstd, const float* gamma,
                               const float* beta, float* y, int G, int group, int C, int H, int W,
                               int stats_per_group, void* stream);
int afsl_gbn_relu_pool_bwd_f32(const float* x, const float* mean, const float* rstd, const float* gamma,
                               const float* beta, const float* d_y, float* d_x, float* sums, int G,
                               int group, int C, int H, int W, int stats_per_group, void* stream);

/* ---------------------------------------------------------------------------
 * Encoder stage 1 fused: Conv3x3(1 -> 64, pad 1) + grouped BatchNorm + ReLU +
 * MaxPool(3,3) without materialising the full-resolution activation.  Replaces the
 * first conv_block of conv_encoder (models/main_modules.py:43-81) per group of
 * `group` consecutive samples (one encoder call of the reference, :18-23).
 *   x [G*group,1,H,W] -> y [G*group,64,H/3,W/3]
 * moments: per group and part the 9 shifted sums and 45 shifted products of the input
 *   (double [G,parts,54]: S_0..S_8, then R_kl for k <= l row by row), obtained from the 13
 *   image autocorrelations at displacements in [-2,2]^2 plus edge-row / edge-column terms; mean / variance
 *   of every conv channel follow from them on the host.
 * fwd: z = a*u + b with u the bias-free conv output; a, b (and mean, rstd of u in
 *   bwd) are [G,64] (per_group = 1) or [64].
 * bwd: partial [G,parts,64,11] = per channel (sum dz, sum dz*xhat, T_0..T_8) with
 *   T_k = sum dz * x(p_argmax + tap k); weight / BatchNorm gradients are assembled
 *   from these and the moments on the host.  The input gets no gradient.
 * channels_last: y, argmax and d_y are [G*group,H/3,W/3,64] instead of [G*group,64,H/3,W/3].
 * argmax [G*group,64,H/3,W/3] bytes [opt]: written by fwd (window element 0..8 that won the
 *   max, 15 where the ReLU zeroed the output); when passed to bwd only that element is
 *   recomputed instead of the whole 3x3 window of convolution outputs.
 * ------------------------------------------------------------------------- */
int afsl_stage1_channels(void);
int afsl_stage1_acc_slots(void);
int afsl_stage1_moments_f64(const float* x, double* moments, int parts, int G, int group, int H, int W,
                            void* stream);
int afsl_stage1_fwd_f32(const float* x, const float* weight, const float* a, const float* b, float* y,
                        unsigned char* argmax, int G, int group, int H, int W, int per_group, int channels_last,
                        void* stream);
int afsl_stage1_bwd_f32(const float* x, const float* weight, const float* a, const float* b, const float* mean,
                        const float* rstd, const float* d_y, const unsigned char* argmax, float* partial, int parts,
                        int G, int group, int H, int W, int per_group, int channels_last, void* stream);

/* ---------------------------------------------------------------------------
 * Channels-last (NHWC) variants of the grouped BatchNorm + ReLU + MaxPool kernels:
 * x physically [G*group,H,W,C], y / d_y [G*group,H/3,W/3,C], C a multiple of 4.
 * cuDNN's sm_100 convolutions are NHWC-native; with channels-last activations no
 * layout-conversion kernels run around them.  The per-(group, channel) reductions are
 * two-stage: partial is caller-provided workspace, double [G,parts,C,2], with
 * parts = afsl_gbn_nhwc_parts(G).  channels_last = 1 in afsl_stage1_{fwd,bwd}_f32
 * makes stage 1 write y / argmax (and read d_y) in this layout.
 * ------------------------------------------------------------------------- */
int afsl_gbn_nhwc_parts(int G);
int afsl_gbn_stats_nhwc_f32(const float* x, double* partial, int parts, float* mean, float* rstd,
                            float* var_biased, int G, int group, int C, int H, int W, float eps, void* stream);
int afsl_gbn_relu_pool_nhwc_fwd_f32(const float* x, const float* mean, const float* rstd, const float* gamma,
                                    const float* beta, float* y, int G, int group, int C, int H, int W,
                                    int stats_per_group, void* stream);
int afsl_gbn_relu_pool_nhwc_bwd_f32(const float* x, const float* mean, const float* rstd, const float* gamma,
                                    const float* beta, const float* d_y, float* d_x, double* partial, int parts,
                                    float* sums, int G, int group, int C, int H, int W, int stats_per_group,
                                    void* stream);

/* ---------------------------------------------------------------------------
 * Glue kernels around the encoder stages (each replaces a chain of tiny eager ops).
 * bn_running_update: the G momentum updates of BatchNorm's running statistics that G
 *   separate module calls would make, in group order: running_mean <- (1-m) rm + m (mean_g + shift),
 *   running_var <- (1-m) rv + m var_g * unbias; mean, var_biased [G,C]; shift [C] [opt] (a
 *   convolution bias folded out of the statistics); num_batches_tracked [opt] += G.
 * stage1_finalize: moments [G,parts,54] (afsl_stage1_moments_f64) -> S [G,9], R [G,9,9] (double)
 *   and per (group, conv-1 channel) mean_u, var (biased), rstd, a = gamma*rstd, b = beta - mean_u*a.
 * stage1_dw: partial [G,parts,64,11] (afsl_stage1_bwd_f32) + S, R -> d_w [64,9], d_gamma, d_beta [64]
 *   and d_bias [64] [opt] (zero under batch statistics).
 * ------------------------------------------------------------------------- */
int afsl_bn_running_update_f32(const float* mean, const float* var_biased, const float* shift,
                               float* running_mean, float* running_var, long long* num_batches_tracked,
                               float momentum, float unbias, int G, int C, void* stream);
int afsl_stage1_finalize_f64(const double* moments, int parts, const float* weight, const float* gamma,
                             const float* beta, float eps, double count, double* S, double* R, float* mean_u,
                             float* var, float* rstd, float* a, float* b, int G, void* stream);
int afsl_stage1_dw_f32(const float* partial, int parts, int G, const double* S, const double* R,
                       const float* weight, const float* a, const float* mean_u, const float* rstd,
                       double count, int per_group, float* d_w, float* d_gamma, float* d_beta, float* d_bias,
                       void* stream);

/* ---------------------------------------------------------------------------
 * Layout switch of the encoder wrapper: y[n][B][A] = x[n][A][B].  With A = H*W, B = C it turns the
 * channels-last output of the fused first block into the NCHW tensor on which cuDNN's fp32 (TF32 off) convolutions
 * of the following conv_blocks (models/main_modules.py:43-60) run fastest on sm_100; with A = C, B = H*W it is its
 * backward.  Replaces Tensor.contiguous() (3.7 ms for the 3.6 GB stage-1 output of a 32-episode step).
 * ------------------------------------------------------------------------- */
int afsl_transpose_f32(const float* x, float* y, int n, int A, int B, void* stream);

/* ---------------------------------------------------------------------------
 * F2: the Linear layers of ProjectionHead (models/main_modules.py:231-255: fc1 -> ReLU -> fc2 -> L2 normalise;
 * the closing normalisation is afsl_l2_normalize_*).  Replaces torch.nn.functional.linear / cuBLAS on this path.
 * fwd: y[M,N] = x[M,K] . w[N,K]^T + bias[N] [opt], ReLU when relu != 0.
 * bwd: with g = d_y masked by (y_relu > 0) when y_relu (the forward's ReLU output) is given:
 *      d_x[M,K] = g . w [opt], d_w[N,K] = g^T . x [opt], d_bias[N] = column sums of g [opt].
 *      workspace: afsl_linear_bwd_workspace_floats(M,N,K) floats (row-split partial sums, reduced in fixed order).
 * ------------------------------------------------------------------------- */
int afsl_linear_fwd_f32(const float* x, const float* w, const float* bias, float* y, int M, int N, int K, int relu,
                        void* stream);
long long afsl_linear_bwd_workspace_floats(int M, int N, int K);
int afsl_linear_bwd_f32(const float* x, const float* w, const float* y_relu, const float* d_y, float* d_x, float* d_w,
                        float* d_bias, float* workspace, int M, int N, int K, void* stream);

/* ---------------------------------------------------------------------------
 * f-4: waveform front end (datasets/batch_creation.py:138-143,215-218 with the MelSpectrogram of src/train_test.py:123-129):
 * wave [N,L] (16 kHz) -> out [N,1,128,T], T = L/hop + 1:
 *   STFT (n_fft = 1024, window [1024] as given, centred frames, reflect padding) -> |X|^2 -> mel filters in band form
 *   (filter m: fb_len[m] weights fb_w[fb_off[m] ..] on bins fb_start[m] ..) -> 10 log10(mel + eps) -> (x - mean) / std.
 * One launch; the window and the filterbank are the reference transform's own tensors.
 * ------------------------------------------------------------------------- */
int afsl_logmel_f32(const float* wave, const float* window, const int32_t* fb_start, const int32_t* fb_len,
                    const int32_t* fb_off, const float* fb_w, float* out, int N, int L, int T, int n_fft, int hop,
                    int n_mels, float eps, float mean, float std, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AFSL_H */
