// Parameter block shared by the prototype-head kernels (proto_head.cu: one CTA per episode, any shape;
// proto_head_warp.cu: one warp per episode with the prototypes in registers, small W*D).
#pragma once

#include "afsl_common.cuh"

namespace afsl {

struct HeadParams {
  // forward inputs
  const float* support;     // [E,Ns,D] or null (then protos_in is used)
  const int32_t* s_labels;  // [E,Ns]
  const float* protos_in;   // [E,W,D] or null
  const float* queries;     // [rows,D] or null (prototype-only call)
  const int32_t* q_labels;  // [rows] or null
  const int32_t* q_offsets; // [E+1] or null
  // forward outputs (nullable)
  float* protos_out;
  float* scores;
  float* loss;
  int32_t* pred;
  float* posterior;
  int32_t* correct;
  // backward
  const float* d_loss;          // [E]
  const float* d_scores;        // [rows,W] or null
  const float* d_protos_extra;  // [E,W,D] or null
  float* d_support;             // [E,Ns,D] or null
  float* d_protos;              // [E,W,D] or null
  float* d_queries;             // [rows,D] or null
  int E, Ns, Nq, W, D;
  long long* dbg;               // proto_head_tma.cu timeline buffer (AFSL_HEAD_DBG), else null
  int tile_rows, ring_stages;   // proto_head_tma.cu: rows per k-block tile of a ring stage, ring depth (set by its launcher)
  int l2_prefetch;              // proto_head_tma.cu: prefetch every box into L2 one task ahead (AFSL_HEAD_L2PF=0 disables)
};

// proto_head_warp.cu: launches the warp-per-episode kernels when the shape fits them; *handled says whether it did
int launch_head_warp(const HeadParams& p, bool bwd, cudaStream_t stream, const char* name, bool* handled);
// proto_head_wide.cu: forward launches of many-way episodes (W >= 8, D in {128,256}); same contract
int launch_head_wide(const HeadParams& p, bool bwd, cudaStream_t stream, const char* name, bool* handled);
// proto_head_mma.cu: forward launches of fixed-size many-way tasks (8 <= W <= 24, 25 < Nq <= 128, D in {64,128,256}) with the
// query x prototype contraction on tcgen05 tensor cores in 3-pass split TF32; same contract
int launch_head_mma(const HeadParams& p, bool bwd, cudaStream_t stream, const char* name, bool* handled);
// proto_head_tma.cu: the same shapes with every byte arriving through a TMA-filled shared-memory ring (the default)
int launch_head_tma(const HeadParams& p, bool bwd, cudaStream_t stream, const char* name, bool* handled);

}  // namespace afsl
