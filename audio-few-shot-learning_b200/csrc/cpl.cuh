// Parameter block shared by the CPL kernels (cpl.cu: one CTA per episode, any shape; cpl_warp.cu: one warp
// per episode, 5-way at Dp <= 256).
#pragma once

#include "afsl_common.cuh"

namespace afsl {

struct CplParams {
  const float* protos;     // [E,W,D]
  const float* queries;    // [E,Nq,D]
  const int32_t* labels;   // [E,Nq]
  const uint32_t* keep;    // [E,Nq,words] or null
  float temperature;
  float* loss;             // [E]          (forward)
  const float* d_loss;     // [E]          (backward)
  float* d_protos;         // [E,W,D]
  float* d_queries;        // [E,Nq,D]
  int E, Nq, W, D;
};

// cpl_warp.cu: launches the warp-per-episode kernels when the shape fits them; *handled says whether it did
int launch_cpl_warp(const CplParams& p, bool bwd, cudaStream_t stream, const char* name, bool* handled);

}  // namespace afsl
