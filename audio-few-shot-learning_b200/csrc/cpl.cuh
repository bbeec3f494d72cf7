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
  float* sim_out;          // [E,W,Nq] C = cos / T and
  float* qinv_out;         // [E,Nq] 1 / |q| (negative: clamped norm) saved by the forward for the backward, or null
  const float* sim_in;     // the same two, given to the backward (warp family only): no recomputation walk over Q
  const float* qinv_in;
  int E, Nq, W, D;
};

// cpl_warp.cu: launches the warp-per-episode kernels when the shape fits them; *handled says whether it did
int launch_cpl_warp(const CplParams& p, bool bwd, cudaStream_t stream, const char* name, bool* handled);
// whether the warp family takes this shape (the saved-similarity entry points exist for it only)
bool cpl_warp_supported(int Nq, int W, int D);

}  // namespace afsl
