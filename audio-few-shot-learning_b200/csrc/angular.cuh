// Parameter block shared by the angular-loss kernels (angular.cu: one CTA per episode, any shape;
// angular_warp.cu: one warp per episode, Dp = 64 and at most 32 pooled rows; angular_tc.cu: tensor cores, anchors branch).
#pragma once

#include "afsl_common.cuh"

namespace afsl {

struct AngParams {
  const float* protos;    // [E,W,D]
  const float* queries;   // [E,Nq,D]
  const int32_t* labels;  // [E,Nq]
  float miner_angle;      // radians
  float miner_tan;        // tan(miner_angle): atan(r) > angle  <=>  r > tan(angle) for angle in (-pi/2, pi/2)
  int miner_never;        // angle >= pi/2: atan never exceeds it, nothing is mined
  float t2;               // tan^2(alpha)
  int anchors, normalize_ref;
  float* loss;            // [E]
  const float* d_loss;    // [E]      (backward)
  float* d_protos;        // [E,W,D]
  float* d_queries;       // [E,Nq,D]
  int E, Nq, W, D;
  long long* dbg;         // angular_tc.cu timeline buffer (AFSL_ANGULAR_DBG), else null
};

// angular_warp.cu: launches the warp-per-episode kernel when the shape fits it; *handled says whether it did
int launch_angular_warp(const AngParams& p, bool bwd, cudaStream_t stream, const char* name, bool* handled);
// angular_tc.cu: the tensor-core kernel (anchors branch, Dp = 64, W <= 8, W + Nq <= 31), four episodes per tcgen05 tile
int launch_angular_tc(const AngParams& p, bool bwd, cudaStream_t stream, const char* name, bool* handled);

}  // namespace afsl
