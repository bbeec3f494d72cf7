// Batched matrix transpose y[n][b][a] = x[n][a][b] (fp32): the layout switch between the channels-last activations the
// fused stage-1 kernel emits ([N][H*W][C]) and the NCHW tensors ([N][C][H*W]) on which cuDNN's fp32 (TF32 off)
// convolution kernels of stages 2-4 run ~25 % faster on sm_100 than their channels-last ones (tools/conv_fp32_probe.py:
// 57.7 against 77.0 ms per 6400-sample step, forward + backward); torch's own copy takes 3.7 ms for the stage-1 output
// (3.6 GB), this kernel moves the same bytes near the HBM rate.  Part of the encoder wrapper, not of the head.
//
// 64 x 64 tiles through shared memory (row stride 65 floats), 128-bit global accesses on both sides when A and B are
// multiples of 4; with B = 64 (a channels-last pixel row) a tile is one contiguous 16 KB run of the input.
#include "afsl_common.cuh"

namespace afsl {
namespace {

constexpr int kT = 64;
constexpr int kTrThreads = 256;

template <bool kVec>
__global__ void __launch_bounds__(kTrThreads) transpose_kernel(const float* __restrict__ x, float* __restrict__ y, int A, int B,
                                                               int tiles_a, int tiles_b) {
  __shared__ float tile[kT][kT + 1];
  const long long blk = blockIdx.x;
  const int tb = (int)(blk % tiles_b);
  const int ta = (int)((blk / tiles_b) % tiles_a);
  const long long n = blk / ((long long)tiles_a * tiles_b);
  const int a0 = ta * kT, b0 = tb * kT;
  const float* xs = x + n * (long long)A * B;
  float* ys = y + n * (long long)A * B;
  const int t = threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = t + kTrThreads * k, r = idx >> 4, c4 = (idx & 15) * 4;
    const int a = a0 + r, b = b0 + c4;
    if (a < A) {
      if (kVec && b + 3 < B) {
        const float4 v = ldg_stream(reinterpret_cast<const float4*>(xs + (long long)a * B + b));
        tile[r][c4] = v.x; tile[r][c4 + 1] = v.y; tile[r][c4 + 2] = v.z; tile[r][c4 + 3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (b + j < B) tile[r][c4 + j] = xs[(long long)a * B + b + j];
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = t + kTrThreads * k, r = idx >> 4, c4 = (idx & 15) * 4;
    const int b = b0 + r, a = a0 + c4;
    if (b < B) {
      if (kVec && a + 3 < A) {
        stg_stream(reinterpret_cast<float4*>(ys + (long long)b * A + a),
                   make_float4(tile[c4][r], tile[c4 + 1][r], tile[c4 + 2][r], tile[c4 + 3][r]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (a + j < A) ys[(long long)b * A + a + j] = tile[c4 + j][r];
      }
    }
  }
}

}  // namespace
}  // namespace afsl

extern "C" int afsl_transpose_f32(const float* x, float* y, int n, int A, int B, void* stream) {
  AFSL_REQUIRE(x && y, "afsl_transpose_f32: null pointer");
  AFSL_REQUIRE(n >= 0 && A > 0 && B > 0, "afsl_transpose_f32: n=%d A=%d B=%d", n, A, B);
  if (n == 0) return AFSL_OK;
  const int tiles_a = (A + afsl::kT - 1) / afsl::kT, tiles_b = (B + afsl::kT - 1) / afsl::kT;
  const long long blocks = (long long)n * tiles_a * tiles_b;
  AFSL_REQUIRE(blocks < (1ll << 31), "afsl_transpose_f32: %lld tiles exceed the grid limit", blocks);
  const bool vec = A % 4 == 0 && B % 4 == 0 && afsl::aligned16(x) && afsl::aligned16(y);
  if (vec)
    afsl::transpose_kernel<true><<<(unsigned)blocks, afsl::kTrThreads, 0, (cudaStream_t)stream>>>(x, y, A, B, tiles_a, tiles_b);
  else
    afsl::transpose_kernel<false><<<(unsigned)blocks, afsl::kTrThreads, 0, (cudaStream_t)stream>>>(x, y, A, B, tiles_a, tiles_b);
  AFSL_CHECK_LAUNCH("afsl_transpose_f32");
  return AFSL_OK;
}
