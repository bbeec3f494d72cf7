// Small per-(group, channel) "glue" kernels that replace chains of tiny eager ops around the encoder kernels
// (each chain was 10-40 launches of 2-5 us kernels per call; ~600 of a training step's 650 launches were such).
//
//   bn_running_update : the G momentum updates of BatchNorm's running statistics that G separate module calls
//                       would have made (models/main_modules.py:18-23: one encoder call per 25-sample set),
//                       applied in group order, one thread per channel.
//   stage1_finalize   : per group, the 54 input moments -> mean / variance / rstd of the 64 conv-1 channels and
//                       the folded affine (a, b), in double (stage1.cu header: mean_c = w_c.S / m,
//                       E[u^2] = w_c^T R w_c / m).
//   stage1_dw         : conv-1 weight / BatchNorm gradients from the backward partials and the moments:
//                       dW_c[k] = sum_g a_gc [ T_gck - m1 S_gk - m2 rstd (sum_l w_cl R_glk - mean S_gk) ].
#include "afsl_common.cuh"

namespace afsl {
namespace {

__global__ void bn_running_update_kernel(const float* __restrict__ mean, const float* __restrict__ var_biased,
                                         const float* __restrict__ shift, float* running_mean, float* running_var,
                                         long long* num_batches, float momentum, float unbias, int G, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    float rm = running_mean[c], rv = running_var[c];
    const float sh = shift ? shift[c] : 0.f;
#pragma unroll 8
    for (int g = 0; g < G; ++g) {          // exactly the sequence of updates of G separate calls (loads batched by the unroll)
      rm = (1.f - momentum) * rm + momentum * (mean[(size_t)g * C + c] + sh);
      rv = (1.f - momentum) * rv + momentum * (var_biased[(size_t)g * C + c] * unbias);
    }
    running_mean[c] = rm;
    running_var[c] = rv;
  }
  if (num_batches && blockIdx.x == 0 && threadIdx.x == 0) *num_batches += G;
}

constexpr int kC1 = 64;     // conv-1 output channels

// grid = G, block = 64: moments [G, parts, 54] -> S [G,9], R [G,9,9] (double), mean_u / var / rstd / a / b [G,64]
__global__ void stage1_finalize_kernel(const double* __restrict__ moments, int parts, const float* __restrict__ w9,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                       double m, double* S, double* R, float* mean_u, float* var, float* rstd, float* a,
                                       float* b) {
  __shared__ double s54[54];
  __shared__ double r81[81];
  const int g = blockIdx.x, c = threadIdx.x;
  if (c < 54) {
    double t = 0.0;
    for (int p = 0; p < parts; ++p) t += moments[((size_t)g * parts + p) * 54 + c];
    s54[c] = t;
  }
  __syncthreads();
  // upper triangle (row by row after the 9 sums) -> symmetric 9x9
  for (int i = c; i < 81; i += blockDim.x) {
    int k = i / 9, l = i - k * 9;
    if (l < k) { const int t = k; k = l; l = t; }
    const int off = 9 + k * 9 - k * (k - 1) / 2 + (l - k);
    r81[i] = s54[off];
  }
  __syncthreads();
  if (c < 9) S[(size_t)g * 9 + c] = s54[c];
  for (int i = c; i < 81; i += blockDim.x) R[(size_t)g * 81 + i] = r81[i];
  if (c < kC1) {
    double w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w[k] = (double)w9[c * 9 + k];
    double mu = 0.0, e2 = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      mu += w[k] * s54[k];
      double row = 0.0;
#pragma unroll
      for (int l = 0; l < 9; ++l) row += r81[k * 9 + l] * w[l];
      e2 += w[k] * row;
    }
    mu /= m;
    e2 /= m;
    double v = e2 - mu * mu;
    if (v < 0.0) v = 0.0;
    const double rs = 1.0 / sqrt(v + (double)eps);
    const double aa = (double)gamma[c] * rs;
    const size_t o = (size_t)g * kC1 + c;
    mean_u[o] = (float)mu;
    var[o] = (float)v;
    rstd[o] = (float)rs;
    a[o] = (float)aa;
    b[o] = (float)((double)beta[c] - mu * aa);
  }
}

// grid = 64 channels, block = (11, 32): partial [G, parts, 64, 11] -> d_w [64,9], d_gamma [64], d_beta [64] (+ d_bias when
// the statistics are running ones).  threadIdx.x = k (k < 9: tap; 9: gamma; 10: beta / bias), threadIdx.y = slice of the
// groups (g = slice, slice + 32, ...); the 32 slice sums are added in slice order (deterministic).
// per_group: batch statistics (a, mean_u, rstd are [G,64]); else eval ([64]).
constexpr int kDwSlices = 32;

__global__ void __launch_bounds__(11 * kDwSlices) stage1_dw_kernel(
    const float* __restrict__ partial, int parts, int G, const double* __restrict__ S, const double* __restrict__ R,
    const float* __restrict__ w9, const float* __restrict__ a, const float* __restrict__ mean_u,
    const float* __restrict__ rstd, double m, int per_group, float* d_w, float* d_gamma, float* d_beta, float* d_bias) {
  __shared__ double red[kDwSlices][11], red_bias[kDwSlices];
  const int c = blockIdx.x, k = threadIdx.x, slice = threadIdx.y;
  double acc = 0.0, acc_bias = 0.0;
  for (int g = slice; g < G; g += kDwSlices) {
    double s1 = 0.0, s2 = 0.0, t = 0.0;
    for (int p = 0; p < parts; ++p) {
      const float* src = partial + (((size_t)g * parts + p) * kC1 + c) * 11;
      s1 += (double)src[0];
      s2 += (double)src[1];
      if (k < 9) t += (double)src[2 + k];
    }
    const int idx = per_group ? g * kC1 + c : c;
    const double ag = (double)a[idx];
    if (k < 9) {
      if (per_group) {
        const double m1 = s1 / m, m2 = s2 / m;
        double wr = 0.0;
#pragma unroll
        for (int l = 0; l < 9; ++l) wr += (double)w9[c * 9 + l] * R[(size_t)g * 81 + l * 9 + k];
        const double sk = S[(size_t)g * 9 + k];
        const double corr = (double)rstd[idx] * (wr - (double)mean_u[idx] * sk);
        acc += ag * (t - m1 * sk - m2 * corr);
      } else {
        acc += ag * t;
      }
    } else if (k == 9) {
      acc += s2;
    } else {
      acc += s1;
      acc_bias += ag * s1;
    }
  }
  red[slice][k] = acc;
  if (k == 10) red_bias[slice] = acc_bias;
  __syncthreads();
  if (slice != 0) return;
  double tot = 0.0, tot_bias = 0.0;
  for (int q = 0; q < kDwSlices; ++q) tot += red[q][k];
  if (k < 9) d_w[c * 9 + k] = (float)tot;
  else if (k == 9) d_gamma[c] = (float)tot;
  else {
    for (int q = 0; q < kDwSlices; ++q) tot_bias += red_bias[q];
    d_beta[c] = (float)tot;
    if (d_bias) d_bias[c] = per_group ? 0.f : (float)tot_bias;   // batch statistics remove any per-channel constant
  }
}

}  // namespace
}  // namespace afsl

extern "C" int afsl_bn_running_update_f32(const float* mean, const float* var_biased, const float* shift, float* running_mean,
                                           float* running_var, long long* num_batches_tracked, float momentum, float unbias,
                                           int G, int C, void* stream) {
  using namespace afsl;
  AFSL_REQUIRE(mean && var_biased && running_mean && running_var, "afsl_bn_running_update_f32: null pointer");
  AFSL_REQUIRE(G > 0 && C > 0, "afsl_bn_running_update_f32: bad sizes G=%d C=%d", G, C);
  bn_running_update_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mean, var_biased, shift, running_mean, running_var,
                                                                            num_batches_tracked, momentum, unbias, G, C);
  AFSL_CHECK_LAUNCH("afsl_bn_running_update_f32");
  return AFSL_OK;
}

extern "C" int afsl_stage1_finalize_f64(const double* moments, int parts, const float* weight, const float* gamma,
                                         const float* beta, float eps, double count, double* S, double* R, float* mean_u,
                                         float* var, float* rstd, float* a, float* b, int G, void* stream) {
  using namespace afsl;
  AFSL_REQUIRE(moments && weight && gamma && beta && S && R && mean_u && var && rstd && a && b && parts > 0 && G > 0,
               "afsl_stage1_finalize_f64: null pointer / sizes");
  stage1_finalize_kernel<<<G, 64, 0, (cudaStream_t)stream>>>(moments, parts, weight, gamma, beta, eps, count, S, R, mean_u, var,
                                                            rstd, a, b);
  AFSL_CHECK_LAUNCH("afsl_stage1_finalize_f64");
  return AFSL_OK;
}

extern "C" int afsl_stage1_dw_f32(const float* partial, int parts, int G, const double* S, const double* R, const float* weight,
                                   const float* a, const float* mean_u, const float* rstd, double count, int per_group,
                                   float* d_w, float* d_gamma, float* d_beta, float* d_bias, void* stream) {
  using namespace afsl;
  AFSL_REQUIRE(partial && weight && a && d_w && d_gamma && d_beta && parts > 0 && G > 0, "afsl_stage1_dw_f32: null pointer / sizes");
  AFSL_REQUIRE(!per_group || (S && R && mean_u && rstd), "afsl_stage1_dw_f32: batch statistics need S, R, mean_u, rstd");
  stage1_dw_kernel<<<kC1, dim3(11, kDwSlices), 0, (cudaStream_t)stream>>>(partial, parts, G, S, R, weight, a, mean_u, rstd, count, per_group, d_w,
                                                        d_gamma, d_beta, d_bias);
  AFSL_CHECK_LAUNCH("afsl_stage1_dw_f32");
  return AFSL_OK;
}
