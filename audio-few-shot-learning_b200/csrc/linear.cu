// Linear layers of the projection head (ProjectionHead.fc1 / fc2, models/main_modules.py:231-255:
// normalize(fc2(relu(fc1(x))))) as libafsl kernels, forward and backward, fp32 on the CUDA cores.
//
//   forward   y[M,N]  = x[M,K] . w[N,K]^T + b[N]            (optionally ReLU)
//   backward  g       = dy (.) [y > 0]                        (ReLU mask from the forward output, applied on load)
//             dx[M,K] = g[M,N] . w[N,K]
//             dw[N,K] = g[M,N]^T . x[M,K]                     (contraction over the rows: split over M, fixed-order reduce)
//             db[N]   = column sums of g                      (the same split / reduce)
//
// One register-tiled SGEMM (64 x 64 x 16 tiles, 256 threads, 4 x 4 outputs per thread, operands staged in shared memory,
// accumulation in ascending k with fmaf) serves the three products through operand strides.  The projection runs on
// M = E (Nq + W) rows with shared 256 x 512 x 256 weights; at the training step's size (960 rows) it is launch-bound, at
// E = 16384 it is fp32-pipe bound.  (A tcgen05 split-TF32 version needs MN-major operand descriptors for the two
// transposed products; the K-major building blocks are in tc_common.cuh.)
#include "afsl_common.cuh"

namespace afsl {
namespace {

constexpr int kBM = 64, kBN = 64, kBK = 16;
constexpr int kGemmThreads = 256;

// C[m, n] = sum_k A(m, k) B(k, n): A(m, k) = a[m * a_rs + k * a_cs], B(k, n) = b[k * b_rs + n * b_cs].
// mask_a (optional, same indexing as A): A(m, k) is taken as 0 where mask_a <= 0 (ReLU gate); mask_c the same for C.
// gridDim.z splits the k range; split z writes its partial tile to c + z * c_split.
struct GemmArgs {
  const float* a; const float* b; float* c;
  const float* bias;       // [n] added in the epilogue (split 0 only) or null
  const float* mask_a;     // gate on A or null
  const float* mask_c;     // gate on C (same layout as C) or null
  long long a_rs, a_cs, b_rs, b_cs, c_rs;
  long long c_split;       // elements between split outputs
  int M, N, K;
  int k_per_split;
  int relu;                // epilogue ReLU
};

__global__ void __launch_bounds__(kGemmThreads) sgemm_kernel(const GemmArgs g) {
  __shared__ float As[kBK][kBM + 4];
  __shared__ float Bs[kBK][kBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;                 // 16 x 16 threads, 4 x 4 outputs each
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const int k_lo = blockIdx.z * g.k_per_split;
  const int k_hi = min(g.K, k_lo + g.k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader mapping: the fastest-varying thread index follows the operand's contiguous dimension
  const bool a_k_contig = g.a_cs == 1;                    // A row-major over k
  const bool b_n_contig = g.b_cs == 1;
  for (int k0 = k_lo; k0 < k_hi; k0 += kBK) {
#pragma unroll
    for (int it = 0; it < (kBM * kBK) / kGemmThreads; ++it) {
      const int idx = tid + it * kGemmThreads;
      const int kk = a_k_contig ? (idx & (kBK - 1)) : (idx / kBM);
      const int mm = a_k_contig ? (idx / kBK) : (idx & (kBM - 1));
      const int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < g.M && k < k_hi) {
        const long long off = (long long)m * g.a_rs + (long long)k * g.a_cs;
        v = g.a[off];
        if (g.mask_a && !(g.mask_a[off] > 0.f)) v = 0.f;
      }
      As[kk][mm] = v;
    }
#pragma unroll
    for (int it = 0; it < (kBN * kBK) / kGemmThreads; ++it) {
      const int idx = tid + it * kGemmThreads;
      const int kk = b_n_contig ? (idx / kBN) : (idx & (kBK - 1));
      const int nn = b_n_contig ? (idx & (kBN - 1)) : (idx / kBK);
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < g.N && k < k_hi) ? g.b[(long long)k * g.b_rs + (long long)n * g.b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* c = g.c + (long long)blockIdx.z * g.c_split;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias && blockIdx.z == 0) v += g.bias[n];
      if (g.relu) v = fmaxf(v, 0.f);
      const long long off = (long long)m * g.c_rs + n;
      if (g.mask_c && !(g.mask_c[off] > 0.f)) v = 0.f;
      c[off] = v;
    }
  }
}

// out[i] = sum over splits (ascending) of part[s * stride + i]
__global__ void reduce_splits_kernel(const float* __restrict__ part, float* __restrict__ out, long long n, long long stride, int splits) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(long long)z * stride + i];
  out[i] = s;
}

// partial column sums of g = dy (.) [y > 0]: block (z = row split, x = 32-column group); rows in ascending order per thread
// group, then a fixed-order combine of the 8 row lanes
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dy, const float* __restrict__ mask, float* __restrict__ part,
                                                     int M, int N, int rows_per_split) {
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  const int r_lo = blockIdx.z * rows_per_split, r_hi = min(M, r_lo + rows_per_split);
  float s = 0.f;
  if (col < N)
    for (int r = r_lo + rl; r < r_hi; r += 8) {
      const long long off = (long long)r * N + col;
      float v = dy[off];
      if (mask && !(mask[off] > 0.f)) v = 0.f;
      s += v;
    }
  red[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x & 31];
    part[(long long)blockIdx.z * N + col] = t;
  }
}

int launch_gemm(const GemmArgs& g, int splits, cudaStream_t stream, const char* name) {
  dim3 grid((g.N + kBN - 1) / kBN, (g.M + kBM - 1) / kBM, splits);
  sgemm_kernel<<<grid, kGemmThreads, 0, stream>>>(g);
  AFSL_CHECK_LAUNCH(name);
  return AFSL_OK;
}

int row_splits(int M) {
  int s = (M + 4095) / 4096;          // ~4096 rows per split: enough CTAs at large M, one split at the training step's size
  return s < 1 ? 1 : (s > 64 ? 64 : s);
}

}  // namespace
}  // namespace afsl

using namespace afsl;

extern "C" int afsl_linear_fwd_f32(const float* x, const float* w, const float* bias, float* y, int M, int N, int K, int relu,
                                    void* stream) {
  AFSL_REQUIRE(x && w && y, "afsl_linear_fwd_f32: null pointer");
  AFSL_REQUIRE(M >= 0 && N > 0 && K > 0, "afsl_linear_fwd_f32: bad sizes M=%d N=%d K=%d", M, N, K);
  if (M == 0) return AFSL_OK;
  GemmArgs g{};
  g.a = x; g.a_rs = K; g.a_cs = 1;               // A(m, k) = x[m, k]
  g.b = w; g.b_rs = 1; g.b_cs = K;               // B(k, n) = w[n, k]
  g.c = y; g.c_rs = N; g.bias = bias;
  g.M = M; g.N = N; g.K = K; g.k_per_split = K; g.relu = relu;
  return launch_gemm(g, 1, (cudaStream_t)stream, "afsl_linear_fwd_f32");
}

extern "C" long long afsl_linear_bwd_workspace_floats(int M, int N, int K) {
  const long long s = row_splits(M);
  return s * ((long long)N * K + N);
}

extern "C" int afsl_linear_bwd_f32(const float* x, const float* w, const float* y_relu, const float* d_y, float* d_x, float* d_w,
                                    float* d_bias, float* workspace, int M, int N, int K, void* stream) {
  AFSL_REQUIRE(x && w && d_y && workspace, "afsl_linear_bwd_f32: null pointer");
  AFSL_REQUIRE(M > 0 && N > 0 && K > 0, "afsl_linear_bwd_f32: bad sizes M=%d N=%d K=%d", M, N, K);
  cudaStream_t st = (cudaStream_t)stream;
  if (d_x) {                                       // dx[M,K] = g[M,N] . w[N,K]
    GemmArgs g{};
    g.a = d_y; g.mask_a = y_relu; g.a_rs = N; g.a_cs = 1;
    g.b = w; g.b_rs = K; g.b_cs = 1;
    g.c = d_x; g.c_rs = K;
    g.M = M; g.N = K; g.K = N; g.k_per_split = N;
    if (int rc = launch_gemm(g, 1, st, "afsl_linear_bwd_f32")) return rc;
  }
  const int splits = row_splits(M);
  const int rows_per_split = (M + splits - 1) / splits;
  if (d_w) {                                       // dw[N,K] = g^T . x, rows split over gridDim.z
    GemmArgs g{};
    g.a = d_y; g.mask_a = y_relu; g.a_rs = 1; g.a_cs = N;          // A(n, m) = g[m, n]
    g.b = x; g.b_rs = K; g.b_cs = 1;                                // B(m, k) = x[m, k]
    g.c = splits == 1 ? d_w : workspace; g.c_rs = K; g.c_split = (long long)N * K;
    g.M = N; g.N = K; g.K = M; g.k_per_split = rows_per_split;
    if (int rc = launch_gemm(g, splits, st, "afsl_linear_bwd_f32")) return rc;
    if (splits > 1) {
      const long long n = (long long)N * K;
      reduce_splits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(workspace, d_w, n, n, splits);
      AFSL_CHECK_LAUNCH("afsl_linear_bwd_f32");
    }
  }
  if (d_bias) {
    float* part = workspace + (long long)splits * N * K;
    dim3 grid((N + 31) / 32, 1, splits);
    colsum_kernel<<<grid, 256, 0, st>>>(d_y, y_relu, splits == 1 ? d_bias : part, M, N, rows_per_split);
    AFSL_CHECK_LAUNCH("afsl_linear_bwd_f32");
    if (splits > 1) {
      reduce_splits_kernel<<<(N + 255) / 256, 256, 0, st>>>(part, d_bias, N, N, splits);
      AFSL_CHECK_LAUNCH("afsl_linear_bwd_f32");
    }
  }
  return AFSL_OK;
}
