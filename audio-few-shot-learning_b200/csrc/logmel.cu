// Waveform front end: 16 kHz waveform -> log-mel spectrogram in dB -> global z-normalisation, ONE kernel.
//
// Reference: datasets/batch_creation.py:138-143 and :215-218 (mel_spec_function_gpu) with the MelSpectrogram of
// src/train_test.py:123-129: torchaudio MelSpectrogram(sample_rate 16000, n_fft 1024, hop 512, n_mels 128, power 2) =
// STFT (periodic Hann window of n_fft samples, centre frames, reflect padding, one-sided) -> |X|^2 -> [n_freq x n_mels]
// triangular filterbank; then 20/2 * log10(mel + float32 eps), (x - mean) / std with the dataset's statistics, and a
// channel axis: [N, n_samples] -> [N, 1, 128, T], T = n_samples / hop + 1 (80 000 samples -> 157 frames).
//
// The eager chain is pad + unfold + cuFFT + abs^2 + a cuBLAS matmul + three elementwise kernels, ~8 launches and four
// intermediate tensors; here a CTA takes 8 consecutive frames of one clip: reflect-indexed window load, radix-2 FFT of the
// 1024 real samples in shared memory (twiddles from sincospi, computed once per CTA), power spectrum, the mel filters in
// their sparse band form (the caller passes the reference transform's own window and filterbank, so both are the
// reference's numbers), dB, normalisation, and a staged [128 x 8] output tile so that the [N,1,128,T] rows are written in
// 32-byte runs.  The output is the tensor the SpecAugment kernel reads.
#include "afsl_common.cuh"

namespace afsl {
namespace {

constexpr int kNfft = 1024, kHalf = kNfft / 2, kBins = kHalf + 1, kMels = 128;
constexpr int kFramesPerCta = 8;
constexpr int kMelThreads = 256;

struct MelParams {
  const float* wave;      // [N, L]
  const float* window;    // [1024]
  const int32_t* fb_start;  // [128] first frequency bin of each mel filter
  const int32_t* fb_len;    // [128] number of bins
  const int32_t* fb_off;    // [128] offset into fb_w
  const float* fb_w;        // packed filter weights
  float* out;             // [N, 1, 128, T]
  int N, L, T, hop;
  float eps, mean, std;
};

__device__ __forceinline__ int bitrev10(int i) { return (int)(__brev((unsigned)i) >> 22); }

__global__ void __launch_bounds__(kMelThreads) logmel_kernel(const MelParams p) {
  __shared__ float tw_c[kHalf], tw_s[kHalf];     // e^{-2 pi i k / 1024}
  __shared__ float win[kNfft];
  __shared__ float re[kNfft], im[kNfft];
  __shared__ float power[kBins + 3];
  __shared__ float tile[kMels][kFramesPerCta + 1];
  const int tid = threadIdx.x;
  const int n = blockIdx.y, t0 = blockIdx.x * kFramesPerCta;
  for (int k = tid; k < kHalf; k += kMelThreads) {
    float s, c;
    sincospif((float)k / (float)kHalf, &s, &c);   // angle = 2 pi k / 1024
    tw_c[k] = c;
    tw_s[k] = -s;
  }
  for (int i = tid; i < kNfft; i += kMelThreads) win[i] = p.window[i];
  const float* x = p.wave + (size_t)n * p.L;
  const int frames = min(kFramesPerCta, p.T - t0);
  for (int f = 0; f < frames; ++f) {
    const int t = t0 + f;
    __syncthreads();
    // frame t = samples [t hop - 512, t hop + 512) of the reflect-padded clip, windowed, stored in bit-reversed order
    for (int i = tid; i < kNfft; i += kMelThreads) {
      int idx = t * p.hop - kHalf + i;
      if (idx < 0) idx = -idx;
      if (idx >= p.L) idx = 2 * (p.L - 1) - idx;
      const int j = bitrev10(i);
      re[j] = x[idx] * win[i];
      im[j] = 0.f;
    }
    __syncthreads();
    // 10 radix-2 decimation-in-time stages, 512 butterflies each (two per thread)
#pragma unroll 1
    for (int s = 0; s < 10; ++s) {
      const int half = 1 << s;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int b = tid + u * kMelThreads;
        const int j = b & (half - 1);
        const int lo = ((b >> s) << (s + 1)) + j, hi = lo + half;
        const int k = j << (9 - s);
        const float wr = tw_c[k], wi = tw_s[k];
        const float xr = re[hi], xi = im[hi];
        const float tr = fmaf(wr, xr, -wi * xi), ti = fmaf(wr, xi, wi * xr);
        const float ar = re[lo], ai = im[lo];
        re[lo] = ar + tr; im[lo] = ai + ti;
        re[hi] = ar - tr; im[hi] = ai - ti;
      }
      __syncthreads();
    }
    for (int k = tid; k < kBins; k += kMelThreads) power[k] = fmaf(re[k], re[k], im[k] * im[k]);
    __syncthreads();
    if (tid < kMels) {
      const int st = p.fb_start[tid], len = p.fb_len[tid];
      const float* w = p.fb_w + p.fb_off[tid];
      float acc = 0.f;
      for (int k = 0; k < len; ++k) acc = fmaf(power[st + k], w[k], acc);
      const float db = 10.f * log10f(acc + p.eps);
      tile[tid][f] = __fdiv_rn(db - p.mean, p.std);
    }
  }
  __syncthreads();
  // [128 x frames] tile -> out[n, 0, m, t0 ..]: two threads per mel row, 4 frames each
  {
    const int m = tid >> 1, f0 = (tid & 1) * 4;
    float* dst = p.out + ((size_t)n * kMels + m) * p.T + t0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (f0 + k < frames) dst[f0 + k] = tile[m][f0 + k];
  }
}

}  // namespace
}  // namespace afsl

using namespace afsl;

extern "C" int afsl_logmel_f32(const float* wave, const float* window, const int32_t* fb_start, const int32_t* fb_len,
                                const int32_t* fb_off, const float* fb_w, float* out, int N, int L, int T, int n_fft, int hop,
                                int n_mels, float eps, float mean, float std, void* stream) {
  AFSL_REQUIRE(wave && window && fb_start && fb_len && fb_off && fb_w && out, "afsl_logmel_f32: null pointer");
  AFSL_REQUIRE(n_fft == kNfft && n_mels == kMels, "afsl_logmel_f32: supports n_fft=1024, n_mels=128 (got %d, %d)", n_fft, n_mels);
  AFSL_REQUIRE(hop > 0 && L > kHalf && T == L / hop + 1, "afsl_logmel_f32: bad sizes L=%d hop=%d T=%d", L, hop, T);
  AFSL_REQUIRE(N >= 0 && N <= 65535, "afsl_logmel_f32: N=%d clips per launch (max 65535)", N);
  if (N == 0) return AFSL_OK;
  MelParams p{wave, window, fb_start, fb_len, fb_off, fb_w, out, N, L, T, hop, eps, mean, std};
  dim3 grid((T + kFramesPerCta - 1) / kFramesPerCta, N);
  logmel_kernel<<<grid, kMelThreads, 0, (cudaStream_t)stream>>>(p);
  AFSL_CHECK_LAUNCH("afsl_logmel_f32");
  return AFSL_OK;
}
