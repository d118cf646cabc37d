// lidfe_stft_fbank.cuh -- the wav2vec-exp FBank variant (row f4, ref: wav2vec-exp/s3prl_model.py:174-204): n_fft = 2 * hop
// (640 / 320 in the reference's model), periodic Hann, no centring, HTK mel, 10 log10, ONE mean / std per utterance.
// 640 = 5 * 2^7 is no size of the fused 512-point FFT kernel; the transform runs as a GEMM instead -- frames x windowed DFT
// basis -- on the windowed-GEMM machinery of the resampler (lidfe_resample_tc.cuh: tcgen05, 3 x TF32): "phase" 2 b is
// w[k] cos(2 pi b k / n_fft), phase 2 b + 1 is -w[k] sin(...), hop = n_fft / 2, no left padding.  The two kernels here turn its
// output rows (re, im interleaved) into power -> mel -> dB and accumulate / apply the per-utterance scalar statistics.
#pragma once
#include "lidfe_kernels.cuh"

namespace lidfe {

constexpr int kSfWarps = 8;          // frames per CTA of stft_mel_db_kernel
constexpr int kSfMaxBins = 1025;     // n_fft <= 2048

struct StftMelParams {
  const float* g;                    // [sum frames][nw] DFT rows: (re, im) of bin b at 2 b, 2 b + 1
  const long long* g_off;            // [B] first element of utterance i in g
  const long long* frames;           // [B]
  float* out;                        // [rows][n_mels]
  const long long* out_row;          // [B] first output row of utterance i
  const float* melT;                 // [n_mels][n_bins] filter bank, transposed (ta: functional/functional.py:492-588)
  const int* mel_lo;                 // [n_mels] first / last bin with a non-zero weight
  const int* mel_hi;
  double* stats;                     // [B][2] sum, sum of squares of the utterance's dB values (zeroed by the caller)
  int nw, n_bins, n_mels;
  float amin;                        // 1e-10
};

__global__ void __launch_bounds__(kSfWarps * 32) stft_mel_db_kernel(const __grid_constant__ StftMelParams P) {
  __shared__ float s_pow[kSfWarps][kSfMaxBins + 3];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long T = P.frames[b];
  const long long t = static_cast<long long>(blockIdx.x) * kSfWarps + warp;
  if (t >= T) return;
  const float2* row = reinterpret_cast<const float2*>(P.g + P.g_off[b] + t * P.nw);      // nw is even, g_off a multiple of 2
  float* pw = s_pow[warp];
  for (int k = lane; k < P.n_bins; k += 32) {
    const float2 z = __ldg(row + k);
    pw[k] = fmaf(z.x, z.x, z.y * z.y);                  // |X[k]|^2       ta: functional/functional.py:135-141 (power=2)
  }
  __syncwarp();
  double s1 = 0.0, s2 = 0.0;
  float* o = P.out + (P.out_row[b] + t) * P.n_mels;
  for (int m = lane; m < P.n_mels; m += 32) {
    const float* w = P.melT + static_cast<long long>(m) * P.n_bins;
    float acc = 0.f;
    for (int k = P.mel_lo[m]; k <= P.mel_hi[m]; ++k) acc = fmaf(pw[k], __ldg(w + k), acc);
    const float db = 10.0f * log10f(fmaxf(acc, P.amin));   // amplitude_to_DB(multiplier 10, amin, db_multiplier 0, no top_db)
    o[m] = db;
    s1 += static_cast<double>(db);
    s2 += static_cast<double>(db) * static_cast<double>(db);
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  }
  if (lane == 0) {
    atomicAdd(&P.stats[2 * b], s1);
    atomicAdd(&P.stats[2 * b + 1], s2);
  }
}

struct ScalarNormParams {
  float* out;
  const long long* out_row;
  const long long* frames;
  const double* stats;
  int n_mels;
};

// (x - mean) / (std + 1e-9), one mean and one unbiased std over all frames x n_mels values of the utterance
// (ref: wav2vec-exp/s3prl_model.py:203-204 torch.std_mean(spec))
__global__ void __launch_bounds__(256) scalar_norm_kernel(const __grid_constant__ ScalarNormParams P) {
  const int b = blockIdx.y;
  const long long n = P.frames[b] * P.n_mels;
  const double s = P.stats[2 * b], q = P.stats[2 * b + 1];
  const double mu = s / static_cast<double>(n);
  double var = (q - s * mu) / static_cast<double>(n - 1);          // n == 1 -> NaN, as torch.std of one value
  var = var > 0.0 ? var : (var == var ? 0.0 : var);
  const float mean = static_cast<float>(mu);
  const float lo = static_cast<float>(mu - static_cast<double>(mean));
  const float inv = static_cast<float>(1.0 / (sqrt(var) + 1e-9));
  float* o = P.out + P.out_row[b] * P.n_mels;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    o[i] = ((o[i] - mean) - lo) * inv;
}

}  // namespace lidfe
