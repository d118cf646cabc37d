// lidfe_kernels.cuh -- sm_100a kernels of the speech-lid front-end.
//
// Path (ref: lid/audio_processor.py:41-69 -> ta: compliance/kaldi.py:154-217,591-633):
//   frame (400 samples, hop 160) -> DC removal -> pre-emphasis (replicate-left) -> Povey window ->
//   zero-pad to 512 -> |rfft|^2 -> sparse triangular mel -> log(max(., eps)) [-> DCT + lifter] ->
//   [global CMVN] -> [SpecAugment zero-fill] -> out, plus per-utterance / global sum & sum-of-squares.
//
// Mapping: one persistent CTA of 8 warps loops over tiles of <= 32 consecutive frames of one utterance.
// The tile's 160*F+240 samples are staged into shared memory by ONE TMA bulk copy (cp.async.bulk +
// mbarrier), double buffered so the next tile's copy overlaps this tile's math.  A half-warp (16 lanes)
// owns one frame: the 512-point real FFT is a 256-point complex FFT of z[n] = x[2n] + i x[2n+1] done as
// 16 x 16 (two in-register radix-4x4 16-point DFTs per lane, one shared-memory transpose in between),
// followed by the real-FFT split which pairs lane t with lane 16-t through warp shuffles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lidfe {

constexpr int kFrameLen = 400;
constexpr int kFrameShift = 160;
constexpr int kFftLen = 512;
constexpr int kBins = kFftLen / 2 + 1;            // 257
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kFramesPerRound = kWarps * 2;       // one frame per half-warp
constexpr int kTileFrames = 32;
constexpr int kTileSamples = kFrameShift * kTileFrames + (kFrameLen - kFrameShift);   // 5360
constexpr int kTileSamplesPad = kTileSamples + 32;
constexpr int kMaxMels = 80;
constexpr int kBands = kMaxMels / 16;             // lane t owns mel bins t + 16*b
constexpr int kRowStride = 18;                    // float2 per transpose row (16 + 2 pad -> LDS.128 conflict free)
constexpr int kScratchFloats = 16 * kRowStride * 2;   // 576 floats per frame
constexpr int kLogmelOff = 272;                   // log-mel staging (MFCC) lives after the 257 power bins
constexpr int kMaxMasks = 8;

struct Tile {
  long long wav_off;    // first sample of the tile's first frame in the packed buffer
  long long out_row;    // output row of the tile's first frame (or first zero-fill row)
  int nframes;          // 1..kTileFrames; 0 -> zero-fill tile
  int utt;              // utterance index
  int t0;               // index of the first frame inside its utterance (time masks)
  int aux;              // zero-fill tiles: number of rows to clear.  else: 1 if TMA-eligible (16B aligned)
};

struct FbankParams {
  const void* wav;
  float* out;
  long long out_ld;
  const Tile* tiles;
  int n_tiles;
  // constant tables (device)
  const float* window;      // [512] zero padded
  const float2* tw1;        // [16][16]  W256^(K1*t)
  const float2* tw2;        // [8][16]   W512^(t+16i)
  const float* mel_w;       // [kBands][maxt][16]
  const int* mel_k0;        // [kBands*16]
  const float* dct;         // [n_mels][n_ceps]
  const float* lifter;      // [n_ceps] (ones when no liftering)
  int mel_maxt;             // taps stride per band
  int band_taps[kBands];    // max taps per band
  int n_mels, n_ceps, n_out;
  float preemph, log_floor, in_scale;
  int remove_dc;
  // epilogue
  const int* masks;         // [B][n_masks][4]
  int n_masks;
  int mode;                 // LIDFE_CMVN_*
  const double* stats_in;   // [2*n_out+1]
  double* stats_out;        // [2*n_out+1]
  double* utt_stats;        // [B][2][n_out]
};

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// (a.x + i a.y) * (w.x + i w.y)
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  return make_float2(fmaf(a.x, w.x, -(a.y * w.y)), fmaf(a.x, w.y, a.y * w.x));
}

// forward 4-point DFT, outputs in natural order
__device__ __forceinline__ void radix4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, t3 = a1 - a3;
  a0 = t0 + t2;
  a2 = t0 - t2;
  a1 = make_float2(t1.x + t3.y, t1.y - t3.x);
  a3 = make_float2(t1.x - t3.y, t1.y + t3.x);
}
// same with a3 == 0 on input (zero padding of the 400-sample frame to 512)
__device__ __forceinline__ void radix4_z3(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = a0 + a2, t1 = a0 - a2;
  float2 t = a1;
  a0 = t0 + t;
  a2 = t0 - t;
  a1 = make_float2(t1.x + t.y, t1.y - t.x);
  a3 = make_float2(t1.x - t.y, t1.y + t.x);
}

// forward 16-point DFT in registers: 4x4 radix-4.  In: v[n] natural.  Out: v[p] = X[rev4(p)],
// rev4(p) = (p >> 2) + 4 * (p & 3).  kTailZero: v[13], v[14], v[15] are known zeros.
template <bool kTailZero>
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  constexpr float kC1 = 0.92387953251128674f;   // cos(pi/8)
  constexpr float kS1 = 0.38268343236508977f;   // sin(pi/8)
  constexpr float kH = 0.70710678118654752f;    // sqrt(1/2)
  radix4(v[0], v[4], v[8], v[12]);
  if (kTailZero) {
    radix4_z3(v[1], v[5], v[9], v[13]);
    radix4_z3(v[2], v[6], v[10], v[14]);
    radix4_z3(v[3], v[7], v[11], v[15]);
  } else {
    radix4(v[1], v[5], v[9], v[13]);
    radix4(v[2], v[6], v[10], v[14]);
    radix4(v[3], v[7], v[11], v[15]);
  }
  // v[n2 + 4*k1] *= W16^(n2*k1)
  v[5] = cmul(v[5], make_float2(kC1, -kS1));                               // W^1
  v[9] = make_float2((v[9].x + v[9].y) * kH, (v[9].y - v[9].x) * kH);      // W^2
  v[13] = cmul(v[13], make_float2(kS1, -kC1));                             // W^3
  v[6] = make_float2((v[6].x + v[6].y) * kH, (v[6].y - v[6].x) * kH);      // W^2
  v[10] = make_float2(v[10].y, -v[10].x);                                  // W^4 = -i
  v[14] = make_float2((v[14].y - v[14].x) * kH, -(v[14].x + v[14].y) * kH);  // W^6
  v[7] = cmul(v[7], make_float2(kS1, -kC1));                               // W^3
  v[11] = make_float2((v[11].y - v[11].x) * kH, -(v[11].x + v[11].y) * kH);  // W^6
  v[15] = cmul(v[15], make_float2(-kC1, kS1));                             // W^9
  radix4(v[0], v[1], v[2], v[3]);
  radix4(v[4], v[5], v[6], v[7]);
  radix4(v[8], v[9], v[10], v[11]);
  radix4(v[12], v[13], v[14], v[15]);
}
__host__ __device__ constexpr int rev4(int p) { return (p >> 2) + 4 * (p & 3); }

template <typename TIn>
struct InTraits;
template <>
struct InTraits<float> {
  static __device__ __forceinline__ float2 ld2(const float* p, float) { return *reinterpret_cast<const float2*>(p); }
  static __device__ __forceinline__ float ld1(const float* p, float) { return *p; }
};
template <>
struct InTraits<short> {
  static __device__ __forceinline__ float2 ld2(const short* p, float s) {
    short2 v = *reinterpret_cast<const short2*>(p);
    return make_float2(static_cast<float>(v.x) * s, static_cast<float>(v.y) * s);
  }
  static __device__ __forceinline__ float ld1(const short* p, float s) { return static_cast<float>(*p) * s; }
};

// shared-memory carve-up (dynamic)
template <typename TIn>
struct SmemLayout {
  static constexpr int kInBytes = ((kTileSamplesPad * (int)sizeof(TIn)) + 127) / 128 * 128;
  static constexpr int off_in0 = 0;
  static constexpr int off_in1 = kInBytes;
  static constexpr int off_scratch = 2 * kInBytes;                                 // [kFramesPerRound][576] floats
  static constexpr int off_window = off_scratch + kFramesPerRound * kScratchFloats * 4;   // [512]
  static constexpr int off_tw1 = off_window + 512 * 4;                             // [16][16] float2
  static constexpr int off_tw2 = off_tw1 + 256 * 8;                                // [8][16] float2
  static constexpr int off_k0 = off_tw2 + 128 * 8;                                 // [80] int
  static constexpr int off_norm = off_k0 + kMaxMels * 4;                           // [2][80] float mean, inv_std
  static constexpr int off_stats = off_norm + 2 * kMaxMels * 4;                    // [2 bufs][2][80] float
  static constexpr int off_gstats = off_stats + 4 * kMaxMels * 4;                  // [2*80+1] double (8B aligned)
  static constexpr int off_masks = off_gstats + (2 * kMaxMels + 2) * 8;            // [kMaxMasks][4] int
  static constexpr int off_bar = off_masks + kMaxMasks * 16;                       // 2 mbarriers
  static constexpr int off_melw = off_bar + 16;                                    // [kBands][maxt][16] float, then dct, lifter
};

// ------------------------------------------------------------------------------------------------
// the fused front-end kernel
// ------------------------------------------------------------------------------------------------
template <typename TIn, bool kMfcc>
__global__ void __launch_bounds__(kThreads, 2) fbank_kernel(const __grid_constant__ FbankParams P) {
  using L = SmemLayout<TIn>;
  extern __shared__ __align__(128) unsigned char smem[];
  TIn* sm_in[2] = {reinterpret_cast<TIn*>(smem + L::off_in0), reinterpret_cast<TIn*>(smem + L::off_in1)};
  float* sm_scratch = reinterpret_cast<float*>(smem + L::off_scratch);
  float* sm_window = reinterpret_cast<float*>(smem + L::off_window);
  float2* sm_tw1 = reinterpret_cast<float2*>(smem + L::off_tw1);
  float2* sm_tw2 = reinterpret_cast<float2*>(smem + L::off_tw2);
  int* sm_k0 = reinterpret_cast<int*>(smem + L::off_k0);
  float* sm_norm = reinterpret_cast<float*>(smem + L::off_norm);
  float* sm_stats = reinterpret_cast<float*>(smem + L::off_stats);
  double* sm_gstats = reinterpret_cast<double*>(smem + L::off_gstats);
  int* sm_masks = reinterpret_cast<int*>(smem + L::off_masks);
  uint64_t* sm_bar = reinterpret_cast<uint64_t*>(smem + L::off_bar);
  float* sm_melw = reinterpret_cast<float*>(smem + L::off_melw);
  float* sm_dct = sm_melw + kBands * P.mel_maxt * 16;
  float* sm_lifter = sm_dct + (kMfcc ? P.n_mels * P.n_ceps : 0);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int t = lane & 15;       // lane inside the frame's half-warp
  const int half = lane >> 4;    // which of the warp's two frames
  const int n_out = P.n_out;
  const bool want_stats = (P.mode == 1) || (P.mode == 3);

  // ---- one-time table staging ----------------------------------------------------------------
  for (int i = tid; i < 512; i += kThreads) sm_window[i] = P.window[i];
  for (int i = tid; i < 256; i += kThreads) sm_tw1[i] = P.tw1[i];
  for (int i = tid; i < 128; i += kThreads) sm_tw2[i] = P.tw2[i];
  for (int i = tid; i < kMaxMels; i += kThreads) sm_k0[i] = P.mel_k0[i];
  for (int i = tid; i < kBands * P.mel_maxt * 16; i += kThreads) sm_melw[i] = P.mel_w[i];
  if (kMfcc) {
    for (int i = tid; i < P.n_mels * P.n_ceps; i += kThreads) sm_dct[i] = P.dct[i];
    for (int i = tid; i < P.n_ceps; i += kThreads) sm_lifter[i] = P.lifter[i];
  }
  for (int i = tid; i < 4 * kMaxMels; i += kThreads) sm_stats[i] = 0.f;
  for (int i = tid; i < 2 * kMaxMels + 2; i += kThreads) sm_gstats[i] = 0.0;
  if (P.mode == 2 && tid < n_out) {
    // finalise the all-reduced sums: mean, 1/(std + 1e-9) (unbiased)
    const double n = P.stats_in[2 * n_out];
    const double mean = P.stats_in[tid] / n;
    double var = (P.stats_in[n_out + tid] - P.stats_in[tid] * mean) / (n - 1.0);
    var = var > 0.0 ? var : 0.0;
    sm_norm[tid] = static_cast<float>(mean);
    sm_norm[kMaxMels + tid] = static_cast<float>(1.0 / (sqrt(var) + 1e-9));
  }
  if (tid == 0) {
    mbar_init(&sm_bar[0], 1);
    mbar_init(&sm_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---- tile staging ---------------------------------------------------------------------------
  auto stage_tile = [&](int tile_idx, int buf) {
    if (tile_idx >= P.n_tiles) return;
    const Tile tl = P.tiles[tile_idx];
    if (tl.nframes == 0) return;
    const int nsamp = kFrameShift * tl.nframes + (kFrameLen - kFrameShift);
    const TIn* src = reinterpret_cast<const TIn*>(P.wav) + tl.wav_off;
    if (tl.aux) {
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t bytes = nsamp * (uint32_t)sizeof(TIn);
        mbar_expect_tx(&sm_bar[buf], bytes);
        tma_bulk_g2s(sm_in[buf], src, bytes, &sm_bar[buf]);
      }
    } else {
      for (int i = tid; i < nsamp; i += kThreads) sm_in[buf][i] = src[i];
    }
  };

  uint32_t phase[2] = {0u, 0u};
  float* my_scratch = sm_scratch + (warp * 2 + half) * kScratchFloats;
  float2* my_T = reinterpret_cast<float2*>(my_scratch);
  const int src_lane = (lane & 16) | ((16 - t) & 15);

  stage_tile(blockIdx.x, 0);

  int it = 0;
  for (int tile_idx = blockIdx.x; tile_idx < P.n_tiles; tile_idx += gridDim.x, ++it) {
    const int buf = it & 1;
    const Tile tl = P.tiles[tile_idx];
    // prefetch the next tile into the other buffer (its previous reader finished before the
    // __syncthreads that closed the previous iteration)
    stage_tile(tile_idx + gridDim.x, buf ^ 1);

    if (tl.nframes == 0) {
      // zero-fill tile: pad_sequence's zeros (ref: lid/raw_datasets.py:347-350)
      const long long total = static_cast<long long>(tl.aux) * n_out;
      for (long long i = tid; i < total; i += kThreads) {
        const long long r = i / n_out;
        const int d = static_cast<int>(i - r * n_out);
        P.out[(tl.out_row + r) * P.out_ld + d] = 0.f;
      }
      __syncthreads();
      continue;
    }

    if (tid < P.n_masks * 4) sm_masks[tid] = P.masks[static_cast<long long>(tl.utt) * P.n_masks * 4 + tid];
    if (tl.aux) {
      mbar_wait(&sm_bar[buf], phase[buf]);
      phase[buf] ^= 1u;
    }
    __syncthreads();   // masks + (fallback path) generic stores visible
    const TIn* in = sm_in[buf];
    float* stats = sm_stats + buf * 2 * kMaxMels;

    float s1[kBands], s2[kBands];
#pragma unroll
    for (int b = 0; b < kBands; ++b) s1[b] = s2[b] = 0.f;

    for (int r0 = 0; r0 < tl.nframes; r0 += kFramesPerRound) {
      const int fl = r0 + warp * 2 + half;        // frame inside the tile
      if (r0 + warp * 2 >= tl.nframes) break;     // warp-uniform
      const bool active = fl < tl.nframes;
      const TIn* fr = in + kFrameShift * (active ? fl : r0 + warp * 2);

      // ---- load, DC removal, pre-emphasis, window (ta: compliance/kaldi.py:183-204) ---------------
      float2 v[16];
      float pv[13];
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 13; ++j) {
        const int n = t + 16 * j;
        if (j < 12 || t < 8) {
          v[j] = InTraits<TIn>::ld2(fr + 2 * n, P.in_scale);
          pv[j] = InTraits<TIn>::ld1(fr + (n == 0 ? 0 : 2 * n - 1), P.in_scale);
          sum += v[j].x;
          sum += v[j].y;
        } else {
          v[j] = make_float2(0.f, 0.f);
          pv[j] = 0.f;
        }
      }
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = P.remove_dc ? __fdiv_rn(sum, static_cast<float>(kFrameLen)) : 0.f;
      const float c = P.preemph;
#pragma unroll
      for (int j = 0; j < 13; ++j) {
        const int n = t + 16 * j;
        const float2 w = *reinterpret_cast<const float2*>(sm_window + 2 * n);
        const float te = __fsub_rn(v[j].x, mean);
        const float to = __fsub_rn(v[j].y, mean);
        const float tp = __fsub_rn(pv[j], mean);
        // x[j] - c * x[j-1]: product and difference rounded separately, as the reference's two tensor ops
        const float se = __fsub_rn(te, __fmul_rn(c, tp));
        const float so = __fsub_rn(to, __fmul_rn(c, te));
        v[j] = make_float2(__fmul_rn(se, w.x), __fmul_rn(so, w.y));
      }
      if (t >= 8) v[12] = make_float2(0.f, 0.f);
      v[13] = v[14] = v[15] = make_float2(0.f, 0.f);

      // ---- stage 1: 16-point DFT over j, twiddle W256^(K1*t), transpose through shared ------------
      fft16<true>(v);
      __syncwarp();   // previous round's readers of my_T are done
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const int K1 = rev4(p);
        float2 y = v[p];
        if (K1 != 0) y = cmul(y, sm_tw1[K1 * 16 + t]);
        my_T[K1 * kRowStride + t] = y;
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 two = *reinterpret_cast<const float4*>(my_T + t * kRowStride + 2 * q);
        v[2 * q] = make_float2(two.x, two.y);
        v[2 * q + 1] = make_float2(two.z, two.w);
      }
      // ---- stage 2: v[p] = Z[t + 16*rev4(p)] --------------------------------------------------------
      fft16<false>(v);

      // ---- real-FFT split + power; lane t pairs with lane 16-t ------------------------------------
      float2 pb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 s = v[rev4(15 - i)];
        pb[i].x = __shfl_sync(0xffffffffu, s.x, src_lane);
        pb[i].y = __shfl_sync(0xffffffffu, s.y, src_lane);
      }
      if (t == 0) {
        // lane 0 pairs k=16i with 256-16i = its own Z[16-i]; k=0 pairs with itself (DC / Nyquist)
#pragma unroll
        for (int i = 7; i >= 1; --i) pb[i] = pb[i - 1];
        pb[0] = v[0];
      }
      __syncwarp();   // all lanes finished reading my_T before the power bins overwrite it
      float* my_P = my_scratch;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 a = v[rev4(i)];
        const float2 b = pb[i];
        const float2 e2 = make_float2(a.x + b.x, a.y - b.y);          // 2E = a + conj(b)
        const float2 o2 = make_float2(a.y + b.y, b.x - a.x);          // 2O = -i (a - conj(b))
        const float2 tw = cmul(o2, sm_tw2[i * 16 + t]);               // W512^k * 2O
        const float2 xa = e2 + tw;                                    // 2 X[k]
        const float2 xb = e2 - tw;                                    // 2 conj(X[256-k])
        const int k = t + 16 * i;
        my_P[k] = 0.25f * fmaf(xa.x, xa.x, xa.y * xa.y);
        my_P[256 - k] = 0.25f * fmaf(xb.x, xb.x, xb.y * xb.y);
      }
      if (t == 0) my_P[128] = fmaf(v[2].x, v[2].x, v[2].y * v[2].y);   // Z[128] = v[rev4(8)]
      __syncwarp();

      // ---- sparse triangular mel + log (ta: compliance/kaldi.py:621-633) ---------------------------
      float val[kBands];
#pragma unroll
      for (int b = 0; b < kBands; ++b) {
        const int m = t + 16 * b;
        float acc = 0.f;
        const int k0 = sm_k0[m];
        const float* wp = sm_melw + (b * P.mel_maxt) * 16 + t;
        const int nt = P.band_taps[b];
        for (int i = 0; i < nt; ++i) {
          const int k = min(k0 + i, kBins - 1);
          acc = fmaf(my_P[k], wp[i * 16], acc);
        }
        val[b] = logf(fmaxf(acc, P.log_floor));
      }

      // ---- MFCC: DCT-II + lifter (ta: compliance/kaldi.py:648-666,786-796) --------------------------
      if (kMfcc) {
        float* my_L = my_scratch + kLogmelOff;
#pragma unroll
        for (int b = 0; b < kBands; ++b)
          if (t + 16 * b < P.n_mels) my_L[t + 16 * b] = val[b];
        __syncwarp();
#pragma unroll
        for (int b = 0; b < kBands; ++b) {
          const int cidx = t + 16 * b;
          float acc = 0.f;
          if (cidx < P.n_ceps) {
            for (int m = 0; m < P.n_mels; ++m) acc = fmaf(my_L[m], sm_dct[m * P.n_ceps + cidx], acc);
            acc = __fmul_rn(acc, sm_lifter[cidx]);
          }
          val[b] = acc;
        }
      }

      // ---- epilogue: stats, global CMVN, SpecAugment zero-fill, store -------------------------------
      if (active) {
        const int tf = tl.t0 + fl;   // frame index inside the utterance
        bool row_masked = false;
        if (P.mode != 3) {
          for (int q = 0; q < P.n_masks; ++q) row_masked |= (tf >= sm_masks[4 * q] && tf < sm_masks[4 * q + 1]);
        }
        float* orow = P.out + (tl.out_row + fl) * P.out_ld;
#pragma unroll
        for (int b = 0; b < kBands; ++b) {
          const int d = t + 16 * b;
          if (d < n_out) {
            float x = val[b];
            if (want_stats) {
              s1[b] += x;
              s2[b] = fmaf(x, x, s2[b]);
            }
            if (P.mode == 2) x = (x - sm_norm[d]) * sm_norm[kMaxMels + d];
            if (P.mode != 1 && P.mode != 3) {
              bool z = row_masked;
              for (int q = 0; q < P.n_masks; ++q) z |= (d >= sm_masks[4 * q + 2] && d < sm_masks[4 * q + 3]);
              if (z) x = 0.f;
            }
            orow[d] = x;
          }
        }
      }
    }

    if (want_stats) {
#pragma unroll
      for (int b = 0; b < kBands; ++b) {
        s1[b] += __shfl_xor_sync(0xffffffffu, s1[b], 16);
        s2[b] += __shfl_xor_sync(0xffffffffu, s2[b], 16);
        const int d = t + 16 * b;
        if (half == 0 && d < n_out) {
          atomicAdd(&stats[d], s1[b]);
          atomicAdd(&stats[kMaxMels + d], s2[b]);
        }
      }
    }
    __syncthreads();   // everyone is done with sm_in[buf], sm_masks and has published its stats
    if (want_stats && tid < 2 * kMaxMels) {
      const int which = tid / kMaxMels, d = tid - which * kMaxMels;
      if (d < n_out) {
        const float s = stats[tid];
        stats[tid] = 0.f;
        if (P.mode == 1) {
          atomicAdd(&P.utt_stats[(static_cast<long long>(tl.utt) * 2 + which) * n_out + d], static_cast<double>(s));
        } else {
          sm_gstats[which * n_out + d] += static_cast<double>(s);
        }
      }
    }
    if (P.mode == 3 && tid == 2 * kMaxMels) sm_gstats[2 * n_out] += static_cast<double>(tl.nframes);
  }

  if (P.mode == 3) {
    __syncthreads();
    for (int i = tid; i < 2 * n_out + 1; i += kThreads) {
      const int slot = (i == 2 * n_out) ? i : i;   // same indexing on both sides
      double vsum = sm_gstats[slot];
      if (vsum != 0.0) atomicAdd(&P.stats_out[i], vsum);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// CMVN apply + masks (second pass of per-utterance / global CMVN).  One CTA per tile.
// ------------------------------------------------------------------------------------------------
struct ApplyParams {
  float* feats;
  long long ld;
  const Tile* tiles;
  int n_tiles;
  int n_out;
  const int* masks;
  int n_masks;
  const double* utt_stats;     // [B][2][n_out] or NULL
  const long long* utt_frames; // [B]
  const double* glob_stats;    // [2*n_out+1] or NULL
  int normalize;               // 0 -> masks only (standalone SpecAugment application)
};

__global__ void __launch_bounds__(256) cmvn_apply_kernel(const __grid_constant__ ApplyParams P) {
  __shared__ float s_mean[kMaxMels], s_inv[kMaxMels];
  __shared__ int s_masks[kMaxMasks * 4];
  const int tid = threadIdx.x;
  for (int tile_idx = blockIdx.x; tile_idx < P.n_tiles; tile_idx += gridDim.x) {
    const Tile tl = P.tiles[tile_idx];
    if (tl.nframes == 0) continue;
    __syncthreads();
    if (!P.normalize) {
      if (tid < P.n_out) {
        s_mean[tid] = 0.f;
        s_inv[tid] = 1.f;
      }
    } else if (tid < P.n_out) {
      double n, s, ss;
      if (P.utt_stats) {
        n = static_cast<double>(P.utt_frames[tl.utt]);
        s = P.utt_stats[(static_cast<long long>(tl.utt) * 2 + 0) * P.n_out + tid];
        ss = P.utt_stats[(static_cast<long long>(tl.utt) * 2 + 1) * P.n_out + tid];
      } else {
        n = P.glob_stats[2 * P.n_out];
        s = P.glob_stats[tid];
        ss = P.glob_stats[P.n_out + tid];
      }
      const double mean = s / n;
      double var = (ss - s * mean) / (n - 1.0);   // n == 1 -> NaN, as torch.std of one sample
      var = var > 0.0 ? var : (var == var ? 0.0 : var);
      s_mean[tid] = static_cast<float>(mean);
      s_inv[tid] = static_cast<float>(1.0 / (sqrt(var) + 1e-9));
    }
    if (tid < P.n_masks * 4) s_masks[tid] = P.masks[static_cast<long long>(tl.utt) * P.n_masks * 4 + tid];
    __syncthreads();
    const int total = tl.nframes * P.n_out;
    for (int i = tid; i < total; i += blockDim.x) {
      const int r = i / P.n_out;
      const int d = i - r * P.n_out;
      float* p = P.feats + (tl.out_row + r) * P.ld + d;
      float x = *p;
      if (P.normalize) x = (x - s_mean[d]) * s_inv[d];
      const int tf = tl.t0 + r;
      bool z = false;
      for (int q = 0; q < P.n_masks; ++q)
        z |= (tf >= s_masks[4 * q] && tf < s_masks[4 * q + 1]) || (d >= s_masks[4 * q + 2] && d < s_masks[4 * q + 3]);
      *p = z ? 0.f : x;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// waveform-level stages (ref: lid/audio_processor.py:108-115,129-134).  One CTA per utterance.
// ------------------------------------------------------------------------------------------------
struct WaveParams {
  const float* in;
  float* out;
  const long long* offsets;   // [B]
  const long long* lengths;   // [B]
  int normalize;
  float dither;
  const float* noise;
  float preemph;
};

__global__ void __launch_bounds__(512) wave_stages_kernel(const __grid_constant__ WaveParams P) {
  __shared__ double s_red[2][16];
  __shared__ float s_mean, s_div;
  const int u = blockIdx.x;
  const long long off = P.offsets[u], n = P.lengths[u];
  const float* x = P.in + off;
  const float* nz = P.noise ? P.noise + off : nullptr;
  float* y = P.out + off;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mean = 0.f, div = 1.f;
  if (P.normalize) {
    // two-pass mean / unbiased variance in fp64 (torch.std_mean accumulates in fp32 with a cascade)
    double s = 0.0;
    for (long long i = tid; i < n; i += blockDim.x) s += x[i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_red[0][warp] = s;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[0][w];
    const double m = tot / static_cast<double>(n);
    double q = 0.0;
    for (long long i = tid; i < n; i += blockDim.x) {
      const double dlt = x[i] - m;
      q += dlt * dlt;
    }
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) s_red[1][warp] = q;
    __syncthreads();
    double qt = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) qt += s_red[1][w];
    if (tid == 0) {
      s_mean = static_cast<float>(m);
      s_div = static_cast<float>(sqrt(qt / static_cast<double>(n - 1))) + 1e-6f;
    }
    __syncthreads();
    mean = s_mean;
    div = s_div;
  }
  auto stage1 = [&](long long i) -> float {
    float v = x[i];
    if (P.normalize) v = __fdiv_rn(__fsub_rn(v, mean), div);
    if (P.dither != 0.f) v = __fadd_rn(v, __fmul_rn(P.dither, nz[i]));
    return v;
  };
  for (long long i = tid; i < n; i += blockDim.x) {
    float v = stage1(i);
    if (P.preemph != 0.f && i > 0) v = __fsub_rn(v, __fmul_rn(P.preemph, stage1(i - 1)));
    y[i] = v;
  }
}

}  // namespace lidfe
