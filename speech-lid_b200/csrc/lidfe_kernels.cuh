// lidfe_kernels.cuh -- sm_100a kernels of the speech-lid front-end.
//
// Path (ref: lid/audio_processor.py:41-69 -> ta: compliance/kaldi.py:154-217,591-633):
//   frame (400 samples, hop 160) -> DC removal -> pre-emphasis (replicate-left) -> Povey window ->
//   zero-pad to 512 -> |rfft|^2 -> sparse triangular mel -> log(max(., eps)) [-> DCT + lifter] ->
//   [global CMVN] -> [SpecAugment zero-fill] -> out, plus per-utterance / global sum & sum-of-squares.
//
// Mapping ("frame-pair packed"): a persistent CTA of 4 warps (4 CTAs per SM) walks a contiguous range of tiles of
// <= 16 consecutive frames of one utterance.  A tile's 160*F+240 samples are staged into shared memory by ONE TMA
// bulk copy (cp.async.bulk + mbarrier) into a single buffer that the last warp to have consumed the previous tile
// refills, so the copy overlaps most of the current tile's math and no CTA-wide barrier is needed per tile.
// A half-warp (16 lanes) owns TWO adjacent frames (A, B) and carries them as the two halves of Blackwell's
// packed f32x2 registers: every add / mul / fma of the FFT is one FADD2 / FMUL2 / FFMA2 for both frames, and
// every table value (window, twiddles, mel weights) is loaded once and broadcast to both.  The 512-point real
// FFT is a 256-point complex FFT of z[n] = x[2n] + i x[2n+1] done as 16 x 16 (two in-register radix-4x4
// 16-point DFTs per lane, one shared-memory transpose in between) followed by the real-FFT split, which pairs
// lane t with lane 16-t through warp shuffles; the mel stage walks every power bin once ("segment form").
// The kernel is latency / issue bound with the shared-memory + shuffle pipe at ~70 %, the FMA pipe and the
// instruction issue at ~50 %; the other kernels of the path (CMVN apply, the MFCC GEMM, the
// waveform stages) follow further down.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <type_traits>
#include <stdint.h>

// Development only: -DLIDFE_ABL=<bits> builds ABLATED variants of fbank_kernel (wrong results) that tools/ablate.py times
// against the full kernel to price each stage inside the real pipeline.  The product is always built with 0.
#ifndef LIDFE_ABL
#define LIDFE_ABL 0
#endif
#ifndef LIDFE_UNIT_SHORTCUT
#define LIDFE_UNIT_SHORTCUT 1    // reference call (DC removal + pre-emphasis 1.0): x[n] - x[n-1] without the frame mean
#endif

namespace lidfe {

constexpr int kFrameLen = 400;
constexpr int kFrameShift = 160;
constexpr int kFftLen = 512;
constexpr int kBins = kFftLen / 2 + 1;            // 257
constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kTileFrames = kWarps * 4;           // two frames per half-warp
constexpr int kTileSamples = kFrameShift * kTileFrames + (kFrameLen - kFrameShift);   // 2800
constexpr int kTileSamplesPad = kTileSamples + 32;
constexpr int kMaxMels = 80;
constexpr int kBands = kMaxMels / 16;             // lane t owns output dims t + 16*b
constexpr int kRowStride = 18;                    // (A,B) pairs per transpose row: 16 + 2 pad -> LDS.128 conflict free
constexpr int kPlaneFloats = 16 * kRowStride * 2; // one plane (re or im) of a frame pair: 576 floats
// per half-warp scratch: ONE transpose plane (re, then im, go through it one after the other), reused for the 257 power
// pairs (floats 0..513); behind them the parked feature pairs of the statistics pass, and for the in-kernel MFCC
// epilogue the log-mel staging area
constexpr int kLogmelOff = 520;                   // MFCC log-mel staging (80 pairs)
__host__ __device__ constexpr int vals_off(bool mfcc) { return mfcc ? 688 : 520; }         // parked feature pairs (80 pairs)
__host__ __device__ constexpr int scratch_floats(bool mfcc) { return mfcc ? 864 : 704; }
constexpr int kMaxMasks = 8;
// mel steps per band of the two banks the reference uses, for the fully unrolled kernel variants:
// kind 1 = Kaldi 80 x 257 (20 Hz .. 8 kHz, mel = 1127 ln(1 + f/700)), kind 2 = HTK 80 x 257 (0 .. 8 kHz, 2595 log10)
__host__ __device__ constexpr int std_taps(int kind, int b) {
  return kind == 1 ? (b == 0 ? 2 : b == 1 ? 2 : b == 2 ? 4 : b == 3 ? 6 : 9)
                   : (b == 0 ? 1 : b == 1 ? 3 : b == 2 ? 4 : b == 3 ? 6 : 10);
}

struct Tile {
  long long wav_off;    // first sample of the tile's first frame in the packed buffer
  long long out_row;    // output row of the tile's first frame (or first zero-fill row)
  int nframes;          // 1..kTileFrames; 0 -> zero-fill tile
  int utt;              // utterance index
  int t0;               // index of the first frame inside its utterance (time masks)
  int aux;              // zero-fill tiles: number of rows to clear.  else: 1 if TMA-eligible (16B aligned)
};
// A SPAN is the unit of work a CTA claims: up to kMaxSpanTiles consecutive tiles of ONE utterance, described like a
// tile whose nframes may exceed kTileFrames (the kernel cuts it into tiles itself: tile i starts 16 i frames / 2560 i
// samples / 16 i rows further), or a run of zero-fill rows (nframes == 0, aux = rows).  Spans are listed utterance-
// major and claimed in that order, so utterances complete progressively and the CTA that completes one can normalise
// it while its rows are still in L2 (per-utterance CMVN / top_db in ONE kernel).
typedef Tile Span;
constexpr int kMaxSpanTiles = 8;

struct FbankParams {
  const void* wav;
  float* out;
  long long out_ld;
  const Span* spans;
  int n_spans;
  // warp-autonomous kernel (lidfe_fbank_warp.cuh): spans = runs of consecutive frames of one utterance
  const Span* wspans;
  int n_wspans;
  const int* w_first;       // [warps + 1]: warp g's static spans are [w_first[g], w_first[g + 1])
  int n_wstatic;            // spans [n_wstatic, n_wspans) are the pool, claimed one at a time (sched[0])
  int w_tab_bytes;          // shared-memory bytes in front of the per-warp areas (tables' mbarrier, tables, global-CMVN constants)
  // dynamic schedule + per-utterance completion (all self-cleaning: the kernel leaves them as it found them)
  int* sched;               // [0] span claim counter, [1] CTA exit counter
  int* utt_done;            // [2][B_cap] frames of utterance i whose statistics have been handed over
  // per-utterance second stage (CMVN / top_db).  The rows of every utterance are cut into ITEMS of <= apply_rows rows,
  // listed utterance-major.  Every WARP of a service CTA (blockIdx.x >= n_compute; they do nothing else) and, once the
  // spans have run out, every warp of the compute CTAs claims items one after the other (one atomic each), waits until
  // the item's utterance has all its frames handed over, finalises the utterance's constants and rewrites the rows in
  // place from L2.  The per-utterance bookkeeping is double buffered by launch parity: launch e works on half e & 1 and
  // puts the other half (launch e - 1's, fully consumed) back to rest, so nothing is reset behind a reader's back.
  int* ictl;                // [0] next item to claim
  const int4* items;        // [n_items] (utt, first row, rows, frames of the utterance)
  int n_items;
  int n_compute;            // CTAs [0, n_compute) process spans; the rest are service CTAs (per-utterance modes only)
  int k_static;             // every compute CTA starts with k_static CONSECUTIVE spans (CTA b: b * k_static ...), the rest are claimed
  int parity;               // launch parity of this plan; B_cap = stride between the halves of the per-utterance arrays
  int b_cap;
  int n_utts;
  long long* dbg_buf;       // [grid][8] per-CTA timeline (LIDFE_DBG & 16), else NULL
  int dbg;                  // development switches (LIDFE_DBG): 1 skip apply_rows, 2 skip the stats atomics, 4 skip fences, 8 no help at span ends
  const long long* utt_frames;      // [B]
  // warp kernel, fused per-utterance second stage (lidfe_fbank_warp.cuh): the warp whose hand-over completes an
  // utterance publishes that utterance's items in the ready queue; warps take them at span ends and when out of spans
  int* wq;                          // [0] entries published, [1] entries claimed, [4 + i] item index + 1 (0 = not yet written)
  const int* utt_first_item;        // [B + 1] first item of every utterance (items are utterance-major)
  const long long* utt_out_row;     // [B]
  const long long* utt_first_tile;  // [B] (tile-blocked log-mel workspace of the two-kernel MFCC path)
  // constant tables: ONE device blob laid out exactly like the kernel's shared-memory table area, so a single TMA
  // bulk copy stages it:  window[416] | tw1[16][16] float2 = W256^(K1*t) | tw2[8][16] float2 = W512^(t+16i) |
  // mel_k0[80] int | mel (wa, wb) pairs per band [steps_b][16] (x 0.25: the power bins are left scaled by 4) |
  // dct[n_mels][n_ceps] | lifter[n_ceps]   (each section padded to 16 bytes)
  const void* const_blob;
  int const_bytes;
  int band_taps[kBands];    // mel steps per band (segment plan, see build_mel_plan in lidfe_abi.cu)
  int n_mels, n_ceps, n_out;
  float preemph, log_floor, log_of_floor, in_scale;   // log_of_floor = the reference's log of the floor, rounded on the host
  float log_scale;          // lg2(x) * log_scale: ln 2 (natural log) or 10 log10(2) (dB)
  int center, pad;          // LIDFE_FRAMING_CENTER: frames centred at 160 f over the constant-padded, reflect-extended signal
  const long long* utt_offsets;  // [B] first sample of each utterance (edge tiles of CENTER framing)
  const long long* utt_lengths;  // [B]
  unsigned* utt_max;        // [2][B_cap] order-preserving encoding of the utterance's max feature (LIDFE_POST_TOPDB); 0 at rest
  unsigned* utt_min;        // [2][B_cap] same for its min (lets the clamp pass be skipped); 0xffffffff at rest
  float top_db;
  int remove_dc;
  // normalize_wav fused into the sample load (row f2): (x * in_scale - mean) / div per utterance, or NULL
  const float2* utt_wnorm;  // [B] (mean, std + 1e-6) from wave_stats
  float dither;             // > 0: x += dither * U[0,1), Philox keyed by (seed, utterance, sample) (ref: lid/audio_processor.py:129), KALDI framing
  unsigned long long seed;
  // epilogue
  const int* masks;         // [B][n_masks][4]
  int n_masks;
  int mode;                 // LIDFE_CMVN_*
  const double* stats_in;   // [2*n_out+1]
  double* stats_out;        // [2*n_out+1]
  double* utt_stats;        // [2][B_cap][2][n_out], zero at rest
  // two-kernel MFCC path: `out` is the log-mel workspace in the tile-blocked layout mfcc_dct_kernel reads,
  // [tile][n_out / 4][16 frames][4 dims] (a 16-byte chunk = 4 consecutive dims of one frame; the 16 frames of a tile
  // sit next to each other so that a half-warp of the DCT kernel loads 256 contiguous bytes)
  int ws_blocked;
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).  The samples are read once:
// they go through L2 with an evict-first policy so that the feature rows written by this kernel stay resident for the
// normalisation pass that follows.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_bulk_g2s_plain(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// order-preserving float -> uint map (atomicMax on floats of either sign); 0 is below every finite value
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---- packed f32x2 arithmetic: .x = frame A, .y = frame B (SASS: FADD2 / FMUL2 / FFMA2) ---------------
typedef float2 f2;
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 bc(float s) { return make_float2(s, s); }
__device__ __forceinline__ f2 neg2(f2 a) { return make_float2(-a.x, -a.y); }
// one MUFU.LG2: the argument is always above the log floor, a normal number (lidfe_create checks), so the denormal
// pre-scaling __log2f wraps around the instruction (6 more instructions per value) is dead weight here
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// log2(x) * scale for a normal positive x.  MUFU.LG2 alone is good to 2^-22 RELATIVE to the result (|log2| ~ 10-25 here:
// ~4e-6 in the natural log, several times the reference's 1-ulp logf and enough to fail the "no worse than the reference"
// metric in the bins where nothing else goes wrong); on a mantissa in [sqrt(1/2), sqrt(2)) its error is 2^-22 ABSOLUTE,
// so the exponent is split off first: log2(x) = e + log2(m).
#ifndef LIDFE_LOG_MODE
#define LIDFE_LOG_MODE 2
#endif
__device__ __forceinline__ float log2_scaled(float x, float scale) {
#if LIDFE_LOG_MODE == 0
  return lg2_ftz(x) * scale;
#elif LIDFE_LOG_MODE == 1
  return log2f(x) * scale;
#else
  const int b = __float_as_int(x);
  const int e = (b - 0x3f3504f3) >> 23;
  const float m = __int_as_float(b - (e << 23));
  return fmaf(static_cast<float>(e), scale, lg2_ftz(m) * scale);
#endif
}

// (R + iI) *= (wr + i wi), both frames at once, one scalar twiddle
__device__ __forceinline__ void cmul2(f2& R, f2& I, float wr, float wi) {
  const f2 nr = fma2(R, bc(wr), mul2(I, bc(-wi)));
  const f2 ni = fma2(R, bc(wi), mul2(I, bc(wr)));
  R = nr;
  I = ni;
}

// forward 4-point DFT on (R,I) pairs, outputs in natural order
__device__ __forceinline__ void radix4(f2& r0, f2& i0, f2& r1, f2& i1, f2& r2, f2& i2, f2& r3, f2& i3) {
  const f2 t0r = add2(r0, r2), t0i = add2(i0, i2), t1r = sub2(r0, r2), t1i = sub2(i0, i2);
  const f2 t2r = add2(r1, r3), t2i = add2(i1, i3), t3r = sub2(r1, r3), t3i = sub2(i1, i3);
  r0 = add2(t0r, t2r);
  i0 = add2(t0i, t2i);
  r2 = sub2(t0r, t2r);
  i2 = sub2(t0i, t2i);
  r1 = add2(t1r, t3i);
  i1 = sub2(t1i, t3r);
  r3 = sub2(t1r, t3i);
  i3 = add2(t1i, t3r);
}
// same with input 3 == 0 (zero padding of the 400-sample frame to 512)
__device__ __forceinline__ void radix4_z3(f2& r0, f2& i0, f2& r1, f2& i1, f2& r2, f2& i2, f2& r3, f2& i3) {
  const f2 t0r = add2(r0, r2), t0i = add2(i0, i2), t1r = sub2(r0, r2), t1i = sub2(i0, i2);
  const f2 ur = r1, ui = i1;
  r0 = add2(t0r, ur);
  i0 = add2(t0i, ui);
  r2 = sub2(t0r, ur);
  i2 = sub2(t0i, ui);
  r1 = add2(t1r, ui);
  i1 = sub2(t1i, ur);
  r3 = sub2(t1r, ui);
  i3 = add2(t1i, ur);
}

// forward 16-point DFT in registers: 4x4 radix-4.  In: natural order.  Out: position p holds X[rev4(p)],
// rev4(p) = (p >> 2) + 4 * (p & 3).  kTailZero: inputs 13, 14, 15 are known zeros.
template <bool kTailZero>
__device__ __forceinline__ void fft16(f2 (&R)[16], f2 (&I)[16]) {
  constexpr float kC1 = 0.92387953251128674f;   // cos(pi/8)
  constexpr float kS1 = 0.38268343236508977f;   // sin(pi/8)
  constexpr float kH = 0.70710678118654752f;    // sqrt(1/2)
  radix4(R[0], I[0], R[4], I[4], R[8], I[8], R[12], I[12]);
  if (kTailZero) {
    radix4_z3(R[1], I[1], R[5], I[5], R[9], I[9], R[13], I[13]);
    radix4_z3(R[2], I[2], R[6], I[6], R[10], I[10], R[14], I[14]);
    radix4_z3(R[3], I[3], R[7], I[7], R[11], I[11], R[15], I[15]);
  } else {
    radix4(R[1], I[1], R[5], I[5], R[9], I[9], R[13], I[13]);
    radix4(R[2], I[2], R[6], I[6], R[10], I[10], R[14], I[14]);
    radix4(R[3], I[3], R[7], I[7], R[11], I[11], R[15], I[15]);
  }
  // element n2 + 4*k1 *= W16^(n2*k1)
  cmul2(R[5], I[5], kC1, -kS1);                                            // W^1
  {                                                                        // W^2 = h(1 - i)
    const f2 r = mul2(add2(R[9], I[9]), bc(kH)), i = mul2(sub2(I[9], R[9]), bc(kH));
    R[9] = r; I[9] = i;
  }
  cmul2(R[13], I[13], kS1, -kC1);                                          // W^3
  {
    const f2 r = mul2(add2(R[6], I[6]), bc(kH)), i = mul2(sub2(I[6], R[6]), bc(kH));
    R[6] = r; I[6] = i;
  }
  {                                                                        // W^4 = -i
    const f2 r = I[10], i = neg2(R[10]);
    R[10] = r; I[10] = i;
  }
  {                                                                        // W^6 = h(-1 - i)
    const f2 r = mul2(sub2(I[14], R[14]), bc(kH)), i = mul2(add2(R[14], I[14]), bc(-kH));
    R[14] = r; I[14] = i;
  }
  cmul2(R[7], I[7], kS1, -kC1);                                            // W^3
  {
    const f2 r = mul2(sub2(I[11], R[11]), bc(kH)), i = mul2(add2(R[11], I[11]), bc(-kH));
    R[11] = r; I[11] = i;
  }
  cmul2(R[15], I[15], -kC1, kS1);                                          // W^9
  radix4(R[0], I[0], R[1], I[1], R[2], I[2], R[3], I[3]);
  radix4(R[4], I[4], R[5], I[5], R[6], I[6], R[7], I[7]);
  radix4(R[8], I[8], R[9], I[9], R[10], I[10], R[11], I[11]);
  radix4(R[12], I[12], R[13], I[13], R[14], I[14], R[15], I[15]);
}
__host__ __device__ constexpr int rev4(int p) { return (p >> 2) + 4 * (p & 3); }

template <typename TIn>
struct InTraits;
template <>
struct InTraits<float> {
  static __device__ __forceinline__ float2 ld2(const float* p, float) { return *reinterpret_cast<const float2*>(p); }
};
template <>
struct InTraits<short> {
  static __device__ __forceinline__ float2 ld2(const short* p, float s) {
    const short2 v = *reinterpret_cast<const short2*>(p);
    return make_float2(static_cast<float>(v.x) * s, static_cast<float>(v.y) * s);
  }
};

// shared-memory carve-up (dynamic)
template <typename TIn, bool kMfcc>
struct SmemLayout {
  static constexpr int kInBytes = ((kTileSamplesPad * (int)sizeof(TIn)) + 127) / 128 * 128;
  static constexpr int off_in0 = 0;                                                // ONE input buffer (see the tile loop)
  static constexpr int off_scratch = kInBytes;                                     // [2*kWarps][scratch_floats] floats
  static constexpr int off_norm = off_scratch + 2 * kWarps * scratch_floats(kMfcc) * 4;   // [80] float2 (mean, inv_std)
  static constexpr int off_lo = off_norm + kMaxMels * 8;                           // [80] float: -(mean - (float)mean) * inv_std
  static constexpr int off_acc = off_lo + kMaxMels * 4;                            // fp64: [kWarps][2][80] per-warp sums + frame count + wmax/wmin
  static constexpr int off_masks = off_acc + (kWarps * 2 * kMaxMels + 6) * 8;      // [kMaxMasks][4] int: the current span's utterance
  static constexpr int off_bar = off_masks + kMaxMasks * 16;                       // mbarriers: full @0, tables @16; arrival counter @32; ctl ints @36..
  static constexpr int off_span = off_bar + 64;                                    // [2] Span descriptors: current, next (prefetched by cp.async)
  // constant tables, one contiguous block = the device blob (see FbankParams::const_blob)
  static constexpr int off_window = off_span + 2 * 32;                             // [416]
  static constexpr int off_tw1 = off_window + 416 * 4;                             // [16][16] float2
  static constexpr int off_tw2 = off_tw1 + 256 * 8;                                // [8][16] float2
  static constexpr int off_k0 = off_tw2 + 128 * 8;                                 // [80] int
  static constexpr int off_melw = off_k0 + kMaxMels * 4;                           // sum(band_taps)*16 float2, then dct, lifter
};

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float4 ld_cg_f4(const float* p) {   // L2 load: the rows were written by other SMs
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_cg_f(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// x / d with r = 1 / d rounded to nearest: quotient, residual, correction (the division sequence without its scaling
// fix-ups; samples and 1/std are far from the exponent limits).  Shared by wave_stages_kernel and the fused load, so
// the two normalize_wav paths produce the same bits.
__device__ __forceinline__ float div_by(float x, float d, float r) {
  const float q = __fmul_rn(x, r);
  return __fmaf_rn(__fmaf_rn(-d, q, x), r, q);
}

// Philox4x32-10 (Salmon et al., SC'11): counter-based, so the dither of sample i of utterance u is a pure function of
// (seed, u, i) -- no state, no host-drawn noise buffer crossing PCIe, the same value whichever thread asks for it.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// U[0,1) with torch.rand's float32 resolution (24 random bits * 2^-24), keyed by (seed, utterance, sample)
__device__ __forceinline__ float dither_uniform(unsigned long long seed, int utt, long long i) {
  const unsigned long long blk = static_cast<unsigned long long>(i) >> 2;
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(blk), static_cast<uint32_t>(blk >> 32), static_cast<uint32_t>(utt), 0u),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  const int e = static_cast<int>(i & 3);
  const uint32_t v = e == 0 ? r.x : e == 1 ? r.y : e == 2 ? r.z : r.w;
  return static_cast<float>(v >> 8) * 5.9604644775390625e-08f;
}

// the same stream for the even/odd pair (i, i + 1), i even: one Philox block covers both
__device__ __forceinline__ float2 dither_pair(unsigned long long seed, int utt, long long i) {
  const unsigned long long blk = static_cast<unsigned long long>(i) >> 2;
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(blk), static_cast<uint32_t>(blk >> 32), static_cast<uint32_t>(utt), 0u),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  const bool hi = (i & 2) != 0;
  return make_float2(static_cast<float>((hi ? r.z : r.x) >> 8) * 5.9604644775390625e-08f,
                     static_cast<float>((hi ? r.w : r.y) >> 8) * 5.9604644775390625e-08f);
}

// ------------------------------------------------------------------------------------------------
// Per-utterance second stage, run by the CTA that completed the utterance's statistics, on rows that are still in L2:
// per-utterance CMVN (x - mean) / (std + 1e-9) or AmplitudeToDB's clamp at max - top_db, then the SpecAugment zero-fill.
// kind: 1 = CMVN, 2 = top_db clamp.  Fast path (80 dims, 16-byte aligned rows): thread i owns float4 i, i + 128, ... of
// a 32-row block; 32 rows x 20 float4 = 640 = 5 x 128, so the five columns a thread meets repeat block after block and
// their constants live in registers.
// ------------------------------------------------------------------------------------------------
// Rewrites rows [r0, r0 + rows) of utterance utt in place, ONE WARP.  kind 1: (x - mean) * inv + lo per dim (constants
// in shared memory, float4 triples per 4 dims), kind 2: max(x, db_floor); then the SpecAugment zero-fill.  Fast path (80
// dims, 16-byte aligned rows): lane l owns float4 l, l + 32, ... of the block, 8 loads in flight per lane.
__device__ __forceinline__ void apply_rows_warp(float* __restrict__ base, long long out_ld, int n_out, int nm, bool fast, int r0, int rows,
                                             int kind, float db_floor, const float4* __restrict__ sm_c4, const int* __restrict__ sm_masks) {
  // (everything by value: through a reference to the kernel's parameter block every store was preceded by a generic
  //  re-load of out_ld -- the stores might alias it -- and a round of 8 stores cost 8 dependent memory latencies)
  const int lane = threadIdx.x & 31;
  if (fast) {
    // Groups of 8 rows = 160 float4 = 5 per lane: lane l owns float4 l + 32 m (m = 0..4) of every group, i.e. five FIXED
    // (row offset, column) pairs -- no division in the loop, the frequency-mask verdicts are five 4-bit constants, the
    // normalisation constants five fixed shared-memory addresses.  Two groups (10 loads) are in flight per lane.
    int roff[5], col[5];
    unsigned fz = 0u;
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      const int i = lane + 32 * m;
      roff[m] = i / 20;
      col[m] = i - roff[m] * 20;
      for (int q = 0; q < nm; ++q) {
        const int f0 = sm_masks[4 * q + 2], f1 = sm_masks[4 * q + 3];
#pragma unroll
        for (int e = 0; e < 4; ++e) fz |= (4 * col[m] + e >= f0 && 4 * col[m] + e < f1) ? (1u << (4 * m + e)) : 0u;
      }
    }
    // time masks as (start, end) rows relative to this block; at most 4 are kept in registers
    int t0[4], t1[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      t0[q] = (q < nm) ? sm_masks[4 * q] - r0 : 0;
      t1[q] = (q < nm) ? sm_masks[4 * q + 1] - r0 : 0;
    }
    const long long ld4 = out_ld >> 2;
    float4* g = reinterpret_cast<float4*>(base);
    for (int rb = 0; rb < rows; rb += 16) {
      float4 x[10];
#pragma unroll
      for (int u = 0; u < 10; ++u) {
        const int row = rb + (u / 5) * 8 + roff[u % 5];
        if (row < rows) x[u] = g[row * ld4 + col[u % 5]];
      }
#pragma unroll
      for (int u = 0; u < 10; ++u) {
        const int m = u % 5;
        const int row = rb + (u / 5) * 8 + roff[m];
        if (row >= rows) continue;
        float4 v = x[u];
        if (kind == 1) {
          const float4 mean = sm_c4[col[m]], inv = sm_c4[20 + col[m]], lo = sm_c4[40 + col[m]];
          v.x = fmaf(v.x - mean.x, inv.x, lo.x);      // (x - mean_hi) * inv - mean_lo * inv
          v.y = fmaf(v.y - mean.y, inv.y, lo.y);
          v.z = fmaf(v.z - mean.z, inv.z, lo.z);
          v.w = fmaf(v.w - mean.w, inv.w, lo.w);
        } else {
          v.x = fmaxf(v.x, db_floor);
          v.y = fmaxf(v.y, db_floor);
          v.z = fmaxf(v.z, db_floor);
          v.w = fmaxf(v.w, db_floor);
        }
        bool zr = false;
#pragma unroll
        for (int q = 0; q < 4; ++q) zr |= (row >= t0[q] && row < t1[q]);
        for (int q = 4; q < nm; ++q) zr |= (r0 + row >= sm_masks[4 * q] && r0 + row < sm_masks[4 * q + 1]);
        const unsigned z = zr ? 0xfu : ((fz >> (4 * m)) & 0xfu);
        v.x = (z & 1u) ? 0.f : v.x;
        v.y = (z & 2u) ? 0.f : v.y;
        v.z = (z & 4u) ? 0.f : v.z;
        v.w = (z & 8u) ? 0.f : v.w;
        g[row * ld4 + col[m]] = v;
      }
    }
  } else {
    const float* sm_c = reinterpret_cast<const float*>(sm_c4);     // mean[80] | inv[80] | lo[80]
    const int total = rows * n_out;
    for (int i = lane; i < total; i += 32) {
      const int rw = i / n_out;
      const int d = i - rw * n_out;
      float* p = base + static_cast<long long>(rw) * out_ld + d;
      float xv = *p;
      if (kind == 1) xv = fmaf(xv - sm_c[d], sm_c[kMaxMels + d], sm_c[2 * kMaxMels + d]);
      else xv = fmaxf(xv, db_floor);
      const int tf = r0 + rw;
      bool z = false;
      for (int q = 0; q < nm; ++q)
        z |= (tf >= sm_masks[4 * q] && tf < sm_masks[4 * q + 1]) || (d >= sm_masks[4 * q + 2] && d < sm_masks[4 * q + 3]);
      *p = z ? 0.f : xv;
    }
  }
}
__device__ __forceinline__ int ld_volatile_i(const int* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void st_volatile_i(int* p, int v) {
  asm volatile("st.volatile.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------------
// the fused front-end kernel
// ------------------------------------------------------------------------------------------------
template <typename TIn, bool kMfcc, int kStdMel>
__global__ void __launch_bounds__(kThreads, 4) fbank_kernel(const __grid_constant__ FbankParams P) {
  using L = SmemLayout<TIn, kMfcc>;
  constexpr int kScratchFloats = scratch_floats(kMfcc);
  constexpr int kValsOff = vals_off(kMfcc);
  extern __shared__ __align__(128) unsigned char smem[];
  TIn* const sm_in = reinterpret_cast<TIn*>(smem + L::off_in0);
  float* const sm_scratch = reinterpret_cast<float*>(smem + L::off_scratch);
  const float* const sm_window = reinterpret_cast<const float*>(smem + L::off_window);
  const float2* const sm_tw1 = reinterpret_cast<const float2*>(smem + L::off_tw1);
  const float2* const sm_tw2 = reinterpret_cast<const float2*>(smem + L::off_tw2);
  const int* const sm_k0 = reinterpret_cast<const int*>(smem + L::off_k0);
  float2* const sm_norm = reinterpret_cast<float2*>(smem + L::off_norm);
  float* const sm_lo = reinterpret_cast<float*>(smem + L::off_lo);
  double* const sm_acc = reinterpret_cast<double*>(smem + L::off_acc);
  float* const sm_wmax = reinterpret_cast<float*>(sm_acc + kWarps * 2 * kMaxMels + 2);   // [kWarps] max, then [kWarps] min
  int* const sm_masks = reinterpret_cast<int*>(smem + L::off_masks);
  uint64_t* const sm_bar = reinterpret_cast<uint64_t*>(smem + L::off_bar);
  int* const sm_arrivals = reinterpret_cast<int*>(smem + L::off_bar + 32);   // warps that have consumed the current tile's samples (running count)
  volatile int* const sm_ctl = reinterpret_cast<volatile int*>(smem + L::off_bar + 36);   // [0], [1] next span (ping-pong with the span slots), [2..6] service_item scratch, [5] also the exit flag
  Span* const sm_span = reinterpret_cast<Span*>(smem + L::off_span);
  float* const sm_melw = reinterpret_cast<float*>(smem + L::off_melw);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int t = lane & 15;       // lane inside the half-warp
  const int half = lane >> 4;    // which frame pair of the warp
  // the two unrolled variants are only selected for 80 mel bins, so without the DCT epilogue the output width is a
  // compile-time constant there (no per-band bound checks, immediate store offsets)
  const int n_out = (kStdMel != 0 && !kMfcc) ? 80 : P.n_out;
  const int mode = P.mode;
  const bool want_stats = !(LIDFE_ABL & 1024) && ((mode == 1) || (mode == 3));

  int taps[kBands], tap_off[kBands + 1];
  tap_off[0] = 0;
#pragma unroll
  for (int b = 0; b < kBands; ++b) {
    taps[b] = kStdMel ? std_taps(kStdMel, b) : P.band_taps[b];
    tap_off[b + 1] = tap_off[b] + taps[b];
  }
  float* const sm_dct = sm_melw + tap_off[kBands] * 32;
  float* const sm_lifter = sm_dct + (kMfcc ? ((P.n_mels * P.n_ceps + 3) & ~3) : 0);   // sections padded to 16 B

  // ---- one-time staging: all constant tables arrive with ONE TMA bulk copy while the CTA sets up the rest ------
  if (tid == 0) {
    mbar_init(&sm_bar[0], 1);            // full: samples of the staged tile have landed
    mbar_init(&sm_bar[2], 1);            // constant tables have landed
    *sm_arrivals = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&sm_bar[2], static_cast<uint32_t>(P.const_bytes));
    tma_bulk_g2s_plain(smem + L::off_window, P.const_blob, static_cast<uint32_t>(P.const_bytes), &sm_bar[2]);
  }
  // Work is claimed span by span in utterance-major order: the first span is this CTA's index, every later one comes
  // from a global counter (claimed one span ahead by thread 0, so the atomic's latency hides behind a whole span; the
  // claimed span's descriptor is fetched into the spare slot by cp.async).  Dynamic claims keep the CTAs balanced
  // when some of them stop to normalise an utterance.
  long long t_start = 0, n_sp = 0, n_blk = 0;
  if (P.dbg_buf && tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
  int span_idx = blockIdx.x * P.k_static;
  int pend = 0;                                                                   // thread 0: the span after this one
  if (tid < 2 && span_idx < P.n_spans) reinterpret_cast<int4*>(&sm_span[0])[tid] = __ldg(reinterpret_cast<const int4*>(P.spans + span_idx) + tid);
  // Every compute CTA first walks k_static consecutive spans of its own (neighbouring spans share their utterance, so
  // per-utterance sums leave the CTA about twice, and no atomics are spent), then claims the remaining ~15 % of the
  // spans one by one from a global counter, which keeps the CTAs balanced to the end.  (Service CTAs take none.)
  const int n_compute = (P.mode == 1 || P.mode == 4) ? P.n_compute : static_cast<int>(gridDim.x);
  const int n_static = n_compute * P.k_static;
  int k_mine = 1;                                                                 // thread 0: spans taken so far
  if (tid == 0 && static_cast<int>(blockIdx.x) < n_compute) pend = (P.k_static > 1) ? span_idx + 1 : n_static + atomicAdd(&P.sched[0], 1);
  for (int i = tid; i < kWarps * 2 * kMaxMels + 2; i += kThreads) sm_acc[i] = 0.0;
  if (tid < kWarps) { sm_wmax[tid] = -INFINITY; sm_wmax[kWarps + tid] = INFINITY; }
  if (mode == 2 && tid < n_out) {
    // finalise the all-reduced sums: mean, 1/(std + 1e-9) (unbiased)
    const double n = P.stats_in[2 * n_out];
    const double mean = P.stats_in[tid] / n;
    double var = (P.stats_in[n_out + tid] - P.stats_in[tid] * mean) / (n - 1.0);
    var = var > 0.0 ? var : 0.0;
    sm_norm[tid] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / (sqrt(var) + 1e-9)));
  }
  __syncthreads();   // mbarriers initialised, first span descriptor in place

  // Producer side of the sample staging.  There is ONE input buffer and no CTA-wide barrier per tile: every warp bumps a
  // shared arrival counter once it has consumed its samples (right after the framing stage, ~15 % into a tile); the
  // warp that arrives last knows the buffer is free and immediately stages the span's next tile (TMA bulk copy), so the
  // copy overlaps the remaining ~85 % of the current tile.  Consumers wait on the "full" mbarrier.  Every tile of a span
  // takes exactly kWarps arrivals (zero-fill rows are spans of their own, handled between two CTA barriers).
  const bool stage_masks = (mode == 0 || mode == 2) && P.n_masks > 0;
  auto stage_tile = [&](const Tile& tl) {          // called by exactly one warp per tile
    const int nsamp = kFrameShift * tl.nframes + (kFrameLen - kFrameShift);
    const TIn* src = reinterpret_cast<const TIn*>(P.wav) + tl.wav_off;
    TIn* dst = sm_in;
    if (__builtin_expect(tl.aux != 0, 1)) {
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t bytes = nsamp * (uint32_t)sizeof(TIn);
        mbar_expect_tx(&sm_bar[0], bytes);
        tma_bulk_g2s(dst, src, bytes, &sm_bar[0], l2_evict_first_policy());
      }
    } else {
      if (!P.center) {
        for (int i = lane; i < nsamp; i += 32) dst[i] = src[i];
      } else {
        // edge tile of the centred framing: index u of the constant-padded signal p (length L = N + 2 pad), mirrored
        // once at either end like torch.stft(center=True, pad_mode="reflect"); zeros inside the constant padding
        const long long N = P.utt_lengths[tl.utt];
        const TIn* x = reinterpret_cast<const TIn*>(P.wav) + P.utt_offsets[tl.utt];
        const long long Lp = N + 2 * P.pad;
        const long long u0 = static_cast<long long>(kFrameShift) * tl.t0 - (kFrameLen / 2);
        for (int i = lane; i < nsamp; i += 32) {
          long long u = u0 + i;
          u = u < 0 ? -u : (u >= Lp ? 2 * (Lp - 1) - u : u);
          const long long r = u - P.pad;
          dst[i] = (r >= 0 && r < N) ? x[r] : static_cast<TIn>(0);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm_bar[0]);         // element-wise staging: plain arrival completes the phase
    }
  };
  // one arrival per warp and tile; true for the warp that completes the round
  auto arrive_is_last = [&]() -> bool {
    int last = 0;
    if (lane == 0) last = ((atomicAdd(sm_arrivals, 1) % kWarps) == kWarps - 1);
    return __shfl_sync(0xffffffffu, last, 0) != 0;
  };

  uint32_t phase = 0u;
  float* const my_scratch = sm_scratch + (warp * 2 + half) * kScratchFloats;
  f2* const T_pl = reinterpret_cast<f2*>(my_scratch);      // the transpose plane
  f2* const my_P = reinterpret_cast<f2*>(my_scratch);
  const int partner = (lane & 16) | ((16 - t) & 15);
  const int up_lane = (lane & 16) | ((t - 1) & 15);

  mbar_wait(&sm_bar[2], 0u);
  int k0[kBands];
#pragma unroll
  for (int b = 0; b < kBands; ++b) k0[b] = sm_k0[t + 16 * b];

  double frames_acc = 0.0;    // mode 3: frames this CTA has added to its sums (thread kThreads - 1)
  int frames_held = 0, held_utt = -1;   // per-utterance modes: frames in the CTA's accumulators, and whose they are
  int cur = 0;                // which sm_span slot holds the current span

  // Second stage of the per-utterance modes (see FbankParams::items), one warp at a time, no CTA barrier anywhere: the
  // warp takes the item it claimed one item ago (and claims the next right away, so that the atomic's and the
  // descriptor's round trips hide behind the rows), waits for the utterance, finalises its constants from the
  // handed-over sums (fp64) into its own scratch, and rewrites the rows.  The warp that meets an utterance's FIRST item
  // also puts the other launch parity's bookkeeping of that utterance back to rest.
  const int par = P.parity & 1;
  int* const done_cur = P.utt_done + par * P.b_cap;
  unsigned* const max_cur = P.utt_max + par * P.b_cap;
  unsigned* const min_cur = P.utt_min + par * P.b_cap;
  double* const stats_cur = P.utt_stats + static_cast<long long>(par) * P.b_cap * 2 * n_out;
  auto service_warp = [&]() {
    float4* const c4 = reinterpret_cast<float4*>(sm_scratch + warp * 2 * kScratchFloats);   // [60] float4: mean | inv | lo
    float* const cf = reinterpret_cast<float*>(c4);
    int* const wmasks = reinterpret_cast<int*>(cf + 3 * kMaxMels);                            // [kMaxMasks][4]
    int nxt = 0;
    if (lane == 0) nxt = atomicAdd(&P.ictl[0], 1);
    for (;;) {
      int4 item = make_int4(-1, 0, 0, 0);
      if (lane == 0) {
        const int it = nxt;
        if (it < P.n_items) {
          nxt = atomicAdd(&P.ictl[0], 1);                        // the item after this one: in flight during the rows
          item = __ldg(&P.items[it]);
          while (ld_volatile_i(&done_cur[item.x]) != item.w) __nanosleep(100);
          __threadfence();                                       // acquire: the utterance's sums and rows
        }
      }
      item.x = __shfl_sync(0xffffffffu, item.x, 0);
      item.y = __shfl_sync(0xffffffffu, item.y, 0);
      item.z = __shfl_sync(0xffffffffu, item.z, 0);
      item.w = __shfl_sync(0xffffffffu, item.w, 0);
      if (item.x < 0) break;
      const int utt = item.x;
      float db_floor = 0.f;
      bool skip = false;
      if (mode == 1) {
        for (int d = lane; d < n_out; d += 32) {
          const double n = static_cast<double>(item.w);
          const double sm = __ldcg(stats_cur + (static_cast<long long>(utt) * 2 + 0) * n_out + d);
          const double ss = __ldcg(stats_cur + (static_cast<long long>(utt) * 2 + 1) * n_out + d);
          const double mu = sm / n;
          double var = (ss - sm * mu) / (n - 1.0);   // n == 1 -> NaN, as torch.std of one sample
          var = var > 0.0 ? var : (var == var ? 0.0 : var);
          const float mean = static_cast<float>(mu);
          cf[d] = mean;
          cf[kMaxMels + d] = static_cast<float>(1.0 / (sqrt(var) + 1e-9));
          cf[2 * kMaxMels + d] = static_cast<float>(-(mu - static_cast<double>(mean)) / (sqrt(var) + 1e-9));   // keeps x - mean accurate when std << |mean|
        }
      } else {
        const float mx = ord2f(__ldcg(&max_cur[utt])), mn = ord2f(__ldcg(&min_cur[utt]));
        db_floor = mx - P.top_db;
        skip = (mn >= db_floor) && P.n_masks == 0;      // nothing below max - top_db: the clamp is the identity
      }
      if (P.n_masks > 0 && lane < P.n_masks * 4) wmasks[lane] = P.masks[static_cast<long long>(utt) * P.n_masks * 4 + lane];
      if (item.y == 0) {   // first item of the utterance: the previous launch's half goes back to rest
        const int op = par ^ 1;
        double* os = P.utt_stats + (static_cast<long long>(op) * P.b_cap + utt) * 2 * n_out;
        for (int e = lane; e < 2 * n_out; e += 32) os[e] = 0.0;
        if (lane == 0) {
          P.utt_done[op * P.b_cap + utt] = 0;
          P.utt_max[op * P.b_cap + utt] = 0u;
          P.utt_min[op * P.b_cap + utt] = 0xffffffffu;
        }
      }
      __syncwarp();
      if (!skip && !(P.dbg & 1)) {
        const bool fast = (n_out == 80) && (P.out_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.out) & 15) == 0);
        apply_rows_warp(P.out + (P.utt_out_row[utt] + item.y) * P.out_ld, P.out_ld, n_out, P.n_masks, fast, item.y, item.z,
                        mode == 1 ? 1 : 2, db_floor, c4, wmasks);
      }
      __syncwarp();                                              // the constants are free again
      ++n_blk;
    }
  };
  const bool per_utt_mode = (mode == 1 || mode == 4);
  const bool service = per_utt_mode && static_cast<int>(blockIdx.x) >= P.n_compute;
  if (service || static_cast<int>(blockIdx.x) >= n_compute) span_idx = P.n_spans;
  while (span_idx < P.n_spans) {
    const Span sp = sm_span[cur];
    bool fetched = false;     // thread 0: the next span's descriptor copy has been issued
    auto fetch_next = [&]() {
      if (tid == 0 && !fetched) {
        fetched = true;
        if (pend < P.n_spans) {
          cp_async16(reinterpret_cast<int4*>(&sm_span[cur ^ 1]), reinterpret_cast<const int4*>(P.spans + pend));
          cp_async16(reinterpret_cast<int4*>(&sm_span[cur ^ 1]) + 1, reinterpret_cast<const int4*>(P.spans + pend) + 1);
        }
        cp_async_commit();
      }
    };

    if (__builtin_expect(sp.nframes == 0, 0)) {
      // zero-fill span: pad_sequence's zeros (ref: lid/raw_datasets.py:347-350)
      fetch_next();
      const long long total = P.ws_blocked ? 0 : static_cast<long long>(sp.aux) * n_out;   // (the DCT kernel fills them)
      for (long long i = tid; i < total; i += kThreads) {
        const long long r = i / n_out;
        const int d = static_cast<int>(i - r * n_out);
        P.out[(sp.out_row + r) * P.out_ld + d] = 0.f;
      }
    } else {
      // ---- per-span set-up: the utterance's mask table, normalize_wav constants, edge-tile geometry ----------
      unsigned dim_masked = 0u;   // bit b: output dim t+16b is inside a frequency mask of the span's utterance
      if (stage_masks) {
        if (tid < P.n_masks * 4) sm_masks[tid] = P.masks[static_cast<long long>(sp.utt) * P.n_masks * 4 + tid];
        __syncthreads();
        for (int q = 0; q < P.n_masks; ++q) {
          const int f0 = sm_masks[4 * q + 2], f1 = sm_masks[4 * q + 3];
#pragma unroll
          for (int b = 0; b < kBands; ++b) dim_masked |= (t + 16 * b >= f0 && t + 16 * b < f1) ? (1u << b) : 0u;
        }
      }
      const int* const wm = sm_masks;
      const bool wnorm = P.utt_wnorm != nullptr;
      float wn_mean = 0.f, wn_div = 1.f, wn_rcp = 1.f;
      if (wnorm) {
        const float2 wn = P.utt_wnorm[sp.utt];
        wn_mean = wn.x;
        wn_div = wn.y;
        wn_rcp = __frcp_rn(wn.y);
      }
      const long long utt_len = P.center ? P.utt_lengths[sp.utt] : 0;
      const long long tile0 = P.ws_blocked ? P.utt_first_tile[sp.utt] + sp.t0 / kTileFrames : 0;
      const int n_tiles = (sp.nframes + kTileFrames - 1) / kTileFrames;
      auto tile_of = [&](int ti) {
        Tile tl;
        tl.wav_off = sp.wav_off + static_cast<long long>(ti) * (kTileFrames * kFrameShift);
        tl.out_row = sp.out_row + ti * kTileFrames;
        tl.nframes = min(kTileFrames, sp.nframes - ti * kTileFrames);
        tl.utt = sp.utt;
        tl.t0 = sp.t0 + ti * kTileFrames;
        tl.aux = sp.aux;
        if (P.center) {    // tiles that touch the constant padding / the reflection are staged element-wise
          const long long rel = static_cast<long long>(tl.t0) * kFrameShift - kFrameLen / 2 - P.pad;
          const long long nsamp = static_cast<long long>(kFrameShift) * tl.nframes + (kFrameLen - kFrameShift);
          if (!(rel >= 0 && rel + nsamp <= utt_len)) tl.aux = 0;
        }
        return tl;
      };
      if (warp == 0 && (!(LIDFE_ABL & 1) || n_sp == 0)) stage_tile(tile_of(0));

      for (int ti = 0; ti < n_tiles; ++ti) {
        const Tile tl = tile_of(ti);
        const long long tile_idx = tile0 + ti;
        // consumer side: wait for this tile's samples
        if (!(LIDFE_ABL & 1) || (n_sp == 0 && ti == 0)) {
          mbar_wait(&sm_bar[0], phase);
          phase ^= 1u;
        }
        const TIn* in = sm_in;

        f2 val[kBands];
        const int flA = warp * 4 + half * 2;          // frame A inside the tile; B = A + 1
        const bool actA = flA < tl.nframes, actB = flA + 1 < tl.nframes;
        {
          // Frame B starts 160 samples = 5 x 32 after frame A, so lane t's element j of B is its element j+5 of A:
          // 18 loads cover both frames.  A pair whose frame B lies beyond the utterance (odd tail) still reads valid
          // shared memory; an entirely dead pair recomputes frame 0.  Neither is stored nor counted.
          const TIn* fr = in + kFrameShift * (actA ? flA : 0);

          // ---- load, DC removal, pre-emphasis, window (ta: compliance/kaldi.py:183-204) ---------------
          f2 R[16], I[16];     // R[j] = (re_A, re_B), I[j] = (im_A, im_B) of z[t + 16 j]
          {
            float2 x[18];
    #pragma unroll
            for (int j = 0; j < 18; ++j) {
              const int n = t + 16 * j;
              if (LIDFE_ABL & 8) x[j] = make_float2(__int_as_float(0x3f000000 + ((tl.t0 + n) << 8)), __int_as_float(0x3f800000 - ((tl.t0 + j) << 9)));
              else x[j] = (j < 17 || t < 8) ? InTraits<TIn>::ld2(fr + 2 * n, P.in_scale) : make_float2(0.f, 0.f);
            }
            if (wnorm) {   // normalize_wav while loading (row f2): (x - mean) / (std + 1e-6), the division as in wave_stages_kernel
#pragma unroll
              for (int j = 0; j < 18; ++j) {
                x[j].x = div_by(__fsub_rn(x[j].x, wn_mean), wn_div, wn_rcp);
                x[j].y = div_by(__fsub_rn(x[j].y, wn_mean), wn_div, wn_rcp);
              }
            }
            if (P.dither != 0.f) {   // in-kernel dither: every frame that loads sample i adds the same noise to it
              const long long s0 = static_cast<long long>(kFrameShift) * (tl.t0 + (actA ? flA : 0));
#pragma unroll
              for (int j = 0; j < 18; ++j) {
                const float2 u = dither_pair(P.seed, tl.utt, s0 + 2 * (t + 16 * j));
                x[j].x = __fadd_rn(x[j].x, __fmul_rn(P.dither, u.x));
                x[j].y = __fadd_rn(x[j].y, __fmul_rn(P.dither, u.y));
              }
            }
            float mA = 0.f, mB = 0.f;
            if (!(LIDFE_ABL & 2048) && ((kStdMel == 1 && !LIDFE_UNIT_SHORTCUT) || (kStdMel == 0 && P.remove_dc))) {
              f2 sA = make_float2(0.f, 0.f), sB = make_float2(0.f, 0.f);
    #pragma unroll
              for (int j = 0; j < 13; ++j) {
                if (j < 12 || t < 8) {
                  sA = add2(sA, x[j]);
                  sB = add2(sB, x[j + 5]);
                }
              }
              f2 sum = make_float2(sA.x + sA.y, sB.x + sB.y);
    #pragma unroll
              for (int o = 8; o >= 1; o >>= 1) {
                sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o);
                sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
              }
              mA = __fdiv_rn(sum.x, static_cast<float>(kFrameLen));
              mB = __fdiv_rn(sum.y, static_cast<float>(kFrameLen));
            }
            const float c = P.preemph;
            // One pass over the 13 sample pairs: (A,B) pairs are formed by the mean subtraction itself (scalar FADDs write
            // straight into the pair halves); x[2n-1] - mean lives in lane t-1 (same j), lane 0 takes lane 15's value of
            // step j-1, and the very first sample of the frame replicates itself (ta: compliance/kaldi.py:193-198).
            // x[j] - c * x[j-1]: product and difference rounded separately, as the reference's two tensor ops; with the
            // reference's c == 1.0 the product is exact, so that (warp-uniform) variant skips the multiplies.
            auto frame_pass = [&](auto unit_tag) {
              constexpr bool kUnit = decltype(unit_tag)::value;
              f2 to_prev = make_float2(0.f, 0.f);
    #pragma unroll
              for (int j = 0; j < 13; ++j) {
                const int n = t + 16 * j;
                const float2 w = (LIDFE_ABL & 256) ? make_float2(0.5f, 0.25f) : *reinterpret_cast<const float2*>(sm_window + 2 * n);
                const f2 te = make_float2(__fsub_rn(x[j].x, mA), __fsub_rn(x[j + 5].x, mB));   // x[2n]   - mean
                const f2 to = make_float2(__fsub_rn(x[j].y, mA), __fsub_rn(x[j + 5].y, mB));   // x[2n+1] - mean
                const f2 send = (t == 15) ? to_prev : to;
                f2 tp;
                if (LIDFE_ABL & 128) tp = send;
                else {
                  tp.x = __shfl_sync(0xffffffffu, send.x, up_lane);
                  tp.y = __shfl_sync(0xffffffffu, send.y, up_lane);
                }
                if (j == 0 && t == 0) tp = te;
                to_prev = to;
                const f2 se = kUnit ? sub2(te, tp) : sub2(te, mul2(tp, bc(c)));
                const f2 so = kUnit ? sub2(to, te) : sub2(to, mul2(te, bc(c)));
                R[j] = mul2(se, bc(w.x));
                I[j] = mul2(so, bc(w.y));
              }
            };
            // The framing flavour is fixed per kernel variant (kStdMel: 1 = the reference's Kaldi call, DC removal +
            // coefficient 1.0; 2 = its torch.stft call, window only; 0 = anything else), so that every instantiation
            // carries one copy of this loop: the tile body has to stay inside the 32 KB instruction cache.
            if (kStdMel == 2) {
              // window only; the (A,B) pairs are formed by the scalar multiplies themselves
    #pragma unroll
              for (int j = 0; j < 13; ++j) {
                const float2 w = *reinterpret_cast<const float2*>(sm_window + 2 * (t + 16 * j));
                R[j] = make_float2(__fmul_rn(x[j].x, w.x), __fmul_rn(x[j + 5].x, w.x));
                I[j] = make_float2(__fmul_rn(x[j].y, w.y), __fmul_rn(x[j + 5].y, w.y));
              }
            } else if (kStdMel == 1 && LIDFE_UNIT_SHORTCUT) {
              // the reference's call: (x[n] - m) - (x[n-1] - m) = x[n] - x[n-1] rounded once, y[0] = 0 -- the frame mean only
              // enters through round-off (see lidfe_fbank_warp.cuh, same arithmetic: both kernels give the same bits)
              float pA = 0.f, pB = 0.f;
    #pragma unroll
              for (int j = 0; j < 13; ++j) {
                const int n = t + 16 * j;
                const float2 w = *reinterpret_cast<const float2*>(sm_window + 2 * n);
                const float sA = (t == 15) ? pA : x[j].y, sB = (t == 15) ? pB : x[j + 5].y;
                float qA = __shfl_sync(0xffffffffu, sA, up_lane);
                float qB = __shfl_sync(0xffffffffu, sB, up_lane);
                if (j == 0 && t == 0) { qA = x[0].x; qB = x[5].x; }
                pA = x[j].y;
                pB = x[j + 5].y;
                const f2 se = make_float2(__fsub_rn(x[j].x, qA), __fsub_rn(x[j + 5].x, qB));
                const f2 so = make_float2(__fsub_rn(x[j].y, x[j].x), __fsub_rn(x[j + 5].y, x[j + 5].x));
                R[j] = mul2(se, bc(w.x));
                I[j] = mul2(so, bc(w.y));
              }
            } else if (kStdMel == 1) {
              frame_pass(std::true_type{});
            } else {
              frame_pass(std::false_type{});
            }
            if (t >= 8) R[12] = I[12] = make_float2(0.f, 0.f);
            R[13] = R[14] = R[15] = I[13] = I[14] = I[15] = make_float2(0.f, 0.f);
          }

          // ---- stage 1: 16-point DFT over j, twiddle W256^(K1*t), transpose through shared ------------
          if (!(LIDFE_ABL & 64)) fft16<true>(R, I);
          __syncwarp();   // previous tile's readers of this scratch are done; every lane has consumed its samples
          if (!(LIDFE_ABL & 1) && arrive_is_last() && ti + 1 < n_tiles) stage_tile(tile_of(ti + 1));   // the input buffer is free: next tile's TMA overlaps the rest
          if (ti == 0) fetch_next();
    #pragma unroll
          for (int p = 0; p < 16; ++p) {
            const int K1 = rev4(p);
            if (K1 != 0) {
              const float2 w = (LIDFE_ABL & 512) ? make_float2(0.6f, 0.8f) : sm_tw1[K1 * 16 + t];
              cmul2(R[p], I[p], w.x, w.y);
            }
            if (!(LIDFE_ABL & 2)) T_pl[K1 * kRowStride + t] = R[p];
          }
          if (!(LIDFE_ABL & 2)) {
          __syncwarp();
    #pragma unroll
          for (int q = 0; q < 8; ++q) {                      // real parts back, transposed
            const float4 a = *reinterpret_cast<const float4*>(T_pl + t * kRowStride + 2 * q);
            R[2 * q] = make_float2(a.x, a.y);
            R[2 * q + 1] = make_float2(a.z, a.w);
          }
          __syncwarp();   // one plane: the imaginary parts go through it after the real parts have been read back
    #pragma unroll
          for (int p = 0; p < 16; ++p) T_pl[rev4(p) * kRowStride + t] = I[p];
          __syncwarp();
    #pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b = *reinterpret_cast<const float4*>(T_pl + t * kRowStride + 2 * q);
            I[2 * q] = make_float2(b.x, b.y);
            I[2 * q + 1] = make_float2(b.z, b.w);
          }
          }   // LIDFE_ABL & 2
          // ---- stage 2: position p holds Z[t + 16*rev4(p)] ----------------------------------------------
          if (!(LIDFE_ABL & 64)) fft16<false>(R, I);
          __syncwarp();   // all lanes finished reading the transpose planes before the power bins overwrite them

          // ---- real-FFT split + power; lane t pairs with lane 16-t ------------------------------------
          f2 abl_psum = make_float2(0.f, 0.f);
          if (t == 0 && !(LIDFE_ABL & 4)) my_P[128] = mul2(fma2(R[2], R[2], mul2(I[2], I[2])), bc(4.f));   // 4|Z[128]|^2, Z[128] at rev4(8)
    #pragma unroll
          for (int i = 0; i < 8; ++i) {
            // partner's Z[(16-t) + 16 (15-i)]; lane 0 pairs k=16i with 256-16i = its own Z[16 (16-i)], k=0 with itself
            const int ps = rev4(15 - i);
            f2 br, bi;
            if (LIDFE_ABL & 32) { br = R[ps]; bi = I[ps]; }
            else {
              br.x = __shfl_sync(0xffffffffu, R[ps].x, partner);
              br.y = __shfl_sync(0xffffffffu, R[ps].y, partner);
              bi.x = __shfl_sync(0xffffffffu, I[ps].x, partner);
              bi.y = __shfl_sync(0xffffffffu, I[ps].y, partner);
            }
            if (t == 0) {
              const int own = (i == 0) ? 0 : rev4(16 - i);
              br = R[own];
              bi = I[own];
            }
            const f2 ar = R[rev4(i)], ai = I[rev4(i)];
            const f2 e2r = add2(ar, br), e2i = sub2(ai, bi);          // 2E = a + conj(b)
            f2 o2r = add2(ai, bi), o2i = sub2(br, ar);                // 2O = -i (a - conj(b))
            const float2 w = sm_tw2[i * 16 + t];
            cmul2(o2r, o2i, w.x, w.y);                                // W512^k * 2O
            const f2 xar = add2(e2r, o2r), xai = add2(e2i, o2i);      // 2 X[k]
            const f2 xbr = sub2(e2r, o2r), xbi = sub2(e2i, o2i);      // 2 conj(X[256-k])
            const int k = t + 16 * i;
            if (LIDFE_ABL & 4) {
              abl_psum = add2(abl_psum, add2(fma2(xar, xar, mul2(xai, xai)), fma2(xbr, xbr, mul2(xbi, xbi))));
            } else {
              my_P[k] = fma2(xar, xar, mul2(xai, xai));                 // 4 |X[k]|^2   (the 1/4 lives in the mel weights)
              my_P[256 - k] = fma2(xbr, xbr, mul2(xbi, xbi));
            }
          }
          __syncwarp();

          // ---- sparse triangular mel + log (ta: compliance/kaldi.py:621-633) ---------------------------
          // Segment form: this lane walks the bins between the centres of filters d = t + 16 b and d + 1 once, with the
          // down-slope weight of its own filter (wa) and the up-slope weight of the next one (wb); the wb sum travels one
          // lane up (lane 15's goes to lane 0 of the next band).
          if (LIDFE_ABL & 4) {
#pragma unroll
            for (int b = 0; b < kBands; ++b) val[b] = make_float2(abl_psum.x + b, abl_psum.y - b);
          } else {
            f2 carry15 = make_float2(0.f, 0.f);                   // lane 0: what lane 15 accumulated for it in the last band
            const int src = (lane & 16) | ((t - 1) & 15);
    #pragma unroll
            for (int b = 0; b < kBands; ++b) {
              f2 own = make_float2(0.f, 0.f), nxt = make_float2(0.f, 0.f);
              const f2* pp = my_P + k0[b];
              const float2* wp = reinterpret_cast<const float2*>(sm_melw) + tap_off[b] * 16 + t;
              if (kStdMel) {
    #pragma unroll
                for (int i = 0; i < std_taps(kStdMel, b); ++i) {
                  const f2 p = pp[i];
                  const float2 w = wp[i * 16];
                  own = fma2(p, bc(w.x), own);
                  nxt = fma2(p, bc(w.y), nxt);
                }
              } else {
    #pragma unroll 2
                for (int i = 0; i < taps[b]; ++i) {
                  const f2 p = pp[i];
                  const float2 w = wp[i * 16];
                  own = fma2(p, bc(w.x), own);
                  nxt = fma2(p, bc(w.y), nxt);
                }
              }
              f2 got;
              got.x = __shfl_sync(0xffffffffu, nxt.x, src);
              got.y = __shfl_sync(0xffffffffu, nxt.y, src);
              const f2 acc = add2(own, t == 0 ? carry15 : got);
              carry15 = got;
              // lg2.approx (abs. error ~1e-7 in the log) except at the floor, where the reference's log(eps) is returned exactly
              val[b] = make_float2(acc.x <= P.log_floor ? P.log_of_floor : log2_scaled(acc.x, P.log_scale),
                                   acc.y <= P.log_floor ? P.log_of_floor : log2_scaled(acc.y, P.log_scale));
            }
          }

          // ---- MFCC: DCT-II + lifter (ta: compliance/kaldi.py:648-666,786-796) --------------------------
          if (kMfcc) {
            f2* my_L = reinterpret_cast<f2*>(my_scratch + kLogmelOff);
    #pragma unroll
            for (int b = 0; b < kBands; ++b)
              if (t + 16 * b < P.n_mels) my_L[t + 16 * b] = val[b];
            __syncwarp();
            f2 acc[kBands];
    #pragma unroll
            for (int b = 0; b < kBands; ++b) acc[b] = make_float2(0.f, 0.f);
            const int nc = P.n_ceps;
    #pragma unroll 4
            for (int m = 0; m < P.n_mels; ++m) {
              const f2 l = my_L[m];
              const float* drow = sm_dct + m * nc + t;
    #pragma unroll
              for (int b = 0; b < kBands; ++b)
                if (16 * b < nc) acc[b] = fma2(l, bc((t + 16 * b < nc) ? drow[16 * b] : 0.f), acc[b]);
            }
    #pragma unroll
            for (int b = 0; b < kBands; ++b) val[b] = mul2(acc[b], bc((t + 16 * b < nc) ? sm_lifter[t + 16 * b] : 0.f));
          }

          // ---- epilogue: global CMVN, SpecAugment zero-fill, store ---------------------------------------
          {
            const int tfA = tl.t0 + flA;   // frame index inside the utterance
            bool rowA = false, rowB = false;
            unsigned dm = 0u;
            if (mode == 0 || mode == 2) {
              dm = dim_masked;
              for (int q = 0; q < P.n_masks; ++q) {
                const int m0 = wm[4 * q], m1 = wm[4 * q + 1];
                rowA |= (tfA >= m0 && tfA < m1);
                rowB |= (tfA + 1 >= m0 && tfA + 1 < m1);
              }
            }
            if (P.ws_blocked) {
              // tile-blocked log-mel workspace: dim d = t + 16 b of frame r lives at ((d / 4) * 16 + r) * 4 + d % 4 inside
              // the tile's block (no masks / normalisation on this path: they belong to the DCT kernel's epilogue)
              float* o = P.out + static_cast<long long>(tile_idx) * (kTileFrames * n_out) + (t >> 2) * 64 + (t & 3) + flA * 4;
    #pragma unroll
              for (int b = 0; b < kBands; ++b) {
                if (t + 16 * b < n_out) {
                  if (actA) o[256 * b] = val[b].x;
                  if (actB) o[256 * b + 4] = val[b].y;
                }
              }
            } else if (mode != 2 && !(mode == 0 && P.n_masks > 0) && tl.nframes == kTileFrames) {
              // full tile, nothing to apply here (no masks given, or a statistics mode: masks and normalisation happen in
              // the second pass)
              float* orow = P.out + (tl.out_row + flA) * P.out_ld + t;
    #pragma unroll
              for (int b = 0; b < kBands; ++b) {
                if (t + 16 * b < n_out && (!(LIDFE_ABL & 16) || val[b].x == 123.456f)) {
                  orow[16 * b] = val[b].x;
                  orow[P.out_ld + 16 * b] = val[b].y;
                }
              }
            } else {
              float* orow = P.out + (tl.out_row + flA) * P.out_ld + t;
    #pragma unroll
              for (int b = 0; b < kBands; ++b) {
                const int d = t + 16 * b;
                if (d < n_out) {
                  f2 x = val[b];
                  if (mode == 2) {
                    const float2 nm = sm_norm[d];
                    x = mul2(sub2(x, bc(nm.x)), bc(nm.y));
                  }
                  const bool dz = (dm >> b) & 1u;
                  if ((LIDFE_ABL & 16) && x.x != 123.456f) continue;
                  if (actA) orow[16 * b] = (dz || rowA) ? 0.f : x.x;
                  if (actB) orow[P.out_ld + 16 * b] = (dz || rowB) ? 0.f : x.y;
                }
              }
            }
          }
        }

        // ---- statistics.  Each warp parks the (A,B) feature pairs of its 4 frames in its own scratch, re-reads them
        //      dim-major and adds sum and sum of squares, in fp64 (x*x is exact there), to ITS OWN accumulators in
        //      shared memory (lane-owned dims: no atomics, no cross-warp traffic per tile).  The 4 warps' accumulators
        //      are combined when the span ends (per-utterance CMVN) or when the CTA runs out of work (global sums).
        if (want_stats) {
          f2* my_V = reinterpret_cast<f2*>(my_scratch + kValsOff);
#pragma unroll
          for (int b = 0; b < kBands; ++b)
            if (t + 16 * b < n_out) my_V[t + 16 * b] = val[b];
          __syncwarp();
          double* wacc = sm_acc + warp * 2 * kMaxMels;
          const float* v0 = sm_scratch + (warp * 2 + 0) * kScratchFloats + kValsOff;
          const float* v1 = v0 + kScratchFloats;
          const int f0 = warp * 4;
          for (int d = lane; d < n_out; d += 32) {
            const f2 p0 = *reinterpret_cast<const f2*>(v0 + 2 * d), p1 = *reinterpret_cast<const f2*>(v1 + 2 * d);
            double a1 = wacc[d], a2 = wacc[kMaxMels + d];
            if (f0 + 3 < tl.nframes) {   // all four frames live (every tile but an utterance's last): no predicates
              const double x0 = p0.x, x1 = p0.y, x2 = p1.x, x3 = p1.y;       // short dependency chains: a tree per tile
              a1 += (x0 + x1) + (x2 + x3);
              a2 += fma(x1, x1, x0 * x0) + fma(x3, x3, x2 * x2);
            } else {
              if (f0 + 0 < tl.nframes) { const double x = p0.x; a1 += x; a2 = fma(x, x, a2); }
              if (f0 + 1 < tl.nframes) { const double x = p0.y; a1 += x; a2 = fma(x, x, a2); }
              if (f0 + 2 < tl.nframes) { const double x = p1.x; a1 += x; a2 = fma(x, x, a2); }
            }
            wacc[d] = a1;
            wacc[kMaxMels + d] = a2;
          }
        }
        if (mode == 4) {   // AmplitudeToDB(top_db): running max / min of this warp's live features (ta: functional/functional.py:391-403)
          float m = -INFINITY, mn = INFINITY;
#pragma unroll
          for (int b = 0; b < kBands; ++b)
            if (t + 16 * b < n_out) {
              if (actA) { m = fmaxf(m, val[b].x); mn = fminf(mn, val[b].x); }
              if (actB) { m = fmaxf(m, val[b].y); mn = fminf(mn, val[b].y); }
            }
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
          }
          if (lane == 0) {
            sm_wmax[warp] = fmaxf(sm_wmax[warp], m);
            sm_wmax[kWarps + warp] = fminf(sm_wmax[kWarps + warp], mn);
          }
        }
      }   // tiles of the span
      if (mode == 3 && tid == kThreads - 1) frames_acc += static_cast<double>(sp.nframes);
    }

    // ---- end of span.  Thread 0 publishes the next span (claimed one span ago; its descriptor sits in the spare
    //      slot) and claims the one after.  For per-utterance CMVN / top_db the span's statistics are handed over to
    //      global memory; the CTA whose hand-over completes the utterance normalises it on the spot. ------------------
    if (tid == 0) {
      cp_async_wait<0>();
      sm_ctl[cur] = pend;
    }
    __syncthreads();   // every warp is done with the span: statistics complete, descriptor + claim visible
    // per-utterance modes: sums / extrema leave the CTA when its next span belongs to another utterance (or is a
    // zero-fill span, or there is none)
    if (sp.nframes != 0) frames_held += sp.nframes;
    bool hand_over = false;
    if ((mode == 1 || mode == 4) && frames_held > 0) {
      const int nxt_idx = sm_ctl[cur];
      hand_over = nxt_idx >= P.n_spans;
      if (!hand_over) {
        const Span nx = sm_span[cur ^ 1];
        hand_over = (nx.nframes == 0) || (nx.utt != held_utt);
      }
    }
    if (sp.nframes != 0) held_utt = sp.utt;
    if (hand_over) {
      if (mode == 1) {
        for (int e = tid; e < 2 * kMaxMels; e += kThreads) {
          const int which = e / kMaxMels, d = e - which * kMaxMels;
          double a = 0.0;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) {
            a += sm_acc[w * 2 * kMaxMels + e];
            sm_acc[w * 2 * kMaxMels + e] = 0.0;
          }
          if (d < n_out && a != 0.0 && !(P.dbg & 2)) atomicAdd(&stats_cur[(static_cast<long long>(held_utt) * 2 + which) * n_out + d], a);
        }
      } else if (tid == 0) {
        float m = sm_wmax[0], mn = sm_wmax[kWarps];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) { m = fmaxf(m, sm_wmax[w]); mn = fminf(mn, sm_wmax[kWarps + w]); }
        atomicMax(&max_cur[held_utt], f2ord(m));
        atomicMin(&min_cur[held_utt], f2ord(mn));
#pragma unroll
        for (int w = 0; w < kWarps; ++w) { sm_wmax[w] = -INFINITY; sm_wmax[kWarps + w] = INFINITY; }
      }
      if (P.n_items > 0) {   // in-kernel second stage: announce the frames (a separate second kernel needs no count)
        __syncthreads();     // every thread's feature rows and sums are issued: thread 0's fence covers them (bar.sync + cumulativity)
        if (tid == 0) {
          if (!(P.dbg & 4)) __threadfence();
          atomicAdd(&done_cur[held_utt], frames_held);     // no return value needed: the service warps watch the count
        }
      } else {
        __syncthreads();     // the accumulators are zero again before any warp adds the next span
      }
      frames_held = 0;
    }
    ++n_sp;
    span_idx = sm_ctl[cur];
    if (tid == 0 && span_idx < P.n_spans) {
      ++k_mine;
      pend = (k_mine < P.k_static) ? span_idx + 1 : n_static + atomicAdd(&P.sched[0], 1);
    }
    cur ^= 1;
  }

  long long t_spans = 0, n_iter = 0;
  if (P.dbg_buf && tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_spans));
  // ---- out of spans (or a service CTA from the start): every warp takes items until all of them have been claimed ----
  if (per_utt_mode) {
    service_warp();
    __syncthreads();   // every warp of the CTA has made its last claim before the CTA counts itself out
  }

  // ---- out of work: global sums (mode 3) leave the CTA once; the last CTA to get here puts the claim counter back ----
  if (mode == 3) {
    __syncthreads();
    for (int e = tid; e < 2 * kMaxMels; e += kThreads) {
      const int which = e / kMaxMels, d = e - which * kMaxMels;
      double a = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) a += sm_acc[w * 2 * kMaxMels + e];
      if (d < n_out && a != 0.0) atomicAdd(&P.stats_out[which * n_out + d], a);
    }
    if (tid == kThreads - 1 && frames_acc != 0.0) atomicAdd(&P.stats_out[2 * n_out], frames_acc);
  }
  if (P.dbg_buf && tid == 0) {
    long long t_end;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
    long long* o = P.dbg_buf + 8 * blockIdx.x;
    o[0] = t_start; o[1] = t_spans; o[2] = t_end; o[3] = n_sp; o[4] = n_blk; o[5] = n_iter;
  }
  // the last CTA to get here puts the schedule (and the per-utterance bookkeeping) back to rest for the next launch
  if (tid == 0) sm_ctl[5] = (atomicAdd(&P.sched[1], 1) == static_cast<int>(gridDim.x) - 1) ? 1 : 0;
  __syncthreads();
  if (sm_ctl[5] && tid == 0) {
    P.ictl[0] = 0;
    P.sched[0] = 0;
    P.sched[1] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// CMVN apply + masks (second pass of per-utterance / global CMVN; standalone SpecAugment application).
// grid = (utterances, row chunks of kApplyRows): every CTA finalises its utterance's mean / inv-std once and
// streams 64 rows with 128-bit accesses, several loads in flight per thread.
// ------------------------------------------------------------------------------------------------
constexpr int kApplyRowsDefault = 192;   // rows of one utterance per CTA (tunable: LIDFE_APPLY_ROWS)

struct ApplyParams {
  float* feats;
  long long ld;
  int n_out;
  const int* masks;
  int n_masks;
  const double* utt_stats;       // [B][2][n_out] or NULL
  const long long* utt_frames;   // [B]
  const long long* utt_out_row;  // [B]
  const double* glob_stats;      // [2*n_out+1] or NULL
  int normalize;                 // 0 -> masks only, 1 -> (x - mean) * inv_std, 2 -> max(x, utt_max - top_db)
  int rows_per_cta;
  const unsigned* utt_max;       // [B] (normalize == 2)
  const unsigned* utt_min;       // [B] or NULL (normalize == 2): lets blocks of an utterance with nothing below the floor return at once
  float top_db;
  // the OTHER launch parity's bookkeeping of every utterance, put back to rest here (the two-kernel flavour of the
  // per-utterance modes; see FbankParams::items), or NULL
  double* rest_stats;            // [B][2][n_out]
  int* rest_done;                // [B]
  unsigned* rest_max;            // [B]
  unsigned* rest_min;            // [B]
};

__global__ void __launch_bounds__(256) cmvn_apply_kernel(const __grid_constant__ ApplyParams P) {
  __shared__ float2 s_norm[kMaxMels];       // (mean, 1/(std + 1e-9))
  __shared__ float s_lo[kMaxMels];          // -(mean - (float)mean) / (std + 1e-9): keeps x - mean accurate when std << |mean|
  __shared__ int s_masks[kMaxMasks * 4];
  const int tid = threadIdx.x;
  const int utt = blockIdx.x;
  const long long T = P.utt_frames[utt];
  const long long r0 = static_cast<long long>(blockIdx.y) * P.rows_per_cta;
  if (r0 >= T) return;
  const int rows = static_cast<int>(T - r0 < P.rows_per_cta ? T - r0 : P.rows_per_cta);
  float* base = P.feats + (P.utt_out_row[utt] + r0) * P.ld;
  const bool vec = (P.n_out % 4 == 0) && (P.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.feats) & 15) == 0);

  // Every thread owns ONE float4 column (4 output dims) and walks down the rows: its normalisation constants and its
  // frequency-mask verdict live in registers, so the per-element work is a load, 8 flops, a row test and a store.
  const int nvec = vec ? (P.n_out >> 2) : 1;
  const int rstep = 256 / nvec;                 // rows covered by one pass of the CTA
  const int c = tid % nvec;
  const int rlane = tid / nvec;
  const bool live = vec && rlane < rstep;
  constexpr int kU = 4;
  float4 x[kU];
  if (live) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int r = rlane + u * rstep;
      if (r < rows) x[u] = *reinterpret_cast<const float4*>(base + static_cast<long long>(r) * P.ld + 4 * c);
    }
  }

  const float db_floor = (P.normalize == 2) ? ord2f(P.utt_max[utt]) - P.top_db : 0.f;
  if (blockIdx.y == 0) {
    if (P.rest_stats && tid < 2 * P.n_out) P.rest_stats[static_cast<long long>(utt) * 2 * P.n_out + tid] = 0.0;
    if (P.rest_done && tid == 0) {
      P.rest_done[utt] = 0;
      P.rest_max[utt] = 0u;
      P.rest_min[utt] = 0xffffffffu;
    }
  }
  // AmplitudeToDB's clamp only moves values below max - top_db: an utterance that holds none needs no pass at all
  if (P.normalize == 2 && P.n_masks == 0 && P.utt_min && ord2f(P.utt_min[utt]) >= db_floor) return;
  if (tid < P.n_out) {
    float mean = 0.f, inv = 1.f, lo = 0.f;
    if (P.normalize == 1) {
      double n, sm, ss;
      if (P.utt_stats) {
        n = static_cast<double>(T);
        sm = P.utt_stats[(static_cast<long long>(utt) * 2 + 0) * P.n_out + tid];
        ss = P.utt_stats[(static_cast<long long>(utt) * 2 + 1) * P.n_out + tid];
      } else {
        n = P.glob_stats[2 * P.n_out];
        sm = P.glob_stats[tid];
        ss = P.glob_stats[P.n_out + tid];
      }
      const double mu = sm / n;
      double var = (ss - sm * mu) / (n - 1.0);   // n == 1 -> NaN, as torch.std of one sample
      var = var > 0.0 ? var : (var == var ? 0.0 : var);
      mean = static_cast<float>(mu);
      inv = static_cast<float>(1.0 / (sqrt(var) + 1e-9));
      lo = static_cast<float>(-(mu - static_cast<double>(mean)) / (sqrt(var) + 1e-9));
    }
    s_norm[tid] = make_float2(mean, inv);
    s_lo[tid] = lo;
  }
  if (tid < P.n_masks * 4) s_masks[tid] = P.masks[static_cast<long long>(utt) * P.n_masks * 4 + tid];
  __syncthreads();

  if (vec) {
    if (!live) return;
    const int d = 4 * c;
    const float2 n0 = s_norm[d], n1 = s_norm[d + 1], n2 = s_norm[d + 2], n3 = s_norm[d + 3];
    const float l0 = s_lo[d], l1 = s_lo[d + 1], l2 = s_lo[d + 2], l3 = s_lo[d + 3];
    bool z0 = false, z1 = false, z2 = false, z3 = false;
    for (int q = 0; q < P.n_masks; ++q) {
      const int f0 = s_masks[4 * q + 2], f1 = s_masks[4 * q + 3];
      z0 |= (d + 0 >= f0 && d + 0 < f1);
      z1 |= (d + 1 >= f0 && d + 1 < f1);
      z2 |= (d + 2 >= f0 && d + 2 < f1);
      z3 |= (d + 3 >= f0 && d + 3 < f1);
    }
    // time bounds of the first kRegMasks masks live in registers (every shipped config has <= 2 rows), the rest in shared
    // memory: registers decide this kernel's occupancy
    constexpr int kRegMasks = 4;
    int t0[kRegMasks], t1[kRegMasks];
#pragma unroll
    for (int q = 0; q < kRegMasks; ++q) {
      t0[q] = (q < P.n_masks) ? s_masks[4 * q] : 0;
      t1[q] = (q < P.n_masks) ? s_masks[4 * q + 1] : 0;
    }
    for (int rb = 0; rb < rows; rb += kU * rstep) {
      if (rb > 0) {
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int r = rb + rlane + u * rstep;
          if (r < rows) x[u] = *reinterpret_cast<const float4*>(base + static_cast<long long>(r) * P.ld + d);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int r = rb + rlane + u * rstep;
        if (r >= rows) continue;
        float4 v = x[u];
        if (P.normalize == 1) {
          v.x = fmaf(v.x - n0.x, n0.y, l0);      // (x - mean_hi) * inv - mean_lo * inv
          v.y = fmaf(v.y - n1.x, n1.y, l1);
          v.z = fmaf(v.z - n2.x, n2.y, l2);
          v.w = fmaf(v.w - n3.x, n3.y, l3);
        } else if (P.normalize == 2) {
          v.x = fmaxf(v.x, db_floor);
          v.y = fmaxf(v.y, db_floor);
          v.z = fmaxf(v.z, db_floor);
          v.w = fmaxf(v.w, db_floor);
        }
        const int tf = static_cast<int>(r0) + r;
        bool zr = false;
#pragma unroll
        for (int q = 0; q < kRegMasks; ++q) zr |= (tf >= t0[q] && tf < t1[q]);
        for (int q = kRegMasks; q < P.n_masks; ++q) zr |= (tf >= s_masks[4 * q] && tf < s_masks[4 * q + 1]);
        v.x = (zr || z0) ? 0.f : v.x;
        v.y = (zr || z1) ? 0.f : v.y;
        v.z = (zr || z2) ? 0.f : v.z;
        v.w = (zr || z3) ? 0.f : v.w;
        *reinterpret_cast<float4*>(base + static_cast<long long>(r) * P.ld + d) = v;
      }
    }
  } else {
    const int total = rows * P.n_out;
    for (int i = tid; i < total; i += 256) {
      const int rw = i / P.n_out;
      const int d = i - rw * P.n_out;
      float* p = base + static_cast<long long>(rw) * P.ld + d;
      float xv = *p;
      if (P.normalize == 1) xv = fmaf(xv - s_norm[d].x, s_norm[d].y, s_lo[d]);
      else if (P.normalize == 2) xv = fmaxf(xv, db_floor);
      const int tf = static_cast<int>(r0) + rw;
      bool z = false;
      for (int q = 0; q < P.n_masks; ++q)
        z |= (tf >= s_masks[4 * q] && tf < s_masks[4 * q + 1]) || (d >= s_masks[4 * q + 2] && d < s_masks[4 * q + 3]);
      *p = z ? 0.f : xv;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// MFCC as a second kernel: ceps = (log-mel @ DCT) * lifter  (ta: compliance/kaldi.py:648-666,786-796).
// The in-kernel DCT epilogue of fbank_kernel re-reads the 80x40 matrix from shared memory for every frame pair (+65 %
// shared-memory wavefronts on a kernel that is bound by exactly that pipe); when no statistics are needed (cmvn none /
// global apply) fbank_kernel writes the log-mels to a tile-blocked workspace (FbankParams::ws_blocked) and this FP32
// GEMM [rows x n_mels] . [n_mels x n_ceps] finishes the job.
//   * a warp works on up to 4 tiles (64 frames) at a time: lane l owns frame l % 16 of each of them and the column
//     half l / 16, i.e. a register tile of 4 rows x 20 columns (80 accumulators, packed FFMA2; 16 warps per SM).  A
//     DCT-row load is one 16-byte address per half-warp (5 LDS.128 feed 40 FFMA2), so the FMA pipe is the limit;
//   * the log-mels arrive through a private 4-stage cp.async ring (16 B per row and stage, 3 stages in flight).  In the
//     blocked layout the 16 lanes of a half-warp copy 256 contiguous bytes per row and stage (a row-major workspace
//     cost 28 shared-memory wavefronts per cp.async instruction instead of 4); every thread only reads what it copied
//     itself, so nothing is synchronised after the table load;
//   * tiles are dealt out evenly: every warp gets n_tiles / n_warps consecutive tiles, give or take one (a fixed 4-tile
//     unit would leave 1.35 units per warp on cfg3, i.e. two rounds where 1.35 are needed) and takes them 4, 2 and 1
//     at a time.
// Summation order is m = 0..n_mels-1 with fused multiply-adds, the same as a scalar loop.
// ------------------------------------------------------------------------------------------------
constexpr int kDctThreads = 128;
constexpr int kDctMaxTiles = 4;          // tiles (rows per thread) per pass
constexpr int kDctStages = 4;            // cp.async ring depth
constexpr int kDctMaxCeps = 40;          // n_ceps <= 40 takes this path
constexpr int kDctCols = kDctMaxCeps / 2;   // accumulator columns per thread (the other half lives 16 lanes away)

struct DctParams {
  const float* logmel;      // workspace [n_tiles][n_mels / 4][16][4]
  float* out;
  long long out_ld;
  const Tile* tiles;
  int n_tiles;
  int n_mels, n_ceps;       // n_mels % 4 == 0
  const float* dct;         // [n_mels][n_ceps] device copy
  const float* lifter;      // [n_ceps]
  const int* masks;
  int n_masks;
  int mode;                 // 0 none, 2 global apply
  const double* stats_in;   // [2*n_ceps+1] (mode 2)
};


// One pass: this thread's frame (row g) of TM consecutive tiles, columns c0 .. c0+19.  `ring` is the thread's slot of
// stage 0 / row 0; stage s, row i lives at ring[(s * kDctMaxTiles + i) * kDctThreads].
template <int TM>
__device__ __forceinline__ void dct_pass(const DctParams& P, int tile0, int g, int c0, const float* s_dct,
                                         const float* s_lift, const float2* s_norm, float4* ring, bool vec_out) {
  const int nm = P.n_mels, nc = P.n_ceps;
  const int nsteps = nm >> 2;
  const float* src = P.logmel + static_cast<long long>(tile0) * (kTileFrames * nm) + g * 4;   // + i * 16 nm + s * 64

  f2 acc[TM][kDctCols / 2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < kDctCols / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);
  s_dct += c0;                                          // this thread's column half

  auto issue = [&](int s) {
    if (s < nsteps) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
        cp_async16(ring + ((s % kDctStages) * kDctMaxTiles + i) * kDctThreads, src + i * (kTileFrames * nm) + s * 64);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < kDctStages - 1; ++s) issue(s);
#pragma unroll 1
  for (int s = 0; s < nsteps; ++s) {
    issue(s + kDctStages - 1);
    cp_async_wait<kDctStages - 1>();                    // step s has landed (only this thread reads it)
    float4 cur[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) cur[i] = ring[((s % kDctStages) * kDctMaxTiles + i) * kDctThreads];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4* d4 = reinterpret_cast<const float4*>(s_dct + (4 * s + kk) * kDctMaxCeps);
      f2 l[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) l[i] = bc(kk == 0 ? cur[i].x : kk == 1 ? cur[i].y : kk == 2 ? cur[i].z : cur[i].w);
#pragma unroll
      for (int j = 0; j < kDctCols / 4; ++j) {
        const float4 d = d4[j];
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          acc[i][2 * j] = fma2(l[i], make_float2(d.x, d.y), acc[i][2 * j]);
          acc[i][2 * j + 1] = fma2(l[i], make_float2(d.z, d.w), acc[i][2 * j + 1]);
        }
      }
    }
  }
  cp_async_wait<0>();

  // lifter, global CMVN, SpecAugment zero-fill, store
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const Tile tl = P.tiles[tile0 + i];
    if (g >= tl.nframes) continue;                      // short last tile of an utterance / zero-fill tile
    const int tf = tl.t0 + g;
    const int* mk = P.masks + static_cast<long long>(tl.utt) * P.n_masks * 4;
    unsigned cm = 0u;                                   // bit k: column c0 + k is zero-filled (time or frequency mask)
    for (int qm = 0; qm < P.n_masks; ++qm) {
      if (tf >= mk[4 * qm] && tf < mk[4 * qm + 1]) cm = 0xffffffffu;
      const int lo = max(mk[4 * qm + 2] - c0, 0), hi = min(mk[4 * qm + 3] - c0, kDctCols);
      if (hi > lo) cm |= ((1u << hi) - 1u) & ~((1u << lo) - 1u);
    }
    float* o = P.out + (tl.out_row + g) * P.out_ld + c0;
#pragma unroll
    for (int j = 0; j < kDctCols / 4; ++j) {
      float x[4] = {acc[i][2 * j].x, acc[i][2 * j].y, acc[i][2 * j + 1].x, acc[i][2 * j + 1].y};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = c0 + 4 * j + e;
        float v = __fmul_rn(x[e], s_lift[c]);
        if (P.mode == 2) v = (v - s_norm[c].x) * s_norm[c].y;
        x[e] = ((cm >> (4 * j + e)) & 1u) ? 0.f : v;
      }
      if (vec_out) {
        if (c0 + 4 * j < nc) *reinterpret_cast<float4*>(o + 4 * j) = make_float4(x[0], x[1], x[2], x[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c0 + 4 * j + e < nc) o[4 * j + e] = x[e];
      }
    }
  }
}

__global__ void __launch_bounds__(kDctThreads, 4) mfcc_dct_kernel(const __grid_constant__ DctParams P) {
  extern __shared__ __align__(16) float dsm[];
  const int nm = P.n_mels, nc = P.n_ceps;
  float4* s_ring = reinterpret_cast<float4*>(dsm);      // [kDctStages][kDctMaxTiles][kDctThreads] float4
  float* s_dct = dsm + kDctStages * kDctMaxTiles * kDctThreads * 4;   // [nm][kDctMaxCeps] zero padded
  float* s_lift = s_dct + nm * kDctMaxCeps;             // [kDctMaxCeps]
  float2* s_norm = reinterpret_cast<float2*>(s_lift + kDctMaxCeps);   // [kDctMaxCeps] (mean, inv)
  const int tid = threadIdx.x;
  if (nc == kDctMaxCeps && (reinterpret_cast<uintptr_t>(P.dct) & 15) == 0) {
    for (int i = tid; i < nm * (kDctMaxCeps / 4); i += kDctThreads)
      reinterpret_cast<float4*>(s_dct)[i] = __ldg(reinterpret_cast<const float4*>(P.dct) + i);
  } else {
    for (int i = tid; i < nm * kDctMaxCeps; i += kDctThreads) {
      const int m = i / kDctMaxCeps, c = i - m * kDctMaxCeps;
      s_dct[i] = c < nc ? P.dct[m * nc + c] : 0.f;
    }
  }
  if (tid < kDctMaxCeps) {
    s_lift[tid] = tid < nc ? P.lifter[tid] : 0.f;
    float mean = 0.f, inv = 1.f;
    if (P.mode == 2 && tid < nc) {
      const double n = P.stats_in[2 * nc];
      const double mu = P.stats_in[tid] / n;
      double var = (P.stats_in[nc + tid] - P.stats_in[tid] * mu) / (n - 1.0);
      var = var > 0.0 ? var : 0.0;
      mean = static_cast<float>(mu);
      inv = static_cast<float>(1.0 / (sqrt(var) + 1e-9));
    }
    s_norm[tid] = make_float2(mean, inv);
  }
  __syncthreads();

  const bool vec_out = (nc % 4 == 0) && (P.out_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.out) & 15) == 0);
  const long long gtid = static_cast<long long>(blockIdx.x) * kDctThreads + tid;
  const long long nthreads = static_cast<long long>(gridDim.x) * kDctThreads;

  // zero-fill tiles (pad_sequence's zeros): 4 threads per tile
  for (long long q = gtid; q < static_cast<long long>(P.n_tiles) * 4; q += nthreads) {
    const Tile tl = P.tiles[q >> 2];
    if (tl.nframes != 0) continue;
    for (int r = static_cast<int>(q & 3); r < tl.aux; r += 4) {
      float* o = P.out + (tl.out_row + r) * P.out_ld;
      if (vec_out) {
        for (int c = 0; c < nc; c += 4) *reinterpret_cast<float4*>(o + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        for (int c = 0; c < nc; ++c) o[c] = 0.f;
      }
    }
  }

  // frame tiles: the same number of consecutive tiles for every warp, taken 4 / 2 / 1 at a time
  const int lane = tid & 31;
  const int g = lane & 15, c0 = (lane >> 4) * kDctCols;
  const long long warp = gtid >> 5, nwarps = nthreads >> 5;
  const long long base = P.n_tiles / nwarps, rem = P.n_tiles - base * nwarps;   // the first `rem` warps take one more
  long long t = warp * base + (warp < rem ? warp : rem);
  const long long t_end = t + base + (warp < rem ? 1 : 0);
  float4* ring = s_ring + tid;
  while (t + 4 <= t_end) {
    dct_pass<4>(P, static_cast<int>(t), g, c0, s_dct, s_lift, s_norm, ring, vec_out);
    t += 4;
  }
  if (t + 2 <= t_end) {
    dct_pass<2>(P, static_cast<int>(t), g, c0, s_dct, s_lift, s_norm, ring, vec_out);
    t += 2;
  }
  if (t < t_end) dct_pass<1>(P, static_cast<int>(t), g, c0, s_dct, s_lift, s_norm, ring, vec_out);
}

// ------------------------------------------------------------------------------------------------
// waveform-level stages (ref: lid/audio_processor.py:108-115,129-134).  One 8-CTA cluster per utterance.
// ------------------------------------------------------------------------------------------------
struct WaveParams {
  const void* in;             // float32 or int16 samples
  int in_i16;                 // 1 -> int16 PCM, converted as (float)s * in_scale before anything else
  float in_scale;
  float* out;                 // NULL: statistics only (the fused load of fbank_kernel applies them, row f2)
  float2* norm_out;           // [B] (mean, std + 1e-6) of every utterance, or NULL
  const long long* offsets;   // [B]
  const long long* lengths;   // [B]
  int normalize;
  float dither;
  const float* noise;         // the U[0,1) draw (parity with the reference's torch.rand_like), or NULL: Philox below
  unsigned long long seed;
  float preemph;
};

// One thread-block CLUSTER of kWaveCluster CTAs per utterance: every CTA owns a contiguous slice of the samples, the
// partial sums of the two-pass mean / variance meet through distributed shared memory (each CTA reads its peers'
// partials after a cluster barrier and adds them in rank order, so all of them hold the same bits), and the slice is
// re-read from L2 for the elementwise pass.  (One CTA per utterance left 3/4 of the SMs idle on a 32-utterance chunk
// and serialised 128 000 fp64 adds per thread block.)
constexpr int kWaveCluster = 8;
constexpr int kWaveThreads = 512;

__global__ void __cluster_dims__(kWaveCluster, 1, 1) __launch_bounds__(kWaveThreads)
    wave_stages_kernel(const __grid_constant__ WaveParams P) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double s_red[kWaveThreads / 32];
  __shared__ double s_part[2];                           // this CTA's partial sum / partial sum of squared deviations
  const int u = blockIdx.x / kWaveCluster;
  const int rank = static_cast<int>(cluster.block_rank());
  const long long off = P.offsets[u], n = P.lengths[u];
  const float* xf = reinterpret_cast<const float*>(P.in) + off;
  const short* xs = reinterpret_cast<const short*>(P.in) + off;
  auto ld = [&](long long i) -> float { return P.in_i16 ? static_cast<float>(xs[i]) * P.in_scale : xf[i]; };
  const float* nz = P.noise ? P.noise + off : nullptr;
  float* y = P.out ? P.out + off : nullptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long per = (n + kWaveCluster - 1) / kWaveCluster;
  const long long lo = per * rank < n ? per * rank : n, hi = lo + per < n ? lo + per : n;

  // CTA-wide fp64 sum of `v`, left in s_part[slot]
  auto cta_sum = [&](double v, int slot) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kWaveThreads / 32; ++w) t += s_red[w];
      s_part[slot] = t;
    }
  };
  auto cluster_total = [&](int slot) -> double {
    cluster.sync();                                      // every CTA's partial is written (and visible cluster-wide)
    double t = 0.0;
    for (int r = 0; r < kWaveCluster; ++r) t += cluster.map_shared_rank(s_part, r)[slot];
    return t;
  };

  float mean = 0.f, div = 1.f, rcp = 1.f;
  if (P.normalize) {
    // two-pass mean / unbiased variance in fp64 (torch.std_mean accumulates in fp32 with a cascade)
    double s = 0.0;
    for (long long i = lo + tid; i < hi; i += kWaveThreads) s += ld(i);
    cta_sum(s, 0);
    const double m = cluster_total(0) / static_cast<double>(n);
    double q = 0.0;
    for (long long i = lo + tid; i < hi; i += kWaveThreads) {
      const double dlt = ld(i) - m;
      q += dlt * dlt;
    }
    cta_sum(q, 1);
    const double qt = cluster_total(1);
    mean = static_cast<float>(m);
    div = static_cast<float>(sqrt(qt / static_cast<double>(n - 1))) + 1e-6f;
    rcp = __frcp_rn(div);
    if (P.norm_out && rank == 0 && tid == 0) P.norm_out[u] = make_float2(mean, div);
  }
  if (y) {
    auto stage1 = [&](long long i) -> float {
      float v = ld(i);
      if (P.normalize) v = div_by(__fsub_rn(v, mean), div, rcp);
      if (P.dither != 0.f) v = __fadd_rn(v, __fmul_rn(P.dither, nz ? nz[i] : dither_uniform(P.seed, u, i)));
      return v;
    };
    for (long long i = lo + tid; i < hi; i += kWaveThreads) {
      float v = stage1(i);
      if (P.preemph != 0.f && i > 0) v = __fsub_rn(v, __fmul_rn(P.preemph, stage1(i - 1)));
      y[i] = v;
    }
  }
  cluster.sync();                                        // no CTA exits while a peer may still read its partials
}

// ------------------------------------------------------------------------------------------------
// Polyphase sinc resampler (row f4: the DataProcessor in front of the models, ref: lid/ConformerLangModel.py:131-169 ->
// ta: functional/functional.py _apply_sinc_resample_kernel).  With orig / new reduced by their gcd, output sample
// i * new + p is  sum_k W[p][k] * xpad[i * orig + k],  xpad = x with `width` zeros in front and width + orig behind:
// a GEMM [new x K] . [K x frames] whose right-hand columns are overlapping windows of the waveform.  One CTA computes
// 16 consecutive frames i of one utterance for all phases p: the 16 windows sit in shared memory (one row each, so a
// 4-tap LDS.128 is a pure broadcast), every thread owns one phase and 16 accumulators, and the weights stream from the
// L2-resident transposed table [K][new] with coalesced loads -- 128 FFMA per 16 LDS.128 and 4 LDG.
// ------------------------------------------------------------------------------------------------
constexpr int kRsFrames = 16;
constexpr int kRsThreads = 160;

struct ResampleParams {
  const float* in;              // packed waveforms
  const long long* in_off;      // [B]
  const long long* in_len;      // [B]
  float* out;
  const long long* out_off;     // [B]
  const long long* out_len;     // [B] samples to write (<= ceil(new * in_len / orig))
  const float* wt;              // [K4][new] transposed kernel, rows K..K4-1 zero
  int orig, nw, K, K4, width;
};

__global__ void __launch_bounds__(kRsThreads) resample_kernel(const __grid_constant__ ResampleParams P) {
  extern __shared__ __align__(16) float rs_x[];           // [kRsFrames][K4]
  const int b = blockIdx.y;
  const long long n_in = P.in_len[b], n_out = P.out_len[b];
  const long long f0 = static_cast<long long>(blockIdx.x) * kRsFrames;
  if (f0 * P.nw >= n_out) return;
  const float* x = P.in + P.in_off[b];
  const int tid = threadIdx.x;
  for (int i = tid; i < kRsFrames * P.K4; i += kRsThreads) {
    const int f = i / P.K4, k = i - f * P.K4;
    const long long j = (f0 + f) * P.orig + k - P.width;   // index into the unpadded waveform
    rs_x[i] = (k < P.K && j >= 0 && j < n_in) ? x[j] : 0.f;
  }
  __syncthreads();
  float* y = P.out + P.out_off[b];
  for (int pc = 0; pc < P.nw; pc += kRsThreads) {
    const int p = pc + tid;
    if (p >= P.nw) continue;
    float acc[kRsFrames];
#pragma unroll
    for (int f = 0; f < kRsFrames; ++f) acc[f] = 0.f;
    const float* w = P.wt + p;
#pragma unroll 1
    for (int k = 0; k < P.K4; k += 4) {
      const float w0 = __ldg(w + static_cast<long long>(k) * P.nw), w1 = __ldg(w + static_cast<long long>(k + 1) * P.nw);
      const float w2 = __ldg(w + static_cast<long long>(k + 2) * P.nw), w3 = __ldg(w + static_cast<long long>(k + 3) * P.nw);
#pragma unroll
      for (int f = 0; f < kRsFrames; ++f) {
        const float4 xv = *reinterpret_cast<const float4*>(rs_x + f * P.K4 + k);
        acc[f] = fmaf(w0, xv.x, acc[f]);
        acc[f] = fmaf(w1, xv.y, acc[f]);
        acc[f] = fmaf(w2, xv.z, acc[f]);
        acc[f] = fmaf(w3, xv.w, acc[f]);
      }
    }
#pragma unroll
    for (int f = 0; f < kRsFrames; ++f) {
      const long long o = (f0 + f) * P.nw + p;
      if (o < n_out) y[o] = acc[f];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The same resampler on the tensor cores: the GEMM [new x K] . [K x frames] as 3xTF32 (hi*hi + hi*lo + lo*hi with
// hi = tf32(v), lo = v - hi: ~21 mantissa bits per product, fp32 accumulation) with mma.sync.m16n8k8.  A CTA of 5 warps
// computes 32 frames of one utterance for all phases: the 32 windows sit in shared memory as fp32 (row stride = 4 mod 32
// floats, so the B-fragment loads -- lane (g, tig) reads window g, tap tig -- are conflict free), every warp takes two
// 16-phase M tiles at a time against the four 8-frame N tiles, loads its A fragments straight from the L2-resident
// fp32 table [new][K8] and splits both operands in registers.  Used when new % 16 == 0 (160 and 320 for the
// reference's two rates); the FP32 kernel above remains the general path.
// ------------------------------------------------------------------------------------------------
constexpr int kRmFrames = 32;
constexpr int kRmWarps = 5;

struct ResampleMmaParams {
  const float* in;
  const long long* in_off;
  const long long* in_len;
  float* out;
  const long long* out_off;
  const long long* out_len;
  const float* w;               // [nw][K8] row-major fp32, taps K..K8-1 zero
  int orig, nw, K, K8, KS, width;   // KS = shared-memory row stride (K8 + 4)
};

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  lo = __float_as_uint(v - __uint_as_float(hi));          // the MMA reads its top 19 bits
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kRmWarps * 32) resample_mma_kernel(const __grid_constant__ ResampleMmaParams P) {
  extern __shared__ __align__(16) float rm_x[];           // [kRmFrames][KS]
  const int b = blockIdx.y;
  const long long n_in = P.in_len[b], n_out = P.out_len[b];
  const long long f0 = static_cast<long long>(blockIdx.x) * kRmFrames;
  if (f0 * P.nw >= n_out) return;
  const float* x = P.in + P.in_off[b];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  for (int f = warp; f < kRmFrames; f += kRmWarps) {     // one window per warp and turn: coalesced, no divisions
    const long long j0 = (f0 + f) * P.orig - P.width;
    float* row = rm_x + f * P.KS;
    for (int k = lane; k < P.KS; k += 32) {
      const long long j = j0 + k;
      row[k] = (k < P.K && j >= 0 && j < n_in) ? x[j] : 0.f;
    }
  }
  __syncthreads();
  float* y = P.out + P.out_off[b];
  const int m_tiles = P.nw >> 4;
  for (int mt0 = warp * 2; mt0 < m_tiles; mt0 += kRmWarps * 2) {       // two M tiles (32 phases) per pass
    const int n_m = (mt0 + 1 < m_tiles) ? 2 : 1;
    float acc[2][kRmFrames / 8][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < kRmFrames / 8; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[m][n][e] = 0.f;
    // running pointers: rows g and g+8 of the two M tiles (a missing second tile aliases the first), window g of the
    // four N tiles
    const float* wa = P.w + static_cast<long long>(mt0 * 16 + g) * P.K8 + tig;
    const float* wb = wa + 8 * P.K8;
    const float* wc = wa + static_cast<long long>(n_m == 2 ? 16 : 0) * P.K8;
    const float* wd = wc + 8 * P.K8;
    const float* xr = rm_x + g * P.KS + tig;
    const int xs = 8 * P.KS;
#pragma unroll 1
    for (int k = 0; k < P.K8; k += 8, wa += 8, wb += 8, wc += 8, wd += 8, xr += 8) {
      uint32_t ah[2][4], al[2][4];
      split_tf32(__ldg(wa), ah[0][0], al[0][0]); split_tf32(__ldg(wb), ah[0][1], al[0][1]);
      split_tf32(__ldg(wa + 4), ah[0][2], al[0][2]); split_tf32(__ldg(wb + 4), ah[0][3], al[0][3]);
      split_tf32(__ldg(wc), ah[1][0], al[1][0]); split_tf32(__ldg(wd), ah[1][1], al[1][1]);
      split_tf32(__ldg(wc + 4), ah[1][2], al[1][2]); split_tf32(__ldg(wd + 4), ah[1][3], al[1][3]);
      uint32_t bh[kRmFrames / 8][2], bl[kRmFrames / 8][2];
#pragma unroll
      for (int n = 0; n < kRmFrames / 8; ++n) {
        split_tf32(xr[n * xs], bh[n][0], bl[n][0]);            // (tap k + tig, frame 8 n + g)
        split_tf32(xr[n * xs + 4], bh[n][1], bl[n][1]);
      }
      // three sweeps over the 8 accumulator tiles, small terms first: consecutive MMAs never touch the same tile
#pragma unroll
      for (int n = 0; n < kRmFrames / 8; ++n)
#pragma unroll
        for (int m = 0; m < 2; ++m) mma_tf32(acc[m][n], al[m], bh[n][0], bh[n][1]);
#pragma unroll
      for (int n = 0; n < kRmFrames / 8; ++n)
#pragma unroll
        for (int m = 0; m < 2; ++m) mma_tf32(acc[m][n], ah[m], bl[n][0], bl[n][1]);
#pragma unroll
      for (int n = 0; n < kRmFrames / 8; ++n)
#pragma unroll
        for (int m = 0; m < 2; ++m) mma_tf32(acc[m][n], ah[m], bh[n][0], bh[n][1]);
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      if (m >= n_m) continue;
#pragma unroll
      for (int n = 0; n < kRmFrames / 8; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int phase = (mt0 + m) * 16 + g + ((e & 2) ? 8 : 0);
          const long long frame = f0 + n * 8 + 2 * tig + (e & 1);
          const long long o = frame * P.nw + phase;
          if (o < n_out) y[o] = acc[m][n][e];
        }
    }
  }
}

}  // namespace lidfe
