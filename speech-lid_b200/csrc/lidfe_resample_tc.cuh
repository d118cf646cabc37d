// lidfe_resample_tc.cuh -- the polyphase resampler GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), row f4.
//
// out[frame * nw + p] = sum_k W[p][k] * xpad[frame * orig - width + k]        (ref: lid/ConformerLangModel.py:131-178 ->
// ta: functional/functional.py _apply_sinc_resample_kernel: a strided conv1d = a GEMM whose left operand is an im2col
// view of the waveform).  One CTA computes 128 consecutive frames of one utterance for N phases:
//   D[128 frames x N phases] (fp32, in TMEM) = A[128 x K] . B[N x K]^T,  both operands K-major in shared memory,
// as 3 x TF32: every fp32 value is split once into hi = tf32(v) and lo = v - hi, and D += A_lo.B_hi + A_hi.B_lo + A_hi.B_hi
// (the product of the two low parts is below fp32 resolution) -- the same decomposition, to the same 1e-5 parity bar,
// as resample_mma_kernel, but with tcgen05.mma issued by one thread, accumulators in tensor memory, and the operands
// staged through a 3-stage mbarrier pipeline:
//   * B (the FIR bank) is constant: the host lays every (phase tile, 32-tap block) out as the exact shared-memory image
//     -- rows of 128 bytes, 16-byte chunks XOR-swizzled by (row % 8), hi image then lo image -- so one 1-D TMA bulk copy
//     (cp.async.bulk, complete_tx on the stage's mbarrier) brings a block in;
//   * A cannot come by TMA: its rows are windows of the waveform `orig` samples apart (441 * 4 bytes is no multiple of 16),
//     so four builder warps gather them -- one coalesced 128-byte load per row and block, split, two conflict-free
//     STS.32 into the same swizzled layout -- and publish the stage with fence.proxy.async + mbarrier arrive;
//   * one lane of the MMA warp waits for both, issues 4 k-steps x 3 tcgen05.mma (M 128, N, K 8) per block and hands the
//     stage back with tcgen05.commit; a last commit tells the builder warps, which double as the epilogue, that the
//     accumulators are complete: tcgen05.ld (32 lanes x 32 columns per warp and turn) -> registers -> global.
// SASS of this kernel: UTCHMMA / UTCBAR / LDTM / UBLKCP (profiles/r2_resample_tc_sass.txt).
#pragma once
#include "lidfe_kernels.cuh"

namespace lidfe {

constexpr int kTcM = 128;            // frames per CTA = rows of A = TMEM lanes
constexpr int kTcKB = 32;            // taps per pipeline stage = one 128-byte swizzle row
constexpr int kTcThreads = 256;      // warp 0: B loader, warp 1: MMA issuer + TMEM owner, warps 4-7: A builders + epilogue
constexpr int kTcABytes = kTcM * kTcKB * 4;          // 16 KB per (hi | lo) image of A

struct ResampleTcParams {
  const float* in;
  const long long* in_off;
  const long long* in_len;
  float* out;
  const long long* out_off;
  const long long* out_len;
  const unsigned char* wimg;         // [n_tiles][KB][2][N * 128 bytes] shared-memory images of the FIR bank (hi, lo)
  int orig, nw, K, KB, width;        // KB = ceil(K / 32) tap blocks
  int N, stages, tmem_cols;          // phases per CTA (multiple of 32, <= 256), pipeline depth, allocated TMEM columns
};

__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity) {
  // bounded: a descriptor mistake must end in a trap the host sees, not in a hung GPU
  for (uint32_t spins = 0;; ++spins) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
    if (spins > (1u << 24)) __trap();
  }
}

// shared-memory matrix descriptor, K-major, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);      // start address
  d |= static_cast<uint64_t>(1u) << 16;                        // leading byte offset (unused with swizzle): 1
  d |= static_cast<uint64_t>(1024u >> 4) << 32;                // stride byte offset: next 8-row group
  d |= static_cast<uint64_t>(1u) << 46;                        // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2u) << 61;                        // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kTcThreads, 1) resample_tc_kernel(const __grid_constant__ ResampleTcParams P) {
  extern __shared__ unsigned char tc_smem_raw[];
  const int b = blockIdx.y, nt = blockIdx.z;
  const long long n_in = P.in_len[b], n_out = P.out_len[b];
  const long long f0 = static_cast<long long>(blockIdx.x) * kTcM;
  if (f0 * P.nw >= n_out) return;                                  // (the whole CTA: nothing has been set up yet)

  // 1024-byte aligned stage buffers (the swizzle works on absolute shared-memory address bits)
  unsigned char* const base = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  const int b_bytes = P.N * 128;                                   // one (hi | lo) image of B
  const int stage_bytes = 2 * kTcABytes + 2 * b_bytes;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(base + P.stages * stage_bytes);
  uint64_t* const full_a = bars;                                   // [stages] 128 builder threads
  uint64_t* const full_b = bars + P.stages;                        // [stages] TMA bytes
  uint64_t* const empty = bars + 2 * P.stages;                     // [stages] tcgen05.commit
  uint64_t* const acc_full = bars + 3 * P.stages;                  // accumulators complete
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * P.stages + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&full_a[s], kTcM);
      mbar_init(&full_b[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                                                 // TMEM: allocated and later freed by this warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(static_cast<uint32_t>(P.tmem_cols)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ---- B loader: one bulk copy per tap block (hi and lo images are adjacent in the host-built table) -------------
    if (lane == 0) {
      const unsigned char* src = P.wimg + static_cast<long long>(nt) * P.KB * 2 * b_bytes;
      for (int kb = 0; kb < P.KB; ++kb) {
        const int s = kb % P.stages, u = kb / P.stages;
        if (u > 0) tc_wait(&empty[s], static_cast<uint32_t>((u - 1) & 1));
        mbar_expect_tx(&full_b[s], static_cast<uint32_t>(2 * b_bytes));
        tma_bulk_g2s_plain(base + s * stage_bytes + 2 * kTcABytes, src + static_cast<long long>(kb) * 2 * b_bytes,
                           static_cast<uint32_t>(2 * b_bytes), &full_b[s]);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: one lane ------------------------------------------------------------------------------------
    if (lane == 0) {
      // instruction descriptor: D fp32, A and B tf32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(P.N >> 3) << 17) |
                             (static_cast<uint32_t>(kTcM >> 4) << 24);
      for (int kb = 0; kb < P.KB; ++kb) {
        const int s = kb % P.stages, u = kb / P.stages;
        tc_wait(&full_a[s], static_cast<uint32_t>(u & 1));
        tc_wait(&full_b[s], static_cast<uint32_t>(u & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(base + s * stage_bytes), a_lo = a_hi + kTcABytes;
        const uint32_t b_hi = a_hi + 2 * kTcABytes, b_lo = b_hi + b_bytes;
#pragma unroll
        for (int j = 0; j < kTcKB / 8; ++j) {                      // K = 8 tf32 = 32 bytes per instruction
          const uint32_t off = 32u * j;
          tc_mma_tf32(tmem_d, tc_smem_desc(a_lo + off), tc_smem_desc(b_hi + off), idesc, (kb | j) ? 1u : 0u);
          tc_mma_tf32(tmem_d, tc_smem_desc(a_hi + off), tc_smem_desc(b_lo + off), idesc, 1u);
          tc_mma_tf32(tmem_d, tc_smem_desc(a_hi + off), tc_smem_desc(b_hi + off), idesc, 1u);
        }
        tc_commit(&empty[s]);                                      // the stage is free once these MMAs have read it
      }
      tc_commit(acc_full);
    }
  } else if (warp >= 4) {
    // ---- A builders: warp q gathers rows 32 q .. 32 q + 31 of every tap block -------------------------------------
    const int q = warp - 4;
    const float* x = P.in + P.in_off[b];
    for (int kb = 0; kb < P.KB; ++kb) {
      const int s = kb % P.stages, u = kb / P.stages;
      const int k = kb * kTcKB + lane;                             // this lane's tap
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {                               // 32 independent coalesced loads in flight
        const long long j = (f0 + 32 * q + i) * P.orig - P.width + k;
        v[i] = (k < P.K && j >= 0 && j < n_in) ? __ldg(x + j) : 0.f;
      }
      if (u > 0) tc_wait(&empty[s], static_cast<uint32_t>((u - 1) & 1));
      unsigned char* const a_hi = base + s * stage_bytes;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int row = 32 * q + i;
        uint32_t hi, lo;
        split_tf32(v[i], hi, lo);
        const int off = row * 128 + ((((lane >> 2) ^ (row & 7)) << 4) | ((lane & 3) << 2));
        *reinterpret_cast<uint32_t*>(a_hi + off) = hi;
        *reinterpret_cast<uint32_t*>(a_hi + kTcABytes + off) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
      mbar_arrive(&full_a[s]);
    }
    // ---- epilogue: TMEM lane = frame, column = phase ------------------------------------------------------------------
    tc_wait(acc_full, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* const y = P.out + P.out_off[b];
    const long long o_row = (f0 + 32 * q + lane) * P.nw + static_cast<long long>(nt) * P.N;
    for (int c = 0; c < P.N; c += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem_d + (static_cast<uint32_t>(32 * q) << 16) + static_cast<uint32_t>(c);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const long long o = o_row + c;
      if (o + 32 <= n_out && ((reinterpret_cast<uintptr_t>(y + o) & 15) == 0)) {
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          *reinterpret_cast<float4*>(y + o + e) = make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]),
                                                              __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (o + e < n_out) y[o + e] = __uint_as_float(r[e]);
      }
    }
  }

  // ---- teardown: every tcgen05 operation of this CTA is complete before the columns go back ------------------------
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(static_cast<uint32_t>(P.tmem_cols)) : "memory");
  }
}

}  // namespace lidfe
