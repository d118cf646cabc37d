// Micro-benchmarks that pin the per-SM ceilings the front-end kernel is designed against (B200, sm_100a):
// scalar vs packed (f32x2) FP32 add / fma issue rate, warp shuffle rate, shared-memory load rate.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <string>

#define ITER 2048
#define NACC 8

__global__ void k_fadd(float* out, float seed) {
  float a[NACC];
  for (int i = 0; i < NACC; ++i) a[i] = seed + i + threadIdx.x;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = a[i] + seed;
  }
  float s = 0; for (int i = 0; i < NACC; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, float seed) {
  float a[NACC];
  for (int i = 0; i < NACC; ++i) a[i] = seed + i + threadIdx.x;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = fmaf(a[i], seed, 0.5f);
  }
  float s = 0; for (int i = 0; i < NACC; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fadd2(float* out, float seed) {
  float2 a[NACC];
  const float2 sd = make_float2(seed, seed * 0.5f);
  for (int i = 0; i < NACC; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = __fadd2_rn(a[i], sd);
  }
  float s = 0; for (int i = 0; i < NACC; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float seed) {
  float2 a[NACC];
  const float2 sd = make_float2(seed, seed * 0.5f), c = make_float2(0.5f, 0.25f);
  for (int i = 0; i < NACC; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = __ffma2_rn(a[i], sd, c);
  }
  float s = 0; for (int i = 0; i < NACC; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// half FADD2, half scalar FADD interleaved: can they dual-issue / share the pipe?
__global__ void k_mix_fadd2_fadd(float* out, float seed) {
  float2 a[NACC / 2]; float b[NACC / 2];
  const float2 sd = make_float2(seed, seed * 0.5f);
  for (int i = 0; i < NACC / 2; ++i) { a[i] = make_float2(seed + i + threadIdx.x, seed - i); b[i] = seed * i; }
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC / 2; ++i) { a[i] = __fadd2_rn(a[i], sd); b[i] = b[i] + seed; }
  }
  float s = 0; for (int i = 0; i < NACC / 2; ++i) s += a[i].x + a[i].y + b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_shfl(float* out, float seed) {
  float a[NACC];
  for (int i = 0; i < NACC; ++i) a[i] = seed + i + threadIdx.x;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 3));
  }
  float s = 0; for (int i = 0; i < NACC; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int W>   // W = words per load (1, 2, 4)
__global__ void k_lds(float* out, float seed) {
  __shared__ __align__(16) float sm[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  float s = 0;
  int idx = (threadIdx.x * W) & 8191;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      const int j = (idx + i * 256 * W) & 8191;
      if (W == 1) s += sm[j];
      if (W == 2) { float2 v = *reinterpret_cast<float2*>(&sm[j]); s += v.x + v.y; }
      if (W == 4) { float4 v = *reinterpret_cast<float4*>(&sm[j]); s += v.x + v.y + v.z + v.w; }
    }
    idx = (idx + 32 * W) & 8191;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// shuffle + LDS + FADD mixed: do SHFL and LDS share the same data pipe?
__global__ void k_shfl_lds(float* out, float seed) {
  __shared__ float sm[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  float a[NACC / 2]; float s = 0;
  for (int i = 0; i < NACC / 2; ++i) a[i] = seed + i + threadIdx.x;
  int idx = threadIdx.x;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC / 2; ++i) {
      a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 3));
      s += sm[(idx + i * 256) & 8191];
    }
    idx = (idx + 32) & 8191;
  }
  for (int i = 0; i < NACC / 2; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// LDS.W where the two half-warps read the SAME addresses (table loads of the fbank kernel): is the duplicate free?
template <int W>
__global__ void k_lds_dup(float* out, float seed) {
  __shared__ __align__(16) float sm[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  float s = 0;
  int idx = ((threadIdx.x & 15) * W) & 8191;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      const int j = (idx + i * 256 * W) & 8191;
      if (W == 1) s += sm[j];
      if (W == 2) { float2 v = *reinterpret_cast<float2*>(&sm[j]); s += v.x + v.y; }
      if (W == 4) { float4 v = *reinterpret_cast<float4*>(&sm[j]); s += v.x + v.y + v.z + v.w; }
    }
    idx = (idx + 16 * W) & 8191;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int W>
__global__ void k_sts(float* out, float seed) {
  __shared__ __align__(16) float sm[8192];
  int idx = (threadIdx.x * W) & 8191;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      const int j = (idx + i * 256 * W) & 8191;
      if (W == 1) sm[j] = seed;
      if (W == 2) *reinterpret_cast<float2*>(&sm[j]) = make_float2(seed, seed);
      if (W == 4) *reinterpret_cast<float4*>(&sm[j]) = make_float4(seed, seed, seed, seed);
    }
    idx = (idx + 32 * W) & 8191;
    seed += 1.f;
  }
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
}
__global__ void k_dfma(float* out, float seed) {
  double a[NACC];
  const double sd = seed;
  for (int i = 0; i < NACC; ++i) a[i] = seed + i + threadIdx.x;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = fma(a[i], sd, 0.5);
  }
  double s = 0; for (int i = 0; i < NACC; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
__global__ void k_lg2(float* out, float seed) {
  float a[NACC];
  for (int i = 0; i < NACC; ++i) a[i] = seed + i + threadIdx.x;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  }
  float s = 0; for (int i = 0; i < NACC; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FFMA2 interleaved 1:1 with LDS.64: does the load issue in the second cycle of the packed op?
__global__ void k_ffma2_lds64(float* out, float seed) {
  __shared__ __align__(16) float sm[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  float2 a[NACC];
  const float2 sd = make_float2(seed, seed * 0.5f), c = make_float2(0.5f, 0.25f);
  for (int i = 0; i < NACC; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
  float s = 0;
  int idx = (threadIdx.x * 2) & 8191;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      a[i] = __ffma2_rn(a[i], sd, c);
      if ((i & 3) == 0) { float2 v = *reinterpret_cast<float2*>(&sm[(idx + i * 512) & 8191]); s += v.x; }
    }
    idx = (idx + 64) & 8191;
  }
  for (int i = 0; i < NACC; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FFMA2 + IADD3 interleaved 1:1 (integer / address work in the shadow of packed FP)
__global__ void k_ffma2_int(float* out, float seed) {
  float2 a[NACC]; int b[NACC];
  const float2 sd = make_float2(seed, seed * 0.5f), c = make_float2(0.5f, 0.25f);
  for (int i = 0; i < NACC; ++i) { a[i] = make_float2(seed + i + threadIdx.x, seed - i); b[i] = threadIdx.x + i; }
  const int k = (int)seed + 3;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) { a[i] = __ffma2_rn(a[i], sd, c); b[i] = (b[i] ^ k) + it; }
  }
  float s = 0; for (int i = 0; i < NACC; ++i) s += a[i].x + a[i].y + b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char* name, F launch, double ops_per_thread_iter, int threads, int blocks, int sms, double clk_ghz) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(e0); for (int r = 0; r < 5; ++r) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double total = ops_per_thread_iter * ITER * (double)threads * blocks;
  const double per_sm_clk = total / (ms * 1e-3) / sms / (clk_ghz * 1e9);
  printf("%-18s %8.3f ms  %8.1f thread-ops/clk/SM (at %.3f GHz nominal)  = %6.2f warp-instr/clk/SM\n", name, ms, per_sm_clk, clk_ghz,
         per_sm_clk / 32.0);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount; const double clk = p.clockRate * 1e-6;
  printf("%s SMs=%d clockRate=%.3f GHz\n", p.name, sms, clk);
  const int threads = 256, blocks = sms * 8;
  float* out; cudaMalloc(&out, sizeof(float) * threads * blocks);
  run("FADD", [&] { k_fadd<<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("FFMA", [&] { k_ffma<<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("FADD2 (instr)", [&] { k_fadd2<<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("FFMA2 (instr)", [&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("FADD2+FADD (instr)", [&] { k_mix_fadd2_fadd<<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("SHFL", [&] { k_shfl<<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("LDS.32", [&] { k_lds<1><<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("LDS.64", [&] { k_lds<2><<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("LDS.128", [&] { k_lds<4><<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("SHFL+LDS.32 (instr)", [&] { k_shfl_lds<<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("LDS.32 dup-halves", [&] { k_lds_dup<1><<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("LDS.64 dup-halves", [&] { k_lds_dup<2><<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("LDS.128 dup-halves", [&] { k_lds_dup<4><<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("STS.32", [&] { k_sts<1><<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("STS.64", [&] { k_sts<2><<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("STS.128", [&] { k_sts<4><<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("DFMA", [&] { k_dfma<<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("MUFU.LG2", [&] { k_lg2<<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("FFMA2 (+LDS.64 1:4)", [&] { k_ffma2_lds64<<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  run("FFMA2 (+2 int ops)", [&] { k_ffma2_int<<<blocks, threads>>>(out, 1.0001f); }, NACC, threads, blocks, sms, clk);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
