#include <cuda_runtime.h>
#include <cstdio>
// one warp streams `rows` rows of 80 floats: 10 float4 loads in flight, (x-m)*s, store.  variants by template.
template <int MODE>
__global__ void k(float* base, int rows, long long* cyc) {
  const int lane = threadIdx.x & 31;
  int roff[5], col[5];
  for (int m = 0; m < 5; ++m) { int i = lane + 32 * m; roff[m] = i / 20; col[m] = i - roff[m] * 20; }
  float4* g = reinterpret_cast<float4*>(base);
  long long t0 = clock64();
  for (int rb = 0; rb < rows; rb += 16) {
    float4 x[10];
#pragma unroll
    for (int u = 0; u < 10; ++u) {
      const int row = rb + (u / 5) * 8 + roff[u % 5];
      if (row < rows) x[u] = g[row * 20 + col[u % 5]];
    }
#pragma unroll
    for (int u = 0; u < 10; ++u) {
      const int row = rb + (u / 5) * 8 + roff[u % 5];
      if (row >= rows) continue;
      float4 v = x[u];
      v.x = (v.x - 1.f) * 0.5f; v.y = (v.y - 1.f) * 0.5f; v.z = (v.z - 1.f) * 0.5f; v.w = (v.w - 1.f) * 0.5f;
      if (MODE == 0) g[row * 20 + col[u % 5]] = v;
      if (MODE == 1) { if (v.x == 123.456f) g[row * 20 + col[u % 5]] = v; }
    }
  }
  long long t1 = clock64();
  if (lane == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  const int rows = 798;
  float* d; long long* c; cudaMalloc(&d, 64 << 20); cudaMalloc(&c, 8 * 1024);
  cudaMemset(d, 0, 64 << 20);
  long long h[1024];
  for (int mode = 0; mode < 2; ++mode)
    for (int nb : {1, 148, 592, 2368}) {
      for (int r = 0; r < 2; ++r) {
        if (mode == 0) k<0><<<nb, 32>>>(d, rows, c); else k<1><<<nb, 32>>>(d, rows, c);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, c, 8 * (nb < 1024 ? nb : 1024), cudaMemcpyDeviceToHost);
      printf("mode %d (0=load+store,1=load only) blocks %4d (same rows): cycles/round %.0f  (%.2f us per 798 rows)\n", mode, nb, h[0] / 50.0, h[0] / 1.965e3);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
