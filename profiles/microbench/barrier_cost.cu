// Cost of CTA-wide phase boundaries on B200 (192 threads = 6 warps, one CTA on one SM):
//   a) bare __syncthreads()              b) one global store by thread 0 before the barrier
//   c) cp.async.wait_all + barrier       d) barrier + smem broadcast + 8 dependent DFMA per thread
//   e) one warp works 400 cycles, the others wait at the barrier (skew)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, int* g, int iters) {
  __shared__ double sh[64];
  const int tid = threadIdx.x;
  if (tid < 64) sh[tid] = tid * 1e-3;
  __syncthreads();
  long long t0, t1;
  double acc = tid;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) __syncthreads();
  t1 = clock64(); if (tid == 0) out[0] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { if (tid == 0) g[i & 1023] = i; __syncthreads(); }
  t1 = clock64(); if (tid == 0) out[1] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { asm volatile("cp.async.wait_all;" ::: "memory"); __syncthreads(); }
  t1 = clock64(); if (tid == 0) out[2] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    __syncthreads();
    double b = sh[i & 63];
    for (int q = 0; q < 8; ++q) acc = fma(acc, b, 1e-9);
    if (tid == (i & 63)) sh[i & 63] = acc * 1e-30 + 1e-3;
  }
  t1 = clock64(); if (tid == 0) out[3] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (tid < 32) { for (int q = 0; q < 50; ++q) acc = fma(acc, 1.0000001, 1e-9); }
    __syncthreads();
  }
  t1 = clock64(); if (tid == 0) out[4] = t1 - t0;
  // store to global each iteration by thread 0 WITHOUT barrier dependence: 3 stores like the trace
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { if (tid == 0) { g[i & 1023] = i; g[1024 + (i & 1023)] = i; g[2048 + (i & 1023)] = i; } __syncthreads(); }
  t1 = clock64(); if (tid == 0) out[5] = t1 - t0;
  if (acc == 12345.678) out[7] = 1;
}
int main() {
  long long* out; int* g; cudaMalloc(&out, 64); cudaMalloc(&g, 4096 * 4);
  const int iters = 2000;
  k<<<1, 192>>>(out, g, iters); cudaDeviceSynchronize();
  k<<<1, 192>>>(out, g, iters); cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
  const char* n[] = {"bare barrier", "1 STG by tid0 + barrier", "cp.async.wait_all + barrier", "barrier + LDS + 8 dep DFMA", "one warp 50 dep DFMA + barrier", "3 STG by tid0 + barrier"};
  for (int i = 0; i < 6; ++i) printf("%-32s %.1f cycles/iter\n", n[i], (double)h[i] / iters);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
