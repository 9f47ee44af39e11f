#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, int iters, double b, double c) {
  double a = out[threadIdx.x];
  long long t0, t1;
  // dependent DFMA chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { a = __dadd_rn(a, c); a = __dadd_rn(a, c); a = __dadd_rn(a, c); a = __dadd_rn(a, c); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { a = __dmul_rn(a, b); a = __dmul_rn(a, b); a = __dmul_rn(a, b); a = __dmul_rn(a, b); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { a = sqrt(a + 2.0); a = sqrt(a + 2.0); a = sqrt(a + 2.0); a = sqrt(a + 2.0); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { a = 1.0 / (a + 2.0); a = 1.0 / (a + 2.0); a = 1.0 / (a + 2.0); a = 1.0 / (a + 2.0); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { double s, co; sincos(a, &s, &co); a = s + co; sincos(a, &s, &co); a = s + co; sincos(a, &s, &co); a = s + co; sincos(a, &s, &co); a = s + co; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // 4 independent chains
  double a1 = a + 1, a2 = a + 2, a3 = a + 3;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { a = fma(a, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
  out[threadIdx.x] = a + a1 + a2 + a3;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 1024 * 8);
  const char* names[] = {"dfma dep", "dadd dep", "dmul dep", "sqrt(+add) dep", "rcp(+add) dep", "sincos(+add) dep", "dfma 4 chains"};
  for (int threads : {32, 128, 256}) {
    lat<<<1, threads>>>(out, cyc, 1024, 1.0000001, 1e-9); cudaDeviceSynchronize();
    lat<<<1, threads>>>(out, cyc, 1024, 1.0000001, 1e-9); cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
    for (int k = 0; k < 7; ++k) printf("threads %3d %-18s %.1f cycles/op\n", threads, names[k], h[k] / 4096.0);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
