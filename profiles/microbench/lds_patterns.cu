// lds_patterns.cu — how many shared-memory wavefronts (SM cycles of the 128 B/clk crossbar) one
// warp-wide LDS.64 / LDS.128 / STS.64 costs for the lane->address patterns the fused EKF kernels use.
// Four warps (one per SM sub-partition) issue independent loads back to back; the steady-state
// SM cycles per warp instruction is the wavefront count of that pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_patterns lds_patterns.cu && ./lds_patterns
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void lds128(unsigned long long& a, unsigned long long& b, unsigned addr) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ void lds64(unsigned long long& a, unsigned addr) {
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(a) : "r"(addr) : "memory");
}

// pattern -> byte offset of this lane's element (unit = the access width)
__device__ int pattern_index(int p, int lane) {
  switch (p) {
    case 0: return 0;                    // all lanes one address
    case 1: return lane >> 4;            // 2 distinct, adjacent
    case 2: return lane >> 3;            // 4 distinct (by quarter warp), adjacent
    case 3: return lane & 3;             // 4 distinct (interleaved), adjacent
    case 4: return lane & 7;             // 8 distinct (interleaved), adjacent
    case 5: return lane >> 2;            // 8 distinct (groups of 4 lanes), adjacent
    case 6: return lane & 15;            // 16 distinct (interleaved)
    case 7: return lane >> 1;            // 16 distinct (pairs)
    case 8: return lane;                 // 32 distinct, consecutive
    case 9: return (lane & 7) * 13;      // 8 distinct, scattered (odd stride)
    case 10: return (lane >> 3) * 9 + (lane & 7) * 0 + 100;   // 4 distinct, scattered
    case 11: return (lane & 7) + 8 * 3 * (lane >> 3);         // 4 groups of 8 consecutive, groups 24 apart
    case 12: return (lane & 3) * 5 + (lane >> 2) * 0;         // 4 distinct stride 5
    case 13: return (lane & 7) * 16;     // 8 distinct, same bank (worst case)
    default: return lane;
  }
}

template <int WIDTH>   // 8 or 16 bytes
__global__ void probe(long long* cyc, double* sink, int pattern, int iters) {
  extern __shared__ __align__(16) double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 0.0;   // zeros: loaded values feed the next address
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + pattern_index(pattern, lane) * WIDTH;
  // eight independent load chains per lane; every loaded value (zero at run time, unknown at
  // compile time) is added to the chain's next address, so no load can leave the loop
  unsigned addr[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) addr[k] = base + k * 4096;
  unsigned long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      unsigned long long a, b = 0;
      if (WIDTH == 16) lds128(a, b, addr[k]);
      else lds64(a, addr[k]);
      addr[k] += (unsigned)a;
      acc[k] ^= b;
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  sink[threadIdx.x] = (double)(acc[0] ^ acc[1] ^ acc[2] ^ acc[3] ^ acc[4] ^ acc[5] ^ acc[6] ^ acc[7]);
}

int main() {
  long long* cyc;
  double* sink;
  cudaMalloc(&cyc, 64);
  cudaMalloc(&sink, 1024 * 8);
  cudaFuncSetAttribute(probe<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(probe<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const char* names[] = {"1 address", "2 distinct (halves)", "4 distinct (quarters)", "4 distinct (lane&3)",
                         "8 distinct (lane&7)", "8 distinct (lane>>2)", "16 distinct (lane&15)", "16 distinct (lane>>1)",
                         "32 consecutive", "8 distinct stride 13", "4 distinct stride 9", "4x8 consecutive, 24 apart",
                         "4 distinct stride 5", "8 distinct same bank"};
  const int iters = 2048;
  for (int width : {8, 16})
    for (int threads : {32, 128, 256}) {
      for (int p = 0; p < 14; ++p) {
        for (int rep = 0; rep < 2; ++rep) {
          if (width == 8) probe<8><<<1, threads, 65536>>>(cyc, sink, p, iters);
          else probe<16><<<1, threads, 65536>>>(cyc, sink, p, iters);
          cudaDeviceSynchronize();
        }
        long long h = 0;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double per = (double)h / (iters * 8.0 * (threads / 32));
        printf("LDS.%-3d warps %d  %-28s %.2f SM-cycles per warp instruction\n", width * 8, threads / 32, names[p], per);
      }
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
