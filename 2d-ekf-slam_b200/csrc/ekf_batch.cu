// ekf_batch.cu — regime A: batched small maps, one CTA per filter (sm_100a).
//
// ekf_batch_run_kernel: the fused multi-step path. A CTA pulls its filter's covariance slab from
// HBM into shared memory with one bulk async copy (cp.async.bulk, mbarrier completion), runs all
// T iterations of the slam.cpp:130-182 loop with P resident on chip (propagate -> optional
// compass -> one update per measurement; step records streamed in with cp.async one step
// ahead), and pushes the slab back with one bulk store. Filters are independent: no
// inter-CTA communication, so a batch shards across GPUs by filter range.
//
// ekf_percall_kernel: one reference call (doPropagation / doUpdate / doUpdateCompass) for every
// filter of the batch, covariance addressed in HBM.
#include <cstdio>
#include "ekf_cta.cuh"
#include "ekf_internal.h"

namespace {

constexpr int kThreads = 256;

// ---- PTX helpers: mbarrier + bulk async copy (TMA engine, non-tensor form) ------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Shared-memory carve-up (all offsets multiples of 16 bytes).
struct SmemLayout {
  size_t P, x, W, rec, scratch, bar, total;
};
__host__ __device__ inline SmemLayout smem_layout(int cap_n, int ld, int L) {
  SmemLayout s;
  size_t o = 0;
  s.P = o;       o += (size_t)cap_n * ld * sizeof(double);
  s.x = o;       o += (size_t)((cap_n + 1) & ~1) * sizeof(double);
  s.W = o;       o += (size_t)cap_n * sizeof(double2);
  s.rec = o;     o += (size_t)2 * ((L + 1) & ~1) * sizeof(double);
  s.scratch = o; o += (sizeof(CtaScratch) + 15) & ~(size_t)15;
  s.bar = o;     o += 16;
  s.total = o;
  return s;
}

struct RunArgs {
  EkfState st;
  EkfRunIO io;
  EkfConst k;
};

__global__ void __launch_bounds__(kThreads, 2) ekf_batch_run_kernel(const RunArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int cap_n = a.st.cap_n, ld = a.st.ld, L = a.io.L, T = a.io.T, M = a.io.M;
  const SmemLayout lay = smem_layout(cap_n, ld, L);
  double* Ps = reinterpret_cast<double*>(smem + lay.P);
  double* xs = reinterpret_cast<double*>(smem + lay.x);
  double2* Ws = reinterpret_cast<double2*>(smem + lay.W);
  double* recbuf = reinterpret_cast<double*>(smem + lay.rec);
  CtaScratch* sc = reinterpret_cast<CtaScratch*>(smem + lay.scratch);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + lay.bar);
  const int Lp = (L + 1) & ~1;
  const int tid = threadIdx.x;
  const uint32_t slab_bytes = (uint32_t)(a.st.slab * sizeof(double));

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  uint32_t phase = 0;

  for (int f = blockIdx.x; f < a.st.F; f += gridDim.x) {
    double* gP = a.st.P + (size_t)f * a.st.slab;
    double* gx = a.st.x + (size_t)f * a.st.xs;
    const double* grec = a.io.records + (size_t)f * T * L;
    if (tid == 0) {
      bulk_wait_read0();   // the previous filter's store has finished reading Ps
      mbar_expect_tx(bar, slab_bytes);
      bulk_g2s(Ps, gP, slab_bytes, bar);
    }
    for (int i = tid; i < cap_n; i += kThreads) xs[i] = gx[i];
    for (int i = tid; i < L; i += kThreads) cp_async8(recbuf + i, grec + i);
    int n_lm = a.st.nlm[f];
    int dropped = 0;
    cp_async_wait_all();
    mbar_wait(bar, phase);
    phase ^= 1;
    __syncthreads();

    for (int t = 0; t < T; ++t) {
      const double* cur = recbuf + (size_t)(t & 1) * Lp;
      if (t + 1 < T) {
        double* nxt = recbuf + (size_t)((t + 1) & 1) * Lp;
        const double* g = grec + (size_t)(t + 1) * L;
        for (int i = tid; i < L; i += kThreads) cp_async8(nxt + i, g + i);
      }
      // slam.cpp:136
      cta_propagate(Ps, ld, xs, 3 + 2 * n_lm, cur[0], cur[1], cur[2], sc, a.k);
      // slam.cpp:144-147
      if (cur[6] != 0.0) cta_update_compass(Ps, ld, xs, 3 + 2 * n_lm, cur[3], cur[4], Ws, sc, a.k);
      // slam.cpp:150-171
      const int nz = min((int)cur[5], (L - 8) / 6);   // never read past the record's measurement slots
      for (int m = 0; m < M; ++m) {
        UpdateOut o;
        if (m < nz) {
          const double* zr = cur + 8 + 6 * m;
          o = cta_update(Ps, ld, xs, n_lm, n_lm, a.st.cap_lm, zr[0], zr[1], zr + 2, Ws, sc, a.k);
          dropped |= (o.decision == EKF_DEC_DROPPED);
        } else {
          o.decision = EKF_DEC_NONE; o.index = -1; o.mahal = 0.0;
        }
        if (tid == 0) {
          const size_t oi = ((size_t)f * T + t) * M + m;
          if (a.io.decision) a.io.decision[oi] = o.decision;
          if (a.io.index) a.io.index[oi] = o.index;
          if (a.io.mahal) a.io.mahal[oi] = o.mahal;
        }
      }
      if (a.io.pose_trace && tid < 3) a.io.pose_trace[((size_t)f * T + t) * 3 + tid] = xs[tid];   // slam.cpp:181
      cp_async_wait_all();
      __syncthreads();
    }

    // write back
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(gP, Ps, slab_bytes);
      bulk_commit();
      a.st.nlm[f] = n_lm;
      if (dropped) a.st.status[f] |= 1;
    }
    for (int i = tid; i < cap_n; i += kThreads) gx[i] = xs[i];
    __syncthreads();   // xs is reloaded by the next iteration
  }
  if (tid == 0) bulk_wait0();
}

struct PercallArgs {
  EkfState st;
  EkfPercallIO io;
  EkfConst k;
};

template <int OP>
__global__ void __launch_bounds__(kThreads) ekf_percall_kernel(const PercallArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int cap_n = a.st.cap_n, ld = a.st.ld;
  double* xs = reinterpret_cast<double*>(smem);
  double2* Ws = reinterpret_cast<double2*>(smem + (size_t)((cap_n + 1) & ~1) * sizeof(double));
  CtaScratch* sc = reinterpret_cast<CtaScratch*>(reinterpret_cast<unsigned char*>(Ws) + (size_t)cap_n * sizeof(double2));
  const int tid = threadIdx.x;
  const int f = blockIdx.x;
  double* P = a.st.P + (size_t)f * a.st.slab;
  double* gx = a.st.x + (size_t)f * a.st.xs;
  int n_lm = a.st.nlm[f];
  for (int i = tid; i < cap_n; i += kThreads) xs[i] = gx[i];
  __syncthreads();
  if (OP == EKF_OP_PROPAGATE) {
    cta_propagate(P, ld, xs, 3 + 2 * n_lm, a.io.vel[f], a.io.rotvel[f], a.io.dt[(size_t)f * a.io.dt_stride], sc, a.k);
  } else if (OP == EKF_OP_COMPASS) {
    if (!a.io.cvalid || a.io.cvalid[f]) cta_update_compass(P, ld, xs, 3 + 2 * n_lm, a.io.cz[f], a.io.cR[f], Ws, sc, a.k);
  } else {
    int dropped = 0;
    const int n_gate = n_lm;   // Update.cpp:26: the gating bound is read once per doUpdate call
    for (int m = 0; m < a.io.n_z; ++m) {
      const double* zr = a.io.zr + ((size_t)f * a.io.n_z + m) * 6;
      const UpdateOut o = cta_update(P, ld, xs, n_lm, n_gate, a.st.cap_lm, zr[0], zr[1], zr + 2, Ws, sc, a.k);
      dropped |= (o.decision == EKF_DEC_DROPPED);
      if (tid == 0) {
        const size_t oi = (size_t)f * a.io.n_z + m;
        if (a.io.decision) a.io.decision[oi] = o.decision;
        if (a.io.index) a.io.index[oi] = o.index;
        if (a.io.mahal) a.io.mahal[oi] = o.mahal;
      }
    }
    if (tid == 0) {
      a.st.nlm[f] = n_lm;
      if (dropped) a.st.status[f] |= 1;
    }
  }
  __syncthreads();
  for (int i = tid; i < cap_n; i += kThreads) gx[i] = xs[i];
}

// DFMA-chain microbenchmark: 8 independent chains per thread, 2 flop per DFMA.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
  double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
      a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace

size_t ekf_batch_smem_bytes(int cap_n, int ld, int L) { return smem_layout(cap_n, ld, L).total; }

int ekf_batch_max_landmarks(size_t smem_optin) {
  int best = 0;
  for (int N = 1; N < 4096; ++N) {
    const int cap_n = 3 + 2 * N, ld = (cap_n + 1) & ~1;
    if (ekf_batch_smem_bytes(cap_n, ld, EKF_RECORD_LEN_MAX) <= smem_optin) best = N;
    else break;
  }
  return best;
}

cudaError_t ekf_batch_prepare(int cap_n, int ld, int max_L, int sm_count, int* grid_cap) {
  const size_t bytes = ekf_batch_smem_bytes(cap_n, ld, max_L);
  // The attribute is per function AND per device, and several handles of different capacity may be
  // alive on one GPU: keep a high-water mark per device and only ever raise it (a later, smaller
  // handle must not lower the limit under an earlier, larger one).
  static size_t configured_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  size_t& configured = configured_dev[dev & 63];
  if (bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(ekf_batch_run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    configured = bytes;
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ekf_batch_run_kernel, kThreads, bytes);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidConfiguration;
  *grid_cap = per_sm * sm_count;
  return cudaSuccess;
}

cudaError_t ekf_batch_run(const EkfState& st, const EkfRunIO& io, const EkfConst& k, int grid_cap,
                          cudaStream_t stream) {
  RunArgs a{st, io, k};
  const size_t bytes = ekf_batch_smem_bytes(st.cap_n, st.ld, io.L);
  const int grid = st.F < grid_cap ? st.F : grid_cap;
  ekf_batch_run_kernel<<<grid, kThreads, bytes, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t ekf_batch_percall(const EkfState& st, const EkfPercallIO& io, EkfOp op, const EkfConst& k,
                              cudaStream_t stream) {
  PercallArgs a{st, io, k};
  const size_t bytes = (size_t)((st.cap_n + 1) & ~1) * sizeof(double) + (size_t)st.cap_n * sizeof(double2) +
                       sizeof(CtaScratch) + 16;
  switch (op) {
    case EKF_OP_PROPAGATE: ekf_percall_kernel<EKF_OP_PROPAGATE><<<st.F, kThreads, bytes, stream>>>(a); break;
    case EKF_OP_UPDATE: ekf_percall_kernel<EKF_OP_UPDATE><<<st.F, kThreads, bytes, stream>>>(a); break;
    case EKF_OP_COMPASS: ekf_percall_kernel<EKF_OP_COMPASS><<<st.F, kThreads, bytes, stream>>>(a); break;
  }
  return cudaGetLastError();
}

cudaError_t ekf_fp64_peak(double* flops_per_s, cudaStream_t stream) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  cudaError_t e = cudaMalloc(&out, sizeof(double) * blocks * threads);
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  fp64_peak_kernel<<<blocks, threads, 0, stream>>>(out, 64);   // warm-up
  float best_ms = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0, stream);
    fp64_peak_kernel<<<blocks, threads, 0, stream>>>(out, iters);
    cudaEventRecord(e1, stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best_ms) best_ms = ms;
  }
  e = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (e != cudaSuccess) return e;
  const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;   // 8 chains x 8 unroll = 64 DFMA / iter
  *flops_per_s = flops / (best_ms * 1e-3);
  return cudaSuccess;
}
