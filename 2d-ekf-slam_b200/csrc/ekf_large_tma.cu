// ekf_large_tma.cu — regime B covariance downdate, TMA-staged variant (sm_100a).
//
// Same arithmetic as large_downdate in ekf_large.cu (P_ij += u_i . W_j, one read + one write of the
// dense n x n covariance), but the covariance is tile-streamed through shared memory by the TMA
// engine: persistent CTAs, one producer thread issuing cp.async.bulk.tensor.2d loads (SASS
// UTMALDG) into a 4-stage ring of 128x16 FP64 tiles with mbarrier completion, four consumer warps
// updating the tile in place in shared memory (one row per thread: conflict-free), and
// cp.async.bulk.tensor.2d stores (UTMASTG) back to the same coordinates. The tensor map covers the
// filter's whole capacity slab; tiles beyond the live dimension are not visited, rows/columns
// inside a boundary tile but beyond n see W = 0 (unchanged), out-of-bounds parts are zero-filled
// on load and clipped on store.
#include <cuda.h>
#include <cstdlib>
#include <cuda_runtime.h>

#include "ekf_cta.cuh"
#include "ekf_internal.h"
#include "ekf_pdl.cuh"

namespace {

constexpr int TR = 128, TC = 16;
// Ring depth. Two CTAs per SM in both uses. One GPU: six stages (192 KB of tiles in flight per SM - measured
// best from N = 1,000 to 10,000). A shard of a look-ahead run: four stages (128 KB per SM) - with more than
// that in shared memory the side stream's small kernels are not scheduled beside the sweep until most of it
// is over (profiles/r02_shard_timeline.txt).
constexpr int kStagesWide = 6, kStagesSlim = 4;
constexpr int kConsumers = 128, kThreadsTma = kConsumers + 32;
constexpr uint32_t kTileBytes = TR * TC * sizeof(double);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0),
               "r"(c1), "r"(smem_u32(src))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void named_barrier(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// What the downdate needs to know; filled on the device by large_decide / large_compass_setup
// (ekf_large.cu keeps the full structure, this is the prefix both translation units agree on).
struct TmaParams {
  const int* decision;     // null for the compass variant
  const int* n_dim;        // live dimension n
  const double* m0;        // -sign(d0)
  const double* m1;        // -sign(d1) (unused for rank 1)
  const double2* W;
  int* nlm_out;            // for the New bookkeeping the plain kernel also does (null: look-ahead run, done earlier)
  const int* n_lm;
  int reverse;             // walk the tiles back to front (see ekf_large_tma_downdate)
  int c0, c1;              // the tensor map covers columns [c0, c1) of the matrix (a shard's slab; 0, INT_MAX: all)
  int early_trigger;       // programmatic dependent launch: let the next kernel be scheduled at once
};

template <int RANK, bool COMPASS, int STAGES>
__global__ void __launch_bounds__(kThreadsTma) large_downdate_tma(const TmaParams q, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem[];
  double* tiles = reinterpret_cast<double*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * kTileBytes);
  uint64_t* done = full + STAGES;
  if (q.early_trigger) ekf_pdl_trigger();
  ekf_pdl_wait();
  if (!COMPASS) {
    const int dec = *q.decision;
    if (dec != EKF_DEC_OLD) {
      if (dec == EKF_DEC_NEW && q.nlm_out && blockIdx.x == 0 && threadIdx.x == 0) *q.nlm_out = *q.n_lm + 1;
      return;
    }
  }
  const int n = *q.n_dim;
  const double m0 = *q.m0, m1 = RANK == 2 ? *q.m1 : 0.0;
  const int ncols = (q.c1 < n ? q.c1 : n) - q.c0;     // live columns of this slab
  if (ncols <= 0) return;
  const int n_rt = (n + TR - 1) / TR, n_ct = (ncols + TC - 1) / TC;
  const long n_tiles = (long)n_rt * n_ct;
  const long first = blockIdx.x;
  const long count = first < n_tiles ? (n_tiles - first + gridDim.x - 1) / gridDim.x : 0;
  const int tid = threadIdx.x;
  const long t_sign = q.reverse ? -1 : 1, t_base = q.reverse ? n_tiles - 1 - first : first;
  auto tile_of = [&](long k) { return t_base + t_sign * k * (long)gridDim.x; };
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], 1); }
    fence_barrier_init();
  }
  __syncthreads();

  if (tid >= kConsumers) {
    // ---- producer: one thread drives the TMA engine ------------------------------------------------
    if (tid == kConsumers) {
      auto issue = [&](long k) {
        const long tile = tile_of(k);
        const int rt = (int)(tile % n_rt), ct = (int)(tile / n_rt);
        const int s = (int)(k % STAGES);
        mbar_expect_tx(&full[s], kTileBytes);
        tma_load_2d(tiles + (size_t)s * TR * TC, &tmap, rt * TR, ct * TC, &full[s]);
      };
      for (long k = 0; k < count && k < STAGES; ++k) issue(k);
      for (long k = 0; k < count; ++k) {
        const int s = (int)(k % STAGES);
        mbar_wait(&done[s], (uint32_t)((k / STAGES) & 1));     // consumers finished tile k in stage s
        const long tile = tile_of(k);
        const int rt = (int)(tile % n_rt), ct = (int)(tile / n_rt);
        tma_store_2d(&tmap, rt * TR, ct * TC, tiles + (size_t)s * TR * TC);
        bulk_commit();
        if (k + STAGES < count) {
          bulk_wait_read0();                                     // the store has drained stage s
          issue(k + STAGES);
        }
      }
      bulk_wait0();
    }
    return;
  }

  // ---- consumers: one tile row per thread, updated in place in shared memory ------------------------
  const double2* __restrict__ W = q.W;
  for (long k = 0; k < count; ++k) {
    const int s = (int)(k % STAGES);
    const long tile = tile_of(k);
    const int rt = (int)(tile % n_rt), ct = (int)(tile / n_rt);
    const double2 wi = W[rt * TR + tid];
    const double u0 = m0 * wi.x, u1 = m1 * wi.y;
    double2 wj[TC];
#pragma unroll
    for (int c = 0; c < TC; ++c) wj[c] = W[q.c0 + ct * TC + c];
    mbar_wait(&full[s], (uint32_t)((k / STAGES) & 1));
    double* t = tiles + (size_t)s * TR * TC + tid;
#pragma unroll
    for (int c = 0; c < TC; ++c) {
      double v = t[c * TR];
      if (RANK == 2) v = fma(u1, wj[c].y, v);
      v = fma(u0, wj[c].x, v);
      t[c * TR] = v;
    }
    fence_proxy_async();                 // generic-proxy writes -> visible to the TMA store
    named_barrier(1, kConsumers);
    if (tid == 0) mbar_arrive(&done[s]);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

size_t ekf_large_tma_map_bytes() { return sizeof(CUtensorMap); }

// Tensor map of one filter's capacity slab: dim0 = rows (contiguous, extent ld), dim1 = columns
// (extent cap_n, stride ld*8 bytes), box 128 x 16.
cudaError_t ekf_large_tma_encode(void* map_out, double* P, int cap_n, int ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return cudaErrorNotSupported;
  const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)cap_n};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {TR, TC};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(reinterpret_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, P, dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int STAGES>
static cudaError_t tma_prepare(int sm_count, int* grid) {
  const size_t bytes = (size_t)STAGES * kTileBytes + 2 * STAGES * sizeof(uint64_t);
  cudaError_t e = cudaFuncSetAttribute(large_downdate_tma<2, false, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(large_downdate_tma<1, true, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, large_downdate_tma<2, false, STAGES>, kThreadsTma, bytes);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidConfiguration;
  if (per_sm > 2) per_sm = 2;
  *grid = per_sm * sm_count;
  return cudaSuccess;
}
cudaError_t ekf_large_tma_prepare(int sm_count, int* grid, bool slim) {
  return slim ? tma_prepare<kStagesSlim>(sm_count, grid) : tma_prepare<kStagesWide>(sm_count, grid);
}

template <int STAGES>
static cudaError_t tma_launch(const TmaParams& q, const CUtensorMap* m, int grid, bool compass, cudaStream_t s) {
  const size_t bytes = (size_t)STAGES * kTileBytes + 2 * STAGES * sizeof(uint64_t);
  if (compass) return ekf_launch_pdl(large_downdate_tma<1, true, STAGES>, grid, kThreadsTma, bytes, s, q, *m);
  return ekf_launch_pdl(large_downdate_tma<2, false, STAGES>, grid, kThreadsTma, bytes, s, q, *m);
}

cudaError_t ekf_large_tma_downdate(const EkfLargeTmaArgs& t, const void* map, int grid, bool compass, cudaStream_t s) {
  TmaParams q;
  q.decision = t.decision; q.n_dim = t.n_dim; q.m0 = t.m0; q.m1 = t.m1; q.W = t.W; q.nlm_out = t.nlm_out; q.n_lm = t.n_lm;
  q.reverse = t.reverse;
  q.c0 = t.c0; q.c1 = t.c1 > t.c0 ? t.c1 : 0x7fffffff;
  q.early_trigger = !t.no_early_trigger;
  const CUtensorMap* m = reinterpret_cast<const CUtensorMap*>(map);
  return t.slim ? tma_launch<kStagesSlim>(q, m, grid, compass, s) : tma_launch<kStagesWide>(q, m, grid, compass, s);
}
