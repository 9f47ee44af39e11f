// ekf_internal.h — host-side declarations shared by the C ABI (ekf_api.cu) and the kernel
// translation units (ekf_batch.cu, ekf_large.cu). Not part of the public interface.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "ekf_small.cuh"

#define EKF_MAX_MEAS 16                              // record capacity the kernels are sized for
#define EKF_RECORD_LEN_MAX (8 + 6 * EKF_MAX_MEAS)

// Device-resident state of a handle. x: [F][xs], P: [F][slab] (column-major, leading dim ld).
struct EkfState {
  double* x;
  double* P;
  int* nlm;
  int* status;      // sticky per-filter flags (bit 0: a New was dropped at capacity)
  int F, cap_lm, cap_n, ld;
  size_t xs;        // doubles per filter in x (cap_n rounded up to even)
  size_t slab;      // doubles per filter in P (cap_n * ld)
};

// Step-record stream + optional trace outputs of the fused path (all device pointers).
struct EkfRunIO {
  const double* records;   // [F][T][L]
  int T, M, L;
  int* decision;           // [F][T][M] or null
  int* index;              // [F][T][M] or null
  double* mahal;           // [F][T][M] or null
  double* pose_trace;      // [F][T][3] or null
  // Growth beyond a kernel's tile capacity (ekf_stile.cu): a filter whose map outgrows the kernel that
  // runs it is written back and "parked"; resume[f] = 1 + t*(M+1) + m says where (step t propagated,
  // next measurement m; -1 = before step 0). A continuation launch (continuation = 1) of a kernel with
  // larger tiles visits only the parked filters and picks up there. null = no parking (New at capacity
  // is dropped and flagged).
  int* resume;             // [F] or null
  int continuation;        // 0: primary (parks); 1: resumes the parked filters; 2: runs the filters marked -2 from the start
};

struct EkfPercallIO {
  // propagate
  const double* vel; const double* rotvel; const double* dt; int dt_stride;
  // update: zr [F][n_z][6] = {z0,z1,R00,R10,R01,R11} (the record's measurement slot layout);
  // outputs [F][n_z] or null
  int n_z; const double* zr;
  int* decision; int* index; double* mahal;
  // compass
  const double* cz; const double* cR; const uint8_t* cvalid;
};

enum EkfOp { EKF_OP_PROPAGATE = 0, EKF_OP_UPDATE = 1, EKF_OP_COMPASS = 2 };

// ---- regime A: one CTA per filter (ekf_batch.cu) ------------------------------------------------
size_t ekf_batch_smem_bytes(int cap_n, int ld, int L);
// Largest landmark capacity whose covariance fits in one CTA's shared memory on `device`.
int ekf_batch_max_landmarks(size_t smem_optin);
cudaError_t ekf_batch_prepare(int cap_n, int ld, int max_L, int sm_count, int* grid_cap);
cudaError_t ekf_batch_run(const EkfState& st, const EkfRunIO& io, const EkfConst& k, int grid_cap,
                          cudaStream_t stream);
cudaError_t ekf_batch_percall(const EkfState& st, const EkfPercallIO& io, EkfOp op, const EkfConst& k,
                              cudaStream_t stream);

// Register-tile variant of the fused path (ekf_tile.cu): covariance half in registers as 8x8 tiles.
int ekf_tile_max_landmarks();
cudaError_t ekf_tile_phase_cycles(long long* out);   // profiling aid, see ekf_tile.cu
cudaError_t ekf_tile_run(const EkfState& st, const EkfRunIO& io, const EkfConst& k, int sm_count, cudaStream_t stream);

// Shared-memory tiled-triangle variant of the fused path (ekf_stile.cu): four filters per SM.
int ekf_stile_max_landmarks();
int ekf_stile_ctas_per_sm(int cap_lm);
cudaError_t ekf_stile_timestamps(long long* out128);
cudaError_t ekf_stile_run(const EkfState& st, const EkfRunIO& io, const EkfConst& k, int sm_count, cudaStream_t stream,
                          int tile_cap = 0);
int ekf_stile_fast_landmarks();   // capacity of the four-filters-per-SM instance
cudaError_t ekf_stile_mark_grown(const int* nlm, int* resume, int F, cudaStream_t stream);

// Deferred-downdate variant (ekf_dtile.cu): eager strip / diagonal blocks, P_LL tiles swept once per
// three updates, exact gating with two lanes per landmark; four filters per SM, <= 50 landmarks.
int ekf_dtile_max_landmarks();
int ekf_dtile_ctas_per_sm();
cudaError_t ekf_dtile_timestamps(long long* out64);
cudaError_t ekf_dtile_run(const EkfState& st, const EkfRunIO& io, const EkfConst& k, int sm_count, cudaStream_t stream);

// ---- regime B: whole grid per filter, covariance streamed from HBM (ekf_large.cu) ---------------
struct EkfLargeWork {      // device scratch owned by the handle
  double2* W;              // [cap_n + 2]  downdate vectors
  double* cand_val;        // [grid]   per-CTA gating minima
  int* cand_idx;           // [grid]
  double* small;           // LargeSmall (setup, winner, decision), ekf_large_small_doubles() doubles
  int grid;                // CTAs of the downdate sweep (multiple of the SM count)
  // TMA-staged downdate (ekf_large_tma.cu): one tensor map per filter (host copies), persistent grid
  const unsigned char* tmaps;   // [F][ekf_large_tma_map_bytes()] or null
  int tma_grid;
  int use_tma;
  // look-ahead runs (ekf_large.cu header): O(n) cache of what the next operation's gating reads, a side
  // stream for the chain that works on the cache alone, and the two events that join the streams
  double* strip;           // [3][lds]
  double* diag;            // [cap_lm][4]
  int lds;
  unsigned* sweep_seq;     // host counter of sweeps launched (direction of the next one)
  int snake;
  int la;                  // 1: ekf_large_run overlaps gating / decision with the previous sweep
  cudaStream_t s_side;
  cudaEvent_t ev_a, ev_b;
};

// Device addresses the TMA downdate reads its control state from (fields of LargeSmall).
struct EkfLargeTmaArgs {
  const int* decision;     // null for the compass variant
  const int* n_dim;
  const double* m0;
  const double* m1;
  const double2* W;
  int* nlm_out;
  const int* n_lm;
  // Consecutive sweeps walk the tiles in opposite directions: what one sweep touched last is what the next
  // touches first, so a covariance of about the size of the 126 MB L2 is largely served from it.
  int reverse;
  // A column-sharded map (ekf_shard.cu): the tensor map covers columns [c0, c1) only; 0, 0 = the whole matrix.
  int c0, c1;
  int no_early_trigger;    // 1: dependents are released when the sweep ends (chains without events, ekf_pdl.cuh)
  int slim;                // 1: the four-stage instance (leaves shared memory for kernels running beside the sweep)
};
size_t ekf_large_tma_map_bytes();
cudaError_t ekf_large_tma_encode(void* map_out, double* P, int cap_n, int ld);
cudaError_t ekf_large_tma_prepare(int sm_count, int* grid, bool slim = false);
cudaError_t ekf_large_tma_downdate(const EkfLargeTmaArgs& t, const void* map, int grid, bool compass, cudaStream_t s);

// Optional CUDA-event sampling of the dominant (downdate) kernel: every `every`-th launch is
// bracketed by ev0[k]/ev1[k] until `cap` pairs are used.
struct EkfLargeTiming {
  cudaEvent_t* ev0;
  cudaEvent_t* ev1;
  int cap, used, every, seen;
};
cudaError_t ekf_large_prepare(int sm_count, int* grid);
size_t ekf_large_small_doubles();
int ekf_large_launches_per(EkfOp op);
cudaError_t ekf_large_percall(const EkfState& st, int filter, const EkfPercallIO& io, EkfOp op, const EkfConst& k,
                              const EkfLargeWork& wk, EkfLargeTiming* tm, cudaStream_t stream);
// has_compass / n_z: host mirrors [T] of the record flags of this filter.
cudaError_t ekf_large_run(const EkfState& st, int filter, const EkfRunIO& io, const uint8_t* has_compass,
                          const uint8_t* n_z, const EkfConst& k, const EkfLargeWork& wk, EkfLargeTiming* tm,
                          cudaStream_t stream, long long* launches);

// ---- misc ---------------------------------------------------------------------------------------
cudaError_t ekf_fp64_peak(double* flops_per_s, cudaStream_t stream);
