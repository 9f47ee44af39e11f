/* ekf_slam_b200.h — C ABI of the B200-native EKF-SLAM filter core (libekf_slam_b200.so).
 *
 * This library replaces ONE path of kentsommer/2D-EKF-SLAM: the filter arithmetic in
 * odometry/kalmanfilter.cpp, odometry/Propagate.cpp and odometry/Update.cpp. Every entry point
 * cites the reference interface it stands in for. The host-side C++ class `KalmanFilter`
 * (2d-ekf-slam_b200/host/kalmanfilter.h) keeps the reference's call surface
 * (odometry/kalmanfilter.h:24-32) on top of these functions, so slam.cpp:130-182 can drive it
 * unchanged; INTEGRATION.md shows the binding.
 *
 * All arithmetic is IEEE FP64 on the GPU (hand-written sm_100a kernels, no tensor cores, no CPU
 * fallback: every call fails with EKF_ERR_NO_DEVICE / EKF_ERR_CUDA when no B200 is usable).
 *
 * Conventions
 *   - state x: [X, Y, Phi, L1x, L1y, L2x, L2y, ...], n = 3 + 2*n_landmarks (Update.cpp:106);
 *   - covariance P: column-major, leading dimension `ld` >= n, bit-symmetric at call boundaries
 *     (the reference symmetrises after every operation: Propagate.cpp:66-67, Update.cpp:193-194,
 *     kalmanfilter.cpp:123-124);
 *   - 2x2 matrices (R) are column-major: {R00, R10, R01, R11} (Eigen's data() order);
 *   - a handle owns `n_filters` independent filters on one device; batched arrays are indexed
 *     [filter] first; host pointers unless stated otherwise;
 *   - calls are stream-ordered on the handle's stream and return after enqueueing unless they
 *     have host outputs; ekf_sync() waits and reports sticky device-side errors;
 *   - not thread-safe per handle; distinct handles (e.g. one per GPU) may be used concurrently.
 */
#ifndef EKF_SLAM_B200_H
#define EKF_SLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------ */
#define EKF_OK 0
#define EKF_ERR_CUDA 1          /* a CUDA runtime call failed; see ekf_last_error()            */
#define EKF_ERR_BAD_ARG 2
#define EKF_ERR_CAPACITY 3      /* a "New" association arrived with the map full; it was dropped */
#define EKF_ERR_NO_DEVICE 4     /* no CUDA device / not sm_100                                  */
#define EKF_ERR_UNSUPPORTED 5

/* ---- data-association decision codes (Update.cpp:152,181,191) ------------------------------ */
#define EKF_DECISION_NONE (-1)  /* no measurement in this slot                                  */
#define EKF_DECISION_NEW 0      /* Opt_i==0 || Mahal > Gamma_max : state augmented              */
#define EKF_DECISION_OLD 1      /* Mahal < Gamma_min : gain + state + covariance update         */
#define EKF_DECISION_IGNORE 2   /* Gamma_min <= Mahal <= Gamma_max                              */
#define EKF_DECISION_DROPPED 3  /* would be New but the map is at capacity (EKF_ERR_CAPACITY)   */

/* ---- regimes ----------------------------------------------------------------------------- */
#define EKF_REGIME_AUTO 0
#define EKF_REGIME_BATCH 1      /* one CTA per filter; covariance resident in shared memory     */
#define EKF_REGIME_LARGE 2      /* whole grid per filter; covariance streamed from HBM          */

/* ---- fused-kernel variants of the batch regime (same arithmetic, bit-identical results) ------- */
#define EKF_BATCH_KERNEL_AUTO 0   /* STILE when max_landmarks <= 62 (measured fastest), else SMEM  */
#define EKF_BATCH_KERNEL_SMEM 1   /* covariance resident in shared memory                         */
#define EKF_BATCH_KERNEL_TILE 2   /* covariance's lower block triangle resident in registers      */
#define EKF_BATCH_KERNEL_STILE 3  /* lower block triangle tiled in shared memory, 4 filters per SM */
#define EKF_BATCH_KERNEL_DTILE 4  /* as STILE with the O(n^2) downdate deferred (tiles swept once per
                                     three updates, strip / diagonal blocks updated eagerly), exact
                                     gating with two lanes per landmark; max_landmarks <= 50         */

typedef struct ekf_handle_s* ekf_handle;

/* Every tuning constant of the reference is a literal; they become defaulted fields here
 * (defaults are the reference literals, bit for bit). */
typedef struct ekf_config {
  double sigma_v;        /* 0.01          kalmanfilter.cpp:28  */
  double sigma_w;        /* 0.04          kalmanfilter.cpp:29  */
  double deg2rad_pi;     /* 3.141592654   kalmanfilter.cpp:19  */
  double two_pi;         /* 6.283185307   kalmanfilter.cpp:99  */
  double cond_max;       /* 80            Update.cpp:131       */
  double mahal_init;     /* 999999999999  kalmanfilter.h:17    */
  int32_t gamma_max;     /* 50            kalmanfilter.cpp:67  */
  int32_t gamma_min;     /* 10            kalmanfilter.cpp:68  */
  int32_t regime;        /* EKF_REGIME_*                        */
  int32_t batch_kernel;  /* EKF_BATCH_KERNEL_* : which fused kernel the batch regime runs */
} ekf_config;

void ekf_default_config(ekf_config* cfg);

/* ---- step records (input of the fused multi-step path) -------------------------------------
 * One record drives one iteration of slam.cpp:130-182 for one filter:
 *   doPropagation(dt) with the robot reporting getVel()=vel_mm_s, getRotVel()=rotvel_deg_s
 *   (kalmanfilter.cpp:17-20), then doUpdateCompass(compass_z, compass_R) if has_compass
 *   (slam.cpp:144-147), then one doUpdate per measurement, in order (slam.cpp:150-171).
 * Layout, in doubles:  [0] vel_mm_s [1] rotvel_deg_s [2] dt [3] compass_z [4] compass_R
 *                      [5] n_z (0..max_meas) [6] has_compass (0/1) [7] reserved (0)
 *                      then max_meas x { z0, z1, R00, R10, R01, R11 }
 * Records are stored [filter][step][EKF_RECORD_LEN(max_meas)]. */
#define EKF_RECORD_HEADER 8
#define EKF_RECORD_LEN(max_meas) (EKF_RECORD_HEADER + 6 * (max_meas))

/* ---- lifetime ------------------------------------------------------------------------------ */
/* Replaces KalmanFilter::KalmanFilter (kalmanfilter.cpp:4-12) for n_filters filters at once:
 * x = 0 (3), P = 0 (3x3), no landmarks. max_landmarks is the per-filter capacity N_cap. */
int ekf_create(ekf_handle* out, int device, int n_filters, int max_landmarks, const ekf_config* cfg);
int ekf_destroy(ekf_handle h);
int ekf_reset(ekf_handle h);                       /* back to kalmanfilter.cpp:7-11 */
/* The reference grows its state without bound (Update.cpp:158-177, kalmanfilter.cpp:76-84: state and
 * covariance are re-allocated after every update). Here capacity is fixed per handle; ekf_resize
 * replaces *h by a handle of capacity new_max_landmarks (>= every filter's current landmark count) on
 * the same device with the same configuration and the same state (device-to-device copy), and destroys
 * the old one. The regime is chosen again for the new capacity. The drop-in class calls it before an
 * update that could overflow, which restores the reference's behaviour up to device memory. */
int ekf_resize(ekf_handle* h, int new_max_landmarks);
int ekf_n_filters(ekf_handle h);
int ekf_max_landmarks(ekf_handle h);
int ekf_regime(ekf_handle h);                      /* the regime actually selected */
int ekf_set_batch_kernel(ekf_handle h, int batch_kernel);   /* EKF_BATCH_KERNEL_*, for later fused runs */
/* Which O(n^2) covariance sweep a large-regime handle runs: 1 = large_downdate_tma (TMA-staged tiles,
 * the default), 0 = large_downdate (plain double2 sweep, EKF_LARGE_TMA=0 at creation), -1 = not a
 * large-regime handle. */
int ekf_large_downdate_kernel(ekf_handle h);

/* ---- state access (also checkpoint / state injection) -------------------------------------- */
/* The reference keeps state/covariance private (kalmanfilter.h:37-38); these are the harness
 * hooks SURVEY.md 8b asks for. P must be bit-symmetric (EKF_ERR_BAD_ARG otherwise). */
int ekf_set_state(ekf_handle h, int filter, int n_landmarks, const double* x, const double* P, int ld);
int ekf_get_state(ekf_handle h, int filter, int* n_landmarks, double* x, double* P, int ld);
/* A sub-block P(r0:r0+nr, c0:c0+nc) of one filter's covariance, column-major into out (ld_out >= nr).
 * Used by the drop-in class for the covRun.txt row the reference writes (kalmanfilter.cpp:51). */
int ekf_get_cov_block(ekf_handle h, int filter, int r0, int c0, int nr, int nc, double* out, int ld_out);
/* Public mirrors X, Y, Phi, Num_Landmarks (kalmanfilter.h:24-27) for every filter:
 * xyphi[n_filters][3], n_landmarks[n_filters] (either may be NULL). Synchronises. */
int ekf_get_pose(ekf_handle h, double* xyphi, int32_t* n_landmarks);

/* ---- per-call surface (P goes through HBM once per loop iteration) ----------------------------- */
/* KalmanFilter::doPropagation (kalmanfilter.cpp:15-62 -> Propagate.cpp:15-75), minus the two
 * ofstream side effects. Arrays of n_filters; dt may be a single value when dt_stride == 0.
 * Batch regime: the inputs are copied and the call is HELD BACK - if the next call on the handle is
 * ekf_update with n_z = 1 (the slam.cpp:136 -> :170 sequence) both run in ONE launch of the fused
 * kernel (covariance read and written once for the pair, no host round trip unless outputs are
 * requested); any other call first runs the propagation on its own. Results do not depend on which
 * way a call was executed. */
int ekf_propagate(ekf_handle h, const double* vel_mm_s, const double* rotvel_deg_s, const double* dt,
                  int dt_stride);
/* KalmanFilter::doUpdate (kalmanfilter.cpp:64-90 -> Update.cpp:22-204) with n_z measurements per
 * filter, processed sequentially as Update.cpp:80-195 does - including its quirk that the gating
 * bound n_lm is read once at call entry (Update.cpp:26): a landmark added by measurement j is not a
 * candidate for measurements j+1..n_z of the same call (slam.cpp:150-171 itself always calls with
 * n_z = 1, which is what ekf_run's M slots per step are). z[n_filters][n_z][2],
 * R[n_filters][n_z][4]. Optional outputs [n_filters][n_z] (NULL to skip; non-NULL synchronises):
 * decision (EKF_DECISION_*), lm_index (state index Opt_i of the associated landmark, or of the
 * new landmark for New), mahal (Mahal_dist after the gating loop). */
int ekf_update(ekf_handle h, int n_z, const double* z, const double* R, int32_t* decision, int32_t* lm_index,
               double* mahal);
/* KalmanFilter::doUpdateCompass (kalmanfilter.cpp:96-130). z[n_filters], R[n_filters];
 * valid[n_filters] may be NULL (= all valid). */
int ekf_update_compass(ekf_handle h, const double* z, const double* R, const uint8_t* valid);

/* ---- fused multi-step path (covariance stays on chip for all steps) --------------------------- */
typedef struct ekf_run_outputs {
  int32_t* decision;    /* [F][T][M] or NULL */
  int32_t* lm_index;    /* [F][T][M] or NULL */
  double* mahal;        /* [F][T][M] or NULL */
  double* pose_trace;   /* [F][T][3] or NULL : X,Y,Phi after each step (slam.cpp:181)  */
  double* final_pose;   /* [F][3]    or NULL */
  int32_t* final_nlm;   /* [F]       or NULL */
} ekf_run_outputs;

/* End-to-end: copies `records` (host; pinned memory from ekf_host_alloc makes the copy async)
 * to the device, runs n_steps iterations of the slam.cpp loop for every filter, copies the
 * requested outputs back and synchronises. */
int ekf_run(ekf_handle h, int n_steps, int max_meas, const double* records, const ekf_run_outputs* out);
/* Same, split so inputs can stay resident in HBM: upload once, run many times. */
int ekf_upload_records(ekf_handle h, int n_steps, int max_meas, const double* records);
int ekf_run_resident(ekf_handle h, int want_trace, int want_pose_trace);   /* enqueue only */
int ekf_download_outputs(ekf_handle h, const ekf_run_outputs* out);      /* synchronises */

/* ---- plumbing -------------------------------------------------------------------------------- */
int ekf_sync(ekf_handle h);                         /* waits; returns sticky EKF_ERR_CAPACITY etc. */
/* Waits, then counts the filters whose sticky capacity flag is set (a New association was dropped since
 * the last ekf_reset / ekf_set_state / clear) and, if `clear`, clears the flags. */
int ekf_capacity_flags(ekf_handle h, int* n_flagged, int clear);
const char* ekf_last_error(ekf_handle h);           /* h may be NULL: last create() failure        */
void* ekf_host_alloc(size_t bytes);                 /* pinned host memory (cudaMallocHost)         */
void ekf_host_free(void* p);
int ekf_timer_start(ekf_handle h);                  /* cudaEventRecord on the handle's stream      */
int ekf_timer_stop(ekf_handle h, float* ms);        /* records, synchronises, elapsed milliseconds */
long long ekf_kernel_launches(ekf_handle h);        /* kernels this handle has launched so far     */
/* Average device time of the dominant kernel since the last call (ms) and how many launches that
 * covers; measured with CUDA events around each launch on the handle's stream. */
int ekf_kernel_time(ekf_handle h, float* avg_ms, int* n_launches);
int ekf_device_count(int* n_devices);                /* CUDA devices visible to the process */
int ekf_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin,
                    size_t* total_mem);
/* DFMA-chain microbenchmark on `device`: achieved FP64 FLOP/s (2 flop per DFMA). Used as the
 * measured FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 entry). */
int ekf_measure_fp64_peak(int device, double* flops_per_s);
/* Profiling aid for the register-tile fused kernel, per device (the current one): the first call
 * (out16 may be NULL) enables per-phase cycle accumulation by CTA 0 of every following launch; later
 * calls read and clear the 16 counters {scalar chains, covariance propagate, gating, column publish,
 * gain rows, downdate, step epilogue, unused, then 8 finer probes along thread 0 (only in
 * -DEKF_FINE_TIMING builds)}. */
int ekf_debug_phase_cycles(long long* out16);
/* Profiling aid for the shared-memory tiled fused kernel (-DEKF_STILE_TIMING builds, zeros
 * otherwise): clock64 at which lane 0 of each warp of CTA 0 arrived at each barrier of one step
 * (first filter, step 500), out128 = [8 warps][16 barriers]. */
int ekf_debug_stile_timestamps(long long* out128);
/* Same for the deferred-downdate kernel (-DEKF_DTILE_TIMING builds): out64 = [4 warps][16 stamps]. */
int ekf_debug_dtile_timestamps(long long* out64);
/* Test hook (no GPU needed): the chunk boundaries ekf_run() would use to pipeline n_filters filters when
 * `wave` filters are co-resident on the device. begin[0..n] with begin[n] = n_filters; capacity >= 49;
 * returns n (>= 1) or -EKF_ERR_BAD_ARG. */
int ekf_debug_pipeline_chunks(long long n_filters, long long wave, long long* begin, int capacity);

/* ---- one large map sharded over several GPUs ---------------------------- */
/* The covariance of ONE map is split by columns over n_shards devices (one process drives them;
 * every pair needs CUDA peer access = NVLink on an NVSwitch node). Same reference calls as above,
 * same bits as the single-GPU large regime for every shard count: the state vector, P_RR and the
 * landmark count are replicated, the O(n^2) covariance downdate (Update.cpp:188,193-194) runs on
 * each shard's own columns, and the gain kernel stores its rows of K / W / x straight into every
 * peer's memory before the sweep (the one exchange step the path has). devices may repeat an
 * ordinal (several shards on one GPU), which is how the logic is tested on a single-GPU box. */
#define EKF_SHARDED_MAX_SHARDS 8
typedef struct ekf_sharded_s* ekf_sharded;
int ekf_sharded_create(ekf_sharded* out, int n_shards, const int* devices, int max_landmarks, const ekf_config* cfg);
int ekf_sharded_destroy(ekf_sharded m);
/* kalmanfilter.cpp:7-11 */
int ekf_sharded_reset(ekf_sharded m);
int ekf_sharded_n_shards(ekf_sharded m);
int ekf_sharded_max_landmarks(ekf_sharded m);
/* Column range [c0, c1) of the covariance held by one shard. */
int ekf_sharded_columns(ekf_sharded m, int shard, int* c0, int* c1);
/* As ekf_set_state / ekf_get_state / ekf_get_pose for the one map (P column-major, bit-symmetric). */
int ekf_sharded_set_state(ekf_sharded m, int n_landmarks, const double* x, const double* P, int ld);
int ekf_sharded_get_state(ekf_sharded m, int* n_landmarks, double* x, double* P, int ld);
int ekf_sharded_get_pose(ekf_sharded m, double* xyphi, int32_t* n_landmarks);
/* Test hook: the replicas one shard holds (x[0:n], P_RR column-major, landmark count). */
int ekf_sharded_get_replica(ekf_sharded m, int shard, int* n_landmarks, double* x, double* PRR9);
/* KalmanFilter::doPropagation / doUpdate (n_z <= 16, gating bound frozen at call entry as
 * Update.cpp:26) / doUpdateCompass for the sharded map; each call synchronises. */
int ekf_sharded_propagate(ekf_sharded m, double vel_mm_s, double rotvel_deg_s, double dt);
int ekf_sharded_update(ekf_sharded m, int n_z, const double* z, const double* R, int32_t* decision, int32_t* lm_index,
                       double* mahal);
int ekf_sharded_update_compass(ekf_sharded m, double z, double R);
/* slam.cpp:127-182 for n_steps step records of the one map (layout above, F = 1): all kernels and
 * exchange steps of all steps are enqueued without a host round trip. By default the run overlaps
 * propagate / gating / decision of the next operation with the covariance sweep of the current one
 * (a replicated O(n) cache on a side stream per shard, exchange steps as flags in peer memory, the
 * TMA-staged sweep); EKF_SHARD_LOOKAHEAD=0 in the environment at ekf_sharded_create() selects the chain
 * with an event exchange per gating pass and per gain (also what the per-call functions above run, and
 * what a map runs whose shards share a device - flags need one GPU per shard). Same results either way. */
int ekf_sharded_run(ekf_sharded m, int n_steps, int max_meas, const double* records, const ekf_run_outputs* out);
/* How ekf_sharded_run sweeps the covariance: 2 = look-ahead run with the TMA-staged sweep, 1 = look-ahead
 * run with the plain double2 sweep (EKF_LARGE_TMA=0), 0 = event chain with the plain sweep. */
int ekf_sharded_run_mode(ekf_sharded m);
/* Device time of the last ekf_sharded_run (events on shard 0 bracketing cross-shard barriers) and
 * of one covariance downdate sampled mid-run on shard 0 (0 if none ran). */
int ekf_sharded_last_run_ms(ekf_sharded m, float* run_ms, float* downdate_ms);
long long ekf_sharded_kernel_launches(ekf_sharded m);
const char* ekf_sharded_last_error(ekf_sharded m);

#ifdef __cplusplus
}
#endif
#endif /* EKF_SLAM_B200_H */
