// ekf_api.cu — the C ABI of include/ekf_slam_b200.h: handle lifetime, device memory, staging of
// host buffers, stream-ordered launches of the regime A / regime B kernels. No arithmetic of the
// filter happens on the host, and nothing here falls back to a CPU path: without a usable sm_100
// device every entry point reports EKF_ERR_NO_DEVICE / EKF_ERR_CUDA.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ekf_internal.h"
#include "ekf_slam_b200.h"

namespace {

std::string g_create_error;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;   // elements
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

constexpr int kTimingPairs = 64;
constexpr int kMaxChunks = 48;

}  // namespace

struct ekf_handle_s {
  int device = 0;
  int regime = EKF_REGIME_BATCH;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;   // copy engines of the pipelined end-to-end path
  cudaStream_t s_grow = nullptr;                  // continuation launches of the growth path, beside the primary launch
  cudaEvent_t ev_grow0 = nullptr, ev_grow1 = nullptr;
  cudaEvent_t ev_in[kMaxChunks], ev_k[kMaxChunks], ev_done = nullptr;
  EkfConst k{};
  ekf_config cfg{};
  EkfState st{};
  int grid_cap = 0;           // regime A: co-resident CTAs
  EkfLargeWork wk{};          // regime B scratch
  unsigned sweep_seq = 0;
  // staging
  DevBuf<double> in;          // per-call inputs
  DevBuf<int> o_dec, o_idx;
  DevBuf<double> o_mah;
  DevBuf<uint8_t> in_valid;
  // fused path
  DevBuf<double> records;
  DevBuf<int> t_dec, t_idx;
  DevBuf<double> t_mah, t_pose;
  DevBuf<int> resume;         // [F] park codes of the two-launch growth path (launch_batch_kernel)
  // per-call surface of the batch regime: a doPropagation call is held back until the next call shows
  // whether it can ride in one launch with the doUpdate that follows it (slam.cpp:136,170)
  bool prop_pending = false;
  std::vector<double> pend;   // [3][F]: vel, rotvel, dt of the held-back call
  // [F][14] step records of the fused propagate+update launch: a ring of pinned staging buffers, each with
  // its own device buffer, copied on the upload stream so the copy of call k+1 overlaps the kernel of call k
  DevBuf<double> pc_rec[4];
  double* stage[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t stage_ev[4] = {nullptr, nullptr, nullptr, nullptr};    // the copy out of stage[i] is done
  cudaEvent_t stage_kev[4] = {nullptr, nullptr, nullptr, nullptr};   // the kernel that read pc_rec[i] is done
  int stage_i = 0;
  int rec_T = 0, rec_M = 0, rec_L = 0;
  bool have_trace = false, have_pose_trace = false;
  std::vector<uint8_t> flag_compass, flag_nz;   // host mirror [F][T] of record flags (regime B)
  std::vector<unsigned char> tmaps;             // tensor maps of the TMA-staged downdate (regime B)
  // timing
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;     // user timer
  cudaEvent_t kev0[kTimingPairs], kev1[kTimingPairs];
  int kev_used = 0;
  EkfLargeTiming ltm{};
  long long launches = 0;
  std::string err;
};

namespace {

int fail(ekf_handle h, int code, const std::string& msg) {
  if (h) h->err = msg;
  else g_create_error = msg;
  return code;
}

#define EKF_CK(h, call)                                                                          \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return fail(h, EKF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));         \
  } while (0)

int check_status(ekf_handle h) {
  std::vector<int> s(h->st.F);
  EKF_CK(h, cudaMemcpyAsync(s.data(), h->st.status, sizeof(int) * h->st.F, cudaMemcpyDeviceToHost, h->stream));
  EKF_CK(h, cudaStreamSynchronize(h->stream));
  for (int f = 0; f < h->st.F; ++f)
    if (s[f] & 1) return fail(h, EKF_ERR_CAPACITY, "filter " + std::to_string(f) + ": landmark capacity exceeded, a New association was dropped");
  return EKF_OK;
}

// Which fused kernel a batch-regime run uses: 1 = shared-memory full matrix, 2 = register tiles,
// 3 = shared-memory tiled triangle, 4 = tiled triangle with the deferred downdate.
int pick_batch_kernel(ekf_handle h) {
  const int want = h->cfg.batch_kernel;
  const int cap = h->st.cap_lm;
  if (want == EKF_BATCH_KERNEL_SMEM) return 1;
  if (want == EKF_BATCH_KERNEL_TILE) return cap <= ekf_tile_max_landmarks() ? 2 : -1;
  if (want == EKF_BATCH_KERNEL_STILE) return cap <= ekf_stile_max_landmarks() ? 3 : -1;
  if (want == EKF_BATCH_KERNEL_DTILE) return cap <= ekf_dtile_max_landmarks() ? 4 : -1;
  if (cap <= ekf_stile_max_landmarks()) return 3;   // measured fastest
  return 1;
}

cudaError_t launch_batch_kernel(ekf_handle h, int kern, const EkfState& st, const EkfRunIO& io) {
  if (kern == 4) return ekf_dtile_run(st, io, h->k, h->sm_count, h->stream);
  if (kern == 3 && h->cfg.batch_kernel == EKF_BATCH_KERNEL_AUTO && st.cap_lm > ekf_stile_fast_landmarks()) {
    // Capacity beyond the four-filters-per-SM tile size (Update.cpp:158-177 grows the map without bound):
    // every filter starts in the fast small-tile instance; the few whose map outgrows it are written
    // back ("parked") and finished, from the measurement where they stopped, by the instance sized for
    // the handle's capacity. Same arithmetic in both, so the result does not depend on where a filter ran.
    // Maps that are already beyond the fast tiles when the run starts (they grew in an earlier run) are
    // marked and run by a continuation launch on a second stream BESIDE the primary launch, so their lap
    // (one CTA per filter for the whole lap) is not a serial tail; only a filter that outgrows the tiles
    // during this run waits for the primary launch to finish.
    cudaError_t e = h->resume.reserve((size_t)h->st.F);
    if (e != cudaSuccess) return e;
    if (!h->s_grow) {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      if ((e = cudaStreamCreateWithPriority(&h->s_grow, cudaStreamNonBlocking, hi)) != cudaSuccess) return e;
      cudaEventCreateWithFlags(&h->ev_grow0, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&h->ev_grow1, cudaEventDisableTiming);
    }
    EkfRunIO io2 = io;
    io2.resume = h->resume.p + (st.nlm - h->st.nlm);
    if ((e = ekf_stile_mark_grown(st.nlm, io2.resume, st.F, h->stream)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(h->ev_grow0, h->stream)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(h->s_grow, h->ev_grow0, 0)) != cudaSuccess) return e;
    io2.continuation = 2;
    if ((e = ekf_stile_run(st, io2, h->k, h->sm_count, h->s_grow, 0)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(h->ev_grow1, h->s_grow)) != cudaSuccess) return e;
    io2.continuation = 0;
    if ((e = ekf_stile_run(st, io2, h->k, h->sm_count, h->stream, ekf_stile_fast_landmarks())) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(h->stream, h->ev_grow1, 0)) != cudaSuccess) return e;
    io2.continuation = 1;
    h->launches += 3;
    return ekf_stile_run(st, io2, h->k, h->sm_count, h->stream, 0);
  }
  if (kern == 3) return ekf_stile_run(st, io, h->k, h->sm_count, h->stream);
  if (kern == 2) return ekf_tile_run(st, io, h->k, h->sm_count, h->stream);
  return ekf_batch_run(st, io, h->k, h->grid_cap, h->stream);
}

void kernel_event_begin(ekf_handle h) {
  if (h->kev_used < kTimingPairs) cudaEventRecord(h->kev0[h->kev_used], h->stream);
}
void kernel_event_end(ekf_handle h) {
  if (h->kev_used < kTimingPairs) cudaEventRecord(h->kev1[h->kev_used++], h->stream);
}

int launch_run(ekf_handle h, bool want_trace, bool want_pose) {
  if (!h->records.p || h->rec_T <= 0) return fail(h, EKF_ERR_BAD_ARG, "no records uploaded");
  const size_t FT = (size_t)h->st.F * h->rec_T, FTM = FT * (h->rec_M > 0 ? h->rec_M : 1);
  EkfRunIO io{};
  io.records = h->records.p;
  io.T = h->rec_T; io.M = h->rec_M > 0 ? h->rec_M : 1; io.L = h->rec_L;   // M >= 1: a slot with n_z = 0 is reported as NONE
  h->have_trace = want_trace;
  h->have_pose_trace = want_pose;
  if (want_trace) {
    EKF_CK(h, h->t_dec.reserve(FTM));
    EKF_CK(h, h->t_idx.reserve(FTM));
    EKF_CK(h, h->t_mah.reserve(FTM));
    io.decision = h->t_dec.p; io.index = h->t_idx.p; io.mahal = h->t_mah.p;
    if (h->regime == EKF_REGIME_LARGE) {   // slots without a measurement are not visited by any kernel
      EKF_CK(h, cudaMemsetAsync(h->t_dec.p, 0xFF, FTM * sizeof(int), h->stream));
      EKF_CK(h, cudaMemsetAsync(h->t_idx.p, 0xFF, FTM * sizeof(int), h->stream));
      EKF_CK(h, cudaMemsetAsync(h->t_mah.p, 0, FTM * sizeof(double), h->stream));
    }
  }
  if (want_pose) {
    EKF_CK(h, h->t_pose.reserve(FT * 3));
    io.pose_trace = h->t_pose.p;
  }
  if (h->regime == EKF_REGIME_BATCH) {
    const int kern = pick_batch_kernel(h);
    if (kern < 0)
      return fail(h, EKF_ERR_UNSUPPORTED, "the requested fused kernel does not support max_landmarks = " + std::to_string(h->st.cap_lm) + " (TILE / STILE: <= " + std::to_string(ekf_tile_max_landmarks()) + ", DTILE: <= " + std::to_string(ekf_dtile_max_landmarks()) + ")");
    kernel_event_begin(h);
    EKF_CK(h, launch_batch_kernel(h, kern, h->st, io));
    kernel_event_end(h);
    h->launches += 1;
  } else {
    for (int f = 0; f < h->st.F; ++f)
      EKF_CK(h, ekf_large_run(h->st, f, io, h->flag_compass.data() + (size_t)f * h->rec_T,
                              h->flag_nz.data() + (size_t)f * h->rec_T, h->k, h->wk, &h->ltm, h->stream,
                              &h->launches));
  }
  return EKF_OK;
}

int download(ekf_handle h, const ekf_run_outputs* out) {
  const size_t F = h->st.F, FT = F * h->rec_T, FTM = FT * (h->rec_M > 0 ? h->rec_M : 1);
  if (out) {
    if ((out->decision || out->lm_index || out->mahal) && !h->have_trace)
      return fail(h, EKF_ERR_BAD_ARG, "trace requested but the run did not record one");
    if (out->pose_trace && !h->have_pose_trace)
      return fail(h, EKF_ERR_BAD_ARG, "pose trace requested but the run did not record one");
    if (out->decision) EKF_CK(h, cudaMemcpyAsync(out->decision, h->t_dec.p, FTM * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (out->lm_index) EKF_CK(h, cudaMemcpyAsync(out->lm_index, h->t_idx.p, FTM * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (out->mahal) EKF_CK(h, cudaMemcpyAsync(out->mahal, h->t_mah.p, FTM * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (out->pose_trace) EKF_CK(h, cudaMemcpyAsync(out->pose_trace, h->t_pose.p, FT * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (out->final_pose)
      EKF_CK(h, cudaMemcpy2DAsync(out->final_pose, 3 * sizeof(double), h->st.x, h->st.xs * sizeof(double),
                                  3 * sizeof(double), F, cudaMemcpyDeviceToHost, h->stream));
    if (out->final_nlm) EKF_CK(h, cudaMemcpyAsync(out->final_nlm, h->st.nlm, F * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  }
  return check_status(h);
}

// Runs a held-back doPropagation call on its own (the per-call propagate kernel).
int flush_pending(ekf_handle h) {
  if (!h->prop_pending) return EKF_OK;
  h->prop_pending = false;
  const size_t F = h->st.F;
  EKF_CK(h, h->in.reserve(3 * F));
  EKF_CK(h, cudaMemcpyAsync(h->in.p, h->pend.data(), 3 * F * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EkfPercallIO io{};
  io.vel = h->in.p; io.rotvel = h->in.p + F; io.dt = h->in.p + 2 * F; io.dt_stride = 1;
  EKF_CK(h, ekf_batch_percall(h->st, io, EKF_OP_PROPAGATE, h->k, h->stream));
  h->launches += 1;
  EKF_CK(h, cudaStreamSynchronize(h->stream));   // h->pend / the staging buffer are reusable
  return EKF_OK;
}
#define EKF_FLUSH(h)                              \
  do {                                            \
    const int rc__ = flush_pending(h);            \
    if (rc__ != EKF_OK) return rc__;              \
  } while (0)

}  // namespace

extern "C" {

void ekf_default_config(ekf_config* cfg) {
  cfg->sigma_v = 0.01;
  cfg->sigma_w = 0.04;
  cfg->deg2rad_pi = 3.141592654;
  cfg->two_pi = 6.283185307;
  cfg->cond_max = 80.0;
  cfg->mahal_init = 999999999999.0;
  cfg->gamma_max = 50;
  cfg->gamma_min = 10;
  cfg->regime = EKF_REGIME_AUTO;
  cfg->batch_kernel = EKF_BATCH_KERNEL_AUTO;
}

int ekf_device_count(int* n_devices) {
  if (!n_devices) return EKF_ERR_BAD_ARG;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return EKF_ERR_NO_DEVICE;
  *n_devices = n;
  return EKF_OK;
}

int ekf_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin, size_t* total_mem) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return EKF_ERR_NO_DEVICE;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return EKF_ERR_CUDA;
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (smem_optin) *smem_optin = p.sharedMemPerBlockOptin;
  if (total_mem) *total_mem = p.totalGlobalMem;
  return EKF_OK;
}

int ekf_create(ekf_handle* out, int device, int n_filters, int max_landmarks, const ekf_config* cfg_in) {
  if (!out) return fail(nullptr, EKF_ERR_BAD_ARG, "out is NULL");
  *out = nullptr;
  if (n_filters < 1 || max_landmarks < 1) return fail(nullptr, EKF_ERR_BAD_ARG, "n_filters and max_landmarks must be >= 1");
  int sms = 0, maj = 0, mnr = 0;
  size_t smem_optin = 0, total = 0;
  int rc = ekf_device_info(device, &sms, &maj, &mnr, &smem_optin, &total);
  if (rc != EKF_OK) return fail(nullptr, EKF_ERR_NO_DEVICE, "no usable CUDA device " + std::to_string(device) + " (this library has no CPU fallback)");
  if (maj != 10) return fail(nullptr, EKF_ERR_NO_DEVICE, "device is sm_" + std::to_string(maj * 10 + mnr) + "; this library is built for sm_100a (B200) only");
  ekf_config cfg;
  if (cfg_in) cfg = *cfg_in;
  else ekf_default_config(&cfg);

  ekf_handle h = new ekf_handle_s();
  h->device = device;
  h->sm_count = sms;
  h->cfg = cfg;
  h->k = EkfConst{cfg.sigma_v, cfg.sigma_w, cfg.deg2rad_pi, cfg.two_pi, cfg.cond_max, cfg.mahal_init, cfg.gamma_max, cfg.gamma_min};
  auto bail = [&](int code, const std::string& msg) {
    g_create_error = msg;
    ekf_destroy(h);
    return code;
  };
  if (cudaSetDevice(device) != cudaSuccess) return bail(EKF_ERR_CUDA, "cudaSetDevice failed");
  const int fit = ekf_batch_max_landmarks(smem_optin);
  int regime = cfg.regime;
  if (regime == EKF_REGIME_AUTO) regime = max_landmarks <= fit ? EKF_REGIME_BATCH : EKF_REGIME_LARGE;
  if (regime == EKF_REGIME_BATCH && max_landmarks > fit)
    return bail(EKF_ERR_UNSUPPORTED, "max_landmarks " + std::to_string(max_landmarks) + " does not fit the shared-memory-resident regime (limit " + std::to_string(fit) + ")");
  if (regime != EKF_REGIME_BATCH && regime != EKF_REGIME_LARGE) return bail(EKF_ERR_BAD_ARG, "bad regime");
  h->regime = regime;

  EkfState& st = h->st;
  st.F = n_filters;
  st.cap_lm = max_landmarks;
  st.cap_n = 3 + 2 * max_landmarks;
  // leading dimension: even (double2 rows); regime B pads columns to 128 bytes
  st.ld = regime == EKF_REGIME_LARGE ? ((st.cap_n + 15) & ~15) : ((st.cap_n + 1) & ~1);
  st.xs = (size_t)((st.cap_n + 1) & ~1);
  st.slab = (size_t)st.cap_n * st.ld;
  const size_t need = (size_t)n_filters * (st.slab + st.xs) * sizeof(double);
  if (need > total) return bail(EKF_ERR_UNSUPPORTED, "state does not fit device memory");
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMalloc(&st.x, (size_t)n_filters * st.xs * sizeof(double))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMalloc(&st.P, (size_t)n_filters * st.slab * sizeof(double))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMalloc(&st.nlm, (size_t)n_filters * sizeof(int))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMalloc(&st.status, (size_t)n_filters * sizeof(int))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  cudaEventCreate(&h->ev0);
  cudaEventCreate(&h->ev1);
  cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming);
  for (int i = 0; i < kMaxChunks; ++i) {
    cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming);
  }
  for (int i = 0; i < kTimingPairs; ++i) {
    cudaEventCreate(&h->kev0[i]);
    cudaEventCreate(&h->kev1[i]);
  }
  if (regime == EKF_REGIME_BATCH) {
    if ((e = ekf_batch_prepare(st.cap_n, st.ld, EKF_RECORD_LEN_MAX, sms, &h->grid_cap)) != cudaSuccess)
      return bail(EKF_ERR_CUDA, std::string("ekf_batch_prepare: ") + cudaGetErrorString(e));
  } else {
    if ((e = ekf_large_prepare(sms, &h->wk.grid)) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
    const size_t w_count = (size_t)st.cap_n + 512;   // zero tail: boundary tiles of the TMA sweep read past n
    if ((e = cudaMalloc(&h->wk.W, w_count * sizeof(double2))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
    if ((e = cudaMalloc(&h->wk.cand_val, (size_t)h->wk.grid * sizeof(double))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
    if ((e = cudaMalloc(&h->wk.cand_idx, (size_t)h->wk.grid * sizeof(int))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
    if ((e = cudaMalloc(&h->wk.small, ekf_large_small_doubles() * sizeof(double))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
    cudaMemset(h->wk.W, 0, w_count * sizeof(double2));
    {
      // look-ahead runs (ekf_large.cu): on unless EKF_LARGE_LOOKAHEAD=0 keeps every kernel on one stream
      const char* env = getenv("EKF_LARGE_LOOKAHEAD");
      h->wk.la = env ? (atoi(env) != 0) : 1;
      env = getenv("EKF_LARGE_SNAKE");         // consecutive sweeps in opposite directions (L2 reuse)
      h->wk.snake = env ? (atoi(env) != 0) : 1;
      h->wk.sweep_seq = &h->sweep_seq;
      h->wk.lds = (st.cap_n + 2 + 7) & ~7;
      if ((e = cudaMalloc(&h->wk.strip, 3 * (size_t)h->wk.lds * sizeof(double))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
      if ((e = cudaMalloc(&h->wk.diag, 4 * ((size_t)st.cap_lm + 1) * sizeof(double))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
      int lo = 0, hi = 0;                      // the side chain's small grids go in front of the sweep's CTAs
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      if ((e = cudaStreamCreateWithPriority(&h->wk.s_side, cudaStreamNonBlocking, hi)) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
      cudaEventCreateWithFlags(&h->wk.ev_a, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&h->wk.ev_b, cudaEventDisableTiming);
    }
    {
      // TMA-staged downdate: on unless EKF_LARGE_TMA=0 selects the plain double2 sweep. No silent
      // fallback: if the tensor maps cannot be encoded the handle is not created (ekf_large_downdate_kernel()
      // reports which sweep a handle runs).
      const char* env = getenv("EKF_LARGE_TMA");
      h->wk.use_tma = env ? atoi(env) : 1;
      if (h->wk.use_tma) {
        const size_t mb = ekf_large_tma_map_bytes();
        h->tmaps.resize((size_t)n_filters * mb);
        bool ok = ekf_large_tma_prepare(sms, &h->wk.tma_grid) == cudaSuccess;
        for (int f = 0; ok && f < n_filters; ++f)
          ok = ekf_large_tma_encode(h->tmaps.data() + (size_t)f * mb, st.P + (size_t)f * st.slab, st.cap_n, st.ld) == cudaSuccess;
        if (!ok)
          return bail(EKF_ERR_CUDA, "the tensor maps of the TMA-staged covariance sweep could not be encoded "
                                    "(set EKF_LARGE_TMA=0 to run the plain double2 sweep instead)");
        h->wk.tmaps = h->tmaps.data();
      }
    }
    h->ltm.ev0 = h->kev0;
    h->ltm.ev1 = h->kev1;
    h->ltm.cap = kTimingPairs;
    h->ltm.every = 1;
  }
  rc = ekf_reset(h);
  if (rc != EKF_OK) {
    const std::string msg = h->err;
    return bail(rc, msg);
  }
  *out = h;
  return EKF_OK;
}

int ekf_destroy(ekf_handle h) {
  if (!h) return EKF_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->st.x); cudaFree(h->st.P); cudaFree(h->st.nlm); cudaFree(h->st.status);
  cudaFree(h->wk.W); cudaFree(h->wk.cand_val); cudaFree(h->wk.cand_idx); cudaFree(h->wk.small);
  cudaFree(h->wk.strip); cudaFree(h->wk.diag);
  if (h->wk.s_side) cudaStreamDestroy(h->wk.s_side);
  if (h->wk.ev_a) cudaEventDestroy(h->wk.ev_a);
  if (h->wk.ev_b) cudaEventDestroy(h->wk.ev_b);
  h->in.release(); h->o_dec.release(); h->o_idx.release(); h->o_mah.release(); h->in_valid.release();
  h->resume.release();
  for (int i = 0; i < 4; ++i) h->pc_rec[i].release();
  if (h->s_grow) {
    cudaStreamDestroy(h->s_grow);
    cudaEventDestroy(h->ev_grow0);
    cudaEventDestroy(h->ev_grow1);
  }
  for (int i = 0; i < 4; ++i) {
    if (h->stage[i]) cudaFreeHost(h->stage[i]);
    if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]);
    if (h->stage_kev[i]) cudaEventDestroy(h->stage_kev[i]);
  }
  h->records.release(); h->t_dec.release(); h->t_idx.release(); h->t_mah.release(); h->t_pose.release();
  if (h->ev0) {
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaEventDestroy(h->ev_done);
    for (int i = 0; i < kMaxChunks; ++i) {
      cudaEventDestroy(h->ev_in[i]);
      cudaEventDestroy(h->ev_k[i]);
    }
    cudaStreamDestroy(h->s_in);
    cudaStreamDestroy(h->s_out);
    for (int i = 0; i < kTimingPairs; ++i) {
      cudaEventDestroy(h->kev0[i]);
      cudaEventDestroy(h->kev1[i]);
    }
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return EKF_OK;
}

int ekf_reset(ekf_handle h) {
  if (!h) return EKF_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  h->prop_pending = false;                       // a held-back propagate of the old state is moot
  const EkfState& st = h->st;
  EKF_CK(h, cudaMemsetAsync(st.x, 0, (size_t)st.F * st.xs * sizeof(double), h->stream));
  EKF_CK(h, cudaMemsetAsync(st.P, 0, (size_t)st.F * st.slab * sizeof(double), h->stream));
  EKF_CK(h, cudaMemsetAsync(st.nlm, 0, (size_t)st.F * sizeof(int), h->stream));
  EKF_CK(h, cudaMemsetAsync(st.status, 0, (size_t)st.F * sizeof(int), h->stream));
  if (h->wk.W) EKF_CK(h, cudaMemsetAsync(h->wk.W, 0, ((size_t)st.cap_n + 512) * sizeof(double2), h->stream));
  return EKF_OK;
}

int ekf_resize(ekf_handle* hp, int new_max_landmarks) {
  if (!hp || !*hp) return EKF_ERR_BAD_ARG;
  ekf_handle h = *hp;
  cudaSetDevice(h->device);
  EKF_FLUSH(h);
  const EkfState& so = h->st;
  std::vector<int> nl(so.F);
  EKF_CK(h, cudaMemcpyAsync(nl.data(), so.nlm, sizeof(int) * so.F, cudaMemcpyDeviceToHost, h->stream));
  EKF_CK(h, cudaStreamSynchronize(h->stream));
  int need = 0;
  for (int f = 0; f < so.F; ++f) need = nl[f] > need ? nl[f] : need;
  if (new_max_landmarks < need || new_max_landmarks < 1)
    return fail(h, EKF_ERR_BAD_ARG, "ekf_resize: new capacity " + std::to_string(new_max_landmarks) + " is below the largest map (" + std::to_string(need) + " landmarks)");
  ekf_config cfg = h->cfg;
  cfg.regime = EKF_REGIME_AUTO;                  // the regime follows the capacity
  ekf_handle g = nullptr;
  const int rc = ekf_create(&g, h->device, so.F, new_max_landmarks, &cfg);
  if (rc != EKF_OK) return fail(h, rc, "ekf_resize: " + g_create_error);
  const EkfState& sn = g->st;
  for (int f = 0; f < so.F; ++f) {
    const size_t n = 3 + 2 * (size_t)nl[f];
    EKF_CK(h, cudaMemcpyAsync(sn.x + (size_t)f * sn.xs, so.x + (size_t)f * so.xs, n * sizeof(double), cudaMemcpyDeviceToDevice, g->stream));
    EKF_CK(h, cudaMemcpy2DAsync(sn.P + (size_t)f * sn.slab, (size_t)sn.ld * sizeof(double), so.P + (size_t)f * so.slab,
                                (size_t)so.ld * sizeof(double), n * sizeof(double), n, cudaMemcpyDeviceToDevice, g->stream));
  }
  EKF_CK(h, cudaMemcpyAsync(sn.nlm, so.nlm, sizeof(int) * so.F, cudaMemcpyDeviceToDevice, g->stream));
  EKF_CK(h, cudaMemcpyAsync(sn.status, so.status, sizeof(int) * so.F, cudaMemcpyDeviceToDevice, g->stream));
  EKF_CK(h, cudaStreamSynchronize(g->stream));
  g->launches = h->launches;
  ekf_destroy(h);
  *hp = g;
  return EKF_OK;
}

int ekf_n_filters(ekf_handle h) { return h ? h->st.F : 0; }
int ekf_max_landmarks(ekf_handle h) { return h ? h->st.cap_lm : 0; }
int ekf_regime(ekf_handle h) { return h ? h->regime : 0; }
int ekf_large_downdate_kernel(ekf_handle h) {
  if (!h || h->regime != EKF_REGIME_LARGE) return -1;
  return h->wk.use_tma ? 1 : 0;
}
int ekf_set_batch_kernel(ekf_handle h, int batch_kernel) {
  if (!h || batch_kernel < EKF_BATCH_KERNEL_AUTO || batch_kernel > EKF_BATCH_KERNEL_DTILE) return EKF_ERR_BAD_ARG;
  h->cfg.batch_kernel = batch_kernel;
  return EKF_OK;
}

int ekf_set_state(ekf_handle h, int filter, int n_landmarks, const double* x, const double* P, int ld) {
  if (!h || !x || !P) return EKF_ERR_BAD_ARG;
  const EkfState& st = h->st;
  const int n = 3 + 2 * n_landmarks;
  if (filter < 0 || filter >= st.F || n_landmarks < 0 || n_landmarks > st.cap_lm || ld < n)
    return fail(h, EKF_ERR_BAD_ARG, "ekf_set_state: bad filter / n_landmarks / ld");
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < j; ++i) {
      const double a = P[i + (size_t)j * ld], b = P[j + (size_t)i * ld];
      if (!(a == b || (a != a && b != b)))
        return fail(h, EKF_ERR_BAD_ARG, "ekf_set_state: P must be bit-symmetric (the reference symmetrises after every operation)");
    }
  cudaSetDevice(h->device);
  EKF_FLUSH(h);
  EKF_CK(h, cudaMemcpyAsync(st.x + (size_t)filter * st.xs, x, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  EKF_CK(h, cudaMemcpy2DAsync(st.P + (size_t)filter * st.slab, (size_t)st.ld * sizeof(double), P, (size_t)ld * sizeof(double),
                              (size_t)n * sizeof(double), n, cudaMemcpyHostToDevice, h->stream));
  EKF_CK(h, cudaMemcpyAsync(st.nlm + filter, &n_landmarks, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  EKF_CK(h, cudaMemsetAsync(st.status + filter, 0, sizeof(int), h->stream));   // a fresh map has dropped nothing
  if (h->wk.W) EKF_CK(h, cudaMemsetAsync(h->wk.W, 0, ((size_t)st.cap_n + 512) * sizeof(double2), h->stream));
  EKF_CK(h, cudaStreamSynchronize(h->stream));
  return EKF_OK;
}

int ekf_get_state(ekf_handle h, int filter, int* n_landmarks, double* x, double* P, int ld) {
  if (!h) return EKF_ERR_BAD_ARG;
  const EkfState& st = h->st;
  if (filter < 0 || filter >= st.F) return fail(h, EKF_ERR_BAD_ARG, "ekf_get_state: bad filter");
  cudaSetDevice(h->device);
  EKF_FLUSH(h);
  int nl = 0;
  EKF_CK(h, cudaMemcpyAsync(&nl, st.nlm + filter, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  EKF_CK(h, cudaStreamSynchronize(h->stream));
  const int n = 3 + 2 * nl;
  if (n_landmarks) *n_landmarks = nl;
  if (x) EKF_CK(h, cudaMemcpyAsync(x, st.x + (size_t)filter * st.xs, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  if (P) {
    if (ld < n) return fail(h, EKF_ERR_BAD_ARG, "ekf_get_state: ld < n");
    EKF_CK(h, cudaMemcpy2DAsync(P, (size_t)ld * sizeof(double), st.P + (size_t)filter * st.slab, (size_t)st.ld * sizeof(double),
                                (size_t)n * sizeof(double), n, cudaMemcpyDeviceToHost, h->stream));
  }
  EKF_CK(h, cudaStreamSynchronize(h->stream));
  return EKF_OK;
}

int ekf_get_cov_block(ekf_handle h, int filter, int r0, int c0, int nr, int nc, double* out, int ld_out) {
  if (!h || !out) return EKF_ERR_BAD_ARG;
  const EkfState& st = h->st;
  if (filter < 0 || filter >= st.F || r0 < 0 || c0 < 0 || nr < 1 || nc < 1 || r0 + nr > st.cap_n || c0 + nc > st.cap_n ||
      ld_out < nr)
    return fail(h, EKF_ERR_BAD_ARG, "ekf_get_cov_block: bad block");
  cudaSetDevice(h->device);
  EKF_FLUSH(h);
  EKF_CK(h, cudaMemcpy2DAsync(out, (size_t)ld_out * sizeof(double),
                              st.P + (size_t)filter * st.slab + r0 + (size_t)c0 * st.ld, (size_t)st.ld * sizeof(double),
                              (size_t)nr * sizeof(double), nc, cudaMemcpyDeviceToHost, h->stream));
  EKF_CK(h, cudaStreamSynchronize(h->stream));
  return EKF_OK;
}

int ekf_get_pose(ekf_handle h, double* xyphi, int32_t* n_landmarks) {
  if (!h) return EKF_ERR_BAD_ARG;
  const EkfState& st = h->st;
  cudaSetDevice(h->device);
  EKF_FLUSH(h);
  if (xyphi)
    EKF_CK(h, cudaMemcpy2DAsync(xyphi, 3 * sizeof(double), st.x, st.xs * sizeof(double), 3 * sizeof(double), st.F,
                                cudaMemcpyDeviceToHost, h->stream));
  if (n_landmarks) EKF_CK(h, cudaMemcpyAsync(n_landmarks, st.nlm, sizeof(int) * st.F, cudaMemcpyDeviceToHost, h->stream));
  EKF_CK(h, cudaStreamSynchronize(h->stream));
  return EKF_OK;
}

int ekf_propagate(ekf_handle h, const double* vel_mm_s, const double* rotvel_deg_s, const double* dt, int dt_stride) {
  if (!h || !vel_mm_s || !rotvel_deg_s || !dt) return EKF_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  const size_t F = h->st.F;
  if (h->regime == EKF_REGIME_BATCH) {
    // Held back: if the next call is doUpdate with one measurement (slam.cpp:136 -> :170) both ride in ONE
    // launch of the fused kernel (covariance on chip for the pair, read and written once); any other call
    // runs this propagate on its own first (flush_pending).
    EKF_FLUSH(h);
    h->pend.resize(3 * F);
    memcpy(h->pend.data(), vel_mm_s, F * sizeof(double));
    memcpy(h->pend.data() + F, rotvel_deg_s, F * sizeof(double));
    for (size_t f = 0; f < F; ++f) h->pend[2 * F + f] = dt[dt_stride ? f : 0];
    h->prop_pending = true;
    return EKF_OK;
  }
  EKF_CK(h, h->in.reserve(3 * F));
  EKF_CK(h, cudaMemcpyAsync(h->in.p, vel_mm_s, F * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EKF_CK(h, cudaMemcpyAsync(h->in.p + F, rotvel_deg_s, F * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EKF_CK(h, cudaMemcpyAsync(h->in.p + 2 * F, dt, (dt_stride ? F : 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EkfPercallIO io{};
  io.vel = h->in.p; io.rotvel = h->in.p + F; io.dt = h->in.p + 2 * F; io.dt_stride = dt_stride ? 1 : 0;
  if (h->regime == EKF_REGIME_BATCH) {
    EKF_CK(h, ekf_batch_percall(h->st, io, EKF_OP_PROPAGATE, h->k, h->stream));
    h->launches += 1;
  } else {
    for (int f = 0; f < h->st.F; ++f) {
      EKF_CK(h, ekf_large_percall(h->st, f, io, EKF_OP_PROPAGATE, h->k, h->wk, nullptr, h->stream));
      h->launches += ekf_large_launches_per(EKF_OP_PROPAGATE);
    }
  }
  // the staging buffer is reused by the next call: the copy above must have been consumed
  EKF_CK(h, cudaStreamSynchronize(h->stream));
  return EKF_OK;
}

int ekf_update(ekf_handle h, int n_z, const double* z, const double* R, int32_t* decision, int32_t* lm_index, double* mahal) {
  if (!h || n_z < 0 || (n_z > 0 && (!z || !R))) return EKF_ERR_BAD_ARG;
  if (n_z == 0) return EKF_OK;
  cudaSetDevice(h->device);
  const size_t F = h->st.F, FZ = F * n_z;
  if (h->prop_pending && n_z == 1 && h->regime == EKF_REGIME_BATCH && pick_batch_kernel(h) > 0) {
    // doPropagation + doUpdate of one loop iteration in one fused launch (T = 1 step records)
    h->prop_pending = false;
    const int L = EKF_RECORD_LEN(1);
    const int si = h->stage_i;
    h->stage_i = (si + 1) & 3;
    if (!h->stage[si]) {
      EKF_CK(h, cudaMallocHost(&h->stage[si], F * L * sizeof(double)));
      EKF_CK(h, cudaEventCreateWithFlags(&h->stage_ev[si], cudaEventDisableTiming));
      EKF_CK(h, cudaEventCreateWithFlags(&h->stage_kev[si], cudaEventDisableTiming));
      EKF_CK(h, h->pc_rec[si].reserve(F * L));
    } else {
      EKF_CK(h, cudaEventSynchronize(h->stage_ev[si]));   // the copy that last read this staging buffer is done
    }
    double* rec = h->stage[si];
    for (size_t f = 0; f < F; ++f) {
      double* r = rec + f * L;
      r[0] = h->pend[f]; r[1] = h->pend[F + f]; r[2] = h->pend[2 * F + f];
      r[3] = 0.0; r[4] = 0.0; r[5] = 1.0; r[6] = 0.0; r[7] = 0.0;
      r[8] = z[2 * f]; r[9] = z[2 * f + 1];
      for (int c = 0; c < 4; ++c) r[10 + c] = R[4 * f + c];
    }
    EKF_CK(h, cudaStreamWaitEvent(h->s_in, h->stage_kev[si], 0));   // (no-op until recorded) four calls ago
    EKF_CK(h, cudaMemcpyAsync(h->pc_rec[si].p, rec, F * L * sizeof(double), cudaMemcpyHostToDevice, h->s_in));
    EKF_CK(h, cudaEventRecord(h->stage_ev[si], h->s_in));
    EKF_CK(h, cudaStreamWaitEvent(h->stream, h->stage_ev[si], 0));
    EkfRunIO io{};
    io.records = h->pc_rec[si].p; io.T = 1; io.M = 1; io.L = L;
    const bool want = decision || lm_index || mahal;
    if (want) {
      EKF_CK(h, h->o_dec.reserve(F));
      EKF_CK(h, h->o_idx.reserve(F));
      EKF_CK(h, h->o_mah.reserve(F));
      io.decision = h->o_dec.p; io.index = h->o_idx.p; io.mahal = h->o_mah.p;
    }
    kernel_event_begin(h);
    EKF_CK(h, launch_batch_kernel(h, pick_batch_kernel(h), h->st, io));
    kernel_event_end(h);
    EKF_CK(h, cudaEventRecord(h->stage_kev[si], h->stream));
    h->launches += 1;
    if (!want) return EKF_OK;                        // fully asynchronous: no host round trip
    if (decision) EKF_CK(h, cudaMemcpyAsync(decision, h->o_dec.p, F * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (lm_index) EKF_CK(h, cudaMemcpyAsync(lm_index, h->o_idx.p, F * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (mahal) EKF_CK(h, cudaMemcpyAsync(mahal, h->o_mah.p, F * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EKF_CK(h, cudaStreamSynchronize(h->stream));
    if (decision)
      for (size_t q = 0; q < F; ++q)
        if (decision[q] == EKF_DECISION_DROPPED) return fail(h, EKF_ERR_CAPACITY, "landmark capacity exceeded, a New association was dropped");
    return EKF_OK;
  }
  EKF_FLUSH(h);
  std::vector<double> zr(FZ * 6);
  for (size_t q = 0; q < FZ; ++q) {
    zr[6 * q + 0] = z[2 * q + 0];
    zr[6 * q + 1] = z[2 * q + 1];
    for (int c = 0; c < 4; ++c) zr[6 * q + 2 + c] = R[4 * q + c];
  }
  EKF_CK(h, h->in.reserve(FZ * 6));
  EKF_CK(h, h->o_dec.reserve(FZ));
  EKF_CK(h, h->o_idx.reserve(FZ));
  EKF_CK(h, h->o_mah.reserve(FZ));
  EKF_CK(h, cudaMemcpyAsync(h->in.p, zr.data(), FZ * 6 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EkfPercallIO io{};
  io.n_z = n_z; io.zr = h->in.p;
  io.decision = h->o_dec.p; io.index = h->o_idx.p; io.mahal = h->o_mah.p;
  if (h->regime == EKF_REGIME_BATCH) {
    EKF_CK(h, ekf_batch_percall(h->st, io, EKF_OP_UPDATE, h->k, h->stream));
    h->launches += 1;
  } else {
    for (int f = 0; f < h->st.F; ++f) {
      EKF_CK(h, ekf_large_percall(h->st, f, io, EKF_OP_UPDATE, h->k, h->wk, &h->ltm, h->stream));
      h->launches += (long long)n_z * ekf_large_launches_per(EKF_OP_UPDATE);
    }
  }
  if (decision) EKF_CK(h, cudaMemcpyAsync(decision, h->o_dec.p, FZ * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (lm_index) EKF_CK(h, cudaMemcpyAsync(lm_index, h->o_idx.p, FZ * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (mahal) EKF_CK(h, cudaMemcpyAsync(mahal, h->o_mah.p, FZ * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  EKF_CK(h, cudaStreamSynchronize(h->stream));   // zr (host) and the staging buffer are reusable
  if (decision)
    for (size_t q = 0; q < FZ; ++q)
      if (decision[q] == EKF_DECISION_DROPPED) return fail(h, EKF_ERR_CAPACITY, "landmark capacity exceeded, a New association was dropped");
  return EKF_OK;
}

int ekf_update_compass(ekf_handle h, const double* z, const double* R, const uint8_t* valid) {
  if (!h || !z || !R) return EKF_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  EKF_FLUSH(h);
  const size_t F = h->st.F;
  EKF_CK(h, h->in.reserve(2 * F));
  EKF_CK(h, cudaMemcpyAsync(h->in.p, z, F * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EKF_CK(h, cudaMemcpyAsync(h->in.p + F, R, F * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  EkfPercallIO io{};
  io.cz = h->in.p; io.cR = h->in.p + F;
  if (valid) {
    EKF_CK(h, h->in_valid.reserve(F));
    EKF_CK(h, cudaMemcpyAsync(h->in_valid.p, valid, F, cudaMemcpyHostToDevice, h->stream));
    io.cvalid = h->in_valid.p;
  }
  if (h->regime == EKF_REGIME_BATCH) {
    EKF_CK(h, ekf_batch_percall(h->st, io, EKF_OP_COMPASS, h->k, h->stream));
    h->launches += 1;
  } else {
    for (int f = 0; f < h->st.F; ++f) {
      if (valid && !valid[f]) continue;
      EKF_CK(h, ekf_large_percall(h->st, f, io, EKF_OP_COMPASS, h->k, h->wk, nullptr, h->stream));
      h->launches += ekf_large_launches_per(EKF_OP_COMPASS);
    }
  }
  EKF_CK(h, cudaStreamSynchronize(h->stream));
  return EKF_OK;
}

int ekf_upload_records(ekf_handle h, int n_steps, int max_meas, const double* records) {
  if (!h || !records || n_steps < 1 || max_meas < 0 || max_meas > EKF_MAX_MEAS)
    return fail(h, EKF_ERR_BAD_ARG, "ekf_upload_records: bad arguments (max_meas <= " + std::to_string(EKF_MAX_MEAS) + ")");
  cudaSetDevice(h->device);
  const int L = EKF_RECORD_LEN(max_meas);
  const size_t count = (size_t)h->st.F * n_steps * L;
  EKF_CK(h, h->records.reserve(count));
  EKF_CK(h, cudaMemcpyAsync(h->records.p, records, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->rec_T = n_steps; h->rec_M = max_meas; h->rec_L = L;
  if (h->regime == EKF_REGIME_LARGE) {
    const size_t FT = (size_t)h->st.F * n_steps;
    h->flag_compass.resize(FT);
    h->flag_nz.resize(FT);
    for (size_t q = 0; q < FT; ++q) {
      const double* rec = records + q * L;
      h->flag_compass[q] = rec[6] != 0.0;
      int nz = (int)rec[5];
      h->flag_nz[q] = (uint8_t)(nz < 0 ? 0 : nz > max_meas ? max_meas : nz);
    }
  }
  return EKF_OK;
}

int ekf_run_resident(ekf_handle h, int want_trace, int want_pose_trace) {
  if (!h) return EKF_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  EKF_FLUSH(h);
  return launch_run(h, want_trace != 0, want_pose_trace != 0);
}

int ekf_download_outputs(ekf_handle h, const ekf_run_outputs* out) {
  if (!h) return EKF_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  return download(h, out);
}

// Chunk boundaries of the pipelined run. Sizes in waves of co-resident CTAs: 1, 1, 2, 2, 4, 4, 8, 8, ... 8, 8, 4,
// 4, 2, 2, 1, 1 - the first copy in (nothing to overlap it with) and the last copy out (nothing left to hide it
// behind) are short, the chunks between are long enough to amortise their launch and the tail of their last
// wave. begin[0..n] (n <= kMaxChunks) are filter indices, begin[n] = F; returns n.
static int pipeline_chunks(size_t F, size_t wave, size_t* begin) {
  if (wave < 1) wave = 1;
  const size_t waves = (F + wave - 1) / wave;
  size_t cap_waves = 8;
  int n_chunks = 0;
  for (;;) {
    size_t front[kMaxChunks], back[kMaxChunks];      // ramp sizes up from both ends towards the middle
    int nf = 0, nb = 0;
    size_t left = waves, w = 1;
    bool twice = false, at_front = true;
    while (left > 0 && nf + nb < kMaxChunks) {
      const size_t take = w < left ? w : left;
      if (at_front) front[nf++] = take; else back[nb++] = take;
      left -= take;
      if (!at_front) { if (twice && w < cap_waves) w *= 2; twice = !twice; }
      at_front = !at_front;
    }
    if (left == 0) {
      size_t f = 0;
      for (int c = 0; c < nf; ++c) { begin[n_chunks++] = f; f += front[c] * wave; }
      for (int c = nb - 1; c >= 0; --c) { begin[n_chunks++] = f; f += back[c] * wave; }
      break;
    }
    if (cap_waves >= waves) {                        // more waves than any ramp of kMaxChunks chunks covers: equal chunks
      const size_t per = (waves + kMaxChunks - 1) / kMaxChunks;
      for (size_t w0 = 0; w0 < waves; w0 += per) begin[n_chunks++] = w0 * wave;
      break;
    }
    cap_waves *= 2;                                  // too many chunks for the event ring: longer ones
  }
  begin[n_chunks] = F;
  return n_chunks;
}

// End-to-end run of the batch regime, pipelined over chunks of filters: the H2D copy of chunk c+1
// and the D2H copy of chunk c-1 overlap the kernel of chunk c (three streams, two copy engines).
// Filters are independent, so a chunk is just a sub-range of the batch.
static int run_pipelined(ekf_handle h, int n_steps, int max_meas, const double* records, const ekf_run_outputs* out) {
  const EkfState& st = h->st;
  const int L = EKF_RECORD_LEN(max_meas), M = max_meas > 0 ? max_meas : 1, T = n_steps;
  const size_t F = st.F, FT = F * T, FTM = FT * M;
  EKF_CK(h, h->records.reserve(FT * L));
  h->rec_T = T; h->rec_M = max_meas; h->rec_L = L;
  const bool want_trace = out && (out->decision || out->lm_index || out->mahal);
  const bool want_pose = out && out->pose_trace;
  h->have_trace = want_trace;
  h->have_pose_trace = want_pose;
  if (want_trace) {
    EKF_CK(h, h->t_dec.reserve(FTM));
    EKF_CK(h, h->t_idx.reserve(FTM));
    EKF_CK(h, h->t_mah.reserve(FTM));
  }
  if (want_pose) EKF_CK(h, h->t_pose.reserve(FT * 3));
  const int kern = pick_batch_kernel(h);
  if (kern < 0)
    return fail(h, EKF_ERR_UNSUPPORTED, "the requested fused kernel does not support max_landmarks = " + std::to_string(h->st.cap_lm) + " (TILE / STILE: <= " + std::to_string(ekf_tile_max_landmarks()) + ", DTILE: <= " + std::to_string(ekf_dtile_max_landmarks()) + ")");
  // chunk = a multiple of the co-resident CTA count (2 filters per CTA), at most kMaxChunks chunks
  const size_t wave = (size_t)(kern == 4 ? ekf_dtile_ctas_per_sm() * h->sm_count
                               : kern == 3 ? ekf_stile_ctas_per_sm(h->cfg.batch_kernel == EKF_BATCH_KERNEL_AUTO ? 1 : st.cap_lm) * h->sm_count
                                           : (kern == 2 ? 2 * h->sm_count : h->grid_cap));
  size_t begin[kMaxChunks + 1];
  const int n_chunks = pipeline_chunks(F, wave, begin);
  EKF_CK(h, cudaEventRecord(h->ev_done, h->stream));          // order after earlier work on the handle
  EKF_CK(h, cudaStreamWaitEvent(h->s_in, h->ev_done, 0));
  for (int c = 0; c < n_chunks; ++c) {
    const size_t f0 = begin[c], nf = begin[c + 1] - begin[c];
    EKF_CK(h, cudaMemcpyAsync(h->records.p + f0 * T * L, records + f0 * T * L, nf * T * L * sizeof(double),
                              cudaMemcpyHostToDevice, h->s_in));
    EKF_CK(h, cudaEventRecord(h->ev_in[c], h->s_in));
    EKF_CK(h, cudaStreamWaitEvent(h->stream, h->ev_in[c], 0));
    EkfState sub = st;
    sub.F = (int)nf;
    sub.x = st.x + f0 * st.xs;
    sub.P = st.P + f0 * st.slab;
    sub.nlm = st.nlm + f0;
    sub.status = st.status + f0;
    EkfRunIO io{};
    io.records = h->records.p + f0 * T * L;
    io.T = T; io.M = M; io.L = L;
    if (want_trace) { io.decision = h->t_dec.p + f0 * T * M; io.index = h->t_idx.p + f0 * T * M; io.mahal = h->t_mah.p + f0 * T * M; }
    if (want_pose) io.pose_trace = h->t_pose.p + f0 * T * 3;
    kernel_event_begin(h);
    EKF_CK(h, launch_batch_kernel(h, kern, sub, io));
    kernel_event_end(h);
    h->launches += 1;
    EKF_CK(h, cudaEventRecord(h->ev_k[c], h->stream));
    EKF_CK(h, cudaStreamWaitEvent(h->s_out, h->ev_k[c], 0));
    if (out) {
      cudaStream_t so = h->s_out;
      if (out->decision) EKF_CK(h, cudaMemcpyAsync(out->decision + f0 * T * M, h->t_dec.p + f0 * T * M, nf * T * M * sizeof(int), cudaMemcpyDeviceToHost, so));
      if (out->lm_index) EKF_CK(h, cudaMemcpyAsync(out->lm_index + f0 * T * M, h->t_idx.p + f0 * T * M, nf * T * M * sizeof(int), cudaMemcpyDeviceToHost, so));
      if (out->mahal) EKF_CK(h, cudaMemcpyAsync(out->mahal + f0 * T * M, h->t_mah.p + f0 * T * M, nf * T * M * sizeof(double), cudaMemcpyDeviceToHost, so));
      if (out->pose_trace) EKF_CK(h, cudaMemcpyAsync(out->pose_trace + f0 * T * 3, h->t_pose.p + f0 * T * 3, nf * T * 3 * sizeof(double), cudaMemcpyDeviceToHost, so));
      if (out->final_pose)
        EKF_CK(h, cudaMemcpy2DAsync(out->final_pose + f0 * 3, 3 * sizeof(double), st.x + f0 * st.xs, st.xs * sizeof(double),
                                    3 * sizeof(double), nf, cudaMemcpyDeviceToHost, so));
      if (out->final_nlm) EKF_CK(h, cudaMemcpyAsync(out->final_nlm + f0, st.nlm + f0, nf * sizeof(int), cudaMemcpyDeviceToHost, so));
    }
  }
  EKF_CK(h, cudaStreamSynchronize(h->s_out));
  EKF_CK(h, cudaStreamSynchronize(h->stream));
  return check_status(h);
}

int ekf_run(ekf_handle h, int n_steps, int max_meas, const double* records, const ekf_run_outputs* out) {
  if (h) {
    cudaSetDevice(h->device);
    EKF_FLUSH(h);
  }
  if (h && records && h->regime == EKF_REGIME_BATCH && n_steps >= 1 && max_meas >= 0 && max_meas <= EKF_MAX_MEAS) {
    cudaSetDevice(h->device);
    return run_pipelined(h, n_steps, max_meas, records, out);
  }
  int rc = ekf_upload_records(h, n_steps, max_meas, records);
  if (rc != EKF_OK) return rc;
  const bool want_trace = out && (out->decision || out->lm_index || out->mahal);
  const bool want_pose = out && out->pose_trace;
  rc = launch_run(h, want_trace, want_pose);
  if (rc != EKF_OK) return rc;
  return download(h, out);
}

int ekf_sync(ekf_handle h) {
  if (!h) return EKF_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  EKF_FLUSH(h);
  EKF_CK(h, cudaStreamSynchronize(h->stream));
  return check_status(h);
}

int ekf_capacity_flags(ekf_handle h, int* n_flagged, int clear) {
  if (!h || !n_flagged) return EKF_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  EKF_FLUSH(h);
  std::vector<int> s(h->st.F);
  EKF_CK(h, cudaMemcpyAsync(s.data(), h->st.status, sizeof(int) * h->st.F, cudaMemcpyDeviceToHost, h->stream));
  EKF_CK(h, cudaStreamSynchronize(h->stream));
  int n = 0;
  for (int f = 0; f < h->st.F; ++f) n += (s[f] & 1);
  *n_flagged = n;
  if (clear) EKF_CK(h, cudaMemsetAsync(h->st.status, 0, sizeof(int) * h->st.F, h->stream));
  return EKF_OK;
}

const char* ekf_last_error(ekf_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void* ekf_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
void ekf_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int ekf_timer_start(ekf_handle h) {
  if (!h) return EKF_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  EKF_CK(h, cudaEventRecord(h->ev0, h->stream));
  return EKF_OK;
}
int ekf_timer_stop(ekf_handle h, float* ms) {
  if (!h || !ms) return EKF_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  EKF_FLUSH(h);
  EKF_CK(h, cudaEventRecord(h->ev1, h->stream));
  EKF_CK(h, cudaEventSynchronize(h->ev1));
  EKF_CK(h, cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return EKF_OK;
}

long long ekf_kernel_launches(ekf_handle h) { return h ? h->launches : 0; }

int ekf_kernel_time(ekf_handle h, float* avg_ms, int* n_launches) {
  if (!h || !avg_ms || !n_launches) return EKF_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  EKF_CK(h, cudaStreamSynchronize(h->stream));
  int& used = h->regime == EKF_REGIME_BATCH ? h->kev_used : h->ltm.used;
  double sum = 0;
  for (int i = 0; i < used; ++i) {
    float ms = 0;
    EKF_CK(h, cudaEventElapsedTime(&ms, h->kev0[i], h->kev1[i]));
    sum += ms;
  }
  *n_launches = used;
  *avg_ms = used ? (float)(sum / used) : 0.f;
  used = 0;
  h->ltm.seen = 0;
  return EKF_OK;
}

int ekf_debug_phase_cycles(long long* out8) {
  return ekf_tile_phase_cycles(out8) == cudaSuccess ? EKF_OK : EKF_ERR_CUDA;
}

int ekf_debug_pipeline_chunks(long long n_filters, long long wave, long long* begin, int capacity) {
  if (n_filters < 1 || wave < 1 || !begin || capacity < kMaxChunks + 1) return -EKF_ERR_BAD_ARG;
  size_t b[kMaxChunks + 1];
  const int n = pipeline_chunks((size_t)n_filters, (size_t)wave, b);
  for (int c = 0; c <= n; ++c) begin[c] = (long long)b[c];
  return n;
}

int ekf_debug_dtile_timestamps(long long* out64) {
  if (!out64) return EKF_ERR_BAD_ARG;
  return ekf_dtile_timestamps(out64) == cudaSuccess ? EKF_OK : EKF_ERR_CUDA;
}

int ekf_debug_stile_timestamps(long long* out128) {
  if (!out128) return EKF_ERR_BAD_ARG;
  return ekf_stile_timestamps(out128) == cudaSuccess ? EKF_OK : EKF_ERR_CUDA;
}

int ekf_measure_fp64_peak(int device, double* flops_per_s) {
  if (!flops_per_s) return EKF_ERR_BAD_ARG;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return EKF_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return EKF_ERR_CUDA;
  return ekf_fp64_peak(flops_per_s, nullptr) == cudaSuccess ? EKF_OK : EKF_ERR_CUDA;
}

}  // extern "C"
