// TEST INFRASTRUCTURE ONLY (checker + CPU baseline; never on the product path).
//
// C-ABI harness around the REFERENCE's own KalmanFilter (kentsommer/2D-EKF-SLAM,
// odometry/kalmanfilter.{h,cpp}, Propagate.cpp, Update.cpp), whose translation units are
// compiled unmodified, where they lie under /root/reference, over the stand-in headers in
// oracle/shim/ (see oracle/Makefile). Outputs go to oracle/_ref/ only.
//
// What this file adds, and nothing else:
//   * a slam.cpp:130-182 shaped driver (propagate -> optional compass -> one doUpdate per
//     measurement) reading the step-record format shared with the CUDA library
//     (include/ekf_slam_b200.h, "step record");
//   * observation of the data-association decision without editing the reference: the
//     "New "/"Old "/"Ignore " tokens Update.cpp:154,183,191 print to std::cout are captured by a
//     thread-local streambuf, and the per-landmark cond / Mahalanobis values of the gating loop
//     (Update.cpp:127-147) are captured by the shim's trace hooks, from which Opt_i and
//     Mahal_dist are re-derived with the same strict-'>' rule (Update.cpp:140) and cross-checked
//     against the column index the reference then reads at Update.cpp:186;
//   * a std::thread batch runner used as the CPU baseline (one reference filter per task).
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <streambuf>
#include <string>
#include <thread>
#include <vector>

#define private public  // reach state/covariance/Propagate/Update; TUs themselves are untouched
#include "kalmanfilter.h"
#undef private

namespace {

// ---- std::cout capture: tokens go to a per-thread string (or nowhere) -----------------------
thread_local std::string* tls_tokens = nullptr;

class TokenBuf : public std::streambuf {
 protected:
  std::streamsize xsputn(const char* s, std::streamsize n) override {
    if (tls_tokens) tls_tokens->append(s, static_cast<size_t>(n));
    return n;
  }
  int overflow(int c) override {
    if (tls_tokens && c != EOF) tls_tokens->push_back(static_cast<char>(c));
    return c == EOF ? 0 : c;
  }
};
TokenBuf g_token_buf;
std::streambuf* g_saved_cout = nullptr;
std::atomic<int> g_installed{0};

void install_cout_capture() {
  int expected = 0;
  if (g_installed.compare_exchange_strong(expected, 1)) g_saved_cout = std::cout.rdbuf(&g_token_buf);
}

// ---- shim trace sink: what the gating loop computed ------------------------------------------
struct GateSink : Eigen::ShimTraceSink {
  std::vector<double> cond;      // one per landmark visited (Update.cpp:128)
  std::vector<double> scalars;   // one per landmark that passed the cond gate (Update.cpp:136)
  long k_col = -1;               // start column of P_min.block(0,Opt_i,stateSize,2) (Update.cpp:186)
  long state_size = 0;
  void reset(long n) { cond.clear(); scalars.clear(); k_col = -1; state_size = n; }
  void on_scalar(double v) override { scalars.push_back(v); }
  void on_svd(double s0, double s1) override { cond.push_back(s0 / s1); }
  void on_block(long rows, long cols, long r0, long c0, long nr, long nc) override {
    // Update.cpp:186 is the only read with r0==0, nr==stateSize, nc==2, c0>=3 on an n x n matrix
    // (Update.cpp:175 writes into an (n+2) x (n+2) matrix; Update.cpp:116 has nr==3).
    if (rows == state_size && cols == state_size && r0 == 0 && nr == state_size && nc == 2 && c0 >= 3 &&
        state_size > 3)
      k_col = c0;
  }
};

}  // namespace

extern "C" {

struct RefFilter {
  ArRobot robot;
  KalmanFilter* ekf;
  std::ofstream cov_file, known_file;  // never opened: kalmanfilter.cpp:51-61 writes fail fast
  GateSink sink;
  std::string tokens;
};

// decision codes shared with include/ekf_slam_b200.h
enum { REF_NEW = 0, REF_OLD = 1, REF_IGNORE = 2 };

struct RefTrace {
  int32_t decision;   // REF_NEW / REF_OLD / REF_IGNORE, from the stdout token
  int32_t opt_i;      // state index of the associated landmark (3,5,7,..), 0 = none
  double mahal;       // Mahal_dist after the gating loop (INF literal if none)
  int32_t n_cond_skipped;
  int32_t k_col;      // column the reference read at Update.cpp:186 (-1 unless Old)
  double margin_gmin; // |mahal - 10|, |mahal - 50|, min |cond - 80| : knife-edge evidence
  double margin_gmax;
  double margin_cond;
};

RefFilter* ref_create(void) {
  install_cout_capture();
  RefFilter* f = new RefFilter();
  f->ekf = new KalmanFilter(&f->robot);
  return f;
}

void ref_destroy(RefFilter* f) {
  if (!f) return;
  delete f->ekf->state;       // the reference has no destructor (kalmanfilter.h:21-43)
  delete f->ekf->covariance;
  delete f->ekf;
  delete f;
}

int ref_dim(RefFilter* f) { return static_cast<int>(f->ekf->state->size()); }

// x: n doubles; P: n*n doubles, column-major, leading dimension n.
void ref_get_state(RefFilter* f, double* x, double* P) {
  const long n = f->ekf->state->size();
  for (long i = 0; i < n; ++i) x[i] = (*f->ekf->state)(i);
  for (long j = 0; j < n; ++j)
    for (long i = 0; i < n; ++i) P[i + j * n] = (*f->ekf->covariance)(i, j);
}

void ref_set_state(RefFilter* f, int n, const double* x, const double* P) {
  delete f->ekf->state;
  delete f->ekf->covariance;
  f->ekf->state = new Eigen::VectorXd(n);
  f->ekf->covariance = new Eigen::MatrixXd(n, n);
  for (long i = 0; i < n; ++i) (*f->ekf->state)(i) = x[i];
  for (long j = 0; j < n; ++j)
    for (long i = 0; i < n; ++i) (*f->ekf->covariance)(i, j) = P[i + j * n];
  f->ekf->Num_Landmarks = (n - 3) / 2;
  f->ekf->X = x[0];
  f->ekf->Y = x[1];
  f->ekf->Phi = x[2];
}

void ref_get_pose(RefFilter* f, double* xyphi, int* num_landmarks) {
  xyphi[0] = f->ekf->X;
  xyphi[1] = f->ekf->Y;
  xyphi[2] = f->ekf->Phi;
  if (num_landmarks) *num_landmarks = f->ekf->Num_Landmarks;
}

// kalmanfilter.cpp:15-62 through the public surface; the robot stub supplies mm/s and deg/s.
void ref_propagate(RefFilter* f, double vel_mm_s, double rotvel_deg_s, double dt) {
  f->robot.vel_mm_s = vel_mm_s;
  f->robot.rotvel_deg_s = rotvel_deg_s;
  f->ekf->doPropagation(dt, f->cov_file, f->known_file);
}

void ref_update_compass(RefFilter* f, double z, double R) { f->ekf->doUpdateCompass(z, R); }

// One doUpdate call with a single 2x1 measurement (slam.cpp:152-170). R is column-major 2x2.
// trace may be NULL (then nothing is observed and std::cout output is discarded).
int ref_update(RefFilter* f, const double* z, const double* R, RefTrace* trace) {
  Eigen::MatrixXd z_chunk(2, 1), R_chunk(2, 2);
  z_chunk(0, 0) = z[0];
  z_chunk(1, 0) = z[1];
  R_chunk(0, 0) = R[0];
  R_chunk(1, 0) = R[1];
  R_chunk(0, 1) = R[2];
  R_chunk(1, 1) = R[3];
  if (!trace) {
    f->ekf->doUpdate(z_chunk, R_chunk);
    return 0;
  }
  const long n = f->ekf->state->size();
  f->sink.reset(n);
  f->tokens.clear();
  Eigen::shim_trace_sink() = &f->sink;
  tls_tokens = &f->tokens;
  f->ekf->doUpdate(z_chunk, R_chunk);
  tls_tokens = nullptr;
  Eigen::shim_trace_sink() = nullptr;

  int rc = 0;
  if (f->tokens == "New ") trace->decision = REF_NEW;
  else if (f->tokens == "Old ") trace->decision = REF_OLD;
  else if (f->tokens == "Ignore ") trace->decision = REF_IGNORE;
  else { trace->decision = -1; rc = 1; }

  // Re-derive Opt_i / Mahal_dist from the values the reference itself computed.
  const long n_lm = (n - 3) / 2;
  double mahal = INF;
  int opt_i = 0, skipped = 0;
  size_t s = 0;
  double m_cond = 1e300;
  if (static_cast<long>(f->sink.cond.size()) != n_lm) rc |= 2;
  for (long i = 1; i <= n_lm && i <= static_cast<long>(f->sink.cond.size()); ++i) {
    const double cond = f->sink.cond[i - 1];
    const double dc = std::fabs(cond - 80.0);
    if (dc < m_cond) m_cond = dc;
    if (cond >= 80) { ++skipped; continue; }           // Update.cpp:131
    if (s >= f->sink.scalars.size()) { rc |= 4; break; }
    const double temp = f->sink.scalars[s++];
    if (mahal > temp) { mahal = temp; opt_i = static_cast<int>(2 * i + 1); }  // Update.cpp:140-143
  }
  if (s != f->sink.scalars.size()) rc |= 8;
  trace->opt_i = opt_i;
  trace->mahal = mahal;
  trace->n_cond_skipped = skipped;
  trace->k_col = static_cast<int32_t>(f->sink.k_col);
  trace->margin_gmin = std::fabs(mahal - 10.0);
  trace->margin_gmax = std::fabs(mahal - 50.0);
  trace->margin_cond = m_cond;
  // cross-checks: the token must agree with the re-derived decision, and the column the
  // reference actually used for K must be the re-derived Opt_i.
  int expect = (opt_i == 0 || mahal > 50) ? REF_NEW : (mahal < 10 ? REF_OLD : REF_IGNORE);
  if (expect != trace->decision) rc |= 16;
  if (trace->decision == REF_OLD && f->sink.k_col != opt_i) rc |= 32;
  return rc;
}

// Pure-function access to the private Propagate / Update: state in, Set out.
// x_io/P_io hold capacity for (n+2*n_z) entries; returns the new dimension.
void ref_call_propagate(int n, double* x_io, double* P_io, double v_m, double w_m, const double* Q, double dt) {
  install_cout_capture();
  ArRobot robot;
  KalmanFilter kf(&robot);
  Eigen::VectorXd x(n);
  Eigen::MatrixXd P(n, n), Qm(2, 2);
  for (int i = 0; i < n; ++i) x(i) = x_io[i];
  for (int k = 0; k < n * n; ++k) P(k) = P_io[k];
  for (int k = 0; k < 4; ++k) Qm(k) = Q[k];
  Eigen::MatrixXd Set = kf.Propagate(x, P, v_m, w_m, Qm, dt);
  for (int i = 0; i < n; ++i) x_io[i] = Set(i, 0);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) P_io[i + j * n] = Set(i, j + 1);
  delete kf.state;
  delete kf.covariance;
}

int ref_call_update(int n, double* x_io, double* P_io, int n_z, const double* z_chunk, const double* R_chunk,
                    int gamma_max, int gamma_min) {
  install_cout_capture();
  ArRobot robot;
  KalmanFilter kf(&robot);
  Eigen::VectorXd x(n);
  Eigen::MatrixXd P(n, n), zc(2, n_z), Rc(2, 2 * n_z);
  for (int i = 0; i < n; ++i) x(i) = x_io[i];
  for (int k = 0; k < n * n; ++k) P(k) = P_io[k];
  for (int k = 0; k < 2 * n_z; ++k) zc(k) = z_chunk[k];
  for (int k = 0; k < 4 * n_z; ++k) Rc(k) = R_chunk[k];
  Eigen::MatrixXd Set = kf.Update(x, P, zc, Rc, gamma_max, gamma_min);
  const int m = static_cast<int>(Set.rows());
  for (int i = 0; i < m; ++i) x_io[i] = Set(i, 0);
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < m; ++i) P_io[i + j * m] = Set(i, j + 1);
  delete kf.state;
  delete kf.covariance;
  return m;
}

// slam.cpp:158-167, restated with the same expressions over the same matrix type:
// feature in mm (robot frame) -> z (m) and R_chunk = G*diag(0.0025,0.0001)*G^T. R_out column-major.
void ref_measurement_from_feature(double fx_mm, double fy_mm, double* z_out, double* R_out) {
  Eigen::MatrixXd R(2, 2), R_chunk(2, 2), G(2, 2);
  double fx = fx_mm / 1000.0;
  double fy = fy_mm / 1000.0;
  double dist = sqrt(fx * fx + fy * fy);
  double bearing = atan2(fy, fx);
  R << 0.0025, 0, 0, 0.0001;
  G << cos(bearing), -dist * sin(bearing), sin(bearing), dist * cos(bearing);
  R_chunk = G * R * G.transpose();
  z_out[0] = fx;
  z_out[1] = fy;
  for (int k = 0; k < 4; ++k) R_out[k] = R_chunk(k);
}

// The reference's log files for one filter driven by n_steps step records (layout below), written
// with the reference's own statements where they live in the compiled TUs (covRun / knownfeaturesRun:
// kalmanfilter.cpp:51-61, with OPEN streams this time) and with slam.cpp's expressions restated here
// for the two files main() writes (featuresRun: slam.cpp:172-177, odomRun: slam.cpp:181). scanRun
// needs laser readings and is not produced. Returns 0, or -1 if a file cannot be opened.
// scan_x / scan_y / scan_r (optional): [n_steps][n_beams] laser readings (getLocalX/Y in mm, getRange) of
// every step, for the scanRun.txt dump of slam.cpp:184-203.
int ref_run_logged_scans(int n_steps, int max_meas, const double* inputs, const char* dir, const double* scan_x,
                         const double* scan_y, const uint32_t* scan_r, int n_beams);
int ref_run_logged(int n_steps, int max_meas, const double* inputs, const char* dir) {
  return ref_run_logged_scans(n_steps, max_meas, inputs, dir, nullptr, nullptr, nullptr, 0);
}
int ref_run_logged_scans(int n_steps, int max_meas, const double* inputs, const char* dir, const double* scan_x,
                         const double* scan_y, const uint32_t* scan_r, int n_beams) {
  const int L = 8 + 6 * max_meas;
  const std::string d(dir);
  std::ofstream scanFile;
  if (scan_r) scanFile.open((d + "/scanRun.txt").c_str());
  double loopTime = 0.0;
  std::ofstream odomFile((d + "/odomRun.txt").c_str()), featuresFile((d + "/featuresRun.txt").c_str());
  std::ofstream covFile((d + "/covRun.txt").c_str()), knownfeaturesFile((d + "/knownfeaturesRun.txt").c_str());
  if (!odomFile.is_open() || !featuresFile.is_open() || !covFile.is_open() || !knownfeaturesFile.is_open()) return -1;
  install_cout_capture();    // "New / Old / Ignore" go nowhere (tls_tokens stays null)
  ArRobot robot;
  KalmanFilter* ekf = new KalmanFilter(&robot);
  for (int t = 0; t < n_steps; ++t) {
    const double* rec = inputs + static_cast<size_t>(t) * L;
    robot.vel_mm_s = rec[0];
    robot.rotvel_deg_s = rec[1];
    ekf->doPropagation(rec[2], covFile, knownfeaturesFile);
    if (rec[6] != 0.0) ekf->doUpdateCompass(rec[3], rec[4]);
    const int nz = static_cast<int>(rec[5]);
    for (int m = 0; m < nz && m < max_meas; ++m) {
      const double* zr = rec + 8 + 6 * m;
      Eigen::MatrixXd z_chunk(2, 1), R_chunk(2, 2);
      z_chunk(0, 0) = zr[0];
      z_chunk(1, 0) = zr[1];
      R_chunk(0, 0) = zr[2];
      R_chunk(1, 0) = zr[3];
      R_chunk(0, 1) = zr[4];
      R_chunk(1, 1) = zr[5];
      ekf->doUpdate(z_chunk, R_chunk);
      const double fx = zr[0], fy = zr[1];
      const double newX = fx * cos(ekf->Phi) - fy * sin(ekf->Phi);
      const double newY = fx * sin(ekf->Phi) + fy * cos(ekf->Phi);
      featuresFile << newX + ekf->X << " " << newY + ekf->Y << std::endl;
    }
    odomFile << ekf->X << " " << ekf->Y << std::endl;
    // slam.cpp:184-203: at most one laser scan per second, in the world frame of the filter's pose
    loopTime += rec[2];
    if (scan_r && loopTime > 1.0) {
      for (int i = 0; i < n_beams; i++) {
        const size_t q = static_cast<size_t>(t) * n_beams + i;
        if (scan_r[q] > 7000) continue;
        double fx = scan_x[q] / 1000.0;
        double fy = scan_y[q] / 1000.0;
        double newX = fx * cos(ekf->Phi) - fy * sin(ekf->Phi);
        double newY = fx * sin(ekf->Phi) + fy * cos(ekf->Phi);
        scanFile << newX + ekf->X << " " << newY + ekf->Y << std::endl;
      }
      loopTime = 0.0;
    }
  }
  delete ekf->state;         // the reference has no destructor (kalmanfilter.h:21-43)
  delete ekf->covariance;
  delete ekf;
  return 0;
}

// ---- batch runner (CPU baseline and bulk golden generation) ---------------------------------
// Step record layout (doubles), identical to include/ekf_slam_b200.h:
//   [0] vel_mm_s [1] rotvel_deg_s [2] dt [3] compass_z [4] compass_R [5] n_z [6] has_compass [7] 0
//   then max_meas x { z0, z1, R00, R10, R01, R11 }.
// inputs: [n_filters][n_steps][8 + 6*max_meas]. Optional outputs (NULL to skip):
//   decision/index: int32 [F][T][M] (-1 where no measurement), mahal: double [F][T][M],
//   pose_trace: double [F][T][3], final_pose: double [F][3], final_nlm: int32 [F].
// Returns the seconds spent on the TIMED steps (t >= warm_steps; warm_steps = 0 times everything):
// each worker thread accumulates the time of its own timed steps and the slowest worker is
// reported, so  (n_filters * (n_steps - warm_steps)) / return value  is the multi-threaded
// throughput with maps already built during the untimed warm-up steps. Negative = a harness
// cross-check failed.
double ref_run_batch(int n_filters, int n_steps, int max_meas, const double* inputs, int n_threads,
                     int32_t* decision, int32_t* index, double* mahal, double* pose_trace, double* final_pose,
                     int32_t* final_nlm, double* final_x, double* final_P, int final_ld, int warm_steps) {
  install_cout_capture();
  const long L = 8 + 6L * max_meas;
  const bool want_trace = decision || index || mahal;
  if (n_threads < 1) n_threads = 1;
  std::atomic<int> next{0};
  std::atomic<int> bad{0};
  std::vector<double> worker_secs(static_cast<size_t>(n_threads), 0.0);
  auto worker = [&](int wid) {
    for (;;) {
      const int f = next.fetch_add(1);
      if (f >= n_filters) break;
      RefFilter* rf = ref_create();
      auto tstart = std::chrono::steady_clock::now();
      for (int t = 0; t < n_steps; ++t) {
        if (t == warm_steps) tstart = std::chrono::steady_clock::now();
        const double* rec = inputs + (static_cast<long>(f) * n_steps + t) * L;
        ref_propagate(rf, rec[0], rec[1], rec[2]);
        if (rec[6] != 0.0) ref_update_compass(rf, rec[3], rec[4]);
        const int nz = static_cast<int>(rec[5]);
        for (int m = 0; m < max_meas; ++m) {
          const long o = (static_cast<long>(f) * n_steps + t) * max_meas + m;
          if (m < nz) {
            const double* zr = rec + 8 + 6 * m;
            if (want_trace) {
              RefTrace tr;
              if (ref_update(rf, zr, zr + 2, &tr)) bad.fetch_add(1);
              if (decision) decision[o] = tr.decision;
              if (index) index[o] = tr.decision == REF_NEW ? ref_dim(rf) - 2 : tr.opt_i;
              if (mahal) mahal[o] = tr.mahal;
            } else {
              ref_update(rf, zr, zr + 2, nullptr);
            }
          } else {
            if (decision) decision[o] = -1;
            if (index) index[o] = -1;
            if (mahal) mahal[o] = 0.0;
          }
        }
        if (pose_trace) ref_get_pose(rf, pose_trace + (static_cast<long>(f) * n_steps + t) * 3, nullptr);
      }
      if (n_steps > warm_steps)
        worker_secs[static_cast<size_t>(wid)] +=
            std::chrono::duration<double>(std::chrono::steady_clock::now() - tstart).count();
      if (final_pose) ref_get_pose(rf, final_pose + 3L * f, nullptr);
      if (final_nlm) final_nlm[f] = rf->ekf->Num_Landmarks;
      if (final_x && final_P) {
        const int n = ref_dim(rf);
        if (n <= final_ld) {
          double* x = final_x + static_cast<long>(f) * final_ld;
          double* P = final_P + static_cast<long>(f) * final_ld * final_ld;
          for (int i = 0; i < n; ++i) x[i] = (*rf->ekf->state)(i);
          for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) P[i + static_cast<long>(j) * final_ld] = (*rf->ekf->covariance)(i, j);
        } else {
          bad.fetch_add(1);
        }
      }
      ref_destroy(rf);
    }
  };
  std::vector<std::thread> pool;
  for (int i = 1; i < n_threads; ++i) pool.emplace_back(worker, i);
  worker(0);
  for (auto& th : pool) th.join();
  double secs = 0.0;
  for (double w : worker_secs) secs = w > secs ? w : secs;
  return bad.load() ? -secs : secs;
}

int ref_hardware_threads(void) { return static_cast<int>(std::thread::hardware_concurrency()); }

}  // extern "C"
