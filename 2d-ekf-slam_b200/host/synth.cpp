// Synthetic odometry + landmark-measurement driver (host only). See include/ekf_synth.h.
// Stands in for the reference's robot / laser front-end (slam.cpp:54-118,141-167): it produces
// what ArRobot::getVel()/getRotVel() and FeatureDetector::getFeatures() would have handed to the
// filter, as step records both the CPU reference and the CUDA core consume bit-for-bit.
#include "ekf_synth.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

const double kPi = 3.14159265358979323846;
const int kHeader = 8;  // EKF_RECORD_HEADER in ekf_slam_b200.h

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// Counter-based standard normals: two per (seed, filter, step, stream) key.
struct Normal2 { double a, b; };
inline Normal2 normal_pair(uint64_t seed, uint64_t f, uint64_t t, uint64_t stream) {
  uint64_t k = splitmix64(seed ^ splitmix64(f * 0x100000001B3ull + 0x51ull));
  k = splitmix64(k ^ splitmix64(t + 0x7F4A7C15ull));
  k = splitmix64(k ^ (stream * 0xD6E8FEB86659FD93ull));
  const uint64_t r1 = splitmix64(k), r2 = splitmix64(k ^ 0xA5A5A5A5A5A5A5A5ull);
  const double u1 = (static_cast<double>(r1 >> 11) + 1.0) * (1.0 / 9007199254740993.0);  // (0,1)
  const double u2 = static_cast<double>(r2 >> 11) * (1.0 / 9007199254740992.0);          // [0,1)
  const double m = std::sqrt(-2.0 * std::log(u1));
  return {m * std::cos(2.0 * kPi * u2), m * std::sin(2.0 * kPi * u2)};
}

struct World {
  std::vector<double> lx, ly;
  double cx, cy, rho, v, w;
};

World make_world(const ekf_synth_config& c) {
  World wd;
  const int T = c.steps_per_lap;
  wd.rho = c.radius;
  wd.cx = wd.rho * std::sin(kPi / T);
  wd.cy = wd.rho * std::cos(kPi / T);
  wd.v = 2.0 * wd.rho * std::sin(kPi / T) / c.dt;  // polygon side / dt
  wd.w = (2.0 * kPi / T) / c.dt;
  wd.lx.resize(c.n_landmarks);
  wd.ly.resize(c.n_landmarks);
  for (int k = 0; k < c.n_landmarks; ++k) {
    const double g = 2.0 * kPi * (k + 0.5) / c.n_landmarks;
    const double r = wd.rho + ((k & 1) ? c.ring_offset : -c.ring_offset);
    wd.lx[k] = wd.cx + r * std::sin(g);
    wd.ly[k] = wd.cy - r * std::cos(g);
  }
  return wd;
}

// Pose after t propagation steps: vertex t of the polygon, heading along the next side.
inline void true_pose(const ekf_synth_config& c, const World& wd, long t, double* p) {
  const int T = c.steps_per_lap;
  const long tl = ((t % T) + T) % T;
  const double beta = 2.0 * kPi * static_cast<double>(tl) / T - kPi / T;
  p[0] = wd.cx + wd.rho * std::sin(beta);
  p[1] = wd.cy - wd.rho * std::cos(beta);
  p[2] = 2.0 * kPi * static_cast<double>(t) / T;  // heading accumulates like the filter's Phi
}

bool config_ok(const ekf_synth_config* c) {
  return c && c->n_landmarks >= 0 && c->steps_per_lap >= 3 && c->max_meas >= 0 && c->dt > 0 && c->radius > 0;
}

void gen_filters(const ekf_synth_config& c, const World& wd, long f0, int fa, int fb, long t0, int nt, double* out,
                 int32_t* lm_ids) {
  const int M = c.max_meas;
  const int L = kHeader + 6 * M;
  std::vector<int> vis;
  std::vector<double> vr, vb;
  for (int fi = fa; fi < fb; ++fi) {
    const uint64_t f = static_cast<uint64_t>(f0 + fi);
    for (int ti = 0; ti < nt; ++ti) {
      const long t = t0 + ti;
      double* rec = out + (static_cast<long>(fi) * nt + ti) * L;
      std::memset(rec, 0, sizeof(double) * L);
      // odometry the robot would report for the move t -> t+1
      const Normal2 on = normal_pair(c.seed, f, static_cast<uint64_t>(t), 0);
      const double v_m = wd.v + c.sigma_v * wd.v * on.a;
      const double w_m = wd.w + c.sigma_w * wd.v * on.b;
      rec[0] = v_m * 1000.0;                 // getVel(): mm/s (kalmanfilter.cpp:18,26)
      rec[1] = w_m * 180.0 / 3.141592654;    // getRotVel(): deg/s (kalmanfilter.cpp:19)
      rec[2] = c.dt;
      double pose[3];
      true_pose(c, wd, t + 1, pose);
      if (c.compass_every > 0 && (t % c.compass_every) == 0) {
        const Normal2 cn = normal_pair(c.seed, f, static_cast<uint64_t>(t), 1);
        double zc = std::fmod(pose[2] + c.sigma_compass * cn.a, 2.0 * kPi);
        if (zc < 0) zc += 2.0 * kPi;
        rec[3] = zc;
        rec[4] = c.compass_R;
        rec[6] = 1.0;
      }
      // visible landmarks at the post-move pose
      vis.clear(); vr.clear(); vb.clear();
      const double cp = std::cos(pose[2]), sp = std::sin(pose[2]);
      for (int k = 0; k < c.n_landmarks; ++k) {
        const double dx = wd.lx[k] - pose[0], dy = wd.ly[k] - pose[1];
        const double rx = cp * dx + sp * dy, ry = -sp * dx + cp * dy;
        const double d = std::sqrt(rx * rx + ry * ry), b = std::atan2(ry, rx);
        if (d >= c.min_range + 4.0 * c.sigma_range && d <= c.max_range - 4.0 * c.sigma_range &&
            std::fabs(b) <= 0.5 * c.fov - 4.0 * c.sigma_bearing) {
          vis.push_back(k); vr.push_back(d); vb.push_back(b);
        }
      }
      const int nv = static_cast<int>(vis.size());
      const int nz = std::min(nv, M);
      rec[5] = static_cast<double>(nz);
      for (int m = 0; m < M; ++m) {
        int32_t id = -1;
        if (m < nz) {
          const int pick = static_cast<int>((static_cast<long>(t) * M + m) % nv);
          id = vis[pick];
          const Normal2 mn = normal_pair(c.seed, f, static_cast<uint64_t>(t), 2 + static_cast<uint64_t>(m));
          const double d_m = vr[pick] + c.sigma_range * mn.a;
          const double b_m = vb[pick] + c.sigma_bearing * mn.b;
          // the corner feature the detector would emit: mm, robot frame (featuredetector.h:16-19)
          const double fx_mm = d_m * std::cos(b_m) * 1000.0, fy_mm = d_m * std::sin(b_m) * 1000.0;
          ekf_synth_measurement_from_feature(fx_mm, fy_mm, rec + kHeader + 6 * m, rec + kHeader + 6 * m + 2);
        }
        if (lm_ids) lm_ids[(static_cast<long>(fi) * nt + ti) * M + m] = id;
      }
    }
  }
}

}  // namespace

extern "C" {

void ekf_synth_default_config(ekf_synth_config* cfg, int n_landmarks) {
  std::memset(cfg, 0, sizeof *cfg);
  cfg->n_landmarks = n_landmarks;
  cfg->steps_per_lap = 1000;
  cfg->max_meas = 1;
  cfg->compass_every = 0;
  cfg->dt = 0.2;
  cfg->radius = std::max(5.0, 0.2 * n_landmarks);
  cfg->ring_offset = 3.0;
  cfg->sigma_v = 0.01;
  cfg->sigma_w = 0.04;
  cfg->sigma_range = 0.05;
  cfg->sigma_bearing = 0.01;
  cfg->min_range = 1.0;
  cfg->max_range = 8.0;
  cfg->fov = kPi;
  cfg->sigma_compass = std::sqrt(0.0005);
  cfg->compass_R = 0.0005;
  cfg->seed = 0x2D5EEDull;
}

int ekf_synth_record_len(const ekf_synth_config* cfg) { return kHeader + 6 * cfg->max_meas; }

void ekf_synth_world(const ekf_synth_config* cfg, double* lm_xy) {
  if (!config_ok(cfg)) return;
  const World wd = make_world(*cfg);
  for (int k = 0; k < cfg->n_landmarks; ++k) {
    lm_xy[2 * k] = wd.lx[k];
    lm_xy[2 * k + 1] = wd.ly[k];
  }
}

void ekf_synth_true_pose(const ekf_synth_config* cfg, long t, double* xyphi) {
  if (!config_ok(cfg)) return;
  const World wd = make_world(*cfg);
  true_pose(*cfg, wd, t, xyphi);
}

int ekf_synth_scan(const ekf_synth_config* cfg, long t, double* local_x_mm, double* local_y_mm, uint32_t* range_mm) {
  if (!config_ok(cfg) || !local_x_mm || !local_y_mm || !range_mm) return 0;
  const World wd = make_world(*cfg);
  double pose[3];
  true_pose(*cfg, wd, t, pose);
  const double half = 0.15, max_m = 8.191;
  for (int b = 0; b < EKF_SYNTH_SCAN_BEAMS; ++b) {
    const double ang = (b - 90) * kPi / 180.0;
    const double dx = std::cos(pose[2] + ang), dy = std::sin(pose[2] + ang);
    double best = max_m;
    for (int k = 0; k < cfg->n_landmarks; ++k) {
      // slab test of the ray against the axis-aligned square around landmark k
      const double ox = wd.lx[k] - pose[0], oy = wd.ly[k] - pose[1];
      if (ox * ox + oy * oy > (max_m + 1.0) * (max_m + 1.0)) continue;
      double t0 = 0.0, t1 = best;
      bool hit = true;
      const double o[2] = {ox, oy}, d[2] = {dx, dy};
      for (int a = 0; a < 2 && hit; ++a) {
        if (std::fabs(d[a]) < 1e-12) {
          hit = std::fabs(o[a]) <= half;
        } else {
          double ta = (o[a] - half) / d[a], tb = (o[a] + half) / d[a];
          if (ta > tb) std::swap(ta, tb);
          t0 = std::max(t0, ta);
          t1 = std::min(t1, tb);
          hit = t0 <= t1;
        }
      }
      if (hit && t0 > 1e-9 && t0 < best) best = t0;
    }
    const double r_mm = std::floor(best * 1000.0 + 0.5);
    const uint32_t r = best >= max_m ? 8191u : static_cast<uint32_t>(r_mm);
    range_mm[b] = r;
    local_x_mm[b] = r * std::cos(ang);
    local_y_mm[b] = r * std::sin(ang);
  }
  return EKF_SYNTH_SCAN_BEAMS;
}

// slam.cpp:158-167. Evaluated in the reference's order: R_chunk = (G*R)*G^T with sequential
// two-term inner sums seeded by the first product.
void ekf_synth_measurement_from_feature(double fx_mm, double fy_mm, double* z, double* R) {
  const double fx = fx_mm / 1000.0, fy = fy_mm / 1000.0;
  const double dist = std::sqrt(fx * fx + fy * fy);
  const double bearing = std::atan2(fy, fx);
  const double g00 = std::cos(bearing), g01 = -dist * std::sin(bearing);
  const double g10 = std::sin(bearing), g11 = dist * std::cos(bearing);
  const double r00 = 0.0025, r01 = 0.0, r10 = 0.0, r11 = 0.0001;
  const double t00 = g00 * r00 + g01 * r10, t01 = g00 * r01 + g01 * r11;
  const double t10 = g10 * r00 + g11 * r10, t11 = g10 * r01 + g11 * r11;
  z[0] = fx;
  z[1] = fy;
  R[0] = t00 * g00 + t01 * g01;  // (0,0)
  R[1] = t10 * g00 + t11 * g01;  // (1,0)
  R[2] = t00 * g10 + t01 * g11;  // (0,1)
  R[3] = t10 * g10 + t11 * g11;  // (1,1)
}

int ekf_synth_generate(const ekf_synth_config* cfg, long f0, int nf, long t0, int nt, double* out, int32_t* lm_ids,
                       int n_threads) {
  if (!config_ok(cfg) || nf < 0 || nt < 0 || !out) return 1;
  const World wd = make_world(*cfg);
  if (n_threads <= 0) n_threads = static_cast<int>(std::thread::hardware_concurrency());
  n_threads = std::max(1, std::min(n_threads, nf));
  if (n_threads == 1) {
    gen_filters(*cfg, wd, f0, 0, nf, t0, nt, out, lm_ids);
    return 0;
  }
  std::vector<std::thread> pool;
  for (int i = 0; i < n_threads; ++i) {
    const int a = static_cast<int>(static_cast<long>(nf) * i / n_threads);
    const int b = static_cast<int>(static_cast<long>(nf) * (i + 1) / n_threads);
    pool.emplace_back([&, a, b]() { gen_filters(*cfg, wd, f0, a, b, t0, nt, out, lm_ids); });
  }
  for (auto& th : pool) th.join();
  return 0;
}

}  // extern "C"
