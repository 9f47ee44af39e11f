/* ekf_synth.h — synthetic odometry + landmark-measurement driver (libekf_synth.so, host only).
 *
 * Replaces the reference's robot / serial-port / LMS-200 I/O (slam.cpp:54-118, features/,
 * movement/) with a deterministic generator, so the SAME input bits feed the reference CPU
 * filter and the CUDA filter core. It emits step records in the layout documented in
 * ekf_slam_b200.h ("step records"): what ArRobot::getVel()/getRotVel() would have returned
 * (kalmanfilter.cpp:17-20, mm/s and deg/s), the loop dt (slam.cpp:132-135), an optional
 * structural-compass reading (slam.cpp:144-147) and up to max_meas corner features converted to
 * (z, R) exactly as slam.cpp:158-167 does.
 *
 * World (SURVEY.md 8d recipe): the robot drives a closed regular polygon of `steps_per_lap`
 * sides inscribed around a circle of radius `radius`, i.e. exactly the unicycle Euler model the
 * filter integrates (Propagate.cpp:33-37), starting at the filter origin (0,0,0)
 * (kalmanfilter.cpp:10); the trajectory is exactly periodic, so one lap of records can be
 * replayed for any number of laps. Landmarks sit at equal angles, alternating radius-ring_offset
 * / radius+ring_offset. Sensor envelope from the reference: 180 degree field of view
 * (slam.cpp:90), 1 m minimum feature distance (featuredetector.h:28), 8 m range
 * (houghtransform.h:20). Noise: v_m = v + N(0,(sigma_v v)^2), w_m = w + N(0,(sigma_w v)^2)
 * (matching Q, kalmanfilter.cpp:28-37); range sigma 0.05 m and bearing sigma 0.01 rad (matching R,
 * slam.cpp:165). Visible landmarks are measured round-robin, max_meas per step.
 * Randomness is counter-based (seed, filter, step, stream), so any sub-range of filters/steps can
 * be generated independently and reproducibly.
 */
#ifndef EKF_SYNTH_H
#define EKF_SYNTH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ekf_synth_config {
  int32_t n_landmarks;     /* landmarks in the world */
  int32_t steps_per_lap;   /* polygon sides T */
  int32_t max_meas;        /* measurements per step (record capacity) */
  int32_t compass_every;   /* 0 = no compass; k = a compass reading every k-th step */
  double dt;               /* 0.2 s */
  double radius;           /* circle radius (m) */
  double ring_offset;      /* 3 m */
  double sigma_v, sigma_w; /* 0.01, 0.04 (scaled by v) */
  double sigma_range, sigma_bearing; /* 0.05 m, 0.01 rad */
  double min_range, max_range;       /* 1 m, 8 m */
  double fov;              /* pi */
  double sigma_compass;    /* sqrt(0.0005) */
  double compass_R;        /* 0.0005 (slam.cpp:146) */
  uint64_t seed;
} ekf_synth_config;

/* radius = max(5, 0.2*n_landmarks); steps_per_lap = 1000; max_meas = 1; no compass. */
void ekf_synth_default_config(ekf_synth_config* cfg, int n_landmarks);
int ekf_synth_record_len(const ekf_synth_config* cfg);
/* lm_xy[n_landmarks][2] */
void ekf_synth_world(const ekf_synth_config* cfg, double* lm_xy);
/* True pose at global step index t (pose AFTER t propagation steps): xyphi[3]. */
void ekf_synth_true_pose(const ekf_synth_config* cfg, long t, double* xyphi);
/* Records for filters [f0,f0+nf) and steps [t0,t0+nt) -> out[nf][nt][record_len].
 * lm_ids (optional) [nf][nt][max_meas]: which world landmark each measurement came from (-1 none).
 * Returns 0, or non-zero on bad arguments. n_threads <= 0 = all hardware threads. */
int ekf_synth_generate(const ekf_synth_config* cfg, long f0, int nf, long t0, int nt, double* out, int32_t* lm_ids,
                       int n_threads);
/* What the LMS-200 would return at the TRUE pose of global step t (slam.cpp:90: 181 beams over the
 * front 180 degrees in 1-degree steps, beam 0 at -90 degrees): every landmark is a square pillar of
 * side 0.3 m centred on it, a beam returns the distance to the nearest pillar face, or 8191 mm when
 * nothing is within 8.191 m. local_x_mm / local_y_mm are ArSensorReading::getLocalX/Y (x forward, y
 * left), range_mm is getRange (integer millimetres). Arrays of EKF_SYNTH_SCAN_BEAMS. Deterministic.
 * Used for the scanRun.txt log (slam.cpp:184-203). Returns the number of beams written. */
#define EKF_SYNTH_SCAN_BEAMS 181
int ekf_synth_scan(const ekf_synth_config* cfg, long t, double* local_x_mm, double* local_y_mm, uint32_t* range_mm);
/* slam.cpp:158-167: corner feature (mm, robot frame) -> z (m), R (column-major 2x2). */
void ekf_synth_measurement_from_feature(double fx_mm, double fy_mm, double* z, double* R);

#ifdef __cplusplus
}
#endif
#endif /* EKF_SYNTH_H */
