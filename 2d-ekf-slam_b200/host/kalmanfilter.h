// kalmanfilter.h — drop-in replacement for the reference's odometry/kalmanfilter.h.
//
// Same class name, same public members and methods (odometry/kalmanfilter.h:24-32), so the SLAM
// loop of slam.cpp:127-182 compiles and runs against it unchanged; the arithmetic of
// odometry/kalmanfilter.cpp, Propagate.cpp and Update.cpp runs on a B200 behind the C ABI of
// include/ekf_slam_b200.h. Header-only on purpose: like the reference it includes <Eigen/Dense>
// and "Aria.h" and therefore compiles against whatever Eigen / ARIA (or stand-ins) the
// application uses; only (i,j), size() are needed from the matrix type.
//
// Differences a maintainer should know about:
//   * the reference grows its state without bound (Update.cpp:158-177); a GPU handle has a landmark
//     capacity. The class keeps the reference's behaviour: before an update that could overflow it
//     doubles the capacity (ekf_resize: new handle, device-to-device copy of the state), so "New " is
//     never refused while device memory lasts. Growth can be switched off (setGrowth(false)): then a
//     New association arriving with the map full is dropped, "Full " is printed instead of "New ", and
//     status() reports EKF_ERR_CAPACITY. The sharded mode (one map over several GPUs) has a fixed capacity;
//   * a third constructor takes a list of device ordinals and runs the filter as ONE map whose covariance
//     is column-sharded over those GPUs (ekf_sharded_*, NVLink exchange of the gain rows): for maps too
//     large or too slow for one GPU. Same call surface, same results;
//   * the reference never frees its state (no destructor); this class releases the GPU handle;
//   * one filter per object is a latency-bound use of a GPU: batches of filters should use the
//     C ABI directly (ekf_create with n_filters > 1, ekf_run).
#ifndef KALMANFILTER_H
#define KALMANFILTER_H

#include <Eigen/Dense>
#include <cmath>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "Aria.h"
#include "ekf_slam_b200.h"

#define INF 999999999999
#define PI 3.141592653589793238462643383279502884197169399375105820974944592307816406286

class KalmanFilter {
 public:
  double X = 0.0;
  double Y = 0.0;
  double Phi = 0.0;
  int Num_Landmarks = 0;

  static const int kDefaultMaxLandmarks = 256;

  explicit KalmanFilter(ArRobot* robot) : KalmanFilter(robot, kDefaultMaxLandmarks, 0) {}
  KalmanFilter(ArRobot* robot, int max_landmarks, int device) : robot(robot), capacity(max_landmarks) {
    const int rc = ekf_create(&handle, device, 1, max_landmarks, nullptr);
    if (rc != EKF_OK) throw std::runtime_error(std::string("ekf_create: ") + ekf_last_error(nullptr));
  }
  // One map column-sharded over several GPUs (devices may repeat an ordinal).
  KalmanFilter(ArRobot* robot, int max_landmarks, const std::vector<int>& devices) : robot(robot), capacity(max_landmarks) {
    const int rc = ekf_sharded_create(&sharded, static_cast<int>(devices.size()), devices.data(), max_landmarks, nullptr);
    if (rc != EKF_OK) throw std::runtime_error(std::string("ekf_sharded_create: ") + ekf_sharded_last_error(nullptr));
    grow = false;
  }
  ~KalmanFilter() {
    if (handle) ekf_destroy(handle);
    if (sharded) ekf_sharded_destroy(sharded);
  }
  KalmanFilter(const KalmanFilter&) = delete;
  KalmanFilter& operator=(const KalmanFilter&) = delete;

  // kalmanfilter.cpp:15-62
  void doPropagation(double dt, std::ofstream& covFile, std::ofstream& knownfeaturesFile) {
    robot->lock();
    double V = robot->getVel();         // mm/s; converted on the device exactly as :18,26
    double RTV = robot->getRotVel();    // deg/s; converted on the device exactly as :19
    robot->unlock();
    if (sharded) check(ekf_sharded_propagate(sharded, V, RTV, dt));
    else check(ekf_propagate(handle, &V, &RTV, &dt, 0));
    refresh();
    if (covFile.is_open()) {            // :51  P00 P01 P10 P11
      double b[4];
      if (sharded) {
        double prr[9];
        int nl = 0;
        std::vector<double> x(3 + 2 * static_cast<size_t>(capacity));
        check(ekf_sharded_get_replica(sharded, 0, &nl, x.data(), prr));
        b[0] = prr[0]; b[1] = prr[1]; b[2] = prr[3]; b[3] = prr[4];
      } else {
        check(ekf_get_cov_block(handle, 0, 0, 0, 2, 2, b, 2));
      }
      covFile << b[0] << " " << b[2] << " " << b[1] << " " << b[3] << std::endl;
    }
    if (knownfeaturesFile.is_open() && Num_Landmarks > 0) {   // :53-61, index stride as in the reference
      std::vector<double> x(3 + 2 * static_cast<size_t>(Num_Landmarks));
      int nl = 0;
      if (sharded) check(ekf_sharded_get_state(sharded, &nl, x.data(), nullptr, 0));
      else check(ekf_get_state(handle, 0, &nl, x.data(), nullptr, 0));
      for (int i = 1; i < nl; i++) knownfeaturesFile << x[3 + i] << " " << x[4 + i] << std::endl;
    }
  }

  // kalmanfilter.cpp:64-90 -> Update.cpp:22-204. z_chunk is 2 x n_z, R_chunk is 2 x 2n_z.
  void doUpdate(Eigen::MatrixXd z_chunk, Eigen::MatrixXd R_chunk) {
    const int n_z = static_cast<int>(z_chunk.size() / 2);
    if (n_z <= 0) return;
    std::vector<double> z(2 * static_cast<size_t>(n_z)), R(4 * static_cast<size_t>(n_z));
    for (int j = 0; j < n_z; ++j) {
      z[2 * j + 0] = z_chunk(0, j);
      z[2 * j + 1] = z_chunk(1, j);
      R[4 * j + 0] = R_chunk(0, 2 * j);       // column-major 2x2 block j (Update.cpp:86)
      R[4 * j + 1] = R_chunk(1, 2 * j);
      R[4 * j + 2] = R_chunk(0, 2 * j + 1);
      R[4 * j + 3] = R_chunk(1, 2 * j + 1);
    }
    decisions.assign(n_z, EKF_DECISION_NONE);
    indices.assign(n_z, -1);
    mahal.assign(n_z, 0.0);
    if (grow && handle && Num_Landmarks + n_z > capacity) {
      // every measurement of this call could start a landmark (Update.cpp:152-178): make room first
      int want = capacity;
      while (want < Num_Landmarks + n_z) want *= 2;
      check(ekf_resize(&handle, want));
      capacity = want;
    }
    const int rc = sharded ? ekf_sharded_update(sharded, n_z, z.data(), R.data(), decisions.data(), indices.data(), mahal.data())
                           : ekf_update(handle, n_z, z.data(), R.data(), decisions.data(), indices.data(), mahal.data());
    last_status = rc;
    if (rc != EKF_OK && rc != EKF_ERR_CAPACITY) check(rc);
    for (int j = 0; j < n_z; ++j) {           // the tokens Update.cpp:154,183,191 print
      switch (decisions[j]) {
        case EKF_DECISION_NEW: std::cout << "New "; break;
        case EKF_DECISION_OLD: std::cout << "Old "; break;
        case EKF_DECISION_IGNORE: std::cout << "Ignore "; break;
        default: std::cout << "Full "; break;
      }
    }
    refresh();
  }

  // kalmanfilter.cpp:96-130
  void doUpdateCompass(double z, double R) {
    if (sharded) check(ekf_sharded_update_compass(sharded, z, R));
    else check(ekf_update_compass(handle, &z, &R, nullptr));
    refresh();
  }

  // ---- additions (not in the reference) ---------------------------------------------------------
  int status() const { return last_status; }                         // EKF_OK or EKF_ERR_CAPACITY
  const std::vector<int32_t>& lastDecisions() const { return decisions; }
  const std::vector<int32_t>& lastLandmarkIndices() const { return indices; }   // Opt_i per measurement
  const std::vector<double>& lastMahalanobis() const { return mahal; }
  ekf_handle nativeHandle() const { return handle; }                 // null in the sharded mode
  ekf_sharded nativeShardedHandle() const { return sharded; }        // null in the single-GPU mode
  int maxLandmarks() const { return capacity; }                      // current capacity (grows on demand)
  void setGrowth(bool on) { grow = on && handle != nullptr; }

 private:
  ArRobot* robot;
  ekf_handle handle = nullptr;
  ekf_sharded sharded = nullptr;
  int capacity = 0;
  bool grow = true;
  int last_status = EKF_OK;
  std::vector<int32_t> decisions, indices;
  std::vector<double> mahal;

  void check(int rc) {
    if (rc != EKF_OK)
      throw std::runtime_error(std::string("ekf_slam_b200: ") + (sharded ? ekf_sharded_last_error(sharded) : ekf_last_error(handle)));
  }
  void refresh() {   // X, Y, Phi, Num_Landmarks mirrors (kalmanfilter.cpp:46-48,85-89,127-129)
    double p[3];
    int32_t nl = 0;
    if (sharded) check(ekf_sharded_get_pose(sharded, p, &nl));
    else check(ekf_get_pose(handle, p, &nl));
    X = p[0];
    Y = p[1];
    Phi = p[2];
    Num_Landmarks = nl;
  }
};

#endif  // KALMANFILTER_H
