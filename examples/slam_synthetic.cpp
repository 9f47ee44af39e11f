// slam_synthetic.cpp — the reference's SLAM loop (slam.cpp:127-182) driving the GPU filter core
// through the drop-in KalmanFilter class, with the robot / laser I/O replaced by the synthetic
// driver (include/ekf_synth.h). Everything between "Enter SLAM loop" and the odometry log line is
// the reference's call sequence: doPropagation, optional doUpdateCompass, one doUpdate per
// feature with R built from the feature as slam.cpp:158-167 does.
//
// Build (tests/test_dropin.py does this; any <Eigen/Dense> + Aria.h pair works, here the
// stand-ins under oracle/shim are used because real Eigen / ARIA are not installed):
//   g++ -std=c++11 -O2 -Ioracle/shim -I2d-ekf-slam_b200/host -Iinclude examples/slam_synthetic.cpp \
//       -L2d-ekf-slam_b200/lib -lekf_slam_b200 -lekf_synth -o slam_synthetic
// Usage: slam_synthetic [n_landmarks] [n_steps] [max_landmarks] [log_dir]  -> one line per step on
//   stdout: "Update: <tokens><Num_Landmarks>" lines as slam.cpp:169-171 prints, then "odom X Y Phi".
//   With log_dir the reference's log files are written there in its own text formats
//   (odomRun.txt slam.cpp:181, featuresRun.txt slam.cpp:172-177, covRun.txt and knownfeaturesRun.txt
//   kalmanfilter.cpp:51-61, scanRun.txt slam.cpp:184-203 from synthetic LMS-200 scans of the landmark
//   world), so plot.py / RealTimePlotting.m read a synthetic run unchanged.
//   A fifth argument "shards=<k>" runs the filter as ONE map column-sharded over k shards (devices 0..).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "ekf_synth.h"
#include "kalmanfilter.h"

int main(int argc, char** argv) {
  const int n_landmarks = argc > 1 ? std::atoi(argv[1]) : 20;
  const int n_steps = argc > 2 ? std::atoi(argv[2]) : 200;
  const int max_landmarks = argc > 3 ? std::atoi(argv[3]) : n_landmarks + 4;

  ekf_synth_config cfg;
  ekf_synth_default_config(&cfg, n_landmarks);
  cfg.steps_per_lap = n_steps;
  cfg.max_meas = 2;
  cfg.compass_every = 10;
  const int L = ekf_synth_record_len(&cfg);
  std::vector<double> rec(static_cast<size_t>(n_steps) * L);
  ekf_synth_generate(&cfg, 0, 1, 0, n_steps, rec.data(), nullptr, 1);

  ArRobot robot;
  std::ofstream odomFile, featuresFile, covFile, knownfeaturesFile, scanFile;   // closed unless log_dir is given
  if (argc > 4 && std::string(argv[4]) != "-") {
    const std::string dir(argv[4]);
    scanFile.open((dir + "/scanRun.txt").c_str());
    odomFile.open((dir + "/odomRun.txt").c_str());
    featuresFile.open((dir + "/featuresRun.txt").c_str());
    covFile.open((dir + "/covRun.txt").c_str());
    knownfeaturesFile.open((dir + "/knownfeaturesRun.txt").c_str());
  }
  std::cout.precision(17);

  // Initialize the kalman filter (slam.cpp:127)
  KalmanFilter* ekf;
  if (argc > 5 && std::string(argv[5]).rfind("shards=", 0) == 0) {
    const int k = std::atoi(argv[5] + 7);
    int n_dev = 1;
    ekf_device_count(&n_dev);
    std::vector<int> devices;
    for (int s = 0; s < k; ++s) devices.push_back(s % (n_dev > 0 ? n_dev : 1));
    ekf = new KalmanFilter(&robot, max_landmarks, devices);
  } else {
    ekf = new KalmanFilter(&robot, max_landmarks, 0);
  }
  double loopTime = 0.0;

  // Enter SLAM loop (slam.cpp:130)
  for (int t = 0; t < n_steps; ++t) {
    const double* r = &rec[static_cast<size_t>(t) * L];
    robot.vel_mm_s = r[0];
    robot.rotvel_deg_s = r[1];
    double dt = r[2];
    ekf->doPropagation(dt, covFile, knownfeaturesFile);              // slam.cpp:136

    if (r[6] != 0.0) ekf->doUpdateCompass(r[3], r[4]);               // slam.cpp:144-147

    const int nz = static_cast<int>(r[5]);
    for (int i = 0; i < nz; i++) {                                   // slam.cpp:150-171
      const double* zr = r + 8 + 6 * i;
      Eigen::MatrixXd z_chunk(2, 1);
      Eigen::MatrixXd R_chunk(2, 2);
      z_chunk << zr[0], zr[1];
      R_chunk << zr[2], zr[4], zr[3], zr[5];   // row-major fill of the column-major record slot
      std::cout << "Update: ";
      ekf->doUpdate(z_chunk, R_chunk);
      std::cout << ekf->Num_Landmarks << std::endl;
      if (featuresFile.is_open()) {                                  // slam.cpp:172-177
        const double fx = zr[0], fy = zr[1];
        const double newX = fx * cos(ekf->Phi) - fy * sin(ekf->Phi);
        const double newY = fx * sin(ekf->Phi) + fy * cos(ekf->Phi);
        featuresFile << newX + ekf->X << " " << newY + ekf->Y << std::endl;
      }
    }
    if (odomFile.is_open()) odomFile << ekf->X << " " << ekf->Y << std::endl;          // slam.cpp:181
    std::cout << "odom " << ekf->X << " " << ekf->Y << " " << ekf->Phi << std::endl;   // slam.cpp:181

    // If haven't saved laser scan in over a second, save the laser scan (slam.cpp:184-203)
    loopTime += dt;
    if (loopTime > 1.0) {
      if (scanFile.is_open()) {
        double lx[EKF_SYNTH_SCAN_BEAMS], ly[EKF_SYNTH_SCAN_BEAMS];
        uint32_t range[EKF_SYNTH_SCAN_BEAMS];
        const int nb = ekf_synth_scan(&cfg, t + 1, lx, ly, range);   // the robot has made t+1 moves
        for (int i = 0; i < nb; i++) {
          if (range[i] > 7000) continue;
          double fx = lx[i] / 1000.0;
          double fy = ly[i] / 1000.0;
          double newX = fx * cos(ekf->Phi) - fy * sin(ekf->Phi);
          double newY = fx * sin(ekf->Phi) + fy * cos(ekf->Phi);
          scanFile << newX + ekf->X << " " << newY + ekf->Y << std::endl;
        }
      }
      loopTime = 0.0;
    }
  }
  delete ekf;
  return 0;
}
