// hough_synthetic.cpp — the measurement front-end through the C ABI (include/ekf_hough_b200.h):
// a handful of synthetic LMS-200 scans of a rectangular room corner go through
// ekf_hough_get_features (Hough lines, line segments, corners, structural compass), i.e. the body of
// FeatureDetector::getFeatures (features/featuredetector.cpp:16-70) for a batch of scans.
//
// Build: g++ -std=c++11 -O2 -Iinclude examples/hough_synthetic.cpp -L2d-ekf-slam_b200/lib
//            -lekf_slam_b200 -Wl,-rpath,2d-ekf-slam_b200/lib -o hough_synthetic
// Usage: hough_synthetic [n_scans]  -> per scan: "scan s: L lines, F features, compass C" and the features.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ekf_hough_b200.h"

int main(int argc, char** argv) {
  const int S = argc > 1 ? std::atoi(argv[1]) : 4;
  const int P = 181;
  std::vector<double> x(static_cast<size_t>(S) * P), y(x.size()), phi(S, 0.0), offset(S, 100.0), compass(S);
  std::vector<uint32_t> range(x.size());
  // the robot looks into a corner: a wall 3 m ahead (x = 3000) and a wall 2.5 m to the left
  // (y = 2500); scan s is taken after turning by 5 degrees * s
  for (int s = 0; s < S; ++s) {
    const double turn = 5.0 * s * 3.141592654 / 180.0;
    for (int i = 0; i < P; ++i) {
      const double a = (i - 90) * 3.141592654 / 180.0;          // beam angle in the robot frame
      const double w = a + turn;                                // ... in the room frame
      double d = 8191.0;
      if (std::cos(w) > 1e-6) d = std::fmin(d, 3000.0 / std::cos(w));
      if (std::sin(w) > 1e-6) d = std::fmin(d, 2500.0 / std::sin(w));
      const uint32_t r = static_cast<uint32_t>(std::lround(d));
      range[static_cast<size_t>(s) * P + i] = r;
      x[static_cast<size_t>(s) * P + i] = r * std::cos(a);
      y[static_cast<size_t>(s) * P + i] = r * std::sin(a);
    }
  }
  ekf_hough h = nullptr;
  if (ekf_hough_create(&h, 0, S) != 0) {
    std::fprintf(stderr, "ekf_hough_create: %s\n", ekf_hough_last_error(nullptr));
    return 1;
  }
  const int max_feats = 8, max_lines = 16;
  std::vector<ekf_feature> feats(static_cast<size_t>(S) * max_feats);
  std::vector<ekf_hough_line> lines(static_cast<size_t>(S) * max_lines);
  std::vector<int32_t> n_feats(S), n_lines(S);
  const int rc = ekf_hough_get_features(h, S, P, x.data(), y.data(), range.data(), phi.data(), offset.data(), feats.data(),
                                        max_feats, n_feats.data(), compass.data(), lines.data(), max_lines, n_lines.data(),
                                        nullptr, 0, nullptr);
  if (rc != 0) {
    std::fprintf(stderr, "ekf_hough_get_features: %s\n", ekf_hough_last_error(h));
    return 1;
  }
  for (int s = 0; s < S; ++s) {
    std::printf("scan %d: %d lines, %d features, compass %.6f\n", s, n_lines[s], n_feats[s], compass[s]);
    for (int f = 0; f < n_feats[s] && f < max_feats; ++f)
      std::printf("  feature %.1f %.1f\n", feats[static_cast<size_t>(s) * max_feats + f].x, feats[static_cast<size_t>(s) * max_feats + f].y);
  }
  ekf_hough_destroy(h);
  return 0;
}
