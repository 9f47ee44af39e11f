// TEST INFRASTRUCTURE ONLY. Stand-in for MobileRobots ARIA's "Aria.h" (absent from this image).
// The EKF hot path touches exactly four ArRobot methods (odometry/kalmanfilter.cpp:17-20):
// lock(), getVel() [mm/s], getRotVel() [deg/s], unlock(). The synthetic driver sets the two
// velocity fields before each doPropagation call.
#ifndef EKF_SHIM_ARIA_H
#define EKF_SHIM_ARIA_H

class ArRobot {
 public:
  double vel_mm_s = 0.0;
  double rotvel_deg_s = 0.0;
  int lock() { return 0; }
  int unlock() { return 0; }
  double getVel() const { return vel_mm_s; }
  double getRotVel() const { return rotvel_deg_s; }
};

// The Hough front-end (features/houghtransform.cpp:240-250) reads three things of a laser return:
// its range [mm] and its position in the robot frame [mm].
class ArSensorReading {
 public:
  ArSensorReading() {}
  ArSensorReading(double x_mm, double y_mm, unsigned int range_mm) : x_(x_mm), y_(y_mm), range_(range_mm) {}
  unsigned int getRange() const { return range_; }
  double getLocalX() const { return x_; }
  double getLocalY() const { return y_; }

 private:
  double x_ = 0.0, y_ = 0.0;
  unsigned int range_ = 0;
};

// FeatureDetector (features/featuredetector.{h,cpp}) additionally names the laser device and a
// time stamp; only FeatureDetector::getFeatures touches them, and the harness calls the stages
// behind it directly, so these only have to exist.
#include <vector>
class ArTime {
 public:
  bool isAt(const ArTime&) const { return false; }
};
class ArSick {
 public:
  int lockDevice() { return 0; }
  int unlockDevice() { return 0; }
  std::vector<ArSensorReading>* getRawReadingsAsVector() { return &readings_; }
  ArTime getLastReadingTime() const { return ArTime(); }
  std::vector<ArSensorReading> readings_;
};

#endif  // EKF_SHIM_ARIA_H
