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

#endif  // EKF_SHIM_ARIA_H
