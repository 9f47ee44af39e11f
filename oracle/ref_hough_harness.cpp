// TEST INFRASTRUCTURE ONLY - never linked into the product.
//
// C ABI over the reference's own Hough front-end (features/houghtransform.{h,cpp}), compiled
// UNMODIFIED from /root/reference over the stand-in Aria.h (oracle/shim). Used to pin the C
// restatement (oracle/hough_oracle.c) and to generate the CPU baseline of the Hough leg.
// HoughTransform::getLines (houghtransform.cpp:40) is declared int and flows off its end. g++ 13
// compiles that to an unreachable point; the TU is built with -funreachable-traps, which makes the
// point a trap instruction placed AFTER all of the function's work and the destruction of its
// locals, and call_get_lines() below resumes from the SIGILL with siglongjmp. The reference source
// stays unmodified and every line it pushed into the caller's vector is kept.
#include <csetjmp>
#include <csignal>
#include <cstring>
#include <vector>

#define private public
#include "houghtransform.h"
#undef private

namespace {

thread_local sigjmp_buf tls_resume;
thread_local volatile int tls_armed = 0;

void on_sigill(int) {
  if (tls_armed) siglongjmp(tls_resume, 1);
  std::signal(SIGILL, SIG_DFL);   // not ours: die the normal way
  std::raise(SIGILL);
}

void call_get_lines(HoughTransform& h, std::vector<ArSensorReading>* readings, std::vector<struct houghLine>* lines) {
  static const bool installed = [] {
    struct sigaction sa;
    std::memset(&sa, 0, sizeof(sa));
    sa.sa_handler = on_sigill;
    sa.sa_flags = SA_NODEFER;
    sigaction(SIGILL, &sa, nullptr);
    return true;
  }();
  (void)installed;
  if (sigsetjmp(tls_resume, 1) == 0) {
    tls_armed = 1;
    h.getLines(readings, lines);
  }
  tls_armed = 0;
}

}  // namespace

extern "C" {

// The trigonometric tables the reference constructor builds (houghtransform.cpp:5-18), 180 floats each.
void refh_tables(float* cos_out, float* sin_out) {
  HoughTransform h;
  std::memcpy(cos_out, h.COS_ARRAY, sizeof(float) * HoughTransform::THETA_SIZE);
  std::memcpy(sin_out, h.SIN_ARRAY, sizeof(float) * HoughTransform::THETA_SIZE);
}

int refh_theta_size(void) { return HoughTransform::THETA_SIZE; }
int refh_radius_size(void) { return HoughTransform::RADIUS_SIZE; }
int refh_num_peaks(void) { return HoughTransform::NUM_PEAKS; }

// One scan through HoughTransform::getLines. x, y [mm] and range [mm] of n readings.
// lines_out: up to max_lines triples (radius, theta, weight); returns the number of lines.
// grid_out (THETA_SIZE*RADIUS_SIZE bytes) and peaks_out (NUM_PEAKS ints) may be NULL; they are the
// accumulator after performHoughTransform and the array getPeaks fills (recomputed here through
// the same private methods, on a second object, so getLines itself runs untouched).
int refh_get_lines(int n, const double* x, const double* y, const unsigned int* range, double* lines_out,
                   int max_lines, unsigned char* grid_out, int* peaks_out) {
  std::vector<ArSensorReading> readings;
  readings.reserve(n);
  for (int i = 0; i < n; ++i) readings.push_back(ArSensorReading(x[i], y[i], range[i]));
  std::vector<struct houghLine> lines;
  {
    HoughTransform h;
    call_get_lines(h, &readings, &lines);
  }
  if (grid_out || peaks_out) {
    HoughTransform h;
    h.performHoughTransform(&readings);
    if (grid_out) std::memcpy(grid_out, h.houghGrid, (size_t)HoughTransform::THETA_SIZE * HoughTransform::RADIUS_SIZE);
    if (peaks_out) {
      int peaks[HoughTransform::NUM_PEAKS] = {0};
      h.getPeaks(HoughTransform::NUM_PEAKS, peaks);
      std::memcpy(peaks_out, peaks, sizeof(peaks));
    }
  }
  const int m = (int)lines.size() < max_lines ? (int)lines.size() : max_lines;
  for (int i = 0; i < m; ++i) {
    lines_out[3 * i + 0] = lines[i].radius;
    lines_out[3 * i + 1] = lines[i].theta;
    lines_out[3 * i + 2] = lines[i].weight;
  }
  return (int)lines.size();
}

// Timed loop for the CPU baseline: n_scans scans of n readings each, one after another on this
// thread; returns the number of lines found in total (so the work cannot be optimised away).
long refh_run_scans(int n_scans, int n, const double* x, const double* y, const unsigned int* range) {
  long total = 0;
  HoughTransform h;
  std::vector<ArSensorReading> readings(n);
  for (int s = 0; s < n_scans; ++s) {
    for (int i = 0; i < n; ++i) readings[i] = ArSensorReading(x[(size_t)s * n + i], y[(size_t)s * n + i], range[(size_t)s * n + i]);
    std::vector<struct houghLine> lines;
    call_get_lines(h, &readings, &lines);
    h.clearHoughGrid();                       // featuredetector.cpp:40-41
    total += (long)lines.size();
  }
  return total;
}

}  // extern "C"

// ---- the rest of the measurement front-end: FeatureDetector (features/featuredetector.cpp) ---------
// FeatureDetector::getFeatures (featuredetector.cpp:16-70) is: getLines, clearHoughGrid,
// fitLineSegments, extractCorners, getStructCompass. It cannot be called as a whole here (its
// getLines call ends in the trap described above, inside a frame we cannot resume), so the harness
// runs the same five stages in the same order through the class's own (private) methods.
#define private public
#include "featuredetector.h"
#undef private

extern "C" {

// One scan through the stages of getFeatures. Outputs: features (x, y) pairs [max_feats],
// returns their number; segments_out (optional) 7 doubles each: radius, theta, startX, startY,
// endX, endY, numPoints [max_segs], *n_segs; compass_out: getStructCompass(lines, cur_phi) with the
// detector's COMPASS_OFFSET preset to *offset_io (100.0 = unset) and written back.
int reff_get_features(int n, const double* x, const double* y, const unsigned int* range, double cur_phi,
                      double* offset_io, double* feats_out, int max_feats, double* segments_out, int max_segs,
                      int* n_segs, double* compass_out) {
  std::vector<ArSensorReading> readings;
  readings.reserve(n);
  for (int i = 0; i < n; ++i) readings.push_back(ArSensorReading(x[i], y[i], range[i]));
  ArSick sick;
  FeatureDetector fd(&sick);
  if (offset_io) fd.COMPASS_OFFSET = *offset_io;
  std::vector<struct houghLine> lines;
  call_get_lines(*fd.hough, &readings, &lines);                       // featuredetector.cpp:39-40
  fd.hough->clearHoughGrid();                                        // :41
  std::vector<FeatureDetector::lineSegment> segs;
  fd.fitLineSegments(&readings, &lines, &segs);                      // :44-45
  std::vector<Feature> feats;
  fd.extractCorners(&feats, &segs);                                  // :60
  const double compass = fd.getStructCompass(&lines, cur_phi);       // :69
  if (compass_out) *compass_out = compass;
  if (offset_io) *offset_io = fd.COMPASS_OFFSET;
  if (n_segs) *n_segs = (int)segs.size();
  if (segments_out)
    for (int i = 0; i < (int)segs.size() && i < max_segs; ++i) {
      double* o = segments_out + 7 * i;
      o[0] = segs[i].radius; o[1] = segs[i].theta; o[2] = segs[i].startX; o[3] = segs[i].startY;
      o[4] = segs[i].endX; o[5] = segs[i].endY; o[6] = segs[i].numPoints;
    }
  for (int i = 0; i < (int)feats.size() && i < max_feats; ++i) {
    feats_out[2 * i] = feats[i].x;
    feats_out[2 * i + 1] = feats[i].y;
  }
  return (int)feats.size();
}

}  // extern "C"
