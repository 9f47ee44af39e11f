/* ekf_hough_b200.h - C ABI of the B200-native Hough line extractor, the measurement front-end
 * that feeds the EKF-SLAM filter core. It replaces ONE call of the reference,
 *
 *     int HoughTransform::getLines(std::vector<ArSensorReading>* readings,
 *                                  std::vector<struct houghLine>* lines)
 *         (features/houghtransform.h:32, features/houghtransform.cpp:40-236; called from
 *          FeatureDetector::getFeatures, features/featuredetector.cpp:39-41)
 *
 * for a BATCH of laser scans, all on the GPU: accumulate (houghtransform.cpp:240-256), the
 * streaming 200-peak selection (:260-280, order-dependent and reproduced exactly), and the integer
 * peak grouping / merging / line conversion (:58-236). Results are bit-identical to the
 * reference's: same accumulator bytes, same peak array, same doubles in the lines.
 * Implemented in the same shared library as ekf_slam_b200.h; sm_100a only, no CPU fallback. */
#ifndef EKF_HOUGH_B200_H
#define EKF_HOUGH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* houghtransform.h:20-30 */
#define EKF_HOUGH_MAX_DIST 8000     /* mm: returns farther than this are skipped (houghtransform.cpp:245) */
#define EKF_HOUGH_DISTANCE 10       /* mm per radius bin */
#define EKF_HOUGH_THETA_SIZE 180
#define EKF_HOUGH_RADIUS_SIZE 1601
#define EKF_HOUGH_NUM_PEAKS 200
#define EKF_HOUGH_MAX_POINTS 200    /* readings per scan this implementation accepts (an LMS-200 scan has 181) */

typedef struct ekf_hough_s* ekf_hough;

/* struct houghLine (houghtransform.h:10-14) */
typedef struct ekf_hough_line {
  double radius;   /* mm */
  double theta;    /* rad */
  double weight;
} ekf_hough_line;

/* Feature.x, Feature.y (featuredetector.h:16-19): a corner in the robot frame, mm. */
typedef struct ekf_feature {
  double x, y;
} ekf_feature;

#define EKF_HOUGH_NO_COMPASS 100.0   /* FeatureDetector::NO_COMPASS (featuredetector.h:25) */

/* Status codes are the EKF_* codes of ekf_slam_b200.h (0 = ok). */
int ekf_hough_create(ekf_hough* out, int device, int max_scans);
int ekf_hough_destroy(ekf_hough h);

/* The COS_ARRAY / SIN_ARRAY tables of the HoughTransform constructor (houghtransform.cpp:5-18),
 * 180 floats each; host only. */
void ekf_hough_tables(float* cos_out, float* sin_out);

/* getLines for n_scans scans of n_points readings each. Host arrays: x, y [n_scans][n_points]
 * (ArSensorReading::getLocalX/Y, mm), range [n_scans][n_points] (getRange, mm).
 * Outputs: lines [n_scans][max_lines] and n_lines [n_scans] (the full count, even if larger than
 * max_lines); optional (NULL to skip) peaks / values [n_scans][200] - the peaks array of
 * houghtransform.cpp:46-47 and the accumulator value at each - and grid
 * [n_scans][180*1601] - the accumulator itself (debug / parity; 288 KB per scan).
 * A reading whose radius bin falls outside the accumulator (impossible for consistent x, y, range)
 * is dropped where the reference would write out of bounds. */
int ekf_hough_get_lines(ekf_hough h, int n_scans, int n_points, const double* x, const double* y,
                        const uint32_t* range, ekf_hough_line* lines, int max_lines, int32_t* n_lines,
                        int32_t* peaks, int32_t* values, uint8_t* grid);

/* FeatureDetector::getFeatures (featuredetector.cpp:16-70) minus the device lock / time-stamp checks,
 * for n_scans scans: getLines, fitLineSegments (:74-226), extractCorners (:230-292) and, when
 * compass != NULL, getStructCompass (:297-362) - all on the GPU. feats [n_scans][max_feats] and
 * n_feats [n_scans] (full count) are the corner features slam.cpp:150-160 turns into (z, R).
 * cur_phi [n_scans] is the filter heading handed to getStructCompass (NULL = 0); compass_offset
 * [n_scans] is the detector's COMPASS_OFFSET, in/out (100.0 = not yet set; NULL = always unset);
 * compass [n_scans] receives the compass value or EKF_HOUGH_NO_COMPASS. Optional outputs: lines /
 * n_lines as above, segments [n_scans][max_segs][7] = {radius, theta, startX, startY, endX, endY,
 * numPoints} and n_segs. Integer and IEEE arithmetic as the reference's, except that sin / cos of the
 * line angles come from the device's double routines. */
int ekf_hough_get_features(ekf_hough h, int n_scans, int n_points, const double* x, const double* y,
                           const uint32_t* range, const double* cur_phi, double* compass_offset, ekf_feature* feats,
                           int max_feats, int32_t* n_feats, double* compass, ekf_hough_line* lines, int max_lines,
                           int32_t* n_lines, double* segments, int max_segs, int32_t* n_segs);

/* The same in three steps, for timing with the inputs resident in HBM: upload, run (kernels only,
 * asynchronous; at most max_lines lines per scan are kept), download (synchronises). */
int ekf_hough_upload(ekf_hough h, int n_scans, int n_points, const double* x, const double* y,
                     const uint32_t* range);
int ekf_hough_run_resident(ekf_hough h, int max_lines);
int ekf_hough_download(ekf_hough h, ekf_hough_line* lines, int max_lines, int32_t* n_lines, int32_t* peaks,
                       int32_t* values);
/* Device time of the kernel launches since the last call (CUDA events on the handle's stream). */
int ekf_hough_kernel_time(ekf_hough h, float* total_ms, int* n_launches);
int ekf_hough_sync(ekf_hough h);

/* Host only: peaks + values -> lines (houghtransform.cpp:58-236), the function the second kernel
 * runs per scan. Returns the number of lines. */
int ekf_hough_lines_from_peaks(const int32_t* peaks, const int32_t* values, ekf_hough_line* lines, int max_lines);

const char* ekf_hough_last_error(ekf_hough h);

#ifdef __cplusplus
}
#endif
#endif /* EKF_HOUGH_B200_H */
