/* TEST INFRASTRUCTURE ONLY - never linked into the product (only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this library; it is built into libhough_oracle.so).
 *
 * Plain-C restatement of the stages of FeatureDetector::getFeatures that follow the Hough
 * transform (features/featuredetector.cpp:74-362): fitLineSegments, extractCorners and
 * getStructCompass. PARITY STATUS: pinned by execution against the reference's own translation unit
 * (oracle/_ref/libhough_ref.so: reff_get_features) in tests/test_hough_oracle.py - segments,
 * corner features and the compass value bit for bit on seeded scans.
 *
 * Constants: featuredetector.h:26-36. */
#include <math.h>
#include <string.h>

#define F_MAX_DIST 8000
#define F_MIN_DIST (1000 * 1000)
#define F_MIN_POINTS 3
#define F_POINT_DIST 600
#define F_CORNER_DIST 90000
#define F_NO_COMPASS 100.0
#define F_MAX_ITEMS 512

typedef struct {
  double radius, theta;
  double sx, sy, ex, ey;   /* start / end point */
  int n_points;
  int next;                /* index of the next (older) segment of the same line, -1 = none */
} Seg;

/* featuredetector.cpp:74-226. lines: n_lines triples (radius, theta, weight). segs_out: 7 doubles
 * per segment (radius, theta, startX, startY, endX, endY, numPoints). Returns the segment count. */
int features_oracle_segments(int n, const double* x, const double* y, const unsigned int* range, const double* lines,
                             int n_lines, double* segs_out, int max_segs) {
  float sn[F_MAX_ITEMS], cs[F_MAX_ITEMS];
  int head[F_MAX_ITEMS];
  Seg pool[F_MAX_ITEMS];
  int n_pool = 0;
  if (n_lines > F_MAX_ITEMS) n_lines = F_MAX_ITEMS;
  for (int l = 0; l < n_lines; ++l) {
    sn[l] = (float)sin(lines[3 * l + 1]);
    cs[l] = (float)cos(lines[3 * l + 1]);
    head[l] = -1;
  }
  for (int r = 0; r < n; ++r) {
    if (range[r] > (unsigned int)F_MAX_DIST) continue;
    const double px = x[r], py = y[r];
    double best = 1000000.0;
    int line = 0;
    for (int l = 0; l < n_lines; ++l) {              /* closest line, the first one on ties */
      const double rad = px * cs[l] + py * sn[l];
      const double diff = fabs(lines[3 * l] - rad);
      if (diff < best) { best = diff; line = l; }
    }
    if (best > F_POINT_DIST) continue;
    /* walk the line's segments, newest first; "horizontal-ish" lines are ordered by x, the others by y */
    const int by_x = fabs(sn[line]) > fabs(cs[line]);
    const double v = by_x ? px : py;
    int s = head[line];
    while (s >= 0) {
      Seg* q = &pool[s];
      const double sv = by_x ? q->sx : q->sy, ev = by_x ? q->ex : q->ey;
      if (v <= sv && v >= ev) { q->n_points++; break; }
      if (v > sv && fabs(v - sv) <= F_POINT_DIST) { q->sx = px; q->sy = py; q->n_points++; break; }
      if (v < ev && fabs(v - ev) <= F_POINT_DIST) { q->ex = px; q->ey = py; q->n_points++; break; }
      s = q->next;
    }
    if (s < 0 && n_pool < F_MAX_ITEMS) {
      Seg* q = &pool[n_pool];
      q->theta = lines[3 * line + 1]; q->radius = lines[3 * line];
      q->n_points = 1;
      q->sx = q->ex = px; q->sy = q->ey = py;
      q->next = head[line];
      head[line] = n_pool++;
    }
  }
  int count = 0;
  for (int l = 0; l < n_lines; ++l)
    for (int s = head[l]; s >= 0; s = pool[s].next) {
      if (pool[s].n_points <= F_MIN_POINTS) continue;
      if (count < max_segs) {
        double* o = segs_out + 7 * count;
        o[0] = pool[s].radius; o[1] = pool[s].theta; o[2] = pool[s].sx; o[3] = pool[s].sy;
        o[4] = pool[s].ex; o[5] = pool[s].ey; o[6] = pool[s].n_points;
      }
      ++count;
    }
  return count;
}

/* featuredetector.cpp:230-292. segs: 7 doubles each as above. feats_out: (x, y) pairs. */
int features_oracle_corners(const double* segs, int n_segs, double* feats_out, int max_feats) {
  float sn[F_MAX_ITEMS], cs[F_MAX_ITEMS];
  const double corner_theta = 22.0 * 3.141592654 / 180.0;
  if (n_segs > F_MAX_ITEMS) n_segs = F_MAX_ITEMS;
  for (int i = 0; i < n_segs; ++i) {
    sn[i] = (float)sin(segs[7 * i + 1]);
    cs[i] = (float)cos(segs[7 * i + 1]);
  }
  int count = 0;
  for (int i = 0; i < n_segs; ++i) {
    const double* a = segs + 7 * i;
    for (int j = i + 1; j < n_segs; ++j) {
      const double* b = segs + 7 * j;
      double dth = fabs(a[1] - b[1]);
      if (dth > 3.141592654) dth = fabs(dth - 6.283185307);
      if (dth > 1.570796327) dth = fabs(dth - 3.141592654);
      if (dth < corner_theta) continue;
      const float detf = cs[i] * sn[j] - sn[i] * cs[j];          /* float arithmetic, as the reference's arrays are float */
      const double det = detf;
      const double cx = (a[0] * sn[j] - b[0] * sn[i]) / det;
      const double cy = (b[0] * cs[i] - a[0] * cs[j]) / det;
      double dx, dy;
      dx = a[2] - cx; dy = a[3] - cy; const int start1 = (dx * dx + dy * dy) < F_CORNER_DIST;
      dx = a[4] - cx; dy = a[5] - cy; const int end1 = (dx * dx + dy * dy) < F_CORNER_DIST;
      dx = b[2] - cx; dy = b[3] - cy; const int start2 = (dx * dx + dy * dy) < F_CORNER_DIST;
      dx = b[4] - cx; dy = b[5] - cy; const int end2 = (dx * dx + dy * dy) < F_CORNER_DIST;
      if ((start1 || end1) && (start2 || end2) && (cx * cx + cy * cy) > F_MIN_DIST) {
        if (count < max_feats) { feats_out[2 * count] = cx; feats_out[2 * count + 1] = cy; }
        ++count;
      }
    }
  }
  return count;
}

/* featuredetector.cpp:297-362. offset_io: the detector's COMPASS_OFFSET (100.0 = not yet set). */
double features_oracle_compass(const double* lines, int n_lines, double cur_phi, double* offset_io) {
  double g_theta[F_MAX_ITEMS], g_weight[F_MAX_ITEMS];
  int ng = 0;
  const double thresh = 10 * 3.141592654 / 180.0;
  for (int i = 0; i < n_lines; ++i) {
    const double th = lines[3 * i + 1] - 1.570796327 * floor(lines[3 * i + 1] / 1.570796327);
    const double w = lines[3 * i + 2];
    int merged = 0;
    for (int j = 0; j < ng; ++j) {                   /* no break: a line joins EVERY group it is close to */
      const double mean = g_theta[j] / g_weight[j];
      if (fabs(th - mean) < thresh) { g_theta[j] += th * w; g_weight[j] += w; merged = 1; }
    }
    if (!merged && ng < F_MAX_ITEMS) { g_theta[ng] = th * w; g_weight[ng] = w; ++ng; }
  }
  double best_theta = 0.0, best_w = 0.0;
  for (int j = 0; j < ng; ++j)
    if (g_weight[j] > best_w) { best_theta = g_theta[j]; best_w = g_weight[j]; }
  if (best_w == 0.0) return F_NO_COMPASS;
  double cardinal = -(best_theta / best_w);
  if (*offset_io == 100.0) *offset_io = cardinal;
  cardinal -= *offset_io;
  cardinal -= 1.570796327 * floor(cardinal / 1.570796327);
  cur_phi -= 6.283185307 * floor(cur_phi / 6.283185307);
  /* :346-361: six candidate errors; the result is the first candidate that is no worse than every
   * LATER one (earlier ones are not compared again), candidates 5 and 6 standing for the roll-overs */
  static const double shift[5] = {0.0, 1.570796327, 3.141592654, 4.71238898, 6.283185307};
  static const double add[6] = {0.0, 1.570796327, 3.141592654, 4.71238898, 0.0, 4.71238898};
  double err[6];
  for (int k = 0; k < 5; ++k) err[k] = fabs((cur_phi - cardinal) - shift[k]);
  err[5] = fabs((cur_phi - cardinal) + 1.570796327);
  for (int k = 0; k < 6; ++k) {
    int best = 1;
    for (int j = k + 1; j < 6; ++j) best = best && err[k] <= err[j];
    if (best) return add[k] == 0.0 ? cardinal : cardinal + add[k];
  }
  return cardinal;   /* not reached: k = 5 always qualifies */
}
