/* TEST INFRASTRUCTURE ONLY - never linked into the product (only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this library).
 *
 * Plain-C restatement of the reference's Hough line extractor, HoughTransform::getLines
 * (features/houghtransform.{h,cpp}), the measurement front-end listed as the third "next" row of
 * SURVEY.md 8(f). PARITY STATUS: pinned by execution - tests/test_hough_oracle.py compares the
 * accumulator, the peak array and the lines bit for bit with the reference's own translation unit
 * (oracle/_ref/libhough_ref.so, built by oracle/Makefile from /root/reference) on seeded scans.
 *
 * Constants: houghtransform.h:20-30. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum {
  H_MAX_DIST = 8000,                               /* mm */
  H_DISTANCE = 10,                                 /* mm per radius bin */
  H_THETA = 180,
  H_RADIUS = 2 * H_MAX_DIST / H_DISTANCE + 1,      /* 1601 */
  H_ADD = H_RADIUS / 2,                            /* 800 */
  H_PEAKS = 200,
  H_MERGE_THETA = 30,
  H_MERGE_RADIUS = 5
};

/* houghtransform.cpp:5-18: theta accumulates in float, cos/sin are the double functions of it,
 * stored as float. */
void hough_oracle_tables(float* cos_out, float* sin_out) {
  const float d_theta = (float)(3.141592654 / H_THETA);
  float theta = 0.0f;
  for (int i = 0; i < H_THETA; ++i) {
    cos_out[i] = (float)cos((double)theta);
    sin_out[i] = (float)sin((double)theta);
    theta += d_theta;
  }
}

/* houghtransform.cpp:240-256. grid: H_THETA*H_RADIUS bytes, incremented in place. */
void hough_oracle_accumulate(int n, const double* x, const double* y, const unsigned int* range,
                             const float* cos_tab, const float* sin_tab, unsigned char* grid) {
  for (int i = 0; i < n; ++i) {
    if (range[i] > (unsigned int)H_MAX_DIST) continue;
    for (int t = 0; t < H_THETA; ++t) {
      int radius = (int)round(x[i] * (double)cos_tab[t] + y[i] * (double)sin_tab[t]);
      radius /= H_DISTANCE;                        /* truncates toward zero */
      radius += H_ADD;
      grid[t * H_RADIUS + radius]++;               /* unsigned char: wraps at 256 */
    }
  }
}

/* houghtransform.cpp:260-280: streaming selection, order-dependent. peaks: H_PEAKS ints, all
 * zero on entry (houghtransform.cpp:46). */
void hough_oracle_peaks(const unsigned char* grid, int* peaks) {
  int mindex = 0;
  for (int cell = 0; cell < H_THETA * H_RADIUS; ++cell) {
    const int v = grid[cell];
    if (v > grid[peaks[mindex]]) {
      peaks[mindex] = cell;
      for (int i = 0; i < H_PEAKS; ++i)
        if (grid[peaks[i]] < grid[peaks[mindex]]) mindex = i;
    }
  }
}

typedef struct {
  int max_r, min_r, max_t, min_t;
  int radius, theta, weight, n_points;             /* radius, theta: weighted sums */
} HoughGroup;

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* houghtransform.cpp:58-236: peaks -> groups -> merged groups -> lines (radius mm, theta rad,
 * weight). Returns the number of lines; at most max_lines triples are written. */
int hough_oracle_lines(const unsigned char* grid, const int* peaks, double* lines_out, int max_lines) {
  HoughGroup g[H_PEAKS];
  int n_groups = 0;
  for (int p = 0; p < H_PEAKS; ++p) {              /* :66-112, greedy, in peak-array order */
    const int r = peaks[p] % H_RADIUS, t = peaks[p] / H_RADIUS, w = grid[peaks[p]];
    if (r <= 0) continue;
    int merged = 0;
    for (int j = 0; j < n_groups && !merged; ++j) {
      HoughGroup* q = &g[j];
      const int t_in = t < q->max_t && t > q->min_t, r_in = r < q->max_r && r > q->min_r;
      const int near_t = abs(q->max_t - t) < H_MERGE_THETA || abs(q->min_t - t) < H_MERGE_THETA || t_in;
      const int near_r = abs(q->max_r - r) < H_MERGE_RADIUS || abs(q->min_r - r) < H_MERGE_RADIUS || r_in;
      if (near_t && near_r) {
        q->max_r = imax(r, q->max_r); q->min_r = imin(r, q->min_r);
        q->max_t = imax(t, q->max_t); q->min_t = imin(t, q->min_t);
        q->radius += r * w; q->theta += t * w; q->weight += w; q->n_points++;
        merged = 1;
      }
    }
    if (!merged) {
      HoughGroup* q = &g[n_groups++];
      q->max_r = q->min_r = r; q->max_t = q->min_t = t;
      q->weight = w; q->n_points = 1; q->radius = r * w; q->theta = t * w;
    }
  }
  for (int i = 0; i < n_groups; ++i) {             /* :118-129: mirror negative radii */
    HoughGroup* q = &g[i];
    if (q->radius < H_ADD * q->weight) {
      q->radius = 2 * H_ADD * q->weight - q->radius;
      q->max_r = 2 * H_ADD - q->max_r; q->min_r = 2 * H_ADD - q->min_r;
      q->theta -= H_THETA * q->weight;
      q->max_t -= H_THETA; q->min_t -= H_THETA;
    }
  }
  int parent[H_PEAKS];                             /* :158-190 (the reference keeps these in a char array) */
  for (int i = 0; i < n_groups; ++i) parent[i] = -1;
  for (int i = 0; i < n_groups; ++i) {
    const HoughGroup* a = &g[i];
    for (int j = i + 1; j < n_groups; ++j) {
      const HoughGroup* b = &g[j];
      const int near_t = abs(b->max_t - a->min_t) < H_MERGE_THETA || abs(b->min_t - a->max_t) < H_MERGE_THETA ||
                         (a->max_t > b->min_t && a->min_t < b->max_t);
      const int near_r = abs(b->max_r - a->min_r) < H_MERGE_RADIUS || abs(b->min_r - a->max_r) < H_MERGE_RADIUS ||
                         (a->max_r > b->min_r && a->min_r < b->max_r);
      if (near_t && near_r) parent[j] = i;
    }
  }
  for (int i = 0; i < n_groups; ++i) {             /* :194-211: fold every non-root into its root */
    if (parent[i] == -1) continue;
    int j = i;
    while (parent[j] != -1) j = parent[j];
    HoughGroup* dst = &g[j];
    const HoughGroup* src = &g[i];
    dst->max_r = imax(src->max_r, dst->max_r); dst->min_r = imin(src->min_r, dst->min_r);
    dst->max_t = imax(src->max_t, dst->max_t); dst->min_t = imin(src->min_t, dst->min_t);
    dst->radius += src->radius; dst->theta += src->theta; dst->weight += src->weight; dst->n_points += src->n_points;
  }
  int n_lines = 0;
  for (int i = 0; i < n_groups; ++i) {             /* :215-233 */
    if (parent[i] != -1) continue;
    const HoughGroup* q = &g[i];
    double theta = q->theta / (double)q->weight;
    theta *= 3.141592654 / H_THETA;
    double radius = q->radius / (double)q->weight;
    radius -= H_ADD;
    radius *= H_DISTANCE;
    const double weight = q->weight / (double)q->n_points;
    if (n_lines < max_lines) {
      lines_out[3 * n_lines + 0] = radius;
      lines_out[3 * n_lines + 1] = theta;
      lines_out[3 * n_lines + 2] = weight;
    }
    ++n_lines;
  }
  return n_lines;
}

/* HoughTransform::getLines for one scan. grid_out (H_THETA*H_RADIUS bytes) and peaks_out (H_PEAKS
 * ints) may be NULL. */
int hough_oracle_get_lines(int n, const double* x, const double* y, const unsigned int* range, const float* cos_tab,
                           const float* sin_tab, double* lines_out, int max_lines, unsigned char* grid_out,
                           int* peaks_out) {
  unsigned char* grid = grid_out ? grid_out : (unsigned char*)malloc((size_t)H_THETA * H_RADIUS);
  int peaks[H_PEAKS];
  memset(grid, 0, (size_t)H_THETA * H_RADIUS);
  memset(peaks, 0, sizeof(peaks));
  hough_oracle_accumulate(n, x, y, range, cos_tab, sin_tab, grid);
  hough_oracle_peaks(grid, peaks);
  const int n_lines = hough_oracle_lines(grid, peaks, lines_out, max_lines);
  if (peaks_out) memcpy(peaks_out, peaks, sizeof(peaks));
  if (!grid_out) free(grid);
  return n_lines;
}

/* Lines from a peak array and the accumulator values AT the peaks (what the CUDA path returns):
 * the grouping stage only ever reads grid[peaks[p]]. values: H_PEAKS ints. */
int hough_oracle_lines_from_peaks(const int* peaks, const int* values, double* lines_out, int max_lines) {
  /* a sparse stand-in grid: only the peak cells are read */
  unsigned char* grid = (unsigned char*)calloc((size_t)H_THETA * H_RADIUS, 1);
  for (int p = 0; p < H_PEAKS; ++p) grid[peaks[p]] = (unsigned char)values[p];
  const int n = hough_oracle_lines(grid, peaks, lines_out, max_lines);
  free(grid);
  return n;
}

/* Timed loop for the CPU baseline: n_scans scans of n readings each on the calling thread. */
long hough_oracle_run_scans(int n_scans, int n, const double* x, const double* y, const unsigned int* range) {
  float c[H_THETA], s[H_THETA];
  double lines[3 * H_PEAKS];
  unsigned char* grid = (unsigned char*)malloc((size_t)H_THETA * H_RADIUS);
  long total = 0;
  hough_oracle_tables(c, s);
  for (int k = 0; k < n_scans; ++k)
    total += hough_oracle_get_lines(n, x + (size_t)k * n, y + (size_t)k * n, range + (size_t)k * n, c, s, lines, H_PEAKS,
                                    grid, NULL);
  free(grid);
  return total;
}
