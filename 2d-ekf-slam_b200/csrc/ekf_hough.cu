// ekf_hough.cu — batched Hough line extraction for LMS-200 scans (sm_100a), the measurement
// front-end of the reference (features/houghtransform.cpp), C ABI in include/ekf_hough_b200.h.
//
// One CTA per scan, the accumulator on chip. The reference's accumulator is 180 x 1601 bytes
// (288 KB), more than an SM's shared memory, and the selection stage below is serial per scan, so
// a scan is processed in PARTS slices of 180/PARTS angles, small enough that several CTAs (= several
// serial selection streams) share an SM; nothing of the accumulator ever goes to HBM unless the
// caller asks for it.
//   vote      houghtransform.cpp:240-256: one (reading, angle) pair per thread,
//             radius = (int)round(x*cos + y*sin) / 10 + 800 in the reference's arithmetic
//             (double products of the float tables, no fma), a byte-wide increment done as a
//             32-bit shared-memory atomic on the containing word (counts stay below 256)
//   compact   the cells of the slice that can still enter the peak array - count above the current
//             minimum slot, which never decreases, so everything else (all zero cells included) is
//             dropped exactly - in cell order, packed as (cell << 8 | count): two passes over a
//             contiguous chunk per thread around a block-wide exclusive scan
//   select    houghtransform.cpp:260-280 is a STREAMING top-200 selection whose result (which
//             cells, and in which slots) depends on the visiting order; the grouping stage that
//             follows is greedy in slot order, so the slots must come out exactly as the
//             reference leaves them. One warp walks the compacted list 32 candidates at a time: a ballot finds the
//             candidates above the current minimum, each is placed into the minimum slot and the
//             new minimum slot is found with one REDUX over (count << 8 | slot) keys - the
//             reference's rescan ("first slot holding a strictly smaller count, else stay").
//             The 200 slots live in the registers of that warp (7 per lane).
// While warp 0 runs the selection of slice s, the other warps already zero the accumulator and cast
// the votes of slice s+1.
// Peak grouping, merging and the conversion to (radius, theta, weight) lines
// (houghtransform.cpp:58-236) is integer work on <= 200 items per scan: a second kernel runs it
// with one warp per scan (a sequential version of the same logic is exported for host use).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ekf_hough_b200.h"
#include "ekf_slam_b200.h"

namespace {

#ifndef EKF_HOUGH_PARTS
#define EKF_HOUGH_PARTS 30
#endif
#ifndef EKF_HOUGH_THREADS
#define EKF_HOUGH_THREADS 128
#endif
constexpr int kThreads = EKF_HOUGH_THREADS;
constexpr int TH = EKF_HOUGH_THETA_SIZE, RS = EKF_HOUGH_RADIUS_SIZE, ADD = RS / 2, PK = EKF_HOUGH_NUM_PEAKS;
constexpr int PARTS = EKF_HOUGH_PARTS, TPP = TH / PARTS;   // angles per slice
static_assert(TH % PARTS == 0, "slices of equal size");
constexpr int ACC_BYTES = TPP * RS;
constexpr int ACC_WORDS = (ACC_BYTES + 3) / 4;
constexpr int ACC_WORDS_PAD = (ACC_WORDS + 31) / 32 * 32;
constexpr int MAXP = EKF_HOUGH_MAX_POINTS;
// Votes go into byte-wide cells through 32-bit atomics on the containing word: a cell can receive at
// most one vote per reading, so it cannot carry into its neighbour as long as a scan has < 256 readings.
static_assert(EKF_HOUGH_MAX_POINTS <= 255, "byte-wide Hough cells would carry into the neighbouring cell");
constexpr int CAND_CAP = MAXP * TPP;                       // a half cannot hold more non-zero cells than votes
constexpr int SLOTS_PER_LANE = (PK + 31) / 32;             // 7

struct HoughSmem {
  unsigned int acc[ACC_WORDS_PAD];
  unsigned int cand[CAND_CAP];
  double px[MAXP], py[MAXP];
  float cs[TH], sn[TH];
  unsigned char valid[(MAXP + 15) / 16 * 16];
  int warp_sum[kThreads / 32];
  int n_cand;
  int grid0;
  int minval;   // count in the current minimum slot after the last selection pass
};

struct HoughArgs {
  const double* x;
  const double* y;
  const unsigned int* range;
  const float* cos_tab;
  const float* sin_tab;
  int n_scans, n_points;
  int* peaks;            // [n_scans][PK]
  int* values;           // [n_scans][PK]
  unsigned char* grid;   // [n_scans][TH*RS] or null
};

__device__ __forceinline__ void named_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void zero_acc(HoughSmem& sm, int t0, int nt) {
  for (int w = t0; w < ACC_WORDS_PAD; w += nt) sm.acc[w] = 0u;
}

// houghtransform.cpp:240-256 for the angles of one slice. A thread keeps one angle (its cos / sin
// as doubles in registers) and walks every G-th reading; round() is spelled out (truncate, then
// adjust by the exact fractional part: half away from zero, as round() does).
__device__ __forceinline__ void vote(HoughSmem& sm, int part, int n_points, int t0, int nt) {
  const int G = nt / TPP;
  const int g = t0 / TPP, tl = t0 - g * TPP;
  if (g >= G) return;
  const int t = part * TPP + tl;
  const double c = (double)sm.cs[t], s = (double)sm.sn[t];
  unsigned int* row = sm.acc;
  const int row0 = tl * RS + ADD;
  for (int p = g; p < n_points; p += G) {
    if (!sm.valid[p]) continue;
    const double rho = sm.px[p] * c + sm.py[p] * s;
    const double tr = trunc(rho);
    const double fr = rho - tr;                      // exact
    int r = (int)tr + (fr >= 0.5 ? 1 : 0) - (fr <= -0.5 ? 1 : 0);
    r /= EKF_HOUGH_DISTANCE;
    if ((unsigned)(r + ADD) < (unsigned)RS) {
      const int b = row0 + r;
      atomicAdd(&row[b >> 2], 1u << ((b & 3) * 8));
    }
  }
}

__global__ void __launch_bounds__(kThreads) hough_scan_kernel(const HoughArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HoughSmem& sm = *reinterpret_cast<HoughSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_points = a.n_points;
  for (int i = tid; i < TH; i += kThreads) { sm.cs[i] = a.cos_tab[i]; sm.sn[i] = a.sin_tab[i]; }

  for (int scan = blockIdx.x; scan < a.n_scans; scan += gridDim.x) {
    const size_t pbase = (size_t)scan * n_points;
    for (int i = tid; i < n_points; i += kThreads) {
      sm.px[i] = a.x[pbase + i];
      sm.py[i] = a.y[pbase + i];
      sm.valid[i] = a.range[pbase + i] <= (unsigned int)EKF_HOUGH_MAX_DIST;   // houghtransform.cpp:245
    }
    zero_acc(sm, tid, kThreads);
    __syncthreads();
    vote(sm, 0, n_points, tid, kThreads);
    __syncthreads();

    // the 200 slots of houghtransform.cpp:46 (warp 0 only): slot s = k*32 + lane
    int pv[SLOTS_PER_LANE], pi[SLOTS_PER_LANE];
    int mindex = 0, minval = 0;

    for (int part = 0; part < PARTS; ++part) {
      // ---- compact the cells of this slice that can still enter the peak array, in cell order -----
      // Each thread owns a contiguous run of 16-byte groups; most groups are all zero and cost one
      // 128-bit load and a test. Counts are compared four at a time (per-byte SIMD compare).
      constexpr int GROUPS = ACC_WORDS_PAD / 4;
      constexpr int GPT = (GROUPS + kThreads - 1) / kThreads;
      const int g0 = tid * GPT, g1 = (g0 + GPT < GROUPS) ? g0 + GPT : GROUPS;
      // a cell can only ever enter the peak array if its count exceeds the current minimum slot
      const unsigned int thr = part == 0 ? (sm.acc[0] & 0xFFu) : (unsigned int)sm.minval;
      const unsigned int thr4 = thr * 0x01010101u;
      const uint4* acc4 = reinterpret_cast<const uint4*>(sm.acc);
      int cnt = 0;
      for (int g = g0; g < g1; ++g) {
        const uint4 v = acc4[g];
        if ((v.x | v.y | v.z | v.w) == 0u) continue;
        cnt += (__popc(__vcmpgtu4(v.x, thr4)) + __popc(__vcmpgtu4(v.y, thr4)) + __popc(__vcmpgtu4(v.z, thr4)) +
                __popc(__vcmpgtu4(v.w, thr4))) >> 3;
      }
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      if (lane == 31) sm.warp_sum[warp] = incl;
      if (part == 0 && tid == 0) sm.grid0 = (int)(sm.acc[0] & 0xFFu);
      __syncthreads();
      int off = incl - cnt;
      for (int w = 0; w < warp; ++w) off += sm.warp_sum[w];
      if (tid == kThreads - 1) sm.n_cand = off + cnt;
      const unsigned int cell0 = (unsigned int)(part * ACC_BYTES);
      if (cnt) {
        for (int g = g0; g < g1; ++g) {
          const uint4 v = acc4[g];
          if ((v.x | v.y | v.z | v.w) == 0u) continue;
          const unsigned int vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            unsigned int m = __vcmpgtu4(vv[j], thr4) & 0x01010101u;   // bit 8k set: byte k passes
            while (m) {
              const int k = (__ffs(m) - 1) >> 3;
              m &= m - 1u;
              sm.cand[off++] = ((cell0 + 16u * g + 4u * j + k) << 8) | ((vv[j] >> (8 * k)) & 0xFFu);
            }
          }
        }
      }
      // the per-word bounds the debug dump below uses
      constexpr int WPT = GPT * 4;
      const int w0 = g0 * 4, w1 = (g1 * 4 < ACC_WORDS) ? g1 * 4 : ACC_WORDS;
      (void)WPT;
      if (a.grid) {   // debug / parity: the accumulator itself
        unsigned char* g = a.grid + (size_t)scan * TH * RS + (size_t)part * ACC_BYTES;
        for (int w = w0; w < w1; ++w) {
          const unsigned int v = sm.acc[w];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (4 * w + k < ACC_BYTES) g[4 * w + k] = (unsigned char)(v >> (8 * k));
        }
      }
      __syncthreads();

      if (warp == 0) {
        // ---- houghtransform.cpp:260-280 over the compacted stream ----------------------------------
        if (part == 0) {
          const int first = sm.grid0;    // every slot starts at cell 0 (houghtransform.cpp:46)
#pragma unroll
          for (int k = 0; k < SLOTS_PER_LANE; ++k) { pv[k] = first; pi[k] = 0; }
          mindex = 0;
          minval = first;
        }
        const int n_cand = sm.n_cand;
        for (int base = 0; base < n_cand; base += 32) {
          const unsigned int c = base + lane < n_cand ? sm.cand[base + lane] : 0u;
          const int v = (int)(c & 0xFFu), cell = (int)(c >> 8);
          unsigned int pend = __ballot_sync(0xffffffffu, v > minval);
          while (pend) {
            const int src = __ffs(pend) - 1;
            const int v_s = __shfl_sync(0xffffffffu, v, src);
            const int cell_s = __shfl_sync(0xffffffffu, cell, src);
            const int owner = mindex & 31, kk = mindex >> 5;
            unsigned int lkey = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < SLOTS_PER_LANE; ++k) {
              if (lane == owner && k == kk) { pv[k] = v_s; pi[k] = cell_s; }
              const int slot = k * 32 + lane;
              if (slot < PK) {
                const unsigned int key = ((unsigned int)pv[k] << 8) | (unsigned int)slot;
                lkey = key < lkey ? key : lkey;
              }
            }
            const unsigned int gkey = __reduce_min_sync(0xffffffffu, lkey);
            const int gmin = (int)(gkey >> 8);
            if (gmin < v_s) { mindex = (int)(gkey & 0xFFu); minval = gmin; }   // first slot with a strictly smaller count
            else minval = v_s;                                                 // none: the slot just written stays the minimum
            pend = __ballot_sync(0xffffffffu, v > minval) & ~((2u << src) - 1u);
          }
        }
        if (lane == 0) sm.minval = minval;
      } else if (part + 1 < PARTS) {
        // ---- meanwhile: next half's votes ------------------------------------------------------------
        zero_acc(sm, tid - 32, kThreads - 32);
        named_barrier(1, kThreads - 32);
        vote(sm, part + 1, n_points, tid - 32, kThreads - 32);
      }
      __syncthreads();
    }
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < SLOTS_PER_LANE; ++k) {
        const int slot = k * 32 + lane;
        if (slot < PK) {
          a.peaks[(size_t)scan * PK + slot] = pi[k];
          a.values[(size_t)scan * PK + slot] = pv[k];
        }
      }
    }
  }
}

// ---- houghtransform.cpp:58-236, one scan (host and device) ---------------------------------------
struct Group {
  int hi_r, lo_r, hi_t, lo_t;
  int sum_r, sum_t, weight, count;
};

__host__ __device__ inline int iabs(int v) { return v < 0 ? -v : v; }
__host__ __device__ inline bool close_to(int hi, int lo, int v, int tol) {
  return iabs(hi - v) < tol || iabs(lo - v) < tol || (v < hi && v > lo);
}

__host__ __device__ int lines_from_peaks(const int32_t* peaks, const int32_t* values, ekf_hough_line* lines, int max_lines) {
  Group g[PK];
  int ng = 0;
  for (int p = 0; p < PK; ++p) {                       // greedy clustering in slot order (:66-112)
    const int r = peaks[p] % RS, t = peaks[p] / RS, w = values[p];
    if (r <= 0) continue;
    int j = 0;
    for (; j < ng; ++j)
      if (close_to(g[j].hi_t, g[j].lo_t, t, 30) && close_to(g[j].hi_r, g[j].lo_r, r, 5)) break;
    if (j == ng) {
      Group& q = g[ng++];
      q.hi_r = q.lo_r = r;
      q.hi_t = q.lo_t = t;
      q.sum_r = r * w;
      q.sum_t = t * w;
      q.weight = w;
      q.count = 1;
    } else {
      Group& q = g[j];
      if (r > q.hi_r) q.hi_r = r;
      if (r < q.lo_r) q.lo_r = r;
      if (t > q.hi_t) q.hi_t = t;
      if (t < q.lo_t) q.lo_t = t;
      q.sum_r += r * w;
      q.sum_t += t * w;
      q.weight += w;
      q.count += 1;
    }
  }
  for (int i = 0; i < ng; ++i) {                       // negative radii -> the opposite normal (:118-129)
    Group& q = g[i];
    if (q.sum_r < ADD * q.weight) {
      q.sum_r = 2 * ADD * q.weight - q.sum_r;
      q.hi_r = 2 * ADD - q.hi_r;
      q.lo_r = 2 * ADD - q.lo_r;
      q.sum_t -= TH * q.weight;
      q.hi_t -= TH;
      q.lo_t -= TH;
    }
  }
  int root_of[PK];                                     // :158-190 (the last matching earlier group wins)
  for (int j = 0; j < ng; ++j) {
    root_of[j] = -1;
    for (int i = 0; i < j; ++i) {
      const Group &u = g[i], &v = g[j];
      const bool t_ok = iabs(v.hi_t - u.lo_t) < 30 || iabs(v.lo_t - u.hi_t) < 30 || (u.hi_t > v.lo_t && u.lo_t < v.hi_t);
      const bool r_ok = iabs(v.hi_r - u.lo_r) < 5 || iabs(v.lo_r - u.hi_r) < 5 || (u.hi_r > v.lo_r && u.lo_r < v.hi_r);
      if (t_ok && r_ok) root_of[j] = i;
    }
  }
  for (int i = 0; i < ng; ++i) {                       // :194-211
    if (root_of[i] < 0) continue;
    int j = i;
    while (root_of[j] >= 0) j = root_of[j];
    Group& d = g[j];
    const Group& s = g[i];
    if (s.hi_r > d.hi_r) d.hi_r = s.hi_r;
    if (s.lo_r < d.lo_r) d.lo_r = s.lo_r;
    if (s.hi_t > d.hi_t) d.hi_t = s.hi_t;
    if (s.lo_t < d.lo_t) d.lo_t = s.lo_t;
    d.sum_r += s.sum_r;
    d.sum_t += s.sum_t;
    d.weight += s.weight;
    d.count += s.count;
  }
  int n = 0;
  for (int i = 0; i < ng; ++i) {                       // :215-233
    if (root_of[i] >= 0) continue;
    if (n < max_lines) {
      ekf_hough_line& L = lines[n];
      L.theta = g[i].sum_t / (double)g[i].weight;
      L.theta *= 3.141592654 / TH;
      L.radius = g[i].sum_r / (double)g[i].weight;
      L.radius -= ADD;
      L.radius *= EKF_HOUGH_DISTANCE;
      L.weight = g[i].weight / (double)g[i].count;
    }
    ++n;
  }
  return n;
}

// houghtransform.cpp:58-236 with one WARP per scan: the peaks are visited in slot order (the
// clustering is greedy, so that loop stays sequential), but every peak is tested against all
// existing groups at once (one group per lane, a ballot picks the first match, which is where the
// reference's inner loop breaks), and the group-against-group pass and the line conversion run one
// group per lane. Same integer arithmetic as lines_from_peaks above, which stays the host version.
constexpr int kLinesWarps = 4;
struct LinesSmem {
  int hi_r[PK], lo_r[PK], hi_t[PK], lo_t[PK], sum_r[PK], sum_t[PK], weight[PK], count[PK];
  int root_of[PK];
  int pr[PK], pt[PK], pw[PK];
};

__global__ void __launch_bounds__(kLinesWarps * 32) hough_lines_kernel(const int* __restrict__ peaks, const int* __restrict__ values,
                                                                      ekf_hough_line* __restrict__ lines, int* __restrict__ n_lines,
                                                                      int max_lines, int n_scans) {
  __shared__ LinesSmem smem[kLinesWarps];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int s = blockIdx.x * kLinesWarps + wib;
  if (s >= n_scans) return;
  LinesSmem& g = smem[wib];
  for (int p = lane; p < PK; p += 32) {
    const int cell = peaks[(size_t)s * PK + p];
    g.pr[p] = cell % RS;
    g.pt[p] = cell / RS;
    g.pw[p] = values[(size_t)s * PK + p];
  }
  __syncwarp();
  int ng = 0;
  for (int p = 0; p < PK; ++p) {                       // :66-112
    const int r = g.pr[p], t = g.pt[p], w = g.pw[p];
    if (r <= 0) continue;
    int found = -1;
    for (int base = 0; base < ng && found < 0; base += 32) {
      const int j = base + lane;
      bool m = false;
      if (j < ng) m = close_to(g.hi_t[j], g.lo_t[j], t, 30) && close_to(g.hi_r[j], g.lo_r[j], r, 5);
      const unsigned int hit = __ballot_sync(0xffffffffu, m);
      if (hit) found = base + __ffs(hit) - 1;
    }
    if (lane == 0) {
      if (found < 0) {
        g.hi_r[ng] = g.lo_r[ng] = r;
        g.hi_t[ng] = g.lo_t[ng] = t;
        g.sum_r[ng] = r * w;
        g.sum_t[ng] = t * w;
        g.weight[ng] = w;
        g.count[ng] = 1;
      } else {
        if (r > g.hi_r[found]) g.hi_r[found] = r;
        if (r < g.lo_r[found]) g.lo_r[found] = r;
        if (t > g.hi_t[found]) g.hi_t[found] = t;
        if (t < g.lo_t[found]) g.lo_t[found] = t;
        g.sum_r[found] += r * w;
        g.sum_t[found] += t * w;
        g.weight[found] += w;
        g.count[found] += 1;
      }
    }
    if (found < 0) ++ng;
    __syncwarp();
  }
  for (int i = lane; i < ng; i += 32) {                // :118-129
    if (g.sum_r[i] < ADD * g.weight[i]) {
      g.sum_r[i] = 2 * ADD * g.weight[i] - g.sum_r[i];
      g.hi_r[i] = 2 * ADD - g.hi_r[i];
      g.lo_r[i] = 2 * ADD - g.lo_r[i];
      g.sum_t[i] -= TH * g.weight[i];
      g.hi_t[i] -= TH;
      g.lo_t[i] -= TH;
    }
  }
  __syncwarp();
  for (int j = lane; j < ng; j += 32) {                // :158-190, the last matching earlier group wins
    int root = -1;
    const int vhi_t = g.hi_t[j], vlo_t = g.lo_t[j], vhi_r = g.hi_r[j], vlo_r = g.lo_r[j];
    for (int i = 0; i < j; ++i) {
      const int uhi_t = g.hi_t[i], ulo_t = g.lo_t[i], uhi_r = g.hi_r[i], ulo_r = g.lo_r[i];
      const bool t_ok = iabs(vhi_t - ulo_t) < 30 || iabs(vlo_t - uhi_t) < 30 || (uhi_t > vlo_t && ulo_t < vhi_t);
      const bool r_ok = iabs(vhi_r - ulo_r) < 5 || iabs(vlo_r - uhi_r) < 5 || (uhi_r > vlo_r && ulo_r < vhi_r);
      if (t_ok && r_ok) root = i;
    }
    g.root_of[j] = root;
  }
  __syncwarp();
  if (lane == 0) {                                     // :194-211
    for (int i = 0; i < ng; ++i) {
      if (g.root_of[i] < 0) continue;
      int j = i;
      while (g.root_of[j] >= 0) j = g.root_of[j];
      if (g.hi_r[i] > g.hi_r[j]) g.hi_r[j] = g.hi_r[i];
      if (g.lo_r[i] < g.lo_r[j]) g.lo_r[j] = g.lo_r[i];
      if (g.hi_t[i] > g.hi_t[j]) g.hi_t[j] = g.hi_t[i];
      if (g.lo_t[i] < g.lo_t[j]) g.lo_t[j] = g.lo_t[i];
      g.sum_r[j] += g.sum_r[i];
      g.sum_t[j] += g.sum_t[i];
      g.weight[j] += g.weight[i];
      g.count[j] += g.count[i];
    }
  }
  __syncwarp();
  int n = 0;
  ekf_hough_line* out = lines + (size_t)s * max_lines;
  for (int base = 0; base < ng; base += 32) {          // :215-233, roots in group order
    const int i = base + lane;
    const bool is_root = i < ng && g.root_of[i] < 0;
    const unsigned int roots = __ballot_sync(0xffffffffu, is_root);
    const int pos = n + __popc(roots & ((1u << lane) - 1u));
    if (is_root && pos < max_lines) {
      ekf_hough_line L;
      L.theta = g.sum_t[i] / (double)g.weight[i];
      L.theta *= 3.141592654 / TH;
      L.radius = g.sum_r[i] / (double)g.weight[i];
      L.radius -= ADD;
      L.radius *= EKF_HOUGH_DISTANCE;
      L.weight = g.weight[i] / (double)g.count[i];
      out[pos] = L;
    }
    n += __popc(roots);
  }
  if (lane == 0) n_lines[s] = n;
}

// ---- FeatureDetector stages behind getLines (featuredetector.cpp:74-362), one warp per scan ----------
// fitLineSegments: the closest line of every reading is found in parallel (one reading per lane);
// the segment lists are per line and only ever touched by readings of that line, so each lane then
// walks the readings of its own line(s) in scan order (the list logic is sequential by nature).
// extractCorners: for every segment i the pairs (i, j > i) go one per lane, a ballot keeps the
// features in the reference's (i, j) order. getStructCompass: a handful of lines, one lane.
// sin / cos of the line angles are the device's double routines rounded to float; they can differ
// from glibc's where the double result sits within an ulp of a float rounding boundary (~1e-8 per
// value) - everything else is the reference's arithmetic (float products for the determinant).
constexpr int kFeatWarps = 2;
constexpr int F_POINT_DIST = 600, F_MIN_POINTS = 3, F_CORNER_DIST = 90000, F_MIN_DIST = 1000 * 1000;
struct FeatSmem {
  double px[MAXP], py[MAXP];
  double rad[PK], th[PK];
  double s_sx[MAXP], s_sy[MAXP], s_ex[MAXP], s_ey[MAXP];
  float sn[PK], cs[PK];
  int head[PK];
  int good_off[PK];
  int pt_line[MAXP];
  int s_line[MAXP], s_np[MAXP], s_next[MAXP];
  int order[MAXP];
  int pool_count;
};

struct FeatArgs {
  const double* x;
  const double* y;
  const unsigned int* range;
  const ekf_hough_line* lines;   // [n_scans][lines_stride]
  const int* n_lines;
  int lines_stride;
  int n_scans, n_points;
  const double* cur_phi;         // [n_scans] or null
  double* offset;                // [n_scans] in/out or null
  ekf_feature* feats;            // [n_scans][max_feats]
  int* n_feats;
  int max_feats;
  double* compass;               // [n_scans] or null
  double* segs;                  // [n_scans][max_segs][7] or null
  int* n_segs;
  int max_segs;
};

__global__ void __launch_bounds__(kFeatWarps * 32) hough_features_kernel(const FeatArgs a) {
  __shared__ FeatSmem smem[kFeatWarps];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int s = blockIdx.x * kFeatWarps + wib;
  if (s >= a.n_scans) return;
  FeatSmem& g = smem[wib];
  const int n_points = a.n_points;
  int nl = a.n_lines[s];
  if (nl > a.lines_stride) nl = a.lines_stride;
  const ekf_hough_line* L = a.lines + (size_t)s * a.lines_stride;
  for (int l = lane; l < nl; l += 32) {
    g.rad[l] = L[l].radius;
    g.th[l] = L[l].theta;
    g.sn[l] = (float)sin(L[l].theta);            // featuredetector.cpp:88-89
    g.cs[l] = (float)cos(L[l].theta);
    g.head[l] = -1;
  }
  if (lane == 0) g.pool_count = 0;
  __syncwarp();
  // ---- closest line of every reading (:97-118) -----------------------------------------------------
  for (int p = lane; p < n_points; p += 32) {
    const double px = a.x[(size_t)s * n_points + p], py = a.y[(size_t)s * n_points + p];
    g.px[p] = px;
    g.py[p] = py;
    int line = -1;
    if (a.range[(size_t)s * n_points + p] <= (unsigned int)EKF_HOUGH_MAX_DIST) {
      double best = 1000000.0;
      int arg = 0;
      for (int l = 0; l < nl; ++l) {
        const double r = px * (double)g.cs[l] + py * (double)g.sn[l];
        const double d = fabs(g.rad[l] - r);
        if (d < best) { best = d; arg = l; }
      }
      if (!(best > (double)F_POINT_DIST)) line = arg;
    }
    g.pt_line[p] = line;
  }
  __syncwarp();
  // ---- segment lists, one line per lane, readings in scan order (:120-201) -------------------------
  for (int line = lane; line < nl; line += 32) {
    const bool by_x = fabsf(g.sn[line]) > fabsf(g.cs[line]);
    for (int p = 0; p < n_points; ++p) {
      if (g.pt_line[p] != line) continue;
      const double px = g.px[p], py = g.py[p];
      const double v = by_x ? px : py;
      int q = g.head[line];
      while (q >= 0) {
        const double sv = by_x ? g.s_sx[q] : g.s_sy[q], ev = by_x ? g.s_ex[q] : g.s_ey[q];
        if (v <= sv && v >= ev) { g.s_np[q]++; break; }
        if (v > sv && fabs(v - sv) <= (double)F_POINT_DIST) { g.s_sx[q] = px; g.s_sy[q] = py; g.s_np[q]++; break; }
        if (v < ev && fabs(v - ev) <= (double)F_POINT_DIST) { g.s_ex[q] = px; g.s_ey[q] = py; g.s_np[q]++; break; }
        q = g.s_next[q];
      }
      if (q < 0) {
        q = atomicAdd(&g.pool_count, 1);
        g.s_line[q] = line;
        g.s_np[q] = 1;
        g.s_sx[q] = g.s_ex[q] = px;
        g.s_sy[q] = g.s_ey[q] = py;
        g.s_next[q] = g.head[line];
        g.head[line] = q;
      }
    }
    int good = 0;                                  // :205-221 keeps segments with more than MIN_POINTS readings
    for (int q = g.head[line]; q >= 0; q = g.s_next[q]) good += g.s_np[q] > F_MIN_POINTS;
    g.good_off[line] = good;
  }
  __syncwarp();
  int n_segs = 0;
  if (lane == 0) {                                 // lines in order, each list newest first
    for (int l = 0; l < nl; ++l) { const int c = g.good_off[l]; g.good_off[l] = n_segs; n_segs += c; }
  }
  n_segs = __shfl_sync(0xffffffffu, n_segs, 0);
  __syncwarp();
  for (int line = lane; line < nl; line += 32) {
    int o = g.good_off[line];
    for (int q = g.head[line]; q >= 0; q = g.s_next[q])
      if (g.s_np[q] > F_MIN_POINTS) g.order[o++] = q;
  }
  __syncwarp();
  if (a.segs) {
    double* so = a.segs + (size_t)s * a.max_segs * 7;
    for (int k = lane; k < n_segs && k < a.max_segs; k += 32) {
      const int q = g.order[k], l = g.s_line[q];
      so[7 * k + 0] = g.rad[l]; so[7 * k + 1] = g.th[l]; so[7 * k + 2] = g.s_sx[q]; so[7 * k + 3] = g.s_sy[q];
      so[7 * k + 4] = g.s_ex[q]; so[7 * k + 5] = g.s_ey[q]; so[7 * k + 6] = (double)g.s_np[q];
    }
  }
  if (a.n_segs && lane == 0) a.n_segs[s] = n_segs;
  // ---- corners (:230-292): pairs (i, j > i) in order ---------------------------------------------------
  const double corner_theta = 22.0 * 3.141592654 / 180.0;
  ekf_feature* fo = a.feats + (size_t)s * a.max_feats;
  int n_feats = 0;
  for (int i = 0; i < n_segs; ++i) {
    const int qi = g.order[i], li = g.s_line[qi];
    const double r1 = g.rad[li], t1 = g.th[li];
    const float sni = g.sn[li], csi = g.cs[li];
    for (int base = i + 1; base < n_segs; base += 32) {
      const int j = base + lane;
      bool hit = false;
      double cx = 0.0, cy = 0.0;
      if (j < n_segs) {
        const int qj = g.order[j], lj = g.s_line[qj];
        double dth = fabs(t1 - g.th[lj]);
        if (dth > 3.141592654) dth = fabs(dth - 6.283185307);
        if (dth > 1.570796327) dth = fabs(dth - 3.141592654);
        if (!(dth < corner_theta)) {
          const float snj = g.sn[lj], csj = g.cs[lj];
          const double det = (double)__fsub_rn(__fmul_rn(csi, snj), __fmul_rn(sni, csj));   // float arithmetic (float arrays)
          const double r2 = g.rad[lj];
          cx = (r1 * (double)snj - r2 * (double)sni) / det;
          cy = (r2 * (double)csi - r1 * (double)csj) / det;
          double dx, dy;
          dx = g.s_sx[qi] - cx; dy = g.s_sy[qi] - cy; const bool start1 = (dx * dx + dy * dy) < (double)F_CORNER_DIST;
          dx = g.s_ex[qi] - cx; dy = g.s_ey[qi] - cy; const bool end1 = (dx * dx + dy * dy) < (double)F_CORNER_DIST;
          dx = g.s_sx[qj] - cx; dy = g.s_sy[qj] - cy; const bool start2 = (dx * dx + dy * dy) < (double)F_CORNER_DIST;
          dx = g.s_ex[qj] - cx; dy = g.s_ey[qj] - cy; const bool end2 = (dx * dx + dy * dy) < (double)F_CORNER_DIST;
          hit = (start1 || end1) && (start2 || end2) && (cx * cx + cy * cy) > (double)F_MIN_DIST;
        }
      }
      const unsigned int hits = __ballot_sync(0xffffffffu, hit);
      const int pos = n_feats + __popc(hits & ((1u << lane) - 1u));
      if (hit && pos < a.max_feats) { fo[pos].x = cx; fo[pos].y = cy; }
      n_feats += __popc(hits);
    }
  }
  if (lane == 0) a.n_feats[s] = n_feats;
  // ---- structural compass (:297-362) ------------------------------------------------------------------
  if (a.compass && lane == 0) {
    // groups reuse the segment scratch (no longer needed): s_sx = weighted angle sums, s_sy = weights
    double* g_theta = g.s_sx;
    double* g_weight = g.s_sy;
    int ng = 0;
    const double thresh = 10 * 3.141592654 / 180.0;
    for (int i = 0; i < nl; ++i) {
      const double th = g.th[i] - 1.570796327 * floor(g.th[i] / 1.570796327);
      const double w = L[i].weight;
      bool merged = false;
      for (int j = 0; j < ng; ++j) {                 // no break: a line joins every group it is close to
        const double mean = g_theta[j] / g_weight[j];
        if (fabs(th - mean) < thresh) { g_theta[j] += th * w; g_weight[j] += w; merged = true; }
      }
      if (!merged && ng < MAXP) { g_theta[ng] = th * w; g_weight[ng] = w; ++ng; }
    }
    double best_theta = 0.0, best_w = 0.0;
    for (int j = 0; j < ng; ++j)
      if (g_weight[j] > best_w) { best_theta = g_theta[j]; best_w = g_weight[j]; }
    double result = 100.0;                           // NO_COMPASS
    if (best_w != 0.0) {
      double cardinal = -(best_theta / best_w);
      double off = a.offset ? a.offset[s] : 100.0;
      if (off == 100.0) off = cardinal;
      if (a.offset) a.offset[s] = off;
      cardinal -= off;
      cardinal -= 1.570796327 * floor(cardinal / 1.570796327);
      double phi = a.cur_phi ? a.cur_phi[s] : 0.0;
      phi -= 6.283185307 * floor(phi / 6.283185307);
      // six candidate errors (:346-351); the answer is the first candidate that is no worse than every
      // later one (:356-361), candidates 5 and 6 being the roll-overs of 1 and 4
      const double shift[5] = {0.0, 1.570796327, 3.141592654, 4.71238898, 6.283185307};
      const double add[6] = {0.0, 1.570796327, 3.141592654, 4.71238898, 0.0, 4.71238898};
      double err[6];
      for (int k = 0; k < 5; ++k) err[k] = fabs((phi - cardinal) - shift[k]);
      err[5] = fabs((phi - cardinal) + 1.570796327);
      for (int k = 0; k < 6; ++k) {
        bool best = true;
        for (int j = k + 1; j < 6; ++j) best = best && err[k] <= err[j];
        if (best) { result = add[k] == 0.0 ? cardinal : cardinal + add[k]; break; }
      }
    }
    a.compass[s] = result;
  }
}

std::string g_hough_create_error;
constexpr int kHoughChunks = 16;

}  // namespace

struct ekf_hough_s {
  int device = 0, sm_count = 0, max_scans = 0, ctas_per_sm = 1;
  cudaStream_t stream = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;          // copy streams of the pipelined end-to-end path
  cudaEvent_t ev_in[kHoughChunks], ev_k[kHoughChunks], ev_done = nullptr;
  double* d_x = nullptr;
  double* d_y = nullptr;
  unsigned int* d_range = nullptr;
  float* d_cos = nullptr;
  float* d_sin = nullptr;
  int* d_peaks = nullptr;
  int* d_values = nullptr;
  unsigned char* d_grid = nullptr;
  size_t grid_cap = 0;
  ekf_hough_line* d_lines = nullptr;
  int* d_nlines = nullptr;
  // feature stages (allocated on first use, grown on demand)
  ekf_feature* d_feats = nullptr;
  size_t feats_cap = 0;
  double* d_segs = nullptr;
  size_t segs_cap = 0;
  int* d_nfeats = nullptr;
  int* d_nsegs = nullptr;
  double* d_phi = nullptr;
  double* d_off = nullptr;
  double* d_compass = nullptr;
  int lines_cap = 0;          // lines per scan the device buffer holds
  int run_max_lines = 0;      // of the last run
  int n_scans = 0, n_points = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float kernel_ms = 0.f;
  int kernel_launches = 0;
  bool pending_event = false;
  std::string err;
};

namespace {

int hfail(ekf_hough h, int code, const std::string& msg) {
  if (h) h->err = msg;
  else g_hough_create_error = msg;
  return code;
}

#define HG_CK(h, call)                                                                           \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return hfail(h, EKF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));        \
  } while (0)

int collect_kernel_time(ekf_hough h) {
  if (h->pending_event) {
    float ms = 0.f;
    HG_CK(h, cudaEventSynchronize(h->ev1));
    HG_CK(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->kernel_ms += ms;
    h->kernel_launches += 1;
    h->pending_event = false;
  }
  return EKF_OK;
}

int launch(ekf_hough h, bool want_grid, int max_lines) {
  if (h->n_scans < 1) return hfail(h, EKF_ERR_BAD_ARG, "no scans uploaded");
  if (max_lines < 1) max_lines = 1;
  if (max_lines > PK) max_lines = PK;
  if (max_lines > h->lines_cap) {
    cudaFree(h->d_lines);
    h->d_lines = nullptr;
    h->lines_cap = 0;
    HG_CK(h, cudaMalloc(&h->d_lines, (size_t)h->max_scans * max_lines * sizeof(ekf_hough_line)));
    h->lines_cap = max_lines;
  }
  h->run_max_lines = max_lines;
  int rc = collect_kernel_time(h);
  if (rc != EKF_OK) return rc;
  HoughArgs a;
  a.x = h->d_x; a.y = h->d_y; a.range = h->d_range; a.cos_tab = h->d_cos; a.sin_tab = h->d_sin;
  a.n_scans = h->n_scans; a.n_points = h->n_points;
  a.peaks = h->d_peaks; a.values = h->d_values;
  a.grid = nullptr;
  if (want_grid) {
    const size_t need = (size_t)h->n_scans * TH * RS;
    if (need > h->grid_cap) {
      cudaFree(h->d_grid);
      h->d_grid = nullptr;
      h->grid_cap = 0;
      HG_CK(h, cudaMalloc(&h->d_grid, need));
      h->grid_cap = need;
    }
    a.grid = h->d_grid;
  }
  const int cap = h->ctas_per_sm * h->sm_count;                            // persistent grid: all CTAs co-resident
  const int grid = h->n_scans < cap ? h->n_scans : cap;
  HG_CK(h, cudaEventRecord(h->ev0, h->stream));
  hough_scan_kernel<<<grid, kThreads, sizeof(HoughSmem), h->stream>>>(a);
  hough_lines_kernel<<<(h->n_scans + kLinesWarps - 1) / kLinesWarps, kLinesWarps * 32, 0, h->stream>>>(
      h->d_peaks, h->d_values, h->d_lines, h->d_nlines, max_lines, h->n_scans);
  HG_CK(h, cudaGetLastError());
  HG_CK(h, cudaEventRecord(h->ev1, h->stream));
  h->pending_event = true;
  return EKF_OK;
}

}  // namespace

extern "C" {

void ekf_hough_tables(float* cos_out, float* sin_out) {
  // houghtransform.cpp:5-18: the angle accumulates in float; cos / sin are the double functions
  const float step = (float)(3.141592654 / TH);
  float theta = 0.0f;
  for (int i = 0; i < TH; ++i) {
    cos_out[i] = (float)std::cos((double)theta);
    sin_out[i] = (float)std::sin((double)theta);
    theta += step;
  }
}

int ekf_hough_lines_from_peaks(const int32_t* peaks, const int32_t* values, ekf_hough_line* lines, int max_lines) {
  if (!peaks || !values || (!lines && max_lines > 0)) return 0;
  return lines_from_peaks(peaks, values, lines, max_lines);
}

int ekf_hough_create(ekf_hough* out, int device, int max_scans) {
  if (!out) return hfail(nullptr, EKF_ERR_BAD_ARG, "out is NULL");
  *out = nullptr;
  if (max_scans < 1) return hfail(nullptr, EKF_ERR_BAD_ARG, "max_scans must be >= 1");
  int sms = 0, maj = 0, mnr = 0;
  size_t smem_optin = 0;
  if (ekf_device_info(device, &sms, &maj, &mnr, &smem_optin, nullptr) != EKF_OK)
    return hfail(nullptr, EKF_ERR_NO_DEVICE, "no usable CUDA device " + std::to_string(device) + " (this library has no CPU fallback)");
  if (maj != 10) return hfail(nullptr, EKF_ERR_NO_DEVICE, "device is sm_" + std::to_string(maj * 10 + mnr) + "; this library is built for sm_100a (B200) only");
  if (smem_optin < sizeof(HoughSmem)) return hfail(nullptr, EKF_ERR_UNSUPPORTED, "not enough shared memory per block");
  ekf_hough h = new ekf_hough_s();
  h->device = device;
  h->sm_count = sms;
  h->max_scans = max_scans;
  auto bail = [&](int code, const std::string& msg) {
    g_hough_create_error = msg;
    ekf_hough_destroy(h);
    return code;
  };
  cudaError_t e;
  if (cudaSetDevice(device) != cudaSuccess) return bail(EKF_ERR_CUDA, "cudaSetDevice failed");
  if ((e = cudaFuncSetAttribute(hough_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HoughSmem))) != cudaSuccess)
    return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  cudaFuncSetAttribute(hough_scan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->ctas_per_sm, hough_scan_kernel, kThreads, sizeof(HoughSmem))) != cudaSuccess)
    return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  if (h->ctas_per_sm < 1) h->ctas_per_sm = 1;
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  const size_t np = (size_t)max_scans * MAXP;
  if ((e = cudaMalloc(&h->d_x, np * sizeof(double))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMalloc(&h->d_y, np * sizeof(double))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMalloc(&h->d_range, np * sizeof(unsigned int))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMalloc(&h->d_peaks, (size_t)max_scans * PK * sizeof(int))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMalloc(&h->d_values, (size_t)max_scans * PK * sizeof(int))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMalloc(&h->d_nlines, (size_t)max_scans * sizeof(int))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMalloc(&h->d_cos, TH * sizeof(float))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMalloc(&h->d_sin, TH * sizeof(float))) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
  float c[TH], s[TH];
  ekf_hough_tables(c, s);
  cudaMemcpy(h->d_cos, c, sizeof(c), cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_sin, s, sizeof(s), cudaMemcpyHostToDevice);
  cudaEventCreate(&h->ev0);
  cudaEventCreate(&h->ev1);
  cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming);
  for (int i = 0; i < kHoughChunks; ++i) {
    cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming);
  }
  *out = h;
  return EKF_OK;
}

int ekf_hough_destroy(ekf_hough h) {
  if (!h) return EKF_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_x); cudaFree(h->d_y); cudaFree(h->d_range); cudaFree(h->d_peaks); cudaFree(h->d_values);
  cudaFree(h->d_cos); cudaFree(h->d_sin); cudaFree(h->d_grid); cudaFree(h->d_lines); cudaFree(h->d_nlines);
  cudaFree(h->d_feats); cudaFree(h->d_segs); cudaFree(h->d_nfeats); cudaFree(h->d_nsegs); cudaFree(h->d_phi); cudaFree(h->d_off);
  cudaFree(h->d_compass);
  if (h->ev0) { cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); }
  if (h->ev_done) {
    cudaEventDestroy(h->ev_done);
    for (int i = 0; i < kHoughChunks; ++i) { cudaEventDestroy(h->ev_in[i]); cudaEventDestroy(h->ev_k[i]); }
    cudaStreamDestroy(h->s_in);
    cudaStreamDestroy(h->s_out);
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return EKF_OK;
}

int ekf_hough_upload(ekf_hough h, int n_scans, int n_points, const double* x, const double* y, const uint32_t* range) {
  if (!h || !x || !y || !range || n_scans < 1 || n_scans > h->max_scans || n_points < 1 || n_points > MAXP)
    return hfail(h, EKF_ERR_BAD_ARG, "ekf_hough_upload: 1 <= n_scans <= max_scans and 1 <= n_points <= " + std::to_string(MAXP) + " required");
  cudaSetDevice(h->device);
  const size_t np = (size_t)n_scans * n_points;
  HG_CK(h, cudaMemcpyAsync(h->d_x, x, np * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  HG_CK(h, cudaMemcpyAsync(h->d_y, y, np * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  HG_CK(h, cudaMemcpyAsync(h->d_range, range, np * sizeof(unsigned int), cudaMemcpyHostToDevice, h->stream));
  h->n_scans = n_scans;
  h->n_points = n_points;
  return EKF_OK;
}

int ekf_hough_run_resident(ekf_hough h, int max_lines) {
  if (!h) return EKF_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  return launch(h, false, max_lines);
}

int ekf_hough_download(ekf_hough h, ekf_hough_line* lines, int max_lines, int32_t* n_lines, int32_t* peaks, int32_t* values) {
  if (!h || max_lines < 0 || (max_lines > 0 && !lines)) return EKF_ERR_BAD_ARG;
  if (h->run_max_lines < 1) return hfail(h, EKF_ERR_BAD_ARG, "ekf_hough_download: nothing has been run");
  cudaSetDevice(h->device);
  const size_t S = (size_t)h->n_scans, cnt = S * PK;
  const int dev_lines = h->run_max_lines, take = max_lines < dev_lines ? max_lines : dev_lines;
  if (lines && take > 0)
    HG_CK(h, cudaMemcpy2DAsync(lines, (size_t)max_lines * sizeof(ekf_hough_line), h->d_lines, (size_t)dev_lines * sizeof(ekf_hough_line),
                               (size_t)take * sizeof(ekf_hough_line), S, cudaMemcpyDeviceToHost, h->stream));
  if (n_lines) HG_CK(h, cudaMemcpyAsync(n_lines, h->d_nlines, S * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (peaks) HG_CK(h, cudaMemcpyAsync(peaks, h->d_peaks, cnt * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (values) HG_CK(h, cudaMemcpyAsync(values, h->d_values, cnt * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  HG_CK(h, cudaStreamSynchronize(h->stream));
  return EKF_OK;
}

// End-to-end path without the accumulator dump: chunks of scans go through copy-in, the two
// kernels and copy-out on three streams, so the transfers of neighbouring chunks overlap the
// kernels (scans are independent; a chunk is a multiple of the co-resident CTA count). With
// pageable host buffers the copies serialise in the driver; ekf_host_alloc() gives pinned ones.
static int get_lines_pipelined(ekf_hough h, int n_scans, int n_points, const double* x, const double* y,
                               const uint32_t* range, ekf_hough_line* lines, int max_lines, int32_t* n_lines,
                               int32_t* peaks, int32_t* values) {
  if (!x || !y || !range || n_scans < 1 || n_scans > h->max_scans || n_points < 1 || n_points > MAXP)
    return hfail(h, EKF_ERR_BAD_ARG, "ekf_hough_get_lines: 1 <= n_scans <= max_scans and 1 <= n_points <= " + std::to_string(MAXP) + " required");
  if (max_lines < 0 || (max_lines > 0 && !lines)) return EKF_ERR_BAD_ARG;
  int dev_lines = max_lines < 1 ? 1 : (max_lines > PK ? PK : max_lines);
  if (dev_lines > h->lines_cap) {
    cudaFree(h->d_lines);
    h->d_lines = nullptr;
    h->lines_cap = 0;
    HG_CK(h, cudaMalloc(&h->d_lines, (size_t)h->max_scans * dev_lines * sizeof(ekf_hough_line)));
    h->lines_cap = dev_lines;
  }
  h->run_max_lines = dev_lines;
  h->n_scans = n_scans;
  h->n_points = n_points;
  const int wave = h->ctas_per_sm * h->sm_count;
  int chunk = 2 * wave;
  while ((n_scans + chunk - 1) / chunk > kHoughChunks) chunk += wave;
  const int n_chunks = (n_scans + chunk - 1) / chunk;
  const int take = max_lines < dev_lines ? max_lines : dev_lines;
  HG_CK(h, cudaEventRecord(h->ev_done, h->stream));
  HG_CK(h, cudaStreamWaitEvent(h->s_in, h->ev_done, 0));
  HG_CK(h, cudaStreamWaitEvent(h->s_out, h->ev_done, 0));
  for (int c = 0; c < n_chunks; ++c) {
    const size_t s0 = (size_t)c * chunk;
    const int ns = (int)((s0 + chunk <= (size_t)n_scans) ? chunk : n_scans - s0);
    const size_t p0 = s0 * n_points, np = (size_t)ns * n_points;
    HG_CK(h, cudaMemcpyAsync(h->d_x + p0, x + p0, np * sizeof(double), cudaMemcpyHostToDevice, h->s_in));
    HG_CK(h, cudaMemcpyAsync(h->d_y + p0, y + p0, np * sizeof(double), cudaMemcpyHostToDevice, h->s_in));
    HG_CK(h, cudaMemcpyAsync(h->d_range + p0, range + p0, np * sizeof(unsigned int), cudaMemcpyHostToDevice, h->s_in));
    HG_CK(h, cudaEventRecord(h->ev_in[c], h->s_in));
    HG_CK(h, cudaStreamWaitEvent(h->stream, h->ev_in[c], 0));
    HoughArgs a;
    a.x = h->d_x + p0; a.y = h->d_y + p0; a.range = h->d_range + p0; a.cos_tab = h->d_cos; a.sin_tab = h->d_sin;
    a.n_scans = ns; a.n_points = n_points;
    a.peaks = h->d_peaks + s0 * PK; a.values = h->d_values + s0 * PK;
    a.grid = nullptr;
    const int grid = ns < wave ? ns : wave;
    hough_scan_kernel<<<grid, kThreads, sizeof(HoughSmem), h->stream>>>(a);
    hough_lines_kernel<<<(ns + kLinesWarps - 1) / kLinesWarps, kLinesWarps * 32, 0, h->stream>>>(
        a.peaks, a.values, h->d_lines + s0 * dev_lines, h->d_nlines + s0, dev_lines, ns);
    HG_CK(h, cudaGetLastError());
    HG_CK(h, cudaEventRecord(h->ev_k[c], h->stream));
    HG_CK(h, cudaStreamWaitEvent(h->s_out, h->ev_k[c], 0));
    if (lines && take > 0)
      HG_CK(h, cudaMemcpy2DAsync(lines + s0 * max_lines, (size_t)max_lines * sizeof(ekf_hough_line), h->d_lines + s0 * dev_lines,
                                 (size_t)dev_lines * sizeof(ekf_hough_line), (size_t)take * sizeof(ekf_hough_line), ns,
                                 cudaMemcpyDeviceToHost, h->s_out));
    if (n_lines) HG_CK(h, cudaMemcpyAsync(n_lines + s0, h->d_nlines + s0, (size_t)ns * sizeof(int), cudaMemcpyDeviceToHost, h->s_out));
    if (peaks) HG_CK(h, cudaMemcpyAsync(peaks + s0 * PK, h->d_peaks + s0 * PK, (size_t)ns * PK * sizeof(int), cudaMemcpyDeviceToHost, h->s_out));
    if (values) HG_CK(h, cudaMemcpyAsync(values + s0 * PK, h->d_values + s0 * PK, (size_t)ns * PK * sizeof(int), cudaMemcpyDeviceToHost, h->s_out));
  }
  HG_CK(h, cudaStreamSynchronize(h->s_out));
  HG_CK(h, cudaStreamSynchronize(h->stream));
  return EKF_OK;
}

int ekf_hough_get_lines(ekf_hough h, int n_scans, int n_points, const double* x, const double* y, const uint32_t* range,
                        ekf_hough_line* lines, int max_lines, int32_t* n_lines, int32_t* peaks, int32_t* values,
                        uint8_t* grid) {
  if (!h) return EKF_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  if (!grid) return get_lines_pipelined(h, n_scans, n_points, x, y, range, lines, max_lines, n_lines, peaks, values);
  int rc = ekf_hough_upload(h, n_scans, n_points, x, y, range);
  if (rc != EKF_OK) return rc;
  rc = launch(h, grid != nullptr, max_lines);
  if (rc != EKF_OK) return rc;
  if (grid) HG_CK(h, cudaMemcpyAsync(grid, h->d_grid, (size_t)n_scans * TH * RS, cudaMemcpyDeviceToHost, h->stream));
  return ekf_hough_download(h, lines, max_lines, n_lines, peaks, values);
}

int ekf_hough_get_features(ekf_hough h, int n_scans, int n_points, const double* x, const double* y, const uint32_t* range,
                           const double* cur_phi, double* compass_offset, ekf_feature* feats, int max_feats, int32_t* n_feats,
                           double* compass, ekf_hough_line* lines, int max_lines, int32_t* n_lines, double* segments,
                           int max_segs, int32_t* n_segs) {
  if (!h || !feats || !n_feats || max_feats < 1 || (segments && max_segs < 1) || (lines && max_lines < 1))
    return hfail(h, EKF_ERR_BAD_ARG, "ekf_hough_get_features: feats, n_feats and max_feats >= 1 required");
  cudaSetDevice(h->device);
  int rc = ekf_hough_upload(h, n_scans, n_points, x, y, range);
  if (rc != EKF_OK) return rc;
  rc = launch(h, false, PK);                       // every line is kept: the segment stage needs them all
  if (rc != EKF_OK) return rc;
  const size_t S = (size_t)n_scans, MS = (size_t)h->max_scans;
#define HF_CK(call) HG_CK(h, call)
  if (!h->d_nfeats) {
    HF_CK(cudaMalloc(&h->d_nfeats, MS * sizeof(int)));
    HF_CK(cudaMalloc(&h->d_nsegs, MS * sizeof(int)));
    HF_CK(cudaMalloc(&h->d_phi, MS * sizeof(double)));
    HF_CK(cudaMalloc(&h->d_off, MS * sizeof(double)));
    HF_CK(cudaMalloc(&h->d_compass, MS * sizeof(double)));
  }
  if (MS * max_feats > h->feats_cap) {
    cudaFree(h->d_feats);
    h->d_feats = nullptr;
    h->feats_cap = 0;
    HF_CK(cudaMalloc(&h->d_feats, MS * max_feats * sizeof(ekf_feature)));
    h->feats_cap = MS * max_feats;
  }
  if (segments && MS * max_segs * 7 > h->segs_cap) {
    cudaFree(h->d_segs);
    h->d_segs = nullptr;
    h->segs_cap = 0;
    HF_CK(cudaMalloc(&h->d_segs, MS * max_segs * 7 * sizeof(double)));
    h->segs_cap = MS * max_segs * 7;
  }
  ekf_feature* d_feats = h->d_feats;
  int* d_nfeats = h->d_nfeats;
  int* d_nsegs = h->d_nsegs;
  double* d_phi = compass ? h->d_phi : nullptr;
  double* d_off = compass ? h->d_off : nullptr;
  double* d_compass = compass ? h->d_compass : nullptr;
  double* d_segs = segments ? h->d_segs : nullptr;
  if (compass) {
    if (cur_phi) HF_CK(cudaMemcpyAsync(d_phi, cur_phi, S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    else HF_CK(cudaMemsetAsync(d_phi, 0, S * sizeof(double), h->stream));
    if (compass_offset) HF_CK(cudaMemcpyAsync(d_off, compass_offset, S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    else {
      std::vector<double> unset(S, 100.0);
      HF_CK(cudaMemcpyAsync(d_off, unset.data(), S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      HF_CK(cudaStreamSynchronize(h->stream));
    }
  }
  FeatArgs a;
  a.x = h->d_x; a.y = h->d_y; a.range = h->d_range;
  a.lines = h->d_lines; a.n_lines = h->d_nlines; a.lines_stride = h->run_max_lines;
  a.n_scans = n_scans; a.n_points = n_points;
  a.cur_phi = d_phi; a.offset = d_off;
  a.feats = d_feats; a.n_feats = d_nfeats; a.max_feats = max_feats;
  a.compass = d_compass;
  a.segs = d_segs; a.n_segs = d_nsegs; a.max_segs = max_segs;
  hough_features_kernel<<<(n_scans + kFeatWarps - 1) / kFeatWarps, kFeatWarps * 32, 0, h->stream>>>(a);
  HF_CK(cudaGetLastError());
  HF_CK(cudaMemcpyAsync(feats, d_feats, S * max_feats * sizeof(ekf_feature), cudaMemcpyDeviceToHost, h->stream));
  HF_CK(cudaMemcpyAsync(n_feats, d_nfeats, S * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (n_segs) HF_CK(cudaMemcpyAsync(n_segs, d_nsegs, S * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (segments) HF_CK(cudaMemcpyAsync(segments, d_segs, S * max_segs * 7 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (compass) {
    HF_CK(cudaMemcpyAsync(compass, d_compass, S * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (compass_offset) HF_CK(cudaMemcpyAsync(compass_offset, d_off, S * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  HF_CK(cudaStreamSynchronize(h->stream));
#undef HF_CK
  if (lines || n_lines) return ekf_hough_download(h, lines, lines ? max_lines : 0, n_lines, nullptr, nullptr);
  return EKF_OK;
}

int ekf_hough_kernel_time(ekf_hough h, float* total_ms, int* n_launches) {
  if (!h || !total_ms || !n_launches) return EKF_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  int rc = collect_kernel_time(h);
  if (rc != EKF_OK) return rc;
  *total_ms = h->kernel_ms;
  *n_launches = h->kernel_launches;
  h->kernel_ms = 0.f;
  h->kernel_launches = 0;
  return EKF_OK;
}

int ekf_hough_sync(ekf_hough h) {
  if (!h) return EKF_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  HG_CK(h, cudaStreamSynchronize(h->stream));
  return EKF_OK;
}

const char* ekf_hough_last_error(ekf_hough h) { return h ? h->err.c_str() : g_hough_create_error.c_str(); }

}  // extern "C"
