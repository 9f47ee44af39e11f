// ekf_large.cu — regime B: one large map, covariance in HBM, whole grid per filter (sm_100a).
//
// Each reference call becomes a short chain of stream-ordered kernels whose control flow is
// decided ON THE DEVICE (no host round trip between gating and the covariance update):
//   doPropagation (kalmanfilter.cpp:15-48, Propagate.cpp:15-75)
//       large_prop_setup   1 thread : odometry -> Q, Phi, G; x update; 3x3 robot block
//       large_prop_strip   grid     : P_RL <- Phi*P_RL and its mirror, O(n)
//   doUpdate (Update.cpp:80-195), per measurement
//       large_gate         grid     : H, S, cond, Mahalanobis per landmark + per-CTA argmin
//       large_decide       1 CTA    : global argmin (lowest index wins ties), New/Old/Ignore,
//                                     S^-1, L D L^T of S, or the new landmark's blocks
//       large_gain         grid     : Old: gain rows, state correction, downdate vectors W
//                                     New: two new columns + mirror rows, O(n)
//       large_downdate     grid     : Old: P_ij += u_i . W_j over the dense n x n covariance —
//                                     the HBM-bound kernel (one read + one write of P)
//   doUpdateCompass (kalmanfilter.cpp:96-130): large_compass_setup, large_compass_gain,
//       large_downdate<1>.
// The arithmetic is the same code the batch regime uses (ekf_small.cuh), so decisions follow
// the reference's operation order; the downdate uses the bit-symmetric two-fma form described
// in ekf_cta.cuh.
//
// Look-ahead (the LA = true instantiations, used by ekf_large_run): of everything the dense sweep
// of operation k writes, the gating / decision chain of operation k+1 reads only O(n) entries - the
// three robot columns and the 2x2 diagonal blocks. Those are kept in a small cache (strip[3][lds],
// diag[lm][4]) that the gain kernel itself brings up to date with the same two-fma expression the
// sweep applies to P, so propagate, gating and the decision of operation k+1 run on a second stream
// WHILE the sweep of operation k is still streaming P; only the gain kernel (which needs two full,
// swept columns of P) stays between two consecutive sweeps. P's own robot rows / columns are stale
// during such a run and are written back from the cache when it ends (la_store); the diagonal blocks in
// P always equal the cache (same operations, same bits). The control block is double-buffered so the
// decision of operation k+1 never overwrites what the sweep of operation k is reading.
#include "ekf_cta.cuh"
#include "ekf_internal.h"
#include "ekf_la.cuh"
#include "ekf_pdl.cuh"

namespace {

constexpr int kThreads = 256;
// Side-stream kernels of a look-ahead run share the SMs with the persistent CTAs of the sweep, which leave
// about a third of the register file free: small CTAs, so they are scheduled while the sweep is running
// instead of after its first CTAs retire.
constexpr int kSideThreads = 64;

struct LargeSmall {
  PropSetup prop;
  UpdateSetup upd;
  int decision, opt_i, n, n_lm;
  int gate_nlm;   // gating bound frozen at doUpdate entry (Update.cpp:26), see large_gate
  double mahal;
  double res[2], S[4], Si[4], h3[2];
  double l, sq0, sq1, m0, m1;
  double nl[2], PLL[4], h3n[2];
  double cres, cS, csq, cm0;
  double2 Wp[3];  // look-ahead: downdate vectors of the pose rows (computed with the decision)
};

struct LargeArgs {
  EkfState st;
  int f;
  EkfConst k;
  LargeSmall* sm;
  double2* W;
  double* cand_val;
  int* cand_idx;
  int n_cand;
  double* strip;  // look-ahead cache: strip[r*lds + i] = P(i, r), r = 0..2 (includes the 3x3 robot block)
  double* diag;   // look-ahead cache: diag[4*lm + q] = the 2x2 block of landmark lm, column-major
  int lds;
  int la;         // 1: look-ahead run (the sweep leaves the landmark count alone)
  int reverse;    // plain sweep: walk the tiles back to front
};

__device__ __forceinline__ double* filt_P(const LargeArgs& a) { return a.st.P + (size_t)a.f * a.st.slab; }
__device__ __forceinline__ double* filt_x(const LargeArgs& a) { return a.st.x + (size_t)a.f * a.st.xs; }
__device__ __forceinline__ LaCache la_cache(const LargeArgs& a) { return LaCache{a.strip, a.diag, a.lds}; }

// ---- propagate ---------------------------------------------------------------------------------
// The 3x3 robot block lives in P (element (i,j) at P[i + j*ld]) or, in a look-ahead run, in the cache.
template <bool LA>
__device__ __forceinline__ double* prr_elem(const LargeArgs& a, double* P, int i, int j) {
  return LA ? a.strip + (size_t)j * a.lds + i : P + i + (size_t)j * a.st.ld;
}

template <bool LA>
__global__ void large_prop_setup(const LargeArgs a, const double* vel, const double* rot, const double* dt) {
  ekf_pdl_entry();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double* P = filt_P(a);
  double* x = filt_x(a);
  PropSetup p;
  ekf_build_prop(p, *vel, *rot, *dt, x[2], a.k);
  a.sm->prop = p;
  const double xm0 = p.v * p.c, xm1 = p.v * p.s, xm2 = p.w;   // Propagate.cpp:33-37
  x[0] = x[0] + p.dt * xm0;
  x[1] = x[1] + p.dt * xm1;
  x[2] = x[2] + p.dt * xm2;
  double PRR[9];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) PRR[i + 3 * j] = *prr_elem<LA>(a, P, i, j);
  ekf_prop_prr(p, PRR);
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) *prr_elem<LA>(a, P, i, j) = PRR[i + 3 * j];
}

template <bool LA>
__global__ void __launch_bounds__(kThreads) large_prop_strip(const LargeArgs a) {
  __shared__ PropSetup ps;
  ekf_pdl_entry();
  if (threadIdx.x == 0) ps = a.sm->prop;
  __syncthreads();
  double* P = filt_P(a);
  const int ld = a.st.ld;
  const int n = 3 + 2 * a.st.nlm[a.f];
  for (int j = 3 + blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    if (LA) {                                   // the cache is the strip; P's copy is written back by la_store
      double* c = a.strip + j;
      double a0 = c[0], a1 = c[a.lds], a2 = c[2 * (size_t)a.lds];
      ekf_prop_col(ps, a0, a1, a2);
      c[0] = a0; c[a.lds] = a1; c[2 * (size_t)a.lds] = a2;
      continue;
    }
    double a0 = P[j], a1 = P[j + (size_t)ld], a2 = P[j + (size_t)2 * ld];   // mirror rows: coalesced
    ekf_prop_col(ps, a0, a1, a2);
    P[j] = a0;
    P[j + (size_t)ld] = a1;
    P[j + (size_t)2 * ld] = a2;
    double* c = P + (size_t)j * ld;
    c[0] = a0; c[1] = a1; c[2] = a2;
  }
}

// ---- update: gating ------------------------------------------------------------------------------
template <bool LA>
__device__ __forceinline__ void load_gate_inputs(const LargeArgs& a, const double* P, int ld, int Li, double* p, double* pll) {
  if constexpr (LA) {
    la_gate_inputs(la_cache(a), Li, p, pll);
  } else {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      p[0 + 2 * j] = P[Li + (size_t)j * ld];
      p[1 + 2 * j] = P[Li + 1 + (size_t)j * ld];
    }
    pll[0] = P[Li + (size_t)Li * ld];
    pll[1] = P[Li + 1 + (size_t)Li * ld];
    pll[2] = P[Li + (size_t)(Li + 1) * ld];
    pll[3] = P[Li + 1 + (size_t)(Li + 1) * ld];
  }
}

// chunk_pos: 0 = a doUpdate call of its own (gating bound = live landmark count); 1 = first
// measurement of an n_z > 1 call (same bound, and block 0 records it); 2 = later measurement of that
// call: Update.cpp:26 read n_lm once, so landmarks added since the call began are not candidates.
template <bool LA>
__global__ void __launch_bounds__(kThreads) large_gate(const LargeArgs a, const double* zr, int chunk_pos) {
  __shared__ CtaScratch sc;
  ekf_pdl_entry();
  const double* P = filt_P(a);
  const double* x = filt_x(a);
  const int ld = a.st.ld;
  const int n_lm = chunk_pos == 2 ? a.sm->gate_nlm : a.st.nlm[a.f];
  if (chunk_pos == 1 && blockIdx.x == 0 && threadIdx.x == 0) a.sm->gate_nlm = n_lm;
  if (threadIdx.x == 0) {
    double PRR[9];
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) PRR[i + 3 * j] = *prr_elem<LA>(a, const_cast<double*>(P), i, j);
    UpdateSetup u;
    ekf_build_setup(u, x[2], x[0], x[1], PRR, zr[0], zr[1], zr + 2);
    sc.upd = u;
    if (blockIdx.x == 0) a.sm->upd = u;
  }
  __syncthreads();
  double best = INFINITY;
  int best_idx = INT_MAX;
  for (int lm = blockIdx.x * blockDim.x + threadIdx.x; lm < n_lm; lm += gridDim.x * blockDim.x) {
    const int Li = 3 + 2 * lm;
    double p[6], pll[4];
    load_gate_inputs<LA>(a, P, ld, Li, p, pll);
    GateResult g;
    ekf_gate_landmark(sc.upd, x[Li], x[Li + 1], p, pll, a.k.cond_max, g);
    const bool valid = !g.skip && (a.k.mahal_init > g.d2);
    if (valid && g.d2 < best) { best = g.d2; best_idx = Li; }
  }
  cta_argmin(best, best_idx, &sc);
  if (threadIdx.x == 0) {
    a.cand_val[blockIdx.x] = best;
    a.cand_idx[blockIdx.x] = best_idx;
  }
}

// ---- update: decision ------------------------------------------------------------------------------
// Look-ahead runs: the decision thread also finishes the three pose rows of the operation (gain, state,
// downdate vectors, the 3x3 block of the cache) - it has every input in the cache, and it keeps the
// gain kernel, the only kernel left between two sweeps, free of any cross-CTA dependence.
template <bool LA>
__global__ void __launch_bounds__(kThreads) large_decide(const LargeArgs a, const double* zr, int* out_decision,
                                                        int* out_index, double* out_mahal) {
  __shared__ CtaScratch sc;
  ekf_pdl_entry();
  double val = INFINITY;
  int idx = INT_MAX;
  for (int c = threadIdx.x; c < a.n_cand; c += blockDim.x) {
    const double v = a.cand_val[c];
    const int i = a.cand_idx[c];
    if (v < val || (v == val && i < idx)) { val = v; idx = i; }
  }
  cta_argmin(val, idx, &sc);
  if (threadIdx.x != 0) return;
  LargeSmall* sm = a.sm;
  const double* P = filt_P(a);
  double* x = filt_x(a);
  const int ld = a.st.ld;
  const int n_lm = a.st.nlm[a.f];
  const int n = 3 + 2 * n_lm;
  const int opt_i = (idx == INT_MAX) ? 0 : idx;
  const double mahal = (idx == INT_MAX) ? a.k.mahal_init : val;
  int decision = ekf_decide(opt_i, mahal, a.k);
  int index = opt_i;
  const UpdateSetup& u = sm->upd;
  if (decision == EKF_DEC_OLD) {
    double p[6], pll[4];
    load_gate_inputs<LA>(a, P, ld, opt_i, p, pll);
    GateResult g;
    ekf_gate_landmark(u, x[opt_i], x[opt_i + 1], p, pll, a.k.cond_max, g);   // same bits as the gating pass
    sm->res[0] = g.res0; sm->res[1] = g.res1;
    for (int q = 0; q < 4; ++q) sm->S[q] = g.S[q];
    sm->h3[0] = g.h3_0; sm->h3[1] = g.h3_1;
    ekf_inv2(g.S, sm->Si);
    const double d0 = g.S[0], l = g.S[1] / g.S[0], d1 = g.S[3] - l * g.S[1];
    sm->l = l;
    sm->sq0 = sqrt(fabs(d0));
    sm->sq1 = sqrt(fabs(d1));
    sm->m0 = d0 < 0 ? 1.0 : -1.0;
    sm->m1 = d1 < 0 ? 1.0 : -1.0;
    if (LA) la_pose_rows_old(la_cache(a), sm, x, opt_i);   // the pose rows of large_gain's Old branch + the cache's 3x3 block
  } else if (decision == EKF_DEC_NEW) {
    if (n_lm >= a.st.cap_lm) {
      decision = EKF_DEC_DROPPED;
      index = -1;
      a.st.status[a.f] |= 1;
    } else {
      const double c = u.c, s = u.s, z0 = zr[0], z1 = zr[1];
      const double Cz0 = c * z0 + (-s) * z1, Cz1 = s * z0 + c * z1;   // Update.cpp:155
      const double nl0 = u.x0 + Cz0, nl1 = u.x1 + Cz1;
      const double dn0 = nl0 - u.x0, dn1 = nl1 - u.x1;
      const double h30 = u.mCtJ[0] * dn0 + u.mCtJ[2] * dn1;
      const double h31 = u.mCtJ[1] * dn0 + u.mCtJ[3] * dn1;
      const double HR[6] = {u.mCt[0], u.mCt[1], u.mCt[2], u.mCt[3], h30, h31};
      double a1[6], t1[4], in[4], b1[4];
      for (int j = 0; j < 3; ++j) {
        a1[0 + 2 * j] = u.q[0 + 2 * j] + h30 * u.PRR[2 + 3 * j];
        a1[1 + 2 * j] = u.q[1 + 2 * j] + h31 * u.PRR[2 + 3 * j];
      }
      for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 2; ++i) t1[i + 2 * j] = (a1[i] * HR[j] + a1[i + 2] * HR[j + 2]) + a1[i + 4] * HR[j + 4];
      for (int q = 0; q < 4; ++q) in[q] = t1[q] + u.R[q];
      const double Cm[4] = {u.Ct[0], u.Ct[2], u.Ct[1], u.Ct[3]};
      for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 2; ++i) b1[i + 2 * j] = Cm[i] * in[0 + 2 * j] + Cm[i + 2] * in[1 + 2 * j];
      for (int j = 0; j < 2; ++j)       // Update.cpp:168
        for (int i = 0; i < 2; ++i) sm->PLL[i + 2 * j] = b1[i] * u.Ct[0 + 2 * j] + b1[i + 2] * u.Ct[1 + 2 * j];
      sm->nl[0] = nl0; sm->nl[1] = nl1;
      sm->h3n[0] = h30; sm->h3n[1] = h31;
      index = n;
      if (LA) la_pose_rows_new(la_cache(a), sm, x, a.st.nlm + a.f, n, n_lm);
    }
  }
  sm->decision = decision;
  sm->opt_i = opt_i;
  sm->n = n;
  sm->n_lm = n_lm;
  sm->mahal = mahal;
  if (out_decision) *out_decision = decision;
  if (out_index) *out_index = index;
  if (out_mahal) *out_mahal = mahal;
}

// ---- update: gain / augmentation (O(n)) ----------------------------------------------------------
__global__ void __launch_bounds__(kThreads) large_gain(const LargeArgs a) {
  ekf_pdl_entry();
  const LargeSmall* sm = a.sm;
  const int decision = sm->decision;
  if (decision != EKF_DEC_OLD && decision != EKF_DEC_NEW) return;
  double* P = filt_P(a);
  double* x = filt_x(a);
  const int ld = a.st.ld, n = sm->n;
  const UpdateSetup& u = sm->upd;
  const int i0 = blockIdx.x * kThreads + threadIdx.x, stride = gridDim.x * kThreads;
  if (decision == EKF_DEC_OLD) {
    const int opt_i = sm->opt_i;
    const double h00 = u.mCt[0], h01 = u.mCt[2], h02 = sm->h3[0];
    const double h10 = u.mCt[1], h11 = u.mCt[3], h12 = sm->h3[1];
    const double c00 = u.Ct[0], c10 = u.Ct[2], c01 = u.Ct[1], c11 = u.Ct[3];
    const double s00 = sm->Si[0], s10 = sm->Si[1], s01 = sm->Si[2], s11 = sm->Si[3];
    const double r0 = sm->res[0], r1 = sm->res[1], l = sm->l, sq0 = sm->sq0, sq1 = sm->sq1;
    const int n_even = (n + 1) & ~1;
    for (int i = i0; i < n_even; i += stride) {
      if (i >= n) { a.W[i] = make_double2(0.0, 0.0); continue; }   // pad row of the double2 sweep
      const double p0 = P[i], p1 = P[i + (size_t)ld], p2 = P[i + (size_t)2 * ld];
      const double pa = P[i + (size_t)opt_i * ld], pb = P[i + (size_t)(opt_i + 1) * ld];
      const double A0 = (p0 * h00 + p1 * h01) + p2 * h02;
      const double A1 = (p0 * h10 + p1 * h11) + p2 * h12;
      const double B0 = pa * c00 + pb * c10;
      const double B1 = pa * c01 + pb * c11;
      const double M0 = A0 + B0, M1 = A1 + B1;
      const double K0 = M0 * s00 + M1 * s10;      // Update.cpp:186
      const double K1 = M0 * s01 + M1 * s11;
      x[i] = x[i] + (K0 * r0 + K1 * r1);          // :187
      a.W[i] = make_double2(sq0 * fma(l, K1, K0), sq1 * K1);
    }
  } else {
    const double h00 = u.mCt[0], h01 = u.mCt[2], h02 = sm->h3n[0];
    const double h10 = u.mCt[1], h11 = u.mCt[3], h12 = sm->h3n[1];
    const double ct00 = u.Ct[0], ct10 = u.Ct[1], ct01 = u.Ct[2], ct11 = u.Ct[3];
    for (int i = i0; i < n; i += stride) {        // Update.cpp:169,175-176
      const double q0 = -P[i], q1 = -P[i + (size_t)ld], q2 = -P[i + (size_t)2 * ld];
      const double t0 = (q0 * h00 + q1 * h01) + q2 * h02;
      const double t1 = (q0 * h10 + q1 * h11) + q2 * h12;
      const double o0 = t0 * ct00 + t1 * ct10;
      const double o1 = t0 * ct01 + t1 * ct11;
      P[i + (size_t)n * ld] = o0;
      P[i + (size_t)(n + 1) * ld] = o1;
      P[n + (size_t)i * ld] = o0;
      P[n + 1 + (size_t)i * ld] = o1;
    }
    if (i0 == 0) {
      const double off = 0.5 * (sm->PLL[2] + sm->PLL[1]);
      P[n + (size_t)n * ld] = sm->PLL[0];
      P[n + 1 + (size_t)n * ld] = off;
      P[n + (size_t)(n + 1) * ld] = off;
      P[n + 1 + (size_t)(n + 1) * ld] = sm->PLL[3];
      x[n] = sm->nl[0];
      x[n + 1] = sm->nl[1];
    }
  }
}

// Look-ahead form of large_gain: one landmark (two rows) per thread, pose rows already done by
// large_decide<true>. Besides what large_gain does, the thread brings the landmark's cache entries (its
// three strip columns' worth of robot-row elements and its 2x2 diagonal block) up to date with exactly
// the fma pair the sweep applies to the same elements of P.
__global__ void __launch_bounds__(kThreads) large_gain_la(const LargeArgs a) {
  ekf_pdl_entry();
  const LargeSmall* sm = a.sm;
  const int decision = sm->decision;
  if (decision != EKF_DEC_OLD && decision != EKF_DEC_NEW) return;
  double* P = filt_P(a);
  double* x = filt_x(a);
  const LaCache c = la_cache(a);
  const int ld = a.st.ld, n = sm->n, n_lm = sm->n_lm, lds = a.lds;
  const int lm0 = blockIdx.x * kThreads + threadIdx.x, stride = gridDim.x * kThreads;
  if (decision == EKF_DEC_OLD) {
    const int opt_i = sm->opt_i;
    const LaGain q = la_gain_coef(sm);
    const double m0 = sm->m0, m1 = sm->m1;
    const double2 wp[3] = {sm->Wp[0], sm->Wp[1], sm->Wp[2]};
    if (lm0 == 0) {
      a.W[0] = wp[0]; a.W[1] = wp[1]; a.W[2] = wp[2];
      a.W[n] = make_double2(0.0, 0.0);            // pad row of the double2 sweep (n is odd)
    }
    for (int lm = lm0; lm < n_lm; lm += stride) {
      double2 w[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = 3 + 2 * lm + e;
        double dx;
        la_gain_row(q, a.strip[i], a.strip[lds + i], a.strip[2 * (size_t)lds + i], P[i + (size_t)opt_i * ld],
                    P[i + (size_t)(opt_i + 1) * ld], dx, w[e]);
        x[i] = x[i] + dx;                           // Update.cpp:187
        a.W[i] = w[e];
      }
      la_cache_landmark(c, lm, w, wp, m0, m1);
    }
  } else {
    const LaNew q = la_new_coef(sm);
    for (int lm = lm0; lm < n_lm; lm += stride) {   // Update.cpp:169,175-176 (rows 0..2: large_decide<true>)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = 3 + 2 * lm + e;
        double o0, o1;
        la_new_row(q, a.strip[i], a.strip[lds + i], a.strip[2 * (size_t)lds + i], o0, o1);
        P[i + (size_t)n * ld] = o0;
        P[i + (size_t)(n + 1) * ld] = o1;
        P[n + (size_t)i * ld] = o0;
        P[n + 1 + (size_t)i * ld] = o1;
      }
    }
    if (lm0 == 0) {
      const double off = 0.5 * (sm->PLL[2] + sm->PLL[1]);
      P[n + (size_t)n * ld] = sm->PLL[0];
      P[n + 1 + (size_t)n * ld] = off;
      P[n + (size_t)(n + 1) * ld] = off;
      P[n + 1 + (size_t)(n + 1) * ld] = sm->PLL[3];
    }
  }
}

// ---- the HBM-bound kernel: dense symmetric rank-RANK downdate --------------------------------------
// P_ij <- P_ij + u_i0*W_j0 + u_i1*W_j1,  u = (m0*W_0, m1*W_1). Column-major P: each thread owns
// two consecutive rows (one 16-byte double2) of a 512-row panel and walks CB columns with all
// loads issued before the first use. Algorithmic traffic: one read + one write of P = 16 n^2 bytes.
constexpr int kCB = 8;
template <int RANK, bool COMPASS>
__global__ void __launch_bounds__(kThreads) large_downdate(const LargeArgs a) {
  ekf_pdl_entry();
  const LargeSmall* sm = a.sm;
  if (!COMPASS) {
    if (sm->decision != EKF_DEC_OLD) {
      if (sm->decision == EKF_DEC_NEW && !a.la && blockIdx.x == 0 && threadIdx.x == 0) a.st.nlm[a.f] = sm->n_lm + 1;
      return;
    }
  }
  double* P = filt_P(a);
  const int ld = a.st.ld;
  const int n = sm->n;      // set by large_decide / large_compass_setup
  const double m0 = COMPASS ? sm->cm0 : sm->m0, m1 = COMPASS ? 0.0 : sm->m1;
  const int n_even = (n + 1) & ~1;
  const int rows_per_panel = 2 * kThreads;
  const int n_panels = (n_even + rows_per_panel - 1) / rows_per_panel;
  const int n_cb = (n + kCB - 1) / kCB;
  const long n_tiles = (long)n_panels * n_cb;
  const double2* __restrict__ W = a.W;
  for (long tk = blockIdx.x; tk < n_tiles; tk += gridDim.x) {
    const long tile = a.reverse ? n_tiles - 1 - tk : tk;
    const int panel = (int)(tile % n_panels), cb = (int)(tile / n_panels);
    const int i = panel * rows_per_panel + 2 * threadIdx.x;
    if (i >= n_even) continue;
    const double2 wa = W[i], wb = W[i + 1];
    const double ua0 = m0 * wa.x, ua1 = m1 * wa.y, ub0 = m0 * wb.x, ub1 = m1 * wb.y;
    const int j0 = cb * kCB;
    double2* base = reinterpret_cast<double2*>(P + i + (size_t)j0 * ld);
    const size_t cstride = (size_t)ld / 2;   // ld is even
    if (j0 + kCB <= n) {
      double2 v[kCB];
#pragma unroll
      for (int j = 0; j < kCB; ++j) v[j] = base[j * cstride];
#pragma unroll
      for (int j = 0; j < kCB; ++j) {
        const double2 wj = W[j0 + j];
        if (RANK == 2) { v[j].x = fma(ua1, wj.y, v[j].x); v[j].y = fma(ub1, wj.y, v[j].y); }
        v[j].x = fma(ua0, wj.x, v[j].x);
        v[j].y = fma(ub0, wj.x, v[j].y);
      }
#pragma unroll
      for (int j = 0; j < kCB; ++j) base[j * cstride] = v[j];
    } else {
      for (int j = 0; j0 + j < n; ++j) {
        double2 v = base[j * cstride];
        const double2 wj = W[j0 + j];
        if (RANK == 2) { v.x = fma(ua1, wj.y, v.x); v.y = fma(ub1, wj.y, v.y); }
        v.x = fma(ua0, wj.x, v.x);
        v.y = fma(ub0, wj.x, v.y);
        base[j * cstride] = v;
      }
    }
  }
}

// ---- compass ---------------------------------------------------------------------------------------
template <bool LA>
__global__ void large_compass_setup(const LargeArgs a, const double* z, const double* R) {
  ekf_pdl_entry();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double* P = filt_P(a);
  double* x = filt_x(a);
  LargeSmall* sm = a.sm;
  sm->cres = ekf_compass_residual(x[2], *z, a.k);
  const double S = *prr_elem<LA>(a, P, 2, 2) + *R;
  sm->cS = S;
  sm->csq = sqrt(fabs(S));
  sm->cm0 = S < 0 ? 1.0 : -1.0;
  sm->n = 3 + 2 * a.st.nlm[a.f];
  if (LA) la_compass_pose_rows(la_cache(a), sm, x);   // pose rows of large_compass_gain + the cache's 3x3 block
}

__global__ void __launch_bounds__(kThreads) large_compass_gain(const LargeArgs a) {
  ekf_pdl_entry();
  const LargeSmall* sm = a.sm;
  const double* P = filt_P(a);
  double* x = filt_x(a);
  const int ld = a.st.ld, n = 3 + 2 * a.st.nlm[a.f];
  const double res = sm->cres, invS = 1 / sm->cS, sq = sm->csq;
  const int n_even = (n + 1) & ~1;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n_even; i += gridDim.x * kThreads) {
    if (i >= n) { a.W[i] = make_double2(0.0, 0.0); continue; }
    const double Ki = invS * P[i + (size_t)2 * ld];
    x[i] = x[i] + res * Ki;
    a.W[i] = make_double2(sq * Ki, 0.0);
  }
}

// Look-ahead form of large_compass_gain: landmark rows only, cache kept current (rank 1).
__global__ void __launch_bounds__(kThreads) large_compass_gain_la(const LargeArgs a) {
  ekf_pdl_entry();
  const LargeSmall* sm = a.sm;
  double* x = filt_x(a);
  const LaCache c = la_cache(a);
  const int n = sm->n, n_lm = (n - 3) / 2;
  const int lm0 = blockIdx.x * kThreads + threadIdx.x;
  if (lm0 == 0) {
    a.W[0] = sm->Wp[0]; a.W[1] = sm->Wp[1]; a.W[2] = sm->Wp[2];
    a.W[n] = make_double2(0.0, 0.0);
  }
  for (int lm = lm0; lm < n_lm; lm += gridDim.x * kThreads) {
    double wv[2];
    la_compass_landmark(c, sm, x, lm, wv);
    a.W[3 + 2 * lm] = make_double2(wv[0], 0.0);
    a.W[4 + 2 * lm] = make_double2(wv[1], 0.0);
  }
}

// ---- look-ahead cache <-> P -----------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) large_la_load(const LargeArgs a) {
  const double* P = filt_P(a);
  const int ld = a.st.ld, n = 3 + 2 * a.st.nlm[a.f];
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    a.strip[i] = P[i];
    a.strip[a.lds + i] = P[i + (size_t)ld];
    a.strip[2 * (size_t)a.lds + i] = P[i + (size_t)2 * ld];
    if (i >= 3) {                                   // i = Li + e: column e of the landmark's 2x2 block
      const int Li = 3 + 2 * ((i - 3) >> 1), e = (i - 3) & 1;
      a.diag[2 * (size_t)(Li - 3) + 2 * e + 0] = P[Li + (size_t)i * ld];
      a.diag[2 * (size_t)(Li - 3) + 2 * e + 1] = P[Li + 1 + (size_t)i * ld];
    }
  }
}
// Robot rows and columns of P from the cache (the diagonal blocks of P are current, see the header).
__global__ void __launch_bounds__(kThreads) large_la_store(const LargeArgs a) {
  double* P = filt_P(a);
  const int ld = a.st.ld, n = 3 + 2 * a.st.nlm[a.f];
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const double v0 = a.strip[i], v1 = a.strip[a.lds + i], v2 = a.strip[2 * (size_t)a.lds + i];
    P[i] = v0;
    P[i + (size_t)ld] = v1;
    P[i + (size_t)2 * ld] = v2;
    if (i >= 3) {
      double* c = P + (size_t)i * ld;
      c[0] = v0; c[1] = v1; c[2] = v2;
    }
  }
}

LargeArgs make_args(const EkfState& st, int f, const EkfConst& k, const EkfLargeWork& wk, int n_cand) {
  LargeArgs a;
  a.st = st;
  a.f = f;
  a.k = k;
  a.sm = reinterpret_cast<LargeSmall*>(wk.small);
  a.W = wk.W;
  a.cand_val = wk.cand_val;
  a.cand_idx = wk.cand_idx;
  a.n_cand = n_cand;
  a.strip = wk.strip;
  a.diag = wk.diag;
  a.lds = wk.lds;
  a.la = 0;
  a.reverse = 0;
  return a;
}

int gate_grid(const EkfState& st, const EkfLargeWork& wk) {
  int g = (st.cap_lm + kThreads - 1) / kThreads;
  if (g > wk.grid) g = wk.grid;
  return g < 1 ? 1 : g;
}
int row_grid(const EkfState& st, const EkfLargeWork& wk) {
  int g = (st.cap_n + kThreads - 1) / kThreads;
  if (g > wk.grid) g = wk.grid;
  return g < 1 ? 1 : g;
}

// The dense sweep of one operation (rank 2 after a landmark update, rank 1 after a compass update).
void launch_sweep(LargeArgs a, const EkfLargeWork& wk, bool compass, EkfLargeTiming* tm, cudaStream_t s) {
  const bool sample = !compass && tm && tm->used < tm->cap && (tm->seen++ % tm->every) == 0;
  if (sample) cudaEventRecord(tm->ev0[tm->used], s);
  const int reverse = wk.snake ? (int)((*wk.sweep_seq)++ & 1) : 0;
  a.reverse = reverse;
  if (wk.use_tma) {
    EkfLargeTmaArgs t{&a.sm->decision, &a.sm->n, &a.sm->m0, &a.sm->m1, a.W, a.la ? nullptr : a.st.nlm + a.f, &a.sm->n_lm, reverse};
    if (compass) t = EkfLargeTmaArgs{nullptr, &a.sm->n, &a.sm->cm0, &a.sm->cm0, a.W, nullptr, nullptr, reverse};
    ekf_large_tma_downdate(t, wk.tmaps + (size_t)a.f * ekf_large_tma_map_bytes(), wk.tma_grid, compass, s);
  } else if (compass) {
    ekf_launch_pdl(large_downdate<1, true>, wk.grid, kThreads, 0, s, a);
  } else {
    ekf_launch_pdl(large_downdate<2, false>, wk.grid, kThreads, 0, s, a);
  }
  if (sample) cudaEventRecord(tm->ev1[tm->used++], s);
}

void launch_propagate(const LargeArgs& a, const EkfState& st, const EkfLargeWork& wk, const double* vel,
                      const double* rot, const double* dt, cudaStream_t s) {
  ekf_launch_pdl(large_prop_setup<false>, 1, 32, 0, s, a, vel, rot, dt);
  ekf_launch_pdl(large_prop_strip<false>, row_grid(st, wk), kThreads, 0, s, a);
}
void launch_update(LargeArgs a, const EkfState& st, const EkfLargeWork& wk, const double* zr, int chunk_pos, int* dec,
                   int* idx, double* mah, EkfLargeTiming* tm, cudaStream_t s) {
  const int gg = gate_grid(st, wk);
  a.n_cand = gg;
  ekf_launch_pdl(large_gate<false>, gg, kThreads, 0, s, a, zr, chunk_pos);
  ekf_launch_pdl(large_decide<false>, 1, kThreads, 0, s, a, zr, dec, idx, mah);
  ekf_launch_pdl(large_gain, row_grid(st, wk), kThreads, 0, s, a);
  launch_sweep(a, wk, false, tm, s);
}
void launch_compass(const LargeArgs& a, const EkfState& st, const EkfLargeWork& wk, const double* z, const double* R,
                    cudaStream_t s) {
  ekf_launch_pdl(large_compass_setup<false>, 1, 32, 0, s, a, z, R);
  ekf_launch_pdl(large_compass_gain, row_grid(st, wk), kThreads, 0, s, a);
  launch_sweep(a, wk, true, nullptr, s);
}

}  // namespace

constexpr size_t kSmallStride = (sizeof(LargeSmall) + 15) / 16 * 16;   // two control blocks (look-ahead)
size_t ekf_large_small_doubles() { return 2 * kSmallStride / 8; }

cudaError_t ekf_large_prepare(int sm_count, int* grid) {
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, large_downdate<2, false>, kThreads, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  *grid = per_sm * sm_count;
  return cudaSuccess;
}

int ekf_large_launches_per(EkfOp op) { return op == EKF_OP_PROPAGATE ? 2 : op == EKF_OP_UPDATE ? 4 : 3; }

cudaError_t ekf_large_percall(const EkfState& st, int filter, const EkfPercallIO& io, EkfOp op, const EkfConst& k,
                              const EkfLargeWork& wk, EkfLargeTiming* tm, cudaStream_t stream) {
  const LargeArgs a = make_args(st, filter, k, wk, 0);
  if (op == EKF_OP_PROPAGATE) {
    launch_propagate(a, st, wk, io.vel + filter, io.rotvel + filter, io.dt + (size_t)filter * io.dt_stride, stream);
  } else if (op == EKF_OP_COMPASS) {
    launch_compass(a, st, wk, io.cz + filter, io.cR + filter, stream);
  } else {
    for (int m = 0; m < io.n_z; ++m) {
      const size_t oi = (size_t)filter * io.n_z + m;
      launch_update(a, st, wk, io.zr + oi * 6, io.n_z == 1 ? 0 : (m == 0 ? 1 : 2), io.decision ? io.decision + oi : nullptr,
                    io.index ? io.index + oi : nullptr, io.mahal ? io.mahal + oi : nullptr, tm, stream);
    }
  }
  return cudaGetLastError();
}

// Fused-run equivalent for one filter: the host enqueues the kernel chain of every step; the
// per-step flags (has_compass, n_z) come from the host mirror the C ABI keeps of the records.
//
// With look-ahead (wk.la, header comment) the chain is split over two streams. `stream` carries only
// what has to sit between two sweeps: [gain of operation k] [sweep k]. wk.s_side carries everything
// that works on the cache alone: [propagate] [gating k+1] [decision k+1], started as soon as gain k has
// brought the cache up to date (event ev_a) and joined again before gain k+1 (event ev_b).
static cudaError_t large_run_lookahead(const EkfState& st, int filter, const EkfRunIO& io, const uint8_t* has_compass,
                                       const uint8_t* n_z, const EkfConst& k, const EkfLargeWork& wk,
                                       EkfLargeTiming* tm, cudaStream_t A, long long* launches) {
  LargeArgs a = make_args(st, filter, k, wk, 0);
  a.la = 1;
  cudaStream_t B = wk.s_side;
  LargeSmall* const small0 = a.sm;
  auto ctl = [&](long op) { return reinterpret_cast<LargeSmall*>(reinterpret_cast<unsigned char*>(small0) + (op & 1) * kSmallStride); };
  const int rg = row_grid(st, wk), gg = gate_grid(st, wk);   // one landmark per thread: the gating grid fits both
  auto side_grid = [&](int count) {
    int g = (count + kSideThreads - 1) / kSideThreads;
    if (g > wk.grid) g = wk.grid;
    return g < 1 ? 1 : g;
  };
  const int sg_rows = side_grid(st.cap_n), sg_lm = side_grid(st.cap_lm);
  large_la_load<<<rg, kThreads, 0, A>>>(a);
  cudaEventRecord(wk.ev_a, A);
  cudaStreamWaitEvent(B, wk.ev_a, 0);
  *launches += 1;
  long op = 0;
  bool after_wait = true;          // the next side-stream kernel follows a cross-stream wait: plain launch
  auto side = [&](auto kern, int grid, int threads, auto... args) {
    if (after_wait) kern<<<grid, threads, 0, B>>>(args...);
    else ekf_launch_pdl(kern, grid, threads, 0, B, args...);
    after_wait = false;
    *launches += 1;
  };
  for (int t = 0; t < io.T; ++t) {
    const double* rec = io.records + ((size_t)filter * io.T + t) * io.L;
    a.sm = ctl(op);
    side(large_prop_setup<true>, 1, 32, a, rec + 0, rec + 1, rec + 2);
    side(large_prop_strip<true>, sg_rows, kSideThreads, a);
    if (has_compass[t]) {
      side(large_compass_setup<true>, 1, 32, a, rec + 3, rec + 4);
      cudaEventRecord(wk.ev_b, B);
      cudaStreamWaitEvent(A, wk.ev_b, 0);
      large_compass_gain_la<<<gg, kThreads, 0, A>>>(a);
      cudaEventRecord(wk.ev_a, A);
      launch_sweep(a, wk, true, nullptr, A);
      cudaStreamWaitEvent(B, wk.ev_a, 0);
      after_wait = true;
      *launches += 2;
      a.sm = ctl(++op);
    }
    for (int m = 0; m < io.M; ++m) {
      if (m >= n_z[t]) continue;
      const size_t oi = ((size_t)filter * io.T + t) * io.M + m;
      a.n_cand = sg_lm;
      side(large_gate<true>, sg_lm, kSideThreads, a, rec + 8 + 6 * m, 0);
      side(large_decide<true>, 1, kSideThreads, a, rec + 8 + 6 * m, io.decision ? io.decision + oi : nullptr,
           io.index ? io.index + oi : nullptr, io.mahal ? io.mahal + oi : nullptr);
      cudaEventRecord(wk.ev_b, B);
      cudaStreamWaitEvent(A, wk.ev_b, 0);
      large_gain_la<<<gg, kThreads, 0, A>>>(a);
      cudaEventRecord(wk.ev_a, A);
      launch_sweep(a, wk, false, tm, A);
      cudaStreamWaitEvent(B, wk.ev_a, 0);
      after_wait = true;
      *launches += 2;
      a.sm = ctl(++op);
    }
    if (io.pose_trace)             // the pose is written by side-stream kernels only
      cudaMemcpyAsync(io.pose_trace + ((size_t)filter * io.T + t) * 3, st.x + (size_t)filter * st.xs,
                      3 * sizeof(double), cudaMemcpyDeviceToDevice, B);
  }
  cudaEventRecord(wk.ev_b, B);
  cudaStreamWaitEvent(A, wk.ev_b, 0);
  large_la_store<<<rg, kThreads, 0, A>>>(a);
  *launches += 1;
  return cudaGetLastError();
}

cudaError_t ekf_large_run(const EkfState& st, int filter, const EkfRunIO& io, const uint8_t* has_compass,
                          const uint8_t* n_z, const EkfConst& k, const EkfLargeWork& wk, EkfLargeTiming* tm,
                          cudaStream_t stream, long long* launches) {
  if (wk.la) return large_run_lookahead(st, filter, io, has_compass, n_z, k, wk, tm, stream, launches);
  const LargeArgs a = make_args(st, filter, k, wk, 0);
  for (int t = 0; t < io.T; ++t) {
    const double* rec = io.records + ((size_t)filter * io.T + t) * io.L;
    launch_propagate(a, st, wk, rec + 0, rec + 1, rec + 2, stream);
    *launches += 2;
    if (has_compass[t]) {
      launch_compass(a, st, wk, rec + 3, rec + 4, stream);
      *launches += 3;
    }
    for (int m = 0; m < io.M; ++m) {
      const size_t oi = ((size_t)filter * io.T + t) * io.M + m;
      if (m < n_z[t]) {
        launch_update(a, st, wk, rec + 8 + 6 * m, 0, io.decision ? io.decision + oi : nullptr,
                      io.index ? io.index + oi : nullptr, io.mahal ? io.mahal + oi : nullptr, tm, stream);
        *launches += 4;
      }
    }
    if (io.pose_trace)
      cudaMemcpyAsync(io.pose_trace + ((size_t)filter * io.T + t) * 3, st.x + (size_t)filter * st.xs,
                      3 * sizeof(double), cudaMemcpyDeviceToDevice, stream);
  }
  return cudaGetLastError();
}
