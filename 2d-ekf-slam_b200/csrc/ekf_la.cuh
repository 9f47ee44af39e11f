// ekf_la.cuh — the look-ahead cache of the large-map regimes (ekf_large.cu, ekf_shard.cu).
//
// Of everything the dense covariance sweep of operation k writes, the gating / decision chain of
// operation k+1 (Update.cpp:80-150) reads only O(n) entries: the three robot columns P(:,0:3) and the
// 2x2 diagonal block of every landmark. They are kept in a small cache,
//     strip[r*lds + i] = P(i, r)   r = 0..2, i = 0..n-1   (rows 0..2 of it are the 3x3 robot block)
//     diag[4*lm + q]               the block of landmark lm, column-major {(0,0),(1,0),(0,1),(1,1)}
// which is brought up to date with exactly the fma pair the sweep applies to the same elements of P
//     v <- fma(u_i1, w_j.y, v);  v <- fma(u_i0, w_j.x, v)      u_i = (m0*w_i.x, m1*w_i.y)
// so cache and P never differ by a bit, and propagate / gating / decision can run while the sweep is
// still streaming P. The functions here are the pieces both regimes share; `Small` is the regime's
// control block (LargeSmall / ShardSmall: same member names).
#pragma once
#include "ekf_small.cuh"

struct LaCache {
  double* strip;
  double* diag;
  int lds;
};

__device__ __forceinline__ double* la_prr(const LaCache& c, int i, int j) { return c.strip + (size_t)j * c.lds + i; }

// Inputs of ekf_gate_landmark for the landmark whose state entries start at Li.
__device__ __forceinline__ void la_gate_inputs(const LaCache& c, int Li, double* p, double* pll) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    p[0 + 2 * j] = c.strip[(size_t)j * c.lds + Li];
    p[1 + 2 * j] = c.strip[(size_t)j * c.lds + Li + 1];
  }
  const double2* d = reinterpret_cast<const double2*>(c.diag + 2 * (Li - 3));   // 4 doubles per landmark
  const double2 d0 = d[0], d1 = d[1];
  pll[0] = d0.x; pll[1] = d0.y; pll[2] = d1.x; pll[3] = d1.y;
}

// Everything one gain row of an Old update needs besides the row's five covariance entries.
struct LaGain {
  double h00, h01, h02, h10, h11, h12, c00, c10, c01, c11, s00, s10, s01, s11, r0, r1, l, sq0, sq1;
};
template <class Small>
__device__ __forceinline__ LaGain la_gain_coef(const Small* sm) {
  const UpdateSetup& u = sm->upd;
  LaGain q;
  q.h00 = u.mCt[0]; q.h01 = u.mCt[2]; q.h02 = sm->h3[0];
  q.h10 = u.mCt[1]; q.h11 = u.mCt[3]; q.h12 = sm->h3[1];
  q.c00 = u.Ct[0]; q.c10 = u.Ct[2]; q.c01 = u.Ct[1]; q.c11 = u.Ct[3];
  q.s00 = sm->Si[0]; q.s10 = sm->Si[1]; q.s01 = sm->Si[2]; q.s11 = sm->Si[3];
  q.r0 = sm->res[0]; q.r1 = sm->res[1]; q.l = sm->l; q.sq0 = sm->sq0; q.sq1 = sm->sq1;
  return q;
}
// Row i of K = P H^T S^-1 (Update.cpp:186), its state correction (:187) and its downdate vector.
// p0..p2 = P(i, 0..2), pa / pb = P(i, Opt_i) / P(i, Opt_i + 1).
__device__ __forceinline__ void la_gain_row(const LaGain& q, double p0, double p1, double p2, double pa, double pb,
                                            double& dx, double2& w) {
  const double A0 = (p0 * q.h00 + p1 * q.h01) + p2 * q.h02;
  const double A1 = (p0 * q.h10 + p1 * q.h11) + p2 * q.h12;
  const double B0 = pa * q.c00 + pb * q.c10;
  const double B1 = pa * q.c01 + pb * q.c11;
  const double M0 = A0 + B0, M1 = A1 + B1;
  const double K0 = M0 * q.s00 + M1 * q.s10;
  const double K1 = M0 * q.s01 + M1 * q.s11;
  dx = K0 * q.r0 + K1 * q.r1;
  w = make_double2(q.sq0 * fma(q.l, K1, K0), q.sq1 * K1);
}

// The decision thread finishes the three pose rows of an Old update: state, downdate vectors (kept in
// sm->Wp for the gain kernel) and the 3x3 block of the cache. Call after sm's S^-1 / LDL^T fields are set.
template <class Small>
__device__ __forceinline__ void la_pose_rows_old(const LaCache& c, Small* sm, double* x, int opt_i) {
  const LaGain q = la_gain_coef(sm);
  double2 w[3];
  for (int i = 0; i < 3; ++i) {
    const double p0 = c.strip[i], p1 = c.strip[c.lds + i], p2 = c.strip[2 * (size_t)c.lds + i];
    const double pa = c.strip[(size_t)i * c.lds + opt_i], pb = c.strip[(size_t)i * c.lds + opt_i + 1];
    double dx;
    la_gain_row(q, p0, p1, p2, pa, pb, dx, w[i]);
    x[i] = x[i] + dx;
    sm->Wp[i] = w[i];
  }
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) {
      double* e = la_prr(c, i, j);
      double v = *e;
      v = fma(sm->m1 * w[i].y, w[j].y, v);
      v = fma(sm->m0 * w[i].x, w[j].x, v);
      *e = v;
    }
}

// One row of the two new columns of a New update (Update.cpp:169,175-176): P(i, n), P(i, n+1) from P(i, 0..2).
struct LaNew {
  double h00, h01, h02, h10, h11, h12, ct00, ct10, ct01, ct11;
};
template <class Small>
__device__ __forceinline__ LaNew la_new_coef(const Small* sm) {
  const UpdateSetup& u = sm->upd;
  LaNew q;
  q.h00 = u.mCt[0]; q.h01 = u.mCt[2]; q.h02 = sm->h3n[0];
  q.h10 = u.mCt[1]; q.h11 = u.mCt[3]; q.h12 = sm->h3n[1];
  q.ct00 = u.Ct[0]; q.ct10 = u.Ct[1]; q.ct01 = u.Ct[2]; q.ct11 = u.Ct[3];
  return q;
}
__device__ __forceinline__ void la_new_row(const LaNew& q, double p0, double p1, double p2, double& o0, double& o1) {
  const double q0 = -p0, q1 = -p1, q2 = -p2;
  const double t0 = (q0 * q.h00 + q1 * q.h01) + q2 * q.h02;
  const double t1 = (q0 * q.h10 + q1 * q.h11) + q2 * q.h12;
  o0 = t0 * q.ct00 + t1 * q.ct10;
  o1 = t0 * q.ct01 + t1 * q.ct11;
}
// The decision thread's share of a New update: pose rows of the new columns, the new diagonal block, the
// new state entries and the landmark count - cache / vector writes only, nothing a running sweep touches.
// Call after sm->PLL, nl, h3n are set.
template <class Small>
__device__ __forceinline__ void la_pose_rows_new(const LaCache& c, const Small* sm, double* x, int* nlm, int n, int n_lm) {
  const LaNew q = la_new_coef(sm);
  for (int i = 0; i < 3; ++i) {
    double o0, o1;
    la_new_row(q, c.strip[i], c.strip[c.lds + i], c.strip[2 * (size_t)c.lds + i], o0, o1);
    c.strip[(size_t)i * c.lds + n] = o0;
    c.strip[(size_t)i * c.lds + n + 1] = o1;
  }
  const double off = 0.5 * (sm->PLL[2] + sm->PLL[1]);
  double* d = c.diag + 4 * (size_t)n_lm;
  d[0] = sm->PLL[0]; d[1] = off; d[2] = off; d[3] = sm->PLL[3];
  x[n] = sm->nl[0];
  x[n + 1] = sm->nl[1];
  *nlm = n_lm + 1;
}

// Cache entries of landmark lm after a rank-2 downdate with vectors w[0], w[1] (its two rows) and wp (pose rows).
__device__ __forceinline__ void la_cache_landmark(const LaCache& c, int lm, const double2* w, const double2* wp, double m0,
                                                  double m1) {
  const int Li = 3 + 2 * lm;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const double u0 = m0 * w[e].x, u1 = m1 * w[e].y;
    double* s = c.strip + Li + e;
#pragma unroll
    for (int r = 0; r < 3; ++r) s[(size_t)r * c.lds] = fma(u0, wp[r].x, fma(u1, wp[r].y, s[(size_t)r * c.lds]));
  }
  double2* d = reinterpret_cast<double2*>(c.diag + 4 * (size_t)lm);
  double2 d0 = d[0], d1 = d[1];               // (Li,Li) (Li+1,Li) | (Li,Li+1) (Li+1,Li+1)
  const double ua0 = m0 * w[0].x, ua1 = m1 * w[0].y, ub0 = m0 * w[1].x, ub1 = m1 * w[1].y;
  d0.x = fma(ua0, w[0].x, fma(ua1, w[0].y, d0.x));
  d0.y = fma(ub0, w[0].x, fma(ub1, w[0].y, d0.y));
  d1.x = fma(ua0, w[1].x, fma(ua1, w[1].y, d1.x));
  d1.y = fma(ub0, w[1].x, fma(ub1, w[1].y, d1.y));
  d[0] = d0; d[1] = d1;
}

// ---- compass (kalmanfilter.cpp:96-130), rank 1 ----------------------------------------------------------
// Pose rows by the set-up thread; call after sm->cres, cS, csq, cm0 are set.
template <class Small>
__device__ __forceinline__ void la_compass_pose_rows(const LaCache& c, Small* sm, double* x) {
  const double res = sm->cres, invS = 1 / sm->cS, sq = sm->csq;
  double2 w[3];
  for (int i = 0; i < 3; ++i) {
    const double Ki = invS * c.strip[2 * (size_t)c.lds + i];
    x[i] = x[i] + res * Ki;
    w[i] = make_double2(sq * Ki, 0.0);
    sm->Wp[i] = w[i];
  }
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) {
      double* e = la_prr(c, i, j);
      *e = fma(sm->cm0 * w[i].x, w[j].x, *e);
    }
}
// Both rows of landmark lm: gain, state, downdate vector (returned in wv, stored by the caller) and cache.
template <class Small>
__device__ __forceinline__ void la_compass_landmark(const LaCache& c, const Small* sm, double* x, int lm, double* wv) {
  const double res = sm->cres, invS = 1 / sm->cS, sq = sm->csq, m0 = sm->cm0;
  const double w0 = sm->Wp[0].x, w1 = sm->Wp[1].x, w2 = sm->Wp[2].x;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int i = 3 + 2 * lm + e;
    const double p0 = c.strip[i], p1 = c.strip[c.lds + i], p2 = c.strip[2 * (size_t)c.lds + i];
    const double Ki = invS * p2;
    x[i] = x[i] + res * Ki;
    wv[e] = sq * Ki;
    const double u0 = m0 * wv[e];
    c.strip[i] = fma(u0, w0, p0);
    c.strip[c.lds + i] = fma(u0, w1, p1);
    c.strip[2 * (size_t)c.lds + i] = fma(u0, w2, p2);
  }
  double2* d = reinterpret_cast<double2*>(c.diag + 4 * (size_t)lm);
  double2 d0 = d[0], d1 = d[1];
  const double ua = m0 * wv[0], ub = m0 * wv[1];
  d0.x = fma(ua, wv[0], d0.x);
  d0.y = fma(ub, wv[0], d0.y);
  d1.x = fma(ua, wv[1], d1.x);
  d1.y = fma(ub, wv[1], d1.y);
  d[0] = d0; d[1] = d1;
}
