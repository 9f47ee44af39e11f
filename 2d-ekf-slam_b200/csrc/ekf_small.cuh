// ekf_small.cuh — per-thread small-matrix arithmetic of the EKF-SLAM filter core (FP64).
//
// Everything here follows the OPERATION ORDER of the reference expressions it replaces
// (kentsommer/2D-EKF-SLAM, odometry/Update.cpp:85-148, Propagate.cpp:33-53,
// kalmanfilter.cpp:26-37,96-118): chained matrix products associate left-to-right and every
// inner sum is sequential, seeded with its first product. This translation unit is compiled
// with -fmad=false, so a*b + c*d is DMUL, DMUL, DADD exactly as written: the data-association
// decisions (cond >= 80, Mahalanobis argmin, Gamma thresholds) are computed from the same
// rounded intermediates the reference produces; only cos/sin (CUDA libdevice vs glibc, <= 2 ulp)
// can differ. The O(n^2) covariance work is elsewhere (ekf_cta.cuh) and uses explicit fma().
#pragma once
#include <cstdint>

struct EkfConst {
  double sigma_v, sigma_w, deg2rad_pi, two_pi, cond_max, mahal_init;
  int gamma_max, gamma_min;
};

// Per-update quantities that do not depend on the landmark (Update.cpp:89-95,113-114), built
// once per measurement by one thread and broadcast through shared memory.
struct UpdateSetup {
  double c, s;          // cos(phi), sin(phi)
  double Ct[4];         // H_Li = C^T          column-major {(0,0),(1,0),(0,1),(1,1)}
  double mCt[4];        // -1.0*C^T            (first two columns of every H_R)
  double mCtJ[4];       // (-1.0*C^T)*J
  double PRR[9];        // P_min.block(0,0,3,3), column-major
  double q[6];          // HR(:,0:2)*PRR(0:2,:) partial sums: q(i,j)=mCt(i,0)*PRR(0,j)+mCt(i,1)*PRR(1,j)
  double x0, x1;        // G_pR_hat
  double z0, z1;        // measurement
  double R[4];          // measurement covariance, column-major
};

// The heading-dependent part of UpdateSetup (one sincos), separable so a kernel can compute it
// once per measurement on one thread and let every gating lane complete the rest.
struct UpdateTrig {
  double c, s;
  double Ct[4], mCt[4], mCtJ[4];
};

__device__ __forceinline__ void ekf_build_trig_sc(UpdateTrig& t, double s, double c) {
  t.c = c; t.s = s;
  // C << cos,-sin,sin,cos (Update.cpp:90); C^T
  t.Ct[0] = c;  t.Ct[1] = -s; t.Ct[2] = s;  t.Ct[3] = c;
  for (int k = 0; k < 4; ++k) t.mCt[k] = -1.0 * t.Ct[k];
  // J << 0,-1,1,0 (Update.cpp:73): J(0,0)=0 J(1,0)=1 J(0,1)=-1 J(1,1)=0
  const double J00 = 0.0, J10 = 1.0, J01 = -1.0, J11 = 0.0;
  t.mCtJ[0] = t.mCt[0] * J00 + t.mCt[2] * J10;
  t.mCtJ[1] = t.mCt[1] * J00 + t.mCt[3] * J10;
  t.mCtJ[2] = t.mCt[0] * J01 + t.mCt[2] * J11;
  t.mCtJ[3] = t.mCt[1] * J01 + t.mCt[3] * J11;
}

__device__ __forceinline__ void ekf_build_trig(UpdateTrig& t, double phi) {
  double s, c;
  sincos(phi, &s, &c);
  ekf_build_trig_sc(t, s, c);
}

__device__ __forceinline__ void ekf_complete_setup(UpdateSetup& u, const UpdateTrig& t, double x0, double x1,
                                                   const double* PRR, double z0, double z1, const double* R) {
  u.c = t.c; u.s = t.s;
  for (int k = 0; k < 4; ++k) { u.Ct[k] = t.Ct[k]; u.mCt[k] = t.mCt[k]; u.mCtJ[k] = t.mCtJ[k]; }
  for (int k = 0; k < 9; ++k) u.PRR[k] = PRR[k];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 2; ++i) u.q[i + 2 * j] = u.mCt[i] * PRR[0 + 3 * j] + u.mCt[i + 2] * PRR[1 + 3 * j];
  u.x0 = x0; u.x1 = x1; u.z0 = z0; u.z1 = z1;
  for (int k = 0; k < 4; ++k) u.R[k] = R[k];
}

__device__ __forceinline__ void ekf_build_setup(UpdateSetup& u, double phi, double x0, double x1, const double* PRR,
                                                double z0, double z1, const double* R) {
  UpdateTrig t;
  ekf_build_trig(t, phi);
  ekf_complete_setup(u, t, x0, x1, PRR, z0, z1, R);
}

struct GateResult {
  double res0, res1;    // z - z_hat
  double S[4];          // symmetrised innovation covariance, column-major
  double h3_0, h3_1;    // third column of H_R
  bool skip;            // cond >= cond_max (Update.cpp:131)
  double d2;            // res^T S^-1 res (only meaningful when !skip)
  double Si[4];         // S^-1 as ekf_inv2 returns it (the same expressions): {i00, i10, i01, i11}
};

// ---- one iteration of the landmark loop, Update.cpp:103-136, in four separable pieces ------------
// p[0..5] = P(Li+r, j) for r=0..1, j=0..2 stored as p[r + 2*j] (this is P_LiR; P_RLi is its exact
// transpose because P is bit-symmetric); pll = {P(Li,Li), P(Li+1,Li), P(Li,Li+1), P(Li+1,Li+1)}.
// S = (((t1 + t2) + t3) + t4) + R in exactly that order; the pieces can run on different lanes.
struct GatePre {
  double res0, res1;   // z - z_hat
  double HR[6];        // H_R (2x3), column-major
};

__device__ __forceinline__ void ekf_gate_prelude(const UpdateSetup& u, double lx, double ly, GatePre& g) {
  const double d0 = lx - u.x0, d1 = ly - u.x1;
  const double zh0 = u.Ct[0] * d0 + u.Ct[2] * d1;
  const double zh1 = u.Ct[1] * d0 + u.Ct[3] * d1;
  g.res0 = u.z0 - zh0;
  g.res1 = u.z1 - zh1;
  const double h30 = u.mCtJ[0] * d0 + u.mCtJ[2] * d1;
  const double h31 = u.mCtJ[1] * d0 + u.mCtJ[3] * d1;
  g.HR[0] = u.mCt[0]; g.HR[1] = u.mCt[1]; g.HR[2] = u.mCt[2]; g.HR[3] = u.mCt[3];
  g.HR[4] = h30; g.HR[5] = h31;
}

// t12 = H_R*P_RR*H_R^T + H_Li*P_LiR*H_R^T
__device__ __forceinline__ void ekf_gate_terms12(const UpdateSetup& u, const GatePre& g, const double* p, double* t12) {
  const double* HR = g.HR;
  const double h30 = HR[4], h31 = HR[5];
  // a1 = H_R*P_RR (2x3): ((HR(i,0)*P(0,j) + HR(i,1)*P(1,j)) + HR(i,2)*P(2,j))
  double a1[6];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    a1[0 + 2 * j] = u.q[0 + 2 * j] + h30 * u.PRR[2 + 3 * j];
    a1[1 + 2 * j] = u.q[1 + 2 * j] + h31 * u.PRR[2 + 3 * j];
  }
  double t1[4], t2[4];
  // t1 = a1*H_R^T (2x2): t(i,j) = a(i,0)*HR(j,0) + a(i,1)*HR(j,1) + a(i,2)*HR(j,2)
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 2; ++i)
      t1[i + 2 * j] = (a1[i] * HR[j] + a1[i + 2] * HR[j + 2]) + a1[i + 4] * HR[j + 4];
  // a2 = H_Li*P_LiR (2x3), P_LiR(r,j) = p[r + 2*j]
  double a2[6];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < 2; ++i) a2[i + 2 * j] = u.Ct[i] * p[0 + 2 * j] + u.Ct[i + 2] * p[1 + 2 * j];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 2; ++i)
      t2[i + 2 * j] = (a2[i] * HR[j] + a2[i + 2] * HR[j + 2]) + a2[i + 4] * HR[j + 4];
#pragma unroll
  for (int k = 0; k < 4; ++k) t12[k] = t1[k] + t2[k];
}

// t3 = H_R*P_RLi*H_Li^T, t4 = H_Li*P_LiLi*H_Li^T
__device__ __forceinline__ void ekf_gate_terms34(const UpdateSetup& u, const GatePre& g, const double* p,
                                                 const double* pll, double* t3, double* t4) {
  const double* HR = g.HR;
  // a3 = H_R*P_RLi (2x2), P_RLi(r,j) = P(r, Li+j) = p[j + 2*r]
  double a3[4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 2; ++i)
      a3[i + 2 * j] = (HR[i] * p[j + 0] + HR[i + 2] * p[j + 2]) + HR[i + 4] * p[j + 4];
  // t3 = a3*H_Li^T ; H_Li^T = C, C(k,j): C(0,0)=c C(1,0)=s C(0,1)=-s C(1,1)=c = {Ct[0],Ct[2],Ct[1],Ct[3]}
  const double Cm[4] = {u.Ct[0], u.Ct[2], u.Ct[1], u.Ct[3]};
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 2; ++i) t3[i + 2 * j] = a3[i] * Cm[0 + 2 * j] + a3[i + 2] * Cm[1 + 2 * j];
  // a4 = H_Li*P_LiLi ; t4 = a4*H_Li^T
  double a4[4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 2; ++i) a4[i + 2 * j] = u.Ct[i] * pll[0 + 2 * j] + u.Ct[i + 2] * pll[1 + 2 * j];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 2; ++i) t4[i + 2 * j] = a4[i] * Cm[0 + 2 * j] + a4[i + 2] * Cm[1 + 2 * j];
}

// Everything of one landmark-loop iteration behind the raw innovation covariance: Sraw[k] =
// ((t12[k] + t3[k]) + t4[k]) + R[k] (column-major 2x2). Symmetrisation, condition gate, Mahalanobis
// distance (Update.cpp:123-136). Shared by the one-lane-per-landmark path (ekf_gate_finish) and the
// four-lanes-per-landmark path of ekf_dtile.cu, so both produce the same bits.
__device__ __forceinline__ void ekf_gate_from_S(const GatePre& pre, const double* Sraw, double cond_max, GateResult& g) {
  g.res0 = pre.res0; g.res1 = pre.res1;
  g.h3_0 = pre.HR[4]; g.h3_1 = pre.HR[5];
  // S = 0.5*(S + S^T) (Update.cpp:123-124)
  const double s01 = 0.5 * (Sraw[2] + Sraw[1]);
  const double s10 = 0.5 * (Sraw[1] + Sraw[2]);
  g.S[0] = 0.5 * (Sraw[0] + Sraw[0]);
  g.S[3] = 0.5 * (Sraw[3] + Sraw[3]);
  g.S[2] = s01;
  g.S[1] = s10;
  // condition number via 2x2 singular values (Update.cpp:127-128): cond = (Q+R)/|Q-R| with
  // Q = sqrt(q2), R = sqrt(r2). Only the comparison cond >= cond_max is ever used (Update.cpp:131).
  // With rho2 = min(q2,r2)/max(q2,r2), cond >= c  <=>  rho2 >= ((c-1)/(c+1))^2; far from that
  // boundary (relative margin 1e-9, ~1e6 times the rounding error of the exact evaluation) the
  // comparison is decided from q2, r2 alone; inside the band, and for NaN / zero radicands, the
  // reference's sqrt/sqrt/divide sequence is evaluated. The outcome is the same either way.
  const double a = g.S[0], c = g.S[1], b = g.S[2], d = g.S[3];
  const double E = (a + d) * 0.5, F = (a - d) * 0.5, G = (c + b) * 0.5, H = (c - b) * 0.5;
  const double q2 = E * E + H * H, r2 = F * F + G * G;
  {
    const bool ordered = (q2 == q2) && (r2 == r2) && cond_max > 1.0;   // NaN goes down the exact path
    const double lo = fmin(q2, r2), hi = fmax(q2, r2);
    const double rr = (cond_max - 1.0) / (cond_max + 1.0);   // uniform: hoisted out of the landmark loop
    const double thr = rr * rr;
    if (ordered && lo < (thr * (1.0 - 1e-9)) * hi) {
      g.skip = false;
    } else if (ordered && lo > (thr * (1.0 + 1e-9)) * hi) {
      g.skip = true;
    } else {
      const double Q = sqrt(q2), Rr = sqrt(r2);
      const double cond = (Q + Rr) / fabs(Q - Rr);
      g.skip = cond >= cond_max;
    }
  }
  // Mahalanobis distance (Update.cpp:135-136)
  const double det = a * d - b * c;
  const double invdet = 1.0 / det;
  const double i00 = d * invdet, i10 = -c * invdet, i01 = -b * invdet, i11 = a * invdet;
  const double r0 = g.res0 * i00 + g.res1 * i10;
  const double r1 = g.res0 * i01 + g.res1 * i11;
  g.d2 = r0 * g.res0 + r1 * g.res1;
  g.Si[0] = i00; g.Si[1] = i10; g.Si[2] = i01; g.Si[3] = i11;
}

__device__ __forceinline__ void ekf_gate_finish(const UpdateSetup& u, const GatePre& pre, const double* t12,
                                                const double* t3, const double* t4, double cond_max, GateResult& g) {
  double S[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) S[k] = ((t12[k] + t3[k]) + t4[k]) + u.R[k];
  ekf_gate_from_S(pre, S, cond_max, g);
}

// Element k = i + 2j of the raw innovation covariance of one landmark, the way ekf_gate_terms12 /
// ekf_gate_terms34 / ekf_gate_finish evaluate it (same expressions, same order), for a lane that
// owns only that element. Operands are picked with selects so nothing is indexed dynamically.
__device__ __forceinline__ double ekf_gate_S_element(const UpdateSetup& u, const GatePre& g, const double* p,
                                                     const double* pll, int k) {
  const bool i1 = (k & 1) != 0, j1 = (k & 2) != 0;
  const double* HR = g.HR;
  const double HRj0 = j1 ? HR[1] : HR[0], HRj2 = j1 ? HR[3] : HR[2], HRj4 = j1 ? HR[5] : HR[4];
  const double HRi0 = i1 ? HR[1] : HR[0], HRi2 = i1 ? HR[3] : HR[2], HRi4 = i1 ? HR[5] : HR[4];
  const double Cti = i1 ? u.Ct[1] : u.Ct[0], Cti2 = i1 ? u.Ct[3] : u.Ct[2];
  const double Cm0j = j1 ? u.Ct[1] : u.Ct[0], Cm1j = j1 ? u.Ct[3] : u.Ct[2];   // Cm[0+2j], Cm[1+2j]
  // t1: a1[i+2c] = q[i+2c] + h3_i*PRR[2+3c]
  const double a1_0 = (i1 ? u.q[1] : u.q[0]) + HRi4 * u.PRR[2];
  const double a1_2 = (i1 ? u.q[3] : u.q[2]) + HRi4 * u.PRR[5];
  const double a1_4 = (i1 ? u.q[5] : u.q[4]) + HRi4 * u.PRR[8];
  const double t1 = (a1_0 * HRj0 + a1_2 * HRj2) + a1_4 * HRj4;
  // t2: a2[i+2c] = Ct[i]*p[0+2c] + Ct[i+2]*p[1+2c]
  const double a2_0 = Cti * p[0] + Cti2 * p[1];
  const double a2_2 = Cti * p[2] + Cti2 * p[3];
  const double a2_4 = Cti * p[4] + Cti2 * p[5];
  const double t2 = (a2_0 * HRj0 + a2_2 * HRj2) + a2_4 * HRj4;
  const double t12 = t1 + t2;
  // t3: a3[i+2c] = (HR[i]*p[c] + HR[i+2]*p[c+2]) + HR[i+4]*p[c+4], c = 0..1
  const double a3_0 = (HRi0 * p[0] + HRi2 * p[2]) + HRi4 * p[4];
  const double a3_2 = (HRi0 * p[1] + HRi2 * p[3]) + HRi4 * p[5];
  const double t3 = a3_0 * Cm0j + a3_2 * Cm1j;
  // t4: a4[i+2c] = Ct[i]*pll[0+2c] + Ct[i+2]*pll[1+2c]
  const double a4_0 = Cti * pll[0] + Cti2 * pll[1];
  const double a4_2 = Cti * pll[2] + Cti2 * pll[3];
  const double t4 = a4_0 * Cm0j + a4_2 * Cm1j;
  const double Rk = j1 ? (i1 ? u.R[3] : u.R[2]) : (i1 ? u.R[1] : u.R[0]);
  return ((t12 + t3) + t4) + Rk;
}

__device__ __forceinline__ void ekf_gate_landmark(const UpdateSetup& u, double lx, double ly, const double* p,
                                                  const double* pll, double cond_max, GateResult& g) {
  GatePre pre;
  double t12[4], t3[4], t4[4];
  ekf_gate_prelude(u, lx, ly, pre);
  ekf_gate_terms12(u, pre, p, t12);
  ekf_gate_terms34(u, pre, p, pll, t3, t4);
  ekf_gate_finish(u, pre, t12, t3, t4, cond_max, g);
}

__device__ __forceinline__ void ekf_inv2(const double* S, double* Si) {
  const double a = S[0], c = S[1], b = S[2], d = S[3];
  const double det = a * d - b * c;
  const double invdet = 1.0 / det;
  Si[0] = d * invdet;
  Si[1] = -c * invdet;
  Si[2] = -b * invdet;
  Si[3] = a * invdet;
}

// Three-way decision, Update.cpp:152,181,191. 0 = New, 1 = Old, 2 = Ignore.
__device__ __forceinline__ int ekf_decide(int opt_i, double mahal, const EkfConst& k) {
  if (opt_i == 0 || mahal > (double)k.gamma_max) return 0;
  if (mahal < (double)k.gamma_min) return 1;
  return 2;
}

// Odometry read + Q (kalmanfilter.cpp:17-37) and the scalars of Propagate.cpp:33-48.
struct PropSetup {
  double v, w, dt;
  double Q[4];       // column-major
  double c, s;       // cos/sin of the pre-propagation heading
  double phi02, phi12;  // Phi_R(0,2), Phi_R(1,2)
  double g00, g10, g21; // G(0,0), G(1,0), G(2,1)
};

// The part of doPropagation's scalars that depends on the odometry record only (kalmanfilter.cpp:17-37):
// it can be evaluated before the heading is known.
__device__ __forceinline__ void ekf_build_prop_pre(PropSetup& p, double vel_mm_s, double rotvel_deg_s, double dt,
                                                   const EkfConst& k) {
  const double RTV = rotvel_deg_s * k.deg2rad_pi / 180.0;   // kalmanfilter.cpp:19
  p.v = vel_mm_s / 1000.0;                                   // :26
  p.w = RTV;
  p.dt = dt;
  // Q = (v*v)*Q0*Q0 with Q0 = [sigma_v 0; 0 sigma_w] (kalmanfilter.cpp:35-37)
  const double vv = p.v * p.v;
  const double Q0[4] = {k.sigma_v, 0.0, 0.0, k.sigma_w};
  double A[4];
  for (int i = 0; i < 4; ++i) A[i] = vv * Q0[i];
  for (int j = 0; j < 2; ++j)
    for (int i = 0; i < 2; ++i) p.Q[i + 2 * j] = A[i] * Q0[0 + 2 * j] + A[i + 2] * Q0[1 + 2 * j];
}
// The heading-dependent part (Propagate.cpp:42-48); (s, c) = sincos of the pre-propagation heading.
__device__ __forceinline__ void ekf_build_prop_trig(PropSetup& p, double s, double c) {
  const double dt = p.dt;
  p.s = s; p.c = c;
  p.phi02 = -dt * p.v * p.s;   // Propagate.cpp:42
  p.phi12 = dt * p.v * p.c;    // :43
  p.g00 = -dt * p.c;           // :46
  p.g10 = -dt * p.s;           // :47
  p.g21 = -dt;                 // :48
}
// (s, c) = sincos of the pre-propagation heading, computed by the caller.
__device__ __forceinline__ void ekf_build_prop_sc(PropSetup& p, double vel_mm_s, double rotvel_deg_s, double dt,
                                                  double s, double c, const EkfConst& k) {
  ekf_build_prop_pre(p, vel_mm_s, rotvel_deg_s, dt, k);
  ekf_build_prop_trig(p, s, c);
}

__device__ __forceinline__ void ekf_build_prop(PropSetup& p, double vel_mm_s, double rotvel_deg_s, double dt,
                                               double ori, const EkfConst& k) {
  double s, c;
  sincos(ori, &s, &c);
  ekf_build_prop_sc(p, vel_mm_s, rotvel_deg_s, dt, s, c, k);
}

// One element M(i,j) of Phi*P_RR*Phi^T + G*Q*G^T (Propagate.cpp:53), with Phi (3x3), G (3x2),
// Q (2x2) and the old P_RR (3x3) given column-major. The same expression whether one thread
// evaluates all nine elements or nine lanes evaluate one each.
__device__ __forceinline__ double ekf_prop_prr_elem(const double* Phi, const double* G, const double* Q,
                                                    const double* PRR, int i, int j) {
  // T1(i,k) = (Phi*PRR)(i,k), k = 0..2
  const double t10 = (Phi[i] * PRR[0] + Phi[i + 3] * PRR[1]) + Phi[i + 6] * PRR[2];
  const double t11 = (Phi[i] * PRR[3] + Phi[i + 3] * PRR[4]) + Phi[i + 6] * PRR[5];
  const double t12 = (Phi[i] * PRR[6] + Phi[i + 3] * PRR[7]) + Phi[i + 6] * PRR[8];
  // T2(i,j) = sum_k T1(i,k)*Phi(j,k)
  const double t2 = (t10 * Phi[j] + t11 * Phi[j + 3]) + t12 * Phi[j + 6];
  // T3(i,c) = (G*Q)(i,c); T4(i,j) = T3(i,0)*G(j,0) + T3(i,1)*G(j,1)
  const double t30 = G[i] * Q[0] + G[i + 3] * Q[1];
  const double t31 = G[i] * Q[2] + G[i + 3] * Q[3];
  const double t4 = t30 * G[j] + t31 * G[j + 3];
  return t2 + t4;
}

// P_RR <- Phi*P_RR*Phi^T + G*Q*G^T, then the 3x3 part of 0.5*(P+P^T) (Propagate.cpp:53,66-67).
// PRR column-major 3x3, in place.
__device__ __forceinline__ void ekf_prop_prr(const PropSetup& p, double* PRR) {
  const double Phi[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, p.phi02, p.phi12, 1.0};
  const double G[6] = {p.g00, p.g10, 0.0, 0.0, 0.0, p.g21};
  double M[9];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) M[i + 3 * j] = ekf_prop_prr_elem(Phi, G, p.Q, PRR, i, j);
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) PRR[i + 3 * j] = 0.5 * (M[i + 3 * j] + M[j + 3 * i]);
}

// One column of P_RL <- Phi*P_RL (Propagate.cpp:56), literal 3-term sums with the 1/0 entries.
__device__ __forceinline__ void ekf_prop_col(const PropSetup& p, double& a0, double& a1, double& a2) {
  const double o0 = (1.0 * a0 + 0.0 * a1) + p.phi02 * a2;
  const double o1 = (0.0 * a0 + 1.0 * a1) + p.phi12 * a2;
  const double o2 = (0.0 * a0 + 0.0 * a1) + 1.0 * a2;
  a0 = o0; a1 = o1; a2 = o2;
}

// Compass residual selection, kalmanfilter.cpp:98-110.
__device__ __forceinline__ double ekf_compass_residual(double phi, double z, const EkfConst& k) {
  double z_hat = phi;
  z_hat -= k.two_pi * floor(z_hat / k.two_pi);
  const double res1 = z - z_hat;
  const double res2 = z - k.two_pi - z_hat;
  const double res3 = z + k.two_pi - z_hat;
  if ((fabs(res1) <= fabs(res2)) && (fabs(res1) <= fabs(res3))) return res1;
  if (fabs(res2) <= fabs(res3)) return res2;
  return res3;
}
