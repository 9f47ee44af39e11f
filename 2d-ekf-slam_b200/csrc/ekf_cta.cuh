// ekf_cta.cuh — CTA-cooperative EKF-SLAM operations: one thread block owns one filter.
//
// The covariance P (column-major, leading dimension ld) is addressed through a plain pointer:
// the fused batch kernel passes shared memory (covariance resident on chip for all steps), the
// per-call kernels pass global memory. x lives in shared memory in both cases.
//
// Reference semantics: odometry/Propagate.cpp:15-75, odometry/Update.cpp:22-204,
// odometry/kalmanfilter.cpp:96-130. The O(n) parts (gating, gain, augmentation, strip
// propagation) keep the reference's operation order (ekf_small.cuh). The O(n^2) covariance
// downdate  P <- sym(P - K S K^T)  is evaluated as  P_ij - sum_a s_a W_ia W_ja  with
// K S K^T = (K L) D (K L)^T (L D L^T = S, W_a = sqrt|d_a| (K L)_a, s_a = sign d_a): two fma()
// per element whose result is bit-symmetric by construction (products commute), so the
// reference's separate 0.5*(P+P^T) pass (Update.cpp:193-194) is not needed. It differs from
// the reference's  0.5*((P_ij - T_ij) + (P_ij - T_ji))  by O(eps |T_ij|).
#pragma once
#include <climits>
#include <cmath>
#include "ekf_small.cuh"

#define EKF_DEC_NONE (-1)
#define EKF_DEC_NEW 0
#define EKF_DEC_OLD 1
#define EKF_DEC_IGNORE 2
#define EKF_DEC_DROPPED 3

struct CtaScratch {
  PropSetup prop;
  UpdateSetup upd;
  double red_val[32];
  int red_idx[32];
  double bc_val;
  int bc_idx;
  // winner of the gating loop (Opt_res, Opt_S, Opt_H_R third column)
  double res[2], S[4], h3[2];
  // New branch
  double nl[2], PLL[4], h3n[2];
  // downdate
  double m0, m1;
  // compass
  double cres, cS;
};

// (value, index) lexicographic minimum over the CTA; every thread gets the result.
// Reproduces the sequential strict-'>' rule of Update.cpp:140: lowest index wins ties.
__device__ __forceinline__ void cta_argmin(double& val, int& idx, CtaScratch* sc) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, val, o);
    const int oi = __shfl_down_sync(0xffffffffu, idx, o);
    if (ov < val || (ov == val && oi < idx)) { val = ov; idx = oi; }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sc->red_val[w] = val; sc->red_idx[w] = idx; }
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    double v = l < nw ? sc->red_val[l] : INFINITY;
    int i = l < nw ? sc->red_idx[l] : INT_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, v, o);
      const int oi = __shfl_down_sync(0xffffffffu, i, o);
      if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    if (l == 0) { sc->bc_val = v; sc->bc_idx = i; }
  }
  __syncthreads();
  val = sc->bc_val;
  idx = sc->bc_idx;
}

// P_ij <- P_ij + u_i0*W_j0 + u_i1*W_j1 for i,j < n, with u = (m0*W_0, m1*W_1), m = -sign(d).
// Ws[i] = (W_i0, W_i1) in shared memory. Each lane owns one row of a 32-row chunk and walks the
// columns: P accesses are contiguous across the warp, W_j is a broadcast load.
template <int RANK>
__device__ __forceinline__ void cta_downdate(double* __restrict__ P, int ld, int n, const double2* __restrict__ Ws,
                                             double m0, double m1) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int nchunks = (n + 31) >> 5;
  constexpr int CB = 8;  // columns per work item
  const int ncb = (n + CB - 1) / CB;
  for (int item = warp; item < nchunks * ncb; item += nwarps) {
    const int chunk = item % nchunks, cb = item / nchunks;
    const int i = chunk * 32 + lane;
    if (i < n) {
      const double2 wi = Ws[i];
      const double u0 = m0 * wi.x, u1 = m1 * wi.y;
      double* col = P + i + (size_t)(cb * CB) * ld;
      const int jn = min(CB, n - cb * CB);
      if (jn == CB) {
        double v[CB];
#pragma unroll
        for (int j = 0; j < CB; ++j) v[j] = col[(size_t)j * ld];
#pragma unroll
        for (int j = 0; j < CB; ++j) {
          const double2 wj = Ws[cb * CB + j];
          double t = v[j];
          if (RANK == 2) t = fma(u1, wj.y, t);
          t = fma(u0, wj.x, t);
          col[(size_t)j * ld] = t;
        }
      } else {
        for (int j = 0; j < jn; ++j) {
          const double2 wj = Ws[cb * CB + j];
          double t = col[(size_t)j * ld];
          if (RANK == 2) t = fma(u1, wj.y, t);
          t = fma(u0, wj.x, t);
          col[(size_t)j * ld] = t;
        }
      }
    }
  }
}

// KalmanFilter::doPropagation + Propagate (kalmanfilter.cpp:15-48, Propagate.cpp:15-75).
// Ends with a barrier: P and xs are consistent for every thread on return.
__device__ __forceinline__ void cta_propagate(double* __restrict__ P, int ld, double* __restrict__ xs, int n,
                                              double vel_mm_s, double rotvel_deg_s, double dt, CtaScratch* sc,
                                              const EkfConst& k) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) {
    PropSetup p;
    ekf_build_prop(p, vel_mm_s, rotvel_deg_s, dt, xs[2], k);
    sc->prop = p;
    // Propagate.cpp:33-37
    const double xm0 = p.v * p.c, xm1 = p.v * p.s, xm2 = p.w;
    xs[0] = xs[0] + dt * xm0;
    xs[1] = xs[1] + dt * xm1;
    xs[2] = xs[2] + dt * xm2;
  }
  __syncthreads();
  if (tid == nt - 1) {  // 3x3 robot block (a different warp than the strip threads when n is small)
    double PRR[9];
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) PRR[i + 3 * j] = P[i + (size_t)j * ld];
    ekf_prop_prr(sc->prop, PRR);
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) P[i + (size_t)j * ld] = PRR[i + 3 * j];
  }
  // P_RL <- Phi*P_RL, P_LR <- P_RL^T (Propagate.cpp:56-60). Read the strip through its mirror
  // P(j,0..2) (contiguous across threads), write both.
  for (int j = 3 + tid; j < n; j += nt) {
    double a0 = P[j], a1 = P[j + (size_t)ld], a2 = P[j + (size_t)2 * ld];
    ekf_prop_col(sc->prop, a0, a1, a2);
    P[j] = a0;
    P[j + (size_t)ld] = a1;
    P[j + (size_t)2 * ld] = a2;
    double* c = P + (size_t)j * ld;
    c[0] = a0; c[1] = a1; c[2] = a2;
  }
  __syncthreads();
}

// KalmanFilter::doUpdateCompass (kalmanfilter.cpp:96-130).
__device__ __forceinline__ void cta_update_compass(double* __restrict__ P, int ld, double* __restrict__ xs, int n,
                                                   double z, double R, double2* __restrict__ Ws, CtaScratch* sc,
                                                   const EkfConst& k) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) {
    sc->cres = ekf_compass_residual(xs[2], z, k);
    sc->cS = P[2 + (size_t)2 * ld] + R;   // :114
  }
  __syncthreads();
  const double res = sc->cres, S = sc->cS;
  const double invS = 1 / S;
  const double sq = sqrt(fabs(S));
  for (int i = tid; i < n; i += nt) {
    const double Ki = invS * P[i + (size_t)2 * ld];   // :118
    xs[i] = xs[i] + res * Ki;                         // :121
    Ws[i] = make_double2(sq * Ki, 0.0);
  }
  __syncthreads();
  cta_downdate<1>(P, ld, n, Ws, S < 0 ? 1.0 : -1.0, 0.0);   // :122-124
  __syncthreads();
}

struct UpdateOut {
  int decision;   // EKF_DEC_*
  int index;      // Opt_i, or the new landmark's state index for New
  double mahal;   // Mahal_dist after the gating loop
};

// One measurement of KalmanFilter::Update (Update.cpp:80-195). n_lm is updated in place
// (uniformly by every thread). n_gate is the gating loop's bound: Update.cpp:26 reads n_lm ONCE per
// call, so inside one doUpdate(z_chunk) with n_z > 1 a landmark added by measurement j is not a
// candidate for j+1..n_z (n_gate = n_lm at call entry); separate calls pass n_gate = n_lm.
// Returns the same UpdateOut in every thread.
__device__ __forceinline__ UpdateOut cta_update(double* __restrict__ P, int ld, double* __restrict__ xs, int& n_lm,
                                                int n_gate, int cap_lm, double z0, double z1, const double* Rm,
                                                double2* __restrict__ Ws, CtaScratch* sc, const EkfConst& k) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n = 3 + 2 * n_lm;
  if (tid == 0) {
    double PRR[9];
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) PRR[i + 3 * j] = P[i + (size_t)j * ld];
    UpdateSetup u;
    ekf_build_setup(u, xs[2], xs[0], xs[1], PRR, z0, z1, Rm);
    sc->upd = u;
  }
  __syncthreads();

  // ---- gating loop, Update.cpp:103-148: one landmark per thread --------------------------------
  double best = INFINITY;
  int best_idx = INT_MAX;
  double b_res0 = 0, b_res1 = 0, b_S0 = 0, b_S1 = 0, b_S2 = 0, b_S3 = 0, b_h0 = 0, b_h1 = 0;
  for (int lm = tid; lm < n_gate; lm += nt) {
    const int Li = 3 + 2 * lm;
    double p[6], pll[4];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      p[0 + 2 * j] = P[Li + (size_t)j * ld];
      p[1 + 2 * j] = P[Li + 1 + (size_t)j * ld];
    }
    pll[0] = P[Li + (size_t)Li * ld];
    pll[1] = P[Li + 1 + (size_t)Li * ld];
    pll[2] = P[Li + (size_t)(Li + 1) * ld];
    pll[3] = P[Li + 1 + (size_t)(Li + 1) * ld];
    GateResult g;
    ekf_gate_landmark(sc->upd, xs[Li], xs[Li + 1], p, pll, k.cond_max, g);
    const bool valid = !g.skip && (k.mahal_init > g.d2);   // :131, :140 vs INF
    if (valid && g.d2 < best) {
      best = g.d2; best_idx = Li;
      b_res0 = g.res0; b_res1 = g.res1;
      b_S0 = g.S[0]; b_S1 = g.S[1]; b_S2 = g.S[2]; b_S3 = g.S[3];
      b_h0 = g.h3_0; b_h1 = g.h3_1;
    }
  }
  double val = best;
  int idx = best_idx;
  cta_argmin(val, idx, sc);
  UpdateOut out;
  const int opt_i = (idx == INT_MAX) ? 0 : idx;
  out.mahal = (idx == INT_MAX) ? k.mahal_init : val;
  out.decision = ekf_decide(opt_i, out.mahal, k);
  out.index = opt_i;

  if (out.decision == EKF_DEC_OLD) {
    // ---- Update.cpp:181-189 ------------------------------------------------------------------
    if (best_idx == idx) {   // the thread that gated the winning landmark publishes Opt_*
      sc->res[0] = b_res0; sc->res[1] = b_res1;
      sc->S[0] = b_S0; sc->S[1] = b_S1; sc->S[2] = b_S2; sc->S[3] = b_S3;
      sc->h3[0] = b_h0; sc->h3[1] = b_h1;
    }
    __syncthreads();
    const UpdateSetup& u = sc->upd;
    const double S0 = sc->S[0], S1 = sc->S[1], S2 = sc->S[2], S3 = sc->S[3];
    double Si[4];
    {
      const double Sm[4] = {S0, S1, S2, S3};
      ekf_inv2(Sm, Si);
    }
    const double r0 = sc->res[0], r1 = sc->res[1];
    const double h00 = u.mCt[0], h01 = u.mCt[2], h02 = sc->h3[0];
    const double h10 = u.mCt[1], h11 = u.mCt[3], h12 = sc->h3[1];
    const double c00 = u.Ct[0], c10 = u.Ct[2], c01 = u.Ct[1], c11 = u.Ct[3];   // C = (C^T)^T
    // S = L D L^T
    const double d0 = S0, l = S1 / S0, d1 = S3 - l * S1;
    const double sq0 = sqrt(fabs(d0)), sq1 = sqrt(fabs(d1));
    for (int i = tid; i < n; i += nt) {
      const double p0 = P[i], p1 = P[i + (size_t)ld], p2 = P[i + (size_t)2 * ld];
      const double pa = P[i + (size_t)opt_i * ld], pb = P[i + (size_t)(opt_i + 1) * ld];
      const double A0 = (p0 * h00 + p1 * h01) + p2 * h02;
      const double A1 = (p0 * h10 + p1 * h11) + p2 * h12;
      const double B0 = pa * c00 + pb * c10;
      const double B1 = pa * c01 + pb * c11;
      const double M0 = A0 + B0, M1 = A1 + B1;
      const double K0 = M0 * Si[0] + M1 * Si[1];
      const double K1 = M0 * Si[2] + M1 * Si[3];
      xs[i] = xs[i] + (K0 * r0 + K1 * r1);   // :187
      Ws[i] = make_double2(sq0 * fma(l, K1, K0), sq1 * K1);
    }
    __syncthreads();
    cta_downdate<2>(P, ld, n, Ws, d0 < 0 ? 1.0 : -1.0, d1 < 0 ? 1.0 : -1.0);   // :188,193-194
    __syncthreads();
  } else if (out.decision == EKF_DEC_NEW) {
    // ---- Update.cpp:152-178 ------------------------------------------------------------------
    if (n_lm >= cap_lm) {
      out.decision = EKF_DEC_DROPPED;
      out.index = -1;
      return out;
    }
    const UpdateSetup& u = sc->upd;
    if (tid == 0) {
      const double c = u.c, s = u.s;
      // newLand = G_pR_hat + C*z (:155)
      const double Cz0 = c * z0 + (-s) * z1, Cz1 = s * z0 + c * z1;
      const double nl0 = u.x0 + Cz0, nl1 = u.x1 + Cz1;
      const double dn0 = nl0 - u.x0, dn1 = nl1 - u.x1;
      const double h30 = u.mCtJ[0] * dn0 + u.mCtJ[2] * dn1;
      const double h31 = u.mCtJ[1] * dn0 + u.mCtJ[3] * dn1;
      const double HR[6] = {u.mCt[0], u.mCt[1], u.mCt[2], u.mCt[3], h30, h31};
      double a1[6], t1[4], in[4], b1[4], PLL[4];
      for (int j = 0; j < 3; ++j) {
        a1[0 + 2 * j] = u.q[0 + 2 * j] + h30 * u.PRR[2 + 3 * j];
        a1[1 + 2 * j] = u.q[1 + 2 * j] + h31 * u.PRR[2 + 3 * j];
      }
      for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 2; ++i) t1[i + 2 * j] = (a1[i] * HR[j] + a1[i + 2] * HR[j + 2]) + a1[i + 4] * HR[j + 4];
      for (int q = 0; q < 4; ++q) in[q] = t1[q] + u.R[q];
      const double Cm[4] = {u.Ct[0], u.Ct[2], u.Ct[1], u.Ct[3]};   // C = H_Li^T
      for (int j = 0; j < 2; ++j)       // b1 = H_Li^T * in
        for (int i = 0; i < 2; ++i) b1[i + 2 * j] = Cm[i] * in[0 + 2 * j] + Cm[i + 2] * in[1 + 2 * j];
      for (int j = 0; j < 2; ++j)       // PLL = b1 * H_Li
        for (int i = 0; i < 2; ++i) PLL[i + 2 * j] = b1[i] * u.Ct[0 + 2 * j] + b1[i + 2] * u.Ct[1 + 2 * j];
      sc->nl[0] = nl0; sc->nl[1] = nl1;
      sc->h3n[0] = h30; sc->h3n[1] = h31;
      sc->PLL[0] = PLL[0]; sc->PLL[1] = PLL[1]; sc->PLL[2] = PLL[2]; sc->PLL[3] = PLL[3];
    }
    __syncthreads();
    const double h00 = u.mCt[0], h01 = u.mCt[2], h02 = sc->h3n[0];
    const double h10 = u.mCt[1], h11 = u.mCt[3], h12 = sc->h3n[1];
    const double ct00 = u.Ct[0], ct10 = u.Ct[1], ct01 = u.Ct[2], ct11 = u.Ct[3];
    // P_RLi = -P[:,0:3]*H_R^T*H_Li (:169), written as two new columns and their mirror rows
    for (int i = tid; i < n; i += nt) {
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
    if (tid == 0) {
      const double off = 0.5 * (sc->PLL[2] + sc->PLL[1]);   // :193-194 on the new 2x2 block
      P[n + (size_t)n * ld] = sc->PLL[0];
      P[n + 1 + (size_t)n * ld] = off;
      P[n + (size_t)(n + 1) * ld] = off;
      P[n + 1 + (size_t)(n + 1) * ld] = sc->PLL[3];
      xs[n] = sc->nl[0];
      xs[n + 1] = sc->nl[1];
    }
    out.index = n;
    n_lm += 1;
    __syncthreads();
  }
  return out;
}
