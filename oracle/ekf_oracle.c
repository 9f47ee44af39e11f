/* TEST INFRASTRUCTURE ONLY: CPU oracle for the EKF-SLAM filter core. Never linked into, imported
 * by, or executed from the product path (2d-ekf-slam_b200/). Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker.
 *
 * Plain-C restatement of kentsommer/2D-EKF-SLAM's filter arithmetic:
 *   odometry/kalmanfilter.cpp:15-62   doPropagation (unit conversion, Q)
 *   odometry/Propagate.cpp:15-75      Propagate
 *   odometry/kalmanfilter.cpp:64-90   doUpdate (Gamma_max=50, Gamma_min=10)
 *   odometry/Update.cpp:22-204        Update (gating, New / Old / Ignore, symmetrise)
 *   odometry/kalmanfilter.cpp:96-130  doUpdateCompass
 *   slam.cpp:158-167                  measurement covariance R from a corner feature
 *
 * PARITY STATUS: pinned by execution, not by published vectors. The reference ships no tests,
 * golden vectors or known-answer values, and its Eigen dependency is un-vendored and un-pinned
 * (kalmanfilter.h:8; Makefile:2). This restatement is therefore pinned against the reference's
 * own translation units compiled unmodified over oracle/shim/ (oracle/_ref/libekf_ref.so) and
 * run in this container: tests/test_oracle_vs_ref.py requires BIT-IDENTICAL state, covariance,
 * decisions and Mahalanobis distances on seeded sequences, and tests/golden/ holds vectors
 * generated from that library by tests/golden/make_golden.py. Relative to a build against real
 * Eigen the only unpinned degrees of freedom are the ones oracle/shim/Eigen/Dense documents
 * (<=3-term summation order, 2x2 SVD / inverse internals: O(1e-16) relative).
 *
 * Arithmetic contract (must match oracle/shim/Eigen/Dense): column-major; C=A*B evaluates
 * C(i,j)=A(i,0)*B(0,j) then C(i,j)=C(i,j)+A(i,k)*B(k,j), k=1..; chained products associate
 * left-to-right; no FMA (build with -ffp-contract=off).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define EKF_INF 999999999999.0 /* kalmanfilter.h:17 */

enum { ORACLE_NEW = 0, ORACLE_OLD = 1, ORACLE_IGNORE = 2 };

typedef struct {
  int32_t decision;
  int32_t opt_i;
  double mahal;
  int32_t n_cond_skipped;
  int32_t k_col;
  double margin_gmin, margin_gmax, margin_cond;
} OracleTrace; /* same layout as RefTrace in oracle/ref_harness.cpp */

/* C(m x n) = A(m x k) * B(k x n), column-major with leading dimensions. */
static void mm(int m, int k, int n, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) {
      double acc = 0.0;
      if (k > 0) {
        acc = A[i] * B[(size_t)j * ldb];
        for (int p = 1; p < k; ++p) acc = acc + A[i + (size_t)p * lda] * B[p + (size_t)j * ldb];
      }
      C[i + (size_t)j * ldc] = acc;
    }
}
static void tr(int m, int n, const double* A, int lda, double* T, int ldt) { /* T = A^T */
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) T[j + (size_t)i * ldt] = A[i + (size_t)j * lda];
}
static void inv2(const double* S, double* Si) { /* shim inverse(): invdet form */
  const double a = S[0], c = S[1], b = S[2], d = S[3];
  const double det = a * d - b * c;
  const double invdet = 1.0 / det;
  Si[0] = d * invdet;
  Si[1] = -c * invdet;
  Si[2] = -b * invdet;
  Si[3] = a * invdet;
}
static double cond2(const double* S) { /* shim JacobiSVD: sigma0/sigma1 */
  const double a = S[0], c = S[1], b = S[2], d = S[3];
  const double E = (a + d) * 0.5, F = (a - d) * 0.5, G = (c + b) * 0.5, H = (c - b) * 0.5;
  const double Q = sqrt(E * E + H * H), R = sqrt(F * F + G * G);
  return (Q + R) / fabs(Q - R);
}
/* P <- 0.5*(P + P^T), literally over the whole n x n matrix (Propagate.cpp:66-67, Update.cpp:193-194,
 * kalmanfilter.cpp:123-124). */
static void symmetrise(int n, double* P, int ld) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i <= j; ++i) {
      const double a = P[i + (size_t)j * ld], b = P[j + (size_t)i * ld];
      const double u = 0.5 * (a + b), l = 0.5 * (b + a);
      P[i + (size_t)j * ld] = u;
      P[j + (size_t)i * ld] = l;
    }
}

/* kalmanfilter.cpp:15-48 + Propagate.cpp:15-75. x: n, P: n x n (ld), in place. */
void ekf_oracle_propagate(int n, double* x, double* P, int ld, double vel_mm_s, double rotvel_deg_s, double dt) {
  const double RTV = rotvel_deg_s * 3.141592654 / 180.0; /* kalmanfilter.cpp:19 */
  const double v = vel_mm_s / 1000.0;                    /* :26 */
  const double w = RTV;
  /* Q = (v*v)*Q*Q, Q = diag(sigma_v, sigma_w) (kalmanfilter.cpp:28-37) */
  const double Q0[4] = {0.01, 0.0, 0.0, 0.04};
  double A[4], Q[4];
  const double vv = v * v;
  for (int k = 0; k < 4; ++k) A[k] = vv * Q0[k];
  mm(2, 2, 2, A, 2, Q0, 2, Q, 2);

  const double ori = x[2]; /* Propagate.cpp:19 */
  const double c = cos(ori), s = sin(ori);
  /* Propagate.cpp:33-37 */
  const double xm0 = v * c, xm1 = v * s, xm2 = w;
  x[0] = x[0] + dt * xm0;
  x[1] = x[1] + dt * xm1;
  x[2] = x[2] + dt * xm2;
  /* Propagate.cpp:42-48 (column-major 3x3 and 3x2) */
  const double Phi[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, -dt * v * s, dt * v * c, 1.0};
  const double G[6] = {-dt * c, -dt * s, 0.0, 0.0, 0.0, -dt};
  double PhiT[9], GT[6], T1[9], T2[9], T3[6], T4[9];
  tr(3, 3, Phi, 3, PhiT, 3);
  tr(3, 2, G, 3, GT, 2);
  /* P_RR = Phi*P_RR*Phi^T + G*Q*G^T (Propagate.cpp:53) */
  double PRR[9];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) PRR[i + 3 * j] = P[i + (size_t)j * ld];
  mm(3, 3, 3, Phi, 3, PRR, 3, T1, 3);
  mm(3, 3, 3, T1, 3, PhiT, 3, T2, 3);
  mm(3, 2, 2, G, 3, Q, 2, T3, 3);
  mm(3, 2, 3, T3, 3, GT, 2, T4, 3);
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) P[i + (size_t)j * ld] = T2[i + 3 * j] + T4[i + 3 * j];
  /* P_RL = Phi*P_RL ; P_LR = P_RL^T (Propagate.cpp:56-60); P_LL unchanged (:63) */
  for (int j = 3; j < n; ++j) {
    double col[3], out[3];
    for (int i = 0; i < 3; ++i) col[i] = P[i + (size_t)j * ld];
    mm(3, 3, 1, Phi, 3, col, 3, out, 3);
    for (int i = 0; i < 3; ++i) {
      P[i + (size_t)j * ld] = out[i];
      P[j + (size_t)i * ld] = out[i];
    }
  }
  symmetrise(n, P, ld); /* Propagate.cpp:66-67 */
}

/* slam.cpp:158-167. R_out column-major. */
void ekf_oracle_measurement_from_feature(double fx_mm, double fy_mm, double* z_out, double* R_out) {
  const double fx = fx_mm / 1000.0, fy = fy_mm / 1000.0;
  const double dist = sqrt(fx * fx + fy * fy);
  const double bearing = atan2(fy, fx);
  const double R[4] = {0.0025, 0.0, 0.0, 0.0001};
  const double G[4] = {cos(bearing), sin(bearing), -dist * sin(bearing), dist * cos(bearing)};
  double GT[4], T[4];
  tr(2, 2, G, 2, GT, 2);
  mm(2, 2, 2, G, 2, R, 2, T, 2);
  mm(2, 2, 2, T, 2, GT, 2, R_out, 2);
  z_out[0] = fx;
  z_out[1] = fy;
}

/* Update.cpp:80-195, the body of the j-loop for ONE measurement. n_lm is the gating loop's bound:
 * Update.cpp:26 reads it once per call, so for j > 1 of an n_z > 1 call it is the landmark count at
 * call entry, not (n-3)/2. x: capacity >= n+2, P: ld >= n+2, in place. R column-major 2x2.
 * Returns the new dimension (n or n+2), or -1 if a New would exceed cap_n (state untouched). */
static int update_one(int n, int n_lm, double* x, double* P, int ld, int cap_n, const double* z, const double* R,
                      int gamma_max, int gamma_min, OracleTrace* trace) {
  const double phi = x[2];
  const double cphi = cos(phi), sphi = sin(phi);
  const double C[4] = {cphi, sphi, -sphi, cphi};   /* C << cos,-sin,sin,cos (row-major fill) */
  const double J[4] = {0.0, 1.0, -1.0, 0.0};       /* J << 0,-1,1,0 */
  double Ct[4];
  tr(2, 2, C, 2, Ct, 2);                           /* H_Li = C^T (Update.cpp:95) */
  const double* HLi = Ct;
  double HLiT[4];
  tr(2, 2, HLi, 2, HLiT, 2);
  double mCt[4];
  for (int k = 0; k < 4; ++k) mCt[k] = -1.0 * Ct[k]; /* -1.0*C^T (Update.cpp:113) */
  double mCtJ[4];
  mm(2, 2, 2, mCt, 2, J, 2, mCtJ, 2);              /* (-1.0*C^T)*J (Update.cpp:114) */
  double PRR[9];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) PRR[i + 3 * j] = P[i + (size_t)j * ld];

  double mahal = EKF_INF;
  int opt_i = 0, skipped = 0;
  double opt_res[2] = {0, 0}, opt_S[4] = {0, 0, 0, 0}, opt_HR[6] = {0, 0, 0, 0, 0, 0};
  double m_cond = 1e300;

  for (int i = 1; i <= n_lm; ++i) { /* Update.cpp:103-148 */
    const int Li = 2 * i + 1;
    const double d[2] = {x[Li] - x[0], x[Li + 1] - x[1]};
    double zhat[2], res[2], HR[6], h3[2];
    mm(2, 2, 1, Ct, 2, d, 2, zhat, 2);
    res[0] = z[0] - zhat[0];
    res[1] = z[1] - zhat[1];
    HR[0] = mCt[0]; HR[1] = mCt[1]; HR[2] = mCt[2]; HR[3] = mCt[3];
    mm(2, 2, 1, mCtJ, 2, d, 2, h3, 2);
    HR[4] = h3[0]; HR[5] = h3[1];
    double HRT[6];
    tr(2, 3, HR, 2, HRT, 3);
    double PRLi[6], PLiR[6], PLiLi[4];
    for (int jj = 0; jj < 2; ++jj)
      for (int ii = 0; ii < 3; ++ii) PRLi[ii + 3 * jj] = P[ii + (size_t)(Li + jj) * ld];
    for (int jj = 0; jj < 3; ++jj)
      for (int ii = 0; ii < 2; ++ii) PLiR[ii + 2 * jj] = P[(Li + ii) + (size_t)jj * ld];
    for (int jj = 0; jj < 2; ++jj)
      for (int ii = 0; ii < 2; ++ii) PLiLi[ii + 2 * jj] = P[(Li + ii) + (size_t)(Li + jj) * ld];
    /* S = H_R*P_RR*H_R^T + H_Li*P_LiR*H_R^T + H_R*P_RLi*H_Li^T + H_Li*P_LiLi*H_Li^T + R  (:122) */
    double a1[6], t1[4], a2[6], t2[4], a3[4], t3[4], a4[4], t4[4], S[4];
    mm(2, 3, 3, HR, 2, PRR, 3, a1, 2);
    mm(2, 3, 2, a1, 2, HRT, 3, t1, 2);
    mm(2, 2, 3, HLi, 2, PLiR, 2, a2, 2);
    mm(2, 3, 2, a2, 2, HRT, 3, t2, 2);
    mm(2, 3, 2, HR, 2, PRLi, 3, a3, 2);
    mm(2, 2, 2, a3, 2, HLiT, 2, t3, 2);
    mm(2, 2, 2, HLi, 2, PLiLi, 2, a4, 2);
    mm(2, 2, 2, a4, 2, HLiT, 2, t4, 2);
    for (int k = 0; k < 4; ++k) S[k] = (((t1[k] + t2[k]) + t3[k]) + t4[k]) + R[k];
    { /* S = 0.5*(S + S^T) (:123-124) */
      const double s01 = 0.5 * (S[2] + S[1]), s10 = 0.5 * (S[1] + S[2]);
      S[0] = 0.5 * (S[0] + S[0]);
      S[3] = 0.5 * (S[3] + S[3]);
      S[2] = s01;
      S[1] = s10;
    }
    const double cond = cond2(S); /* :127-128 */
    {
      const double dc = fabs(cond - 80.0);
      if (dc < m_cond) m_cond = dc;
    }
    if (cond >= 80) { ++skipped; continue; } /* :131 */
    double Si[4], r1[2], temp;
    inv2(S, Si);
    /* temp = res^T * S^-1 * res (:136): (1x2 * 2x2) * 2x1 */
    r1[0] = res[0] * Si[0] + res[1] * Si[1];
    r1[1] = res[0] * Si[2] + res[1] * Si[3];
    temp = r1[0] * res[0] + r1[1] * res[1];
    if (mahal > temp) { /* :140-147 */
      mahal = temp;
      opt_i = Li;
      opt_res[0] = res[0]; opt_res[1] = res[1];
      memcpy(opt_S, S, sizeof S);
      memcpy(opt_HR, HR, sizeof HR);
    }
  }

  int decision, n_out = n, k_col = -1;
  if (opt_i == 0 || mahal > gamma_max) { /* New: Update.cpp:152-178 */
    decision = ORACLE_NEW;
    if (n + 2 > cap_n) return -1;
    double Cz[2], nl[2], dn[2], HR[6], h3[2], HRT[6];
    mm(2, 2, 1, C, 2, z, 2, Cz, 2);
    nl[0] = x[0] + Cz[0];
    nl[1] = x[1] + Cz[1];
    dn[0] = nl[0] - x[0];
    dn[1] = nl[1] - x[1];
    HR[0] = mCt[0]; HR[1] = mCt[1]; HR[2] = mCt[2]; HR[3] = mCt[3];
    mm(2, 2, 1, mCtJ, 2, dn, 2, h3, 2);
    HR[4] = h3[0]; HR[5] = h3[1];
    tr(2, 3, HR, 2, HRT, 3);
    /* P_LiLi = H_Li^T * (H_R*P_RR*H_R^T + R) * H_Li (:168) */
    double a1[6], t1[4], in[4], b1[4], PLL[4];
    mm(2, 3, 3, HR, 2, PRR, 3, a1, 2);
    mm(2, 3, 2, a1, 2, HRT, 3, t1, 2);
    for (int k = 0; k < 4; ++k) in[k] = t1[k] + R[k];
    mm(2, 2, 2, HLiT, 2, in, 2, b1, 2);
    mm(2, 2, 2, b1, 2, HLi, 2, PLL, 2);
    /* P_RLi = -P[:,0:3] * H_R^T * H_Li (:169), n x 2 */
    for (int i = 0; i < n; ++i) {
      double row[3], t[2], o[2];
      for (int k = 0; k < 3; ++k) row[k] = -P[i + (size_t)k * ld];
      mm(1, 3, 2, row, 1, HRT, 3, t, 1);
      mm(1, 2, 2, t, 1, HLi, 2, o, 1);
      P[i + (size_t)n * ld] = o[0];
      P[i + (size_t)(n + 1) * ld] = o[1];
      P[n + (size_t)i * ld] = o[0];
      P[(n + 1) + (size_t)i * ld] = o[1];
    }
    for (int jj = 0; jj < 2; ++jj)
      for (int ii = 0; ii < 2; ++ii) P[(n + ii) + (size_t)(n + jj) * ld] = PLL[ii + 2 * jj];
    x[n] = nl[0];
    x[n + 1] = nl[1];
    n_out = n + 2;
  } else if (mahal < gamma_min) { /* Old: Update.cpp:181-189 */
    decision = ORACLE_OLD;
    k_col = opt_i;
    double HRT[6], Si[4];
    tr(2, 3, opt_HR, 2, HRT, 3);
    inv2(opt_S, Si);
    double* K = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    double* W = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    for (int i = 0; i < n; ++i) {
      double r3[3], r2[2], A[2], B[2], M[2], k2[2];
      for (int k = 0; k < 3; ++k) r3[k] = P[i + (size_t)k * ld];
      for (int k = 0; k < 2; ++k) r2[k] = P[i + (size_t)(opt_i + k) * ld];
      mm(1, 3, 2, r3, 1, HRT, 3, A, 1);
      mm(1, 2, 2, r2, 1, HLiT, 2, B, 1);
      M[0] = A[0] + B[0];
      M[1] = A[1] + B[1];
      mm(1, 2, 2, M, 1, Si, 2, k2, 1);
      K[i] = k2[0];
      K[i + n] = k2[1];
    }
    for (int i = 0; i < n; ++i) { /* x = x + K*res (:187) */
      const double kr = K[i] * opt_res[0] + K[i + n] * opt_res[1];
      x[i] = x[i] + kr;
    }
    mm(n, 2, 2, K, n, opt_S, 2, W, n); /* K*S */
    for (int j = 0; j < n; ++j)        /* P = P - (K*S)*K^T (:188) */
      for (int i = 0; i < n; ++i) {
        const double t = W[i] * K[j] + W[i + n] * K[j + n];
        P[i + (size_t)j * ld] = P[i + (size_t)j * ld] - t;
      }
    free(K);
    free(W);
  } else {
    decision = ORACLE_IGNORE; /* :191 */
  }
  symmetrise(n_out, P, ld); /* :193-194 */

  if (trace) {
    trace->decision = decision;
    trace->opt_i = opt_i;
    trace->mahal = mahal;
    trace->n_cond_skipped = skipped;
    trace->k_col = k_col;
    trace->margin_gmin = fabs(mahal - 10.0);
    trace->margin_gmax = fabs(mahal - 50.0);
    trace->margin_cond = m_cond;
  }
  return n_out;
}

/* One doUpdate call with a single measurement (what slam.cpp:150-171 issues per feature). */
int ekf_oracle_update(int n, double* x, double* P, int ld, int cap_n, const double* z, const double* R,
                      int gamma_max, int gamma_min, OracleTrace* trace) {
  return update_one(n, (n - 3) / 2, x, P, ld, cap_n, z, R, gamma_max, gamma_min, trace);
}

/* One doUpdate call with n_z measurements (z_chunk 2 x n_z, R_chunk 2 x 2n_z, both column-major):
 * sequential, but with the gating bound frozen at call entry (Update.cpp:26) - a landmark added by
 * measurement j is not a candidate for measurements j+1..n_z of the same call. traces: n_z entries
 * or NULL. Returns the new dimension, or -1 on capacity (state as left by the measurements before). */
int ekf_oracle_update_chunk(int n, double* x, double* P, int ld, int cap_n, int n_z, const double* z_chunk,
                            const double* R_chunk, int gamma_max, int gamma_min, OracleTrace* traces) {
  const int n_lm = (n - 3) / 2;
  for (int j = 0; j < n_z; ++j) {
    const int n2 = update_one(n, n_lm, x, P, ld, cap_n, z_chunk + 2 * j, R_chunk + 4 * j, gamma_max, gamma_min,
                              traces ? traces + j : NULL);
    if (n2 < 0) return -1;
    n = n2;
  }
  return n;
}

/* kalmanfilter.cpp:96-130 */
void ekf_oracle_update_compass(int n, double* x, double* P, int ld, double z, double R) {
  double z_hat = x[2];
  z_hat -= 6.283185307 * floor(z_hat / 6.283185307);
  const double res1 = z - z_hat;
  const double res2 = z - 6.283185307 - z_hat;
  const double res3 = z + 6.283185307 - z_hat;
  double res;
  if ((fabs(res1) <= fabs(res2)) && (fabs(res1) <= fabs(res3))) res = res1;
  else if (fabs(res2) <= fabs(res3)) res = res2;
  else res = res3;
  const double S = P[2 + (size_t)2 * ld] + R;
  const double invS = 1 / S;
  double* K = (double*)malloc(sizeof(double) * (size_t)n);
  double* SK = (double*)malloc(sizeof(double) * (size_t)n);
  for (int i = 0; i < n; ++i) K[i] = invS * P[i + (size_t)2 * ld];
  for (int i = 0; i < n; ++i) x[i] = x[i] + res * K[i];
  for (int i = 0; i < n; ++i) SK[i] = S * K[i];
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) P[i + (size_t)j * ld] = P[i + (size_t)j * ld] - SK[i] * K[j];
  symmetrise(n, P, ld);
  free(K);
  free(SK);
}

/* ---- sequence / batch driver over the shared step-record format --------------------------------
 * record (doubles): [0] vel_mm_s [1] rotvel_deg_s [2] dt [3] compass_z [4] compass_R [5] n_z
 * [6] has_compass [7] 0, then max_meas x {z0,z1,R00,R10,R01,R11}; inputs [F][T][8+6*max_meas]. */
typedef struct {
  int n_filters, n_steps, max_meas, cap_lm;
  const double* inputs;
  int32_t *decision, *index, *final_nlm;
  double *mahal, *pose_trace, *final_pose, *final_x, *final_P;
  int final_ld;
  int warm_steps;
  volatile int next;
  volatile int bad;
  double slowest;
  pthread_mutex_t mu;
} BatchJob;

static double now_s(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static void* batch_worker(void* arg) {
  BatchJob* jb = (BatchJob*)arg;
  const long L = 8 + 6L * jb->max_meas;
  const int cap_n = 3 + 2 * jb->cap_lm;
  double* x = (double*)malloc(sizeof(double) * (size_t)cap_n);
  double* P = (double*)malloc(sizeof(double) * (size_t)cap_n * cap_n);
  double my_secs = 0.0;
  for (;;) {
    pthread_mutex_lock(&jb->mu);
    const int f = jb->next++;
    pthread_mutex_unlock(&jb->mu);
    if (f >= jb->n_filters) break;
    int n = 3;
    memset(x, 0, sizeof(double) * (size_t)cap_n);
    memset(P, 0, sizeof(double) * (size_t)cap_n * cap_n);
    double tstart = now_s();
    for (int t = 0; t < jb->n_steps; ++t) {
      if (t == jb->warm_steps) tstart = now_s();
      const double* rec = jb->inputs + ((long)f * jb->n_steps + t) * L;
      ekf_oracle_propagate(n, x, P, cap_n, rec[0], rec[1], rec[2]);
      if (rec[6] != 0.0) ekf_oracle_update_compass(n, x, P, cap_n, rec[3], rec[4]);
      const int nz = (int)rec[5];
      for (int m = 0; m < jb->max_meas; ++m) {
        const long o = ((long)f * jb->n_steps + t) * jb->max_meas + m;
        if (m < nz) {
          OracleTrace tr_;
          const double* zr = rec + 8 + 6 * m;
          const int n2 = ekf_oracle_update(n, x, P, cap_n, cap_n, zr, zr + 2, 50, 10, &tr_);
          if (n2 < 0) {
            pthread_mutex_lock(&jb->mu);
            jb->bad++;
            pthread_mutex_unlock(&jb->mu);
            if (jb->decision) jb->decision[o] = 3;
            if (jb->index) jb->index[o] = -1;
            if (jb->mahal) jb->mahal[o] = 0.0;
            continue;
          }
          if (jb->decision) jb->decision[o] = tr_.decision;
          if (jb->index) jb->index[o] = tr_.decision == ORACLE_NEW ? n : tr_.opt_i;
          if (jb->mahal) jb->mahal[o] = tr_.mahal;
          n = n2;
        } else {
          if (jb->decision) jb->decision[o] = -1;
          if (jb->index) jb->index[o] = -1;
          if (jb->mahal) jb->mahal[o] = 0.0;
        }
      }
      if (jb->pose_trace) memcpy(jb->pose_trace + ((long)f * jb->n_steps + t) * 3, x, 3 * sizeof(double));
    }
    if (jb->n_steps > jb->warm_steps) my_secs += now_s() - tstart;
    if (jb->final_pose) memcpy(jb->final_pose + 3L * f, x, 3 * sizeof(double));
    if (jb->final_nlm) jb->final_nlm[f] = (n - 3) / 2;
    if (jb->final_x && jb->final_P && n <= jb->final_ld) {
      double* fx = jb->final_x + (long)f * jb->final_ld;
      double* fP = jb->final_P + (long)f * jb->final_ld * jb->final_ld;
      for (int i = 0; i < n; ++i) fx[i] = x[i];
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) fP[i + (long)j * jb->final_ld] = P[i + (size_t)j * cap_n];
    }
  }
  pthread_mutex_lock(&jb->mu);
  if (my_secs > jb->slowest) jb->slowest = my_secs;
  pthread_mutex_unlock(&jb->mu);
  free(x);
  free(P);
  return NULL;
}

/* Same contract as ref_run_batch in oracle/ref_harness.cpp (returns the slowest worker's time over
 * the steps t >= warm_steps), plus a landmark capacity. */
double ekf_oracle_run_batch(int n_filters, int n_steps, int max_meas, int cap_lm, const double* inputs,
                            int n_threads, int32_t* decision, int32_t* index, double* mahal, double* pose_trace,
                            double* final_pose, int32_t* final_nlm, double* final_x, double* final_P,
                            int final_ld, int warm_steps) {
  BatchJob jb;
  memset(&jb, 0, sizeof jb);
  jb.n_filters = n_filters; jb.n_steps = n_steps; jb.max_meas = max_meas; jb.cap_lm = cap_lm;
  jb.inputs = inputs; jb.decision = decision; jb.index = index; jb.mahal = mahal;
  jb.pose_trace = pose_trace; jb.final_pose = final_pose; jb.final_nlm = final_nlm;
  jb.final_x = final_x; jb.final_P = final_P; jb.final_ld = final_ld;
  jb.warm_steps = warm_steps;
  pthread_mutex_init(&jb.mu, NULL);
  if (n_threads < 1) n_threads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
  for (int i = 1; i < n_threads; ++i) pthread_create(&th[i], NULL, batch_worker, &jb);
  batch_worker(&jb);
  for (int i = 1; i < n_threads; ++i) pthread_join(th[i], NULL);
  free(th);
  pthread_mutex_destroy(&jb.mu);
  const double secs = jb.slowest;
  return jb.bad ? -secs : secs;
}
