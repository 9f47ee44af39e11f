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
#include "ekf_cta.cuh"
#include "ekf_internal.h"
#include "ekf_pdl.cuh"

namespace {

constexpr int kThreads = 256;

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
};

__device__ __forceinline__ double* filt_P(const LargeArgs& a) { return a.st.P + (size_t)a.f * a.st.slab; }
__device__ __forceinline__ double* filt_x(const LargeArgs& a) { return a.st.x + (size_t)a.f * a.st.xs; }

// ---- propagate ---------------------------------------------------------------------------------
__global__ void large_prop_setup(const LargeArgs a, const double* vel, const double* rot, const double* dt) {
  ekf_pdl_entry();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double* P = filt_P(a);
  double* x = filt_x(a);
  const int ld = a.st.ld;
  PropSetup p;
  ekf_build_prop(p, *vel, *rot, *dt, x[2], a.k);
  a.sm->prop = p;
  const double xm0 = p.v * p.c, xm1 = p.v * p.s, xm2 = p.w;   // Propagate.cpp:33-37
  x[0] = x[0] + p.dt * xm0;
  x[1] = x[1] + p.dt * xm1;
  x[2] = x[2] + p.dt * xm2;
  double PRR[9];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) PRR[i + 3 * j] = P[i + (size_t)j * ld];
  ekf_prop_prr(p, PRR);
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) P[i + (size_t)j * ld] = PRR[i + 3 * j];
}

__global__ void __launch_bounds__(kThreads) large_prop_strip(const LargeArgs a) {
  __shared__ PropSetup ps;
  ekf_pdl_entry();
  if (threadIdx.x == 0) ps = a.sm->prop;
  __syncthreads();
  double* P = filt_P(a);
  const int ld = a.st.ld;
  const int n = 3 + 2 * a.st.nlm[a.f];
  for (int j = 3 + blockIdx.x * kThreads + threadIdx.x; j < n; j += gridDim.x * kThreads) {
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
__device__ __forceinline__ void load_gate_inputs(const double* P, int ld, int Li, double* p, double* pll) {
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

// chunk_pos: 0 = a doUpdate call of its own (gating bound = live landmark count); 1 = first
// measurement of an n_z > 1 call (same bound, and block 0 records it); 2 = later measurement of that
// call: Update.cpp:26 read n_lm once, so landmarks added since the call began are not candidates.
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
      for (int i = 0; i < 3; ++i) PRR[i + 3 * j] = P[i + (size_t)j * ld];
    UpdateSetup u;
    ekf_build_setup(u, x[2], x[0], x[1], PRR, zr[0], zr[1], zr + 2);
    sc.upd = u;
    if (blockIdx.x == 0) a.sm->upd = u;
  }
  __syncthreads();
  double best = INFINITY;
  int best_idx = INT_MAX;
  for (int lm = blockIdx.x * kThreads + threadIdx.x; lm < n_lm; lm += gridDim.x * kThreads) {
    const int Li = 3 + 2 * lm;
    double p[6], pll[4];
    load_gate_inputs(P, ld, Li, p, pll);
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
__global__ void __launch_bounds__(kThreads) large_decide(const LargeArgs a, const double* zr, int* out_decision,
                                                        int* out_index, double* out_mahal) {
  __shared__ CtaScratch sc;
  ekf_pdl_entry();
  double val = INFINITY;
  int idx = INT_MAX;
  for (int c = threadIdx.x; c < a.n_cand; c += kThreads) {
    const double v = a.cand_val[c];
    const int i = a.cand_idx[c];
    if (v < val || (v == val && i < idx)) { val = v; idx = i; }
  }
  cta_argmin(val, idx, &sc);
  if (threadIdx.x != 0) return;
  LargeSmall* sm = a.sm;
  const double* P = filt_P(a);
  const double* x = filt_x(a);
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
    load_gate_inputs(P, ld, opt_i, p, pll);
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
      if (sm->decision == EKF_DEC_NEW && blockIdx.x == 0 && threadIdx.x == 0) a.st.nlm[a.f] = sm->n_lm + 1;
      return;
    }
  }
  double* P = filt_P(a);
  const int ld = a.st.ld;
  const int n = COMPASS ? 3 + 2 * a.st.nlm[a.f] : sm->n;
  const double m0 = COMPASS ? sm->cm0 : sm->m0, m1 = COMPASS ? 0.0 : sm->m1;
  const int n_even = (n + 1) & ~1;
  const int rows_per_panel = 2 * kThreads;
  const int n_panels = (n_even + rows_per_panel - 1) / rows_per_panel;
  const int n_cb = (n + kCB - 1) / kCB;
  const long n_tiles = (long)n_panels * n_cb;
  const double2* __restrict__ W = a.W;
  for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
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
__global__ void large_compass_setup(const LargeArgs a, const double* z, const double* R) {
  ekf_pdl_entry();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double* P = filt_P(a);
  const double* x = filt_x(a);
  LargeSmall* sm = a.sm;
  sm->cres = ekf_compass_residual(x[2], *z, a.k);
  const double S = P[2 + (size_t)2 * a.st.ld] + *R;
  sm->cS = S;
  sm->csq = sqrt(fabs(S));
  sm->cm0 = S < 0 ? 1.0 : -1.0;
  sm->n = 3 + 2 * a.st.nlm[a.f];
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

void launch_propagate(const LargeArgs& a, const EkfState& st, const EkfLargeWork& wk, const double* vel,
                      const double* rot, const double* dt, cudaStream_t s) {
  ekf_launch_pdl(large_prop_setup, 1, 32, 0, s, a, vel, rot, dt);
  ekf_launch_pdl(large_prop_strip, row_grid(st, wk), kThreads, 0, s, a);
}
void launch_update(LargeArgs a, const EkfState& st, const EkfLargeWork& wk, const double* zr, int chunk_pos, int* dec,
                   int* idx, double* mah, EkfLargeTiming* tm, cudaStream_t s) {
  const int gg = gate_grid(st, wk);
  a.n_cand = gg;
  ekf_launch_pdl(large_gate, gg, kThreads, 0, s, a, zr, chunk_pos);
  ekf_launch_pdl(large_decide, 1, kThreads, 0, s, a, zr, dec, idx, mah);
  ekf_launch_pdl(large_gain, row_grid(st, wk), kThreads, 0, s, a);
  const bool sample = tm && tm->used < tm->cap && (tm->seen++ % tm->every) == 0;
  if (sample) cudaEventRecord(tm->ev0[tm->used], s);
  if (wk.use_tma) {
    EkfLargeTmaArgs t{&a.sm->decision, &a.sm->n, &a.sm->m0, &a.sm->m1, a.W, a.st.nlm + a.f, &a.sm->n_lm};
    ekf_large_tma_downdate(t, wk.tmaps + (size_t)a.f * ekf_large_tma_map_bytes(), wk.tma_grid, false, s);
  } else {
    ekf_launch_pdl(large_downdate<2, false>, wk.grid, kThreads, 0, s, a);
  }
  if (sample) cudaEventRecord(tm->ev1[tm->used++], s);
}
void launch_compass(const LargeArgs& a, const EkfState& st, const EkfLargeWork& wk, const double* z, const double* R,
                    cudaStream_t s) {
  ekf_launch_pdl(large_compass_setup, 1, 32, 0, s, a, z, R);
  ekf_launch_pdl(large_compass_gain, row_grid(st, wk), kThreads, 0, s, a);
  if (wk.use_tma) {
    EkfLargeTmaArgs t{nullptr, &a.sm->n, &a.sm->cm0, &a.sm->cm0, a.W, nullptr, nullptr};
    ekf_large_tma_downdate(t, wk.tmaps + (size_t)a.f * ekf_large_tma_map_bytes(), wk.tma_grid, true, s);
  } else {
    ekf_launch_pdl(large_downdate<1, true>, wk.grid, kThreads, 0, s, a);
  }
}

}  // namespace

size_t ekf_large_small_doubles() { return (sizeof(LargeSmall) + 7) / 8; }

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
cudaError_t ekf_large_run(const EkfState& st, int filter, const EkfRunIO& io, const uint8_t* has_compass,
                          const uint8_t* n_z, const EkfConst& k, const EkfLargeWork& wk, EkfLargeTiming* tm,
                          cudaStream_t stream, long long* launches) {
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
