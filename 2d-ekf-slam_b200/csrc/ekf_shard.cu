// ekf_shard.cu — one large map whose covariance is sharded over several GPUs (sm_100a, NVLink P2P).
//
// the only place NVLink helps this path. P is bit-symmetric at every call
// boundary (the reference symmetrises after every operation: Propagate.cpp:66-67,
// Update.cpp:193-194, kalmanfilter.cpp:123-124), so "row i of P" and "column i of P" hold the same
// bits. Shard s therefore stores ALL rows of its own column range [c0, c1) (boundaries on landmark
// pairs) and reads every "row i at the 5 gain columns" (Update.cpp:186) from the top / the Opt_i
// rows of its own column i. What is replicated on every shard: the state vector x, the landmark
// count, the 3x3 robot block P_RR and the small per-operation control block. Per reference call:
//
//   doPropagation (Propagate.cpp:15-75): NO exchange. Every shard advances its replicas of x and
//       P_RR with the same arithmetic; the strip P_RL of the own columns is local; shard 0 also
//       owns columns 0..2 and recomputes the mirror rows from its own copy.
//   doUpdate (Update.cpp:80-195), per measurement, two exchange steps:
//       gate   own landmarks -> the shard's best candidate (value, index and the 12 inputs the
//              decision needs) is STORED INTO EVERY PEER's candidate table over NVLink
//       -- exchange 1 --
//       decide every shard reduces the G candidates identically (lowest index wins ties)
//       gain   gain rows / state entries / downdate vectors W_i of the own rows i, stored into
//              every peer's x and W (the "all-gather of K" is the gain kernel's own epilogue:
//              16 n bytes per shard pair, no separate collective kernel)
//       -- exchange 2 --
//       downdate  the HBM-bound sweep over the own columns only: 16 n^2 / G bytes per shard
//   doUpdateCompass (kalmanfilter.cpp:96-130): setup, exchange, gain (stores into peers),
//       exchange, rank-1 downdate.
// An exchange step of this chain (the per-call functions, runs with EKF_SHARD_LOOKAHEAD=0 or with shards
// that share a device) is stream-ordered: every shard records an event after its producer kernel and
// every other shard's stream waits on it; no kernel spins on a peer.
//
// ekf_sharded_run additionally overlaps the O(n) chain with the sweep (look-ahead, ekf_la.cuh): every
// shard keeps a full replica of the O(n) cache (robot columns + diagonal blocks), so gating and the
// decision need NO exchange at all and run on a side stream while the sweep of the previous operation is
// still streaming the shard's slab; what stays between two sweeps is the gain kernel (own rows, peer
// stores of W and x) and its exchange - flags in peer memory that one-warp kernels poll (bounded), one GPU
// per shard required. Compass updates need no exchange either (every input is in the cache). See
// run_shard_thread_la.
//
// Arithmetic is ekf_small.cuh (the reference's operation order) and the bit-symmetric two-fma
// downdate of ekf_cta.cuh, i.e. results are bit-identical to the single-GPU regime B and
// independent of the shard count.
#include <cuda_runtime.h>

#include <atomic>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "ekf_cta.cuh"
#include "ekf_internal.h"
#include "ekf_la.cuh"
#include "ekf_pdl.cuh"
#include "ekf_slam_b200.h"

namespace {

constexpr int kThreads = 256;
constexpr int kSideThreads = 64;   // side-stream kernels of a look-ahead run: small CTAs that fit beside the sweep's
constexpr int kMaxShards = EKF_SHARDED_MAX_SHARDS;
constexpr int kCB = 8;

struct ShardSmall {
  PropSetup prop;
  UpdateSetup upd;
  int decision, opt_i, n, n_lm;
  int gate_nlm;
  unsigned int gate_done;
  double mahal;
  double res[2], S[4], Si[4], h3[2];
  double l, sq0, sq1, m0, m1;
  double nl[2], PLL[4], h3n[2];
  double cres, cS, csq, cm0;
  double PRR[9];   // replica of P(0:3,0:3), column-major
  double2 Wp[3];   // look-ahead runs: downdate vectors of the pose rows (computed with the decision)
};

// A shard's best gating candidate plus everything the decision needs about it, so that no shard
// reads another shard's covariance (and nobody reads x while its owners are rewriting it).
struct ShardCand {
  double val;
  int idx, pad;
  double lx, ly;
  double p[6], pll[4];
};

struct ShardArgs {
  double* P;          // own slab: element (i, j) at P[i + (j - c0) * ld], c0 <= j < c1
  double* x;          // replica [cap_n]
  int* nlm;           // replica
  int* status;        // replica
  ShardSmall* sm;
  double2* W;         // full-length downdate vectors (gathered)
  double* cand_val;   // per-CTA gating minima (local)
  int* cand_idx;
  ShardCand* xcand;   // [n_shards] candidates of every shard (gathered)
  int ld, c0, c1, shard, n_shards;
  int cap_lm;
  int bounds[kMaxShards + 1];
  double2* W_all[kMaxShards];
  double* x_all[kMaxShards];
  ShardCand* xcand_all[kMaxShards];
  EkfConst k;
  // look-ahead runs: full replica of the O(n) cache (ekf_la.cuh) and every shard's copy of it (filled once
  // per run by peer stores); la = 1 tells the sweep that the New bookkeeping is done elsewhere
  double* strip;
  double* diag;
  int lds, la;
  double* strip_all[kMaxShards];
  double* diag_all[kMaxShards];
  // flags of a look-ahead run, in this shard's memory and written by every shard (peer stores):
  // flags_d[t] = number of operations whose decision shard t has finished, flags_g[t] = ... whose gain
  // kernel (W / x stores into every shard) shard t has finished. gain_count: last-CTA counter of the gain kernels.
  unsigned* flags_d;
  unsigned* flags_g;
  unsigned* gain_count;
  unsigned* flags_d_all[kMaxShards];
  unsigned* flags_g_all[kMaxShards];
  unsigned long long poll_ns;   // how long a poll waits before it gives up (2 s; EKF_SHARD_POLL_MS overrides)
};

__device__ __forceinline__ double* scol(const ShardArgs& a, int j) { return a.P + (size_t)(j - a.c0) * a.ld; }

// ---- propagate ---------------------------------------------------------------------------------
__global__ void shard_prop_setup(const ShardArgs a, const double* in3) {
  ekf_pdl_entry();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double* x = a.x;
  PropSetup p;
  ekf_build_prop(p, in3[0], in3[1], in3[2], x[2], a.k);
  a.sm->prop = p;
  const double xm0 = p.v * p.c, xm1 = p.v * p.s, xm2 = p.w;   // Propagate.cpp:33-37
  x[0] = x[0] + p.dt * xm0;
  x[1] = x[1] + p.dt * xm1;
  x[2] = x[2] + p.dt * xm2;
  double PRR[9];
  for (int q = 0; q < 9; ++q) PRR[q] = a.sm->PRR[q];
  ekf_prop_prr(p, PRR);
  for (int q = 0; q < 9; ++q) a.sm->PRR[q] = PRR[q];
  if (a.c0 == 0)
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) a.P[i + (size_t)j * a.ld] = PRR[i + 3 * j];
}

__global__ void __launch_bounds__(kThreads) shard_prop_strip(const ShardArgs a) {
  ekf_pdl_entry();
  __shared__ PropSetup ps;
  if (threadIdx.x == 0) ps = a.sm->prop;
  __syncthreads();
  const int n = 3 + 2 * a.nlm[0];
  const int g0 = blockIdx.x * kThreads + threadIdx.x, stride = gridDim.x * kThreads;
  // P_RL of the own columns: the three entries on top of column j (Propagate.cpp:56-57)
  const int jlo = a.c0 < 3 ? 3 : a.c0, jhi = a.c1 < n ? a.c1 : n;
  for (int j = jlo + g0; j < jhi; j += stride) {
    double* c = scol(a, j);
    double a0 = c[0], a1 = c[1], a2 = c[2];
    ekf_prop_col(ps, a0, a1, a2);
    c[0] = a0; c[1] = a1; c[2] = a2;
  }
  // P_LR: rows j of columns 0..2 live on shard 0; same inputs (bit-symmetric P), same arithmetic
  if (a.c0 == 0) {
    double* P = a.P;
    const int ld = a.ld;
    for (int j = 3 + g0; j < n; j += stride) {
      double a0 = P[j], a1 = P[j + (size_t)ld], a2 = P[j + (size_t)2 * ld];
      ekf_prop_col(ps, a0, a1, a2);
      P[j] = a0;
      P[j + (size_t)ld] = a1;
      P[j + (size_t)2 * ld] = a2;
    }
  }
}

// ---- update: gating over the own landmarks + exchange 1 ----------------------------------------
__device__ __forceinline__ void shard_gate_inputs(const ShardArgs& a, int Li, double* p, double* pll) {
  const double* ca = scol(a, Li);
  const double* cb = scol(a, Li + 1);
#pragma unroll
  for (int j = 0; j < 3; ++j) {      // P_LiR(r, j) == P_RLi(j, r): the top of the landmark's own columns
    p[0 + 2 * j] = ca[j];
    p[1 + 2 * j] = cb[j];
  }
  pll[0] = ca[Li];
  pll[1] = ca[Li + 1];
  pll[2] = cb[Li];
  pll[3] = cb[Li + 1];
}

// chunk_pos as in ekf_large.cu: 0 separate doUpdate call, 1 first / 2 later measurement of an
// n_z > 1 call (gating bound frozen at call entry, Update.cpp:26).
__global__ void __launch_bounds__(kThreads) shard_gate(const ShardArgs a, const double* zr, int chunk_pos) {
  ekf_pdl_entry();
  __shared__ CtaScratch sc;
  __shared__ bool last;
  const double* x = a.x;
  const int n_lm = chunk_pos == 2 ? a.sm->gate_nlm : a.nlm[0];
  if (chunk_pos == 1 && blockIdx.x == 0 && threadIdx.x == 0) a.sm->gate_nlm = n_lm;
  if (threadIdx.x == 0) {
    UpdateSetup u;
    ekf_build_setup(u, x[2], x[0], x[1], a.sm->PRR, zr[0], zr[1], zr + 2);
    sc.upd = u;
    if (blockIdx.x == 0) a.sm->upd = u;
  }
  __syncthreads();
  const int lm_lo = a.c0 < 3 ? 0 : (a.c0 - 3) / 2;
  const int lm_cap = (a.c1 - 3) / 2;
  const int lm_hi = lm_cap < n_lm ? lm_cap : n_lm;
  double best = INFINITY;
  int best_idx = INT_MAX;
  for (int lm = lm_lo + blockIdx.x * kThreads + threadIdx.x; lm < lm_hi; lm += gridDim.x * kThreads) {
    const int Li = 3 + 2 * lm;
    double p[6], pll[4];
    shard_gate_inputs(a, Li, p, pll);
    GateResult g;
    ekf_gate_landmark(sc.upd, x[Li], x[Li + 1], p, pll, a.k.cond_max, g);
    const bool valid = !g.skip && (a.k.mahal_init > g.d2);
    if (valid && g.d2 < best) { best = g.d2; best_idx = Li; }
  }
  cta_argmin(best, best_idx, &sc);
  if (threadIdx.x == 0) {
    a.cand_val[blockIdx.x] = best;
    a.cand_idx[blockIdx.x] = best_idx;
    __threadfence();
    last = atomicAdd(&a.sm->gate_done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  // the last CTA to finish reduces the shard and stores its candidate into every shard's table
  __threadfence();
  double v = INFINITY;
  int ix = INT_MAX;
  for (int c = threadIdx.x; c < (int)gridDim.x; c += kThreads) {
    const double v2 = __ldcg(a.cand_val + c);
    const int i2 = __ldcg(a.cand_idx + c);
    if (v2 < v || (v2 == v && i2 < ix)) { v = v2; ix = i2; }
  }
  cta_argmin(v, ix, &sc);
  if (threadIdx.x == 0) {
    ShardCand c;
    c.val = v; c.idx = ix; c.pad = 0;
    c.lx = c.ly = 0.0;
    for (int q = 0; q < 6; ++q) c.p[q] = 0.0;
    for (int q = 0; q < 4; ++q) c.pll[q] = 0.0;
    if (ix != INT_MAX) {
      shard_gate_inputs(a, ix, c.p, c.pll);
      c.lx = x[ix];
      c.ly = x[ix + 1];
    }
    for (int t = 0; t < a.n_shards; ++t) a.xcand_all[t][a.shard] = c;
    a.sm->gate_done = 0;
  }
}

// ---- update: decision (every shard, identically) -------------------------------------------------
__global__ void shard_decide(const ShardArgs a, const double* zr, int* out_decision, int* out_index, double* out_mahal) {
  ekf_pdl_entry();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ShardSmall* sm = a.sm;
  double val = INFINITY;
  int idx = INT_MAX, who = 0;
  for (int s = 0; s < a.n_shards; ++s) {     // shards are in column order: strict '<' keeps the lowest index
    const double v = a.xcand[s].val;
    const int i = a.xcand[s].idx;
    if (v < val || (v == val && i < idx)) { val = v; idx = i; who = s; }
  }
  const int n_lm = a.nlm[0];
  const int n = 3 + 2 * n_lm;
  const int opt_i = (idx == INT_MAX) ? 0 : idx;
  const double mahal = (idx == INT_MAX) ? a.k.mahal_init : val;
  int decision = ekf_decide(opt_i, mahal, a.k);
  int index = opt_i;
  const UpdateSetup& u = sm->upd;
  if (decision == EKF_DEC_OLD) {
    const ShardCand& c = a.xcand[who];
    GateResult g;
    ekf_gate_landmark(u, c.lx, c.ly, c.p, c.pll, a.k.cond_max, g);   // same bits as the gating pass
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
    if (n_lm >= a.cap_lm) {
      decision = EKF_DEC_DROPPED;
      index = -1;
      a.status[0] |= 1;
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

// ---- update: gain / augmentation of the own rows + exchange 2 ------------------------------------
__global__ void __launch_bounds__(kThreads) shard_gain(const ShardArgs a) {
  ekf_pdl_entry();
  const ShardSmall* sm = a.sm;
  const int decision = sm->decision;
  if (decision != EKF_DEC_OLD && decision != EKF_DEC_NEW) return;
  const int n = sm->n;
  const UpdateSetup& u = sm->upd;
  const int g0 = blockIdx.x * kThreads + threadIdx.x, stride = gridDim.x * kThreads;
  const int ihi = a.c1 < n ? a.c1 : n;
  if (decision == EKF_DEC_OLD) {
    const int opt_i = sm->opt_i;
    const double h00 = u.mCt[0], h01 = u.mCt[2], h02 = sm->h3[0];
    const double h10 = u.mCt[1], h11 = u.mCt[3], h12 = sm->h3[1];
    const double c00 = u.Ct[0], c10 = u.Ct[2], c01 = u.Ct[1], c11 = u.Ct[3];
    const double s00 = sm->Si[0], s10 = sm->Si[1], s01 = sm->Si[2], s11 = sm->Si[3];
    const double r0 = sm->res[0], r1 = sm->res[1], l = sm->l, sq0 = sm->sq0, sq1 = sm->sq1;
    for (int i = a.c0 + g0; i < ihi; i += stride) {
      const double* c = scol(a, i);     // row i of P at columns {0,1,2,Opt_i,Opt_i+1} == column i at those rows
      const double p0 = c[0], p1 = c[1], p2 = c[2];
      const double pa = c[opt_i], pb = c[opt_i + 1];
      const double A0 = (p0 * h00 + p1 * h01) + p2 * h02;
      const double A1 = (p0 * h10 + p1 * h11) + p2 * h12;
      const double B0 = pa * c00 + pb * c10;
      const double B1 = pa * c01 + pb * c11;
      const double M0 = A0 + B0, M1 = A1 + B1;
      const double K0 = M0 * s00 + M1 * s10;      // Update.cpp:186
      const double K1 = M0 * s01 + M1 * s11;
      const double xi = a.x[i] + (K0 * r0 + K1 * r1);   // :187
      const double2 w = make_double2(sq0 * fma(l, K1, K0), sq1 * K1);
      for (int t = 0; t < a.n_shards; ++t) {      // the all-gather: peer stores over NVLink
        a.x_all[t][i] = xi;
        a.W_all[t][i] = w;
      }
    }
    if (g0 == 0) a.W[n] = make_double2(0.0, 0.0);   // pad row of the double2 sweep (n is odd)
  } else {
    const double h00 = u.mCt[0], h01 = u.mCt[2], h02 = sm->h3n[0];
    const double h10 = u.mCt[1], h11 = u.mCt[3], h12 = sm->h3n[1];
    const double ct00 = u.Ct[0], ct10 = u.Ct[1], ct01 = u.Ct[2], ct11 = u.Ct[3];
    int owner = 0;
    while (owner + 1 < a.n_shards && n >= a.bounds[owner + 1]) ++owner;
    for (int i = a.c0 + g0; i < ihi; i += stride) {   // Update.cpp:169,175-176
      double* c = scol(a, i);
      const double q0 = -c[0], q1 = -c[1], q2 = -c[2];
      const double t0 = (q0 * h00 + q1 * h01) + q2 * h02;
      const double t1 = (q0 * h10 + q1 * h11) + q2 * h12;
      const double o0 = t0 * ct00 + t1 * ct10;
      const double o1 = t0 * ct01 + t1 * ct11;
      c[n] = o0;                                      // mirror rows n, n+1 of the own column
      c[n + 1] = o1;
      a.W_all[owner][i] = make_double2(o0, o1);       // rows i of the two new columns, to their owner
    }
    if (g0 == 0) {
      a.x[n] = sm->nl[0];
      a.x[n + 1] = sm->nl[1];
    }
  }
}

// ---- the HBM-bound kernel over the own columns ---------------------------------------------------
template <int RANK, bool COMPASS>
__global__ void __launch_bounds__(kThreads) shard_downdate(const ShardArgs a) {
  if (!a.la) ekf_pdl_trigger();   // look-ahead runs: no early trigger, or the whole future chain queues up behind the sweep (ekf_pdl.cuh)
  ekf_pdl_wait();
  ShardSmall* sm = a.sm;
  const int n = sm->n;
  if (!COMPASS && sm->decision != EKF_DEC_OLD) {
    if (sm->decision == EKF_DEC_NEW && !a.la) {
      if (a.c0 <= n && n < a.c1) {                    // owner of the two new columns: rows from W
        double* ca = scol(a, n);
        double* cb = scol(a, n + 1);
        for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
          const double2 w = a.W[i];
          ca[i] = w.x;
          cb[i] = w.y;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
          const double off = 0.5 * (sm->PLL[2] + sm->PLL[1]);
          ca[n] = sm->PLL[0];
          ca[n + 1] = off;
          cb[n] = off;
          cb[n + 1] = sm->PLL[3];
        }
      }
      if (blockIdx.x == 0 && threadIdx.x == 0) a.nlm[0] = sm->n_lm + 1;
    }
    return;
  }
  const double m0 = COMPASS ? sm->cm0 : sm->m0, m1 = COMPASS ? 0.0 : sm->m1;
  const double2* __restrict__ W = a.W;
  if (!a.la && blockIdx.x == 0 && threadIdx.x < 9) {  // the replica of P_RR, same two fma
    const int i = threadIdx.x % 3, j = threadIdx.x / 3;
    const double2 wi = W[i], wj = W[j];
    double v = sm->PRR[i + 3 * j];
    if (RANK == 2) v = fma(m1 * wi.y, wj.y, v);
    v = fma(m0 * wi.x, wj.x, v);
    sm->PRR[i + 3 * j] = v;
  }
  const int jlo = a.c0, jhi = a.c1 < n ? a.c1 : n;
  if (jhi <= jlo) return;
  const int ld = a.ld;
  const int n_even = (n + 1) & ~1;
  const int rows_per_panel = 2 * kThreads;
  const int n_panels = (n_even + rows_per_panel - 1) / rows_per_panel;
  const int n_cb = (jhi - jlo + kCB - 1) / kCB;
  const long n_tiles = (long)n_panels * n_cb;
  for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int panel = (int)(tile % n_panels), cb = (int)(tile / n_panels);
    const int i = panel * rows_per_panel + 2 * threadIdx.x;
    if (i >= n_even) continue;
    const double2 wa = W[i], wb = W[i + 1];
    const double ua0 = m0 * wa.x, ua1 = m1 * wa.y, ub0 = m0 * wb.x, ub1 = m1 * wb.y;
    const int j0 = jlo + cb * kCB;
    double2* base = reinterpret_cast<double2*>(scol(a, j0) + i);
    const size_t cstride = (size_t)ld / 2;   // ld is even
    if (j0 + kCB <= jhi) {
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
      for (int j = 0; j0 + j < jhi; ++j) {
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
__global__ void shard_compass_setup(const ShardArgs a, const double* zR) {
  ekf_pdl_entry();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ShardSmall* sm = a.sm;
  sm->cres = ekf_compass_residual(a.x[2], zR[0], a.k);
  const double S = sm->PRR[8] + zR[1];
  sm->cS = S;
  sm->csq = sqrt(fabs(S));
  sm->cm0 = S < 0 ? 1.0 : -1.0;
  sm->n = 3 + 2 * a.nlm[0];
}

__global__ void __launch_bounds__(kThreads) shard_compass_gain(const ShardArgs a) {
  ekf_pdl_entry();
  const ShardSmall* sm = a.sm;
  const int n = sm->n;
  const double res = sm->cres, invS = 1 / sm->cS, sq = sm->csq;
  const int g0 = blockIdx.x * kThreads + threadIdx.x, stride = gridDim.x * kThreads;
  const int ihi = a.c1 < n ? a.c1 : n;
  for (int i = a.c0 + g0; i < ihi; i += stride) {
    const double Ki = invS * scol(a, i)[2];            // P(i,2) == P(2,i)
    const double xi = a.x[i] + res * Ki;
    const double2 w = make_double2(sq * Ki, 0.0);
    for (int t = 0; t < a.n_shards; ++t) {
      a.x_all[t][i] = xi;
      a.W_all[t][i] = w;
    }
  }
  if (g0 == 0) a.W[n] = make_double2(0.0, 0.0);
}

// ---- look-ahead runs (ekf_la.cuh): kernels that work on the replicated O(n) cache ------------------
__device__ __forceinline__ LaCache shard_cache(const ShardArgs& a) { return LaCache{a.strip, a.diag, a.lds}; }

// Exchange steps of a look-ahead run are flags, not events: a producer kernel stores its data, fences, and
// its last CTA stores the operation count into the flag word it owns on every shard; a consumer polls the
// flag words in its OWN memory. Compared with cudaStreamWaitEvent on G-1 peer events this costs one poll
// instead of G-1 serially processed stream waits, and the host threads need no barrier. Polls are bounded
// (about two seconds): on expiry status bit 2 is set and the run continues, so a lost peer cannot hang a GPU.
__device__ __forceinline__ unsigned long long shard_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// dbg: four words behind the last-CTA counter; the first poll that gives up leaves {site, wanted count, whose
// flag, value seen} there for the error message.
__device__ __forceinline__ void shard_poll(const unsigned* flags, int lo, int hi, unsigned want, int* status, unsigned* dbg,
                                           unsigned site, unsigned long long limit_ns) {
  const unsigned long long t0 = shard_now_ns();
  for (int t = lo; t < hi; ++t) {
    const volatile unsigned* f = flags + t;
    while (*f < want) {
      if (shard_now_ns() - t0 > limit_ns) {
        if ((atomicOr(status, 2) & 2) == 0) { dbg[0] = site; dbg[1] = want; dbg[2] = (unsigned)t; dbg[3] = *f; }
        break;
      }
    }
  }
  __threadfence_system();
}
// Called by every thread of the kernel after its stores; the last CTA to arrive publishes `done` in the
// flag word of this shard on shards [lo, hi).
__device__ __forceinline__ void shard_signal_last(const ShardArgs& a, unsigned* const* flag_all, int lo, int hi, unsigned done) {
  __shared__ bool last_cta;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last_cta = atomicAdd(a.gain_count, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last_cta || threadIdx.x != 0) return;
  *a.gain_count = 0;
  __threadfence_system();
  for (int t = lo; t < hi; ++t) *const_cast<volatile unsigned*>(flag_all[t] + a.shard) = done;
}
// One warp that waits for flag words [lo, hi) (side stream, and the main stream between gain and sweep).
__global__ void shard_poll_kernel(const ShardArgs a, int which, int lo, int hi, unsigned want) {
  ekf_pdl_wait();          // dependents are released by the exit, i.e. after the poll (ekf_pdl.cuh)
  if (threadIdx.x == 0) shard_poll(which ? a.flags_g : a.flags_d, lo, hi, want, a.status, a.gain_count + 1, 10 + which, a.poll_ns);
}

// Start of a run: every shard stores the cache entries of its own columns into every shard's cache.
__global__ void __launch_bounds__(kThreads) shard_la_load(const ShardArgs a) {
  const int n = 3 + 2 * a.nlm[0];
  const int ihi = a.c1 < n ? a.c1 : n;
  for (int i = a.c0 + blockIdx.x * kThreads + threadIdx.x; i < ihi; i += gridDim.x * kThreads) {
    const double* c = scol(a, i);
    const double v0 = c[0], v1 = c[1], v2 = c[2];          // P(r, i) == P(i, r)
    double d0 = 0.0, d1 = 0.0;
    int slot = 0;
    if (i >= 3) {                                           // column e of the landmark's 2x2 block
      const int Li = 3 + 2 * ((i - 3) >> 1), e = (i - 3) & 1;
      d0 = c[Li];
      d1 = c[Li + 1];
      slot = 2 * (Li - 3) + 2 * e;
    }
    for (int t = 0; t < a.n_shards; ++t) {
      double* st = a.strip_all[t];
      st[i] = v0;
      st[a.lds + i] = v1;
      st[2 * (size_t)a.lds + i] = v2;
      if (i >= 3) { a.diag_all[t][slot] = d0; a.diag_all[t][slot + 1] = d1; }
    }
  }
}
// End of a run: the robot rows of the own columns, shard 0's robot columns and the P_RR replica of the
// per-call surface from the cache (the rest of the slab is current).
__global__ void __launch_bounds__(kThreads) shard_la_store(const ShardArgs a) {
  const int n = 3 + 2 * a.nlm[0];
  const int g0 = blockIdx.x * kThreads + threadIdx.x, stride = gridDim.x * kThreads;
  const int ihi = a.c1 < n ? a.c1 : n;
  for (int i = a.c0 + g0; i < ihi; i += stride) {
    double* c = scol(a, i);
    c[0] = a.strip[i];
    c[1] = a.strip[a.lds + i];
    c[2] = a.strip[2 * (size_t)a.lds + i];
  }
  if (a.c0 == 0)
    for (int j = 3 + g0; j < n; j += stride) {
      a.P[j] = a.strip[j];
      a.P[j + (size_t)a.ld] = a.strip[a.lds + j];
      a.P[j + (size_t)2 * a.ld] = a.strip[2 * (size_t)a.lds + j];
    }
  if (g0 < 9) a.sm->PRR[g0] = *la_prr(shard_cache(a), g0 % 3, g0 / 3);
}

__global__ void shard_prop_setup_la(const ShardArgs a, const double* in3) {
  ekf_pdl_entry();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double* x = a.x;
  const LaCache c = shard_cache(a);
  PropSetup p;
  ekf_build_prop(p, in3[0], in3[1], in3[2], x[2], a.k);
  a.sm->prop = p;
  const double xm0 = p.v * p.c, xm1 = p.v * p.s, xm2 = p.w;   // Propagate.cpp:33-37
  x[0] = x[0] + p.dt * xm0;
  x[1] = x[1] + p.dt * xm1;
  x[2] = x[2] + p.dt * xm2;
  double PRR[9];
  for (int q = 0; q < 9; ++q) PRR[q] = *la_prr(c, q % 3, q / 3);
  ekf_prop_prr(p, PRR);
  for (int q = 0; q < 9; ++q) *la_prr(c, q % 3, q / 3) = PRR[q];
}

__global__ void __launch_bounds__(kThreads) shard_prop_strip_la(const ShardArgs a) {
  ekf_pdl_entry();
  __shared__ PropSetup ps;
  if (threadIdx.x == 0) ps = a.sm->prop;
  __syncthreads();
  const int n = 3 + 2 * a.nlm[0];
  for (int j = 3 + blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    double* c = a.strip + j;
    double a0 = c[0], a1 = c[a.lds], a2 = c[2 * (size_t)a.lds];
    ekf_prop_col(ps, a0, a1, a2);
    c[0] = a0; c[a.lds] = a1; c[2 * (size_t)a.lds] = a2;
  }
}

// Gating over ALL landmarks on every shard (the cache is replicated): no exchange, per-CTA candidates.
__global__ void __launch_bounds__(kThreads) shard_gate_la(const ShardArgs a, const double* zr) {
  ekf_pdl_entry();
  __shared__ CtaScratch sc;
  const double* x = a.x;
  const LaCache c = shard_cache(a);
  const int n_lm = a.nlm[0];
  if (threadIdx.x == 0) {
    double PRR[9];
    for (int q = 0; q < 9; ++q) PRR[q] = *la_prr(c, q % 3, q / 3);
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
    la_gate_inputs(c, Li, p, pll);
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

// Decision on every shard, identically; the decision thread also finishes the pose rows (ekf_la.cuh).
__global__ void __launch_bounds__(kThreads) shard_decide_la(const ShardArgs a, const double* zr, int n_cand, unsigned done,
                                                           int* out_decision, int* out_index, double* out_mahal) {
  ekf_pdl_entry();
  __shared__ CtaScratch sc;
  double val = INFINITY;
  int idx = INT_MAX;
  for (int q = threadIdx.x; q < n_cand; q += blockDim.x) {
    const double v = a.cand_val[q];
    const int i = a.cand_idx[q];
    if (v < val || (v == val && i < idx)) { val = v; idx = i; }
  }
  cta_argmin(val, idx, &sc);
  if (threadIdx.x != 0) return;
  ShardSmall* sm = a.sm;
  double* x = a.x;
  const LaCache c = shard_cache(a);
  const int n_lm = a.nlm[0];
  const int n = 3 + 2 * n_lm;
  const int opt_i = (idx == INT_MAX) ? 0 : idx;
  const double mahal = (idx == INT_MAX) ? a.k.mahal_init : val;
  int decision = ekf_decide(opt_i, mahal, a.k);
  int index = opt_i;
  const UpdateSetup& u = sm->upd;
  if (decision == EKF_DEC_OLD) {
    double p[6], pll[4];
    la_gate_inputs(c, opt_i, p, pll);
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
    la_pose_rows_old(c, sm, x, opt_i);
  } else if (decision == EKF_DEC_NEW) {
    if (n_lm >= a.cap_lm) {
      decision = EKF_DEC_DROPPED;
      index = -1;
      a.status[0] |= 1;
    } else {
      const double cs = u.c, sn = u.s, z0 = zr[0], z1 = zr[1];
      const double Cz0 = cs * z0 + (-sn) * z1, Cz1 = sn * z0 + cs * z1;   // Update.cpp:155
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
      la_pose_rows_new(c, sm, x, a.nlm, n, n_lm);
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
  // this shard's gating has read x: every shard's gain kernel may now overwrite its entries of it
  __threadfence_system();
  for (int t = 0; t < a.n_shards; ++t) *const_cast<volatile unsigned*>(a.flags_d_all[t] + a.shard) = done;
}

// The one kernel left between two sweeps: gain rows / state / downdate vectors of the OWN rows, stored
// into every shard (exchange 2). Row i of P at the gain columns: the robot part from the cache, the
// Opt_i part from the shard's own column i (rows Opt_i, Opt_i+1 - swept, never stale).
__global__ void __launch_bounds__(kThreads) shard_gain_la(const ShardArgs a, unsigned done) {
  ekf_pdl_wait();
  if (threadIdx.x == 0) shard_poll(a.flags_d, 0, a.n_shards, done, a.status, a.gain_count + 1, 1, a.poll_ns);   // every shard's decision is made
  __syncthreads();
  const ShardSmall* sm = a.sm;
  const int decision = sm->decision;
  ekf_pdl_trigger();       // the spin is over: the one-warp poll kernel behind may be scheduled
  const int n = sm->n, lds = a.lds;
  const int g0 = blockIdx.x * kThreads + threadIdx.x, stride = gridDim.x * kThreads;
  const int ilo = a.c0 < 3 ? 3 : a.c0, ihi = a.c1 < n ? a.c1 : n;
  if (decision == EKF_DEC_OLD) {
    const int opt_i = sm->opt_i;
    const LaGain q = la_gain_coef(sm);
    if (g0 == 0) {                                  // pose rows (every shard has them) and the pad row, locally
      a.W[0] = sm->Wp[0]; a.W[1] = sm->Wp[1]; a.W[2] = sm->Wp[2];
      a.W[n] = make_double2(0.0, 0.0);
    }
    for (int i = ilo + g0; i < ihi; i += stride) {
      const double* col = scol(a, i);
      double dx;
      double2 w;
      la_gain_row(q, a.strip[i], a.strip[lds + i], a.strip[2 * (size_t)lds + i], col[opt_i], col[opt_i + 1], dx, w);
      const double xi = a.x[i] + dx;                // Update.cpp:187
      for (int t = 0; t < a.n_shards; ++t) {        // the all-gather: peer stores over NVLink
        a.x_all[t][i] = xi;
        a.W_all[t][i] = w;
      }
    }
  } else if (decision == EKF_DEC_NEW) {
    // Update.cpp:169,175-176. Every input is in the replicated cache, so nothing is exchanged: each shard
    // writes rows n, n+1 of its own columns, the owner of the new columns writes them whole.
    const LaNew q = la_new_coef(sm);
    for (int i = ilo + g0; i < ihi; i += stride) {
      double o0, o1;
      la_new_row(q, a.strip[i], a.strip[lds + i], a.strip[2 * (size_t)lds + i], o0, o1);
      double* col = scol(a, i);
      col[n] = o0;
      col[n + 1] = o1;
    }
    if (a.c0 <= n && n < a.c1) {
      double* ca = scol(a, n);
      double* cb = scol(a, n + 1);
      for (int i = 3 + g0; i < n; i += stride) {
        double o0, o1;
        la_new_row(q, a.strip[i], a.strip[lds + i], a.strip[2 * (size_t)lds + i], o0, o1);
        ca[i] = o0;
        cb[i] = o1;
      }
      if (g0 == 0) {
        const double off = 0.5 * (sm->PLL[2] + sm->PLL[1]);
        ca[n] = sm->PLL[0];
        ca[n + 1] = off;
        cb[n] = off;
        cb[n + 1] = sm->PLL[3];
      }
    }
  }
  shard_signal_last(a, a.flags_g_all, 0, a.n_shards, done);   // exchange 2
}

// After exchange 2 every shard holds the full W: bring the whole cache replica up to date (O(n)).
__global__ void __launch_bounds__(kThreads) shard_cache_update(const ShardArgs a) {
  ekf_pdl_entry();
  const ShardSmall* sm = a.sm;
  if (sm->decision != EKF_DEC_OLD) return;
  const LaCache c = shard_cache(a);
  const int n_lm = sm->n_lm;
  const double m0 = sm->m0, m1 = sm->m1;
  const double2 wp[3] = {sm->Wp[0], sm->Wp[1], sm->Wp[2]};
  for (int lm = blockIdx.x * blockDim.x + threadIdx.x; lm < n_lm; lm += gridDim.x * blockDim.x) {
    const double2 w[2] = {a.W[3 + 2 * lm], a.W[4 + 2 * lm]};
    la_cache_landmark(c, lm, w, wp, m0, m1);
  }
}

__global__ void shard_compass_setup_la(const ShardArgs a, const double* zR, unsigned done) {
  ekf_pdl_entry();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ShardSmall* sm = a.sm;
  const LaCache c = shard_cache(a);
  sm->cres = ekf_compass_residual(a.x[2], zR[0], a.k);
  const double S = *la_prr(c, 2, 2) + zR[1];
  sm->cS = S;
  sm->csq = sqrt(fabs(S));
  sm->cm0 = S < 0 ? 1.0 : -1.0;
  sm->n = 3 + 2 * a.nlm[0];
  la_compass_pose_rows(c, sm, a.x);
  __threadfence();
  *const_cast<volatile unsigned*>(a.flags_d + a.shard) = done;    // local: the compass path has no cross-shard step
}
// Compass gain: every input is in the replicated cache, every shard computes all rows - no exchange.
__global__ void __launch_bounds__(kThreads) shard_compass_gain_la(const ShardArgs a, unsigned done) {
  ekf_pdl_wait();          // no early trigger: the sweep behind must not occupy the device while this spins
  if (threadIdx.x == 0) shard_poll(a.flags_d, a.shard, a.shard + 1, done, a.status, a.gain_count + 1, 2, a.poll_ns);
  __syncthreads();
  const ShardSmall* sm = a.sm;
  const LaCache c = shard_cache(a);
  const int n = sm->n, n_lm = (n - 3) / 2;
  const int lm0 = blockIdx.x * kThreads + threadIdx.x;
  if (lm0 == 0) {
    a.W[0] = sm->Wp[0]; a.W[1] = sm->Wp[1]; a.W[2] = sm->Wp[2];
    a.W[n] = make_double2(0.0, 0.0);
  }
  for (int lm = lm0; lm < n_lm; lm += gridDim.x * kThreads) {
    double wv[2];
    la_compass_landmark(c, sm, a.x, lm, wv);
    a.W[3 + 2 * lm] = make_double2(wv[0], 0.0);
    a.W[4 + 2 * lm] = make_double2(wv[1], 0.0);
  }
  shard_signal_last(a, a.flags_g_all, a.shard, a.shard + 1, done);
}

// ---- host side ---------------------------------------------------------------------------------------
struct Shard {
  int device = 0, sms = 0, grid = 0;
  int c0 = 0, c1 = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev = nullptr;            // exchange steps of the per-call surface (single host thread)
  cudaEvent_t evx[2] = {nullptr, nullptr};   // exchange steps of ekf_sharded_run (one host thread per shard)
  double* P = nullptr;
  double* x = nullptr;
  int* nlm = nullptr;
  int* status = nullptr;
  ShardSmall* sm = nullptr;
  double2* W = nullptr;
  double* cand_val = nullptr;
  int* cand_idx = nullptr;
  ShardCand* xcand = nullptr;
  double* records = nullptr;
  size_t rec_cap = 0;
  double* stage = nullptr;   // per-call inputs [64 + 6*EKF_MAX_MEAS]
  ShardArgs args;
  // look-ahead runs: side stream, the O(n) cache replica, the second W buffer, and the argument blocks of
  // even / odd operations (control block and W buffer alternate)
  cudaStream_t side = nullptr;
  cudaEvent_t evc = nullptr, evl = nullptr;     // main -> side at the start of a run, side -> main at its end
  double* strip = nullptr;
  double* diag = nullptr;
  double2* W2 = nullptr;
  unsigned* flags = nullptr;   // [kMaxShards] decisions, [kMaxShards] gains, [1] last-CTA counter
  // look-ahead runs sweep the slab with the TMA-staged kernel of ekf_large_tma.cu (its five-warp CTAs leave
  // room on every SM sub-partition for the side stream's warps; the 8-warp double2 sweep does not)
  std::vector<unsigned char> tmap;
  int tma_grid = 0;
  ShardArgs args_la[2];
};

std::string g_shard_create_error;

}  // namespace

struct ekf_sharded_s {
  int G = 0;
  Shard sh[kMaxShards];
  EkfConst k{};
  ekf_config cfg{};
  int cap_lm = 0, cap_n = 0, ld = 0, lds = 0;
  int lookahead = 1;         // ekf_sharded_run overlaps the O(n) chain with the sweep (EKF_SHARD_LOOKAHEAD=0: off)
  int use_tma = 1;           // look-ahead runs: TMA-staged sweep (EKF_LARGE_TMA=0: the plain double2 sweep)
  size_t w_count = 0;
  // outputs of the fused run (device of shard 0)
  int* t_dec = nullptr;
  int* t_idx = nullptr;
  double* t_mah = nullptr;
  double* t_pose = nullptr;
  size_t t_cap = 0, pose_cap = 0;
  int* o_dec = nullptr;      // per-call outputs [EKF_MAX_MEAS]
  int* o_idx = nullptr;
  double* o_mah = nullptr;
  cudaEvent_t t0 = nullptr, t1 = nullptr, d0 = nullptr, d1 = nullptr;
  float last_ms = 0.f, last_downdate_ms = 0.f;
  long long launches = 0;
  std::string err;
};

namespace {

int sfail(ekf_sharded m, int code, const std::string& msg) {
  if (m) m->err = msg;
  else g_shard_create_error = msg;
  return code;
}

#define SH_CK(m, call)                                                                           \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return sfail(m, EKF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));        \
  } while (0)

// One exchange step: every stream waits for every other shard's work enqueued so far.
int exchange(ekf_sharded m) {
  if (m->G == 1) return EKF_OK;
  for (int s = 0; s < m->G; ++s) {
    SH_CK(m, cudaSetDevice(m->sh[s].device));
    SH_CK(m, cudaEventRecord(m->sh[s].ev, m->sh[s].stream));
  }
  for (int s = 0; s < m->G; ++s)
    for (int t = 0; t < m->G; ++t)
      if (t != s) SH_CK(m, cudaStreamWaitEvent(m->sh[s].stream, m->sh[t].ev, 0));
  return EKF_OK;
}

int rows_grid(const Shard& s, int count) {
  int g = (count + kThreads - 1) / kThreads;
  if (g > s.grid) g = s.grid;
  return g < 1 ? 1 : g;
}

int enqueue_propagate(ekf_sharded m, const double* const* in3) {
  for (int s = 0; s < m->G; ++s) {
    Shard& sh = m->sh[s];
    SH_CK(m, cudaSetDevice(sh.device));
    ekf_launch_pdl(shard_prop_setup, 1, 32, 0, sh.stream, sh.args, in3[s]);
    ekf_launch_pdl(shard_prop_strip, rows_grid(sh, s == 0 ? m->cap_n : sh.c1 - sh.c0), kThreads, 0, sh.stream, sh.args);
  }
  m->launches += 2 * m->G;
  return EKF_OK;
}

int enqueue_update(ekf_sharded m, const double* const* zr, int chunk_pos, int* dec, int* idx, double* mah, bool time_downdate) {
  for (int s = 0; s < m->G; ++s) {
    Shard& sh = m->sh[s];
    SH_CK(m, cudaSetDevice(sh.device));
    ekf_launch_pdl(shard_gate, rows_grid(sh, (sh.c1 - sh.c0) / 2), kThreads, 0, sh.stream, sh.args, zr[s], chunk_pos);
  }
  int rc = exchange(m);
  if (rc != EKF_OK) return rc;
  for (int s = 0; s < m->G; ++s) {
    Shard& sh = m->sh[s];
    SH_CK(m, cudaSetDevice(sh.device));
    ekf_launch_pdl(shard_decide, 1, 32, 0, sh.stream, sh.args, zr[s], s == 0 ? dec : nullptr, s == 0 ? idx : nullptr, s == 0 ? mah : nullptr);
    ekf_launch_pdl(shard_gain, rows_grid(sh, sh.c1 - sh.c0), kThreads, 0, sh.stream, sh.args);
  }
  rc = exchange(m);
  if (rc != EKF_OK) return rc;
  for (int s = 0; s < m->G; ++s) {
    Shard& sh = m->sh[s];
    SH_CK(m, cudaSetDevice(sh.device));
    if (s == 0 && time_downdate) SH_CK(m, cudaEventRecord(m->d0, sh.stream));
    ekf_launch_pdl(shard_downdate<2, false>, sh.grid, kThreads, 0, sh.stream, sh.args);
    if (s == 0 && time_downdate) SH_CK(m, cudaEventRecord(m->d1, sh.stream));
  }
  m->launches += 4 * m->G;
  return EKF_OK;
}

int enqueue_compass(ekf_sharded m, const double* const* zR) {
  for (int s = 0; s < m->G; ++s) {
    Shard& sh = m->sh[s];
    SH_CK(m, cudaSetDevice(sh.device));
    ekf_launch_pdl(shard_compass_setup, 1, 32, 0, sh.stream, sh.args, zR[s]);
  }
  int rc = exchange(m);
  if (rc != EKF_OK) return rc;
  for (int s = 0; s < m->G; ++s) {
    Shard& sh = m->sh[s];
    SH_CK(m, cudaSetDevice(sh.device));
    ekf_launch_pdl(shard_compass_gain, rows_grid(sh, sh.c1 - sh.c0), kThreads, 0, sh.stream, sh.args);
  }
  rc = exchange(m);
  if (rc != EKF_OK) return rc;
  for (int s = 0; s < m->G; ++s) {
    Shard& sh = m->sh[s];
    SH_CK(m, cudaSetDevice(sh.device));
    ekf_launch_pdl(shard_downdate<1, true>, sh.grid, kThreads, 0, sh.stream, sh.args);
  }
  m->launches += 3 * m->G;
  return EKF_OK;
}

int sync_all(ekf_sharded m) {
  for (int s = 0; s < m->G; ++s) {
    SH_CK(m, cudaSetDevice(m->sh[s].device));
    SH_CK(m, cudaStreamSynchronize(m->sh[s].stream));
  }
  return EKF_OK;
}

int check_capacity(ekf_sharded m) {
  int st = 0;
  SH_CK(m, cudaSetDevice(m->sh[0].device));
  SH_CK(m, cudaMemcpy(&st, m->sh[0].status, sizeof(int), cudaMemcpyDeviceToHost));
  if (st & 1) return sfail(m, EKF_ERR_CAPACITY, "landmark capacity exceeded, a New association was dropped");
  return EKF_OK;
}

// ---- ekf_sharded_run: one host thread per shard ------------------------------------------------------
// With G shards a step is ~6 G launches plus two exchange steps of G records and G (G-1) waits;
// issued from one thread that is more host time than the GPUs need for the step at G = 8. Each
// shard's stream is therefore fed by its own host thread. An exchange step needs every record
// call to precede the waits on it (cudaStreamWaitEvent captures the event's state at call time):
// record, host barrier, waits. Two alternating events per shard make one barrier per exchange
// enough (an event is re-recorded only after the NEXT barrier, by which time every wait on its
// previous record has been issued).
struct SpinBarrier {
  std::atomic<int> count{0};
  std::atomic<int> gen{0};
  int n = 1;
  void wait() {
    const int g = gen.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
      count.store(0, std::memory_order_relaxed);
      gen.fetch_add(1, std::memory_order_release);
    } else {
      int spins = 0;
      while (gen.load(std::memory_order_acquire) == g)
        if (++spins > 4096) std::this_thread::yield();
    }
  }
};

struct RunCtx {
  ekf_sharded m = nullptr;
  int T = 0, L = 0, M = 1, max_meas = 0;
  const double* hrec = nullptr;
  bool want_trace = false, want_pose = false, timed = false;
  SpinBarrier bar;
  cudaError_t err[kMaxShards];
};

#define TH_CK(call)                                      \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess && c.err[s] == cudaSuccess) c.err[s] = e__;   \
  } while (0)

// Errors are recorded, never returned early: every thread must reach every barrier.
void run_shard_thread(RunCtx& c, int s) {
  ekf_sharded m = c.m;
  Shard& sh = m->sh[s];
  const int G = m->G;
  int xk = 0;
  auto xchg = [&]() {
    if (G == 1) return;
    TH_CK(cudaEventRecord(sh.evx[xk & 1], sh.stream));
    c.bar.wait();
    for (int t = 0; t < G; ++t)
      if (t != s) TH_CK(cudaStreamWaitEvent(sh.stream, m->sh[t].evx[xk & 1], 0));
    ++xk;
  };
  TH_CK(cudaSetDevice(sh.device));
  xchg();                                   // uploads and resets of every shard precede the first peer store
  if (s == 0) TH_CK(cudaEventRecord(m->t0, sh.stream));
  const int g_strip = rows_grid(sh, s == 0 ? m->cap_n : sh.c1 - sh.c0);
  const int g_rows = rows_grid(sh, sh.c1 - sh.c0);
  const int g_gate = rows_grid(sh, (sh.c1 - sh.c0) / 2);
  bool timed = false;
  for (int t = 0; t < c.T; ++t) {
    const double* hrec = c.hrec + (size_t)t * c.L;
    const double* rec = sh.records + (size_t)t * c.L;
    ekf_launch_pdl(shard_prop_setup, 1, 32, 0, sh.stream, sh.args, rec);
    ekf_launch_pdl(shard_prop_strip, g_strip, kThreads, 0, sh.stream, sh.args);
    if (hrec[6] != 0.0) {
      ekf_launch_pdl(shard_compass_setup, 1, 32, 0, sh.stream, sh.args, rec + 3);
      xchg();
      ekf_launch_pdl(shard_compass_gain, g_rows, kThreads, 0, sh.stream, sh.args);
      xchg();
      ekf_launch_pdl(shard_downdate<1, true>, sh.grid, kThreads, 0, sh.stream, sh.args);
    }
    int nz = (int)hrec[5];
    nz = nz < 0 ? 0 : nz > c.max_meas ? c.max_meas : nz;
    for (int q = 0; q < nz; ++q) {
      const double* zr = rec + 8 + 6 * q;
      const size_t oi = (size_t)t * c.M + q;
      const bool out = s == 0 && c.want_trace;
      ekf_launch_pdl(shard_gate, g_gate, kThreads, 0, sh.stream, sh.args, zr, 0);
      xchg();
      ekf_launch_pdl(shard_decide, 1, 32, 0, sh.stream, sh.args, zr, out ? m->t_dec + oi : nullptr, out ? m->t_idx + oi : nullptr,
                                          out ? m->t_mah + oi : nullptr);
      ekf_launch_pdl(shard_gain, g_rows, kThreads, 0, sh.stream, sh.args);
      xchg();
      const bool time_this = s == 0 && !timed && t >= c.T / 2;   // one downdate sampled mid-run on shard 0
      if (time_this) TH_CK(cudaEventRecord(m->d0, sh.stream));
      ekf_launch_pdl(shard_downdate<2, false>, sh.grid, kThreads, 0, sh.stream, sh.args);
      if (time_this) TH_CK(cudaEventRecord(m->d1, sh.stream));
      timed = timed || time_this;
    }
    if (s == 0 && c.want_pose)
      TH_CK(cudaMemcpyAsync(m->t_pose + (size_t)t * 3, sh.x, 3 * sizeof(double), cudaMemcpyDeviceToDevice, sh.stream));
  }
  xchg();
  if (s == 0) {
    TH_CK(cudaEventRecord(m->t1, sh.stream));
    c.timed = timed;
  }
  TH_CK(cudaGetLastError());
  TH_CK(cudaStreamSynchronize(sh.stream));
}

// Look-ahead run (header comment, ekf_la.cuh). Per shard two streams: A = sh.stream carries what has to sit
// between two sweeps, B = sh.side the chain that works on the cache replica alone. Operation k (1-based count
// `done`), landmark update:
//   B: gating k, decision k (-> flags_d = done on every shard)
//   A: gain k (waits for flags_d of EVERY shard: a peer's gain overwrites x entries its gating reads;
//              stores W / x into every shard, -> flags_g = done on every shard)
//   A: poll (flags_g of every shard: W and x complete), sweep k
//   B: poll (the same), cache update k, then propagate / gating of operation k+1 ...
// A compass operation has no cross-shard step (own flags only). Operations alternate between two control
// blocks and two W buffers, so the chain of operation k+1 never touches what the sweep of operation k reads.
// The two streams and the G host threads are coupled by the flags alone: no event, no host barrier per step.
// (Tools that serialise kernel execution cannot run this: EKF_SHARD_LOOKAHEAD=0 selects the event chain.)
void run_shard_thread_la(RunCtx& c, int s) {
  ekf_sharded m = c.m;
  Shard& sh = m->sh[s];
  const int G = m->G;
  cudaStream_t A = sh.stream, B = sh.side;
  int xk = 0;
  auto xchg = [&]() {                       // full exchange on the main streams (start / end of the run)
    if (G == 1) return;
    TH_CK(cudaEventRecord(sh.evx[xk & 1], A));
    c.bar.wait();
    for (int t = 0; t < G; ++t)
      if (t != s) TH_CK(cudaStreamWaitEvent(A, m->sh[t].evx[xk & 1], 0));
    ++xk;
  };
  TH_CK(cudaSetDevice(sh.device));
  TH_CK(cudaMemsetAsync(sh.flags, 0, (2 * kMaxShards + 8) * sizeof(unsigned), A));
  xchg();                                   // uploads and resets of every shard precede the first peer store
  if (s == 0) TH_CK(cudaEventRecord(m->t0, A));
  const int n_all = m->cap_n, lm_all = m->cap_lm;
  const int g_own = rows_grid(sh, sh.c1 - sh.c0), g_n = rows_grid(sh, n_all), g_lm = rows_grid(sh, lm_all);
  auto side_grid = [&](int count) {
    int g = (count + kSideThreads - 1) / kSideThreads;
    if (g > sh.grid) g = sh.grid;
    return g < 1 ? 1 : g;
  };
  const int sg_n = side_grid(n_all), sg_lm = side_grid(lm_all);
  shard_la_load<<<g_own, kThreads, 0, A>>>(sh.args_la[0]);
  xchg();                                   // every shard's cache replica is complete
  TH_CK(cudaEventRecord(sh.evc, A));
  TH_CK(cudaStreamWaitEvent(B, sh.evc, 0));
  unsigned op = 0;
  bool timed = false;
  auto sweep = [&](const ShardArgs& a, bool compass) {
    if (m->use_tma) {
      EkfLargeTmaArgs q{&a.sm->decision, &a.sm->n, &a.sm->m0, &a.sm->m1, a.W, nullptr, &a.sm->n_lm, 0, sh.c0, sh.c1, 1, 1};
      if (compass) q = EkfLargeTmaArgs{nullptr, &a.sm->n, &a.sm->cm0, &a.sm->cm0, a.W, nullptr, nullptr, 0, sh.c0, sh.c1, 1, 1};
      TH_CK(ekf_large_tma_downdate(q, sh.tmap.data(), sh.tma_grid, compass, A));
    } else if (compass) {
      ekf_launch_pdl(shard_downdate<1, true>, sh.grid, kThreads, 0, A, a);
    } else {
      ekf_launch_pdl(shard_downdate<2, false>, sh.grid, kThreads, 0, A, a);
    }
  };
  // EKF_SHARD_TIMELINE=1: CUDA events around every kernel of two mid-run steps of shard 0, printed to stderr
  const bool dbg = s == 0 && getenv("EKF_SHARD_TIMELINE") && atoi(getenv("EKF_SHARD_TIMELINE")) != 0;
  std::vector<cudaEvent_t> dev;
  std::vector<std::string> dname;
  auto mark = [&](cudaStream_t st, const char* name, int t) {
    if (!dbg || t < c.T / 2 || t >= c.T / 2 + 2) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    dev.push_back(e);
    dname.push_back(std::string(st == A ? "A " : "B ") + name + " t=" + std::to_string(t));
  };
  for (int t = 0; t < c.T; ++t) {
    const double* hrec = c.hrec + (size_t)t * c.L;
    const double* rec = sh.records + (size_t)t * c.L;
    mark(B, "before prop_setup", t);
    ekf_launch_pdl(shard_prop_setup_la, 1, 32, 0, B, sh.args_la[op & 1], rec);
    mark(B, "after prop_setup", t);
    ekf_launch_pdl(shard_prop_strip_la, sg_n, kSideThreads, 0, B, sh.args_la[op & 1]);
    if (hrec[6] != 0.0) {
      const ShardArgs& a = sh.args_la[op & 1];
      const unsigned done = op + 1;
      ekf_launch_pdl(shard_compass_setup_la, 1, 32, 0, B, a, rec + 3, done);
      ekf_launch_pdl(shard_compass_gain_la, g_lm, kThreads, 0, A, a, done);
      sweep(a, true);
      ekf_launch_pdl(shard_poll_kernel, 1, 32, 0, B, a, 1, s, s + 1, done);   // own compass gain has updated cache and x
      ++op;
    }
    int nz = (int)hrec[5];
    nz = nz < 0 ? 0 : nz > c.max_meas ? c.max_meas : nz;
    for (int q = 0; q < nz; ++q) {
      const ShardArgs& a = sh.args_la[op & 1];
      const unsigned done = op + 1;
      const double* zr = rec + 8 + 6 * q;
      const size_t oi = (size_t)t * c.M + q;
      const bool out = s == 0 && c.want_trace;
      mark(B, "after prop_strip", t);
      ekf_launch_pdl(shard_gate_la, sg_lm, kSideThreads, 0, B, a, zr);
      mark(B, "after gate", t);
      ekf_launch_pdl(shard_decide_la, 1, kSideThreads, 0, B, a, zr, sg_lm, done, out ? m->t_dec + oi : nullptr,
                     out ? m->t_idx + oi : nullptr, out ? m->t_mah + oi : nullptr);
      mark(B, "after decide", t);
      mark(A, "before gain", t);
      ekf_launch_pdl(shard_gain_la, g_own, kThreads, 0, A, a, done);
      mark(A, "after gain", t);
      ekf_launch_pdl(shard_poll_kernel, 1, 32, 0, A, a, 1, 0, G, done);
      mark(A, "after poll", t);
      const bool time_this = s == 0 && !timed && t >= c.T / 2;   // one downdate sampled mid-run on shard 0
      if (time_this) TH_CK(cudaEventRecord(m->d0, A));
      sweep(a, false);
      if (time_this) TH_CK(cudaEventRecord(m->d1, A));
      timed = timed || time_this;
      mark(A, "after sweep", t);
      ekf_launch_pdl(shard_poll_kernel, 1, 32, 0, B, a, 1, 0, G, done);
      mark(B, "after poll", t);
      ekf_launch_pdl(shard_cache_update, sg_lm, kSideThreads, 0, B, a);
      mark(B, "after cache_update", t);
      ++op;
    }
    if (s == 0 && c.want_pose)               // the pose is written by side-stream kernels only
      TH_CK(cudaMemcpyAsync(m->t_pose + (size_t)t * 3, sh.x, 3 * sizeof(double), cudaMemcpyDeviceToDevice, B));
  }
  TH_CK(cudaEventRecord(sh.evl, B));
  TH_CK(cudaStreamWaitEvent(A, sh.evl, 0));
  shard_la_store<<<g_n, kThreads, 0, A>>>(sh.args);
  xchg();
  if (s == 0) {
    TH_CK(cudaEventRecord(m->t1, A));
    c.timed = timed;
  }
  TH_CK(cudaGetLastError());
  TH_CK(cudaStreamSynchronize(B));
  TH_CK(cudaStreamSynchronize(A));
  for (size_t i = 0; i < dev.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, dev[0], dev[i]);
    fprintf(stderr, "[shard timeline] %9.1f us  %s\n", ms * 1e3, dname[i].c_str());
    if (i) cudaEventDestroy(dev[i]);
  }
  if (!dev.empty()) cudaEventDestroy(dev[0]);
}
#undef TH_CK

// Copy the same small host block into every shard's staging buffer at offset off.
int stage_all(ekf_sharded m, const double* host, size_t count, size_t off) {
  for (int s = 0; s < m->G; ++s) {
    SH_CK(m, cudaSetDevice(m->sh[s].device));
    SH_CK(m, cudaMemcpyAsync(m->sh[s].stage + off, host, count * sizeof(double), cudaMemcpyHostToDevice, m->sh[s].stream));
  }
  return EKF_OK;
}

}  // namespace

extern "C" {

int ekf_sharded_create(ekf_sharded* out, int n_shards, const int* devices, int max_landmarks, const ekf_config* cfg_in) {
  if (!out) return sfail(nullptr, EKF_ERR_BAD_ARG, "out is NULL");
  *out = nullptr;
  if (n_shards < 1 || n_shards > kMaxShards || !devices || max_landmarks < 1)
    return sfail(nullptr, EKF_ERR_BAD_ARG, "ekf_sharded_create: 1 <= n_shards <= " + std::to_string(kMaxShards) + ", devices and max_landmarks >= 1 required");
  if (max_landmarks < n_shards) return sfail(nullptr, EKF_ERR_BAD_ARG, "ekf_sharded_create: fewer landmarks than shards");
  for (int s = 0; s < n_shards; ++s) {
    int sms = 0, maj = 0, mnr = 0;
    if (ekf_device_info(devices[s], &sms, &maj, &mnr, nullptr, nullptr) != EKF_OK)
      return sfail(nullptr, EKF_ERR_NO_DEVICE, "no usable CUDA device " + std::to_string(devices[s]) + " (this library has no CPU fallback)");
    if (maj != 10) return sfail(nullptr, EKF_ERR_NO_DEVICE, "device is sm_" + std::to_string(maj * 10 + mnr) + "; this library is built for sm_100a (B200) only");
  }
  for (int s = 0; s < n_shards; ++s)
    for (int t = 0; t < n_shards; ++t) {
      if (devices[s] == devices[t]) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, devices[s], devices[t]) != cudaSuccess || !can)
        return sfail(nullptr, EKF_ERR_UNSUPPORTED, "devices " + std::to_string(devices[s]) + " and " + std::to_string(devices[t]) + " have no peer access");
      cudaSetDevice(devices[s]);
      cudaError_t e = cudaDeviceEnablePeerAccess(devices[t], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return sfail(nullptr, EKF_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
      cudaGetLastError();
    }
  ekf_config cfg;
  if (cfg_in) cfg = *cfg_in;
  else ekf_default_config(&cfg);
  ekf_sharded m = new ekf_sharded_s();
  m->G = n_shards;
  m->cfg = cfg;
  m->k = EkfConst{cfg.sigma_v, cfg.sigma_w, cfg.deg2rad_pi, cfg.two_pi, cfg.cond_max, cfg.mahal_init, cfg.gamma_max, cfg.gamma_min};
  m->cap_lm = max_landmarks;
  m->cap_n = 3 + 2 * max_landmarks;
  m->ld = (m->cap_n + 15) & ~15;
  m->w_count = (size_t)m->cap_n + 512;     // zero tail: boundary tiles of the TMA sweep read past n
  m->lds = (m->cap_n + 2 + 7) & ~7;
  {
    const char* env = getenv("EKF_SHARD_LOOKAHEAD");
    m->lookahead = env ? (atoi(env) != 0) : 1;
    env = getenv("EKF_LARGE_TMA");
    m->use_tma = env ? (atoi(env) != 0) : 1;
    // The look-ahead run couples the shards through flags that kernels poll, which needs the kernels of
    // different shards to make progress independently of each other. Shards on their own GPUs do. Shards
    // that share a device (a convenience for testing the exchange logic on one GPU) do not reliably: late in
    // long runs a shard's side stream stopped being scheduled while another shard's kernel polled
    // (profiles/dbg_cap.py 6 0,0,0 - never with one shard per device), so they run the event chain.
    for (int s = 0; s < n_shards; ++s)
      for (int t = 0; t < s; ++t)
        if (devices[s] == devices[t]) m->lookahead = 0;
  }
  auto bail = [&](int code, const std::string& msg) {
    g_shard_create_error = msg;
    ekf_sharded_destroy(m);
    return code;
  };
  int bounds[kMaxShards + 1];
  bounds[0] = 0;
  for (int s = 1; s < n_shards; ++s) bounds[s] = 3 + 2 * (int)(((long long)max_landmarks * s) / n_shards);
  bounds[n_shards] = m->cap_n;
  for (int s = 0; s < n_shards; ++s) {
    Shard& sh = m->sh[s];
    sh.device = devices[s];
    sh.c0 = bounds[s];
    sh.c1 = bounds[s + 1];
    ekf_device_info(sh.device, &sh.sms, nullptr, nullptr, nullptr, nullptr);
    if (cudaSetDevice(sh.device) != cudaSuccess) return bail(EKF_ERR_CUDA, "cudaSetDevice failed");
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, shard_downdate<2, false>, kThreads, 0);
    if (e != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    sh.grid = per_sm * sh.sms;
    const size_t slab = (size_t)(sh.c1 - sh.c0) * m->ld;
#define SH_ALLOC(ptr, bytes)                                                                     \
  if ((e = cudaMalloc(&(ptr), (bytes))) != cudaSuccess)                                          \
    return bail(EKF_ERR_CUDA, std::string("cudaMalloc(" #ptr "): ") + cudaGetErrorString(e));
    if ((e = cudaStreamCreateWithFlags(&sh.stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
    if ((e = cudaEventCreateWithFlags(&sh.ev, cudaEventDisableTiming)) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
    if ((e = cudaEventCreateWithFlags(&sh.evx[0], cudaEventDisableTiming)) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
    if ((e = cudaEventCreateWithFlags(&sh.evx[1], cudaEventDisableTiming)) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
    SH_ALLOC(sh.P, slab * sizeof(double));
    SH_ALLOC(sh.x, (size_t)(m->cap_n + 1) * sizeof(double));
    SH_ALLOC(sh.nlm, sizeof(int));
    SH_ALLOC(sh.status, sizeof(int));
    SH_ALLOC(sh.sm, 2 * sizeof(ShardSmall));
    SH_ALLOC(sh.W, m->w_count * sizeof(double2));
    SH_ALLOC(sh.W2, m->w_count * sizeof(double2));
    SH_ALLOC(sh.strip, 3 * (size_t)m->lds * sizeof(double));
    SH_ALLOC(sh.diag, 4 * ((size_t)max_landmarks + 1) * sizeof(double));
    SH_ALLOC(sh.flags, (2 * kMaxShards + 8) * sizeof(unsigned));
    {
      // Load every kernel of the look-ahead run now (CUDA loads a kernel lazily at its first launch, and that
      // load may have to wait for kernels that are running): no first launch happens while a kernel polls.
      cudaFuncAttributes fa;
      const void* kernels[] = {(const void*)shard_poll_kernel, (const void*)shard_la_load, (const void*)shard_la_store,
                               (const void*)shard_prop_setup_la, (const void*)shard_prop_strip_la, (const void*)shard_gate_la,
                               (const void*)shard_decide_la, (const void*)shard_gain_la, (const void*)shard_cache_update,
                               (const void*)shard_compass_setup_la, (const void*)shard_compass_gain_la,
                               (const void*)shard_downdate<2, false>, (const void*)shard_downdate<1, true>};
      for (const void* kf : kernels)
        if ((e = cudaFuncGetAttributes(&fa, kf)) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
    }
    if (m->use_tma) {
      sh.tmap.resize(ekf_large_tma_map_bytes());
      if (ekf_large_tma_prepare(sh.sms, &sh.tma_grid, true) != cudaSuccess ||
          ekf_large_tma_encode(sh.tmap.data(), sh.P, sh.c1 - sh.c0, m->ld) != cudaSuccess)
        return bail(EKF_ERR_CUDA, "the tensor map of the TMA-staged covariance sweep could not be encoded "
                                  "(set EKF_LARGE_TMA=0 to run the plain double2 sweep instead)");
    }
    {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      if ((e = cudaStreamCreateWithPriority(&sh.side, cudaStreamNonBlocking, hi)) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
      for (cudaEvent_t* ev : {&sh.evc, &sh.evl})
        if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return bail(EKF_ERR_CUDA, cudaGetErrorString(e));
    }
    SH_ALLOC(sh.cand_val, (size_t)sh.grid * sizeof(double));
    SH_ALLOC(sh.cand_idx, (size_t)sh.grid * sizeof(int));
    SH_ALLOC(sh.xcand, (size_t)kMaxShards * sizeof(ShardCand));
    SH_ALLOC(sh.stage, (size_t)(64 + 6 * EKF_MAX_MEAS) * sizeof(double));
    if (s == 0) {
      SH_ALLOC(m->o_dec, EKF_MAX_MEAS * sizeof(int));
      SH_ALLOC(m->o_idx, EKF_MAX_MEAS * sizeof(int));
      SH_ALLOC(m->o_mah, EKF_MAX_MEAS * sizeof(double));
      cudaEventCreate(&m->t0);
      cudaEventCreate(&m->t1);
      cudaEventCreate(&m->d0);
      cudaEventCreate(&m->d1);
    }
#undef SH_ALLOC
  }
  for (int s = 0; s < n_shards; ++s) {
    Shard& sh = m->sh[s];
    ShardArgs& a = sh.args;
    a.P = sh.P; a.x = sh.x; a.nlm = sh.nlm; a.status = sh.status; a.sm = sh.sm; a.W = sh.W;
    a.cand_val = sh.cand_val; a.cand_idx = sh.cand_idx; a.xcand = sh.xcand;
    a.ld = m->ld; a.c0 = sh.c0; a.c1 = sh.c1; a.shard = s; a.n_shards = n_shards; a.cap_lm = max_landmarks;
    for (int t = 0; t <= kMaxShards; ++t) a.bounds[t] = t <= n_shards ? bounds[t] : m->cap_n;
    for (int t = 0; t < kMaxShards; ++t) {
      const Shard& o = m->sh[t < n_shards ? t : 0];
      a.W_all[t] = o.W; a.x_all[t] = o.x; a.xcand_all[t] = o.xcand;
    }
    a.k = m->k;
    a.strip = sh.strip; a.diag = sh.diag; a.lds = m->lds; a.la = 0;
    a.flags_d = sh.flags; a.flags_g = sh.flags + kMaxShards; a.gain_count = sh.flags + 2 * kMaxShards;
    {
      const char* env = getenv("EKF_SHARD_POLL_MS");
      a.poll_ns = (env && atoll(env) > 0 ? (unsigned long long)atoll(env) : 2000ull) * 1000000ull;
    }
    for (int t = 0; t < kMaxShards; ++t) {
      const Shard& o = m->sh[t < n_shards ? t : 0];
      a.strip_all[t] = o.strip; a.diag_all[t] = o.diag;
      a.flags_d_all[t] = o.flags; a.flags_g_all[t] = o.flags + kMaxShards;
    }
    for (int b = 0; b < 2; ++b) {
      ShardArgs& l = sh.args_la[b];
      l = a;
      l.la = 1;
      l.sm = sh.sm + b;
      l.W = b ? sh.W2 : sh.W;
      for (int t = 0; t < kMaxShards; ++t) {
        const Shard& o = m->sh[t < n_shards ? t : 0];
        l.W_all[t] = b ? o.W2 : o.W;
      }
    }
  }
  int rc = ekf_sharded_reset(m);
  if (rc != EKF_OK) {
    const std::string msg = m->err;
    return bail(rc, msg);
  }
  *out = m;
  return EKF_OK;
}

int ekf_sharded_destroy(ekf_sharded m) {
  if (!m) return EKF_OK;
  for (int s = 0; s < m->G; ++s) {
    Shard& sh = m->sh[s];
    cudaSetDevice(sh.device);
    if (sh.stream) cudaStreamSynchronize(sh.stream);
  }
  for (int s = 0; s < m->G; ++s) {
    Shard& sh = m->sh[s];
    cudaSetDevice(sh.device);
    cudaFree(sh.P); cudaFree(sh.x); cudaFree(sh.nlm); cudaFree(sh.status); cudaFree(sh.sm); cudaFree(sh.W);
    cudaFree(sh.cand_val); cudaFree(sh.cand_idx); cudaFree(sh.xcand); cudaFree(sh.records); cudaFree(sh.stage);
    cudaFree(sh.W2); cudaFree(sh.strip); cudaFree(sh.diag); cudaFree(sh.flags);
    for (cudaEvent_t ev : {sh.evc, sh.evl})
      if (ev) cudaEventDestroy(ev);
    if (sh.side) { cudaStreamSynchronize(sh.side); cudaStreamDestroy(sh.side); }
    if (s == 0) {
      cudaFree(m->t_dec); cudaFree(m->t_idx); cudaFree(m->t_mah); cudaFree(m->t_pose);
      cudaFree(m->o_dec); cudaFree(m->o_idx); cudaFree(m->o_mah);
      if (m->t0) { cudaEventDestroy(m->t0); cudaEventDestroy(m->t1); cudaEventDestroy(m->d0); cudaEventDestroy(m->d1); }
    }
    if (sh.ev) cudaEventDestroy(sh.ev);
    if (sh.evx[0]) cudaEventDestroy(sh.evx[0]);
    if (sh.evx[1]) cudaEventDestroy(sh.evx[1]);
    if (sh.stream) cudaStreamDestroy(sh.stream);
  }
  delete m;
  return EKF_OK;
}

int ekf_sharded_reset(ekf_sharded m) {
  if (!m) return EKF_ERR_BAD_ARG;
  for (int s = 0; s < m->G; ++s) {
    Shard& sh = m->sh[s];
    SH_CK(m, cudaSetDevice(sh.device));
    SH_CK(m, cudaMemsetAsync(sh.P, 0, (size_t)(sh.c1 - sh.c0) * m->ld * sizeof(double), sh.stream));
    SH_CK(m, cudaMemsetAsync(sh.x, 0, (size_t)(m->cap_n + 1) * sizeof(double), sh.stream));
    SH_CK(m, cudaMemsetAsync(sh.nlm, 0, sizeof(int), sh.stream));
    SH_CK(m, cudaMemsetAsync(sh.status, 0, sizeof(int), sh.stream));
    SH_CK(m, cudaMemsetAsync(sh.sm, 0, 2 * sizeof(ShardSmall), sh.stream));
    SH_CK(m, cudaMemsetAsync(sh.W, 0, m->w_count * sizeof(double2), sh.stream));
    SH_CK(m, cudaMemsetAsync(sh.W2, 0, m->w_count * sizeof(double2), sh.stream));
    SH_CK(m, cudaMemsetAsync(sh.xcand, 0, (size_t)kMaxShards * sizeof(ShardCand), sh.stream));
  }
  return sync_all(m);
}

int ekf_sharded_n_shards(ekf_sharded m) { return m ? m->G : 0; }
int ekf_sharded_max_landmarks(ekf_sharded m) { return m ? m->cap_lm : 0; }
int ekf_sharded_columns(ekf_sharded m, int shard, int* c0, int* c1) {
  if (!m || shard < 0 || shard >= m->G) return EKF_ERR_BAD_ARG;
  if (c0) *c0 = m->sh[shard].c0;
  if (c1) *c1 = m->sh[shard].c1;
  return EKF_OK;
}

int ekf_sharded_set_state(ekf_sharded m, int n_landmarks, const double* x, const double* P, int ld) {
  if (!m || !x || !P) return EKF_ERR_BAD_ARG;
  const int n = 3 + 2 * n_landmarks;
  if (n_landmarks < 0 || n_landmarks > m->cap_lm || ld < n) return sfail(m, EKF_ERR_BAD_ARG, "ekf_sharded_set_state: bad n_landmarks / ld");
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < j; ++i) {
      const double a = P[i + (size_t)j * ld], b = P[j + (size_t)i * ld];
      if (!(a == b || (a != a && b != b)))
        return sfail(m, EKF_ERR_BAD_ARG, "ekf_sharded_set_state: P must be bit-symmetric (the reference symmetrises after every operation)");
    }
  double PRR[9];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) PRR[i + 3 * j] = P[i + (size_t)j * ld];
  int rc = ekf_sharded_reset(m);
  if (rc != EKF_OK) return rc;
  for (int s = 0; s < m->G; ++s) {
    Shard& sh = m->sh[s];
    SH_CK(m, cudaSetDevice(sh.device));
    SH_CK(m, cudaMemcpyAsync(sh.x, x, sizeof(double) * n, cudaMemcpyHostToDevice, sh.stream));
    const int jhi = sh.c1 < n ? sh.c1 : n;
    if (jhi > sh.c0)
      SH_CK(m, cudaMemcpy2DAsync(sh.P, (size_t)m->ld * sizeof(double), P + (size_t)sh.c0 * ld, (size_t)ld * sizeof(double),
                                 (size_t)n * sizeof(double), jhi - sh.c0, cudaMemcpyHostToDevice, sh.stream));
    SH_CK(m, cudaMemcpyAsync(sh.nlm, &n_landmarks, sizeof(int), cudaMemcpyHostToDevice, sh.stream));
    SH_CK(m, cudaMemcpyAsync(reinterpret_cast<char*>(sh.sm) + offsetof(ShardSmall, PRR), PRR, sizeof(PRR), cudaMemcpyHostToDevice, sh.stream));
  }
  return sync_all(m);
}

int ekf_sharded_get_state(ekf_sharded m, int* n_landmarks, double* x, double* P, int ld) {
  if (!m) return EKF_ERR_BAD_ARG;
  int rc = sync_all(m);
  if (rc != EKF_OK) return rc;
  int nl = 0;
  SH_CK(m, cudaSetDevice(m->sh[0].device));
  SH_CK(m, cudaMemcpy(&nl, m->sh[0].nlm, sizeof(int), cudaMemcpyDeviceToHost));
  const int n = 3 + 2 * nl;
  if (n_landmarks) *n_landmarks = nl;
  if (x) SH_CK(m, cudaMemcpy(x, m->sh[0].x, sizeof(double) * n, cudaMemcpyDeviceToHost));
  if (P) {
    if (ld < n) return sfail(m, EKF_ERR_BAD_ARG, "ekf_sharded_get_state: ld < n");
    for (int s = 0; s < m->G; ++s) {
      Shard& sh = m->sh[s];
      const int jhi = sh.c1 < n ? sh.c1 : n;
      if (jhi <= sh.c0) continue;
      SH_CK(m, cudaSetDevice(sh.device));
      SH_CK(m, cudaMemcpy2D(P + (size_t)sh.c0 * ld, (size_t)ld * sizeof(double), sh.P, (size_t)m->ld * sizeof(double),
                            (size_t)n * sizeof(double), jhi - sh.c0, cudaMemcpyDeviceToHost));
    }
  }
  return EKF_OK;
}

/* Debug / test hook: shard s's replicas (x[0:n], P_RR, landmark count) - they must be identical on every shard. */
int ekf_sharded_get_replica(ekf_sharded m, int shard, int* n_landmarks, double* x, double* PRR9) {
  if (!m || shard < 0 || shard >= m->G) return EKF_ERR_BAD_ARG;
  int rc = sync_all(m);
  if (rc != EKF_OK) return rc;
  Shard& sh = m->sh[shard];
  SH_CK(m, cudaSetDevice(sh.device));
  int nl = 0;
  SH_CK(m, cudaMemcpy(&nl, sh.nlm, sizeof(int), cudaMemcpyDeviceToHost));
  if (n_landmarks) *n_landmarks = nl;
  if (x) SH_CK(m, cudaMemcpy(x, sh.x, sizeof(double) * (3 + 2 * nl), cudaMemcpyDeviceToHost));
  if (PRR9) SH_CK(m, cudaMemcpy(PRR9, reinterpret_cast<char*>(sh.sm) + offsetof(ShardSmall, PRR), 9 * sizeof(double), cudaMemcpyDeviceToHost));
  return EKF_OK;
}

int ekf_sharded_get_pose(ekf_sharded m, double* xyphi, int32_t* n_landmarks) {
  if (!m) return EKF_ERR_BAD_ARG;
  int rc = sync_all(m);
  if (rc != EKF_OK) return rc;
  SH_CK(m, cudaSetDevice(m->sh[0].device));
  if (xyphi) SH_CK(m, cudaMemcpy(xyphi, m->sh[0].x, 3 * sizeof(double), cudaMemcpyDeviceToHost));
  if (n_landmarks) SH_CK(m, cudaMemcpy(n_landmarks, m->sh[0].nlm, sizeof(int), cudaMemcpyDeviceToHost));
  return EKF_OK;
}

int ekf_sharded_propagate(ekf_sharded m, double vel_mm_s, double rotvel_deg_s, double dt) {
  if (!m) return EKF_ERR_BAD_ARG;
  const double in3[3] = {vel_mm_s, rotvel_deg_s, dt};
  int rc = stage_all(m, in3, 3, 0);
  if (rc != EKF_OK) return rc;
  const double* ptr[kMaxShards];
  for (int s = 0; s < m->G; ++s) ptr[s] = m->sh[s].stage;
  rc = enqueue_propagate(m, ptr);
  if (rc != EKF_OK) return rc;
  return sync_all(m);
}

int ekf_sharded_update(ekf_sharded m, int n_z, const double* z, const double* R, int32_t* decision, int32_t* lm_index, double* mahal) {
  if (!m || n_z < 0 || n_z > EKF_MAX_MEAS || (n_z > 0 && (!z || !R)))
    return sfail(m, EKF_ERR_BAD_ARG, "ekf_sharded_update: bad arguments (n_z <= " + std::to_string(EKF_MAX_MEAS) + ")");
  if (n_z == 0) return EKF_OK;
  double zr[6 * EKF_MAX_MEAS];
  for (int q = 0; q < n_z; ++q) {
    zr[6 * q + 0] = z[2 * q + 0];
    zr[6 * q + 1] = z[2 * q + 1];
    for (int c = 0; c < 4; ++c) zr[6 * q + 2 + c] = R[4 * q + c];
  }
  int rc = stage_all(m, zr, (size_t)6 * n_z, 64);
  if (rc != EKF_OK) return rc;
  for (int q = 0; q < n_z; ++q) {
    const double* ptr[kMaxShards];
    for (int s = 0; s < m->G; ++s) ptr[s] = m->sh[s].stage + 64 + 6 * q;
    rc = enqueue_update(m, ptr, n_z == 1 ? 0 : (q == 0 ? 1 : 2), m->o_dec + q, m->o_idx + q, m->o_mah + q, false);
    if (rc != EKF_OK) return rc;
  }
  rc = sync_all(m);
  if (rc != EKF_OK) return rc;
  SH_CK(m, cudaSetDevice(m->sh[0].device));
  int dec[EKF_MAX_MEAS];
  SH_CK(m, cudaMemcpy(dec, m->o_dec, n_z * sizeof(int), cudaMemcpyDeviceToHost));
  if (decision) memcpy(decision, dec, n_z * sizeof(int));
  if (lm_index) SH_CK(m, cudaMemcpy(lm_index, m->o_idx, n_z * sizeof(int), cudaMemcpyDeviceToHost));
  if (mahal) SH_CK(m, cudaMemcpy(mahal, m->o_mah, n_z * sizeof(double), cudaMemcpyDeviceToHost));
  for (int q = 0; q < n_z; ++q)
    if (dec[q] == EKF_DECISION_DROPPED) return sfail(m, EKF_ERR_CAPACITY, "landmark capacity exceeded, a New association was dropped");
  return EKF_OK;
}

int ekf_sharded_update_compass(ekf_sharded m, double z, double R) {
  if (!m) return EKF_ERR_BAD_ARG;
  const double zR[2] = {z, R};
  int rc = stage_all(m, zR, 2, 8);
  if (rc != EKF_OK) return rc;
  const double* ptr[kMaxShards];
  for (int s = 0; s < m->G; ++s) ptr[s] = m->sh[s].stage + 8;
  rc = enqueue_compass(m, ptr);
  if (rc != EKF_OK) return rc;
  return sync_all(m);
}

int ekf_sharded_run(ekf_sharded m, int n_steps, int max_meas, const double* records, const ekf_run_outputs* out) {
  if (!m || !records || n_steps < 1 || max_meas < 0 || max_meas > EKF_MAX_MEAS)
    return sfail(m, EKF_ERR_BAD_ARG, "ekf_sharded_run: bad arguments (max_meas <= " + std::to_string(EKF_MAX_MEAS) + ")");
  const int L = EKF_RECORD_LEN(max_meas), M = max_meas > 0 ? max_meas : 1, T = n_steps;
  const size_t count = (size_t)T * L, TM = (size_t)T * M;
  const bool want_trace = out && (out->decision || out->lm_index || out->mahal);
  const bool want_pose = out && out->pose_trace;
  for (int s = 0; s < m->G; ++s) {       // the step records are a few hundred bytes per step: every shard gets them
    Shard& sh = m->sh[s];
    SH_CK(m, cudaSetDevice(sh.device));
    if (sh.rec_cap < count) {
      cudaFree(sh.records);
      sh.records = nullptr;
      sh.rec_cap = 0;
      SH_CK(m, cudaMalloc(&sh.records, count * sizeof(double)));
      sh.rec_cap = count;
    }
    SH_CK(m, cudaMemcpyAsync(sh.records, records, count * sizeof(double), cudaMemcpyHostToDevice, sh.stream));
  }
  Shard& s0 = m->sh[0];
  SH_CK(m, cudaSetDevice(s0.device));
  if (want_trace) {
    if (m->t_cap < TM) {
      cudaFree(m->t_dec); cudaFree(m->t_idx); cudaFree(m->t_mah);
      m->t_dec = m->t_idx = nullptr; m->t_mah = nullptr; m->t_cap = 0;
      SH_CK(m, cudaMalloc(&m->t_dec, TM * sizeof(int)));
      SH_CK(m, cudaMalloc(&m->t_idx, TM * sizeof(int)));
      SH_CK(m, cudaMalloc(&m->t_mah, TM * sizeof(double)));
      m->t_cap = TM;
    }
    SH_CK(m, cudaMemsetAsync(m->t_dec, 0xFF, TM * sizeof(int), s0.stream));
    SH_CK(m, cudaMemsetAsync(m->t_idx, 0xFF, TM * sizeof(int), s0.stream));
    SH_CK(m, cudaMemsetAsync(m->t_mah, 0, TM * sizeof(double), s0.stream));
  }
  if (want_pose && m->pose_cap < (size_t)T * 3) {
    cudaFree(m->t_pose);
    m->t_pose = nullptr;
    m->pose_cap = 0;
    SH_CK(m, cudaMalloc(&m->t_pose, (size_t)T * 3 * sizeof(double)));
    m->pose_cap = (size_t)T * 3;
  }
  RunCtx c;
  c.m = m;
  c.T = T; c.L = L; c.M = M; c.max_meas = max_meas;
  c.hrec = records;
  c.want_trace = want_trace;
  c.want_pose = want_pose;
  c.bar.n = m->G;
  for (int s = 0; s < kMaxShards; ++s) c.err[s] = cudaSuccess;
  void (*body)(RunCtx&, int) = m->lookahead ? run_shard_thread_la : run_shard_thread;
  if (m->G == 1) {
    body(c, 0);
  } else {
    std::vector<std::thread> th;
    for (int s = 1; s < m->G; ++s) th.emplace_back(body, std::ref(c), s);
    body(c, 0);
    for (auto& t : th) t.join();
  }
  long long per_step = m->lookahead ? 2 : 0;      // look-ahead: cache load / store
  for (int t = 0; t < T; ++t) {
    const double* hrec = records + (size_t)t * L;
    int nz = (int)hrec[5];
    nz = nz < 0 ? 0 : nz > max_meas ? max_meas : nz;
    per_step += 2 + (hrec[6] != 0.0 ? (m->lookahead ? 4 : 3) : 0) + (m->lookahead ? 7 : 4) * nz;
  }
  m->launches += per_step * m->G;
  for (int s = 0; s < m->G; ++s)
    if (c.err[s] != cudaSuccess)
      return sfail(m, EKF_ERR_CUDA, "ekf_sharded_run, shard " + std::to_string(s) + ": " + cudaGetErrorString(c.err[s]));
  for (int s = 0; s < m->G; ++s) {
    int st = 0;
    SH_CK(m, cudaSetDevice(m->sh[s].device));
    SH_CK(m, cudaMemcpy(&st, m->sh[s].status, sizeof(int), cudaMemcpyDeviceToHost));
    if (st & 2) {
      unsigned dbg[4] = {0, 0, 0, 0};
      cudaMemcpy(dbg, m->sh[s].flags + 2 * kMaxShards + 1, sizeof(dbg), cudaMemcpyDeviceToHost);
      return sfail(m, EKF_ERR_CUDA, "ekf_sharded_run: shard " + std::to_string(s) + " gave up waiting for a peer (exchange flag not set within 2 s: wait site " +
                   std::to_string(dbg[0]) + ", operation " + std::to_string(dbg[1]) + ", flag of shard " + std::to_string(dbg[2]) + " was " + std::to_string(dbg[3]) +
                   "); the map state is undefined");
    }
  }
  SH_CK(m, cudaSetDevice(s0.device));
  SH_CK(m, cudaEventElapsedTime(&m->last_ms, m->t0, m->t1));
  m->last_downdate_ms = 0.f;
  if (c.timed) SH_CK(m, cudaEventElapsedTime(&m->last_downdate_ms, m->d0, m->d1));
  if (out) {
    if (out->decision) SH_CK(m, cudaMemcpy(out->decision, m->t_dec, TM * sizeof(int), cudaMemcpyDeviceToHost));
    if (out->lm_index) SH_CK(m, cudaMemcpy(out->lm_index, m->t_idx, TM * sizeof(int), cudaMemcpyDeviceToHost));
    if (out->mahal) SH_CK(m, cudaMemcpy(out->mahal, m->t_mah, TM * sizeof(double), cudaMemcpyDeviceToHost));
    if (out->pose_trace) SH_CK(m, cudaMemcpy(out->pose_trace, m->t_pose, (size_t)T * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (out->final_pose) SH_CK(m, cudaMemcpy(out->final_pose, s0.x, 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (out->final_nlm) SH_CK(m, cudaMemcpy(out->final_nlm, s0.nlm, sizeof(int), cudaMemcpyDeviceToHost));
  }
  return check_capacity(m);
}

int ekf_sharded_last_run_ms(ekf_sharded m, float* run_ms, float* downdate_ms) {
  if (!m) return EKF_ERR_BAD_ARG;
  if (run_ms) *run_ms = m->last_ms;
  if (downdate_ms) *downdate_ms = m->last_downdate_ms;
  return EKF_OK;
}

long long ekf_sharded_kernel_launches(ekf_sharded m) { return m ? m->launches : 0; }

int ekf_sharded_run_mode(ekf_sharded m) { return !m ? -1 : !m->lookahead ? 0 : m->use_tma ? 2 : 1; }

const char* ekf_sharded_last_error(ekf_sharded m) { return m ? m->err.c_str() : g_shard_create_error.c_str(); }

}  // extern "C"
