// ekf_tile.cu — regime A, register-tile variant of the fused multi-step kernel (sm_100a).
//
// One CTA per filter, but the covariance does not live in shared memory: its lower block
// triangle is held in REGISTERS as 8x8 FP64 tiles, one tile per thread (NT(NT+1)/2 tile threads,
// 64 doubles = 128 registers each), for all T steps of the run. Shared memory only carries what
// the O(n) phases exchange: the state x, the first three columns of P ("strip"), the 2x2 diagonal
// landmark blocks, the two gain columns of the associated landmark, and the downdate vectors W.
//
// Internally the state is padded with one dummy entry after the robot pose
// ([X,Y,Phi,pad,L1x,L1y,...]) so every landmark pair is 2-aligned and never straddles a tile;
// the pad row/column of P is identically zero. External layout (C ABI, HBM) is unchanged.
//
// Per step (slam.cpp:130-182 order), with a dedicated helper warp for the scalar chains:
//   helper lane 0 : odometry -> Q, Phi, G, x update (kalmanfilter.cpp:17-37, Propagate.cpp:33-48)
//   tile threads  : column-0 tiles apply Phi to their strip rows in registers, tile (0,0) does the
//                   3x3 robot block (Propagate.cpp:53-60), strip / diagonal blocks re-published
//   helper warp   : gating, two landmarks per lane (Update.cpp:103-148), warp-shuffle argmin,
//                   decision, S^-1 and L D L^T of S (or the New-landmark blocks)
//   tile threads  : publish the two covariance columns of the associated landmark
//   all threads   : one state row each: gain, state correction, W (Update.cpp:186-187)
//   tile threads  : P_tile += u_rows (x) W_cols, 128 fma per thread, no shared-memory traffic for P
//                   (Update.cpp:188,193-194 in the bit-symmetric form described in ekf_cta.cuh)
// Arithmetic is shared with the other kernels (ekf_small.cuh), so results are bit-identical to
// the shared-memory-resident kernel in ekf_batch.cu.
#include <cstdlib>
#include "ekf_cta.cuh"
#include "ekf_internal.h"

namespace {

template <int NT>
struct TileCfg {
  static constexpr int NTILES = NT * (NT + 1) / 2;
  static constexpr int TW = (NTILES + 31) / 32 * 32;   // threads in tile warps
  static constexpr int THREADS = TW + 32;              // + helper warp
  static constexpr int NI = 8 * NT;                    // padded internal dimension
  static constexpr int MAX_LM = (NI - 4) / 2;
  static constexpr int MINB = NT == 13 ? 2 : 1;   // 128 threads x 255 registers: two CTAs per SM
};

template <int NT>
struct TileSmem {
  double xs[TileCfg<NT>::NI];
  double s0[TileCfg<NT>::NI], s1[TileCfg<NT>::NI], s2[TileCfg<NT>::NI];   // P(r,0..2)
  double d00[TileCfg<NT>::NI / 2], d10[TileCfg<NT>::NI / 2], d11[TileCfg<NT>::NI / 2];
  double ca[TileCfg<NT>::NI], cb[TileCfg<NT>::NI];                        // P(r,Li), P(r,Li+1)
  double2 W[TileCfg<NT>::NI];
  double rec[2][EKF_RECORD_LEN_MAX];
  PropSetup prop;
  UpdateSetup upd;
  double res[2], S[4], Si[4], h3[2];
  double l, sq0, sq1, m0, m1;
  double nl[2], PLL[4], h3n[2];
  double cres, cS;
  double mahal;
  int decision, opt_i;   // opt_i: INTERNAL state index of the associated landmark (0 = none)
};

struct RunArgs {
  EkfState st;
  EkfRunIO io;
  EkfConst k;
  int* sm_slots;   // [>= number of SMs], zeroed before the launch: arrival order of CTAs per SM
  int rot;         // role rotation (in warps) applied to every second CTA of an SM
};

__device__ __forceinline__ int ext_index(int r) { return r < 3 ? r : r - 1; }   // internal -> external
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// strip rows (tiles in block column 0) and 2x2 diagonal blocks (diagonal tiles) -> shared memory
template <int NT>
__device__ __forceinline__ void publish(TileSmem<NT>& sm, const double (&p)[8][8], bool is_tile, int I, int J) {
  if (is_tile && J == 0) {
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      sm.s0[8 * I + a] = p[a][0];
      sm.s1[8 * I + a] = p[a][1];
      sm.s2[8 * I + a] = p[a][2];
    }
  }
  if (is_tile && I == J) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sm.d00[4 * I + q] = p[2 * q][2 * q];
      sm.d10[4 * I + q] = p[2 * q + 1][2 * q];
      sm.d11[4 * I + q] = p[2 * q + 1][2 * q + 1];
    }
  }
}

template <int NT, int RANK>
__device__ __forceinline__ void tile_downdate(TileSmem<NT>& sm, double (&p)[8][8], int I, int J, double m0, double m1) {
  double u0[8], u1[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const double2 wi = sm.W[8 * I + a];
    u0[a] = m0 * wi.x;
    u1[a] = m1 * wi.y;
  }
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const double2 wj = sm.W[8 * J + b];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      double t = p[a][b];
      if (RANK == 2) t = fma(u1[a], wj.y, t);
      t = fma(u0[a], wj.x, t);
      p[a][b] = t;
    }
  }
}

// Gate one landmark from the shared-memory strip / diagonal blocks (internal index Li).
template <int NT>
__device__ __forceinline__ void gate_from_smem(const TileSmem<NT>& sm, int Li, GateResult& g) {
  double pp[6], pll[4];
  pp[0] = sm.s0[Li]; pp[1] = sm.s0[Li + 1];
  pp[2] = sm.s1[Li]; pp[3] = sm.s1[Li + 1];
  pp[4] = sm.s2[Li]; pp[5] = sm.s2[Li + 1];
  const int pr = Li >> 1;
  pll[0] = sm.d00[pr]; pll[1] = sm.d10[pr]; pll[2] = sm.d10[pr]; pll[3] = sm.d11[pr];
  ekf_gate_landmark(sm.upd, sm.xs[Li], sm.xs[Li + 1], pp, pll, g);
}

template <int NT>
__global__ void __launch_bounds__(TileCfg<NT>::THREADS, TileCfg<NT>::MINB) ekf_batch_tile_kernel(const RunArgs a) {
  using C = TileCfg<NT>;
  __shared__ __align__(16) TileSmem<NT> sm;
  __shared__ int s_slot;
  // Warp w of a CTA issues on sub-partition w % 4. The helper warp carries most of the FP64
  // issue slots of a step (scalar chains + gating), so co-resident CTAs rotate their warp roles:
  // their helper warps then sit on different sub-partitions.
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    s_slot = atomicAdd(&a.sm_slots[smid], 1);
  }
  __syncthreads();
  constexpr int NW = TileCfg<NT>::THREADS / 32;
  const int vwarp = ((threadIdx.x >> 5) + ((s_slot & 1) ? a.rot : 0)) % NW;
  const int tid = vwarp * 32 + (threadIdx.x & 31), lane = tid & 31;
  const bool is_tile = tid < C::NTILES;
  const bool is_helper = tid >= C::TW;
  int I = 0, J = 0;
  if (is_tile) {
    int t = tid;
    while (t > I) { t -= I + 1; ++I; }
    J = t;
  }
  const int ld = a.st.ld, L = a.io.L, T = a.io.T, M = a.io.M;
  const EkfConst& k = a.k;
  double p[8][8];

  for (int f = blockIdx.x; f < a.st.F; f += gridDim.x) {
    double* gP = a.st.P + (size_t)f * a.st.slab;
    double* gx = a.st.x + (size_t)f * a.st.xs;
    const double* grec = a.io.records + (size_t)f * T * L;
    int n_lm = a.st.nlm[f];
    int dropped = 0;
    {
      const int n_int = 4 + 2 * n_lm;
      if (is_tile) {
#pragma unroll
        for (int b = 0; b < 8; ++b)
#pragma unroll
          for (int aa = 0; aa < 8; ++aa) {
            const int r = 8 * I + aa, c = 8 * J + b;
            const bool live = r != 3 && c != 3 && r < n_int && c < n_int;
            p[aa][b] = live ? gP[ext_index(r) + (size_t)ext_index(c) * ld] : 0.0;
          }
      }
      for (int r = tid; r < C::NI; r += C::THREADS) sm.xs[r] = (r != 3 && r < n_int) ? gx[ext_index(r)] : 0.0;
      for (int i = tid; i < L; i += C::THREADS) cp_async8(&sm.rec[0][i], grec + i);
      publish<NT>(sm, p, is_tile, I, J);
      cp_async_wait_all();
    }
    __syncthreads();

    for (int t = 0; t < T; ++t) {
      const double* cur = sm.rec[t & 1];
      if (t + 1 < T) {
        const double* g = grec + (size_t)(t + 1) * L;
        for (int i = tid; i < L; i += C::THREADS) cp_async8(&sm.rec[(t + 1) & 1][i], g + i);
      }
      // ---- doPropagation (slam.cpp:136) ---------------------------------------------------------
      if (is_helper && lane == 0) {
        PropSetup ps;
        ekf_build_prop(ps, cur[0], cur[1], cur[2], sm.xs[2], k);
        sm.prop = ps;
        const double xm0 = ps.v * ps.c, xm1 = ps.v * ps.s, xm2 = ps.w;   // Propagate.cpp:33-37
        sm.xs[0] = sm.xs[0] + ps.dt * xm0;
        sm.xs[1] = sm.xs[1] + ps.dt * xm1;
        sm.xs[2] = sm.xs[2] + ps.dt * xm2;
      }
      __syncthreads();
      if (is_tile && J == 0) {
        if (I == 0) {
          double PRR[9];
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i) PRR[i + 3 * j] = p[i][j];
          ekf_prop_prr(sm.prop, PRR);                       // Propagate.cpp:53 (+ :66-67 on the 3x3)
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i) p[i][j] = PRR[i + 3 * j];
#pragma unroll
          for (int aa = 4; aa < 8; ++aa) {                  // Propagate.cpp:56-60 for rows 4..7
            ekf_prop_col(sm.prop, p[aa][0], p[aa][1], p[aa][2]);
            p[0][aa] = p[aa][0]; p[1][aa] = p[aa][1]; p[2][aa] = p[aa][2];
          }
        } else {
#pragma unroll
          for (int aa = 0; aa < 8; ++aa) ekf_prop_col(sm.prop, p[aa][0], p[aa][1], p[aa][2]);
        }
      }
      publish<NT>(sm, p, is_tile, I, J);
      __syncthreads();

      // ---- doUpdateCompass (slam.cpp:144-147, kalmanfilter.cpp:96-130) ---------------------------
      if (cur[6] != 0.0) {
        if (is_helper && lane == 0) {
          sm.cres = ekf_compass_residual(sm.xs[2], cur[3], k);
          sm.cS = sm.s2[2] + cur[4];
        }
        __syncthreads();
        {
          const double res = sm.cres, S = sm.cS, invS = 1 / S, sq = sqrt(fabs(S));
          for (int r = tid; r < C::NI; r += C::THREADS) {
            const double Ki = invS * sm.s2[r];
            sm.xs[r] = sm.xs[r] + res * Ki;
            sm.W[r] = make_double2(sq * Ki, 0.0);
          }
        }
        __syncthreads();
        if (is_tile && 8 * I < 4 + 2 * n_lm) tile_downdate<NT, 1>(sm, p, I, J, sm.cS < 0 ? 1.0 : -1.0, 0.0);
        publish<NT>(sm, p, is_tile, I, J);
        __syncthreads();
      }

      // ---- doUpdate per measurement (slam.cpp:150-171, Update.cpp:80-195) -----------------------
      const int nz = (int)cur[5];
      for (int m = 0; m < M; ++m) {
        int decision = EKF_DEC_NONE, index = -1;
        double mahal = 0.0;
        if (m < nz) {
          const double* zr = cur + 8 + 6 * m;
          if (is_helper) {
            if (lane == 0) {
              double PRR[9];
#pragma unroll
              for (int i = 0; i < 3; ++i) { PRR[i] = sm.s0[i]; PRR[i + 3] = sm.s1[i]; PRR[i + 6] = sm.s2[i]; }
              UpdateSetup u;
              ekf_build_setup(u, sm.xs[2], sm.xs[0], sm.xs[1], PRR, zr[0], zr[1], zr + 2);
              sm.upd = u;
            }
            __syncwarp();
            // gating loop, Update.cpp:103-148: landmarks lane, lane+32, ...
            double best = INFINITY;
            int best_idx = INT_MAX;
            double b_res0 = 0, b_res1 = 0, b_S0 = 0, b_S1 = 0, b_S2 = 0, b_S3 = 0, b_h0 = 0, b_h1 = 0;
            for (int lm = lane; lm < n_lm; lm += 32) {
              const int Li = 4 + 2 * lm;
              GateResult g;
              gate_from_smem<NT>(sm, Li, g);
              const bool valid = !(g.cond >= k.cond_max) && (k.mahal_init > g.d2);
              if (valid && g.d2 < best) {
                best = g.d2; best_idx = Li;
                b_res0 = g.res0; b_res1 = g.res1;
                b_S0 = g.S[0]; b_S1 = g.S[1]; b_S2 = g.S[2]; b_S3 = g.S[3];
                b_h0 = g.h3_0; b_h1 = g.h3_1;
              }
            }
            double val = best;
            int idx = best_idx;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {     // lowest index wins ties (Update.cpp:140)
              const double ov = __shfl_xor_sync(0xffffffffu, val, o);
              const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
              if (ov < val || (ov == val && oi < idx)) { val = ov; idx = oi; }
            }
            const int opt_i = (idx == INT_MAX) ? 0 : idx;
            const double mh = (idx == INT_MAX) ? k.mahal_init : val;
            int dec = ekf_decide(opt_i, mh, k);
            if (dec == EKF_DEC_NEW && n_lm >= a.st.cap_lm) dec = EKF_DEC_DROPPED;
            if (dec == EKF_DEC_OLD && best_idx == idx) {
              // Opt_res, Opt_S, Opt_H_R of the winner, plus S^-1 and L D L^T of S
              sm.res[0] = b_res0; sm.res[1] = b_res1;
              sm.S[0] = b_S0; sm.S[1] = b_S1; sm.S[2] = b_S2; sm.S[3] = b_S3;
              sm.h3[0] = b_h0; sm.h3[1] = b_h1;
              const double Sm[4] = {b_S0, b_S1, b_S2, b_S3};
              double Si[4];
              ekf_inv2(Sm, Si);
              sm.Si[0] = Si[0]; sm.Si[1] = Si[1]; sm.Si[2] = Si[2]; sm.Si[3] = Si[3];
              const double d0 = b_S0, l = b_S1 / b_S0, d1 = b_S3 - l * b_S1;
              sm.l = l;
              sm.sq0 = sqrt(fabs(d0));
              sm.sq1 = sqrt(fabs(d1));
              sm.m0 = d0 < 0 ? 1.0 : -1.0;
              sm.m1 = d1 < 0 ? 1.0 : -1.0;
            }
            if (lane == 0) {
              sm.decision = dec;
              sm.opt_i = opt_i;
              sm.mahal = mh;
              if (dec == EKF_DEC_NEW) {
                const UpdateSetup& u = sm.upd;
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
                  for (int i = 0; i < 2; ++i)
                    t1[i + 2 * j] = (a1[i] * HR[j] + a1[i + 2] * HR[j + 2]) + a1[i + 4] * HR[j + 4];
                for (int q = 0; q < 4; ++q) in[q] = t1[q] + u.R[q];
                const double Cm[4] = {u.Ct[0], u.Ct[2], u.Ct[1], u.Ct[3]};
                for (int j = 0; j < 2; ++j)
                  for (int i = 0; i < 2; ++i) b1[i + 2 * j] = Cm[i] * in[0 + 2 * j] + Cm[i + 2] * in[1 + 2 * j];
                for (int j = 0; j < 2; ++j)       // Update.cpp:168
                  for (int i = 0; i < 2; ++i)
                    sm.PLL[i + 2 * j] = b1[i] * u.Ct[0 + 2 * j] + b1[i + 2] * u.Ct[1 + 2 * j];
                sm.nl[0] = nl0; sm.nl[1] = nl1;
                sm.h3n[0] = h30; sm.h3n[1] = h31;
              }
            }
          }
          __syncthreads();
          decision = sm.decision;
          mahal = sm.mahal;
          const int Li = sm.opt_i;
          index = Li ? Li - 1 : 0;   // external state index

          if (decision == EKF_DEC_OLD) {
            // ---- publish the two covariance columns of landmark Li -------------------------------
            const int JL = Li >> 3, c = Li & 7;
            if (is_tile && J == JL) {
#define EKF_PUB_COL(C0)                                                   \
  _Pragma("unroll") for (int aa = 0; aa < 8; ++aa) {                      \
    sm.ca[8 * I + aa] = p[aa][C0];                                        \
    sm.cb[8 * I + aa] = p[aa][C0 + 1];                                    \
  }
              if (c == 0) { EKF_PUB_COL(0) } else if (c == 2) { EKF_PUB_COL(2) } else if (c == 4) { EKF_PUB_COL(4) } else { EKF_PUB_COL(6) }
#undef EKF_PUB_COL
            }
            if (is_tile && I == JL && J < JL) {
#define EKF_PUB_ROW(R0)                                                   \
  _Pragma("unroll") for (int b = 0; b < 8; ++b) {                         \
    sm.ca[8 * J + b] = p[R0][b];                                          \
    sm.cb[8 * J + b] = p[R0 + 1][b];                                      \
  }
              if (c == 0) { EKF_PUB_ROW(0) } else if (c == 2) { EKF_PUB_ROW(2) } else if (c == 4) { EKF_PUB_ROW(4) } else { EKF_PUB_ROW(6) }
#undef EKF_PUB_ROW
            }
            __syncthreads();
            // ---- gain, state correction, downdate vectors (Update.cpp:186-187) --------------------
            {
              const UpdateSetup& u = sm.upd;
              const double h00 = u.mCt[0], h01 = u.mCt[2], h02 = sm.h3[0];
              const double h10 = u.mCt[1], h11 = u.mCt[3], h12 = sm.h3[1];
              const double c00 = u.Ct[0], c10 = u.Ct[2], c01 = u.Ct[1], c11 = u.Ct[3];
              const double si0 = sm.Si[0], si1 = sm.Si[1], si2 = sm.Si[2], si3 = sm.Si[3];
              const double r0 = sm.res[0], r1 = sm.res[1], l = sm.l, sq0 = sm.sq0, sq1 = sm.sq1;
              for (int r = tid; r < C::NI; r += C::THREADS) {
                const double p0 = sm.s0[r], p1 = sm.s1[r], p2 = sm.s2[r];
                const double pa = sm.ca[r], pb = sm.cb[r];
                const double A0 = (p0 * h00 + p1 * h01) + p2 * h02;
                const double A1 = (p0 * h10 + p1 * h11) + p2 * h12;
                const double B0 = pa * c00 + pb * c10;
                const double B1 = pa * c01 + pb * c11;
                const double M0 = A0 + B0, M1 = A1 + B1;
                const double K0 = M0 * si0 + M1 * si1;
                const double K1 = M0 * si2 + M1 * si3;
                sm.xs[r] = sm.xs[r] + (K0 * r0 + K1 * r1);
                sm.W[r] = make_double2(sq0 * fma(l, K1, K0), sq1 * K1);
              }
            }
            __syncthreads();
            // ---- covariance downdate in registers (Update.cpp:188,193-194) -------------------------
            if (is_tile && 8 * I < 4 + 2 * n_lm) tile_downdate<NT, 2>(sm, p, I, J, sm.m0, sm.m1);
            publish<NT>(sm, p, is_tile, I, J);
            __syncthreads();
          } else if (decision == EKF_DEC_NEW) {
            // ---- state augmentation (Update.cpp:152-178) -----------------------------------------
            const int r0i = 4 + 2 * n_lm;            // internal index of the new landmark
            const int In = r0i >> 3, a0 = r0i & 7;
            if (is_tile && I == In) {
              const UpdateSetup& u = sm.upd;
              const double h00 = u.mCt[0], h01 = u.mCt[2], h02 = sm.h3n[0];
              const double h10 = u.mCt[1], h11 = u.mCt[3], h12 = sm.h3n[1];
              const double ct00 = u.Ct[0], ct10 = u.Ct[1], ct01 = u.Ct[2], ct11 = u.Ct[3];
              double o0[8], o1[8];
#pragma unroll
              for (int b = 0; b < 8; ++b) {          // P_RLi = -P[:,0:3]*H_R^T*H_Li (:169), column 8J+b
                const int j = 8 * J + b;
                const double q0 = -sm.s0[j], q1 = -sm.s1[j], q2 = -sm.s2[j];
                const double t0 = (q0 * h00 + q1 * h01) + q2 * h02;
                const double t1 = (q0 * h10 + q1 * h11) + q2 * h12;
                o0[b] = t0 * ct00 + t1 * ct10;
                o1[b] = t0 * ct01 + t1 * ct11;
              }
              const double off = 0.5 * (sm.PLL[2] + sm.PLL[1]);   // :193-194 on the new 2x2 block
#define EKF_NEW_ROWS(A0)                                                                       \
  _Pragma("unroll") for (int b = 0; b < 8; ++b) {                                              \
    if (8 * J + b < r0i) { p[A0][b] = o0[b]; p[A0 + 1][b] = o1[b]; }                           \
  }                                                                                            \
  if (J == In) {                                                                               \
    _Pragma("unroll") for (int b = 0; b < A0; ++b) { p[b][A0] = o0[b]; p[b][A0 + 1] = o1[b]; } \
    p[A0][A0] = sm.PLL[0]; p[A0 + 1][A0] = off; p[A0][A0 + 1] = off; p[A0 + 1][A0 + 1] = sm.PLL[3]; \
  }
              if (a0 == 0) { EKF_NEW_ROWS(0) } else if (a0 == 2) { EKF_NEW_ROWS(2) } else if (a0 == 4) { EKF_NEW_ROWS(4) } else { EKF_NEW_ROWS(6) }
#undef EKF_NEW_ROWS
            }
            if (tid == 0) {
              sm.xs[r0i] = sm.nl[0];
              sm.xs[r0i + 1] = sm.nl[1];
            }
            index = r0i - 1;
            n_lm += 1;
            publish<NT>(sm, p, is_tile, I, J);
            __syncthreads();
          } else if (decision == EKF_DEC_DROPPED) {
            dropped = 1;
            index = -1;
          }
        }
        if (tid == 0) {
          const size_t oi = ((size_t)f * T + t) * M + m;
          if (a.io.decision) a.io.decision[oi] = decision;
          if (a.io.index) a.io.index[oi] = index;
          if (a.io.mahal) a.io.mahal[oi] = mahal;
        }
      }
      if (a.io.pose_trace && tid < 3) a.io.pose_trace[((size_t)f * T + t) * 3 + tid] = sm.xs[tid];   // slam.cpp:181
      cp_async_wait_all();
      __syncthreads();
    }

    // ---- write back to HBM (external layout, both triangles) -------------------------------------
    {
      const int n_int = 4 + 2 * n_lm;
      if (is_tile) {
#pragma unroll
        for (int b = 0; b < 8; ++b)
#pragma unroll
          for (int aa = 0; aa < 8; ++aa) {
            const int r = 8 * I + aa, c = 8 * J + b;
            if (r != 3 && c != 3 && r < n_int && c < n_int) {
              gP[ext_index(r) + (size_t)ext_index(c) * ld] = p[aa][b];
              gP[ext_index(c) + (size_t)ext_index(r) * ld] = p[aa][b];
            }
          }
      }
      for (int r = tid; r < n_int; r += C::THREADS)
        if (r != 3) gx[ext_index(r)] = sm.xs[r];
      if (tid == 0) {
        a.st.nlm[f] = n_lm;
        if (dropped) a.st.status[f] |= 1;
      }
    }
    __syncthreads();
  }
}

template <int NT>
cudaError_t launch_tile(RunArgs a, int sm_count, cudaStream_t stream) {
  using C = TileCfg<NT>;
  static int grid_cap = 0;
  static int* slots = nullptr;
  static int rot = -1;
  if (!slots) {
    cudaError_t e = cudaMalloc(&slots, 1024 * sizeof(int));
    if (e != cudaSuccess) return e;
  }
  if (rot < 0) {
    const char* env = getenv("EKF_TILE_ROT");
    rot = env ? atoi(env) : 2;
  }
  cudaMemsetAsync(slots, 0, 1024 * sizeof(int), stream);
  a.sm_slots = slots;
  a.rot = rot;
  if (!grid_cap) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ekf_batch_tile_kernel<NT>, C::THREADS, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    grid_cap = per_sm * sm_count;
  }
  const int grid = a.st.F < grid_cap ? a.st.F : grid_cap;
  ekf_batch_tile_kernel<NT><<<grid, C::THREADS, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace

int ekf_tile_max_landmarks() { return TileCfg<16>::MAX_LM; }

cudaError_t ekf_tile_run(const EkfState& st, const EkfRunIO& io, const EkfConst& k, int sm_count, cudaStream_t stream) {
  RunArgs a{st, io, k, nullptr, 0};
  if (st.cap_lm <= TileCfg<13>::MAX_LM) return launch_tile<13>(a, sm_count, stream);
  if (st.cap_lm <= TileCfg<14>::MAX_LM) return launch_tile<14>(a, sm_count, stream);
  if (st.cap_lm <= TileCfg<15>::MAX_LM) return launch_tile<15>(a, sm_count, stream);
  if (st.cap_lm <= TileCfg<16>::MAX_LM) return launch_tile<16>(a, sm_count, stream);
  return cudaErrorInvalidValue;
}
