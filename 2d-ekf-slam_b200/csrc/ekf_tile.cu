// ekf_tile.cu — regime A, register-tile variant of the fused multi-step kernel (sm_100a).
//
// One CTA per filter, persistent over filters. The covariance does not live in shared memory:
// its lower block triangle is held in REGISTERS as 4x8 FP64 tiles, one tile per thread
// (NB(NB+1) tile threads: 182 for up to 50 landmarks, 64 registers of P each), for all T steps
// of the run. Shared memory (~12 KB) only carries what the O(n) phases exchange: the state x, the
// first three columns of P ("strip"), the 2x2 diagonal landmark blocks, the two gain columns of
// the associated landmark, the downdate vectors W and the gating partial sums.
//
// Internally the state is padded with one dummy entry after the robot pose
// ([X,Y,Phi,pad,L1x,L1y,...]) so every landmark pair is 2-aligned and never straddles a tile;
// the pad row/column of P is identically zero. External layout (C ABI, HBM) is unchanged.
//
// Every warp holds tiles AND takes part in the O(n) phases, so the latency-bound scalar chains
// of a step are spread over the four sub-partitions of the SM instead of serialising on one warp:
//   step start    warp NW-1 lane 0: odometry -> Q, Phi, G, new pose (kalmanfilter.cpp:17-37,
//                 Propagate.cpp:33-48);  warp NW-2 lane 0: sincos of the post-propagation heading
//                 and the rotation blocks of Update.cpp:89-95,113-114 (both need only x and the record)
//   propagate     column-0 tiles apply Phi to their strip rows in registers, tile (0,0) does the
//                 3x3 robot block (Propagate.cpp:53-60); strip / diagonal blocks re-published
//   gating        Update.cpp:103-148, one landmark per lane on TWO warp groups working in parallel:
//                 group A evaluates H_R P_RR H_R^T + H_Li P_LiR H_R^T, group B the other two terms of
//                 S; A then adds them in the reference's order, symmetrises, applies the cond gate
//                 and the Mahalanobis test; warp-shuffle argmin per warp (lowest index wins ties),
//                 one candidate slot per warp, S^-1 and L D L^T of S precomputed per candidate
//   update        tile threads publish the two covariance columns of the associated landmark; one
//                 state row per thread: gain, state correction, W (Update.cpp:186-187); then
//                 P_tile += u_rows (x) W_cols, 64 fma per thread, no shared-memory traffic for P
//                 (Update.cpp:188,193-194 in the bit-symmetric form described in ekf_cta.cuh)
// Arithmetic is shared with the other kernels (ekf_small.cuh), so results are bit-identical to
// the shared-memory-resident kernel in ekf_batch.cu.
#include <type_traits>

#include "ekf_cta.cuh"
#include "ekf_internal.h"

namespace {

template <int NB>
struct TileCfg {
  static constexpr int NI = 8 * NB;                    // padded internal dimension
  static constexpr int NTILES = NB * (NB + 1);         // tiles (I,J) with J <= I/2
  static constexpr int THREADS = (NTILES + 31) / 32 * 32;
  static constexpr int NW = THREADS / 32;
  static constexpr int MAX_LM = (NI - 4) / 2;
  static constexpr int GA = (MAX_LM + 31) / 32;        // warps per gating group
  static constexpr int LMP = GA * 32;                  // padded landmark slots
  static constexpr int MINB = NB == 13 ? 2 : 1;        // 192 threads x 168 registers: two CTAs per SM
  static_assert(2 * GA <= NW - 2, "gating groups and scalar warps must be distinct warps");
};

struct Candidate {          // best landmark of one gating warp (Opt_* of Update.cpp:140-147)
  double val;               // Mahalanobis distance (INFINITY: none)
  int idx;                  // internal state index (INT_MAX: none)
  int pad;
  double res[2], S[4], h3[2];
};

struct Post {               // derived from the winning S while the gain columns are being published
  double Si[4];             // S^-1 (Update.cpp:186)
  double l, sq0, sq1, m0, m1;   // S = L D L^T: l, sqrt|d|, -sign(d)
};

template <int NB>
struct TileSmem {
  using C = TileCfg<NB>;
  double xs[C::NI];
  double s0[C::NI], s1[C::NI], s2[C::NI];              // P(r,0..2)
  double d00[C::NI / 2], d10[C::NI / 2], d11[C::NI / 2];
  double ca[C::NI], cb[C::NI];                          // P(r,Li), P(r,Li+1)
  double2 W[C::NI];
  double t34[8][C::LMP];                                // group B partial sums t3[4], t4[4] per landmark
  double rec[2][EKF_RECORD_LEN_MAX];
  PropSetup prop;
  double PhiS[9], GS[6];                                // Phi_R, G column-major (Propagate.cpp:42-48)
  double xnew[3];
  double prr_new[9];                                    // propagated 3x3 robot block, for tile (0,0)
  UpdateSetup upd;                                      // landmark-independent part of the update
  Candidate cand[C::GA];
  Post post;
  double nl[2], PLL[4], h3n[2];
  double cres, cS;
};

struct RunArgs {
  EkfState st;
  EkfRunIO io;
  EkfConst k;
  long long* phase_cycles;   // optional [8]: cycles spent per phase, accumulated by CTA 0 (profiling aid)
};

__device__ __forceinline__ int ext_index(int r) { return r < 3 ? r : r - 1; }   // internal -> external
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void named_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Run f(std::integral_constant<int, v>) for the run-time value v in {0, STEP, 2*STEP, ...} < LIMIT:
// register tiles can only be indexed with compile-time constants.
template <int STEP, int LIMIT, int V = 0, typename F>
__device__ __forceinline__ void static_switch(int v, F&& f) {
  if constexpr (V < LIMIT) {
    if (v == V) f(std::integral_constant<int, V>{});
    else static_switch<STEP, LIMIT, V + STEP>(v, f);
  }
}

// strip rows (tiles in block column 0) and 2x2 diagonal blocks (diagonal tiles) -> shared memory
template <int NB>
__device__ __forceinline__ void publish_strip(TileSmem<NB>& sm, const double (&p)[4][8], int I) {
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    sm.s0[4 * I + a] = p[a][0];
    sm.s1[4 * I + a] = p[a][1];
    sm.s2[4 * I + a] = p[a][2];
  }
}

template <int NB>
__device__ __forceinline__ void publish(TileSmem<NB>& sm, const double (&p)[4][8], bool is_tile, int I, int J) {
  if (is_tile && J == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      sm.s0[4 * I + a] = p[a][0];
      sm.s1[4 * I + a] = p[a][1];
      sm.s2[4 * I + a] = p[a][2];
    }
  }
  if (is_tile && J == (I >> 1)) {   // the tile holding the diagonal 4x4 block of row block I
    static_switch<4, 8>(4 * (I & 1), [&](auto CO) {
      constexpr int co = decltype(CO)::value;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        sm.d00[2 * I + q] = p[2 * q][co + 2 * q];
        sm.d10[2 * I + q] = p[2 * q + 1][co + 2 * q];
        sm.d11[2 * I + q] = p[2 * q + 1][co + 2 * q + 1];
      }
    });
  }
}

// W is stored transposed by 8-row block, W[(r & 7) * NB + (r >> 3)]: the tile threads of a warp
// differ in their block indices, so for a fixed in-block offset their loads hit consecutive
// 16-byte words (no bank conflicts) and equal indices broadcast.
template <int NB>
__device__ __forceinline__ int widx(int r) { return (r & 7) * NB + (r >> 3); }

template <int NB, int RANK>
__device__ __forceinline__ void tile_downdate(TileSmem<NB>& sm, double (&p)[4][8], int I, int J, double m0, double m1) {
  double2 wi[4], wj[8];
#pragma unroll
  for (int a = 0; a < 4; ++a) wi[a] = sm.W[(4 * (I & 1) + a) * NB + (I >> 1)];
#pragma unroll
  for (int b = 0; b < 8; ++b) wj[b] = sm.W[b * NB + J];
  double u0[4], u1[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    u0[a] = m0 * wi[a].x;
    u1[a] = m1 * wi[a].y;
  }
#pragma unroll
  for (int b = 0; b < 8; ++b) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double t = p[a][b];
      if (RANK == 2) t = fma(u1[a], wj[b].y, t);
      t = fma(u0[a], wj[b].x, t);
      p[a][b] = t;
    }
  }
}

template <int NB>
__global__ void __launch_bounds__(TileCfg<NB>::THREADS, TileCfg<NB>::MINB) ekf_batch_tile_kernel(const RunArgs a) {
  using C = TileCfg<NB>;
  __shared__ __align__(16) TileSmem<NB> sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_tile = tid < C::NTILES;
  // The last warp has THREADS - NTILES >= 10 lanes without a tile: they carry the scalar work.
  constexpr int SP0 = C::NTILES - (C::NW - 1) * 32;           // first spare lane of the last warp
  const bool helper_warp = warp == C::NW - 1;
  const bool sc_prop = helper_warp && lane == SP0;            // scalar chain 1: propagate set-up
  const bool sc_trig = helper_warp && lane == SP0 + 1;        // scalar chain 2: heading trig for the update
  int I = 0, J = 0;
  if (is_tile) {                                    // tid = sum_{i<I}(i/2+1) + J
    int t = tid;
    while (t > (I >> 1)) { t -= (I >> 1) + 1; ++I; }
    J = t;
  }
  const int ld = a.st.ld, L = a.io.L, T = a.io.T, M = a.io.M;
  const EkfConst& k = a.k;
  double p[4][8];
#ifdef EKF_TILE_TIMING   // profiling builds only (make EXTRA=-DEKF_TILE_TIMING): the probes cost registers
  const bool timing = a.phase_cycles != nullptr && blockIdx.x == 0 && tid == 0;
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = 0;
#else
  constexpr bool timing = false;
  (void)timing;
#endif
#ifdef EKF_FINE_TIMING
  long long facc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long fprev = 0;
#define EKF_FINE0 if (timing) fprev = clock64();
#define EKF_FINE(i)                                   \
  if (timing) {                                       \
    const long long now = clock64();                  \
    facc[i] += now - fprev;                           \
    fprev = now;                                      \
  }
#else
#define EKF_FINE0
#define EKF_FINE(i)
#endif
#ifdef EKF_TILE_TIMING
#define EKF_PHASE(i)                                  \
  if (timing) {                                       \
    const long long now = clock64();                  \
    tacc[i] += now - tprev;                           \
    tprev = now;                                      \
  }
#else
#define EKF_PHASE(i)
#endif

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
          for (int aa = 0; aa < 4; ++aa) {
            const int r = 4 * I + aa, c = 8 * J + b;
            const bool live = r != 3 && c != 3 && r < n_int && c < n_int;
            p[aa][b] = live ? gP[ext_index(r) + (size_t)ext_index(c) * ld] : 0.0;
          }
      }
      for (int r = tid; r < C::NI; r += C::THREADS) sm.xs[r] = (r != 3 && r < n_int) ? gx[ext_index(r)] : 0.0;
      for (int i = tid; i < L; i += C::THREADS) cp_async8(&sm.rec[0][i], grec + i);
      publish<NB>(sm, p, is_tile, I, J);
      cp_async_wait_all();
    }
    __syncthreads();

    // Scalar chains of one step (kalmanfilter.cpp:17-37, Propagate.cpp:33-48, Update.cpp:89-95):
    // thread sc_prop builds Q, Phi, G and the propagated pose, thread sc_trig the rotation blocks
    // for the post-propagation heading and the first measurement of the step.
    auto scalar_chains = [&](const double* rec) {
      if (sc_prop || sc_trig) {
        // the two lanes share one sincos instruction stream: lane SP0 takes the pre-propagation
        // heading, lane SP0+1 the heading the update will see (same expression as the pose
        // update below, so the same bits)
        const double RTV = rec[1] * k.deg2rad_pi / 180.0;
        const double phi = sc_prop ? sm.xs[2] : sm.xs[2] + rec[2] * RTV;
        double sn, cs;
        sincos(phi, &sn, &cs);
        if (sc_prop) {
          PropSetup ps;
          ekf_build_prop_sc(ps, rec[0], rec[1], rec[2], sn, cs, k);
          sm.prop = ps;
          sm.PhiS[0] = 1.0; sm.PhiS[1] = 0.0; sm.PhiS[2] = 0.0;           // Propagate.cpp:42-44
          sm.PhiS[3] = 0.0; sm.PhiS[4] = 1.0; sm.PhiS[5] = 0.0;
          sm.PhiS[6] = ps.phi02; sm.PhiS[7] = ps.phi12; sm.PhiS[8] = 1.0;
          sm.GS[0] = ps.g00; sm.GS[1] = ps.g10; sm.GS[2] = 0.0;            // :46-48
          sm.GS[3] = 0.0; sm.GS[4] = 0.0; sm.GS[5] = ps.g21;
          const double xm0 = ps.v * ps.c, xm1 = ps.v * ps.s, xm2 = ps.w;   // Propagate.cpp:33-37
          sm.xnew[0] = sm.xs[0] + ps.dt * xm0;
          sm.xnew[1] = sm.xs[1] + ps.dt * xm1;
          sm.xnew[2] = sm.xs[2] + ps.dt * xm2;
        } else {
          UpdateTrig tg;
          ekf_build_trig_sc(tg, sn, cs);
          UpdateSetup& u = sm.upd;
          u.c = tg.c; u.s = tg.s;
          for (int q = 0; q < 4; ++q) { u.Ct[q] = tg.Ct[q]; u.mCt[q] = tg.mCt[q]; u.mCtJ[q] = tg.mCtJ[q]; }
          if ((int)rec[5] > 0) {                         // first measurement of the step
            u.z0 = rec[8]; u.z1 = rec[9];
            for (int q = 0; q < 4; ++q) u.R[q] = rec[10 + q];
          }
        }
      }
    };
    bool scalar_done = false;

#ifdef EKF_TILE_TIMING
    if (timing) tprev = clock64();
#endif
    for (int t = 0; t < T; ++t) {
      const double* cur = sm.rec[t & 1];
      if (t + 1 < T) {
        const double* g = grec + (size_t)(t + 1) * L;
        for (int i = tid; i < L; i += C::THREADS) cp_async8(&sm.rec[(t + 1) & 1][i], g + i);
      }
      // ---- doPropagation (slam.cpp:136): the two scalar chains run on different warps. They only
      // need the final state of the previous step and the record, so when the previous step ended
      // with an Old update they have already run, overlapped with that step's covariance downdate.
      const int nz = min((int)cur[5], (L - 8) / 6);   // never read past the record's measurement slots
      if (!scalar_done) {
        scalar_chains(cur);
        __syncthreads();
      }
      scalar_done = false;
      EKF_PHASE(0)   // scalar chains (when not overlapped)
      if (sc_prop) {
        sm.xs[0] = sm.xnew[0]; sm.xs[1] = sm.xnew[1]; sm.xs[2] = sm.xnew[2];
        sm.upd.x0 = sm.xnew[0]; sm.upd.x1 = sm.xnew[1];
      }
      if (helper_warp) {
        // 3x3 robot block, one element per spare lane (Propagate.cpp:53, then :66-67), straight from
        // and back to the shared-memory strip; the q partial sums of the update follow by shuffle.
        const int e = (lane + 36 - SP0) % 9, i = e % 3, j = e / 3;   // lanes SP0..SP0+8 map to e = 0..8
        double PRR[9];
#pragma unroll
        for (int r = 0; r < 3; ++r) { PRR[r] = sm.s0[r]; PRR[r + 3] = sm.s1[r]; PRR[r + 6] = sm.s2[r]; }
        const double mij = ekf_prop_prr_elem(sm.PhiS, sm.GS, sm.prop.Q, PRR, i, j);
        const double mji = __shfl_sync(0xffffffffu, mij, SP0 + j + 3 * i);
        const double pn = 0.5 * (mij + mji);
        // q(qi,qj) = mCt(qi,0)*P(0,qj) + mCt(qi,1)*P(1,qj)   (ekf_complete_setup)
        const int qe = e % 6, qi = qe % 2, qj = qe / 2;
        const double p0j = __shfl_sync(0xffffffffu, pn, SP0 + 0 + 3 * qj);
        const double p1j = __shfl_sync(0xffffffffu, pn, SP0 + 1 + 3 * qj);
        const double qv = sm.upd.mCt[qi] * p0j + sm.upd.mCt[qi + 2] * p1j;
        if (lane >= SP0 && lane < SP0 + 9) {
          sm.prr_new[e] = pn;
          sm.upd.PRR[e] = pn;
          double* col = j == 0 ? sm.s0 : (j == 1 ? sm.s1 : sm.s2);
          col[i] = pn;
          if (e < 6) sm.upd.q[qe] = qv;
        }
      }
      // P_RL <- Phi*P_RL (Propagate.cpp:56-60), one strip row per thread, in shared memory; the
      // column-0 tiles take the result back into their registers after the barrier.
      for (int r = 4 + tid; r < C::NI; r += C::THREADS) {
        double a0 = sm.s0[r], a1 = sm.s1[r], a2 = sm.s2[r];
        ekf_prop_col(sm.prop, a0, a1, a2);
        sm.s0[r] = a0; sm.s1[r] = a1; sm.s2[r] = a2;
      }
      __syncthreads();
      EKF_PHASE(1)   // covariance propagate + publish
      if (tid == 0) {   // tile (0,0) takes the propagated robot block into its registers
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int i = 0; i < 3; ++i) p[i][j] = sm.prr_new[i + 3 * j];
      } else if (is_tile && J == 0) {   // the other column-0 tiles take their propagated strip rows
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) {
          p[aa][0] = sm.s0[4 * I + aa];
          p[aa][1] = sm.s1[4 * I + aa];
          p[aa][2] = sm.s2[4 * I + aa];
        }
      }
      bool setup_valid = true;   // sm.upd describes the current state and measurement 0

      // ---- doUpdateCompass (slam.cpp:144-147, kalmanfilter.cpp:96-130) ---------------------------
      if (cur[6] != 0.0) {
        if (tid == 0) {
          sm.cres = ekf_compass_residual(sm.xs[2], cur[3], k);
          sm.cS = sm.s2[2] + cur[4];
        }
        __syncthreads();
        {
          const double res = sm.cres, S = sm.cS, invS = 1 / S, sq = sqrt(fabs(S));
          for (int r = tid; r < C::NI; r += C::THREADS) {
            const double Ki = invS * sm.s2[r];
            sm.xs[r] = sm.xs[r] + res * Ki;
            sm.W[widx<NB>(r)] = make_double2(sq * Ki, 0.0);
          }
        }
        __syncthreads();
        if (is_tile && 4 * I < 4 + 2 * n_lm) tile_downdate<NB, 1>(sm, p, I, J, sm.cS < 0 ? 1.0 : -1.0, 0.0);
        publish<NB>(sm, p, is_tile, I, J);
        setup_valid = false;                 // pose and P_RR changed
        __syncthreads();
      }

      // ---- doUpdate per measurement (slam.cpp:150-171, Update.cpp:80-195) -----------------------
      for (int m = 0; m < M; ++m) {
        int decision = EKF_DEC_NONE, index = -1;
        double mahal = 0.0;
        if (m < nz) {
          const double* zr = cur + 8 + 6 * m;
          if (!setup_valid) {   // compass update or an earlier measurement of this step moved the state
            if (sc_trig) {
              double PRR[9];
              for (int i = 0; i < 3; ++i) { PRR[i] = sm.s0[i]; PRR[i + 3] = sm.s1[i]; PRR[i + 6] = sm.s2[i]; }
              UpdateSetup u;
              ekf_build_setup(u, sm.xs[2], sm.xs[0], sm.xs[1], PRR, zr[0], zr[1], zr + 2);
              sm.upd = u;
            }
            __syncthreads();
          }
          setup_valid = false;  // whatever happens next, a further measurement needs a fresh set-up
          // ---- gating: group A (warps 0..GA-1) and group B (warps GA..2GA-1), one landmark per lane --
          EKF_FINE0
          if (warp < 2 * C::GA) {
            const bool groupA = warp < C::GA;
            const int lm = (groupA ? warp : warp - C::GA) * 32 + lane;
            const int Li = 4 + 2 * lm;
            const bool have = lm < n_lm;
            const UpdateSetup& u = sm.upd;
            GatePre pre;
            double pp[6];
            if (have) {
              pp[0] = sm.s0[Li]; pp[1] = sm.s0[Li + 1];
              pp[2] = sm.s1[Li]; pp[3] = sm.s1[Li + 1];
              pp[4] = sm.s2[Li]; pp[5] = sm.s2[Li + 1];
              ekf_gate_prelude(u, sm.xs[Li], sm.xs[Li + 1], pre);
            }
            EKF_FINE(0)   // loads + prelude
            double t12[4];
            if (groupA) {
              if (have) ekf_gate_terms12(u, pre, pp, t12);
            } else if (have) {
              const int pr = Li >> 1;
              const double pll[4] = {sm.d00[pr], sm.d10[pr], sm.d10[pr], sm.d11[pr]};
              double t3[4], t4[4];
              ekf_gate_terms34(u, pre, pp, pll, t3, t4);
#pragma unroll
              for (int q = 0; q < 4; ++q) { sm.t34[q][lm] = t3[q]; sm.t34[4 + q][lm] = t4[q]; }
            }
            EKF_FINE(1)   // terms
            named_barrier(1, 2 * C::GA * 32);
            EKF_FINE(2)   // group barrier
            if (groupA) {
              double val = INFINITY;
              int idx = INT_MAX;
              GateResult g;
              if (have) {
                double t3[4], t4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { t3[q] = sm.t34[q][lm]; t4[q] = sm.t34[4 + q][lm]; }
                ekf_gate_finish(u, pre, t12, t3, t4, k.cond_max, g);
                const bool valid = !g.skip && (k.mahal_init > g.d2);   // Update.cpp:131,140
                if (valid) { val = g.d2; idx = Li; }
              }
              EKF_FINE(3)   // finish
              // warp argmin with lowest-index tie break (Update.cpp:140) on an order-preserving integer
              // key of the distance: three REDUX operations instead of five shuffle rounds.
              const int my_idx = idx;
              {
                const double v = val + 0.0;                        // -0 -> +0 (they compare equal)
                const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
                const unsigned long long key = (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
                const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
                const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
                const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
                const bool best = hi == mhi && lo == mlo;
                idx = (int)__reduce_min_sync(0xffffffffu, best ? (unsigned)my_idx : (unsigned)INT_MAX);
              }
              Candidate& cd = sm.cand[warp];
              if (idx == INT_MAX) {
                if (lane == 0) { cd.val = INFINITY; cd.idx = INT_MAX; }
              } else if (my_idx == idx) {
                cd.val = val; cd.idx = idx;
                cd.res[0] = g.res0; cd.res[1] = g.res1;
                cd.S[0] = g.S[0]; cd.S[1] = g.S[1]; cd.S[2] = g.S[2]; cd.S[3] = g.S[3];
                cd.h3[0] = g.h3_0; cd.h3[1] = g.h3_1;
              }
            }
          }
          EKF_FINE(5)   // candidate store
          __syncthreads();
          EKF_FINE(6)   // CTA barrier
          EKF_PHASE(2)   // gating
          // ---- decision (Update.cpp:152,181,191), uniform over the CTA -----------------------------
          int wsel = 0;
          {
            double val = sm.cand[0].val;
            int idx = sm.cand[0].idx;
#pragma unroll
            for (int w = 1; w < C::GA; ++w) {
              const double ov = sm.cand[w].val;
              const int oi = sm.cand[w].idx;
              if (ov < val || (ov == val && oi < idx)) { val = ov; idx = oi; wsel = w; }
            }
            const int opt_i = (idx == INT_MAX) ? 0 : idx;
            mahal = (idx == INT_MAX) ? k.mahal_init : val;
            decision = ekf_decide(opt_i, mahal, k);
            if (decision == EKF_DEC_NEW && n_lm >= a.st.cap_lm) decision = EKF_DEC_DROPPED;
            index = opt_i ? opt_i - 1 : 0;   // external state index
          }
          const Candidate& cd = sm.cand[wsel];

          if (decision == EKF_DEC_OLD) {
            if (tid == C::THREADS - 1) {   // S^-1 and L D L^T of the winning S, off the critical path
              const double Sm[4] = {cd.S[0], cd.S[1], cd.S[2], cd.S[3]};
              double Si[4];
              ekf_inv2(Sm, Si);
              Post& po = sm.post;
              po.Si[0] = Si[0]; po.Si[1] = Si[1]; po.Si[2] = Si[2]; po.Si[3] = Si[3];
              const double d0 = Sm[0], l = Sm[1] / Sm[0], d1 = Sm[3] - l * Sm[1];
              po.l = l;
              po.sq0 = sqrt(fabs(d0));
              po.sq1 = sqrt(fabs(d1));
              po.m0 = d0 < 0 ? 1.0 : -1.0;
              po.m1 = d1 < 0 ? 1.0 : -1.0;
            }
            // ---- publish the two covariance columns of landmark Li -------------------------------
            const int Li = cd.idx;
            const int JL = Li >> 3, c = Li & 7, IL = Li >> 2, a0 = Li & 3;
            if (is_tile && J == JL && 4 * I + 3 >= Li) {     // rows >= Li: column c of the tiles below
              static_switch<2, 8>(c, [&](auto CC) {
                constexpr int c0 = decltype(CC)::value;
#pragma unroll
                for (int aa = 0; aa < 4; ++aa)
                  if (4 * I + aa >= Li) { sm.ca[4 * I + aa] = p[aa][c0]; sm.cb[4 * I + aa] = p[aa][c0 + 1]; }
              });
            }
            if (is_tile && I == IL) {                        // rows < Li: row a0 of the tiles to the left
              static_switch<2, 4>(a0, [&](auto AA) {
                constexpr int r0 = decltype(AA)::value;
#pragma unroll
                for (int b = 0; b < 8; ++b)
                  if (8 * J + b < Li) { sm.ca[8 * J + b] = p[r0][b]; sm.cb[8 * J + b] = p[r0 + 1][b]; }
              });
            }
            __syncthreads();
            EKF_PHASE(3)   // decision + column publish
            // ---- gain, state correction, downdate vectors (Update.cpp:186-187) --------------------
            {
              const UpdateSetup& u = sm.upd;
              const Post& po = sm.post;
              const double h00 = u.mCt[0], h01 = u.mCt[2], h02 = cd.h3[0];
              const double h10 = u.mCt[1], h11 = u.mCt[3], h12 = cd.h3[1];
              const double c00 = u.Ct[0], c10 = u.Ct[2], c01 = u.Ct[1], c11 = u.Ct[3];
              const double si0 = po.Si[0], si1 = po.Si[1], si2 = po.Si[2], si3 = po.Si[3];
              const double r0 = cd.res[0], r1 = cd.res[1], l = po.l, sq0 = po.sq0, sq1 = po.sq1;
              const int n_int = 4 + 2 * n_lm;
              for (int r = tid; r < C::NI; r += C::THREADS) {
                double2 w = make_double2(0.0, 0.0);
                if (r < n_int) {
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
                  w = make_double2(sq0 * fma(l, K1, K0), sq1 * K1);
                }
                sm.W[widx<NB>(r)] = w;
              }
            }
            cp_async_wait_all();   // next step's record (prefetched at step start) is visible after the barrier
            __syncthreads();
            EKF_PHASE(4)   // gain rows
            // x is final for this step if this was its last measurement: run the next step's scalar
            // chains now, on two threads, while every warp does its covariance downdate.
            if (m == nz - 1 && t + 1 < T) {
              scalar_chains(sm.rec[(t + 1) & 1]);
              scalar_done = true;
            }
            // ---- covariance downdate in registers (Update.cpp:188,193-194) -------------------------
            if (is_tile && 4 * I < 4 + 2 * n_lm) tile_downdate<NB, 2>(sm, p, I, J, sm.post.m0, sm.post.m1);
            publish<NB>(sm, p, is_tile, I, J);
            __syncthreads();
            EKF_PHASE(5)   // downdate + publish
          } else if (decision == EKF_DEC_NEW) {
            // ---- state augmentation (Update.cpp:152-178) -----------------------------------------
            const int r0i = 4 + 2 * n_lm;            // internal index of the new landmark
            if (tid == 0) {
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
            __syncthreads();
            const int In = r0i >> 2, a0 = r0i & 3, c0n = r0i & 7;
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
              static_switch<2, 4>(a0, [&](auto AA) {
                constexpr int r0 = decltype(AA)::value;
#pragma unroll
                for (int b = 0; b < 8; ++b)
                  if (8 * J + b < r0i) { p[r0][b] = o0[b]; p[r0 + 1][b] = o1[b]; }
              });
              if (J == (In >> 1)) {
                static_switch<2, 4>(a0, [&](auto AA) {
                  constexpr int r0 = decltype(AA)::value;
                  static_switch<2, 8>(c0n, [&](auto CC) {
                    constexpr int cc = decltype(CC)::value;
                    p[r0][cc] = sm.PLL[0]; p[r0 + 1][cc] = off; p[r0][cc + 1] = off; p[r0 + 1][cc + 1] = sm.PLL[3];
                  });
                });
              }
            }
            if (tid == 0) {
              sm.xs[r0i] = sm.nl[0];
              sm.xs[r0i + 1] = sm.nl[1];
            }
            index = r0i - 1;
            n_lm += 1;
            publish<NB>(sm, p, is_tile, I, J);
            __syncthreads();
          } else {
            if (decision == EKF_DEC_DROPPED) { dropped = 1; index = -1; }
            __syncthreads();   // the candidate slots are rewritten by the next gating pass
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
      EKF_PHASE(6)   // trace output, record prefetch wait, end-of-step barrier
    }
#ifdef EKF_TILE_TIMING
    if (timing)
      for (int i = 0; i < 8; ++i) a.phase_cycles[i] += tacc[i];
#endif
#ifdef EKF_FINE_TIMING
    if (timing)
      for (int i = 0; i < 8; ++i) a.phase_cycles[8 + i] += facc[i];
#endif

    // ---- write back to HBM (external layout, both triangles) -------------------------------------
    {
      const int n_int = 4 + 2 * n_lm;
      if (is_tile) {
#pragma unroll
        for (int b = 0; b < 8; ++b)
#pragma unroll
          for (int aa = 0; aa < 4; ++aa) {
            const int r = 4 * I + aa, c = 8 * J + b;
            if (r != 3 && c != 3 && r < n_int && c <= r) {
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

constexpr int kPhaseCounters = 16;
long long* g_phase_cycles[64] = {nullptr};   // one buffer per device, set by ekf_tile_phase_cycles()
inline long long*& phase_buf() {
  int dev = 0;
  cudaGetDevice(&dev);
  return g_phase_cycles[dev & 63];
}

template <int NB>
cudaError_t launch_tile(RunArgs a, int sm_count, cudaStream_t stream) {
  using C = TileCfg<NB>;
  static int grid_cap = 0;   // occupancy is a property of the kernel binary; every device here is a B200
  a.phase_cycles = phase_buf();
  if (!grid_cap) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ekf_batch_tile_kernel<NB>, C::THREADS, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    grid_cap = per_sm * sm_count;
  }
  const int grid = a.st.F < grid_cap ? a.st.F : grid_cap;
  ekf_batch_tile_kernel<NB><<<grid, C::THREADS, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace

int ekf_tile_max_landmarks() { return TileCfg<16>::MAX_LM; }

// Profiling aid: enable per-phase cycle accumulation (CTA 0 of every following launch) and read
// the eight counters back. Passing out == nullptr only enables.
cudaError_t ekf_tile_phase_cycles(long long* out) {
  long long*& buf = phase_buf();
  if (!buf) {
    cudaError_t e = cudaMalloc(&buf, kPhaseCounters * sizeof(long long));
    if (e != cudaSuccess) return e;
    cudaMemset(buf, 0, kPhaseCounters * sizeof(long long));
  }
  if (out) {
    cudaError_t e = cudaMemcpy(out, buf, kPhaseCounters * sizeof(long long), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return e;
    cudaMemset(buf, 0, kPhaseCounters * sizeof(long long));
  }
  return cudaSuccess;
}

cudaError_t ekf_tile_run(const EkfState& st, const EkfRunIO& io, const EkfConst& k, int sm_count, cudaStream_t stream) {
  RunArgs a{st, io, k, nullptr};
  if (st.cap_lm <= TileCfg<13>::MAX_LM) return launch_tile<13>(a, sm_count, stream);
  if (st.cap_lm <= TileCfg<14>::MAX_LM) return launch_tile<14>(a, sm_count, stream);
  if (st.cap_lm <= TileCfg<15>::MAX_LM) return launch_tile<15>(a, sm_count, stream);
  if (st.cap_lm <= TileCfg<16>::MAX_LM) return launch_tile<16>(a, sm_count, stream);
  return cudaErrorInvalidValue;
}
