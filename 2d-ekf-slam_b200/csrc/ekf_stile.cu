// ekf_stile.cu — regime A, fused multi-step kernel with the covariance in SHARED MEMORY as a
// tiled lower block triangle (sm_100a). One CTA per filter, persistent over filters.
//
// The north star's "one CTA per filter instance, covariance resident in shared memory", laid
// out so that FOUR filters fit on an SM at once (N_cap <= 50): only the lower block triangle of
// the (padded) covariance is stored, as NB(NB+1) tiles of 4x8 doubles in plane-major order
// T[a + 4b][tile] (48.1 KB with the padded plane stride, instead of 84.9 KB for the full matrix).
// Element (r, c), r >= c, lives in tile (r/4, c/8); symmetric accesses swap indices. The O(n)
// phases read the entries they need directly from this storage (no staging copies); the O(n^2)
// downdate is one thread per tile: 32 conflict-free loads, 64 fma, 32 stores. Registers hold
// nothing across phases, so the kernel runs at 72 registers per thread (224 threads: six tile warps
// and a helper warp) and four CTAs fill both the shared memory and the register file of an SM:
// with ~8000 cycles of dependent FP64 latency per filter-step, throughput is filters-in-flight per SM.
// The covariance enters and leaves shared memory column by column (contiguous HBM accesses).
//
// Internally the state is padded with one dummy entry after the robot pose
// ([X,Y,Phi,pad,L1x,L1y,...]) so every landmark pair is 2-aligned and never straddles a tile;
// the pad row/column of P is identically zero. External layout (C ABI, HBM) is unchanged.
//
// Phase structure per step (slam.cpp:130-182 order), same arithmetic as the other kernels
// (ekf_small.cuh), bit-identical results:
//   scalar chains   two lanes of the helper warp share one sincos stream: odometry -> Q, Phi,
//                   G, new pose (kalmanfilter.cpp:17-37, Propagate.cpp:33-48) and the rotation blocks
//                   of the update (Update.cpp:89-95); they run during the previous step's downdate
//   propagate       nine helper lanes do the 3x3 robot block (Propagate.cpp:53,66-67) and the q
//                   partial sums; one thread per strip row applies Phi (Propagate.cpp:56-60)
//   gating          Update.cpp:103-148, one landmark per lane on two warp groups in parallel
//                   (A: H_R P_RR H_R^T + H_Li P_LiR H_R^T, B: the other two terms of S), REDUX
//                   warp argmin, lowest index wins ties
//   gain            one state row per thread: P H^T (before S^-1 is known), then K, x, W
//                   (Update.cpp:186-187); S^-1 and L D L^T of S come from a helper lane meanwhile
//   downdate        one tile per thread, P_tile += u_rows (x) W_cols (Update.cpp:188,193-194 in the
//                   bit-symmetric form described in ekf_cta.cuh)
#include "ekf_cta.cuh"
#include "ekf_internal.h"

namespace {

#ifdef EKF_STILE_TIMING
// Profiling builds only: lane 0 of every warp of CTA 0 stamps its ARRIVAL at each barrier of one
// step (first filter, step 500), so the phase critical paths can be read as max-over-warps deltas.
__device__ long long g_stile_ts[8][16];
#define STILE_TS(kk)                                                                              \
  do {                                                                                            \
    if (blockIdx.x == 0 && f == 0 && t == 500 && (threadIdx.x & 31) == 0) g_stile_ts[threadIdx.x >> 5][kk] = clock64(); \
  } while (0)
#else
#define STILE_TS(kk) do { } while (0)
#endif

template <int NB>
struct STileCfg {
  static constexpr int NI = 8 * NB;                    // padded internal dimension
  static constexpr int NTILES = NB * (NB + 1);         // tiles (I,J), I = r/4, J = c/8, J <= I/2
  // Plane stride of the tile storage T[a + 4b][PS] in doubles: the smallest value >= NTILES that is
  // 4 or 12 (mod 16). One tile per lane is conflict-free for any stride; the row-wise accesses of
  // the O(n) phases (lane = row r: plane r&3, tile r>>2; lane = landmark: planes {0,2} / {1,3})
  // are conflict-free only when 4 consecutive planes land 4 double-banks apart.
  static constexpr int PS = NTILES + ((4 - NTILES % 8) % 8 + 8) % 8;
  // One warp more than the tiles need: the helper warp (scalar chains, robot block, S^-1) owns no
  // tile, so its sincos chain runs beside the other warps' downdate instead of in front of its own
  // (measured +4 %; 72 registers per thread at four CTAs per SM).
  static constexpr int THREADS = (NTILES + 31) / 32 * 32 + 32;
  static constexpr int NW = THREADS / 32;
  static constexpr int MAX_LM = (NI - 4) / 2;
  static constexpr int GA = (MAX_LM + 31) / 32;        // warps per gating group
  static constexpr int LMP = GA * 32;
  static constexpr int SP0 = 0;                        // first helper lane of the last warp
  static constexpr int MINB = NB == 13 ? 4 : (NB == 14 ? 3 : 2);
  static_assert(2 * GA <= NW - 1, "gating groups must not use the helper warp");
  static_assert(NTILES <= (NW - 1) * 32 && SP0 + 9 <= 32, "the helper warp owns no tile");
  static_assert(NI <= THREADS, "one state row per thread");
};

struct Candidate {          // best landmark of one gating warp (Opt_* of Update.cpp:140-147)
  double val;
  int idx;
  int pad;
  double res[2], S[4], h3[2];
};

struct Post {
  double Si[4];
  double l, sq0, sq1, m0, m1;
};

template <int NB>
struct STileSmem {          // everything except the covariance tiles and the record buffers
  using C = STileCfg<NB>;
  double xs[C::NI];
  double2 W[C::NI];
  double t34[8][C::LMP];
  PropSetup prop;
  double PhiS[9], GS[6];
  double xnew[3];
  UpdateSetup upd;
  Candidate cand[C::GA];
  Post post;
  double nl[2], PLL[4], h3n[2];
  double cres, cS;
};

struct RunArgs {
  EkfState st;
  EkfRunIO io;
  EkfConst k;
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

// Tiles are numbered column block by column block: column block J holds row blocks I = 2J..2NB-1,
// so tile (I, J) is number J*(2NB) - J*(J-1) + (I - 2J). Consecutive threads then share J (their
// W_col loads are warp-uniform broadcasts) and the strip tiles (J = 0) are tiles 0..2NB-1 in row
// order (the gating / gain accesses to P(r, 0..2) are conflict-free).
template <int NB>
__device__ __forceinline__ int tile_number(int I, int J) { return J * (2 * NB) - J * (J - 1) + (I - 2 * J); }
// index of P(r, c), r >= c, in the plane-major tile storage
template <int NB>
__device__ __forceinline__ int pidx_lower(int r, int c) {
  return ((r & 3) + 4 * (c & 7)) * STileCfg<NB>::PS + tile_number<NB>(r >> 2, c >> 3);
}
template <int NB>
__device__ __forceinline__ int pidx(int r, int c) { return r >= c ? pidx_lower<NB>(r, c) : pidx_lower<NB>(c, r); }
template <int NB>
__device__ __forceinline__ int widx(int r) { return (r & 7) * NB + (r >> 3); }   // see ekf_tile.cu

template <int NB, int RANK>
__device__ __forceinline__ void tile_downdate(double* __restrict__ Tt, const double2* __restrict__ W, int I, int J,
                                              double m0, double m1) {
  constexpr int NT = STileCfg<NB>::PS;   // plane stride
  double2 wi[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) wi[a] = W[(4 * (I & 1) + a) * NB + (I >> 1)];
  double u0[4], u1[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    u0[a] = m0 * wi[a].x;
    u1[a] = m1 * wi[a].y;
  }
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const double2 wj = W[b * NB + J];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double t = Tt[(a + 4 * b) * NT];
      if (RANK == 2) t = fma(u1[a], wj.y, t);
      t = fma(u0[a], wj.x, t);
      Tt[(a + 4 * b) * NT] = t;
    }
  }
}

// MODE 0: plain (a New association at capacity is dropped and flagged). MODE 1: the fast small-tile
// instance of the growth path - a filter whose map outgrows the tiles is written back and parked
// (io.resume). MODE 2: the continuation - visits only the parked filters and resumes them at the
// measurement where they stopped. Separate instantiations, so the plain and the parking kernels do
// not carry the resume logic.
template <int NB, int MODE>
__global__ void __launch_bounds__(STileCfg<NB>::THREADS, STileCfg<NB>::MINB) ekf_batch_stile_kernel(const RunArgs a) {
  using C = STileCfg<NB>;
  constexpr bool CAN_PARK = MODE == 1, CAN_RESUME = MODE == 2;
  constexpr int NT = C::PS, SP0 = C::SP0;   // NT: plane stride of the tile storage
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* T = reinterpret_cast<double*>(smem_raw);                               // [32][NT]
  STileSmem<NB>& sm = *reinterpret_cast<STileSmem<NB>*>(smem_raw + (size_t)32 * NT * sizeof(double));
  double* recbuf = reinterpret_cast<double*>(smem_raw + (size_t)32 * NT * sizeof(double) +
                                             ((sizeof(STileSmem<NB>) + 15) & ~(size_t)15));   // [2][Lp]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_tile = tid < C::NTILES;
  const bool helper_warp = warp == C::NW - 1;
  const bool sc_prop = helper_warp && lane == SP0;
  const bool sc_trig = helper_warp && lane == SP0 + 1;
  const bool post_lane = helper_warp && lane == SP0 + 9;
  int I = 0, J = 0;
  if (is_tile) {                                     // inverse of tile_number
    int t = tid;
    while (t >= 2 * NB - 2 * J) { t -= 2 * NB - 2 * J; ++J; }
    I = 2 * J + t;
  }
  double* Tt = T + tid;                              // this thread's tile: element (a,b) at Tt[(a+4b)*NT]
  const int rowbase = pidx_lower<NB>(tid < C::NI ? tid : 0, 0);   // P(tid, 0); P(tid, j) = + 4*j*NT
  const int ld = a.st.ld, L = a.io.L, T_steps = a.io.T, M = a.io.M;
  const int Lp = (L + 1) & ~1;
  const EkfConst& k = a.k;

  for (int f = blockIdx.x; f < a.st.F; f += gridDim.x) {
    double* gP = a.st.P + (size_t)f * a.st.slab;
    double* gx = a.st.x + (size_t)f * a.st.xs;
    const double* grec = a.io.records + (size_t)f * T_steps * L;
    int n_lm = a.st.nlm[f];
    int dropped = 0;
    int t_begin = 0, m_begin = 0, parked = 0;
    bool resumed = false;
    if (CAN_RESUME) {
      const int code = a.io.resume[f];
      if (code == 0) continue;                           // finished (or being run) by the primary launch
      if (a.io.continuation == 2 ? code > 0 : code < 0) continue;   // 2: only maps that were too large at launch; 1: only parked ones
      if (code > 0) { t_begin = (code - 1) / (M + 1); m_begin = (code - 1) % (M + 1); resumed = true; }
    }
    if (CAN_PARK && n_lm > C::MAX_LM) continue;          // too large for these tiles at launch: marked, the concurrent continuation runs it
    {
      const int n_int = 4 + 2 * n_lm;
      // Column by column (a warp per column, a lane per row): the HBM reads are contiguous runs and
      // the scatter into the plane-major tiles is conflict-free. Column c of the stored triangle
      // starts at row 8*(c/8) (its diagonal tile is stored whole); dead entries are zeroed.
      // The copies are cp.async (global -> shared without passing through registers), so a lane has
      // its whole share of the covariance in flight at once instead of eight loads per pass.
      for (int c = warp; c < C::NI; c += C::NW) {
        const bool col_live = c != 3 && c < n_int;
        const double* gc = gP + (size_t)ext_index(c) * ld;
        const int rlo = c & ~7, cb = 4 * (c & 7) * NT, cj = c >> 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int r = rlo + lane + 32 * k;
          if (r < C::NI) {
            double* dst = &T[(r & 3) * NT + cb + tile_number<NB>(r >> 2, cj)];
            if (col_live && r != 3 && r < n_int) cp_async8(dst, gc + ext_index(r));
            else *dst = 0.0;
          }
        }
      }
      for (int r = tid; r < C::NI; r += C::THREADS) sm.xs[r] = (r != 3 && r < n_int) ? gx[ext_index(r)] : 0.0;
      for (int i = tid; i < L; i += C::THREADS) cp_async8(recbuf + (size_t)(t_begin & 1) * Lp + i, grec + (size_t)t_begin * L + i);   // t_begin = 0 unless resumed
      cp_async_wait_all();
    }
    __syncthreads();

    auto scalar_chains = [&](const double* rec) {
      if (sc_prop || sc_trig) {
        const double RTV = rec[1] * k.deg2rad_pi / 180.0;
        const double phi = sc_prop ? sm.xs[2] : sm.xs[2] + rec[2] * RTV;   // same expression as the pose update
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
          if ((int)rec[5] > 0) {
            u.z0 = rec[8]; u.z1 = rec[9];
            for (int q = 0; q < 4; ++q) u.R[q] = rec[10 + q];
          }
        }
      }
    };
    bool scalar_done = false;

    for (int t = CAN_RESUME ? t_begin : 0; t < T_steps && !(CAN_PARK && parked); ++t) {
      const double* cur = recbuf + (size_t)(t & 1) * Lp;
      const bool skip_front = CAN_RESUME && resumed && t == t_begin;   // a resumed filter: this step is already propagated
      if (t + 1 < T_steps) {
        const double* g = grec + (size_t)(t + 1) * L;
        double* nxt = recbuf + (size_t)((t + 1) & 1) * Lp;
        for (int i = tid; i < L; i += C::THREADS) cp_async8(nxt + i, g + i);
      }
      const int nz = min((int)cur[5], (L - 8) / 6);   // never read past the record's measurement slots
      bool rec_ready = false;
      STILE_TS(0);
      if (!skip_front) {
      // ---- doPropagation (slam.cpp:136) ------------------------------------------------------------
      if (!scalar_done) {
        scalar_chains(cur);
        __syncthreads();
      }
      scalar_done = false;
      if (sc_prop) {
        sm.xs[0] = sm.xnew[0]; sm.xs[1] = sm.xnew[1]; sm.xs[2] = sm.xnew[2];
        sm.upd.x0 = sm.xnew[0]; sm.upd.x1 = sm.xnew[1];
      }
      if (helper_warp) {
        // 3x3 robot block, one element per spare lane (Propagate.cpp:53, then :66-67)
        const int e = (lane + 36 - SP0) % 9, i = e % 3, j = e / 3;
        double PRR[9];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int r = 0; r < 3; ++r) PRR[r + 3 * c] = T[(r + 4 * c) * NT];      // tile (0,0)
        const double mij = ekf_prop_prr_elem(sm.PhiS, sm.GS, sm.prop.Q, PRR, i, j);
        const double mji = __shfl_sync(0xffffffffu, mij, SP0 + j + 3 * i);
        const double pn = 0.5 * (mij + mji);
        const int qe = e % 6, qi = qe % 2, qj = qe / 2;
        const double p0j = __shfl_sync(0xffffffffu, pn, SP0 + 0 + 3 * qj);
        const double p1j = __shfl_sync(0xffffffffu, pn, SP0 + 1 + 3 * qj);
        const double qv = sm.upd.mCt[qi] * p0j + sm.upd.mCt[qi + 2] * p1j;
        if (lane >= SP0 && lane < SP0 + 9) {
          T[(i + 4 * j) * NT] = pn;            // both triangles of the 3x3 block live in tile (0,0)
          sm.upd.PRR[e] = pn;
          if (e < 6) sm.upd.q[qe] = qv;
        }
      }
      if (tid >= 4 && tid < C::NI) {           // P_RL <- Phi*P_RL (Propagate.cpp:56-60), strip row tid
        double a0 = T[rowbase], a1 = T[rowbase + 4 * NT], a2 = T[rowbase + 8 * NT];
        ekf_prop_col(sm.prop, a0, a1, a2);
        T[rowbase] = a0; T[rowbase + 4 * NT] = a1; T[rowbase + 8 * NT] = a2;
      }
      STILE_TS(1);
      __syncthreads();
      }   // !skip_front

      bool setup_valid = !skip_front;
      // ---- doUpdateCompass (slam.cpp:144-147, kalmanfilter.cpp:96-130) ---------------------------
      if (!skip_front && cur[6] != 0.0) {
        if (tid == 0) {
          sm.cres = ekf_compass_residual(sm.xs[2], cur[3], k);
          sm.cS = T[(2 + 4 * 2) * NT] + cur[4];
        }
        __syncthreads();
        if (tid < C::NI) {
          const double res = sm.cres, S = sm.cS, invS = 1 / S, sq = sqrt(fabs(S));
          const double Ki = invS * T[pidx<NB>(tid, 2)];
          sm.xs[tid] = sm.xs[tid] + res * Ki;
          sm.W[widx<NB>(tid)] = make_double2(sq * Ki, 0.0);
        }
        __syncthreads();
        if (is_tile && 4 * I < 4 + 2 * n_lm) tile_downdate<NB, 1>(Tt, sm.W, I, J, sm.cS < 0 ? 1.0 : -1.0, 0.0);
        setup_valid = false;
        __syncthreads();
      }

      // ---- doUpdate per measurement (slam.cpp:150-171, Update.cpp:80-195) -----------------------
      for (int m = (CAN_RESUME && skip_front) ? m_begin : 0; m < M; ++m) {
        int decision = EKF_DEC_NONE, index = -1;
        double mahal = 0.0;
        if (m < nz) {
          const double* zr = cur + 8 + 6 * m;
          if (!setup_valid) {
            if (sc_trig) {
              double PRR[9];
              for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r) PRR[r + 3 * c] = T[(r + 4 * c) * NT];
              UpdateSetup u;
              ekf_build_setup(u, sm.xs[2], sm.xs[0], sm.xs[1], PRR, zr[0], zr[1], zr + 2);
              sm.upd = u;
            }
            __syncthreads();
          }
          setup_valid = false;
          // ---- gating: group A (warps 0..GA-1), group B (warps GA..2GA-1), one landmark per lane ---
          if (warp < 2 * C::GA) {
            const bool groupA = warp < C::GA;
            const int lm = (groupA ? warp : warp - C::GA) * 32 + lane;
            const int Li = 4 + 2 * lm;
            const bool have = lm < n_lm;
            const UpdateSetup& u = sm.upd;
            GatePre pre;
            double pp[6];
            if (have) {
              const int sb = pidx_lower<NB>(Li, 0);          // P(Li,0); rows Li, Li+1 share a tile
#pragma unroll
              for (int j = 0; j < 3; ++j) { pp[2 * j] = T[sb + 4 * j * NT]; pp[2 * j + 1] = T[sb + NT + 4 * j * NT]; }
              ekf_gate_prelude(u, sm.xs[Li], sm.xs[Li + 1], pre);
            }
            double t12[4];
            if (groupA) {
              if (have) ekf_gate_terms12(u, pre, pp, t12);
            } else if (have) {
              const int db = pidx_lower<NB>(Li, Li);         // P(Li,Li); (Li+1,Li) = +NT; (Li+1,Li+1) = +5NT
              const double p10 = T[db + NT];
              const double pll[4] = {T[db], p10, p10, T[db + 5 * NT]};
              double t3[4], t4[4];
              ekf_gate_terms34(u, pre, pp, pll, t3, t4);
#pragma unroll
              for (int q = 0; q < 4; ++q) { sm.t34[q][lm] = t3[q]; sm.t34[4 + q][lm] = t4[q]; }
            }
            STILE_TS(2);
            named_barrier(1, 2 * C::GA * 32);
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
              const int my_idx = idx;
              {   // warp argmin, lowest index wins ties (Update.cpp:140): order-preserving integer key + REDUX
                const double v = val + 0.0;
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
          STILE_TS(3);
          __syncthreads();
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
            if (decision == EKF_DEC_NEW && n_lm >= a.st.cap_lm) decision = EKF_DEC_DROPPED;   // (tiles are sized for >= cap_lm unless CAN_PARK)
            if (CAN_PARK && decision == EKF_DEC_NEW && n_lm >= C::MAX_LM)
              parked = 1 + t * (M + 1) + m;   // the map outgrows this kernel's tiles: a kernel with larger tiles continues here
            index = opt_i ? opt_i - 1 : 0;   // external state index
          }
          if (CAN_PARK && parked) break;     // nothing of this measurement has been applied (the decision is CTA-uniform)
          const Candidate& cd = sm.cand[wsel];
          STILE_TS(9);

          if (decision == EKF_DEC_OLD) {
            const int Li = cd.idx;
            const int n_int = 4 + 2 * n_lm;
            if (post_lane) {   // S^-1 and L D L^T of the winning S, while the row threads form P H^T
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
            // ---- M = P[:,0:3] H_R^T + P[:,Li:Li+2] H_Li^T, one row per thread (Update.cpp:186) --------
            double M0 = 0.0, M1 = 0.0;
            const bool row = tid < n_int;
            if (row) {
              const UpdateSetup& u = sm.upd;
              const double h00 = u.mCt[0], h01 = u.mCt[2], h02 = cd.h3[0];
              const double h10 = u.mCt[1], h11 = u.mCt[3], h12 = cd.h3[1];
              const double c00 = u.Ct[0], c10 = u.Ct[2], c01 = u.Ct[1], c11 = u.Ct[3];
              double p0, p1, p2;
              if (tid >= 3) { p0 = T[rowbase]; p1 = T[rowbase + 4 * NT]; p2 = T[rowbase + 8 * NT]; }
              else { p0 = T[pidx<NB>(tid, 0)]; p1 = T[pidx<NB>(tid, 1)]; p2 = T[pidx<NB>(tid, 2)]; }
              const double pa = T[pidx<NB>(tid, Li)], pb = T[pidx<NB>(tid, Li + 1)];
              const double A0 = (p0 * h00 + p1 * h01) + p2 * h02;
              const double A1 = (p0 * h10 + p1 * h11) + p2 * h12;
              const double B0 = pa * c00 + pb * c10;
              const double B1 = pa * c01 + pb * c11;
              M0 = A0 + B0;
              M1 = A1 + B1;
            }
            STILE_TS(10);
            cp_async_wait_all();   // next step's record (prefetched at step start) is visible after the barrier
            rec_ready = true;
            STILE_TS(4);
            __syncthreads();
            // ---- gain, state correction, downdate vectors (Update.cpp:186-187) --------------------
            if (tid < C::NI) {
              double2 w = make_double2(0.0, 0.0);
              if (row) {
                const Post& po = sm.post;
                const double K0 = M0 * po.Si[0] + M1 * po.Si[1];
                const double K1 = M0 * po.Si[2] + M1 * po.Si[3];
                sm.xs[tid] = sm.xs[tid] + (K0 * cd.res[0] + K1 * cd.res[1]);
                w = make_double2(po.sq0 * fma(po.l, K1, K0), po.sq1 * K1);
              }
              sm.W[widx<NB>(tid)] = w;
            }
            STILE_TS(5);
            __syncthreads();
            // x is final for this step if this was its last measurement: run the next step's scalar
            // chains now, on two spare lanes, while every warp does its covariance downdate.
            if (m == nz - 1 && t + 1 < T_steps) {
              scalar_chains(recbuf + (size_t)((t + 1) & 1) * Lp);
              scalar_done = true;
            }
            // ---- covariance downdate (Update.cpp:188,193-194) -----------------------------------------
            if (is_tile && 4 * I < n_int) tile_downdate<NB, 2>(Tt, sm.W, I, J, sm.post.m0, sm.post.m1);
            STILE_TS(6);
            __syncthreads();
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
            if (tid < r0i) {                    // P_RLi = -P[:,0:3]*H_R^T*H_Li (:169), existing row tid
              const UpdateSetup& u = sm.upd;
              const double h00 = u.mCt[0], h01 = u.mCt[2], h02 = sm.h3n[0];
              const double h10 = u.mCt[1], h11 = u.mCt[3], h12 = sm.h3n[1];
              const double ct00 = u.Ct[0], ct10 = u.Ct[1], ct01 = u.Ct[2], ct11 = u.Ct[3];
              const double q0 = -T[pidx<NB>(tid, 0)], q1 = -T[pidx<NB>(tid, 1)], q2 = -T[pidx<NB>(tid, 2)];
              const double t0 = (q0 * h00 + q1 * h01) + q2 * h02;
              const double t1 = (q0 * h10 + q1 * h11) + q2 * h12;
              const int nb = pidx_lower<NB>(r0i, tid);         // P(r0i, tid); (r0i+1, tid) = + NT
              T[nb] = t0 * ct00 + t1 * ct10;
              T[nb + NT] = t0 * ct01 + t1 * ct11;
            }
            if (tid == 0) {
              const double off = 0.5 * (sm.PLL[2] + sm.PLL[1]);   // :193-194 on the new 2x2 block
              const int db = pidx_lower<NB>(r0i, r0i);
              T[db] = sm.PLL[0];
              T[db + NT] = off;
              T[db + 5 * NT] = sm.PLL[3];
              sm.xs[r0i] = sm.nl[0];
              sm.xs[r0i + 1] = sm.nl[1];
            }
            index = r0i - 1;
            n_lm += 1;
            __syncthreads();
          } else {
            if (decision == EKF_DEC_DROPPED) { dropped = 1; index = -1; }
            __syncthreads();   // the candidate slots are rewritten by the next gating pass
          }
        }
        if (tid == 0) {
          const size_t oi = ((size_t)f * T_steps + t) * M + m;
          if (a.io.decision) a.io.decision[oi] = decision;
          if (a.io.index) a.io.index[oi] = index;
          if (a.io.mahal) a.io.mahal[oi] = mahal;
        }
      }
      if (a.io.pose_trace && sc_prop && !(CAN_PARK && parked)) {   // slam.cpp:181; by the lane that overwrites the pose at the next step start
        double* pt = a.io.pose_trace + ((size_t)f * T_steps + t) * 3;
        pt[0] = sm.xs[0]; pt[1] = sm.xs[1]; pt[2] = sm.xs[2];
      }
      // The next step's record (prefetched at step start) must be visible to every thread. An Old
      // update already waited for it in front of one of its barriers and ended on a barrier (the
      // decision is CTA-uniform), so only the other paths pay for this one.
      if (!rec_ready) {
        cp_async_wait_all();
        STILE_TS(7);
        __syncthreads();
      }
      STILE_TS(8);
    }

    // ---- write back to HBM (external layout, both triangles) -------------------------------------
    {
      const int n_int = 4 + 2 * n_lm;
      // full columns of the external matrix, a warp per column and a lane per row (contiguous HBM
      // writes); entries above the diagonal come from the mirrored position of the stored triangle
      for (int c = warp; c < n_int; c += C::NW) {
        if (c == 3) continue;
        double* gc = gP + (size_t)ext_index(c) * ld;
#pragma unroll
        for (int k = 0; k < (C::NI + 31) / 32; ++k) {
          const int r = lane + 32 * k;
          if (r != 3 && r < n_int) gc[ext_index(r)] = T[pidx<NB>(r, c)];
        }
      }
      for (int r = tid; r < n_int; r += C::THREADS)
        if (r != 3) gx[ext_index(r)] = sm.xs[r];
      if (tid == 0) {
        a.st.nlm[f] = n_lm;
        if (dropped) a.st.status[f] |= 1;
        if (MODE != 0) a.io.resume[f] = parked;
      }
    }
    __syncthreads();
  }
}

template <int NB>
size_t stile_smem_bytes(int L) {
  return (size_t)32 * STileCfg<NB>::PS * sizeof(double) + ((sizeof(STileSmem<NB>) + 15) & ~(size_t)15) +
         (size_t)2 * ((L + 1) & ~1) * sizeof(double);
}

template <int NB, int MODE>
cudaError_t launch_stile(const RunArgs& a, int sm_count, cudaStream_t stream) {
  using C = STileCfg<NB>;
  const size_t bytes = stile_smem_bytes<NB>(a.io.L);
  // function attributes are per device: keep one configuration per device of this process
  static size_t configured_dev[64] = {0};
  static int grid_cap_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  size_t& configured = configured_dev[dev];
  int& grid_cap = grid_cap_dev[dev];
  if (bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(ekf_batch_stile_kernel<NB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ekf_batch_stile_kernel<NB, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ekf_batch_stile_kernel<NB, MODE>, C::THREADS, bytes);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    grid_cap = per_sm * sm_count;
    configured = bytes;
  }
  const int grid = a.st.F < grid_cap ? a.st.F : grid_cap;
  ekf_batch_stile_kernel<NB, MODE><<<grid, C::THREADS, bytes, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace

// Profiling builds (-DEKF_STILE_TIMING): the barrier-arrival stamps of one step, [warp][barrier].
cudaError_t ekf_stile_timestamps(long long* out128) {
#ifdef EKF_STILE_TIMING
  return cudaMemcpyFromSymbol(out128, g_stile_ts, sizeof(long long) * 128);
#else
  for (int i = 0; i < 128; ++i) out128[i] = 0;
  return cudaSuccess;
#endif
}

int ekf_stile_max_landmarks() { return STileCfg<16>::MAX_LM; }

// CTAs that run at once per SM for a given capacity (the filter chunking of the pipelined path uses it)
int ekf_stile_ctas_per_sm(int cap_lm) { return cap_lm <= STileCfg<13>::MAX_LM ? 4 : (cap_lm <= STileCfg<14>::MAX_LM ? 3 : 2); }

// tile_cap: the landmark count the tiles are sized for (0 = the handle's capacity). A smaller value
// runs the faster small-tile instance; with io.resume set, filters that outgrow it are parked for a
// continuation launch with tile_cap = 0.
cudaError_t ekf_stile_run(const EkfState& st, const EkfRunIO& io, const EkfConst& k, int sm_count, cudaStream_t stream,
                          int tile_cap) {
  RunArgs a{st, io, k};
  const int cap = tile_cap > 0 ? tile_cap : st.cap_lm;
  if (io.resume && !io.continuation) {                  // primary launch of the growth path: the fast instance, parking
    if (cap <= STileCfg<13>::MAX_LM) return launch_stile<13, 1>(a, sm_count, stream);
    return cudaErrorInvalidValue;
  }
  if (io.resume) {                                       // continuation: tiles for the handle's capacity
    if (cap <= STileCfg<14>::MAX_LM) return launch_stile<14, 2>(a, sm_count, stream);
    if (cap <= STileCfg<15>::MAX_LM) return launch_stile<15, 2>(a, sm_count, stream);
    if (cap <= STileCfg<16>::MAX_LM) return launch_stile<16, 2>(a, sm_count, stream);
    return cudaErrorInvalidValue;
  }
  if (cap <= STileCfg<13>::MAX_LM) return launch_stile<13, 0>(a, sm_count, stream);
  if (cap <= STileCfg<14>::MAX_LM) return launch_stile<14, 0>(a, sm_count, stream);
  if (cap <= STileCfg<15>::MAX_LM) return launch_stile<15, 0>(a, sm_count, stream);
  if (cap <= STileCfg<16>::MAX_LM) return launch_stile<16, 0>(a, sm_count, stream);
  return cudaErrorInvalidValue;
}
int ekf_stile_fast_landmarks() { return STileCfg<13>::MAX_LM; }

namespace {
__global__ void stile_mark_grown(const int* __restrict__ nlm, int* __restrict__ resume, int F, int fast_cap) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < F) resume[f] = nlm[f] > fast_cap ? -2 : 0;
}
}  // namespace
// resume[f] = -2 for the filters whose map is already beyond the fast tiles (they are run, from the start
// of the lap, by a continuation launch with io.continuation = 2 beside the primary launch), else 0.
cudaError_t ekf_stile_mark_grown(const int* nlm, int* resume, int F, cudaStream_t stream) {
  stile_mark_grown<<<(F + 255) / 256, 256, 0, stream>>>(nlm, resume, F, STileCfg<13>::MAX_LM);
  return cudaGetLastError();
}
