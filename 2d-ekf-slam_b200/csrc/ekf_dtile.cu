// ekf_dtile.cu — regime A, fused multi-step kernel with a DEFERRED covariance downdate (sm_100a).
// One CTA (128 threads = one warp per SM sub-partition) per filter, persistent over filters, four
// CTAs per SM, T steps with the covariance resident in shared memory.
//
// What is different from ekf_stile.cu (same arithmetic, bit-identical results):
//
//  * Deferred downdate. The landmark-landmark block P_LL is only MODIFIED by the rank-2 downdates
//    (Update.cpp:188,193-194) and only READ when one of the two landmarks is the associated one
//    (the two gain columns of Update.cpp:186) - the gating loop (Update.cpp:103-148) needs just the
//    robot strip P_LR, P_RR and the 2x2 diagonal blocks. So P_LL is kept as "stored tiles + up to
//    three pending rank-2 updates": each Old update appends its downdate vector W to a history ring
//    in shared memory, the strip / P_RR / diagonal blocks are updated eagerly (O(n)), the two gain
//    columns are corrected on the fly when read, and the tiles are swept once per three updates,
//    applying the pending updates to every element in order. Every element sees the same fma
//    sequence as with an immediate sweep, so the bits are the same; the O(n^2) shared-memory
//    traffic per update drops by three.
//  * Storage. Pose, P_RR (3x3), strip SR (3 x 2N), the 2x2 diagonal blocks Dd and the 2x2 block
//    between the two landmarks of an aligned pair Do live in small arrays ("eager" entries). The
//    rest of the lower triangle of P_LL lives in 4x12 tiles (row blocks of 4, column blocks of 12;
//    tile (I,J) exists for I >= 3J+1), plane-major T[(a+4b)*PS + b + slot]: a sweep thread owns one
//    tile (conflict-free for any plane), the column walk of the gain phase (consecutive rows) and
//    the row walk (consecutive columns) both spread over the banks (PS = 12 mod 16 and the +b skew).
//    108 tiles for 50 landmarks: four warps sweep them, one tile per lane.
//  * Gating: every landmark in the reference's exact operation order (ekf_small.cuh), two lanes per
//    landmark on all four warps - lane i evaluates row i of S (two elements, ekf_gate_S_element), the
//    pair exchanges them by shuffle and both finish the condition gate and the Mahalanobis distance;
//    S^-1 (a by-product of the distance) and the L D L^T factors of S are evaluated by every lane, so
//    the winner of the argmin writes a complete decision block and no serial "winner" phase follows.
//    The strip propagation (Propagate.cpp:56-60) rides on the gating lanes' loads.
//  * Fixed warp roles: warp 2 runs the landmark-independent scalar chains of the NEXT step (sincos of
//    the two headings, Phi, G, P_RR, pose) beside the gain phase of warps 1 and 3, so every SM
//    sub-partition's instruction cache sees one role's code.
//
// Phase structure per step (slam.cpp:130-182 order):
//   warp 2        record-only scalars (kalmanfilter.cpp:17-37, one step ahead), sincos of the two headings,
//                 Phi, G, new pose (Propagate.cpp:33-48), P_RR <- Phi P_RR Phi^T + G Q G^T (Propagate.cpp:53,
//                 66-67), the landmark-independent part of the update (Update.cpp:89-95)
//   all warps     (every third update) sweep of the tiles; gating with the strip propagation folded in
//                 (Propagate.cpp:56-60, Update.cpp:103-148), warp argmin, decision (Update.cpp:152,181,191)
//   warp 2        pose rows of the gain, pose correction, P_RR downdate - then on to the next step's chain
//   warps 1, 3    gain column correction, K, x, W (Update.cpp:186-187), eager downdate of strip, diagonal
//                 and pair blocks, one landmark (two rows) per thread
#include "ekf_cta.cuh"
#include "ekf_internal.h"

namespace {

constexpr int KH = 4;            // history ring: the tiles are swept when three updates are pending
constexpr int DT_THREADS = 128;

#ifdef EKF_DTILE_TIMING
__device__ long long g_dtile_ts[4][16];
#define DTILE_TS(kk)                                                                              \
  do {                                                                                            \
    if (blockIdx.x == 0 && f == 0 && t == 501 && (threadIdx.x & 31) == 0) g_dtile_ts[threadIdx.x >> 5][kk] = clock64(); \
  } while (0)
#else
#define DTILE_TS(kk) do { } while (0)
#endif

template <int RB>                // RB row blocks of 4 rows = 2*RB landmarks
struct DCfg {
  static constexpr int NQ = 4 * RB;                   // rows / columns of P_LL
  static constexpr int NL = 2 * RB;                   // landmark capacity
  static constexpr int NP = RB;                       // aligned landmark pairs
  static constexpr int NJ = (RB - 2) / 3 + 1;         // column blocks that own tiles
  static constexpr int NT = NJ * (RB - 1) - 3 * NJ * (NJ - 1) / 2;   // tiles
  static constexpr int PS = NT + ((4 - NT % 8) + 8) % 8;   // plane stride: NT rounded up to 4 or 12 (mod 16)
  static constexpr int TSIZE = 47 * PS + 11 + NT;     // doubles of tile storage (48 planes, skew b)
  // History vectors are stored row-in-block major: entry of row q at (q&3)*HP + (q>>2). Consecutive row
  // blocks (the sweep: one tile per lane) and consecutive rows (the gain phase) both read conflict-free
  // 16-byte words when HP = 2 (mod 8).
  static constexpr int HP = RB + ((2 - RB % 8) + 8) % 8;
  static constexpr int HN = 4 * HP;
  static_assert(PS % 16 == 4 || PS % 16 == 12, "plane stride must be 4 or 12 (mod 16)");
  static_assert(NT <= DT_THREADS, "one tile per thread");
  static_assert(NQ + 3 <= DT_THREADS, "one state row per thread");
  static_assert(NL <= 64, "two landmarks per gating lane");
};

template <int RB>
__device__ __forceinline__ int dt_hidx(int q) { return (q & 3) * DCfg<RB>::HP + (q >> 2); }
template <int RB>
__device__ __forceinline__ int dt_slot(int I, int J) { return J * (RB - 1) - 3 * J * (J - 1) / 2 + I - 3 * J - 1; }
// storage index of P_LL(r, c), r > c, r and c in different aligned pairs
template <int RB>
__device__ __forceinline__ int dt_rowterm(int r) { return (r & 3) * DCfg<RB>::PS + (r >> 2); }
template <int RB>
__device__ __forceinline__ int dt_colterm(int c) {
  const int J = (c * 43) >> 9;            // c / 12 for c < 128
  const int b = c - 12 * J;
  return 4 * b * DCfg<RB>::PS + b + dt_slot<RB>(0, J);
}
template <int RB>
__device__ __forceinline__ int dt_addr(int r, int c) { return dt_rowterm<RB>(r) + dt_colterm<RB>(c); }


// What the winner of a gating warp publishes: everything the gain phase needs (Opt_* of Update.cpp:140-147
// plus S^-1, the L D L^T factors of S and the address terms of the associated column pair).
struct DCand {
  double val;
  int idx, c0;                   // state index Opt_i (INT_MAX: none), first P_LL row of the landmark
  int rt_c0, ct_c0, sb_c0, pad;  // row / column address terms of that row, first tile slot of its column block
  double res[2], S[4], Si[4], h3[2];
  double l, sq0, sq1;
  unsigned sm0, sm1;             // sign-bit masks of the downdate (u = mask ^ W)
  double Ct[4], mCt[4];          // H_Li and the first two columns of H_R of THIS update (sm.upd is rewritten early)
  double xr[4];                  // pose after the update (Update.cpp:187, rows 0..2)
  double2 WR[3];                 // downdate vectors of the three pose rows
};

template <int RB>
struct DSmem {
  using C = DCfg<RB>;
  double2 H[KH][C::HN];          // pending downdate vectors, entry of row q at dt_hidx(q)
  double SR[3][C::NQ];           // strip: SR[j][q] = P(3+q, j)
  double Dd[3][C::NL];           // P(2l,2l), P(2l+1,2l), P(2l+1,2l+1)
  double Do[4][C::NP];           // Do[i+2j][m] = P(4m+2+i, 4m+j)
  double xl[C::NQ];              // landmark part of the state
  double2 WR[3];                 // downdate vectors of the three pose rows of the current update
  double xr[4];                  // pose
  double PRR[9];                 // column-major, both triangles
  PropSetup prop;
  PropSetup pre[2];              // record-only part of PropSetup (v, w, dt, Q) of steps t, t+1 (by step parity)
  double pre_dphi[2];            // dt * RTV of those steps
  UpdateSetup upd;
  DCand cand[4];                 // one per gating warp
  double nl[2], h3n[2], PLL[4];  // New branch
  double cres, cS;               // compass
  unsigned hs0[KH], hs1[KH];     // sign masks of the pending updates
  int hrank[KH];                 // 2, or 1 for a compass update
};

struct DRunArgs {
  EkfState st;
  EkfRunIO io;
  EkfConst k;
};

__device__ __forceinline__ uint32_t dt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dt_cp_async8(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dt_smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void dt_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void dt_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void dt_cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ double dt_flip(double v, unsigned mask) {
  return __hiloint2double(__double2hiint(v) ^ (int)mask, __double2loint(v));
}
// order-preserving 64-bit key of a double (total order, -0 < +0, NaNs at the ends)
__device__ __forceinline__ unsigned long long dt_key(double v) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  return (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
}
__device__ __forceinline__ unsigned long long dt_warp_min_key(unsigned long long key) {
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
  return ((unsigned long long)mhi << 32) | mlo;
}

// ---- tile sweep: apply the pending downdates to this thread's tile, in order ---------------------
// One copy in the binary (code footprint matters: the SM sub-partitions' instruction caches are the bottleneck of
// the scalar phases). A third of a tile
// (4 rows x 4 columns) at a time in registers, the three column groups in a rolled loop: column
// 12J + 4g + b of a history vector sits at b*HP + 3J + g, so the group only shifts the address. Per
// pending update the eight history words are loaded back to back, then 32 fma on 16 independent elements.
template <int RB>
__device__ __noinline__ void dt_sweep_tile(double* __restrict__ Tt, const DSmem<RB>& sm, int I, int J, int from, int to) {
  constexpr int PS = DCfg<RB>::PS, HP = DCfg<RB>::HP;
#pragma unroll 1
  for (int g = 0; g < 3; ++g) {
    double* Tg = Tt + 4 * g * (4 * PS + 1);            // element (a, 4g + b) at Tg[(a + 4b) * PS + b]
    double t[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int a = 0; a < 4; ++a) t[a][b] = Tg[(a + 4 * b) * PS + b];
#pragma unroll 1
    for (int en = from; en < to; ++en) {
      const int p = en & (KH - 1);
      const double2* W = sm.H[p];
      const unsigned s0 = sm.hs0[p], s1 = sm.hs1[p];
      const bool rank2 = sm.hrank[p] == 2;
      double2 wi[4], wj[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) wi[a] = W[a * HP + I];               // rows 4I..4I+3
#pragma unroll
      for (int b = 0; b < 4; ++b) wj[b] = W[b * HP + 3 * J + g];       // columns 12J + 4g + b
      if (rank2 && (s0 & s1 & 0x80000000u)) {
        // positive definite S (the normal case): u = -W, the negation rides on the fma operand
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int a = 0; a < 4; ++a) t[a][b] = fma(-wi[a].x, wj[b].x, fma(-wi[a].y, wj[b].y, t[a][b]));
      } else {
        double u0[4], u1[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          u0[a] = dt_flip(wi[a].x, s0);
          u1[a] = dt_flip(wi[a].y, s1);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            double v = t[a][b];
            if (rank2) v = fma(u1[a], wj[b].y, v);
            t[a][b] = fma(u0[a], wj[b].x, v);
          }
      }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int a = 0; a < 4; ++a) Tg[(a + 4 * b) * PS + b] = t[a][b];
  }
}

// Record-only scalars of a step (kalmanfilter.cpp:17-37), one lane; evaluated one step ahead.
template <int RB>
__device__ __noinline__ void dt_helper_pre(DSmem<RB>& sm, const double* rec, int slot, const EkfConst& k) {
  PropSetup ps;
  ekf_build_prop_pre(ps, rec[0], rec[1], rec[2], k);
  sm.pre[slot] = ps;
  const double RTV = rec[1] * k.deg2rad_pi / 180.0;
  sm.pre_dphi[slot] = rec[2] * RTV;
}

template <int RB>
__global__ void __launch_bounds__(DT_THREADS, 4) ekf_batch_dtile_kernel(const DRunArgs a) {
  using C = DCfg<RB>;
  constexpr int NQ = C::NQ, HP = C::HP, PS = C::PS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DSmem<RB>& sm = *reinterpret_cast<DSmem<RB>*>(smem_raw);
  double* T = reinterpret_cast<double*>(smem_raw + ((sizeof(DSmem<RB>) + 15) & ~(size_t)15));
  double* recbuf = T + ((C::TSIZE + 1) & ~1);                                                  // [3][Lp]: records of steps t, t+1, t+2
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool helper = warp == 2;                     // scalar chains / P_RR / pose rows
  const bool rows_warp = warp == 1 || warp == 3;     // gain phase: one landmark (two P_LL rows) per thread
  const bool is_tile = tid < C::NT;
  int I = 0, J = 0;
  if (is_tile) {                                     // inverse of dt_slot
    int s = tid;
    while (s >= RB - 1 - 3 * J) { s -= RB - 1 - 3 * J; ++J; }
    I = 3 * J + 1 + s;
  }
  double* Tt = T + tid;                              // this thread's tile
  const int q = tid;                                 // this thread's P_LL row in the row-wise phases (tid < NQ)
  const int hq = dt_hidx<RB>(q < NQ ? q : 0);        // its slot in a history vector
  // gating: two lanes per landmark
  const int gl = tid >> 1, gi = tid & 1;
  // gain phase
  const int lrow = (warp == 1 ? 0 : 32) + lane;      // landmark of this thread
  const int q0 = 2 * lrow < NQ ? 2 * lrow : 0;       // its first P_LL row
  const int rt_l = dt_rowterm<RB>(q0), ct_l = dt_colterm<RB>(q0);
  const int hq0 = dt_hidx<RB>(q0);
  const int ld = a.st.ld, L = a.io.L, T_steps = a.io.T, M = a.io.M;
  const int Lp = (L + 1) & ~1;
  const EkfConst& k = a.k;
  const int cap_lm = a.st.cap_lm < C::NL ? a.st.cap_lm : C::NL;

  // storage of P_LL(r, c), r >= c
  auto pll_ref = [&](int r, int c) -> double* {
    if ((r >> 1) == (c >> 1)) return &sm.Dd[(r & 1) + (c & 1)][r >> 1];
    if ((r >> 2) == (c >> 2)) return &sm.Do[(r & 1) + 2 * (c & 1)][r >> 2];
    return T + dt_addr<RB>(r, c);
  };

  for (int f = blockIdx.x; f < a.st.F; f += gridDim.x) {
    double* gP = a.st.P + (size_t)f * a.st.slab;
    double* gx = a.st.x + (size_t)f * a.st.xs;
    const double* grec = a.io.records + (size_t)f * T_steps * L;
    int n_lm = a.st.nlm[f];
    int cnt = 0, ap = 0;          // history entries created / applied to the tiles
    int dropped = 0;
    // ---- load: HBM (external layout, lower triangle, column by column) -> shared memory ------------
    {
      const int n = 3 + 2 * n_lm;
      for (int i = tid; i < C::TSIZE; i += DT_THREADS) T[i] = 0.0;
      for (int i = tid; i < 3 * NQ; i += DT_THREADS) (&sm.SR[0][0])[i] = 0.0;
      for (int i = tid; i < 3 * C::NL; i += DT_THREADS) (&sm.Dd[0][0])[i] = 0.0;
      for (int i = tid; i < 4 * C::NP; i += DT_THREADS) (&sm.Do[0][0])[i] = 0.0;
      if (tid < NQ) sm.xl[tid] = tid < 2 * n_lm ? gx[3 + tid] : 0.0;
      if (tid < 3) sm.xr[tid] = gx[tid];
      for (int i = tid; i < L; i += DT_THREADS) dt_cp_async8(recbuf + i, grec + i);
      __syncthreads();
      for (int c = warp; c < n; c += DT_THREADS / 32) {
        const double* gc = gP + (size_t)c * ld;
        for (int r = c + lane; r < n; r += 32) {
          const double v = gc[r];
          if (c < 3) {
            if (r < 3) { sm.PRR[r + 3 * c] = v; sm.PRR[c + 3 * r] = v; }
            else sm.SR[c][r - 3] = v;
          } else {
            *pll_ref(r - 3, c - 3) = v;
          }
        }
      }
      dt_cp_async_wait_all();
    }
    __syncthreads();

    // Apply the pending history entries to this thread's tile; CTA-wide variant for the places that
    // need the tiles current (New associations, a full history ring, end of run).
    auto sweep_own = [&]() {
      if (is_tile && 4 * I < 2 * n_lm && ap < cnt) dt_sweep_tile<RB>(Tt, sm, I, J, ap, cnt);
    };
    auto flush_all = [&]() {
      sweep_own();
      ap = cnt;
      __syncthreads();
    };
    // lanes 0..8 of the helper warp: P_RR(i,j) += u_i . W_j of the history entry in `slot`
    auto prr_eager = [&](int slot) {
      if (lane < 9) {
        const int i = lane % 3, j = lane / 3;
        const double2 wi = sm.WR[i], wj = sm.WR[j];
        double v = sm.PRR[i + 3 * j];
        if (sm.hrank[slot] == 2) v = fma(dt_flip(wi.y, sm.hs1[slot]), wj.y, v);
        v = fma(dt_flip(wi.x, sm.hs0[slot]), wj.x, v);
        sm.PRR[i + 3 * j] = v;
      }
      __syncwarp();
    };
    // Warp 0: records are fetched three steps ahead and their record-only scalars evaluated two steps
    // ahead (in its idle time during the gain phase), so the helper chain never waits for either: the
    // CTA barriers of the step in between order them.
    auto fetch_record = [&](int step) {                   // -> ring slot step % 3
      if (step < T_steps) {
        const double* g = grec + (size_t)step * L;
        double* dst = recbuf + (size_t)(step % 3) * Lp;
        for (int i = lane; i < L; i += 32) dt_cp_async8(dst + i, g + i);
      }
      dt_cp_async_commit();
    };
    auto warp0_pre = [&](int step) {
      if (step < T_steps) {
        if (lane == 0) dt_helper_pre<RB>(sm, recbuf + (size_t)(step % 3) * Lp, step & 1, k);
        __syncwarp();
      }
    };
    // ---- helper warp: everything of a step that does not depend on the landmarks -------------------
    // Reads the pose / P_RR left by the previous update; doPropagation's heading-dependent scalars, the
    // propagated pose and P_RR, and the landmark-independent part of the first measurement's update.
    auto helper_chain = [&](const double* rec, int tpar) {
      // operands that do not depend on the headings are loaded first
      const int e = lane % 9, i = e % 3, j = e / 3;
      double PRR[9], Q[4];
#pragma unroll
      for (int w = 0; w < 9; ++w) PRR[w] = sm.PRR[w];
#pragma unroll
      for (int w = 0; w < 4; ++w) Q[w] = sm.pre[tpar].Q[w];
      double sn = 0.0, cs = 1.0;
      PropSetup ps;
      if (lane < 2) {
        const double phi = lane == 0 ? sm.xr[2] : sm.xr[2] + sm.pre_dphi[tpar];   // same expression as the pose update
        if (lane == 0) ps = sm.pre[tpar];
        sincos(phi, &sn, &cs);
        if (lane == 0) {
          ekf_build_prop_trig(ps, sn, cs);
          sm.prop.phi02 = ps.phi02;                                        // all the strip propagation needs (Propagate.cpp:56)
          sm.prop.phi12 = ps.phi12;
          const double xm0 = ps.v * ps.c, xm1 = ps.v * ps.s, xm2 = ps.w;   // Propagate.cpp:33-37
          const double n0 = sm.xr[0] + ps.dt * xm0, n1 = sm.xr[1] + ps.dt * xm1, n2 = sm.xr[2] + ps.dt * xm2;
          sm.xr[0] = n0; sm.xr[1] = n1; sm.xr[2] = n2;
          sm.upd.x0 = n0; sm.upd.x1 = n1;
        } else {
          UpdateTrig tg;
          ekf_build_trig_sc(tg, sn, cs);
          UpdateSetup& u = sm.upd;
          u.c = tg.c; u.s = tg.s;
          for (int w = 0; w < 4; ++w) { u.Ct[w] = tg.Ct[w]; u.mCt[w] = tg.mCt[w]; u.mCtJ[w] = tg.mCtJ[w]; }
          if ((int)rec[5] > 0) {
            u.z0 = rec[8]; u.z1 = rec[9];
            for (int w = 0; w < 4; ++w) u.R[w] = rec[10 + w];
          }
        }
      }
      // Phi_R, G (Propagate.cpp:42-48) from lane 0, the update's rotation from lane 1, by shuffle
      const double phi02 = __shfl_sync(0xffffffffu, ps.phi02, 0), phi12 = __shfl_sync(0xffffffffu, ps.phi12, 0);
      const double g00 = __shfl_sync(0xffffffffu, ps.g00, 0), g10 = __shfl_sync(0xffffffffu, ps.g10, 0);
      const double g21 = __shfl_sync(0xffffffffu, ps.g21, 0);
      const double c1 = __shfl_sync(0xffffffffu, cs, 1), s1 = __shfl_sync(0xffffffffu, sn, 1);
      {
        // 3x3 robot block, one element per lane (Propagate.cpp:53, then :66-67)
        const double Phi[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, phi02, phi12, 1.0};
        const double G[6] = {g00, g10, 0.0, 0.0, 0.0, g21};
        // the select-free form of ekf_prop_prr_elem(Phi, G, Q, PRR, i, j) for a lane-dependent (i, j)
        const double Pi0 = i == 0 ? Phi[0] : (i == 1 ? Phi[1] : Phi[2]);
        const double Pi3 = i == 0 ? Phi[3] : (i == 1 ? Phi[4] : Phi[5]);
        const double Pi6 = i == 0 ? Phi[6] : (i == 1 ? Phi[7] : Phi[8]);
        const double Pj0 = j == 0 ? Phi[0] : (j == 1 ? Phi[1] : Phi[2]);
        const double Pj3 = j == 0 ? Phi[3] : (j == 1 ? Phi[4] : Phi[5]);
        const double Pj6 = j == 0 ? Phi[6] : (j == 1 ? Phi[7] : Phi[8]);
        const double Gi0 = i == 0 ? G[0] : (i == 1 ? G[1] : G[2]), Gi3 = i == 0 ? G[3] : (i == 1 ? G[4] : G[5]);
        const double Gj0 = j == 0 ? G[0] : (j == 1 ? G[1] : G[2]), Gj3 = j == 0 ? G[3] : (j == 1 ? G[4] : G[5]);
        const double t10 = (Pi0 * PRR[0] + Pi3 * PRR[1]) + Pi6 * PRR[2];
        const double t11 = (Pi0 * PRR[3] + Pi3 * PRR[4]) + Pi6 * PRR[5];
        const double t12 = (Pi0 * PRR[6] + Pi3 * PRR[7]) + Pi6 * PRR[8];
        const double t2 = (t10 * Pj0 + t11 * Pj3) + t12 * Pj6;
        const double t30 = Gi0 * Q[0] + Gi3 * Q[1];
        const double t31 = Gi0 * Q[2] + Gi3 * Q[3];
        const double t4 = t30 * Gj0 + t31 * Gj3;
        const double mij = t2 + t4;
        const double mji = __shfl_sync(0xffffffffu, mij, j + 3 * i);
        const double pn = 0.5 * (mij + mji);
        const int qe = e % 6, qi = qe % 2, qj = qe / 2;
        const double p0j = __shfl_sync(0xffffffffu, pn, 0 + 3 * qj);
        const double p1j = __shfl_sync(0xffffffffu, pn, 1 + 3 * qj);
        // mCt = -1.0 * C^T, C^T = {c, -s, s, c} (ekf_build_trig_sc): mCt[qi], mCt[qi + 2]
        const double mct_a = qi == 0 ? -1.0 * c1 : -1.0 * (-s1);
        const double mct_b = qi == 0 ? -1.0 * s1 : -1.0 * c1;
        const double qv = mct_a * p0j + mct_b * p1j;
        if (lane < 9) {
          sm.PRR[e] = pn;
          sm.upd.PRR[e] = pn;
          if (e < 6) sm.upd.q[qe] = qv;
        }
      }
    };

    if (warp == 0) {                                     // records 1, 2 on their way; scalars of steps 0 and 1
      fetch_record(1);
      fetch_record(2);
      warp0_pre(0);
      dt_cp_async_wait_1();                              // record 1 has landed
      __syncwarp();
      warp0_pre(1);
    }
    __syncthreads();
    bool setup_valid = true;
    bool sync_first = false;      // the previous step ended on a New association: its row threads still read P_RR

    for (int t = 0; t < T_steps; ++t) {
      const double* cur = recbuf + (size_t)(t % 3) * Lp;
      // ---- helper warp: landmark-independent part of this step (the one call site of helper_chain).
      // After an Old update the helper warp gets here straight from its pose rows, so this runs beside
      // the gain phase of the previous step on warps 1 and 3.
      if (sync_first) __syncthreads();
      if (helper) {
        if (t > 0 && a.io.pose_trace && lane == 0) {     // slam.cpp:181 for the previous step, before the pose is propagated
          double* pt = a.io.pose_trace + ((size_t)f * T_steps + t - 1) * 3;
          pt[0] = sm.xr[0]; pt[1] = sm.xr[1]; pt[2] = sm.xr[2];
        }
        __syncwarp();
        helper_chain(cur, t & 1);
      }
      sync_first = false;
      setup_valid = true;
      DTILE_TS(0);
      __syncthreads();                                   // B1: helper results of this step (and its record) visible
      DTILE_TS(1);
      const int nz = min((int)cur[5], (L - 8) / 6);   // never read past the record's measurement slots
      // The strip propagation (Propagate.cpp:56-60) rides on the gating lanes' loads of the first measurement
      // unless something else needs the propagated strip first (compass) or nothing gates (no measurement).
      bool fold_prop = nz > 0 && cur[6] == 0.0;
      if (!fold_prop) {
        if (q < 2 * n_lm) {
          double a0 = sm.SR[0][q], a1 = sm.SR[1][q], a2 = sm.SR[2][q];
          ekf_prop_col(sm.prop, a0, a1, a2);
          sm.SR[0][q] = a0; sm.SR[1][q] = a1; sm.SR[2][q] = a2;
        }
        if (nz > 0) __syncthreads();
      }
      // ---- doUpdateCompass (slam.cpp:144-147, kalmanfilter.cpp:96-130): a rank-1 pending update -----
      if (cur[6] != 0.0) {
        if (cnt - ap > KH - 1) flush_all();              // no free history slot
        const int sl = cnt & (KH - 1);
        if (tid == 0) {
          sm.cres = ekf_compass_residual(sm.xr[2], cur[3], k);
          sm.cS = sm.PRR[8] + cur[4];
          sm.hs0[sl] = sm.cS < 0 ? 0u : 0x80000000u;
          sm.hs1[sl] = 0u;
          sm.hrank[sl] = 1;
        }
        __syncthreads();
        {
          const double res = sm.cres, S = sm.cS, invS = 1 / S, sq = sqrt(fabs(S));
          if (tid < NQ) {
            double2 w = make_double2(0.0, 0.0);
            if (q < 2 * n_lm) {
              const double Ki = invS * sm.SR[2][q];
              sm.xl[q] = sm.xl[q] + res * Ki;
              w = make_double2(sq * Ki, 0.0);
            }
            sm.H[sl][hq] = w;
          } else if (tid < NQ + 3) {
            const int r = tid - NQ;
            const double Ki = invS * sm.PRR[r + 6];
            sm.xr[r] = sm.xr[r] + res * Ki;
            sm.WR[r] = make_double2(sq * Ki, 0.0);
          }
        }
        __syncthreads();
        if (q < 2 * n_lm) {                              // eager rank-1 downdate of strip and diagonal blocks
          const unsigned s0 = sm.hs0[sl];
          const double2* Hc = sm.H[sl];
          const double u0 = dt_flip(Hc[hq].x, s0);
#pragma unroll
          for (int j = 0; j < 3; ++j) sm.SR[j][q] = fma(u0, sm.WR[j].x, sm.SR[j][q]);
          const int aa = q & 3, l = q >> 1, m = q >> 2;
          if ((aa & 1) == 0) sm.Dd[0][l] = fma(u0, Hc[hq].x, sm.Dd[0][l]);
          else {
            sm.Dd[1][l] = fma(u0, Hc[hq - HP].x, sm.Dd[1][l]);
            sm.Dd[2][l] = fma(u0, Hc[hq].x, sm.Dd[2][l]);
          }
          if (aa >= 2) {
            sm.Do[(aa & 1)][m] = fma(u0, Hc[m].x, sm.Do[(aa & 1)][m]);
            sm.Do[(aa & 1) + 2][m] = fma(u0, Hc[HP + m].x, sm.Do[(aa & 1) + 2][m]);
          }
        }
        if (helper) prr_eager(sl);
        cnt += 1;
        setup_valid = false;
        __syncthreads();
      }

      // ---- doUpdate per measurement (slam.cpp:150-171, Update.cpp:80-195) -------------------------
      for (int m = 0; m < M; ++m) {
        if (m >= nz) {
          if (tid == 0) {
            const size_t oi = ((size_t)f * T_steps + t) * M + m;
            if (a.io.decision) a.io.decision[oi] = EKF_DEC_NONE;
            if (a.io.index) a.io.index[oi] = -1;
            if (a.io.mahal) a.io.mahal[oi] = 0.0;
          }
          continue;
        }
        const double* zr = cur + 8 + 6 * m;
        if (!setup_valid) {                                // pose / P_RR changed since the helper built sm.upd
          __syncthreads();                                 // the previous update (or the compass) is complete
          if (tid == 0) {
            UpdateSetup u;
            double PRR[9];
            for (int w = 0; w < 9; ++w) PRR[w] = sm.PRR[w];
            ekf_build_setup(u, sm.xr[2], sm.xr[0], sm.xr[1], PRR, zr[0], zr[1], zr + 2);
            sm.upd = u;
          }
          __syncthreads();
        }
        // ---- every third update: bring the tiles up to date (nothing reads them before the gain phase)
        if (cnt - ap >= KH - 1) {
          sweep_own();
          ap = cnt;
        }
        DTILE_TS(2);
        // ================= gating: two lanes per landmark, reference operation order ===================
        {
          const UpdateSetup& u = sm.upd;
          const bool have = gl < n_lm;
          double val = INFINITY;
          int my_idx = INT_MAX;
          GateResult g;
          GatePre pre;
          double ll = 0.0, sq0 = 0.0, sq1 = 0.0, d0 = 0.0, d1 = 0.0;
          double Sa = 0.0, Sb = 0.0;
          double pxa = 0.0, pxb = 0.0;
          double2 pwa = make_double2(0.0, 0.0), pwb = pwa;
          double2 s0 = make_double2(0.0, 0.0), s1 = s0, s2 = s0;
          if (have) {
            const double2 xy = *reinterpret_cast<const double2*>(&sm.xl[2 * gl]);
            s0 = *reinterpret_cast<const double2*>(&sm.SR[0][2 * gl]);
            s1 = *reinterpret_cast<const double2*>(&sm.SR[1][2 * gl]);
            s2 = *reinterpret_cast<const double2*>(&sm.SR[2][2 * gl]);
            const double p10 = sm.Dd[1][gl];
            const double pll[4] = {sm.Dd[0][gl], p10, p10, sm.Dd[2][gl]};
            if (fold_prop) {                               // P_RL <- Phi*P_RL for this landmark's two rows (Propagate.cpp:56-60)
              ekf_prop_col(sm.prop, s0.x, s1.x, s2.x);
              ekf_prop_col(sm.prop, s0.y, s1.y, s2.y);
            }
            const double p[6] = {s0.x, s0.y, s1.x, s1.y, s2.x, s2.y};
            ekf_gate_prelude(u, xy.x, xy.y, pre);
            // lane gi owns row gi of S: elements k = gi (column 0) and k = gi + 2 (column 1)
            Sa = ekf_gate_S_element(u, pre, p, pll, gi);
            Sb = ekf_gate_S_element(u, pre, p, pll, gi + 2);
          }
          // the pair exchanges its rows of S (every lane takes part; the loads above are complete here)
          const double Oa = __shfl_xor_sync(0xffffffffu, Sa, 1), Ob = __shfl_xor_sync(0xffffffffu, Sb, 1);
          if (have) {
            if (fold_prop) {                               // lane gi stores the propagated strip row gi
              sm.SR[0][2 * gl + gi] = gi ? s0.y : s0.x;
              sm.SR[1][2 * gl + gi] = gi ? s1.y : s1.x;
              sm.SR[2][2 * gl + gi] = gi ? s2.y : s2.x;
            }
            const double Sraw[4] = {gi ? Oa : Sa, gi ? Sa : Oa, gi ? Ob : Sb, gi ? Sb : Ob};
            ekf_gate_from_S(pre, Sraw, k.cond_max, g);
            // L D L^T of S (what the downdate needs if this landmark wins), beside the gate's own latencies
            d0 = g.S[0]; ll = g.S[1] / g.S[0]; d1 = g.S[3] - ll * g.S[1];
            sq0 = sqrt(fabs(d0)); sq1 = sqrt(fabs(d1));
            const bool valid = !g.skip && (k.mahal_init > g.d2);   // Update.cpp:131,140
            if (valid) { val = g.d2; my_idx = 3 + 2 * gl; }
            // Pose rows of the gain, pose correction and their downdate vectors (Update.cpp:186-187) as if this
            // landmark won: lane 0 of the pair takes rows 0 and 1, lane 1 row 2 (the heading the next step's
            // scalar chain starts from), so nothing serial follows the argmin.
            {
              const double h00 = u.mCt[0], h01 = u.mCt[2], h02 = g.h3_0;
              const double h10 = u.mCt[1], h11 = u.mCt[3], h12 = g.h3_1;
              const double c00 = u.Ct[0], c10 = u.Ct[2], c01 = u.Ct[1], c11 = u.Ct[3];
#pragma unroll
              for (int rr = 0; rr < 2; ++rr) {
                if (rr == 1 && gi == 1) break;
                const int r = gi ? 2 : rr;
                const double p0 = u.PRR[r], p1 = u.PRR[r + 3], p2 = u.PRR[r + 6];
                const double pa = r == 0 ? s0.x : (r == 1 ? s1.x : s2.x);   // P(r, c0)   = SR[r][2 gl]
                const double pb = r == 0 ? s0.y : (r == 1 ? s1.y : s2.y);   // P(r, c0+1) = SR[r][2 gl + 1]
                const double A0 = (p0 * h00 + p1 * h01) + p2 * h02;
                const double A1 = (p0 * h10 + p1 * h11) + p2 * h12;
                const double B0 = pa * c00 + pb * c10;
                const double B1 = pa * c01 + pb * c11;
                const double M0 = A0 + B0, M1 = A1 + B1;
                const double K0 = M0 * g.Si[0] + M1 * g.Si[1];
                const double K1 = M0 * g.Si[2] + M1 * g.Si[3];
                const double xn = sm.xr[r] + (K0 * g.res0 + K1 * g.res1);
                const double2 wr = make_double2(sq0 * fma(ll, K1, K0), sq1 * K1);
                if (rr == 0) { pxa = xn; pwa = wr; } else { pxb = xn; pwb = wr; }
              }
            }
          }
          int idx;
          {   // warp argmin, lowest index wins ties (Update.cpp:140)
            const unsigned long long key = dt_key(val + 0.0);
            const unsigned long long mk = dt_warp_min_key(key);
            idx = (int)__reduce_min_sync(0xffffffffu, key == mk ? (unsigned)my_idx : (unsigned)INT_MAX);
          }
          DCand& cd = sm.cand[warp];
          if (idx == INT_MAX) {
            if (lane == 0) { cd.val = INFINITY; cd.idx = INT_MAX; }
          } else if (my_idx == idx && gi == 1) {
            cd.xr[2] = pxa; cd.WR[2] = pwa;
          } else if (my_idx == idx) {
            cd.xr[0] = pxa; cd.WR[0] = pwa; cd.xr[1] = pxb; cd.WR[1] = pwb;
            const int c0 = 2 * gl;
            cd.val = val; cd.idx = idx; cd.c0 = c0;
            cd.rt_c0 = dt_rowterm<RB>(c0); cd.ct_c0 = dt_colterm<RB>(c0);
            cd.res[0] = g.res0; cd.res[1] = g.res1;
            cd.S[0] = g.S[0]; cd.S[1] = g.S[1]; cd.S[2] = g.S[2]; cd.S[3] = g.S[3];
            cd.Si[0] = g.Si[0]; cd.Si[1] = g.Si[1]; cd.Si[2] = g.Si[2]; cd.Si[3] = g.Si[3];
            cd.h3[0] = g.h3_0; cd.h3[1] = g.h3_1;
            cd.l = ll; cd.sq0 = sq0; cd.sq1 = sq1;
            cd.sm0 = d0 < 0 ? 0u : 0x80000000u; cd.sm1 = d1 < 0 ? 0u : 0x80000000u;
#pragma unroll
            for (int w = 0; w < 4; ++w) { cd.Ct[w] = u.Ct[w]; cd.mCt[w] = u.mCt[w]; }
          }
        }
        fold_prop = false;
        DTILE_TS(3);
        __syncthreads();                                 // B2: candidates published
        DTILE_TS(4);
        // ---- decision (Update.cpp:152,181,191), uniform over the CTA ----------------------------------
        int wsel = 0;
        int decision, index;
        double mahal;
        {
          double val = sm.cand[0].val;
          int idx = sm.cand[0].idx;
#pragma unroll
          for (int w = 1; w < 4; ++w) {
            const double ov = sm.cand[w].val;
            const int oi = sm.cand[w].idx;
            if (ov < val || (ov == val && oi < idx)) { val = ov; idx = oi; wsel = w; }
          }
          const int opt_i = (idx == INT_MAX) ? 0 : idx;
          mahal = (idx == INT_MAX) ? k.mahal_init : val;
          decision = ekf_decide(opt_i, mahal, k);
          if (decision == EKF_DEC_NEW && n_lm >= cap_lm) decision = EKF_DEC_DROPPED;
          index = opt_i;
          if (decision == EKF_DEC_NEW) index = 3 + 2 * n_lm;
          if (decision == EKF_DEC_DROPPED) index = -1;
        }
        const DCand& dc = sm.cand[wsel];
        if (tid == 0) {
          const size_t oi = ((size_t)f * T_steps + t) * M + m;
          if (a.io.decision) a.io.decision[oi] = decision;
          if (a.io.index) a.io.index[oi] = index;
          if (a.io.mahal) a.io.mahal[oi] = mahal;
        }
        setup_valid = false;
        if (decision == EKF_DEC_OLD) {
          if (cnt - ap > KH - 1) flush_all();              // no free history slot (compass / several measurements per step)
          const int sl = cnt & (KH - 1);
          if (helper) {
            // ---- the winner's pose, the history entry's signs, P_RR downdate; then straight on to the next
            // step's chain (top of the loop)
            if (lane < 3) { sm.xr[lane] = dc.xr[lane]; sm.WR[lane] = dc.WR[lane]; }
            if (lane == 0) { sm.hs0[sl] = dc.sm0; sm.hs1[sl] = dc.sm1; sm.hrank[sl] = 2; }
            __syncwarp();
            prr_eager(sl);
          }
          if (rows_warp) {
            // ---- gain rows, state, downdate vectors (Update.cpp:186-187): one landmark (two rows) per thread
            const bool live = lrow < n_lm;
            double2 wA = make_double2(0.0, 0.0), wB = make_double2(0.0, 0.0);
            double uA0 = 0.0, uA1 = 0.0, uB0 = 0.0, uB1 = 0.0;
            double2 s0 = make_double2(0.0, 0.0), s1 = s0, s2 = s0;
            if (live) {
              const int c0 = dc.c0, hc0 = dt_hidx<RB>(c0);
              s0 = *reinterpret_cast<const double2*>(&sm.SR[0][q0]);
              s1 = *reinterpret_cast<const double2*>(&sm.SR[1][q0]);
              s2 = *reinterpret_cast<const double2*>(&sm.SR[2][q0]);
              double paA, pbA, paB, pbB;                 // P(q0, c0), P(q0, c0+1), P(q0+1, c0), P(q0+1, c0+1)
              if ((q0 >> 2) == (c0 >> 2)) {              // same aligned pair: eager entries, already current
                if (q0 == c0) {
                  const double d1v = sm.Dd[1][lrow];
                  paA = sm.Dd[0][lrow]; pbA = d1v; paB = d1v; pbB = sm.Dd[2][lrow];
                } else {
                  const int m2 = q0 >> 2;
                  const double e0 = sm.Do[0][m2], e1 = sm.Do[1][m2], e2 = sm.Do[2][m2], e3 = sm.Do[3][m2];
                  if (c0 & 2) { paA = e0; pbA = e1; paB = e2; pbB = e3; }   // this landmark is the pair's first: P(4m+j, 4m+2+i) = Do[i+2j]
                  else { paA = e0; pbA = e2; paB = e1; pbB = e3; }          // this landmark is the second: P(4m+2+i, 4m+j) = Do[i+2j]
                }
              } else {
                // stored tile entries + the pending downdates, in order (the deferred sweep's fma sequence)
                if (q0 > c0) {
                  const int ad = rt_l + dc.ct_c0;        // (q0, c0)
                  paA = T[ad]; pbA = T[ad + 4 * PS + 1]; paB = T[ad + PS]; pbB = T[ad + 5 * PS + 1];
                } else {
                  const int ad = dc.rt_c0 + ct_l;        // (c0, q0)
                  paA = T[ad]; pbA = T[ad + PS]; paB = T[ad + 4 * PS + 1]; pbB = T[ad + 5 * PS + 1];
                }
                for (int en = ap; en < cnt; ++en) {
                  const int p = en & (KH - 1);
                  const double2 wqa = sm.H[p][hq0], wqb = sm.H[p][hq0 + HP], wa = sm.H[p][hc0], wb = sm.H[p][hc0 + HP];
                  const unsigned m0 = sm.hs0[p], m1 = sm.hs1[p];
                  const double va0 = dt_flip(wqa.x, m0), va1 = dt_flip(wqa.y, m1);
                  const double vb0 = dt_flip(wqb.x, m0), vb1 = dt_flip(wqb.y, m1);
                  if (sm.hrank[p] == 2) {
                    paA = fma(va1, wa.y, paA); pbA = fma(va1, wb.y, pbA);
                    paB = fma(vb1, wa.y, paB); pbB = fma(vb1, wb.y, pbB);
                  }
                  paA = fma(va0, wa.x, paA); pbA = fma(va0, wb.x, pbA);
                  paB = fma(vb0, wa.x, paB); pbB = fma(vb0, wb.x, pbB);
                }
              }
              const double h00 = dc.mCt[0], h01 = dc.mCt[2], h02 = dc.h3[0];
              const double h10 = dc.mCt[1], h11 = dc.mCt[3], h12 = dc.h3[1];
              const double c00 = dc.Ct[0], c10 = dc.Ct[2], c01 = dc.Ct[1], c11 = dc.Ct[3];
              {   // row q0
                const double A0 = (s0.x * h00 + s1.x * h01) + s2.x * h02;
                const double A1 = (s0.x * h10 + s1.x * h11) + s2.x * h12;
                const double B0 = paA * c00 + pbA * c10;
                const double B1 = paA * c01 + pbA * c11;
                const double M0 = A0 + B0, M1 = A1 + B1;
                const double K0 = M0 * dc.Si[0] + M1 * dc.Si[1];
                const double K1 = M0 * dc.Si[2] + M1 * dc.Si[3];
                sm.xl[q0] = sm.xl[q0] + (K0 * dc.res[0] + K1 * dc.res[1]);
                wA = make_double2(dc.sq0 * fma(dc.l, K1, K0), dc.sq1 * K1);
                uA0 = dt_flip(wA.x, dc.sm0); uA1 = dt_flip(wA.y, dc.sm1);
              }
              {   // row q0 + 1
                const double A0 = (s0.y * h00 + s1.y * h01) + s2.y * h02;
                const double A1 = (s0.y * h10 + s1.y * h11) + s2.y * h12;
                const double B0 = paB * c00 + pbB * c10;
                const double B1 = paB * c01 + pbB * c11;
                const double M0 = A0 + B0, M1 = A1 + B1;
                const double K0 = M0 * dc.Si[0] + M1 * dc.Si[1];
                const double K1 = M0 * dc.Si[2] + M1 * dc.Si[3];
                sm.xl[q0 + 1] = sm.xl[q0 + 1] + (K0 * dc.res[0] + K1 * dc.res[1]);
                wB = make_double2(dc.sq0 * fma(dc.l, K1, K0), dc.sq1 * K1);
                uB0 = dt_flip(wB.x, dc.sm0); uB1 = dt_flip(wB.y, dc.sm1);
              }
              // eager downdate of the landmark's own 2x2 block (Update.cpp:188,193-194)
              sm.Dd[0][lrow] = fma(uA0, wA.x, fma(uA1, wA.y, sm.Dd[0][lrow]));
              sm.Dd[1][lrow] = fma(uB0, wA.x, fma(uB1, wA.y, sm.Dd[1][lrow]));
              sm.Dd[2][lrow] = fma(uB0, wB.x, fma(uB1, wB.y, sm.Dd[2][lrow]));
            }
            if (lrow < C::NL) {                          // zeros for landmarks beyond the live map
              sm.H[sl][hq0] = wA;
              sm.H[sl][hq0 + HP] = wB;
            }
            {   // 2x2 block between the two landmarks of an aligned pair: rows of the odd one, columns of the even one
              const double pAx = __shfl_up_sync(0xffffffffu, wA.x, 1), pAy = __shfl_up_sync(0xffffffffu, wA.y, 1);
              const double pBx = __shfl_up_sync(0xffffffffu, wB.x, 1), pBy = __shfl_up_sync(0xffffffffu, wB.y, 1);
              if (live && (lrow & 1)) {
                const int m2 = lrow >> 1;
                sm.Do[0][m2] = fma(uA0, pAx, fma(uA1, pAy, sm.Do[0][m2]));   // (4m+2, 4m)
                sm.Do[1][m2] = fma(uB0, pAx, fma(uB1, pAy, sm.Do[1][m2]));   // (4m+3, 4m)
                sm.Do[2][m2] = fma(uA0, pBx, fma(uA1, pBy, sm.Do[2][m2]));   // (4m+2, 4m+1)
                sm.Do[3][m2] = fma(uB0, pBx, fma(uB1, pBy, sm.Do[3][m2]));   // (4m+3, 4m+1)
              }
            }
            if (live) {                                  // eager downdate of the strip rows
              const double2 w0 = dc.WR[0], w1 = dc.WR[1], w2 = dc.WR[2];
              *reinterpret_cast<double2*>(&sm.SR[0][q0]) = make_double2(fma(uA0, w0.x, fma(uA1, w0.y, s0.x)), fma(uB0, w0.x, fma(uB1, w0.y, s0.y)));
              *reinterpret_cast<double2*>(&sm.SR[1][q0]) = make_double2(fma(uA0, w1.x, fma(uA1, w1.y, s1.x)), fma(uB0, w1.x, fma(uB1, w1.y, s1.y)));
              *reinterpret_cast<double2*>(&sm.SR[2][q0]) = make_double2(fma(uA0, w2.x, fma(uA1, w2.y, s2.x)), fma(uB0, w2.x, fma(uB1, w2.y, s2.y)));
            }
          }
          DTILE_TS(6);
          cnt += 1;
        } else if (decision == EKF_DEC_NEW) {
          // ---- state augmentation (Update.cpp:152-178); pending downdates are applied first ----------
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
            for (int w2 = 0; w2 < 4; ++w2) in[w2] = t1[w2] + u.R[w2];
            const double Cm[4] = {u.Ct[0], u.Ct[2], u.Ct[1], u.Ct[3]};
            for (int j = 0; j < 2; ++j)
              for (int i = 0; i < 2; ++i) b1[i + 2 * j] = Cm[i] * in[0 + 2 * j] + Cm[i + 2] * in[1 + 2 * j];
            for (int j = 0; j < 2; ++j)       // Update.cpp:168
              for (int i = 0; i < 2; ++i)
                sm.PLL[i + 2 * j] = b1[i] * u.Ct[0 + 2 * j] + b1[i + 2] * u.Ct[1 + 2 * j];
            sm.nl[0] = nl0; sm.nl[1] = nl1;
            sm.h3n[0] = h30; sm.h3n[1] = h31;
          }
          flush_all();                                   // tiles current; the New blocks visible
          const UpdateSetup& u = sm.upd;
          const int qn = 2 * n_lm;                       // P_LL rows of the new landmark
          const bool pose_row = tid >= NQ && tid < NQ + 3;
          if (q < qn || pose_row) {                      // P_xL = -P[:,0:3]*H_R^T*H_Li (:169), existing row
            const double h00 = u.mCt[0], h01 = u.mCt[2], h02 = sm.h3n[0];
            const double h10 = u.mCt[1], h11 = u.mCt[3], h12 = sm.h3n[1];
            const double ct00 = u.Ct[0], ct10 = u.Ct[1], ct01 = u.Ct[2], ct11 = u.Ct[3];
            double e0, e1, e2;
            if (pose_row) { const int r = tid - NQ; e0 = sm.PRR[r]; e1 = sm.PRR[r + 3]; e2 = sm.PRR[r + 6]; }
            else { e0 = sm.SR[0][q]; e1 = sm.SR[1][q]; e2 = sm.SR[2][q]; }
            const double q0v = -e0, q1v = -e1, q2v = -e2;
            const double t0 = (q0v * h00 + q1v * h01) + q2v * h02;
            const double t1 = (q0v * h10 + q1v * h11) + q2v * h12;
            const double v0 = t0 * ct00 + t1 * ct10, v1 = t0 * ct01 + t1 * ct11;
            if (pose_row) { const int r = tid - NQ; sm.SR[r][qn] = v0; sm.SR[r][qn + 1] = v1; }
            else { *pll_ref(qn, q) = v0; *pll_ref(qn + 1, q) = v1; }
          }
          if (tid == 0) {
            const double off = 0.5 * (sm.PLL[2] + sm.PLL[1]);   // :193-194 on the new 2x2 block
            sm.Dd[0][n_lm] = sm.PLL[0];
            sm.Dd[1][n_lm] = off;
            sm.Dd[2][n_lm] = sm.PLL[3];
            sm.xl[qn] = sm.nl[0];
            sm.xl[qn + 1] = sm.nl[1];
          }
          n_lm += 1;
          sync_first = true;
        } else if (decision == EKF_DEC_DROPPED) {
          dropped = 1;
        }
      }
      // ---- warp 0 (idle during the gain phase): record-only scalars of the next step -----------------
      if (warp == 0) {
        fetch_record(t + 3);                             // into the slot of this step's record (not read any more)
        dt_cp_async_wait_1();                            // record t+2 has landed
        __syncwarp();
        warp0_pre(t + 2);
      }
      DTILE_TS(8);
    }

    // ---- flush the pending downdates, write back to HBM (external layout, both triangles) ---------
    __syncthreads();
    if (a.io.pose_trace && tid == 0 && T_steps > 0) {    // slam.cpp:181 for the last step
      double* pt = a.io.pose_trace + ((size_t)f * T_steps + T_steps - 1) * 3;
      pt[0] = sm.xr[0]; pt[1] = sm.xr[1]; pt[2] = sm.xr[2];
    }
    flush_all();
    {
      const int n = 3 + 2 * n_lm;
      for (int c = warp; c < n; c += DT_THREADS / 32) {
        double* gc = gP + (size_t)c * ld;
        for (int r = lane; r < n; r += 32) {
          const int hi = r > c ? r : c, lo = r > c ? c : r;
          double v;
          if (hi < 3) v = sm.PRR[hi + 3 * lo];
          else if (lo < 3) v = sm.SR[lo][hi - 3];
          else v = *pll_ref(hi - 3, lo - 3);
          gc[r] = v;
        }
      }
      if (tid < 2 * n_lm) gx[3 + tid] = sm.xl[tid];
      if (tid < 3) gx[tid] = sm.xr[tid];
      if (tid == 0) {
        a.st.nlm[f] = n_lm;
        if (dropped) a.st.status[f] |= 1;
      }
    }
    __syncthreads();
  }
}

template <int RB>
size_t dtile_smem_bytes(int L) {
  return ((sizeof(DSmem<RB>) + 15) & ~(size_t)15) + (size_t)((DCfg<RB>::TSIZE + 1) & ~1) * sizeof(double) +
         (size_t)3 * ((L + 1) & ~1) * sizeof(double);
}

template <int RB>
cudaError_t launch_dtile(const DRunArgs& a, int sm_count, cudaStream_t stream) {
  const size_t bytes = dtile_smem_bytes<RB>(a.io.L);
  static size_t configured_dev[64] = {0};      // function attributes are per device
  static int grid_cap_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  size_t& configured = configured_dev[dev];
  int& grid_cap = grid_cap_dev[dev];
  if (bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(ekf_batch_dtile_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ekf_batch_dtile_kernel<RB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ekf_batch_dtile_kernel<RB>, DT_THREADS, bytes);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    grid_cap = per_sm * sm_count;
    configured = bytes;
  }
  const int grid = a.st.F < grid_cap ? a.st.F : grid_cap;
  ekf_batch_dtile_kernel<RB><<<grid, DT_THREADS, bytes, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace

cudaError_t ekf_dtile_timestamps(long long* out64) {
#ifdef EKF_DTILE_TIMING
  return cudaMemcpyFromSymbol(out64, g_dtile_ts, sizeof(long long) * 64);
#else
  for (int i = 0; i < 64; ++i) out64[i] = 0;
  return cudaSuccess;
#endif
}

int ekf_dtile_max_landmarks() { return DCfg<25>::NL; }
int ekf_dtile_ctas_per_sm() { return 4; }

cudaError_t ekf_dtile_run(const EkfState& st, const EkfRunIO& io, const EkfConst& k, int sm_count, cudaStream_t stream) {
  DRunArgs a{st, io, k};
  if (st.cap_lm <= DCfg<25>::NL) return launch_dtile<25>(a, sm_count, stream);
  return cudaErrorInvalidValue;
}
