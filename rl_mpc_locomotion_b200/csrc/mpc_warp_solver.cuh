// mpc_warp_solver.cuh — one warp owns one robot's stance-force QP end to end.
//
// What it computes (reference file:line it replaces):
//   * the condensed single-rigid-body QP of ConvexMpc::ComputeContactForces
//     (reference/MPC_Controller/convex_MPC/mpc_osqp.cc:593-692) — built on chip from the
//     96-float problem record, never materialised (see "QP structure" below);
//   * OSQP 0.6.0 exactly as the reference drives it (mpc_osqp.cc:705-712, :757-796 and
//     reference/extern/osqp/src/{osqp.c:288-630, auxil.c, scaling.c, polish.c}): Ruiz
//     equilibration (10 passes), rho vector with 1e3*rho on equality rows, ADMM with
//     alpha=1.6 / sigma=1e-6, residual check + rho adaptation every 25 iterations,
//     eps_abs=eps_rel=1e-3, polish (delta=1e-6, 3 refinement steps), unscale, negate.
//   The reference's loosely converged output depends on that whole trajectory (the true
//   minimiser is up to ~100 N away on the benchmark distribution), so reproducing the
//   *algorithm* in fp64 is the only way to return what the reference returns.
//
// QP structure: with E = [I_w^-1 [r_i]x ; I/m] (6x12, "wrench map"),
//   P = alpha*I + 2 (I_h (x) E)^T (Gamma (x) Q1 + N (x) Q2) (I_h (x) E),
//   Gamma_ij = sum_{k>=max(i,j)} c_{k-i} c_{k-j}, c_m = dt^2 (m + 1/2), N_ij = h - max(i,j),
//   Q1 = blkdiag(T^T W_rpy T, W_pos), Q2 = dt^2 blkdiag(W_omega, W_vel),
// and the friction constraints are 5x3 blocks per (step, leg).  Hence the ADMM matrix
//   K = Pbar + sigma*I + Abar^T diag(rho) Abar = Delta + U^T Q U
// is "3x3-block-diagonal + rank 6h", and K^-1 is applied through the 6h x 6h SPD matrix
//   Hm = blkdiag(F_k^-1) + Q,  F_k = U_k Delta_k^-1 U_k^T            (Woodbury)
// whose Cholesky factor is inverted once per factorisation so that every ADMM iteration
// is two conflict-free triangular mat-vecs out of shared memory (no dependent chain).
//
// Written in the warp-program dialect of warp_prog.h: compiles for sm_100a (product) and,
// with -DMPC_EMULATE, as a sequential lane-by-lane emulation for GPU-less CI.
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/mpc_b200.h"
#include "warp_prog.h"

// Development instrumentation (-DMPC_STAGE_TIMING): cycles between the hand-off points of one ADMM
// iteration, accumulated by lane 0 of every CTA into a global table (see tools/stage_timing.py).
#if defined(MPC_STAGE_TIMING) && defined(__CUDACC__) && !defined(MPC_EMULATE)
__device__ unsigned long long g_stage_cycles[16];
#define MPC_T0() long long t_stage_ = clock64()
#define MPC_TICK(slot) do { const long long t_now_ = clock64(); if (lane == 0) atomicAdd(&g_stage_cycles[slot], (unsigned long long)(t_now_ - t_stage_)); t_stage_ = t_now_; } while (0)
#else
#define MPC_T0() ((void)0)
#define MPC_TICK(slot) ((void)0)
#endif

namespace mpc {

constexpr int NL = MPC_NL;  // threads per robot (1 or 2 warps), see warp_prog.h

struct SolveParams {
  int max_iter;         // osqp MAX_ITER 4000
  int check_every;      // osqp CHECK_TERMINATION 25 (== adaptive_rho_interval, mpc_osqp.cc:710)
  int scaling_iters;    // osqp SCALING 10
  int polish;           // mpc_osqp.cc:709
  int polish_refine;    // osqp POLISH_REFINE_ITER 3
  int adaptive_rho;     // osqp ADAPTIVE_RHO 1
  int fixed_iters;      // >0: run exactly this many iterations, no termination test
  int reserved;
  double dt, alpha;     // mpc_osqp.cc:512-513
  double rho0, sigma, relax, eps_abs, eps_rel, delta, rho_tol_ratio;
};

// per-robot diagnostic record (doubles) written when d_info != nullptr
enum { INFO_ITERS = 0, INFO_RHO_UPDATES, INFO_RHO, INFO_POLISH, INFO_PRI, INFO_DUA,
       INFO_C, INFO_NFREE, INFO_STRIDE = 8 };

constexpr double kGravity = 9.8;    // mpc_osqp.cc:54
constexpr double kMaxScale = 10.0;  // mpc_osqp.cc:55
constexpr double kMinScale = 0.1;   // mpc_osqp.cc:56
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoEqFactor = 1e3, kRhoTol = 1e-4;
constexpr double kMinScaling = 1e-4, kMaxScaling = 1e4;

template <int H, int PCAP_ = 6 * H>
struct Dims {
  static constexpr int NL = MPC_NL;              // lanes (threads) cooperating on one robot
  static constexpr int NV = 12 * H;             // variables
  static constexpr int NLS = 4 * H;             // (step, leg) pairs
  static constexpr int NSLOT = (NLS + NL - 1) / NL; // leg-steps per lane
  static constexpr int NW = 6 * H;              // wrench-space dimension
  static constexpr int NWS = (NW + NL - 1) / NL;
  static constexpr int NVS = (NV + NL - 1) / NL;
  // polish capacity: free coordinates the on-chip polish can hold.  6H covers every contact table
  // with <= 2H stance leg-steps (all of the reference's gaits); the overflow kernel uses 12H.
  static constexpr int PCAP = PCAP_ < NW ? NW : PCAP_;
  static constexpr int NTRI = PCAP * (PCAP + 1) / 2;  // packed lower triangle (row i at i(i+1)/2; slot i of row i unused)
  // shared-memory layout (doubles)
  static constexpr int OFF_WP = 0;                       // [NTRI] Hm -> L -> W = L^-1, row-packed
  static constexpr int OFF_SCR = OFF_WP + NTRI;          // scratch: vecA vecB vt vg vw Finv
  static constexpr int OFF_VECA = OFF_SCR;               // [NV]
  static constexpr int OFF_VECB = OFF_VECA + NV;         // [NV]
  static constexpr int OFF_VT = OFF_VECB + NV;           // [NW]
  static constexpr int OFF_VG = OFF_VT + NW;             // [NW]
  static constexpr int OFF_VW = OFF_VG + NW;             // [NW]
  static constexpr int OFF_FINV = OFF_VW + NW;           // [H][36]
  static constexpr int SCR_MIN = 2 * NV + 3 * NW + 36 * H;
  static constexpr int SCR_POLISH = 12 * PCAP + PCAP / 2 + 8;   // Mf, A1f, kf of the polish
  static constexpr int SCR_SIZE = SCR_MIN > SCR_POLISH ? SCR_MIN : ((SCR_POLISH + 1) & ~1);
  static constexpr int OFF_DINV = OFF_SCR + SCR_SIZE;    // [PCAP] diag: H_ii -> 1/L_ii = W_ii
  static constexpr int OFF_COL = OFF_DINV + PCAP;        // [2][PCAP] column staging (inversion, polish)
  static constexpr int OFF_GAM = OFF_COL + 2 * PCAP;     // [H][H]
  static constexpr int OFF_NN = OFF_GAM + H * H;         // [H][H]
  static constexpr int WBS = 13;                         // row stride of Wb: odd, so the six rows read
                                                         // by the lanes of a step sit in distinct banks
  static constexpr int OFF_WB = OFF_NN + H * H;          // [6][WBS] balanced wrench map (12 used per row)
  static constexpr int OFF_Q1S = OFF_WB + 6 * WBS + (6 * WBS & 1);  // [6][6] Q1'
  static constexpr int OFF_Q2S = OFF_Q1S + 36;           // [6] diag of Q2' (+2 pad)
  static constexpr int OFF_DSM = OFF_Q2S + 8;            // [max(NV, 2 PCAP)] D (scaling); polish vectors
  // G_k = F_k^-1 U_k ([NW][GS], GS odd: conflict-free rows) lets ksolve() go from s to g = F^-1 U s
  // in one stage.  Only where the extra 6.2 KB do not cost a resident CTA (h = 10: still 6 per SM).
  static constexpr bool FUSE_G = (H <= 10);
  static constexpr int GS = 13;
  static constexpr int OFF_G = OFF_DSM + (NV > 2 * PCAP ? NV : 2 * PCAP);
  static constexpr int SMEM_DOUBLES = OFF_G + (FUSE_G ? NW * GS + (NW * GS & 1) : 0);
  static_assert(SCR_POLISH <= SCR_SIZE, "polish scratch does not fit");
  static_assert(NLS * 21 <= NTRI && 288 + 72 <= NTRI, "staging does not fit in the triangle buffer");
};

// Row-packed lower triangle: element (i, k), k < i, lives at i(i+1)/2 + k (row i owns i+1 slots,
// the last one is unused - the diagonal is kept in dinv[]).  Triangular numbers t_n = n(n+1)/2,
// n = 16m..16m+15, are a complete residue system mod 16, so the 16 rows of a half-warp hit 16
// different 8-byte bank pairs at the same column (row walks are conflict-free) and a fixed row
// is contiguous in k (column walks across lanes touch consecutive words).
MPC_LANEFN int tri(int i) { return (i * (i + 1)) >> 1; }
// ----------------------------------------------------------------- lane-local helpers
MPC_LANEFN double limit_scaling(double v) {  // scaling.c:7-14
  v = v < kMinScaling ? 1.0 : v;
  return v > kMaxScaling ? kMaxScaling : v;
}

// 1/sqrt(x): one MUFU + Newton steps on the device instead of a square root and a division
#if defined(__CUDACC__) && !defined(MPC_EMULATE)
#define MPC_RSQRT(x) rsqrt(x)
#else
#define MPC_RSQRT(x) (1.0 / std::sqrt(x))
#endif

// max / clamp without fmax's NaN-quieting sequence (DSETP + 2 selects instead of ~7 instructions);
// identical results whenever the running value `m` / the bounds are not NaN, which holds at every
// use (a NaN candidate is dropped exactly as fmax would drop it; a NaN t passes through clampd and
// is caught by the finite check on the output)
MPC_LANEFN double maxd(double m, double v) { return v > m ? v : m; }
MPC_LANEFN double clampd(double t, double lo, double hi) { return t < lo ? lo : (t > hi ? hi : t); }

// inverse of a symmetric 3x3 given as (xx,xy,xz,yy,yz,zz)
MPC_LANEFN void inv3sym(const double a[6], double o[6]) {
  const double c00 = a[3] * a[5] - a[4] * a[4];
  const double c01 = a[2] * a[4] - a[1] * a[5];
  const double c02 = a[1] * a[4] - a[2] * a[3];
  const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  const double id = 1.0 / det;
  o[0] = c00 * id;
  o[1] = c01 * id;
  o[2] = c02 * id;
  o[3] = (a[0] * a[5] - a[2] * a[2]) * id;
  o[4] = (a[1] * a[2] - a[0] * a[4]) * id;
  o[5] = (a[0] * a[3] - a[1] * a[1]) * id;
}
MPC_LANEFN void mul3sym(const double m[6], const double v[3], double o[3]) {
  o[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  o[1] = m[1] * v[0] + m[3] * v[1] + m[4] * v[2];
  o[2] = m[2] * v[0] + m[4] * v[1] + m[5] * v[2];
}

// inverse of a 6x6 SPD matrix (full row-major in, full row-major out) via Cholesky
MPC_LANEFN void inv6spd(const double* a, double* out) {
  double L[6][6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = a[j * 6 + j];
#pragma unroll
    for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
    const double id = MPC_RSQRT(s);
    L[j][j] = id;  // store the reciprocal on the diagonal
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double t = a[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k];
      L[i][j] = t * id;
    }
  }
  double Wi[6][6];  // W = L^-1 (lower)
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    Wi[c][c] = L[c][c];
#pragma unroll
    for (int i = c + 1; i < 6; ++i) {
      double t = 0;
#pragma unroll
      for (int k = c; k < i; ++k) t += L[i][k] * Wi[k][c];
      Wi[i][c] = -t * L[i][i];
    }
  }
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) {
      double t = 0;
#pragma unroll
      for (int k = r; k < 6; ++k) t += Wi[k][r] * Wi[k][c];
      out[r * 6 + c] = t;
      out[c * 6 + r] = t;
    }
}

// ---------------------------------------------------------------------------------------
// Dense factorisation kernels of the wrench-space / polish systems (shared memory only).
// ---------------------------------------------------------------------------------------
constexpr int kFactorBlock = 6;
#ifndef MPC_FACTOR_UNROLL
#define MPC_FACTOR_UNROLL 4   // k-loop unrolling of the factorisation dot products (code size vs loads in flight)
#endif
constexpr int kFactorUnroll = MPC_FACTOR_UNROLL;

  // Blocked dense Cholesky (left-looking by block columns of width CB) of the leading n x n block,
  // n a multiple of CB, held as: strictly-lower triangle in Wp (row-packed), diagonal in colb[0..n).
  // On exit Wp holds L (strictly lower), dinv[i] = 1/L_ii.  Row i is owned by lane i % NL.
  // Per block column: (1) every row at or below the block subtracts the contribution of the
  // finished columns from its CB entries - CB independent dot products that share the load of the
  // row's own element; the rows of the diagonal block park theirs in `stage`; (2) every lane
  // factorises the CB x CB diagonal block redundantly in registers (same latency as one lane doing
  // it, no third barrier) while forward-substituting its own row.  Two barriers per block column.
template <int PCAP>
MPC_NOINLINE_FN void cholesky_blocked(double* Wp, double* colb, double* dinv, int n MPC_LANE_ARG) {
  constexpr int NL = MPC_NL, CB = kFactorBlock;
  {
    double* const hdiag = colb;
    double* const stage = colb + PCAP;  // [CB(CB+1)/2] updated diagonal block, row-packed with diagonal
    for (int J = 0; J < n; J += CB) {
      const double* const pj = Wp + tri(J);  // row J+c starts at pj + c*J + c(c+1)/2
      // once at most 32 rows remain, the two warps share each row: CB/2 columns per lane
      const bool split = (NL == 64) && (n - J <= 32);
      LANES_BEGIN
      if (!split) {
        for (int i = lane; i < n; i += NL) {
          if (i < J) continue;
          double* const rowi = Wp + tri(i);
          const int r = i - J;  // < CB: a row of the diagonal block
          double acc[CB];
#pragma unroll
          for (int c = 0; c < CB; ++c) acc[c] = (c < r) ? rowi[J + c] : (c == r ? hdiag[i] : 0.0);
#pragma unroll(kFactorUnroll)
          for (int k = 0; k < J; ++k) {
            const double li = rowi[k];
#pragma unroll
            for (int c = 0; c < CB; ++c) acc[c] -= li * pj[c * J + (c * (c + 1)) / 2 + k];
          }
          if (r < CB) {
#pragma unroll
            for (int c = 0; c < CB; ++c)
              if (c <= r) stage[tri(r) + c] = acc[c];
          } else {
#pragma unroll
            for (int c = 0; c < CB; ++c) rowi[J + c] = acc[c];
          }
        }
      } else {
        constexpr int CH = CB / 2;
        const int i = J + (lane & 31), c0 = (lane >> 5) * CH;
        if (i < n) {
          double* const rowi = Wp + tri(i);
          const int r = i - J;
          double acc[CH];
          int off[CH];
#pragma unroll
          for (int q = 0; q < CH; ++q) {
            const int c = c0 + q;
            off[q] = c * J + (c * (c + 1)) / 2;
            acc[q] = (c < r) ? rowi[J + c] : (c == r ? hdiag[i] : 0.0);
          }
#pragma unroll(kFactorUnroll)
          for (int k = 0; k < J; ++k) {
            const double li = rowi[k];
#pragma unroll
            for (int q = 0; q < CH; ++q) acc[q] -= li * pj[off[q] + k];
          }
#pragma unroll
          for (int q = 0; q < CH; ++q) {
            const int c = c0 + q;
            if (r >= CB) rowi[J + c] = acc[q];
            else if (c <= r) stage[tri(r) + c] = acc[q];
          }
        }
      }
      LANES_END
      WSYNC();
      LANES_BEGIN
      for (int i = lane; i < n; i += NL) {
        if (i < J) continue;
        double* const rowi = Wp + tri(i);
        const int r = i - J;
        double S[CB * (CB + 1) / 2], x[CB];
#pragma unroll
        for (int e = 0; e < CB * (CB + 1) / 2; ++e) S[e] = stage[e];
#pragma unroll
        for (int c = 0; c < CB; ++c) x[c] = (r >= CB) ? rowi[J + c] : (c <= r ? stage[tri(r) + c] : 0.0);
#pragma unroll
        for (int c = 0; c < CB; ++c) {  // right-looking on the register copy; x rides along
          const double inv = MPC_RSQRT(S[tri(c) + c]);
          if (c == r) dinv[i] = inv;
          x[c] *= inv;
#pragma unroll
          for (int r1 = c + 1; r1 < CB; ++r1) {
            S[tri(r1) + c] *= inv;
            x[r1] -= x[c] * S[tri(r1) + c];
          }
#pragma unroll
          for (int r1 = c + 1; r1 < CB; ++r1)
#pragma unroll
            for (int r2 = c + 1; r2 <= r1; ++r2) S[tri(r1) + r2] -= S[tri(r1) + c] * S[tri(r2) + c];
        }
#pragma unroll
        for (int c = 0; c < CB; ++c)
          if (J + c < i) rowi[J + c] = x[c];
      }
      LANES_END
      WSYNC();
    }
  }
}

  // in-place inverse of the lower factor: W = L^-1 (strictly-lower part in Wp, W_ii = dinv[i]), by
  // block columns J = n-CB .. 0:  W_IJ = -(sum_{K>J} W_IK L_KJ) L_JJ^-1.  A row accumulates its CB
  // sums (independent chains), back-substitutes through L_JJ and holds the result in registers
  // across the single barrier of the step (every other row's L_KJ must have been read before its
  // owner overwrites it); the next step reads only columns left of J, so no second barrier.
template <int PCAP>
MPC_NOINLINE_FN void invert_lower_blocked(double* Wp, const double* dinv, int n MPC_LANE_ARG) {
  constexpr int NL = MPC_NL, CB = kFactorBlock;
  constexpr int NRS = (PCAP + NL - 1) / NL;  // rows of the dense factor per lane
  {
    PER_LANE_ARR(double, Tw, [NRS][CB]);
    for (int J = n - CB; J >= 0; J -= CB) {
      const double* const pj = Wp + tri(J) + J;  // L_JJ[c2][c] = pj[c2*J + c2(c2+1)/2 + c]
      LANES_BEGIN
#pragma unroll
      for (int s = 0; s < NRS; ++s) {
        const int i = lane + NL * s;
        if (i >= J && i < n) {
          const double* const rowi = Wp + tri(i);
          const int r = i - J;
          double S[CB];
          if (r < CB) {  // a row of L_JJ^-1 itself: e_r L_JJ^-1
#pragma unroll
            for (int c = 0; c < CB; ++c) S[c] = (c == r) ? -1.0 : 0.0;
          } else {
            const double di = dinv[i];
#pragma unroll
            for (int c = 0; c < CB; ++c) S[c] = di * rowi[J + c];
            const double* p = Wp + tri(J + CB) + J;  // &L[k][J]
#pragma unroll(kFactorUnroll)
            for (int k = J + CB; k < i; ++k) {
              const double w = rowi[k];
#pragma unroll
              for (int c = 0; c < CB; ++c) S[c] += w * p[c];
              p += k + 1;
            }
          }
#pragma unroll
          for (int c = CB - 1; c >= 0; --c) {  // T L_JJ = -S
            double t = -S[c];
#pragma unroll
            for (int c2 = c + 1; c2 < CB; ++c2) t -= _(Tw)[s][c2] * pj[c2 * J + (c2 * (c2 + 1)) / 2 + c];
            _(Tw)[s][c] = t * dinv[J + c];
          }
        }
      }
      LANES_END
      WSYNC();
      LANES_BEGIN
#pragma unroll
      for (int s = 0; s < NRS; ++s) {
        const int i = lane + NL * s;
        if (i >= J && i < n) {
          double* const rowi = Wp + tri(i);
#pragma unroll
          for (int c = 0; c < CB; ++c)
            if (J + c < i) rowi[J + c] = _(Tw)[s][c];
        }
      }
      LANES_END
    }
    WSYNC();
  }
}


// ---------------------------------------------------------------------------------------
// The solver.  `rec` = problem record (float, MPC_REC_* layout), `rp` = robot constant row.
// `sm` = this warp's shared-memory block (Dims<H>::SMEM_DOUBLES doubles); Gamma/N tables at
// OFF_GAM/OFF_NN must have been filled by fill_tables<H>() before the first call.
// Outputs: grf0[12] (negated horizon-0 forces), grf_full[12H] or null, status, info or null.
// ---------------------------------------------------------------------------------------
template <int H, int PCAP_ = 6 * H>
MPC_FN void fill_tables(double* sm, double dt
#if defined(__CUDACC__) && !defined(MPC_EMULATE)
                        , const int lane
#endif
) {
  using D = Dims<H, PCAP_>;
  LANES_BEGIN
  for (int e = lane; e < H * H; e += NL) {
    const int i = e / H, j = e % H;
    const int k0 = i > j ? i : j;
    double g = 0;
    for (int k = k0; k < H; ++k)
      g += (dt * dt * ((k - i) + 0.5)) * (dt * dt * ((k - j) + 0.5));
    sm[D::OFF_GAM + e] = g;
    sm[D::OFF_NN + e] = (double)(H - k0);
  }
  LANES_END
  WSYNC();
}

// Phases of one solve (template mask PH): the monolithic kernels run PH_ALL; the phase-split batch path
// runs the three phases as three code-homogeneous kernels that hand the robot's state over through a
// small global-memory workspace row `ws` (WsLayout<H>, ~9 KB at h = 10, L2-resident for a batch):
//   PH_SETUP   structure + q + Ruiz scaling                      -> ws: c, Wb/Q1'/Q2', per-leg-step q, D, E, contact
//   PH_ADMM    rho, factorisations, ADMM loop to termination     -> ws: x, z, y, status, iterations, rho, residuals
//   PH_POLISH  polish (+ acceptance)                             -> outputs
// Every phase that ends a robot's solve (ADMM when no polish is due, POLISH otherwise) writes the outputs.
enum { PH_SETUP = 1, PH_ADMM = 2, PH_POLISH = 4, PH_ALL = 7 };
enum { WS_C = 0, WS_STAGE = 1, WS_STATUS = 2, WS_ITER = 3, WS_RHO = 4, WS_RHO_UPDATES = 5, WS_PRI = 6, WS_DUA = 7, WS_HDR = 8 };
constexpr double kWsStageAdmm = 1.0, kWsStagePolish = 2.0, kWsStageDone = 3.0;
template <int H>
struct WsLayout {
  using D = Dims<H>;
  static constexpr int SM_BLOCK = D::OFF_DSM - D::OFF_WB;       // Wb, Q1', Q2' as they sit in shared memory
  static constexpr int OFF_SM = WS_HDR;
  static constexpr int OFF_LS = OFF_SM + SM_BLOCK;               // [12][NLS]: q(3) D(3) E(5) contact(1), lane-contiguous
  static constexpr int OFF_XZY = OFF_LS + 12 * D::NLS;           // [13][NLS]: x(3) z(5) y(5)
  static constexpr int STRIDE = (OFF_XZY + 13 * D::NLS + 1) & ~1;
};

template <int H, bool WARM = false, int PCAP_ = 6 * H, int PH = PH_ALL>
MPC_FN void solve_robot(const float* rec, const float* rp, const SolveParams& prm, double* sm,
                        float* grf0, float* grf_full, int32_t* status_out, float* resid_out,
                        double* info_out, double* dbg_out, double* warm, double* ws
#if defined(__CUDACC__) && !defined(MPC_EMULATE)
                        , const int lane
#endif
) {
  using D = Dims<H, PCAP_>;
  using WL = WsLayout<H>;
  constexpr bool DO_SETUP = (PH & PH_SETUP) != 0, DO_ADMM = (PH & PH_ADMM) != 0, DO_POLISH = (PH & PH_POLISH) != 0;
  static_assert(PH == PH_ALL || PCAP_ == 6 * H, "the phase-split path uses the fast polish capacity only");
  constexpr int PCAP = D::PCAP;
  // `warm` (MPC_WARM_STRIDE(H) doubles in global memory, or null): OSQP's persistent workspace
  // state between consecutive solves of the same robot (settings.warm_start, mpc_osqp.cc:708,
  // :760-778): [0] flag (0 = none yet, 1 / 2 = which q slot holds the previous problem's q),
  // [1] rho, then x[12h], z[20h], y[20h] (scaled iterates, left as they are when the data is
  // re-scaled, exactly as osqp_update_P_A does) and two slots of UNSCALED q[12h]: osqp_update_P_A
  // re-runs scale_data before osqp_update_lin_cost installs the new q, so the cost normalisation
  // c sees the OLD q.  The new q goes to the other slot and the flag flips only when the solve
  // completes, so a robot handed to the overflow kernel is re-solved from the same state.
  // WARM is a compile-time switch so that the cold-start kernel carries none of this code.
  bool warm_valid = false;
  int q_prev_off = 0, q_new_off = 0;
  if constexpr (WARM) {
    warm_valid = warm != nullptr && warm[0] != 0.0;
    const bool prev_in_b = warm != nullptr && warm[0] == 2.0;
    q_prev_off = 2 + 52 * H + (prev_in_b ? 12 * H : 0);
    q_new_off = 2 + 52 * H + (prev_in_b ? 0 : (warm_valid ? 12 * H : 0));
  }
  constexpr int NS = D::NSLOT, NW = D::NW, NV = D::NV;
  double* const Wp = sm + D::OFF_WP;
  double* const vecA = sm + D::OFF_VECA;
  double* const vecB = sm + D::OFF_VECB;
  double* const vt = sm + D::OFF_VT;
  double* const vg = sm + D::OFF_VG;
  double* const vw = sm + D::OFF_VW;
  double* const Finv = sm + D::OFF_FINV;
  double* const dinv = sm + D::OFF_DINV;
  double* const colb = sm + D::OFF_COL;
  double* const q1s = sm + D::OFF_Q1S;
  double* const q2s = sm + D::OFF_Q2S;
  const double* const gam = sm + D::OFF_GAM;
  const double* const nnm = sm + D::OFF_NN;
  double* const Wb = sm + D::OFF_WB;
  constexpr int WBS = D::WBS;
  double* const Dsm = sm + D::OFF_DSM;

  const double dt = prm.dt, alpha = prm.alpha;

  // ------------------------------------------------------------------ uniform inputs
  const double mass = rp[0];
  const double mu0 = rec[MPC_REC_MU + 0], mu1 = rec[MPC_REC_MU + 1];
  const double mu2 = rec[MPC_REC_MU + 2], mu3 = rec[MPC_REC_MU + 3];
  const double fz_max = mass * kGravity * kMaxScale;            // mpc_osqp.cc:685-686
  const double fz_min = mass * kGravity * kMinScale;
  const double f_ub = (mu0 + 1.0) * fz_max;                     // mpc_osqp.cc:466-467
  const double roll = rec[MPC_REC_RPY], pitch = rec[MPC_REC_RPY + 1], yaw = rec[MPC_REC_RPY + 2];
  const double sr = sin(roll), cr = cos(roll), sp = sin(pitch), cp = cos(pitch);
  const double sy = sin(yaw), cy = cos(yaw);
  const double tp = tan(pitch);
  // T: angular velocity -> rpy rate (mpc_osqp.cc:306-312)
  const double T00 = cy / cp, T01 = sy / cp, T10 = -sy, T11 = cy, T20 = cy * tp, T21 = sy * tp;
  const double w_rpy0 = rec[MPC_REC_WEIGHTS + 0], w_rpy1 = rec[MPC_REC_WEIGHTS + 1],
               w_rpy2 = rec[MPC_REC_WEIGHTS + 2];
  double balance;  // torque-row balancing factor s (conditioning only)

  // ------------------------------------------------------------------ per-lane state
  PER_LANE_ARR(double, wbl, [6][3]);   // balanced wrench map block of this lane's leg
  PER_LANE_ARR(double, xv, [NS][3]);
  PER_LANE_ARR(double, zv, [NS][5]);
  PER_LANE_ARR(double, yv, [NS][5]);
  PER_LANE_ARR(double, qb, [NS][3]);   // scaled linear cost
  PER_LANE_ARR(double, Dv, [NS][3]);   // Ruiz D
  PER_LANE_ARR(double, Ev, [NS][5]);   // Ruiz E
  PER_LANE_ARR(double, di, [NS][6]);   // Delta^-1 (sym 3x3)
  PER_LANE_ARR(double, rhoj, [NS]);    // rho of this leg-step's rows
  PER_LANE_ARR(double, cj, [NS]);      // contact flag as 0/1 double
  PER_LANE_ARR(int, valid, [NS]);
  PER_LANE_ARR(double, tmp3, [NS][3]);
  PER_LANE_ARR(double, xt, [NS][3]);
  PER_LANE(double, red);               // reduction operand
  PER_LANE(double, red2);

  double cscale = 1.0;  // OSQP's cost scaling c
  if constexpr (DO_SETUP) {
  // ------------------------------------------------------------------ S1: structure
  LANES_BEGIN
  const int leg = lane & 3;
  // R = Rz Ry Rx (mpc_osqp.cc:283-291); Rf = Rx Ry Rz (mpc_osqp.cc:606-617, Q3)
  const double R[3][3] = {{cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr},
                          {sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr},
                          {-sp, cp * sr, cp * cr}};
  const double Rf[3][3] = {{cp * cy, -cp * sy, sp},
                           {sr * sp * cy + cr * sy, -sr * sp * sy + cr * cy, -sr * cp},
                           {-cr * sp * cy + sr * sy, cr * sp * sy + sr * cy, cr * cp}};
  // inertia_(inertia.data()): column-major read of the row-major list (mpc_osqp.cc:516)
  double In[3][3], Ii[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) In[i][j] = rp[1 + j * 3 + i];
  {
    const double c00 = In[1][1] * In[2][2] - In[1][2] * In[2][1];
    const double c01 = In[1][2] * In[2][0] - In[1][0] * In[2][2];
    const double c02 = In[1][0] * In[2][1] - In[1][1] * In[2][0];
    const double idet = 1.0 / (In[0][0] * c00 + In[0][1] * c01 + In[0][2] * c02);
    Ii[0][0] = c00 * idet;
    Ii[0][1] = (In[0][2] * In[2][1] - In[0][1] * In[2][2]) * idet;
    Ii[0][2] = (In[0][1] * In[1][2] - In[0][2] * In[1][1]) * idet;
    Ii[1][0] = c01 * idet;
    Ii[1][1] = (In[0][0] * In[2][2] - In[0][2] * In[2][0]) * idet;
    Ii[1][2] = (In[0][2] * In[1][0] - In[0][0] * In[1][2]) * idet;
    Ii[2][0] = c02 * idet;
    Ii[2][1] = (In[0][1] * In[2][0] - In[0][0] * In[2][1]) * idet;
    Ii[2][2] = (In[0][0] * In[1][1] - In[0][1] * In[1][0]) * idet;
  }
  double RI[3][3], Iw[3][3];  // Iw = R Ii R^T (mpc_osqp.cc:670-671)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      RI[i][j] = R[i][0] * Ii[0][j] + R[i][1] * Ii[1][j] + R[i][2] * Ii[2][j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Iw[i][j] = RI[i][0] * R[j][0] + RI[i][1] * R[j][1] + RI[i][2] * R[j][2];
  const double px = rec[MPC_REC_FOOTPOS + 3 * leg], py = rec[MPC_REC_FOOTPOS + 3 * leg + 1],
               pz = rec[MPC_REC_FOOTPOS + 3 * leg + 2];
  const double rx = Rf[0][0] * px + Rf[0][1] * py + Rf[0][2] * pz;
  const double ry = Rf[1][0] * px + Rf[1][1] * py + Rf[1][2] * pz;
  const double rz = Rf[2][0] * px + Rf[2][1] * py + Rf[2][2] * pz;
  // Bw = Iw * skew(r) (mpc_osqp.cc:293-297, 324-336)
  const double sk[3][3] = {{0, -rz, ry}, {rz, 0, -rx}, {-ry, rx, 0}};
  double fro = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const double v = Iw[a][0] * sk[0][b] + Iw[a][1] * sk[1][b] + Iw[a][2] * sk[2][b];
      _(wbl)[a][b] = v;
      _(wbl)[3 + a][b] = (a == b) ? 1.0 / mass : 0.0;
      fro += v * v;
    }
  _(red) = lane < 4 ? fro : 0.0;
  LANES_END
  {
    const double fro_all = WSUM(red);
    balance = sqrt((12.0 / (mass * mass)) / fro_all);
  }
  const double ib = 1.0 / balance, ib2 = ib * ib;
  // Q1' = S^-1 blkdiag(T^T W_rpy T, W_pos) S^-1 ; Q2' = S^-1 dt^2 blkdiag(W_w, W_v) S^-1
  double Q1t[6];  // (xx,xy,xz,yy,yz,zz) of T^T W T / s^2
  {
    const double a0 = w_rpy0, a1 = w_rpy1, a2 = w_rpy2;
    // T columns: col0 = (T00,T10,T20), col1 = (T01,T11,T21), col2 = (0,0,1)
    Q1t[0] = (T00 * a0 * T00 + T10 * a1 * T10 + T20 * a2 * T20) * ib2;
    Q1t[1] = (T00 * a0 * T01 + T10 * a1 * T11 + T20 * a2 * T21) * ib2;
    Q1t[2] = (T20 * a2) * ib2;
    Q1t[3] = (T01 * a0 * T01 + T11 * a1 * T11 + T21 * a2 * T21) * ib2;
    Q1t[4] = (T21 * a2) * ib2;
    Q1t[5] = a2 * ib2;
  }
  const double q1p[3] = {(double)rec[MPC_REC_WEIGHTS + 3], (double)rec[MPC_REC_WEIGHTS + 4],
                         (double)rec[MPC_REC_WEIGHTS + 5]};
  const double q2d[6] = {rec[MPC_REC_WEIGHTS + 6] * dt * dt * ib2, rec[MPC_REC_WEIGHTS + 7] * dt * dt * ib2,
                         rec[MPC_REC_WEIGHTS + 8] * dt * dt * ib2, rec[MPC_REC_WEIGHTS + 9] * dt * dt,
                         rec[MPC_REC_WEIGHTS + 10] * dt * dt,      rec[MPC_REC_WEIGHTS + 11] * dt * dt};
  // Q1'[c][c'] accessor
  auto Q1at = [&](int c, int c2) -> double {
    if (c < 3 && c2 < 3) {
      const int lo = c < c2 ? c : c2, hi = c < c2 ? c2 : c;
      return Q1t[lo == 0 ? hi : (lo == 1 ? 2 + hi : 5)];
    }
    return (c == c2) ? q1p[c - 3] : 0.0;
  };

  LANES_BEGIN
  const int leg = lane & 3;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) _(wbl)[a][b] *= balance;
  if (lane < 4) {
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
      for (int b = 0; b < 3; ++b) Wb[c * WBS + 3 * leg + b] = _(wbl)[c][b];
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int j = lane + NL * s;
    _(valid)[s] = j < D::NLS;
    _(cj)[s] = _(valid)[s] ? (double)rec[MPC_REC_CONTACT + j] : 0.0;
  }
  // per-step tracking-error terms a_k (-> vg) and b_k (-> vw), unbalanced wrench coords
  if (lane < H) {
    const int k = lane;
    const double t = dt * (k + 1);
    const double g = -kGravity;
    double e[12];
    const double om[3] = {(double)rec[MPC_REC_ANGVEL], (double)rec[MPC_REC_ANGVEL + 1],
                          (double)rec[MPC_REC_ANGVEL + 2]};
    const double ve[3] = {(double)rec[MPC_REC_COM_VEL], (double)rec[MPC_REC_COM_VEL + 1],
                          (double)rec[MPC_REC_COM_VEL + 2]};
    const double po[3] = {(double)rec[MPC_REC_COM_POS], (double)rec[MPC_REC_COM_POS + 1],
                          (double)rec[MPC_REC_COM_POS + 2]};
    const double nr[3] = {(double)rec[MPC_REC_NORMAL], (double)rec[MPC_REC_NORMAL + 1],
                          (double)rec[MPC_REC_NORMAL + 2]};
    const double dv[3] = {(double)rec[MPC_REC_DES_VEL], (double)rec[MPC_REC_DES_VEL + 1],
                          (double)rec[MPC_REC_DES_VEL + 2]};
    const double dw[3] = {(double)rec[MPC_REC_DES_ANGVEL], (double)rec[MPC_REC_DES_ANGVEL + 1],
                          (double)rec[MPC_REC_DES_ANGVEL + 2]};
    // desired states (mpc_osqp.cc:635-659)
    const double des[12] = {(double)rec[MPC_REC_DES_RPY], (double)rec[MPC_REC_DES_RPY + 1], yaw + t * dw[2],
                            t * dv[0] + po[0], t * dv[1] + po[1], (double)rec[MPC_REC_DES_POS + 2],
                            dw[0], dw[1], dw[2], dv[0], dv[1], 0.0};
    if (k < H - 1) {  // a_qp block k = A_exp^(k+1); the LAST block stays zero (mpc_osqp.cc:361)
      e[0] = roll + t * (T00 * om[0] + T01 * om[1]) - des[0];
      e[1] = pitch + t * (T10 * om[0] + T11 * om[1]) - des[1];
      e[2] = yaw + t * (T20 * om[0] + T21 * om[1] + om[2]) - des[2];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        e[3 + a] = po[a] + t * ve[a] + 0.5 * t * t * nr[a] * g - des[3 + a];
        e[6 + a] = om[a] - des[6 + a];
        e[9 + a] = ve[a] + t * nr[a] * g - des[9 + a];
      }
    } else {
#pragma unroll
      for (int a = 0; a < 12; ++a) e[a] = -des[a];
    }
    const double we0 = w_rpy0 * e[0], we1 = w_rpy1 * e[1], we2 = w_rpy2 * e[2];
    vg[6 * k + 0] = T00 * we0 + T10 * we1 + T20 * we2;
    vg[6 * k + 1] = T01 * we0 + T11 * we1 + T21 * we2;
    vg[6 * k + 2] = we2;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      vg[6 * k + 3 + a] = rec[MPC_REC_WEIGHTS + 3 + a] * e[3 + a];
      vw[6 * k + a] = dt * rec[MPC_REC_WEIGHTS + 6 + a] * e[6 + a];
      vw[6 * k + 3 + a] = dt * rec[MPC_REC_WEIGHTS + 9 + a] * e[9 + a];
    }
  }
  LANES_END
  WSYNC();
  // acc_j = sum_{k>=j} c_{k-j} a_k + b_k ; balanced (rows 0..2 divided by s) -> vt
  LANES_BEGIN
  for (int i = lane; i < NW; i += NL) {
    const int j = i / 6, c = i % 6;
    double acc = 0;
    for (int k = j; k < H; ++k) acc += (dt * dt * ((k - j) + 0.5)) * vg[6 * k + c] + vw[6 * k + c];
    vt[i] = c < 3 ? acc * ib : acc;
  }
  LANES_END
  WSYNC();
  // q = 2 Wb'^T acc' (mpc_osqp.cc:681-683); initialise D = E = 1
  LANES_BEGIN
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int j = lane + NL * s;
    const int k = (j >> 2) < H ? (j >> 2) : 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double q = 0;
#pragma unroll
      for (int c = 0; c < 6; ++c) q += Wb[c * WBS + 3 * (lane & 3) + a] * vt[6 * k + c];
      const double qn = _(valid)[s] ? 2.0 * q : 0.0;
      _(qb)[s][a] = qn;
      if (dbg_out && _(valid)[s]) dbg_out[3 * j + a] = qn;   // unscaled q for builder tests
      if constexpr (WARM) {
        if (warm && _(valid)[s]) {
          // the workspace keeps the UNSCALED q; scale_data below must see the OLD one (if any)
          if (warm_valid) _(qb)[s][a] = warm[q_prev_off + 3 * j + a];
          warm[q_new_off + 3 * j + a] = qn;
        }
      }
      _(Dv)[s][a] = 1.0;
      _(xv)[s][a] = 0.0;
    }
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      _(Ev)[s][r] = 1.0;
      _(zv)[s][r] = 0.0;
      _(yv)[s][r] = 0.0;
    }
  }
  for (int v = lane; v < NV; v += NL) Dsm[v] = 1.0;
  LANES_END
  WSYNC();
  // ------------------------------------------------------------------ S2: Ruiz scaling
  // G1 = Wb'^T Q1' Wb', G2 = Wb'^T Q2' Wb' (12x12) staged in the (still unused) triangle buffer.
  double* const G1 = Wp;
  double* const G2 = Wp + 144;
  double* const Y1 = Wp + 288;  // Q1' Wb' (6x12)
  LANES_BEGIN
  for (int e = lane; e < 36; e += NL) q1s[e] = Q1at(e / 6, e % 6);
  if (lane < 6) q2s[lane] = q2d[lane];
  for (int e = lane; e < 72; e += NL) {
    const int c = e / 12, b = e % 12;
    double y1 = 0;
    if (c < 3) {
#pragma unroll
      for (int c2 = 0; c2 < 3; ++c2) y1 += Q1at(c, c2) * Wb[c2 * WBS + b];
    } else {
      y1 = q1p[c - 3] * Wb[c * WBS + b];
    }
    Y1[e] = y1;
  }
  LANES_END
  WSYNC();
  LANES_BEGIN
  for (int e = lane; e < 144; e += NL) {
    const int a = e / 12, b = e % 12;
    double g1 = 0, g2 = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      g1 += Wb[c * WBS + a] * Y1[c * 12 + b];
      g2 += Wb[c * WBS + a] * q2d[c] * Wb[c * WBS + b];
    }
    G1[e] = g1;
    G2[e] = g2;
  }
  LANES_END
  WSYNC();
  if (dbg_out) {  // dense P (unscaled) for builder tests: dbg_out[NV + i*NV + v]
    LANES_BEGIN
    for (int v = lane; v < NV; v += NL) {
      const int k2 = v / 12, b = v % 12;
#pragma unroll 1
      for (int k = 0; k < H; ++k)
#pragma unroll 1
        for (int a = 0; a < 12; ++a) {
          double p = 2.0 * (gam[k * H + k2] * G1[a * 12 + b] + nnm[k * H + k2] * G2[a * 12 + b]);
          if (12 * k + a == v) p += alpha;
          dbg_out[NV + (12 * k + a) * NV + v] = p;
        }
    }
    LANES_END
  }

  // column inf-norms of (D P D) WITHOUT the c factor -> vecA   (scaling.c:30-43, lin_alg.c:358)
  // (code size matters more than ILP here: the kernel's phases compete for the instruction cache,
  // so the k loop is kept rolled and the lambda has ONE call site)
  auto col_norms_P = [&]() {
    LANES_BEGIN
#pragma unroll 1
    for (int v = lane; v < NV; v += NL) {
      const int k2 = v / 12, b = v % 12;
      double g1[12], g2[12];
#pragma unroll
      for (int a = 0; a < 12; ++a) {
        g1[a] = 2.0 * G1[a * 12 + b];
        g2[a] = 2.0 * G2[a * 12 + b];
      }
      double m0 = 0, m1 = 0;
#pragma unroll 1
      for (int k = 0; k < H; ++k) {
        const double gk = gam[k * H + k2], nk = nnm[k * H + k2];
        const double* dk = Dsm + 12 * k;
#pragma unroll
        for (int a = 0; a < 12; a += 2) {
          m0 = maxd(m0, dk[a] * fabs(gk * g1[a] + nk * g2[a]));
          m1 = maxd(m1, dk[a + 1] * fabs(gk * g1[a + 1] + nk * g2[a + 1]));
        }
      }
      // diagonal entry carries + alpha (P_vv >= 0, so |P_vv + alpha| only grows the max)
      const double pd = 2.0 * (gam[k2 * H + k2] * G1[b * 12 + b] + nnm[k2 * H + k2] * G2[b * 12 + b]) + alpha;
      vecA[v] = maxd(maxd(m0, m1), Dsm[v] * pd) * Dsm[v];
    }
    LANES_END
    WSYNC();
  };
  // Ruiz passes: col_norms -> D/E update -> col_norms -> cost normalisation -> D/E update -> ...
  // written as one loop whose top is the only col_norms_P() call site.
  if (prm.scaling_iters > 0) {
    for (int it = 0;; ++it) {
      col_norms_P();
      if (it > 0) {
        // cost normalisation of pass it-1 (scaling.c:117-146)
        LANES_BEGIN
        double ssum = 0, qmax = 0;
        for (int v = lane; v < NV; v += NL) ssum += cscale * vecA[v];
    #pragma unroll
        for (int s = 0; s < NS; ++s)
          if (_(valid)[s])
    #pragma unroll
            for (int a = 0; a < 3; ++a) qmax = maxd(qmax, fabs(_(qb)[s][a]));
        _(red) = ssum;
        _(red2) = qmax;
        LANES_END
        {
          const double mean = WSUM(red) / (double)NV;
          const double nq = limit_scaling(WMAX(red2));
          const double ctemp = 1.0 / limit_scaling(fmax(mean, nq));
          cscale *= ctemp;
          LANES_BEGIN
    #pragma unroll
          for (int s = 0; s < NS; ++s)
    #pragma unroll
            for (int a = 0; a < 3; ++a) _(qb)[s][a] *= ctemp;
          LANES_END
        }
        }
      if (it == prm.scaling_iters) break;
      // D_temp, E_temp from the current scaled P, A (scaling.c:78-93)
      LANES_BEGIN
  #pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int j = lane + NL * s;
        if (_(valid)[s]) {
          const double* E = _(Ev)[s];
          const double* Dd = _(Dv)[s];
          const double am0 = fabs(mu0), am1 = fabs(mu1), am2 = fabs(mu2), am3 = fabs(mu3);
          double colA[3];
          colA[0] = maxd(E[0], E[1]) * Dd[0];
          colA[1] = maxd(E[2], E[3]) * Dd[1];
          colA[2] = maxd(maxd(maxd(E[0] * am0, E[1] * am1), maxd(E[2] * am2, E[3] * am3)), E[4]) * Dd[2];
          double rowA[5];
          rowA[0] = E[0] * maxd(Dd[0], am0 * Dd[2]);
          rowA[1] = E[1] * maxd(Dd[0], am1 * Dd[2]);
          rowA[2] = E[2] * maxd(Dd[1], am2 * Dd[2]);
          rowA[3] = E[3] * maxd(Dd[1], am3 * Dd[2]);
          rowA[4] = E[4] * Dd[2];
  #pragma unroll
          for (int a = 0; a < 3; ++a) {
            const double nrm = maxd(cscale * vecA[3 * j + a], colA[a]);
            const double dtmp = MPC_RSQRT(limit_scaling(nrm));
            _(Dv)[s][a] *= dtmp;
            _(qb)[s][a] *= dtmp;
            Dsm[3 * j + a] = _(Dv)[s][a];
          }
  #pragma unroll
          for (int r = 0; r < 5; ++r) _(Ev)[s][r] *= MPC_RSQRT(limit_scaling(rowA[r]));
        }
      }
      LANES_END
      WSYNC();
    }
  }
  }  // DO_SETUP
  if constexpr (DO_SETUP && !DO_ADMM) {   // hand the scaled problem over to the ADMM kernel
    LANES_BEGIN
    for (int e = lane; e < WL::SM_BLOCK; e += NL) ws[WL::OFF_SM + e] = sm[D::OFF_WB + e];
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (_(valid)[s]) {
        const int j = lane + NL * s;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          ws[WL::OFF_LS + a * D::NLS + j] = _(qb)[s][a];
          ws[WL::OFF_LS + (3 + a) * D::NLS + j] = _(Dv)[s][a];
        }
#pragma unroll
        for (int r = 0; r < 5; ++r) ws[WL::OFF_LS + (6 + r) * D::NLS + j] = _(Ev)[s][r];
        ws[WL::OFF_LS + 11 * D::NLS + j] = _(cj)[s];
      }
    if (lane == 0) {
      ws[WS_C] = cscale;
      ws[WS_STAGE] = kWsStageAdmm;
    }
    LANES_END
    return;
  }
  if constexpr (!DO_SETUP) {              // ... and pick it up
    cscale = ws[WS_C];
    LANES_BEGIN
    for (int e = lane; e < WL::SM_BLOCK; e += NL) sm[D::OFF_WB + e] = ws[WL::OFF_SM + e];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int j = lane + NL * s;
      _(valid)[s] = j < D::NLS;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        _(qb)[s][a] = _(valid)[s] ? ws[WL::OFF_LS + a * D::NLS + j] : 0.0;
        _(Dv)[s][a] = _(valid)[s] ? ws[WL::OFF_LS + (3 + a) * D::NLS + j] : 1.0;
        _(xv)[s][a] = 0.0;
      }
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        _(Ev)[s][r] = _(valid)[s] ? ws[WL::OFF_LS + (6 + r) * D::NLS + j] : 1.0;
        _(zv)[s][r] = 0.0;
        _(yv)[s][r] = 0.0;
      }
      _(cj)[s] = _(valid)[s] ? ws[WL::OFF_LS + 11 * D::NLS + j] : 0.0;
    }
    LANES_END
    WSYNC();
  }
  const double cinv = 1.0 / cscale;
  if (dbg_out) {
    LANES_BEGIN
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (_(valid)[s]) {
        const int j = lane + NL * s;
#pragma unroll
        for (int a = 0; a < 3; ++a) dbg_out[NV + NV * NV + 3 * j + a] = _(Dv)[s][a];
#pragma unroll
        for (int r = 0; r < 5; ++r) dbg_out[NV + NV * NV + NV + 5 * j + r] = _(Ev)[s][r];
      }
    if (lane == 0) dbg_out[NV + NV * NV + NV + 20 * H] = cscale;
    LANES_END
  }

  // ------------------------------------------------------------------ helpers
  // scaled bounds of row r of a slot (mpc_osqp.cc:449-477 scaled by E, osqp.c:scale_data)
  auto lo_of = [&](const double* E, double c, int r) -> double { return r == 4 ? E[4] * (fz_min * c) : 0.0; };
  auto hi_of = [&](const double* E, double c, int r) -> double {
    return r == 4 ? E[4] * (fz_max * c) : E[r] * (f_ub * c);
  };
  // Abar v : rows of a leg-step (A block mpc_osqp.cc:437-447 scaled E A D)
  auto A_mul = [&](const double* E, const double* Dd, const double* v, double* o) {
    const double dx = Dd[0] * v[0], dy = Dd[1] * v[1], dz = Dd[2] * v[2];
    o[0] = E[0] * (-dx + mu0 * dz);
    o[1] = E[1] * (dx + mu1 * dz);
    o[2] = E[2] * (-dy + mu2 * dz);
    o[3] = E[3] * (dy + mu3 * dz);
    o[4] = E[4] * dz;
  };
  auto At_mul = [&](const double* E, const double* Dd, const double* w, double* o) {
    const double w0 = E[0] * w[0], w1 = E[1] * w[1], w2 = E[2] * w[2], w3 = E[3] * w[3], w4 = E[4] * w[4];
    o[0] = Dd[0] * (w1 - w0);
    o[1] = Dd[1] * (w3 - w2);
    o[2] = Dd[2] * (mu0 * w0 + mu1 * w1 + mu2 * w2 + mu3 * w3 + w4);
  };

  constexpr int CB = kFactorBlock;
  static_assert(PCAP % CB == 0 && NW % CB == 0, "factor sizes must be multiples of the block width");
  auto cholesky = [&](int n) { cholesky_blocked<PCAP>(Wp, colb, dinv, n MPC_LANE_PASS); };
  auto invert_lower = [&](int n) { invert_lower_blocked<PCAP>(Wp, dinv, n MPC_LANE_PASS); };

  // factor(): Delta^-1 per leg-step, F_k^-1, Hm, Cholesky, W = L^-1 into the upper triangle.
  // wrow = per-row weights are rho_j on every row of a leg-step (ADMM factorisation).
  auto factor = [&](double sig) {
    double* const Fs = Wp;  // [NLS][21] staging (the triangle is rebuilt below)
    LANES_BEGIN
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int j = lane + NL * s;
      if (_(valid)[s]) {
        const double* E = _(Ev)[s];
        const double* Dd = _(Dv)[s];
        const double w = _(rhoj)[s];
        const double e0 = w * E[0] * E[0], e1 = w * E[1] * E[1], e2 = w * E[2] * E[2],
                     e3 = w * E[3] * E[3], e4 = w * E[4] * E[4];
        double dl[6];
        const double base = cscale * alpha;
        dl[0] = base * Dd[0] * Dd[0] + sig + (e0 + e1) * Dd[0] * Dd[0];
        dl[1] = 0.0;
        dl[2] = (-e0 * mu0 + e1 * mu1) * Dd[0] * Dd[2];
        dl[3] = base * Dd[1] * Dd[1] + sig + (e2 + e3) * Dd[1] * Dd[1];
        dl[4] = (-e2 * mu2 + e3 * mu3) * Dd[1] * Dd[2];
        dl[5] = base * Dd[2] * Dd[2] + sig +
                (e0 * mu0 * mu0 + e1 * mu1 * mu1 + e2 * mu2 * mu2 + e3 * mu3 * mu3 + e4) * Dd[2] * Dd[2];
        inv3sym(dl, _(di)[s]);
        // X = wbl * diag(D) (6x3), Y = X Delta^-1, F contribution = Y X^T (sym 6x6 -> 21)
        double X[6][3], Y[6][3];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
#pragma unroll
          for (int a = 0; a < 3; ++a) X[c][a] = Wb[c * WBS + 3 * (lane & 3) + a] * Dd[a];
          mul3sym(_(di)[s], X[c], Y[c]);
        }
        int e = 0;
#pragma unroll
        for (int c = 0; c < 6; ++c)
#pragma unroll
          for (int c2 = 0; c2 <= c; ++c2) {
            Fs[j * 21 + e] = Y[c][0] * X[c2][0] + Y[c][1] * X[c2][1] + Y[c][2] * X[c2][2];
            ++e;
          }
      }
    }
    LANES_END
    WSYNC();
    // F_k = sum over the 4 legs; invert (lane k < H)
    LANES_BEGIN
    if (lane < H) {
      const int k = lane;
      double F[36], Fi[36];
      int e = 0;
#pragma unroll
      for (int c = 0; c < 6; ++c)
#pragma unroll
        for (int c2 = 0; c2 <= c; ++c2) {
          const double v = Fs[(4 * k + 0) * 21 + e] + Fs[(4 * k + 1) * 21 + e] + Fs[(4 * k + 2) * 21 + e] +
                           Fs[(4 * k + 3) * 21 + e];
          F[c * 6 + c2] = v;
          F[c2 * 6 + c] = v;
          ++e;
        }
      inv6spd(F, Fi);
#pragma unroll
      for (int q = 0; q < 36; ++q) Finv[k * 36 + q] = Fi[q];
    }
    LANES_END
    WSYNC();
    // Hm = blkdiag(Finv) + 2c (Gamma (x) Q1' + N (x) Q2')  (strictly-lower rows + diagonal)
    LANES_BEGIN
    for (int i = lane; i < NW; i += NL) {
      const int k = i / 6, c = i % 6;
      if constexpr (D::FUSE_G) {  // row (k, c) of G = F_k^-1 U
        double* const grow = sm + D::OFF_G + i * D::GS;
#pragma unroll
        for (int b = 0; b < 12; ++b) {
          double g = 0;
#pragma unroll
          for (int c2 = 0; c2 < 6; ++c2) g += Finv[k * 36 + c * 6 + c2] * Wb[c2 * WBS + b];
          grow[b] = g;
        }
      }
      double* const row = Wp + tri(i);
      const double* q1r = q1s + c * 6;
      for (int k2 = 0; k2 <= k; ++k2) {
        const double g2 = 2.0 * cscale * gam[k * H + k2], n2 = 2.0 * cscale * nnm[k * H + k2] * q2s[c];
        const int cend = (k2 == k) ? c : 5;
        for (int c2 = 0; c2 <= cend; ++c2) {
          double v = g2 * q1r[c2] + (c == c2 ? n2 : 0.0);
          if (k == k2) v += Finv[k * 36 + c * 6 + c2];
          const int i2 = 6 * k2 + c2;
          if (i2 < i) row[i2] = v;
          else colb[i] = v;      // input diagonal for cholesky() (hdiag)
        }
      }
    }
    LANES_END
    WSYNC();
    cholesky(NW);
    invert_lower(NW);
  };

  // ksolve(): xt = K^-1 tmp3  (K = Delta + U^T Q U) through the wrench-space system
  auto ksolve = [&]() {
    MPC_T0();
    LANES_BEGIN
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (_(valid)[s]) {
        const int j = lane + NL * s;
        double sv[3];
        mul3sym(_(di)[s], _(tmp3)[s], sv);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          _(xt)[s][a] = sv[a];
          vecA[3 * j + a] = _(Dv)[s][a] * sv[a];
        }
      }
    }
    LANES_END
    WSYNC();
    MPC_TICK(0);
    if constexpr (D::FUSE_G) {
      LANES_BEGIN  // g = (F^-1 U) s in one stage
      for (int i = lane; i < NW; i += NL) {
        const int k = i / 6;
        const double* const grow = sm + D::OFF_G + i * D::GS;
        double g0 = 0, g1 = 0;
#pragma unroll
        for (int b = 0; b < 12; b += 2) {
          g0 += grow[b] * vecA[12 * k + b];
          g1 += grow[b + 1] * vecA[12 * k + b + 1];
        }
        vg[i] = g0 + g1;
      }
      LANES_END
      WSYNC();
    MPC_TICK(1);
    } else {
      LANES_BEGIN  // t = U s
      for (int i = lane; i < NW; i += NL) {
        const int k = i / 6, c = i % 6;
        double t = 0;
#pragma unroll
        for (int b = 0; b < 12; ++b) t += Wb[c * WBS + b] * vecA[12 * k + b];
        vt[i] = t;
      }
      LANES_END
      WSYNC();
    MPC_TICK(2);
      LANES_BEGIN  // g = F^-1 t
      for (int i = lane; i < NW; i += NL) {
        const int k = i / 6, c = i % 6;
        double g = 0;
#pragma unroll
        for (int c2 = 0; c2 < 6; ++c2) g += Finv[k * 36 + c * 6 + c2] * vt[6 * k + c2];
        vg[i] = g;
      }
      LANES_END
      WSYNC();
    MPC_TICK(3);
    }
    LANES_BEGIN  // y = W g   (row walk; rows at the same column hit distinct banks, see tri())
    for (int i = lane; i < NW; i += NL) {
      const double* rowi = Wp + tri(i);
      double a0 = dinv[i] * vg[i], a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int k = 0;
      for (; k + 3 < i; k += 4) {
        const double2 g01 = *reinterpret_cast<const double2*>(vg + k);
        const double2 g23 = *reinterpret_cast<const double2*>(vg + k + 2);
        a0 += rowi[k] * g01.x;
        a1 += rowi[k + 1] * g01.y;
        a2 += rowi[k + 2] * g23.x;
        a3 += rowi[k + 3] * g23.y;
      }
      if (k < i) {  // last 1..3 columns as one predicated body (no scalar tail loop)
        const double2 g01 = *reinterpret_cast<const double2*>(vg + k);
        a0 += rowi[k] * g01.x;
        if (k + 1 < i) a1 += rowi[k + 1] * g01.y;
        if (k + 2 < i) a2 += rowi[k + 2] * vg[k + 2];
      }
      vw[i] = (a0 + a1) + (a2 + a3);
    }
    LANES_END
    WSYNC();
    MPC_TICK(5);
    LANES_BEGIN  // w = W^T y (column walk).  All lanes of a warp visit the SAME row i at the same
    // time (lanes whose column is not yet below the diagonal are predicated off): W[i][c] is then
    // contiguous across the lanes (no bank conflicts) and y[i] is a broadcast.  A warp issues its
    // longest column's trip count anyway, so the uniform loop costs no extra steps.
    static_assert(NW % 4 == 0, "column walk takes rows four at a time");
    for (int c = lane; c < NW; c += NL) {
      double a0 = dinv[c] * vw[c], a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int i = c & ~31;                       // first column of this warp: a multiple of 4
      const double* p = Wp + tri(i) + c;     // &W[i][c] (dereferenced only where c < i)
      for (; i < NW; i += 4) {
        const double2 y01 = *reinterpret_cast<const double2*>(vw + i);
        const double2 y23 = *reinterpret_cast<const double2*>(vw + i + 2);
        if (c < i) a0 += p[0] * y01.x;
        if (c < i + 1) a1 += p[i + 1] * y01.y;
        if (c < i + 2) a2 += p[2 * i + 3] * y23.x;
        if (c < i + 3) a3 += p[3 * i + 6] * y23.y;
        p += 4 * i + 10;
      }
      vt[c] = (a0 + a1) + (a2 + a3);
    }
    LANES_END
    WSYNC();
    MPC_TICK(6);
    LANES_BEGIN  // v = g - F^-1 w  -> vecB (NW <= NV)
    for (int i = lane; i < NW; i += NL) {
      const int k = i / 6, c = i % 6;
      double v = vg[i];
#pragma unroll
      for (int c2 = 0; c2 < 6; ++c2) v -= Finv[k * 36 + c * 6 + c2] * vt[6 * k + c2];
      vecB[i] = v;
    }
    LANES_END
    WSYNC();
    MPC_TICK(7);
    LANES_BEGIN  // x = s - Delta^-1 U^T v
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (_(valid)[s]) {
        const int j = lane + NL * s;
        const int k = j >> 2;
        double u[3], du[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          double t = 0;
#pragma unroll
          for (int c = 0; c < 6; ++c) t += Wb[c * WBS + 3 * (lane & 3) + a] * vecB[6 * k + c];
          u[a] = _(Dv)[s][a] * t;
        }
        mul3sym(_(di)[s], u, du);
#pragma unroll
        for (int a = 0; a < 3; ++a) _(xt)[s][a] -= du[a];
      }
    }
    LANES_END
    WSYNC();
    MPC_TICK(8);
  };

  // apply_P(): xt = Pbar * tmp3   (structured mat-vec, used by residuals and polish)
  auto apply_P = [&]() {
    LANES_BEGIN
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (_(valid)[s]) {
        const int j = lane + NL * s;
#pragma unroll
        for (int a = 0; a < 3; ++a) vecA[3 * j + a] = _(Dv)[s][a] * _(tmp3)[s][a];
      }
    LANES_END
    WSYNC();
    LANES_BEGIN
    for (int i = lane; i < NW; i += NL) {
      const int k = i / 6, c = i % 6;
      double t = 0;
#pragma unroll
      for (int b = 0; b < 12; ++b) t += Wb[c * WBS + b] * vecA[12 * k + b];
      vt[i] = t;
    }
    LANES_END
    WSYNC();
    LANES_BEGIN
    for (int i = lane; i < NW; i += NL) {
      const int k = i / 6, c = i % 6;
      double a1 = 0;
#pragma unroll
      for (int c2 = 0; c2 < 6; ++c2) a1 += q1s[c * 6 + c2] * vt[6 * k + c2];
      vg[i] = a1;
      vw[i] = q2s[c] * vt[i];
    }
    LANES_END
    WSYNC();
    LANES_BEGIN
    for (int i = lane; i < NW; i += NL) {
      const int k = i / 6, c = i % 6;
      double o = 0;
      for (int k2 = 0; k2 < H; ++k2) o += gam[k * H + k2] * vg[6 * k2 + c] + nnm[k * H + k2] * vw[6 * k2 + c];
      vecB[i] = 2.0 * cscale * o;
    }
    LANES_END
    WSYNC();
    LANES_BEGIN
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (_(valid)[s]) {
        const int j = lane + NL * s;
        const int k = j >> 2;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          double t = 0;
#pragma unroll
          for (int c = 0; c < 6; ++c) t += Wb[c * WBS + 3 * (lane & 3) + a] * vecB[6 * k + c];
          const double d = _(Dv)[s][a];
          _(xt)[s][a] = cscale * alpha * d * d * _(tmp3)[s][a] + d * t;
        }
      }
    LANES_END
    WSYNC();
  };

  // ------------------------------------------------------------------ S3: rho, factor
  double rho = fmin(fmax(prm.rho0, kRhoMin), kRhoMax);  // auxil.c:82
  int iter = 0, status = MPC_STATUS_MAX_ITER, rho_updates = 0;
  double pri_res = 0, dua_res = 0;
  PER_LANE_ARR(double, Axv, [NS][5]);
  PER_LANE_ARR(double, Atyv, [NS][3]);
  PER_LANE_ARR(double, Pxv, [NS][3]);
  PER_LANE_ARR(double, nrm, [14]);
  if constexpr (DO_ADMM) {
  if constexpr (WARM) if (warm_valid) {
    rho = warm[1];   // settings->rho keeps the value adapted during the previous solve
    LANES_BEGIN
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (_(valid)[s]) {
        const int j = lane + NL * s;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          // osqp_update_lin_cost: q = c * (D .* q_new)   (osqp.c:752-782)
          _(qb)[s][a] = (_(Dv)[s][a] * warm[q_new_off + 3 * j + a]) * cscale;
          _(xv)[s][a] = warm[2 + 3 * j + a];
        }
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          _(zv)[s][r] = warm[2 + 12 * H + 5 * j + r];
          _(yv)[s][r] = warm[2 + 32 * H + 5 * j + r];
        }
      }
    LANES_END
  }
  LANES_BEGIN
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    // equality rows <=> swing leg (u - l < RHO_TOL on the SCALED bounds, auxil.c:90-93);
    // all five rows of a leg-step share the type unless E makes a stance row tiny
    const double* E = _(Ev)[s];
    const bool eq = (hi_of(E, _(cj)[s], 4) - lo_of(E, _(cj)[s], 4) < kRhoTol);
    _(rhoj)[s] = eq ? kRhoEqFactor * rho : rho;
  }
  LANES_END
  bool need_factor = true;   // factor() has ONE call site (top of the loop): one register allocation

  // ------------------------------------------------------------------ S4: ADMM
  const double ar = prm.relax, sigma = prm.sigma;
  const int max_iter = prm.fixed_iters > 0 ? prm.fixed_iters : prm.max_iter;
  double rinv_rho = 1.0 / rho, rinv_eq = 1.0 / (kRhoEqFactor * rho);
  int until_check = prm.check_every;
  for (iter = 1; iter <= max_iter; ++iter) {
    if (need_factor) {
      factor(prm.sigma);
      need_factor = false;
      rinv_rho = 1.0 / rho;
      rinv_eq = 1.0 / (kRhoEqFactor * rho);
    }
    MPC_T0();
    LANES_BEGIN  // rhs = sigma x - q + A^T (rho z - y)      (auxil.c:149-162)
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (_(valid)[s]) {
        double w[5], at[3];
#pragma unroll
        for (int r = 0; r < 5; ++r) w[r] = _(rhoj)[s] * _(zv)[s][r] - _(yv)[s][r];
        At_mul(_(Ev)[s], _(Dv)[s], w, at);
#pragma unroll
        for (int a = 0; a < 3; ++a) _(tmp3)[s][a] = sigma * _(xv)[s][a] - _(qb)[s][a] + at[a];
      }
    }
    LANES_END
    MPC_TICK(12);
    ksolve();
    { MPC_T0();
    LANES_BEGIN  // update_x, update_z, update_y (auxil.c:188-228)
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (_(valid)[s]) {
        double zt[5];
        A_mul(_(Ev)[s], _(Dv)[s], _(xt)[s], zt);
        const double rinv = (_(rhoj)[s] == rho) ? rinv_rho : rinv_eq;   // 1 / rho_j without the division
#pragma unroll
        for (int a = 0; a < 3; ++a) _(xv)[s][a] = ar * _(xt)[s][a] + (1.0 - ar) * _(xv)[s][a];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const double zr = ar * zt[r] + (1.0 - ar) * _(zv)[s][r];
          const double lo = lo_of(_(Ev)[s], _(cj)[s], r), hi = hi_of(_(Ev)[s], _(cj)[s], r);
          const double zn = clampd(zr + rinv * _(yv)[s][r], lo, hi);
          _(yv)[s][r] += _(rhoj)[s] * (zr - zn);
          _(zv)[s][r] = zn;
        }
      }
    }
    LANES_END
    MPC_TICK(13); }
    // check_termination every check_every iterations (a countdown instead of iter % check_every)
    const bool check = prm.fixed_iters > 0 ? (iter == max_iter) : (--until_check == 0);
    if (!check) continue;
    until_check = prm.check_every;
    // ---- residuals (auxil.c:243-362), unscaled norms for termination, scaled for rho
    LANES_BEGIN
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int a = 0; a < 3; ++a) _(tmp3)[s][a] = _(xv)[s][a];
    LANES_END
    apply_P();
    LANES_BEGIN
#pragma unroll
    for (int q = 0; q < 14; ++q) _(nrm)[q] = 0.0;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (_(valid)[s]) {
        A_mul(_(Ev)[s], _(Dv)[s], _(xv)[s], _(Axv)[s]);
        At_mul(_(Ev)[s], _(Dv)[s], _(yv)[s], _(Atyv)[s]);
#pragma unroll
        for (int a = 0; a < 3; ++a) _(Pxv)[s][a] = _(xt)[s][a];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const double ei = 1.0 / _(Ev)[s][r];
          const double rp_ = _(Axv)[s][r] - _(zv)[s][r];
          _(nrm)[0] = maxd(_(nrm)[0], fabs(ei * rp_));
          _(nrm)[2] = maxd(_(nrm)[2], fabs(ei * _(zv)[s][r]));
          _(nrm)[3] = maxd(_(nrm)[3], fabs(ei * _(Axv)[s][r]));
          _(nrm)[7] = maxd(_(nrm)[7], fabs(rp_));
          _(nrm)[8] = maxd(_(nrm)[8], fabs(_(zv)[s][r]));
          _(nrm)[9] = maxd(_(nrm)[9], fabs(_(Axv)[s][r]));
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const double dii = 1.0 / _(Dv)[s][a];
          const double rd = _(qb)[s][a] + _(Pxv)[s][a] + _(Atyv)[s][a];
          _(nrm)[1] = maxd(_(nrm)[1], fabs(dii * rd));
          _(nrm)[4] = maxd(_(nrm)[4], fabs(dii * _(qb)[s][a]));
          _(nrm)[5] = maxd(_(nrm)[5], fabs(dii * _(Atyv)[s][a]));
          _(nrm)[6] = maxd(_(nrm)[6], fabs(dii * _(Pxv)[s][a]));
          _(nrm)[10] = maxd(_(nrm)[10], fabs(rd));
          _(nrm)[11] = maxd(_(nrm)[11], fabs(_(qb)[s][a]));
          _(nrm)[12] = maxd(_(nrm)[12], fabs(_(Atyv)[s][a]));
          _(nrm)[13] = maxd(_(nrm)[13], fabs(_(Pxv)[s][a]));
        }
      }
    }
    LANES_END
    double nr_[14];
    WMAXV(nr_, nrm);
    pri_res = nr_[0];
    dua_res = cinv * nr_[1];
    if (!(pri_res <= 1e30) || !(dua_res <= 1e30)) {   // diverged / NaN: OSQP_NON_CVX (auxil.c:697-704)
      status = MPC_STATUS_NON_FINITE;
      break;
    }
    if (prm.fixed_iters <= 0) {
      const double eps_prim = prm.eps_abs + prm.eps_rel * fmax(nr_[2], nr_[3]);
      const double eps_dual = prm.eps_abs + prm.eps_rel * cinv * fmax(fmax(nr_[4], nr_[5]), nr_[6]);
      if (pri_res < eps_prim && dua_res < eps_dual) {  // auxil.c:735-747
        status = MPC_STATUS_SOLVED;
        break;
      }
    } else {
      status = MPC_STATUS_SOLVED;
      break;
    }
    if (prm.adaptive_rho && iter < max_iter) {  // compute_rho_estimate / adapt_rho (auxil.c:13-77)
      const double pr = nr_[7] / (fmax(nr_[8], nr_[9]) + 1e-10);
      const double dr = nr_[10] / (fmax(fmax(nr_[11], nr_[12]), nr_[13]) + 1e-10);
      double rho_new = rho * sqrt(pr / (dr + 1e-10));
      rho_new = fmin(fmax(rho_new, kRhoMin), kRhoMax);
      if (rho_new > rho * prm.rho_tol_ratio || rho_new < rho / prm.rho_tol_ratio) {
        rho = rho_new;
        ++rho_updates;
        LANES_BEGIN
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const double* E = _(Ev)[s];
          const bool eq = (hi_of(E, _(cj)[s], 4) - lo_of(E, _(cj)[s], 4) < kRhoTol);
          _(rhoj)[s] = eq ? kRhoEqFactor * rho : rho;
        }
        LANES_END
        need_factor = true;
      }
    }
  }
  if (iter > max_iter) iter = max_iter;
  }  // DO_ADMM
  if constexpr (DO_ADMM && !DO_POLISH) {
    if (prm.polish && status == MPC_STATUS_SOLVED) {   // the polish kernel finishes this robot
      LANES_BEGIN
#pragma unroll
      for (int s = 0; s < NS; ++s)
        if (_(valid)[s]) {
          const int j = lane + NL * s;
#pragma unroll
          for (int a = 0; a < 3; ++a) ws[WL::OFF_XZY + a * D::NLS + j] = _(xv)[s][a];
#pragma unroll
          for (int r = 0; r < 5; ++r) {
            ws[WL::OFF_XZY + (3 + r) * D::NLS + j] = _(zv)[s][r];
            ws[WL::OFF_XZY + (8 + r) * D::NLS + j] = _(yv)[s][r];
          }
        }
      if (lane == 0) {
        ws[WS_STATUS] = status;
        ws[WS_ITER] = iter;
        ws[WS_RHO] = rho;
        ws[WS_RHO_UPDATES] = rho_updates;
        ws[WS_PRI] = pri_res;
        ws[WS_DUA] = dua_res;
        ws[WS_STAGE] = kWsStagePolish;
      }
      LANES_END
      return;
    }
    LANES_BEGIN
    if (lane == 0) ws[WS_STAGE] = kWsStageDone;
    LANES_END
  }
  if constexpr (!DO_ADMM) {
    if constexpr (WARM) if (warm_valid) {   // the ADMM kernel installed the NEW q (osqp_update_lin_cost); so must we
      LANES_BEGIN
#pragma unroll
      for (int s = 0; s < NS; ++s)
        if (_(valid)[s]) {
          const int j = lane + NL * s;
#pragma unroll
          for (int a = 0; a < 3; ++a) _(qb)[s][a] = (_(Dv)[s][a] * warm[q_new_off + 3 * j + a]) * cscale;
        }
      LANES_END
    }
    status = (int)ws[WS_STATUS];
    iter = (int)ws[WS_ITER];
    rho = ws[WS_RHO];
    rho_updates = (int)ws[WS_RHO_UPDATES];
    pri_res = ws[WS_PRI];
    dua_res = ws[WS_DUA];
    LANES_BEGIN
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (_(valid)[s]) {
        const int j = lane + NL * s;
#pragma unroll
        for (int a = 0; a < 3; ++a) _(xv)[s][a] = ws[WL::OFF_XZY + a * D::NLS + j];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          _(zv)[s][r] = ws[WL::OFF_XZY + (3 + r) * D::NLS + j];
          _(yv)[s][r] = ws[WL::OFF_XZY + (8 + r) * D::NLS + j];
        }
      }
    LANES_END
  }

  // ------------------------------------------------------------------ S5: polish
  int polish_status = 0, nfree_total = 0;
  const bool polish_due = DO_POLISH && prm.polish && status == MPC_STATUS_SOLVED;
  if (polish_due) {
    // polish.c:20-95 active-set guess; per leg-step null-space basis of the active rows
    PER_LANE_ARR(double, Yb, [NS][3][3]);   // orthonormal basis: first rk = row space, rest = null space
    PER_LANE_ARR(double, pp, [NS][3]);      // particular (pinned) part of x
    PER_LANE_ARR(int, rk, [NS]);
    PER_LANE_ARR(int, act, [NS]);           // bit r: row active; bit 8+r: upper (else lower)
    PER_LANE_ARR(int, foff, [NS]);
    PER_LANE(int, cnt);
    PER_LANE(int, off);
    LANES_BEGIN
    _(cnt) = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      _(rk)[s] = 0;
      _(act)[s] = 0;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        _(pp)[s][a] = 0;
#pragma unroll
        for (int b = 0; b < 3; ++b) _(Yb)[s][a][b] = (a == b) ? 1.0 : 0.0;
      }
      if (_(valid)[s]) {
        const double* E = _(Ev)[s];
        const double* Dd = _(Dv)[s];
        double rows[5][3], bt[5];
        int na = 0;
        const double arow[5][3] = {{-E[0] * Dd[0], 0.0, E[0] * mu0 * Dd[2]},
                                   {E[1] * Dd[0], 0.0, E[1] * mu1 * Dd[2]},
                                   {0.0, -E[2] * Dd[1], E[2] * mu2 * Dd[2]},
                                   {0.0, E[3] * Dd[1], E[3] * mu3 * Dd[2]},
                                   {0.0, 0.0, E[4] * Dd[2]}};
        double rmax = 0;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const double lo = lo_of(E, _(cj)[s], r), hi = hi_of(E, _(cj)[s], r);
          const bool low = (_(zv)[s][r] - lo < -_(yv)[s][r]);
          const bool upp = (hi - _(zv)[s][r] < _(yv)[s][r]);
          if (low || upp) {
            _(act)[s] |= (1 << r) | ((low ? 0 : 1) << (8 + r));
            rows[na][0] = arow[r][0]; rows[na][1] = arow[r][1]; rows[na][2] = arow[r][2];
            bt[na] = low ? lo : hi;
            rmax = maxd(rmax, arow[r][0] * arow[r][0] + arow[r][1] * arow[r][1] + arow[r][2] * arow[r][2]);  // squared
            ++na;
          }
        }
        // modified Gram-Schmidt over the active rows -> orthonormal row-space basis
        double Yr[3][3];
        int rank = 0;
        for (int q = 0; q < na && rank < 3; ++q) {
          double v[3] = {rows[q][0], rows[q][1], rows[q][2]};
          for (int p = 0; p < rank; ++p) {
            const double d = v[0] * Yr[p][0] + v[1] * Yr[p][1] + v[2] * Yr[p][2];
            v[0] -= d * Yr[p][0]; v[1] -= d * Yr[p][1]; v[2] -= d * Yr[p][2];
          }
          const double nv2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
          if (nv2 > 1e-18 * rmax) {   // |v| > 1e-9 * max row norm
            const double inv = MPC_RSQRT(nv2);
            Yr[rank][0] = v[0] * inv; Yr[rank][1] = v[1] * inv; Yr[rank][2] = v[2] * inv;
            ++rank;
          }
        }
        // complete to an orthonormal basis of R^3 (null space of the active rows)
        if (rank == 0) {
          // identity already
        } else if (rank == 1) {
          const double* y = Yr[0];
          int ax = 0;
          if (fabs(y[1]) < fabs(y[ax])) ax = 1;
          if (fabs(y[2]) < fabs(y[ax])) ax = 2;
          double e[3] = {0, 0, 0};
          e[ax] = 1.0;
          double n1[3] = {y[1] * e[2] - y[2] * e[1], y[2] * e[0] - y[0] * e[2], y[0] * e[1] - y[1] * e[0]};
          const double l1 = MPC_RSQRT(n1[0] * n1[0] + n1[1] * n1[1] + n1[2] * n1[2]);
          n1[0] *= l1; n1[1] *= l1; n1[2] *= l1;
          const double n2[3] = {y[1] * n1[2] - y[2] * n1[1], y[2] * n1[0] - y[0] * n1[2], y[0] * n1[1] - y[1] * n1[0]};
#pragma unroll
          for (int a = 0; a < 3; ++a) { _(Yb)[s][0][a] = y[a]; _(Yb)[s][1][a] = n1[a]; _(Yb)[s][2][a] = n2[a]; }
        } else if (rank == 2) {
          const double* y1 = Yr[0];
          const double* y2 = Yr[1];
          const double n1[3] = {y1[1] * y2[2] - y1[2] * y2[1], y1[2] * y2[0] - y1[0] * y2[2], y1[0] * y2[1] - y1[1] * y2[0]};
#pragma unroll
          for (int a = 0; a < 3; ++a) { _(Yb)[s][0][a] = y1[a]; _(Yb)[s][1][a] = y2[a]; _(Yb)[s][2][a] = n1[a]; }
        } else {
#pragma unroll
          for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int a = 0; a < 3; ++a) _(Yb)[s][p][a] = Yr[p][a];
        }
        _(rk)[s] = rank;
        // particular solution: least squares of rows*(Y eta) = bt in the row-space basis
        if (rank > 0) {
          double BtB[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Btb[3] = {0, 0, 0};
          for (int q = 0; q < na; ++q) {
            double bq[3];
            for (int p = 0; p < rank; ++p)
              bq[p] = rows[q][0] * _(Yb)[s][p][0] + rows[q][1] * _(Yb)[s][p][1] + rows[q][2] * _(Yb)[s][p][2];
            for (int p = 0; p < rank; ++p) {
              Btb[p] += bq[p] * bt[q];
              for (int p2 = 0; p2 < rank; ++p2) BtB[p][p2] += bq[p] * bq[p2];
            }
          }
          for (int p = rank; p < 3; ++p) BtB[p][p] = 1.0;
          const double sym[6] = {BtB[0][0], BtB[0][1], BtB[0][2], BtB[1][1], BtB[1][2], BtB[2][2]};
          double isym[6], eta[3];
          inv3sym(sym, isym);
          mul3sym(isym, Btb, eta);
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            double v = 0;
            for (int p = 0; p < rank; ++p) v += _(Yb)[s][p][a] * eta[p];
            _(pp)[s][a] = v;
          }
        }
        _(cnt) += 3 - rank;
      }
    }
    LANES_END
    WEXSCAN_I(off, cnt);
    nfree_total = WSUM_I(cnt);
    if (nfree_total <= PCAP) {
      const int nf = nfree_total;
      double* const Mf = sm + D::OFF_SCR;            // [nf][6]  m_f = U nu_f
      double* const A1f = Mf + 6 * PCAP;             // [nf][6]  Q1' m_f
      int* const kf = reinterpret_cast<int*>(A1f + 6 * PCAP);  // [nf] step index
      // gP = Pbar p + qbar  (needs apply_P before the scratch is overwritten)
      LANES_BEGIN
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int a = 0; a < 3; ++a) _(tmp3)[s][a] = _(pp)[s][a];
      LANES_END
      apply_P();
      PER_LANE_ARR(double, bxi, [NS][3]);   // -N^T (q + P p) per free coordinate of the slot
      LANES_BEGIN
      int o = _(off);
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        _(foff)[s] = o;
        if (_(valid)[s]) {
          const int j = lane + NL * s;
          const int k = j >> 2;
          for (int p = _(rk)[s]; p < 3; ++p) {
            const double* nu = _(Yb)[s][p];
            double gdot = 0;
#pragma unroll
            for (int a = 0; a < 3; ++a) gdot += nu[a] * (_(qb)[s][a] + _(xt)[s][a]);
            _(bxi)[s][p] = -gdot;
          }
          o += 3 - _(rk)[s];
          (void)k;
        }
      }
      LANES_END
      WSYNC();
      LANES_BEGIN  // m_f, Q1' m_f into scratch (scratch == vecA.. region, apply_P is done)
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (_(valid)[s]) {
          const int j = lane + NL * s;
          const int k = j >> 2;
          int f = _(foff)[s];
          for (int p = _(rk)[s]; p < 3; ++p, ++f) {
            const double* nu = _(Yb)[s][p];
            double m[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              m[c] = 0;
#pragma unroll
              for (int a = 0; a < 3; ++a) m[c] += Wb[c * WBS + 3 * (lane & 3) + a] * _(Dv)[s][a] * nu[a];
              Mf[6 * f + c] = m[c];
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              double a1 = 0;
#pragma unroll
              for (int c2 = 0; c2 < 6; ++c2) a1 += q1s[c * 6 + c2] * m[c2];
              A1f[6 * f + c] = a1;
            }
            kf[f] = k;
          }
        }
      }
      LANES_END
      WSYNC();
      // M = K_f + delta*I: wrench part (strictly-lower rows in Wp, diagonal in dinv)
      LANES_BEGIN
      for (int f = lane; f < nf; f += NL) {
        double m[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) m[c] = Mf[6 * f + c];
        const int k = kf[f];
        double* const row = Wp + tri(f);
        for (int f2 = 0; f2 <= f; ++f2) {
          const int k2 = kf[f2];
          double d1 = 0, d2 = 0;
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            d1 += m[c] * A1f[6 * f2 + c];
            d2 += m[c] * q2s[c] * Mf[6 * f2 + c];
          }
          const double v = 2.0 * cscale * (gam[k * H + k2] * d1 + nnm[k * H + k2] * d2);
          if (f2 == f) colb[f] = v + prm.delta;   // hdiag of cholesky()
          else row[f2] = v;
        }
      }
      LANES_END
      WSYNC();
      // same-leg-step term nu^T diag(c alpha D^2) nu' added by the owner lane
      LANES_BEGIN
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (_(valid)[s]) {
          const int f0 = _(foff)[s];
          for (int p = _(rk)[s]; p < 3; ++p)
            for (int p2 = _(rk)[s]; p2 <= p; ++p2) {
              double v = 0;
#pragma unroll
              for (int a = 0; a < 3; ++a)
                v += _(Yb)[s][p][a] * cscale * alpha * _(Dv)[s][a] * _(Dv)[s][a] * _(Yb)[s][p2][a];
              const int f = f0 + (p - _(rk)[s]), f2 = f0 + (p2 - _(rk)[s]);
              if (f == f2) colb[f] += v;
              else Wp[tri(f) + f2] += v;
            }
        }
      }
      LANES_END
      WSYNC();
      // pad to a multiple of the factorisation's block width with identity rows
      const int nfp = ((nf + CB - 1) / CB) * CB;
      LANES_BEGIN
      for (int f = nf + lane; f < nfp; f += NL) {
        double* const row = Wp + tri(f);
        for (int f2 = 0; f2 < f; ++f2) row[f2] = 0.0;
        colb[f] = 1.0;
      }
      LANES_END
      WSYNC();
      cholesky(nfp);
      // polish vectors: Dsm is free here (D is also held in registers), colb holds xi_0
      double* const pv = Dsm;        // [nf] right-hand side / solution of the triangular solves
      double* const pxi = Dsm + PCAP;  // [nf] current xi
      double* const pxi0 = colb;     // [nf] first regularised solve (colb is free again after invert_lower)
      invert_lower(nfp);            // W = L^-1: every solve below is two barrier-light mat-vecs
      auto tri_solve = [&]() {      // pv <- W^T (W pv) = (L L^T)^-1 pv
        double* const ty = colb + PCAP;
        LANES_BEGIN
        for (int i = lane; i < nf; i += NL) {
          const double* rowi = Wp + tri(i);
          double a0 = dinv[i] * pv[i], a1 = 0.0;
          int k = 0;
          for (; k + 1 < i; k += 2) {
            a0 += rowi[k] * pv[k];
            a1 += rowi[k + 1] * pv[k + 1];
          }
          if (k < i) a0 += rowi[k] * pv[k];
          ty[i] = a0 + a1;
        }
        LANES_END
        WSYNC();
        LANES_BEGIN
        for (int c = lane; c < nf; c += NL) {
          double a0 = dinv[c] * ty[c];
          const double* p = Wp + tri(c + 1) + c;
          for (int i = c + 1; i < nf; ++i) {
            a0 += p[0] * ty[i];
            p += i + 1;
          }
          pv[c] = a0;
        }
        LANES_END
        WSYNC();
      };
      // first solve: (K_f + delta) xi_0 = b
      LANES_BEGIN
#pragma unroll
      for (int s = 0; s < NS; ++s)
        if (_(valid)[s])
          for (int p = _(rk)[s]; p < 3; ++p) pv[_(foff)[s] + (p - _(rk)[s])] = _(bxi)[s][p];
      LANES_END
      WSYNC();
      tri_solve();
      LANES_BEGIN
      for (int f = lane; f < nf; f += NL) {
        pxi[f] = pv[f];
        pxi0[f] = pv[f];
      }
      LANES_END
      WSYNC();
      // iterative refinement (polish.c:127-175): xi <- xi + M^-1 (b - K_f xi) with M = K_f + delta I
      // which is algebraically xi <- xi_0 + delta * M^-1 xi (no copy of K_f needed)
      for (int itr = 0; itr < prm.polish_refine; ++itr) {
        LANES_BEGIN
        for (int f = lane; f < nf; f += NL) pv[f] = pxi[f];
        LANES_END
        WSYNC();
        tri_solve();
        LANES_BEGIN
        for (int f = lane; f < nf; f += NL) pxi[f] = pxi0[f] + prm.delta * pv[f];
        LANES_END
        WSYNC();
      }
      // x_pol = p + N xi ; g = Pbar x_pol + qbar ; multipliers ; acceptance (polish.c:300-330)
      PER_LANE_ARR(double, xp, [NS][3]);
      LANES_BEGIN
#pragma unroll
      for (int s = 0; s < NS; ++s) {
#pragma unroll
        for (int a = 0; a < 3; ++a) _(xp)[s][a] = _(pp)[s][a];
        if (_(valid)[s])
          for (int p = _(rk)[s]; p < 3; ++p) {
            const double x_ = pxi[_(foff)[s] + (p - _(rk)[s])];
#pragma unroll
            for (int a = 0; a < 3; ++a) _(xp)[s][a] += _(Yb)[s][p][a] * x_;
          }
#pragma unroll
        for (int a = 0; a < 3; ++a) _(tmp3)[s][a] = _(xp)[s][a];
      }
      LANES_END
      WSYNC();
      // restore D in shared memory (Dsm was used as polish vector space) before apply_P
      LANES_BEGIN
#pragma unroll
      for (int s = 0; s < NS; ++s)
        if (_(valid)[s])
#pragma unroll
          for (int a = 0; a < 3; ++a) Dsm[3 * (lane + NL * s) + a] = _(Dv)[s][a];
      LANES_END
      WSYNC();
      apply_P();
      PER_LANE_ARR(double, ypv, [NS][5]);
      LANES_BEGIN
      double npri = 0, ndua = 0;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (_(valid)[s]) {
          const double* E = _(Ev)[s];
          const double* Dd = _(Dv)[s];
          const double arow[5][3] = {{-E[0] * Dd[0], 0.0, E[0] * mu0 * Dd[2]},
                                     {E[1] * Dd[0], 0.0, E[1] * mu1 * Dd[2]},
                                     {0.0, -E[2] * Dd[1], E[2] * mu2 * Dd[2]},
                                     {0.0, E[3] * Dd[1], E[3] * mu3 * Dd[2]},
                                     {0.0, 0.0, E[4] * Dd[2]}};
          double g[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) g[a] = _(qb)[s][a] + _(xt)[s][a];
          // y_act = B (B^T B)^-1 (-Y^T g), B = A_act Y (rank columns)
          const int rank = _(rk)[s];
          double yred[5] = {0, 0, 0, 0, 0};
          if (rank > 0) {
            double BtB[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, rhs3[3] = {0, 0, 0};
            double Bq[5][3];
#pragma unroll
            for (int r = 0; r < 5; ++r) {
              for (int p = 0; p < 3; ++p) Bq[r][p] = 0;
              if (_(act)[s] & (1 << r)) {
                for (int p = 0; p < rank; ++p) {
                  Bq[r][p] = arow[r][0] * _(Yb)[s][p][0] + arow[r][1] * _(Yb)[s][p][1] + arow[r][2] * _(Yb)[s][p][2];
                }
                for (int p = 0; p < rank; ++p)
                  for (int p2 = 0; p2 < rank; ++p2) BtB[p][p2] += Bq[r][p] * Bq[r][p2];
              }
            }
            for (int p = 0; p < rank; ++p)
              rhs3[p] = -(_(Yb)[s][p][0] * g[0] + _(Yb)[s][p][1] * g[1] + _(Yb)[s][p][2] * g[2]);
            for (int p = rank; p < 3; ++p) BtB[p][p] = 1.0;
            const double sym[6] = {BtB[0][0], BtB[0][1], BtB[0][2], BtB[1][1], BtB[1][2], BtB[2][2]};
            double isym[6], tt[3];
            inv3sym(sym, isym);
            mul3sym(isym, rhs3, tt);
#pragma unroll
            for (int r = 0; r < 5; ++r)
              if (_(act)[s] & (1 << r)) yred[r] = Bq[r][0] * tt[0] + Bq[r][1] * tt[1] + Bq[r][2] * tt[2];
          }
          double zp[5];
          A_mul(E, Dd, _(xp)[s], zp);
          double yp[5];
#pragma unroll
          for (int r = 0; r < 5; ++r) {  // project_normalcone (proj.c:17-31)
            const double lo = lo_of(E, _(cj)[s], r), hi = hi_of(E, _(cj)[s], r);
            const double t = zp[r] + yred[r];
            const double zc = fmin(fmax(t, lo), hi);
            yp[r] = t - zc;
            _(ypv)[s][r] = yp[r];
            npri = fmax(npri, fabs((zp[r] - zc) / E[r]));
            zp[r] = zc;
          }
          double aty[3];
          At_mul(E, Dd, yp, aty);
#pragma unroll
          for (int a = 0; a < 3; ++a) ndua = fmax(ndua, fabs((g[a] + aty[a]) / Dd[a]));
#pragma unroll
          for (int r = 0; r < 5; ++r) _(Axv)[s][r] = zp[r];  // polished z kept for warm start
        }
      }
      _(red) = npri;
      _(red2) = ndua;
      LANES_END
      const double pri_p = WMAX(red), dua_p = cinv * WMAX(red2);
      const bool ok = (pri_p < pri_res && dua_p < dua_res) || (pri_p < pri_res && dua_res < 1e-10) ||
                      (dua_p < dua_res && pri_res < 1e-10);
      if (ok) {
        polish_status = 1;
        pri_res = pri_p;
        dua_res = dua_p;
        LANES_BEGIN
#pragma unroll
        for (int s = 0; s < NS; ++s) {
#pragma unroll
          for (int a = 0; a < 3; ++a) _(xv)[s][a] = _(xp)[s][a];
#pragma unroll
          for (int r = 0; r < 5; ++r) {
            _(zv)[s][r] = _(Axv)[s][r];
            _(yv)[s][r] = _(ypv)[s][r];
          }
        }
        LANES_END
      } else {
        polish_status = -1;
      }
    } else {
      polish_status = -1;  // free set larger than this instantiation's polish capacity
      status = MPC_STATUS_POLISH_OVERFLOW;   // the overflow kernel (PCAP = 12H) re-solves this robot
    }
  }

  // ------------------------------------------------------------------ S6: output
  // NaN / Inf inputs slip through OSQP's inf-norms (comparisons with NaN are false) and would
  // come back as NaN forces; report them instead
  LANES_BEGIN
  double bad = 0.0;
#pragma unroll
  for (int s = 0; s < NS; ++s)
    if (_(valid)[s])
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const double v = _(Dv)[s][a] * _(xv)[s][a];
        if (!(fabs(v) <= 1e30)) bad = 1.0;
      }
  _(red) = bad;
  LANES_END
  if (WMAX(red) != 0.0) status = MPC_STATUS_NON_FINITE;
  // unscale (scaling.c:183-199) and negate (mpc_osqp.cc:789-790)
  LANES_BEGIN
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    if (_(valid)[s]) {
      const int j = lane + NL * s;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        // no solution (OSQP_NON_CVX): the reference returns an empty vector; we return zeros
        const float f = status == MPC_STATUS_NON_FINITE ? 0.0f : (float)(-_(Dv)[s][a] * _(xv)[s][a]);
        if (j < 4) grf0[3 * j + a] = f;
        if (grf_full) grf_full[3 * j + a] = f;
      }
    }
  }
  if constexpr (WARM) if (warm && status != MPC_STATUS_POLISH_OVERFLOW && status != MPC_STATUS_NON_FINITE) {
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (_(valid)[s]) {
        const int j = lane + NL * s;
#pragma unroll
        for (int a = 0; a < 3; ++a) warm[2 + 3 * j + a] = _(xv)[s][a];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          warm[2 + 12 * H + 5 * j + r] = _(zv)[s][r];
          warm[2 + 32 * H + 5 * j + r] = _(yv)[s][r];
        }
      }
    if (lane == 0) {
      warm[0] = (q_new_off == 2 + 52 * H) ? 1.0 : 2.0;   // the slot just written becomes "previous"
      warm[1] = rho;
    }
  }
  if (lane == 0) {
    if (status_out) *status_out = status;
    if (resid_out) {
      resid_out[0] = (float)pri_res;
      resid_out[1] = (float)dua_res;
    }
    if (info_out) {
      info_out[INFO_ITERS] = iter;
      info_out[INFO_RHO_UPDATES] = rho_updates;
      info_out[INFO_RHO] = rho;
      info_out[INFO_POLISH] = polish_status;
      info_out[INFO_PRI] = pri_res;
      info_out[INFO_DUA] = dua_res;
      info_out[INFO_C] = cscale;
      info_out[INFO_NFREE] = nfree_total;
    }
  }
  LANES_END
  WSYNC();
}

}  // namespace mpc
