// emu_main.cc — GPU-less development/CI harness (TEST INFRASTRUCTURE, never shipped).
//
// Compiles the warp program of rl_mpc_locomotion_b200/csrc/mpc_warp_solver.cuh with
// -DMPC_EMULATE (32 lanes executed sequentially, see warp_prog.h) so that the kernel LOGIC
// can be checked against the oracle on a machine without a GPU.  Nothing in the package,
// bench.py or the C-ABI library links or loads this file.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../rl_mpc_locomotion_b200/csrc/mpc_warp_solver.cuh"
#include "../../rl_mpc_locomotion_b200/csrc/mpc_tick.cuh"

template <int H>
static void run(const float* rec, int n, const float* rp, const int* rt, const mpc::SolveParams& prm,
                float* grf0, float* grf_full, int32_t* status, double* info, double* dbg, double* warm) {
  using D = mpc::Dims<H>;
  using DB = mpc::Dims<H, 12 * H>;
  std::vector<double> sm(D::SMEM_DOUBLES, 0.0), smb(DB::SMEM_DOUBLES, 0.0);
  mpc::fill_tables<H>(sm.data(), prm.dt);
  mpc::fill_tables<H, 12 * H>(smb.data(), prm.dt);
  const int stride = MPC_REC_STRIDE(H);
  const int dbg_stride = D::NV + D::NV * D::NV + D::NV + 20 * H + 1;
  for (int r = 0; r < n; ++r) {
    const float* rpr = rp + (size_t)(rt ? rt[r] : 0) * MPC_ROBOT_PARAMS_STRIDE;
    int32_t st = 0;
    mpc::solve_robot<H, true>(rec + (size_t)r * stride, rpr, prm, sm.data(), grf0 + (size_t)r * 12,
                        grf_full ? grf_full + (size_t)r * 12 * H : nullptr, &st,
                        nullptr, info ? info + (size_t)r * mpc::INFO_STRIDE : nullptr,
                        dbg ? dbg + (size_t)r * dbg_stride : nullptr,
                        warm ? warm + (size_t)r * MPC_WARM_STRIDE(H) : nullptr);
    if (st == MPC_STATUS_POLISH_OVERFLOW)   // the overflow kernel's job on the device
      mpc::solve_robot<H, true, 12 * H>(rec + (size_t)r * stride, rpr, prm, smb.data(), grf0 + (size_t)r * 12,
                          grf_full ? grf_full + (size_t)r * 12 * H : nullptr, &st,
                          nullptr, info ? info + (size_t)r * mpc::INFO_STRIDE : nullptr,
                          dbg ? dbg + (size_t)r * dbg_stride : nullptr,
                          warm ? warm + (size_t)r * MPC_WARM_STRIDE(H) : nullptr);
    if (status) status[r] = st;
  }
}

extern "C" int mpc_emu_solve(const float* rec, int n, int horizon, const float* rp, const int* rt,
                             const mpc::SolveParams* prm, float* grf0, float* grf_full,
                             int32_t* status, double* info, double* dbg, double* warm) {
  if (horizon == 10) run<10>(rec, n, rp, rt, *prm, grf0, grf_full, status, info, dbg, warm);
  else if (horizon == 16) run<16>(rec, n, rp, rt, *prm, grf0, grf_full, status, info, dbg, warm);
  else return -1;
  return 0;
}

// One emulated control tick for n robots (same sequence as mpc_tick_kernel in mpc_b200.cu).
extern "C" int mpc_emu_tick(const mpc::SolveParams* prm_in, const mpc::TickParams* tp, const float* dof,
                            const float* body, const float* est, const float* cmd, int has_weights, const int* gait,
                            const float* rp, const float* geom, const int* rt, int n, float* state,
                            float* torques, float* grf0, float* records, int32_t* solved, double* warm) {
  constexpr int H = 10;
  using D = mpc::Dims<H>;
  mpc::SolveParams prm = *prm_in;
  prm.dt = (double)tp->dt * tp->iters_between_mpc;
  std::vector<double> sm(D::SMEM_DOUBLES, 0.0);
  const int stride = MPC_REC_STRIDE(H);
  for (int r = 0; r < n; ++r) {
    const int type = rt ? rt[r] : 0;
    float rec[stride];
    std::memset(rec, 0, sizeof(rec));
    float legs[4 * mpc::LG_STRIDE];
    std::memset(legs, 0, sizeof(legs));
    float grf[12] = {0};
    int32_t st_code = 0;
    float* st = state + (size_t)r * MPC_TICK_STATE_STRIDE;
    const float* g = geom + (size_t)type * MPC_GEOM_STRIDE;
    float est_local[MPC_EST_STRIDE];
    const float* e = est ? est + (size_t)r * MPC_EST_STRIDE : est_local;
    float rp16[2] = {0, 0};
    if (!est) mpc::estimator_update(body + (size_t)r * 13, st + MPC_ST_NORMAL, est_local, rp16);
    const int do_solve = mpc::tick_pre<H>(dof + (size_t)r * 24, e, cmd + (size_t)r * MPC_CMD_STRIDE, has_weights,
                                          gait ? gait[r] : 0, g, *tp, st, rec, legs);
    if (!est) reinterpret_cast<int*>(st)[MPC_ST_UNSAFE] = mpc::locomotion_unsafe(rp16, legs + mpc::LG_P, mpc::LG_STRIDE);
    if (records) std::memcpy(records + (size_t)r * stride, rec, sizeof(rec));
    if (do_solve) {
      mpc::fill_tables<H>(sm.data(), prm.dt);
      double info[mpc::INFO_STRIDE] = {0};
      mpc::solve_robot<H, true>(rec, rp + (size_t)type * MPC_ROBOT_PARAMS_STRIDE, prm, sm.data(), grf, nullptr, &st_code,
                          nullptr, info, nullptr, warm ? warm + (size_t)r * MPC_WARM_STRIDE(H) : nullptr);
      reinterpret_cast<int*>(st)[MPC_ST_ITERS] = (int)info[mpc::INFO_ITERS];
      if (grf0) std::memcpy(grf0 + (size_t)r * 12, grf, sizeof(grf));
    }
    mpc::tick_post(e, g, (do_solve && st_code == MPC_STATUS_SOLVED) ? grf : nullptr, st, legs, torques + (size_t)r * 12);
    if (solved) solved[r] = st_code;
  }
  return 0;
}

extern "C" void mpc_emu_estimator(const float* body, const float* normal, float* est, int n) {
  for (int r = 0; r < n; ++r) mpc::estimator_update(body + 13 * r, normal + 3 * r, est + MPC_EST_STRIDE * r);
}

// Direct access to the two dense factorisation lane programs for tests/test_warp_emu.py:
// A (n x n SPD, row-major, n a multiple of 6, n <= pcap in {60, 120}) -> W = chol(A)^-1 (lower, row-major).
template <int PCAP>
static void chol_inv_t(const double* A, int n, double* W) {
  std::vector<double> Wp((size_t)PCAP * (PCAP + 1) / 2 + 8, 0.0), colb(2 * PCAP, 0.0), dinv(PCAP, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < i; ++k) Wp[mpc::tri(i) + k] = A[(size_t)i * n + k];
    colb[i] = A[(size_t)i * n + i];
  }
  mpc::cholesky_blocked<PCAP>(Wp.data(), colb.data(), dinv.data(), n);
  mpc::invert_lower_blocked<PCAP>(Wp.data(), dinv.data(), n);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k) W[(size_t)i * n + k] = k < i ? Wp[mpc::tri(i) + k] : (k == i ? dinv[i] : 0.0);
}

extern "C" int mpc_emu_chol_inv(const double* A, int n, int pcap, double* W) {
  if (n <= 0 || n % 6 || n > pcap) return -1;
  if (pcap == 60) chol_inv_t<60>(A, n, W);
  else if (pcap == 120) chol_inv_t<120>(A, n, W);
  else return -1;
  return 0;
}
