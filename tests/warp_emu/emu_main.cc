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
                        warm ? warm + (size_t)r * MPC_WARM_STRIDE(H) : nullptr, nullptr);
    if (st == MPC_STATUS_POLISH_OVERFLOW)   // the overflow kernel's job on the device
      mpc::solve_robot<H, true, 12 * H>(rec + (size_t)r * stride, rpr, prm, smb.data(), grf0 + (size_t)r * 12,
                          grf_full ? grf_full + (size_t)r * 12 * H : nullptr, &st,
                          nullptr, info ? info + (size_t)r * mpc::INFO_STRIDE : nullptr,
                          dbg ? dbg + (size_t)r * dbg_stride : nullptr,
                          warm ? warm + (size_t)r * MPC_WARM_STRIDE(H) : nullptr, nullptr);
    if (status) status[r] = st;
  }
}

// The phase-split batch path (three kernels on the device): SETUP -> ADMM -> POLISH through the workspace row.
template <int H>
static void run_split(const float* rec, int n, const float* rp, const int* rt, const mpc::SolveParams& prm,
                      float* grf0, float* grf_full, int32_t* status, double* info, double* warm) {
  using D = mpc::Dims<H>;
  using DB = mpc::Dims<H, 12 * H>;
  using WL = mpc::WsLayout<H>;
  std::vector<double> smb(DB::SMEM_DOUBLES, 0.0), ws(WL::STRIDE, 0.0);
  mpc::fill_tables<H, 12 * H>(smb.data(), prm.dt);
  const int stride = MPC_REC_STRIDE(H);
  for (int r = 0; r < n; ++r) {
    const float* rpr = rp + (size_t)(rt ? rt[r] : 0) * MPC_ROBOT_PARAMS_STRIDE;
    const float* rc = rec + (size_t)r * stride;
    float* g0 = grf0 + (size_t)r * 12;
    float* gf = grf_full ? grf_full + (size_t)r * 12 * H : nullptr;
    double* inf = info ? info + (size_t)r * mpc::INFO_STRIDE : nullptr;
    double* wm = warm ? warm + (size_t)r * MPC_WARM_STRIDE(H) : nullptr;
    int32_t st = 0;
    // every phase runs in a FRESH shared-memory image: nothing but the workspace row crosses a kernel boundary
    { std::vector<double> sm(D::SMEM_DOUBLES, -7.0); mpc::fill_tables<H>(sm.data(), prm.dt);
      mpc::solve_robot<H, true, 6 * H, mpc::PH_SETUP>(rc, rpr, prm, sm.data(), g0, gf, &st, nullptr, inf, nullptr, wm, ws.data()); }
    { std::vector<double> sm(D::SMEM_DOUBLES, -7.0); mpc::fill_tables<H>(sm.data(), prm.dt);
      mpc::solve_robot<H, true, 6 * H, mpc::PH_ADMM>(rc, rpr, prm, sm.data(), g0, gf, &st, nullptr, inf, nullptr, wm, ws.data()); }
    if (ws[mpc::WS_STAGE] == mpc::kWsStagePolish) {
      std::vector<double> sm(D::SMEM_DOUBLES, -7.0); mpc::fill_tables<H>(sm.data(), prm.dt);
      mpc::solve_robot<H, true, 6 * H, mpc::PH_POLISH>(rc, rpr, prm, sm.data(), g0, gf, &st, nullptr, inf, nullptr, wm, ws.data());
    }
    if (st == MPC_STATUS_POLISH_OVERFLOW)
      mpc::solve_robot<H, true, 12 * H>(rc, rpr, prm, smb.data(), g0, gf, &st, nullptr, inf, nullptr, wm, nullptr);
    if (status) status[r] = st;
  }
}

extern "C" int mpc_emu_solve_split(const float* rec, int n, int horizon, const float* rp, const int* rt,
                                   const mpc::SolveParams* prm, float* grf0, float* grf_full,
                                   int32_t* status, double* info, double* warm) {
  if (horizon == 10) run_split<10>(rec, n, rp, rt, *prm, grf0, grf_full, status, info, warm);
  else if (horizon == 16) run_split<16>(rec, n, rp, rt, *prm, grf0, grf_full, status, info, warm);
  else return -1;
  return 0;
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
                            const float* body, const float* est, const float* cmd, int has_weights,
                            const float* actions, const int* gait,
                            const float* rp, const float* geom, const int* rt, int n, const int* active, float* state,
                            float* torques, float* grf0, float* records, int32_t* solved, double* warm) {
  constexpr int H = 10;
  using D = mpc::Dims<H>;
  mpc::SolveParams prm = *prm_in;
  prm.dt = (double)tp->dt * tp->iters_between_mpc;
  std::vector<double> sm(D::SMEM_DOUBLES, 0.0);
  const int stride = MPC_REC_STRIDE(H);
  const int cstride = tp->cmd_stride > 0 ? tp->cmd_stride : MPC_CMD_STRIDE;
  for (int r = 0; r < n; ++r) {
    const int gate = active ? active[r] : MPC_TICK_RUN;
    if (gate == MPC_TICK_SKIP) continue;
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
    float cmd_s[16] = {0};
    for (int i = 0; i < 16; ++i) {
      if (i < 3) cmd_s[i] = cmd[(size_t)r * cstride + i];
      else if (actions) {
        if (i < 15) {
          volatile float prod = actions[(size_t)r * 12 + i - 3] * tp->act_scale[i - 3];   // two roundings, no FMA
          cmd_s[i] = prod + tp->act_const[i - 3];
        }
      } else if (has_weights && cstride >= MPC_CMD_STRIDE) cmd_s[i] = cmd[(size_t)r * cstride + i];
    }
    const bool with_weights = actions != nullptr || (has_weights && cstride >= MPC_CMD_STRIDE);
    if (!est) mpc::estimator_update(body + (size_t)r * 13, st + MPC_ST_NORMAL, est_local, rp16);
    mpc::tick_legs(dof + (size_t)r * 24, g, st, rec, legs);
    const int unsafe = est ? 0 : mpc::locomotion_unsafe(rp16, legs + mpc::LG_P, mpc::LG_STRIDE);
    if (gate == MPC_TICK_CHECK_ONLY || (gate == MPC_TICK_CHECK_RUN && unsafe)) {
      for (int i = 0; i < 12; ++i) torques[(size_t)r * 12 + i] = 0.0f;
      reinterpret_cast<int*>(st)[MPC_ST_UNSAFE] = unsafe;
      if (solved) solved[r] = 0;
      continue;
    }
    const int do_solve = mpc::tick_pre<H>(e, cmd_s, with_weights ? cmd_s + 3 : nullptr, gait ? gait[r] : 0, g, *tp, st,
                                          rec, legs);
    if (!est) reinterpret_cast<int*>(st)[MPC_ST_UNSAFE] = unsafe;
    if (records) std::memcpy(records + (size_t)r * stride, rec, sizeof(rec));
    if (do_solve) {
      mpc::fill_tables<H>(sm.data(), prm.dt);
      double info[mpc::INFO_STRIDE] = {0};
      mpc::solve_robot<H, true>(rec, rp + (size_t)type * MPC_ROBOT_PARAMS_STRIDE, prm, sm.data(), grf, nullptr, &st_code,
                          nullptr, info, nullptr, warm ? warm + (size_t)r * MPC_WARM_STRIDE(H) : nullptr, nullptr);
      reinterpret_cast<int*>(st)[MPC_ST_ITERS] = (int)info[mpc::INFO_ITERS];
      if (grf0) std::memcpy(grf0 + (size_t)r * 12, grf, sizeof(grf));
    }
    const bool apply = do_solve && (st_code == MPC_STATUS_SOLVED || st_code == MPC_STATUS_POLISH_OVERFLOW);
    mpc::tick_post(e, g, apply ? grf : nullptr, st, legs, torques + (size_t)r * 12);
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
