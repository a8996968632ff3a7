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

template <int H>
static void run(const float* rec, int n, const float* rp, const int* rt, const mpc::SolveParams& prm,
                float* grf0, float* grf_full, int32_t* status, double* info, double* dbg) {
  using D = mpc::Dims<H>;
  std::vector<double> sm(D::SMEM_DOUBLES, 0.0);
  mpc::fill_tables<H>(sm.data(), prm.dt);
  const int stride = MPC_REC_STRIDE(H);
  const int dbg_stride = D::NV + D::NV * D::NV + D::NV + 20 * H + 1;
  for (int r = 0; r < n; ++r) {
    const float* rpr = rp + (size_t)(rt ? rt[r] : 0) * MPC_ROBOT_PARAMS_STRIDE;
    mpc::solve_robot<H>(rec + (size_t)r * stride, rpr, prm, sm.data(), grf0 + (size_t)r * 12,
                        grf_full ? grf_full + (size_t)r * 12 * H : nullptr, status ? status + r : nullptr,
                        nullptr, info ? info + (size_t)r * mpc::INFO_STRIDE : nullptr,
                        dbg ? dbg + (size_t)r * dbg_stride : nullptr);
  }
}

extern "C" int mpc_emu_solve(const float* rec, int n, int horizon, const float* rp, const int* rt,
                             const mpc::SolveParams* prm, float* grf0, float* grf_full,
                             int32_t* status, double* info, double* dbg) {
  if (horizon == 10) run<10>(rec, n, rp, rt, *prm, grf0, grf_full, status, info, dbg);
  else if (horizon == 16) run<16>(rec, n, rp, rt, *prm, grf0, grf_full, status, info, dbg);
  else return -1;
  return 0;
}
