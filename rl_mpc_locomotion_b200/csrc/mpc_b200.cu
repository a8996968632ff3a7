// mpc_b200.cu — sm_100a kernels + the C-ABI of include/mpc_b200.h.
//
// One CTA = MPC_NL threads (two warps) = one robot.  The hardware block scheduler load-balances the
// variable per-robot iteration counts (OSQP terminates after 50..300 iterations depending
// on the state).  The robot's 384-byte problem record is staged HBM -> shared memory with
// one TMA bulk copy (cp.async.bulk, mbarrier completion); everything else the solver
// touches lives in registers / shared memory until the 12 horizon-0 forces are written.
//
// There is NO CPU fallback in this library: without a CUDA device every entry point
// returns MPC_B200_ENOGPU.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "../../include/mpc_b200.h"
#include "mpc_warp_solver.cuh"
#include "mpc_tick.cuh"

namespace {

thread_local char g_err[512] = "";

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return MPC_B200_ECUDA;
}
#define CK(call)                                   \
  do {                                             \
    cudaError_t e_ = (call);                       \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
  } while (0)

// ---------------------------------------------------------------- TMA bulk copy helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

#ifndef MPC_MIN_CTAS
#define MPC_MIN_CTAS 6   // resident robots per SM the shared-memory-walk build is tuned for (168 regs: no hot-loop spills)
#endif
// h = 16: 56 KB of shared memory per robot cap the SM at 4 resident CTAs anyway - ask for no more, so that
// ptxas may use 255 registers instead of 168
template <int H>
constexpr int min_ctas() { return H <= 10 ? MPC_MIN_CTAS : (MPC_MIN_CTAS < 4 ? MPC_MIN_CTAS : 4); }

template <int H, int PCAP = 6 * H>
struct KernelCfg {
  using D = mpc::Dims<H, PCAP>;
  static constexpr int REC_FLOATS = MPC_REC_STRIDE(H);
  static constexpr int REC_BYTES = REC_FLOATS * 4;
  // [solver doubles][record floats (16B aligned)][robot params 12 floats][mbarrier]
  static constexpr int OFF_REC = D::SMEM_DOUBLES * 8;
  static constexpr int OFF_RP = OFF_REC + ((REC_BYTES + 15) & ~15);
  static constexpr int OFF_BAR = OFF_RP + 64;
  static constexpr int SMEM_BYTES = OFF_BAR + 16;
};

// PCAP = polish capacity (free coordinates).  The fast instantiation (6H) covers every contact
// table with <= 2H stance leg-steps; robots beyond that leave MPC_STATUS_POLISH_OVERFLOW and are
// re-solved by the ONLY_OVERFLOW instantiation (PCAP = 12H, larger shared-memory footprint).
template <int H, bool WARM, int PCAP, bool ONLY_OVERFLOW>
__global__ void __launch_bounds__(MPC_NL, (PCAP > 6 * H ? 1 : min_ctas<H>()))
mpc_solve_kernel(const mpc::SolveParams prm, const float* __restrict__ records,
                 const float* __restrict__ robot_params, const int32_t* __restrict__ robot_type, int n_types,
                 int n_robots, float* __restrict__ grf0, float* __restrict__ grf_full,
                 int32_t* __restrict__ status, float* __restrict__ resid, double* __restrict__ info,
                 double* __restrict__ dbg, double* __restrict__ warm, const int32_t* __restrict__ order) {
  using K = KernelCfg<H, PCAP>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  float* rec_s = reinterpret_cast<float*>(smem_raw + K::OFF_REC);
  float* rp_s = reinterpret_cast<float*>(smem_raw + K::OFF_RP);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + K::OFF_BAR);
  const int lane = threadIdx.x;
  if ((int)blockIdx.x >= n_robots) return;
  const int robot = order ? order[blockIdx.x] : (int)blockIdx.x;   // scheduling hint only
  if (robot < 0 || robot >= n_robots) return;                      // not a permutation entry: skipped
  if constexpr (ONLY_OVERFLOW) {
    if (status[robot] != MPC_STATUS_POLISH_OVERFLOW) return;   // uniform per CTA
  }
  const int type = robot_type ? robot_type[robot] : 0;
  if (type < 0 || type >= n_types) {   // bad robot type: report, never index robot_params with it
    if (lane < 12) grf0[(size_t)robot * 12 + lane] = 0.0f;
    if (grf_full)
      for (int i = lane; i < 12 * H; i += MPC_NL) grf_full[(size_t)robot * 12 * H + i] = 0.0f;
    if (status && lane == 0) status[robot] = MPC_STATUS_BAD_INDEX;
    return;
  }

  // stage the robot's state block with one TMA bulk copy while the tables are built
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_load_1d(rec_s, records + (size_t)robot * K::REC_FLOATS, K::REC_BYTES, bar);
  }
  if (lane < MPC_ROBOT_PARAMS_STRIDE) rp_s[lane] = robot_params[(size_t)type * MPC_ROBOT_PARAMS_STRIDE + lane];
  mpc::fill_tables<H, PCAP>(sm, prm.dt, lane);
  WSYNC();
  mbar_wait(bar, 0);

  constexpr int dbg_stride = 12 * H + 144 * H * H + 12 * H + 20 * H + 1;
  mpc::solve_robot<H, WARM, PCAP>(rec_s, rp_s, prm, sm, grf0 + (size_t)robot * 12,
                      grf_full ? grf_full + (size_t)robot * 12 * H : nullptr,
                      status ? status + robot : nullptr, resid ? resid + (size_t)robot * 2 : nullptr,
                      info ? info + (size_t)robot * mpc::INFO_STRIDE : nullptr,
                      dbg ? dbg + (size_t)robot * dbg_stride : nullptr,
                      warm ? warm + (size_t)robot * MPC_WARM_STRIDE(H) : nullptr, nullptr, lane);
}


// ---------------------------------------------------------------------------------------
// Phase-split batch path: the same solve as mpc_solve_kernel, cut at its two natural seams into three
// code-homogeneous kernels (SETUP: structure + Ruiz scaling; ADMM: factorisations + iterations;
// POLISH) that hand each robot over through a ~9 KB workspace row in global memory (L2-resident for
// a batch).  Why: the monolithic kernel's ~200 KB of SASS is shared by CTAs that sit in different
// phases (OSQP's iteration counts de-synchronise them), which thrashes the SM's 32 KB instruction
// cache; per-phase kernels keep each SM in one code region, get a register allocation and a
// shared-memory footprint of their own (more resident robots in the ADMM phase, which is where the
// time goes), and only the ADMM kernel's duration depends on the data.  Results are bit-identical
// to the monolithic kernel (same lane program, tests/test_warp_emu.py::test_phase_split_*).
// ---------------------------------------------------------------------------------------
#ifndef MPC_ADMM_CTAS
#define MPC_ADMM_CTAS 6
#endif
template <int PH>
constexpr int phase_ctas() { return PH == mpc::PH_ADMM ? MPC_ADMM_CTAS : MPC_MIN_CTAS; }

template <int H, bool WARM, int PH>
__global__ void __launch_bounds__(MPC_NL, phase_ctas<PH>())
mpc_phase_kernel(const mpc::SolveParams prm, const float* __restrict__ records,
                 const float* __restrict__ robot_params, const int32_t* __restrict__ robot_type, int n_types,
                 int n_robots, float* __restrict__ grf0, float* __restrict__ grf_full,
                 int32_t* __restrict__ status, float* __restrict__ resid, double* __restrict__ info,
                 double* __restrict__ warm, double* __restrict__ ws_all, const int32_t* __restrict__ order) {
  using K = KernelCfg<H>;
  using WL = mpc::WsLayout<H>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  float* rec_s = reinterpret_cast<float*>(smem_raw + K::OFF_REC);
  float* rp_s = reinterpret_cast<float*>(smem_raw + K::OFF_RP);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + K::OFF_BAR);
  const int lane = threadIdx.x;
  if ((int)blockIdx.x >= n_robots) return;
  const int robot = (PH == mpc::PH_ADMM && order) ? order[blockIdx.x] : (int)blockIdx.x;   // scheduling hint only
  if (robot < 0 || robot >= n_robots) return;
  double* const ws = ws_all + (size_t)robot * WL::STRIDE;
  const int type = robot_type ? robot_type[robot] : 0;
  if (type < 0 || type >= n_types) {   // bad robot type: reported by the SETUP kernel, skipped by the others
    if (PH == mpc::PH_SETUP) {
      if (lane < 12) grf0[(size_t)robot * 12 + lane] = 0.0f;
      if (grf_full)
        for (int i = lane; i < 12 * H; i += MPC_NL) grf_full[(size_t)robot * 12 * H + i] = 0.0f;
      if (status && lane == 0) status[robot] = MPC_STATUS_BAD_INDEX;
    }
    return;
  }
  if (PH == mpc::PH_ADMM && ws[mpc::WS_STAGE] != mpc::kWsStageAdmm) return;       // uniform per CTA
  if (PH == mpc::PH_POLISH && ws[mpc::WS_STAGE] != mpc::kWsStagePolish) return;
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_load_1d(rec_s, records + (size_t)robot * K::REC_FLOATS, K::REC_BYTES, bar);
  }
  if (lane < MPC_ROBOT_PARAMS_STRIDE) rp_s[lane] = robot_params[(size_t)type * MPC_ROBOT_PARAMS_STRIDE + lane];
  mpc::fill_tables<H>(sm, prm.dt, lane);
  WSYNC();
  mbar_wait(bar, 0);
  mpc::solve_robot<H, WARM, 6 * H, PH>(rec_s, rp_s, prm, sm, grf0 + (size_t)robot * 12,
                      grf_full ? grf_full + (size_t)robot * 12 * H : nullptr,
                      status ? status + robot : nullptr, resid ? resid + (size_t)robot * 2 : nullptr,
                      info ? info + (size_t)robot * mpc::INFO_STRIDE : nullptr, nullptr,
                      warm ? warm + (size_t)robot * MPC_WARM_STRIDE(H) : nullptr, ws, lane);
}


// ---------------------------------------------------------------------------------------
// Fused Locomotion tick: tick_pre -> [solve_robot] -> tick_post, one CTA per robot.
// ---------------------------------------------------------------------------------------
template <int H>
struct TickCfg {
  using K = KernelCfg<H>;
  static constexpr int OFF_STATE = K::SMEM_BYTES;                    // float[96]
  static constexpr int OFF_LEGS = OFF_STATE + MPC_TICK_STATE_STRIDE * 4;  // float[4][32]
  static constexpr int OFF_IN = OFF_LEGS + 4 * mpc::LG_STRIDE * 4;   // dof[24] est[18] cmd[16] geom[24] grf[12]
  static constexpr int SMEM_BYTES = OFF_IN + (24 + 18 + 16 + 24 + 12 + 2) * 4;
};

template <int H, bool WARM>
__global__ void __launch_bounds__(MPC_NL, min_ctas<H>())
mpc_tick_kernel(const mpc::SolveParams prm, const mpc::TickParams tp, const float* __restrict__ dof,
                const float* __restrict__ body, const float* __restrict__ est, const float* __restrict__ cmd,
                int has_weights, const float* __restrict__ actions,
                const int32_t* __restrict__ gait, const float* __restrict__ robot_params,
                const float* __restrict__ robot_geom, const int32_t* __restrict__ robot_type, int n_types, int n_robots,
                const int32_t* __restrict__ active,
                float* __restrict__ state, float* __restrict__ torques, float* __restrict__ grf0,
                float* __restrict__ records, int32_t* __restrict__ solved, double* __restrict__ warm,
                const int32_t* __restrict__ order) {
  using K = KernelCfg<H>;
  using T = TickCfg<H>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  float* rec_s = reinterpret_cast<float*>(smem_raw + K::OFF_REC);
  float* rp_s = reinterpret_cast<float*>(smem_raw + K::OFF_RP);
  float* st_s = reinterpret_cast<float*>(smem_raw + T::OFF_STATE);
  float* legs_s = reinterpret_cast<float*>(smem_raw + T::OFF_LEGS);
  float* dof_s = reinterpret_cast<float*>(smem_raw + T::OFF_IN);
  float* est_s = dof_s + 24;
  float* cmd_s = est_s + 18;
  float* geom_s = cmd_s + 16;
  float* grf_s = geom_s + 24;
  int32_t* stat_s = reinterpret_cast<int32_t*>(grf_s + 12);
  const int lane = threadIdx.x;
  if ((int)blockIdx.x >= n_robots) return;
  const int robot = order ? order[blockIdx.x] : (int)blockIdx.x;   // scheduling hint only
  if (robot < 0 || robot >= n_robots) return;
  const int gate = active ? active[robot] : MPC_TICK_RUN;          // the FSM shell's per-robot gate
  if (gate == MPC_TICK_SKIP) return;
  const int type = robot_type ? robot_type[robot] : 0;
  if (type < 0 || type >= n_types) {
    if (lane < 12) torques[(size_t)robot * 12 + lane] = 0.0f;
    if (solved && lane == 0) solved[robot] = MPC_STATUS_BAD_INDEX;
    return;
  }
  for (int i = lane; i < MPC_TICK_STATE_STRIDE; i += MPC_NL) st_s[i] = state[(size_t)robot * MPC_TICK_STATE_STRIDE + i];
  if (lane < 24) {
    dof_s[lane] = dof[(size_t)robot * 24 + lane];
    geom_s[lane] = robot_geom[(size_t)type * MPC_GEOM_STRIDE + lane];
  }
  if (est && lane < 18) est_s[lane] = est[(size_t)robot * MPC_EST_STRIDE + lane];
  // commands: vx, vy, yaw rate [+ 13 weights]; with `actions` the weights are actions * scale + const
  // (two float32 roundings as torch.mul(...).add(...), aliengo.py:237-245; 13th weight 0, :253)
  const int cstride = tp.cmd_stride > 0 ? tp.cmd_stride : MPC_CMD_STRIDE;
  if (lane < 16) {
    float v = 0.0f;
    if (lane < 3) v = cmd[(size_t)robot * cstride + lane];
    else if (actions) v = lane < 15 ? __fadd_rn(__fmul_rn(actions[(size_t)robot * 12 + lane - 3], tp.act_scale[lane - 3]),
                                                 tp.act_const[lane - 3]) : 0.0f;
    else if (has_weights && cstride >= MPC_CMD_STRIDE) v = cmd[(size_t)robot * cstride + lane];
    cmd_s[lane] = v;
  }
  const bool with_weights = actions != nullptr || (has_weights && cstride >= MPC_CMD_STRIDE);
  if (lane < MPC_ROBOT_PARAMS_STRIDE) rp_s[lane] = robot_params[(size_t)type * MPC_ROBOT_PARAMS_STRIDE + lane];
  if (lane == 0) stat_s[0] = 0;
  for (int i = lane; i < K::REC_FLOATS; i += MPC_NL) rec_s[i] = 0.0f;
  __syncthreads();
  if (!est) {  // StateEstimator.update on the device (uses the previous tick's ground normal)
    if (lane == 0) mpc::estimator_update(body + (size_t)robot * 13, st_s + MPC_ST_NORMAL, est_s, grf_s);
    __syncthreads();
  }
  const bool have_rp = !est;   // grf_s[0..1] temporarily holds the world roll / pitch (float16 values)
  const float roll16 = have_rp ? grf_s[0] : 0.0f, pitch16 = have_rp ? grf_s[1] : 0.0f;
  const int g = gait ? gait[robot] : 0;
  mpc::tick_legs(dof_s, geom_s, st_s, rec_s, legs_s, lane);
  // FSM_State_Locomotion.checkTransition -> locomotionSafe() comes BEFORE the control step
  // (ControlFSM.py:84-95): an unsafe robot is not ticked, its commands stay zero
  int unsafe = 0;
  if (have_rp) {
    const float rp2[2] = {roll16, pitch16};
    unsafe = mpc::locomotion_unsafe(rp2, legs_s + mpc::LG_P, mpc::LG_STRIDE);   // uniform (shared-memory inputs)
  }
  if (gate == MPC_TICK_CHECK_ONLY || (gate == MPC_TICK_CHECK_RUN && unsafe)) {
    if (lane < 12) torques[(size_t)robot * 12 + lane] = 0.0f;
    if (lane == 0) reinterpret_cast<int*>(state)[(size_t)robot * MPC_TICK_STATE_STRIDE + MPC_ST_UNSAFE] = unsafe;
    if (solved && lane == 0) solved[robot] = 0;
    return;
  }
  const int do_solve = mpc::tick_pre<H>(est_s, cmd_s, with_weights ? cmd_s + 3 : nullptr, g, geom_s, tp, st_s, rec_s,
                                        legs_s, lane);
  if (have_rp && lane == 0) reinterpret_cast<int*>(st_s)[MPC_ST_UNSAFE] = unsafe;
  if (records)
    for (int i = lane; i < K::REC_FLOATS; i += MPC_NL) records[(size_t)robot * K::REC_FLOATS + i] = rec_s[i];
  if (do_solve) {
    mpc::fill_tables<H>(sm, prm.dt, lane);
    __shared__ double info_s[mpc::INFO_STRIDE];
    mpc::solve_robot<H, WARM>(rec_s, rp_s, prm, sm, grf_s, nullptr, stat_s, nullptr, info_s, nullptr,
                              warm ? warm + (size_t)robot * MPC_WARM_STRIDE(H) : nullptr, nullptr, lane);
    __syncthreads();
    if (lane == 0) reinterpret_cast<int*>(st_s)[MPC_ST_ITERS] = (int)info_s[mpc::INFO_ITERS];
    if (grf0 && lane < 12) grf0[(size_t)robot * 12 + lane] = grf_s[lane];
  }
  // fallback policy: forces of a solve that did not reach OSQP_SOLVED are not applied, the previous
  // MPC tick's f_ff stays in force (the reference would raise on its empty result, mpc_osqp.cc:786-794).
  // MPC_STATUS_POLISH_OVERFLOW = OSQP_SOLVED whose polish did not fit this kernel's buffer (no reference
  // gait gets there): the converged, un-polished ADMM solution is applied.
  const bool apply = do_solve && (stat_s[0] == MPC_STATUS_SOLVED || stat_s[0] == MPC_STATUS_POLISH_OVERFLOW);
  mpc::tick_post(est_s, geom_s, apply ? grf_s : nullptr, st_s, legs_s, torques + (size_t)robot * 12, lane);
  for (int i = lane; i < MPC_TICK_STATE_STRIDE; i += MPC_NL) state[(size_t)robot * MPC_TICK_STATE_STRIDE + i] = st_s[i];
  if (solved && lane == 0) solved[robot] = stat_s[0];
}

// StateEstimator.update alone (debug / parity hook): one thread per robot
__global__ void estimator_kernel(const float* __restrict__ body, const float* __restrict__ normal,
                                 float* __restrict__ est, int n) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) mpc::estimator_update(body + (size_t)r * 13, normal + (size_t)r * 3, est + (size_t)r * MPC_EST_STRIDE);
}

template <int H>
int launch_tick(const mpc::SolveParams& prm, const mpc::TickParams& tp, const float* dof, const float* body,
                const float* est, const float* cmd, int has_weights, const float* actions, const int32_t* gait,
                const float* rp, const float* geom, const int32_t* rt, int n_types, int n, const int32_t* active,
                float* state, float* torques, float* grf0, float* records,
                int32_t* solved, double* warm, const int32_t* order, cudaStream_t stream) {
  using T = TickCfg<H>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(mpc_tick_kernel<H, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(mpc_tick_kernel<H, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(max dynamic smem, tick)");
  if (n <= 0) return MPC_B200_OK;
  if (warm)
    mpc_tick_kernel<H, true><<<n, MPC_NL, T::SMEM_BYTES, stream>>>(prm, tp, dof, body, est, cmd, has_weights, actions,
                                                                  gait, rp, geom, rt, n_types, n, active, state, torques,
                                                                  grf0, records, solved, warm, order);
  else
    mpc_tick_kernel<H, false><<<n, MPC_NL, T::SMEM_BYTES, stream>>>(prm, tp, dof, body, est, cmd, has_weights, actions,
                                                                   gait, rp, geom, rt, n_types, n, active, state, torques,
                                                                   grf0, records, solved, nullptr, order);
  CK(cudaGetLastError());
  return MPC_B200_OK;
}

mpc::SolveParams to_params(const MpcB200Settings& s) {
  mpc::SolveParams p;
  memset(&p, 0, sizeof(p));
  p.max_iter = s.max_iter;
  p.check_every = s.check_every;
  p.scaling_iters = s.scaling_iters;
  p.polish = s.polish;
  p.polish_refine = s.polish_refine;
  p.adaptive_rho = (s.mode == MPC_MODE_FIXED) ? 0 : s.adaptive_rho;
  p.fixed_iters = (s.mode == MPC_MODE_FIXED) ? s.admm_iters : 0;
  p.dt = s.dt;
  p.alpha = s.alpha;
  p.rho0 = s.rho;
  p.sigma = s.sigma;
  p.relax = s.relax;
  p.eps_abs = s.eps_abs;
  p.eps_rel = s.eps_rel;
  p.delta = s.delta;
  p.rho_tol_ratio = s.rho_tol_ratio;
  return p;
}

int check_settings(const MpcB200Settings* s) {
  if (!s) {
    snprintf(g_err, sizeof(g_err), "settings is NULL");
    return MPC_B200_EINVAL;
  }
  if (s->horizon != 10 && s->horizon != 16) {
    snprintf(g_err, sizeof(g_err), "horizon %d not built (supported: 10, 16)", s->horizon);
    return MPC_B200_EINVAL;
  }
  if (s->mode == MPC_MODE_FIXED && s->admm_iters <= 0) {
    snprintf(g_err, sizeof(g_err), "MPC_MODE_FIXED needs admm_iters > 0");
    return MPC_B200_EINVAL;
  }
  if (s->mode != MPC_MODE_FIXED && s->mode != MPC_MODE_OSQP) {
    snprintf(g_err, sizeof(g_err), "unknown mode %d", s->mode);
    return MPC_B200_EINVAL;
  }
  if (s->kernel_path < MPC_PATH_AUTO || s->kernel_path > MPC_PATH_SPLIT) {
    snprintf(g_err, sizeof(g_err), "unknown kernel_path %d", s->kernel_path);
    return MPC_B200_EINVAL;
  }
  if (s->check_every <= 0 || s->max_iter <= 0 || !(s->dt > 0) || !(s->rho > 0)) {
    snprintf(g_err, sizeof(g_err), "invalid settings (check_every/max_iter/dt/rho must be > 0)");
    return MPC_B200_EINVAL;
  }
  return MPC_B200_OK;
}

template <int H, bool WARM>
int launch_w(const mpc::SolveParams& prm, const float* rec, const float* rp, const int32_t* rt, int nt, int n, float* grf0,
             float* grf_full, int32_t* status, float* resid, double* info, double* dbg, double* warm,
             const int32_t* order, int path, cudaStream_t stream) {
  using K = KernelCfg<H>;
  using KB = KernelCfg<H, 12 * H>;
  using WL = mpc::WsLayout<H>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    auto set = [](const void* f, int bytes) {
      if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    };
    set((const void*)mpc_solve_kernel<H, WARM, 6 * H, false>, K::SMEM_BYTES);
    set((const void*)mpc_solve_kernel<H, WARM, 12 * H, true>, KB::SMEM_BYTES);
    set((const void*)mpc_phase_kernel<H, WARM, mpc::PH_SETUP>, K::SMEM_BYTES);
    set((const void*)mpc_phase_kernel<H, WARM, mpc::PH_ADMM>, K::SMEM_BYTES);
    set((const void*)mpc_phase_kernel<H, WARM, mpc::PH_POLISH>, K::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(max dynamic smem)");
  if (n <= 0) return MPC_B200_OK;
  // the overflow pass needs the per-robot status even when the caller does not want it
  int32_t* st = status;
  if (!st) CK(cudaMallocAsync((void**)&st, (size_t)n * sizeof(int32_t), stream));
  // the phase-split path is opt-in (measured slower than the monolithic kernel, see include/mpc_b200.h);
  // the debug hook needs the monolithic kernel
  const bool split = !dbg && path == MPC_PATH_SPLIT;
  if (split) {
    // the workspace comes from the device's stream-ordered pool; keep freed blocks in the pool (the
    // default release threshold of 0 hands them back to the driver at every synchronisation, which
    // would turn each call's cudaMallocAsync into a fresh 37 MB allocation)
    static std::once_flag pool_once[64];
    int dev = 0;
    CK(cudaGetDevice(&dev));
    std::call_once(pool_once[dev & 63], [dev] {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
    });
    double* ws = nullptr;
    CK(cudaMallocAsync((void**)&ws, (size_t)n * WL::STRIDE * sizeof(double), stream));
    mpc_phase_kernel<H, WARM, mpc::PH_SETUP><<<n, MPC_NL, K::SMEM_BYTES, stream>>>(prm, rec, rp, rt, nt, n, grf0, grf_full, st,
                                                                                  resid, info, warm, ws, nullptr);
    CK(cudaGetLastError());
    mpc_phase_kernel<H, WARM, mpc::PH_ADMM><<<n, MPC_NL, K::SMEM_BYTES, stream>>>(prm, rec, rp, rt, nt, n, grf0, grf_full, st,
                                                                                 resid, info, warm, ws, order);
    CK(cudaGetLastError());
    if (prm.polish) {
      mpc_phase_kernel<H, WARM, mpc::PH_POLISH><<<n, MPC_NL, K::SMEM_BYTES, stream>>>(prm, rec, rp, rt, nt, n, grf0, grf_full,
                                                                                     st, resid, info, warm, ws, nullptr);
      CK(cudaGetLastError());
    }
    CK(cudaFreeAsync(ws, stream));
  } else {
    // development knob (occupancy experiments): extra dynamic shared memory per CTA
    static const int extra = [] {
      const char* e = getenv("MPC_B200_DEBUG_EXTRA_SMEM");
      const int v = e ? atoi(e) : 0;
      if (v > 0) cudaFuncSetAttribute((const void*)mpc_solve_kernel<H, WARM, 6 * H, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES + v);
      return v > 0 ? v : 0;
    }();
    mpc_solve_kernel<H, WARM, 6 * H, false><<<n, MPC_NL, K::SMEM_BYTES + extra, stream>>>(prm, rec, rp, rt, nt, n, grf0, grf_full,
                                                                                         st, resid, info, dbg, warm, order);
    CK(cudaGetLastError());
  }
  // robots whose polish did not fit the fast kernel's buffer (> 2H stance leg-steps): every other
  // CTA of this launch exits at once
  if (prm.polish) {
    mpc_solve_kernel<H, WARM, 12 * H, true><<<n, MPC_NL, KB::SMEM_BYTES, stream>>>(prm, rec, rp, rt, nt, n, grf0, grf_full, st,
                                                                                 resid, info, dbg, warm, order);
    CK(cudaGetLastError());
  }
  if (!status) CK(cudaFreeAsync(st, stream));
  return MPC_B200_OK;
}

template <int H>
int launch(const mpc::SolveParams& prm, const float* rec, const float* rp, const int32_t* rt, int nt, int n,
           float* grf0, float* grf_full, int32_t* status, float* resid, double* info, double* dbg, double* warm,
           const int32_t* order, int path, cudaStream_t stream) {
  if (warm)
    return launch_w<H, true>(prm, rec, rp, rt, nt, n, grf0, grf_full, status, resid, info, dbg, warm, order, path, stream);
  return launch_w<H, false>(prm, rec, rp, rt, nt, n, grf0, grf_full, status, resid, info, dbg, nullptr, order, path, stream);
}

int dispatch(const MpcB200Settings* s, const float* rec, const float* rp, const int32_t* rt, int nt, int n,
             float* grf0, float* grf_full, int32_t* status, float* resid, double* info, double* dbg, double* warm,
             const int32_t* order, cudaStream_t stream) {
  const mpc::SolveParams prm = to_params(*s);
  if (s->horizon == 10)
    return launch<10>(prm, rec, rp, rt, nt, n, grf0, grf_full, status, resid, info, dbg, warm, order, s->kernel_path, stream);
  return launch<16>(prm, rec, rp, rt, nt, n, grf0, grf_full, status, resid, info, dbg, warm, order, s->kernel_path, stream);
}

// grow-only device scratch for the host-pointer convenience entry points
struct HostPath {
  std::mutex mu;
  cudaStream_t stream = nullptr;
  void* buf = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (!stream) CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (bytes > cap) {
      if (buf) cudaFree(buf);
      buf = nullptr;
      cap = 0;
      CK(cudaMalloc(&buf, bytes));
      cap = bytes;
    }
    return MPC_B200_OK;
  }
} g_host;

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

extern "C" {

void mpc_b200_default_settings(MpcB200Settings* s) {
  memset(s, 0, sizeof(*s));
  s->horizon = 10;              // ConvexMPCLocomotion.py:27
  s->mode = MPC_MODE_OSQP;
  s->admm_iters = 50;
  s->max_iter = 4000;           // osqp constants.h:61
  s->check_every = 25;          // constants.h:83, mpc_osqp.cc:710
  s->scaling_iters = 10;        // constants.h:85
  s->adaptive_rho = 1;          // constants.h:105
  s->polish = 1;                // mpc_osqp.cc:709
  s->polish_refine = 3;         // constants.h:78
  s->dt = 0.02;                 // RobotRunnerMin.py:21-22 (dtMPC = 0.01 * 2)
  s->alpha = 1e-5;              // Parameters.py:6
  s->rho = 0.1;                 // constants.h:59
  s->sigma = 1e-6;              // constants.h:60
  s->relax = 1.6;               // constants.h:66
  s->eps_abs = 1e-3;            // mpc_osqp.cc:711
  s->eps_rel = 1e-3;            // mpc_osqp.cc:712
  s->delta = 1e-6;              // constants.h:76
  s->rho_tol_ratio = 5.0;       // constants.h:110
}

const char* mpc_b200_version(void) { return "rl_mpc_locomotion_b200 0.1.0 (sm_100a)"; }
const char* mpc_b200_last_error(void) { return g_err; }

int mpc_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int mpc_b200_kernel_info(int horizon, int* smem_bytes, int* robots_per_cta, int* regs_per_thread) {
  if (horizon != 10 && horizon != 16) return MPC_B200_EINVAL;
  if (smem_bytes) *smem_bytes = horizon == 10 ? KernelCfg<10>::SMEM_BYTES : KernelCfg<16>::SMEM_BYTES;
  if (robots_per_cta) *robots_per_cta = 1;  /* MPC_NL threads cooperate on it */
  if (regs_per_thread) {
    *regs_per_thread = -1;
    cudaFuncAttributes a;
    cudaError_t e = horizon == 10 ? cudaFuncGetAttributes(&a, mpc_solve_kernel<10, false, 60, false>)
                                  : cudaFuncGetAttributes(&a, mpc_solve_kernel<16, false, 96, false>);
    if (e == cudaSuccess) *regs_per_thread = a.numRegs;
    else cudaGetLastError();
  }
  return MPC_B200_OK;
}

int mpc_b200_solve_batch(const MpcB200Settings* settings, const float* d_records,
                         const float* d_robot_params, const int32_t* d_robot_type, int n_types, int n_robots,
                         float* d_grf0, float* d_grf_full, int32_t* d_status, float* d_resid,
                         double* d_info, double* d_warm, const int32_t* d_order, void* stream) {
  int rc = check_settings(settings);
  if (rc) return rc;
  if (n_robots < 0 || n_types < 1 || (n_robots > 0 && (!d_records || !d_robot_params || !d_grf0))) {
    snprintf(g_err, sizeof(g_err), "null pointer / negative batch / n_types < 1");
    return MPC_B200_EINVAL;
  }
  if (mpc_b200_device_count() <= 0) {
    snprintf(g_err, sizeof(g_err), "no CUDA device: this library has no CPU path");
    return MPC_B200_ENOGPU;
  }
  if (n_robots == 0) return MPC_B200_OK;
  return dispatch(settings, d_records, d_robot_params, d_robot_type, n_types, n_robots, d_grf0, d_grf_full, d_status,
                  d_resid, d_info, nullptr, d_warm, d_order, (cudaStream_t)stream);
}

int mpc_b200_solve_batch_host(const MpcB200Settings* settings, const float* h_records,
                              const float* h_robot_params, const int32_t* h_robot_type, int n_types, int n_robots,
                              float* h_grf0, float* h_grf_full, int32_t* h_status) {
  int rc = check_settings(settings);
  if (rc) return rc;
  if (!h_records || !h_robot_params || !h_grf0 || n_robots < 0 || n_types < 1) {
    snprintf(g_err, sizeof(g_err), "null pointer / negative batch / n_types < 1");
    return MPC_B200_EINVAL;
  }
  if (h_robot_type)
    for (int i = 0; i < n_robots; ++i)
      if (h_robot_type[i] < 0 || h_robot_type[i] >= n_types) {
        snprintf(g_err, sizeof(g_err), "robot_type[%d] = %d outside [0, %d)", i, h_robot_type[i], n_types);
        return MPC_B200_EINVAL;
      }
  if (mpc_b200_device_count() <= 0) {
    snprintf(g_err, sizeof(g_err), "no CUDA device: this library has no CPU path");
    return MPC_B200_ENOGPU;
  }
  if (n_robots == 0) return MPC_B200_OK;
  const int h = settings->horizon;
  const size_t b_rec = align256((size_t)n_robots * MPC_REC_STRIDE(h) * 4);
  const size_t b_rp = align256((size_t)n_types * MPC_ROBOT_PARAMS_STRIDE * 4);
  const size_t b_rt = align256((size_t)n_robots * 4);
  const size_t b_g0 = align256((size_t)n_robots * 12 * 4);
  const size_t b_gf = align256((size_t)n_robots * 12 * h * 4);
  const size_t b_st = align256((size_t)n_robots * 4);
  std::lock_guard<std::mutex> lock(g_host.mu);
  rc = g_host.ensure(b_rec + b_rp + b_rt + b_g0 + b_gf + b_st);
  if (rc) return rc;
  char* p = (char*)g_host.buf;
  float* d_rec = (float*)p; p += b_rec;
  float* d_rp = (float*)p; p += b_rp;
  int32_t* d_rt = (int32_t*)p; p += b_rt;
  float* d_g0 = (float*)p; p += b_g0;
  float* d_gf = (float*)p; p += b_gf;
  int32_t* d_st = (int32_t*)p;
  cudaStream_t st = g_host.stream;
  CK(cudaMemcpyAsync(d_rec, h_records, (size_t)n_robots * MPC_REC_STRIDE(h) * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_rp, h_robot_params, (size_t)n_types * MPC_ROBOT_PARAMS_STRIDE * 4,
                     cudaMemcpyHostToDevice, st));
  if (h_robot_type) CK(cudaMemcpyAsync(d_rt, h_robot_type, (size_t)n_robots * 4, cudaMemcpyHostToDevice, st));
  rc = dispatch(settings, d_rec, d_rp, h_robot_type ? d_rt : nullptr, n_types, n_robots, d_g0,
                h_grf_full ? d_gf : nullptr, h_status ? d_st : nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h_grf0, d_g0, (size_t)n_robots * 12 * 4, cudaMemcpyDeviceToHost, st));
  if (h_grf_full) CK(cudaMemcpyAsync(h_grf_full, d_gf, (size_t)n_robots * 12 * h * 4, cudaMemcpyDeviceToHost, st));
  if (h_status) CK(cudaMemcpyAsync(h_status, d_st, (size_t)n_robots * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return MPC_B200_OK;
}

int mpc_b200_tick(const MpcB200Settings* settings, const MpcB200TickParams* tick, const float* d_dof,
                  const float* d_body, const float* d_est, const float* d_cmd, int has_weights,
                  const float* d_actions, const int32_t* d_gait,
                  const float* d_robot_params, const float* d_robot_geom, const int32_t* d_robot_type, int n_types,
                  int n_robots, const int32_t* d_active, float* d_state, float* d_torques, float* d_grf0,
                  float* d_records, int32_t* d_solved, double* d_warm, const int32_t* d_order, void* stream) {
  int rc = check_settings(settings);
  if (rc) return rc;
  if (!tick || tick->iters_between_mpc <= 0 || !(tick->dt > 0)) {
    snprintf(g_err, sizeof(g_err), "invalid tick parameters");
    return MPC_B200_EINVAL;
  }
  if (n_robots < 0 || n_types < 1 || (n_robots > 0 && (!d_dof || (!d_est == !d_body) || !d_cmd || !d_robot_params ||
                                                       !d_robot_geom || !d_state || !d_torques))) {
    snprintf(g_err, sizeof(g_err), "null pointer / negative batch / n_types < 1");
    return MPC_B200_EINVAL;
  }
  const int cstride = tick->cmd_stride > 0 ? tick->cmd_stride : MPC_CMD_STRIDE;
  if (cstride != 3 && cstride != MPC_CMD_STRIDE) {
    snprintf(g_err, sizeof(g_err), "cmd_stride must be 3 or %d", MPC_CMD_STRIDE);
    return MPC_B200_EINVAL;
  }
  if (has_weights && !d_actions && cstride != MPC_CMD_STRIDE) {
    snprintf(g_err, sizeof(g_err), "has_weights needs cmd_stride == %d (or pass d_actions)", MPC_CMD_STRIDE);
    return MPC_B200_EINVAL;
  }
  if (d_active && !d_body) {
    snprintf(g_err, sizeof(g_err), "d_active (FSM gate with locomotionSafe) needs d_body");
    return MPC_B200_EINVAL;
  }
  if (mpc_b200_device_count() <= 0) {
    snprintf(g_err, sizeof(g_err), "no CUDA device: this library has no CPU path");
    return MPC_B200_ENOGPU;
  }
  if (n_robots == 0) return MPC_B200_OK;
  mpc::SolveParams prm = to_params(*settings);
  // the MPC timestep is dt * iterationsBetweenMPC (ConvexMPCLocomotion.py:58)
  prm.dt = (double)tick->dt * tick->iters_between_mpc;
  mpc::TickParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.dt = tick->dt;
  tp.iters_between_mpc = tick->iters_between_mpc;
  tp.flat_ground = tick->flat_ground;
  tp.bonus_swing = tick->bonus_swing;
  tp.cmd_stride = cstride;
  for (int i = 0; i < 12; ++i) {
    tp.act_scale[i] = tick->act_scale[i];
    tp.act_const[i] = tick->act_const[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (settings->horizon == 10)
    return launch_tick<10>(prm, tp, d_dof, d_body, d_est, d_cmd, has_weights, d_actions, d_gait, d_robot_params,
                           d_robot_geom, d_robot_type, n_types, n_robots, d_active, d_state, d_torques, d_grf0, d_records,
                           d_solved, d_warm, d_order, st);
  return launch_tick<16>(prm, tp, d_dof, d_body, d_est, d_cmd, has_weights, d_actions, d_gait, d_robot_params,
                         d_robot_geom, d_robot_type, n_types, n_robots, d_active, d_state, d_torques, d_grf0, d_records,
                         d_solved, d_warm, d_order, st);
}

#ifdef MPC_STAGE_TIMING
int mpc_b200_debug_stage_cycles(unsigned long long* out16, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, g_stage_cycles, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_stage_cycles, z, sizeof(z));
  }
  return 0;
}
#endif

int mpc_b200_debug_estimator(const float* h_body, const float* h_normal, float* h_est, int n) {
  if (!h_body || !h_normal || !h_est || n < 0) {
    snprintf(g_err, sizeof(g_err), "null pointer / negative batch");
    return MPC_B200_EINVAL;
  }
  if (mpc_b200_device_count() <= 0) {
    snprintf(g_err, sizeof(g_err), "no CUDA device: this library has no CPU path");
    return MPC_B200_ENOGPU;
  }
  if (n == 0) return MPC_B200_OK;
  float *d_body = nullptr, *d_normal = nullptr, *d_est = nullptr;
  CK(cudaMalloc(&d_body, (size_t)n * 13 * 4));
  CK(cudaMalloc(&d_normal, (size_t)n * 3 * 4));
  CK(cudaMalloc(&d_est, (size_t)n * MPC_EST_STRIDE * 4));
  int rc = MPC_B200_OK;
  cudaError_t e = cudaMemcpy(d_body, h_body, (size_t)n * 13 * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_normal, h_normal, (size_t)n * 3 * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    estimator_kernel<<<(n + 63) / 64, 64>>>(d_body, d_normal, d_est, n);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(h_est, d_est, (size_t)n * MPC_EST_STRIDE * 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) rc = cuda_fail(e, "debug estimator");
  cudaFree(d_body);
  cudaFree(d_normal);
  cudaFree(d_est);
  return rc;
}

int mpc_b200_debug_build_qp(const MpcB200Settings* settings, const float* h_record,
                            const float* h_robot_params, double* h_P, double* h_q, double* h_D,
                            double* h_E, double* h_c) {
  int rc = check_settings(settings);
  if (rc) return rc;
  if (mpc_b200_device_count() <= 0) {
    snprintf(g_err, sizeof(g_err), "no CUDA device: this library has no CPU path");
    return MPC_B200_ENOGPU;
  }
  const int h = settings->horizon, nv = 12 * h;
  const size_t ndbg = (size_t)nv + (size_t)nv * nv + nv + 20 * h + 1;
  float *d_rec = nullptr, *d_rp = nullptr, *d_g0 = nullptr;
  double* d_dbg = nullptr;
  CK(cudaMalloc(&d_rec, MPC_REC_STRIDE(h) * 4));
  CK(cudaMalloc(&d_rp, MPC_ROBOT_PARAMS_STRIDE * 4));
  CK(cudaMalloc(&d_g0, 12 * 4));
  CK(cudaMalloc(&d_dbg, ndbg * 8));
  CK(cudaMemcpy(d_rec, h_record, MPC_REC_STRIDE(h) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_rp, h_robot_params, MPC_ROBOT_PARAMS_STRIDE * 4, cudaMemcpyHostToDevice));
  rc = dispatch(settings, d_rec, d_rp, nullptr, 1, 1, d_g0, nullptr, nullptr, nullptr, nullptr, d_dbg, nullptr, nullptr,
                nullptr);
  if (rc == MPC_B200_OK) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = cuda_fail(e, "debug kernel");
  }
  if (rc == MPC_B200_OK) {
    if (h_q) cudaMemcpy(h_q, d_dbg, nv * 8, cudaMemcpyDeviceToHost);
    if (h_P) cudaMemcpy(h_P, d_dbg + nv, (size_t)nv * nv * 8, cudaMemcpyDeviceToHost);
    if (h_D) cudaMemcpy(h_D, d_dbg + nv + (size_t)nv * nv, nv * 8, cudaMemcpyDeviceToHost);
    if (h_E) cudaMemcpy(h_E, d_dbg + nv + (size_t)nv * nv + nv, 20 * h * 8, cudaMemcpyDeviceToHost);
    if (h_c) cudaMemcpy(h_c, d_dbg + ndbg - 1, 8, cudaMemcpyDeviceToHost);
  }
  cudaFree(d_rec);
  cudaFree(d_rp);
  cudaFree(d_g0);
  cudaFree(d_dbg);
  return rc;
}

}  // extern "C"
