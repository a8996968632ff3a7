/*
 * mpc_b200.h — C-ABI of the B200-native batched convex-MPC stance-force solver.
 *
 * This is the drop-in boundary for the reference's native QP path:
 *   reference/MPC_Controller/convex_MPC/mpc_osqp.cc:178-279  class ConvexMpc
 *   reference/MPC_Controller/convex_MPC/mpc_osqp.cc:578-591  ComputeContactForces(13 args)
 *   reference/MPC_Controller/convex_MPC/mpc_osqp.cc:952-983  PYBIND11_MODULE(mpc_osqp)
 * and (batched extension, no reference counterpart) for the per-robot loop in
 *   reference/RL_Environment/tasks/aliengo.py:227-263        pre_physics_step
 *
 * Everything here is plain C: device/host pointers, sizes, an opaque stream.
 * No torch types.  All entry points return 0 on success or a negative
 * MPC_B200_E* code; none of them synchronises the device unless documented.
 *
 * ---- problem record (one per robot, float32, AoS so that one warp pulls one
 * robot's whole state block with a single TMA bulk copy) --------------------
 * The fields are exactly the 13 arguments of ConvexMpc::ComputeContactForces
 * (mpc_osqp.cc:578-591), in the reference's argument order, followed by the
 * flattened horizon contact table (index i*4+leg, mpc_osqp.cc:598-601).
 */
#ifndef MPC_B200_H
#define MPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* float offsets inside one problem record */
#define MPC_REC_WEIGHTS      0   /* 13  qp_weights                      */
#define MPC_REC_COM_POS     13   /*  3  com_position                    */
#define MPC_REC_COM_VEL     16   /*  3  com_velocity                    */
#define MPC_REC_RPY         19   /*  3  com_roll_pitch_yaw              */
#define MPC_REC_NORMAL      22   /*  3  ground_normal_vec               */
#define MPC_REC_ANGVEL      25   /*  3  com_angular_velocity            */
#define MPC_REC_FOOTPOS     28   /* 12  foot_positions_body_frame       */
#define MPC_REC_MU          40   /*  4  foot_friction_coeffs            */
#define MPC_REC_DES_POS     44   /*  3  desired_com_position            */
#define MPC_REC_DES_VEL     47   /*  3  desired_com_velocity            */
#define MPC_REC_DES_RPY     50   /*  3  desired_com_roll_pitch_yaw      */
#define MPC_REC_DES_ANGVEL  53   /*  3  desired_com_angular_velocity    */
#define MPC_REC_CONTACT     56   /* 4*h foot_contact_states (0/1)       */
/* record stride in floats: 56 + 4*h rounded up to a multiple of 4 (16 B)  */
#define MPC_REC_STRIDE(h)   ((56 + 4 * (h) + 3) & ~3)

/* per-robot-type constant block (float32): what ConvexMpc's ctor receives
 * (mpc_osqp.cc:508-527): mass, inertia[9] row-major, then 2 pad floats.     */
#define MPC_ROBOT_PARAMS_STRIDE 12

#define MPC_B200_MAX_HORIZON 16

/* per-robot OSQP workspace state carried between consecutive solves (float64):
 * flag, rho, x[12h], z[20h], y[20h] (scaled iterates), two slots of unscaled q[12h]            */
#define MPC_WARM_STRIDE(h)  (2 + 76 * (h))

/* error codes */
#define MPC_B200_OK            0
#define MPC_B200_EINVAL       -1   /* bad argument (null ptr, horizon, ...)  */
#define MPC_B200_ECUDA        -2   /* CUDA runtime error (see last_error)    */
#define MPC_B200_ENOGPU       -3   /* no CUDA device: there is NO CPU path   */
#define MPC_B200_ESMEM        -4   /* problem does not fit in shared memory  */

/* per-robot status written by the kernel (int32) */
#define MPC_STATUS_SOLVED          1   /* mirrors OSQP_SOLVED (constants.h)   */
#define MPC_STATUS_MAX_ITER       -2   /* mirrors OSQP_MAX_ITER_REACHED       */
#define MPC_STATUS_NON_FINITE     -7   /* residuals diverged / NaN (mirrors OSQP_NON_CVX): forces zeroed,
                                          warm state untouched; the reference returns an empty vector   */
#define MPC_STATUS_POLISH_OVERFLOW -90 /* more free coordinates than the fast kernel's polish buffer (> 2h stance
                                          leg-steps).  mpc_b200_solve_batch re-solves such robots with the
                                          overflow kernel inside the same call and never returns this code;
                                          mpc_b200_tick (no overflow pass) reports it in d_solved and applies
                                          the UN-polished ADMM solution (see mpc_b200_tick)              */
#define MPC_STATUS_BAD_INDEX      -91  /* d_robot_type[robot] outside [0, n_types): forces zeroed, robot skipped */

/* Solver modes */
#define MPC_MODE_OSQP   0  /* reproduce the reference's OSQP run step for step (fp64): Ruiz
                              scaling, adaptive rho, termination every 25 its, polish.  This
                              is what mpc_osqp.cc:690-796 returns on the same inputs.        */
#define MPC_MODE_FIXED  1  /* same iteration, but exactly `admm_iters` ADMM iterations and no
                              termination test / rho adaptation (BASELINE config 5 sweeps)   */

/* Kernel organisation of mpc_b200_solve_batch (the answers are bit-identical either way):
 * MONO  = one launch, a robot's whole solve in one CTA (the shipped default);
 * SPLIT = three code-homogeneous launches (setup + Ruiz scaling | factorisations + ADMM | polish)
 *         handing robots over through a ~9 KB workspace row each in global memory.  Measured 10 % SLOWER
 *         than MONO at equal occupancy on B200 (three launch tails instead of one outweigh the
 *         instruction-cache relief, profiles/r2_experiments.md); kept selectable for experiments.
 * AUTO  = MONO. */
#define MPC_PATH_AUTO   0
#define MPC_PATH_MONO   1
#define MPC_PATH_SPLIT  2

/* Solver settings.  mpc_b200_default_settings() gives the reference configuration
 * (osqp include/constants.h:59-110 overridden by mpc_osqp.cc:705-712).            */
typedef struct MpcB200Settings {
  int   horizon;        /* planning_horizon (mpc_osqp.cc:511): 10 or 16             */
  int   mode;           /* MPC_MODE_*                                                */
  int   admm_iters;     /* MPC_MODE_FIXED: iteration count                           */
  int   max_iter;       /* MPC_MODE_OSQP: osqp MAX_ITER (4000)                       */
  int   check_every;    /* osqp CHECK_TERMINATION == adaptive_rho_interval (25)      */
  int   scaling_iters;  /* osqp SCALING (10)                                         */
  int   adaptive_rho;   /* osqp ADAPTIVE_RHO (1)                                     */
  int   polish;         /* mpc_osqp.cc:709 (1)                                       */
  int   polish_refine;  /* osqp POLISH_REFINE_ITER (3)                               */
  int   kernel_path;    /* MPC_PATH_*: how mpc_b200_solve_batch runs the solve (results identical) */
  int   reserved_i[2];
  double dt;            /* timestep (mpc_osqp.cc:512), MPC dt = 0.02                 */
  double alpha;         /* force regulariser (mpc_osqp.cc:513), 1e-5                 */
  double rho;           /* osqp RHO 0.1                                              */
  double sigma;         /* osqp SIGMA 1e-6                                           */
  double relax;         /* osqp ALPHA 1.6                                            */
  double eps_abs;       /* mpc_osqp.cc:711  1e-3                                     */
  double eps_rel;       /* mpc_osqp.cc:712  1e-3                                     */
  double delta;         /* osqp DELTA 1e-6 (polish regularisation)                   */
  double rho_tol_ratio; /* osqp ADAPTIVE_RHO_TOLERANCE 5                             */
  double reserved_d[3];
} MpcB200Settings;

void mpc_b200_default_settings(MpcB200Settings* s);

/* library / device probes (host only, never touch the GPU context lazily) */
const char* mpc_b200_version(void);
const char* mpc_b200_last_error(void);
int  mpc_b200_device_count(void);

/* Bytes of dynamic shared memory one CTA of the solve kernel needs for the
 * given horizon, and the number of robots (warps) per CTA it was built for. */
int  mpc_b200_kernel_info(int horizon, int* smem_bytes, int* robots_per_cta,
                          int* regs_per_thread);

/*
 * Batched solve: n_robots independent QPs, one warp per robot.
 *   d_records      [n_robots, MPC_REC_STRIDE(h)] float32, device
 *   d_robot_params [n_types, MPC_ROBOT_PARAMS_STRIDE] float32, device
 *   d_robot_type   [n_robots] int32 device, or NULL (all robots type 0); an entry outside
 *                  [0, n_types) gives that robot MPC_STATUS_BAD_INDEX and zero forces (checked on
 *                  the device; the host-pointer entry point returns MPC_B200_EINVAL instead)
 *   n_types        rows of d_robot_params (>= 1)
 *   d_grf0         [n_robots, 12] float32 device: horizon-0 contact forces,
 *                  NEGATED QP solution exactly as mpc_osqp.cc:789-790 returns
 *   d_grf_full     [n_robots, 12*h] float32 device or NULL (full horizon)
 *   d_status       [n_robots] int32 device or NULL
 *   d_resid        [n_robots, 2] float32 device or NULL (primal, dual residual)
 *   d_info         [n_robots, 8] float64 device or NULL: iterations, rho updates,
 *                  final rho, polish status, pri_res, dua_res, cost scale, n_free
 *   d_warm         [n_robots, MPC_WARM_STRIDE(h)] float64 device or NULL: OSQP's persistent
 *                  workspace state (settings.warm_start = true, mpc_osqp.cc:708, :760-778).
 *                  Rows with flag 0 start cold (= osqp_setup on the first call); every solve
 *                  leaves its iterates / rho / q for the next one.  NULL = cold start always.
 *   d_order        [n_robots] int32 device or NULL: launch order (CTA b solves robot d_order[b]); must be
 *                  a permutation of 0..n_robots-1 (entries outside that range are skipped on the device).
 *                  Purely a scheduling hint - results are identical: OSQP's iteration count varies
 *                  50..350 per robot and the launch ends with its slowest robot, so a control loop
 *                  can pass last tick's robots sorted by descending iteration count
 *                  (longest-processing-time-first) to shorten the tail.  NULL = index order.
 *   stream         cudaStream_t as void* (NULL = legacy default stream)
 * Asynchronous w.r.t. the host.
 */
int mpc_b200_solve_batch(const MpcB200Settings* settings,
                         const float* d_records,
                         const float* d_robot_params,
                         const int32_t* d_robot_type,
                         int n_types,
                         int n_robots,
                         float* d_grf0,
                         float* d_grf_full,
                         int32_t* d_status,
                         float* d_resid,
                         double* d_info,
                         double* d_warm,
                         const int32_t* d_order,
                         void* stream);

/*
 * Host-buffer convenience used by the ConvexMpc compatibility class and the
 * e2e benchmark leg: copies records H2D, solves, copies forces D2H on an
 * internal stream and synchronises.  Same argument meaning as above but all
 * pointers are HOST pointers (pinned or pageable).
 */
int mpc_b200_solve_batch_host(const MpcB200Settings* settings,
                              const float* h_records,
                              const float* h_robot_params,
                              const int32_t* h_robot_type,
                              int n_types,
                              int n_robots,
                              float* h_grf0,
                              float* h_grf_full,
                              int32_t* h_status);

/* ---------------------------------------------------------------------------------------
 * Locomotion tick fused around the QP (one launch per control tick, one CTA per robot):
 * leg FK/Jacobian, gait phases + horizon table, CoM height / ground normal, foot placement,
 * swing Bezier, QP marshalling, [QP solve when the MPC is due], Cartesian PD + J^T f torque
 * map.  Replaces the per-robot Python loop
 *   reference/RL_Environment/tasks/aliengo.py:252-256 -> RobotRunnerMin.run (RobotRunnerMin.py:54-76)
 *   -> ConvexMPCLocomotion.run (ConvexMPCLocomotion.py:222-378) -> LegController.updateCommand.
 * StateEstimator.update() (StateEstimator.py:57-97, float16 rotations reproduced bit for bit)
 * runs in the same launch when `d_body` (Isaac root_states rows) is given instead of `d_est`.
 * ------------------------------------------------------------------------------------- */
/* robot geometry / default-weight row (float32), reference Quadruped.py:16-73 */
#define MPC_GEOM_ABAD     0
#define MPC_GEOM_HIP      1
#define MPC_GEOM_KNEE     2
#define MPC_GEOM_LOCX     3   /* _abadLocation */
#define MPC_GEOM_LOCY     4
#define MPC_GEOM_LOCZ     5
#define MPC_GEOM_HEIGHT   6   /* _bodyHeight */
#define MPC_GEOM_MU       7
#define MPC_GEOM_WEIGHTS  8   /* 13 default MPC weights */
#define MPC_GEOM_STRIDE  24
/* StateEstimator.update() outputs per robot (float32): vBody, omegaBody, rpyBody, ground_R_body_frame (row-major) */
#define MPC_EST_VBODY      0
#define MPC_EST_OMEGABODY  3
#define MPC_EST_RPYBODY    6
#define MPC_EST_GROUNDR    9
#define MPC_EST_STRIDE    18
/* commands per robot (float32): vx, vy, yaw rate [, 13 MPC weights]; the row stride is
 * MpcB200TickParams.cmd_stride = 3 (Isaac `commands[N,3]`, aliengo.py:233) or MPC_CMD_STRIDE */
#define MPC_CMD_STRIDE    16
/* per-robot FSM gate of the tick (d_active), see mpc_b200_tick */
#define MPC_TICK_SKIP        0   /* robot not in Locomotion: nothing read, nothing written                  */
#define MPC_TICK_RUN         1   /* run the tick (FSM_OperatingMode.TEST, ControlFSM.py:80-81)               */
#define MPC_TICK_CHECK_RUN   2   /* NORMAL mode: locomotionSafe() first; unsafe -> zero torques, state untouched
                                    except MPC_ST_UNSAFE (ControlFSM.py:84-95: the tick is not run)         */
#define MPC_TICK_CHECK_ONLY  3   /* evaluate locomotionSafe() only (a transition was requested): zero torques  */
/* persistent per-robot tick state (float32 row; the first six slots hold int32 bit patterns) */
#define MPC_ST_COUNTER     0   /* int  iterationCounter                         */
#define MPC_ST_FIRSTRUN    1   /* int                                           */
#define MPC_ST_FIRSTSWING  2   /* int[4]                                        */
#define MPC_ST_SWINGREM    6   /* [4]  swingTimeRemaining                       */
#define MPC_ST_P0         10   /* [4][3] swing start                            */
#define MPC_ST_PF         22   /* [4][3] swing target                           */
#define MPC_ST_TRAJP      34   /* [4][3] last swing position                    */
#define MPC_ST_TRAJV      46   /* [4][3] last swing velocity                    */
#define MPC_ST_FFF        58   /* [4][3] f_ff of the last MPC solve             */
#define MPC_ST_POSZ       70   /* estimator CoM height                          */
#define MPC_ST_CPHASE     71   /* [4]  estimator contact phase                  */
#define MPC_ST_HIST       75   /* [4][3] foot contact history                   */
#define MPC_ST_NORMAL     87   /* [3]  ground normal (yaw frame)                */
#define MPC_ST_UNSAFE     90   /* int  locomotionSafe() == False this tick (FSM_State_Locomotion.py:104-136);
                                        only evaluated when d_body is given                */
#define MPC_ST_ITERS      91   /* int  OSQP iterations of this robot's last MPC solve (scheduling hint for
                                        d_order, see mpc_b200_tick)                        */
#define MPC_TICK_STATE_STRIDE 96

typedef struct MpcB200TickParams {
  float dt;                /* Parameters.controller_dt (0.01)                       */
  int   iters_between_mpc; /* int(27/(1000*dt)) = 2 (RobotRunnerMin.py:21-22)        */
  int   flat_ground;       /* Parameters.flat_ground                                */
  float bonus_swing;       /* Parameters.cmpc_bonus_swing                           */
  int   cmd_stride;        /* floats per row of d_cmd: 3 or MPC_CMD_STRIDE (0 = MPC_CMD_STRIDE) */
  int   reserved[3];
  float act_scale[12];     /* Parameters.MPC_param_scale (Parameters.py:25-28)      */
  float act_const[12];     /* Parameters.MPC_param_const (Parameters.py:30-33)      */
} MpcB200TickParams;

/*
 * One control tick for n_robots robots (device pointers, asynchronous on `stream`).
 *   d_dof    [n, 12, 2] float32: joint position / velocity interleaved (Isaac dof_state layout)
 *   d_body   [n, 13] float32 root states (pos3, quat xyzw, lin vel3, ang vel3; aliengo.py:248)
 *            or NULL; when given, StateEstimator.update runs on the device
 *   d_est    [n, MPC_EST_STRIDE] float32 or NULL: precomputed StateEstimator.update outputs
 *            (exactly one of d_body / d_est must be non-NULL)
 *   d_cmd    [n, tick->cmd_stride] float32: vx, vy, yaw rate (+ 13 weights when the stride is 16)
 *   has_weights != 0 (needs cmd_stride 16, no d_actions): MPC weights from d_cmd[:, 3:16]
 *   d_actions [n, 12] float32 policy actions in [-1, 1] or NULL.  When given, the MPC weights are
 *            actions * tick->act_scale + tick->act_const (float32 multiply then add, 13th weight 0) -
 *            the rescale of pre_physics_step (aliengo.py:237-245, 253) done on the device.
 *            Neither d_actions nor has_weights: the robot's default weights (Quadruped.py:30-34).
 *   d_gait   [n] int32 gait ids (ConvexMPCLocomotion.py:227-241) or NULL (trot)
 *   d_robot_params [n_types, 12], d_robot_geom [n_types, MPC_GEOM_STRIDE], d_robot_type [n] or NULL
 *            (entries outside [0, n_types): robot skipped, d_solved = MPC_STATUS_BAD_INDEX, zero torques)
 *   d_active [n] int32 MPC_TICK_* or NULL (all MPC_TICK_RUN): the FSM shell's per-robot gate
 *            (RobotRunnerFSM.py:55-69 -> ControlFSM.runFSM, ControlFSM.py:78-104).  The safety check
 *            needs d_body (the world roll / pitch come from StateEstimator.update).
 *   d_state  [n, MPC_TICK_STATE_STRIDE] float32, read and written
 *   d_torques [n, 12] float32 out
 *   d_grf0   [n, 12] float32 out or NULL: written for robots whose MPC fired this tick
 *   d_records[n, MPC_REC_STRIDE(h)] float32 out or NULL: the QP record built this tick (tests)
 *   d_solved [n] int32 out or NULL: OSQP status when the MPC fired, 0 otherwise; a robot whose
 *            solve did not end in MPC_STATUS_SOLVED keeps the feed-forward forces of its previous
 *            MPC tick (hold-last fallback; d_grf0 still shows what the failed solve produced).
 *            MPC_STATUS_POLISH_OVERFLOW (contact tables with > 2h stance leg-steps, which none of
 *            the reference's gaits produces - tests/test_host.py checks every gait) can appear here:
 *            the tick has no overflow pass, it applies the converged but UN-polished ADMM solution
 *            and leaves d_warm untouched for that robot.
 *   d_warm   [n, MPC_WARM_STRIDE(h)] float64 in/out or NULL: see mpc_b200_solve_batch
 *   d_order  [n] int32 or NULL: launch order (CTA b runs robot d_order[b]), a pure scheduling
 *            hint as in mpc_b200_solve_batch; consecutive MPC solves of a robot need similar
 *            iteration counts, so sorting the robots by descending MPC_ST_ITERS of the previous
 *            MPC tick shortens the tail of the launch
 */
int mpc_b200_tick(const MpcB200Settings* settings, const MpcB200TickParams* tick,
                  const float* d_dof, const float* d_body, const float* d_est, const float* d_cmd,
                  int has_weights, const float* d_actions,
                  const int32_t* d_gait, const float* d_robot_params, const float* d_robot_geom,
                  const int32_t* d_robot_type, int n_types, int n_robots, const int32_t* d_active,
                  float* d_state, float* d_torques,
                  float* d_grf0, float* d_records, int32_t* d_solved, double* d_warm,
                  const int32_t* d_order, void* stream);

/* Debug / test hook: StateEstimator.update (StateEstimator.py:57-97) alone on the device, exactly the
 * function the tick kernel runs when d_body is given.  h_body [n,13] root states, h_normal [n,3] the
 * previous tick's ground normal (yaw frame), h_est [n, MPC_EST_STRIDE] out.  Host pointers; synchronises.
 * The parity tests compare rpyBody and ground_R_body_frame (float16 in the reference) BIT FOR BIT.      */
int mpc_b200_debug_estimator(const float* h_body, const float* h_normal, float* h_est, int n);

/* Debug / test hook: run ONE record through the solve kernel and copy back the
 * quantities the kernel never materialises in production: the dense Hessian
 * P[n*n] (row-major, n = 12h) and gradient q[n] of the condensed QP
 * (mpc_osqp.cc:678-683) evaluated from the on-chip closed form, and OSQP's
 * Ruiz scaling D[n], E[20h], c.  Used by the parity tests to compare the
 * on-device builder with the oracle builder entry by entry.  Host pointers;
 * synchronises.                                                             */
int mpc_b200_debug_build_qp(const MpcB200Settings* settings,
                            const float* h_record,
                            const float* h_robot_params,
                            double* h_P, double* h_q,
                            double* h_D, double* h_E, double* h_c);

#ifdef __cplusplus
}
#endif
#endif /* MPC_B200_H */
