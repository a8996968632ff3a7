/*
 * mpc_oracle.h — TEST INFRASTRUCTURE ONLY (the checker, never the product).
 *
 * CPU oracle for the convex-MPC stance-force QP of silvery107/rl-mpc-locomotion.
 *   - QP builder: an Eigen-free C++ restatement of
 *       reference/MPC_Controller/convex_MPC/mpc_osqp.cc:283-688
 *     (bug-for-bug, see SURVEY.md App. B: truncated a_qp, Rx*Ry*Rz foot rotation,
 *     friction-index misuse, negated output).
 *   - QP solver: the reference's own vendored OSQP 0.6.0, compiled UNMODIFIED from
 *     /root/reference/extern/osqp into oracle/_ref/libosqp_ref.so (oracle/Makefile),
 *     driven with the reference's exact settings and glue (mpc_osqp.cc:690-796).
 *
 * PARITY STATUS: pinned to the reference's own source.  The reference holds no golden
 * vectors / tests for this path and vendors Eigen, qpOASES and pybind11 as git submodules
 * whose contents are absent from /root/reference (the image has no Eigen either), so
 * mpc_osqp.cc cannot be built against its real dependencies here.  It IS built, unmodified
 * and from where it lies, against oracle/eigen_standin/ (our restatement of the dense Eigen
 * subset it uses, qpOASES declarations, an empty pybind11/eigen.h) and the unmodified OSQP:
 * oracle/_ref/refmod/mpc_osqp*.so (oracle/Makefile).  tests/test_reference_source.py holds
 * this restatement, the golden vectors and the warm-start path to that build (2.5e-8 N on
 * the bench distributions) and replays the unmodified reference Python through it.  What
 * remains ours in that build: dense products, the 3x3 inverse, the quaternion formulas and
 * the Pade matrix exponential underneath the reference's lines.  The builder is further
 * cross-checked against two independent NumPy/SciPy transcriptions (oracle/np_builder.py,
 * scipy.linalg.expm) in tests/test_oracle.py.  The *solver* half is the real reference code.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 */
#ifndef MPC_ORACLE_H
#define MPC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct MpcOracle MpcOracle;

/* mirrors ConvexMpc::ConvexMpc (mpc_osqp.cc:508-574); inertia row-major 3x3 */
MpcOracle* mpc_oracle_create(double mass, const double* inertia9, int num_legs,
                             int planning_horizon, double timestep, double alpha);
void mpc_oracle_destroy(MpcOracle* o);
/* mirrors ConvexMpc::ResetSolver (mpc_osqp.cc:576) */
void mpc_oracle_reset_solver(MpcOracle* o);

/* solver modes */
#define MPC_ORACLE_COLD 0 /* fresh OSQP workspace every call (golden vectors)      */
#define MPC_ORACLE_WARM 1 /* reference behaviour: setup once, then osqp_update_*  */
#define MPC_ORACLE_TRUE 2 /* fresh workspace, eps 1e-10, polish: "true" minimiser  */

/*
 * mirrors ConvexMpc::ComputeContactForces (mpc_osqp.cc:578-796, OSQP branch).
 * All 13 arguments in the reference order, as doubles.  out_forces must hold
 * 3*num_legs*horizon doubles and receives the NEGATED QP solution
 * (mpc_osqp.cc:789-790).  Returns the OSQP status_val (1 == OSQP_SOLVED); on
 * any other status out_forces is untouched (the reference returns an empty
 * vector, mpc_osqp.cc:791-794).
 */
int mpc_oracle_compute_contact_forces(
    MpcOracle* o, int mode, const double* qp_weights13, const double* com_position,
    const double* com_velocity, const double* com_roll_pitch_yaw,
    const double* ground_normal_vec, const double* com_angular_velocity,
    const double* foot_contact_states, const double* foot_positions_body_frame,
    const double* foot_friction_coeffs, const double* desired_com_position,
    const double* desired_com_velocity, const double* desired_com_roll_pitch_yaw,
    const double* desired_com_angular_velocity, double* out_forces);

/* Only build the QP (mpc_osqp.cc:593-692) without solving. */
int mpc_oracle_build_qp(
    MpcOracle* o, const double* qp_weights13, const double* com_position,
    const double* com_velocity, const double* com_roll_pitch_yaw,
    const double* ground_normal_vec, const double* com_angular_velocity,
    const double* foot_contact_states, const double* foot_positions_body_frame,
    const double* foot_friction_coeffs, const double* desired_com_position,
    const double* desired_com_velocity, const double* desired_com_roll_pitch_yaw,
    const double* desired_com_angular_velocity);

/* Copy out the last built QP: P[n*n] row-major, q[n], C[m*n] row-major dense,
 * l[m], u[m]; n = 3*legs*h, m = 5*legs*h.  Any pointer may be NULL. */
void mpc_oracle_get_qp(const MpcOracle* o, double* P, double* q, double* C,
                       double* l, double* u);
/* Intermediate matrices for builder tests: a_exp[13*13], b_exp[13*12] row-major. */
void mpc_oracle_get_exponentials(const MpcOracle* o, double* a_exp, double* b_exp);

/* info of the last solve */
int    mpc_oracle_last_iters(const MpcOracle* o);
int    mpc_oracle_last_polish_status(const MpcOracle* o);
double mpc_oracle_last_solve_time(const MpcOracle* o); /* setup+update+solve+polish, s */

/*
 * Batch driver over packed float32 problem records (layout: include/mpc_b200.h).
 * Every record is promoted to double and pushed through ONE MpcOracle per call
 * (cold: fresh workspace per QP; warm: the reference's persistent-workspace path).
 *   robot_params: [n_types][12] float32 (mass, inertia[9], pad2); robot_type may be NULL
 *   grf_full:     [n][12*h] doubles or NULL;  grf0: [n][12] doubles or NULL
 *   status:       [n] ints or NULL
 * Returns the number of QPs whose status != OSQP_SOLVED.
 */
int mpc_oracle_solve_records(const float* records, int n, int horizon, double timestep,
                             double alpha, const float* robot_params,
                             const int* robot_type, int mode, double* grf0,
                             double* grf_full, int* status);

#ifdef __cplusplus
}
#endif
#endif
