// mpc_osqp_pybind.cc — TEST INFRASTRUCTURE ONLY.
//
// A Python extension NAMED `mpc_osqp` with the reference module's exact surface
// (reference/MPC_Controller/convex_MPC/mpc_osqp.cc:950-983: class ConvexMpc with the
// 7-argument ctor, compute_contact_forces(13 sequences) -> list, reset_solver(), enum
// QPSolverName{OSQP,QPOASES} exported, __version__, TEST == 42) backed by the CPU oracle
// (mpc_oracle.cc + the unmodified vendored OSQP).  It exists so that the UNMODIFIED
// reference Python (`/root/reference/MPC_Controller/**`) can be imported in this
// container to generate golden tick fixtures (tests/golden/make_golden.py) and to run
// BASELINE config 1.  qpOASES is absent from the reference checkout, so both enum
// values route to OSQP (same unique minimiser: the QP is strictly convex).
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdlib>
#include <vector>

#include "mpc_oracle.h"

namespace py = pybind11;

enum QPSolverName { OSQP, QPOASES };

class ConvexMpc {
 public:
  ConvexMpc(double mass, const std::vector<double>& inertia, int num_legs,
            int planning_horizon, double timestep, double alpha, QPSolverName)
      : horizon_(planning_horizon), legs_(num_legs) {
    o_ = mpc_oracle_create(mass, inertia.data(), num_legs, planning_horizon, timestep, alpha);
    const char* m = std::getenv("MPC_ORACLE_MODE");
    mode_ = (m && m[0] == 'c') ? MPC_ORACLE_COLD : MPC_ORACLE_WARM;
  }
  ~ConvexMpc() { mpc_oracle_destroy(o_); }
  ConvexMpc(const ConvexMpc&) = delete;

  std::vector<double> ComputeContactForces(
      const std::vector<double>& qp_weights, std::vector<double> com_position,
      std::vector<double> com_velocity, std::vector<double> com_roll_pitch_yaw,
      std::vector<double> ground_normal_vec, std::vector<double> com_angular_velocity,
      std::vector<double> foot_contact_states, std::vector<double> foot_positions_body_frame,
      std::vector<double> foot_friction_coeffs, std::vector<double> desired_com_position,
      std::vector<double> desired_com_velocity, std::vector<double> desired_com_roll_pitch_yaw,
      std::vector<double> desired_com_angular_velocity) {
    std::vector<double> out(3 * legs_ * horizon_, 0.0);
    const int st = mpc_oracle_compute_contact_forces(
        o_, mode_, qp_weights.data(), com_position.data(), com_velocity.data(),
        com_roll_pitch_yaw.data(), ground_normal_vec.data(), com_angular_velocity.data(),
        foot_contact_states.data(), foot_positions_body_frame.data(),
        foot_friction_coeffs.data(), desired_com_position.data(), desired_com_velocity.data(),
        desired_com_roll_pitch_yaw.data(), desired_com_angular_velocity.data(), out.data());
    if (st != 1) return {};  // mpc_osqp.cc:791-794: empty vector on failure
    return out;
  }
  void ResetSolver() { mpc_oracle_reset_solver(o_); }

 private:
  MpcOracle* o_;
  int horizon_, legs_, mode_;
};

PYBIND11_MODULE(mpc_osqp, m) {
  m.doc() = "CPU oracle stand-in for the reference mpc_osqp module (test infrastructure)";
  py::enum_<QPSolverName>(m, "QPSolverName")
      .value("OSQP", OSQP, "OSQP")
      .value("QPOASES", QPOASES, "QPOASES")
      .export_values();
  py::class_<ConvexMpc>(m, "ConvexMpc")
      .def(py::init<double, const std::vector<double>&, int, int, double, double, QPSolverName>())
      .def("compute_contact_forces", &ConvexMpc::ComputeContactForces)
      .def("reset_solver", &ConvexMpc::ResetSolver);
  m.attr("__version__") = "oracle";
  m.attr("TEST") = py::int_(int(42));
}
