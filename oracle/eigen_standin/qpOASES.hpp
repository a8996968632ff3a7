// qpOASES.hpp stand-in: the reference vendors qpOASES as a git submodule (extern/qpoases) that is absent
// from /root/reference.  SURVEY §2.1: the QPOASES branch of mpc_osqp.cc is out of scope (source absent,
// OSQP is what the reference runs); these declarations only let the UNMODIFIED file compile.  Selecting
// QPSolverName::QPOASES on the compiled module fails: init() reports an error and the solution is zero.
// TEST INFRASTRUCTURE (see Eigen/Core in this directory)
#pragma once
#include "qpOASES/Types.hpp"

namespace qpOASES {
struct Options {
  void setToMPC() {}
  PrintLevel printLevel = PL_NONE;
};
class QProblem {
 public:
  QProblem(int_t, int_t) {}
  void setOptions(const Options&) {}
  int init(const real_t*, const real_t*, const real_t*, const real_t*, const real_t*, const real_t*,
           const real_t*, int_t&) { return 1; }
  int getPrimalSolution(real_t*) const { return 1; }
};
}  // namespace qpOASES
