// pybind11/eigen.h stand-in: the reference's PYBIND11_MODULE binds std::vector<double> signatures only,
// no Eigen type crosses the Python boundary, so the real header's type casters are not needed (and could
// not compile without Eigen).  TEST INFRASTRUCTURE (see Eigen/Core in this directory)
#pragma once
#include <pybind11/pybind11.h>
