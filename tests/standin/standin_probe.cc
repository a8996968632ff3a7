// Exercises oracle/eigen_standin (the stand-in for the Eigen subset the reference's mpc_osqp.cc uses) on fixed
// inputs and prints the results as JSON; tests/test_eigen_standin.py compares them with NumPy / SciPy.
// TEST INFRASTRUCTURE.  Build: g++ -std=c++17 -Ioracle/eigen_standin tests/standin/standin_probe.cc
#include <Eigen/Dense>
#include <Eigen/SparseCore>
#include <unsupported/Eigen/MatrixFunctions>
#include <cstdio>
#include <vector>

using namespace Eigen;

static void dump(const char* name, const Ref<double>& m, bool last = false) {
  std::printf("\"%s\": [", name);
  for (int i = 0; i < m.rows(); ++i) {
    std::printf("[");
    for (int j = 0; j < m.cols(); ++j) std::printf("%.17g%s", m(i, j), j + 1 < m.cols() ? "," : "");
    std::printf("]%s", i + 1 < m.rows() ? "," : "");
  }
  std::printf("]%s\n", last ? "" : ",");
}

int main() {
  std::printf("{\n");
  // deterministic pseudo-random fill
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffffff) / double(0x1000000) - 0.5; };
  MatrixXd A(7, 5), B(5, 6);
  for (int i = 0; i < 7; ++i) for (int j = 0; j < 5; ++j) A(i, j) = rnd();
  for (int i = 0; i < 5; ++i) for (int j = 0; j < 6; ++j) B(i, j) = rnd();
  dump("A", A); dump("B", B);
  dump("AB", A * B);
  dump("ABt_scaled", 2 * (A * B).transpose() - (A * B).transpose() * 0.25);
  // blocks: lvalue assignment, +=, transpose of a block, comma initialiser on a block, row <- 3-vector
  MatrixXd C(6, 6);
  C.setZero();
  C.block<2, 3>(1, 2) = A.block(3, 1, 2, 3);
  C.block(4, 0, 2, 2) += A.block<2, 2>(0, 0).transpose();
  C.block<2, 2>(0, 0) << 1, 2.5, -3, 4;
  Vector3d v3(0.3, -1.2, 2.0);
  MatrixXd F(2, 3);
  F.setZero();
  F.row(1) = v3;
  dump("C", C); dump("F", F);
  // fixed-size from a coefficient array (column-major read), inverse, Map views
  const double raw[9] = {2, 0.1, 0.3, -0.2, 1.5, 0.4, 0.05, -0.6, 3};
  const Matrix3d M3(raw);
  dump("M3", M3); dump("M3inv", M3.inverse());
  std::vector<double> buf = {1, 2, 3, 4, 5, 6};
  const MatrixXd Mt = Map<const MatrixXd>(buf.data(), 2, 3).transpose();
  dump("Mt", Mt);
  std::vector<double> out(4, 0.0);
  Map<VectorXd> ov(out.data(), 4);
  const std::vector<double> neg = {1, -2, 3, -4};
  ov = -Map<const VectorXd>(neg.data(), 4);
  dump("ov", ov);
  // replicate / asDiagonal / cwiseMax / cwiseMin
  const std::vector<double> w = {1, 2, 3};
  const Map<const VectorXd> wv(w.data(), 3);
  const MatrixXd D = wv.replicate(2, 1).asDiagonal();
  dump("D", D);
  dump("clip", wv.cwiseMax(1.5).cwiseMin(2.5));
  // geometry: yaw * pitch * roll -> quaternion -> matrix, rotation of a row vector
  const double rpy[3] = {0.31, -0.22, 1.05};
  const AngleAxisd roll(rpy[0], Vector3d::UnitX()), pitch(rpy[1], Vector3d::UnitY()), yaw(rpy[2], Vector3d::UnitZ());
  const Quaterniond q = yaw * pitch * roll;
  dump("Rzyx", q.matrix());
  const Quaterniond q2 = AngleAxisd(rpy[0], Vector3d::UnitX()) * AngleAxisd(rpy[1], Vector3d::UnitY()) * AngleAxisd(rpy[2], Vector3d::UnitZ());
  MatrixXd P(1, 3);
  P.row(0) = q2 * F.row(1);
  dump("rot_xyz_v", P);
  // matrix exponential: small norm (Pade 3/5), mid (7/9), large (13 + squarings), and the nilpotent MPC block
  MatrixXd E1(4, 4), E2(4, 4), E3(4, 4);
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { const double r = rnd(); E1(i, j) = 0.01 * r; E2(i, j) = 1.2 * r; E3(i, j) = 9.0 * r; }
  dump("E1", E1); dump("E2", E2); dump("E3", E3);
  dump("expE1", E1.exp()); dump("expE2", E2.exp()); dump("expE3", E3.exp());
  // sparse: sparseView keeps the non-zeros (column-major), triangularView<Upper> keeps row <= col
  MatrixXd S(4, 4);
  S.setZero();
  S(0, 0) = 1; S(2, 0) = -2; S(1, 1) = 3; S(0, 2) = 4; S(3, 2) = 5; S(3, 3) = 1e-300; S(1, 3) = 7;
  const SparseMatrix<double, ColMajor, long long> sp = S.sparseView();
  const SparseMatrix<double, ColMajor, long long> up = sp.triangularView<Upper>();
  dump("S", S);
  auto dump_csc = [](const char* name, const SparseMatrix<double, ColMajor, long long>& m) {
    std::printf("\"%s\": {\"indptr\": [", name);
    for (int c = 0; c <= m.cols(); ++c) std::printf("%lld%s", m.outerIndexPtr()[c], c < m.cols() ? "," : "");
    std::printf("], \"indices\": [");
    for (long long k = 0; k < m.nonZeros(); ++k) std::printf("%lld%s", m.innerIndexPtr()[k], k + 1 < m.nonZeros() ? "," : "");
    std::printf("], \"data\": [");
    for (long long k = 0; k < m.nonZeros(); ++k) std::printf("%.17g%s", m.valuePtr()[k], k + 1 < m.nonZeros() ? "," : "");
    std::printf("]},\n");
  };
  dump_csc("S_csc", sp);
  dump_csc("S_upper_csc", up);
  dump("I23", MatrixXd::Identity(2, 3), true);
  std::printf("}\n");
  return 0;
}
