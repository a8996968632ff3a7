// mpc_oracle.cc — TEST INFRASTRUCTURE ONLY (see mpc_oracle.h for the parity status).
//
// Eigen-free restatement of the reference QP builder
//   reference/MPC_Controller/convex_MPC/mpc_osqp.cc:283-688
// plus the reference's OSQP glue (mpc_osqp.cc:690-796) driving the UNMODIFIED
// vendored OSQP 0.6.0 (oracle/_ref/libosqp_ref.so, built by oracle/Makefile from
// /root/reference/extern/osqp).  Dense matrices are tiny row-major arrays; clarity
// beats speed here.  Every function cites the reference lines it follows.
//
// Third-party arithmetic restated because the dependency is absent from
// /root/reference: Eigen 3.3 `MatrixBase::exp()` (unsupported/Eigen/MatrixFunctions,
// submodule extern/eigen3 @ branch 3.3, empty here), called at mpc_osqp.cc:348.
// Its published algorithm (Higham 2005, "The scaling and squaring method for the
// matrix exponential revisited": Pade 3/5/7/9/13 chosen by the 1-norm, then
// repeated squaring) is restated in expm_pade() below and cross-checked against
// scipy.linalg.expm in tests/test_oracle.py.

#include "mpc_oracle.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {
#include "ctrlc.h"
#include "osqp.h"
}

#include "../include/mpc_b200.h"  // record layout only

namespace {

constexpr int k3Dim = 3;            // mpc_osqp.cc:53
constexpr double kGravity = 9.8;    // mpc_osqp.cc:54
constexpr double kMaxScale = 10;    // mpc_osqp.cc:55
constexpr double kMinScale = 0.1;   // mpc_osqp.cc:56
constexpr int kStateDim = 13;       // mpc_osqp.cc:180-181
constexpr int kConstraintDim = 5;   // mpc_osqp.cc:184

// ---------------------------------------------------------------- tiny dense
struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
  void set_zero() { std::fill(a.begin(), a.end(), 0.0); }
};

Mat eye(int n) {
  Mat m(n, n);
  for (int i = 0; i < n; ++i) m(i, i) = 1.0;
  return m;
}

Mat mul(const Mat& x, const Mat& y) {
  assert(x.c == y.r);
  Mat z(x.r, y.c);
  for (int i = 0; i < x.r; ++i)
    for (int k = 0; k < x.c; ++k) {
      const double v = x(i, k);
      if (v == 0.0) continue;
      for (int j = 0; j < y.c; ++j) z(i, j) += v * y(k, j);
    }
  return z;
}

Mat transpose(const Mat& x) {
  Mat z(x.c, x.r);
  for (int i = 0; i < x.r; ++i)
    for (int j = 0; j < x.c; ++j) z(j, i) = x(i, j);
  return z;
}

Mat block(const Mat& x, int i0, int j0, int nr, int nc) {
  Mat z(nr, nc);
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j) z(i, j) = x(i0 + i, j0 + j);
  return z;
}

void set_block(Mat& x, int i0, int j0, const Mat& b) {
  for (int i = 0; i < b.r; ++i)
    for (int j = 0; j < b.c; ++j) x(i0 + i, j0 + j) = b(i, j);
}

Mat add(const Mat& x, const Mat& y) {
  Mat z(x.r, x.c);
  for (size_t i = 0; i < x.a.size(); ++i) z.a[i] = x.a[i] + y.a[i];
  return z;
}

Mat axpby(double a, const Mat& x, double b, const Mat& y) {
  Mat z(x.r, x.c);
  for (size_t i = 0; i < x.a.size(); ++i) z.a[i] = a * x.a[i] + b * y.a[i];
  return z;
}

double norm1(const Mat& x) {  // max absolute column sum
  double best = 0;
  for (int j = 0; j < x.c; ++j) {
    double s = 0;
    for (int i = 0; i < x.r; ++i) s += std::fabs(x(i, j));
    best = std::max(best, s);
  }
  return best;
}

// Solve A X = B by LU with partial pivoting (dense, small).
Mat lu_solve(Mat A, Mat B) {
  const int n = A.r;
  for (int k = 0; k < n; ++k) {
    int p = k;
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(A(i, k)) > std::fabs(A(p, k))) p = i;
    if (p != k) {
      for (int j = 0; j < n; ++j) std::swap(A(k, j), A(p, j));
      for (int j = 0; j < B.c; ++j) std::swap(B(k, j), B(p, j));
    }
    const double piv = A(k, k);
    for (int i = k + 1; i < n; ++i) {
      const double f = A(i, k) / piv;
      if (f == 0.0) continue;
      for (int j = k; j < n; ++j) A(i, j) -= f * A(k, j);
      for (int j = 0; j < B.c; ++j) B(i, j) -= f * B(k, j);
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    for (int j = 0; j < B.c; ++j) {
      double s = B(k, j);
      for (int i = k + 1; i < n; ++i) s -= A(k, i) * B(i, j);
      B(k, j) = s / A(k, k);
    }
  }
  return B;
}

// Scaling-and-squaring Pade matrix exponential (Higham 2005) — the published
// algorithm behind Eigen 3.3's MatrixBase::exp() that mpc_osqp.cc:348 calls.
Mat expm_pade(const Mat& Ain) {
  const int n = Ain.r;
  const Mat I = eye(n);
  Mat A = Ain;
  const double l1 = norm1(A);
  Mat U, V;
  int squarings = 0;
  auto lin = [&](std::initializer_list<std::pair<double, const Mat*>> terms) {
    Mat z(n, n);
    for (auto& t : terms)
      for (size_t i = 0; i < z.a.size(); ++i) z.a[i] += t.first * t.second->a[i];
    return z;
  };
  if (l1 < 1.495585217958292e-2) {
    const double b[] = {120., 60., 12., 1.};
    Mat A2 = mul(A, A);
    U = mul(A, lin({{b[3], &A2}, {b[1], &I}}));
    V = lin({{b[2], &A2}, {b[0], &I}});
  } else if (l1 < 2.539398330063230e-1) {
    const double b[] = {30240., 15120., 3360., 420., 30., 1.};
    Mat A2 = mul(A, A), A4 = mul(A2, A2);
    U = mul(A, lin({{b[5], &A4}, {b[3], &A2}, {b[1], &I}}));
    V = lin({{b[4], &A4}, {b[2], &A2}, {b[0], &I}});
  } else if (l1 < 9.504178996162932e-1) {
    const double b[] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
    Mat A2 = mul(A, A), A4 = mul(A2, A2), A6 = mul(A4, A2);
    U = mul(A, lin({{b[7], &A6}, {b[5], &A4}, {b[3], &A2}, {b[1], &I}}));
    V = lin({{b[6], &A6}, {b[4], &A4}, {b[2], &A2}, {b[0], &I}});
  } else if (l1 < 2.097847961257068e+0) {
    const double b[] = {17643225600., 8821612800., 2075673600., 302702400., 30270240.,
                        2162160.,     110880.,     3960.,       90.,        1.};
    Mat A2 = mul(A, A), A4 = mul(A2, A2), A6 = mul(A4, A2), A8 = mul(A6, A2);
    U = mul(A, lin({{b[9], &A8}, {b[7], &A6}, {b[5], &A4}, {b[3], &A2}, {b[1], &I}}));
    V = lin({{b[8], &A8}, {b[6], &A6}, {b[4], &A4}, {b[2], &A2}, {b[0], &I}});
  } else {
    const double maxnorm = 5.371920351148152;
    int e = 0;
    std::frexp(l1 / maxnorm, &e);
    squarings = std::max(0, e);
    const double scale = std::ldexp(1.0, -squarings);
    for (auto& v : A.a) v *= scale;
    const double b[] = {64764752532480000., 32382376266240000., 7771770303897600.,
                        1187353796428800.,  129060195264000.,   10559470521600.,
                        670442572800.,      33522128640.,       1323241920.,
                        40840800.,          960960.,            16380.,
                        182.,               1.};
    Mat A2 = mul(A, A), A4 = mul(A2, A2), A6 = mul(A4, A2);
    Mat t = mul(A6, lin({{b[13], &A6}, {b[11], &A4}, {b[9], &A2}}));
    t = add(t, lin({{b[7], &A6}, {b[5], &A4}, {b[3], &A2}, {b[1], &I}}));
    U = mul(A, t);
    Mat v = mul(A6, lin({{b[12], &A6}, {b[10], &A4}, {b[8], &A2}}));
    V = add(v, lin({{b[6], &A6}, {b[4], &A4}, {b[2], &A2}, {b[0], &I}}));
  }
  Mat R = lu_solve(axpby(1.0, V, -1.0, U), add(V, U));
  for (int s = 0; s < squarings; ++s) R = mul(R, R);
  return R;
}

// ------------------------------------------------ 3x3 helpers (Eigen Geometry)
Mat rot_x(double t) {
  Mat m = eye(3);
  m(1, 1) = std::cos(t); m(1, 2) = -std::sin(t);
  m(2, 1) = std::sin(t); m(2, 2) = std::cos(t);
  return m;
}
Mat rot_y(double t) {
  Mat m = eye(3);
  m(0, 0) = std::cos(t); m(0, 2) = std::sin(t);
  m(2, 0) = -std::sin(t); m(2, 2) = std::cos(t);
  return m;
}
Mat rot_z(double t) {
  Mat m = eye(3);
  m(0, 0) = std::cos(t); m(0, 1) = -std::sin(t);
  m(1, 0) = std::sin(t); m(1, 1) = std::cos(t);
  return m;
}

// mpc_osqp.cc:283-291 — q = yaw * pitch * roll ; q.matrix()
Mat ConvertRpyToRot(const double* rpy) {
  return mul(rot_z(rpy[2]), mul(rot_y(rpy[1]), rot_x(rpy[0])));
}

// mpc_osqp.cc:293-297
Mat ConvertToSkewSymmetric(const double* v) {
  Mat s(3, 3);
  s(0, 1) = -v[2]; s(0, 2) = v[1];
  s(1, 0) = v[2];  s(1, 2) = -v[0];
  s(2, 0) = -v[1]; s(2, 1) = v[0];
  return s;
}

Mat inverse3(const Mat& m) {
  const double a = m(0, 0), b = m(0, 1), c = m(0, 2);
  const double d = m(1, 0), e = m(1, 1), f = m(1, 2);
  const double g = m(2, 0), h = m(2, 1), i = m(2, 2);
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  Mat r(3, 3);
  r(0, 0) = (e * i - f * h) / det; r(0, 1) = (c * h - b * i) / det; r(0, 2) = (b * f - c * e) / det;
  r(1, 0) = (f * g - d * i) / det; r(1, 1) = (a * i - c * g) / det; r(1, 2) = (c * d - a * f) / det;
  r(2, 0) = (d * h - e * g) / det; r(2, 1) = (b * g - a * h) / det; r(2, 2) = (a * e - b * d) / det;
  return r;
}

// mpc_osqp.cc:299-322
void CalculateAMat(const double* rpy, const double* ground_normal_vec, Mat* a_mat_ptr) {
  const double cos_yaw = std::cos(rpy[2]);
  const double sin_yaw = std::sin(rpy[2]);
  const double cos_pitch = std::cos(rpy[1]);
  const double tan_pitch = std::tan(rpy[1]);
  Mat& a_mat = *a_mat_ptr;
  a_mat(0, 6) = cos_yaw / cos_pitch; a_mat(0, 7) = sin_yaw / cos_pitch; a_mat(0, 8) = 0;
  a_mat(1, 6) = -sin_yaw;            a_mat(1, 7) = cos_yaw;             a_mat(1, 8) = 0;
  a_mat(2, 6) = cos_yaw * tan_pitch; a_mat(2, 7) = sin_yaw * tan_pitch; a_mat(2, 8) = 1;
  a_mat(3, 9) = 1;
  a_mat(4, 10) = 1;
  a_mat(5, 11) = 1;
  a_mat(9, 12) = ground_normal_vec[0];
  a_mat(10, 12) = ground_normal_vec[1];
  a_mat(11, 12) = ground_normal_vec[2];
}

// mpc_osqp.cc:324-336
void CalculateBMat(double inv_mass, const Mat& inv_inertia, const Mat& foot_positions,
                   Mat* b_mat_ptr) {
  const int num_legs = foot_positions.r;
  Mat& b_mat = *b_mat_ptr;
  for (int i = 0; i < num_legs; ++i) {
    const double fp[3] = {foot_positions(i, 0), foot_positions(i, 1), foot_positions(i, 2)};
    set_block(b_mat, 6, i * k3Dim, mul(inv_inertia, ConvertToSkewSymmetric(fp)));
    b_mat(9, i * k3Dim) = inv_mass;
    b_mat(10, i * k3Dim + 1) = inv_mass;
    b_mat(11, i * k3Dim + 2) = inv_mass;
  }
}

// mpc_osqp.cc:338-351
void CalculateExponentials(const Mat& a_mat, const Mat& b_mat, double timestep,
                           Mat* ab_mat_ptr, Mat* a_exp_ptr, Mat* b_exp_ptr) {
  Mat& ab_mat = *ab_mat_ptr;
  const int action_dim = b_mat.c;
  for (int i = 0; i < kStateDim; ++i) {
    for (int j = 0; j < kStateDim; ++j) ab_mat(i, j) = a_mat(i, j) * timestep;
    for (int j = 0; j < action_dim; ++j) ab_mat(i, kStateDim + j) = b_mat(i, j) * timestep;
  }
  Mat ab_exp = expm_pade(ab_mat);
  *a_exp_ptr = block(ab_exp, 0, 0, kStateDim, kStateDim);
  *b_exp_ptr = block(ab_exp, 0, kStateDim, kStateDim, action_dim);
}

// mpc_osqp.cc:353-435 (including the `i < horizon - 1` loop bound at :361 that
// leaves the last a_qp block zero — SURVEY App. B Q1)
void CalculateQpMats(const Mat& a_exp, const Mat& b_exp, const Mat& qp_weights_single,
                     const Mat& alpha_single, int horizon, Mat* a_qp_ptr, Mat* anb_aux_ptr,
                     Mat* b_qp_ptr, Mat* p_mat_ptr) {
  const int state_dim = kStateDim;
  Mat& a_qp = *a_qp_ptr;
  set_block(a_qp, 0, 0, a_exp);
  for (int i = 1; i < horizon - 1; ++i)
    set_block(a_qp, i * state_dim, 0,
              mul(a_exp, block(a_qp, (i - 1) * state_dim, 0, state_dim, state_dim)));

  const int action_dim = b_exp.c;
  Mat& anb_aux = *anb_aux_ptr;
  set_block(anb_aux, 0, 0, b_exp);
  for (int i = 1; i < horizon; ++i)
    set_block(anb_aux, i * state_dim, 0,
              mul(a_exp, block(anb_aux, (i - 1) * state_dim, 0, state_dim, action_dim)));

  Mat& b_qp = *b_qp_ptr;
  for (int i = 0; i < horizon; ++i) {
    set_block(b_qp, i * state_dim, i * action_dim, b_exp);
    for (int j = 0; j < i; ++j) {
      const int power = i - j;
      set_block(b_qp, i * state_dim, j * action_dim,
                block(anb_aux, power * state_dim, 0, state_dim, action_dim));
    }
  }

  auto anb = [&](int k) { return block(anb_aux, k * state_dim, 0, state_dim, action_dim); };
  Mat& p_mat = *p_mat_ptr;
  for (int i = horizon - 1; i >= 0; --i) {
    Mat blk = mul(mul(transpose(anb(horizon - i - 1)), qp_weights_single), b_exp);
    set_block(p_mat, i * action_dim, (horizon - 1) * action_dim, blk);
    if (i != horizon - 1)
      set_block(p_mat, (horizon - 1) * action_dim, i * action_dim, transpose(blk));
  }
  for (int i = horizon - 2; i >= 0; --i) {
    const Mat si_t_w = mul(transpose(anb(horizon - i - 1)), qp_weights_single);
    set_block(p_mat, i * action_dim, i * action_dim,
              add(block(p_mat, (i + 1) * action_dim, (i + 1) * action_dim, action_dim, action_dim),
                  mul(si_t_w, anb(horizon - i - 1))));
    for (int j = i + 1; j < horizon - 1; ++j) {
      Mat blk = add(block(p_mat, (i + 1) * action_dim, (j + 1) * action_dim, action_dim, action_dim),
                    mul(si_t_w, anb(horizon - j - 1)));
      set_block(p_mat, i * action_dim, j * action_dim, blk);
      set_block(p_mat, j * action_dim, i * action_dim, transpose(blk));
    }
  }
  for (auto& v : p_mat.a) v *= 2.0;
  for (int i = 0; i < horizon; ++i)
    for (int a = 0; a < action_dim; ++a)
      for (int b = 0; b < action_dim; ++b)
        p_mat(i * action_dim + a, i * action_dim + b) += alpha_single(a, b);
}

// mpc_osqp.cc:437-447 (friction_coeff[0..3] index the four ROWS, Q2)
void UpdateConstraintsMatrix(const double* friction_coeff, int horizon, int num_legs,
                             Mat* constraint_ptr) {
  Mat& c = *constraint_ptr;
  for (int i = 0; i < horizon * num_legs; ++i) {
    const int r = i * kConstraintDim, col = i * k3Dim;
    const double blk[5][3] = {{-1, 0, friction_coeff[0]},
                              {1, 0, friction_coeff[1]},
                              {0, -1, friction_coeff[2]},
                              {0, 1, friction_coeff[3]},
                              {0, 0, 1}};
    for (int a = 0; a < 5; ++a)
      for (int b = 0; b < 3; ++b) c(r + a, col + b) = blk[a][b];
  }
}

// mpc_osqp.cc:449-477
void CalculateConstraintBounds(const Mat& contact_state, double fz_max, double fz_min,
                               double friction_coeff, int horizon,
                               std::vector<double>* lb, std::vector<double>* ub) {
  const int num_legs = contact_state.c;
  for (int i = 0; i < horizon; ++i) {
    for (int j = 0; j < num_legs; ++j) {
      const int row = (i * num_legs + j) * kConstraintDim;
      (*lb)[row] = 0; (*lb)[row + 1] = 0; (*lb)[row + 2] = 0; (*lb)[row + 3] = 0;
      (*lb)[row + 4] = fz_min * contact_state(i, j);
      const double friction_ub = (friction_coeff + 1) * fz_max * contact_state(i, j);
      (*ub)[row] = friction_ub; (*ub)[row + 1] = friction_ub;
      (*ub)[row + 2] = friction_ub; (*ub)[row + 3] = friction_ub;
      (*ub)[row + 4] = fz_max * contact_state(i, j);
    }
  }
}

// Eigen `dense.sparseView()` (mpc_osqp.cc:696-700): CSC, exact zeros dropped.
struct Csc {
  std::vector<c_int> p, i;
  std::vector<c_float> x;
};
Csc dense_to_csc(const Mat& m, bool upper_only) {
  Csc s;
  s.p.assign(m.c + 1, 0);
  for (int j = 0; j < m.c; ++j) {
    const int imax = upper_only ? std::min(j, m.r - 1) : m.r - 1;
    for (int i = 0; i <= imax; ++i) {
      if (m(i, j) != 0.0) {
        s.i.push_back(i);
        s.x.push_back(m(i, j));
      }
    }
    s.p[j + 1] = (c_int)s.x.size();
  }
  return s;
}

}  // namespace

// ------------------------------------------------------------------ the class
struct MpcOracle {
  // mpc_osqp.cc:218-277 members
  double mass_, inv_mass_;
  Mat inertia_, inv_inertia_;
  int num_legs_, planning_horizon_;
  double timestep_;
  Mat alpha_single_;
  int action_dim_;
  std::vector<double> state_, desired_states_;
  Mat contact_states_, foot_positions_base_, foot_positions_world_;
  Mat a_mat_, b_mat_, ab_concatenated_, a_exp_, b_exp_;
  Mat a_qp_, b_qp_, p_mat_, anb_aux_, constraint_;
  std::vector<double> q_vec_, constraint_lb_, constraint_ub_, qp_solution_;
  ::OSQPWorkspace* workspace_ = nullptr;
  bool initial_run_ = true;
  // bookkeeping
  int last_iters_ = 0, last_polish_ = 0;
  double last_time_ = 0;
};

extern "C" {

// mpc_osqp.cc:508-574
MpcOracle* mpc_oracle_create(double mass, const double* inertia9, int num_legs,
                             int planning_horizon, double timestep, double alpha) {
  MpcOracle* o = new MpcOracle();
  o->mass_ = mass;
  o->inv_mass_ = 1 / mass;
  // `inertia_(inertia.data())`: Eigen column-major read of the row-major list
  o->inertia_ = Mat(3, 3);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o->inertia_(i, j) = inertia9[j * 3 + i];
  o->inv_inertia_ = inverse3(o->inertia_);
  o->num_legs_ = num_legs;
  o->planning_horizon_ = planning_horizon;
  o->timestep_ = timestep;
  o->action_dim_ = num_legs * k3Dim;
  o->alpha_single_ = Mat(o->action_dim_, o->action_dim_);
  for (int i = 0; i < o->action_dim_; ++i) o->alpha_single_(i, i) = alpha;
  const int h = planning_horizon, ad = o->action_dim_;
  o->state_.assign(kStateDim, 0.0);
  o->desired_states_.assign(kStateDim * h, 0.0);
  o->contact_states_ = Mat(h, num_legs);
  o->foot_positions_base_ = Mat(num_legs, k3Dim);
  o->foot_positions_world_ = Mat(num_legs, k3Dim);
  o->a_mat_ = Mat(kStateDim, kStateDim);
  o->b_mat_ = Mat(kStateDim, ad);
  o->ab_concatenated_ = Mat(kStateDim + ad, kStateDim + ad);
  o->a_exp_ = Mat(kStateDim, kStateDim);
  o->b_exp_ = Mat(kStateDim, ad);
  o->a_qp_ = Mat(kStateDim * h, kStateDim);
  o->b_qp_ = Mat(kStateDim * h, ad * h);
  o->p_mat_ = Mat(ad * h, ad * h);
  o->q_vec_.assign(ad * h, 0.0);
  o->anb_aux_ = Mat(kStateDim * h, ad);
  o->constraint_ = Mat(kConstraintDim * num_legs * h, ad * h);
  o->constraint_lb_.assign(kConstraintDim * num_legs * h, 0.0);
  o->constraint_ub_.assign(kConstraintDim * num_legs * h, 0.0);
  o->qp_solution_.assign(ad * h, 0.0);
  return o;
}

void mpc_oracle_destroy(MpcOracle* o) {
  if (!o) return;
  osqp_cleanup(o->workspace_);  // mpc_osqp.cc:192
  delete o;
}

void mpc_oracle_reset_solver(MpcOracle* o) { o->initial_run_ = true; }  // :576

// mpc_osqp.cc:593-688
int mpc_oracle_build_qp(MpcOracle* o, const double* qp_weights, const double* com_position,
                        const double* com_velocity, const double* com_roll_pitch_yaw,
                        const double* ground_normal_vec, const double* com_angular_velocity,
                        const double* foot_contact_states_flattened,
                        const double* foot_positions_body_frame,
                        const double* foot_friction_coeffs, const double* desired_com_position,
                        const double* desired_com_velocity,
                        const double* desired_com_roll_pitch_yaw,
                        const double* desired_com_angular_velocity) {
  const int h = o->planning_horizon_, legs = o->num_legs_;
  // :594-595 AsBlockDiagonalMat
  Mat qp_weights_single(kStateDim, kStateDim);
  for (int i = 0; i < kStateDim; ++i) qp_weights_single(i, i) = qp_weights[i];
  // :598-601 Map(num_legs x horizon, col-major).transpose()
  Mat foot_contact_states(h, legs);
  for (int i = 0; i < h; ++i)
    for (int j = 0; j < legs; ++j)
      foot_contact_states(i, j) = foot_contact_states_flattened[i * legs + j];
  // :606-617 com_rotation = Rx * Ry * Rz (note the order, Q3)
  const Mat com_rotation = mul(rot_x(com_roll_pitch_yaw[0]),
                               mul(rot_y(com_roll_pitch_yaw[1]), rot_z(com_roll_pitch_yaw[2])));
  for (int i = 0; i < legs; ++i)
    for (int k = 0; k < 3; ++k) o->foot_positions_base_(i, k) = foot_positions_body_frame[i * 3 + k];
  for (int i = 0; i < legs; ++i)
    for (int r = 0; r < 3; ++r) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += com_rotation(r, k) * o->foot_positions_base_(i, k);
      o->foot_positions_world_(i, r) = s;
    }
  const double com_x = com_position[0], com_y = com_position[1], com_z = com_position[2];
  // :630-633
  double* st = o->state_.data();
  st[0] = com_roll_pitch_yaw[0]; st[1] = com_roll_pitch_yaw[1]; st[2] = com_roll_pitch_yaw[2];
  st[3] = com_x; st[4] = com_y; st[5] = com_z;
  st[6] = com_angular_velocity[0]; st[7] = com_angular_velocity[1]; st[8] = com_angular_velocity[2];
  st[9] = com_velocity[0]; st[10] = com_velocity[1]; st[11] = com_velocity[2];
  st[12] = -kGravity;
  // :635-659
  for (int i = 0; i < h; ++i) {
    double* d = &o->desired_states_[i * kStateDim];
    d[0] = desired_com_roll_pitch_yaw[0];
    d[1] = desired_com_roll_pitch_yaw[1];
    d[2] = com_roll_pitch_yaw[2] + o->timestep_ * (i + 1) * desired_com_angular_velocity[2];
    d[3] = o->timestep_ * (i + 1) * desired_com_velocity[0] + com_x;
    d[4] = o->timestep_ * (i + 1) * desired_com_velocity[1] + com_y;
    d[5] = desired_com_position[2];
    d[6] = desired_com_angular_velocity[0];
    d[7] = desired_com_angular_velocity[1];
    d[8] = desired_com_angular_velocity[2];
    d[9] = desired_com_velocity[0];
    d[10] = desired_com_velocity[1];
    d[11] = 0;
    d[12] = -kGravity;
  }
  // :667-676
  CalculateAMat(com_roll_pitch_yaw, ground_normal_vec, &o->a_mat_);
  const Mat rotation = ConvertRpyToRot(com_roll_pitch_yaw);
  const Mat inv_inertia_world = mul(mul(rotation, o->inv_inertia_), transpose(rotation));
  CalculateBMat(o->inv_mass_, inv_inertia_world, o->foot_positions_world_, &o->b_mat_);
  CalculateExponentials(o->a_mat_, o->b_mat_, o->timestep_, &o->ab_concatenated_, &o->a_exp_,
                        &o->b_exp_);
  CalculateQpMats(o->a_exp_, o->b_exp_, qp_weights_single, o->alpha_single_, h, &o->a_qp_,
                  &o->anb_aux_, &o->b_qp_, &o->p_mat_);
  // :681-683  state_diff = a_qp * state - desired ; q = 2 * b_qp^T * (W * state_diff)
  const int sd = kStateDim * h, n = o->action_dim_ * h;
  std::vector<double> wdiff(sd);
  for (int r = 0; r < sd; ++r) {
    double s = 0;
    for (int k = 0; k < kStateDim; ++k) s += o->a_qp_(r, k) * st[k];
    wdiff[r] = qp_weights[r % kStateDim] * (s - o->desired_states_[r]);
  }
  for (int j = 0; j < n; ++j) {
    double s = 0;
    for (int r = 0; r < sd; ++r) s += o->b_qp_(r, j) * wdiff[r];
    o->q_vec_[j] = 2 * s;
  }
  // :685-688
  CalculateConstraintBounds(foot_contact_states, o->mass_ * kGravity * kMaxScale,
                            o->mass_ * kGravity * kMinScale, foot_friction_coeffs[0], h,
                            &o->constraint_lb_, &o->constraint_ub_);
  // :691-692
  UpdateConstraintsMatrix(foot_friction_coeffs, h, legs, &o->constraint_);
  o->contact_states_ = foot_contact_states;
  return 0;
}

void mpc_oracle_get_qp(const MpcOracle* o, double* P, double* q, double* C, double* l,
                       double* u) {
  if (P) std::memcpy(P, o->p_mat_.a.data(), o->p_mat_.a.size() * sizeof(double));
  if (q) std::memcpy(q, o->q_vec_.data(), o->q_vec_.size() * sizeof(double));
  if (C) std::memcpy(C, o->constraint_.a.data(), o->constraint_.a.size() * sizeof(double));
  if (l) std::memcpy(l, o->constraint_lb_.data(), o->constraint_lb_.size() * sizeof(double));
  if (u) std::memcpy(u, o->constraint_ub_.data(), o->constraint_ub_.size() * sizeof(double));
}

void mpc_oracle_get_exponentials(const MpcOracle* o, double* a_exp, double* b_exp) {
  if (a_exp) std::memcpy(a_exp, o->a_exp_.a.data(), o->a_exp_.a.size() * sizeof(double));
  if (b_exp) std::memcpy(b_exp, o->b_exp_.a.data(), o->b_exp_.a.size() * sizeof(double));
}

int mpc_oracle_last_iters(const MpcOracle* o) { return o->last_iters_; }
int mpc_oracle_last_polish_status(const MpcOracle* o) { return o->last_polish_; }
double mpc_oracle_last_solve_time(const MpcOracle* o) { return o->last_time_; }

// mpc_osqp.cc:690-796 (OSQP branch)
int mpc_oracle_compute_contact_forces(
    MpcOracle* o, int mode, const double* qp_weights, const double* com_position,
    const double* com_velocity, const double* com_roll_pitch_yaw,
    const double* ground_normal_vec, const double* com_angular_velocity,
    const double* foot_contact_states, const double* foot_positions_body_frame,
    const double* foot_friction_coeffs, const double* desired_com_position,
    const double* desired_com_velocity, const double* desired_com_roll_pitch_yaw,
    const double* desired_com_angular_velocity, double* out_forces) {
  mpc_oracle_build_qp(o, qp_weights, com_position, com_velocity, com_roll_pitch_yaw,
                      ground_normal_vec, com_angular_velocity, foot_contact_states,
                      foot_positions_body_frame, foot_friction_coeffs, desired_com_position,
                      desired_com_velocity, desired_com_roll_pitch_yaw,
                      desired_com_angular_velocity);

  Csc objective_upper = dense_to_csc(o->p_mat_, /*upper_only=*/true);   // :696-697, :727-729
  Csc constraint_matrix = dense_to_csc(o->constraint_, false);          // :699-700
  const c_int num_variables = o->constraint_.c;
  const c_int num_constraints = o->constraint_.r;

  ::OSQPSettings settings;  // :704-712
  osqp_set_default_settings(&settings);
  settings.verbose = false;
  settings.warm_start = true;
  settings.polish = true;
  settings.adaptive_rho_interval = 25;
  settings.eps_abs = 1e-3;
  settings.eps_rel = 1e-3;
  if (mode == MPC_ORACLE_TRUE) {
    settings.eps_abs = 1e-10;
    settings.eps_rel = 1e-10;
    settings.max_iter = 200000;
    settings.polish_refine_iter = 10;
  }

  std::vector<double> lo = o->constraint_lb_, hi = o->constraint_ub_;  // :720-721
  for (auto& v : lo) v = std::max(v, (double)-OSQP_INFTY);
  for (auto& v : hi) v = std::min(v, (double)OSQP_INFTY);

  ::OSQPData data;  // :723-753
  data.n = num_variables;
  data.m = num_constraints;
  ::csc P = {objective_upper.p[num_variables], num_variables, num_variables,
             objective_upper.p.data(), objective_upper.i.data(), objective_upper.x.data(), -1};
  ::csc A = {constraint_matrix.p[num_variables], num_constraints, num_variables,
             constraint_matrix.p.data(), constraint_matrix.i.data(), constraint_matrix.x.data(), -1};
  data.P = &P;
  data.A = &A;
  data.q = o->q_vec_.data();
  data.l = lo.data();
  data.u = hi.data();

  if (mode != MPC_ORACLE_WARM && o->workspace_) {
    osqp_cleanup(o->workspace_);
    o->workspace_ = nullptr;
  }
  double t_setup = 0;
  if (o->workspace_ == nullptr) {  // :757-759
    if (osqp_setup(&o->workspace_, &data, &settings) != 0) {
      o->workspace_ = nullptr;
      return -100;
    }
    o->initial_run_ = false;
    t_setup = o->workspace_->info->setup_time;
  } else {  // :760-778 (return codes ignored, as the reference does)
    osqp_update_P_A(o->workspace_, objective_upper.x.data(), OSQP_NULL,
                    (c_int)objective_upper.x.size(), constraint_matrix.x.data(), OSQP_NULL,
                    (c_int)constraint_matrix.x.size());
    osqp_update_lin_cost(o->workspace_, o->q_vec_.data());
    osqp_update_bounds(o->workspace_, lo.data(), hi.data());
  }
  if (osqp_solve(o->workspace_) != 0) {  // :780-784
    if (osqp_is_interrupted()) return -101;
  }
  const OSQPInfo* info = o->workspace_->info;
  o->last_iters_ = (int)info->iter;
  o->last_polish_ = (int)info->status_polish;
  o->last_time_ = t_setup + info->update_time + info->solve_time + info->polish_time;
  const int status = (int)info->status_val;
  if (status == OSQP_SOLVED) {  // :788-790
    for (c_int i = 0; i < num_variables; ++i) {
      o->qp_solution_[i] = -o->workspace_->solution->x[i];
      if (out_forces) out_forces[i] = o->qp_solution_[i];
    }
  }
  return status;
}

int mpc_oracle_solve_records(const float* records, int n, int horizon, double timestep,
                             double alpha, const float* robot_params, const int* robot_type,
                             int mode, double* grf0, double* grf_full, int* status_out) {
  const int stride = MPC_REC_STRIDE(horizon);
  const int nv = 12 * horizon;
  int failures = 0;
  // one persistent solver instance per robot TYPE (the reference keeps one ConvexMpc per
  // robot; warm mode therefore shares a workspace across the robots of a type, which only
  // changes the warm-start point, not the pattern)
  std::vector<MpcOracle*> solvers;
  std::vector<double> out(nv);
  for (int r = 0; r < n; ++r) {
    const int t = robot_type ? robot_type[r] : 0;
    if ((int)solvers.size() <= t) solvers.resize(t + 1, nullptr);
    if (!solvers[t]) {
      const float* rp = robot_params + (size_t)t * MPC_ROBOT_PARAMS_STRIDE;
      double inertia[9];
      for (int k = 0; k < 9; ++k) inertia[k] = rp[1 + k];
      solvers[t] = mpc_oracle_create(rp[0], inertia, 4, horizon, timestep, alpha);
    }
    const float* rec = records + (size_t)r * stride;
    double d[MPC_REC_CONTACT + 4 * MPC_B200_MAX_HORIZON];
    for (int k = 0; k < MPC_REC_CONTACT + 4 * horizon; ++k) d[k] = rec[k];
    std::fill(out.begin(), out.end(), 0.0);
    const int st = mpc_oracle_compute_contact_forces(
        solvers[t], mode, d + MPC_REC_WEIGHTS, d + MPC_REC_COM_POS, d + MPC_REC_COM_VEL,
        d + MPC_REC_RPY, d + MPC_REC_NORMAL, d + MPC_REC_ANGVEL, d + MPC_REC_CONTACT,
        d + MPC_REC_FOOTPOS, d + MPC_REC_MU, d + MPC_REC_DES_POS, d + MPC_REC_DES_VEL,
        d + MPC_REC_DES_RPY, d + MPC_REC_DES_ANGVEL, out.data());
    if (st != OSQP_SOLVED) ++failures;
    if (status_out) status_out[r] = st;
    if (grf0) std::memcpy(grf0 + (size_t)r * 12, out.data(), 12 * sizeof(double));
    if (grf_full) std::memcpy(grf_full + (size_t)r * nv, out.data(), nv * sizeof(double));
  }
  for (auto* s : solvers) mpc_oracle_destroy(s);
  return failures;
}

}  // extern "C"
