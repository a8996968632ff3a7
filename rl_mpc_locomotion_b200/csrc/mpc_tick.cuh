// mpc_tick.cuh — the Locomotion tick around the QP, fused into the solve launch.
//
// Device restatement (lane-program dialect, see warp_prog.h) of the per-robot control tick
// the reference runs in Python, for every robot, every 10 ms:
//   LegController.updateData / computeLegJacobianAndPosition   reference/MPC_Controller/common/LegController.py:89-106,135-171
//   ConvexMPCLocomotion.run                                    reference/MPC_Controller/convex_MPC/ConvexMPCLocomotion.py:222-378
//     OffsetDurationGait (phase, contact/swing progress, horizon table)   convex_MPC/Gait.py:26-84
//     StateEstimator CoM height + ground normal                 common/StateEstimator.py:109-144
//     foot-placement heuristic, swing Bezier                    ConvexMPCLocomotion.py:272-311, common/FootSwingTrajectory.py:54-70
//     solveDenseMPC argument marshalling                         ConvexMPCLocomotion.py:128-187
//   LegController.updateCommand (Cartesian PD + J^T f torque map)           LegController.py:108-132
// tick_pre() builds the QP's problem record straight into shared memory (it never touches
// HBM), the caller runs solve_robot() when the MPC is due (every iterationsBetweenMPC ticks,
// ConvexMPCLocomotion.py:217-220), tick_post() turns forces + swing trajectories into joint
// torques.  StateEstimator.update() (quaternion -> fp16 rotation, body-frame velocities,
// rpyBody; StateEstimator.py:57-97) is estimator_update() below and runs in the same launch
// when the caller passes raw root states instead of precomputed estimator outputs.
//
// Arithmetic follows the reference's dtypes: float32 arrays, double where the reference mixes
// in Python floats, and fp16-rounded sin/cos in the yaw-correction rotation
// (orientation_tools.py:13,20-36: DTYPE = np.float16).
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/mpc_b200.h"
#include "warp_prog.h"

#if defined(__CUDACC__) && !defined(MPC_EMULATE)
#include <cuda_fp16.h>
#endif

namespace mpc {

struct TickParams {
  float dt;               // controller dt, Parameters.controller_dt = 0.01
  int iters_between_mpc;  // int(27 / (1000 * dt)) = 2, RobotRunnerMin.py:21-22
  int flat_ground;        // Parameters.flat_ground
  float bonus_swing;      // Parameters.cmpc_bonus_swing
  int cmd_stride;         // floats per robot in the command tensor: 3 (vx, vy, yaw rate) or 16 (+ 13 weights)
  int reserved[3];
  float act_scale[12];    // Parameters.MPC_param_scale  (aliengo.py:237-245: weights = actions * scale + const)
  float act_const[12];    // Parameters.MPC_param_const
};

// value rounded through IEEE half precision (numpy float16 storage)
MPC_LANEFN float round_f16(double v) {
#if defined(__CUDACC__) && !defined(MPC_EMULATE)
  return __half2float(__double2half(v));
#else
  // software round-to-nearest-even to 11 significant bits (normal range only; |v| <= 1 here)
  if (v == 0.0 || !std::isfinite(v)) return (float)v;
  int e;
  double m = std::frexp(v, &e);            // v = m * 2^e, 0.5 <= |m| < 1
  if (e < -13) {                           // subnormal half: quantum 2^-24
    return (float)(std::nearbyint(v * 16777216.0) / 16777216.0);
  }
  const double q = std::nearbyint(m * 2048.0) / 2048.0;
  return (float)std::ldexp(q, e);
#endif
}

// gait id -> (offsets, durations) in MPC segments (ConvexMPCLocomotion.py:30-56, ids :227-241)
MPC_LANEFN void gait_leg(int gait, int leg, float* off, float* dur) {
  int o = 0, d = 5;
  switch (gait) {
    case 1: { const int t[4] = {5, 5, 0, 0}; o = t[leg]; d = 4; break; }    // bound
    case 2: { o = 0; d = 4; break; }                                         // pronk
    case 3: { const int t[4] = {5, 0, 5, 0}; o = t[leg]; d = 5; break; }    // pace
    case 5: { const int t[4] = {0, 2, 7, 9}; o = t[leg]; d = 4; break; }    // gallop
    case 6: { const int t[4] = {0, 3, 5, 8}; o = t[leg]; d = 5; break; }    // walk
    case 7: { const int t[4] = {0, 5, 5, 0}; o = t[leg]; d = 4; break; }    // trot run
    default: { const int t[4] = {0, 5, 5, 0}; o = t[leg]; d = 5; break; }   // trot (and any other id)
  }
  *off = (float)o;
  *dur = (float)d;
}

MPC_LANEFN float bez(float x) { return x * x * x + 3.0f * (x * x * (1.0f - x)); }   // interplation.py:4-11
MPC_LANEFN float dbez(float x) { return 6.0f * x * (1.0f - x); }                     // interplation.py:13-20

// ---------------------------------------------------------------------------------------
// StateEstimator.update (reference/MPC_Controller/common/StateEstimator.py:57-97) with the
// reference's mixed float16 / float32 arithmetic reproduced operation by operation
// (math_utils/orientation_tools.py: DTYPE = np.float16; NumPy 2 promotion rules: a Python
// float next to a float16 scalar is rounded to float16, float16 mat-muls and reductions
// accumulate in float32 and round once).  body = [pos3, quat xyzw, v_world3, omega_world3],
// normal = ground_normal_yaw of the previous tick; est = MPC_EST_* outputs.
// ---------------------------------------------------------------------------------------
MPC_LANEFN float h16(float v) { return round_f16((double)v); }
MPC_LANEFN void mm3_f16(const float* a, const float* b, float* o) {   // o = h16(a @ b), row-major
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) o[3 * r + c] = h16(a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c] + a[3 * r + 2] * b[6 + c]);
}
// quat_to_rpy on a quaternion whose w is a Python float and x, y, z are float16 scalars
// (what rot_to_quat returns for tr > 0, orientation_tools.py:120-133, 162-175)
MPC_LANEFN void estimator_update(const float* body, const float* normal, float* est, float* rp16 = nullptr) {
  const float qx = body[3], qy = body[4], qz = body[5], qw = body[6];
  // quat_to_rot (orientation_tools.py:135-151): float32 expressions, stored as float16, transposed
  const float e0 = qw, e1 = qx, e2 = qy, e3 = qz;
  float Rt[9];  // R before the transpose
  Rt[0] = 1 - 2 * (e2 * e2 + e3 * e3); Rt[1] = 2 * (e1 * e2 - e0 * e3); Rt[2] = 2 * (e1 * e3 + e0 * e2);
  Rt[3] = 2 * (e1 * e2 + e0 * e3); Rt[4] = 1 - 2 * (e1 * e1 + e3 * e3); Rt[5] = 2 * (e2 * e3 - e0 * e1);
  Rt[6] = 2 * (e1 * e3 - e0 * e2); Rt[7] = 2 * (e2 * e3 + e0 * e1); Rt[8] = 1 - 2 * (e1 * e1 + e2 * e2);
  float rB[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) rB[3 * r + c] = h16(Rt[3 * c + r]);
  for (int r = 0; r < 3; ++r) {   // vBody = rBody @ vWorld, omegaBody = rBody @ omegaWorld (float32)
    est[MPC_EST_VBODY + r] = rB[3 * r] * body[7] + rB[3 * r + 1] * body[8] + rB[3 * r + 2] * body[9];
    est[MPC_EST_OMEGABODY + r] = rB[3 * r] * body[10] + rB[3 * r + 1] * body[11] + rB[3 * r + 2] * body[12];
  }
  if (rp16) {  // world-frame roll / pitch of quat_to_rpy (float16), used by locomotionSafe()
    rp16[0] = h16(atan2f(2.0f * (qy * qz + qw * qx), qw * qw - qx * qx - qy * qy + qz * qz));
    const double as0 = (double)(-2.0f * (qx * qz - qw * qy));
    rp16[1] = round_f16(asin(as0 < 0.99999 ? as0 : 0.99999));
  }
  // yaw of quat_to_rpy(orientation) (float32 args, float16 store)
  const float yaw16 = h16(atan2f(2.0f * (qx * qy + qw * qz), qw * qw + qx * qx - qy * qy - qz * qz));
  // world_R_yaw = coordinateRotation(Z, yaw) (X(0) and Y(0) are exact identities); we need its transpose
  const float cz = h16((float)cos((double)yaw16)), sz = h16((float)sin((double)yaw16));
  const float WyT[9] = {cz, -sz, 0.0f, sz, cz, 0.0f, 0.0f, 0.0f, 1.0f};
  // yaw_R_ground = get_rot_from_normals([0,0,1], normal) (orientation_tools.py:89-108, incl. the
  // k1*k1 entry and the un-normalised axis); we need its transpose = R_axis_angle
  const float k0 = -normal[1], k1 = normal[0], k2 = 0.0f;
  const float theta = acosf(normal[2]);
  const float ct = (float)cos((double)theta), st_ = (float)sin((double)theta), vt = (float)(1.0 - cos((double)theta));
  float Ra[9];
  Ra[0] = h16(k0 * k0 * vt + ct); Ra[1] = h16(k0 * k1 * vt - k2 * st_); Ra[2] = h16(k0 * k2 * vt + k1 * st_);
  Ra[3] = h16(k0 * k1 * vt + k2 * st_); Ra[4] = h16(k1 * k1 * vt + ct); Ra[5] = h16(k1 * k2 * vt - k0 * st_);
  Ra[6] = h16(k0 * k2 * vt - k1 * st_); Ra[7] = h16(k1 * k2 * vt + k0 * st_); Ra[8] = h16(k1 * k1 * vt + ct);
  // ground_R_body_frame = (rBody @ world_R_yaw.T) @ yaw_R_ground.T ; yaw_R_ground = Ra.T so its .T is Ra
  float M1[9], G[9];
  mm3_f16(rB, WyT, M1);
  mm3_f16(M1, Ra, G);
  for (int e = 0; e < 9; ++e) est[MPC_EST_GROUNDR + e] = G[e];
  // rpyBody = quat_to_rpy(rot_to_quat(G)) (orientation_tools.py:162-203); r = G.T
  const float tr = h16(G[0] + G[4] + G[8]);
  float w_, x_, y_, z_;   // w_ is a Python float (double in the reference) on the tr > 0 branch
  double wd;
  if (tr > 0.0f) {
    const double S = sqrt((double)h16(tr + 1.0f)) * 2.0;
    wd = 0.25 * S;
    const float Sh = h16((float)S);
    x_ = h16(h16(G[5] - G[7]) / Sh);   // (r[2,1] - r[1,2]) / S with r = G.T
    y_ = h16(h16(G[6] - G[2]) / Sh);
    z_ = h16(h16(G[1] - G[3]) / Sh);
  } else {
    // large-rotation branches (never reached inside the locomotion envelope): evaluate in double
    const double r00 = G[0], r11 = G[4], r22 = G[8];
    double S, xx, yy, zz;
    if (r00 > r11 && r00 > r22) { S = sqrt(1.0 + r00 - r11 - r22) * 2; wd = (G[5] - G[7]) / S; xx = 0.25 * S; yy = (G[3] + G[1]) / S; zz = (G[6] + G[2]) / S; }
    else if (r11 > r22) { S = sqrt(1.0 + r11 - r00 - r22) * 2; wd = (G[6] - G[2]) / S; xx = (G[3] + G[1]) / S; yy = 0.25 * S; zz = (G[7] + G[5]) / S; }
    else { S = sqrt(1.0 + r22 - r00 - r11) * 2; wd = (G[1] - G[3]) / S; xx = (G[6] + G[2]) / S; yy = (G[7] + G[5]) / S; zz = 0.25 * S; }
    x_ = h16((float)xx); y_ = h16((float)yy); z_ = h16((float)zz);
  }
  w_ = h16((float)wd);   // a Python float next to float16 scalars is rounded to float16
  const float as_arg = h16(-2.0f * h16(h16(x_ * z_) - h16(w_ * y_)));
  const double as_ = (double)as_arg < 0.99999 ? (double)as_arg : 0.99999;
  const float ww = h16((float)(wd * wd));
  const float num0 = h16(2.0f * h16(h16(y_ * z_) + h16(w_ * x_)));
  const float den0 = h16(h16(h16(ww - h16(x_ * x_)) - h16(y_ * y_)) + h16(z_ * z_));
  const float num2 = h16(2.0f * h16(h16(x_ * y_) + h16(w_ * z_)));
  const float den2 = h16(h16(h16(ww + h16(x_ * x_)) - h16(y_ * y_)) - h16(z_ * z_));
  est[MPC_EST_RPYBODY + 0] = h16(atan2f(num0, den0));
  est[MPC_EST_RPYBODY + 1] = round_f16(asin(as_));
  est[MPC_EST_RPYBODY + 2] = h16(atan2f(num2, den2));
}

// FSM_State_Locomotion.locomotionSafe (FSM_states/FSM_State_Locomotion.py:104-136), including its
// `fabs(bool)` slips: only POSITIVE roll beyond 40 deg and only leg y > 0.18 trip.
MPC_LANEFN int locomotion_unsafe(const float* rp16, const float* legs_p /* 4 x stride */, int stride) {
  const float lim = (float)(40.0 * 3.14159265358979323846 / 180.0);
  if (rp16[0] > lim) return 1;
  if (fabsf(rp16[1]) > lim) return 1;
  for (int l = 0; l < 4; ++l) {
    if (legs_p[l * stride + 2] > 0.0f) return 1;
    if (legs_p[l * stride + 1] > 0.18f) return 1;
  }
  return 0;
}

// per-leg scratch in shared memory (floats)
enum { LG_Q = 0, LG_QD = 3, LG_P = 6, LG_J = 9, LG_V = 18, LG_PFOOT = 21, LG_FOOT = 24, LG_CONTACT = 27,
       LG_SWING = 28, LG_SWINGTIME = 29, LG_STRIDE = 32 };

// ---------------------------------------------------------------------------------------
// tick_legs: LegController.updateData (LegController.py:89-106, 135-171) - leg FK, Jacobian, foot
// velocity, hip-frame foot positions.  Reads the persistent state (CoM height) but does not
// change it, so the FSM shell's locomotionSafe() check can run between this and tick_pre().
// ---------------------------------------------------------------------------------------
MPC_FN void tick_legs(const float* dof, const float* geom, const float* st, float* rec, float* legs MPC_LANE_ARG) {
  const float abad = geom[MPC_GEOM_ABAD], hipl = geom[MPC_GEOM_HIP], kneel = geom[MPC_GEOM_KNEE];
  LANES_BEGIN
  if (lane < 4) {
    const int leg = lane;
    float* L = legs + leg * LG_STRIDE;
    float q[3], qd[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      q[a] = dof[(3 * leg + a) * 2];
      qd[a] = dof[(3 * leg + a) * 2 + 1];
      L[LG_Q + a] = q[a];
      L[LG_QD + a] = qd[a];
    }
    const double side = (leg & 1) ? -1.0 : 1.0;     // utils.py:7
    const double dy = (double)abad * side, dz1 = -(double)hipl, dz2 = -(double)kneel;
    const double s1 = sin((double)q[0]), s2 = sin((double)q[1]), s3 = sin((double)q[2]);
    const double c1 = cos((double)q[0]), c2 = cos((double)q[1]), c3 = cos((double)q[2]);
    const double c23 = c2 * c3 - s2 * s3, s23 = s2 * c3 + c2 * s3;
    float p[3], J[9];
    p[0] = (float)(dz2 * s23 + dz1 * s2);
    p[1] = (float)(dy * c1 - dz1 * c2 * s1 - dz2 * s1 * c23);
    p[2] = (float)(dy * s1 + dz1 * c1 * c2 + dz2 * c1 * c23);
    J[0] = 0.0f;
    J[3] = (float)(-dy * s1 - dz2 * c1 * c23 - dz1 * c1 * c2);
    J[6] = (float)(-dz2 * s1 * c23 + dy * c1 - dz1 * c2 * s1);
    J[1] = (float)(dz2 * c23 + dz1 * c2);
    J[4] = (float)(dz2 * s1 * s23 + dz1 * s1 * s2);
    J[7] = (float)(-dz2 * c1 * s23 - dz1 * c1 * s2);
    J[2] = (float)(dz2 * c23);
    J[5] = (float)(dz2 * s1 * s23);
    J[8] = (float)(-dz2 * c1 * s23);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      L[LG_P + r] = p[r];
      L[LG_V + r] = J[3 * r] * qd[0] + J[3 * r + 1] * qd[1] + J[3 * r + 2] * qd[2];
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) L[LG_J + e] = J[e];
    // foot_positions = hip + p ; pFoot = foot_positions + position (OLD z)   (ConvexMPCLocomotion.py:248-250)
    const float hip[3] = {leg < 2 ? geom[MPC_GEOM_LOCX] : -geom[MPC_GEOM_LOCX],
                          (leg & 1) == 0 ? geom[MPC_GEOM_LOCY] : -geom[MPC_GEOM_LOCY], geom[MPC_GEOM_LOCZ]};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float f = hip[a] + p[a];
      L[LG_FOOT + a] = f;
      L[LG_PFOOT + a] = f + (a == 2 ? st[MPC_ST_POSZ] : 0.0f);
      rec[MPC_REC_FOOTPOS + 3 * leg + a] = f;
    }
  }
  LANES_END
  WSYNC();
}

// ---------------------------------------------------------------------------------------
// tick_pre: everything of the tick between LegController.updateData (tick_legs) and the QP solve.
// `st` (MPC_TICK_STATE_STRIDE floats, shared memory copy of the robot's persistent state) is
// updated in place; `rec` receives the QP problem record; returns (uniformly) whether the MPC
// fires this tick.  `cmd` = vx, vy, yaw rate; `wts` = the 13 MPC weights or null (robot default).
// ---------------------------------------------------------------------------------------
template <int H>
MPC_FN int tick_pre(const float* est, const float* cmd, const float* wts, int gait,
                    const float* geom, const TickParams& tp, float* st, float* rec, float* legs MPC_LANE_ARG) {
  const float abad = geom[MPC_GEOM_ABAD];
  const float body_h = geom[MPC_GEOM_HEIGHT];
  const double dt_mpc = (double)tp.dt * tp.iters_between_mpc;
  int* sti = reinterpret_cast<int*>(st);

  // ---- first-run init, CoM height, contact history, ground normal (StateEstimator.py:99-144)
  LANES_BEGIN
  if (lane == 0) {
    if (sti[MPC_ST_FIRSTRUN]) {   // ConvexMPCLocomotion.py:257-263
      sti[MPC_ST_FIRSTRUN] = 0;
      for (int l = 0; l < 4; ++l)
        for (int a = 0; a < 3; ++a) {
          st[MPC_ST_HIST + 3 * l + a] = a == 2 ? -body_h : legs[l * LG_STRIDE + LG_FOOT + a];
          st[MPC_ST_P0 + 3 * l + a] = legs[l * LG_STRIDE + LG_PFOOT + a];
          st[MPC_ST_PF + 3 * l + a] = legs[l * LG_STRIDE + LG_PFOOT + a];
        }
    }
    // _update_com_position_ground_frame: contact-progress weighted mean of -z in the ground frame
    float csum = 0.0f, hsum = 0.0f;
    for (int l = 0; l < 4; ++l) {
      const float* f = legs + l * LG_STRIDE + LG_FOOT;
      const float zg = est[MPC_EST_GROUNDR + 6] * f[0] + est[MPC_EST_GROUNDR + 7] * f[1] + est[MPC_EST_GROUNDR + 8] * f[2];
      csum += st[MPC_ST_CPHASE + l];
      hsum += (-zg) * st[MPC_ST_CPHASE + l];
    }
    if (csum != 0.0f) st[MPC_ST_POSZ] = hsum / csum;
    if (!tp.flat_ground) {
      // _update_contact_history + lstsq(history, 1) (normal equations in fp64, sgelsd in the reference)
      double AtA[6] = {0, 0, 0, 0, 0, 0}, Atb[3] = {0, 0, 0};
      for (int l = 0; l < 4; ++l) {
        if (st[MPC_ST_CPHASE + l] != 0.0f)
          for (int a = 0; a < 3; ++a) st[MPC_ST_HIST + 3 * l + a] = legs[l * LG_STRIDE + LG_FOOT + a];
        const double x = st[MPC_ST_HIST + 3 * l], y = st[MPC_ST_HIST + 3 * l + 1], z = st[MPC_ST_HIST + 3 * l + 2];
        AtA[0] += x * x; AtA[1] += x * y; AtA[2] += x * z; AtA[3] += y * y; AtA[4] += y * z; AtA[5] += z * z;
        Atb[0] += x; Atb[1] += y; Atb[2] += z;
      }
      const double c00 = AtA[3] * AtA[5] - AtA[4] * AtA[4], c01 = AtA[2] * AtA[4] - AtA[1] * AtA[5],
                   c02 = AtA[1] * AtA[4] - AtA[2] * AtA[3];
      const double det = AtA[0] * c00 + AtA[1] * c01 + AtA[2] * c02;
      double n[3];
      n[0] = (c00 * Atb[0] + c01 * Atb[1] + c02 * Atb[2]) / det;
      n[1] = (c01 * Atb[0] + (AtA[0] * AtA[5] - AtA[2] * AtA[2]) * Atb[1] + (AtA[1] * AtA[2] - AtA[0] * AtA[4]) * Atb[2]) / det;
      n[2] = (c02 * Atb[0] + (AtA[1] * AtA[2] - AtA[0] * AtA[4]) * Atb[1] + (AtA[0] * AtA[3] - AtA[1] * AtA[1]) * Atb[2]) / det;
      float nf[3] = {(float)n[0], (float)n[1], (float)n[2]};
      float nn = sqrtf(nf[0] * nf[0] + nf[1] * nf[1] + nf[2] * nf[2]);
      for (int a = 0; a < 3; ++a) nf[a] /= nn;
      if (nf[2] < 0.0f) for (int a = 0; a < 3; ++a) nf[a] = -nf[a];
      nn = sqrtf(nf[0] * nf[0] + nf[1] * nf[1] + nf[2] * nf[2]);
      for (int a = 0; a < 3; ++a) st[MPC_ST_NORMAL + a] = nf[a] / nn;
    }
  }
  LANES_END
  WSYNC();

  // ---- gait (pre-increment counter), foot placement, horizon table
  const int counter = sti[MPC_ST_COUNTER];
  const double g_iter = fmod((double)counter / tp.iters_between_mpc, 10.0);                       // Gait.py:27
  const float g_phase = (float)((double)(counter % (tp.iters_between_mpc * 10)) / (double)(tp.iters_between_mpc * 10));
  LANES_BEGIN
  if (lane < 4) {
    const int leg = lane;
    float* L = legs + leg * LG_STRIDE;
    float off, dur, off0, dur0;
    gait_leg(gait, leg, &off, &dur);
    gait_leg(gait, 0, &off0, &dur0);                      // __stance / __swing use durations[0] (Gait.py:22-23)
    const float swing_time = (float)dt_mpc * (10.0f - dur0);  // getCurrentSwingTime
    const float stance_time = (float)dt_mpc * dur0;           // getCurrentStanceTime
    L[LG_SWINGTIME] = swing_time;
    float rem;
    if (sti[MPC_ST_FIRSTSWING + leg]) rem = swing_time;
    else rem = (float)((double)st[MPC_ST_SWINGREM + leg] - (double)tp.dt);
    st[MPC_ST_SWINGREM + leg] = rem;
    const float vx_des = cmd[0], vy_des = cmd[1], yaw_rate = cmd[2];
    const float side = (leg & 1) ? -1.0f : 1.0f;
    const float hx = leg < 2 ? geom[MPC_GEOM_LOCX] : -geom[MPC_GEOM_LOCX];
    const float hy = ((leg & 1) == 0 ? geom[MPC_GEOM_LOCY] : -geom[MPC_GEOM_LOCY]) + side * abad;
    const float hz = geom[MPC_GEOM_LOCZ];
    // coordinateRotation(Z, -yaw_rate * stance_time / 2) is a float16 matrix (orientation_tools.py:13,34)
    const double ang = (double)(-yaw_rate * stance_time / 2.0f);
    const float c16 = round_f16(cos(ang)), s16 = round_f16(sin(ang));
    const float pyx = c16 * hx + s16 * hy, pyy = -s16 * hx + c16 * hy, pyz = hz;
    const float vbx = est[MPC_EST_VBODY], vby = est[MPC_EST_VBODY + 1];
    const float posz = st[MPC_ST_POSZ];
    float Pf[3] = {0.0f + (pyx + vx_des * rem), 0.0f + (pyy + vy_des * rem), posz + (pyz + 0.0f * rem)};
    float pfx = vbx * (0.5f + tp.bonus_swing) * stance_time + 0.03f * (vbx - vx_des) +
                (0.5f * posz / 9.81f) * (vby * yaw_rate);
    float pfy = vby * 0.5f * stance_time * (float)dt_mpc + 0.03f * (vby - vy_des) +
                (0.5f * posz / 9.81f) * (-vbx * yaw_rate);
    pfx = fminf(fmaxf(pfx, -0.3f), 0.3f);
    pfy = fminf(fmaxf(pfy, -0.3f), 0.3f);
    Pf[0] += pfx;
    Pf[1] += pfy;
    Pf[2] = -0.003f;
#pragma unroll
    for (int a = 0; a < 3; ++a) st[MPC_ST_PF + 3 * leg + a] = Pf[a];
    // contact / swing progress in float32, reference operation order (Gait.py:30-67)
    const float offf = off / 10.0f, durf = dur / 10.0f;
    float pc = g_phase - offf;
    if (pc < 0.0f) pc += 1.0f;
    L[LG_CONTACT] = pc > durf ? 0.0f : pc / durf;
    float swo = offf + durf;
    if (swo > 1.0f) swo -= 1.0f;
    const float swd = 1.0f - durf;
    float ps = g_phase - swo;
    if (ps < 0.0f) ps += 1.0f;
    L[LG_SWING] = (ps > swd || swd == 0.0f) ? 0.0f : ps / swd;
    // horizon contact table (Gait.py:69-84); horizons beyond 10 cycle the 10-segment gait
    for (int i = 0; i < H; ++i) {
      const double it = fmod((double)i + g_iter + 1.0, 10.0);
      double pr = it - (double)off;
      if (pr < 0.0) pr += 10.0;
      rec[MPC_REC_CONTACT + 4 * i + leg] = pr < (double)dur ? 1.0f : 0.0f;
    }
  }
  LANES_END
  WSYNC();

  // ---- counter++ and the rest of the problem record (ConvexMPCLocomotion.py:128-185, 314)
  const int counter_post = counter + 1;
  LANES_BEGIN
  if (lane == 0) sti[MPC_ST_COUNTER] = counter_post;
  if (lane < 13) rec[MPC_REC_WEIGHTS + lane] = wts ? wts[lane] : geom[MPC_GEOM_WEIGHTS + lane];
  if (lane < 3) {
    rec[MPC_REC_COM_POS + lane] = lane == 2 ? st[MPC_ST_POSZ] : 0.0f;
    rec[MPC_REC_COM_VEL + lane] = est[MPC_EST_VBODY + lane];
    rec[MPC_REC_RPY + lane] = est[MPC_EST_RPYBODY + lane];
    rec[MPC_REC_NORMAL + lane] = tp.flat_ground ? (lane == 2 ? 1.0f : 0.0f) : st[MPC_ST_NORMAL + lane];
    rec[MPC_REC_ANGVEL + lane] = est[MPC_EST_OMEGABODY + lane];
    rec[MPC_REC_DES_POS + lane] = lane == 2 ? body_h : 0.0f;
    rec[MPC_REC_DES_VEL + lane] = lane == 2 ? 0.0f : cmd[lane];
    rec[MPC_REC_DES_RPY + lane] = 0.0f;
    rec[MPC_REC_DES_ANGVEL + lane] = lane == 2 ? cmd[2] : 0.0f;
  }
  if (lane < 4) rec[MPC_REC_MU + lane] = geom[MPC_GEOM_MU];
  LANES_END
  WSYNC();
  return (counter_post % tp.iters_between_mpc) == 0;   // ConvexMPCLocomotion.py:219
}

// ---------------------------------------------------------------------------------------
// tick_post: swing / stance leg commands and the torque map (ConvexMPCLocomotion.py:326-378,
// LegController.py:108-132).  `grf0` = the 12 horizon-0 forces of this tick's solve or null.
// ---------------------------------------------------------------------------------------
MPC_FN void tick_post(const float* est, const float* geom, const float* grf0, float* st, const float* legs,
                      float* torques MPC_LANE_ARG) {
  int* sti = reinterpret_cast<int*>(st);
  LANES_BEGIN
  if (lane < 4) {
    const int leg = lane;
    const float* L = legs + leg * LG_STRIDE;
    if (grf0)
      for (int a = 0; a < 3; ++a) st[MPC_ST_FFF + 3 * leg + a] = grf0[3 * leg + a];
    const float swing = L[LG_SWING], contact = L[LG_CONTACT], swing_time = L[LG_SWINGTIME];
    const float height = geom[MPC_GEOM_HEIGHT] / 3.0f;        // setHeight(body_height / 3)
    float kp[3], kd[3] = {7.0f, 7.0f, 7.0f}, fff[3] = {0.0f, 0.0f, 0.0f}, kdj = 0.0f, cphase = 0.0f;
    if (swing > 0.0f) {
      if (sti[MPC_ST_FIRSTSWING + leg]) {
        sti[MPC_ST_FIRSTSWING + leg] = 0;
        for (int a = 0; a < 3; ++a) st[MPC_ST_P0 + 3 * leg + a] = L[LG_PFOOT + a];
      }
      // computeSwingTrajectoryBezier (FootSwingTrajectory.py:54-70)
      const float* p0 = st + MPC_ST_P0 + 3 * leg;
      const float* pf = st + MPC_ST_PF + 3 * leg;
      const float b = bez(swing), db = dbez(swing);
      for (int a = 0; a < 2; ++a) {
        st[MPC_ST_TRAJP + 3 * leg + a] = p0[a] + b * (pf[a] - p0[a]);
        st[MPC_ST_TRAJV + 3 * leg + a] = db * (pf[a] - p0[a]) / swing_time;
      }
      float zp, zv;
      if (swing < 0.5f) {
        const float x = swing * 2.0f, y0 = p0[2], yf = p0[2] + height;
        zp = y0 + bez(x) * (yf - y0);
        zv = dbez(x) * (yf - y0) * 2.0f / swing_time;
      } else {
        const float x = swing * 2.0f - 1.0f, y0 = p0[2] + height, yf = pf[2];
        zp = y0 + bez(x) * (yf - y0);
        zv = dbez(x) * (yf - y0) * 2.0f / swing_time;
      }
      st[MPC_ST_TRAJP + 3 * leg + 2] = zp;
      st[MPC_ST_TRAJV + 3 * leg + 2] = zv;
      kp[0] = 700.0f; kp[1] = 700.0f; kp[2] = 150.0f;
    } else {
      sti[MPC_ST_FIRSTSWING + leg] = 1;
      kp[0] = kp[1] = kp[2] = 0.0f;
      for (int a = 0; a < 3; ++a) fff[a] = st[MPC_ST_FFF + 3 * leg + a];
      kdj = 0.2f;
      cphase = contact;
    }
    const float hip[3] = {leg < 2 ? geom[MPC_GEOM_LOCX] : -geom[MPC_GEOM_LOCX],
                          (leg & 1) == 0 ? geom[MPC_GEOM_LOCY] : -geom[MPC_GEOM_LOCY], geom[MPC_GEOM_LOCZ]};
    float force[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float pos_a = a == 2 ? st[MPC_ST_POSZ] : 0.0f;
      const float pdes = (st[MPC_ST_TRAJP + 3 * leg + a] - pos_a) - hip[a];
      const float vdes = st[MPC_ST_TRAJV + 3 * leg + a] - est[MPC_EST_VBODY + a];
      force[a] = fff[a] + kp[a] * (pdes - L[LG_P + a]) + kd[a] * (vdes - L[LG_V + a]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float tau = L[LG_J + c] * force[0] + L[LG_J + 3 + c] * force[1] + L[LG_J + 6 + c] * force[2];   // J^T f
      tau += kdj * (0.0f - L[LG_QD + c]);
      torques[3 * leg + c] = tau;
    }
    st[MPC_ST_CPHASE + leg] = cphase;   // setContactPhase (ConvexMPCLocomotion.py:378)
  }
  LANES_END
  WSYNC();
}

}  // namespace mpc
