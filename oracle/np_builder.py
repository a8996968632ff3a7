"""Independent NumPy/SciPy transcriptions of the reference QP builder — TEST INFRASTRUCTURE.

Two restatements used to pin oracle/mpc_oracle.cc (whose builder half the reference's own
tests do not pin, see mpc_oracle.h):

* :func:`build_qp_literal` — a direct dense transcription of
  reference/MPC_Controller/convex_MPC/mpc_osqp.cc:593-692 using ``scipy.linalg.expm`` for
  ``ab_mat.exp()`` (mpc_osqp.cc:348) and the plain ``2*B_qp^T W B_qp + alpha*I`` product
  instead of the block recursion (mpc_osqp.cc:387-434).
* :func:`build_qp_structured` — the closed form the sm_100a kernel evaluates (DESIGN.md
  §"QP structure"): ``[[A,B],[0,0]]*dt`` is nilpotent of index 3, every ``A_exp^k B_exp``
  factors as ``diag(c_k,c_k,dt,dt) * [T*Bw; Bv; Bw; Bv]`` and therefore
  ``P = alpha*I + 2 (I (x) E)^T (Gamma (x) Q1 + N (x) Q2) (I (x) E)`` with ``E = [Bw; Bv]``
  (6x12), ``Q1 = blkdiag(T^T W_rpy T, W_pos)``, ``Q2 = dt^2 blkdiag(W_omega, W_vel)``.

Both keep the reference's quirks (truncated ``a_qp`` :361, ``Rx*Ry*Rz`` foot rotation
:606-617, friction indices :437-447).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm

K_GRAVITY = 9.8      # mpc_osqp.cc:54
K_MAX_SCALE = 10.0   # :55
K_MIN_SCALE = 0.1    # :56


def _rx(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def _ry(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _rz(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def _common(mass, inertia9, kw):
    rpy = np.asarray(kw["com_roll_pitch_yaw"], float)
    inertia = np.asarray(inertia9, float).reshape(3, 3).T          # Eigen col-major read (:516)
    inv_inertia = np.linalg.inv(inertia)
    R = _rz(rpy[2]) @ _ry(rpy[1]) @ _rx(rpy[0])                      # :283-291
    Rf = _rx(rpy[0]) @ _ry(rpy[1]) @ _rz(rpy[2])                     # :606-617
    feet = np.asarray(kw["foot_positions_body_frame"], float).reshape(4, 3) @ Rf.T
    Iw_inv = R @ inv_inertia @ R.T                                   # :670-671
    cy, sy, cp, tp = np.cos(rpy[2]), np.sin(rpy[2]), np.cos(rpy[1]), np.tan(rpy[1])
    T = np.array([[cy / cp, sy / cp, 0], [-sy, cy, 0], [cy * tp, sy * tp, 1]])  # :306-312
    Bw = np.hstack([Iw_inv @ _skew(feet[i]) for i in range(4)])      # 3x12
    Bv = np.hstack([np.eye(3) / mass for _ in range(4)])             # 3x12
    return rpy, T, Bw, Bv


def constraints(mass, h, kw):
    """mpc_osqp.cc:437-477, 685-692 -> dense C[20h,12h], l, u."""
    mu = np.asarray(kw["foot_friction_coeffs"], float)
    contact = np.asarray(kw["foot_contact_states"], float).reshape(h, 4)
    blk = np.array([[-1, 0, mu[0]], [1, 0, mu[1]], [0, -1, mu[2]], [0, 1, mu[3]], [0, 0, 1.0]])
    C = np.kron(np.eye(4 * h), blk)
    fz_max, fz_min = mass * K_GRAVITY * K_MAX_SCALE, mass * K_GRAVITY * K_MIN_SCALE
    c = contact.reshape(-1)
    l = np.zeros((4 * h, 5))
    u = np.zeros((4 * h, 5))
    l[:, 4] = fz_min * c
    u[:, :4] = ((mu[0] + 1) * fz_max * c)[:, None]
    u[:, 4] = fz_max * c
    return C, l.reshape(-1), u.reshape(-1)


def build_qp_literal(mass, inertia9, h, dt, alpha, full_a_qp=False, **kw):
    rpy, T, Bw, Bv = _common(mass, inertia9, kw)
    normal = np.asarray(kw["ground_normal_vec"], float)
    A = np.zeros((13, 13))
    A[0:3, 6:9] = T
    A[3:6, 9:12] = np.eye(3)
    A[9:12, 12] = normal
    B = np.zeros((13, 12))
    B[6:9] = Bw
    B[9:12] = Bv
    M = np.zeros((25, 25))
    M[:13, :13] = A * dt
    M[:13, 13:] = B * dt
    Mexp = expm(M)
    a_exp, b_exp = Mexp[:13, :13], Mexp[:13, 13:]
    a_qp = np.zeros((13 * h, 13))
    a_qp[:13] = a_exp
    # :361 (Q1): the reference's loop stops at horizon-1, the last a_qp block stays zero.  full_a_qp=True is
    # the "corrected" builder the reference does NOT implement (tests show the two differ visibly)
    for i in range(1, h if full_a_qp else h - 1):
        a_qp[13 * i:13 * (i + 1)] = a_exp @ a_qp[13 * (i - 1):13 * i]
    anb = [b_exp]
    for i in range(1, h):
        anb.append(a_exp @ anb[-1])
    b_qp = np.zeros((13 * h, 12 * h))
    for i in range(h):
        for j in range(i + 1):
            b_qp[13 * i:13 * (i + 1), 12 * j:12 * (j + 1)] = anb[i - j]
    w = np.asarray(kw["qp_weights"], float)
    W = np.diag(np.tile(w, h))
    P = 2 * b_qp.T @ W @ b_qp + alpha * np.eye(12 * h)
    pos, vel = np.asarray(kw["com_position"], float), np.asarray(kw["com_velocity"], float)
    om = np.asarray(kw["com_angular_velocity"], float)
    x0 = np.concatenate([rpy, pos, om, vel, [-K_GRAVITY]])
    dv, dw = np.asarray(kw["desired_com_velocity"], float), np.asarray(kw["desired_com_angular_velocity"], float)
    drpy, dpos = np.asarray(kw["desired_com_roll_pitch_yaw"], float), np.asarray(kw["desired_com_position"], float)
    xd = np.zeros(13 * h)
    for i in range(h):                                               # :635-659
        xd[13 * i:13 * (i + 1)] = [drpy[0], drpy[1], rpy[2] + dt * (i + 1) * dw[2],
                                   dt * (i + 1) * dv[0] + pos[0], dt * (i + 1) * dv[1] + pos[1], dpos[2],
                                   dw[0], dw[1], dw[2], dv[0], dv[1], 0.0, -K_GRAVITY]
    q = 2 * b_qp.T @ (W @ (a_qp @ x0 - xd))
    C, l, u = constraints(mass, h, kw)
    return dict(P=P, q=q, C=C, l=l, u=u, a_exp=a_exp, b_exp=b_exp)


def structure(mass, inertia9, h, dt, alpha, **kw):
    """The small factors of the closed form (what the kernel keeps per robot)."""
    rpy, T, Bw, Bv = _common(mass, inertia9, kw)
    w = np.asarray(kw["qp_weights"], float)
    E = np.vstack([Bw, Bv])                                          # 6x12
    Q1 = np.zeros((6, 6))
    Q1[:3, :3] = T.T @ np.diag(w[0:3]) @ T
    Q1[3:, 3:] = np.diag(w[3:6])
    Q2 = dt * dt * np.diag(np.concatenate([w[6:9], w[9:12]]))
    c = dt * dt * (np.arange(h) + 0.5)
    Gam = np.zeros((h, h))
    Nn = np.zeros((h, h))
    for i in range(h):
        for j in range(h):
            k = np.arange(max(i, j), h)
            Gam[i, j] = np.sum(c[k - i] * c[k - j])
            Nn[i, j] = h - max(i, j)
    return dict(E=E, T=T, Q1=Q1, Q2=Q2, Gam=Gam, N=Nn, c=c, rpy=rpy)


def build_qp_structured(mass, inertia9, h, dt, alpha, **kw):
    s = structure(mass, inertia9, h, dt, alpha, **kw)
    E, T, c, rpy = s["E"], s["T"], s["c"], s["rpy"]
    Eh = np.kron(np.eye(h), E)
    Q = np.kron(s["Gam"], s["Q1"]) + np.kron(s["N"], s["Q2"])
    P = alpha * np.eye(12 * h) + 2 * Eh.T @ Q @ Eh
    w = np.asarray(kw["qp_weights"], float)
    pos, vel = np.asarray(kw["com_position"], float), np.asarray(kw["com_velocity"], float)
    om = np.asarray(kw["com_angular_velocity"], float)
    normal = np.asarray(kw["ground_normal_vec"], float)
    dv, dw = np.asarray(kw["desired_com_velocity"], float), np.asarray(kw["desired_com_angular_velocity"], float)
    drpy, dpos = np.asarray(kw["desired_com_roll_pitch_yaw"], float), np.asarray(kw["desired_com_position"], float)
    g = -K_GRAVITY
    err = np.zeros((h, 12))                                          # e_k = A_exp^{k+1} x0 - xdes_k
    for k in range(h):
        t = dt * (k + 1)
        des = np.array([drpy[0], drpy[1], rpy[2] + t * dw[2], t * dv[0] + pos[0], t * dv[1] + pos[1], dpos[2],
                        dw[0], dw[1], dw[2], dv[0], dv[1], 0.0])
        if k < h - 1:
            pred = np.concatenate([rpy + t * (T @ om), pos + t * vel + 0.5 * t * t * normal * g, om,
                                   vel + t * normal * g])
        else:
            pred = np.zeros(12)                                      # Q1: last a_qp block is zero
        err[k] = pred - des
    q = np.zeros(12 * h)
    for j in range(h):
        acc = np.zeros(6)
        for k in range(j, h):
            m = k - j
            acc[:3] += c[m] * (T.T @ (w[0:3] * err[k, 0:3])) + dt * w[6:9] * err[k, 6:9]
            acc[3:] += c[m] * (w[3:6] * err[k, 3:6]) + dt * w[9:12] * err[k, 9:12]
        q[12 * j:12 * (j + 1)] = 2 * E.T @ acc
    C, l, u = constraints(mass, h, kw)
    return dict(P=P, q=q, C=C, l=l, u=u, **s)
