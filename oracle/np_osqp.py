"""Dense NumPy restatement of OSQP 0.6.0 as the reference configures it — TEST INFRASTRUCTURE.

This module documents, in ~200 lines of dense linear algebra, exactly which algorithm the
sm_100a kernel has to reproduce in order to return what the reference's OSQP path returns
(reference/MPC_Controller/convex_MPC/mpc_osqp.cc:690-796 with the vendored
reference/extern/osqp).  It is validated against the REAL OSQP (oracle/_ref/libosqp_ref.so)
in tests/test_oracle.py; it is never used by the product.

Followed sources (all under reference/extern/osqp):
  scaling        src/scaling.c:45-160 (Ruiz, 10 passes, KKT column inf-norms, cost scaling)
  rho vector     src/auxil.c:79-101   (equality rows get 1e3*rho, decided on SCALED bounds)
  ADMM step      src/auxil.c:164-228, lin_sys/direct/qdldl/qdldl_interface.c:334-368
  residuals      src/auxil.c:243-362  (unscaled inf-norms: scaled_termination = 0)
  termination    src/auxil.c:684-790  (every check_termination = 25 iterations)
  adaptive rho   src/auxil.c:13-77, src/osqp.c:489-516 (interval 25, tolerance 5)
  polish         src/polish.c:20-350  (delta = 1e-6, 3 refinement steps)
  unscale        src/scaling.c:183-199
  settings       include/constants.h:59-110 overridden by mpc_osqp.cc:705-712
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

RHO, SIGMA, MAX_ITER, ALPHA = 0.1, 1e-6, 4000, 1.6
RHO_MIN, RHO_MAX, RHO_EQ_OVER_RHO_INEQ, RHO_TOL = 1e-6, 1e6, 1e3, 1e-4
MIN_SCALING, MAX_SCALING, SCALING_ITERS = 1e-4, 1e4, 10
CHECK_TERMINATION, ADAPTIVE_RHO_INTERVAL, ADAPTIVE_RHO_TOLERANCE = 25, 25, 5.0
DELTA, POLISH_REFINE_ITER = 1e-6, 3
OSQP_INFTY = 1e30
EPS_PRIM_INF = EPS_DUAL_INF = 1e-4


def _limit_scaling(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.where(v > MAX_SCALING, MAX_SCALING, v)


@dataclass
class Scaling:
    D: np.ndarray
    E: np.ndarray
    c: float


def scale_data(P, q, A):
    """src/scaling.c:45-160.  Returns (Pbar, qbar, Abar, Scaling)."""
    n, m = P.shape[0], A.shape[0]
    P, q, A = P.copy(), q.copy(), A.copy()
    D, E, c = np.ones(n), np.ones(m), 1.0
    for _ in range(SCALING_ITERS):
        d = np.maximum(np.abs(P).max(axis=0), np.abs(A).max(axis=0) if m else 0.0)
        e = np.abs(A).max(axis=1) if m else np.zeros(0)
        d = 1.0 / np.sqrt(_limit_scaling(d))
        e = 1.0 / np.sqrt(_limit_scaling(e))
        P = d[:, None] * P * d[None, :]
        A = e[:, None] * A * d[None, :]
        q = d * q
        D, E = D * d, E * e
        c_temp = np.abs(P).max(axis=0).mean()
        nq = float(_limit_scaling(np.array([np.abs(q).max()]))[0])
        c_temp = float(_limit_scaling(np.array([max(c_temp, nq)]))[0])
        c_temp = 1.0 / c_temp
        P, q, c = P * c_temp, q * c_temp, c * c_temp
    return P, q, A, Scaling(D, E, c)


def rho_vec_of(l, u, rho):
    """src/auxil.c:79-101 on the SCALED bounds."""
    loose = (l < -OSQP_INFTY * MIN_SCALING) & (u > OSQP_INFTY * MIN_SCALING)
    eq = (u - l < RHO_TOL) & ~loose
    return np.where(loose, RHO_MIN, np.where(eq, RHO_EQ_OVER_RHO_INEQ * rho, rho)), eq.astype(int) - loose.astype(int)


@dataclass
class Result:
    x: np.ndarray            # unscaled primal solution (NOT negated)
    y: np.ndarray
    iters: int
    status: int              # 1 solved, -2 max iter
    polish_status: int       # 1 ok, -1 failed, 0 not run
    rho_updates: int
    rho: float
    info: dict


def solve(P, q, A, l, u, eps_abs=1e-3, eps_rel=1e-3, polish=True, max_iter=MAX_ITER,
          warm=None, linsolve=None):
    """Cold-start osqp_setup + osqp_solve with the reference's settings.

    ``linsolve(K_builder)`` may be supplied to test an alternative (structured) KKT solve.
    """
    n, m = P.shape[0], A.shape[0]
    Pb, qb, Ab, sc = scale_data(P, q, A)
    lb = sc.E * np.maximum(l, -OSQP_INFTY)
    ub = sc.E * np.minimum(u, OSQP_INFTY)
    Dinv, Einv, cinv = 1.0 / sc.D, 1.0 / sc.E, 1.0 / sc.c
    rho = min(max(RHO, RHO_MIN), RHO_MAX)
    rho_vec, ctype = rho_vec_of(lb, ub, rho)

    def factor(rv):
        K = Pb + SIGMA * np.eye(n) + Ab.T @ (rv[:, None] * Ab)
        return np.linalg.cholesky(K)

    def ksolve(Lc, rhs):
        return np.linalg.solve(Lc.T, np.linalg.solve(Lc, rhs))

    Lc = factor(rho_vec)
    x, z, y = np.zeros(n), np.zeros(m), np.zeros(m)
    if warm is not None:
        x, z, y = (w.copy() for w in warm)
    rho_updates, status, it_done = 0, -10, 0
    pri_res = dua_res = np.inf
    Ax = Px = Aty = None
    for it in range(1, max_iter + 1):
        x_prev, z_prev = x, z
        # update_xz_tilde (auxil.c:164-186) with nu eliminated
        rhs = SIGMA * x_prev - qb + Ab.T @ (rho_vec * z_prev - y)
        xt = ksolve(Lc, rhs)
        zt = Ab @ xt
        x = ALPHA * xt + (1 - ALPHA) * x_prev                       # update_x
        zr = ALPHA * zt + (1 - ALPHA) * z_prev
        z = np.clip(zr + y / rho_vec, lb, ub)                       # update_z + project
        y = y + rho_vec * (zr - z)                                  # update_y
        it_done = it
        check = (it % CHECK_TERMINATION == 0)
        if check or it % ADAPTIVE_RHO_INTERVAL == 0:
            Ax, Px, Aty = Ab @ x, Pb @ x, Ab.T @ y
            rp, rd = Ax - z, qb + Px + Aty
            pri_res = np.abs(Einv * rp).max()
            dua_res = cinv * np.abs(Dinv * rd).max()
        if check:
            eps_prim = eps_abs + eps_rel * max(np.abs(Einv * z).max(), np.abs(Einv * Ax).max())
            eps_dual = eps_abs + eps_rel * cinv * max(np.abs(Dinv * qb).max(), np.abs(Dinv * Aty).max(),
                                                      np.abs(Dinv * Px).max())
            if pri_res < eps_prim and dua_res < eps_dual:
                status = 1
                break
        if it % ADAPTIVE_RHO_INTERVAL == 0:
            # compute_rho_estimate (auxil.c:13-55): SCALED residual vectors
            pr = np.abs(rp).max() / (max(np.abs(z).max(), np.abs(Ax).max()) + 1e-10)
            dr = np.abs(rd).max() / (max(np.abs(qb).max(), np.abs(Aty).max(), np.abs(Px).max()) + 1e-10)
            rho_new = min(max(rho * np.sqrt(pr / (dr + 1e-10)), RHO_MIN), RHO_MAX)
            if rho_new > rho * ADAPTIVE_RHO_TOLERANCE or rho_new < rho / ADAPTIVE_RHO_TOLERANCE:
                rho = rho_new
                rho_vec = np.where(ctype == 0, rho, np.where(ctype == 1, RHO_EQ_OVER_RHO_INEQ * rho, rho_vec))
                Lc = factor(rho_vec)
                rho_updates += 1
    if status != 1:
        status = -2
    polish_status = 0
    info = dict(pri_res=pri_res, dua_res=dua_res)
    if polish and status == 1:
        # polish.c:20-95 active-set guess
        low = (z - lb) < -y
        upp = (ub - z) > -np.inf
        upp = (ub - z) < y
        idx_low, idx_upp = np.nonzero(low)[0], np.nonzero(upp)[0]
        Ared = np.vstack([Ab[idx_low], Ab[idx_upp]]) if (len(idx_low) + len(idx_upp)) else np.zeros((0, n))
        mred = Ared.shape[0]
        Kreg = np.block([[Pb + DELTA * np.eye(n), Ared.T], [Ared, -DELTA * np.eye(mred)]])
        Ktrue = np.block([[Pb, Ared.T], [Ared, np.zeros((mred, mred))]])
        b = np.concatenate([-qb, lb[idx_low], ub[idx_upp]])
        sol = np.linalg.solve(Kreg, b)
        for _ in range(POLISH_REFINE_ITER):
            sol = sol + np.linalg.solve(Kreg, b - Ktrue @ sol)
        xp = sol[:n]
        zp = Ab @ xp
        yp = np.zeros(m)
        yp[idx_low] = sol[n:n + len(idx_low)]
        # rows that are both lower- and upper-active: get_ypol_from_yred takes Alow first
        only_upp = np.array([j for j in idx_upp if not low[j]], dtype=int)
        pos = {j: k for k, j in enumerate(idx_upp)}
        for j in only_upp:
            yp[j] = sol[n + len(idx_low) + pos[j]]
        # project_normalcone (proj.c:17-31)
        tmp = zp + yp
        zp = np.clip(tmp, lb, ub)
        yp = tmp - zp
        pri_p = np.abs(Einv * (Ab @ xp - zp)).max()
        dua_p = cinv * np.abs(Dinv * (qb + Pb @ xp + Ab.T @ yp)).max()
        ok = ((pri_p < pri_res and dua_p < dua_res) or (pri_p < pri_res and dua_res < 1e-10)
              or (dua_p < dua_res and pri_res < 1e-10))
        info.update(pri_pol=pri_p, dua_pol=dua_p, n_low=len(idx_low), n_upp=len(idx_upp))
        if ok:
            x, z, y, polish_status = xp, zp, yp, 1
        else:
            polish_status = -1
    return Result(sc.D * x, sc.E * y * cinv, it_done, status, polish_status, rho_updates, rho,
                  dict(info, scaled=(x, z, y), scaling=sc))
