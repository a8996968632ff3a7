"""ctypes front-end of the CPU oracle — TEST INFRASTRUCTURE ONLY (see mpc_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  It never reads /root/reference at run time: it loads the
prebuilt oracle/libmpc_oracle.so (+ oracle/_ref/libosqp_ref.so).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
COLD, WARM, TRUE = 0, 1, 2
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libmpc_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle not built: run `make -C oracle` (needs /root/reference)")
        L = C.CDLL(path)
        dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)
        L.mpc_oracle_create.restype = C.c_void_p
        L.mpc_oracle_create.argtypes = [C.c_double, dp, C.c_int, C.c_int, C.c_double, C.c_double]
        L.mpc_oracle_destroy.argtypes = [C.c_void_p]
        L.mpc_oracle_reset_solver.argtypes = [C.c_void_p]
        L.mpc_oracle_compute_contact_forces.restype = C.c_int
        L.mpc_oracle_compute_contact_forces.argtypes = [C.c_void_p, C.c_int] + [dp] * 14
        L.mpc_oracle_build_qp.restype = C.c_int
        L.mpc_oracle_build_qp.argtypes = [C.c_void_p] + [dp] * 13
        L.mpc_oracle_get_qp.argtypes = [C.c_void_p] + [dp] * 5
        L.mpc_oracle_get_exponentials.argtypes = [C.c_void_p, dp, dp]
        L.mpc_oracle_last_iters.restype = C.c_int
        L.mpc_oracle_last_iters.argtypes = [C.c_void_p]
        L.mpc_oracle_last_polish_status.restype = C.c_int
        L.mpc_oracle_last_polish_status.argtypes = [C.c_void_p]
        L.mpc_oracle_last_solve_time.restype = C.c_double
        L.mpc_oracle_last_solve_time.argtypes = [C.c_void_p]
        L.mpc_oracle_solve_records.restype = C.c_int
        L.mpc_oracle_solve_records.argtypes = [fp, C.c_int, C.c_int, C.c_double, C.c_double, fp,
                                               ip, C.c_int, dp, dp, ip]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


ARG_ORDER = ("qp_weights", "com_position", "com_velocity", "com_roll_pitch_yaw",
             "ground_normal_vec", "com_angular_velocity", "foot_contact_states",
             "foot_positions_body_frame", "foot_friction_coeffs", "desired_com_position",
             "desired_com_velocity", "desired_com_roll_pitch_yaw", "desired_com_angular_velocity")


class OracleMpc:
    """One ConvexMpc-equivalent (mpc_osqp.cc:178-279) backed by the oracle library."""

    def __init__(self, mass, inertia9, horizon=10, dt=0.02, alpha=1e-5, num_legs=4):
        self.h, self.legs = horizon, num_legs
        inertia = np.ascontiguousarray(inertia9, dtype=np.float64).reshape(9)
        self._o = lib().mpc_oracle_create(float(mass), _dp(inertia), num_legs, horizon, float(dt), float(alpha))

    def __del__(self):
        if getattr(self, "_o", None):
            lib().mpc_oracle_destroy(self._o)
            self._o = None

    def _args(self, kw):
        arrs = [np.ascontiguousarray(kw[k], dtype=np.float64).reshape(-1) for k in ARG_ORDER]
        return arrs, [_dp(a) for a in arrs]

    def solve(self, mode=COLD, **kw):
        """-> (status, forces[12h] float64)  (forces = negated QP solution)."""
        arrs, ptrs = self._args(kw)
        out = np.zeros(3 * self.legs * self.h)
        st = lib().mpc_oracle_compute_contact_forces(self._o, mode, *ptrs, _dp(out))
        return st, out

    def build(self, **kw):
        """-> dict(P[n,n], q[n], C[m,n], l[m], u[m])."""
        arrs, ptrs = self._args(kw)
        lib().mpc_oracle_build_qp(self._o, *ptrs)
        n, m = 3 * self.legs * self.h, 5 * self.legs * self.h
        P, q, Cm, l, u = np.zeros((n, n)), np.zeros(n), np.zeros((m, n)), np.zeros(m), np.zeros(m)
        lib().mpc_oracle_get_qp(self._o, _dp(P), _dp(q), _dp(Cm), _dp(l), _dp(u))
        return dict(P=P, q=q, C=Cm, l=l, u=u)

    def exponentials(self):
        a, b = np.zeros((13, 13)), np.zeros((13, 3 * self.legs))
        lib().mpc_oracle_get_exponentials(self._o, _dp(a), _dp(b))
        return a, b

    @property
    def last_iters(self):
        return lib().mpc_oracle_last_iters(self._o)

    @property
    def last_polish_status(self):
        return lib().mpc_oracle_last_polish_status(self._o)


def solve_records(records, horizon, robot_params, robot_type=None, mode=COLD, dt=0.02, alpha=1e-5,
                  full=False):
    """Batch front-end: -> (grf0[N,12] f64, grf_full[N,12h] f64 | None, status[N] i32)."""
    rec = np.ascontiguousarray(records, dtype=np.float32)
    n = rec.shape[0]
    rp = np.ascontiguousarray(robot_params, dtype=np.float32)
    rt = None if robot_type is None else np.ascontiguousarray(robot_type, dtype=np.int32)
    grf0 = np.zeros((n, 12))
    gfull = np.zeros((n, 12 * horizon)) if full else None
    status = np.zeros(n, dtype=np.int32)
    lib().mpc_oracle_solve_records(
        rec.ctypes.data_as(C.POINTER(C.c_float)), n, horizon, dt, alpha,
        rp.ctypes.data_as(C.POINTER(C.c_float)),
        None if rt is None else rt.ctypes.data_as(C.POINTER(C.c_int)), mode, _dp(grf0),
        None if gfull is None else _dp(gfull), status.ctypes.data_as(C.POINTER(C.c_int)))
    return grf0, gfull, status
