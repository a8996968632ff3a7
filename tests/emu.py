"""ctypes front-end of the CPU warp-emulation build (tests/warp_emu) — TEST INFRASTRUCTURE.

Builds tests/warp_emu/libmpc_emu.so on demand from the SAME warp program the CUDA kernel
compiles (rl_mpc_locomotion_b200/csrc/mpc_warp_solver.cuh, -DMPC_EMULATE).  Used by the
CPU test-suite to exercise the kernel logic without a GPU; never by the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "warp_emu", "emu_main.cc")
_SO = os.path.join(_HERE, "warp_emu", "libmpc_emu.so")
_HDRS = [os.path.join(_HERE, "..", "rl_mpc_locomotion_b200", "csrc", f)
         for f in ("mpc_warp_solver.cuh", "warp_prog.h", "mpc_tick.cuh")] + [os.path.join(_HERE, "..", "include", "mpc_b200.h")]


class SolveParams(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("check_every", C.c_int), ("scaling_iters", C.c_int),
                ("polish", C.c_int), ("polish_refine", C.c_int), ("adaptive_rho", C.c_int),
                ("fixed_iters", C.c_int), ("reserved", C.c_int),
                ("dt", C.c_double), ("alpha", C.c_double), ("rho0", C.c_double), ("sigma", C.c_double),
                ("relax", C.c_double), ("eps_abs", C.c_double), ("eps_rel", C.c_double),
                ("delta", C.c_double), ("rho_tol_ratio", C.c_double)]


def default_params(**kw):
    p = SolveParams(4000, 25, 10, 1, 3, 1, 0, 0, 0.02, 1e-5, 0.1, 1e-6, 1.6, 1e-3, 1e-3, 1e-6, 5.0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def build(force=False):
    newest = max(os.path.getmtime(f) for f in [_SRC] + _HDRS)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-DMPC_EMULATE", "-fPIC", "-shared",
                               "-Wno-unknown-pragmas", _SRC, "-o", _SO])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.mpc_emu_solve.restype = C.c_int
    return _lib


def solve(records, horizon, robot_params, robot_type=None, params=None, want_dbg=False, warm=None, split=False):
    """`split=True`: the phase-split path (SETUP / ADMM / POLISH as separate calls through the workspace row)."""
    rec = np.ascontiguousarray(records, dtype=np.float32)
    n = rec.shape[0]
    rp = np.ascontiguousarray(robot_params, dtype=np.float32)
    rt = None if robot_type is None else np.ascontiguousarray(robot_type, dtype=np.int32)
    prm = params or default_params()
    grf0 = np.zeros((n, 12), np.float32)
    full = np.zeros((n, 12 * horizon), np.float32)
    status = np.zeros(n, np.int32)
    info = np.zeros((n, 8))
    nv = 12 * horizon
    dbg = np.zeros((n, nv + nv * nv + nv + 20 * horizon + 1)) if want_dbg else None
    P = C.c_void_p
    if split:
        assert not want_dbg
        rc = lib().mpc_emu_solve_split(P(rec.ctypes.data), n, horizon, P(rp.ctypes.data),
                                       P(rt.ctypes.data) if rt is not None else None, C.byref(prm),
                                       P(grf0.ctypes.data), P(full.ctypes.data), P(status.ctypes.data),
                                       P(info.ctypes.data), P(warm.ctypes.data) if warm is not None else None)
        assert rc == 0
        return dict(grf0=grf0, grf_full=full, status=status, info=info)
    rc = lib().mpc_emu_solve(P(rec.ctypes.data), n, horizon, P(rp.ctypes.data),
                             P(rt.ctypes.data) if rt is not None else None, C.byref(prm),
                             P(grf0.ctypes.data), P(full.ctypes.data), P(status.ctypes.data),
                             P(info.ctypes.data), P(dbg.ctypes.data) if dbg is not None else None,
                             P(warm.ctypes.data) if warm is not None else None)
    assert rc == 0
    out = dict(grf0=grf0, grf_full=full, status=status, info=info)
    if want_dbg:
        out["q"] = dbg[:, :nv]
        out["P"] = dbg[:, nv:nv + nv * nv].reshape(n, nv, nv)
        out["D"] = dbg[:, nv + nv * nv: nv + nv * nv + nv]
        out["E"] = dbg[:, nv + nv * nv + nv: nv + nv * nv + nv + 20 * horizon]
        out["c"] = dbg[:, -1]
    return out


class TickParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("iters_between_mpc", C.c_int), ("flat_ground", C.c_int),
                ("bonus_swing", C.c_float), ("cmd_stride", C.c_int), ("reserved", C.c_int * 3),
                ("act_scale", C.c_float * 12), ("act_const", C.c_float * 12)]


ACT_SCALE = (4, 4, 4, 20, 20, 20, 1, 1, 1, 1, 1, 1)     # reference Parameters.MPC_param_scale
ACT_CONST = (5, 5, 5, 50, 50, 50, 1, 1, 1, 1, 1, 1)     # reference Parameters.MPC_param_const


def tick(state, dof, est, cmd, has_weights, gait, robot_params, robot_geom, robot_type, flat_ground=False,
         params=None, dt=0.01, iters_between_mpc=2, body=None, warm=None, actions=None, active=None):
    """One emulated control tick; `state` [N,96] float32 is updated in place.  Pass `body`
    ([N,13] root states) with est=None to run StateEstimator.update in the emulated kernel;
    `actions` [N,12] -> weights = actions * scale + const; `active` [N] MPC_TICK_* gate."""
    n = state.shape[0]
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    dof, cmd, rp, geom = f32(dof), f32(cmd), f32(robot_params), f32(robot_geom)
    est = None if est is None else f32(est)
    body = None if body is None else f32(body)
    gait = np.ascontiguousarray(gait, dtype=np.int32)
    rt = np.ascontiguousarray(robot_type, dtype=np.int32)
    assert state.dtype == np.float32 and state.flags.c_contiguous
    tp = TickParams(dt, iters_between_mpc, int(flat_ground), 0.0, cmd.shape[1])
    tp.act_scale[:] = ACT_SCALE
    tp.act_const[:] = ACT_CONST
    actions = None if actions is None else f32(actions)
    active = None if active is None else np.ascontiguousarray(active, dtype=np.int32)
    prm = params or default_params()
    torques = np.zeros((n, 12), np.float32)
    grf0 = np.zeros((n, 12), np.float32)
    records = np.zeros((n, 96), np.float32)
    solved = np.zeros(n, np.int32)
    P = C.c_void_p
    rc = lib().mpc_emu_tick(C.byref(prm), C.byref(tp), P(dof.ctypes.data),
                            P(body.ctypes.data) if body is not None else None,
                            P(est.ctypes.data) if est is not None else None, P(cmd.ctypes.data),
                            int(has_weights), P(actions.ctypes.data) if actions is not None else None,
                            P(gait.ctypes.data), P(rp.ctypes.data), P(geom.ctypes.data),
                            P(rt.ctypes.data), n, P(active.ctypes.data) if active is not None else None,
                            P(state.ctypes.data), P(torques.ctypes.data), P(grf0.ctypes.data),
                            P(records.ctypes.data), P(solved.ctypes.data),
                            P(warm.ctypes.data) if warm is not None else None)
    assert rc == 0
    return dict(torques=torques, grf0=grf0, records=records, solved=solved)


def chol_inv(A: np.ndarray, pcap: int = 60) -> np.ndarray:
    """W = chol(A)^-1 through the kernel's blocked Cholesky + triangular inverse lane programs."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    W = np.zeros((n, n))
    rc = lib().mpc_emu_chol_inv(A.ctypes.data_as(C.c_void_p), n, pcap, W.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    return W
