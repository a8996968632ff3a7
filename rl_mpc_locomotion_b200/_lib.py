"""ctypes binding of the C-ABI library ``libmpc_b200.so`` (include/mpc_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
CPU fallback: if the shared object is missing, or no CUDA device is visible, every solve
raises :class:`MpcB200Error`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPC_B200_LIB", os.path.join(_HERE, "libmpc_b200.so"))  # env override: A/B builds

MODE_OSQP, MODE_FIXED = 0, 1
PATH_AUTO, PATH_MONO, PATH_SPLIT = 0, 1, 2     # MPC_PATH_* (Settings.kernel_path)
STATUS_SOLVED, STATUS_MAX_ITER, STATUS_NON_FINITE, STATUS_POLISH_OVERFLOW, STATUS_BAD_INDEX = 1, -2, -7, -90, -91
TICK_SKIP, TICK_RUN, TICK_CHECK_RUN, TICK_CHECK_ONLY = 0, 1, 2, 3     # MPC_TICK_* (d_active gate)
# reference Parameters.MPC_param_scale / MPC_param_const (Parameters.py:25-33)
ACT_SCALE = (4.0, 4.0, 4.0, 20.0, 20.0, 20.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)
ACT_CONST = (5.0, 5.0, 5.0, 50.0, 50.0, 50.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)
EXPORTS = ("mpc_b200_default_settings", "mpc_b200_version", "mpc_b200_last_error",
           "mpc_b200_device_count", "mpc_b200_kernel_info", "mpc_b200_solve_batch",
           "mpc_b200_solve_batch_host", "mpc_b200_debug_build_qp", "mpc_b200_debug_estimator", "mpc_b200_tick")


class MpcB200Error(RuntimeError):
    pass


class Settings(C.Structure):
    """Mirror of ``MpcB200Settings`` (include/mpc_b200.h)."""
    _fields_ = [("horizon", C.c_int), ("mode", C.c_int), ("admm_iters", C.c_int), ("max_iter", C.c_int),
                ("check_every", C.c_int), ("scaling_iters", C.c_int), ("adaptive_rho", C.c_int),
                ("polish", C.c_int), ("polish_refine", C.c_int), ("kernel_path", C.c_int),
                ("reserved_i", C.c_int * 2),
                ("dt", C.c_double), ("alpha", C.c_double), ("rho", C.c_double), ("sigma", C.c_double),
                ("relax", C.c_double), ("eps_abs", C.c_double), ("eps_rel", C.c_double),
                ("delta", C.c_double), ("rho_tol_ratio", C.c_double), ("reserved_d", C.c_double * 3)]


class TickParams(C.Structure):
    """Mirror of ``MpcB200TickParams``."""
    _fields_ = [("dt", C.c_float), ("iters_between_mpc", C.c_int), ("flat_ground", C.c_int),
                ("bonus_swing", C.c_float), ("cmd_stride", C.c_int), ("reserved", C.c_int * 3),
                ("act_scale", C.c_float * 12), ("act_const", C.c_float * 12)]


def tick_params(dt=0.01, iters_between_mpc=2, flat_ground=False, bonus_swing=0.0, cmd_stride=16,
                act_scale=ACT_SCALE, act_const=ACT_CONST) -> TickParams:
    tp = TickParams(dt, iters_between_mpc, int(flat_ground), bonus_swing, cmd_stride)
    tp.act_scale[:] = act_scale
    tp.act_const[:] = act_const
    return tp


_lib = None


def lib():
    """Load the CUDA library (once).  Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MpcB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.mpc_b200_default_settings.argtypes = [C.POINTER(Settings)]
        L.mpc_b200_default_settings.restype = None
        L.mpc_b200_version.restype = C.c_char_p
        L.mpc_b200_last_error.restype = C.c_char_p
        L.mpc_b200_device_count.restype = C.c_int
        L.mpc_b200_kernel_info.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.mpc_b200_kernel_info.restype = C.c_int
        L.mpc_b200_solve_batch.argtypes = [C.POINTER(Settings), vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                           vp, vp]
        L.mpc_b200_solve_batch.restype = C.c_int
        L.mpc_b200_solve_batch_host.argtypes = [C.POINTER(Settings), vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]
        L.mpc_b200_solve_batch_host.restype = C.c_int
        L.mpc_b200_debug_build_qp.argtypes = [C.POINTER(Settings), vp, vp, vp, vp, vp, vp, vp]
        L.mpc_b200_debug_build_qp.restype = C.c_int
        L.mpc_b200_debug_estimator.argtypes = [vp, vp, vp, C.c_int]
        L.mpc_b200_debug_estimator.restype = C.c_int
        L.mpc_b200_tick.argtypes = [C.POINTER(Settings), C.POINTER(TickParams), vp, vp, vp, vp, C.c_int, vp, vp, vp, vp,
                                    vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.mpc_b200_tick.restype = C.c_int
        _lib = L
    return _lib


def default_settings(**overrides) -> Settings:
    s = Settings()
    lib().mpc_b200_default_settings(C.byref(s))
    for k, v in overrides.items():
        if not hasattr(s, k):
            raise AttributeError(f"MpcB200Settings has no field {k!r}")
        setattr(s, k, v)
    return s


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().mpc_b200_last_error().decode()
        raise MpcB200Error(f"{what} failed (code {rc}): {msg}")


def kernel_info(horizon: int) -> dict:
    smem, rpc, regs = C.c_int(), C.c_int(), C.c_int()
    check(lib().mpc_b200_kernel_info(horizon, C.byref(smem), C.byref(rpc), C.byref(regs)), "kernel_info")
    return {"smem_bytes": smem.value, "robots_per_cta": rpc.value, "regs_per_thread": regs.value}
