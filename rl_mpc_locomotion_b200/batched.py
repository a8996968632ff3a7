"""Batched stance-force solve on device tensors (the new entry point).

No reference counterpart: the reference loops ``controller.run`` over robots on the host
(reference/RL_Environment/tasks/aliengo.py:247-258).  Here the whole batch of problem
records (one per robot, the 13 ``compute_contact_forces`` arguments packed as in
include/mpc_b200.h) is solved by one kernel launch, one warp per robot.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .records import ROBOT_PARAMS_STRIDE, rec_stride
from .robots import params_table


# A fixed iteration budget that actually converges: K ADMM iterations at rho = 1e-4 on the
# UNSCALED problem, no termination test / rho adaptation / polish.  On the BASELINE config-2 distribution K = 100
# lands at a median 1.5e-3 N from the TRUE minimiser (90th percentile 0.7-1.7 N, worst 35-87 N depending on the
# sample: the optimal rho varies from QP to QP) - closer than the reference's own loosely converged OSQP output
# (median 8-9 N, worst 120-200 N) - at ~2.3 M solves/s on one B200 (tests/test_warp_emu.py pins the contract).
# NOT the reference-parity mode (that is the default): `BatchedConvexMpc(h, **FIXED_BUDGET_CONVERGING)`.
FIXED_BUDGET_CONVERGING = dict(mode=_lib.MODE_FIXED, admm_iters=100, rho=1e-4, scaling_iters=0, polish=0)


class BatchedConvexMpc:
    """Batched equivalent of ``mpc_osqp.ConvexMpc`` (mpc_osqp.cc:178-279).

    ``robot_params``: ``[n_types, 12]`` float32 rows (mass, inertia[9] row-major, 2 pad) — what
    the reference passes to the ``ConvexMpc`` constructor per robot; defaults to the
    reference's Aliengo/A1/Go1 table (Quadruped.py:16-73).
    """

    def __init__(self, horizon: int = 10, device="cuda", robot_params=None, **settings):
        self.settings = _lib.default_settings(horizon=horizon, **settings)
        self.horizon = horizon
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MpcB200Error("BatchedConvexMpc needs a CUDA device: there is no CPU path")
        rp = params_table() if robot_params is None else np.asarray(robot_params, dtype=np.float32)
        assert rp.ndim == 2 and rp.shape[1] == ROBOT_PARAMS_STRIDE
        self.robot_params = torch.from_numpy(np.ascontiguousarray(rp)).to(self.device)
        self.stride = rec_stride(horizon)

    def warm_buffer(self, n: int) -> torch.Tensor:
        """Fresh (all-cold) OSQP workspace-state tensor for `n` robots: pass it as ``warm=`` to
        consecutive solves of the same robots to get the reference's ``warm_start`` behaviour
        (mpc_osqp.cc:708, :760-778); zero a row to force ``osqp_setup`` semantics for that robot."""
        return torch.zeros((n, 2 + 76 * self.horizon), dtype=torch.float64, device=self.device)

    def solve(self, records: torch.Tensor, robot_type: torch.Tensor | None = None, *, full: bool = False,
              status: bool = True, info: bool = False, out: torch.Tensor | None = None,
              warm: torch.Tensor | None = None, order: torch.Tensor | None = None) -> dict:
        """``records`` [N, stride] float32 CUDA tensor -> dict(grf0[N,12], grf_full?, status?, resid, info?).

        ``grf0`` is the negated horizon-0 QP solution, i.e. exactly ``result[0:12]`` of the
        reference's ``compute_contact_forces`` (mpc_osqp.cc:789-790).  Asynchronous: runs on
        the current torch stream.
        """
        assert records.is_cuda and records.dtype == torch.float32 and records.is_contiguous()
        assert records.dim() == 2 and records.shape[1] == self.stride, (records.shape, self.stride)
        n = records.shape[0]
        dev = records.device
        grf0 = out if out is not None else torch.empty((n, 12), dtype=torch.float32, device=dev)
        res = {"grf0": grf0}
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        gfull = torch.empty((n, 12 * self.horizon), dtype=torch.float32, device=dev) if full else None
        st = torch.empty((n,), dtype=torch.int32, device=dev) if status else None
        resid = torch.empty((n, 2), dtype=torch.float32, device=dev) if status else None
        inf = torch.empty((n, 8), dtype=torch.float64, device=dev) if info else None
        if robot_type is not None:
            assert robot_type.is_cuda and robot_type.dtype == torch.int32 and robot_type.shape == (n,)
        if warm is not None:
            assert warm.is_cuda and warm.dtype == torch.float64 and warm.is_contiguous()
            assert tuple(warm.shape) == (n, 2 + 76 * self.horizon), warm.shape
        if order is not None:   # scheduling hint (longest-first), see lpt_order()
            assert order.is_cuda and order.dtype == torch.int32 and order.shape == (n,) and order.is_contiguous()
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = _lib.lib().mpc_b200_solve_batch(C.byref(self.settings), ptr(records), ptr(self.robot_params),
                                                 ptr(robot_type), int(self.robot_params.shape[0]), n, ptr(grf0),
                                                 ptr(gfull), ptr(st),
                                                 ptr(resid), ptr(inf), ptr(warm), ptr(order), C.c_void_p(stream))
        _lib.check(rc, "mpc_b200_solve_batch")
        if full:
            res["grf_full"] = gfull
        if status:
            res["status"], res["resid"] = st, resid
        if info:
            res["info"] = inf
        return res


def lpt_order(info: torch.Tensor) -> torch.Tensor:
    """Launch order for the NEXT solve of the same robots from THIS solve's ``info``: robots sorted
    by descending OSQP iteration count (longest-processing-time-first).  Consecutive control ticks
    of a robot need similar iteration counts, so this shortens the tail of the launch; results are
    unaffected."""
    return torch.argsort(info[:, 0], descending=True, stable=True).to(torch.int32)


def solve_host(records: np.ndarray, horizon: int = 10, robot_params=None, robot_type=None, full=False,
               **settings) -> dict:
    """Host-buffer path through ``mpc_b200_solve_batch_host`` (H2D + solve + D2H, synchronous)."""
    s = _lib.default_settings(horizon=horizon, **settings)
    rec = np.ascontiguousarray(records, dtype=np.float32)
    n = rec.shape[0]
    assert rec.shape[1] == rec_stride(horizon)
    rp = np.ascontiguousarray(params_table() if robot_params is None else robot_params, dtype=np.float32).reshape(-1, 12)
    rt = None if robot_type is None else np.ascontiguousarray(robot_type, dtype=np.int32)
    grf0 = np.empty((n, 12), np.float32)
    gfull = np.empty((n, 12 * horizon), np.float32) if full else None
    status = np.empty((n,), np.int32)
    vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
    rc = _lib.lib().mpc_b200_solve_batch_host(C.byref(s), vp(rec), vp(rp), vp(rt), int(rp.shape[0]), n, vp(grf0),
                                              vp(gfull), vp(status))
    _lib.check(rc, "mpc_b200_solve_batch_host")
    out = {"grf0": grf0, "status": status}
    if full:
        out["grf_full"] = gfull
    return out


def debug_build_qp(record: np.ndarray, horizon: int, robot_params_row: np.ndarray, **settings) -> dict:
    """Dense P, q and OSQP scaling (D, E, c) as evaluated on the device (tests only)."""
    s = _lib.default_settings(horizon=horizon, **settings)
    nv = 12 * horizon
    rec = np.ascontiguousarray(record, dtype=np.float32)
    rp = np.ascontiguousarray(robot_params_row, dtype=np.float32)
    P, q, D, E, c = np.zeros((nv, nv)), np.zeros(nv), np.zeros(nv), np.zeros(20 * horizon), np.zeros(1)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    rc = _lib.lib().mpc_b200_debug_build_qp(C.byref(s), vp(rec), vp(rp), vp(P), vp(q), vp(D), vp(E), vp(c))
    _lib.check(rc, "mpc_b200_debug_build_qp")
    return dict(P=P, q=q, D=D, E=E, c=float(c[0]))


def solve_host_pinned(records, robot_type, robot_params, horizon, steps, warmup=2, **settings):
    """End-to-end timing helper (bench.py `e2e`): the C-ABI host entry point fed from PINNED
    host buffers; every step copies the records H2D, solves, and copies forces + status D2H.
    Returns (seconds for `steps` steps, h2d bytes per step, d2h bytes per step)."""
    import time
    s = _lib.default_settings(horizon=horizon, **settings)
    n = records.shape[0]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    t_rec, t_rt, t_rp = pin(records.astype(np.float32)), pin(robot_type.astype(np.int32)), pin(
        np.asarray(robot_params, dtype=np.float32))
    t_grf = torch.empty((n, 12), dtype=torch.float32).pin_memory()
    t_st = torch.empty((n,), dtype=torch.int32).pin_memory()
    vp = lambda t: C.c_void_p(t.data_ptr())

    def one():
        rc = _lib.lib().mpc_b200_solve_batch_host(C.byref(s), vp(t_rec), vp(t_rp), vp(t_rt), int(t_rp.shape[0]), n,
                                                  vp(t_grf), None, vp(t_st))
        _lib.check(rc, "mpc_b200_solve_batch_host")

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    h2d = t_rec.numel() * 4 + t_rt.numel() * 4 + t_rp.numel() * 4
    d2h = t_grf.numel() * 4 + t_st.numel() * 4
    return dt, h2d, d2h
