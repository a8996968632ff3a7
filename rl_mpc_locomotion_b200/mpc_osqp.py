"""Drop-in replacement of the reference's native module ``mpc_osqp``.

Same Python surface as reference/MPC_Controller/convex_MPC/mpc_osqp.cc:950-983:
``ConvexMpc(mass, inertia, num_legs, planning_horizon, timestep, alpha, qp_solver_name)``,
``.compute_contact_forces(13 sequences) -> list[float]`` (length ``3*num_legs*horizon``, the
NEGATED QP solution, empty list when the solve did not reach OSQP_SOLVED —
mpc_osqp.cc:786-794), ``.reset_solver()``, enum ``QPSolverName`` with exported values
``OSQP`` / ``QPOASES``, ``__version__``, ``TEST == 42``.

Backed by the sm_100a kernel (batch of one through the C-ABI).  Put
``rl_mpc_locomotion_b200/dropin`` on ``sys.path`` to make ``import mpc_osqp`` resolve here
(INTEGRATION.md).

State across calls, as in the reference object:

* ``OSQP``: the object owns ONE persistent OSQP workspace-state row on the device (``d_warm`` of
  include/mpc_b200.h).  The first call is ``osqp_setup`` + cold solve, every later call is the
  reference's ``osqp_update_P_A`` / ``update_lin_cost`` / ``update_bounds`` / warm-started
  ``osqp_solve`` (mpc_osqp.cc:757-780): warm x / y, the adapted rho, and the cost scaling that still
  sees the previous q.  ``reset_solver()`` only sets the reference's ``initial_run_`` flag
  (mpc_osqp.cc:576), which the OSQP branch never reads (it tests ``workspace_ == 0``,
  mpc_osqp.cc:757): the workspace survives, as here.  A new ``ConvexMpc`` object starts cold.
* ``QPOASES`` (what the reference's Python passes, ConvexMPCLocomotion.py:108): the reference's
  qpOASES branch (mpc_osqp.cc:797-947, source absent from the checkout) is an active-set solver
  that returns the QP's unique minimiser, cold, on every call.  The north-star forbids dispatching
  to it; this enum value runs the same device solver at ``eps_abs = eps_rel = 1e-8`` (no warm
  start), which lands within 1e-2 N of that minimiser (tests/test_gpu_parity.py checks it against
  the oracle's 1e-10 solve) — OSQP's default-tolerance answer is up to ~150 N away from it.

Precision at the boundary: the reference binds ``std::vector<double>`` (mpc_osqp.cc:202-215), but its
own caller passes float32 arrays and one float16 array (ConvexMPCLocomotion.py:128-185).
The device problem record is float32 (one 384-byte TMA transaction per robot), so arguments are
rounded to float32 here.  That is lossless for everything the reference's Python sends; a caller
that passes float64 values which are not float32-representable gets a one-time RuntimeWarning.
"""
from __future__ import annotations

import warnings
from enum import IntEnum

import numpy as np

from . import _lib
from .records import pack_records

__version__ = "b200-0.2.0"
TEST = 42


class QPSolverName(IntEnum):
    OSQP = 0
    QPOASES = 1


OSQP = QPSolverName.OSQP          # py::enum_::export_values() (mpc_osqp.cc:964-967)
QPOASES = QPSolverName.QPOASES

_ARGS = ("qp_weights", "com_position", "com_velocity", "com_roll_pitch_yaw", "ground_normal_vec",
         "com_angular_velocity", "foot_contact_states", "foot_positions_body_frame",
         "foot_friction_coeffs", "desired_com_position", "desired_com_velocity",
         "desired_com_roll_pitch_yaw", "desired_com_angular_velocity")
_warned_f64 = False


def _to_f32(name, v):
    global _warned_f64
    a = np.asarray(v)
    out = np.asarray(a, dtype=np.float32).reshape(-1)
    if not _warned_f64 and a.dtype == np.float64 and not np.array_equal(out.astype(np.float64), a.reshape(-1)):
        _warned_f64 = True
        warnings.warn(f"mpc_osqp.ConvexMpc: argument {name!r} holds float64 values that are not float32-representable; "
                      "the B200 problem record is float32 (the reference's own caller passes float32 / float16)",
                      RuntimeWarning, stacklevel=3)
    return out


class ConvexMpc:
    def __init__(self, mass, inertia, num_legs, planning_horizon, timestep, alpha, qp_solver_name):
        if int(num_legs) != 4:
            raise ValueError("the B200 solver is built for num_legs == 4 (the reference's only use)")
        inertia = np.asarray(list(inertia), dtype=np.float32).reshape(9)
        rp = np.zeros((1, 12), np.float32)
        rp[0, 0] = mass
        rp[0, 1:10] = inertia
        self._h = int(planning_horizon)
        self._solver = QPSolverName(int(qp_solver_name))      # validates the enum value
        kw = dict(dt=float(timestep), alpha=float(alpha))
        if self._solver == QPSolverName.QPOASES:
            kw.update(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000)
        _lib.default_settings(horizon=self._h)                # validates horizon / loads the CUDA library
        if _lib.lib().mpc_b200_device_count() <= 0:
            raise _lib.MpcB200Error("no CUDA device: this library has no CPU path")
        import torch
        from .batched import BatchedConvexMpc
        self._torch = torch
        self._mpc = BatchedConvexMpc(self._h, device="cuda", robot_params=rp, **kw)
        self._warm = self._mpc.warm_buffer(1) if self._solver == QPSolverName.OSQP else None
        self._rec_host = torch.empty((1, self._mpc.stride), dtype=torch.float32).pin_memory()
        self._rec_dev = torch.empty((1, self._mpc.stride), dtype=torch.float32, device=self._mpc.device)

    def compute_contact_forces(self, *args):
        if len(args) != 13:
            raise TypeError(f"compute_contact_forces() takes 13 positional arguments ({len(args)} given)")
        kw = {k: _to_f32(k, v) for k, v in zip(_ARGS, args)}
        table = kw.pop("foot_contact_states")
        self._rec_host.copy_(self._torch.from_numpy(pack_records(self._h, foot_contact_states=table, **kw)))
        self._rec_dev.copy_(self._rec_host, non_blocking=True)
        out = self._mpc.solve(self._rec_dev, full=True, warm=self._warm)
        status = int(out["status"][0])                        # synchronises
        if status != _lib.STATUS_SOLVED:
            return []                                         # mpc_osqp.cc:791-794
        return out["grf_full"][0].tolist()

    def reset_solver(self):
        """mpc_osqp.cc:576: ``initial_run_ = true`` and nothing else - the OSQP branch keys on
        ``workspace_ == 0`` (:757), so the persistent workspace (here: ``_warm``) is NOT dropped."""
        return None
