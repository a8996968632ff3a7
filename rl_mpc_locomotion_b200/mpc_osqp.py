"""Drop-in replacement of the reference's native module ``mpc_osqp``.

Same Python surface as reference/MPC_Controller/convex_MPC/mpc_osqp.cc:950-983:
``ConvexMpc(mass, inertia, num_legs, planning_horizon, timestep, alpha, qp_solver_name)``,
``.compute_contact_forces(13 sequences) -> list[float]`` (length ``3*num_legs*horizon``, the
NEGATED QP solution, empty list when the solve did not reach OSQP_SOLVED —
mpc_osqp.cc:786-794), ``.reset_solver()``, enum ``QPSolverName`` with exported values
``OSQP`` / ``QPOASES``, ``__version__``, ``TEST == 42``.

Backed by the sm_100a kernel (batch of one through the C-ABI host entry point).  Both
enum values select the OSQP-faithful device solver: the reference's qpOASES backend has
the same unique minimiser and the north-star forbids dispatching to it.  Put
``rl_mpc_locomotion_b200/dropin`` on ``sys.path`` to make ``import mpc_osqp`` resolve here
(INTEGRATION.md).
"""
from __future__ import annotations

from enum import IntEnum

import numpy as np

from . import _lib, batched
from .records import pack_records

__version__ = "b200-0.1.0"
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


class ConvexMpc:
    def __init__(self, mass, inertia, num_legs, planning_horizon, timestep, alpha, qp_solver_name):
        if int(num_legs) != 4:
            raise ValueError("the B200 solver is built for num_legs == 4 (the reference's only use)")
        inertia = np.asarray(list(inertia), dtype=np.float32).reshape(9)
        self._rp = np.zeros((1, 12), np.float32)
        self._rp[0, 0] = mass
        self._rp[0, 1:10] = inertia
        self._h = int(planning_horizon)
        self._kw = dict(dt=float(timestep), alpha=float(alpha))
        QPSolverName(int(qp_solver_name))  # validates the enum value
        _lib.default_settings(horizon=self._h)  # validates horizon / loads the CUDA library

    def compute_contact_forces(self, *args):
        if len(args) != 13:
            raise TypeError(f"compute_contact_forces() takes 13 positional arguments ({len(args)} given)")
        kw = {k: np.asarray(v, dtype=np.float32).reshape(-1) for k, v in zip(_ARGS, args)}
        table = kw.pop("foot_contact_states")
        rec = pack_records(self._h, foot_contact_states=table, **kw)
        out = batched.solve_host(rec, self._h, robot_params=self._rp, full=True, **self._kw)
        if int(out["status"][0]) != _lib.STATUS_SOLVED:
            return []                                    # mpc_osqp.cc:791-794
        return [float(v) for v in out["grf_full"][0]]

    def reset_solver(self):                               # mpc_osqp.cc:576 (no other effect)
        return None
