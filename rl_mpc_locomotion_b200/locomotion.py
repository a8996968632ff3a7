"""Batched, device-resident Locomotion tick (the new entry point for the RL environment).

Replaces the host loop ``for idx, controller in enumerate(self.controllers): controller.run(...)``
(reference/RL_Environment/tasks/aliengo.py:252-256) — i.e. ``RobotRunnerMin.run``
(robot_runner/RobotRunnerMin.py:54-76) — by ONE kernel launch per control tick:
StateEstimator.update (float16 rotations reproduced bit for bit), leg FK/Jacobian, gait phases and horizon table, CoM height / ground normal,
foot placement, swing Bezier, QP marshalling, the OSQP-faithful QP solve when the MPC is due,
and the Cartesian-PD + J^T f torque map.  All per-robot controller state
(iterationCounter, firstSwing, swing trajectories, f_ff, estimator height / contact phase /
contact history / ground normal) lives in one ``[N, 96]`` device tensor (``state``) so it can be
checkpointed with ``torch.save`` and reset per robot by overwriting rows.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .robots import TICK_STATE_STRIDE, geom_table, initial_tick_state, params_table


class BatchedLocomotion:
    def __init__(self, robot_type, gait, horizon: int = 10, device="cuda", dt: float = 0.01,
                 iters_between_mpc: int = 2, flat_ground: bool = False, warm_start: bool = False,
                 longest_first: bool = False, **settings):
        """``robot_type`` / ``gait``: per-robot int arrays (RobotType values; gait ids as in
        ConvexMPCLocomotion.run :227-241 — the reference can only express ONE global gait,
        Parameters.cmpc_gait)."""
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MpcB200Error("BatchedLocomotion needs a CUDA device: there is no CPU path")
        self.settings = _lib.default_settings(horizon=horizon, **settings)
        self.tick_params = _lib.TickParams(dt, iters_between_mpc, int(flat_ground), 0.0)
        rt = np.atleast_1d(np.asarray(robot_type, dtype=np.int32))
        self.n = len(rt)
        dev = self.device
        self.robot_type = torch.from_numpy(rt).to(dev)
        self.gait = torch.from_numpy(np.broadcast_to(np.asarray(gait, dtype=np.int32), rt.shape).copy()).to(dev)
        self.robot_params = torch.from_numpy(params_table()).to(dev)
        self.robot_geom = torch.from_numpy(geom_table()).to(dev)
        self._fresh = torch.from_numpy(initial_tick_state(rt)).to(dev)
        self.state = self._fresh.clone()
        self.torques = torch.empty((self.n, 12), dtype=torch.float32, device=dev)
        self.grf0 = torch.zeros((self.n, 12), dtype=torch.float32, device=dev)
        self.solved = torch.zeros((self.n,), dtype=torch.int32, device=dev)
        # OSQP workspace state per robot: the reference's OSQP branch keeps one workspace per
        # ConvexMpc object and warm-starts every tick (mpc_osqp.cc:757-778); off by default
        # (cold start = the golden-vector semantics)
        self.warm = torch.zeros((self.n, 2 + 76 * horizon), dtype=torch.float64, device=dev) if warm_start else None
        self.longest_first = bool(longest_first) and self.n > 1
        self.order = None           # int32 [N] launch order for the next tick, or None (index order)
        self._ticks = 0             # the kernel's iterationCounter of robots that were never reset

    @property
    def unsafe(self) -> torch.Tensor:
        """Per-robot ``not locomotionSafe()`` of the last tick (FSM_State_Locomotion.py:104-136):
        the host FSM shell uses it to hand a robot over to RecoveryStand, as the reference does
        (FSM_State_Locomotion.py:62-81).  Evaluated only when ``root_states`` are passed to run()."""
        return self.state.view(torch.int32)[:, 90] != 0

    def reset(self, idx=None):
        """``controller.reset()`` for the given robots (aliengo.py:333-334); all when None.
        ``ConvexMPCLocomotion.initialize`` builds a new ``ConvexMpc`` (ConvexMPCLocomotion.py:102),
        so the OSQP workspace state is dropped too."""
        if idx is None:
            self.state.copy_(self._fresh)
            self.order = None
            if self.warm is not None:
                self.warm.zero_()
        else:
            self.state[idx] = self._fresh[idx]
            if self.warm is not None:
                self.warm[idx] = 0

    def run(self, dof_state: torch.Tensor, root_states: torch.Tensor | None, commands: torch.Tensor,
            has_weights: bool = True, est: torch.Tensor | None = None,
            records_out: torch.Tensor | None = None) -> torch.Tensor:
        """One control tick = ``RobotRunnerMin.run`` for every robot.  ``dof_state`` [N,12,2] (pos,
        vel) and ``root_states`` [N,13] (pos, quat xyzw, lin vel, ang vel) are the Isaac Gym tensors
        of ``pre_physics_step`` (aliengo.py:247-251); ``commands`` [N,16] = vx, vy, yaw rate, 13
        weights.  Alternatively pass ``root_states=None`` and precomputed ``est`` [N,18]
        (StateEstimator.update outputs).  Returns ``torques`` [N,12] (float32, device)."""
        assert (root_states is None) != (est is None), "give exactly one of root_states / est"
        src, shape2 = (root_states, (self.n, 13)) if est is None else (est, (self.n, 18))
        for t, shape in ((dof_state, (self.n, 12, 2)), (src, shape2), (commands, (self.n, 16))):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape, (t.shape, shape)
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = _lib.lib().mpc_b200_tick(C.byref(self.settings), C.byref(self.tick_params), p(dof_state),
                                          p(root_states), p(est), p(commands), int(has_weights), p(self.gait), p(self.robot_params),
                                          p(self.robot_geom), p(self.robot_type), self.n, p(self.state),
                                          p(self.torques), p(self.grf0), p(records_out), p(self.solved),
                                          p(self.warm), p(self.order), C.c_void_p(stream))
        _lib.check(rc, "mpc_b200_tick")
        if self.longest_first and (self._ticks + 1) % self.tick_params.iters_between_mpc == 0:
            # this was (for robots never reset individually) an MPC tick - the kernel solves when the
            # incremented iterationCounter is a multiple of iters_between_mpc: sort for the next one
            iters = self.state.view(torch.int32)[:, 91]          # MPC_ST_ITERS
            self.order = torch.argsort(iters, descending=True, stable=True).to(torch.int32)
        self._ticks += 1
        return self.torques
