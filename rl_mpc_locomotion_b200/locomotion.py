"""Batched, device-resident Locomotion tick (the new entry point for the RL environment).

Replaces the host loop ``for idx, controller in enumerate(self.controllers): controller.run(...)``
(reference/RL_Environment/tasks/aliengo.py:252-256) — i.e. ``RobotRunnerMin.run``
(robot_runner/RobotRunnerMin.py:54-76) — by ONE kernel launch per control tick:
StateEstimator.update (float16 rotations reproduced bit for bit), leg FK/Jacobian, gait phases and
horizon table, CoM height / ground normal, foot placement, swing Bezier, QP marshalling, the
OSQP-faithful QP solve when the MPC is due, and the Cartesian-PD + J^T f torque map.  All per-robot
controller state (iterationCounter, firstSwing, swing trajectories, f_ff, estimator height /
contact phase / contact history / ground normal) lives in one ``[N, 96]`` device tensor
(``state``) so it can be checkpointed with ``torch.save`` and reset per robot by overwriting rows.

:meth:`BatchedLocomotion.pre_physics_step` is the adapter at the reference's own boundary
(aliengo.py:227-263): it takes the Isaac Gym tensors exactly as ``pre_physics_step`` holds them —
``actions[N,12]``, ``dof_state[N*12,2]``, ``root_states[N,13]``, ``commands[N,3]`` — and does the
``actions * scale + const`` rescale (aliengo.py:237-245) and the ``commands ++ weights ++ [0]``
assembly (:253) inside the kernel.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, shard
from .robots import TICK_STATE_STRIDE, geom_table, initial_tick_state, params_table

# MPC_ST_* slots (include/mpc_b200.h) the reference's reset path really re-initialises:
# ConvexMPCLocomotion.initialize (ConvexMPCLocomotion.py:89-114: iterationCounter, firstSwing,
# firstRun, a new ConvexMpc) + StateEstimator.reset (StateEstimator.py:41-48: contact phase,
# contact history -> re-seeded by firstRun, height, ground normal).  f_ff, the swing trajectories
# and swingTimeRemaining survive a reference reset.
_ST_RESET_COLS = [0, 1, 2, 3, 4, 5, 70, 71, 72, 73, 74] + list(range(75, 90)) + [90, 91]


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
        self.tick_params = _lib.tick_params(dt, iters_between_mpc, flat_ground)
        rt = np.atleast_1d(np.asarray(robot_type, dtype=np.int32))
        self.n = len(rt)
        dev = self.device
        self.robot_type = torch.from_numpy(rt).to(dev)
        self.gait = torch.from_numpy(np.broadcast_to(np.asarray(gait, dtype=np.int32), rt.shape).copy()).to(dev)
        self.robot_params = torch.from_numpy(params_table()).to(dev)
        self.robot_geom = torch.from_numpy(geom_table()).to(dev)
        self._fresh = torch.from_numpy(initial_tick_state(rt)).to(dev)
        self._reset_cols = torch.tensor(_ST_RESET_COLS, dtype=torch.long, device=dev)
        self.state = self._fresh.clone()
        self.torques = torch.zeros((self.n, 12), dtype=torch.float32, device=dev)
        self.grf0 = torch.zeros((self.n, 12), dtype=torch.float32, device=dev)
        self.solved = torch.zeros((self.n,), dtype=torch.int32, device=dev)
        # OSQP workspace state per robot: the reference's OSQP branch keeps one workspace per
        # ConvexMpc object and warm-starts every tick (mpc_osqp.cc:757-778); off by default
        # (cold start = the golden-vector semantics)
        self.warm = torch.zeros((self.n, 2 + 76 * horizon), dtype=torch.float64, device=dev) if warm_start else None
        self.longest_first = bool(longest_first) and self.n > 1
        self.order = None           # int32 [N] launch order for the next tick, or None (index order)

    @property
    def unsafe(self) -> torch.Tensor:
        """Per-robot ``not locomotionSafe()`` of the last tick (FSM_State_Locomotion.py:104-136):
        the FSM shell (runner_fsm.BatchedRunnerFSM) uses it to hand a robot over to RecoveryStand, as
        the reference does (FSM_State_Locomotion.py:62-81).  Evaluated only when ``root_states`` are
        passed to run()."""
        return self.state.view(torch.int32)[:, 90] != 0

    def reset(self, idx=None, hard: bool = False):
        """``controller.reset()`` for the given robots (aliengo.py:333-334); all when None.

        Reference semantics (default): only what ``ConvexMPCLocomotion.initialize`` +
        ``StateEstimator.reset`` re-initialise is reset — iterationCounter, firstRun, firstSwing,
        estimator height / contact phase / contact history / ground normal, and the OSQP workspace
        (``initialize`` builds a new ``ConvexMpc``, ConvexMPCLocomotion.py:102).  The last MPC
        forces ``f_ff``, the swing trajectories and ``swingTimeRemaining`` SURVIVE, exactly as in
        the reference, so the first ticks after a reset produce the reference's torques.
        ``hard=True`` overwrites the whole state row with a freshly constructed controller instead."""
        if idx is None:
            if hard:
                self.state.copy_(self._fresh)
            else:
                self.state[:, self._reset_cols] = self._fresh[:, self._reset_cols]
            self.order = None
            if self.warm is not None:
                self.warm.zero_()
        else:
            idx = torch.as_tensor(idx, device=self.device, dtype=torch.long).reshape(-1)
            if hard:
                self.state[idx] = self._fresh[idx]
            else:
                self.state[idx[:, None], self._reset_cols[None, :]] = self._fresh[idx[:, None], self._reset_cols[None, :]]
            if self.warm is not None:
                self.warm[idx] = 0

    def reset_mask(self, mask: torch.Tensor, hard: bool = False):
        """:meth:`reset` for the robots where ``mask`` [N] (bool, device) is set, without a host
        synchronisation (the FSM shell's Locomotion.onEnter, FSM_State_Locomotion.py:31-38)."""
        if hard:
            self.state = torch.where(mask[:, None], self._fresh, self.state)
        else:
            cols = self._reset_cols
            self.state[:, cols] = torch.where(mask[:, None], self._fresh[:, cols], self.state[:, cols])
        if self.warm is not None:
            self.warm[:, 0] = torch.where(mask, torch.zeros_like(self.warm[:, 0]), self.warm[:, 0])

    # ------------------------------------------------------------------------------------------
    def _launch(self, dof_state, root_states, est, commands, has_weights, actions, records_out, active):
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        self.tick_params.cmd_stride = int(commands.shape[1])
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = _lib.lib().mpc_b200_tick(C.byref(self.settings), C.byref(self.tick_params), p(dof_state),
                                          p(root_states), p(est), p(commands), int(has_weights), p(actions),
                                          p(self.gait), p(self.robot_params), p(self.robot_geom), p(self.robot_type),
                                          int(self.robot_params.shape[0]), self.n, p(active), p(self.state),
                                          p(self.torques), p(self.grf0), p(records_out), p(self.solved),
                                          p(self.warm), p(self.order), C.c_void_p(stream))
        _lib.check(rc, "mpc_b200_tick")
        if self.longest_first:
            # robots whose MPC fired this tick (d_solved != 0) carry a fresh iteration count in
            # MPC_ST_ITERS: sort for the next launch.  Decided on the device - no host counter that a
            # per-robot reset could desynchronise, and no host sync.
            iters = self.state.view(torch.int32)[:, 91]
            self.order = torch.argsort(iters, descending=True, stable=True).to(torch.int32)
        return self.torques

    def run(self, dof_state: torch.Tensor, root_states: torch.Tensor | None, commands: torch.Tensor,
            has_weights: bool = True, est: torch.Tensor | None = None,
            records_out: torch.Tensor | None = None, active: torch.Tensor | None = None) -> torch.Tensor:
        """One control tick = ``RobotRunnerMin.run`` for every robot.  ``dof_state`` [N,12,2] (pos,
        vel) or the flat Isaac view [N*12,2], ``root_states`` [N,13] (pos, quat xyzw, lin vel, ang
        vel) are the Isaac Gym tensors of ``pre_physics_step`` (aliengo.py:247-251); ``commands``
        [N,16] = vx, vy, yaw rate, 13 weights (or [N,3] with ``has_weights=False``).  Alternatively
        pass ``root_states=None`` and precomputed ``est`` [N,18] (StateEstimator.update outputs).
        ``active`` [N] int32 = the FSM shell's MPC_TICK_* gate.  Returns ``torques`` [N,12]."""
        assert (root_states is None) != (est is None), "give exactly one of root_states / est"
        dof_state = self._dof_view(dof_state)
        src, shape2 = (root_states, (self.n, 13)) if est is None else (est, (self.n, 18))
        assert commands.dim() == 2 and commands.shape[1] in (3, 16), commands.shape
        assert not (has_weights and commands.shape[1] != 16), "has_weights needs commands[N,16]"
        for t, shape in ((src, shape2), (commands, (self.n, commands.shape[1]))):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape, (t.shape, shape)
        if active is not None:
            assert active.is_cuda and active.dtype == torch.int32 and tuple(active.shape) == (self.n,)
        return self._launch(dof_state, root_states, est, commands, has_weights, None, records_out, active)

    def pre_physics_step(self, actions: torch.Tensor, dof_state: torch.Tensor, root_states: torch.Tensor,
                         commands: torch.Tensor, records_out: torch.Tensor | None = None,
                         active: torch.Tensor | None = None) -> torch.Tensor:
        """The MPC branch of ``Aliengo.pre_physics_step`` (aliengo.py:227-263) for the whole batch,
        on the device: ``actions`` [N,12] in [-1,1] (the policy output), ``dof_state`` [N*12,2] (or
        [N,12,2]), ``root_states`` [N,13], ``commands`` [N,3] — the tensors the reference copies to
        the host, loops over and copies back.  MPC weights = ``actions * MPC_param_scale +
        MPC_param_const`` with a 13th weight of 0 (aliengo.py:237-245, :253), computed in the kernel
        with the same two float32 roundings as ``torch.mul(...).add(...)``.  Returns ``torques``
        [N,12] float32 on the device (what ``set_dof_actuation_force_tensor`` consumes, :262)."""
        dof_state = self._dof_view(dof_state)
        for t, shape in ((actions, (self.n, 12)), (root_states, (self.n, 13)), (commands, (self.n, 3))):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape, (t.shape, shape)
        if active is not None:
            assert active.is_cuda and active.dtype == torch.int32 and tuple(active.shape) == (self.n,)
        return self._launch(dof_state, root_states, None, commands, False, actions, records_out, active)

    def _dof_view(self, dof_state: torch.Tensor) -> torch.Tensor:
        assert dof_state.is_cuda and dof_state.dtype == torch.float32 and dof_state.is_contiguous()
        assert tuple(dof_state.shape) in ((self.n, 12, 2), (self.n * 12, 2)), dof_state.shape
        return dof_state

    def summary(self, reduce: bool = True) -> torch.Tensor:
        """[n_solved, n_not_solved, 0, 0] of the last tick's MPC solves over ALL ranks: the one
        collective the sharded batch may need, one NCCL all-reduce of 4 doubles when a
        process group is initialised.  Robots whose MPC did not fire (d_solved == 0) are not counted."""
        fired = self.solved[self.solved != 0]
        local = shard.summarize(fired)
        return shard.reduce_summary(local) if reduce else local
