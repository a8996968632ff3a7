"""Batched FSM shell around the device-resident Locomotion tick.

Batched counterpart of the reference's per-robot wiring
  ``RobotRunnerFSM.run``          reference/MPC_Controller/robot_runner/RobotRunnerFSM.py:55-69
  -> ``ControlFSM.runFSM``        FSM_states/ControlFSM.py:78-104   (TEST / NORMAL / TRANSITIONING)
  -> ``FSM_State_Locomotion``     FSM_states/FSM_State_Locomotion.py:44-101, locomotionSafe :104-136
     ``FSM_State_RecoveyrStand``  FSM_states/FSM_State_RecoveryStand.py  (StandUp / FoldLegs / RollOver)
     ``FSM_State_Passive``        FSM_states/FSM_State_Passive.py
  -> ``LegController.updateCommand``  common/LegController.py:108-132

One robot = one row of a handful of small device tensors (current state, operating mode, next
state, RecoveryStand sub-state and counters, the joint-space interpolation start).  The heavy
state (Locomotion) is the fused tick kernel, gated per robot through ``d_active``
(include/mpc_b200.h MPC_TICK_*): robots that are not in Locomotion are skipped by the kernel,
NORMAL-mode robots get ``locomotionSafe()`` evaluated BEFORE their control step exactly as
``checkTransition`` does (ControlFSM.py:84-95) and are not ticked when unsafe.  Passive and
RecoveryStand are joint-space PD targets (``jointPDControl``: kp 80, kd 1, FSM_State.py:47-66) —
a few element-wise torch ops on the same device; nothing ever visits the host, so the whole
FSM tick is asynchronous like the kernel.

Differences from the reference, all forced by batching: ``Parameters.control_mode`` (a process
global the gamepad writes, RL_MPC_Locomotion.py:91) is a per-robot tensor ``control_mode``;
``Parameters.locomotionUnsafe`` (global flag the main loop turns into a fake gamepad event,
RL_MPC_Locomotion.py:103-105) is the per-robot tensor ``locomotion_unsafe`` and, with
``auto_recover=True``, requests RECOVERY_STAND for that robot.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .locomotion import BatchedLocomotion

# reference utils.py:26-36
PASSIVE, LOCOMOTION, RECOVERY_STAND = 0, 4, 6
MODE_TEST, MODE_NORMAL, MODE_TRANSITIONING = 0, 1, 2
STAND_UP, FOLD_LEGS, ROLL_OVER = 0, 1, 2          # FSM_State_RecoveryStand.py:7-9


class BatchedRunnerFSM:
    def __init__(self, loco: BatchedLocomotion, control_mode: int = RECOVERY_STAND, operating_mode: int = MODE_NORMAL,
                 check_safety: bool = True, auto_recover: bool = True):
        """``loco``: the batched Locomotion tick (owns the controller state).  ``control_mode`` /
        ``operating_mode`` / ``check_safety`` are the reference's ``Parameters.control_mode``,
        ``.operatingMode``, ``.FSM_check_safety`` (Parameters.py:35-43: LOCOMOTION/TEST/False when
        bridging to RL, RECOVERY_STAND/NORMAL/True otherwise)."""
        self.loco = loco
        n, dev = loco.n, loco.device
        self.n, self.device = n, dev
        self.check_safety = bool(check_safety)
        self.auto_recover = bool(auto_recover)
        i32 = lambda v: torch.full((n,), v, dtype=torch.int32, device=dev)
        self.control_mode = i32(control_mode)
        self.cur = i32(control_mode)                  # ControlFSM.initialize (ControlFSM.py:62-76)
        self.nxt = i32(control_mode)
        self.mode = i32(operating_mode)
        self.locomotion_unsafe = torch.zeros(n, dtype=torch.bool, device=dev)
        self.passive_iter = i32(0)
        # RecoveryStand (FSM_State_RecoveryStand.py:12-52); controller_dt = 0.01 -> ramp lengths in ticks
        k = loco.tick_params.dt * 100.0
        self.fold_ramp, self.fold_settle = int(45 / k), int(75 / k)
        self.standup_ramp, self.standup_settle = int(30 / k), int(30 / k)
        self.rollover_ramp, self.rollover_settle = int(13 / k), int(15 / k)
        f32 = lambda v: torch.tensor(v, dtype=torch.float32, device=dev).repeat(4)
        self.fold_jpos = f32([0.0, 1.4, -2.7])
        self.stand_jpos = f32([0.0, 0.8, -1.6])
        self.rolling_jpos = torch.tensor([1.3, 3.1, -2.77, 0.0, 1.6, -2.77] * 2, dtype=torch.float32, device=dev)
        self.passive_qdes = f32([0.0, 0.01, 0.01])    # FSM_State_Passive.py:27
        self.rs_flag = i32(FOLD_LEGS)
        self.rs_state_iter = i32(0)
        self.rs_motion_start = i32(0)
        self.rs_initial = torch.zeros((n, 12), dtype=torch.float32, device=dev)
        self.torques = torch.zeros((n, 12), dtype=torch.float32, device=dev)
        self._gate = i32(0)
        # ControlFSM.initialize -> currentState.onEnter() at construction (ControlFSM.py:62-76): the leg
        # data are still zero and rBody is the zero matrix (not "upside down"), so a robot that starts in
        # RecoveryStand interpolates from q = 0 and stands up iff 0.2 < nominal height < 0.45
        self._last_q = torch.zeros((n, 12), dtype=torch.float32, device=dev)
        self._last_root = torch.zeros((n, 13), dtype=torch.float32, device=dev)
        self._last_root[:, 6] = 1.0
        self.initialize()

    def initialize(self, idx=None):
        """``RobotRunnerFSM.reset`` -> ``ControlFSM.initialize`` (RobotRunnerFSM.py:52-53, ControlFSM.py:62-76)
        for the given robots (all when None): current state = requested control mode, its onEnter()
        with the leg / estimator data of the last tick, NORMAL-or-TEST operating mode kept."""
        m = torch.ones(self.n, dtype=torch.bool, device=self.device)
        if idx is not None:
            m = torch.zeros(self.n, dtype=torch.bool, device=self.device)
            m[torch.as_tensor(idx, device=self.device, dtype=torch.long)] = True
        self.cur = torch.where(m, self.control_mode, self.cur).to(torch.int32)
        self.nxt = torch.where(m, self.control_mode, self.nxt).to(torch.int32)
        self.mode = torch.where(m & (self.mode == MODE_TRANSITIONING), MODE_NORMAL, self.mode).to(torch.int32)
        self.loco.reset_mask(m & (self.cur == LOCOMOTION))
        self._enter_recovery(m & (self.cur == RECOVERY_STAND), self._last_q, self._last_root)

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _upside_down(root_states: torch.Tensor) -> torch.Tensor:
        """``rBody[2,2] < 0`` (FSM_State_RecoveryStand.py:139-142); rBody is the float16 matrix of
        quat_to_rot (orientation_tools.py:135-151): entry [2,2] = 1 - 2 (x^2 + y^2)."""
        x, y = root_states[:, 3], root_states[:, 4]
        return (1.0 - 2.0 * (x * x + y * y)).to(torch.float16) < 0

    def _enter_recovery(self, m: torch.Tensor, q: torch.Tensor, root_states: torch.Tensor):
        """FSM_State_RecoveyrStand.onEnter (FSM_State_RecoveryStand.py:54-74) for the robots in `m`."""
        z = self.loco.state[:, 70]                    # estimator position[2] (MPC_ST_POSZ)
        stand = (~self._upside_down(root_states)) & (z > 0.2) & (z < 0.45)
        flag = torch.where(stand, STAND_UP, FOLD_LEGS).to(torch.int32)
        zero = torch.zeros_like(self.rs_state_iter)
        self.rs_flag = torch.where(m, flag, self.rs_flag)
        self.rs_state_iter = torch.where(m, zero, self.rs_state_iter)
        self.rs_motion_start = torch.where(m, zero, self.rs_motion_start)
        self.rs_initial = torch.where(m[:, None], q, self.rs_initial)

    def _interp(self, cur_iter: torch.Tensor, ramp: int, fin: torch.Tensor) -> torch.Tensor:
        """_SetJPosInterPts (FSM_State_RecoveryStand.py:144-153): joint target of this tick."""
        b = torch.where(cur_iter <= ramp, cur_iter.to(torch.float32) / float(ramp), torch.ones_like(cur_iter, dtype=torch.float32))
        return (1.0 - b)[:, None] * self.rs_initial + b[:, None] * fin[None, :]

    def _run_recovery(self, m: torch.Tensor, q: torch.Tensor, root_states: torch.Tensor) -> torch.Tensor:
        """FSM_State_RecoveyrStand.run (:76-87) for the robots in `m` -> joint targets qDes [N,12]."""
        cur_iter = self.rs_state_iter - self.rs_motion_start
        ud = self._upside_down(root_states)
        flag = self.rs_flag
        su, fl, ro = m & (flag == STAND_UP), m & (flag == FOLD_LEGS), m & (flag == ROLL_OVER)
        # _StandUp (:155-172): body_height is the robot constant (>= 0.26), so only UpsideDown can go wrong
        su_bad = su & (cur_iter > math.floor(self.standup_ramp * 0.7)) & ud
        su_ok = su & ~su_bad
        q_su = self._interp(cur_iter, self.standup_ramp, self.stand_jpos)
        # _FoldLegs (:174-189) ramps with rollover_ramp_iter (sic)
        q_fl = self._interp(cur_iter, self.rollover_ramp, self.fold_jpos)
        fl_done = fl & (cur_iter >= self.fold_ramp + self.fold_settle)
        # _RollOver (:191-201)
        q_ro = self._interp(cur_iter, self.rollover_ramp, self.rolling_jpos)
        ro_done = ro & (cur_iter > self.rollover_ramp + self.rollover_settle)
        qdes = torch.zeros_like(q)
        qdes = torch.where(su_ok[:, None], q_su, qdes)
        qdes = torch.where(fl[:, None], q_fl, qdes)
        qdes = torch.where(ro[:, None], q_ro, qdes)
        has_target = su_ok | fl | ro                  # a failed StandUp tick commands nothing (zero torques)
        # sub-state switches
        new_flag = flag
        new_flag = torch.where(su_bad, FOLD_LEGS, new_flag)
        new_flag = torch.where(fl_done & ud, ROLL_OVER, new_flag)
        new_flag = torch.where(fl_done & ~ud, STAND_UP, new_flag)
        new_flag = torch.where(ro_done, FOLD_LEGS, new_flag)
        switch = su_bad | fl_done | ro_done
        init = self.rs_initial
        init = torch.where(su_bad[:, None], q, init)
        init = torch.where(fl_done[:, None], self.fold_jpos[None, :], init)
        init = torch.where(ro_done[:, None], self.rolling_jpos[None, :], init)
        self.rs_initial = init
        self.rs_flag = new_flag.to(torch.int32)
        self.rs_motion_start = torch.where(switch, self.rs_state_iter + 1, self.rs_motion_start)
        self.rs_state_iter = torch.where(m, self.rs_state_iter + 1, self.rs_state_iter)
        return qdes, has_target

    # ------------------------------------------------------------------ one tick
    def run(self, dof_state: torch.Tensor, root_states: torch.Tensor, commands: torch.Tensor,
            actions: torch.Tensor | None = None, has_weights: bool = False) -> torch.Tensor:
        """``RobotRunnerFSM.run`` for every robot: Isaac tensors in (as BatchedLocomotion.run /
        .pre_physics_step), ``torques`` [N,12] out.  Asynchronous, no host synchronisation."""
        n = self.n
        dof = dof_state.reshape(n, 12, 2)
        q, qd = dof[:, :, 0], dof[:, :, 1]
        self._last_q, self._last_root = q, root_states
        cur, mode, cm = self.cur, self.mode, self.control_mode
        test, normal, trans = mode == MODE_TEST, mode == MODE_NORMAL, mode == MODE_TRANSITIONING
        in_loco, in_pas, in_rs = cur == LOCOMOTION, cur == PASSIVE, cur == RECOVERY_STAND

        # ---- Locomotion: the tick kernel behind the per-robot gate
        want_loco = cm == LOCOMOTION
        chk = self.check_safety
        gate = torch.zeros(n, dtype=torch.int32, device=self.device)
        gate = torch.where(in_loco & test, _lib.TICK_RUN, gate)
        gate = torch.where(in_loco & normal & want_loco, _lib.TICK_CHECK_RUN if chk else _lib.TICK_RUN, gate)
        gate = torch.where(in_loco & normal & ~want_loco, _lib.TICK_CHECK_ONLY if chk else _lib.TICK_SKIP, gate)
        self._gate = gate.to(torch.int32)
        if actions is not None:
            self.loco.pre_physics_step(actions, dof_state, root_states, commands, active=self._gate)
        else:
            self.loco.run(dof_state, root_states, commands, has_weights=has_weights, active=self._gate)
        checked = (self._gate == _lib.TICK_CHECK_RUN) | (self._gate == _lib.TICK_CHECK_ONLY)
        unsafe = checked & self.loco.unsafe           # checkTransition: not locomotionSafe()
        kernel_tau = (self._gate != _lib.TICK_SKIP)   # the kernel wrote this robot's torques (zeros when not ticked)

        # ---- checkTransition (NORMAL mode) -> requested next state per robot
        nxt = self.nxt
        loco_n, pas_n, rs_n = in_loco & normal, in_pas & normal, in_rs & normal
        nxt = torch.where(loco_n & ~unsafe & ~want_loco, cm, nxt)          # FSM_State_Locomotion.py:62-76
        nxt = torch.where(loco_n & unsafe, RECOVERY_STAND, nxt)            # :77-79
        self.passive_iter = torch.where(pas_n, self.passive_iter + 1, self.passive_iter)
        nxt = torch.where(pas_n, torch.where(cm == RECOVERY_STAND, RECOVERY_STAND, PASSIVE), nxt)   # Passive :44-57
        nxt = torch.where(rs_n, cm, nxt)                                   # RecoveryStand :96-110
        nxt = nxt.to(torch.int32)
        start_trans = normal & (nxt != cur)
        self.locomotion_unsafe = self.locomotion_unsafe | unsafe
        if self.auto_recover:
            self.control_mode = torch.where(unsafe, RECOVERY_STAND, cm).to(torch.int32)

        # ---- run() of the joint-space states for robots that stay (NORMAL, no transition) or are in TEST
        runs = test | (normal & ~start_trans)
        pas_run, rs_run = in_pas & runs, in_rs & runs
        q_rs, rs_target = self._run_recovery(rs_run, q, root_states)
        pas_target = pas_run & (self.passive_iter < 10)                    # FSM_State_Passive.py:25-32
        qdes = torch.where(pas_target[:, None], self.passive_qdes[None, :], q_rs)
        pd = pas_target | rs_target
        # LegController.updateCommand with only the joint PD terms set (kpJoint 80, kdJoint 1, qdDes 0)
        tau_pd = torch.where(pd[:, None], 80.0 * (qdes - q) + 1.0 * (0.0 - qd), torch.zeros_like(q))
        self.torques = torch.where(kernel_tau[:, None], self.loco.torques, tau_pd)

        # ---- TRANSITIONING robots of THIS tick's start: transition() is complete at once for every
        # pair the reference allows; onExit, switch, onEnter, back to NORMAL (ControlFSM.py:97-104)
        enter = trans
        ent_loco, ent_rs = enter & (self.nxt == LOCOMOTION), enter & (self.nxt == RECOVERY_STAND)
        self.loco.reset_mask(ent_loco)                                     # FSM_State_Locomotion.onEnter :31-38
        self._enter_recovery(ent_rs, q, root_states)
        self.cur = torch.where(enter, self.nxt, cur).to(torch.int32)
        new_mode = torch.where(enter, MODE_NORMAL, mode)
        new_mode = torch.where(start_trans, MODE_TRANSITIONING, new_mode)
        self.mode = new_mode.to(torch.int32)
        self.nxt = torch.where(enter, self.nxt, nxt).to(torch.int32)
        return self.torques
