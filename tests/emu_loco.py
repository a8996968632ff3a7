"""CPU stand-in for rl_mpc_locomotion_b200.locomotion.BatchedLocomotion — TEST INFRASTRUCTURE.

Same attributes / methods the FSM shell (runner_fsm.BatchedRunnerFSM) uses, backed by the
warp-emulation build of the tick kernel (tests/emu.py) and torch CPU tensors, so that the FSM
logic can be checked against the reference fixture without a GPU.  tests/test_gpu_fsm.py runs the
same scenario through the real CUDA path.
"""
import numpy as np
import torch

import emu
from rl_mpc_locomotion_b200 import _lib
from rl_mpc_locomotion_b200.locomotion import _ST_RESET_COLS
from rl_mpc_locomotion_b200.robots import geom_table, initial_tick_state, params_table


class _TP:
    dt = 0.01


class EmuLocomotion:
    def __init__(self, robot_type, gait, flat_ground=False):
        rt = np.atleast_1d(np.asarray(robot_type, dtype=np.int32))
        self.n, self.device = len(rt), torch.device("cpu")
        self.tick_params = _TP()
        self.rt, self.gait, self.flat = rt, np.broadcast_to(np.asarray(gait, np.int32), rt.shape).copy(), flat_ground
        self._fresh = torch.from_numpy(initial_tick_state(rt))
        self.state = self._fresh.clone()
        self.torques = torch.zeros((self.n, 12), dtype=torch.float32)
        self.solved = torch.zeros(self.n, dtype=torch.int32)
        self._cols = torch.tensor(_ST_RESET_COLS, dtype=torch.long)

    @property
    def unsafe(self):
        return self.state.view(torch.int32)[:, 90] != 0

    def reset_mask(self, mask, hard=False):
        self.state[:, self._cols] = torch.where(mask[:, None], self._fresh[:, self._cols], self.state[:, self._cols])

    def _tick(self, dof, root, cmd, has_weights, actions, active):
        st = self.state.numpy()
        out = emu.tick(st, dof.reshape(self.n, 24).numpy(), None, cmd.numpy(), int(has_weights), self.gait, params_table(),
                       geom_table(), self.rt, flat_ground=self.flat, body=root.numpy(),
                       actions=None if actions is None else actions.numpy(),
                       active=None if active is None else active.numpy())
        gate = np.ones(self.n, np.int32) if active is None else active.numpy()
        wrote = torch.from_numpy(gate != _lib.TICK_SKIP)
        self.torques = torch.where(wrote[:, None], torch.from_numpy(out["torques"]), self.torques)
        self.solved = torch.from_numpy(out["solved"])
        return self.torques

    def run(self, dof_state, root_states, commands, has_weights=True, est=None, records_out=None, active=None):
        return self._tick(dof_state, root_states, commands, has_weights, None, active)

    def pre_physics_step(self, actions, dof_state, root_states, commands, records_out=None, active=None):
        return self._tick(dof_state, root_states, commands, False, actions, active)
