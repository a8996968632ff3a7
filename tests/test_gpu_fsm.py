"""GPU: the FSM shell over the real CUDA tick (same scenario as tests/test_fsm.py), incl. a batch in
which the scripted robot tips over while its neighbours keep trotting."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make(rt, gait):
    from rl_mpc_locomotion_b200.locomotion import BatchedLocomotion
    return BatchedLocomotion(rt, gait, device="cuda:0")


def test_gpu_fsm_scenario_matches_reference_python():
    from test_fsm import replay_fsm
    w = replay_fsm(_make, device="cuda:0")
    print("FSM replay on device: max torque error", w)


def test_gpu_fsm_robots_are_independent():
    from test_fsm import replay_fsm
    replay_fsm(_make, device="cuda:0", extra_robots=1)


def test_gpu_fsm_tipped_robot_leaves_the_batch_and_returns():
    """RL-bridge style batch (control mode LOCOMOTION, NORMAL operating mode, safety checks on): robot 3
    rolls past 40 deg for a while -> RecoveryStand hand-off, zero torques on the transition ticks, back
    to Locomotion when requested; the other robots' torques are bit-identical to a run without it."""
    import torch
    from rl_mpc_locomotion_b200 import runner_fsm as F
    n, T = 8, 40
    rng = np.random.default_rng(0)
    stand = np.tile([0.0, 0.8, -1.6], 4)
    def inputs(t, tip):
        r = np.random.default_rng(100 + t)
        dof = np.stack([stand + 0.05 * r.standard_normal((n, 12)), 0.2 * r.standard_normal((n, 12))], -1)
        roll = 0.03 * r.standard_normal(n)
        if tip and 10 <= t < 20:
            roll[3] = 0.9
        quat = np.stack([np.sin(roll / 2), np.zeros(n), np.zeros(n), np.cos(roll / 2)], 1)
        body = np.concatenate([np.tile([0, 0, 0.35], (n, 1)), quat, np.tile([0.4, 0, 0], (n, 1)), np.zeros((n, 3))], 1)
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
        return f(dof), f(body), f(np.tile([0.5, 0.0, 0.1], (n, 1)))
    def run(tip):
        fsm = F.BatchedRunnerFSM(_make(np.zeros(n, np.int32), np.zeros(n, np.int32)), control_mode=F.LOCOMOTION)
        taus, states = [], []
        for t in range(T):
            if t == 24:
                fsm.control_mode[:] = F.LOCOMOTION          # the operator asks robot 3 back into Locomotion
            d, b, c = inputs(t, tip)
            taus.append(fsm.run(d, b, c).clone())
            states.append(fsm.cur.clone())
        return torch.stack(taus).cpu().numpy(), torch.stack(states).cpu().numpy(), fsm
    tau0, st0, _ = run(False)
    tau1, st1, fsm = run(True)
    others = [0, 1, 2, 4, 5, 6, 7]
    assert np.array_equal(tau0[:, others], tau1[:, others]) and (st0 == F.LOCOMOTION).all()
    assert (st1[:10, 3] == F.LOCOMOTION).all() and st1[10, 3] == F.LOCOMOTION      # tick 10: transition starts
    assert (tau1[10, 3] == 0).all() and (tau1[11, 3] == 0).all()                    # not ticked / transition tick
    assert (st1[11:25, 3] == F.RECOVERY_STAND).all()
    assert st1[-1, 3] == F.LOCOMOTION and bool(fsm.locomotion_unsafe[3]) and not bool(fsm.locomotion_unsafe[0])
    assert np.abs(tau1[12:24, 3]).max() > 0                                          # RecoveryStand joint PD
