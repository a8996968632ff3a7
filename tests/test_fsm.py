"""FSM shell parity on the CPU: BatchedRunnerFSM (the batched RobotRunnerFSM / ControlFSM / Passive /
RecoveryStand / Locomotion wiring) replays the scripted scenario recorded from the UNMODIFIED
reference Python (tests/golden/make_fsm_golden.py) and must visit the same FSM states in the same
operating modes and produce the same joint torques on every tick.  The Locomotion tick behind it is
the warp-emulation build here (tests/emu_loco.py); tests/test_gpu_fsm.py repeats it on the device."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def replay_fsm(make_loco, device="cpu", extra_robots=0):
    from rl_mpc_locomotion_b200 import runner_fsm as F
    g = np.load(os.path.join(GOLD, "fsm_golden_aliengo.npz"))
    n = 1 + extra_robots
    loco = make_loco(np.zeros(n, np.int32), np.zeros(n, np.int32))
    fsm = F.BatchedRunnerFSM(loco)      # reference defaults: RECOVERY_STAND, NORMAL, safety checks on
    script = dict(zip(g["script_ticks"].tolist(), g["script_modes"].tolist()))
    worst, t32 = 0.0, lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(device)
    for t in range(len(g["dof"])):
        if t in script:
            fsm.control_mode[:] = script[t]
        dof = t32(np.repeat(g["dof"][t][None], n, 0))
        body = t32(np.repeat(g["body"][t][None], n, 0))
        cmd = t32(np.repeat(g["cmd"][t][None], n, 0))
        if extra_robots:   # bystanders: robot 1 never leaves RecoveryStand, it must not disturb robot 0
            fsm.control_mode[1:] = F.RECOVERY_STAND
        tau = fsm.run(dof, body, cmd).cpu().numpy()
        assert int(fsm.cur[0]) == int(g["state"][t]), (t, int(fsm.cur[0]), int(g["state"][t]))
        assert int(fsm.mode[0]) == int(g["op_mode"][t]), (t, int(fsm.mode[0]), int(g["op_mode"][t]))
        assert int(fsm.control_mode[0]) == int(g["request"][t]), t
        err = np.abs(tau[0] - g["torques"][t]).max()
        assert err < 2e-3, (t, err, int(g["state"][t]))
        worst = max(worst, err)
        if extra_robots:
            assert int(fsm.cur[1]) == F.RECOVERY_STAND
    assert bool(fsm.locomotion_unsafe[0]) and int(g["unsafe"].sum()) == 2
    return worst


def test_fsm_scenario_matches_reference_python():
    from emu_loco import EmuLocomotion
    w = replay_fsm(lambda rt, gait: EmuLocomotion(rt, gait))
    print("FSM replay: max torque error", w)


def test_fsm_robots_are_independent():
    from emu_loco import EmuLocomotion
    replay_fsm(lambda rt, gait: EmuLocomotion(rt, gait), extra_robots=1)
