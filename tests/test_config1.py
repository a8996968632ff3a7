"""BASELINE config 1: 1 Aliengo, trot, horizon 10, 1000 control ticks of the UNMODIFIED reference
``RobotRunnerMin`` (kinematic replay, the reference's persistent OSQP workspace = warm start every
MPC tick), recorded by tests/golden/make_tick_golden.py --warm --config1.

The replay is open-loop by construction (joint and body states are prescribed functions of time, the
torques do not feed back), so the 13 QP arguments of every MPC tick are independent of the solver:
feeding the recorded argument sequence to the drop-in ``ConvexMpc`` and the recorded Isaac tensors
to the fused tick reproduces exactly what running RobotRunnerMin with the drop-in as its
``mpc_osqp`` would compute.  (/root/reference cannot travel to the GPU box and the drop-in needs a
GPU, so the two cannot meet in one process.)"""
import os

import numpy as np
import pytest

import emu
from rl_mpc_locomotion_b200.robots import geom_table, initial_tick_state, params_table

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tick_golden_config1_aliengo_trot_warm_1000.npz")


def replay_config1(tick_fn, ticks=None):
    g = np.load(GOLD)
    state = initial_tick_state([0])
    worst = dict(tau=0.0, fff=0.0, z=0.0)
    k = 0
    T = len(g["dof"]) if ticks is None else ticks
    for t in range(T):
        out = tick_fn(state, g["dof"][t][None], g["body"][t][None], g["cmd"][t][None])
        assert int(state.view(np.int32)[0, 0]) == int(g["counter"][t])
        assert bool(out["solved"][0] == 1) == bool(g["solved"][t]), t
        if g["solved"][t]:
            worst["fff"] = max(worst["fff"], np.abs(out["grf0"][0] - g["qp_out"][k][:12]).max())
            k += 1
        worst["tau"] = max(worst["tau"], np.abs(out["torques"][0] - g["torques"][t]).max())
        worst["z"] = max(worst["z"], abs(state[0, 70] - g["z_post"][t]))
        assert np.allclose(state[0, 71:75], g["contact_phase"][t], atol=1e-6)
    return worst


def test_config1_replay_on_the_emulated_tick():
    warm = np.zeros((1, 2 + 76 * 10))
    def fn(state, dof, body, cmd):
        return emu.tick(state, dof, None, cmd, 1, [0], params_table(), geom_table(), [0], body=body, warm=warm)
    w = replay_config1(fn, ticks=300)      # the first 300 ticks on the CPU (the GPU test runs all 1000)
    print(w)
    assert w["z"] < 1e-6 and w["fff"] < 5e-3 and w["tau"] < 2e-3


@pytest.mark.gpu
def test_config1_replay_1000_ticks_on_the_device():
    import torch
    from rl_mpc_locomotion_b200.locomotion import BatchedLocomotion
    loco = BatchedLocomotion([0], [0], device="cuda:0", warm_start=True)
    t32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
    def fn(state, dof, body, cmd):
        loco.state.copy_(torch.from_numpy(state))
        loco.solved.zero_()
        tau = loco.run(t32(dof).reshape(1, 12, 2), t32(body), t32(cmd), has_weights=True)
        torch.cuda.synchronize()
        state[:] = loco.state.cpu().numpy()
        return dict(torques=tau.cpu().numpy(), grf0=loco.grf0.cpu().numpy(), solved=loco.solved.cpu().numpy())
    w = replay_config1(fn)
    print("config 1, 1000 ticks on the device:", w)
    assert w["z"] < 1e-6 and w["fff"] < 5e-3 and w["tau"] < 2e-3


@pytest.mark.gpu
def test_config1_dropin_convexmpc_replays_the_500_mpc_calls():
    """The drop-in class with the reference's call pattern: ONE ConvexMpc object, 500 consecutive
    compute_contact_forces calls (float32 argument arrays as ConvexMPCLocomotion passes them), warm
    workspace carried inside the object.  Every returned 120-vector vs what the reference OSQP
    workspace returned on the same call."""
    from rl_mpc_locomotion_b200 import mpc_osqp as mpc
    from rl_mpc_locomotion_b200.robots import ROBOTS, RobotType
    g = np.load(GOLD)
    m = ROBOTS[RobotType.ALIENGO]
    cm = mpc.ConvexMpc(m.mass, list(m.inertia9), 4, 10, 0.02, 1e-5, mpc.OSQP)
    sizes = (13, 3, 3, 3, 3, 3, 40, 12, 4, 3, 3, 3, 3)     # reference argument order (mpc_osqp.cc:578-591)
    err0, errf = 0.0, []
    for k in range(len(g["qp_args"])):
        parts = np.split(g["qp_args"][k], np.cumsum(sizes)[:-1])
        f = cm.compute_contact_forces(*parts)
        assert isinstance(f, list) and len(f) == 120
        d = np.abs(np.asarray(f) - g["qp_out"][k])
        err0, _ = max(err0, d[:12].max()), errf.append(d.max())
    errf = np.asarray(errf)
    print(f"drop-in vs reference, 500 warm calls: horizon-0 max {err0:.3e} N, full horizon max {errf.max():.3e} N, "
          f"99% {np.quantile(errf, 0.99):.3e} N")
    assert err0 < 1e-3 and np.quantile(errf, 0.99) < 1e-3 and errf.max() < 5e-3
    # reset_solver() == initial_run_: the next call is a cold solve again = a fresh object's first call
    cm.reset_solver()
    parts = np.split(g["qp_args"][7], np.cumsum(sizes)[:-1])
    a = cm.compute_contact_forces(*parts)
    b = mpc.ConvexMpc(m.mass, list(m.inertia9), 4, 10, 0.02, 1e-5, mpc.OSQP).compute_contact_forces(*parts)
    assert a == b
