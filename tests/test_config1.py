"""BASELINE config 1: 1 Aliengo, trot, horizon 10, 1000 control ticks of the UNMODIFIED reference
``RobotRunnerMin`` (kinematic replay), recorded by tests/golden/make_tick_golden.py --config1 twice:
with a fresh OSQP workspace per solve (cold) and with the reference's persistent workspace (warm start
every MPC tick, mpc_osqp.cc:757-780).

The replay is open-loop by construction (joint and body states are prescribed functions of time, the
torques do not feed back), so the 13 QP arguments of every MPC tick are independent of the solver:
feeding the recorded argument sequence to the drop-in ``ConvexMpc`` and the recorded Isaac tensors
to the fused tick reproduces exactly what running RobotRunnerMin with the drop-in as its
``mpc_osqp`` would compute.  (/root/reference cannot travel to the GPU box and the drop-in needs a
GPU, so the two cannot meet in one process.)

What can and cannot be asserted over 1000 ticks:
* cold: every solve is a function of its own arguments -> parity on ALL 500 solves / 1000 ticks;
* warm: a sequence of warm-started, loosely converged (eps 1e-3) solves is a chaotic map of its own
  rounding history.  The reference itself decorrelates from a copy whose first input was perturbed by
  ONE float64 ulp after ~250 solves (test_reference_warm_sequence_is_chaotic_beyond_250_solves), so
  no re-implementation can track it for 500 solves.  Our polish is algebraically identical to OSQP's
  but rounds differently (~1e-8 on the scaled iterate), which enters the same amplification ~100
  solves earlier: parity is asserted solve-for-solve over the first 60 solves (1e-4 N), then 150
  solves (1e-2 N); after that only the discrete behaviour (every solve reaches OSQP_SOLVED) is."""
import os

import numpy as np
import pytest

import emu
from rl_mpc_locomotion_b200.robots import geom_table, initial_tick_state, params_table

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tick_golden_config1_aliengo_trot_1000.npz")
SIZES = (13, 3, 3, 3, 3, 3, 40, 12, 4, 3, 3, 3, 3)     # reference argument order (mpc_osqp.cc:578-591)


def replay_config1(tick_fn, mode, ticks=None):
    g = np.load(GOLD)
    state = initial_tick_state([0])
    fff, tau = [], []
    k = 0
    T = len(g["dof"]) if ticks is None else ticks
    for t in range(T):
        out = tick_fn(state, g["dof"][t][None], g["body"][t][None], g["cmd"][t][None])
        assert int(state.view(np.int32)[0, 0]) == int(g["counter"][t])
        assert bool(out["solved"][0] == 1) == bool(g["solved"][t]), t
        if g["solved"][t]:
            fff.append(np.abs(out["grf0"][0] - g[f"qp_out_{mode}"][k][:12]).max())
            k += 1
        tau.append(np.abs(out["torques"][0] - g[f"torques_{mode}"][t]).max())
        assert abs(state[0, 70] - g["z_post"][t]) < 1e-6
        assert np.allclose(state[0, 71:75], g["contact_phase"][t], atol=1e-6)
    return np.asarray(fff), np.asarray(tau)


def _emu_fn(warm):
    def fn(state, dof, body, cmd):
        return emu.tick(state, dof, None, cmd, 1, [0], params_table(), geom_table(), [0], body=body, warm=warm)
    return fn


def test_config1_cold_replay_on_the_emulated_tick():
    fff, tau = replay_config1(_emu_fn(None), "cold", ticks=400)     # the GPU test runs all 1000
    print("cold, 400 ticks (emulated): forces", fff.max(), "torques", tau.max())
    assert fff.max() < 5e-3 and tau.max() < 2e-3


def test_config1_warm_replay_prefix_on_the_emulated_tick():
    fff, tau = replay_config1(_emu_fn(np.zeros((1, 2 + 76 * 10))), "warm", ticks=300)
    print("warm, 150 solves (emulated): forces first 60 solves", fff[:60].max(), "all", fff.max(), "torques", tau.max())
    assert fff[:60].max() < 5e-3 and fff.max() < 2e-2 and tau.max() < 5e-3


def test_reference_warm_sequence_is_chaotic_beyond_250_solves(oracle_mod):
    """The reference OSQP workspace driven through the 500 recorded calls twice; the second time the
    weights of the FIRST call only are scaled by (1 + 1e-15).  Identical to 1e-8 N for ~150 solves,
    newton-level differences after ~250: the horizon beyond which 'same output as the reference' is not
    a property any implementation - including a re-run of the reference on another CPU - can have."""
    g = np.load(GOLD)
    rp = params_table()[0].astype(np.float64)
    def run(perturb):
        o = oracle_mod.OracleMpc(rp[0], rp[1:10], horizon=10)
        outs = []
        for k in range(500):
            kw = dict(zip(oracle_mod.ARG_ORDER, np.split(g["qp_args"][k].astype(np.float64), np.cumsum(SIZES)[:-1])))
            if perturb and k == 0:
                kw["qp_weights"] = kw["qp_weights"] * (1 + 1e-15)
            outs.append(oracle_mod.OracleMpc.solve(o, oracle_mod.WARM, **kw)[1])
        return np.asarray(outs)
    a, b = run(False), run(True)
    # the fixture is this very sequence; its arguments are stored in float32 (the device record's precision),
    # which already is a ~1e-8 relative perturbation of the few float64-valued arguments: same horizon
    assert np.abs(a[:150] - g["qp_out_warm"][:150]).max() < 1e-4
    d = np.abs(a - b).max(axis=1)
    print("reference vs 1-ulp-perturbed reference, max |dF| per block of 50 solves:",
          [float("%.2g" % d[i:i + 50].max()) for i in range(0, 500, 50)])
    assert d[:150].max() < 1e-6 and d[250:].max() > 1e-4


@pytest.mark.gpu
def test_config1_replay_1000_ticks_on_the_device():
    import torch
    from rl_mpc_locomotion_b200.locomotion import BatchedLocomotion
    t32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
    def device_fn(loco):
        def fn(state, dof, body, cmd):
            loco.state.copy_(torch.from_numpy(state))
            loco.solved.zero_()
            tau = loco.run(t32(dof).reshape(1, 12, 2), t32(body), t32(cmd), has_weights=True)
            torch.cuda.synchronize()
            state[:] = loco.state.cpu().numpy()
            return dict(torques=tau.cpu().numpy(), grf0=loco.grf0.cpu().numpy(), solved=loco.solved.cpu().numpy())
        return fn
    fff, tau = replay_config1(device_fn(BatchedLocomotion([0], [0], device="cuda:0")), "cold")
    print(f"config 1, cold, 1000 ticks on the device: forces max {fff.max():.3e} N (99% {np.quantile(fff, .99):.3e}), "
          f"torques max {tau.max():.3e} N m")
    assert fff.max() < 5e-3 and tau.max() < 2e-3
    fff, tau = replay_config1(device_fn(BatchedLocomotion([0], [0], device="cuda:0", warm_start=True)), "warm")
    print(f"config 1, warm: forces first 60 solves {fff[:60].max():.3e}, first 150 {fff[:150].max():.3e}, all 500 {fff.max():.3e} N")
    assert fff[:60].max() < 5e-3 and fff[:150].max() < 2e-2


@pytest.mark.gpu
def test_config1_dropin_convexmpc_replays_the_500_mpc_calls():
    """The drop-in class with the reference's call pattern: ONE ConvexMpc object, 500 consecutive
    compute_contact_forces calls (float32 argument arrays as ConvexMPCLocomotion passes them).  With
    the OSQP enum the object carries the warm workspace state (prefix parity, see the module docstring);
    a second pass resets the object's workspace before every call = the cold fixture, all 500 calls."""
    from rl_mpc_locomotion_b200 import mpc_osqp as mpc
    from rl_mpc_locomotion_b200.robots import ROBOTS, RobotType
    g = np.load(GOLD)
    m = ROBOTS[RobotType.ALIENGO]
    cm = mpc.ConvexMpc(m.mass, list(m.inertia9), 4, 10, 0.02, 1e-5, mpc.OSQP)
    args = [np.split(g["qp_args"][k], np.cumsum(SIZES)[:-1]) for k in range(len(g["qp_args"]))]
    errw = []
    for k, parts in enumerate(args):
        f = cm.compute_contact_forces(*parts)
        assert isinstance(f, list) and len(f) == 120
        errw.append(np.abs(np.asarray(f) - g["qp_out_warm"][k]).max())
    errw = np.asarray(errw)
    print(f"drop-in, 500 warm calls: first 60 {errw[:60].max():.3e}, first 150 {errw[:150].max():.3e}, all {errw.max():.3e} N")
    assert errw[:60].max() < 1e-3 and errw[:150].max() < 2e-2
    errc = []
    for k, parts in enumerate(args):
        cm._warm.zero_()                                     # fresh workspace = osqp_setup on this call
        errc.append(np.abs(np.asarray(cm.compute_contact_forces(*parts)) - g["qp_out_cold"][k]).max())
    errc = np.asarray(errc)
    print(f"drop-in, 500 cold calls: horizon max {errc.max():.3e} N, 99% {np.quantile(errc, .99):.3e} N")
    assert np.quantile(errc, 0.99) < 1e-3 and errc.max() < 5e-3
    cm.reset_solver()                                        # mpc_osqp.cc:576: no effect on the OSQP workspace
