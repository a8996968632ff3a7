"""Locomotion tick parity on the CPU: the device tick (warp-emulation build of mpc_tick.cuh +
mpc_warp_solver.cuh) replays the golden fixtures recorded from the UNMODIFIED reference Python
(tests/golden/make_tick_golden.py) tick by tick and must reproduce the QP arguments, the MPC
forces and the joint torques.  tests/test_gpu_tick.py repeats it on the device."""
import os

import numpy as np
import pytest

import emu
from rl_mpc_locomotion_b200.robots import RobotType, geom_table, initial_tick_state, params_table

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_ORDER = {RobotType.ALIENGO: 1, RobotType.A1: 2, RobotType.GO1: 3}   # reference enum auto() values


def replay(name, tick_fn):
    g = np.load(os.path.join(GOLD, f"tick_golden_{name}.npz"))
    ref_robot, gait, flat, with_w = [int(v) for v in g["meta"]]
    rtype = {v: k for k, v in _ORDER.items()}[ref_robot]
    state = initial_tick_state([int(rtype)])
    worst = dict(tau=0.0, rec=0.0, fff=0.0, z=0.0)
    for t in range(len(g["dof"])):
        est = np.concatenate([g["vBody"][t], g["omegaBody"][t], g["rpyBody"][t], g["groundR"][t]])[None]
        out = tick_fn(state, g["dof"][t][None], est, g["cmd"][t][None], with_w, [gait], [int(rtype)], flat)
        assert int(state.view(np.int32)[0, 0]) == int(g["counter"][t])
        assert bool(out["solved"][0] == 1) == bool(g["solved"][t])
        if g["solved"][t]:
            a = g["qp_args"][t]   # reference argument order: ..., contact table (40) BEFORE foot positions
            as_record = np.concatenate([a[:28], a[68:96], a[28:68]])
            worst["rec"] = max(worst["rec"], np.abs(out["records"][0, :96] - as_record).max())
            worst["fff"] = max(worst["fff"], np.abs(out["grf0"][0] - g["qp_out"][t][:12]).max())
        worst["tau"] = max(worst["tau"], np.abs(out["torques"][0] - g["torques"][t]).max())
        worst["z"] = max(worst["z"], abs(state[0, 70] - g["z_post"][t]))
        assert np.allclose(state[0, 71:75], g["contact_phase"][t], atol=1e-6)
    return worst


def _emu_tick(state, dof, est, cmd, with_w, gait, rtype, flat):
    return emu.tick(state, dof, est, cmd, with_w, gait, params_table(), geom_table(), rtype, flat_ground=flat)


@pytest.mark.parametrize("name", ["aliengo_trot", "aliengo_walk_flat", "a1_bound"])
def test_tick_replay_matches_reference_python(name):
    w = replay(name, _emu_tick)
    print(name, w)
    assert w["rec"] < 5e-6       # QP arguments (float32 marshalling, fp32 estimator arithmetic)
    assert w["z"] < 1e-6
    assert w["fff"] < 5e-3       # MPC forces [N]: the QP inputs differ in the last float32 bits
    assert w["tau"] < 2e-3       # joint torques [N m]
