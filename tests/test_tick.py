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


def replay(name, tick_fn, device_estimator=False, warm=None):
    g = np.load(os.path.join(GOLD, f"tick_golden_{name}.npz"))
    ref_robot, gait, flat, with_w = [int(v) for v in g["meta"]]
    rtype = {v: k for k, v in _ORDER.items()}[ref_robot]
    state = initial_tick_state([int(rtype)])
    worst = dict(tau=0.0, rec=0.0, fff=0.0, z=0.0)
    for t in range(len(g["dof"])):
        est = np.concatenate([g["vBody"][t], g["omegaBody"][t], g["rpyBody"][t], g["groundR"][t]])[None]
        if device_estimator:   # raw root states in, StateEstimator.update runs inside the kernel
            est = ("body", g["body"][t][None])
        if warm is not None:
            out = tick_fn(state, g["dof"][t][None], est, g["cmd"][t][None], with_w, [gait], [int(rtype)], flat, warm)
        else:
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


def _emu_tick(state, dof, est, cmd, with_w, gait, rtype, flat, warm=None):
    if isinstance(est, tuple):
        return emu.tick(state, dof, None, cmd, with_w, gait, params_table(), geom_table(), rtype, flat_ground=flat,
                        body=est[1], warm=warm)
    return emu.tick(state, dof, est, cmd, with_w, gait, params_table(), geom_table(), rtype, flat_ground=flat,
                    warm=warm)


@pytest.mark.parametrize("device_estimator", [False, True])
@pytest.mark.parametrize("name", ["aliengo_trot", "aliengo_walk_flat", "a1_bound", "go1_trot_flat"])
def test_tick_replay_matches_reference_python(name, device_estimator):
    w = replay(name, _emu_tick, device_estimator)
    print(name, w)
    assert w["rec"] < 5e-6       # QP arguments (float32 marshalling, fp32 estimator arithmetic)
    assert w["z"] < 1e-6
    assert w["fff"] < 5e-3       # MPC forces [N]: the QP inputs differ in the last float32 bits
    assert w["tau"] < 2e-3       # joint torques [N m]


def test_device_estimator_reproduces_float16_outputs_bit_for_bit():
    """StateEstimator.update restated with the reference's float16/float32 mix: rpyBody and
    ground_R_body_frame (float16 in the reference) must be EXACT, body-frame velocities 1e-6."""
    import ctypes as C
    L = emu.lib()
    for name in ("aliengo_trot", "aliengo_walk_flat", "a1_bound"):
        g = np.load(os.path.join(GOLD, f"tick_golden_{name}.npz"))
        T = len(g["body"])
        body = np.ascontiguousarray(g["body"], np.float32)
        normal = np.zeros((T, 3), np.float32)
        normal[0] = [0, 0, 1]
        normal[1:] = g["normal"][:-1]
        est = np.zeros((T, 18), np.float32)
        L.mpc_emu_estimator(C.c_void_p(body.ctypes.data), C.c_void_p(normal.ctypes.data), C.c_void_p(est.ctypes.data), T)
        assert np.array_equal(est[:, 6:9], g["rpyBody"].astype(np.float32))
        assert np.array_equal(est[:, 9:18], g["groundR"].astype(np.float32))
        assert np.abs(est[:, 0:3] - g["vBody"]).max() < 1e-6 and np.abs(est[:, 3:6] - g["omegaBody"]).max() < 1e-6


def test_tick_replay_with_warm_started_osqp_workspace():
    """The reference's OSQP branch keeps ONE workspace per robot and warm-starts every MPC tick
    (mpc_osqp.cc:757-778); fixture recorded with the oracle module in WARM mode."""
    warm = np.zeros((1, 2 + 76 * 10))
    w = replay("aliengo_trot_warm", _emu_tick, device_estimator=True, warm=warm)
    print(w)
    assert w["rec"] < 5e-6 and w["fff"] < 5e-3 and w["tau"] < 2e-3


def test_locomotion_safe_flag_replicates_reference_quirks():
    """locomotionSafe (FSM_State_Locomotion.py:104-136): roll > 40 deg trips, roll < -40 deg does
    NOT (`fabs(bool)`), |pitch| > 40 deg trips, a foot above the hip trips."""
    def quat(roll, pitch):
        cr, sr, cp, sp = np.cos(roll / 2), np.sin(roll / 2), np.cos(pitch / 2), np.sin(pitch / 2)
        return [sr * cp, cr * sp, -sr * sp, cr * cp]      # yaw = 0, xyzw
    stand = np.tile([0.0, 0.8, -1.6], 4)
    up = stand.copy(); up[1:3] = [3.0, -0.2]               # leg 0 folded above the hip
    cases = [(0.0, 0.0, stand, 0), (0.8, 0.0, stand, 1), (-0.8, 0.0, stand, 0), (0.0, 0.8, stand, 1),
             (0.0, -0.8, stand, 1), (0.0, 0.0, up, 1)]
    for roll, pitch, q, expect in cases:
        state = initial_tick_state([0])
        body = np.array([[0, 0, 0.35] + quat(roll, pitch) + [0.3, 0, 0, 0, 0, 0]], np.float32)
        dof = np.stack([q, np.zeros(12)], 1)[None].astype(np.float32)
        cmd = np.zeros((1, 16), np.float32)
        emu.tick(state, dof, None, cmd, 0, [0], params_table(), geom_table(), [0], body=body)
        assert int(state.view(np.int32)[0, 90]) == expect, (roll, pitch, expect)


def test_failed_solve_holds_last_forces():
    """Fallback policy of the tick: a solve that does not reach OSQP_SOLVED is reported in `solved`
    and NOT applied - the previous MPC tick's feed-forward forces stay in the state row."""
    g = np.load(os.path.join(GOLD, "tick_golden_aliengo_trot.npz"))
    state = initial_tick_state([0])
    def est_row(t):
        return np.concatenate([g["vBody"][t], g["omegaBody"][t], g["rpyBody"][t], g["groundR"][t]]).astype(np.float32)[None]
    args = lambda t: (g["dof"][t][None], est_row(t), g["cmd"][t][None], 1, [0], params_table(), geom_table(), [0])
    emu.tick(state, *args(0))                             # iterationCounter 1: no MPC
    out1 = emu.tick(state, *args(1))                      # MPC tick, solved
    assert out1["solved"][0] == 1
    fff1 = state[0, 58:70].copy()
    assert np.abs(fff1).max() > 1.0
    emu.tick(state, *args(2))                             # non-MPC tick
    out3 = emu.tick(state, *args(3), params=emu.default_params(max_iter=25))   # MPC tick, cut off at 25 iterations
    assert out3["solved"][0] == -2                        # OSQP_MAX_ITER_REACHED
    assert np.abs(out3["grf0"]).max() > 0                 # the failed solve's iterate is still reported
    assert np.array_equal(state[0, 58:70], fff1)          # ... but not applied
    emu.tick(state, *args(4))
    out5 = emu.tick(state, *args(5))
    assert out5["solved"][0] == 1 and not np.array_equal(state[0, 58:70], fff1)
