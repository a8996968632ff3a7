"""GPU tests of the device code paths that the CPU suite can only reach through the warp
emulation (VERDICT r1 "What's weak" 1): the hold-last fallback and locomotionSafe() wiring inside
mpc_tick_kernel, non-finite inputs, robot-type range checks, MPC_MODE_FIXED, the FSM gate, the
on-device action rescale, and StateEstimator.update's float16 outputs BIT FOR BIT on the device
(`__double2half` there, a software rounder in the emulator)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _t32(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()


def test_device_estimator_float16_outputs_bit_exact():
    from rl_mpc_locomotion_b200 import _lib
    for name in ("aliengo_trot", "aliengo_walk_flat", "a1_bound", "go1_trot_flat"):
        g = np.load(os.path.join(GOLD, f"tick_golden_{name}.npz"))
        T = len(g["body"])
        body = np.ascontiguousarray(g["body"], np.float32)
        normal = np.zeros((T, 3), np.float32)
        normal[0] = [0, 0, 1]
        normal[1:] = g["normal"][:-1]
        est = np.zeros((T, 18), np.float32)
        vp = lambda a: C.c_void_p(a.ctypes.data)
        _lib.check(_lib.lib().mpc_b200_debug_estimator(vp(body), vp(normal), vp(est), T), "debug_estimator")
        assert np.array_equal(est[:, 6:9], g["rpyBody"].astype(np.float32)), name          # float16 values, exact
        assert np.array_equal(est[:, 9:18], g["groundR"].astype(np.float32)), name
        assert np.abs(est[:, 0:3] - g["vBody"]).max() < 1e-6 and np.abs(est[:, 3:6] - g["omegaBody"]).max() < 1e-6


def _loco(n=1, **kw):
    from rl_mpc_locomotion_b200.locomotion import BatchedLocomotion
    return BatchedLocomotion(np.zeros(n, np.int32), np.zeros(n, np.int32), device="cuda:0", **kw)


def _golden_inputs(g, t):
    est = np.concatenate([g["vBody"][t], g["omegaBody"][t], g["rpyBody"][t], g["groundR"][t]])[None]
    return _t32(g["dof"][t][None]).reshape(1, 12, 2), _t32(est), _t32(g["cmd"][t][None])


def test_failed_solve_holds_last_forces_on_device():
    """mpc_tick_kernel's fallback (csrc/mpc_b200.cu, `apply`): a solve cut off at max_iter = 25 is
    reported in `solved` (OSQP_MAX_ITER_REACHED) and NOT applied; the previous MPC tick's f_ff stays."""
    import torch
    g = np.load(os.path.join(GOLD, "tick_golden_aliengo_trot.npz"))
    ok, cut = _loco(), _loco(max_iter=25)
    def tick(loco, t):
        d, e, c = _golden_inputs(g, t)
        loco.run(d, None, c, has_weights=True, est=e)
        torch.cuda.synchronize()
    tick(ok, 0); tick(ok, 1)
    assert int(ok.solved[0]) == 1
    fff1 = ok.state[0, 58:70].clone()
    assert float(fff1.abs().max()) > 1.0
    tick(ok, 2)
    cut.state.copy_(ok.state)                                  # same controller, solver cut off at 25 iterations
    tick(cut, 3)
    assert int(cut.solved[0]) == -2
    assert float(cut.grf0.abs().max()) > 0                     # the failed iterate is still reported ...
    assert torch.equal(cut.state[0, 58:70], fff1)              # ... but not applied
    ref = _loco(); ref.state.copy_(ok.state); tick(ref, 3)     # the same tick with the full solve does apply
    assert int(ref.solved[0]) == 1 and not torch.equal(ref.state[0, 58:70], fff1)
    # torques of the cut-off tick = torques computed from the OLD forces
    assert not torch.equal(cut.torques, ref.torques)


def test_locomotion_unsafe_flag_and_fsm_gate_on_device():
    """locomotionSafe() (FSM_State_Locomotion.py:104-136, incl. the fabs(bool) slips) evaluated by the
    tick kernel, and the MPC_TICK_* gate: CHECK_RUN does not tick an unsafe robot (zero torques,
    controller state untouched), CHECK_ONLY never ticks, SKIP touches nothing."""
    import torch
    from rl_mpc_locomotion_b200 import _lib
    def quat(roll, pitch):
        cr, sr, cp, sp = np.cos(roll / 2), np.sin(roll / 2), np.cos(pitch / 2), np.sin(pitch / 2)
        return [sr * cp, cr * sp, -sr * sp, cr * cp]
    stand = np.tile([0.0, 0.8, -1.6], 4)
    up = stand.copy(); up[1:3] = [3.0, -0.2]
    cases = [(0.0, 0.0, stand, 0), (0.8, 0.0, stand, 1), (-0.8, 0.0, stand, 0), (0.0, 0.8, stand, 1),
             (0.0, -0.8, stand, 1), (0.0, 0.0, up, 1)]
    n = len(cases)
    body = _t32([[0, 0, 0.35] + quat(r, p) + [0.3, 0, 0, 0, 0, 0] for r, p, _, _ in cases])
    dof = _t32(np.stack([np.stack([q, np.zeros(12)], 1) for _, _, q, _ in cases]))
    cmd = _t32(np.zeros((n, 3)))
    expect = torch.tensor([e for *_, e in cases], dtype=torch.bool, device="cuda:0")
    loco = _loco(n)
    loco.run(dof, body, cmd, has_weights=False)
    assert torch.equal(loco.unsafe, expect)
    assert (loco.state.view(torch.int32)[:, 0] == 1).all()     # plain RUN ticks everybody
    # CHECK_RUN: safe robots tick, unsafe ones do not
    loco2 = _loco(n)
    gate = torch.full((n,), _lib.TICK_CHECK_RUN, dtype=torch.int32, device="cuda:0")
    loco2.torques.fill_(7.0)
    loco2.run(dof, body, cmd, has_weights=False, active=gate)
    cnt = loco2.state.view(torch.int32)[:, 0]
    assert torch.equal(cnt == 0, expect) and torch.equal(loco2.unsafe, expect)
    assert (loco2.torques[expect] == 0).all() and torch.equal(loco2.torques[~expect], loco.torques[~expect])
    before = loco2.state.clone()
    gate[:] = _lib.TICK_CHECK_ONLY
    loco2.run(dof, body, cmd, has_weights=False, active=gate)
    assert torch.equal(loco2.state, before) and (loco2.torques == 0).all()
    gate[:] = _lib.TICK_SKIP
    loco2.torques.fill_(3.0)
    loco2.run(dof, body, cmd, has_weights=False, active=gate)
    assert torch.equal(loco2.state, before) and (loco2.torques == 3.0).all()


def test_non_finite_and_bad_index_inputs_are_reported():
    import torch
    from rl_mpc_locomotion_b200 import _lib, synth
    from rl_mpc_locomotion_b200.batched import BatchedConvexMpc
    rec, rt, params, _ = synth.make_batch(8, 10, seed=2)
    rec = rec.copy(); rt = rt.copy()
    rec[2, 16] = np.nan                                        # com_velocity.x
    rec[5, 28] = np.inf                                        # a foot position
    rt[6] = 9                                                  # outside the 3-row constant table
    mpc = BatchedConvexMpc(10, device="cuda:0", robot_params=params)
    out = mpc.solve(torch.from_numpy(rec).cuda(), torch.from_numpy(rt).cuda(), full=True)
    st = out["status"].cpu().numpy()
    assert st[2] == _lib.STATUS_NON_FINITE and st[5] == _lib.STATUS_NON_FINITE and st[6] == _lib.STATUS_BAD_INDEX
    assert (np.delete(st, [2, 5, 6]) == 1).all()
    f = out["grf_full"].cpu().numpy()
    assert np.isfinite(f).all() and (f[[2, 5, 6]] == 0).all()
    # a launch-order entry outside the batch is skipped, the others are still solved
    order = torch.tensor([0, 1, 99, 3, 4, 5, 6, 7], dtype=torch.int32, device="cuda:0")
    out2 = mpc.solve(torch.from_numpy(rec).cuda(), torch.from_numpy(rt).cuda(), order=order)
    torch.cuda.synchronize()
    assert torch.equal(out2["grf0"][[0, 1, 3, 4]], out["grf0"][[0, 1, 3, 4]])
    # tick: NaN joint angles -> the solve reports non-finite, the forces are held, nothing crashes
    loco = _loco(2)
    dof = _t32(np.stack([np.stack([np.tile([0.0, 0.8, -1.6], 4), np.zeros(12)], 1)] * 2))
    dof[1, 0, 0] = float("nan")
    body = _t32([[0, 0, 0.35, 0, 0, 0, 1, 0.3, 0, 0, 0, 0, 0]] * 2)
    for _ in range(2):
        loco.run(dof, body, _t32(np.zeros((2, 3))), has_weights=False)
    torch.cuda.synchronize()
    assert int(loco.solved[0]) == 1 and int(loco.solved[1]) == _lib.STATUS_NON_FINITE
    assert (loco.state[1, 58:70] == 0).all()


def test_fixed_iteration_mode_accuracy_contract(oracle_mod):
    """MPC_MODE_FIXED (exactly K ADMM iterations, no termination test / rho adaptation / polish):
    what it is and is not.  On the BASELINE config-2 distribution it is NOT what the reference
    returns and NOT the true minimiser at any practical K - the QP is ill-conditioned (alpha = 1e-5
    regularises a ~60-dim null space).  The numbers asserted here are the ones bench.py prints next
    to every fixed-mode throughput."""
    import torch
    from rl_mpc_locomotion_b200 import _lib, synth
    from rl_mpc_locomotion_b200.batched import BatchedConvexMpc
    rec, rt, params, _ = synth.make_batch(128, 10, seed=1000)
    ref0, _, _ = oracle_mod.solve_records(rec, 10, params, rt, mode=oracle_mod.COLD)
    tru0, _, _ = oracle_mod.solve_records(rec, 10, params, rt, mode=oracle_mod.TRUE)
    d_rec, d_rt = torch.from_numpy(rec).cuda(), torch.from_numpy(rt).cuda()
    med_true, rows = [], []
    for k in (25, 50, 100, 400):
        mpc = BatchedConvexMpc(10, device="cuda:0", robot_params=params, mode=_lib.MODE_FIXED, admm_iters=k, polish=0)
        out = mpc.solve(d_rec, d_rt, info=True)
        f = out["grf0"].cpu().numpy()
        assert (out["info"][:, 0] == k).all() and (out["status"] == 1).all()
        e_ref, e_true = np.abs(f - ref0).max(1), np.abs(f - tru0).max(1)
        rows.append((k, np.median(e_ref), e_ref.max(), np.median(e_true), e_true.max()))
        med_true.append(np.median(e_true))
    for r in rows:
        print("fixed K=%d: |GRF - reference OSQP| median %.3g max %.3g N ; |GRF - true minimiser| median %.3g max %.3g N" % r)
    e_osqp = np.abs(ref0 - tru0).max(1)
    print("reference OSQP vs true minimiser: median %.3g max %.3g N" % (np.median(e_osqp), e_osqp.max()))
    assert med_true[0] > med_true[1] > med_true[2] > med_true[3]       # converges towards the true minimiser ...
    assert rows[1][2] > 1.0 and rows[1][4] > 1.0                        # ... but K = 50 is newtons away from both
    # the emulator runs the same lane program: device and emulation agree
    import emu
    e = emu.solve(rec[:8], 10, params, rt[:8], params=emu.default_params(fixed_iters=50, adaptive_rho=0, polish=0))
    mpc = BatchedConvexMpc(10, device="cuda:0", robot_params=params, mode=_lib.MODE_FIXED, admm_iters=50, polish=0)
    assert np.abs(mpc.solve(d_rec[:8].contiguous(), d_rt[:8].contiguous())["grf0"].cpu().numpy() - e["grf0"]).max() < 1e-3


def test_pre_physics_step_adapter_rescales_actions_on_device():
    """BatchedLocomotion.pre_physics_step(actions[N,12], dof_state[N*12,2], root_states[N,13],
    commands[N,3]) == run() fed with weights = actions * scale + const computed by torch exactly as
    aliengo.py:237-245 does, bit for bit (torques, records, state)."""
    import torch
    from rl_mpc_locomotion_b200 import _lib
    n = 64
    rng = np.random.default_rng(3)
    a, b = _loco(n), _loco(n)
    scale = torch.tensor(_lib.ACT_SCALE, dtype=torch.float, device="cuda:0")
    const = torch.tensor(_lib.ACT_CONST, dtype=torch.float, device="cuda:0")
    stand = np.tile([0.0, 0.8, -1.6], 4)
    ra = torch.zeros((n, 96), dtype=torch.float32, device="cuda:0")
    rb = torch.zeros_like(ra)
    for t in range(4):
        actions = _t32(rng.uniform(-1, 1, (n, 12)))
        dof = _t32(np.stack([stand + 0.1 * rng.standard_normal((n, 12)), 0.3 * rng.standard_normal((n, 12))], -1))
        rpy = 0.05 * rng.standard_normal((n, 3))
        quat = np.stack([rpy[:, 0] / 2, rpy[:, 1] / 2, rpy[:, 2] / 2, np.ones(n)], 1)
        quat /= np.linalg.norm(quat, axis=1, keepdims=True)
        body = _t32(np.concatenate([np.tile([0, 0, 0.35], (n, 1)), quat, [0.5, 0, 0] + 0.1 * rng.standard_normal((n, 3)),
                                    0.2 * rng.standard_normal((n, 3))], 1))
        cmd3 = _t32(rng.uniform(-1, 1, (n, 3)))
        resc = torch.mul(actions, scale).add(const)                       # aliengo.py:237-245
        cmd16 = torch.cat([cmd3, resc, torch.zeros((n, 1), device="cuda:0")], 1).contiguous()   # :253
        ta = a.pre_physics_step(actions, dof.reshape(n * 12, 2), body, cmd3, records_out=ra).clone()
        tb = b.run(dof, body, cmd16, has_weights=True, records_out=rb).clone()
        assert torch.equal(ta, tb) and torch.equal(ra, rb) and torch.equal(a.state, b.state)
    assert (a.solved == 1).all()


def test_reset_keeps_what_the_reference_reset_keeps():
    """controller.reset() (RobotRunnerMin.py:49-52): counters / firstRun / firstSwing / estimator state
    are re-initialised, f_ff and the swing trajectories survive."""
    import torch
    g = np.load(os.path.join(GOLD, "tick_golden_aliengo_trot.npz"))
    loco = _loco()
    for t in range(8):
        d, e, c = _golden_inputs(g, t)
        loco.run(d, None, c, has_weights=True, est=e)
    before = loco.state.clone()
    loco.reset([0])
    sti = loco.state.view(torch.int32)
    assert int(sti[0, 0]) == 0 and int(sti[0, 1]) == 1 and (sti[0, 2:6] == 1).all()
    assert torch.equal(loco.state[0, 6:70], before[0, 6:70])              # swingTimeRemaining, trajectories, f_ff
    assert float(loco.state[0, 70]) == pytest.approx(0.35) and (loco.state[0, 71:75] == 0).all()
    loco.reset(hard=True)
    assert (loco.state[0, 58:70] == 0).all()


def test_qpoases_enum_lands_near_the_true_minimiser(oracle_mod):
    """`mpc_osqp.QPOASES` (what ConvexMPCLocomotion passes): the reference's qpOASES branch returns the
    QP's unique minimiser; the drop-in's tight-tolerance run must be within 1e-2 N of the oracle's
    1e-10 solve, where the default-tolerance OSQP answer is newtons away."""
    from rl_mpc_locomotion_b200 import mpc_osqp as mpc, synth
    from rl_mpc_locomotion_b200.mpc_osqp import _ARGS
    from rl_mpc_locomotion_b200.records import unpack_record
    from rl_mpc_locomotion_b200.robots import ROBOTS, RobotType
    rec, rt, params, _ = synth.make_batch(12, 10, seed=21)
    tru0, tru, _ = oracle_mod.solve_records(rec, 10, params, rt, mode=oracle_mod.TRUE, full=True)
    m = ROBOTS[RobotType.ALIENGO]
    cm = mpc.ConvexMpc(m.mass, list(m.inertia9), 4, 10, 0.02, 1e-5, mpc.QPOASES)
    worst = 0.0
    for r in range(12):
        kw = unpack_record(rec[r], 10)
        f = np.asarray(cm.compute_contact_forces(*[kw[k] for k in _ARGS]))
        assert f.shape == (120,)
        worst = max(worst, np.abs(f - tru[r]).max())
    print("QPOASES enum vs true minimiser:", worst)
    assert worst < 1e-2
