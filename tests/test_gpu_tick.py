"""GPU: the fused Locomotion tick kernel replays the golden fixtures recorded from the
unmodified reference Python (same checks as tests/test_tick.py, on the device), and a batched
run with per-robot gaits matches single-robot runs bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpu_tick_fn():
    import torch
    from rl_mpc_locomotion_b200.locomotion import BatchedLocomotion
    cache = {}

    def fn(state, dof, est, cmd, with_w, gait, rtype, flat):
        key = (tuple(gait), tuple(rtype), bool(flat))
        if key not in cache:
            cache[key] = BatchedLocomotion(rtype, gait, device="cuda:0", flat_ground=bool(flat))
        loco = cache[key]
        loco.state.copy_(torch.from_numpy(state))
        rec = torch.zeros((1, 96), dtype=torch.float32, device="cuda:0")
        loco.solved.zero_()
        t32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
        if isinstance(est, tuple):
            kw = dict(root_states=t32(est[1]))
        else:
            kw = dict(root_states=None, est=t32(est))
        tau = loco.run(t32(dof).reshape(1, 12, 2), commands=t32(cmd), has_weights=bool(with_w), records_out=rec, **kw)
        torch.cuda.synchronize()
        state[:] = loco.state.cpu().numpy()
        return dict(torques=tau.cpu().numpy(), grf0=loco.grf0.cpu().numpy(), records=rec.cpu().numpy(),
                    solved=loco.solved.cpu().numpy())
    return fn


@pytest.mark.parametrize("device_estimator", [False, True])
@pytest.mark.parametrize("name", ["aliengo_trot", "aliengo_walk_flat", "a1_bound", "go1_trot_flat"])
def test_gpu_tick_replay_matches_reference_python(name, device_estimator):
    from test_tick import replay
    w = replay(name, _gpu_tick_fn(), device_estimator)
    print(name, w)
    assert w["rec"] < 5e-6 and w["z"] < 1e-6 and w["fff"] < 5e-3 and w["tau"] < 2e-3


def test_batched_tick_equals_single_robot_ticks():
    import torch
    from rl_mpc_locomotion_b200.locomotion import BatchedLocomotion
    rng = np.random.default_rng(0)
    n = 96
    rtype = rng.integers(0, 3, n)
    gait = rng.choice([0, 6, 1], n)
    big = BatchedLocomotion(rtype, gait, device="cuda:0", longest_first=True)
    stand = np.tile([0.0, 0.8, -1.6], 4)
    outs = []
    for t in range(6):
        dof = np.stack([stand + 0.1 * rng.standard_normal((n, 12)), 0.5 * rng.standard_normal((n, 12))], -1).astype(np.float32)
        est = np.zeros((n, 18), np.float32)
        est[:, 0:3] = [0.4, 0.0, 0.0] + 0.05 * rng.standard_normal((n, 3))
        est[:, 6:9] = 0.02 * rng.standard_normal((n, 3))
        est[:, 9:18] = np.eye(3).reshape(9)
        cmd = np.zeros((n, 16), np.float32)
        cmd[:, 0] = 0.5
        tau = big.run(torch.from_numpy(dof).cuda(), None, torch.from_numpy(cmd).cuda(), has_weights=False,
                      est=torch.from_numpy(est).cuda()).cpu().numpy().copy()
        outs.append((dof, est, cmd, tau))
    # the launch-order hint (longest OSQP solve first) was active from the second MPC tick on
    iters = big.state.view(torch.int32)[:, 91].cpu().numpy()          # MPC_ST_ITERS
    assert big.order is not None and sorted(big.order.cpu().tolist()) == list(range(n))
    assert (iters >= 25).all() and (iters <= 4000).all()
    for r in (0, 17, 95):
        one = BatchedLocomotion([rtype[r]], [gait[r]], device="cuda:0")
        for dof, est, cmd, tau in outs:
            t1 = one.run(torch.from_numpy(dof[r:r + 1]).cuda(), None, torch.from_numpy(cmd[r:r + 1]).cuda(),
                         has_weights=False, est=torch.from_numpy(est[r:r + 1]).cuda()).cpu().numpy()
            assert np.array_equal(t1[0], tau[r])
