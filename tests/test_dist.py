"""world_size-2 gloo test of the multi-GPU host logic on CPU: shard ranges partition the batch,
rank-sharded solves equal the single-batch solve bit for bit (the warp-emulation build stands in
for the GPU kernel here), and the one summary collective reduces correctly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emu
    from rl_mpc_locomotion_b200.shard import reduce_summary, shard_range, summarize
    g = np.load(os.path.join(ROOT, "tests", "golden", "qp_golden_h10.npz"))
    rec, rt, params = g["records"][:31], g["robot_type"][:31], g["robot_params"]
    lo, hi = shard_range(len(rec), rank, world)
    out = emu.solve(rec[lo:hi], 10, params, rt[lo:hi])
    status = torch.from_numpy(out["status"])
    resid = torch.from_numpy(out["info"][:, 4:6].copy())
    total = reduce_summary(summarize(status, resid))
    ret[rank] = (lo, hi, out["grf_full"], total.numpy())
    dist.destroy_process_group()


def test_shard_ranges_partition():
    from rl_mpc_locomotion_b200.shard import shard_range
    for n in (0, 1, 7, 4096, 16385):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_shard_equivalence():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu
    emu.build()
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    g = np.load(os.path.join(ROOT, "tests", "golden", "qp_golden_h10.npz"))
    single = emu.solve(g["records"][:31], 10, g["robot_params"], g["robot_type"][:31])
    cat = np.concatenate([ret[0][2], ret[1][2]])
    assert (ret[0][0], ret[0][1], ret[1][0], ret[1][1]) == (0, 16, 16, 31)
    assert np.array_equal(cat, single["grf_full"])          # bitwise: robots are independent
    for k in (0, 1):
        assert ret[k][3][0] == 31 and ret[k][3][1] == 0
        assert np.isclose(ret[k][3][2], single["info"][:, 4].max())


def _nccl_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from rl_mpc_locomotion_b200 import synth
    from rl_mpc_locomotion_b200.batched import BatchedConvexMpc
    from rl_mpc_locomotion_b200.locomotion import BatchedLocomotion
    from rl_mpc_locomotion_b200.shard import reduce_summary, shard_range, summarize
    rec, rt, params, _ = synth.make_batch(301, 10, seed=5)
    lo, hi = shard_range(len(rec), rank, world)
    mpc = BatchedConvexMpc(10, device=dev, robot_params=params)
    out = mpc.solve(torch.from_numpy(rec[lo:hi]).to(dev), torch.from_numpy(rt[lo:hi]).to(dev), full=True)
    total = reduce_summary(summarize(out["status"], out["resid"]))          # NCCL all-reduce of 4 doubles
    # the tick's summary hook: one MPC period on this rank's shard, then the same collective
    n = hi - lo
    loco = BatchedLocomotion(np.zeros(n, np.int32), np.zeros(n, np.int32), device=dev)
    stand = np.tile([0.0, 0.8, -1.6], 4)
    dof = torch.from_numpy(np.stack([np.tile(stand, (n, 1)), np.zeros((n, 12))], -1).astype(np.float32)).to(dev)
    body = torch.tensor([[0, 0, 0.35, 0, 0, 0, 1, 0.3, 0, 0, 0, 0, 0]] * n, dtype=torch.float32, device=dev)
    cmd = torch.zeros((n, 3), dtype=torch.float32, device=dev)
    for _ in range(2):
        loco.run(dof, body, cmd, has_weights=False)
    tick_total = loco.summary()
    ret[rank] = (lo, hi, out["grf_full"].cpu().numpy(), total.cpu().numpy(), tick_total.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_rank_nccl_summary_and_sharded_solve_on_gpus():
    """The one collective north_star allows, on real GPUs over NCCL (world size 2): shards solved on two
    devices equal the single-device solve bit for bit and the summary reduces across ranks."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    from rl_mpc_locomotion_b200 import synth
    from rl_mpc_locomotion_b200.batched import BatchedConvexMpc
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_nccl_worker, args=(world, 29631, ret), nprocs=world, join=True)
    rec, rt, params, _ = synth.make_batch(301, 10, seed=5)
    mpc = BatchedConvexMpc(10, device="cuda:0", robot_params=params)
    ref = mpc.solve(torch.from_numpy(rec).cuda(), torch.from_numpy(rt).cuda(), full=True)
    full = np.concatenate([ret[r][2] for r in range(world)])
    assert np.array_equal(full, ref["grf_full"].cpu().numpy())
    for r in range(world):
        assert ret[r][3][0] == 301 and ret[r][3][1] == 0 and ret[r][3][2] > 0
        assert np.array_equal(ret[r][3], ret[0][3])
        assert ret[r][4][0] == 301 and ret[r][4][1] == 0
