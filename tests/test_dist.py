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
