"""Monolithic vs phase-split batch path of the SAME library: device time and bit-identity of the outputs.
usage: python tools/ab_split.py [n] [h]      (MPC_B200_LIB selects an alternative build)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rl_mpc_locomotion_b200 import synth, _lib
from rl_mpc_locomotion_b200.batched import BatchedConvexMpc, lpt_order

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
h = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rec, rt, params, _ = synth.make_batch(n, h, seed=1000)
d_rec, d_rt = torch.from_numpy(rec).cuda(), torch.from_numpy(rt).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
outs = {}
for name, path in (("mono", _lib.PATH_MONO), ("split", _lib.PATH_SPLIT)):
    m = BatchedConvexMpc(h, device="cuda:0", robot_params=params, kernel_path=path)
    o = m.solve(d_rec, d_rt, full=True, info=True)
    outs[name] = o
    order = lpt_order(o["info"])
    for label, kw in (("index order", {}), ("longest-first", dict(order=order))):
        for _ in range(3):
            m.solve(d_rec, d_rt, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); m.solve(d_rec, d_rt, **kw); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"{os.path.basename(os.environ.get('MPC_B200_LIB', 'default'))} h={h} n={n} {name:5s} {label:13s}: "
              f"{np.median(ts):.3f} ms (min {min(ts):.3f}) -> {n / np.median(ts) * 1e3:.4e} solves/s", flush=True)
a, b = outs["mono"], outs["split"]
print("bit-identical outputs:", torch.equal(a["grf_full"], b["grf_full"]), torch.equal(a["status"], b["status"]),
      torch.equal(a["info"], b["info"]), "solved", int((b["status"] == 1).sum()), "/", n)
