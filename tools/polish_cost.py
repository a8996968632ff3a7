"""What the polish costs when every robot runs it in lockstep (fixed iteration count) vs inside the
OSQP-faithful mix of phases (development measurement).  usage: python tools/polish_cost.py [n] [h]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rl_mpc_locomotion_b200 import synth
from rl_mpc_locomotion_b200.batched import BatchedConvexMpc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
h = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rec, rt, params, _ = synth.make_batch(n, h, seed=0)
d_rec, d_rt = torch.from_numpy(rec).cuda(), torch.from_numpy(rt).cuda()
for name, kw in (("osqp", {}), ("osqp_nopolish", dict(polish=0)), ("fixed50", dict(mode=1, admm_iters=50, polish=0)),
                 ("fixed50+polish", dict(mode=1, admm_iters=50, polish=1)), ("fixed100", dict(mode=1, admm_iters=100, polish=0)),
                 ("fixed100+polish", dict(mode=1, admm_iters=100, polish=1))):
    m = BatchedConvexMpc(h, device="cuda:0", robot_params=params, **kw)
    for _ in range(3): m.solve(d_rec, d_rt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): m.solve(d_rec, d_rt)
    e1.record(); torch.cuda.synchronize()
    print(f"h={h} n={n} {name:16s}: {e0.elapsed_time(e1)/5:7.3f} ms")
