"""BASELINE config 5: batch sweep 256 -> 65536 robots x ADMM iterations 25/50/100 (fixed mode)
plus the reference-faithful OSQP mode, one GPU.  Writes gpurun_out/sweep_r1.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rl_mpc_locomotion_b200 import synth
from rl_mpc_locomotion_b200.batched import BatchedConvexMpc

rows = []
for n in (256, 1024, 4096, 16384, 65536):
    rec, rt, params, _ = synth.make_batch(n, 10, seed=0)
    d_rec, d_rt = torch.from_numpy(rec).cuda(), torch.from_numpy(rt).cuda()
    for mode, kw in (("osqp", {}), ("fixed25", dict(mode=1, admm_iters=25, polish=0)),
                     ("fixed50", dict(mode=1, admm_iters=50, polish=0)), ("fixed100", dict(mode=1, admm_iters=100, polish=0))):
        m = BatchedConvexMpc(10, device="cuda:0", robot_params=params, **kw)
        for _ in range(3):
            m.solve(d_rec, d_rt)
        torch.cuda.synchronize()
        reps = 5 if n <= 16384 else 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            m.solve(d_rec, d_rt)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rows.append({"robots": n, "mode": mode, "ms_per_launch": ms, "solves_per_s": n / ms * 1e3})
        print(rows[-1])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep_r1.json", "w"), indent=1)
