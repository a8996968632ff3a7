"""Throughput of the monolithic solve kernel as a function of resident CTAs per SM (occupancy limited by
extra dynamic shared memory, MPC_B200_DEBUG_EXTRA_SMEM): is the kernel latency-bound (linear in CTAs)?"""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r)
import numpy as np, torch
from rl_mpc_locomotion_b200 import synth, _lib
from rl_mpc_locomotion_b200.batched import BatchedConvexMpc
rec, rt, params, _ = synth.make_batch(148 * 30, 10, seed=1000)
d_rec, d_rt = torch.from_numpy(rec).cuda(), torch.from_numpy(rt).cuda()
for name, kw in (("osqp", {}), ("fixed105", dict(mode=1, admm_iters=105, polish=0))):
    m = BatchedConvexMpc(10, device="cuda:0", robot_params=params, kernel_path=_lib.PATH_MONO, **kw)
    for _ in range(3): m.solve(d_rec, d_rt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): m.solve(d_rec, d_rt)
    e1.record(); torch.cuda.synchronize()
    print(name, "%%.3f ms" %% (e0.elapsed_time(e1) / 5), end="  ")
print()
''' % root
for ctas, extra in ((6, 0), (5, 6 * 1024), (4, 14 * 1024), (3, 32 * 1024), (2, 70 * 1024), (1, 150 * 1024)):
    env = dict(os.environ, MPC_B200_DEBUG_EXTRA_SMEM=str(extra))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(f"CTAs/SM {ctas} (extra smem {extra}): {out.stdout.strip()} {out.stderr.strip()[-200:]}", flush=True)
