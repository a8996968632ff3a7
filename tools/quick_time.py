"""Quick device timing of the solve kernel (development aid; bench.py is the contract).
usage: python tools/quick_time.py [n] [h]   (MPC_B200_LIB=<path> selects an alternative build)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rl_mpc_locomotion_b200 import synth, _lib
from rl_mpc_locomotion_b200.batched import BatchedConvexMpc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
h = int(sys.argv[2]) if len(sys.argv) > 2 else 10
print(os.environ.get("MPC_B200_LIB", "default lib"), _lib.kernel_info(h))
rec, rt, params, _ = synth.make_batch(n, h, seed=0)
d_rec, d_rt = torch.from_numpy(rec).cuda(), torch.from_numpy(rt).cuda()
modes = (("osqp", {}), ("osqp_nopolish", dict(polish=0)), ("fixed105", dict(mode=1, admm_iters=105, polish=0)),
         ("fixed50", dict(mode=1, admm_iters=50, polish=0)), ("fixed1", dict(mode=1, admm_iters=1, polish=0)),
         ("fixed1_noscale", dict(mode=1, admm_iters=1, polish=0, scaling_iters=0)))
for mode, kw in modes:
    m2 = BatchedConvexMpc(h, device="cuda:0", robot_params=params, **kw)
    for _ in range(3):
        m2.solve(d_rec, d_rt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        out = m2.solve(d_rec, d_rt, info=(mode == "osqp"))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    extra = ""
    if mode == "osqp":
        info = out["info"].cpu().numpy()
        extra = f"; iters mean {info[:,0].mean():.1f} max {info[:,0].max():.0f} rho_upd {info[:,1].mean():.2f} polish ok {np.mean(info[:,3]==1):.2f}"
    print(f"h={h} n={n} {mode:14s}: {ms:7.3f} ms -> {n/ms*1e3:.3e} QP/s{extra}")
