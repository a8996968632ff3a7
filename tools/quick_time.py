"""Quick device timing of the solve kernel (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rl_mpc_locomotion_b200 import synth, _lib
from rl_mpc_locomotion_b200.batched import BatchedConvexMpc

print(_lib.kernel_info(10), _lib.kernel_info(16))
for h, n in ((10, 4096), (10, 16384), (16, 4096)):
    rec, rt, params, _ = synth.make_batch(n, h, seed=0)
    mpc = BatchedConvexMpc(h, device="cuda:0", robot_params=params)
    d_rec, d_rt = torch.from_numpy(rec).cuda(), torch.from_numpy(rt).cuda()
    out = mpc.solve(d_rec, d_rt, info=True)
    torch.cuda.synchronize()
    info = out["info"].cpu().numpy()
    for mode, kw in (("osqp", {}), ("fixed50", dict(mode=1, admm_iters=50, polish=0))):
        m2 = BatchedConvexMpc(h, device="cuda:0", robot_params=params, **kw)
        for _ in range(3):
            m2.solve(d_rec, d_rt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            m2.solve(d_rec, d_rt)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"h={h} n={n} mode={mode}: {ms:.3f} ms/launch -> {n/ms*1e3:.3e} QP/s ; mean iters {info[:,0].mean():.1f} rho_upd {info[:,1].mean():.2f} polish ok {np.mean(info[:,3]==1):.2f}")
