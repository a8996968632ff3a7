"""Per-stage latency of one ADMM iteration (cycles, lane 0's view), from a -DMPC_STAGE_TIMING build.
usage: MPC_B200_LIB=ab/lib_timing_x.so [MPC_B200_DEBUG_EXTRA_SMEM=..] python tools/stage_timing.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rl_mpc_locomotion_b200 import synth, _lib
from rl_mpc_locomotion_b200.batched import BatchedConvexMpc
n, K = 148 * 12, 105
rec, rt, params, _ = synth.make_batch(n, 10, seed=1000)
d_rec, d_rt = torch.from_numpy(rec).cuda(), torch.from_numpy(rt).cuda()
m = BatchedConvexMpc(10, device="cuda:0", robot_params=params, kernel_path=_lib.PATH_MONO, mode=1, admm_iters=K, polish=0)
L = _lib.lib()
out = (C.c_ulonglong * 16)()
m.solve(d_rec, d_rt); L.mpc_b200_debug_stage_cycles(out, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m.solve(d_rec, d_rt); e1.record(); torch.cuda.synchronize()
L.mpc_b200_debug_stage_cycles(out, 1)
names = {0: "s = Dinv rhs", 1: "g = (F^-1 U) s", 4: "w = M g (dense)", 5: "y = W g (row walk)", 6: "w = W^T y (col walk)",
         7: "v = g - F^-1 w", 8: "x = s - Dinv U^T v", 12: "rhs", 13: "update x z y"}
tot = 0
for k, v in names.items():
    c = out[k] / (n * K)
    tot += c
    if out[k]: print(f"  {v:24s} {c:8.1f} cycles / iteration")
print(f"  sum {tot:.0f} cycles / iteration ; kernel {e0.elapsed_time(e1):.3f} ms for {n} robots x {K} iterations "
      f"(extra smem {os.environ.get('MPC_B200_DEBUG_EXTRA_SMEM', '0')})")
