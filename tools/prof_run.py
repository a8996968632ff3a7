"""Tiny driver for ncu captures: one launch of the solve kernel on a small batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_mpc_locomotion_b200 import synth
from rl_mpc_locomotion_b200.batched import BatchedConvexMpc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1480
h = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rec, rt, params, _ = synth.make_batch(n, h, seed=0)
mpc = BatchedConvexMpc(h, device="cuda:0", robot_params=params)
d_rec, d_rt = torch.from_numpy(rec).cuda(), torch.from_numpy(rt).cuda()
for _ in range(2):
    mpc.solve(d_rec, d_rt)
torch.cuda.synchronize()
