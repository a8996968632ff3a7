"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): one solve launch
(cold + warm) and a few fused ticks on 24 robots."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rl_mpc_locomotion_b200 import synth
from rl_mpc_locomotion_b200.batched import BatchedConvexMpc
from rl_mpc_locomotion_b200.locomotion import BatchedLocomotion
from rl_mpc_locomotion_b200.robots import RobotType

n = 24
for h in (10, 16):
    rec, rt, params, _ = synth.make_batch(n, h, seed=5, gaits=(0, 6, 1), robot_types=tuple(RobotType))
    mpc = BatchedConvexMpc(h, device="cuda:0", robot_params=params)
    warm = mpc.warm_buffer(n)
    for _ in range(2):
        out = mpc.solve(torch.from_numpy(rec).cuda(), torch.from_numpy(rt).cuda(), full=True, info=True, warm=warm)
    out = mpc.solve(torch.from_numpy(rec).cuda(), torch.from_numpy(rt).cuda())
    torch.cuda.synchronize()
    print("h", h, "solved", int((out["status"] == 1).sum()), "/", n)
rng = np.random.default_rng(0)
loco = BatchedLocomotion(rng.integers(0, 3, n), rng.choice([0, 6, 1], n), device="cuda:0", warm_start=True)
stand = np.tile([0.0, 0.8, -1.6], 4)
f = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
for t in range(4):
    dof = np.stack([stand + 0.05 * rng.standard_normal((n, 12)), 0.2 * rng.standard_normal((n, 12))], -1)
    body = np.concatenate([np.tile([0, 0, 0.35], (n, 1)), np.tile([0, 0, 0, 1.0], (n, 1)), np.tile([0.4, 0, 0], (n, 1)),
                           0.1 * rng.standard_normal((n, 3))], 1)
    cmd = np.zeros((n, 16)); cmd[:, 0] = 0.5
    tau = loco.run(f(dof), f(body), f(cmd), has_weights=False)
torch.cuda.synchronize()
print("tick ok", float(tau.abs().max()))
