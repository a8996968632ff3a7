"""Device timing of the fused Locomotion tick (BASELINE config 3 shard: mixed trot/walk/bound)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rl_mpc_locomotion_b200.locomotion import BatchedLocomotion

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
warm = len(sys.argv) > 2 and sys.argv[2] == "warm"
rng = np.random.default_rng(0)
gait = rng.choice([0, 6, 1], n)
lf = len(sys.argv) > 3 and sys.argv[3] == "lf"
loco = BatchedLocomotion(np.zeros(n, np.int32), gait, device="cuda:0", warm_start=warm, longest_first=lf)
stand = np.tile([0.0, 0.8, -1.6], 4)


def inputs(t):
    dof = np.stack([stand + 0.1 * np.sin(0.3 * t + rng.uniform(0, 6, (n, 12))), 0.3 * rng.standard_normal((n, 12))], -1)
    rpy = 0.05 * rng.standard_normal((n, 3))
    quat = np.stack([rpy[:, 0] / 2, rpy[:, 1] / 2, rpy[:, 2] / 2, np.ones(n)], 1)
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    body = np.concatenate([np.tile([0, 0, 0.35], (n, 1)), quat, [0.5, 0, 0] + 0.1 * rng.standard_normal((n, 3)),
                           0.2 * rng.standard_normal((n, 3))], 1)
    cmd = np.zeros((n, 16)); cmd[:, 0] = 0.5
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
    return f(dof), f(body), f(cmd)


ins = [inputs(t) for t in range(8)]
for i in range(8):
    loco.run(ins[i][0], ins[i][1], ins[i][2], has_weights=False)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
ev[0].record()
for i in range(40):
    d, b, c = ins[i % 8]
    loco.run(d, b, c, has_weights=False)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(40)]
mpc_ms, plain_ms = np.mean(sorted(ms)[20:]), np.mean(sorted(ms)[:20])
print(f"n={n} warm_start={warm} longest_first={lf}: MPC ticks {mpc_ms:.3f} ms, non-MPC ticks {plain_ms:.3f} ms -> {n * 40 / sum(ms) * 1e3:.3e} robot-ticks/s "
      f"({n / mpc_ms * 1e3:.3e} QP solves/s on MPC ticks); solved {int((loco.solved == 1).sum())}/{n}")
