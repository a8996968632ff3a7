"""Regenerates tests/golden/*.npz — run in the BUILD container (needs oracle/ built, which
needs /root/reference for the OSQP sources).  Committed so the fixtures are reproducible.

  qp_golden_h{10,16}.npz : seeded problem records (float32, include/mpc_b200.h layout) and
      the reference OSQP path's answers for them: forces (negated QP solution, all 12h),
      OSQP status / iteration count / polish status (oracle COLD mode = fresh workspace,
      reference settings mpc_osqp.cc:705-712), plus the tight-tolerance minimiser.
  gait_golden.npz : contact tables / contact+swing progress produced by the UNMODIFIED
      reference/MPC_Controller/convex_MPC/Gait.py for every gait and tick 0..39.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from rl_mpc_locomotion_b200 import synth  # noqa: E402
from rl_mpc_locomotion_b200.records import unpack_record  # noqa: E402
from rl_mpc_locomotion_b200.robots import RobotType  # noqa: E402


def qp_golden(h, n, seed):
    rec, rt, params, meta = synth.make_batch(n, h, seed=seed, gaits=(0, 1, 2, 3, 5, 6, 7),
                                             robot_types=tuple(RobotType))
    forces = np.zeros((n, 12 * h))
    truth = np.zeros((n, 12 * h))
    status, iters, polish = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    solvers = {}
    for r in range(n):
        t = int(rt[r])
        if t not in solvers:
            rp = params[t].astype(np.float64)
            solvers[t] = oracle.OracleMpc(rp[0], rp[1:10], horizon=h)
        o = solvers[t]
        kw = unpack_record(rec[r], h)
        status[r], forces[r] = o.solve(oracle.COLD, **kw)
        iters[r], polish[r] = o.last_iters, o.last_polish_status
        _, truth[r] = o.solve(oracle.TRUE, **kw)
    np.savez_compressed(os.path.join(HERE, f"qp_golden_h{h}.npz"), records=rec, robot_type=rt,
                        robot_params=params, forces=forces, true_minimiser=truth, status=status,
                        iters=iters, polish=polish, gait=meta["gait"], seed=seed)
    print(f"h={h}: {n} QPs, iters {np.bincount(iters // 25)}, polish ok {np.mean(polish == 1):.2f}, "
          f"|osqp - true| max {np.abs(forces - truth).max():.2f} N")


def gait_golden():
    ref = "/root/reference"
    sys.path.insert(0, ref)
    from MPC_Controller.convex_MPC.Gait import OffsetDurationGait  # the reference's own class
    from rl_mpc_locomotion_b200.gait import GAITS
    out = {}
    for gid, (off, dur) in GAITS.items():
        g = OffsetDurationGait(10, np.array(off, dtype=np.float32), np.array(dur, dtype=np.float32), str(gid))
        tables, contact, swing = [], [], []
        for tick in range(40):
            g.setIterations(2, tick)
            contact.append(np.array(g.getContactState()).reshape(4).copy())
            swing.append(np.array(g.getSwingState()).reshape(4).copy())
            tables.append(np.array(g.getMpcTable(), dtype=np.float32).copy())
        out[f"table_{gid}"] = np.stack(tables)
        out[f"contact_{gid}"] = np.stack(contact)
        out[f"swing_{gid}"] = np.stack(swing)
    np.savez_compressed(os.path.join(HERE, "gait_golden.npz"), **out)
    print("gait golden written")


if __name__ == "__main__":
    qp_golden(10, 96, 20260922)
    qp_golden(16, 32, 20260923)
    gait_golden()
