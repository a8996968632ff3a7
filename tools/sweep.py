"""BASELINE config 5: batch sweep 256 -> 65536 robots per GPU x {OSQP-faithful, fixed 25/50/100 ADMM iterations}
on 1 / 2 / 4 / 8 GPUs (weak scaling: every rank owns its own shard; device time, max over ranks).

  python tools/sweep.py                                   # 1 GPU   (SWEEP_SIZES=1024,4096,... trims the size list)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sweep.py

Rank 0 prints one JSON document (also written to gpurun_out/sweep_config5_<N>gpu.json)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from bench import kernel_hash
from rl_mpc_locomotion_b200 import _lib, synth
from rl_mpc_locomotion_b200.batched import BatchedConvexMpc

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
sizes = [int(x) for x in os.environ.get("SWEEP_SIZES", "256,512,1024,2048,4096,8192,16384,32768,65536").split(",")]
modes = [("osqp", {})] + [(f"fixed{k}", dict(mode=_lib.MODE_FIXED, admm_iters=k, polish=0)) for k in (25, 50, 100)]
rec_all, rt_all, params, _ = synth.make_batch(max(sizes), 10, seed=1000 + rank)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []
for n in sizes:
    d_rec, d_rt = torch.from_numpy(rec_all[:n]).to(dev), torch.from_numpy(rt_all[:n]).to(dev)
    d_out = torch.empty((n, 12), dtype=torch.float32, device=dev)
    for name, kw in modes:
        mpc = BatchedConvexMpc(10, device=dev, robot_params=params, **kw)
        for _ in range(3):
            mpc.solve(d_rec, d_rt, out=d_out)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        steps = 6 if n <= 16384 else 3
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for e0, e1 in ev:
            flush.fill_(1)
            e0.record()
            out = mpc.solve(d_rec, d_rt, out=d_out)
            e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        solved = float((out["status"] == 1).float().mean())
        rows.append({"robots_per_gpu": n, "mode": name, "n_gpus": world, "ms_per_step": float(ms) / steps,
                     "solves_per_s": world * n * steps / (float(ms) * 1e-3), "solved_fraction_rank0": solved})
        if rank == 0:
            print(rows[-1], file=sys.stderr, flush=True)
if rank == 0:
    doc = {"workload": "config 5 sweep, Aliengo trot h=10, seed 1000+rank", "kernel_hash": kernel_hash(), "n_gpus": world,
           "timing": "CUDA events per step, L2 flushed between steps, max over ranks", "rows": rows}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(doc, open(os.path.join(ROOT, "gpurun_out", f"sweep_config5_{world}gpu.json"), "w"), indent=1)
    print(json.dumps(doc))
if world > 1:
    dist.destroy_process_group()
