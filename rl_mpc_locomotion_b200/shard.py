"""Multi-GPU plumbing: robots are independent, so the batch shards by contiguous ranges
 with NO collective on the data path.  The only exchange the RL side may want is
one reduction of per-GPU summaries (solved / failed counts, worst residuals) per tick.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple:
    """Contiguous slice [lo, hi) of robots owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def summarize(status: torch.Tensor, resid: torch.Tensor | None = None) -> torch.Tensor:
    """Local summary vector [n_solved, n_not_solved, max_pri_res, max_dua_res] (float64)."""
    solved = (status == 1).sum().to(torch.float64)
    out = torch.zeros(4, dtype=torch.float64, device=status.device)
    out[0], out[1] = solved, status.numel() - solved
    if resid is not None and resid.numel():
        out[2], out[3] = resid[:, 0].max().to(torch.float64), resid[:, 1].max().to(torch.float64)
    return out


def reduce_summary(local: torch.Tensor) -> torch.Tensor:
    """The single collective: counts are summed, residuals maxed, over all ranks (NCCL on GPU
    tensors, gloo on CPU tensors).  No-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    counts, worst = local[:2].clone(), local[2:].clone()
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    return torch.cat([counts, worst])
