"""Multi-GPU partition of the batched workload.

The only workload that shards is the batch of independent perturbed-load ACOPF instances:
instance ``b`` goes to rank ``b // ceil(B / world)`` (contiguous blocks), the sparsity
pattern is replicated, and NOTHING is exchanged while the instances are being solved.  The
single collective is one all-gather of ``(status:int32, iters:int32, objective:float64)`` =
16 bytes per instance at the end (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

RESULT_DTYPE = np.dtype([("status", "<i4"), ("iters", "<i4"), ("objective", "<f8")])


def shard_range(batch: int, rank: int, world: int):
    """Contiguous block of instance ids owned by ``rank``: [lo, hi)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = -(-batch // world)
    lo = min(rank * per, batch)
    hi = min(lo + per, batch)
    return lo, hi


def pack_results(status, iters, objective):
    out = np.zeros(len(status), dtype=RESULT_DTYPE)
    out["status"], out["iters"], out["objective"] = status, iters, objective
    return out


def gather_results(local: np.ndarray, batch: int, device=None):
    """All-gather the per-instance result records of every rank (16 B per instance).

    Works with any initialised ``torch.distributed`` backend; without a process group it
    returns ``local`` unchanged (single process).
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-batch // world)
    buf = np.zeros(per, dtype=RESULT_DTYPE)
    buf["status"] = -999  # padding marker
    buf[: local.shape[0]] = local
    t = torch.from_numpy(buf.view(np.uint8).copy())
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    rec = np.concatenate([o.cpu().numpy().view(RESULT_DTYPE) for o in outs])
    return rec[rec["status"] != -999][:batch]
