"""Multi-GPU sharding of an ensemble: one process per GPU, no data-path collective.

Replicates are independent, replicate g always draws from Philox stream g, so a
rank only needs its range [g_begin, g_end) — results do not depend on the number of ranks.  The
only collective is the sum of the fixed-bin summary histograms (NCCL all-reduce on GPUs; gloo in
the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(rank: int, world: int, total: int):
    """Contiguous block of replicate numbers for `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank outside 0..world-1")
    base, extra = divmod(int(total), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def histograms_from_summaries(summaries: np.ndarray, nv: int) -> np.ndarray:
    """The three 64-bin histograms of jsb_result_histogram, from a summary array (host side)."""
    out = np.zeros((3, 64), dtype=np.int64)
    width = max(1, (int(nv) + 63) // 64)
    a = np.minimum(summaries["n_alive"].astype(np.int64) // width, 63)
    c = np.minimum(summaries["n_clones"].astype(np.int64), 63)
    it = summaries["n_iterations"].astype(np.uint64)
    lg = np.zeros(len(it), dtype=np.int64)
    v = it.copy()
    while True:
        m = v > 1
        if not m.any():
            break
        lg[m] += 1
        v[m] >>= np.uint64(1)
    np.add.at(out[0], a, 1)
    np.add.at(out[1], c, 1)
    np.add.at(out[2], np.minimum(lg, 63), 1)
    return out


def allreduce_histograms(hist: np.ndarray, device=None) -> np.ndarray:
    """Sum over ranks with torch.distributed (backend nccl on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return hist
    t = torch.from_numpy(np.ascontiguousarray(hist, dtype=np.int64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
