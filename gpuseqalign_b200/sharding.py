"""Host-side sharding of pair batches over the GPUs of one box (BASELINE config 3).

Pairs are independent objects: rank r aligns a contiguous range of pairs, chosen so that the ranks get (nearly) equal
numbers of matrix cells; there is no data-path collective, only a gather of the 4-byte scores (torch.distributed,
NCCL on the GPUs, gloo in the CPU tests).  The reference has nothing like it (single process, device 0,
benchmark.cpp:179,406).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def partition_pairs(lenY: np.ndarray, lenX: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous ranges [lo, hi) per rank, balanced by sum(lenY * lenX) (ties: pairs are never split)."""
    n = len(lenY)
    if world < 1:
        raise ValueError("world must be >= 1")
    cells = np.asarray(lenY, dtype=np.float64) * np.asarray(lenX, dtype=np.float64) + 1.0     # +1: empty pairs still cost a slot
    csum = np.concatenate([[0.0], np.cumsum(cells)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(csum, target, side="left"))
        k = max(bounds[-1], min(n, k))
        bounds.append(k)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard_batch(letters: np.ndarray, offY, lenY, offX, lenX, lo: int, hi: int):
    """The sub-batch [lo, hi) with its own compact letter pool (what one rank uploads to its GPU)."""
    offY = np.asarray(offY, dtype=np.uint64)[lo:hi]; offX = np.asarray(offX, dtype=np.uint64)[lo:hi]
    lenY = np.asarray(lenY, dtype=np.uint32)[lo:hi]; lenX = np.asarray(lenX, dtype=np.uint32)[lo:hi]
    k = hi - lo
    lens = np.empty(2 * k, dtype=np.uint64)
    lens[0::2] = lenY; lens[1::2] = lenX
    starts = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    pool = np.empty(int(starts[-1]), dtype=np.uint8)
    newY = starts[0:-1:2].copy(); newX = starts[1::2].copy()
    for p in range(k):
        pool[int(newY[p]): int(newY[p]) + int(lenY[p])] = letters[int(offY[p]): int(offY[p]) + int(lenY[p])]
        pool[int(newX[p]): int(newX[p]) + int(lenX[p])] = letters[int(offX[p]): int(offX[p]) + int(lenX[p])]
    return pool, newY, lenY, newX, lenX


def gather_scores(local_scores: np.ndarray, ranges: List[Tuple[int, int]], rank: int, world: int):
    """All ranks receive the full score vector.  Uses the default torch.distributed process group."""
    import torch
    import torch.distributed as dist
    n = ranges[-1][1]
    if world == 1:
        return np.asarray(local_scores, dtype=np.int32)
    longest = max(hi - lo for lo, hi in ranges)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.zeros(longest, dtype=torch.int32, device=dev)
    buf[: len(local_scores)] = torch.from_numpy(np.asarray(local_scores, dtype=np.int32)).to(dev)
    out = [torch.zeros(longest, dtype=torch.int32, device=dev) for _ in range(world)]
    dist.all_gather(out, buf)
    full = np.empty(n, dtype=np.int32)
    for r, (lo, hi) in enumerate(ranges):
        full[lo:hi] = out[r][: hi - lo].cpu().numpy()
    return full
