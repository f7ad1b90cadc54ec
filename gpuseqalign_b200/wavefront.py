"""One very long pair across the GPUs of one box: the host side of the column-block wavefront (BASELINE configs 4, 5).

One process per GPU (torch.distributed).  The only things that cross the process group are the 64-byte CUDA IPC
handles of the receive buffers (setup), a barrier, and the final 4-byte score; the border columns themselves move
GPU-to-GPU as peer stores issued by the fill kernel (csrc/nw_fill.cuh), not through NCCL.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def _connect(engine, handle: bytes, rank: int, world: int, group=None):
    """Exchange the IPC handles of the receive buffers and map the right-hand neighbour's."""
    if world == 1:
        engine.wave_connect(None)
        return
    import torch.distributed as dist
    handles: list = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    engine.wave_connect(handles[(rank + 1) % world])
    dist.barrier(group=group)                       # every receive buffer exists and is mapped before anyone pushes into it


def wave_setup(engine, y: np.ndarray, x: np.ndarray, *, rank: int = 0, world: int = 1, block_cols: int = 2048, params=None, group=None):
    """Upload + connect, without a fill: callers that re-run the resident pair (benchmarks) then loop over
    ``engine.wave_fill(epoch)`` / ``engine.wave_fetch()`` with a fresh epoch and a barrier per run."""
    _connect(engine, engine.wave_upload(y, x, rank, world, block_cols, params), rank, world, group)


def scan_setup(engine, y: np.ndarray, x: np.ndarray, *, rank: int = 0, world: int = 1, group=None):
    """Same for the prefix-max scorer: then ``engine.scan_fill(epoch)`` / ``engine.scan_fetch()``."""
    _connect(engine, engine.scan_upload(y, x, rank, world), rank, world, group)


def wave_align(engine, y: np.ndarray, x: np.ndarray, *, rank: int = 0, world: int = 1, block_cols: int = 2048,
               epoch: int = 1, params=None, group=None) -> int:
    """Score of NW(y, x); every rank passes the same y, x, block_cols and epoch and gets the same score back."""
    wave_setup(engine, y, x, rank=rank, world=world, block_cols=block_cols, params=params, group=group)
    if world == 1:
        engine.wave_fill(epoch)
        return engine.wave_fetch()
    import torch
    import torch.distributed as dist
    engine.wave_fill(epoch)
    score = engine.wave_fetch()
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([score if score is not None else -(2 ** 62)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def wave_trace_setup(engine, y: np.ndarray, x: np.ndarray, *, rank: int = 0, world: int = 1, block_cols: int = 2048, tracer: int = 0,
                     params=None, group=None):
    """wave_setup for a fill that is followed by a traceback on rank `tracer` (world > 1): every rank keeps the header rows and
    snapshots of its column blocks, the tracer maps them all.  block_cols must be a multiple of the snapshot spacing (512 columns for
    pairs longer than 65 536, 256 below, or params.tile_cols)."""
    import torch.distributed as dist
    if world < 2:
        raise ValueError("wave_trace_setup is for world > 1; on one GPU use Engine.align(y, x, keep_headers=True) and Engine.trace()")
    engine.wave_keep_headers(True)
    try:
        handle = engine.wave_upload(y, x, rank, world, block_cols, params)
    finally:
        engine.wave_keep_headers(False)
    hs = engine.wave_export_headers()
    _connect(engine, handle, rank, world, group)
    handles: list = [None] * world
    dist.all_gather_object(handles, hs, group=group)
    if rank == tracer:
        engine.wave_connect_headers([h[0] for h in handles], [h[1] for h in handles])
    dist.barrier(group=group)


def wave_trace(engine, *, rank: int = 0, world: int = 1, tracer: int = 0, cap: int = 0, group=None):
    """After wave_fill + wave_fetch on every rank: rank `tracer` pulls the headers the traceback's corridor can touch over NVLink, runs
    the traceback and -- when the path left the corridor -- repeats it on all headers.  Returns (transcript, trace hash, info) on the
    tracer, None elsewhere.  Collective: contains the barrier that separates the fills from the peer copies."""
    import torch.distributed as dist
    dist.barrier(group=group)                      # every rank has synchronised its fill (wave_fetch does)
    out = None
    if rank == tracer:
        engine.wave_gather_headers(False)
        engine.trace_resident()
        info = engine.trace_info()
        if info["corridor_missed"]:
            engine.wave_gather_headers(True)
            engine.trace_resident()
        edit, th = engine.fetch_trace(cap) if cap else engine.fetch_trace()
        out = (edit, th, info)
    dist.barrier(group=group)                      # nobody starts the next fill (and overwrites its headers) while the tracer still reads them
    return out


def scan_align(engine, y: np.ndarray, x: np.ndarray, *, rank: int = 0, world: int = 1, epoch: int = 1, group=None) -> int:
    """Score of NW(y, x) for a matrix with few rows and very many columns (row-parallel prefix max, csrc/nw_scan.cuh);
    the chunks of 4096 columns are dealt to the ranks in contiguous ranges.  Same calling convention as wave_align."""
    scan_setup(engine, y, x, rank=rank, world=world, group=group)
    if world == 1:
        engine.scan_fill(epoch)
        return engine.scan_fetch()
    import torch
    import torch.distributed as dist
    engine.scan_fill(epoch)
    score = engine.scan_fetch()
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([score if score is not None else -(2 ** 62)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())
