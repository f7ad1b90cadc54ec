"""CPU tests of the N > 1 host logic: pair-batch sharding + score gather over torch.distributed (gloo, world_size 2).
The per-rank compute is stood in for by the CPU oracle (tests may use it as the checker); on the GPUs the same
partition / gather code runs around Engine.align_batch (bench.py --workload batch256 --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_balances_cells_and_covers_everything():
    from gpuseqalign_b200.sharding import partition_pairs
    rng = np.random.default_rng(1)
    lenY = rng.integers(0, 600, 1000); lenX = rng.integers(0, 600, 1000)
    for world in (1, 2, 3, 8):
        r = partition_pairs(lenY, lenX, world)
        assert r[0][0] == 0 and r[-1][1] == 1000
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        cells = [float(np.sum(lenY[lo:hi].astype(np.float64) * lenX[lo:hi])) for lo, hi in r]
        assert max(cells) <= 1.15 * (sum(cells) / world) + 600 * 600
    assert partition_pairs(np.array([], dtype=np.int64), np.array([], dtype=np.int64), 4) == [(0, 0)] * 4
    assert partition_pairs(np.array([5]), np.array([5]), 3)[-1][1] == 1


def test_shard_batch_is_self_contained():
    from gpuseqalign_b200.sharding import shard_batch
    from gpuseqalign_b200 import synth
    pool, offY, lenY, offX, lenX = synth.batch_pairs(0, 10, 7, 9)
    sub, oy, ly, ox, lx = shard_batch(pool, offY, lenY, offX, lenX, 3, 8)
    assert sub.size == 5 * 16
    for p in range(5):
        assert np.array_equal(sub[int(oy[p]): int(oy[p]) + 7], pool[int(offY[3 + p]): int(offY[3 + p]) + 7])
        assert np.array_equal(sub[int(ox[p]): int(ox[p]) + 9], pool[int(offX[3 + p]): int(offX[3 + p]) + 9])


def _worker(rank, world, port, q):
    import json
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from gpuseqalign_b200 import synth
    from gpuseqalign_b200.sharding import partition_pairs, shard_batch, gather_scores
    from oracle import pyoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with open(os.path.join(ROOT, "tests", "golden", "scoring.json")) as f:
            subst = np.array(json.load(f)["subst"]["blosum62"], dtype=np.int32)
        rng = np.random.default_rng(9)
        n = 301
        lenY = rng.integers(0, 90, n).astype(np.uint32); lenX = rng.integers(0, 90, n).astype(np.uint32)
        lens = np.empty(2 * n, dtype=np.uint64); lens[0::2] = lenY; lens[1::2] = lenX
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        letters = rng.integers(0, 20, int(offs[-1]) + 1).astype(np.uint8)
        offY, offX = offs[0:-1:2].copy(), offs[1::2].copy()
        ranges = partition_pairs(lenY, lenX, world)
        lo, hi = ranges[rank]
        sub = shard_batch(letters, offY, lenY, offX, lenX, lo, hi)
        local = pyoracle.score_batch(*sub, subst, -11, threads=1)          # stands in for Engine.align_batch on this rank's GPU
        full = gather_scores(local, ranges, rank, world)
        exp = pyoracle.score_batch(letters, offY, lenY, offX, lenX, subst, -11, threads=1)
        q.put((rank, bool(np.array_equal(full, exp)), int(hi - lo)))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_batch_equals_unsharded(oracle):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert sum(k for _, _, k in res) == 301
