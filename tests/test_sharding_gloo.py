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


class _FakeWaveEngine:
    """Stand-in for Engine in the CPU test of the cross-GPU host choreography (gpuseqalign_b200/wavefront.py): records the calls,
    hands out recognisable 64-byte handles, and plays a corridor miss on the first traceback when asked to."""

    def __init__(self, rank, miss_first=False):
        self.rank, self.calls, self.miss_first, self.traces = rank, [], miss_first, 0

    def wave_keep_headers(self, on=True): self.calls.append(("keep", bool(on)))
    def wave_upload(self, y, x, rank, world, block_cols, params=None):
        self.calls.append(("upload", rank, world, block_cols)); return bytes([rank]) * 64
    def scan_upload(self, y, x, rank, world):
        self.calls.append(("scan_upload", rank, world)); return bytes([100 + rank]) * 64
    def wave_connect(self, h): self.calls.append(("connect", None if h is None else h[0]))
    def wave_export_headers(self): return bytes([10 + self.rank]) * 64, bytes([20 + self.rank]) * 64
    def wave_connect_headers(self, hr, sn): self.calls.append(("connect_headers", [h[0] for h in hr], [h[0] for h in sn]))
    def wave_fill(self, epoch): self.calls.append(("fill", epoch))
    def scan_fill(self, epoch): self.calls.append(("scan_fill", epoch))
    def wave_fetch(self): return 1234 if self.rank == 1 else None            # the owner of the last block has the score
    def scan_fetch(self): return 777 if self.rank == 1 else None
    def wave_gather_headers(self, full=False): self.calls.append(("gather", bool(full)))
    def trace_resident(self): self.traces += 1; self.calls.append(("trace",))
    def trace_info(self): return {"corridor_segments": 6, "segments": 100, "corridor_missed": self.miss_first and self.traces == 1}
    def fetch_trace(self, cap=0): return "3=1X", 0xabcdef


def _wave_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from gpuseqalign_b200.wavefront import wave_align, scan_align, wave_trace_setup, wave_trace
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        y = np.zeros(10, np.uint8); x = np.zeros(20, np.uint8)
        e = _FakeWaveEngine(rank)
        s1 = wave_align(e, y, x, rank=rank, world=world, block_cols=64, epoch=5)
        s2 = scan_align(e, y, x, rank=rank, world=world, epoch=6)
        e2 = _FakeWaveEngine(rank, miss_first=True)
        wave_trace_setup(e2, y, x, rank=rank, world=world, block_cols=512)
        e2.wave_fill(7); e2.wave_fetch()
        tr = wave_trace(e2, rank=rank, world=world)
        q.put((rank, s1, s2, e.calls, e2.calls, tr))
    finally:
        dist.destroy_process_group()


def test_two_rank_wavefront_choreography():
    """wave_align / scan_align / wave_trace_setup / wave_trace over gloo with a stand-in engine: every rank maps its RIGHT neighbour's
    receive buffer, every rank gets the score of the last block's owner, only the tracer maps the header buffers, gathers and walks --
    and after a corridor miss gathers everything and walks again -- and nobody is left waiting in a collective."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_wave_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, s1, s2, calls, calls2, tr in res:
        assert (s1, s2) == (1234, 777)
        assert ("connect", (rank + 1) % 2) in calls and ("fill", 5) in calls                    # the right neighbour's wave buffer
        assert ("connect", 100 + (rank + 1) % 2) in calls and ("scan_fill", 6) in calls        # ... and its scan buffer
        assert calls2[0] == ("keep", True) and ("keep", False) in calls2 and ("upload", rank, 2, 512) in calls2
        if rank == 0:
            assert ("connect_headers", [10, 11], [20, 21]) in calls2
            assert [c for c in calls2 if c[0] in ("gather", "trace")] == [("gather", False), ("trace",), ("gather", True), ("trace",)]
            assert tr[0] == "3=1X" and tr[1] == 0xabcdef and tr[2]["corridor_missed"]
        else:
            assert tr is None and not any(c[0] in ("gather", "trace", "connect_headers") for c in calls2)
