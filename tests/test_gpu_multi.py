"""Parity of every N > 1 path on REAL GPUs (one process per GPU, torch.distributed over NCCL), and of the engine against
the reference's own gpu9 / own benchmark executable on the box.

The multi-rank tests skip with a reason on a one-GPU lease (the same protocols run there with world = 1 in
test_gpu_fill.py / test_gpu_big.py); on a multi-GPU lease they spawn 2 ranks (4 when the box has them):
  * column-block wavefront (wave_align): a 3 000 x 20 000 pair against the oracle, cfg5 (200 000^2) against its golden score;
  * cross-GPU fill + traceback from gathered headers (wave_trace): 40 000^2 pairs against the oracle, cfg5 against its golden transcript;
  * prefix-max scorer (scan_align): small shapes against the oracle, cfg4 (2 048 x 4 194 304) against its golden score;
  * sharded align_batch: every rank aligns its shard on ITS GPU, the gathered score vector against the oracle;
  * the strong-scaling shards of the full cfg3 job: per-eighth sha-256 of the scores against tests/golden/batch_golden.json.
"""
import hashlib
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, what, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from gpuseqalign_b200 import Engine, synth
    from gpuseqalign_b200.wavefront import wave_align, scan_align
    from gpuseqalign_b200.sharding import partition_pairs, shard_batch, gather_scores
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    import datetime
    # (a short collective timeout: a rank that fails must not leave the others waiting in a barrier for ten minutes)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank), timeout=datetime.timedelta(seconds=120))
    out = {}
    try:
        with open(os.path.join(ROOT, "tests", "golden", "scoring.json")) as f:
            subst = np.array(json.load(f)["subst"]["blosum62"], dtype=np.int32)
        big = json.load(open(os.path.join(ROOT, "tests", "golden", "big_golden.json")))
        eng = Engine(rank)
        eng.set_scoring(subst, -11)
        epoch = 50
        if what == "wave":
            from oracle import pyoracle
            y = synth.letters(61, 3000); x = synth.letters(62, 20000)
            exp = pyoracle.fill_rolling(y, x, subst, -11)[0]
            for block in (512, 2048):
                epoch += 1
                out[f"small_block{block}"] = (wave_align(eng, y, x, rank=rank, world=world, block_cols=block, epoch=epoch), exp)
            # more rows than columns per block, ragged last block, a pair shorter than one block per rank
            y = synth.letters(63, 5000); x = synth.letters(64, 1500)
            exp = pyoracle.fill_rolling(y, x, subst, -11)[0]
            epoch += 1
            out["tall_ragged"] = (wave_align(eng, y, x, rank=rank, world=world, block_cols=1024, epoch=epoch), exp)
            x = synth.letters(5001, 200000); y = synth.letters(5004, 200000)
            epoch += 1
            out["cfg5"] = (wave_align(eng, y, x, rank=rank, world=world, block_cols=2048, epoch=epoch), big["cfg5_random"]["score"])
        elif what == "wavetrace":
            # cross-GPU fill + traceback on rank 0 from the gathered headers (nwb200_wave_gather_headers): transcript and trace hash
            # against the oracle (moderate pairs: path inside the corridor, and a long indel that leaves it) and cfg5 against its golden
            from oracle import pyoracle
            from gpuseqalign_b200.wavefront import wave_trace_setup, wave_trace
            n = 40000
            x = synth.letters(611, n)
            cases = {"mutated": synth.mutated_copy(x, 612, n), "random": synth.letters(614, n),
                     "long_indel": np.concatenate([x[:19000], x[21500:], synth.letters(613, 2500)])}
            os.environ["NWB200_CORRIDOR"] = "512"
            for name, y in cases.items():
                exp = pyoracle.align_pair(y, x, subst, -11, want_hash=False, want_trace=True) if rank == 0 else None
                for block in (5120, 20480):
                    epoch += 1
                    wave_trace_setup(eng, y, x, rank=rank, world=world, block_cols=block)
                    eng.wave_fill(epoch)
                    score = eng.wave_fetch()
                    if os.environ.get("NWB200_TEST_VERBOSE"): print(f"rank {rank}: {name} block {block}: filled, score {score}", file=sys.stderr, flush=True)
                    tr = wave_trace(eng, rank=rank, world=world)
                    if os.environ.get("NWB200_TEST_VERBOSE"): print(f"rank {rank}: {name} block {block}: traced {tr[2] if tr else None}", file=sys.stderr, flush=True)
                    if rank == 0:
                        out[f"{name}_block{block}_trace"] = ((tr[0], tr[1]), (exp.edit, exp.trace_hash))
                        out[f"{name}_block{block}_missed"] = (tr[2]["corridor_missed"], name == "long_indel")
                    if score is not None:
                        out[f"{name}_block{block}_score"] = (score, int(pyoracle.fill_rolling(y, x, subst, -11)[0]))
            del os.environ["NWB200_CORRIDOR"]
            x = synth.letters(5001, 200000); y = synth.letters(5004, 200000)
            epoch += 1
            wave_trace_setup(eng, y, x, rank=rank, world=world, block_cols=(200000 // world + 511) // 512 * 512)
            eng.wave_fill(epoch)
            score = eng.wave_fetch()
            tr = wave_trace(eng, rank=rank, world=world, cap=1 << 20)
            g5 = big["cfg5_random"]
            if score is not None:
                out["cfg5_score"] = (score, g5["score"])
            if rank == 0:
                out["cfg5_trace"] = ((f"{tr[1]:08x}", hashlib.sha256(tr[0].encode()).hexdigest()), (g5["trace_hash"], g5["edit_sha256"]))
        elif what == "scan":
            from oracle import pyoracle
            for n, m, sy, sx in ((37, 9000, 71, 72), (300, 70000, 73, 74), (1, 5000, 75, 76)):
                y = synth.letters(sy, n); x = synth.letters(sx, m)
                exp = pyoracle.fill_rolling(y, x, subst, -11)[0]
                epoch += 1
                out[f"scan_{n}x{m}"] = (scan_align(eng, y, x, rank=rank, world=world, epoch=epoch), exp)
            y = synth.letters(4001, 2048); x = synth.letters(4002, 4194304)
            epoch += 1
            out["cfg4"] = (scan_align(eng, y, x, rank=rank, world=world, epoch=epoch), big["cfg4"]["score"])
        elif what == "batch":
            from oracle import pyoracle
            rng = np.random.default_rng(9)
            n = 4001
            lenY = rng.integers(0, 300, n).astype(np.uint32); lenX = rng.integers(0, 300, n).astype(np.uint32)
            lens = np.empty(2 * n, dtype=np.uint64); lens[0::2] = lenY; lens[1::2] = lenX
            offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
            letters = rng.integers(0, 20, int(offs[-1]) + 1).astype(np.uint8)
            offY, offX = offs[0:-1:2].copy(), offs[1::2].copy()
            ranges = partition_pairs(lenY, lenX, world)
            lo, hi = ranges[rank]
            local = eng.align_batch(*shard_batch(letters, offY, lenY, offX, lenX, lo, hi))      # on THIS rank's GPU
            full = gather_scores(local, ranges, rank, world)
            exp = pyoracle.score_batch(letters, offY, lenY, offX, lenX, subst, -11, threads=4)
            out["ragged_sharded"] = (int(np.count_nonzero(full != exp)), 0)
            # the strong-scaling shards of the full cfg3 job
            gold = json.load(open(os.path.join(ROOT, "tests", "golden", "batch_golden.json")))
            total = gold["pairs"]; e8 = total // 8
            lo, hi = rank * total // world, (rank + 1) * total // world
            pool, oY, lY, oX, lX = synth.batch_pairs(lo, hi - lo, 256, 256)
            sc = eng.align_batch(pool, oY, lY, oX, lX)
            ok = all(hashlib.sha256(sc[k * e8 - lo:(k + 1) * e8 - lo].tobytes()).hexdigest() == gold["sha256_eighths"][k] for k in range(lo // e8, hi // e8))
            t = torch.tensor([1 if ok else 0], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            out["cfg3_full_job_checksums"] = (int(t.item()), 1)
        eng.close()
        q.put((rank, out, None))
    except Exception as ex:          # reported to the parent, which fails the test
        import traceback
        q.put((rank, out, f"{type(ex).__name__}: {ex}\n{traceback.format_exc()}"))
        os._exit(1)                   # (no orderly shutdown: the other ranks may be waiting in a collective this rank will never join)
    finally:
        dist.destroy_process_group()


def _run(world, what, timeout=300):
    import torch.multiprocessing as mp
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, what, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue as _queue
    res = []
    try:
        for _ in procs:
            res.append(q.get(timeout=timeout))
            if res[-1][2] is not None:
                break                                   # a rank failed: the others cannot finish their collectives
    except _queue.Empty:
        pass
    failed = len(res) < len(procs) or any(r[2] is not None for r in res)
    for p in procs:
        p.join(timeout=5 if failed else 120)
        if p.is_alive():
            p.kill()
    assert res, f"{what}: no rank reported within {timeout} s"
    for rank, out, err in res:
        assert err is None, f"rank {rank}: {err}"
        for k, (got, exp) in out.items():
            assert got == exp, f"rank {rank} {what}/{k}: {got} != {exp}"
    assert len(res) == len(procs), f"{what}: only ranks {[r[0] for r in res]} reported"
    assert all(p.exitcode == 0 for p in procs)
    return res


def _worlds():
    n = _ngpus()
    return [w for w in (2, 4) if w <= n]


@pytest.mark.parametrize("what", ["wave", "wavetrace", "scan", "batch"])
def test_multi_gpu_paths_bit_exact(what, oracle):
    worlds = _worlds()
    if not worlds:
        pytest.skip(f"needs >= 2 GPUs on the box (found {_ngpus()}); the same protocol runs with world = 1 in test_gpu_fill.py / test_gpu_big.py")
    for w in worlds:
        _run(w, what)


# ------------------------------------------------------------------ the reference's own GPU path and executable, on the box
def test_cfg2_bit_exact_with_reference_gpu9(scoring, oracle):
    """north_star: "bit-exact with the reference's own cpu4-mt-diagrow and gpu9 paths": cfg2 (16 384^2) through the UNMODIFIED gpu9
    + NwTrace2_Sparse of oracle/_ref on this GPU against the engine -- score, trace hash, transcript."""
    from gpuseqalign_b200 import Engine, synth
    if not (oracle.ref_available() and oracle.ref().nwref_has_gpu9()):
        pytest.skip("oracle/_ref/libnwref.so (with gpu9) was not prebuilt")
    subst = scoring["subst"]["blosum62"]
    eng = Engine(0)
    try:
        eng.set_scoring(subst, -11)
        x = synth.letters(2001, 16384)
        for y in (synth.letters(2002, 16384), synth.mutated_copy(x, 2003, 16384)):
            ref = oracle.ref_run("gpu9", y, x, subst, -11, want_hash=False, want_trace=True)
            score = eng.align(y, x, keep_headers=True)
            edit, th = eng.trace()
            assert score == ref.score
            assert th == ref.trace_hash
            assert edit == ref.edit
        # a ragged shape with substring-like odd sizes through both as well
        y = synth.letters(81, 1237); x = synth.letters(82, 4099)
        ref = oracle.ref_run("gpu9", y, x, subst, -11, want_hash=True, want_trace=True)
        assert eng.align(y, x, keep_headers=True) == ref.score
        assert eng.trace() == (ref.edit, ref.trace_hash)
        assert eng.score_hash() == ref.score_hash
    finally:
        eng.close()


def test_reference_benchmark_executable_accepts_the_plugin():
    """The reference's OWN `nw` executable (stock benchmark.cpp, all stock algorithms) with NwAlign_B200 registered as one more
    entry: its cross-check (benchmark.cpp:120-147) must accept all 173 pairs of resrc/pair_debug.txt (exit code 0)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "nw_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/nw_b200 was not prebuilt")
    out = os.path.join(ROOT, "gpurun_out", "test_ref_bench_pair_debug.tsv")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run([exe, "-b", "resrc/subst.json", "-r", os.path.join(ROOT, "gpuseqalign_b200", "plugin", "param_b200.json"),
                        "-s", "resrc/seq_generated.fa", "-p", "resrc/pair_debug.txt", "--fCalcTrace", "--fCalcScoreHash", "-o", out],
                       cwd=os.path.join(ROOT, "oracle", "_ref"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = [l.split("\t") for l in open(out).read().splitlines()]
    hdr = rows[0]
    b200 = [dict(zip(hdr, r_)) for r_ in rows[1:] if r_[0] == "NwAlign_B200"]
    assert len(b200) == 173
    assert all(r_["err_step"] == "0" for r_ in b200)


def test_reference_benchmark_debug_output_with_the_plugin():
    """--fPrintTrace / --fPrintScore through the reference's executable: calcDebugTrace folds the path's matrix values into
    trace_hash (nwtrace1_plain.cpp:120-126), so exit code 0 means NwTrace_B200's values equal NwTrace1_Plain's; the printed trace
    of the B200 entry equals cpu4's text byte for byte, its score matrix equals gpu9's (NwPrintScore2_Sparse, nwtrace2_sparse.cpp:346-419:
    setw(4) and a comma) byte for byte and cpu4's (NwPrintScore1_Plain: setw(4) and a blank) value for value."""
    exe = os.path.join(ROOT, "oracle", "_ref", "nw_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/nw_b200 was not prebuilt")
    outdir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(outdir, exist_ok=True)
    pairs = os.path.join(outdir, "test_pairs_debug_small.txt")
    with open(pairs, "w") as f:
        f.write("len1 len1\nlen1 len33\nlen66 len1\nlen40 len60\nlen256 len196\nlen500 len1000\n")
    out = os.path.join(outdir, "test_ref_bench_debug.tsv")
    dbg = os.path.join(outdir, "test_ref_bench_debug.txt")
    r = subprocess.run([exe, "-b", "resrc/subst.json", "-r", os.path.join(ROOT, "gpuseqalign_b200", "plugin", "param_b200.json"),
                        "-s", "resrc/seq_generated.fa", "-p", pairs, "--fCalcTrace", "--fCalcScoreHash", "--fPrintTrace", "--fPrintScore",
                        "--debugPath", dbg, "-o", out],
                       cwd=os.path.join(ROOT, "oracle", "_ref"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    blocks = [b for b in open(dbg).read().split(">results\n") if b.strip()]
    by_alg = {}
    for b in blocks:
        head, _, rest = b.partition("+\n>edit_trace\n")
        cols = head.splitlines()[0].split("\t"); vals = head.splitlines()[1].split("\t")
        row = dict(zip(cols, vals))
        trace_txt, _, score_txt = rest.partition("+\n>score_matrix\n")
        by_alg.setdefault(row["alg_name"], []).append((row["seqY_idx"], row["seqX_idx"], row["trace_hash"], trace_txt, score_txt))
    ref = by_alg["NwAlign_Cpu4_Mt_DiagRow"]; gpu9 = by_alg["NwAlign_Gpu9_Mlsp_DiagDiagDiag"]; got = by_alg["NwAlign_B200"]
    assert len(ref) == len(gpu9) == len(got) == 6
    for a, g, b in zip(ref, gpu9, got):
        assert a[:4] == b[:4]                                        # pair, trace hash (path values folded in), printed trace
        assert g == b                                                # + the printed score matrix, byte for byte
        assert a[4].replace(" \n", "\n").split() == b[4].replace(",", " ").split()
    # peak-memory columns of the B200 rows are filled in (updateNwAlgPeakMemUsage's places, nwalign_shared.cpp:16-24)
    rows = [l.split("\t") for l in open(out).read().splitlines()]
    hdr = rows[0]
    for r_ in rows[1:]:
        d = dict(zip(hdr, r_))
        if d["alg_name"] == "NwAlign_B200":
            assert int(d["glmem_peak_allocs"]) > 0 and int(d["shmem_peak_allocs"]) > 0 and int(d["regmem_peak_allocs"]) > 0
            assert float(d["align.cpy_dev"]) > 0 and float(d["align.cpy_host"]) > 0
