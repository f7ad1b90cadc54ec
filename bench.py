#!/usr/bin/env python
"""bench.py -- GCUPS of the NW linear-gap hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 20 --warmup 5                       # cfg3: the 1 048 576-pair batch (headline), + secondary blocks
    python bench.py --workload pair16k|wave200k|scan4m                   # another BASELINE config as the main line
    python bench.py --impl reference                                     # the reference's own cpu4 path on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Workloads (BASELINE.json configs); a "step" is one pass of the hot path over the whole job:
  batch256 (cfg3, DEFAULT): a FIXED job of 1 048 576 synthetic 256 x 256 pairs, scores; the pairs are sharded over the ranks
            (strong scaling, no data-path collective: pairs are independent objects);
  pair16k  (cfg2): one 16 384 x 16 384 protein pair, score-matrix fill + traceback (rank 0; a lone pair does not shard);
  wave200k (cfg5): one 200 000 x 200 000 pair; N = 1: fill + traceback on one GPU; N > 1: the column-block wavefront, border
            columns pushed GPU-to-GPU over NVLink by the fill kernel itself (peer stores, no NCCL on the data path);
  scan4m   (cfg4): 2 048 x 4 194 304, score only, row-parallel prefix-max scorer, column chunks dealt to the ranks.
The main line carries the chosen workload; the others ride along in `secondary` (fewer steps) unless --no-secondary.
`value` is whole-job GCUPS with the inputs resident in HBM (device time, CUDA events on the engine's stream, max over ranks);
`e2e` is the same through the public call with HOST buffers (H2D + kernels + D2H inside the region, wall clock between
barriers, max over ranks).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SM_COUNT = 148
DPX_PER_CLK_PER_SM = 64.0          # measured VIMNMX3 rate (profiles/microbench_r1.jsonl): 63.96 thread-ops/clk/SM
MIX_CELLS_PER_CLK_PER_SM = 48.3    # measured IDP.4A + VIMNMX3 pair rate (same file): cells/clk/SM of the 2-instruction cell
MIX16_CELLS_PER_CLK_PER_SM = 58.7  # measured rate of the packed cell mix, 0.5 PRMT + IDP.2A + VIMNMX3.U16x2 per two cells, 16 warps/SM (profiles/r2r_microbench_merge_variants.jsonl)
ISSUE_PER_CLK_PER_SM = 128.0       # 4 schedulers x 32 lanes: the hard ceiling of thread-instructions per clock

WORKLOADS = ("batch256", "pair16k", "wave200k", "scan4m")


def load_scoring():
    with open(os.path.join(ROOT, "tests", "golden", "scoring.json")) as f:
        s = json.load(f)
    return np.array(s["subst"]["blosum62"], dtype=np.int32), -11


def load_json(*parts):
    try:
        with open(os.path.join(ROOT, *parts)) as f:
            return json.load(f)
    except Exception:
        return {}


# --------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed region runs (NVML)."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index: int, period_s: float = 0.01):
        self.index, self.period = index, period_s
        self.samples, self.mask, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self._nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        reasons = [n for b, n in self.REASONS.items() if (self.mask & b) and n != "gpu_idle"]
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------------- workload descriptions
def batch_config(total_pairs, world):
    return {"workload": f"cfg3: batch of {total_pairs} synthetic 256x256 pairs, scores only, a fixed job sharded {total_pairs // max(1, world)} pairs/GPU",
            "pairs": total_pairs, "pairs_per_gpu": total_pairs // max(1, world), "seeds": "pair p: X 3e6+2p, Y 3e6+2p+1 (splitmix64)",
            "subst": "blosum62", "gap": -11,
            "l2": "inputs (512 B/pair, 537 MB for the job) exceed L2; flushed between timed steps as well (256 MiB memset)"}


def pair_config(n, m, with_trace=True):
    return {"workload": f"cfg2: single synthetic protein pair {n}x{m}, score" + ("+traceback" if with_trace else " only"),
            "pairs_per_gpu": 1, "seeds": "X 2001, Y 2002 (splitmix64, independent)", "subst": "blosum62", "gap": -11,
            "l2": "flushed between timed steps (256 MiB memset)"}


def wave_config(n, world):
    how = "one GPU: fill + sparse traceback recompute" if world == 1 else \
        f"column-block wavefront over {world} GPUs, border columns as peer stores over NVLink; fill on every GPU + traceback on rank 0 from the headers it pulls over NVLink"
    return {"workload": f"cfg5: single long pair {n}x{n}, {how}", "seeds": "X 5001, Y 5004 (splitmix64, independent)",
            "subst": "blosum62", "gap": -11, "l2": "flushed between timed steps (256 MiB memset)"}


def scan_config(n, m, world):
    return {"workload": f"cfg4: rectangular pair {n}x{m}, score only, row-parallel prefix-max scorer, column chunks dealt to {world} GPU(s)",
            "seeds": "Y 4001, X 4002 (splitmix64)", "subst": "blosum62", "gap": -11, "l2": "flushed between timed steps (256 MiB memset)"}


# --------------------------------------------------------------------------------- reference arm (CPU)
def host_threads():
    """The threads the CPU arm may use: every core this process is allowed on (torchrun exports OMP_NUM_THREADS=1 to its ranks;
    the reference arm is ONE process that owns the host while the other ranks have left)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_pair(y, x, subst, gap, samples, warmup, with_trace=True):
    """cfg2/cfg5-style single pair on the host: the reference's own cpu4-mt-diagrow (+ NwTrace1_Plain) from oracle/_ref when it
    was prebuilt, else the C restatement.  Returns (gcups, kind, cores, ms)."""
    from oracle import pyoracle
    cores = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(cores)
    cells = float(y.size) * float(x.size)
    times = []
    if pyoracle.ref_available():
        kind = "reference"
        for it in range(warmup + samples):
            r = pyoracle.ref_run("cpu4", y, x, subst, gap, want_hash=False, want_trace=with_trace)
            ms = r.laps_ms["align_calc"] + (r.laps_ms["trace_calc"] if with_trace else 0.0)
            if it >= warmup:
                times.append(ms)
    else:
        kind = "port"
        ensure_oracle()
        for it in range(warmup + samples):
            t0 = time.perf_counter()
            pyoracle.align_pair(y, x, subst, gap, want_hash=False, want_trace=with_trace, threads=cores)
            if it >= warmup:
                times.append((time.perf_counter() - t0) * 1e3)
    ms = float(np.mean(times))
    return cells / ms / 1e6, kind, cores, ms


def ensure_oracle():
    from oracle import pyoracle
    if not os.path.exists(pyoracle.ORACLE_SO):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])


def cpu_batch(pool, offY, lenY, offX, lenX, subst, gap, passes, warmup, budget_s):
    """cfg3 on the host.  The reference's own cpu4, called once per pair as benchmark.cpp:406 does, the pairs dealt to every host
    thread by the shim (oracle/ref_shim.cpp nwref_batch_cpu; a 256 x 256 pair is a single cpu4 tile); the
    sample is sized so that all passes fit `budget_s`.  Falls back to the oracle port when oracle/_ref is absent.
    Returns dict(value, kind, cores, ms, sample_pairs, port_gcups, serial_gcups)."""
    from oracle import pyoracle
    ensure_oracle()
    cores = host_threads()
    n = len(lenY)
    cells_per_pair = float(np.mean(lenY.astype(np.float64) * lenX.astype(np.float64)))
    use_ref = pyoracle.ref_available() and hasattr(pyoracle.ref(), "nwref_batch_cpu")

    def run(k, threads=cores):
        if use_ref:
            _, ms = pyoracle.ref_batch_cpu("cpu4", pool, offY[:k], lenY[:k], offX[:k], lenX[:k], subst, gap, threads=threads)
            return ms
        t0 = time.perf_counter()
        pyoracle.score_batch(pool, offY[:k], lenY[:k], offX[:k], lenX[:k], subst, gap, threads=threads)
        return (time.perf_counter() - t0) * 1e3

    k0 = min(n, 4096)
    run(min(n, 512))
    ms0 = run(k0)
    per_pass_s = budget_s / max(1, passes + warmup)
    k = int(min(n, max(k0, k0 * per_pass_s * 1e3 / max(ms0, 1e-3))))
    times = []
    for it in range(warmup + passes):
        ms = run(k)
        if it >= warmup:
            times.append(ms)
    ms = float(np.mean(times))
    out = {"value": k * cells_per_pair / ms / 1e6, "kind": "reference" if use_ref else "port", "cores": cores, "ms": ms, "sample_pairs": k}
    # beside it: the oracle port's rolling-row scorer (OpenMP over pairs) on the same sample, and the reference's strictly serial loop
    t0 = time.perf_counter()
    pyoracle.score_batch(pool, offY[:k], lenY[:k], offX[:k], lenX[:k], subst, gap, threads=cores)
    out["port_gcups"] = k * cells_per_pair / (time.perf_counter() - t0) / 1e9
    if use_ref:
        ks = min(k, 256)
        _, ms1 = pyoracle.ref_batch_cpu("cpu4", pool, offY[:ks], lenY[:ks], offX[:ks], lenX[:ks], subst, gap, threads=1)
        out["serial_gcups"] = ks * cells_per_pair / ms1 / 1e6
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from gpuseqalign_b200 import synth
    subst, gap = load_scoring()
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    world = max(1, args.gpus)
    extra = {}
    if args.workload == "batch256":
        npairs = min(args.pairs, 1 << 18)          # the sample is cut from a prefix of the job
        pool, offY, lenY, offX, lenX = synth.batch_pairs(0, npairs, 256, 256)
        r = cpu_batch(pool, offY, lenY, offX, lenX, subst, gap, steps, warmup, budget_s=100.0)
        g, kind, cores, ms = r["value"], r["kind"], r["cores"], r["ms"]
        what = "cpu4-mt-diagrow blocksz 256 called once per pair (benchmark.cpp:406), pairs dealt to every host thread" if kind == "reference" \
            else "oracle port, rolling-row scorer, OpenMP over pairs"
        sample = f"{steps} x the first {r['sample_pairs']} pairs of the job (256x256, scores only; {what})"
        cfg = batch_config(args.pairs, world)
        cfg["reference_sample_pairs"] = r["sample_pairs"]
        extra = {"port_gcups": r.get("port_gcups"), "serial_gcups": r.get("serial_gcups")}
    elif args.workload == "pair16k":
        n = m = args.len
        x = synth.letters(2001, m); y = synth.letters(2002, n)
        g, kind, cores, ms = cpu_pair(y, x, subst, gap, steps, min(warmup, 2))
        sample = f"{steps} x the full {n}x{m} pair (cpu4-mt-diagrow blocksz 256 fill + NwTrace1_Plain traceback)"
        cfg = pair_config(n, m)
    else:
        # cfg4 / cfg5 cannot run on the reference's cpu4 (int overflow of adjrows*adjcols, 34 / 160 GB: SURVEY.md App. D-1):
        # the largest square it can hold is timed as a proxy and labelled as such
        n = m = 32768
        x = synth.letters(5001, m); y = synth.letters(5004, n)
        steps = min(steps, 5)
        g, kind, cores, ms = cpu_pair(y, x, subst, gap, steps, 1, with_trace=(args.workload == "wave200k"))
        sample = f"PROXY: {steps} x a {n}x{m} pair (the largest square cpu4 can hold; the config itself overflows its int matrix size)"
        cfg = wave_config(200000, world) if args.workload == "wave200k" else scan_config(2048, 4194304, world)
    line = {"impl": "reference", "metric": "GCUPS NW linear-gap", "value": g, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if args.workload != "pair16k" else "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": cfg,
            "cpu_baseline": dict({"value": g, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": sample}, **extra),
            "e2e": {"value": g, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------- this engine
class Ctx:
    """Everything a workload needs: the engine, its stream, the process group helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from gpuseqalign_b200 import Engine
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.subst, self.gap = load_scoring()
        self.eng = Engine(self.local)
        self.eng.set_scoring(self.subst, self.gap)
        self.stream = torch.cuda.ExternalStream(self.eng.stream_ptr(), device=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)      # > 126 MB L2
        self.epoch = 1000

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, v: float, op: str = "max") -> float:
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN}[op])
        return float(t.item())

    def timed_async(self, step, steps, warmup, clocks=False):
        """Device time of `steps` asynchronous steps on the engine's stream (CUDA events, L2 flushed before every step).
        Returns (sum of the per-step times in ms on THIS rank, launches, clock summary or None)."""
        torch = self.torch
        for _ in range(warmup):
            step()
        self.eng.sync()
        l0 = self.eng.launches()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        sampler = ClockSampler(self.local) if clocks else None
        if sampler:
            sampler.__enter__()
        with torch.cuda.stream(self.stream):
            for a, b in ev:
                self.flush.fill_(1)
                a.record(self.stream)
                step()
                b.record(self.stream)
        self.eng.sync()
        if sampler:
            sampler.__exit__()
        self.barrier()
        return sum(a.elapsed_time(b) for a, b in ev), self.eng.launches() - l0, (sampler.summary() if sampler else None)

    def timed_wall(self, step, steps, warmup):
        """Wall time of `steps` synchronous public calls between barriers, max over ranks, in seconds; returns (s, last result)."""
        for _ in range(warmup):
            step()
        self.barrier()
        t0 = time.perf_counter()
        r = None
        for _ in range(steps):
            r = step()
        self.eng.sync()
        self.barrier()
        return self.reduce(time.perf_counter() - t0), r


def bind_to_gpu_numa_node(torch, local):
    """One process per GPU: run (and therefore first-touch / pin host buffers) on the CPUs next to this rank's GPU, so that H2D
    slices do not cross the socket interconnect.  Returns the cpulist used, or None when sysfs does not say."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return txt
    except Exception:
        return None


def wl_batch(cx: Ctx, steps, warmup, main):
    """cfg3: the fixed job of args.pairs pairs, sharded contiguously over the ranks."""
    from gpuseqalign_b200 import synth
    torch, eng, args = cx.torch, cx.eng, cx.args
    total = args.pairs
    lo, hi = cx.rank * total // cx.world, (cx.rank + 1) * total // cx.world
    per = hi - lo
    numa = bind_to_gpu_numa_node(torch, cx.local) if cx.world > 1 else None
    pool, offY, lenY, offX, lenX = synth.batch_pairs(lo, per, 256, 256)
    pool = torch.from_numpy(pool).pin_memory().numpy()          # e2e: H2D from pinned host memory
    scores_pinned = torch.empty(per, dtype=torch.int32).pin_memory().numpy()
    cells_rank = float(per) * 256.0 * 256.0
    cells_job = float(total) * 256.0 * 256.0
    eng.upload_batch(pool, offY, lenY, offX, lenX)
    ms_rank, launches, clocks = cx.timed_async(eng.batch_resident, steps, warmup, clocks=main)
    ms_total = cx.reduce(ms_rank)
    scores = eng.fetch_batch_scores()
    kernel_name = eng.batch_kernel_name() if hasattr(eng, "batch_kernel_name") else "nw_batch2_kernel"
    # parity at full size: checksums of ALL scores against the oracle's (tests/golden/batch_golden.json)
    gold = load_json("tests", "golden", "batch_golden.json")
    parity = None
    if gold.get("pairs") == total and (8 % cx.world) == 0 and total % 8 == 0:
        e8 = total // 8
        ok = all(hashlib.sha256(scores[k * e8 - lo:(k + 1) * e8 - lo].tobytes()).hexdigest() == gold["sha256_eighths"][k]
                 for k in range(lo // e8, hi // e8))
        ok = cx.reduce(1.0 if ok else 0.0, "min") == 1.0
        parity = {"scores_sha256_match_oracle": bool(ok), "golden": "tests/golden/batch_golden.json"}
    # end to end: host buffers in, host scores out
    def step_e2e():
        eng.align_batch(pool, offY, lenY, offX, lenX, out=scores_pinned)
        return 4 * per
    e2e_s, d2h = cx.timed_wall(step_e2e, steps, 2)
    if parity is not None:
        parity["e2e_scores_equal_resident"] = bool(cx.reduce(1.0 if np.array_equal(scores_pinned, scores) else 0.0, "min") == 1.0)
    h2d = pool.size + 24 * per
    # the same call with 5-bit packed letters (nwb200_align_batch_packed5: 8 letters in 5 bytes, packed once, outside the timed region --
    # the host keeps its sequences in that form): 3/8 less H2D, which is what bounds the byte-letter call above
    packed5 = None
    try:
        offs = np.empty(2 * per, dtype=np.int64); lens = np.full(2 * per, 256, dtype=np.int64)
        offs[0::2] = offX; offs[1::2] = offY
        pk, noffs = synth.pack5(pool, offs, lens)
        pk = torch.from_numpy(pk).pin_memory().numpy()
        pX, pY = noffs[0::2].copy(), noffs[1::2].copy()
        scores_pk = torch.empty(per, dtype=torch.int32).pin_memory().numpy()

        def step_pk():
            eng.align_batch_packed5(pk, pY, lenY, pX, lenX, out=scores_pk)
            return 4 * per
        pk_s, _ = cx.timed_wall(step_pk, steps, 2)
        h2d_pk = pk.size + 24 * per
        packed5 = {"value": cells_job * steps / pk_s / 1e9, "unit": "GCUPS", "ms_per_step": pk_s / steps * 1e3,
                   "h2d_bytes_per_step": int(cx.reduce(h2d_pk, "sum")), "h2d_gbs_per_rank": h2d_pk * steps / pk_s / 1e9,
                   "scores_equal_resident": bool(cx.reduce(1.0 if np.array_equal(scores_pk, scores) else 0.0, "min") == 1.0),
                   "note": "nwb200_align_batch_packed5: letters 5-bit packed on the host ahead of time (8 letters in 5 bytes)"}
        eng.upload_batch(pool, offY, lenY, offX, lenX)       # (the resident byte-letter batch again, for whoever runs after this)
    except Exception as ex:                                     # reported, never fatal for the main line
        packed5 = {"error": f"{type(ex).__name__}: {ex}"}
    cfg = batch_config(total, cx.world)
    if numa:
        cfg["cpu_affinity"] = f"rank bound to the CPUs next to its GPU ({numa})"
    h2d_gbs_rank = h2d * steps / e2e_s / 1e9
    return {"value": cells_job * steps / ms_total / 1e6, "ms_per_step": ms_total / steps, "cells_job": cells_job, "cells_rank": cells_rank,
            "kernel_ms": ms_rank / steps, "kernel": kernel_name, "launches": launches, "clocks": clocks, "config": cfg, "scaling": "strong",
            "e2e": {"value": cells_job * steps / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(cx.reduce(h2d, "sum")),
                    "d2h_bytes_per_step": int(cx.reduce(d2h, "sum")), "ms_per_step": e2e_s / steps * 1e3,
                    "h2d_gbs_per_rank": h2d_gbs_rank, "h2d_gbs_all_ranks": cx.reduce(h2d_gbs_rank, "sum"), "packed5": packed5},
            "parity": parity, "dtype": "u16x2 (two pairs per 32-bit register)" if "batch2" in kernel_name else "int32",
            "host": (pool, offY, lenY, offX, lenX)}


def wl_pair(cx: Ctx, steps, warmup, main):
    """cfg2: one pair; does not shard -- as the main workload every rank aligns its own copy (replicas), as a secondary block
    rank 0 alone runs it."""
    from gpuseqalign_b200 import synth
    eng, args = cx.eng, cx.args
    n = m = args.len
    x = synth.letters(2001, m); y = synth.letters(2002, n)
    if not main and cx.rank != 0:
        return None
    with_trace = not args.no_trace
    cells = float(n) * float(m)
    eng.upload_pair(y, x)

    def step():
        eng.fill_resident(True)
        if with_trace:
            eng.trace_resident()

    if main:
        ms_rank, launches, clocks = cx.timed_async(step, steps, warmup, clocks=True)
        ms_total = cx.reduce(ms_rank)
        cells_job = cells * cx.world
    else:
        ms_rank, launches, clocks = _timed_async_local(cx, step, steps, warmup)
        ms_total, cells_job = ms_rank, cells
    score = eng.fetch_score()
    edit, th = eng.fetch_trace() if with_trace else ("", 0)
    lap = eng.timing()
    gold = load_json("tests", "golden", "big_golden.json").get("cfg2_random", {})
    parity = None
    if n == 16384 and gold:
        parity = {"score_matches_oracle": score == gold["score"]}
        if with_trace:
            parity["trace_hash_matches_oracle"] = f"{th:08x}" == gold["trace_hash"]
            parity["edit_sha256_matches_oracle"] = hashlib.sha256(edit.encode()).hexdigest() == gold["edit_sha256"]

    def step_e2e():
        eng.align(y, x, keep_headers=True, with_trace=with_trace)
        if with_trace:
            e, _ = eng.trace()
            return 4 + len(e) + 4
        return 4

    if main:
        e2e_s, d2h = cx.timed_wall(step_e2e, steps, 2)
    else:
        for _ in range(2):
            step_e2e()
        t0 = time.perf_counter()
        for _ in range(steps):
            d2h = step_e2e()
        e2e_s = time.perf_counter() - t0
    fill_ms = lap.get("align_calc", 0.0)
    f_ghz = 1.965
    R_, K_, grp = 4, 2, 8
    nb_ = (n + 32 * R_ - 1) // (32 * R_)
    dep_steps = m + 31 * K_ + (nb_ - 1) * (31 * K_ + grp)
    floor_ms = dep_steps * R_ * 4.47 / (f_ghz * 1e6)
    return {"value": cells_job * steps / ms_total / 1e6, "ms_per_step": ms_total / steps, "cells_job": cells_job, "cells_rank": cells,
            "kernel_ms": fill_ms, "kernel": "nw_fill_kernel", "launches": launches, "clocks": clocks, "config": pair_config(n, m, with_trace),
            "scaling": "weak", "laps_ms_last_step": {k: round(v, 4) for k, v in lap.items()},
            "e2e": {"value": cells_job * steps / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(n + m), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s / steps * 1e3},
            "latency_floor": {"dependent_steps": dep_steps, "clk_per_step_floor": R_ * 4.47, "ms": floor_ms, "frac": floor_ms / fill_ms if fill_ms else None,
                              "note": "critical path of the band wavefront at the DPX dependent-issue latency; the fill kernel's time against it"},
            "parity": parity, "dtype": "int32", "host": (y, x)}


def _timed_async_local(cx, step, steps, warmup):
    """timed_async without barriers (a block that one rank runs alone)."""
    torch, eng = cx.torch, cx.eng
    for _ in range(warmup):
        step()
    eng.sync()
    l0 = eng.launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.cuda.stream(cx.stream):
        for a, b in ev:
            cx.flush.fill_(1)
            a.record(cx.stream)
            step()
            b.record(cx.stream)
    eng.sync()
    return sum(a.elapsed_time(b) for a, b in ev), eng.launches() - l0, None


def _collective_steps(cx, fill, fetch, steps, warmup):
    """Steps of a cross-GPU kernel: every step is one cooperative launch per rank with a fresh epoch, bracketed by barriers (a rank
    must not push epoch e+1 into a neighbour that still reads epoch e).  Returns (device ms summed over the steps, max over ranks
    per step; wall s summed the same way; last score; launches)."""
    eng = cx.eng
    dev_ms, wall_s, score = 0.0, 0.0, None
    l0 = None
    for it in range(warmup + steps):
        if it == warmup:
            l0 = eng.launches()
        cx.epoch += 1
        cx.flush.fill_(1)
        cx.barrier()
        t0 = time.perf_counter()
        fill(cx.epoch)
        s = fetch()
        dt = time.perf_counter() - t0
        if s is not None:
            score = s
        d = cx.reduce(eng.timing()["align_calc"])
        w = cx.reduce(dt)
        if it >= warmup:
            dev_ms += d
            wall_s += w
    cx.barrier()
    return dev_ms, wall_s, score, eng.launches() - (l0 or 0)


def wl_wave(cx: Ctx, steps, warmup, main):
    """cfg5: N = 1: the single-pair engine (fill + traceback); N > 1: the column-block wavefront (score)."""
    from gpuseqalign_b200 import synth
    from gpuseqalign_b200.wavefront import wave_align, wave_setup
    eng, args = cx.eng, cx.args
    n = args.wave_len
    x = synth.letters(5001, n); y = synth.letters(5004, n)
    cells = float(n) * float(n)
    gold = load_json("tests", "golden", "big_golden.json").get("cfg5_random", {}) if n == 200000 else {}
    parity = None
    if cx.world == 1:
        with_trace = not args.no_trace
        eng.upload_pair(y, x)

        def step():
            eng.fill_resident(True)
            if with_trace:
                eng.trace_resident()

        ms_rank, launches, clocks = cx.timed_async(step, steps, warmup, clocks=main)
        score = eng.fetch_score()
        edit, th = eng.fetch_trace(1 << 20) if with_trace else ("", 0)
        lap = eng.timing()
        tinfo = eng.trace_info() if with_trace else None      # corridor maps: segments computed per band, and whether the path left the corridor
        if gold:
            parity = {"score_matches_oracle": score == gold["score"]}
            if with_trace:
                parity["trace_hash_matches_oracle"] = f"{th:08x}" == gold["trace_hash"]
                parity["edit_sha256_matches_oracle"] = hashlib.sha256(edit.encode()).hexdigest() == gold["edit_sha256"]

        def step_e2e():
            eng.align(y, x, keep_headers=True, with_trace=with_trace)
            if with_trace:
                e, _ = eng.trace(1 << 20)
                return 8 + len(e)
            return 4

        e2e_s, d2h = cx.timed_wall(step_e2e, max(1, min(steps, 3)), 1)
        e2e_steps = max(1, min(steps, 3))
        return {"value": cells * steps / ms_rank / 1e6, "ms_per_step": ms_rank / steps, "cells_job": cells, "cells_rank": cells,
                "kernel_ms": lap.get("align_calc", 0.0), "kernel": "nw_fill_kernel", "launches": launches, "clocks": clocks,
                "config": wave_config(n, 1), "scaling": "strong",
                "laps_ms_last_step": {k: round(v, 4) for k, v in lap.items()}, "traceback": tinfo,
                "e2e": {"value": cells * e2e_steps / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(2 * n), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": e2e_s / e2e_steps * 1e3},
                "parity": parity, "dtype": "int32"}
    with_trace = not args.no_trace
    per_rank = (n + cx.world - 1) // cx.world
    block = args.wave_block if args.wave_block > 0 else max(2048, (per_rank + 511) // 512 * 512)     # (a multiple of the snapshot spacing: the traceback needs it)
    if with_trace:
        from gpuseqalign_b200.wavefront import wave_trace_setup, wave_trace
        wave_trace_setup(eng, y, x, rank=cx.rank, world=cx.world, block_cols=block)
    else:
        wave_setup(eng, y, x, rank=cx.rank, world=cx.world, block_cols=block)
    sampler = ClockSampler(cx.local) if main else None
    if sampler:
        sampler.__enter__()
    # every step: one cooperative fill launch per rank with a fresh epoch (device time = max over ranks of the launch), then -- with the
    # traceback -- the pull of the headers and the traceback on rank 0 (wall time between the barriers that bracket it, max over ranks)
    trace_state = {"walls": [], "last": None}

    def fetch():
        s_ = eng.wave_fetch()
        if with_trace:
            t0 = time.perf_counter()
            tr = wave_trace(eng, rank=cx.rank, world=cx.world, cap=1 << 21)
            trace_state["walls"].append(time.perf_counter() - t0)
            if tr is not None:
                trace_state["last"] = tr
        return s_

    dev_ms, wall_s, score, launches = _collective_steps(cx, eng.wave_fill, fetch, steps, warmup)
    if sampler:
        sampler.__exit__()
    score = int(cx.reduce(float(score) if score is not None else -2.0 ** 62))
    trace_ms = 0.0
    tinfo = None
    if with_trace:
        timed = trace_state["walls"][-steps:]                  # (the warm-up steps carry the first peer mapping)
        trace_ms = cx.reduce(sum(timed) / max(1, len(timed)) * 1e3)
    if gold:
        parity = {"score_matches_oracle": score == gold["score"]}
        if with_trace and cx.rank == 0 and trace_state["last"] is not None:
            edit, th, tinfo = trace_state["last"]
            parity["trace_hash_matches_oracle"] = f"{th:08x}" == gold["trace_hash"]
            parity["edit_sha256_matches_oracle"] = hashlib.sha256(edit.encode()).hexdigest() == gold["edit_sha256"]
    # end to end: the whole public call (upload on every rank, handle exchange, fill, score, traceback)
    e2e_steps = max(1, min(steps, 3))
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        cx.epoch += 1
        if with_trace:
            wave_trace_setup(eng, y, x, rank=cx.rank, world=cx.world, block_cols=block)
            eng.wave_fill(cx.epoch)
            eng.wave_fetch()
            wave_trace(eng, rank=cx.rank, world=cx.world, cap=1 << 21)
        else:
            wave_align(eng, y, x, rank=cx.rank, world=cx.world, block_cols=block, epoch=cx.epoch)
    cx.barrier()
    e2e_s = cx.reduce(time.perf_counter() - t0)
    # with the traceback a step is what the wall clock sees between the barrier in front of the fills and the end of the traceback (max over
    # ranks): rank 0's wait for the last rank's fill is inside `gather_and_trace_wall_rank0`, so the laps do not add up
    step_ms = wall_s / steps * 1e3 if with_trace else dev_ms / steps
    trace_dev_ms = cx.reduce(eng.timing().get("trace_calc", 0.0) if (with_trace and cx.rank == 0) else 0.0)
    return {"value": cells / step_ms / 1e6, "ms_per_step": step_ms, "cells_job": cells, "cells_rank": cells / cx.world,
            "kernel_ms": dev_ms / steps, "kernel": "nw_fill_kernel (column-block wavefront)", "launches": launches,
            "clocks": sampler.summary() if sampler else None, "config": dict(wave_config(n, cx.world), block_cols=block), "scaling": "strong",
            "wall_ms_per_step": wall_s / steps * 1e3,
            "laps_ms_last_step": {"align_calc_max_over_ranks": round(dev_ms / steps, 4), "trace_calc_rank0": round(trace_dev_ms, 4),
                                  "gather_and_trace_wall_rank0": round(trace_ms, 4)},
            "traceback": tinfo,
            "e2e": {"value": cells * e2e_steps / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(2 * n * cx.world), "d2h_bytes_per_step": 4 + (n * 2 // 3 if with_trace else 0),
                    "ms_per_step": e2e_s / e2e_steps * 1e3},
            "parity": parity, "dtype": "int32"}


def wl_scan(cx: Ctx, steps, warmup, main):
    """cfg4: the row-parallel prefix-max scorer, column chunks dealt to the ranks in contiguous ranges."""
    from gpuseqalign_b200 import synth
    from gpuseqalign_b200.wavefront import scan_align, scan_setup
    eng, args = cx.eng, cx.args
    n, m = args.scan_rows, args.scan_cols
    y = synth.letters(4001, n); x = synth.letters(4002, m)
    cells = float(n) * float(m)
    gold = load_json("tests", "golden", "big_golden.json").get("cfg4", {}) if (n, m) == (2048, 4194304) else {}
    scan_setup(eng, y, x, rank=cx.rank, world=cx.world)
    sampler = ClockSampler(cx.local) if main else None
    if sampler:
        sampler.__enter__()
    dev_ms, wall_s, score, launches = _collective_steps(cx, eng.scan_fill, eng.scan_fetch, steps, warmup)
    if sampler:
        sampler.__exit__()
    score = int(cx.reduce(float(score) if score is not None else -2.0 ** 62))
    parity = {"score_matches_oracle": score == gold["score"]} if gold else None
    e2e_steps = max(1, min(steps, 3))
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        cx.epoch += 1
        scan_align(eng, y, x, rank=cx.rank, world=cx.world, epoch=cx.epoch)
    cx.barrier()
    e2e_s = cx.reduce(time.perf_counter() - t0)
    return {"value": cells * steps / dev_ms / 1e6, "ms_per_step": dev_ms / steps, "cells_job": cells, "cells_rank": cells / cx.world,
            "kernel_ms": dev_ms / steps, "kernel": "nw_scan_kernel", "launches": launches, "clocks": sampler.summary() if sampler else None,
            "config": scan_config(n, m, cx.world), "scaling": "strong", "wall_ms_per_step": wall_s / steps * 1e3,
            "e2e": {"value": cells * e2e_steps / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int((n + m) * cx.world), "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_s / e2e_steps * 1e3},
            "parity": parity, "dtype": "int32"}


RUNNERS = {"batch256": wl_batch, "pair16k": wl_pair, "wave200k": wl_wave, "scan4m": wl_scan}


def gpu_reference_block(cx, subst, gap):
    """The reference's own gpu9 (nwalign_gpu9_mlsp_diagdiagdiag.cu, recompiled unmodified for sm_100a in oracle/_ref) on this box,
    N = 1: its Stopwatch laps for the cfg2 pair and for a sample of cfg3 pairs called one by one as benchmark.cpp:406 does."""
    from oracle import pyoracle
    from gpuseqalign_b200 import synth
    if not (pyoracle.ref_available() and pyoracle.ref().nwref_has_gpu9()):
        return {"unavailable": "oracle/_ref/libnwref.so (with gpu9) was not prebuilt"}
    out = {"impl": "reference gpu9-mlsp-diagdiagdiag, param_best.json (128, 4, 4, 48), unmodified, -arch sm_100a -maxrregcount 32"}
    x = synth.letters(2001, 16384); y = synth.letters(2002, 16384)
    runs = []
    for it in range(4):
        t0 = time.perf_counter()
        r = pyoracle.ref_run("gpu9", y, x, subst, gap, want_hash=False, want_trace=True)
        wall = (time.perf_counter() - t0) * 1e3
        if it:
            runs.append((r.laps_ms, wall))
    laps = {k: float(np.mean([l[k] for l, _ in runs])) for k in runs[0][0]}
    align_all = sum(v for k, v in laps.items() if k.startswith("align"))
    cells = 16384.0 * 16384.0
    out["pair16k"] = {"laps_ms": {k: round(v, 4) for k, v in laps.items()}, "align_total_ms": align_all, "trace_calc_ms": laps["trace_calc"],
                      "gcups_align_calc": cells / laps["align_calc"] / 1e6, "gcups_align_plus_trace": cells / (align_all + laps["trace_calc"] + laps["trace_alloc"]) / 1e6,
                      "wall_ms": float(np.mean([w for _, w in runs]))}
    k = 256
    pool, offY, lenY, offX, lenX = synth.batch_pairs(0, k, 256, 256)
    calc = tot = 0.0
    for it in range(k + 8):
        p = it % k
        r = pyoracle.ref_run("gpu9", pool[int(offY[p]):int(offY[p]) + 256], pool[int(offX[p]):int(offX[p]) + 256], subst, gap,
                             want_hash=False, want_trace=False)
        if it >= 8:
            calc += r.laps_ms["align_calc"]
            tot += sum(v for kk, v in r.laps_ms.items() if kk.startswith("align"))
    out["batch256"] = {"sample_pairs": k, "align_calc_ms_per_pair": calc / k, "align_total_ms_per_pair": tot / k,
                       "gcups_align_calc": 65536.0 * k / calc / 1e6, "gcups_align_total": 65536.0 * k / tot / 1e6,
                       "note": "one align call per pair (benchmark.cpp:406); the reference has no batch path"}
    return out


def compact(r):
    """A workload's result as a secondary block."""
    if r is None:
        return None
    keep = ("value", "ms_per_step", "kernel", "kernel_ms", "launches", "scaling", "laps_ms_last_step", "traceback", "latency_floor", "parity", "wall_ms_per_step")
    out = {k: r[k] for k in keep if k in r and r[k] is not None}
    out["unit"] = "GCUPS"
    out["workload"] = r["config"]["workload"]
    out["e2e"] = r["e2e"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="batch256", choices=list(WORKLOADS))
    ap.add_argument("--len", type=int, default=16384, help="pair16k: sequence length")
    ap.add_argument("--pairs", type=int, default=1 << 20, help="batch256: pairs in the whole job")
    ap.add_argument("--wave-len", type=int, default=200000, help="wave200k: sequence length")
    ap.add_argument("--wave-block", type=int, default=0, help="wave200k, N > 1: columns per block (0: one block per rank, the measured optimum -- "
                                                                "profiles/r2k_wave_probe_blocks.jsonl; every block boundary re-staggers the band wavefront)")
    ap.add_argument("--scan-rows", type=int, default=2048)
    ap.add_argument("--scan-cols", type=int, default=4194304)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="only the main workload")
    ap.add_argument("--secondary", default="pair16k,wave200k,scan4m,batch256", help="comma list of the blocks that ride along")
    ap.add_argument("--no-trace", action="store_true", help="pair16k / wave200k: fill only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    cx = Ctx(args)
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    res = RUNNERS[args.workload](cx, steps, warmup, True)

    secondary = {}
    if not args.no_secondary:
        for name in [s for s in args.secondary.split(",") if s in WORKLOADS and s != args.workload]:
            try:
                secondary[name] = compact(RUNNERS[name](cx, min(steps, 5), 3, False))
            except Exception as ex:        # a secondary block never takes the main line down; the failure is reported as such
                secondary[name] = {"failed": f"{type(ex).__name__}: {ex}"}
                if cx.world > 1:           # ranks may have left a collective step at different points: stop riding along
                    break

    if cx.rank != 0:
        if cx.world > 1:
            cx.dist.destroy_process_group()
        return 0

    clocks = res["clocks"] or {}
    peaks = load_json("MEASURED_PEAKS.json")
    f_ghz = (clocks.get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0) / 1e3
    peak = SM_COUNT * DPX_PER_CLK_PER_SM * f_ghz            # GCUPS at 1 DPX op (VIMNMX3) per cell
    kernel_ms = res["kernel_ms"]
    achieved = res["cells_rank"] / kernel_ms / 1e6 if kernel_ms else 0.0
    traffic = load_json("profiles", "traffic.json").get(args.workload, {})
    line = {"metric": "GCUPS NW linear-gap", "value": res["value"], "unit": "GCUPS", "n_gpus": cx.world, "steps": steps, "warmup": warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": res["scaling"], "vs_baseline": None,
            "dtype": res["dtype"], "data": "synthetic", "config": res["config"],
            "e2e": res["e2e"], "gpu_launches": int(res["launches"]), "clocks": clocks,
            "roofline": {"bound": "int-issue (DPX VIMNMX3, 1 per cell; not hbm/tensor)", "achieved": achieved, "peak": peak, "unit": "GCUPS",
                         "frac": achieved / peak, "kernel": res["kernel"], "kernel_ms": kernel_ms,
                         # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel at N = 1, from the ncu --set full capture named
                         # beside it (profiles/traffic.json is rewritten with every capture; null when no capture of the current kernel exists)
                         "traffic": (traffic.get("bytes_per_launch") or (traffic.get("bytes_per_pair", 0) * res["config"].get("pairs_per_gpu", 0)) or None),
                         "traffic_source": traffic.get("source"),
                         "peak_source": f"{SM_COUNT} SMs x {DPX_PER_CLK_PER_SM:.0f} VIMNMX3/clk/SM (measured, profiles/microbench_r1.jsonl) x {f_ghz:.3f} GHz",
                         "peak_mix_measured": SM_COUNT * MIX_CELLS_PER_CLK_PER_SM * f_ghz,
                         "peak_issue_slots": SM_COUNT * ISSUE_PER_CLK_PER_SM * f_ghz,
                         "rank": "rank 0's kernel time and cells" if cx.world > 1 else "the one GPU"}}
    if "batch2" in str(res["kernel"]) or "batch3" in str(res["kernel"]):
        # the packed kernel computes TWO cells per VIMNMX3.U16x2: against a roofline of one DPX op per two cells the same rate is half
        # the fraction; both are reported, together with the measured rate of the kernel's own instruction mix
        line["roofline"]["peak_packed16"] = 2.0 * peak
        line["roofline"]["frac_packed16"] = achieved / (2.0 * peak)
        line["roofline"]["peak_mix_measured"] = SM_COUNT * MIX16_CELLS_PER_CLK_PER_SM * f_ghz
        line["roofline"]["frac_of_mix_measured"] = achieved / (SM_COUNT * MIX16_CELLS_PER_CLK_PER_SM * f_ghz)
    for k in ("laps_ms_last_step", "traceback", "latency_floor", "wall_ms_per_step"):
        if res.get(k) is not None:
            line["roofline"][k] = res[k]
    if res.get("parity") is not None:
        line["parity"] = res["parity"]
    if secondary:
        line["secondary"] = secondary
    if cx.world == 1 and not args.no_cpu_baseline:
        try:
            if args.workload == "batch256":
                pool, offY, lenY, offX, lenX = res["host"]
                r = cpu_batch(pool, offY, lenY, offX, lenX, cx.subst, cx.gap, 2, 1, budget_s=18.0)
                what = "the reference's cpu4-mt-diagrow called once per pair (benchmark.cpp:406), pairs dealt to every host thread" if r["kind"] == "reference" \
                    else "oracle port, OpenMP over pairs"
                line["cpu_baseline"] = {"value": r["value"], "unit": "GCUPS", "cores": r["cores"], "kind": r["kind"],
                                        "sample": f"2 x the first {r['sample_pairs']} pairs of the job, {r['ms']:.0f} ms each ({what})",
                                        "port_gcups": r.get("port_gcups"), "reference_serial_loop_gcups": r.get("serial_gcups")}
            elif args.workload == "pair16k":
                y, x = res["host"]
                g, kind, cores, ms = cpu_pair(y, x, cx.subst, cx.gap, 3, 1)
                line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": cores, "kind": kind,
                                        "sample": f"3 x the full {y.size}x{x.size} pair (cpu4-mt-diagrow blocksz 256 fill + NwTrace1_Plain traceback), {ms:.1f} ms each"}
            else:
                from gpuseqalign_b200 import synth
                x = synth.letters(5001, 32768); y = synth.letters(5004, 32768)
                g, kind, cores, ms = cpu_pair(y, x, cx.subst, cx.gap, 2, 1, with_trace=(args.workload == "wave200k"))
                line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": cores, "kind": kind,
                                        "sample": f"PROXY: 2 x a 32768x32768 pair, {ms:.0f} ms each (the config itself overflows cpu4's int matrix size, SURVEY.md App. D-1)"}
        except Exception as ex:      # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": host_threads(), "kind": "port", "sample": f"failed: {ex}"}
        try:
            line["gpu_reference"] = gpu_reference_block(cx, cx.subst, cx.gap)
        except Exception as ex:
            line["gpu_reference"] = {"failed": f"{type(ex).__name__}: {ex}"}
    print(json.dumps(line), flush=True)
    if cx.world > 1:
        cx.dist.destroy_process_group()
    cx.eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
