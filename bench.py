#!/usr/bin/env python
"""bench.py -- GCUPS of the NW linear-gap hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 20 --warmup 3                       # this engine, cfg2 (16k x 16k pair)
    python bench.py --workload batch256                                  # cfg3 (batch of 256 x 256 pairs)
    python bench.py --impl reference                                     # the reference's cpu4 path on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic input:
  pair16k  : one 16 384 x 16 384 protein pair per GPU -- score-matrix fill + traceback (BASELINE.json configs[1]);
  batch256 : a batch of 256 x 256 pairs per GPU, scores only (configs[2]), pairs sharded over the ranks.
`value` is whole-job GCUPS with the inputs resident in HBM (device time, CUDA events on the engine's stream, max
over ranks); `e2e` is the same through the public call with HOST buffers (H2D + kernels + D2H inside the region).
Weak scaling: every rank aligns its own pair(s); there is no data-path collective.
"""
from __future__ import annotations

import argparse
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
MIX16_CELLS_PER_CLK_PER_SM = 54.2  # measured IDP.4A + IDP.2A + VIMNMX3.U16x2 rate (profiles/microbench_r1z.jsonl): the packed batch kernel's
                                   # three instructions per TWO cells


def load_scoring():
    with open(os.path.join(ROOT, "tests", "golden", "scoring.json")) as f:
        s = json.load(f)
    return np.array(s["subst"]["blosum62"], dtype=np.int32), -11


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# --------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed region runs (NVML)."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index: int, period_s: float = 0.02):
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


def pair_config(n, m, with_trace=True):
    return {"workload": f"cfg2: single synthetic protein pair {n}x{m} per GPU, score" + ("+traceback" if with_trace else " only"),
            "pairs_per_gpu": 1, "seeds": "X 2001+10r, Y 2002+10r (splitmix64, independent)", "subst": "blosum62", "gap": -11,
            "l2": "flushed between timed steps (256 MiB memset)"}


def batch_config(total_pairs, per):
    return {"workload": f"cfg3: batch of {total_pairs} synthetic 256x256 pairs, scores only, sharded {per}/GPU",
            "pairs_per_gpu": per, "seeds": "pair p: X 3e6+2p, Y 3e6+2p+1", "subst": "blosum62", "gap": -11,
            "l2": "inputs (512 B/pair) exceed L2 at the full batch; flushed between timed steps as well"}


# --------------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference_pair(y, x, subst, gap, samples, warmup):
    """The reference's own cpu4-mt-diagrow (+ NwTrace1_Plain) from oracle/_ref when it was prebuilt, else the
    C restatement (oracle port).  Returns (gcups, kind, cores, ms_per_sample)."""
    from oracle import pyoracle
    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    cores = int(os.environ["OMP_NUM_THREADS"])            # the threads actually used
    cells = float(y.size) * float(x.size)
    times = []
    if pyoracle.ref_available():
        kind = "reference"
        for it in range(warmup + samples):
            r = pyoracle.ref_run("cpu4", y, x, subst, gap, want_hash=False, want_trace=True)
            ms = r.laps_ms["align_calc"] + r.laps_ms["trace_calc"]
            if it >= warmup:
                times.append(ms)
    else:
        kind = "port"
        import subprocess
        if not os.path.exists(pyoracle.ORACLE_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
        for it in range(warmup + samples):
            t0 = time.perf_counter()
            pyoracle.align_pair(y, x, subst, gap, want_hash=False, want_trace=True, threads=cores)
            if it >= warmup:
                times.append((time.perf_counter() - t0) * 1e3)
    ms = float(np.mean(times))
    return cells / ms / 1e6, kind, cores, ms


def cpu_reference_batch(pool, offY, lenY, offX, lenX, subst, gap, samples, warmup):
    """cfg3 on the host: the oracle port's rolling-row scorer, one pair per thread task (the reference's cpu4
    degenerates to a single tile per 256 x 256 pair)."""
    from oracle import pyoracle
    import subprocess
    if not os.path.exists(pyoracle.ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    cores = os.cpu_count() or 1
    cells = float(np.sum(lenY.astype(np.float64) * lenX.astype(np.float64)))
    times = []
    for it in range(warmup + samples):
        t0 = time.perf_counter()
        pyoracle.score_batch(pool, offY, lenY, offX, lenX, subst, gap, threads=cores)
        if it >= warmup:
            times.append((time.perf_counter() - t0) * 1e3)
    ms = float(np.mean(times))
    return cells / ms / 1e6, "port", cores, ms


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun exports OMP_NUM_THREADS=1 to every rank; this arm is ONE process that owns the host (the other ranks have left):
    # give the reference's OpenMP path all the cores, as at N = 1
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from gpuseqalign_b200 import synth
    subst, gap = load_scoring()
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if args.workload == "pair16k":
        n = m = args.len
        x = synth.letters(2001, m); y = synth.letters(2002, n)
        # bounded: each step is the whole 16k x 16k pair (0.2-0.4 s on 16 cores); cap the step count
        steps = min(steps, 10); warmup = min(warmup, 2)
        g, kind, cores, ms = cpu_reference_pair(y, x, subst, gap, steps, warmup)
        sample = f"{steps} x the full {n}x{m} pair (cpu4-mt-diagrow blocksz 256 fill + NwTrace1_Plain traceback)"
        cfg = pair_config(n, m)
    else:
        npairs = min(args.pairs, 20000)
        pool, offY, lenY, offX, lenX = synth.batch_pairs(0, npairs, 256, 256)
        steps = min(steps, 5); warmup = min(warmup, 1)
        g, kind, cores, ms = cpu_reference_batch(pool, offY, lenY, offX, lenX, subst, gap, steps, warmup)
        sample = f"{steps} x the first {npairs} pairs of the batch (256x256, scores only)"
        cfg = batch_config(args.pairs, args.pairs // max(1, args.gpus))
        cfg["reference_sample_pairs"] = npairs
    line = {"impl": "reference", "metric": "GCUPS NW linear-gap", "value": g, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": g, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": g, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------- this engine
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pair16k", choices=["pair16k", "batch256"])
    ap.add_argument("--len", type=int, default=16384, help="pair16k: sequence length")
    ap.add_argument("--pairs", type=int, default=1 << 20, help="batch256: pairs in the whole job")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-trace", action="store_true", help="pair16k: fill only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from gpuseqalign_b200 import Engine, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # (only where host-to-device traffic matters: the batch workload moves 562 MB per step and rank)
    numa = bind_to_gpu_numa_node(torch, local) if (world > 1 and args.workload == "batch256") else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    subst, gap = load_scoring()
    eng = Engine(local)
    eng.set_scoring(subst, gap)
    stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    if args.workload == "pair16k":
        n = m = args.len
        x = synth.letters(2001 + 10 * rank, m); y = synth.letters(2002 + 10 * rank, n)
        cells_rank = float(n) * float(m)
        with_trace = not args.no_trace
        eng.upload_pair(y, x)

        def step_resident():
            eng.fill_resident(True)
            if with_trace:
                eng.trace_resident()

        def step_e2e():
            s = eng.align(y, x, keep_headers=True, with_trace=with_trace)
            if with_trace:
                e, h = eng.trace()
                return 4 + len(e) + 4
            return 4

        h2d = n + m
        cfg = pair_config(n, m, with_trace)
    else:
        per = args.pairs // world
        first = rank * per
        pool, offY, lenY, offX, lenX = synth.batch_pairs(first, per, 256, 256)
        pool = torch.from_numpy(pool).pin_memory().numpy()          # e2e: H2D from pinned host memory
        cells_rank = float(per) * 256.0 * 256.0
        eng.upload_batch(pool, offY, lenY, offX, lenX)

        def step_resident():
            eng.batch_resident()

        def step_e2e():
            eng.align_batch(pool, offY, lenY, offX, lenX)
            return 4 * per

        h2d = pool.size + 24 * per
        cfg = batch_config(args.pairs, per)

    # ---- device-resident timing ------------------------------------------------------------
    for _ in range(warmup):
        step_resident()
    eng.sync()
    l0 = eng.launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    kernel_ms = []
    barrier()
    with ClockSampler(local) as clk:
        with torch.cuda.stream(stream):
            for a, b in ev:
                flush.fill_(1)
                a.record(stream)
                step_resident()
                b.record(stream)
        eng.sync()
        barrier()
    launches = eng.launches() - l0
    ms_total = sum(a.elapsed_time(b) for a, b in ev)
    ms_total = max_over_ranks(ms_total)
    cells_job = sum_over_ranks(cells_rank)
    value = cells_job * steps / ms_total / 1e6
    if args.workload == "pair16k":
        eng.fetch_score()                       # lands the CUDA-event laps of the last timed step
        if with_trace:
            eng.fetch_trace()
        lap = eng.timing()
        fill_ms = lap.get("align_calc", 0.0)
    else:
        eng.fetch_batch_scores()
        lap = eng.timing()
        fill_ms = lap.get("align_calc", 0.0)

    # ---- end to end through the public call with host buffers ---------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(steps):
        d2h = step_e2e()
    eng.sync()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = cells_job * steps / e2e_s / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    clocks = clk.summary()
    f_ghz = (clocks.get("sm_max_mhz") or measured_peaks().get("sm_max_mhz") or 1965.0) / 1e3
    peak = SM_COUNT * DPX_PER_CLK_PER_SM * f_ghz            # GCUPS at 1 DPX op (VIMNMX3) per cell
    achieved = cells_rank / fill_ms / 1e6 if fill_ms > 0 else 0.0
    line = {"metric": "GCUPS NW linear-gap", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "int-issue (DPX VIMNMX3, 1 per cell; not hbm/tensor)", "achieved": achieved, "peak": peak, "unit": "GCUPS",
                         "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full captures
                         # (profiles/r1y_ncu_fill_cluster16_16k.txt: 35.5 MB for the 16k pair under the profiler's cold caches --
                         # header rows, middle rows, origin maps and snapshots, 59 MB in all, live in L2 otherwise;
                         # profiles/r1z_ncu_batch2_*.txt: 544 B per pair)
                         "traffic": (35489792 if (args.workload == "pair16k" and args.len == 16384 and with_trace) else
                                     (544.2 * per if args.workload == "batch256" else None)),
                         "kernel": "nw_fill_kernel" if args.workload == "pair16k" else
                                   ("nw_batch_kernel" if os.environ.get("NWB200_BATCH_PACKED", "1").startswith("0") else "nw_batch2_kernel"),
                         "kernel_ms": fill_ms,
                         "laps_ms_last_step": {k: round(v, 4) for k, v in lap.items()},
                         "peak_source": f"{SM_COUNT} SMs x {DPX_PER_CLK_PER_SM:.0f} VIMNMX3/clk/SM (measured, profiles/microbench_r1.jsonl) x {f_ghz:.3f} GHz",
                         "peak_mix_measured": SM_COUNT * MIX_CELLS_PER_CLK_PER_SM * f_ghz}}
    if args.workload == "pair16k" and fill_ms > 0:
        # A lone pair is bound by its chain of dependent steps, not by issue slots: bands of 128 rows, a lane
        # R = 4 rows deep, lanes two steps apart, a band following the one above at the lane pipeline + one hand-off group.
        # Floor of a step = its R dependent VIMNMX3 (4.47 clk each, profiles/microbench_r1.jsonl).
        R_, K_, grp = 4, 2, 8
        nb_ = (n + 32 * R_ - 1) // (32 * R_)
        dep_steps = m + 31 * K_ + (nb_ - 1) * (31 * K_ + grp)
        floor_ms = dep_steps * R_ * 4.47 / (f_ghz * 1e6)
        line["roofline"]["latency_floor"] = {"dependent_steps": dep_steps, "clk_per_step_floor": R_ * 4.47, "ms": floor_ms,
                                             "frac": floor_ms / fill_ms,
                                             "note": "critical path of the band wavefront at the DPX dependent-issue latency; the fill kernel's time against it"}
    if numa:
        line["config"]["cpu_affinity"] = f"rank 0 bound to the CPUs next to its GPU ({numa}); every rank does the same"
    if args.workload == "batch256" and line["roofline"]["kernel"] == "nw_batch2_kernel":
        # the packed kernel computes TWO cells per VIMNMX3.U16x2: against a roofline of one DPX op per two cells the same rate is half
        # the fraction; both are reported, together with the measured rate of the kernel's own three-instruction mix
        line["roofline"]["peak_packed16"] = 2.0 * peak
        line["roofline"]["frac_packed16"] = achieved / (2.0 * peak)
        line["roofline"]["peak_mix_measured"] = SM_COUNT * MIX16_CELLS_PER_CLK_PER_SM * f_ghz
        line["roofline"]["frac_of_mix_measured"] = achieved / (SM_COUNT * MIX16_CELLS_PER_CLK_PER_SM * f_ghz)
        line["dtype"] = "u16x2 (two pairs per 32-bit register)"
    if world == 1 and not args.no_cpu_baseline:
        try:
            if args.workload == "pair16k":
                g, kind, cores, ms = cpu_reference_pair(y, x, subst, gap, 3, 1)
                sample = f"3 x the full {n}x{m} pair (cpu4-mt-diagrow blocksz 256 fill + NwTrace1_Plain traceback), {ms:.1f} ms each"
            else:
                k = min(per, 20000)
                g, kind, cores, ms = cpu_reference_batch(pool[: 0] if False else pool, offY[:k], lenY[:k], offX[:k], lenX[:k], subst, gap, 3, 1)
                sample = f"3 x the first {k} pairs of the batch, {ms:.1f} ms each"
            line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": sample}
        except Exception as ex:      # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
