#!/usr/bin/env python
"""Turns an .ncu-rep (one kernel launch, --set full --import-source on) into the short text summary committed under
profiles/: duration, DRAM traffic, issue rate, pipe utilisation, occupancy, stall breakdown of the hot loop and its
instruction mix.   usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [cells_per_launch] > profiles/xxx.txt"""
import collections, csv, io, subprocess, sys

def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1]] + args, capture_output=True, text=True).stdout

def main():
    cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
    h, units, v = raw[0], raw[1], raw[2]
    m = {k: (v[i], units[i]) for i, k in enumerate(h)}
    print(f"kernel: {m.get('Kernel Name', ('?',))[0]}   grid {m.get('launch__grid_size', ('?',))[0]} x block {m.get('launch__block_size', ('?',))[0]}, "
          f"{m.get('launch__registers_per_thread', ('?',))[0]} regs/thread")
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "lts__t_bytes.sum"]
    for k in keys:
        if k in m:
            print(f"  {k:75s} {m[k][0]:>16s} {m[k][1]}")
    if cells and "gpu__time_duration.sum" in m:
        val, unit = m["gpu__time_duration.sum"]
        sec = float(val.replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1, "second": 1}.get(unit, 1e-9)
        print(f"  cells per launch {cells:.4g} -> {cells / sec / 1e9:.1f} GCUPS under the profiler (cold caches, serialised; shares only)")
        if "smsp__inst_executed.sum" in m:
            print(f"  warp instructions per 32 cells: {float(m['smsp__inst_executed.sum'][0].replace(',', '')) / (cells / 32):.2f}")
    src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv"]))))
    hdr, data = src[1], src[2:]
    ix = {k: i for i, k in enumerate(hdr)}
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    cnt = collections.Counter(int(r[ix["Instructions Executed"]] or 0) for r in data if r[ix["Instructions Executed"]])
    cands = [c for c, n in cnt.most_common(6) if c > 0 and n >= 200]
    target = max(cands) if cands else 0
    hot = [r for r in data if r[ix["Instructions Executed"]] and int(r[ix["Instructions Executed"]]) == target]
    n = sum(int(r[ix["# Samples"]] or 0) for r in hot); alln = sum(int(r[ix["# Samples"]] or 0) for r in data)
    print(f"hot loop: {len(hot)} SASS instructions executed {target} times each = {len(hot) / 32:.1f} per step of the unrolled 32-step chunk; "
          f"{n} of {alln} stall samples")
    tot = collections.Counter()
    for r in hot:
        for k in stalls:
            tot[k] += int(r[ix[k]] or 0)
    print("  stall reasons in the hot loop: " + ", ".join(f"{k[6:]} {100 * c / max(n, 1):.1f}%" for k, c in tot.most_common(7)))
    ops = collections.Counter()
    for r in hot:
        s = r[ix["Source"]].split(); op = s[1] if s[0].startswith("@") else s[0]; ops[op.split(".")[0]] += 1
    print("  instruction mix per step: " + ", ".join(f"{k} {c / 32:.2f}" for k, c in ops.most_common(12)))
    allst = collections.Counter()
    for r in data:
        for k in stalls:
            allst[k] += int(r[ix[k]] or 0)
    print("  stall reasons, whole kernel: " + ", ".join(f"{k[6:]} {100 * c / max(alln, 1):.1f}%" for k, c in allst.most_common(6)))

if __name__ == "__main__":
    main()
