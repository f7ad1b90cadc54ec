"""Benchmark driver over the reference's own file formats: what benchmarkAlgs does for one
algorithm entry (benchmark.cpp:328-540) -- subst.json + FASTA + pair list + parameter JSON in, one TSV row per pair and
parameter combination out, with the reference's column names and formats (file_formats.cpp:455-524) -- but with the
engine's contexts, buffers and streams kept alive across pairs instead of re-allocated for every repeat
(benchmark.cpp:436-439).

    python -m gpuseqalign_b200.driver -b subst.json -s seqs.fa -p pairs.txt -o out.tsv [-r params.json]
           [--warmupPerAlign W] [--samplesPerAlign S] [--fCalcTrace] [--fCalcScoreHash] [--verifyTsv ref.tsv] [--devices 0,1,..]

The reference's loop semantics are kept:
  * every combination of the parameter lists of the "NwAlign_B200" entry of the parameter file is run (cartesian product, last
    key fastest: run_types.cpp:69-83; keys rowsPerLane / warpsPerBlock / tileCols / skew as in plugin/param_b200.json);
  * every run is repeated warmup + samples times; successful warm-up runs are discarded, the laps of the sample runs are
    averaged lap by lap over the runs that recorded them, everything else comes from the last run (benchmark.cpp:149-173,434,
    stopwatch.cpp:4-36);
  * the first result seen for a pair (sequence ids and ranges) is the truth every later result of that pair is verified against:
    align_cost, score_hash, trace_hash (benchmark.cpp:120-147) -- later parameter combinations here, and the rows of a reference
    TSV given with --verifyTsv take the place of the reference algorithm that the stock executable runs first;
  * err_step as in benchmark.cpp:473-497: 1 bad parameters (the combination is skipped), 2 align failed, 3 hash, 4 trace,
    5 result mismatch; any mismatch makes the run fail as a whole (benchmark.cpp:532-536).

Score-only runs with default parameters put all short pairs (rows <= 512) through the batch kernel in one call per repeat,
sharded over --devices (one engine per GPU, one host thread each, no communication); transcripts, hashes, parameter sweeps
and longer pairs go through the single-pair kernels.  Nothing here computes on the CPU except formatting.
"""
from __future__ import annotations

import argparse
import csv
import json
import sys
import threading
import time
from typing import Dict, List, Optional

import numpy as np

from . import formats
from .capi import Engine, NwB200Error, NwStat, Params

ALG_NAME = "NwAlign_B200"
PARAM_KEYS = ("rowsPerLane", "warpsPerBlock", "tileCols", "skew")      # plugin/param_b200.json

COLUMNS_HEAD = ["alg_name", "seqY_idx", "seqX_idx", "seqY_id", "seqX_id", "seqY_len", "seqX_len", "subst_name", "gapo_cost",
                "warmup_runs", "sample_runs", "last_run_idx", "alg_params", "err_step", "nw_stat", "cuda_stat", "align_cost"]
COLUMNS_MEM = ["sm_count", "ram_peak_allocs", "glmem_peak_allocs", "shmem_peak_allocs", "locmem_peak_allocs", "regmem_peak_allocs"]
COLUMNS_LAPS = ["align.alloc", "align.cpy_dev", "align.init_hdr", "align.calc_init", "align.calc", "align.cpy_host"]
LAP_COLUMNS = COLUMNS_LAPS + ["hash.calc", "trace.alloc", "trace.calc"]


def tsv_columns(calc_hash: bool, calc_trace: bool) -> List[str]:
    cols = list(COLUMNS_HEAD)
    if calc_hash:
        cols.append("score_hash")
    if calc_trace:
        cols.append("trace_hash")
    cols += COLUMNS_MEM + COLUMNS_LAPS
    if calc_hash:
        cols.append("hash.calc")
    if calc_trace:
        cols += ["trace.alloc", "trace.calc", "edit_trace"]
    return cols


class RunReport(list):
    """The rows of a run (a list of dicts) + the number of results that disagreed with the first result of their pair."""
    calc_errors: int = 0


def combine_repeats(reps: List[dict]) -> dict:
    """combineRepResults (benchmark.cpp:149-173): the last run's row with every lap replaced by its mean over the runs that have it."""
    out = dict(reps[-1])
    for col in LAP_COLUMNS:
        vals = [r[col] for r in reps if col in r]
        if vals:
            out[col] = sum(vals) / len(vals)
    return out


def _compare_key(p: formats.SeqPair):
    return (p.y_id, p.x_id, p.y_range.suffix(), p.x_range.suffix())


def load_truth(tsv_path: str) -> Dict[tuple, tuple]:
    """(align_cost, score_hash, trace_hash) per pair from a TSV of the reference (first row of a pair wins, like the first algorithm)."""
    truth: Dict[tuple, tuple] = {}
    with open(tsv_path) as f:
        for r in csv.DictReader(f, delimiter="\t"):
            yid, _, yr = r["seqY_id"].partition("[")
            xid, _, xr = r["seqX_id"].partition("[")
            key = (yid, xid, "[" + yr if yr else "", "[" + xr if xr else "")
            if key not in truth and r.get("err_step", "0") == "0":
                truth[key] = (int(r["align_cost"]), int(r["score_hash"], 16) if r.get("score_hash") else None,
                              int(r["trace_hash"], 16) if r.get("trace_hash") else None)
    return truth


def _verify(truth: Dict[tuple, tuple], key, row: dict, calc_hash: bool, calc_trace: bool) -> bool:
    """setOrVerifyResult (benchmark.cpp:120-147); fields a side did not compute are not compared."""
    got = (row["align_cost"], row.get("score_hash") if calc_hash else None, row.get("trace_hash") if calc_trace else None)
    if key not in truth:
        truth[key] = got
        return True
    return all(a is None or b is None or a == b for a, b in zip(got, truth[key]))


def _params_of(combo: dict) -> Optional[Params]:
    if not any(combo.get(k, 0) for k in PARAM_KEYS):
        return None
    return Params(combo.get("rowsPerLane", 0), combo.get("warpsPerBlock", 0), combo.get("tileCols", 0), combo.get("skew", 0))


def _batch_scores(engines: List[Engine], letters, short: List[int]):
    """Scores of the short pairs: contiguous shards balanced by cells, one engine (GPU) and host thread per shard."""
    from .sharding import partition_pairs
    lenY = np.array([letters[i][0].size for i in short], dtype=np.uint32)
    lenX = np.array([letters[i][1].size for i in short], dtype=np.uint32)
    ranges = partition_pairs(lenY, lenX, len(engines))
    out = np.zeros(len(short), dtype=np.int32)
    errs: List[Optional[BaseException]] = [None] * len(engines)

    def work(k):
        lo, hi = ranges[k]
        if hi <= lo:
            return
        try:
            lens = np.empty(2 * (hi - lo), dtype=np.uint64)
            lens[0::2] = lenY[lo:hi]; lens[1::2] = lenX[lo:hi]
            offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
            pool = np.concatenate([np.concatenate(letters[short[i]]) for i in range(lo, hi)] + [np.zeros(1, np.uint8)]).astype(np.uint8)
            out[lo:hi] = engines[k].align_batch(pool, offs[0:-1:2].copy(), lenY[lo:hi].copy(), offs[1::2].copy(), lenX[lo:hi].copy())
        except BaseException as ex:      # re-raised by the caller
            errs[k] = ex

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(engines))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errs:
        if e is not None:
            raise e
    return out


def run(subst_path: str, seq_path: str, pair_path: Optional[str], out_path: Optional[str], *, subst_name: str = "blosum62",
        gapo_cost: int = -11, calc_trace: bool = False, calc_hash: bool = False, device: int = 0, engine: Optional[Engine] = None,
        param_path: Optional[str] = None, warmup: int = 0, samples: int = 1, verify_tsv: Optional[str] = None,
        devices: Optional[List[int]] = None) -> RunReport:
    """Aligns every pair of the pair file under every parameter combination; returns the rows (and writes the TSV when out_path is given)."""
    if samples < 1 or warmup < 0:
        raise ValueError("samplesPerAlign must be >= 1 and warmupPerAlign >= 0")
    subst = formats.read_subst(subst_path)
    if subst_name not in subst.subst_map:
        raise formats.FormatError(f"substitution matrix '{subst_name}' not found")
    seqs = formats.read_fasta(seq_path, subst)
    if pair_path:
        pairs = formats.read_pairs(pair_path, seqs)
    else:           # no pair file: every OTHER sequence (rows, Y) against the first one (columns, X) -- cmd_parser.cpp:466-487
        if len(seqs.ids) < 2:
            raise formats.FormatError("at least two sequences are needed when no pair file is given")
        first = seqs.ids[0]
        pairs = [formats.SeqPair(sid, first, formats.SeqRange(), formats.SeqRange()) for sid in seqs.ids[1:]]
    combos = [{}]
    if param_path:
        allp = formats.read_params(param_path)
        if ALG_NAME not in allp:
            raise formats.FormatError(f"{param_path}: no entry for '{ALG_NAME}'")      # cmd_parser.cpp:375-402
        combos = list(formats.param_combinations(allp[ALG_NAME])) or [{}]
    truth = load_truth(verify_tsv) if verify_tsv else {}
    dev_list = devices if devices else [device]
    own = engine is None
    engines = [engine] if engine is not None else [Engine(d) for d in dev_list]
    eng = engines[0]
    report = RunReport()
    try:
        smat = np.asarray(subst.subst_map[subst_name], dtype=np.int32)
        for e in engines:
            e.set_scoring(smat, gapo_cost)
        letters = [formats.pair_letters(p, seqs) for p in pairs]
        idx = {sid: i for i, sid in enumerate(seqs.ids)}

        def base_row(i, p, combo, run_idx):
            y, x = letters[i]
            return {"alg_name": ALG_NAME, "seqY_idx": idx[p.y_id], "seqX_idx": idx[p.x_id],
                    "seqY_id": p.y_id + p.y_range.suffix(), "seqX_id": p.x_id + p.x_range.suffix(),
                    "seqY_len": int(y.size), "seqX_len": int(x.size), "subst_name": subst_name, "gapo_cost": gapo_cost,
                    "warmup_runs": warmup, "sample_runs": samples, "last_run_idx": run_idx,
                    "alg_params": json.dumps(combo, separators=(",", ":")), "err_step": 0, "nw_stat": 0, "cuda_stat": 0, "align_cost": 0,
                    "sm_count": 148, "ram_peak_allocs": 0, "glmem_peak_allocs": 0, "shmem_peak_allocs": 0, "locmem_peak_allocs": 0,
                    "regmem_peak_allocs": 0}

        def fail_row(row, step, ex):
            row["err_step"] = step
            row["nw_stat"] = int(getattr(ex, "stat", NwStat.errorInvalidResult))
            row["cuda_stat"] = int(getattr(ex, "cuda", 0) or 0)

        # ---- score-only, default parameters: all short pairs in one batch call per repeat (sharded over the devices)
        batch_rows: Dict[int, dict] = {}
        short = [i for i, (y, x) in enumerate(letters) if y.size <= 512]
        if short and combos == [{}] and not (calc_trace or calc_hash):
            reps = []
            for r in range(-warmup, samples):
                t0 = time.perf_counter()
                sc = _batch_scores(engines, letters, short)
                per_pair_ms = (time.perf_counter() - t0) * 1e3 / len(short)
                if r >= 0:
                    reps.append((sc, per_pair_ms))
            sc = reps[-1][0]
            mean_ms = sum(t for _, t in reps) / len(reps)
            for k, i in enumerate(short):
                row = base_row(i, pairs[i], {}, samples - 1)
                row["align_cost"] = int(sc[k])
                row["align.calc"] = mean_ms
                batch_rows[i] = row

        for i, p in enumerate(pairs):
            y, x = letters[i]
            key = _compare_key(p)
            for combo in combos:
                if i in batch_rows:
                    row = batch_rows[i]
                else:
                    reps = []
                    for r in range(-warmup, samples):
                        row = base_row(i, p, combo, r)
                        try:
                            row["align_cost"] = eng.align(y, x, keep_headers=calc_trace, params=_params_of(combo))
                            lap = eng.timing()
                            row["align.cpy_dev"], row["align.calc"], row["align.cpy_host"] = lap["align_cpy_dev"], lap["align_calc"], lap["align_cpy_host"]
                            mu = eng.memory_usage()      # where updateNwAlgPeakMemUsage puts them (nwalign_shared.cpp:16-24)
                            row["ram_peak_allocs"], row["glmem_peak_allocs"] = mu["pinned_host_bytes"], mu["device_bytes"]
                            row["shmem_peak_allocs"], row["locmem_peak_allocs"], row["regmem_peak_allocs"] = mu["shared_bytes"], mu["local_bytes"], mu["register_bytes"]
                        except NwB200Error as ex:        # benchmark.cpp:473-483: bad parameters are step 1, anything else step 2
                            fail_row(row, 1 if ex.stat == NwStat.errorInvalidValue else 2, ex)
                        if calc_hash and not row["err_step"]:
                            try:
                                t0 = time.perf_counter()
                                row["score_hash"] = eng.score_hash()
                                row["hash.calc"] = (time.perf_counter() - t0) * 1e3
                            except NwB200Error as ex:
                                fail_row(row, 3, ex)
                        if calc_trace and not row["err_step"]:
                            try:
                                edit, th = eng.trace()
                                lap = eng.timing()
                                row["trace_hash"], row["edit_trace"] = th, edit
                                row["trace.alloc"], row["trace.calc"] = 0.0, lap["trace_calc"] + lap["trace_cpy_host"]
                            except NwB200Error as ex:
                                fail_row(row, 4, ex)
                        if not row["err_step"] and not _verify(truth, key, row, calc_hash, calc_trace):
                            row["err_step"] = 5
                            row["nw_stat"] = int(NwStat.errorInvalidResult)
                            report.calc_errors += 1
                        if r >= 0 or row["err_step"]:          # successful warm-up runs are discarded
                            reps.append(row)
                        if row["err_step"]:
                            break
                    row = combine_repeats(reps)
                if i in batch_rows and not _verify(truth, key, row, False, False):
                    row["err_step"] = 5
                    row["nw_stat"] = int(NwStat.errorInvalidResult)
                    report.calc_errors += 1
                report.append(row)
    finally:
        if own:
            for e in engines:
                e.close()
    if out_path:
        write_tsv(out_path, report, calc_hash, calc_trace)
    return report


def format_field(col: str, v) -> str:
    if col in ("score_hash", "trace_hash"):
        return f"{int(v) & 0xFFFFFFFF:08x}"                     # file_formats.cpp:462-463
    if col in LAP_COLUMNS:
        return f"{float(v):.4f}"                                # file_formats.cpp:464-465
    return str(v)


def write_tsv(path: str, rows, calc_hash: bool, calc_trace: bool):
    cols = tsv_columns(calc_hash, calc_trace)
    with open(path, "w") as f:
        f.write("\t".join(cols) + "\n")
        for r in rows:
            f.write("\t".join(format_field(c, r.get(c, "" if c == "edit_trace" else 0)) for c in cols) + "\n")


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-b", "--substPath", required=True)
    ap.add_argument("-s", "--seqPath", required=True)
    ap.add_argument("-p", "--pairPath")
    ap.add_argument("-o", "--resPath")
    ap.add_argument("-r", "--algParamPath", help="parameter JSON with an 'NwAlign_B200' entry (plugin/param_b200.json)")
    ap.add_argument("--substName", default="blosum62")
    ap.add_argument("--gapoCost", type=int, default=-11)
    ap.add_argument("--warmupPerAlign", type=int, default=0)
    ap.add_argument("--samplesPerAlign", type=int, default=1)
    ap.add_argument("--fCalcTrace", action="store_true")
    ap.add_argument("--fCalcScoreHash", action="store_true")
    ap.add_argument("--verifyTsv", help="TSV of the reference: its first row per pair is the truth the results are verified against")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--devices", help="comma list of GPUs for the batch of short pairs (score-only runs)")
    a = ap.parse_args(argv)
    rows = run(a.substPath, a.seqPath, a.pairPath, a.resPath, subst_name=a.substName, gapo_cost=a.gapoCost,
               calc_trace=a.fCalcTrace, calc_hash=a.fCalcScoreHash, device=a.device, param_path=a.algParamPath,
               warmup=a.warmupPerAlign, samples=a.samplesPerAlign, verify_tsv=a.verifyTsv,
               devices=[int(v) for v in a.devices.split(",")] if a.devices else None)
    print(f"{len(rows)} results" + (f", written to {a.resPath}" if a.resPath else ""))
    if rows.calc_errors:
        print(f"error: {rows.calc_errors} result(s) disagree with the first result of their pair", file=sys.stderr)
        return int(NwStat.errorInvalidResult)                   # benchmark.cpp:532-536
    return 0


if __name__ == "__main__":
    sys.exit(main())
