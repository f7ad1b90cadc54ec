"""Benchmark driver over the reference's own file formats: what benchmarkAlgs does for one
algorithm (benchmark.cpp:328-540) -- subst.json + FASTA + pair list in, one TSV row per pair out with the reference's
column names and formats (file_formats.cpp:455-524) -- but with the engine's contexts, buffers and streams kept alive
across pairs instead of re-allocated for every repeat (benchmark.cpp:436-439).

    python -m gpuseqalign_b200.driver -b subst.json -s seqs.fa -p pairs.txt -o out.tsv [--fCalcTrace] [--fCalcScoreHash]

Short pairs (rows <= 512) are scored by the batch kernel in one call; transcripts, hashes and longer pairs go through
the single-pair kernels.  Nothing here computes on the CPU except formatting.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from typing import List, Optional

import numpy as np

from . import formats
from .capi import Engine

ALG_NAME = "NwAlign_B200"

COLUMNS_HEAD = ["alg_name", "seqY_idx", "seqX_idx", "seqY_id", "seqX_id", "seqY_len", "seqX_len", "subst_name", "gapo_cost",
                "warmup_runs", "sample_runs", "last_run_idx", "alg_params", "err_step", "nw_stat", "cuda_stat", "align_cost"]
COLUMNS_MEM = ["sm_count", "ram_peak_allocs", "glmem_peak_allocs", "shmem_peak_allocs", "locmem_peak_allocs", "regmem_peak_allocs"]
COLUMNS_LAPS = ["align.alloc", "align.cpy_dev", "align.init_hdr", "align.calc_init", "align.calc", "align.cpy_host"]


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


def run(subst_path: str, seq_path: str, pair_path: Optional[str], out_path: Optional[str], *, subst_name: str = "blosum62",
        gapo_cost: int = -11, calc_trace: bool = False, calc_hash: bool = False, device: int = 0, engine: Optional[Engine] = None):
    """Aligns every pair of the pair file; returns the list of row dicts (and writes the TSV when out_path is given)."""
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
    own = engine is None
    eng = engine or Engine(device)
    try:
        eng.set_scoring(np.asarray(subst.subst_map[subst_name], dtype=np.int32), gapo_cost)
        letters = [formats.pair_letters(p, seqs) for p in pairs]
        rows = []
        idx = {sid: i for i, sid in enumerate(seqs.ids)}
        # ---- all scores of the short pairs in one batch call
        short = [i for i, (y, x) in enumerate(letters) if y.size <= 512]
        batch_scores = {}
        t_batch = 0.0
        if short and not (calc_trace or calc_hash):
            lens = np.empty(2 * len(short), dtype=np.uint64)
            lens[0::2] = [letters[i][0].size for i in short]; lens[1::2] = [letters[i][1].size for i in short]
            offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
            pool = np.concatenate([np.concatenate(letters[i]) for i in short]).astype(np.uint8) if short else np.zeros(0, np.uint8)
            t0 = time.perf_counter()
            sc = eng.align_batch(pool, offs[0:-1:2].copy(), lens[0::2].astype(np.uint32), offs[1::2].copy(), lens[1::2].astype(np.uint32))
            t_batch = (time.perf_counter() - t0) * 1e3 / max(1, len(short))
            batch_scores = {i: int(s) for i, s in zip(short, sc)}
        for i, (p, (y, x)) in enumerate(zip(pairs, letters)):
            row = {"alg_name": ALG_NAME, "seqY_idx": idx[p.y_id], "seqX_idx": idx[p.x_id],
                   "seqY_id": p.y_id + p.y_range.suffix(), "seqX_id": p.x_id + p.x_range.suffix(),
                   "seqY_len": int(y.size), "seqX_len": int(x.size), "subst_name": subst_name, "gapo_cost": gapo_cost,
                   "warmup_runs": 0, "sample_runs": 1, "last_run_idx": 0, "alg_params": json.dumps({}, separators=(",", ":")),
                   "err_step": 0, "nw_stat": 0, "cuda_stat": 0,
                   "sm_count": 148, "ram_peak_allocs": 0, "glmem_peak_allocs": 0, "shmem_peak_allocs": 0, "locmem_peak_allocs": 0,
                   "regmem_peak_allocs": 0, "align.alloc": 0.0, "align.cpy_dev": 0.0, "align.init_hdr": 0.0, "align.calc_init": 0.0,
                   "align.calc": 0.0, "align.cpy_host": 0.0}
            if i in batch_scores:
                row["align_cost"] = batch_scores[i]
                row["align.calc"] = t_batch
            else:
                row["align_cost"] = eng.align(y, x, keep_headers=calc_trace)
                lap = eng.timing()
                row["align.cpy_dev"], row["align.calc"], row["align.cpy_host"] = lap["align_cpy_dev"], lap["align_calc"], lap["align_cpy_host"]
                if calc_hash:
                    t0 = time.perf_counter()
                    row["score_hash"] = eng.score_hash()
                    row["hash.calc"] = (time.perf_counter() - t0) * 1e3
                if calc_trace:
                    edit, th = eng.trace()
                    lap = eng.timing()
                    row["trace_hash"], row["edit_trace"] = th, edit
                    row["trace.alloc"], row["trace.calc"] = 0.0, lap["trace_calc"] + lap["trace_cpy_host"]
            rows.append(row)
    finally:
        if own:
            eng.close()
    if out_path:
        write_tsv(out_path, rows, calc_hash, calc_trace)
    return rows


def format_field(col: str, v) -> str:
    if col in ("score_hash", "trace_hash"):
        return f"{int(v) & 0xFFFFFFFF:08x}"                     # file_formats.cpp:462-463
    if col in COLUMNS_LAPS or col in ("hash.calc", "trace.alloc", "trace.calc"):
        return f"{float(v):.4f}"                                # file_formats.cpp:464-465
    return str(v)


def write_tsv(path: str, rows, calc_hash: bool, calc_trace: bool):
    cols = tsv_columns(calc_hash, calc_trace)
    with open(path, "w") as f:
        f.write("\t".join(cols) + "\n")
        for r in rows:
            f.write("\t".join(format_field(c, r.get(c, 0)) for c in cols) + "\n")


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-b", "--substPath", required=True)
    ap.add_argument("-s", "--seqPath", required=True)
    ap.add_argument("-p", "--pairPath")
    ap.add_argument("-o", "--resPath")
    ap.add_argument("--substName", default="blosum62")
    ap.add_argument("--gapoCost", type=int, default=-11)
    ap.add_argument("--fCalcTrace", action="store_true")
    ap.add_argument("--fCalcScoreHash", action="store_true")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    rows = run(a.substPath, a.seqPath, a.pairPath, a.resPath, subst_name=a.substName, gapo_cost=a.gapoCost,
               calc_trace=a.fCalcTrace, calc_hash=a.fCalcScoreHash, device=a.device)
    print(f"{len(rows)} pairs aligned" + (f", results in {a.resPath}" if a.resPath else ""))
    return 0


if __name__ == "__main__":
    sys.exit(main())
