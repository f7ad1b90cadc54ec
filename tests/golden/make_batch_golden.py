#!/usr/bin/env python
"""Generates tests/golden/batch_golden.json: checksums of ALL scores of BASELINE config 3 at full size (1 048 576 synthetic
256 x 256 pairs, seeds of SURVEY.md 8(d)), computed by the CPU oracle's rolling-row scorer (nwo_score_batch, pinned against the
reference's goldens by tests/test_oracle.py and against the reference's own cpu4 per pair on a prefix, below).  About two
minutes of CPU; the GPU tests and bench.py only read the committed JSON.   usage: python tests/golden/make_batch_golden.py"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from gpuseqalign_b200 import synth
from oracle import pyoracle

OUT = os.path.join(ROOT, "tests", "golden", "batch_golden.json")


def main():
    with open(os.path.join(ROOT, "tests", "golden", "scoring.json")) as f:
        subst = np.array(json.load(f)["subst"]["blosum62"], dtype=np.int32)
    total, block = 1 << 20, 1 << 16
    t0 = time.time()
    scores = np.empty(total, dtype=np.int32)
    for lo in range(0, total, block):
        pool, offY, lenY, offX, lenX = synth.batch_pairs(lo, block, 256, 256)
        scores[lo:lo + block] = pyoracle.score_batch(pool, offY, lenY, offX, lenX, subst, -11, threads=os.cpu_count() or 1)
        if lo == 0 and pyoracle.ref_available():       # the reference's own cpu4, pair by pair, on a prefix
            ref, _ = pyoracle.ref_batch_cpu("cpu4", pool, offY[:2048], lenY[:2048], offX[:2048], lenX[:2048], subst, -11,
                                            threads=os.cpu_count() or 1)
            assert (ref == scores[:2048]).all(), "oracle port and reference cpu4 disagree"
        print(lo + block, round(time.time() - t0, 1), flush=True)
    res = {"desc": "cfg3: pair p has X seed 3e6+2p and Y seed 3e6+2p+1 (splitmix64 residues), blosum62, gap -11; int32 scores little-endian",
           "pairs": total, "len_y": 256, "len_x": 256,
           "sha256_all": hashlib.sha256(scores.tobytes()).hexdigest(),
           "sum_all": int(scores.astype(np.int64).sum()),
           # strong-scaling shards hash their own range: prefix digests at every 1/8 of the batch
           "sha256_eighths": [hashlib.sha256(scores[k * (total // 8):(k + 1) * (total // 8)].tobytes()).hexdigest() for k in range(8)],
           "sum_eighths": [int(scores[k * (total // 8):(k + 1) * (total // 8)].astype(np.int64).sum()) for k in range(8)],
           "first_scores": [int(v) for v in scores[:16]],
           "oracle_seconds": round(time.time() - t0, 1)}
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print(res)


if __name__ == "__main__":
    main()
