#!/usr/bin/env python
"""Times transcripts of a batch (cfg3 sample: 4096 pairs of 256 x 256) through nwb200_align_batch with edits, checks a few pairs
against the oracle.  usage: python tools/batch_trace_time.py [pairs]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpuseqalign_b200 import Engine, synth
from oracle import pyoracle

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    with open(os.path.join(ROOT, "tests", "golden", "scoring.json")) as f:
        subst = np.array(json.load(f)["subst"]["blosum62"], dtype=np.int32)
    pool, offY, lenY, offX, lenX = synth.batch_pairs(0, n, 256, 256)
    e = Engine(0)
    e.set_scoring(subst, -11)
    e.align_batch(pool, offY, lenY, offX, lenX, want_trace=True)
    t0 = time.perf_counter(); l0 = e.launches()
    scores, edits, hashes = e.align_batch(pool, offY, lenY, offX, lenX, want_trace=True)
    dt = time.perf_counter() - t0
    for p in range(0, n, max(1, n // 16)):
        y = pool[int(offY[p]): int(offY[p]) + int(lenY[p])]; x = pool[int(offX[p]): int(offX[p]) + int(lenX[p])]
        exp = pyoracle.align_pair(y, x, subst, -11, want_hash=False, want_trace=True)
        assert (scores[p], edits[p], hashes[p]) == (exp.score, exp.edit, exp.trace_hash), p
    print(json.dumps({"pairs": n, "shape": "256x256", "seconds_scores_and_transcripts": dt, "us_per_pair": dt / n * 1e6,
                      "kernel_launches": e.launches() - l0, "checked_vs_oracle": 16}))
    e.close()

if __name__ == "__main__":
    main()
