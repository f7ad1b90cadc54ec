#!/usr/bin/env python
"""One resident batch of N synthetic 256 x 256 pairs through nwb200_batch_resident, a few times (the program ncu is pointed at):
   tools/batch_one.py [pairs=131072] [repeats=3]      env: NWB200_BATCH_PACKED / NWB200_BATCH_VARIANT / NWB200_BATCH_TMA select the kernel"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpuseqalign_b200 import Engine, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
subst = np.array(json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scoring.json")))["subst"]["blosum62"], dtype=np.int32)
e = Engine(0)
e.set_scoring(subst, -11)
pool, oy, ly, ox, lx = synth.batch_pairs(0, n, 256, 256)
e.upload_batch(pool, oy, ly, ox, lx)
best = 1e9
for _ in range(reps):
    e.batch_resident()
    s = e.fetch_batch_scores()
    best = min(best, e.timing()["align_calc"])
print(json.dumps({"pairs": n, "kernel": e.batch_kernel_name() if hasattr(e, "batch_kernel_name") else "", "best_ms": round(best, 4),
                  "GCUPS": round(n * 65536 / best / 1e6, 1), "score_sum": int(s.astype(np.int64).sum())}))
e.close()
