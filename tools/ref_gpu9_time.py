#!/usr/bin/env python
"""Times the UNMODIFIED reference (oracle/_ref/libnwref.so: gpu9 and cpu4) on this box for a synthetic pair."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyoracle
from gpuseqalign_b200 import synth

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    m = int(sys.argv[2]) if len(sys.argv) > 2 else n
    sc = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scoring.json")))
    subst = np.array(sc["subst"]["blosum62"], dtype=np.int32)
    x = synth.letters(2001, m); y = synth.letters(2002, n)
    for alg in ("gpu9", "cpu4"):
        for it in range(4):
            r = pyoracle.ref_run(alg, y, x, subst, -11, want_hash=False, want_trace=True)
            tot = sum(v for k, v in r.laps_ms.items() if k.startswith("align"))
            print(json.dumps({"alg": alg, "n": n, "m": m, "it": it, "score": r.score, "trace_hash": f"{r.trace_hash:08x}",
                              "laps_ms": {k: round(v, 3) for k, v in r.laps_ms.items()},
                              "gcups_align_calc": n * m / r.laps_ms["align_calc"] / 1e6, "gcups_align_total": n * m / tot / 1e6,
                              "cores": os.cpu_count()}), flush=True)

if __name__ == "__main__":
    main()
