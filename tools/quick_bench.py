#!/usr/bin/env python
"""Developer timing loop (not the judged bench): fill-only timings for a few shapes/params."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpuseqalign_b200 import Engine, Params

ITERS = int(os.environ.get('ITERS', '100'))
CONFIGS = os.environ.get('CONFIGS')

def main():
    sc = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scoring.json")))
    subst = np.array(sc["subst"]["blosum62"], dtype=np.int32)
    e = Engine(0)
    e.set_scoring(subst, -11)
    rng = np.random.default_rng(1)
    shapes = [(16384, 16384)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
    for n, m in shapes:
        y = rng.integers(0, 20, n).astype(np.uint8); x = rng.integers(0, 20, m).astype(np.uint8)
        cfgs = [tuple(int(v) for v in c.split(',')) for c in CONFIGS.split(';')] if CONFIGS else None
        for (R, W, K, Bx) in cfgs or [(4, 4, 2, 512), (4, 4, 1, 512), (8, 4, 2, 512), (8, 4, 1, 512), (4, 8, 2, 512), (4, 8, 1, 512), (8, 8, 1, 512)]:
            for keep in ((False, True) if not os.environ.get('KEEP') else (bool(int(os.environ['KEEP'])),)):
                e.upload_pair(y, x, Params(R, W, Bx, K))
                best = 1e9
                for it in range(ITERS):
                    e.fill_resident(keep)
                    s = e.fetch_score()
                    best = min(best, e.timing()["align_calc"])
                print(f"{n}x{m} R={R} W={W} K={K} Bx={Bx} keep={int(keep)} score={s} fill_ms={best:.4f} GCUPS={n*m/best/1e6:.1f}", flush=True)
    e.close()

if __name__ == "__main__":
    main()
