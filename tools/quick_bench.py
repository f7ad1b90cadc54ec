#!/usr/bin/env python
"""Developer timing loop (not the judged bench): fill / trace timings for a few shapes and kernel shapes.
usage: quick_bench.py NxM [NxM ...]   env: CONFIGS="R,W,K,Bx;..." ITERS=20 TRACE=1"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpuseqalign_b200 import Engine, Params, synth

ITERS = int(os.environ.get('ITERS', '20'))
CONFIGS = os.environ.get('CONFIGS')
TRACE = int(os.environ.get('TRACE', '0'))

def main():
    sc = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scoring.json")))
    subst = np.array(sc["subst"]["blosum62"], dtype=np.int32)
    e = Engine(0)
    e.set_scoring(subst, -11)
    if os.environ.get('DBG'):      # bit 1: prefetch distance 1, bit 2: origin maps in a separate launch (see nwb200_debug_band_stamps)
        import ctypes as C
        e._L.nwb200_debug_band_stamps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        e._L.nwb200_debug_band_stamps(e._h, int(os.environ['DBG']), 0, None, 0)
    shapes = [(16384, 16384)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
    for n, m in shapes:
        y = synth.letters(2002, n); x = synth.letters(2001, m)
        if os.environ.get('MUTATED'): y = synth.mutated_copy(x, 2003, n)
        cfgs = [tuple(int(v) for v in c.split(',')) for c in CONFIGS.split(';')] if CONFIGS else [(0, 0, 0, 0)]
        for (R, W, K, Bx) in cfgs:
            e.upload_pair(y, x, Params(R, W, Bx, K))
            best = 1e9; bt = 1e9
            for it in range(ITERS):
                e.fill_resident(True)
                if TRACE: e.trace_resident()
                s = e.fetch_score()
                if TRACE:
                    e.fetch_trace()
                    bt = min(bt, e.timing()["trace_calc"])
                best = min(best, e.timing()["align_calc"])
                if os.environ.get('VERBOSE'): print(f"   it {it}: fill {e.timing()['align_calc']:.4f}", flush=True)
            print(f"{n}x{m} R={R} W={W} K={K} Bx={Bx} score={s} fill_ms={best:.4f} GCUPS={n*m/best/1e6:.1f}" + (f" trace_ms={bt:.4f} {e.trace_info()}" if TRACE else ""), flush=True)
    e.close()

if __name__ == "__main__":
    main()
