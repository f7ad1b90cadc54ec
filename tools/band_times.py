#!/usr/bin/env python
"""Developer aid: per-band start / prologue / end stamps of one fill. usage: band_times.py N M [mode]"""
import os, sys, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpuseqalign_b200 import Engine, Params, synth

def main():
    n, m = int(sys.argv[1]), int(sys.argv[2])
    mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    sc = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scoring.json")))
    e = Engine(0); e.set_scoring(np.array(sc["subst"]["blosum62"], dtype=np.int32), -11)
    y = synth.letters(2002, n); x = synth.letters(2001, m)
    L = e._L
    L.nwb200_debug_band_stamps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.nwb200_debug_band_stamps(e._h, 1 | int(os.environ.get("DBG", "0")), mode, None, 0)
    e.upload_pair(y, x, Params(0, int(os.environ.get("W", "0")), 0, 0))
    for _ in range(3):
        e.fill_resident(True); s = e.fetch_score()
    nb = e.info.trows if hasattr(e, "info") else (n + 127) // 128
    nb = (n + 127) // 128
    out = np.zeros(4 * nb, dtype=np.uint64)
    L.nwb200_debug_band_stamps(e._h, 1 | int(os.environ.get("DBG", "0")), mode, out.ctypes.data_as(C.c_void_p), nb)
    t = out.reshape(nb, 4).astype(np.int64)
    t0 = t[:, 0].min()
    print(f"{n}x{m} mode={mode} score={s} fill_ms={e.timing()['align_calc']:.4f}")
    idx = list(range(min(nb, 14))) + list(range(max(12, nb - 4), nb))
    for b in idx:
        print(f"band {b:4d}: start {(t[b,0]-t0)/1e3:9.1f}us  prologue {(t[b,1]-t0)/1e3:9.1f}us  end {(t[b,2]-t0)/1e3:9.1f}us  dur {(t[b,2]-t[b,1])/1e3:8.1f}us  polls {t[b,3]}")

if __name__ == "__main__":
    main()
