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
    for it in range(40 if os.environ.get("SLOW") else 3):
        e.fill_resident(True); s = e.fetch_score()
        if os.environ.get("SLOW") and it > 2 and e.timing()['align_calc'] > float(os.environ["SLOW"]):
            break
    nb = e.info.trows if hasattr(e, "info") else (n + 127) // 128
    nb = (n + 127) // 128
    out = np.zeros(4 * 3 * nb, dtype=np.uint64)
    L.nwb200_debug_band_stamps(e._h, 1 | int(os.environ.get("DBG", "0")), mode, out.ctypes.data_as(C.c_void_p), 3 * nb)
    t = out.reshape(3 * nb, 4).astype(np.int64)
    t0 = t[:nb, 0].min()
    print(f"{n}x{m} mode={mode} score={s} fill_ms={e.timing()['align_calc']:.4f}")
    idx = list(range(min(nb, 14))) + list(range(max(12, nb - 4), nb))
    for b in idx:
        print(f"band {b:4d}: start {(t[b,0]-t0)/1e3:9.1f}us  prologue {(t[b,1]-t0)/1e3:9.1f}us  end {(t[b,2]-t0)/1e3:9.1f}us  dur {(t[b,2]-t[b,1])/1e3:8.1f}us  polls {t[b,3]}")

    if os.environ.get("MAPS"):
        ends = [((t[nb + u][2] - t0) / 1e3, u) for u in range(2, 2 * nb)]
        ends.sort(reverse=True)
        print("latest map units:", [(f"{e:.0f}us", f"u{u}") for e, u in ends[:12]])
        allu = sorted(ends, key=lambda x: x[1])
        print("map unit end times, every 8th:", [f"{e:.0f}" for e, u in allu[::8]])
        for u in list(range(2, 10)) + list(range(2 * nb - 6, 2 * nb)):
            r = t[nb + u]
            print(f"map unit {u:4d} (band {u // 2}, half {u % 2}): start {(r[0]-t0)/1e3:9.1f}us  end {(r[2]-t0)/1e3:9.1f}us  dur {(r[2]-r[0])/1e3:8.1f}us")

if __name__ == "__main__":
    main()
