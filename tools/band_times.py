#!/usr/bin/env python
"""Developer aid: per-band timeline of one fill (globaltimer stamps written by the kernel)."""
import os, sys, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpuseqalign_b200 import Engine, Params
n, m = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "16384x16384").split("x"))
R, W, K = (int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "4,4,2").split(","))
sc = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scoring.json")))
e = Engine(0); e.set_scoring(np.array(sc["subst"]["blosum62"], dtype=np.int32), -11)
rng = np.random.default_rng(1)
y = rng.integers(0, 20, n).astype(np.uint8); x = rng.integers(0, 20, m).astype(np.uint8)
e.upload_pair(y, x, Params(R, W, 512, K))
L = e._L
L.nwb200_debug_band_stamps.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int]
for it in range(int(os.environ.get("WARM", "20"))):
    e.fill_resident(False); e.fetch_score()
BO = int(os.environ.get("BACKOFF", "1500"))
L.nwb200_debug_band_stamps(e._h, 1 | (BO << 8), None, 0)
e.fill_resident(False); s = e.fetch_score()
out = np.zeros(4 * 4096 + 2400, dtype=np.uint64)
nb = L.nwb200_debug_band_stamps(e._h, 1, out.ctypes.data_as(C.c_void_p), 4096)
t = out[: 4 * nb].reshape(nb, 4).astype(np.int64)
t0 = t[:, 0].min()
print(f"{n}x{m} R={R} W={W} K={K} score={s} fill_ms={e.timing()['align_calc']:.4f} bands={nb}")
for b in range(nb):
    print(f"band {b:3d}: start {(t[b,0]-t0)/1e3:8.1f}us  loop_begin {(t[b,1]-t0)/1e3:8.1f}us  loop_end {(t[b,2]-t0)/1e3:8.1f}us  loop {(t[b,2]-t[b,1])/1e3:8.1f}us  miss_chunks {t[b,3]>>16} repolls {t[b,3]&0xffff}")

ch = out[4 * 4096:].reshape(4, 600).astype(np.int64)
for b in range(min(nb, 4)):
    d = np.diff(ch[b][:530]) / 1e3
    d = d[(d > 0) & (d < 1e6)]
    med = np.median(d)
    big = [(i, round(float(v), 2)) for i, v in enumerate(d) if v > 1.5 * med]
    print(f"band {b}: median chunk {med:.3f}us  mean {d.mean():.3f}us  n_big {len(big)}  big: {big[:40]}")
