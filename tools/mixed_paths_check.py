#!/usr/bin/env python
"""Small invocations of the round-2 code paths in one process (a quick sanity run; also what to point a memory checker at):
corridor traceback (hit and miss), packed 5-bit batch, affine / local variants, TMA letter staging, score rows / path values."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpuseqalign_b200 import Engine, synth

subst = np.array(json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scoring.json")))["subst"]["blosum62"], dtype=np.int32)
e = Engine(0)
e.set_scoring(subst, -11)
n = 40000
x = synth.letters(611, n)
for name, y in (("mutated", synth.mutated_copy(x, 612, n)), ("indel", np.concatenate([x[:19000], x[21500:], synth.letters(613, 2500)]))):
    s = e.align(y, x, keep_headers=True)
    edit, th = e.trace()
    print(name, s, f"{th:08x}", e.trace_info(), flush=True)
y = synth.letters(5, 700); x2 = synth.letters(6, 900)
e.align(y, x2, keep_headers=True); e.trace()
print("rows", e.score_rows(0, 701, 901).sum(), "values", e.trace_values().sum(), e.memory_usage()["device_bytes"] > 0, flush=True)
rng = np.random.default_rng(3)
npairs = 301
lenY = rng.integers(0, 257, npairs).astype(np.uint32); lenX = rng.integers(0, 300, npairs).astype(np.uint32)
lens = np.empty(2 * npairs, dtype=np.uint64); lens[0::2] = lenY; lens[1::2] = lenX
offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
letters = rng.integers(0, 20, int(offs[-1]) + 1).astype(np.uint8)
offY, offX = offs[0:-1:2].copy(), offs[1::2].copy()
ref = e.align_batch(letters, offY, lenY, offX, lenX)
o = np.empty(2 * npairs, dtype=np.int64); l = np.empty(2 * npairs, dtype=np.int64)
o[0::2] = offY; o[1::2] = offX; l[0::2] = lenY; l[1::2] = lenX
packed, no = synth.pack5(letters, o, l)
packed = packed[: packed.size - 64 + 8].copy()          # (only the slack the API promises nothing about: the device pool has its own)
got = e.align_batch_packed5(packed, no[0::2].copy(), lenY, no[1::2].copy(), lenX)
print("packed5 equal", bool(np.array_equal(ref, got)), flush=True)
for v in ("nw_affine", "sw_affine", "sw_linear"):
    print(v, int(e.align_batch_variant(letters, offY, lenY, offX, lenX, v, -11, -1).sum()), flush=True)
os.environ["NWB200_BATCH_PACKED"] = "0"; os.environ["NWB200_BATCH_TMA"] = "1"
pool, oY, lY, oX, lX = synth.batch_pairs(0, 300, 256, 256)
a = e.align_batch(pool, oY, lY, oX, lX)
del os.environ["NWB200_BATCH_TMA"]
b = e.align_batch(pool, oY, lY, oX, lX)
print("tma equal", bool(np.array_equal(a, b)), flush=True)
e.close()
