#!/usr/bin/env python
"""Generates tests/golden/big_golden.json: known answers for BASELINE configs 2, 4 and 5 at FULL size, computed by the
CPU oracle (rolling-row header producer + sparse traceback = the restated gpu9 layout + NwTrace2_Sparse, both pinned
against the reference's golden vectors by tests/test_oracle.py).  Takes several minutes of CPU; the GPU tests only
read the committed JSON.   usage: python tests/golden/make_big_golden.py [cfg2 cfg4 cfg5a cfg5b]"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from gpuseqalign_b200 import synth
from oracle import pyoracle

OUT = os.path.join(ROOT, "tests", "golden", "big_golden.json")

def main():
    want = sys.argv[1:] or ["cfg2", "cfg4", "cfg5a", "cfg5b"]
    with open(os.path.join(ROOT, "tests", "golden", "scoring.json")) as f:
        subst = np.array(json.load(f)["subst"]["blosum62"], dtype=np.int32)
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    def record(name, y, x, trace, desc):
        t0 = time.time()
        By, Bx = 512, 1024
        if trace:
            score, hrow, hcol, _ = pyoracle.fill_rolling(y, x, subst, -11, By, Bx)
            r = pyoracle.trace_sparse(hrow, hcol, By, Bx, y, x, subst, -11)
            assert r.score == score
            e = {"score": int(score), "trace_hash": f"{r.trace_hash:08x}", "edit_len": len(r.edit),
                 "edit_sha256": hashlib.sha256(r.edit.encode()).hexdigest(), "edit_head": r.edit[:64]}
        else:
            score, _, _, _ = pyoracle.fill_rolling(y, x, subst, -11)
            e = {"score": int(score)}
        e.update({"len_y": int(y.size), "len_x": int(x.size), "desc": desc, "oracle_seconds": round(time.time() - t0, 1)})
        res[name] = e
        print(name, e, flush=True)
        with open(OUT, "w") as f:
            json.dump(res, f, indent=1, sort_keys=True)
    if "cfg2" in want:
        x = synth.letters(2001, 16384)
        record("cfg2_random", synth.letters(2002, 16384), x, True, "X seed 2001, Y seed 2002 (independent)")
        record("cfg2_mutated", synth.mutated_copy(x, 2003, 16384), x, True, "X seed 2001, Y = mutated copy of X (seed 2003)")
    if "cfg4" in want:
        record("cfg4", synth.letters(4001, 2048), synth.letters(4002, 4194304), False, "Y seed 4001 (2048), X seed 4002 (4194304), score only")
    if "cfg5a" in want:
        x = synth.letters(5001, 200000)
        record("cfg5_mutated", synth.mutated_copy(x, 5002, 200000), x, True, "X seed 5001, Y = mutated copy of X (seed 5002)")
    if "cfg5b" in want:
        x = synth.letters(5001, 200000)
        record("cfg5_random", synth.letters(5004, 200000), x, True, "X seed 5001, Y seed 5004 (independent)")

if __name__ == "__main__":
    main()
