import os, sys, json
sys.path.insert(0, "/root/repo")
import numpy as np
from gpuseqalign_b200 import Engine, Params, synth
from oracle import pyoracle
sc = json.load(open("/root/repo/tests/golden/scoring.json"))
subst = np.array(sc["subst"]["blosum62"], dtype=np.int32)
e = Engine(0); e.set_scoring(subst, -11)
for (n, m) in [(700, 900), (1300, 5000), (3000, 2900), (128, 70), (513, 33)]:
    y = synth.letters(5, n); x = synth.letters(6, m)
    exp = pyoracle.align_pair(y, x, subst, -11, want_hash=False, want_trace=True)
    for K in (1, 2):
        s = e.align(y, x, params=Params(4, 4, 512, K), with_trace=True)
        ed, th = e.trace()
        print(n, m, "K", K, "ok" if (s == exp.score and ed == exp.edit and th == exp.trace_hash) else "MISMATCH", flush=True)
