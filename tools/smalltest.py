import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpuseqalign_b200 import Engine, Params, synth
sc = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scoring.json")))
subst = np.array(sc["subst"]["blosum62"], dtype=np.int32)
e = Engine(0); e.set_scoring(subst, -11)
n, m = int(sys.argv[1]), int(sys.argv[2])
y = synth.letters(5, n); x = synth.letters(6, m)
print("score", e.align(y, x, keep_headers=(len(sys.argv) > 3)))
