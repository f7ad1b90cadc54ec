python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -8
python - <<'PY'
import numpy as np, json, time, sys
sys.path.insert(0,'.')
from gpuseqalign_b200 import Engine, synth
subst=np.array(json.load(open('tests/golden/scoring.json'))["subst"]["blosum62"],dtype=np.int32)
e=Engine(0); e.set_scoring(subst,-11)
pool,oy,ly,ox,lx=synth.batch_pairs(0,131072,256,256)
e.upload_batch(pool,oy,ly,ox,lx)
for v in ("nw_affine","sw_affine","sw_linear"):
    for _ in range(3):
        e.batch_resident_variant(v,-11,-1); s=e.fetch_batch_scores()
    ms=e.timing()["align_calc"]; print(v, round(ms,3),"ms", round(131072*65536/ms/1e6,1),"GCUPS", int(s[:3].sum()))
e.batch_resident(); s=e.fetch_batch_scores(); print("nw_linear", round(e.timing()["align_calc"],3))
PY
