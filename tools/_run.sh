# scratch script sent to the GPU box by `gpurun -- 'bash tools/_run.sh'` (rewritten per experiment); the validation run of a build:
python -m pytest tests -m gpu -x -q | tail -3
python -c "import __graft_entry__ as g; g.smoke()" | tail -1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo bench rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['roofline']['frac'],4), round(d['e2e']['value'],1), d.get('parity')); print({k:(round(v['value'],1), v.get('parity')) for k,v in d['secondary'].items()})"
