python bench.py > gpurun_out/r3p_bench_default.json 2> gpurun_out/r3p_bench_default.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3p_bench_default.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['e2e']['value'],1), round(d['e2e']['packed5']['value'],1), d.get('parity'))
for k,v in d.get('secondary',{}).items(): print('  ',k, round(v.get('value'),1), round(v.get('ms_per_step'),3), v.get('parity'))
PY
