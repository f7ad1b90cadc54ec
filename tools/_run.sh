timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3m_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r3m_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r3m_bench_default.json 2> gpurun_out/r3m_bench_default.err; echo bench rc=$?
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r3m_bench_N2.json 2> gpurun_out/r3m_bench_N2.err; echo bench2 rc=$?
python - <<'PY'
import json
for f in ('gpurun_out/r3m_bench_default.json','gpurun_out/r3m_bench_N2.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['e2e']['value'],1), round(d['e2e']['packed5']['value'],1), d.get('parity'))
    for k,v in d.get('secondary',{}).items(): print('  ',k, round(v.get('value'),1), round(v.get('ms_per_step'),3), v.get('laps_ms_last_step'), v.get('parity'))
PY
