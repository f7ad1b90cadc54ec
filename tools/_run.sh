python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r2f_tests.log
python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2f_bench_v1.json 2> gpurun_out/r2f_bench_v1.err; echo rc=$?
python -c "
import json
for v in ('v1',):
    d=json.loads(open('gpurun_out/r2f_bench_%s.json'%v).read().strip().splitlines()[-1]); print(v, d['value'], d['roofline']['frac'], d['roofline']['kernel'], d['e2e']['value'], d.get('parity'))
"
ncu --set full --import-source on --clock-control none -k regex:nw_batch3 --launch-skip 3 --launch-count 1 -o gpurun_out/r2f_batch3 python bench.py --pairs 131072 --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2f_ncu.log 2>&1; echo ncu rc=$?
