python -m pytest tests/test_gpu_big.py -m gpu -x -q -k "prefix or cfg4" > gpurun_out/r2o_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r2o_tests.log
NWB200_SCAN_DBG=1 python bench.py --workload scan4m --steps 5 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2o_bench_scan.json 2> gpurun_out/r2o_bench_scan.err; echo rc=$?
grep "scan dbg" gpurun_out/r2o_bench_scan.err | tail -32 | grep "warp  0\|warp 15" | cut -c1-150
python -c "
import json
for v in ('scan',):
    d=json.loads(open('gpurun_out/r2o_bench_%s.json'%v).read().strip().splitlines()[-1]); print(v, d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('parity'))
"
