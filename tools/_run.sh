python -m pytest tests/test_gpu_big.py -m gpu -x -q -k "prefix or cfg4 or scan" 2>&1 | tail -3
for w in 16 8 4; do
NWB200_SCAN_WARPS=$w python bench.py --workload scan4m --steps 5 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2u_scan_w$w.json 2> gpurun_out/r2u_scan_w$w.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r2u_scan_w$w.json').read().strip().splitlines()[-1]); print($w, d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('parity'))"
done
