python -m pytest tests/test_gpu_big.py -m gpu -x -q -k "prefix or cfg4 or scan" 2>&1 | tail -2
for cfg in "4 6" "3 4" "2 3" "4 5" "2 2" "6 8" "3 5" "1 2"; do set -- $cfg
NWB200_SCAN_AHEAD=$1 NWB200_SCAN_LAG=$2 timeout 120 python bench.py --workload scan4m --steps 5 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r3d_scan.json 2> gpurun_out/r3d_scan.err
python -c "
import json
d=json.loads(open('gpurun_out/r3d_scan.json').read().strip().splitlines()[-1]); print('ahead $1 lag $2', round(d['value'],1), round(d['ms_per_step'],3), d.get('parity'))"
done
