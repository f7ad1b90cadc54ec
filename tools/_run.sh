python -m pytest tests/test_formats_driver.py tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d.get('parity'))"
