python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; echo rc=$?; tail -2 gpurun_out/r3a_bench.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r3a_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac']); print(d['e2e'])"
