python -m pytest tests/test_gpu_big.py tests/test_gpu_fill.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload wave200k --steps 5 --warmup 2 --no-secondary --no-cpu-baseline > gpurun_out/r2z_wave.json 2> gpurun_out/r2z_wave.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r2z_wave.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline'].get('laps_ms_last_step'), d['roofline'].get('traceback'), d['e2e']['value'], d.get('parity'))"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/wave_trace_check.py 200000 random 100352 2> gpurun_out/r2z_wt.err | cut -c1-330; echo rc=$?
