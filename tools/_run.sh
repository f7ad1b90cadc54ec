timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload wave200k --steps 5 --warmup 2 --no-secondary --no-cpu-baseline 2> gpurun_out/r3q.err | grep '^{' | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=8', round(d['value'],1), round(d['ms_per_step'],3), d['roofline'].get('laps_ms_last_step'), d['roofline'].get('traceback'), d.get('parity')); open('gpurun_out/r3q_bench_wave200k_N8.json','w').write(json.dumps(d))"
tail -2 gpurun_out/r3q.err | cut -c1-200
