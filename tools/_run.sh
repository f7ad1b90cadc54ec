python -m pytest tests/test_gpu_big.py tests/test_gpu_multi.py -m gpu -x -q -k "prefix or cfg4 or scan" 2>&1 | tail -2
python bench.py --workload scan4m --steps 10 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', round(d['value'],1), round(d['ms_per_step'],3), d.get('parity'))"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload scan4m --steps 10 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2', round(d['value'],1), round(d['ms_per_step'],3), d.get('parity'))"
