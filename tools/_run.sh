python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -3
for v in 1 7; do
NWB200_BATCH_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2q_bench_v$v.json 2> gpurun_out/r2q_bench_v$v.err; echo rc=$?
done
python -c "
import json
for v in (1,7):
    d=json.loads(open('gpurun_out/r2q_bench_v%d.json'%v).read().strip().splitlines()[-1]); print(v, d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'], d['e2e']['value'], d.get('parity'))
"
