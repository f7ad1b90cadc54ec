python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -2
for t in 0 1; do
if [ $t = 1 ]; then export NWB200_BATCH_TICKETS=1; else unset NWB200_BATCH_TICKETS; fi
python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r3g_bench_t$t.json 2> gpurun_out/r3g_bench_t$t.err
python -c "
import json
d=json.loads(open('gpurun_out/r3g_bench_t$t.json').read().strip().splitlines()[-1]); print('tickets=$t', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['e2e']['value'],1), d['e2e']['packed5']['value'], d.get('parity'))"
done
