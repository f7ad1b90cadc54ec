run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/wave_trace_check.py "$@" 2> gpurun_out/r3h.err | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('R=$NWB200_ROWS_PER_LANE', 'block', d['block'], 'fill rank0', d['fill_ms'], 'fill wall', d['fill_wall_ms'], 'gather+trace wall', d['gather_trace_wall_ms'], 'trace', d['trace_calc_ms'], d['trace_hash'])
"; }
export NWB200_ROWS_PER_LANE=8; run 200000 random,random 50176
export NWB200_ROWS_PER_LANE=16; run 200000 random,random 50176
