run() { timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/wave_trace_check.py "$@" >> gpurun_out/r2s_wavetrace.jsonl 2> gpurun_out/r2s_wavetrace.err; echo "rc=$? ($*)"; grep -h "NwB200Error\|Error:" gpurun_out/r2s_wavetrace.err | head -3; }
rm -f gpurun_out/r2s_wavetrace.jsonl
run 40000 random,mutated,long_indel 5120,2048
NWB200_CORRIDOR=512 run 40000 random,long_indel 5120
run 200000 random,random 100352,25088
cat gpurun_out/r2s_wavetrace.jsonl | cut -c1-300
