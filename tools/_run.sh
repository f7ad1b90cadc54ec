python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -3
NWB200_BATCH_PACKED=0 python tools/batch_one.py 262144 5 | tee gpurun_out/r2w_batch32_ldg.json
NWB200_BATCH_PACKED=0 NWB200_BATCH_TMA=1 python tools/batch_one.py 262144 5 | tee gpurun_out/r2w_batch32_tma.json
for v in ldg tma; do
  if [ $v = tma ]; then export NWB200_BATCH_TMA=1; else unset NWB200_BATCH_TMA; fi
  NWB200_BATCH_PACKED=0 ncu --set full --clock-control none --import-source on -k regex:nw_batch_kernel -c 1 -o gpurun_out/r2w_batch32_$v -f python tools/batch_one.py 131072 1 > gpurun_out/r2w_ncu_$v.log 2>&1; echo ncu $v rc=$?
done
