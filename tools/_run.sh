python -m pytest tests/test_gpu_fill.py tests/test_gpu_big.py -m gpu -x -q 2>&1 | tail -2
ITERS=3 TRACE=1 python tools/quick_bench.py 200000x200000 100000x100000 16384x16384 2>&1 | tail -3 | cut -c40-260
NWB200_ROWS_PER_LANE=8 ITERS=3 TRACE=1 python tools/quick_bench.py 200000x200000 2>&1 | tail -1 | cut -c40-260
ITERS=2 TRACE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/r3l.csv python tools/quick_bench.py 200000x200000 > /dev/null 2>&1; grep "hop_kernel\|walk_kernel\|map_kernel" gpurun_out/r3l.csv | awk -F'","' '{print $5, $NF}' | head -6
