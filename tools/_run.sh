timeout 600 python -m pytest tests/test_gpu_fill.py tests/test_gpu_big.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
ITERS=3 TRACE=1 python tools/quick_bench.py 200000x200000 100000x100000 60000x300000 2>&1 | tail -3 | cut -c40-260
