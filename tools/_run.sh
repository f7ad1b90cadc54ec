python -m pytest tests/test_gpu_big.py tests/test_gpu_fill.py -m gpu -x -q 2>&1 | tail -3
ITERS=4 TRACE=1 CONFIGS="0,0,0,0;16,4,2,512;8,4,2,512" python tools/quick_bench.py 200000x200000 300000x300000 2>&1 | tail -6
ITERS=4 TRACE=0 CONFIGS="16,4,2,256" python tools/quick_bench.py 16384x16384 100000x100000 2>&1 | tail -2
