python -m pytest tests/test_gpu_fill.py tests/test_gpu_big.py -m gpu -x -q 2>&1 | tail -8
