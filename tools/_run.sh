python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -3
