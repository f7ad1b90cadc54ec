python -m pytest tests -m gpu -x -q > gpurun_out/r2p_tests.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/r2p_tests.log
( time python bench.py > gpurun_out/r2p_bench_default.json 2> gpurun_out/r2p_bench_default.err ) 2>&1 | tail -4; echo bench rc=$?
tail -c 600 gpurun_out/r2p_bench_default.err
