timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3e_tests_2gpu.log 2>&1; echo tests rc=$?; tail -4 gpurun_out/r3e_tests_2gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
