# scratch script sent to the GPU box by `gpurun -- 'bash tools/_run.sh'` (rewritten per experiment); the validation run of a build:
python -m pytest tests -m gpu -x -q | tail -3
python -c "import __graft_entry__ as g; g.smoke()" | tail -1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo bench rc=$?
