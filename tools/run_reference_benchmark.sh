#!/bin/bash
# Runs the reference's OWN benchmark executable (stock benchmark.cpp / cmd_parser.cpp / every stock algorithm) with the
# B200 engine registered as one more NwAlgorithm entry, on a reference pair list.  The executable cross-checks
# NwAlign_B200 against NwAlign_Cpu4_Mt_DiagRow and gpu9 itself (benchmark.cpp:120-147) and exits non-zero on any
# mismatch.  usage: tools/run_reference_benchmark.sh [pair file under resrc/] [extra flags]
set -e
HERE="$(cd "$(dirname "$0")/.." && pwd)"
PAIRS="${1:-pair_debug.txt}"; shift || true
cd "$HERE/oracle/_ref"
./nw_b200 -b resrc/subst.json -r "$HERE/gpuseqalign_b200/plugin/param_b200.json" -s resrc/seq_generated.fa -p "resrc/$PAIRS" \
    --fCalcTrace --fCalcScoreHash --fWriteProgress -o "$HERE/gpurun_out/ref_bench_${PAIRS%.txt}.tsv" "$@"
