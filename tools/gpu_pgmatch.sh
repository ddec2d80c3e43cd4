#!/bin/bash
# GPU-box session for stage 7 (exact matches between pseudogenomes): parity tests, then the bench tool.
#   bash tools/gpu_pgmatch.sh <tag> "tests small bench"
TAG=${1:-r02}
WHAT=${2:-"tests small"}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $WHAT " == *" $1 "* ]]; }
if has tests; then
    timeout 900 python -m pytest tests/test_gpu_pgmatch.py -x -q -m gpu > $OUT/pytest_pgmatch_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_pgmatch_$TAG.log
    tail -30 $OUT/pytest_pgmatch_$TAG.log
fi
if has small; then     # C2 x 0.05 with the oracle check, and the CPU reference on the same texts
    timeout 600 python tools/pgmatch_bench.py --workload c2 --scale 0.05 --ref-scale 1.0 --check --fixture tests/golden/pgmatch_fullsize_c2_x0.05.json > $OUT/pgmatch_c2s_$TAG.json 2> $OUT/pgmatch_c2s_$TAG.err; echo "small exit $?"
    cat $OUT/pgmatch_c2s_$TAG.json; tail -5 $OUT/pgmatch_c2s_$TAG.err
fi
if has bench; then     # full C2 text (140 Mbp); reference on a tenth
    timeout 900 python tools/pgmatch_bench.py --workload c2 --ref-scale 0.1 --fixture tests/golden/pgmatch_fullsize_c2.json > $OUT/pgmatch_c2_$TAG.json 2> $OUT/pgmatch_c2_$TAG.err; echo "bench exit $?"
    cat $OUT/pgmatch_c2_$TAG.json; tail -5 $OUT/pgmatch_c2_$TAG.err
fi
