#!/bin/bash
# GPU-box session for the partitioned exact pre-filter (pgm_part.cuh).
#   bash tools/gpu_part.sh <tag> "tests c5 c4 c5off e2e sweep"
TAG=${1:-r02}
WHAT=${2:-"tests c5"}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $WHAT " == *" $1 "* ]]; }
if has tests; then
    timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "partitioned_prefilter" > $OUT/pytest_part_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_part_$TAG.log
    tail -30 $OUT/pytest_part_$TAG.log
fi
if has c5; then
    timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --verify > $OUT/bench_c5_n1_part_$TAG.json 2> $OUT/bench_part_$TAG.err; echo "c5 exit $?"
    python tools/show_bench.py $OUT/bench_c5_n1_part_$TAG.json; tail -5 $OUT/bench_part_$TAG.err
fi
if has c5off; then
    PGM_PART_SCAN=0 timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > $OUT/bench_c5_n1_nopart_$TAG.json 2>> $OUT/bench_part_$TAG.err; echo "c5off exit $?"
    python tools/show_bench.py $OUT/bench_c5_n1_nopart_$TAG.json
fi
if has c4; then
    timeout 900 python bench.py --workload c4 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --verify > $OUT/bench_c4_n1_part_$TAG.json 2>> $OUT/bench_part_$TAG.err; echo "c4 exit $?"
    python tools/show_bench.py $OUT/bench_c4_n1_part_$TAG.json
fi
if has sweep; then    # partition size
    for mb in 24 96; do
        PGM_PART_MB=$mb timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > $OUT/bench_c5_n1_part${mb}mb_$TAG.json 2>> $OUT/bench_part_$TAG.err
        echo "PGM_PART_MB=$mb"; python tools/show_bench.py $OUT/bench_c5_n1_part${mb}mb_$TAG.json
    done
fi
if has e2e; then
    timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --verify > $OUT/bench_c5_n1_part_e2e_$TAG.json 2>> $OUT/bench_part_$TAG.err; echo "e2e exit $?"
    python tools/show_bench.py $OUT/bench_c5_n1_part_e2e_$TAG.json
fi
