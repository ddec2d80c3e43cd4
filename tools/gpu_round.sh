#!/bin/bash
# One GPU-box session: parity tests, bench, ncu launch list, ncu full capture of the scan + build kernels.
# Usage (through gpurun): bash tools/gpu_round.sh [tag] [quick]
TAG=${1:-r01}
MODE=${2:-full}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -15 $OUT/pytest_gpu_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
if [ "$MODE" = "quick" ]; then exit 0; fi
timeout 600 python bench.py --workload c1 --no-cpu-baseline > $OUT/bench_c1_$TAG.json 2>> $OUT/bench_$TAG.err
cat $OUT/bench_c1_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:pgm:: -c 200 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'pgm::(scan_kernel|build_table_kernel)' -s 3 -c 3 \
    -f -o $OUT/scan_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
ls -la $OUT
