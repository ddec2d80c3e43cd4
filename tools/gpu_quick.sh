#!/bin/bash
# Quick GPU check: parity tests + bench (+ optional small-scale ncu of the scan kernel with source counters)
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -8 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d.get("e2e",{}).get("value"), d["roofline"]["kernel_ms_per_step"], "cand", d["config"]["candidates_per_step"], "pos", d["config"]["filter_positives_per_step"])
PY
tail -3 $OUT/bench_$TAG.err
if [ "$2" = "ncu" ]; then
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'pgm::(scan_kernel|build_table_kernel)' -s 3 -c 3 \
    -f -o $OUT/scan_s02_$TAG python bench.py --scale 0.2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_s02_$TAG.log 2>&1
echo "ncu scale 0.2 exit $?"; tail -2 $OUT/ncu_s02_$TAG.log | cut -c1-300
fi
