#!/bin/bash
# bench in both matching modes (+ optional ncu launch list of the default run)
TAG=${1:-modes}
OUT=gpurun_out
mkdir -p $OUT
for mode in d i; do
  timeout 600 python bench.py --mode $mode --steps 5 > $OUT/bench_${TAG}_$mode.json 2> $OUT/bench_${TAG}_$mode.err; echo "bench mode $mode exit $?"
  python - <<PY
import json
d=json.load(open("$OUT/bench_${TAG}_$mode.json"))
print("mode $mode value", d["value"], "ms/step", d["ms_per_step"], "e2e", d.get("e2e",{}).get("value"), d["roofline"]["kernel_ms_per_step"], "cand", d["config"]["candidates_per_step"], "pos", d["config"]["filter_positives_per_step"], "matched", d["config"]["matched"], "cpu", d.get("cpu_baseline",{}).get("value"))
PY
  tail -3 $OUT/bench_${TAG}_$mode.err
done
if [ "$2" = "ncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:pgm:: -c 200 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches exit $?"
fi
