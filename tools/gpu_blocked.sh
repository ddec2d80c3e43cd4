#!/bin/bash
# L2-blocked scan pipeline: forced-path parity tests, then bench with the pipeline on (auto) and off.
TAG=${1:-blk}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "blocked" > $OUT/pytest_blocked_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_blocked_$TAG.log
tail -25 $OUT/pytest_blocked_$TAG.log
for mode in 1 0; do
  PGM_BLOCKED_SCAN=$mode timeout 600 python bench.py --no-cpu-baseline --steps 5 > $OUT/bench_${TAG}_m$mode.json 2> $OUT/bench_${TAG}_m$mode.err; echo "bench mode $mode exit $?"
  python - <<PY
import json
d=json.load(open("$OUT/bench_${TAG}_m$mode.json"))
print("mode $mode value", d["value"], "ms/step", d["ms_per_step"], "e2e", d.get("e2e",{}).get("value"), d["roofline"]["kernel_ms_per_step"], "cand", d["config"]["candidates_per_step"], "pos", d["config"]["filter_positives_per_step"], "matched", d["config"]["matched"], "frac", d["roofline"]["frac"])
PY
  tail -3 $OUT/bench_${TAG}_m$mode.err
done
