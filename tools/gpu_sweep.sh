#!/bin/bash
# Tuning sweep on the full C2 workload: prints per-kernel ms for each setting.
TAG=${1:-sweep}
OUT=gpurun_out
mkdir -p $OUT
run() {
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 2 "$@" > $OUT/tmp_sweep.json 2>> $OUT/sweep_$TAG.err
  python - "$@" <<PY
import json,sys
try:
    d=json.load(open("$OUT/tmp_sweep.json"))
    k=d["roofline"]["kernel_ms_per_step"]
    print(" ".join(sys.argv[1:]), "| ms/step", d["ms_per_step"], "scan", k["scan"], "build", k["build_table"], "pos", d["config"]["filter_positives_per_step"], "cand", d["config"]["candidates_per_step"])
except Exception as e:
    print(" ".join(sys.argv[1:]), "FAILED", e)
PY
}
for a in "$@"; do :; done
run --l2-hints 1
run --l2-hints 2
run --l2-hints 0
run --l2-hints 2 --filter-bits 29
run --l2-hints 2 --filter-bits 27
PGM_L2_FETCH=32 run --l2-hints 1
PGM_L2_FETCH=32 run --l2-hints 2
PGM_L2_FETCH=128 run --l2-hints 1
