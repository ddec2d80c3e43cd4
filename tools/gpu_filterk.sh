#!/bin/bash
# adaptive filter: tests, then C2 bench with k = auto / 2 / 3 / 4, then C4 on one GPU
TAG=${1:-fk}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -4 $OUT/pytest_gpu_$TAG.log
for k in 0 2 3 4; do
  PGM_FILTER_K=$k timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 5 > $OUT/bench_${TAG}_k$k.json 2> $OUT/bench_${TAG}_k$k.err; echo "bench k=$k exit $?"
  python - <<PY
import json
d=json.load(open("$OUT/bench_${TAG}_k$k.json"))
print("k=$k value", d["value"], "ms/step", d["ms_per_step"], d["roofline"]["kernel_ms_per_step"]["scan"], d["roofline"]["kernel_ms_per_step"]["build_table"], "pos", d["config"]["filter_positives_per_step"])
PY
done
bash tools/gpu_big.sh $TAG c4 1 --no-e2e
