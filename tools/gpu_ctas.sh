#!/bin/bash
# scan kernel compiled for 5 / 6 resident CTAs per SM (libpgrc_gpu_c5.so, _c6.so built beforehand) vs the default 4
OUT=gpurun_out; mkdir -p $OUT
cp pgrc_b200/libpgrc_gpu.so /tmp/base.so
for n in 4 5 6; do
  if [ $n = 4 ]; then cp /tmp/base.so pgrc_b200/libpgrc_gpu.so; else cp pgrc_b200/libpgrc_gpu_c$n.so pgrc_b200/libpgrc_gpu.so; fi
  for c in $n; do
  timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 5 --ctas-per-sm $c > $OUT/bench_ctas_$n.json 2> $OUT/bench_ctas_$n.err; echo "bench exit $?"
  python - <<PY
import json
d=json.load(open("$OUT/bench_ctas_$n.json"))
print("compiled for $n, grid $c/SM: value", d["value"], "ms/step", d["ms_per_step"], "scan", d["roofline"]["kernel_ms_per_step"]["scan"], "matched", d["config"]["matched"])
PY
  done
done
cp /tmp/base.so pgrc_b200/libpgrc_gpu.so
