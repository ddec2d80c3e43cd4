#!/bin/bash
# Final single-GPU round: all GPU tests, bench (both modes, export), ncu launch list, ncu --set full of the scan kernel at full C2 size.
TAG=${1:-r01s}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -4 $OUT/pytest_gpu_$TAG.log
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke_$TAG.log
timeout 900 python bench.py --export > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
timeout 900 python bench.py --mode i > $OUT/bench_${TAG}_mode_i.json 2>> $OUT/bench_$TAG.err; echo "bench mode i exit $?"
timeout 600 python bench.py --workload c1 --no-cpu-baseline > $OUT/bench_${TAG}_c1.json 2>> $OUT/bench_$TAG.err; echo "bench c1 exit $?"
python - <<PY
import json
for f in ("$OUT/bench_$TAG.json", "$OUT/bench_${TAG}_mode_i.json", "$OUT/bench_${TAG}_c1.json"):
    d=json.load(open(f))
    print(f, "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d.get("e2e",{}).get("value"), "frac", d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"], "cpu", d.get("cpu_baseline",{}).get("value"), d.get("export"))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:pgm:: -c 200 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'pgm::scan_kernel' -s 2 -c 2 \
    -f -o $OUT/scan_full_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
ls -la $OUT | tail -12
