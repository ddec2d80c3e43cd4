#!/bin/bash
# Diagnostics: compute-sanitizer on a tiny run, then ncu --set full on a scaled-down workload (short timeouts).
TAG=${1:-diag}
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_${tool}_$TAG.log 2>&1
  echo "$tool exit $?"; tail -4 $OUT/sanitizer_${tool}_$TAG.log
done
for hints in 0 1; do
  timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'pgm::scan_kernel' -s 2 -c 2 \
      -f -o $OUT/scan_s01_h${hints}_$TAG python bench.py --scale 0.1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --l2-hints $hints > $OUT/ncu_s01_h${hints}_$TAG.log 2>&1
  echo "ncu scale 0.1 hints $hints exit $?"; tail -3 $OUT/ncu_s01_h${hints}_$TAG.log
done
ls -la $OUT | tail -12
