#!/bin/bash
# single-GPU profiling of the routed kernels at the per-rank sizes of an 8-GPU config-5 run (tools/route_one_gpu.py)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02i}
for fb in ${FBS:--1}; do
    FILTER_BITS=$fb timeout 600 python tools/route_one_gpu.py 2 0.25 2 > $OUT/route1gpu_fb${fb}_$TAG.txt 2>&1; tail -4 $OUT/route1gpu_fb${fb}_$TAG.txt
done
if [ -n "$RW" ]; then ROUND_WINDOWS=134217728 timeout 600 python tools/route_one_gpu.py 2 0.25 2 > $OUT/route1gpu_rw128_$TAG.txt 2>&1; tail -3 $OUT/route1gpu_rw128_$TAG.txt; fi
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'route_filter_kernel|route_probe_kernel' -s 8 -c 4 \
    -f -o $OUT/route_kernels_$TAG python tools/route_one_gpu.py 2 0.25 1 > $OUT/ncu_route_$TAG.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/ncu_route_$TAG.log
