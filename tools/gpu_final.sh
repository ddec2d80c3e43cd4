#!/bin/bash
# Last GPU-box session of a round: the whole GPU test suite, the default bench with every leg, the DRAM bytes of the scan at
# config 5 (ncu, explicit metrics) and the reference arm.    bash tools/gpu_final.sh <tag> "tests bench ncuc5 ref launches"
TAG=${1:-r02}
WHAT=${2:-"tests bench ncuc5 ref"}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $WHAT " == *" $1 "* ]]; }
python -c "import bench; print(bench.kernels_sha())" > $OUT/kernels_sha_$TAG.txt 2>/dev/null; cat $OUT/kernels_sha_$TAG.txt
if has tests; then
    timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu_$TAG.log
    tail -6 $OUT/pytest_gpu_$TAG.log
fi
if has smoke; then
    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -8 $OUT/smoke_$TAG.log
fi
if has bench; then
    timeout 1200 python bench.py --verify > $OUT/bench_c5_n1_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
    python tools/show_bench.py $OUT/bench_c5_n1_$TAG.json; tail -3 $OUT/bench_$TAG.err
fi
if has ncuc5; then
    timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
        --kernel-name-base demangled -k regex:'pgm::scan_kernel' -s 4 -c 4 --csv --log-file $OUT/scan_c5_dram_$TAG.csv \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_c5_$TAG.log 2>&1
    echo "ncu c5 exit $?"; tail -5 $OUT/scan_c5_dram_$TAG.csv | cut -c1-300
fi
if has launches; then
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:pgm:: -c 500 --csv \
        --log-file $OUT/launches_c5_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_launches_$TAG.log 2>&1
    echo "ncu launches exit $?"
fi
if has ref; then
    timeout 1200 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; echo "ref exit $?"; cut -c1-1500 $OUT/bench_ref_$TAG.json
fi
