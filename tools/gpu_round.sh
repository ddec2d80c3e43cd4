#!/bin/bash
# One GPU-box session (through gpurun): what each step does is selected by words in $2.
#   bash tools/gpu_round.sh <tag> "routed tests bench ncu"
TAG=${1:-r02}
WHAT=${2:-"tests bench"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
(nproc; free -g | head -2) >> $OUT/gpu_$TAG.txt 2>&1
python -c "import bench; print(bench.kernels_sha())" > $OUT/kernels_sha_$TAG.txt 2>/dev/null
has() { [[ " $WHAT " == *" $1 "* ]]; }
if has routed; then    # the new multi-GPU scheme first (several contexts on this one GPU): fail fast
    timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "routed or sharded" > $OUT/pytest_routed_$TAG.log 2>&1; echo "pytest routed exit $?" | tee -a $OUT/pytest_routed_$TAG.log
    tail -25 $OUT/pytest_routed_$TAG.log
fi
if has tests; then
    timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu_$TAG.log
    tail -25 $OUT/pytest_gpu_$TAG.log
fi
if has bench; then     # the default bench (C5, N = 1) with the full-size parity checks
    timeout 1200 python bench.py --verify --steps 5 --warmup 3 > $OUT/bench_c5_n1_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
    cat $OUT/bench_c5_n1_$TAG.json; tail -5 $OUT/bench_$TAG.err
fi
if has benchall; then  # the other configs at N = 1, with their fixtures
    for w in c1 c2 c3 c4; do
        timeout 900 python bench.py --workload $w --verify --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_${w}_n1_$TAG.json 2>> $OUT/bench_$TAG.err
        cat $OUT/bench_${w}_n1_$TAG.json
    done
fi
if has ref; then
    timeout 1200 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; cat $OUT/bench_ref_$TAG.json
fi
if has c4; then
    timeout 900 python bench.py --workload c4 --verify --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_c4_n1_$TAG.json 2>> $OUT/bench_$TAG.err
    cat $OUT/bench_c4_n1_$TAG.json
fi
if has slices; then    # sweep of the hash-sliced filter at C5 (PGM_FILTER_SLICES = log2 of the slice count)
    for sb in 0 1 2 3; do
        PGM_FILTER_SLICES=$sb timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $OUT/bench_c5_slices${sb}_$TAG.json 2>> $OUT/bench_$TAG.err
        python tools/show_bench.py $OUT/bench_c5_slices${sb}_$TAG.json
    done
fi
if has ncu; then
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:pgm:: -c 400 --csv \
        --log-file $OUT/launches_c5_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_launches_$TAG.log 2>&1
    echo "ncu launches exit $?"
fi
if has ncufull; then   # full captures of the scan kernel: C2 (both passes of one step) and C4 (C5's footprint makes the replays' save/restore slow)
    for w in ${NCU_WORKLOADS:-c2 c4}; do
        timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'pgm::scan_kernel' -s 2 -c 2 \
            -f -o $OUT/scan_${w}_$TAG python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_full_${w}_$TAG.log 2>&1
        echo "ncu full $w exit $?"
    done
fi
if has ncuc5; then     # DRAM bytes of the scan launches at C5 (two metrics only: a full capture would replay 170 ms launches over a 100 GB footprint)
    timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
        --kernel-name-base demangled -k regex:'pgm::scan_kernel' -s 4 -c 4 --csv --log-file $OUT/scan_c5_dram_$TAG.csv \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_c5_$TAG.log 2>&1
    echo "ncu c5 exit $?"; tail -6 $OUT/scan_c5_dram_$TAG.csv
fi
if has c2; then
    timeout 900 python bench.py --workload c2 --verify --steps 20 --warmup 5 > $OUT/bench_c2_n1_$TAG.json 2>> $OUT/bench_$TAG.err
    cat $OUT/bench_c2_n1_$TAG.json
fi
ls -la $OUT | tail -30
