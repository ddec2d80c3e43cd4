#!/bin/bash
# Full-size (C2) ncu capture with an explicit metric list (few replay passes) — the numbers behind roofline.traffic.
TAG=${1:-ncu}
KREGEX=${2:-pgm::(build_|scan_kernel|resolve_kernel|unpack_reads)}
SKIP=${3:-16}
COUNT=${4:-16}
OUT=gpurun_out
mkdir -p $OUT
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__sectors_read.sum,dram__sectors_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_atom.sum,lts__t_sectors_srcunit_tex_op_red.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
timeout 420 ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:"$KREGEX" -s $SKIP -c $COUNT --csv --log-file $OUT/full_metrics_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e $BENCH_ARGS > $OUT/ncu_fullsize_$TAG.log 2>&1
echo "ncu full-size metrics exit $?"; tail -2 $OUT/ncu_fullsize_$TAG.log | cut -c1-200
