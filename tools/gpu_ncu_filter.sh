#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
M=gpu__time_duration.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
for fb in 28 27 26 24; do
timeout 300 ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:'pgm::scan_kernel' -s 2 -c 1 --csv --log-file $OUT/filt_$fb.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --filter-bits $fb > $OUT/filt_$fb.log 2>&1
echo "fb $fb exit $?"
grep -E "scan_kernel" $OUT/filt_$fb.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | tr '\n' ';'; echo
done
