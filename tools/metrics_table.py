#!/usr/bin/env python
"""Per-kernel table of an `ncu --metrics ... --csv` log:  python tools/metrics_table.py gpurun_out/full_metrics_TAG.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
by = collections.OrderedDict()
for r in rows:
    by.setdefault((r[0], r[4].split('(')[0][-30:]), {})[r[-3]] = r[-1]
keys = [('gpu__time_duration.sum', 'ns'), ('dram__sectors_read.sum', 'dramRd'), ('dram__sectors_write.sum', 'dramWr'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('lts__t_sector_hit_rate.pct', 'L2hit%'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'), ('smsp__inst_executed.sum', 'inst'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'), ('launch__registers_per_thread', 'regs'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'longSB'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'barrier'), ('l1tex__data_pipe_lsu_wavefronts.sum', 'L1wf')]
print(f"{'id':>3} {'kernel':30s} " + ' '.join(f'{n:>11s}' for _, n in keys))
for (i, k), m in by.items():
    print(f"{i:>3} {k:30s} " + ' '.join(f"{m.get(x, '-'):>11s}" for x, _ in keys))
