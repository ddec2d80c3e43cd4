#!/usr/bin/env python
"""Summarise an `ncu --set full` capture (read here, no GPU needed) into profiles/:
    python tools/ncu_summary.py gpurun_out/scan_r01a.ncu-rep r01a [workload] [kernel label]
writes profiles/<label>_full_<tag>.json (selected counters per launch) and, for the scan kernel, updates
profiles/scan_traffic.json (dram read+write bytes per launch, averaged over the captured launches, together with the
sha of the kernel sources the capture was taken with: bench.py reports it as roofline.traffic only while that sha is
the one of the sources it runs — run this BEFORE editing pgrc_b200/csrc/pgm_* again, or pass the sha as 5th argument)."""
import hashlib
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes_read.sum.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__t_sectors.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sectors_srcunit_tex_op_atom.sum',
        'lts__t_sectors_srcunit_tex_op_red.sum', 'l1tex__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    workload = sys.argv[3] if len(sys.argv) > 3 else "c2"
    label = sys.argv[4] if len(sys.argv) > 4 else "scan"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out, traffic = [], []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append({k: (d[k] + ' ' + units[hdr.index(k)]).strip() for k in KEEP if k in d})
        traffic.append(to_bytes(d['dram__bytes_read.sum'], units[hdr.index('dram__bytes_read.sum')]) +
                       to_bytes(d['dram__bytes_write.sum'], units[hdr.index('dram__bytes_write.sum')]))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump({"source": f"ncu --set full --clock-control none --import-source on ({os.path.basename(rep)}), workload {workload}",
               "launches": out}, open(os.path.join(ROOT, "profiles", f"{label}_full_{tag}.json"), "w"), indent=1)
    if label == "scan":
        p = os.path.join(ROOT, "profiles", "scan_traffic.json")
        t = json.load(open(p)) if os.path.exists(p) else {}
        h = hashlib.sha256()
        csrc = os.path.join(ROOT, "pgrc_b200", "csrc")
        for f in sorted(os.listdir(csrc)):
            if f == "pgm_kernels.cuh":      # the file scan_kernel (and the other stage-4 kernels) is compiled from
                h.update(open(os.path.join(csrc, f), "rb").read())
        t[workload] = {"dram_bytes_per_launch": int(sum(traffic) / len(traffic)),
                       "kernels_sha": sys.argv[5] if len(sys.argv) > 5 else h.hexdigest()[:16],
                       "source": f"profiles/scan_full_{tag}.json, mean of {len(traffic)} launches"}
        json.dump(t, open(p, "w"), indent=1)
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
