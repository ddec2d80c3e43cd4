#!/usr/bin/env python
"""profiles/scan_traffic.json from an explicit-metric ncu pass (CSV log) over the scan kernel — for configs whose footprint makes
an `ncu --set full` capture impractical (config 5: 170 ms launches over a 100 GB footprint, ~40 replays each):
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
        --kernel-name-base demangled -k regex:'pgm::scan_kernel' -s 4 -c 4 --csv --log-file gpurun_out/scan_c5_dram_<tag>.csv python bench.py ...
    python tools/ncu_csv_traffic.py gpurun_out/scan_c5_dram_<tag>.csv c5 [kernels_sha]
Records the mean DRAM read + write bytes per launch with the sha of pgrc_b200/csrc/pgm_* (bench.py reports the figure as
roofline.traffic only while that sha is the one of the sources it runs) and copies the per-launch rows to profiles/."""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0,
        "nsecond": 1e-6, "second": 1e3, "%": 1.0}


def kernels_sha():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "pgrc_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f == "pgm_kernels.cuh":      # the file scan_kernel (and the other stage-4 kernels) is compiled from
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def main():
    path, workload = sys.argv[1], sys.argv[2]
    sha = sys.argv[3] if len(sys.argv) > 3 else kernels_sha()
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    per = {}
    for r in rows:
        k = per.setdefault(r["ID"], {"kernel": r["Kernel Name"][:90]})
        k[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    launches = list(per.values())
    tr = [l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in launches]
    out = {"source": os.path.basename(path), "workload": workload, "kernels_sha": sha, "launches": launches,
           "dram_bytes_per_launch_mean": sum(tr) / len(tr)}
    tag = os.path.splitext(os.path.basename(path))[0]
    json.dump(out, open(os.path.join(ROOT, "profiles", tag + ".json"), "w"), indent=1)
    p = os.path.join(ROOT, "profiles", "scan_traffic.json")
    t = json.load(open(p)) if os.path.exists(p) else {}
    t[workload] = {"dram_bytes_per_launch": int(sum(tr) / len(tr)), "kernels_sha": sha,
                   "source": f"profiles/{tag}.json, mean of {len(tr)} launches (explicit-metric pass)"}
    json.dump(t, open(p, "w"), indent=1)
    print(json.dumps(t[workload]))


if __name__ == "__main__":
    main()
