#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`):  python tools/launch_table.py <csv> [out.txt]
The times are cold-cache and serialised (profiler): the kernels' SHARES of a step are what is compared with bench.py's
CUDA-event figures."""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = csv.DictReader(rows)
tot, cnt = defaultdict(float), defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"<.*", "", name).replace("pgm::", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
lines = [f"{'kernel':32s} {'launches':>8s} {'ms':>12s} {'share':>7s}"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    lines.append(f"{k:32s} {cnt[k]:8d} {v:12.3f} {100 * v / total:6.1f}%")
lines.append(f"{'total':32s} {sum(cnt.values()):8d} {total:12.3f}")
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
