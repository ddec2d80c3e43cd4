#!/usr/bin/env python
"""Prints the essentials of bench.py JSON lines:  python tools/show_bench.py gpurun_out/bench_*.json"""
import json
import sys

for f in sys.argv[1:]:
    d = None
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line)
    print("==", f)
    if d is None:
        print("   no JSON line")
        continue
    e = d.get("e2e") or {}
    print("   N=%d  %.4g reads/s  %.2f ms/step   e2e %.4g reads/s (%s ms)   roofline frac %s" % (
        d["n_gpus"], d["value"], d["ms_per_step"], e.get("value", 0), e.get("ms_per_step"), d.get("roofline", {}).get("frac")))
    print("   parallelism:", d["config"].get("parallelism"))
    print("   kernels ms/step:", d.get("roofline", {}).get("kernel_ms_per_step"))
    if d.get("verify"):
        print("   verify:", d["verify"])
    det = d.get("detail") or {}
    print("   detail:", {k: det.get(k) for k in ("matched", "route", "candidates_per_step", "filter_positives_per_step")})
    if d.get("cpu_baseline"):
        print("   cpu_baseline:", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:120])
