#!/usr/bin/env python
"""Per-region instruction counts of a kernel from an ncu report (source page, needs -lineinfo):
    python tools/ncu_lines.py report.ncu-rep [top_n] [kernel regex]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
kern = sys.argv[3] if len(sys.argv) > 3 else "scan_kernel"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern, "-c", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
hi = his[0]; end = his[1] if len(his) > 1 else len(rows)
hdr = rows[hi]; ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
per = {}; samp = {}
for r in rows[hi + 1:end]:
    if len(r) <= ie or not r[0].isdigit(): continue
    try: n = int(r[ie]); s = int(r[isamp])
    except ValueError: continue
    per[int(r[0])] = per.get(int(r[0]), 0) + n; samp[int(r[0])] = samp.get(int(r[0]), 0) + s
tot = sum(per.values()); stot = sum(samp.values())
print('total warp instructions', tot, 'samples', stot)
src = open('/root/repo/pgrc_b200/csrc/pgm_kernels.cuh').read().split('\n')
marks = {}
pats = [('seed_hash64', 'uint64_t seed_hash64('), ('helpers', 'small helpers'), ('record_of', 'uint4 *record_of('), ('text kernels', 'text packing'),
        ('window_hash', 'uint64_t window_hash('), ('count_groups', 'int count_groups('), ('scan prologue', 'scan_kernel(const __grid_constant__'),
        ('A1', '// ---- A1'), ('A2 produce', '// ---- A2 + B'), ('B consume', 'const uint32_t half = lane & 1u;'), ('chain', '// hot keys: walk'), ('epilogue', '// counters: warp reduce')]
for i, l in enumerate(src, 1):
    for name, pat in pats:
        if pat in l and name not in marks: marks[name] = i
ks = sorted(marks.items(), key=lambda x: x[1])
for (n, a), (n2, b) in zip(ks, ks[1:] + [('end', len(src) + 1)]):
    v = sum(c for k, c in per.items() if a <= k < b); s = sum(c for k, c in samp.items() if a <= k < b)
    print(f'{n:14s} L{a:4d}-{b - 1:4d}: {v:12d} {100 * v / tot:5.1f}%   samples {100 * s / max(1, stot):5.1f}%')
if topn:
    for ln, v in sorted(per.items(), key=lambda x: -x[1])[:topn]:
        print(f'{v:10d} {100 * v / tot:5.1f}% samp {100 * samp[ln] / max(1, stot):5.1f}% L{ln:4d} {src[ln - 1].strip()[:110]}')
