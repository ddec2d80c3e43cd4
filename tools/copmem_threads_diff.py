#!/usr/bin/env python
"""How much does the reference's mode c (CopMEMReadsApproxMatcher) depend on -t?  Its multithreaded index build
(CopMEMMatcher::processRefMultithreaded, CopMEMMatcher.cpp:271-324) orders the hash buckets differently from the serial one
and races on counts[] (:303); the GPU path reproduces the serial build (-t 1).  This script runs the UNMODIFIED reference
classes (oracle/_ref) on one workload at several thread counts and reports what differs.  CPU only.
    python tools/copmem_threads_diff.py [workload] [scale] > profiles/copmem_threads_r02.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from pgrc_b200 import synth  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
c = synth.scaled_config(w, scale)
p = synth.hashed_params(**c, seed=20261017)
text = synth.hashed_text(p).numpy()
packed = synth.hashed_reads(p, 0, c["n_reads"]).numpy()
asc = synth.unpack_reads_ascii(packed, c["read_len"])
print(f"workload {w} x {scale}: {c['n_reads']} reads x {c['read_len']} bp vs {text.size} bases, reference mode c")
base = None
for t in (1, 1, 2, 4, 8, 8):
    r = oracle.ref_map_reads(text, asc, None, c["read_len"], mode="c", threads=t)
    if base is None:
        base = r
    dp = int((r.pos != base.pos).sum()); dm = int((r.mm != base.mm).sum()); dr = int((r.rc != base.rc).sum())
    print(f"-t {t}: matched {r.matched}  sum of mismatches {int(r.mm[r.mm != 255].astype(np.int64).sum())}  {r.seconds:.1f} s   "
          f"vs the first -t 1 run: {dp} positions, {dm} mismatch counts, {dr} strands differ ({100.0 * dp / len(r.pos):.3f} % of the reads)")
o = oracle.oracle_map_reads(text, packed, None, c["read_len"], mode="c")
print(f"C oracle (serial build, what the GPU path reproduces): {int((o.pos != base.pos).sum())} positions differ from the reference at -t 1")
