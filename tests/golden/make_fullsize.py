#!/usr/bin/env python
"""Generates the full-size parity fixtures of tests/golden/ (run in the build container, CPU only).

The bench workloads come from the counter-based generator (pgrc_b200/csrc/pgs_synth.cu): the same parameters
give the same bytes on this host and on the B200 box.  So results computed HERE travel as small fixtures:

  sample_<cfg>.npz        N_SAMPLE reads (every stride-th read of the workload) matched ALONE by the C oracle
                          (oracle/pgrc_oracle.c, pinned against the reference) against the FULL text of the config:
                          a read's result does not depend on the other reads (SURVEY.md §8 a-R), so the full GPU run
                          must report exactly these (pos, rc, mm) at these read indices.  `bench.py --verify` and
                          tests/test_gpu_fullsize.py compare.  Modes d (all configs) and c (--mode-c configs).
  fullsize_<cfg>_ref.json the UNMODIFIED reference classes (oracle/_ref, DefaultReadsApproxMatcher, mode d) on the
                          WHOLE workload: matched count, per-mismatch histogram and sha256 of the three result
                          arrays.  Feasible for c1-c3 (the reference is serial: minutes); c4/c5 would take hours and
                          more memory than this host has.

usage: python tests/golden/make_fullsize.py [--sample c1 c2 c3 c4 c5] [--ref c1 c2 c3] [--mode-c c1 c2]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

SEED = 20261017          # bench.py's workload seed
N_SAMPLE = 20000
MATCH = dict(seed=38, min_chars_per_mismatch=3)


def sample_indices(n_reads: int) -> np.ndarray:
    n = min(N_SAMPLE, n_reads)
    return (np.arange(n, dtype=np.int64) * (n_reads // n)).astype(np.int64)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sample", nargs="*", default=[])
    ap.add_argument("--ref", nargs="*", default=[])
    ap.add_argument("--mode-c", nargs="*", default=[])
    ap.add_argument("--threads", type=int, default=1)
    args = ap.parse_args()
    import oracle
    from pgrc_b200 import synth
    for cfg in sorted(set(args.sample) | set(args.ref) | set(args.mode_c)):
        c = synth.scaled_config(cfg, 1.0)
        p = synth.hashed_params(**c, seed=SEED)
        L = c["read_len"]
        t0 = time.time()
        text = synth.hashed_text(p).numpy()
        print(f"[{cfg}] text {text.size} bases in {time.time() - t0:.0f} s", flush=True)
        idx = sample_indices(c["n_reads"])
        sample = synth.hashed_reads_at(p, idx)
        for mode, wanted in (("d", cfg in args.sample), ("c", cfg in args.mode_c)):
            if not wanted:
                continue
            t0 = time.time()
            r = oracle.oracle_map_reads(text, sample, None, L, mode=mode, **MATCH)
            print(f"[{cfg}] oracle mode {mode}: {len(idx)} sampled reads vs the full text in {time.time() - t0:.0f} s, matched {r.matched}", flush=True)
            name = f"sample_{cfg}.npz" if mode == "d" else f"sample_{cfg}_mode_{mode}.npz"
            np.savez_compressed(os.path.join(HERE, name), idx=idx, pos=r.pos, rc=r.rc, mm=r.mm, seed=SEED, mode=np.bytes_(mode),
                                genome_len=c["genome_len"], n_reads=c["n_reads"], read_len=L, text_len=int(p.text_len),
                                err_q24=int(p.err_q24), contig=int(p.contig))
        if cfg in args.ref:
            if not oracle.have_ref():
                raise SystemExit("oracle/_ref is not built")
            t0 = time.time()
            packed = synth.hashed_reads(p, 0, c["n_reads"]).numpy()
            ascii_reads = synth.unpack_reads_ascii(packed, L)
            print(f"[{cfg}] {c['n_reads']} reads in {time.time() - t0:.0f} s", flush=True)
            r = oracle.ref_map_reads(text, ascii_reads, None, L, mode="d", threads=args.threads, **MATCH)
            out = {"config": cfg, "seed": SEED, "mode": "d", "reads": c["n_reads"], "text_len": int(p.text_len), "read_len": L,
                   "matched": int(r.matched), "per_mm": [int(x) for x in r.per_mm], "sha256_pos": sha(r.pos), "sha256_rc": sha(r.rc),
                   "sha256_mm": sha(r.mm), "reference_seconds": round(r.seconds, 1), "reference_threads": args.threads,
                   "what": "DefaultReadsApproxMatcher of the unmodified reference (oracle/_ref) on the whole workload of the counter-based generator"}
            # cross-check: the sampled reads matched alone by the oracle agree with the reference's full run
            ro = oracle.oracle_map_reads(text, sample, None, L, mode="d", **MATCH)
            same = bool(np.array_equal(ro.pos, r.pos[idx]) and np.array_equal(ro.rc, r.rc[idx]) and np.array_equal(ro.mm, r.mm[idx]))
            out["sampled_oracle_equals_reference"] = same
            print(f"[{cfg}] reference mode d: matched {r.matched} in {r.seconds:.0f} s; sampled oracle == reference: {same}", flush=True)
            json.dump(out, open(os.path.join(HERE, f"fullsize_{cfg}_ref.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
