"""Generates tests/golden/*.npz — golden vectors for the read-vs-pseudogenome matching path.

The reference ships no fixtures for this path, so the pins are outputs of the
reference's OWN classes (DefaultReadsApproxMatcher / InterleavedReadsApproxMatcher / DefaultReadsExactMatcher, unmodified objects
built from /root/reference by oracle/Makefile) run in the build container through
oracle/ref_harness.cpp.  This script needs oracle/_ref/libpgrc_ref.so, i.e. it only runs where
/root/reference exists; the .npz files it writes are committed and travel to the GPU box.

    python tests/golden/make_golden.py

Each file holds the inputs (text, ASCII reads of the ACGT and the ACGNT set), the matcher
parameters, and the reference's per-read results (readMatchPos, readMatchRC,
readMismatchesCount) plus its counters (matched, better, false, per-mismatch histogram).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from pgrc_b200 import synth  # noqa: E402

# (name, input factory, matcher keyword arguments)
CASES = [
    ("adv_L100_default", lambda: synth.adversarial(201, 100, n_reads=900, text_len=16000), {}),
    ("adv_L150_default", lambda: synth.adversarial(202, 150, n_reads=900, text_len=20000), {}),
    ("adv_L120_seed30", lambda: synth.adversarial(203, 120, n_reads=700, text_len=16000), dict(seed=30)),
    ("adv_L100_seed45", lambda: synth.adversarial(204, 100, n_reads=700, text_len=16000), dict(seed=45)),
    ("adv_L100_exact", lambda: synth.adversarial(205, 100, n_reads=700, text_len=16000), dict(seed=100)),
    ("adv_L100_shortcut", lambda: synth.adversarial(206, 100, n_reads=700, text_len=16000), dict(mode="D")),
    ("adv_L100_prephase_exact", lambda: synth.adversarial(207, 100, n_reads=700, text_len=16000), dict(pre_seed=100)),
    ("adv_L100_prephase50", lambda: synth.adversarial(208, 100, n_reads=700, text_len=16000), dict(pre_seed=50)),
    ("adv_L100_M2", lambda: synth.adversarial(209, 100, n_reads=700, text_len=16000), dict(min_chars_per_mismatch=2)),
    ("adv_L100_norc", lambda: synth.adversarial(210, 100, n_reads=700, text_len=16000), dict(rev_compl=False)),
    ("adv_L64_seed32", lambda: synth.adversarial(211, 64, n_reads=700, text_len=12000), dict(seed=32)),
    ("adv_L255_default", lambda: synth.adversarial(212, 255, n_reads=400, text_len=24000), {}),
    ("c1_shape_small", lambda: synth.workload(40_000, 3_000, 100, 0.001, seed=213, name="c1 shape"), {}),
    ("c2_shape_small", lambda: synth.workload(40_000, 4_000, 150, 0.005, seed=214, n_frac=0.03, name="c2 shape"), {}),
    ("c4_shape_small", lambda: synth.workload(40_000, 4_000, 100, 0.01, seed=215, name="c4 shape"), {}),
    # mode 'i' (InterleavedReadsApproxMatcher, ReadsMatchers.cpp:343-409): strided seeds
    ("ilv_L100_default", lambda: synth.adversarial(221, 100, n_reads=900, text_len=16000), dict(mode="i")),
    ("ilv_L150_default", lambda: synth.adversarial(222, 150, n_reads=900, text_len=20000), dict(mode="i")),
    ("ilv_L120_seed30", lambda: synth.adversarial(223, 120, n_reads=700, text_len=16000), dict(mode="i", seed=30)),
    ("ilv_L100_shortcut", lambda: synth.adversarial(224, 100, n_reads=700, text_len=16000), dict(mode="I")),
    ("ilv_L100_prephase_exact", lambda: synth.adversarial(225, 100, n_reads=700, text_len=16000), dict(mode="i", pre_seed=100)),
    ("ilv_L100_prephase50_i", lambda: synth.adversarial(226, 100, n_reads=700, text_len=16000), dict(mode="i", pre_seed=50, pre_mode="i")),
    ("ilv_L255_default", lambda: synth.adversarial(227, 255, n_reads=400, text_len=24000), dict(mode="i")),
    ("ilv_c2_shape_small", lambda: synth.workload(40_000, 4_000, 150, 0.005, seed=228, n_frac=0.03, name="c2 shape"), dict(mode="i")),
    # mode 'c' (CopMEMReadsApproxMatcher, ReadsMatchers.cpp:411-451; what the release CLI runs), reference at ONE thread:
    # its serial index build (CopMEMMatcher.cpp:171-233) is the deterministic one
    ("cop_L100_default", lambda: synth.adversarial(231, 100, n_reads=900, text_len=16000), dict(mode="c")),
    ("cop_L150_default", lambda: synth.adversarial(232, 150, n_reads=900, text_len=20000), dict(mode="c")),
    ("cop_L120_seed30", lambda: synth.adversarial(233, 120, n_reads=700, text_len=16000), dict(mode="c", seed=30)),
    ("cop_L100_seed64", lambda: synth.adversarial(234, 100, n_reads=700, text_len=16000), dict(mode="c", seed=64)),
    ("cop_L100_shortcut", lambda: synth.adversarial(235, 100, n_reads=700, text_len=16000), dict(mode="C")),
    ("cop_L100_prephase50_c", lambda: synth.adversarial(236, 100, n_reads=700, text_len=16000), dict(mode="c", pre_seed=50, pre_mode="c")),
    ("cop_L100_prephase_exact_d", lambda: synth.adversarial(237, 100, n_reads=700, text_len=16000), dict(mode="c", pre_seed=100, pre_mode="d")),
    ("cop_L255_default", lambda: synth.adversarial(238, 255, n_reads=400, text_len=24000), dict(mode="c")),
    ("cop_c2_shape_small", lambda: synth.workload(40_000, 4_000, 150, 0.005, seed=239, n_frac=0.03, name="c2 shape"), dict(mode="c")),
    ("cop_c1_shape_small", lambda: synth.workload(40_000, 3_000, 100, 0.001, seed=240, name="c1 shape"), dict(mode="c")),
]

DEFAULTS = dict(seed=38, min_chars_per_mismatch=3, mode="d", pre_seed=0, pre_mode="d", rev_compl=True)


def main():
    if not oracle.have_ref():
        raise SystemExit("oracle/_ref/libpgrc_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    for name, make, kw in CASES:
        if os.path.exists(os.path.join(HERE, name + ".npz")) and "--all" not in sys.argv:
            continue   # committed vectors are kept as they are; --all regenerates every case
        inp = make()
        p = dict(DEFAULTS); p.update(kw)
        r = oracle.ref_map_reads(inp.text, inp.lq_reads, inp.n_reads, inp.read_len, threads=1 if "c" in (p["mode"] + p["pre_mode"]).lower() else 0, **p)
        # layout pin (a11): the reference's own packing of these reads
        lq_packed = oracle.pack_reads(inp.lq_reads, inp.read_len, False, use_ref=True)
        n_packed = oracle.pack_reads(inp.n_reads, inp.read_len, True, use_ref=True)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), text=inp.text, lq_reads=inp.lq_reads, n_reads=inp.n_reads,
            lq_packed=lq_packed, n_packed=n_packed, read_len=np.int64(inp.read_len),
            seed=np.int64(p["seed"]), min_chars_per_mismatch=np.int64(p["min_chars_per_mismatch"]),
            mode=np.bytes_(p["mode"]), pre_seed=np.int64(p["pre_seed"]), pre_mode=np.bytes_(p["pre_mode"]),
            rev_compl=np.int64(int(p["rev_compl"])),
            pos=r.pos, rc=r.rc, mm=r.mm, matched=np.int64(r.matched), better=np.int64(r.better),
            false_matches=np.int64(r.false_matches), per_mm=r.per_mm)
        print(f"{name}: {len(r.pos)} reads, matched {r.matched}, better {r.better}, false {r.false_matches}")


if __name__ == "__main__":
    main()
