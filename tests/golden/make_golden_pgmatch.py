"""Generates tests/golden/pgmatch_*.npz — golden vectors for stage 7 (exact matches between pseudogenomes).

The reference ships no fixtures; the pins are outputs of the reference's OWN CopMEMMatcher (constructor +
matchTexts, one thread: its serial index build is the deterministic one) and SimplePgMatcher::markAndRemoveExactMatches,
run in the build container through oracle/ref_harness.cpp.  Needs oracle/_ref/libpgrc_ref.so (i.e. /root/reference);
the .npz files are committed and travel to the GPU box.

    python tests/golden/make_golden_pgmatch.py [--all]

Per file: src, dest, target_len, min_len; resMatches {posSrcText, length, posDestText} in push order for the three ways
SimplePgMatcher calls the matcher — matches_lq (destination reverse-complemented, SimplePgMatcher.cpp:39-41), matches_fw
(no reverse complement, :43) and matches_self (the source against its own reverse complement, :35-36) — and, for the lq
and self calls, the outputs of markAndRemoveExactMatches (mapped sequence, offsets stream, lengths stream).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from pgrc_b200 import synth  # noqa: E402

# (name, (seed, n, n2, keyword arguments of synth.pg_texts), target length, minimal length)
CASES = [
    ("pgmatch_L45_default", (301, 24000, 9000, dict(self_rc=30)), 45, 0xFFFFFFFF),
    ("pgmatch_L45_with_n", (302, 24000, 9000, dict(self_rc=30, n_frac=0.003)), 45, 0xFFFFFFFF),
    ("pgmatch_L24", (303, 16000, 6000, dict(self_rc=20)), 24, 0xFFFFFFFF),
    ("pgmatch_L30", (304, 16000, 6000, dict(self_rc=20)), 30, 0xFFFFFFFF),
    ("pgmatch_L50", (305, 16000, 6000, dict(self_rc=20)), 50, 0xFFFFFFFF),
    ("pgmatch_L64", (306, 20000, 8000, dict(self_rc=20)), 64, 0xFFFFFFFF),
    ("pgmatch_L120", (307, 30000, 12000, dict(self_rc=20)), 120, 0xFFFFFFFF),
    ("pgmatch_L45_min36", (308, 16000, 6000, dict(self_rc=20)), 45, 36),
    ("pgmatch_L64_min28", (309, 16000, 6000, dict(self_rc=20)), 64, 28),
    ("pgmatch_L45_short_dest", (310, 16000, 801, dict()), 45, 0xFFFFFFFF),
    ("pgmatch_L45_plain", (311, 60000, 30000, dict(self_rc=100, adversarial=False, max_copy=600)), 45, 0xFFFFFFFF),
]


def main():
    if not oracle.have_ref():
        raise SystemExit("oracle/_ref/libpgrc_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    for name, (seed, n, n2, kw), L, min_len in CASES:
        path = os.path.join(HERE, name + ".npz")
        if os.path.exists(path) and "--all" not in sys.argv:
            continue
        src, dest = synth.pg_texts(seed, n, n2, **kw)
        out = dict(src=src, dest=dest, target_len=np.uint32(L), min_len=np.uint32(min_len))
        for tag, dis, rc in (("lq", False, True), ("fw", False, False), ("self", True, True)):
            d = src if dis else dest
            q = oracle.reverse_complement(d) if rc else d
            out["matches_" + tag] = oracle.ref_match_texts(src, q, dis, rc, L, min_len, threads=1)
        for tag, dis in (("lq", False), ("self", True)):
            mapped, off, ln = oracle.ref_mark_matches(src, dest, dis, True, L, min_len, threads=1)
            out["mapped_" + tag], out["map_off_" + tag], out["map_len_" + tag] = mapped, off, ln
        np.savez_compressed(path, **out)
        print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.startswith("matches")})


if __name__ == "__main__":
    main()
