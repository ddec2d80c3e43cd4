"""CPU: the STAGED form of mode c's per-read query (copmem_query_staged in the oracle: every distinct alignment of a read verified
once with both mismatch counts in full, then the reference's sequential loop replayed over table indices — the CPU model of the
opt-in warp-per-read kernels, pgrc_b200/csrc/pgm_copmem_warp.cuh) against the sequential transcription, which is pinned against
the reference: same positions, strands, mismatch counts AND the log-only better / false-match counters."""
import numpy as np
import pytest

import oracle
from golden_util import NAMES, load
from pgrc_b200 import synth


def _both(inp_text, lq, nn, L, **kw):
    oracle.set_copmem_staged(False)
    a = oracle.oracle_map_reads(inp_text, lq, nn, L, **kw)
    oracle.set_copmem_staged(True)
    try:
        b = oracle.oracle_map_reads(inp_text, lq, nn, L, **kw)
    finally:
        oracle.set_copmem_staged(False)
    assert np.array_equal(a.pos, b.pos) and np.array_equal(a.rc, b.rc) and np.array_equal(a.mm, b.mm)
    assert (a.matched, a.better, a.false_matches) == (b.matched, b.better, b.false_matches)
    return a


@pytest.mark.parametrize("seed,L", [(401, 100), (402, 150), (403, 120), (404, 64), (405, 255)])
def test_staged_query_equals_the_sequential_one(seed, L):
    inp = synth.adversarial(seed, L, n_reads=1200, text_len=24000)
    for kw in (dict(mode="c"), dict(mode="C"), dict(mode="c", seed=30), dict(mode="c", pre_seed=50, pre_mode="c")):
        if kw.get("seed", 38) <= L:
            assert _both(inp.text, inp.lq_packed, inp.n_packed, L, **kw).matched > 0


@pytest.mark.parametrize("name", [n for n in NAMES if n.startswith("cop_")])
def test_staged_query_on_the_reference_golden_vectors(name):
    g = load(name)
    r = _both(g["text"], g["lq_packed"], g["n_packed"], g["read_len"], **g["params"])
    assert np.array_equal(r.pos, g["pos"]) and np.array_equal(r.mm, g["mm"]) and r.false_matches == int(g["false_matches"])
