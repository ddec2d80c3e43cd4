"""CPU: the plain-C oracle (oracle/pgrc_oracle.c) against the committed golden vectors — outputs of
the reference's own classes (tests/golden/make_golden.py).  Bit-exact, including the log-only
counters betterMatchCount / falseMatchCount, which pins the oracle's event order too."""
import numpy as np
import pytest

import oracle
from golden_util import NAMES, load
from pgrc_b200 import synth


def test_golden_fixtures_present():
    assert len(NAMES) >= 15


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_golden(name):
    g = load(name)
    L = g["read_len"]
    # layout pin (a11): our packers reproduce the reference's SymbolsPackingFacility bytes
    assert np.array_equal(oracle.pack_reads(g["lq_reads"], L, False), g["lq_packed"])
    assert np.array_equal(synth.pack_reads(g["lq_reads"], False), g["lq_packed"])
    if len(g["n_reads"]):
        assert np.array_equal(oracle.pack_reads(g["n_reads"], L, True), g["n_packed"])
        assert np.array_equal(synth.pack_reads(g["n_reads"], True), g["n_packed"])
    r = oracle.oracle_map_reads(g["text"], g["lq_packed"], g["n_packed"], L, **g["params"])
    assert np.array_equal(r.pos, g["pos"])
    assert np.array_equal(r.rc, g["rc"])
    assert np.array_equal(r.mm, g["mm"])
    assert r.matched == int(g["matched"])
    assert r.better == int(g["better"])
    # falseMatchCount (log-only) is the one output of the reference that is not reproducible: its character
    # table is seeded from time() (mersennetwister.cpp:151-175), and besides the table-independent collisions
    # the oracle reproduces, some table-dependent ones fire in a fraction of the runs (observed: +3 in 2 of 12
    # runs on c4_shape_small; with seed length == word size 32, periodic windows such as poly-A collapse to a
    # few parity-determined hash values and the count moves by hundreds).  None of them survives verification.
    if name.startswith("cop_"):   # CopMEM's hash is not time-seeded: its counters are deterministic
        assert int(g["false_matches"]) == r.false_matches
    assert int(g["false_matches"]) >= r.false_matches
    if name.startswith("c") and not name.startswith("cop_"):   # random genomes: only rare accidental extras
        assert int(g["false_matches"]) <= r.false_matches + 16
    exact_only = L == min(g["params"]["seed"], L) and g["params"]["pre_seed"] == 0
    if not exact_only:
        assert np.array_equal(r.per_mm, g["per_mm"])


def test_unpack_roundtrip():
    rng = np.random.default_rng(3)
    for L in (1, 3, 4, 5, 37, 100, 150, 255):
        reads = synth.sample_reads(synth.random_genome(2000, rng), 50, L, 0.05, rng, require_error=False) if L < 2000 else None
        assert np.array_equal(synth.unpack_reads_ascii(synth.pack_reads(reads), L), reads)
