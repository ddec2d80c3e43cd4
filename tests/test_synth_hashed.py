"""CPU: the counter-based workload generator (pgrc_b200/csrc/pgs_synth.cu, host path) and the full-size fixtures made from
it (tests/golden/make_fullsize.py).  The GPU box re-creates the same bytes with the device path (tests/test_gpu_fullsize.py
checks device == host there); here: determinism, range independence, shape, and that the committed config-1 fixtures are
what the oracle computes today for the generator's output."""
import json
import os

import numpy as np

import oracle
from pgrc_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = 20261017


def test_ranges_are_independent_and_deterministic():
    c = synth.scaled_config("c2", 0.002)
    p = synth.hashed_params(**c, seed=7)
    text = synth.hashed_text(p).numpy()
    assert text.size == int(c["genome_len"] * c["copies"]) and set(np.unique(text)) <= set(b"ACGT")
    assert np.array_equal(synth.hashed_text(p, 1234, 5000).numpy(), text[1234:6234])
    reads = synth.hashed_reads(p, 0, c["n_reads"]).numpy()
    assert reads.shape == (c["n_reads"], 38)
    assert np.array_equal(synth.hashed_reads(p, 100, 50).numpy(), reads[100:150])
    assert np.array_equal(synth.hashed_reads_at(p, [3, 999, 77]), reads[[3, 999, 77]])
    p2 = synth.hashed_params(**c, seed=8)
    assert not np.array_equal(synth.hashed_text(p2, 0, 1000).numpy(), text[:1000])
    # every read carries at least one substitution and still maps: the matcher finds (almost) all of them, none exactly... unless repeated
    r = oracle.oracle_map_reads(text, reads, None, 150)
    assert r.matched > 0.98 * c["n_reads"] and r.per_mm[0] <= 2 and r.per_mm[1] > 0.7 * c["n_reads"]
    # both strands occur: a read matches on the RC pass only when none of the ~2.8 copies of its locus has its orientation (0.5^2.8)
    assert 0.05 < r.rc.mean() < 0.5


def test_config1_fixtures_are_current():
    z = np.load(os.path.join(GOLDEN, "sample_c1.npz"))
    c = synth.scaled_config("c1", 1.0)
    assert int(z["seed"]) == SEED and int(z["n_reads"]) == c["n_reads"]
    p = synth.hashed_params(**c, seed=SEED)
    assert int(z["text_len"]) == p.text_len and int(z["err_q24"]) == p.err_q24 and int(z["contig"]) == p.contig
    text = synth.hashed_text(p).numpy()
    sample = synth.hashed_reads_at(p, z["idx"])
    r = oracle.oracle_map_reads(text, sample, None, c["read_len"])
    assert np.array_equal(r.pos, z["pos"]) and np.array_equal(r.rc, z["rc"]) and np.array_equal(r.mm, z["mm"])
    zc = np.load(os.path.join(GOLDEN, "sample_c1_mode_c.npz"))
    rc_ = oracle.oracle_map_reads(text, sample, None, c["read_len"], mode="c")
    assert np.array_equal(rc_.pos, zc["pos"]) and np.array_equal(rc_.mm, zc["mm"])
    ref = json.load(open(os.path.join(GOLDEN, "fullsize_c1_ref.json")))
    assert ref["sampled_oracle_equals_reference"] and ref["reads"] == c["n_reads"] and sum(ref["per_mm"][:255]) == ref["matched"]


def test_all_fullsize_fixtures_present_and_consistent():
    for name in ("c1", "c2", "c3", "c4", "c5"):
        z = np.load(os.path.join(GOLDEN, f"sample_{name}.npz"))
        c = synth.CONFIGS[name]
        assert int(z["n_reads"]) == c["n_reads"] and int(z["genome_len"]) == c["genome_len"] and z["idx"].size == 20000
        assert z["idx"].max() < c["n_reads"] and (z["mm"] != 255).mean() > 0.85
    for name in ("c1", "c2", "c3"):
        ref = json.load(open(os.path.join(GOLDEN, f"fullsize_{name}_ref.json")))
        assert ref["sampled_oracle_equals_reference"] is True
