"""CPU: stage 7 (exact matches between pseudogenomes).  The sequential C oracle (pgo_match_texts) against the committed
golden vectors of the reference's CopMEMMatcher (tests/golden/make_golden_pgmatch.py), live against the reference where
oracle/_ref exists, and the order-free form the CUDA kernels use (tests/cpu_mem_model.py) against the oracle."""
import glob
import os

import numpy as np
import pytest

import cpu_mem_model
import oracle
from pgrc_b200 import synth

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pgmatch_*.npz")))
CALLS = (("lq", False, True), ("fw", False, False), ("self", True, True))


def _query(src, dest, dis, rc):
    d = src if dis else dest
    return oracle.reverse_complement(d) if rc else np.ascontiguousarray(d)


def test_golden_fixtures_present():
    assert len(GOLDEN) >= 10


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    z = np.load(path)
    L, min_len = int(z["target_len"]), int(z["min_len"])
    for tag, dis, rc in CALLS:
        got = oracle.oracle_match_texts(z["src"], _query(z["src"], z["dest"], dis, rc), dis, rc, L, min_len)
        assert np.array_equal(got, z["matches_" + tag]), tag


@pytest.mark.parametrize("path", [p for p in GOLDEN if "min" not in os.path.basename(p)][:6],
                         ids=lambda p: os.path.basename(p)[:-4])
def test_order_free_model_matches_reference_golden(path):
    z = np.load(path)
    L = int(z["target_len"])
    for tag, dis, rc in CALLS:
        got = cpu_mem_model.match_texts(z["src"], _query(z["src"], z["dest"], dis, rc), dis, rc, L)
        assert np.array_equal(got, z["matches_" + tag]), tag


def test_parameters_follow_the_reference():
    # initParams / calcCoprimes (CopMEMMatcher.cpp:69-137): PgRC's default target length 45 gives K = 32, k1 = 4, k2 = 3
    par = {}
    src, dest = synth.pg_texts(3, 5000, 900)
    oracle.oracle_match_texts(src, dest, False, False, 45, params=par)
    assert (par["K"], par["k1"], par["k2"], par["hash_size"]) == (32, 4, 3, 1 << 24)
    assert cpu_mem_model.derive(45, 0xFFFFFFFF, 5000) == (32, 4, 3, 1 << 24)
    for L in (24, 27, 28, 32, 33, 43, 47, 54, 63, 111, 150, 1000):
        oracle.oracle_match_texts(src, dest, False, False, L, params=par)
        assert cpu_mem_model.derive(L, 0xFFFFFFFF, 5000) == (par["K"], par["k1"], par["k2"], par["hash_size"]), L


@pytest.mark.parametrize("seed", range(10))
def test_order_free_model_against_the_oracle(seed):
    n = [4000, 30000][seed % 2]
    n2 = [769, 3000, 9000, 1600][seed % 4]
    L = [45, 24, 30, 50, 64, 120, 33, 47][seed % 8]
    src, dest = synth.pg_texts(500 + seed, n, n2, n_frac=0.002 if seed % 3 == 0 else 0.0, self_rc=30 if seed % 2 else 0)
    total = 0
    for tag, dis, rc in CALLS:
        q = _query(src, dest, dis, rc)
        want = oracle.oracle_match_texts(src, q, dis, rc, L)
        got = cpu_mem_model.match_texts(src, q, dis, rc, L)
        assert got.shape == want.shape and np.array_equal(got, want), (tag, got.shape, want.shape)
        total += len(want)
    assert total > 0


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("seed", range(16))
def test_oracle_against_the_reference_live(seed):
    n = [4000, 30000, 150000][seed % 3]
    n2 = [769, 3000, 12000, 1600, 60000][seed % 5]
    L = [45, 24, 30, 50, 64, 120, 33, 47][seed % 8]
    min_len = [0xFFFFFFFF, 0xFFFFFFFF, max(24, L - 9)][seed % 3]
    src, dest = synth.pg_texts(600 + seed, n, n2, n_frac=0.002 if seed % 4 == 0 else 0.0, self_rc=30 if seed % 2 else 0)
    for tag, dis, rc in CALLS + (("self_fw", True, False),):
        q = _query(src, dest, dis, rc)
        want = oracle.ref_match_texts(src, q, dis, rc, L, min_len, threads=1)
        got = oracle.oracle_match_texts(src, q, dis, rc, L, min_len)
        assert got.shape == want.shape and np.array_equal(got, want), (tag, got.shape, want.shape)


@pytest.mark.parametrize("n_parts", [2, 3, 7])
def test_query_groups_shared_out_over_several_contexts(n_parts):
    """The host logic of pgm_group_mem_match (ranges of groups per context, concatenation, suppression across the seams) in the
    CPU model against the sequential oracle; long copies so that matches span the seams."""
    src, dest = synth.pg_texts(640 + n_parts, 30000, 9000, max_copy=5000, self_rc=40)
    for tag, dis, rc in CALLS:
        q = _query(src, dest, dis, rc)
        want = oracle.oracle_match_texts(src, q, dis, rc, 45)
        got = cpu_mem_model.match_texts(src, q, dis, rc, 45, n_parts=n_parts)
        assert got.shape == want.shape and np.array_equal(got, want), (tag, n_parts)
