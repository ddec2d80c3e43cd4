"""GPU: stage 7 (exact matches between pseudogenomes, pgm_mem_*) through the C ABI against the sequential CPU oracle
(oracle.oracle_match_texts, pinned against the reference's CopMEMMatcher) and the committed golden vectors of the
reference.  Bit-exact: the raw resMatches vector in push order."""
import glob
import json
import os
import sys

import numpy as np
import pytest

import oracle
from pgrc_b200 import matcher, synth
from pgrc_b200._lib import PgmError

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pgmatch_*.npz")))


def _check(src, dest, dis, rc, L, min_len=0xFFFFFFFF, tm=None, device_inputs=False, null_dest=False):
    d = src if dis else dest
    q = oracle.reverse_complement(d) if rc else np.ascontiguousarray(d)
    want = oracle.oracle_match_texts(src, q, dis, rc, L, min_len)
    own = tm is None
    if own:
        tm = matcher.GpuTextMatcher(src, L, min_len)
    try:
        arg = None if null_dest else q
        if device_inputs and arg is not None:
            import torch
            arg = torch.from_numpy(arg).cuda()
        got = tm.match_texts(arg, dis, rc, min_len)
    finally:
        if own:
            tm.close()
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got, want)
    return len(want)


@pytest.mark.parametrize("seed", range(12))
def test_adversarial_texts_against_the_oracle(seed):
    n = [4000, 30000, 120000][seed % 3]
    n2 = [769, 3000, 12000, 1600, 50000][seed % 5]
    L = [45, 24, 30, 50, 64, 120, 33, 47][seed % 8]
    src, dest = synth.pg_texts(700 + seed, n, n2, n_frac=0.002 if seed % 3 == 0 else 0.0, self_rc=40 if seed % 2 else 0)
    with matcher.GpuTextMatcher(src, L) as tm:
        total = 0
        for dis, rc in ((False, True), (False, False), (True, True), (True, False)):
            total += _check(src, dest, dis, rc, L, tm=tm)
        total += _check(src, dest, True, True, L, tm=tm, null_dest=True)      # reverse complement taken from the planes on the GPU
        total += _check(src, dest, False, True, L, tm=tm, device_inputs=True)
    assert total > 0


def test_minimal_match_length_below_the_target_is_refused():
    src, dest = synth.pg_texts(800, 20000, 8000, self_rc=20)
    with pytest.raises(PgmError):
        matcher.GpuTextMatcher(src, 45, 40)
    assert _check(src, dest, False, True, 45, 45) > 0            # (equal to the target: fine)
    assert _check(src, dest, False, True, 45, 1000) > 0          # (above: clamped, CopMEMMatcher.cpp:574-575)


@pytest.mark.parametrize("n2", [0, 10, 31, 32, 44, 45, 46, 767, 768, 769, 800, 801, 802, 803, 1535, 1536, 1537, 1569])
def test_destination_lengths_around_the_group_boundaries(n2):
    # K = 32, k2 = 3: groups of 256 query positions = 768 characters; the main loop needs i1 + K + 768 < N2 + 1
    src, _ = synth.pg_texts(900, 6000, 2000)
    rng = np.random.default_rng(n2)
    a = int(rng.integers(0, 6000 - n2 - 1))
    dest = src[a:a + n2].copy()                      # one long match across every group boundary
    if n2 > 200:
        dest[n2 // 2] = ord("A") if dest[n2 // 2] != ord("A") else ord("C")
    _check(src, dest, False, False, 45)
    _check(src, dest, False, True, 45)


def test_long_repeats_and_periodic_texts():
    rng = np.random.default_rng(5)
    unit = synth.random_genome(7, rng)
    src = np.concatenate([synth.random_genome(5000, rng), np.resize(unit, 4000), synth.random_genome(5000, rng),
                          np.full(3000, ord("A"), np.uint8), synth.random_genome(3000, rng)])
    dest = np.concatenate([np.resize(unit, 2500), synth.random_genome(100, rng), src[2000:9000], np.full(1000, ord("A"), np.uint8),
                           src[-2500:]])
    for L in (45, 30, 64):
        assert _check(src, dest, False, False, L) > 0
        _check(src, dest, False, True, L)
        _check(src, dest, True, True, L)


def test_medium_size_shape_of_a_pseudogenome():
    # 3 Mbp source, 1.5 Mbp destination with ~4 % of it copied from the source: what the LQ-vs-HQ call looks like
    src, dest = synth.pg_texts(42, 3_000_000, 1_500_000, max_copy=1200, n_frac=0.0005, self_rc=2000, adversarial=False)
    with matcher.GpuTextMatcher(src, 45) as tm:
        assert _check(src, dest, False, True, 45, tm=tm) > 1000
        assert _check(src, dest, True, True, 45, tm=tm, null_dest=True) > 500


@pytest.mark.parametrize("n_ctx", [1, 2, 3, 5])
def test_group_of_contexts_shares_the_query_groups(n_ctx):
    """pgm_group_mem_*: every context indexes the source, context r takes the r-th share of the groups of 256 query positions,
    the shares are concatenated and the "covered by the previous match" test runs across the seams.  Real GPUs when the box has
    them, else several contexts on GPU 0.  Long matches that span the seams, short destinations (fewer groups than contexts)."""
    import torch
    have = torch.cuda.device_count()
    devs = [k % have for k in range(n_ctx)]
    src, dest = synth.pg_texts(950 + n_ctx, 60000, 20000, max_copy=6000, self_rc=60)
    with matcher.GpuMatcherGroup(devs) as g:
        g.set_text(src)
        assert g.mem_index(45)[:3] == (32, 4, 3)
        total = 0
        for d, dis, rc in ((dest, False, True), (dest, False, False), (src, True, True), (dest[:700], False, True), (dest[:40], False, False)):
            q = oracle.reverse_complement(d) if rc else np.ascontiguousarray(d)
            want = oracle.oracle_match_texts(src, q, dis, rc, 45)
            got = g.match_texts(None if dis else q, dis, rc)
            assert got.shape == want.shape and np.array_equal(got, want), (n_ctx, dis, rc, got.shape, want.shape)
            total += len(want)
        assert total > 100


@pytest.mark.parametrize("n_parts", [1, 2, 4])
def test_shares_of_one_rank_each_merge_to_the_oracle_result(n_parts):
    """pgm_mem_match_share / pgm_mem_get_share (one process per GPU: rank r computes the r-th share) + the product's merge
    (matcher.merge_text_match_shares), all shares computed here on one GPU one after the other."""
    src, dest = synth.pg_texts(990 + n_parts, 50000, 16000, max_copy=5000, self_rc=50)
    with matcher.GpuTextMatcher(src, 45) as tm:
        for d, dis, rc in ((dest, False, True), (src, True, True), (dest[:500], False, False)):
            q = oracle.reverse_complement(d) if rc else np.ascontiguousarray(d)
            want = oracle.oracle_match_texts(src, q, dis, rc, 45)
            shares = [tm.match_texts_share(None if dis else q, dis, rc, r, n_parts) for r in range(n_parts)]
            got = matcher.merge_text_match_shares(shares, tm.K)
            assert got.shape == want.shape and np.array_equal(got, want), (n_parts, dis, rc)
            assert np.array_equal(tm.match_texts(None if dis else q, dis, rc), want)      # (a plain call after the shares)


@pytest.mark.skipif(not GOLDEN, reason="no pgmatch golden vectors")
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_reference_golden_vectors(path):
    z = np.load(path)
    if int(z["min_len"]) < int(z["target_len"]):
        pytest.skip("minimal length below the target: PGM_ERR_UNSUPPORTED (oracle-only vector)")
    with matcher.GpuTextMatcher(z["src"], int(z["target_len"]), int(z["min_len"])) as tm:
        for tag, dis, rc in (("lq", False, True), ("fw", False, False), ("self", True, True)):
            d = z["src"] if dis else z["dest"]
            q = oracle.reverse_complement(d) if rc else d
            got = tm.match_texts(q, dis, rc, int(z["min_len"]))
            assert np.array_equal(got, z["matches_" + tag]), tag


FULLSIZE = [("c2", 0.05, "pgmatch_fullsize_c2_x0.05.json"), ("c2", 1.0, "pgmatch_fullsize_c2.json"), ("c3", 1.0, "pgmatch_fullsize_c3.json")]


@pytest.mark.parametrize("workload,scale,fixture", FULLSIZE, ids=[f[2][:-5] for f in FULLSIZE])
def test_full_size_texts_against_the_reference_run(workload, scale, fixture):
    """The pseudogenome of a BASELINE config (counter-based generator: the same bytes here and in the build container) against
    its own reverse complement and against a derived destination: count and sha256 of both raw resMatches vectors equal the
    REFERENCE's own one-thread run (tests/golden/make_fullsize_pgmatch.py: 140.6 Mbp and 250 Mbp texts, where the sequential
    oracle agreed with the reference too).  Also through a group of three contexts."""
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import pgmatch_bench
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", fixture)))
    src, dest, dest_rc = pgmatch_bench.stage7_texts(workload, scale, torch.device("cuda", 0))
    assert int(src.numel()) == want["src_bases"] and int(dest_rc.numel()) == want["dest_bases"]
    with matcher.GpuTextMatcher(src, want["target_len"]) as tm:
        assert pgmatch_bench.result_digest(tm.match_texts(None, True, True)) == want["self_rc"]
        assert pgmatch_bench.result_digest(tm.match_texts(dest_rc, False, True)) == want["lq"]
    if scale < 1.0 or workload == "c2":
        with matcher.GpuMatcherGroup([0, 0, 0]) as g:
            g.set_text(src)
            g.mem_index(want["target_len"])
            assert pgmatch_bench.result_digest(g.match_texts(None, True, True)) == want["self_rc"]
            assert pgmatch_bench.result_digest(g.match_texts(dest_rc, False, True)) == want["lq"]


def test_errors_are_loud():
    src, dest = synth.pg_texts(1, 5000, 1000)
    with matcher.GpuReadsMatcher(0) as m:
        m.set_text(src)
        with pytest.raises(PgmError):
            matcher.GpuTextMatcher(None, 20, matcher=m)              # minimal matching length below 24: the reference exits
        tm = matcher.GpuTextMatcher(None, 45, matcher=m)
        with pytest.raises(PgmError):
            tm.match_texts(dest, True, True)                         # dest_is_src with another length
        bad = src.copy(); bad[100] = ord("N")
        m.set_text(bad)
        with pytest.raises(PgmError):
            tm.match_texts(dest, False, True)                        # the index belongs to the previous text
        with pytest.raises(PgmError):
            matcher.GpuTextMatcher(None, 45, matcher=m)              # N in the source text
