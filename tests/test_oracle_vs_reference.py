"""CPU, live: the plain-C oracle against the reference's own classes (oracle/_ref, built from
/root/reference by oracle/Makefile) on fresh seeded inputs.  Skipped where oracle/_ref is absent
(the committed golden vectors in tests/test_oracle_golden.py cover that case)."""
import numpy as np
import pytest

import oracle
from pgrc_b200 import synth

pytestmark = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def _cmp(inp, **kw):
    r = oracle.ref_map_reads(inp.text, inp.lq_reads, inp.n_reads, inp.read_len, **kw)
    o = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, inp.read_len, **kw)
    assert np.array_equal(o.pos, r.pos) and np.array_equal(o.rc, r.rc) and np.array_equal(o.mm, r.mm), inp.name
    assert (o.matched, o.better) == (r.matched, r.better)
    # falseMatchCount: the reference adds table-dependent collisions in some runs (time-seeded table), never fewer
    assert r.false_matches >= o.false_matches
    return o


@pytest.mark.parametrize("seed,L", [(31, 100), (32, 150), (33, 120), (34, 70), (35, 200)])
def test_adversarial(seed, L):
    o = _cmp(synth.adversarial(seed, L, n_reads=1500, text_len=30000))
    if L <= 150:
        assert o.extra["n_cross_strand_skips"] > 0


@pytest.mark.parametrize("kw", [dict(seed=30), dict(seed=33), dict(seed=50), dict(seed=100), dict(mode="D"),
                                dict(pre_seed=100), dict(pre_seed=50, mode="D"), dict(min_chars_per_mismatch=2),
                                dict(min_chars_per_mismatch=7), dict(rev_compl=False)])
def test_parameter_matrix(kw):
    _cmp(synth.adversarial(41, 100, n_reads=1200, text_len=24000), **kw)


def test_workload_shapes():
    _cmp(synth.workload(100_000, 8_000, 100, 0.001, seed=51))
    _cmp(synth.workload(100_000, 12_000, 150, 0.005, seed=52, n_frac=0.02))


def test_device_generator_shape():
    # the bench generator (counter-based, pgs_synth.cu) produces inputs the reference accepts and mostly matches
    c = synth.scaled_config("c2", 0.001)
    _, text, packed = synth.workload_hashed("c2", 0.001, seed=5)
    text, packed = text.numpy(), packed.numpy()
    asc = synth.unpack_reads_ascii(packed, c["read_len"])
    r = oracle.ref_map_reads(text, asc, None, c["read_len"])
    o = oracle.oracle_map_reads(text, packed, None, c["read_len"])
    assert np.array_equal(o.pos, r.pos) and np.array_equal(o.mm, r.mm) and np.array_equal(o.rc, r.rc)
    assert r.matched > 0.95 * len(r.pos)


@pytest.mark.parametrize("seed,L", [(61, 100), (62, 150), (63, 120), (64, 255)])
def test_interleaved_mode_adversarial(seed, L):
    """Mode 'i' (InterleavedReadsApproxMatcher, ReadsMatchers.cpp:343-409): strided seeds, shift j instead of j * seed."""
    o = _cmp(synth.adversarial(seed, L, n_reads=1500, text_len=30000), mode="i")
    assert o.matched > 100


@pytest.mark.parametrize("kw", [dict(mode="i", seed=30), dict(mode="i", seed=33), dict(mode="i", seed=45), dict(mode="I"),
                                dict(mode="i", pre_seed=100), dict(mode="i", pre_seed=50, pre_mode="i"),
                                dict(mode="d", pre_seed=50, pre_mode="i"), dict(mode="I", pre_seed=50, pre_mode="d"),
                                dict(mode="i", min_chars_per_mismatch=2), dict(mode="i", rev_compl=False), dict(mode="i", seed=20)])
def test_interleaved_mode_parameter_matrix(kw):
    _cmp(synth.adversarial(65, 100, n_reads=1200, text_len=24000), **kw)


def test_interleaved_mode_workload_shapes():
    _cmp(synth.workload(100_000, 8_000, 100, 0.001, seed=66), mode="i")
    _cmp(synth.workload(100_000, 12_000, 150, 0.005, seed=67, n_frac=0.02), mode="i")


@pytest.mark.parametrize("pair", [False, True])
@pytest.mark.parametrize("mode", ["d", "i", "c"])
def test_mismatch_lists_of_the_export_step(pair, mode):
    """pgo_mismatch_lists (variants 1 / 2) against the reference's own updateEntry (ReadsMatchers.cpp:548-558) driven as
    exportMatchesInPgOrder drives it; variant 0 (the GPU contract) is the forward fill the other two are derived from."""
    inp = synth.adversarial(91, 100, n_reads=1500, text_len=30000)
    r, off, o, pg, rd = oracle.ref_mismatch_lists(inp.text, inp.lq_reads, inp.n_reads, 100, rev_compl_pair_file=pair, mode=mode)
    assert r.matched > 100 and int(off[-1]) == int(r.mm[r.mm != 255].astype(np.int64).sum()) > 500
    assert (rd == 4).sum() > 0                                          # N reads contribute N mismatches
    want = oracle.oracle_mismatch_lists(inp.text, inp.lq_packed, inp.n_packed, 100, r.pos, r.rc, r.mm, variant=2 if pair else 1)
    for a, b in zip((off, o, pg, rd), want):
        assert np.array_equal(a, b)
    # variant 0 -> variant 1 by the host rule the C++ shim applies: walk the list backwards, complement, mirror the offsets
    f_off, f_o, f_pg, f_rd = oracle.oracle_mismatch_lists(inp.text, inp.lq_packed, inp.n_packed, 100, r.pos, r.rc, r.mm, variant=0)
    assert np.array_equal(f_off, off)
    comp = np.array([3, 2, 1, 0, 4], np.uint8)
    for i in np.nonzero(r.mm != 255)[0][:400]:
        s, e = int(off[i]), int(off[i + 1])
        reversed_fill = (bool(r.rc[i]) != bool(i % 2)) if pair else bool(r.rc[i])
        if reversed_fill:
            assert np.array_equal(o[s:e], (99 - f_o[s:e])[::-1]) and np.array_equal(pg[s:e], comp[f_pg[s:e]][::-1]) and np.array_equal(rd[s:e], comp[f_rd[s:e]][::-1])
        else:
            assert np.array_equal(o[s:e], f_o[s:e]) and np.array_equal(pg[s:e], f_pg[s:e]) and np.array_equal(rd[s:e], f_rd[s:e])


def _cmp_c(inp, **kw):
    """Mode 'c' (CopMEMReadsApproxMatcher) against the reference with ONE thread: its serial index build is the deterministic one."""
    r = oracle.ref_map_reads(inp.text, inp.lq_reads, inp.n_reads, inp.read_len, threads=1, **kw)
    o = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, inp.read_len, **kw)
    bad = np.nonzero((o.pos != r.pos) | (o.rc != r.rc) | (o.mm != r.mm))[0]
    assert bad.size == 0, (inp.name, kw, bad[:5], o.pos[bad[:5]], r.pos[bad[:5]], o.mm[bad[:5]], r.mm[bad[:5]])
    assert (o.matched, o.better, o.false_matches) == (r.matched, r.better, r.false_matches)
    return o


@pytest.mark.parametrize("seed,L", [(111, 100), (112, 150), (113, 120), (114, 64), (115, 255)])
def test_copmem_mode_adversarial(seed, L):
    o = _cmp_c(synth.adversarial(seed, L, n_reads=1500, text_len=30000), mode="c")
    assert o.matched > 100


@pytest.mark.parametrize("kw", [dict(mode="c", seed=30), dict(mode="c", seed=33), dict(mode="c", seed=45), dict(mode="c", seed=64),
                                dict(mode="c", seed=100), dict(mode="C"), dict(mode="c", pre_seed=100, pre_mode="c"),
                                dict(mode="c", pre_seed=50, pre_mode="c"), dict(mode="c", pre_seed=100, pre_mode="d"),
                                dict(mode="d", pre_seed=50, pre_mode="c"), dict(mode="c", min_chars_per_mismatch=2),
                                dict(mode="c", rev_compl=False), dict(mode="c", seed=24), dict(mode="c", seed=120)])
def test_copmem_mode_parameter_matrix(kw):
    _cmp_c(synth.adversarial(116, 100, n_reads=1200, text_len=24000), **kw)


def test_copmem_mode_workload_shapes():
    _cmp_c(synth.workload(100_000, 8_000, 100, 0.001, seed=117), mode="c")
    _cmp_c(synth.workload(100_000, 12_000, 150, 0.005, seed=118, n_frac=0.02), mode="c")
    _cmp_c(synth.workload(100_000, 12_000, 100, 0.01, seed=119), mode="c")


def _random_case(rng):
    """A random small matcher input and a random, valid parameter set over all three matching modes."""
    L = int(rng.choice([40, 50, 64, 75, 100, 101, 125, 150, 200, 255]))
    mode = str(rng.choice(list("dDiIcC")))
    lo_seed = 24 if mode.lower() == "c" else 12
    seed = int(rng.integers(lo_seed, L + 20))
    kw = dict(mode=mode, seed=seed, min_chars_per_mismatch=int(rng.choice([2, 3, 3, 4, 6, 10])), rev_compl=bool(rng.random() < 0.8))
    if rng.random() < 0.4:
        pre_mode = str(rng.choice(list("dDiIcC")))
        kw.update(pre_mode=pre_mode, pre_seed=int(rng.integers(24 if pre_mode.lower() == "c" else 12, L + 20)))
    inp = synth.adversarial(int(rng.integers(1 << 30)), L, n_reads=int(rng.integers(50, 400)), text_len=int(rng.integers(3000, 12000)))
    return inp, kw


@pytest.mark.parametrize("chunk", range(6))
def test_randomized_sweep_over_modes_and_parameters(chunk):
    """60 random (input, parameter) combinations — read lengths 40..255, seeds from 12 (24 for CopMEM) to beyond the read length,
    all mode letters in both phases, with and without the RC pass — oracle against the reference's own classes (one thread)."""
    rng = np.random.default_rng(1000 + chunk)
    for _ in range(10):
        inp, kw = _random_case(rng)
        r = oracle.ref_map_reads(inp.text, inp.lq_reads, inp.n_reads, inp.read_len, threads=1, **kw)
        o = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, inp.read_len, **kw)
        bad = np.nonzero((o.pos != r.pos) | (o.rc != r.rc) | (o.mm != r.mm))[0]
        assert bad.size == 0, (inp.read_len, kw, bad[:5], o.pos[bad[:5]], r.pos[bad[:5]], o.mm[bad[:5]], r.mm[bad[:5]])
        assert o.matched == r.matched
