"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle
(oracle/pgrc_oracle.c, pinned against the reference's own classes) on the same seeded inputs.
Bar: bit-exact positions, strands, mismatch counts, matched count and histogram."""
import numpy as np
import pytest

import oracle
from pgrc_b200 import matcher, synth

pytestmark = pytest.mark.gpu


def _check(inp, **kw):
    okw = dict(seed=kw.get("reads_exact_matching_chars", 38), min_chars_per_mismatch=kw.get("min_chars_per_mismatch", 3),
               mode=kw.get("matching_mode", "d"), pre_seed=kw.get("pre_reads_exact_matching_chars", 0),
               pre_mode=kw.get("pre_matching_mode", "d"), rev_compl=kw.get("rev_compl_pg", True))
    want = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, inp.read_len, **okw)
    got = matcher.map_reads_into_pg(inp.text, inp.lq_packed, inp.n_packed, inp.read_len, **kw)
    bad = np.nonzero((got.pos != want.pos) | (got.rc != want.rc) | (got.mm != want.mm))[0]
    assert bad.size == 0, (f"{inp.name} {kw}: {bad.size} reads differ, first {bad[:5]}: "
                           f"gpu pos/rc/mm {got.pos[bad[:5]]} {got.rc[bad[:5]]} {got.mm[bad[:5]]} "
                           f"oracle {want.pos[bad[:5]]} {want.rc[bad[:5]]} {want.mm[bad[:5]]}")
    assert got.matched == want.matched
    if not (inp.read_len == okw["seed"] and okw["pre_seed"] == 0):
        assert np.array_equal(got.per_mm, want.per_mm)
    return got, want


@pytest.mark.parametrize("seed,L", [(1, 100), (2, 100), (3, 150), (4, 120), (5, 64), (6, 255)])
def test_adversarial_default_params(seed, L):
    got, want = _check(synth.adversarial(seed, L))
    if L <= 150:
        assert want.extra["n_cross_strand_skips"] > 0  # the ":313" quirk is exercised


@pytest.mark.parametrize("kw", [
    dict(reads_exact_matching_chars=30),
    dict(reads_exact_matching_chars=32),
    dict(reads_exact_matching_chars=33),
    dict(reads_exact_matching_chars=45),
    dict(reads_exact_matching_chars=100),                      # exact path
    dict(reads_exact_matching_chars=200),                      # clamped to read length: exact path
    dict(matching_mode="D"),                                   # shortcut mode
    dict(pre_reads_exact_matching_chars=100),                  # exact pre-phase + continuation
    dict(pre_reads_exact_matching_chars=50, reads_exact_matching_chars=38),
    dict(pre_reads_exact_matching_chars=50, matching_mode="D"),
    dict(pre_reads_exact_matching_chars=100, pre_matching_mode="D", matching_mode="D"),
    dict(min_chars_per_mismatch=2),
    dict(min_chars_per_mismatch=10),
    dict(rev_compl_pg=False),
])
def test_adversarial_parameter_matrix(kw):
    for s in (11, 12):
        _check(synth.adversarial(s, 100), **kw)


def test_config1_shape_scaled():
    # BASELINE config 1 shape / 20: SE, 100 bp, 0.1 % substitutions
    _check(synth.workload(250_000, 20_000, 100, 0.001, seed=101, name="c1/20"))


def test_config2_shape_scaled():
    # BASELINE config 2 shape / 100: 150 bp, 0.5 %, 3 seeds per read
    _check(synth.workload(500_000, 100_000, 150, 0.005, seed=102, n_frac=0.02, name="c2/100"))


def test_config4_shape_scaled_high_error():
    _check(synth.workload(400_000, 60_000, 100, 0.01, seed=104, name="c4 scaled"))


def test_edge_cases():
    rng = np.random.default_rng(5)
    g = synth.random_genome(5000, rng)
    reads = synth.sample_reads(g, 300, 100, 0.01, rng)
    # text shorter than a read, than a seed, empty reads sets, single read
    for text in (g[:99], g[:37], g[:38], g[:100], g[:101], g):
        inp = synth.MatcherInputs(np.ascontiguousarray(text), reads, np.zeros((0, 100), np.uint8), 100, f"text{len(text)}")
        _check(inp)
    inp = synth.MatcherInputs(g, reads[:1], np.zeros((0, 100), np.uint8), 100, "one read")
    _check(inp)
    inp = synth.MatcherInputs(g, np.zeros((0, 100), np.uint8), synth.inject_n(reads[:50], rng), 100, "only N reads")
    _check(inp)
    # duplicated reads and low-complexity text: long chains of identical seeds
    dup = np.repeat(reads[:10], 40, axis=0)
    polya = np.full(3000, ord("A"), np.uint8)
    text = np.concatenate([g[:1500], polya, g[1500:3000]])
    areads = np.full((200, 100), ord("A"), np.uint8)
    inp = synth.MatcherInputs(np.ascontiguousarray(text), np.concatenate([dup, areads]), np.zeros((0, 100), np.uint8), 100, "dups")
    _check(inp)


def test_bad_symbol_is_an_error():
    rng = np.random.default_rng(6)
    g = synth.random_genome(4000, rng)
    g[1234] = ord("N")
    reads = synth.sample_reads(synth.random_genome(4000, rng), 10, 100, 0.01, rng)
    with pytest.raises(matcher.PgmError) as e:
        matcher.map_reads_into_pg(g, synth.pack_reads(reads), None, 100)
    assert e.value.status == -5


def test_unsupported_modes_fail_loudly():
    rng = np.random.default_rng(7)
    g = synth.random_genome(4000, rng)
    reads = synth.sample_reads(g, 10, 100, 0.01, rng)
    for kw in (dict(matching_mode="x"), dict(matching_mode="d", pre_reads_exact_matching_chars=50, pre_matching_mode="q"),
               dict(match_prefix_length=50), dict(matching_mode="c", reads_exact_matching_chars=20)):   # (CopMEM: seed < 24)
        with pytest.raises(matcher.PgmError) as e:
            matcher.map_reads_into_pg(g, synth.pack_reads(reads), None, 100, **kw)
        assert e.value.status == -6


def test_text_left_untouched_and_device_inputs():
    import torch
    inp = synth.adversarial(21, 100)
    before = inp.text.copy()
    want = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, 100)
    t = torch.from_numpy(inp.text).cuda()
    lq = torch.from_numpy(inp.lq_packed).cuda()
    nn = torch.from_numpy(inp.n_packed).cuda()
    got = matcher.map_reads_into_pg(t, lq, nn, 100)
    assert np.array_equal(inp.text, before) and np.array_equal(t.cpu().numpy(), before)
    assert np.array_equal(got.pos, want.pos) and np.array_equal(got.rc, want.rc) and np.array_equal(got.mm, want.mm)


def test_context_reuse_and_filter_off():
    inp = synth.adversarial(22, 100)
    inp2 = synth.adversarial(23, 150)
    w1 = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, 100)
    w2 = oracle.oracle_map_reads(inp2.text, inp2.lq_packed, inp2.n_packed, 150)
    with matcher.GpuReadsMatcher(0) as m:
        for fb in (-1, 0, 12):
            m.set_tuning(filter_log2_bits=fb)
            g1 = matcher.map_reads_into_pg(inp.text, inp.lq_packed, inp.n_packed, 100, matcher=m)
            g2 = matcher.map_reads_into_pg(inp2.text, inp2.lq_packed, inp2.n_packed, 150, matcher=m)
            assert np.array_equal(g1.pos, w1.pos) and np.array_equal(g1.mm, w1.mm) and np.array_equal(g1.rc, w1.rc)
            assert np.array_equal(g2.pos, w2.pos) and np.array_equal(g2.mm, w2.mm) and np.array_equal(g2.rc, w2.rc)
        assert m.kernel_launches() > 0


def test_hot_seeds_duplicate_reads_and_homopolymers():
    """Hundreds of reads sharing one seed (duplicates, poly-A): exercises the full-bucket walk and the
    chains behind a slot of the seed table, and the per-warp candidate queues running full."""
    rng = np.random.default_rng(9)
    g = synth.random_genome(30_000, rng)
    text = np.concatenate([g[:10_000], np.full(400, ord("A"), np.uint8), g[10_000:20_000], np.full(300, ord("T"), np.uint8),
                           np.tile(np.frombuffer(b"AC", np.uint8), 200), g[20_000:]])
    base = synth.sample_reads(g, 40, 100, 0.02, rng)
    dup = np.repeat(base, 60, axis=0)                                     # 60 copies of each of 40 reads
    polya = np.full((300, 100), ord("A"), np.uint8)
    polya[np.arange(300), rng.integers(0, 100, 300)] = ord("C")           # one substitution each
    acac = np.tile(np.frombuffer(b"AC", np.uint8), (150, 50))
    reads = np.concatenate([dup, polya, acac, synth.sample_reads(g, 500, 100, 0.01, rng)])
    reads = reads[rng.permutation(len(reads))]
    nn = synth.inject_n(np.repeat(base[:5], 40, axis=0), rng)
    inp = synth.MatcherInputs(np.ascontiguousarray(text), np.ascontiguousarray(reads), nn, 100, "hot seeds")
    got, want = _check(inp)
    assert got.stats["candidates"] > 100_000
    _check(inp, matching_mode="D")
    _check(inp, reads_exact_matching_chars=100)


def _run_sharded_on_one_gpu(inp, world, **kw):
    """The text-sharded path with `world` contexts on ONE GPU, one thread per rank: every context scans its text range and
    the product's own run_plan_sharded / merge_accumulators exchange the per-read accumulators (tests/local_comm.py stands in
    for the NCCL all-reduces).  Exercises the shard geometry of the kernels (halos, RC coordinates of a slice) AND the merge
    protocol, including the in-place reduction of the `touched` flag that resolve_kernel gates on."""
    from local_comm import LocalWorld
    pg_len = inp.text.size
    plan = matcher.MatchPlan.derive(inp.read_len, kw.get("seed", 38), kw.get("min_chars_per_mismatch", 3), kw.get("mode", "d"),
                                    kw.get("pre_seed", 0), kw.get("pre_mode", "d"))

    def rank_body(rank, comm):
        with matcher.GpuReadsMatcher(0, use_torch_stream=True) as m:
            sb, sl, ob, oe = matcher.shard_plan(pg_len, rank, world)
            m.set_text_shard(np.ascontiguousarray(inp.text[sb:sb + sl]), sb, pg_len, ob, oe)
            m.set_reads(inp.lq_packed, inp.n_packed if len(inp.n_reads) else None, inp.read_len)
            matcher.run_plan_sharded(m, plan, kw.get("rev_compl", True), comm)
            return m.get_results()

    return LocalWorld(world).run(rank_body)


def _run_routed_on_one_gpu(inp, world, round_windows=0, exchange=None, **kw):
    """The routed scheme (pgm_route_*) with `world` contexts on ONE GPU, one thread per rank, driven by the product's
    run_plan_routed; returns the per-rank results concatenated in read order (rank g owns read range g)."""
    from local_comm import LocalWorld
    plan = matcher.MatchPlan.derive(inp.read_len, kw.get("seed", 38), kw.get("min_chars_per_mismatch", 3), kw.get("mode", "d"),
                                    kw.get("pre_seed", 0), kw.get("pre_mode", "d"))
    n_lq, n_n = inp.lq_packed.shape[0], (inp.n_packed.shape[0] if len(inp.n_reads) else 0)
    n = n_lq + n_n
    rb = matcher.read_ranges(n, world)

    def rank_body(rank, comm):
        lo, hi = rb[rank], rb[rank + 1]
        lq = inp.lq_packed[min(lo, n_lq):min(hi, n_lq)]          # global read index: LQ reads first, then N reads
        nn = inp.n_packed[max(lo, n_lq) - n_lq:max(hi, n_lq) - n_lq] if n_n else None
        with matcher.GpuReadsMatcher(0, use_torch_stream=True) as m:
            m.set_text(inp.text)
            m.set_reads(np.ascontiguousarray(lq), np.ascontiguousarray(nn) if nn is not None and len(nn) else None, inp.read_len)
            # world 2 / 3: the pipelined schedule with the peer-pull exchange (what bench.py runs on real GPUs; here the peers
            # are contexts of this process, so the pull copies device-to-device without IPC); world 8: plain send / receive
            ex = exchange or ("pull" if world <= 3 else "nccl")
            info = matcher.run_plan_routed(m, plan, kw.get("rev_compl", True), comm, n, round_windows,
                                           comm2=comm.sibling() if ex == "pull" else None, exchange=ex, deep=(world == 3))
            return m.get_results(), info

    outs = LocalWorld(world).run(rank_body)
    res = [o[0] for o in outs]
    return (np.concatenate([r.pos for r in res]), np.concatenate([r.rc for r in res]), np.concatenate([r.mm for r in res]),
            sum(r.matched for r in res), [o[1] for o in outs], res)


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("kw", [dict(), dict(pre_seed=100), dict(mode="D"), dict(seed=45, rev_compl=False), dict(seed=100), dict(pre_seed=50)])
def test_routed_contexts_match_oracle(world, kw):
    """Hash-partitioned seed table + routed windows / candidates: the result is the single-matcher result bit for bit —
    adversarial inputs (hot seeds with chains, N reads, palindromes: the rule-3 accumulators), the scaled config shapes,
    one and several rounds per pass."""
    for inp, rw in ((synth.adversarial(61, 100, n_reads=2000, text_len=30000), 0), (synth.adversarial(62, 150, n_reads=1500, text_len=30000), 4096),
                    (synth.workload(300_000, 40_000, 150, 0.005, seed=63, n_frac=0.03, name="c2 shape"), 65536)):
        want = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, inp.read_len, **kw)
        pos, rc, mm, matched, infos, res = _run_routed_on_one_gpu(inp, world, rw, **kw)
        bad = np.nonzero((pos != want.pos) | (rc != want.rc) | (mm != want.mm))[0]
        assert bad.size == 0, (f"{inp.name} world {world} {kw}: {bad.size} reads differ, first {bad[:5]}: gpu {pos[bad[:5]]} "
                               f"{rc[bad[:5]]} {mm[bad[:5]]} oracle {want.pos[bad[:5]]} {want.rc[bad[:5]]} {want.mm[bad[:5]]}")
        assert matched == want.matched
        if rw and world <= 3:
            assert infos[0]["rounds_per_pass"] > 1
        assert sum(i["sent_bytes"]["windows"] for i in infos) > 0 and sum(r.stats["candidates"] for r in res) > 0


def test_routed_hot_seeds_and_empty_ranks():
    """Duplicated reads (one hash owner gets almost every pattern, chains behind one slot) and more ranks than reads."""
    rng = np.random.default_rng(64)
    g = synth.random_genome(20_000, rng)
    dup = np.repeat(synth.sample_reads(g, 3, 100, 0.01, rng), 3000, axis=0)
    inp = synth.MatcherInputs(g, np.concatenate([dup, synth.sample_reads(g, 2000, 100, 0.01, rng)]), np.zeros((0, 100), np.uint8), 100, "skew")
    want = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, inp.read_len)
    pos, rc, mm, matched, _, _ = _run_routed_on_one_gpu(inp, 4)
    assert np.array_equal(pos, want.pos) and np.array_equal(rc, want.rc) and np.array_equal(mm, want.mm) and matched == want.matched
    tiny = synth.MatcherInputs(g, synth.sample_reads(g, 3, 100, 0.01, rng), np.zeros((0, 100), np.uint8), 100, "tiny")
    want = oracle.oracle_map_reads(tiny.text, tiny.lq_packed, None, 100)
    pos, rc, mm, matched, _, _ = _run_routed_on_one_gpu(tiny, 8)
    assert np.array_equal(pos, want.pos) and np.array_equal(rc, want.rc) and np.array_equal(mm, want.mm)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("kw", [dict(), dict(pre_seed=100), dict(mode="D"), dict(seed=45, rev_compl=False)])
def test_text_sharded_contexts_match_oracle(world, kw):
    for inp in (synth.adversarial(31, 100, n_reads=2000, text_len=30000), synth.workload(300_000, 40_000, 150, 0.005, seed=33, name="c2 shape")):
        want = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, inp.read_len, **kw)
        for got in _run_sharded_on_one_gpu(inp, world, **kw):
            bad = np.nonzero((got.pos != want.pos) | (got.rc != want.rc) | (got.mm != want.mm))[0]
            assert bad.size == 0, (f"{inp.name} world {world} {kw}: {bad.size} reads differ, first {bad[:5]}: gpu {got.pos[bad[:5]]} "
                                   f"{got.rc[bad[:5]]} {got.mm[bad[:5]]} oracle {want.pos[bad[:5]]} {want.rc[bad[:5]]} {want.mm[bad[:5]]}")
            assert got.matched == want.matched


def test_two_step_table_build_forced(monkeypatch):
    """The region-queued table build (used for tables beyond the L2) forced onto small inputs, including hot seeds
    that overflow a region queue and fall back to direct inserts."""
    monkeypatch.setenv("PGM_TWO_STEP_BUILD", "2")
    for inp in (synth.adversarial(41, 100), synth.adversarial(42, 150), synth.workload(200_000, 50_000, 150, 0.005, seed=43, n_frac=0.02, name="c2 shape")):
        _check(inp)
        _check(inp, pre_reads_exact_matching_chars=inp.read_len)
    rng = np.random.default_rng(44)
    g = synth.random_genome(20_000, rng)
    dup = np.repeat(synth.sample_reads(g, 3, 100, 0.01, rng), 4000, axis=0)      # 3 reads x 4000 copies: skewed regions
    inp = synth.MatcherInputs(g, np.concatenate([dup, synth.sample_reads(g, 2000, 100, 0.01, rng)]), np.zeros((0, 100), np.uint8), 100, "skew")
    _check(inp)


@pytest.mark.parametrize("mode", ["2", "3"])
def test_blocked_scan_pipeline_forced(monkeypatch, mode):
    """The L2-blocked scan pipeline (filter -> probe by table region -> verify by read range; used for tables beyond
    the L2) forced onto small inputs.  Mode 3 shrinks the stage queues to 64 entries, so they overflow and the fused
    kernel redoes the pass (idempotent accumulators): both routes must give the oracle's result."""
    monkeypatch.setenv("PGM_BLOCKED_SCAN", mode)
    for inp in (synth.adversarial(51, 100), synth.adversarial(52, 150), synth.adversarial(53, 255), synth.adversarial(54, 64),
                synth.workload(200_000, 50_000, 150, 0.005, seed=55, n_frac=0.02, name="c2 shape"),
                synth.workload(300_000, 40_000, 100, 0.01, seed=56, name="c4 shape")):
        got, _ = _check(inp)
        assert got.stats["candidates"] > 0 and got.stats["filter_positives"] > 0
        _check(inp, pre_reads_exact_matching_chars=inp.read_len)
        _check(inp, matching_mode="D")
    for kw in (dict(reads_exact_matching_chars=30), dict(reads_exact_matching_chars=45), dict(reads_exact_matching_chars=100),
               dict(min_chars_per_mismatch=2), dict(rev_compl_pg=False), dict(pre_reads_exact_matching_chars=50)):
        _check(synth.adversarial(57, 100), **kw)


def test_blocked_scan_pipeline_hot_seeds_and_shards(monkeypatch):
    """Hot seeds (duplicates, homopolymers: chains behind a slot, skewed queues) and text shards through the pipeline."""
    monkeypatch.setenv("PGM_BLOCKED_SCAN", "2")
    rng = np.random.default_rng(58)
    g = synth.random_genome(30_000, rng)
    text = np.concatenate([g[:10_000], np.full(400, ord("A"), np.uint8), g[10_000:20_000], np.full(300, ord("T"), np.uint8), g[20_000:]])
    base = synth.sample_reads(g, 40, 100, 0.02, rng)
    polya = np.full((300, 100), ord("A"), np.uint8)
    polya[np.arange(300), rng.integers(0, 100, 300)] = ord("C")
    reads = np.concatenate([np.repeat(base, 60, axis=0), polya, synth.sample_reads(g, 500, 100, 0.01, rng)])
    reads = reads[rng.permutation(len(reads))]
    nn = synth.inject_n(np.repeat(base[:5], 40, axis=0), rng)
    inp = synth.MatcherInputs(np.ascontiguousarray(text), np.ascontiguousarray(reads), nn, 100, "hot seeds")
    _check(inp)
    inp = synth.adversarial(59, 100, n_reads=2000, text_len=30000)
    want = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, inp.read_len)
    for got in _run_sharded_on_one_gpu(inp, 3):
        assert np.array_equal(got.pos, want.pos) and np.array_equal(got.rc, want.rc) and np.array_equal(got.mm, want.mm)


@pytest.mark.parametrize("mode", ["2", "3"])
def test_partitioned_prefilter_forced(monkeypatch, mode):
    """The partitioned exact pre-filter (pgm_part.cuh: every window queued by table partition, partitions probed in the L2,
    hit bitmap in place of the fused kernel's hash + filter stage; used for pattern sets beyond the Bloom filter) forced onto
    small inputs.  Mode 3 shrinks the partition queues to 64 entries: the windows that do not fit are marked without a probe
    and the fused kernel's own probe stage sorts them out — both routes must give the oracle's result."""
    monkeypatch.setenv("PGM_PART_SCAN", mode)
    for inp in (synth.adversarial(151, 100), synth.adversarial(152, 150), synth.adversarial(153, 255), synth.adversarial(154, 64),
                synth.workload(200_000, 50_000, 150, 0.005, seed=155, n_frac=0.02, name="c2 shape"),
                synth.workload(300_000, 40_000, 100, 0.01, seed=156, name="c4 shape")):
        got, _ = _check(inp)
        assert got.stats["candidates"] > 0 and got.stats["filter_positives"] > 0
        _check(inp, pre_reads_exact_matching_chars=inp.read_len)
        _check(inp, matching_mode="D")
    for kw in (dict(reads_exact_matching_chars=30), dict(reads_exact_matching_chars=45), dict(reads_exact_matching_chars=100),
               dict(min_chars_per_mismatch=2), dict(rev_compl_pg=False), dict(pre_reads_exact_matching_chars=50)):
        _check(synth.adversarial(157, 100), **kw)


def test_partitioned_prefilter_hot_seeds_shards_and_host_text(monkeypatch):
    """Hot seeds (chains behind a slot, one partition far above the average), text shards, the group API and a host text
    that arrives in chunks through the partitioned pre-filter."""
    monkeypatch.setenv("PGM_PART_SCAN", "2")
    rng = np.random.default_rng(158)
    g = synth.random_genome(30_000, rng)
    text = np.concatenate([g[:10_000], np.full(400, ord("A"), np.uint8), g[10_000:20_000], np.full(300, ord("T"), np.uint8), g[20_000:]])
    base = synth.sample_reads(g, 40, 100, 0.02, rng)
    polya = np.full((300, 100), ord("A"), np.uint8)
    polya[np.arange(300), rng.integers(0, 100, 300)] = ord("C")
    reads = np.concatenate([np.repeat(base, 60, axis=0), polya, synth.sample_reads(g, 500, 100, 0.01, rng)])
    reads = reads[rng.permutation(len(reads))]
    nn = synth.inject_n(np.repeat(base[:5], 40, axis=0), rng)
    inp = synth.MatcherInputs(np.ascontiguousarray(text), np.ascontiguousarray(reads), nn, 100, "hot seeds")
    _check(inp)
    inp = synth.adversarial(159, 100, n_reads=2000, text_len=30000)
    want = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, inp.read_len)
    for got in _run_sharded_on_one_gpu(inp, 3):
        assert np.array_equal(got.pos, want.pos) and np.array_equal(got.rc, want.rc) and np.array_equal(got.mm, want.mm)
    # a text of several upload chunks (16 M bases each), host inputs: the forward pass scans while the text arrives
    big = synth.workload(14_000_000, 60_000, 100, 0.01, seed=160, name="40 Mbp host text")
    _check(big)


@pytest.mark.parametrize("seed,L", [(71, 100), (72, 150), (73, 120), (74, 64), (75, 255)])
def test_interleaved_mode_adversarial(seed, L):
    """Mode 'i' (InterleavedReadsApproxMatcher, ReadsMatchers.cpp:343-409): strided seeds, alignment = hit - j."""
    got, want = _check(synth.adversarial(seed, L), matching_mode="i")
    assert want.matched > 50


@pytest.mark.parametrize("kw", [
    dict(matching_mode="i", reads_exact_matching_chars=30),
    dict(matching_mode="i", reads_exact_matching_chars=33),
    dict(matching_mode="i", reads_exact_matching_chars=45),
    dict(matching_mode="i", reads_exact_matching_chars=20),                      # 5 seeds per read
    dict(matching_mode="i", reads_exact_matching_chars=100),                     # exact path
    dict(matching_mode="I"),                                                     # shortcut mode
    dict(matching_mode="i", pre_reads_exact_matching_chars=100),
    dict(matching_mode="i", pre_reads_exact_matching_chars=50, pre_matching_mode="i"),
    dict(matching_mode="d", pre_reads_exact_matching_chars=50, pre_matching_mode="i"),
    dict(matching_mode="I", pre_reads_exact_matching_chars=50, pre_matching_mode="d"),
    dict(matching_mode="i", min_chars_per_mismatch=2),
    dict(matching_mode="i", rev_compl_pg=False),
])
def test_interleaved_mode_parameter_matrix(kw):
    for s in (76, 77):
        _check(synth.adversarial(s, 100), **kw)


def test_interleaved_mode_workloads_edges_and_shards():
    _check(synth.workload(250_000, 20_000, 100, 0.001, seed=78, name="c1/20"), matching_mode="i")
    _check(synth.workload(500_000, 100_000, 150, 0.005, seed=79, n_frac=0.02, name="c2/100"), matching_mode="i")
    rng = np.random.default_rng(80)
    g = synth.random_genome(5000, rng)
    reads = synth.sample_reads(g, 300, 100, 0.01, rng)
    for text in (g[:99], g[:75], g[:76], g[:77], g[:100], g[:101]):       # around the seed span (2 x 38) and the read length
        _check(synth.MatcherInputs(np.ascontiguousarray(text), reads, np.zeros((0, 100), np.uint8), 100, f"text{len(text)}"), matching_mode="i")
    _check(synth.MatcherInputs(g, np.zeros((0, 100), np.uint8), synth.inject_n(reads[:50], rng), 100, "only N reads"), matching_mode="i")
    inp = synth.adversarial(81, 100, n_reads=2000, text_len=30000)
    want = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, inp.read_len, mode="i")
    for got in _run_sharded_on_one_gpu(inp, 3, mode="i"):
        assert np.array_equal(got.pos, want.pos) and np.array_equal(got.rc, want.rc) and np.array_equal(got.mm, want.mm)


@pytest.mark.parametrize("mode", ["d", "i"])
def test_mismatch_lists_match_the_oracle(mode):
    """pgm_get_mismatches (what the reference's export recomputes per matched read on the host, ReadsMatchers.cpp:548-558)
    against the oracle's restatement, on adversarial inputs with N reads, both strands, two-phase and a c2-shaped workload."""
    cases = [(synth.adversarial(95, 100), {}), (synth.adversarial(96, 150), {}), (synth.adversarial(97, 255), {}),
             (synth.adversarial(98, 100), dict(pre_seed=50)), (synth.workload(300_000, 60_000, 150, 0.005, seed=99, n_frac=0.02), {})]
    for inp, kw in cases:
        with matcher.GpuReadsMatcher(0) as m:
            m.set_text(inp.text)
            m.set_reads(inp.lq_packed, inp.n_packed if len(inp.n_reads) else None, inp.read_len)
            got = m.map_reads(38, 3, mode, kw.get("pre_seed", 0), mode)
            off, o, pg, rd = m.get_mismatches()
        want = oracle.oracle_mismatch_lists(inp.text, inp.lq_packed, inp.n_packed, inp.read_len, got.pos, got.rc, got.mm, variant=0)
        assert int(off[-1]) == int(got.mm[got.mm != 255].astype(np.int64).sum()) > 0
        assert np.array_equal(off, want[0]) and np.array_equal(o, want[1]) and np.array_equal(pg, want[2]) and np.array_equal(rd, want[3])
    # no reads matched / empty inputs
    rng = np.random.default_rng(5)
    g = synth.random_genome(3000, rng)
    other = synth.sample_reads(synth.random_genome(3000, rng), 20, 100, 0.0, rng, require_error=False)
    with matcher.GpuReadsMatcher(0) as m:
        m.set_text(g)
        m.set_reads(synth.pack_reads(other), None, 100)
        got = m.map_reads()
        off, o, pg, rd = m.get_mismatches()
    assert got.matched == 0 and not off.any() and o.size == 0


@pytest.mark.parametrize("pair", ["0", "2"])
def test_paired_filter_lookups_forced_on_and_off(monkeypatch, pair):
    """Paired pre-filter lookups (one gather per two adjacent text windows, every pattern entered under both roles): off,
    and forced on regardless of the load — seed lengths around the 32-bit word (exactly-hashed core of 64 - n bases),
    N reads, hot seeds, text shards, the blocked pipeline."""
    monkeypatch.setenv("PGM_FILTER_PAIR", pair)
    for seed, L in ((101, 100), (102, 150), (103, 64), (104, 255)):
        _check(synth.adversarial(seed, L))
    for kw in (dict(reads_exact_matching_chars=20), dict(reads_exact_matching_chars=32), dict(reads_exact_matching_chars=33),
               dict(reads_exact_matching_chars=45), dict(reads_exact_matching_chars=50), dict(reads_exact_matching_chars=51),
               dict(reads_exact_matching_chars=100), dict(matching_mode="D"), dict(pre_reads_exact_matching_chars=50),
               dict(matching_mode="i"), dict(rev_compl_pg=False)):
        _check(synth.adversarial(105, 100), **kw)
    _check(synth.workload(500_000, 100_000, 150, 0.005, seed=106, n_frac=0.02, name="c2/100"))
    inp = synth.adversarial(107, 100, n_reads=2000, text_len=30000)
    want = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, inp.read_len)
    for got in _run_sharded_on_one_gpu(inp, 3):
        assert np.array_equal(got.pos, want.pos) and np.array_equal(got.rc, want.rc) and np.array_equal(got.mm, want.mm)
    monkeypatch.setenv("PGM_BLOCKED_SCAN", "2")
    _check(synth.adversarial(108, 100))


@pytest.mark.parametrize("seed,L", [(121, 100), (122, 150), (123, 120), (124, 64), (125, 255)])
def test_copmem_mode_adversarial(seed, L):
    """Mode 'c' (CopMEMReadsApproxMatcher, what the release CLI runs): text index + per-read query, the reference's results at -t 1."""
    got, want = _check(synth.adversarial(seed, L), matching_mode="c")
    assert want.matched > 50


@pytest.mark.parametrize("kw", [
    dict(matching_mode="c", reads_exact_matching_chars=30),
    dict(matching_mode="c", reads_exact_matching_chars=33),
    dict(matching_mode="c", reads_exact_matching_chars=45),
    dict(matching_mode="c", reads_exact_matching_chars=64),
    dict(matching_mode="c", reads_exact_matching_chars=100),
    dict(matching_mode="c", reads_exact_matching_chars=24),
    dict(matching_mode="C"),
    dict(matching_mode="c", pre_reads_exact_matching_chars=100, pre_matching_mode="c"),
    dict(matching_mode="c", pre_reads_exact_matching_chars=50, pre_matching_mode="c"),
    dict(matching_mode="c", pre_reads_exact_matching_chars=100, pre_matching_mode="d"),
    dict(matching_mode="d", pre_reads_exact_matching_chars=50, pre_matching_mode="c"),
    dict(matching_mode="i", pre_reads_exact_matching_chars=50, pre_matching_mode="c"),
    dict(matching_mode="c", min_chars_per_mismatch=2),
    dict(matching_mode="c", rev_compl_pg=False),
])
def test_copmem_mode_parameter_matrix(kw):
    for s in (126, 127):
        _check(synth.adversarial(s, 100), **kw)


def test_copmem_mode_workloads_and_edges():
    _check(synth.workload(250_000, 20_000, 100, 0.001, seed=128, name="c1/20"), matching_mode="c")
    _check(synth.workload(500_000, 100_000, 150, 0.005, seed=129, n_frac=0.02, name="c2/100"), matching_mode="c")
    _check(synth.workload(400_000, 60_000, 100, 0.01, seed=130, name="c4 scaled"), matching_mode="c")
    rng = np.random.default_rng(131)
    g = synth.random_genome(5000, rng)
    reads = synth.sample_reads(g, 300, 100, 0.01, rng)
    for text in (g[:99], g[:27], g[:28], g[:100], g[:101], g[:130]):       # around K = 28 and the read length
        _check(synth.MatcherInputs(np.ascontiguousarray(text), reads, np.zeros((0, 100), np.uint8), 100, f"text{len(text)}"), matching_mode="c")
    _check(synth.MatcherInputs(g, np.zeros((0, 100), np.uint8), synth.inject_n(reads[:50], rng), 100, "only N reads"), matching_mode="c")
    # hot seeds: more than 13 text positions per hash value (bucket cap) and the false-match budget
    dup = np.repeat(reads[:10], 40, axis=0)
    text = np.concatenate([g[:1500], np.full(3000, ord("A"), np.uint8), g[1500:3000], np.tile(g[200:420], 30)])
    areads = np.full((200, 100), ord("A"), np.uint8)
    areads[np.arange(200), rng.integers(0, 100, 200)] = ord("C")
    rep = synth.sample_reads(np.tile(g[200:420], 3), 300, 100, 0.02, rng)
    inp = synth.MatcherInputs(np.ascontiguousarray(text), np.concatenate([dup, areads, rep]), np.zeros((0, 100), np.uint8), 100, "hot")
    _check(inp, matching_mode="c")


def test_copmem_staged_query_forced(monkeypatch):
    """PGM_CM_WARP=1: mode 'c' through the staged per-read query (pgm_copmem_warp.cuh: a warp per read gathers the candidates of
    all offsets and verifies each distinct alignment once, a thread per read replays the reference's sequential loop over the
    table indices; reads with more than 32 distinct alignments go to the thread-per-read kernel) — same results as the oracle
    on the adversarial inputs, the parameter matrix, N reads, texts around K, hot seeds (table overflow) and a workload."""
    monkeypatch.setenv("PGM_CM_WARP", "1")
    for seed, L in ((121, 100), (122, 150), (123, 120), (124, 64), (125, 255)):
        _check(synth.adversarial(seed, L), matching_mode="c")
    for kw in (dict(matching_mode="c", reads_exact_matching_chars=30), dict(matching_mode="c", reads_exact_matching_chars=64),
               dict(matching_mode="c", reads_exact_matching_chars=24), dict(matching_mode="c", reads_exact_matching_chars=100),
               dict(matching_mode="C"), dict(matching_mode="c", pre_reads_exact_matching_chars=50, pre_matching_mode="c"),
               dict(matching_mode="c", pre_reads_exact_matching_chars=100, pre_matching_mode="d"),
               dict(matching_mode="c", min_chars_per_mismatch=2), dict(matching_mode="c", rev_compl_pg=False)):
        _check(synth.adversarial(126, 100), **kw)
    test_copmem_mode_workloads_and_edges()


@pytest.mark.parametrize("chunk", range(4))
def test_randomized_sweep_over_modes_and_parameters(chunk):
    """40 random (input, parameter) combinations — read lengths 40..255, seeds from 12 (24 for CopMEM) to beyond the read length,
    all mode letters in both phases, with and without the RC pass — CUDA path against the oracle (the same sweep pins the oracle
    against the reference's classes in tests/test_oracle_vs_reference.py)."""
    rng = np.random.default_rng(2000 + chunk)
    for _ in range(10):
        L = int(rng.choice([40, 50, 64, 75, 100, 101, 125, 150, 200, 255]))
        mode = str(rng.choice(list("dDiIcC")))
        kw = dict(matching_mode=mode, reads_exact_matching_chars=int(rng.integers(24 if mode.lower() == "c" else 12, L + 20)),
                  min_chars_per_mismatch=int(rng.choice([2, 3, 3, 4, 6, 10])), rev_compl_pg=bool(rng.random() < 0.8))
        if rng.random() < 0.4:
            pre_mode = str(rng.choice(list("dDiIcC")))
            kw.update(pre_matching_mode=pre_mode,
                      pre_reads_exact_matching_chars=int(rng.integers(24 if pre_mode.lower() == "c" else 12, L + 20)))
        inp = synth.adversarial(int(rng.integers(1 << 30)), L, n_reads=int(rng.integers(50, 400)), text_len=int(rng.integers(3000, 12000)))
        _check(inp, **kw)


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0]])
@pytest.mark.parametrize("kw", [dict(), dict(pre_seed=100), dict(mode="D"), dict(mode="i"), dict(mode="c"), dict(pre_seed=50, pre_mode="i", mode="c")])
def test_group_of_contexts_in_one_process(devices, kw):
    """pgm_group_* (what the C++ host side of PgRC calls): one handle over several contexts — listed on this one GPU more
    than once here, so the whole path (host threads, barriers, peer copies of the exchanges, global read order across the
    LQ / N boundary, result and mismatch-list assembly) runs on a single-GPU box."""
    for inp in (synth.adversarial(71, 100, n_reads=2000, text_len=30000), synth.workload(200_000, 30_000, 150, 0.005, seed=72, n_frac=0.03, name="c2 shape")):
        want = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, inp.read_len, **kw)
        plan = matcher.MatchPlan.derive(inp.read_len, kw.get("seed", 38), 3, kw.get("mode", "d"), kw.get("pre_seed", 0), kw.get("pre_mode", "d"))
        with matcher.GpuMatcherGroup(devices) as g:
            g.set_text(inp.text)
            g.set_reads(inp.lq_packed, inp.n_packed if len(inp.n_reads) else None, inp.read_len)
            g.run_plan(plan)
            got = g.get_results()
            lists = g.get_mismatches()
        bad = np.nonzero((got.pos != want.pos) | (got.rc != want.rc) | (got.mm != want.mm))[0]
        assert bad.size == 0, f"{inp.name} {devices} {kw}: {bad.size} reads differ, first {bad[:5]}"
        assert got.matched == want.matched and np.array_equal(got.per_mm, want.per_mm)
        wl = oracle.oracle_mismatch_lists(inp.text, inp.lq_packed, inp.n_packed, inp.read_len, got.pos, got.rc, got.mm, variant=0)
        assert all(np.array_equal(a, b) for a, b in zip(lists, wl)), "mismatch lists differ from the oracle"


@pytest.mark.parametrize("slices", ["1", "2", "3"])
def test_hash_sliced_filter_forced(monkeypatch, slices):
    """The hash-sliced filter (used for pattern sets far beyond an L2-resident filter: one scan launch per 64 MB slice of a
    4-bits-per-pattern filter) forced onto small inputs: same results, and — the filter being the same bits, only visited
    slice by slice — the same filter positives and candidates as the single-launch scan."""
    inputs = (synth.adversarial(81, 100), synth.adversarial(82, 150), synth.workload(200_000, 50_000, 150, 0.005, seed=83, n_frac=0.02, name="c2 shape"))
    monkeypatch.setenv("PGM_FILTER_PAIR", "0")        # (paired lookups are another filter layout, never combined with slices)
    base = [_check(inp)[0] for inp in inputs]
    monkeypatch.setenv("PGM_FILTER_SLICES", slices)
    for inp, b in zip(inputs, base):
        got, _ = _check(inp)
        assert got.stats["filter_positives"] == b.stats["filter_positives"] and got.stats["candidates"] == b.stats["candidates"]
        _check(inp, pre_reads_exact_matching_chars=inp.read_len)
        _check(inp, matching_mode="D")
