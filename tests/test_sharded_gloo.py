"""CPU, world_size 2, gloo: the host logic of the text-sharded path (shard_plan -> per-rank scan ->
merge_accumulators over torch.distributed -> per-pass decision on every rank) reproduces the oracle.
The per-rank kernels are replaced by the plain-Python model in tests/cpu_shard_model.py; the merge, the
shard plan and the phase plan are the product's own code."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, kw, ret_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from cpu_shard_model import CpuShardMatcher
    from pgrc_b200 import matcher, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inp = synth.adversarial(case["seed"], case["L"], n_reads=case["n_reads"], text_len=case["text_len"])
    m = CpuShardMatcher()
    pg_len = inp.text.size
    sb, sl, ob, oe = matcher.shard_plan(pg_len, rank, world)
    m.set_text_shard(inp.text[sb:sb + sl], sb, pg_len, ob, oe)
    m.set_reads(inp.lq_reads, inp.n_reads, inp.read_len)
    plan = matcher.MatchPlan.derive(inp.read_len, kw.get("seed", 38), kw.get("min_chars_per_mismatch", 3), kw.get("mode", "d"),
                                    kw.get("pre_seed", 0), kw.get("pre_mode", "d"))
    matcher.run_plan_sharded(m, plan, kw.get("rev_compl", True))
    res = m.get_results()
    np.savez(os.path.join(ret_dir, f"rank{rank}.npz"), pos=res.pos, rc=res.rc, mm=res.mm)
    dist.destroy_process_group()


@pytest.mark.parametrize("kw", [dict(), dict(pre_seed=100), dict(mode="D"), dict(seed=45, rev_compl=False), dict(mode="i"),
                                dict(mode="i", pre_seed=50, pre_mode="i")])
def test_text_sharded_two_ranks_gloo(tmp_path, kw):
    import oracle
    from pgrc_b200 import synth
    case = dict(seed=77, L=100, n_reads=240, text_len=5000)
    mp.spawn(_worker, args=(2, _free_port(), case, kw, str(tmp_path)), nprocs=2, join=True)
    inp = synth.adversarial(case["seed"], case["L"], n_reads=case["n_reads"], text_len=case["text_len"])
    want = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, inp.read_len, **kw)
    assert want.matched > 20
    for rank in range(2):
        got = np.load(tmp_path / f"rank{rank}.npz")
        assert np.array_equal(got["pos"], want.pos), f"rank {rank}"
        assert np.array_equal(got["rc"], want.rc) and np.array_equal(got["mm"], want.mm)


def test_match_plan_mirrors_reference_parameter_derivation():
    from pgrc_b200.matcher import MatchPlan, PgmError
    # (ReadsMatchers.cpp:699-713,749-756)
    F = False
    assert MatchPlan.derive(100, 38, 3, "d").phases == [(38, 2, 33, 0, F, F)]
    assert MatchPlan.derive(150, 38, 3, "d").phases == [(38, 3, 50, 0, F, F)]
    assert MatchPlan.derive(100, 38, 3, "D").phases == [(38, 2, 33, 33, F, F)]
    assert MatchPlan.derive(100, 100, 3, "d").phases == [(100, 1, 0, 0, F, F)]
    assert MatchPlan.derive(100, 250, 3, "d").phases == [(100, 1, 0, 0, F, F)]
    assert MatchPlan.derive(100, 38, 3, "d", pre_seed=100).phases == [(100, 1, 0, 0, F, F), (38, 2, 33, 1, True, F)]
    assert MatchPlan.derive(100, 38, 3, "d", pre_seed=50).phases == [(50, 2, 33, 0, F, F), (38, 2, 33, 2, True, F)]
    assert MatchPlan.derive(100, 38, 3, "D", pre_seed=50).phases == [(50, 2, 33, 0, F, F), (38, 2, 33, 33, True, F)]
    # mode 'i': the interleaved matcher wherever the reference constructs it (:728-731, :760-763)
    assert MatchPlan.derive(150, 38, 3, "i").phases == [(38, 3, 50, 0, F, True)]
    assert MatchPlan.derive(100, 100, 3, "i").phases == [(100, 1, 0, 0, F, F)]
    assert MatchPlan.derive(100, 38, 3, "I", pre_seed=50, pre_mode="d").phases == [(50, 2, 33, 0, F, F), (38, 2, 33, 33, True, True)]
    assert MatchPlan.derive(100, 38, 3, "d", pre_seed=50, pre_mode="i").phases == [(50, 2, 33, 0, F, True), (38, 2, 33, 2, True, F)]
    assert MatchPlan.derive(100, 38, 3, "c").phases == [(38, 2, 33, 0, F, "c")]
    assert MatchPlan.derive(100, 100, 3, "c").phases == [(100, 1, 33, 0, F, "c")]
    assert MatchPlan.derive(100, 38, 3, "c", pre_seed=50, pre_mode="d").phases == [(50, 2, 33, 0, F, F), (38, 2, 33, 2, True, "c")]
    for bad in ("x", "dd"):
        with pytest.raises(PgmError):
            MatchPlan.derive(100, 38, 3, bad)


def _gather_worker(rank, world, port, pg_len):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from pgrc_b200 import matcher
    dist.init_process_group("gloo", rank=rank, world_size=world)
    text = torch.from_numpy(np.random.default_rng(1).choice(np.frombuffer(b"ACGT", np.uint8), pg_len))
    b, e, per = matcher.text_share(pg_len, rank, world)
    bufs = {}
    for _ in range(2):   # second call reuses the buffers
        got = matcher.all_gather_text(text[b:e], pg_len, rank, world, "cpu", bufs)
        assert torch.equal(got, text), rank
    dist.destroy_process_group()


@pytest.mark.parametrize("pg_len", [1003, 4096])
def test_text_all_gather_three_ranks_gloo(pg_len):
    """Read-sharded runs: every rank uploads 1/N of the pseudogenome, one all-gather replicates it."""
    mp.spawn(_gather_worker, args=(3, _free_port(), pg_len), nprocs=3, join=True)


def _worker_2d(rank, world, port, case, T, ret_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from cpu_shard_model import CpuShardMatcher
    from pgrc_b200 import matcher, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    group = None
    for r0 in range(0, world, T):                 # rank = r * T + t, as in bench.py --shard 2d
        g = dist.new_group(list(range(r0, r0 + T)))
        if r0 <= rank < r0 + T:
            group = g
    t, r, R = rank % T, rank // T, world // T
    inp = synth.adversarial(case["seed"], case["L"], n_reads=case["n_reads"], text_len=case["text_len"], with_n=False)
    n = len(inp.lq_reads)
    lo, hi = (n * r) // R, (n * (r + 1)) // R
    m = CpuShardMatcher()
    pg_len = inp.text.size
    sb, sl, ob, oe = matcher.shard_plan(pg_len, t, T)
    m.set_text_shard(inp.text[sb:sb + sl], sb, pg_len, ob, oe)
    m.set_reads(inp.lq_reads[lo:hi], None, inp.read_len)
    matcher.run_plan_sharded(m, matcher.MatchPlan.derive(inp.read_len, 38, 3, "d"), True, group)
    res = m.get_results()
    np.savez(os.path.join(ret_dir, f"rank{rank}.npz"), pos=res.pos, rc=res.rc, mm=res.mm, lo=lo, hi=hi)
    dist.destroy_process_group()


def test_2d_sharded_four_ranks_gloo(tmp_path):
    """2 text ranges x 2 read groups: the accumulators are merged inside each group of ranks that share the reads."""
    import oracle
    from pgrc_b200 import synth
    case = dict(seed=78, L=100, n_reads=240, text_len=5000)
    mp.spawn(_worker_2d, args=(4, _free_port(), case, 2, str(tmp_path)), nprocs=4, join=True)
    inp = synth.adversarial(case["seed"], case["L"], n_reads=case["n_reads"], text_len=case["text_len"], with_n=False)
    want = oracle.oracle_map_reads(inp.text, inp.lq_packed, None, inp.read_len)
    assert want.matched > 20
    for rank in range(4):
        got = np.load(tmp_path / f"rank{rank}.npz")
        lo, hi = int(got["lo"]), int(got["hi"])
        assert np.array_equal(got["pos"], want.pos[lo:hi]), f"rank {rank}"
        assert np.array_equal(got["rc"], want.rc[lo:hi]) and np.array_equal(got["mm"], want.mm[lo:hi])


def _worker_routed(rank, world, port, case, kw, round_windows, ret_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from cpu_route_model import CpuRouteMatcher
    from pgrc_b200 import matcher, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inp = synth.adversarial(case["seed"], case["L"], n_reads=case["n_reads"], text_len=case["text_len"])
    reads = list(inp.lq_reads) + list(inp.n_reads)
    rb = matcher.read_ranges(len(reads), world)
    m = CpuRouteMatcher()
    m.set_text(inp.text)
    m.set_reads(reads[rb[rank]:rb[rank + 1]], None, inp.read_len)
    plan = matcher.MatchPlan.derive(inp.read_len, kw.get("seed", 38), kw.get("min_chars_per_mismatch", 3), kw.get("mode", "d"),
                                    kw.get("pre_seed", 0), kw.get("pre_mode", "d"))
    comm = matcher.TorchComm()
    info = matcher.run_plan_routed(m, plan, kw.get("rev_compl", True), comm, len(reads), round_windows,
                                   comm2=comm.sibling() if round_windows else None,      # several rounds: the pipelined schedules
                                   deep=bool(kw.get("pre_seed")))
    res = m.get_results()
    np.savez(os.path.join(ret_dir, f"rank{rank}.npz"), pos=res.pos, rc=res.rc, mm=res.mm, lo=rb[rank], hi=rb[rank + 1], rounds=info["rounds_per_pass"])
    dist.destroy_process_group()


@pytest.mark.parametrize("world,round_windows", [(2, 0), (3, 1024)])
@pytest.mark.parametrize("kw", [dict(), dict(pre_seed=100), dict(mode="D")])
def test_routed_ranks_gloo(tmp_path, world, round_windows, kw):
    """The routed scheme's host logic (run_plan_routed: counts exchange, all-to-all of patterns / windows / candidates over
    send/recv, rounds, local decision) over gloo; the per-rank kernels are the plain-Python model of tests/cpu_route_model.py."""
    import oracle
    from pgrc_b200 import synth
    case = dict(seed=79, L=100, n_reads=200, text_len=4000)
    mp.spawn(_worker_routed, args=(world, _free_port(), case, kw, round_windows, str(tmp_path)), nprocs=world, join=True)
    inp = synth.adversarial(case["seed"], case["L"], n_reads=case["n_reads"], text_len=case["text_len"])
    want = oracle.oracle_map_reads(inp.text, inp.lq_packed, inp.n_packed, inp.read_len, **kw)
    assert want.matched > 20
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}.npz")
        lo, hi = int(got["lo"]), int(got["hi"])
        assert np.array_equal(got["pos"], want.pos[lo:hi]), f"rank {rank}"
        assert np.array_equal(got["rc"], want.rc[lo:hi]) and np.array_equal(got["mm"], want.mm[lo:hi])
        assert (int(got["rounds"]) > 1) == bool(round_windows)


def _worker_pgmatch(rank, world, port, seed, ret_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import oracle
    from cpu_mem_model import CpuTextMatcher
    from pgrc_b200 import matcher, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = matcher.TorchComm()
    src, dest = synth.pg_texts(seed, 20000, 7000, max_copy=4000, self_rc=30)
    tm = CpuTextMatcher(src, 45)
    out = {}
    for tag, dis, rc in (("lq", False, True), ("fw", False, False), ("self", True, True)):
        d = src if dis else dest
        q = oracle.reverse_complement(d) if rc else d
        out[tag] = matcher.match_texts_distributed(tm, world, rank, comm.all_gather_arrays, q, dis, rc)
    np.savez(os.path.join(ret_dir, f"pgmatch_rank{rank}.npz"), **out)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_stage7_shares_merged_over_gloo(tmp_path, world):
    """Stage 7 with one process per GPU (matcher.match_texts_distributed): every rank takes its share of the groups of 256 query
    positions, the shares are all-gathered over torch.distributed and merged (merge_text_match_shares: the "covered by the previous
    match" test across the seams) — the per-rank kernels replaced by tests/cpu_mem_model.py, the host logic the product's own."""
    import oracle
    from pgrc_b200 import synth
    mp.spawn(_worker_pgmatch, args=(world, _free_port(), 970 + world, str(tmp_path)), nprocs=world, join=True)
    src, dest = synth.pg_texts(970 + world, 20000, 7000, max_copy=4000, self_rc=30)
    total = 0
    for tag, dis, rc in (("lq", False, True), ("fw", False, False), ("self", True, True)):
        d = src if dis else dest
        q = oracle.reverse_complement(d) if rc else d
        want = oracle.oracle_match_texts(src, q, dis, rc, 45)
        total += len(want)
        for rank in range(world):
            got = np.load(tmp_path / f"pgmatch_rank{rank}.npz")[tag]
            assert got.shape == want.shape and np.array_equal(got, want), (tag, rank)
    assert total > 20
