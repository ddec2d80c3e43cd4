"""GPU, BASELINE.json's configuration 2 at FULL size (10.45 M reads x 150 bp against a 140.6 Mbp pseudogenome; the
oracle needs minutes for this, the reference's mode d about four): parity through size-independent properties.
  1. every reported alignment is re-counted with plain torch ops: the read (reverse-complemented when RC is set) lies
     at `pos` with exactly `mm` mismatches; the matched count equals the histogram;
  2. the per-read rule does not depend on the other reads: a permutation of the reads permutes the results;
  3. text shards (two contexts, accumulators merged as the NCCL MIN / SUM all-reduces do) give the unsharded result;
  4. the L2-blocked scan pipeline gives the fused kernel's result;
  5. the sampled oracle: 20 000 reads of the full set, matched alone by the CPU oracle against the full text (in the build
     container: tests/golden/sample_<cfg>.npz, made by tests/golden/make_fullsize.py — the counter-based generator gives the
     same bytes there and here), must get the GPU's (pos, rc, mm) of the full run (per-read independence again);
  6. the REFERENCE's own full-size run (DefaultReadsApproxMatcher through oracle/_ref, tests/golden/fullsize_<cfg>_ref.json):
     matched count, per-mismatch histogram and sha256 of the three result arrays are identical — configs 2 and 3 (the
     paired-end shape: 32 M reads x 150 bp against 250 Mbp);
  7. the routed multi-GPU scheme (four contexts on this one GPU) reproduces the single-context result at full size."""
import hashlib
import json
import os


import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 20261017


@pytest.fixture(scope="module")
def c2():
    import torch
    from pgrc_b200 import matcher, synth
    return _run_config("c2")


def _run_config(name):
    import torch
    from pgrc_b200 import matcher, synth
    cfg = synth.scaled_config(name, 1.0)
    _, text, reads = synth.workload_hashed(name, 1.0, SEED, torch.device("cuda", 0))
    n = reads.shape[0]
    out = (torch.empty(n, dtype=torch.uint64, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda"),
           torch.empty(n, dtype=torch.uint8, device="cuda"))
    with matcher.GpuReadsMatcher(0, use_torch_stream=True) as m:
        m.set_text(text)
        m.set_reads(reads, None, cfg["read_len"])
        res = m.map_reads(out=out)
    torch.cuda.synchronize()
    return cfg, text, reads, out, res


def test_every_alignment_recounts(c2):
    from pgrc_b200 import synth
    cfg, text, reads, out, res = c2
    v = synth.check_matches_device(text, reads, cfg["read_len"], out[0], out[1], out[2])
    assert v["bad"] == 0, v
    assert v["matched"] == res.matched == int(res.per_mm[:255].sum())
    assert res.matched > 0.99 * reads.shape[0]
    assert res.stats["patterns_inserted"] == 3 * reads.shape[0]


def test_permuted_reads_permute_results(c2):
    import torch
    from pgrc_b200 import matcher
    cfg, text, reads, out, res = c2
    perm = torch.randperm(reads.shape[0], device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    with matcher.GpuReadsMatcher(0, use_torch_stream=True) as m:
        m.set_text(text)
        m.set_reads(reads[perm].contiguous(), None, cfg["read_len"])
        got = m.map_reads(out=tuple(torch.empty_like(o) for o in out))
    torch.cuda.synchronize()
    assert got.matched == res.matched
    assert torch.equal(got.pos.view(torch.int64), out[0].view(torch.int64)[perm])
    assert torch.equal(got.rc, out[1][perm]) and torch.equal(got.mm, out[2][perm])


def test_text_shards_and_blocked_pipeline_equal_the_fused_single_context(c2, monkeypatch):
    import torch
    from pgrc_b200 import matcher
    cfg, text, reads, out, res = c2
    L, pg_len = cfg["read_len"], text.numel()
    plan = matcher.MatchPlan.derive(L, 38, 3, "d")
    from local_comm import LocalWorld

    def rank_body(rank, comm):
        with matcher.GpuReadsMatcher(0, use_torch_stream=True) as m:
            sb, sl, ob, oe = matcher.shard_plan(pg_len, rank, 2)
            m.set_text_shard(text[sb:sb + sl].contiguous(), sb, pg_len, ob, oe)
            m.set_reads(reads, None, L)
            matcher.run_plan_sharded(m, plan, True, comm)        # the product's own merge (touched reduced in place)
            return m.get_results(tuple(torch.empty_like(o) for o in out))

    for got in LocalWorld(2).run(rank_body):
        assert got.matched == res.matched
        assert torch.equal(got.pos.view(torch.int64), out[0].view(torch.int64)) and torch.equal(got.rc, out[1]) and torch.equal(got.mm, out[2])
    monkeypatch.setenv("PGM_BLOCKED_SCAN", "1")
    with matcher.GpuReadsMatcher(0, use_torch_stream=True) as m:
        m.set_text(text)
        m.set_reads(reads, None, L)
        m.set_profiling(True)
        got = m.map_reads(out=tuple(torch.empty_like(o) for o in out))
        assert m.timings()["scan_probe"][1] == 2          # the pipeline did run (one probe launch per pass)
    assert got.matched == res.matched and got.stats["candidates"] == res.stats["candidates"]
    assert torch.equal(got.pos.view(torch.int64), out[0].view(torch.int64)) and torch.equal(got.rc, out[1]) and torch.equal(got.mm, out[2])


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _check_fixtures(name, out, res):
    """items 5 and 6 of the module docstring"""
    import torch
    z = np.load(os.path.join(GOLDEN, f"sample_{name}.npz"))
    assert int(z["seed"]) == SEED and int(z["n_reads"]) == out[0].numel()
    idx = torch.from_numpy(z["idx"].astype(np.int64)).cuda()
    pos = out[0].view(torch.int64)[idx].cpu().numpy().view(np.uint64)
    rc, mm = out[1][idx].cpu().numpy(), out[2][idx].cpu().numpy()
    bad = np.nonzero((pos != z["pos"]) | (rc != z["rc"]) | (mm != z["mm"]))[0]
    assert bad.size == 0, f"{name}: {bad.size} of {idx.numel()} sampled reads differ from the oracle, first read {z['idx'][bad[:3]]}"
    want = json.load(open(os.path.join(GOLDEN, f"fullsize_{name}_ref.json")))
    sha = lambda t: hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()
    assert res.matched == want["matched"]
    assert [int(x) for x in res.per_mm] == want["per_mm"]
    assert sha(out[0].view(torch.int64)) == want["sha256_pos"], "readMatchPos differs from the reference's full-size run"
    assert sha(out[1]) == want["sha256_rc"] and sha(out[2]) == want["sha256_mm"]


def test_c2_equals_sampled_oracle_and_reference_full_run(c2):
    cfg, text, reads, out, res = c2
    _check_fixtures("c2", out, res)


def test_c3_paired_end_shape_equals_sampled_oracle_and_reference_full_run():
    """BASELINE config 3 (PE, 100 Mbp genome, 2 x 30 M x 150 bp: 32 M LQ reads against a 250 Mbp pseudogenome) at full size."""
    from pgrc_b200 import synth
    cfg, text, reads, out, res = _run_config("c3")
    v = synth.check_matches_device(text, reads, cfg["read_len"], out[0], out[1], out[2])
    assert v["bad"] == 0 and v["matched"] == res.matched
    _check_fixtures("c3", out, res)


def test_c2_routed_four_contexts_equal_the_single_context(c2):
    import torch
    from local_comm import LocalWorld
    from pgrc_b200 import matcher
    cfg, text, reads, out, res = c2
    L, n, world = cfg["read_len"], reads.shape[0], 4
    plan = matcher.MatchPlan.derive(L, 38, 3, "d")
    rb = matcher.read_ranges(n, world)

    def rank_body(rank, comm):
        with matcher.GpuReadsMatcher(0, use_torch_stream=True) as m:
            m.set_text(text)
            m.set_reads(reads[rb[rank]:rb[rank + 1]].contiguous(), None, L)
            matcher.run_plan_routed(m, plan, True, comm, n, 16 << 20, comm2=comm.sibling(), exchange="pull")
            o = tuple(torch.empty(rb[rank + 1] - rb[rank], dtype=t.dtype, device="cuda") for t in out)
            return m.get_results(o)

    got = LocalWorld(world).run(rank_body)
    assert sum(g.matched for g in got) == res.matched
    assert torch.equal(torch.cat([g.pos for g in got]).view(torch.int64), out[0].view(torch.int64))
    assert torch.equal(torch.cat([g.rc for g in got]), out[1]) and torch.equal(torch.cat([g.mm for g in got]), out[2])
