"""GPU, BASELINE.json's configuration 2 at FULL size (10.45 M reads x 150 bp against a 140.6 Mbp pseudogenome; the
oracle needs minutes for this, the reference's mode d about four): parity through size-independent properties.
  1. every reported alignment is re-counted with plain torch ops: the read (reverse-complemented when RC is set) lies
     at `pos` with exactly `mm` mismatches; the matched count equals the histogram;
  2. the per-read rule does not depend on the other reads: a permutation of the reads permutes the results;
  3. text shards (two contexts, accumulators merged as the NCCL MIN / SUM all-reduces do) give the unsharded result;
  4. the L2-blocked scan pipeline gives the fused kernel's result;
  5. the sampled oracle: 20 000 reads drawn from the full set, matched alone by the CPU oracle against the full text,
     must get the GPU's (pos, rc, mm) of the full run (per-read independence again)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 20261017


@pytest.fixture(scope="module")
def c2():
    import torch
    from pgrc_b200 import matcher, synth
    cfg = synth.scaled_config("c2", 1.0)
    text, reads = synth.workload_device(**cfg, seed=SEED, device=torch.device("cuda", 0))
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
    ms = [matcher.GpuReadsMatcher(0, use_torch_stream=True) for _ in range(2)]
    try:
        for rank, m in enumerate(ms):
            sb, sl, ob, oe = matcher.shard_plan(pg_len, rank, 2)
            m.set_text_shard(text[sb:sb + sl].contiguous(), sb, pg_len, ob, oe)
            m.set_reads(reads, None, L)
        for seed_len, parts, max_mm, min_mm, cont, ilv in plan.phases:
            for m in ms:
                m.match_begin(seed_len, parts, max_mm, min_mm, cont, ilv)
            for rev in (False, True):
                for m in ms:
                    m.scan_pass(rev)
                accs = [m.accumulators() for m in ms]
                merged = {"best_key": torch.minimum(accs[0]["best_key"], accs[1]["best_key"]),
                          "first_other_order": torch.minimum(accs[0]["first_other_order"], accs[1]["first_other_order"]),
                          "same_pos_mask": accs[0]["same_pos_mask"] + accs[1]["same_pos_mask"],
                          "same_pos_mm": torch.minimum(accs[0]["same_pos_mm"], accs[1]["same_pos_mm"]),
                          "touched": torch.maximum(accs[0]["touched"], accs[1]["touched"])}
                for m, a in zip(ms, accs):
                    for k, v in merged.items():
                        a[k].copy_(v)
                    m.put_accumulators()
                    m.resolve_pass(rev)
        for m in ms:
            got = m.get_results(tuple(torch.empty_like(o) for o in out))
            assert got.matched == res.matched
            assert torch.equal(got.pos.view(torch.int64), out[0].view(torch.int64)) and torch.equal(got.rc, out[1]) and torch.equal(got.mm, out[2])
    finally:
        for m in ms:
            m.close()
    monkeypatch.setenv("PGM_BLOCKED_SCAN", "1")
    with matcher.GpuReadsMatcher(0, use_torch_stream=True) as m:
        m.set_text(text)
        m.set_reads(reads, None, L)
        m.set_profiling(True)
        got = m.map_reads(out=tuple(torch.empty_like(o) for o in out))
        assert m.timings()["scan_probe"][1] == 2          # the pipeline did run (one probe launch per pass)
    assert got.matched == res.matched and got.stats["candidates"] == res.stats["candidates"]
    assert torch.equal(got.pos.view(torch.int64), out[0].view(torch.int64)) and torch.equal(got.rc, out[1]) and torch.equal(got.mm, out[2])


def test_sampled_reads_against_the_oracle(c2):
    import oracle
    cfg, text, reads, out, res = c2
    rng = np.random.default_rng(3)
    idx = np.sort(rng.choice(reads.shape[0], 20_000, replace=False))
    sub = reads.cpu().numpy()[idx]
    want = oracle.oracle_map_reads(text.cpu().numpy(), sub, None, cfg["read_len"])
    pos, rc, mm = (o.cpu().numpy()[idx] for o in out)
    assert np.array_equal(pos, want.pos) and np.array_equal(rc, want.rc) and np.array_equal(mm, want.mm)
