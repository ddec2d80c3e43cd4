#!/usr/bin/env python
"""Stage 7 (exact matches between pseudogenomes) on a bench-workload text: the HQ pseudogenome against its own reverse
complement (SimplePgMatcher::markAndRemoveExactMatches(true, hqPg, ..., revComplMatching = true), SimplePgMatcher.cpp:196)
and a destination text built from it (stretches of the source, reverse-complemented blockwise, with substitutions), GPU
(pgm_mem_*) and — on a bounded sample — the reference's CopMEMMatcher on the host cores.  Prints one JSON line.

    python tools/pgmatch_bench.py [--workload c2] [--scale 1.0] [--ref-scale 0.1] [--check]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stage7_texts(workload, scale, dev):
    """(source, destination, destination as SimplePgMatcher::exactMatchPg hands it to the matcher, SimplePgMatcher.cpp:39-41) for
    the bench workload at `scale`: the source is the bench text (counter-based generator); the destination has a third of its
    length — blocks of 2000 characters taken from the source at a stride of three blocks, every other one reverse-complemented,
    one substitution per 333 characters at counter-derived positions — and is then reverse-complemented as a whole.  Pure
    integer arithmetic in torch: the same bytes on any device."""
    import torch
    from pgrc_b200 import synth
    cfg = synth.scaled_config(workload, scale)
    gp = synth.hashed_params(**cfg, seed=20261017)
    src = synth.hashed_text(gp, 0, int(gp.text_len), dev)
    comp = torch.arange(256, dtype=torch.uint8, device=dev)
    for a, b in ("AT", "CG", "GC", "TA"):
        comp[ord(a)] = ord(b)
    blk = 2000
    nb = int(src.numel()) // (3 * blk)
    d = src[: nb * 3 * blk].view(nb, 3, blk)[:, 0, :].clone()
    d[1::2] = comp[d[1::2].long()].flip(1)
    dest = d.reshape(-1)
    idx = torch.arange(dest.numel(), device=dev, dtype=torch.int64)
    sub = ((idx * 2654435761) >> 7) % 333 == 0
    nxt = torch.zeros(256, dtype=torch.uint8, device=dev)
    for a, b in ("AC", "CG", "GT", "TA"):
        nxt[ord(a)] = ord(b)
    dest[sub] = nxt[dest[sub].long()]
    return src, dest, comp[dest.long()].flip(0).contiguous()


def result_digest(matches):
    import hashlib
    import numpy as np
    return {"count": int(len(matches)), "sha256": hashlib.sha256(np.ascontiguousarray(matches, dtype=np.uint64).tobytes()).hexdigest()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--ref-scale", type=float, default=0.1, help="fraction of the text the CPU reference is timed on (0 = skip)")
    ap.add_argument("--target", type=int, default=45, help="targetPgMatchLength (PgRC default 45)")
    ap.add_argument("--check", action="store_true", help="compare the GPU result with the sequential oracle (small scales only)")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--fixture", default="", help="JSON made by tests/golden/make_fullsize_pgmatch.py (the REFERENCE's own run of this "
                    "workload at one thread): compare count and sha256 of both result vectors")
    args = ap.parse_args()
    import torch
    from pgrc_b200 import matcher, synth
    dev = torch.device("cuda", 0)
    make_texts = lambda scale: stage7_texts(args.workload, scale, dev)

    src_d, dest_d, dest_rc_d = make_texts(args.scale)
    n = int(src_d.numel())
    torch.cuda.synchronize()

    m = matcher.GpuReadsMatcher(0)
    out = {"workload": args.workload, "scale": args.scale, "src_bases": n, "dest_bases": int(dest_d.numel()), "target_len": args.target}
    m.set_text(src_d)
    m.synchronize()
    def timed(fn):
        best = None
        for _ in range(args.reps):
            torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best, r
    t_index, tm = timed(lambda: matcher.GpuTextMatcher(None, args.target, matcher=m))
    out["params"] = {"K": tm.K, "k1": tm.k1, "k2": tm.k2, "hash_size": tm.hash_size}
    out["index_ms"] = round(t_index * 1e3, 3)
    t_self, r_self = timed(lambda: tm.match_texts(None, True, True))
    t_lq, r_lq = timed(lambda: tm.match_texts(dest_rc_d, False, True))
    out["self_rc"] = {"ms": round(t_self * 1e3, 3), "matches": int(len(r_self)), "gbases_per_s": round(n / t_self / 1e9, 3)}
    out["lq"] = {"ms": round(t_lq * 1e3, 3), "matches": int(len(r_lq)), "gbases_per_s": round(int(dest_d.numel()) / t_lq / 1e9, 3)}
    # host inputs (what the C++ shim passes): the upload of the destination is inside
    dest_rc_h = dest_rc_d.cpu().numpy()
    t_lq_h, r_lq_h = timed(lambda: tm.match_texts(dest_rc_h, False, True))
    assert np.array_equal(r_lq_h, r_lq)
    out["lq_host_input"] = {"ms": round(t_lq_h * 1e3, 3)}
    m.set_profiling(True); m.timings()
    matcher.GpuTextMatcher(None, args.target, matcher=m); tm.match_texts(None, True, True); tm.match_texts(dest_rc_d, False, True)
    out["kernel_ms"] = {k: round(v[0], 3) for k, v in m.timings().items() if v[1]}
    m.set_profiling(False)
    if args.fixture:
        want = json.load(open(args.fixture))
        assert want["workload"] == args.workload and want["scale"] == args.scale and want["target_len"] == args.target
        out["reference_full_run"] = {"fixture": os.path.relpath(args.fixture, ROOT),
                                     "self_equal": result_digest(r_self) == want["self_rc"], "lq_equal": result_digest(r_lq) == want["lq"]}
    if args.check:
        import oracle
        src_h = src_d.cpu().numpy()
        w_self = oracle.oracle_match_texts(src_h, oracle.reverse_complement(src_h), True, True, args.target)
        w_lq = oracle.oracle_match_texts(src_h, dest_rc_h, False, True, args.target)
        out["check"] = {"self_equal": bool(np.array_equal(w_self, r_self)), "lq_equal": bool(np.array_equal(w_lq, r_lq))}
    if args.ref_scale > 0:
        import oracle
        if oracle.have_ref():
            rs, _, rd = make_texts(args.scale * args.ref_scale)      # the same workload shape, scaled
            src_h, dst_h = rs.cpu().numpy(), rd.cpu().numpy()
            k = int(src_h.size)
            ref = {}
            for threads in (1, os.cpu_count() or 1):
                secs = [0, 0]
                r = oracle.ref_match_texts(src_h, oracle.reverse_complement(src_h), True, True, args.target, threads=threads, seconds=secs)
                secs2 = [0, 0]
                r2 = oracle.ref_match_texts(src_h, dst_h, False, True, args.target, threads=threads, seconds=secs2)
                ref[f"t{threads}"] = {"src_bases": int(k), "dest_bases": int(dst_h.size), "index_s": round(secs[0], 3), "self_rc_s": round(secs[1], 3),
                                      "lq_s": round(secs2[1], 3), "self_matches": int(len(r)), "lq_matches": int(len(r2)),
                                      "self_gbases_per_s": round(k / secs[1] / 1e9, 4), "lq_gbases_per_s": round(dst_h.size / secs2[1] / 1e9, 4)}
            out["reference_cpu"] = ref
    tm.close(); m.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
