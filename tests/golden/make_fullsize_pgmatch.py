"""Full-size pin for stage 7: the REFERENCE's own CopMEMMatcher (one thread: its deterministic index build) on the texts
tools/pgmatch_bench.py builds from a bench workload — the source pseudogenome against its own reverse complement and against
the derived destination — reduced to count + sha256 of the raw resMatches vectors.  The GPU box re-creates the same texts
(counter-based generator, integer arithmetic) and `tools/pgmatch_bench.py --fixture` compares.

    python tests/golden/make_fullsize_pgmatch.py [workload=c2] [scale=1.0]
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import oracle  # noqa: E402
import pgmatch_bench  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    import torch
    src, dest, dest_rc = pgmatch_bench.stage7_texts(workload, scale, torch.device("cpu"))
    src, dest_rc = src.numpy(), dest_rc.numpy()
    out = {"workload": workload, "scale": scale, "target_len": 45, "src_bases": int(src.size), "dest_bases": int(dest_rc.size)}
    t0 = time.time()
    secs = [0, 0]
    r_self = oracle.ref_match_texts(src, oracle.reverse_complement(src), True, True, 45, threads=1, seconds=secs)
    out["self_rc"] = pgmatch_bench.result_digest(r_self)
    out["reference_seconds"] = {"index": round(secs[0], 3), "self_rc": round(secs[1], 3)}
    r_lq = oracle.ref_match_texts(src, dest_rc, False, True, 45, threads=1, seconds=secs)
    out["lq"] = pgmatch_bench.result_digest(r_lq)
    out["reference_seconds"]["lq"] = round(secs[1], 3)
    # the sequential oracle agrees with the reference at this size
    o_self = oracle.oracle_match_texts(src, oracle.reverse_complement(src), True, True, 45)
    o_lq = oracle.oracle_match_texts(src, dest_rc, False, True, 45)
    out["oracle_equal"] = bool(pgmatch_bench.result_digest(o_self) == out["self_rc"] and pgmatch_bench.result_digest(o_lq) == out["lq"])
    name = f"pgmatch_fullsize_{workload}.json" if scale == 1.0 else f"pgmatch_fullsize_{workload}_x{scale:g}.json"
    json.dump(out, open(os.path.join(HERE, name), "w"), indent=1)
    print(json.dumps(out), f"({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
