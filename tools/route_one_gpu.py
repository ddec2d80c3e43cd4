#!/usr/bin/env python
"""The routed multi-GPU scheme with W contexts on ONE GPU (threads + tests/local_comm.py instead of NCCL), at a scale where
every context sees what a rank of an 8-GPU run of config 5 sees: W = 2 contexts at scale 0.25 -> 124 M patterns and
0.94 G windows per pass and context.  For profiling the routed kernels with ncu on a single-GPU box:
    ncu --set full -k regex:route_probe -c 2 ... python tools/route_one_gpu.py [world] [scale] [steps]
Prints per-kernel CUDA-event times of the last step."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from local_comm import LocalWorld  # noqa: E402
from pgrc_b200 import matcher, synth  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
fbits = int(os.environ.get("FILTER_BITS", "-1"))
cfg = synth.scaled_config("c5", scale)
p = synth.hashed_params(**cfg, seed=20261017)
dev = torch.device("cuda", 0)
text = synth.hashed_text(p, 0, None, dev)
n = cfg["n_reads"]
rb = matcher.read_ranges(n, world)
reads = [synth.hashed_reads(p, rb[r], rb[r + 1] - rb[r], dev) for r in range(world)]
plan = matcher.MatchPlan.derive(cfg["read_len"], 38, 3, "d")
ms = [matcher.GpuReadsMatcher(0, use_torch_stream=True) for _ in range(world)]
for m in ms:
    m.set_tuning(fbits, 3, 4, 1)


def body(rank, comm):
    m = ms[rank]
    m.set_text(text)
    m.set_reads(reads[rank], None, cfg["read_len"])
    info = matcher.run_plan_routed(m, plan, True, comm, n, int(os.environ.get("ROUND_WINDOWS", "0")))
    return m.get_results(), info


for s in range(steps):
    if s == steps - 1:
        for m in ms:
            m.set_profiling(True); m.timings()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = LocalWorld(world).run(body)
    torch.cuda.synchronize()
    print(f"step {s}: {1e3 * (time.perf_counter() - t0):.1f} ms wall for {world} contexts, matched {sum(o[0].matched for o in out)}", flush=True)
tm = ms[0].timings()
print("context 0 kernels (ms):", {k: round(v[0], 2) for k, v in tm.items() if v[1]})
print("context 0 stats:", out[0][0].stats, out[0][1])
