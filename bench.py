#!/usr/bin/env python
"""bench.py — matching reads/sec of the read-vs-pseudogenome matcher (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c1..c5] [--scale F] [--shard text|reads|2d]

One "step" = one whole matcher invocation (what PgTools::mapReadsIntoPg does between
ReadsMatchers.cpp:714 and :783): text upload/packing, read upload/unpacking, seed-table build,
forward pass, reverse-complement pass, per-read decision, results out.  Prints ONE JSON line.

  value  reads/sec with the inputs (ASCII text, packed reads: the reference's own formats)
         already resident in HBM; results land in device buffers.
  e2e    the same call through the reference-facing API with HOST buffers (pinned): the H2D copy
         of text + reads and the D2H copy of the three result arrays are inside the timed region.
  roofline      the scan kernel (dominant): algorithmic bytes per launch / its CUDA-event time.
  cpu_baseline  the reference's own classes (oracle/_ref) on a bounded sample, on the host cores.

`--impl reference` times the reference's CPU matchers (oracle/_ref harness, unmodified reference
objects) on a bounded sample of the same workload shape; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "read-vs-pseudogenome matching reads/sec"
UNIT = "reads/s"
SEED = 20261017
MATCH_KW = dict(seed=38, min_chars_per_mismatch=3, mode="d")  # PgRCParams defaults (pgrc-params.h:138-146), hash-matcher path


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0, help="scale genome and read count (testing)")
    ap.add_argument("--mode", default="d", choices=["d", "i", "c"], help="matching mode: d = contiguous seeds (the headline), "
                    "i = interleaved seeds (InterleavedReadsApproxMatcher, SURVEY §8(f) rank 3), "
                    "c = CopMEM (CopMEMReadsApproxMatcher, what the release CLI runs; §8(f) rank 1)")
    ap.add_argument("--shard", default="auto", choices=["auto", "text", "reads", "2d"],
                    help="multi-GPU partitioning: text ranges + NCCL min-merge of the per-read keys, read ranges (no collective), or "
                         "2d = --text-shards T text ranges x N/T read groups (merge inside each group of T ranks); auto = reads, "
                         "or 2d with T = 2 for texts beyond 1 Gbase on >= 4 GPUs (the per-position stage then dominates; DESIGN.md §7)")
    ap.add_argument("--text-shards", type=int, default=0, help="T of --shard 2d (0 = auto: 2)")
    ap.add_argument("--cpu-sample", type=float, default=0.05, help="fraction of the workload shape timed on the CPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--export", action="store_true", help="also time the mismatch lists of the export step (SURVEY §8(f) rank 2): "
                    "pgm_get_mismatches into pinned host arrays at full size, next to the reference's own per-read loop "
                    "(updateEntry) on the CPU sample")
    ap.add_argument("--verify", action="store_true", help="after the timed region: re-count every reported alignment on the "
                    "device with plain torch ops (size-independent parity property for the full-size configs)")
    # kernel tuning knobs (pgm_set_tuning); defaults = the library's
    ap.add_argument("--filter-bits", type=int, default=-1)
    ap.add_argument("--slots-per-pattern", type=int, default=3)
    ap.add_argument("--ctas-per-sm", type=int, default=4)
    ap.add_argument("--l2-hints", type=int, default=1)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], None, set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax = float(c[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU legs
def cpu_sample_inputs(args, frac):
    """A smaller workload of the SAME shape (coverage, read length, error rate, text/genome ratio):
    genome and read count scaled by `frac`.  Generated with torch on the CPU."""
    import numpy as np
    from pgrc_b200 import synth
    c = synth.scaled_config(args.workload, args.scale * frac)
    text, packed = synth.workload_device(**c, seed=SEED + 1, device="cpu")
    text, packed = text.numpy(), packed.numpy()
    return c, text, packed, synth.unpack_reads_ascii(packed, c["read_len"])


def run_reference_cpu(args, frac, mode, threads, repeats=1):
    """Times the reference's own matcher classes (oracle/_ref) on the sample.  Returns (reads/s, dict)."""
    import oracle
    c, text, packed, ascii_reads = cpu_sample_inputs(args, frac)
    best = None
    for _ in range(repeats):
        r = oracle.ref_map_reads(text, ascii_reads, None, c["read_len"], seed=MATCH_KW["seed"],
                                 min_chars_per_mismatch=MATCH_KW["min_chars_per_mismatch"], mode=mode, threads=threads)
        best = r.seconds if best is None else min(best, r.seconds)
    info = {"reads": c["n_reads"], "text_bases": int(text.size), "seconds": round(best, 4), "matched": r.matched, "mode": mode,
            "threads": threads}
    return c["n_reads"] / best, info


def run_oracle_cpu(args, frac):
    import oracle
    c, text, packed, _ = cpu_sample_inputs(args, frac)
    t0 = time.perf_counter()
    r = oracle.oracle_map_reads(text, packed, None, c["read_len"], seed=MATCH_KW["seed"],
                                min_chars_per_mismatch=MATCH_KW["min_chars_per_mismatch"], mode=MATCH_KW["mode"])
    dt = time.perf_counter() - t0
    return c["n_reads"] / dt, {"reads": c["n_reads"], "text_bases": int(text.size), "seconds": round(dt, 4), "matched": r.matched}


def cpu_baseline(args) -> dict:
    """The reference's hash-matcher classes (mode 'd': the algorithm the CUDA path replaces; serial by
    construction) on a bounded sample; falls back to the C port if oracle/_ref is absent."""
    import oracle
    frac = args.cpu_sample
    shape = f"{args.workload} shape x {args.scale * frac:g} (genome, reads scaled; same read length, error rate, coverage)"
    if oracle.have_ref():
        cores = (os.cpu_count() or 1) if args.mode == "c" else 1       # mode c is the reference's one multithreaded matcher
        v, info = run_reference_cpu(args, frac, args.mode, cores)
        if args.mode == "c":
            return {"value": round(v, 1), "unit": UNIT, "cores": cores, "kind": "reference",
                    "sample": f"{shape}: {info['reads']} reads vs {info['text_bases']} bases, {info['seconds']} s, reference mode c "
                              f"(CopMEMReadsApproxMatcher, {cores} threads)"}
        return {"value": round(v, 1), "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": f"{shape}: {info['reads']} reads vs {info['text_bases']} bases, {info['seconds']} s, reference mode {args.mode} "
                          f"({'DefaultReadsApproxMatcher' if args.mode == 'd' else 'InterleavedReadsApproxMatcher'}, single-threaded by construction)"}
    v, info = run_oracle_cpu(args, frac)
    return {"value": round(v, 1), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{shape}: {info['reads']} reads vs {info['text_bases']} bases, {info['seconds']} s, oracle/pgrc_oracle.c"}


def reference_arm(args):
    """`--impl reference`: the reference's own CPU matchers on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    cores = os.cpu_count() or 1
    frac = args.cpu_sample
    total = args.steps + args.warmup
    t_start = time.perf_counter()
    line = {"metric": METRIC, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "gpu_launches": 0}
    if oracle.have_ref():
        # multithreaded matcher of the reference (mode 'c', CopMEM: what `./PgRC -i` runs) with all host threads,
        # and the hash-matcher classes (mode 'd', serial) the CUDA path restates; value = the faster of the two
        c, text, packed, ascii_reads = cpu_sample_inputs(args, frac)
        times = {"c": [], "d": []}
        matched = {}
        for i in range(total):
            for mode, thr in (("c", cores), ("d", 1)):
                if mode == "d" and i >= max(2, min(total, 3)):
                    continue   # serial mode d is slow: bounded to a few repetitions
                r = oracle.ref_map_reads(text, ascii_reads, None, c["read_len"], seed=MATCH_KW["seed"],
                                         min_chars_per_mismatch=MATCH_KW["min_chars_per_mismatch"], mode=mode, threads=thr)
                matched[mode] = r.matched
                if i >= args.warmup or mode == "d":
                    times[mode].append(r.seconds)
        tc = sum(times["c"]) / max(1, len(times["c"]))
        td = min(times["d"])
        v_c, v_d = c["n_reads"] / tc, c["n_reads"] / td
        v = max(v_c, v_d)
        kind, used_cores = "reference", (cores if v_c >= v_d else 1)
        sample = (f"{args.workload} shape x {args.scale * frac:g}: {c['n_reads']} reads vs {text.size} bases per step; "
                  f"mode c (CopMEM, {cores} threads): {v_c:.0f} reads/s, matched {matched['c']}; "
                  f"mode d (hash matcher, 1 thread, serial by construction): {v_d:.0f} reads/s, matched {matched['d']}")
        ms = 1e3 * (tc if v_c >= v_d else td)
    else:
        v, info = run_oracle_cpu(args, frac)
        kind, used_cores = "port", 1
        sample = f"{args.workload} shape x {args.scale * frac:g}: {info['reads']} reads vs {info['text_bases']} bases, oracle C port"
        ms = 1e3 * info["seconds"]
    line.update({"value": round(v, 1), "ms_per_step": round(ms, 3),
                 "config": {"workload": workload_name(args), "sample_fraction": args.scale * frac, "timing": "host wall clock inside the reference harness"},
                 "cpu_baseline": {"value": round(v, 1), "unit": UNIT, "cores": used_cores, "kind": kind, "sample": sample},
                 "e2e": {"value": round(v, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "wall_s": round(time.perf_counter() - t_start, 1)})
    print(json.dumps(line), flush=True)


def workload_name(args) -> str:
    from pgrc_b200 import synth
    c = synth.scaled_config(args.workload, args.scale)
    names = {"c1": "configs[0] SE 5 Mbp / 4M x 100 bp / 0.1%", "c2": "configs[1] SE_ORD 50 Mbp / 20M x 150 bp / 0.5%",
             "c3": "configs[2] PE 100 Mbp / 2x30M x 150 bp", "c4": "configs[3] SE 1 Gbp / 200M x 100 bp / 1%",
             "c5": "configs[4] PE_ORD 3 Gbp / 2x300M x 150 bp"}
    s = f"{names[args.workload]}: matcher input {c['n_reads']} LQ reads x {c['read_len']} bp vs ~{int(c['genome_len'] * c['copies'])} bp pseudogenome"
    if args.scale != 1.0:
        s += f" (scaled x{args.scale:g})"
    return s


# ------------------------------------------------------------------------------------------ GPU arm
def ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from pgrc_b200 import matcher, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the matcher has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = synth.scaled_config(args.workload, args.scale)
    if args.shard == "auto":
        # (mode c indexes the whole text on every GPU: it shards by reads only)
        args.shard = "2d" if (world >= 4 and cfg["genome_len"] * cfg["copies"] >= 1e9 and args.mode != "c") else "reads"
    T = 1
    if args.shard == "2d":
        T = args.text_shards or 2
        if world % T or T < 1:
            raise SystemExit(f"--text-shards {T} does not divide --gpus {world}")
        if T == 1:
            args.shard = "reads"
        elif T == world:
            args.shard = "text"
    group = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
        if args.shard == "2d":
            # rank = r * T + t: text range t of T, read group r of world / T; the per-read accumulators are merged inside a group
            for r0 in range(0, world, T):
                g = dist.new_group(list(range(r0, r0 + T)))
                if r0 <= rank < r0 + T:
                    group = g

    L = cfg["read_len"]
    text_d, lq_d = synth.workload_device(**cfg, seed=SEED, device=dev)   # same seed on every rank: replicated inputs
    n_reads, pg_len = lq_d.shape[0], text_d.numel()
    torch.cuda.synchronize()

    m = matcher.GpuReadsMatcher(local, use_torch_stream=True)
    m.set_tuning(args.filter_bits, args.slots_per_pattern, args.ctas_per_sm, args.l2_hints)
    plan = matcher.MatchPlan.derive(L, MATCH_KW["seed"], MATCH_KW["min_chars_per_mismatch"], MATCH_KW["mode"])

    # shard (N > 1)
    t_rank, r_rank, R = rank % T, rank // T, world // T
    if world > 1 and args.shard == "text":
        sb, sl, ob, oe = matcher.shard_plan(pg_len, rank, world)
        my_text_d = text_d[sb:sb + sl].contiguous()
        my_reads_d = lq_d
    elif world > 1 and args.shard == "2d":
        sb, sl, ob, oe = matcher.shard_plan(pg_len, t_rank, T)
        my_text_d = text_d[sb:sb + sl].contiguous()
        lo, hi = (n_reads * r_rank) // R, (n_reads * (r_rank + 1)) // R
        my_reads_d = lq_d[lo:hi].contiguous()
    elif world > 1:
        lo, hi = (n_reads * rank) // world, (n_reads * (rank + 1)) // world
        my_text_d, my_reads_d = text_d, lq_d[lo:hi].contiguous()
    else:
        my_text_d, my_reads_d = text_d, lq_d
    n_mine = my_reads_d.shape[0]
    out_d = (torch.empty(n_mine, dtype=torch.uint64, device=dev), torch.empty(n_mine, dtype=torch.uint8, device=dev),
             torch.empty(n_mine, dtype=torch.uint8, device=dev))

    def step(text, reads, out):
        if world > 1 and args.shard in ("text", "2d"):
            m.set_text_shard(text, sb, pg_len, ob, oe)
            m.set_reads(reads, None, L)
            matcher.run_plan_sharded(m, plan, True, group)
            return m.get_results(out)
        m.set_text(text)
        m.set_reads(reads, None, L)
        return m.map_reads(MATCH_KW["seed"], MATCH_KW["min_chars_per_mismatch"], MATCH_KW["mode"], out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = m.kernel_launches()
        e0.record()
        for _ in range(steps):
            res = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), m.kernel_launches() - l0, res

    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        res = step(my_text_d, my_reads_d, out_d)
    if rank == 0:
        sampler.start()
    dev_ms, launches, res = timed(lambda: step(my_text_d, my_reads_d, out_d), args.steps)
    value = n_reads * args.steps / (dev_ms * 1e-3)

    # end to end: pinned host buffers in, pinned host buffers out
    e2e = None
    if not args.no_e2e:
        gather_text = world > 1 and args.shard in ("reads", "2d")
        gather_reads = world > 1 and args.shard == "2d"
        if gather_text:
            # every rank uploads 1/N of the text over its own PCIe link; an NCCL all-gather over NVLink replicates it
            tb, te, _ = matcher.text_share(pg_len, rank, world)
            text_h = torch.empty(te - tb, dtype=torch.uint8, pin_memory=True); text_h.copy_(text_d[tb:te])
            gbufs = {}
        else:
            text_h = torch.empty(my_text_d.shape, dtype=torch.uint8, pin_memory=True); text_h.copy_(my_text_d)
        if gather_reads:
            # 2d: the T ranks of a read group upload 1/T of the group's packed reads each and all-gather them inside the group
            rbytes = my_reads_d.numel()
            rb, re_, _ = matcher.text_share(rbytes, t_rank, T)
            reads_h = torch.empty(re_ - rb, dtype=torch.uint8, pin_memory=True); reads_h.copy_(my_reads_d.view(-1)[rb:re_])
            rbufs = {}
        else:
            reads_h = torch.empty(my_reads_d.shape, dtype=torch.uint8, pin_memory=True); reads_h.copy_(my_reads_d)
        out_h = (torch.empty(n_mine, dtype=torch.uint64, pin_memory=True), torch.empty(n_mine, dtype=torch.uint8, pin_memory=True),
                 torch.empty(n_mine, dtype=torch.uint8, pin_memory=True))
        def step_e2e():
            text_in, reads_in = text_h, reads_h
            if gather_text:
                text_in = matcher.all_gather_text(text_h, pg_len, rank, world, dev, gbufs)
                if args.shard == "2d":
                    text_in = text_in[sb:sb + sl]
            if gather_reads:
                reads_in = matcher.all_gather_text(reads_h, rbytes, t_rank, T, dev, rbufs, group).view(my_reads_d.shape)
            return step(text_in, reads_in, out_h)
        for _ in range(min(args.warmup, 2)):
            step_e2e()
        e2e_ms, _, res_h = timed(step_e2e, args.steps)
        assert res_h.matched == res.matched
        e2e = {"value": round(n_reads * args.steps / (e2e_ms * 1e-3), 1), "unit": UNIT,
               "h2d_bytes_per_step": int(text_h.numel() + reads_h.numel()), "d2h_bytes_per_step": int(n_mine * 10),
               "ms_per_step": round(e2e_ms / args.steps, 3),
               "upload": ("text: 1/N per rank over PCIe + NCCL all-gather over NVLink" + ("; reads: 1/T per rank of a read group + all-gather inside the group" if gather_reads else "")) if gather_text else "whole inputs per rank"}
    clocks = sampler.stop() if rank == 0 else None

    # roofline of the dominant kernel (scan): separate short run with per-kernel events
    m.set_profiling(True)
    m.timings()
    psteps = 3
    for _ in range(psteps):
        res = step(my_text_d, my_reads_d, out_d)
    tm = m.timings()
    m.set_profiling(False)
    st = res.stats
    # a scan pass = one launch of the fused scan kernel, or the three stage kernels of the L2-blocked pipeline (+ the
    # fused kernel's no-op fallback launch); per-pass time = everything the pass launched
    blocked = tm["scan_filter"][1] > 0
    copmem = tm["copmem_query"][1] > 0
    scan_ms = tm["scan"][0] + tm["scan_filter"][0] + tm["scan_probe"][0] + tm["scan_verify"][0]
    scan_launches = tm["scan_filter"][1] if blocked else tm["scan"][1]
    if copmem:      # mode c: a pass = the text index (hash, scans, scatter, select) + the per-read query kernel
        scan_ms = tm["copmem_index"][0] + tm["copmem_query"][0]
        scan_launches = tm["copmem_query"][1]
    cand_per_launch = st["candidates"] / max(1, len(plan.phases) * 2)
    packed_len = lq_d.shape[1]
    my_pg = my_text_d.numel()
    # algorithmic bytes of ONE scan launch: 2-bit text once + packed read per candidate
    # + one 8-byte table slot per inserted pattern + one 8-byte key per read
    b_alg = my_pg / 4 + cand_per_launch * packed_len + 8 * st["patterns_inserted"] + 8 * n_mine
    if copmem:
        # mode c, one pass: 2-bit text once for the index + 4 bytes per sampled position written and read, the packed read
        # once, 8 bytes of bucket bounds per query offset, and per verified candidate its 4-byte entry + the text window
        K, k1, k2 = 28, 5, 2                                   # CopMEMMatcher parameters at seed 38 (CopMEMMatcher.cpp:71-137)
        lookups = n_mine * ((L - K) // k2 + 1)
        b_alg = my_pg / 4 + 2 * 4 * (my_pg / k1) + n_mine * packed_len + 8 * lookups + cand_per_launch * (4 + packed_len) + 8 * n_mine
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json"))).get((args.workload + ("_blocked" if blocked else "") + ("_mode_c" if copmem else "")) if args.scale == 1.0 else "", None)
    except OSError:
        pass
    achieved = b_alg / (scan_ms / max(1, scan_launches) * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    traffic_gbs = traffic / (scan_ms / max(1, scan_launches) * 1e-3) / 1e9 if (traffic and scan_ms > 0 and world == 1) else None
    roofline = {"bound": "hbm", "kernel": "copmem pass: index kernels + cm_query_kernel" if copmem else
                ("scan pass: scan_kernel<filter stage> + probe_kernel + verify_kernel" if blocked else "scan_kernel"), "achieved": round(achieved, 2), "peak": peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)", "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": traffic if world == 1 else None,
                "traffic_gbs": round(traffic_gbs, 1) if traffic_gbs else None,
                "traffic_frac_of_peak": round(traffic_gbs / peak, 4) if traffic_gbs else None,
                "note": "random 32/64-byte requests cost a full 128-byte DRAM line each (tools/ubench.cu): traffic >> algorithmic bytes",
                "algorithmic_bytes_per_launch": int(b_alg), "text_bytes_per_launch": int(my_pg / 4),
                "scan_ms_per_launch": round(scan_ms / max(1, scan_launches), 4),
                "text_positions_per_s": round(my_pg / (scan_ms / max(1, scan_launches) * 1e-3), 1) if scan_ms > 0 else None,
                "kernel_ms_per_step": {k: round(v[0] / psteps, 4) for k, v in tm.items()}}

    verify = None
    if args.verify:
        res_v = step(my_text_d, my_reads_d, out_d)
        torch.cuda.synchronize()
        verify = synth.check_matches_device(text_d, my_reads_d, L, out_d[0], out_d[1], out_d[2])
        verify["hist_matches_outputs"] = bool(int((out_d[2] != 255).sum().item()) == res_v.matched)
        if world > 1:
            once = args.shard == "reads" or (args.shard == "2d" and t_rank == 0) or (args.shard == "text" and rank == 0)
            vb = torch.tensor([verify["bad"], verify["matched"] if once else 0], device=dev, dtype=torch.int64)
            dist.all_reduce(vb)
            verify["bad"], verify["matched"] = int(vb[0].item()), int(vb[1].item())
    export = None
    if args.export and world == 1:
        import ctypes
        import numpy as np
        step(my_text_d, my_reads_d, out_d)
        total = ctypes.c_uint64()
        m._check(m._lib.pgm_get_mismatches(m._h, None, None, None, 0, ctypes.byref(total)))
        tot = int(total.value)
        off_h = torch.empty(n_mine + 1, dtype=torch.int64, pin_memory=True)
        pos_h = torch.empty(max(tot, 1), dtype=torch.uint8, pin_memory=True)
        sym_h = torch.empty(max(tot, 1), dtype=torch.uint8, pin_memory=True)
        def lists():
            m._check(m._lib.pgm_get_mismatches(m._h, off_h.data_ptr(), pos_h.data_ptr(), sym_h.data_ptr(), tot, ctypes.byref(total)))
        lists()
        m.set_profiling(True); m.timings()
        t0 = time.perf_counter()
        for _ in range(3):
            lists()
        wall_ms = (time.perf_counter() - t0) / 3 * 1e3
        k_ms = m.timings()["mismatches"][0] / 3
        m.set_profiling(False)
        export = {"what": "mismatch lists of all matched reads (offsets + 2 bytes per mismatch) into pinned host arrays",
                  "reads": n_mine, "mismatches": tot, "ms_per_call": round(wall_ms, 3), "kernel_ms": round(k_ms, 4),
                  "reads_per_s": round(n_mine / (wall_ms * 1e-3), 1), "d2h_bytes": int(8 * (n_mine + 1) + 2 * tot)}
        import oracle
        if oracle.have_ref() and not args.no_cpu_baseline:
            c, text_s, packed_s, ascii_s = cpu_sample_inputs(args, args.cpu_sample)
            r, *_ = oracle.ref_mismatch_lists(text_s, ascii_s, None, c["read_len"], mode=args.mode)
            export["cpu_reference"] = {"reads": c["n_reads"], "seconds": round(r.seconds, 4), "reads_per_s": round(c["n_reads"] / r.seconds, 1),
                                       "what": "the reference's updateEntry loop (getRead + reverse complement + compare + addMismatch, "
                                               "ReadsMatchers.cpp:548-558) on the CPU sample, 1 thread (serial in the reference)"}
    matched_total = res.matched
    if world > 1 and args.shard in ("reads", "2d"):
        mt = torch.tensor([res.matched if (args.shard == "reads" or t_rank == 0) else 0], device=dev, dtype=torch.int64)
        dist.all_reduce(mt)
        matched_total = int(mt.item())
    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(dev_ms / args.steps, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u32", "data": "synthetic",
                "config": {"workload": workload_name(args), "reads": n_reads, "read_len": L, "text_bases": pg_len,
                           "seed_len": plan.phases[0][0], "parts": plan.phases[0][1], "max_mismatches": plan.phases[0][2],
                           "matching_mode": MATCH_KW["mode"],
                           "matched": matched_total,
                           "parallelism": "single GPU" if world == 1 else (f"2d-sharded: {T} text ranges x {R} read groups" if args.shard == "2d"
                                                                           else f"{args.shard}-sharded x{world}"),
                           "l2": "inputs (text + reads + seed table) exceed the 126 MB L2; no flush between steps",
                           "candidates_per_step": st["candidates"], "filter_positives_per_step": st["filter_positives"],
                           "table_slots": st["table_slots"],
                           "tuning": {"filter_bits": args.filter_bits, "slots_per_pattern": args.slots_per_pattern,
                                      "ctas_per_sm": args.ctas_per_sm, "l2_hints": args.l2_hints}},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
        if verify is not None:
            line["verify"] = verify
        if export is not None:
            line["export"] = export
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    m.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    MATCH_KW["mode"] = args.mode
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
