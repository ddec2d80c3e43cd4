"""GPU, end to end through the reference's own CLI: `PgRC-dev` built from the unmodified reference sources with the C++
shim of pgrc_b200/host/ (oracle/Makefile target `cli`) compresses the same synthetic FASTQ twice — once with the
reference's CPU matchers (mode d, i or c), once with the GPU matchers behind the same class interface
(PGRC_GPU_MATCHER=1) — and the two .pgrc archives must be byte-identical, for every archive mode of BASELINE.json's
configs (SE, SE_ORD, PE, PE_ORD) and for the two-phase / exact / shortcut parameterisations."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "oracle", "_ref", "PgRC-dev-gpu")


def _fastq(tmp, name, pair=False, **kw):
    out = os.path.join(tmp, name + "_1.fastq")
    cmd = [sys.executable, os.path.join(ROOT, "tools", "make_fastq.py"), out]
    out2 = None
    if pair:
        out2 = os.path.join(tmp, name + "_2.fastq")
        cmd += ["--pair", out2]
    for k, v in kw.items():
        cmd += ["--" + k.replace("_", "-"), str(v)]
    subprocess.run(cmd, check=True)
    return out, out2


def _compress(tmp, tag, gpu, fastq, fastq2, flags, devices=None, env_matcher="1", env_extra=None):
    d = os.path.join(tmp, tag)
    os.makedirs(d)
    env = dict(os.environ)
    for k in ("PGRC_GPU_MATCHER", "PGRC_GPU_DEVICES", "PGRC_GPU_DEVICE", "PGRC_GPU_PGMATCH"):
        env.pop(k, None)
    env.update(env_extra or {})
    if gpu and env_matcher is not None:
        env["PGRC_GPU_MATCHER"] = env_matcher
    if devices:
        env["PGRC_GPU_DEVICES"] = devices
    cmd = [CLI, "-t", "1"] + flags + ["-i", fastq] + ([fastq2] if fastq2 else []) + ["a.pgrc"]
    r = subprocess.run(cmd, cwd=d, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert ("(GPU)" in r.stdout or "on the GPU" in r.stdout) == gpu, r.stdout[-1500:]
    return open(os.path.join(d, "a.pgrc"), "rb").read(), r.stdout


CASES = {
    "SE_100bp": dict(pair=False, flags=["-s", "d38"], gen=dict(genome=300_000, reads=60_000, len=100, err=0.005, n_frac=0.01, seed=1)),
    "SE_ORD_150bp": dict(pair=False, flags=["-o", "-s", "d38"], gen=dict(genome=200_000, reads=40_000, len=150, err=0.005, n_frac=0.01, seed=2)),
    "PE_150bp": dict(pair=True, flags=["-s", "d38"], gen=dict(genome=200_000, reads=20_000, len=150, err=0.005, seed=3)),
    "PE_ORD_150bp": dict(pair=True, flags=["-o", "-s", "d38"], gen=dict(genome=200_000, reads=20_000, len=150, err=0.005, n_frac=0.005, seed=4)),
    "SE_shortcut": dict(pair=False, flags=["-s", "ds38"], gen=dict(genome=200_000, reads=40_000, len=100, err=0.01, seed=5)),
    "SE_two_phase": dict(pair=False, flags=["-l", "d50", "-s", "d33"], gen=dict(genome=200_000, reads=40_000, len=100, err=0.01, seed=6)),
    "SE_exact_prephase": dict(pair=False, flags=["-l", "d100", "-s", "d38"], gen=dict(genome=200_000, reads=40_000, len=100, err=0.003, seed=7)),
    "SE_M2": dict(pair=False, flags=["-M", "2", "-s", "d40"], gen=dict(genome=200_000, reads=40_000, len=120, err=0.02, seed=8)),
    # mode 'i' (InterleavedReadsApproxMatcher): the same shim classes with strided seeds
    "SE_ilv_100bp": dict(pair=False, flags=["-s", "i38"], gen=dict(genome=300_000, reads=60_000, len=100, err=0.005, n_frac=0.01, seed=9)),
    "SE_ORD_ilv_150bp": dict(pair=False, flags=["-o", "-s", "i38"], gen=dict(genome=200_000, reads=40_000, len=150, err=0.005, n_frac=0.01, seed=10)),
    "PE_ilv_150bp": dict(pair=True, flags=["-s", "i38"], gen=dict(genome=200_000, reads=20_000, len=150, err=0.005, seed=11)),
    "SE_ilv_two_phase": dict(pair=False, flags=["-l", "i50", "-s", "i33"], gen=dict(genome=200_000, reads=40_000, len=100, err=0.01, seed=12)),
    # mode 'c' (CopMEMReadsApproxMatcher) — what the CLI runs when no mode is given; -t 1: the reference's serial index build
    "SE_default_cli": dict(pair=False, flags=[], gen=dict(genome=300_000, reads=60_000, len=100, err=0.005, n_frac=0.01, seed=13)),
    "SE_ORD_copmem_150bp": dict(pair=False, flags=["-o", "-s", "c38"], gen=dict(genome=200_000, reads=40_000, len=150, err=0.005, n_frac=0.01, seed=14)),
    "PE_copmem_150bp": dict(pair=True, flags=["-s", "c38"], gen=dict(genome=200_000, reads=20_000, len=150, err=0.005, seed=15)),
    "PE_ORD_default_cli": dict(pair=True, flags=["-o"], gen=dict(genome=200_000, reads=20_000, len=150, err=0.005, n_frac=0.005, seed=16)),
    "SE_copmem_two_phase": dict(pair=False, flags=["-l", "c50", "-s", "c33"], gen=dict(genome=200_000, reads=40_000, len=100, err=0.01, seed=17)),
}


@pytest.mark.skipif(not os.path.exists(CLI), reason="oracle/_ref/PgRC-dev-gpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name", sorted(CASES))
def test_archive_bytes_identical_to_reference_cli(tmp_path, name):
    c = CASES[name]
    f1, f2 = _fastq(str(tmp_path), name, c["pair"], **c["gen"])
    ref, ref_out = _compress(str(tmp_path), "cpu", False, f1, f2, c["flags"])
    got, got_out = _compress(str(tmp_path), "gpu", True, f1, f2, c["flags"])
    assert len(ref) > 1000
    assert "Matched" in ref_out and "Matched" in got_out
    # stage 7 ran on the GPU as well (GpuTextMatcher behind SimplePgMatcher) and pushed the same number of matches per call
    assert "Pseudogenome text index on the GPU" in got_out and "Pseudogenome text index on the GPU" not in ref_out
    found = lambda out: [l.split(" exact matches")[0] for l in out.splitlines() if l.startswith("... found ")]
    assert len(found(ref_out)) == 3 and found(got_out) == found(ref_out)
    assert got == ref, f"{name}: archives differ ({len(got)} vs {len(ref)} bytes)"


@pytest.mark.skipif(not os.path.exists(CLI), reason="oracle/_ref/PgRC-dev-gpu not built (needs /root/reference at build time)")
def test_stage7_can_stay_on_the_cpu(tmp_path):
    """PGRC_GPU_PGMATCH=0: stage 4 on the GPU, stage 7 (SimplePgMatcher's text matcher) the reference's CopMEMMatcher."""
    c = CASES["PE_ORD_150bp"]
    f1, f2 = _fastq(str(tmp_path), "p", c["pair"], **c["gen"])
    ref, _ = _compress(str(tmp_path), "cpu", False, f1, f2, c["flags"])
    got, out = _compress(str(tmp_path), "gpu4", True, f1, f2, c["flags"], env_extra=dict(PGRC_GPU_PGMATCH="0"))
    assert "Pseudogenome text index on the GPU" not in out and "(GPU)" in out
    assert got == ref


MULTI = ["SE_100bp", "SE_ORD_150bp", "PE_ORD_150bp", "SE_two_phase", "SE_exact_prephase", "SE_ilv_100bp", "SE_default_cli"]


@pytest.mark.skipif(not os.path.exists(CLI), reason="oracle/_ref/PgRC-dev-gpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name", MULTI)
def test_archive_identical_with_the_stage_sharded_over_several_device_contexts(tmp_path, name):
    """PGRC_GPU_DEVICES: the C++ host side shards stage 4 over a group of device contexts (pgm_group_*: routed scheme for the
    hash-matcher modes, read ranges for i / c).  Two real GPUs when the box has them, else two contexts on GPU 0."""
    import torch
    c = CASES[name]
    f1, f2 = _fastq(str(tmp_path), name, c["pair"], **c["gen"])
    ref, _ = _compress(str(tmp_path), "cpu", False, f1, f2, c["flags"])
    devs = "0,1" if torch.cuda.device_count() >= 2 else "0,0"
    got, out = _compress(str(tmp_path), "gpu2", True, f1, f2, c["flags"], devices=devs)
    assert "sharded over 2 device contexts" in out
    assert "queries shared out over 2 device contexts" in out          # stage 7 (GpuTextMatcher) over the same device list
    assert got == ref, f"{name}: archives differ with PGRC_GPU_DEVICES={devs}"
    got3, _ = _compress(str(tmp_path), "gpu3", True, f1, f2, c["flags"], devices="0,0,0")
    assert got3 == ref


@pytest.mark.skipif(not os.path.exists(CLI), reason="oracle/_ref/PgRC-dev-gpu not built (needs /root/reference at build time)")
def test_no_silent_cpu_run(tmp_path):
    """A GPU request that cannot be served is an error, never a CPU run: a mistyped PGRC_GPU_MATCHER value, a device that
    does not exist.  (The in-band selection — mode letter g, `-s g38` — needs the one-line `case 'g'` in PgRC.cpp's option
    parser that INTEGRATION.md shows; this binary is built from the unmodified PgRC.cpp.)"""
    import torch
    c = CASES["SE_ORD_150bp"]
    f1, f2 = _fastq(str(tmp_path), "g", c["pair"], **c["gen"])
    for tag, env_add, needle in (("typo", dict(PGRC_GPU_MATCHER="yes"), "not understood"),
                                 ("nodev", dict(PGRC_GPU_MATCHER="1", PGRC_GPU_DEVICE="15"), "GPU matcher")):
        if tag == "nodev" and torch.cuda.device_count() > 15:
            continue
        d = os.path.join(str(tmp_path), tag)
        os.makedirs(d)
        r = subprocess.run([CLI, "-t", "1", "-s", "d38", "-i", f1, "a.pgrc"], cwd=d, env=dict(os.environ, **env_add), capture_output=True, text=True, timeout=600)
        assert r.returncode != 0 and needle in r.stderr, r.stderr[-500:]
        assert "Matched" not in r.stdout
