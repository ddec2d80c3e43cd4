"""Synthetic matcher-level inputs (there is no network; BASELINE.md §3 / SURVEY.md §8(d)).

Everything here is host-side numpy: genomes, pseudogenome-like texts, error-containing
reads, and the reference's packed-read layout.  The shapes follow the
BASELINE configs; the matcher inputs are generated directly (text + the error-containing
subset of reads) because the full PgRC chain cannot produce them on the bench host.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_A, _C, _G, _T, _N = (ord(c) for c in "ACGTN")
_CODE2ASCII = np.array([_A, _C, _G, _T], np.uint8)
_COMP = np.arange(256, dtype=np.uint8)
_COMP[_A], _COMP[_C], _COMP[_G], _COMP[_T] = _T, _G, _C, _A
_ASCII2CODE4 = np.full(256, 255, np.uint8)
_ASCII2CODE4[[_A, _C, _G, _T]] = [0, 1, 2, 3]
_ASCII2CODE5 = np.full(256, 255, np.uint8)
_ASCII2CODE5[[_A, _C, _G, _N, _T]] = [0, 1, 2, 3, 4]


def random_genome(length: int, rng: np.random.Generator) -> np.ndarray:
    """i.i.d. uniform ACGT, ASCII uint8."""
    return _CODE2ASCII[rng.integers(0, 4, size=length, dtype=np.uint8)]


def revcomp(a: np.ndarray) -> np.ndarray:
    """Reverse complement along the last axis (N stays N)."""
    return _COMP[a[..., ::-1]]


def make_pseudogenome(genome: np.ndarray, rng: np.random.Generator, copies: float = 2.8,
                      mean_contig: int = 4000) -> np.ndarray:
    """Pseudogenome-like text: the genome cut into contigs, each emitted in a random
    orientation, repeated until the text is ``copies`` x the genome (the reference's HQ
    pseudogenome is 2.0-2.9 x the genome, both strands mixed)."""
    G = genome.size
    out = []
    remaining = copies
    while remaining > 1e-9:
        frac = min(1.0, remaining)
        span = max(1, int(G * frac))
        lo = 0 if span >= G else int(rng.integers(0, G - span + 1))
        ncut = max(0, span // max(1, mean_contig) - 1)
        cuts = np.unique(rng.integers(lo + 1, lo + span, size=ncut)) if ncut and span > 1 else np.empty(0, np.int64)
        bounds = np.concatenate(([lo], cuts, [lo + span])).astype(np.int64)
        flips = rng.random(bounds.size - 1) < 0.5
        for b, e, f in zip(bounds[:-1], bounds[1:], flips):
            seg = genome[b:e]
            out.append(revcomp(seg) if f else seg)
        remaining -= frac
    return np.ascontiguousarray(np.concatenate(out))


def sample_reads(genome: np.ndarray, n: int, read_len: int, err: float, rng: np.random.Generator,
                 require_error: bool = True, chunk: int = 1 << 20) -> np.ndarray:
    """n reads (n x L ASCII): uniform start, 50 % reverse-complemented, i.i.d. substitutions
    with probability ``err``.  With ``require_error`` every read carries >= 1 substitution —
    the matcher only sees the error-containing (LQ) subset."""
    G = genome.size
    out = np.empty((n, read_len), np.uint8)
    ar = np.arange(read_len, dtype=np.int64)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        start = rng.integers(0, G - read_len + 1, size=m, dtype=np.int64)
        r = genome[start[:, None] + ar[None, :]]
        flip = rng.random(m) < 0.5
        r[flip] = revcomp(r[flip])
        mask = rng.random((m, read_len)) < err
        if require_error:
            none = ~mask.any(axis=1)
            mask[np.nonzero(none)[0], rng.integers(0, read_len, size=int(none.sum()))] = True
        codes = _ASCII2CODE4[r]
        shift = rng.integers(1, 4, size=(m, read_len), dtype=np.uint8)
        codes = np.where(mask, (codes + shift) & 3, codes)
        out[s:s + m] = _CODE2ASCII[codes]
    return out


def inject_n(reads: np.ndarray, rng: np.random.Generator, max_n: int = 3) -> np.ndarray:
    """Replaces 1..max_n random bases of every read by 'N' (reads destined for the ACGNT set)."""
    r = reads.copy()
    n, L = r.shape
    for _ in range(max_n):
        sel = rng.random(n) < (1.0 if _ == 0 else 0.5)
        r[np.nonzero(sel)[0], rng.integers(0, L, size=int(sel.sum()))] = _N
    return r


def pack_reads(reads: np.ndarray, with_n: bool = False) -> np.ndarray:
    """The reference's packed-read layout (SymbolsPackingFacility.cpp:147-185): ACGT -> 4
    bases/byte, first base in the most significant digit, tail padded with A; ACGNT -> 3
    bases/byte, base 5, codes A0 C1 G2 N3 T4."""
    reads = np.ascontiguousarray(reads, np.uint8)
    n, L = reads.shape
    spe, sigma, lut = (3, 5, _ASCII2CODE5) if with_n else (4, 4, _ASCII2CODE4)
    plen = (L + spe - 1) // spe
    codes = lut[reads]
    if codes.size and codes.max() == 255:
        raise ValueError("symbol outside the alphabet")
    pad = plen * spe - L
    if pad:
        codes = np.concatenate([codes, np.zeros((n, pad), np.uint8)], axis=1)
    w = np.array([sigma ** (spe - 1 - j) for j in range(spe)], np.uint16)
    return (codes.reshape(n, plen, spe).astype(np.uint16) * w).sum(axis=2).astype(np.uint8)


@dataclass
class MatcherInputs:
    """What PgRCEncoder::runMappingLQReadsOnHQPg hands to mapReadsIntoPg (pgrc-encoder.cpp:342-366)."""
    text: np.ndarray            # ASCII pseudogenome
    lq_reads: np.ndarray        # n_lq x L ASCII (ACGT)
    n_reads: np.ndarray         # n_n x L ASCII (ACGNT), may be empty
    read_len: int
    name: str = ""

    @property
    def lq_packed(self) -> np.ndarray:
        return pack_reads(self.lq_reads, False)

    @property
    def n_packed(self) -> np.ndarray:
        return pack_reads(self.n_reads, True) if len(self.n_reads) else np.zeros((0, (self.read_len + 2) // 3), np.uint8)


def workload(genome_len: int, n_reads: int, read_len: int, err: float, seed: int,
             copies: float = 2.8, n_frac: float = 0.0, name: str = "") -> MatcherInputs:
    """Matcher-level workload of a BASELINE config shape (scaled by the caller)."""
    rng = np.random.default_rng(seed)
    genome = random_genome(genome_len, rng)
    text = make_pseudogenome(genome, rng, copies=copies)
    n_n = int(round(n_reads * n_frac))
    lq = sample_reads(genome, n_reads - n_n, read_len, err, rng)
    nn = inject_n(sample_reads(genome, n_n, read_len, err, rng), rng) if n_n else np.zeros((0, read_len), np.uint8)
    return MatcherInputs(text, lq, nn, read_len, name)


def adversarial(seed: int, read_len: int = 100, n_reads: int = 3000, text_len: int = 40000,
                with_n: bool = True) -> MatcherInputs:
    """Small adversarial case modelled on the survey's validation harness:
    diverged repeats, reverse-complement copies, near-reverse-palindromes (which trigger the
    coordinate-only skip of ReadsMatchers.cpp:313), low-complexity tracts, a tail shorter than
    a read, reads with 0..L/3+ substitutions, and N reads including N pairs 32 apart (they
    cancel in the 32-bit rotate-xor hash when the seed is longer than 32)."""
    rng = np.random.default_rng(seed)
    L = read_len
    parts = []
    base = random_genome(text_len // 2, rng)
    parts.append(base)
    # diverged repeats of random windows
    for _ in range(12):
        s = int(rng.integers(0, base.size - 3 * L))
        seg = base[s:s + int(rng.integers(L, 3 * L))].copy()
        k = int(rng.integers(0, 6))
        idx = rng.integers(0, seg.size, size=k)
        seg[idx] = _CODE2ASCII[rng.integers(0, 4, size=k)]
        parts.append(seg)
    # reverse-complement copies
    for _ in range(8):
        s = int(rng.integers(0, base.size - 3 * L))
        parts.append(revcomp(base[s:s + int(rng.integers(L, 3 * L))]))
    # near-reverse-palindromes: X + few-diffs + revcomp(X)
    pals = []
    for _ in range(24):
        half = random_genome(int(rng.integers(L // 2 + 5, L + 20)), rng)
        other = revcomp(half).copy()
        k = int(rng.integers(0, 5))
        idx = rng.integers(0, other.size, size=k)
        other[idx] = _CODE2ASCII[rng.integers(0, 4, size=k)]
        pals.append(np.concatenate([half, other]))
    parts.extend(pals)
    # low-complexity tracts
    for unit in ("A", "AC", "AT", "ACG", "T"):
        u = np.frombuffer(unit.encode(), np.uint8)
        parts.append(np.tile(u, (2 * L + 30) // u.size + 1)[:2 * L + 30])
    parts.append(random_genome(text_len // 4, rng))
    parts.append(random_genome(L - 7, rng))            # tail shorter than a read
    order = rng.permutation(len(parts) - 1)
    text = np.ascontiguousarray(np.concatenate([parts[i] for i in order] + [parts[-1]]))

    # reads: windows of the text (either strand) with 0..many substitutions
    n_n = n_reads // 5 if with_n else 0
    n_lq = n_reads - n_n

    def draw(m):
        start = rng.integers(0, text.size - L + 1, size=m)
        r = text[start[:, None] + np.arange(L)[None, :]].copy()
        # every 4th read sits on a near-palindrome, half of those exactly centred on it: the
        # forward and the RC alignment then report the SAME coordinate (ReadsMatchers.cpp:313)
        for i in range(0, m, 4):
            pal = pals[int(rng.integers(0, len(pals)))]
            c = pal.size // 2
            s0 = c - L // 2 if (i // 4) % 2 == 0 else int(rng.integers(max(0, c - L + 8), min(c - 8, pal.size - L) + 1))
            r[i] = pal[s0:s0 + L]
        flip = rng.random(m) < 0.5
        r[flip] = revcomp(r[flip])
        nsub = rng.choice([0, 1, 2, 3, 5, 8, 12, 20, L // 3, L // 3 + 1, L // 2], size=m)
        for i in range(m):
            if nsub[i]:
                idx = rng.choice(L, size=int(nsub[i]), replace=False)
                r[i, idx] = _CODE2ASCII[(_ASCII2CODE4[r[i, idx]] + rng.integers(1, 4, size=idx.size)) & 3]
        return r

    lq = draw(n_lq)
    nn = draw(n_n)
    for i in range(n_n):
        mode = i % 3
        if mode == 0:       # N pair 32 apart inside seed 0: cancels in the hash for seeds > 32
            k = int(rng.integers(0, 6))
            nn[i, [k, k + 32]] = _N
        elif mode == 1:     # single N
            nn[i, int(rng.integers(0, L))] = _N
        else:               # 2-3 random N
            nn[i, rng.choice(L, size=int(rng.integers(2, 4)), replace=False)] = _N
    return MatcherInputs(text, lq, nn, L, f"adversarial(seed={seed},L={L})")


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs at matcher level: genome length, LQ reads handed to the
# matcher, read length, substitution rate, pseudogenome length / genome length.
# C1 and C2 sizes are the reference's own (probed) matcher inputs; C3-C5 are the survey's estimates.
CONFIGS = {
    "c1": dict(genome_len=5_000_000, n_reads=384_497, read_len=100, err=0.001, copies=2.0687),
    "c2": dict(genome_len=50_000_000, n_reads=10_447_991, read_len=150, err=0.005, copies=2.8126),
    "c3": dict(genome_len=100_000_000, n_reads=32_000_000, read_len=150, err=0.005, copies=2.5),
    "c4": dict(genome_len=1_000_000_000, n_reads=126_000_000, read_len=100, err=0.01, copies=2.5),
    "c5": dict(genome_len=3_000_000_000, n_reads=330_000_000, read_len=150, err=0.005, copies=2.5),
}


# ---------------------------------------------------------------------------------------------
# Stage 7 (exact matches between pseudogenomes): a source text and a destination text that
# shares stretches with it.
def pg_texts(seed: int, n: int, n2: int, max_copy: int = 3000, n_frac: float = 0.0, self_rc: int = 0,
             adversarial: bool = True):
    """(src, dest) ASCII uint8.  src: random ACGT with internal repeats (buckets of the text index with several
    entries), tandem repeats and homopolymers (buckets cut at 13 entries), `self_rc` reverse-complement repeats (what the
    HQ-vs-its-own-reverse-complement call finds); dest: random ACGT with copies of stretches of src of 20..max_copy
    characters, forward and reverse-complemented, 0-2 substitutions each, copies that touch both ends of the texts, and
    `n_frac` N symbols."""
    rng = np.random.default_rng(seed)
    src = random_genome(n, rng)
    if adversarial:
        for _ in range(n // 400):
            l = int(rng.integers(40, 600)); a = int(rng.integers(0, n - l)); b = int(rng.integers(0, n - l))
            src[b:b + l] = src[a:a + l].copy()
        for _ in range(n // 3000 + 1):
            l = int(rng.integers(60, 400)); b = int(rng.integers(0, n - l))
            src[b:b + l] = np.resize(random_genome(int(rng.integers(1, 7)), rng), l)
    for _ in range(self_rc):
        l = int(rng.integers(30, 400)); a = int(rng.integers(0, n - l)); b = int(rng.integers(0, n - l))
        src[b:b + l] = revcomp(src[a:a + l].copy())
    dest = random_genome(n2, rng)
    for _ in range(n2 // 250 + 2):
        l = int(rng.integers(20, max(21, min(max_copy, n2)))); a = int(rng.integers(0, max(1, n - l))); l = min(l, n - a)
        b = int(rng.integers(0, n2))
        seg = src[a:a + l].copy()
        if rng.random() < 0.5:
            seg = revcomp(seg)
        for _ in range(int(rng.integers(0, 3))):
            seg[int(rng.integers(0, l))] = _CODE2ASCII[int(rng.integers(0, 4))]
        m = min(l, n2 - b)
        dest[b:b + m] = seg[:m]
    l = min(200, n2, n)
    dest[:l] = src[:l]
    dest[n2 - l:] = src[n - l:]
    if n_frac > 0:
        dest[rng.random(n2) < n_frac] = _N
    return src, dest


def scaled_config(name: str, scale: float = 1.0) -> dict:
    c = dict(CONFIGS[name])
    c["genome_len"] = max(2000, int(c["genome_len"] * scale))
    c["n_reads"] = max(1, int(c["n_reads"] * scale))
    return c


# ---------------------------------------------------------------------------------------------
# Counter-based generator (pgrc_b200/csrc/pgs_synth.cu -> libpgrc_synth.so): every byte is a pure function of
# (seed, index), bit-identical on the host (OpenMP) and on the device (kernel).  This is what the bench and the
# full-size parity tests use: fixtures under tests/golden/ hold what the reference / the oracle computed in the
# build container for inputs that the GPU box re-creates from the same parameters.
import ctypes as _ct
import os as _os

SYNTH_LIB_PATH = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "libpgrc_synth.so")
_synth_lib = None


class PgsParams(_ct.Structure):
    _fields_ = [("seed", _ct.c_uint64), ("genome_len", _ct.c_uint64), ("text_len", _ct.c_uint64), ("contig", _ct.c_uint32),
                ("read_len", _ct.c_uint32), ("err_q24", _ct.c_uint32), ("reserved", _ct.c_uint32)]


def _synth():
    global _synth_lib
    if _synth_lib is None:
        if not _os.path.exists(SYNTH_LIB_PATH):
            raise ImportError(f"{SYNTH_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = _ct.CDLL(SYNTH_LIB_PATH)
        for fn in (lib.pgs_text, lib.pgs_reads):
            fn.restype = _ct.c_int
            fn.argtypes = [_ct.POINTER(PgsParams), _ct.c_void_p, _ct.c_uint64, _ct.c_uint64]
        _synth_lib = lib
    return _synth_lib


def hashed_params(genome_len: int, n_reads: int, read_len: int, err: float, copies: float, seed: int, contig: int = 4000) -> PgsParams:
    return PgsParams(seed, genome_len, int(genome_len * copies), contig, read_len, int(round(err * (1 << 24))), 0)


def hashed_text(p: PgsParams, begin: int = 0, count: int | None = None, device="cpu"):
    """ASCII text [begin, begin + count) as a torch uint8 tensor on `device`."""
    import torch
    count = p.text_len - begin if count is None else count
    out = torch.empty(count, dtype=torch.uint8, device=device)
    if out.is_cuda:
        torch.cuda.current_stream(out.device).synchronize()
    with (torch.cuda.device(out.device) if out.is_cuda else _NullCtx()):
        rc = _synth().pgs_text(_ct.byref(p), out.data_ptr(), begin, count)
        if out.is_cuda:
            torch.cuda.synchronize(out.device)      # the generator launches on the legacy default stream
    if rc:
        raise RuntimeError(f"pgs_text failed ({rc})")
    return out


def hashed_reads(p: PgsParams, first: int, count: int, device="cpu"):
    """Packed reads [first, first + count) (count x ceil(L/4) uint8) on `device`."""
    import torch
    out = torch.empty((count, (p.read_len + 3) // 4), dtype=torch.uint8, device=device)
    if out.is_cuda:
        torch.cuda.current_stream(out.device).synchronize()
    with (torch.cuda.device(out.device) if out.is_cuda else _NullCtx()):
        rc = _synth().pgs_reads(_ct.byref(p), out.data_ptr(), first, count)
        if out.is_cuda:
            torch.cuda.synchronize(out.device)
    if rc:
        raise RuntimeError(f"pgs_reads failed ({rc})")
    return out


def hashed_reads_at(p: PgsParams, indices) -> np.ndarray:
    """Packed reads with the given global indices (host; for the sampled-oracle fixtures)."""
    out = np.empty((len(indices), (p.read_len + 3) // 4), np.uint8)
    lib = _synth()
    for k, r in enumerate(indices):
        rc = lib.pgs_reads(_ct.byref(p), out[k].ctypes.data, int(r), 1)
        if rc:
            raise RuntimeError(f"pgs_reads failed ({rc})")
    return out


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def workload_hashed(name: str, scale: float = 1.0, seed: int = 20261017, device="cpu"):
    """(params, text, packed reads) of a BASELINE config shape from the counter-based generator."""
    c = scaled_config(name, scale)
    p = hashed_params(**c, seed=seed)
    return p, hashed_text(p, 0, None, device), hashed_reads(p, 0, c["n_reads"], device)


def unpack_reads_ascii(packed: np.ndarray, read_len: int) -> np.ndarray:
    """Inverse of pack_reads for the ACGT set (n x L ASCII)."""
    p = np.ascontiguousarray(packed, np.uint8)
    codes = np.stack([(p >> 6) & 3, (p >> 4) & 3, (p >> 2) & 3, p & 3], axis=2).reshape(p.shape[0], -1)[:, :read_len]
    return _CODE2ASCII[codes]


def check_matches_device(text, lq_packed, read_len: int, pos, rc, mm, chunk: int = 1 << 20) -> dict:
    """Size-independent check of the archive-visible outputs, with plain torch ops on the tensors' device (test
    infrastructure, used by the full-size GPU tests and `bench.py --verify`): every matched read, reverse-complemented
    when rc is set, must lie inside the text at `pos` with exactly `mm` mismatches.  Returns counts; `bad` must be 0."""
    import torch
    dev = text.device
    L, n = read_len, lq_packed.shape[0]
    pg_len = text.numel()
    code = torch.full((256,), 255, dtype=torch.uint8, device=dev)
    code[torch.tensor([ord(c) for c in "ACGT"], device=dev)] = torch.arange(4, dtype=torch.uint8, device=dev)
    ar = torch.arange(L, device=dev)
    pos = torch.as_tensor(pos, device=dev).view(torch.int64) if not isinstance(pos, torch.Tensor) else pos.view(torch.int64)
    rc = torch.as_tensor(rc, device=dev)
    mm = torch.as_tensor(mm, device=dev)
    bad = matched = 0
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        m_ok = mm[s:e] != 255
        p = torch.where(m_ok, pos[s:e], torch.zeros_like(pos[s:e]))
        in_text = (p >= 0) & (p + L <= pg_len)
        p = torch.where(in_text, p, torch.zeros_like(p))
        pk = lq_packed[s:e]
        r = torch.stack([(pk >> 6) & 3, (pk >> 4) & 3, (pk >> 2) & 3, pk & 3], dim=2).reshape(e - s, -1)[:, :L]
        flip = rc[s:e] != 0
        r = torch.where(flip[:, None], 3 - r.flip(1), r)                   # the read as it lies on the forward text
        t = code[text[p[:, None] + ar[None, :]].long()]
        d = (r != t).sum(dim=1)
        ok = in_text & (d == mm[s:e].long())
        bad += int((m_ok & ~ok).sum().item())
        matched += int(m_ok.sum().item())
        # an unmatched read carries the sentinels
        bad += int(((~m_ok) & ((pos[s:e] != -1) | flip)).sum().item())
    return {"reads": n, "matched": matched, "bad": bad}
