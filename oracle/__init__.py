"""TEST INFRASTRUCTURE — ctypes bindings for the CPU oracle and the reference harness.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this
package; the product (``pgrc_b200``) never does.

* :func:`oracle_map_reads`  -> ``oracle/libpgrc_oracle.so`` (plain-C restatement, pgrc_oracle.c)
* :func:`ref_map_reads`     -> ``oracle/_ref/libpgrc_ref.so`` (the reference's own classes
  behind oracle/ref_harness.cpp; present only if it was built where /root/reference exists)
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libpgrc_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libpgrc_ref.so")

NOT_MATCHED_POSITION = np.uint64(0xFFFFFFFFFFFFFFFF)
NOT_MATCHED_COUNT = 255


def build(ref: bool = True) -> None:
    """Compile the C restatement and, where /root/reference exists, the reference harness."""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/matching"):
        subprocess.run(["make", "-s", "-j8", "-C", _HERE, "ref"], check=True)
        # the reference's CLI with the C++ GPU shim linked in (tests/test_gpu_cli_archive.py); needs the CUDA library
        if os.path.exists(os.path.join(_HERE, "..", "pgrc_b200", "libpgrc_gpu.so")):
            subprocess.run(["make", "-s", "-j8", "-C", _HERE, "cli"], check=True)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class _Stats(ctypes.Structure):
    _fields_ = [("matched", ctypes.c_uint64), ("better", ctypes.c_uint64),
                ("false_matches", ctypes.c_uint64), ("per_mm", ctypes.c_uint64 * 256),
                ("n_patterns", ctypes.c_uint64), ("n_events", ctypes.c_uint64),
                ("n_verified", ctypes.c_uint64), ("n_cross_strand_skips", ctypes.c_uint64)]


@dataclass
class MatchResult:
    pos: np.ndarray          # uint64, NOT_MATCHED_POSITION when unmatched
    rc: np.ndarray           # uint8 0/1
    mm: np.ndarray           # uint8, 255 when unmatched
    matched: int = 0
    better: int = 0
    false_matches: int = 0
    per_mm: np.ndarray = field(default_factory=lambda: np.zeros(256, np.uint64))
    seconds: float = 0.0
    extra: dict = field(default_factory=dict)


_oracle_lib = None
_ref_lib = None


def _oracle():
    global _oracle_lib
    if _oracle_lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        lib = ctypes.CDLL(ORACLE_SO)
        lib.pgo_pack_reads.restype = ctypes.c_int
        lib.pgo_pack_reads.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p]
        lib.pgo_map_reads.restype = ctypes.c_int
        lib.pgo_map_reads.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint32,
                                      ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                      ctypes.c_uint32, ctypes.c_uint32, ctypes.c_char, ctypes.c_char, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.pgo_mismatch_lists.restype = ctypes.c_int
        lib.pgo_mismatch_lists.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                           ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.pgo_set_copmem_staged.restype = None
        lib.pgo_set_copmem_staged.argtypes = [ctypes.c_int]
        lib.pgo_match_texts.restype = ctypes.c_int
        lib.pgo_match_texts.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
        _oracle_lib = lib
    return _oracle_lib


def _ref():
    global _ref_lib
    if _ref_lib is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libpgrc_ref.so not built (needs /root/reference)")
        lib = ctypes.CDLL(REF_SO)
        lib.pgref_map_reads.restype = ctypes.c_int
        lib.pgref_map_reads.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint32,
                                        ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                        ctypes.c_uint32, ctypes.c_uint32, ctypes.c_char, ctypes.c_char,
                                        ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p]
        lib.pgref_pack_reads.restype = ctypes.c_int
        lib.pgref_pack_reads.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p]
        lib.pgref_mismatch_lists.restype = ctypes.c_int
        lib.pgref_mismatch_lists.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint32,
                                             ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.c_uint32, ctypes.c_uint32, ctypes.c_char, ctypes.c_char,
                                             ctypes.c_int, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p]
        lib.pgref_match_texts.restype = ctypes.c_int
        lib.pgref_match_texts.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64,
                                          ctypes.c_void_p, ctypes.c_void_p]
        lib.pgref_mark_matches.restype = ctypes.c_int
        lib.pgref_mark_matches.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                                           ctypes.c_void_p, ctypes.c_void_p]
        _ref_lib = lib
    return _ref_lib


def _ascii2d(reads, read_len: int) -> np.ndarray:
    a = np.ascontiguousarray(reads, dtype=np.uint8)
    if a.size == 0:
        return a.reshape(0, read_len)
    return a.reshape(-1, read_len)


def pack_reads(reads_ascii: np.ndarray, read_len: int, with_n: bool, use_ref: bool = False) -> np.ndarray:
    """ASCII reads (n x L uint8) -> packed reads with the reference's layout."""
    a = _ascii2d(reads_ascii, read_len)
    n = a.shape[0]
    spe = 3 if with_n else 4
    packed_len = (read_len + spe - 1) // spe
    out = np.zeros((n, packed_len), np.uint8)
    if n:
        fn = _ref().pgref_pack_reads if use_ref else _oracle().pgo_pack_reads
        r = fn(a.ctypes.data, n, read_len, int(with_n), out.ctypes.data)
        if r != packed_len:
            raise ValueError(f"pack_reads failed ({r})")
    return out


def set_copmem_staged(on: bool) -> None:
    """Mode 'c' of oracle_map_reads through the STAGED form of the per-read query (copmem_query_staged: the CPU model of the
    warp-per-read CUDA kernels of pgm_copmem_warp.cuh) instead of the sequential transcription."""
    _oracle().pgo_set_copmem_staged(int(on))


def _mode(c: str) -> bytes:
    return c.encode("ascii")[:1]


def oracle_map_reads(text: np.ndarray, lq_packed: np.ndarray, n_packed: np.ndarray | None, read_len: int,
                     seed: int = 38, min_chars_per_mismatch: int = 3, mode: str = "d",
                     pre_seed: int = 0, pre_mode: str = "d", rev_compl: bool = True) -> MatchResult:
    text = np.ascontiguousarray(text, dtype=np.uint8)
    lq = np.ascontiguousarray(lq_packed, dtype=np.uint8)
    n_lq = lq.shape[0] if lq.size else 0
    nn = np.ascontiguousarray(n_packed, dtype=np.uint8) if n_packed is not None and len(n_packed) else np.zeros((0, 1), np.uint8)
    n_n = nn.shape[0] if nn.size else 0
    n = n_lq + n_n
    pos = np.empty(n, np.uint64); rc = np.empty(n, np.uint8); mm = np.empty(n, np.uint8)
    st = _Stats()
    r = _oracle().pgo_map_reads(text.ctypes.data, text.size, lq.ctypes.data, n_lq, nn.ctypes.data, n_n,
                                read_len, pre_seed, seed, min_chars_per_mismatch, _mode(pre_mode), _mode(mode),
                                int(rev_compl), pos.ctypes.data, rc.ctypes.data, mm.ctypes.data, ctypes.byref(st))
    if r != 0:
        raise RuntimeError(f"pgo_map_reads failed ({r})")
    return MatchResult(pos, rc, mm, st.matched, st.better, st.false_matches,
                       np.frombuffer(bytes(st.per_mm), np.uint64).copy(),
                       extra={"n_patterns": st.n_patterns, "n_events": st.n_events, "n_verified": st.n_verified,
                              "n_cross_strand_skips": st.n_cross_strand_skips})


def ref_map_reads(text: np.ndarray, lq_ascii: np.ndarray, n_ascii: np.ndarray | None, read_len: int,
                  seed: int = 38, min_chars_per_mismatch: int = 3, mode: str = "d",
                  pre_seed: int = 0, pre_mode: str = "d", rev_compl: bool = True, threads: int = 0) -> MatchResult:
    """Runs the reference's own matcher classes (ASCII reads in, as addRead takes them)."""
    # one extra readable byte: the reference reads txt[len] once (HashMatcher.h:62)
    buf = np.zeros(np.asarray(text).size + 1, np.uint8)
    buf[:-1] = np.asarray(text, dtype=np.uint8)
    lq = _ascii2d(lq_ascii, read_len)
    nn = _ascii2d(n_ascii, read_len) if n_ascii is not None and len(n_ascii) else np.zeros((0, read_len), np.uint8)
    n = lq.shape[0] + nn.shape[0]
    pos = np.empty(n, np.uint64); rc = np.empty(n, np.uint8); mm = np.empty(n, np.uint8)
    st = np.zeros(259, np.uint64)
    secs = ctypes.c_double(0.0)
    r = _ref().pgref_map_reads(buf.ctypes.data, buf.size - 1, lq.ctypes.data, lq.shape[0], nn.ctypes.data, nn.shape[0],
                               read_len, pre_seed, seed, min_chars_per_mismatch, _mode(pre_mode), _mode(mode),
                               int(rev_compl), threads, pos.ctypes.data, rc.ctypes.data, mm.ctypes.data,
                               st.ctypes.data, ctypes.byref(secs))
    if r != 0:
        raise RuntimeError(f"pgref_map_reads failed ({r})")
    if not np.array_equal(buf[:-1], np.asarray(text, dtype=np.uint8)):
        raise RuntimeError("reference left the text modified")
    return MatchResult(pos, rc, mm, int(st[0]), int(st[1]), int(st[2]), st[3:].copy(), seconds=secs.value)


_SYM_CODE = np.full(256, 255, np.uint8)
_SYM_CODE[[ord(c) for c in "ACGTN"]] = np.arange(5, dtype=np.uint8)


def oracle_mismatch_lists(text, lq_packed, n_packed, read_len: int, pos, rc, mm, variant: int = 0):
    """Mismatch lists of the export step (pgo_mismatch_lists; AbstractReadsApproxMatcher::updateEntry,
    ReadsMatchers.cpp:548-558): (offsets uint64[n+1], off uint8[], pg_sym uint8[], read_sym uint8[]), symbols as codes
    A C G T N = 0..4.  variant 0: forward fill for every read (the pgm_get_mismatches contract); 1 / 2: updateEntry with
    revComplPairFile false / true."""
    text = np.ascontiguousarray(text, dtype=np.uint8)
    lq = np.ascontiguousarray(lq_packed, dtype=np.uint8)
    n_lq = lq.shape[0] if lq.size else 0
    nn = np.ascontiguousarray(n_packed, dtype=np.uint8) if n_packed is not None and len(n_packed) else np.zeros((0, 1), np.uint8)
    n_n = nn.shape[0] if nn.size else 0
    pos = np.ascontiguousarray(pos, np.uint64); rc = np.ascontiguousarray(rc, np.uint8); mm = np.ascontiguousarray(mm, np.uint8)
    total = int(mm[mm != 255].astype(np.int64).sum())
    off = np.empty(n_lq + n_n + 1, np.uint64)
    o, pg, rd = (np.empty(max(total, 1), np.uint8) for _ in range(3))
    r = _oracle().pgo_mismatch_lists(text.ctypes.data, text.size, lq.ctypes.data, n_lq, nn.ctypes.data, n_n, read_len,
                                     pos.ctypes.data, rc.ctypes.data, mm.ctypes.data, variant,
                                     off.ctypes.data, o.ctypes.data, pg.ctypes.data, rd.ctypes.data)
    if r != 0:
        raise RuntimeError(f"pgo_mismatch_lists failed ({r})")
    return off, o[:total], _SYM_CODE[pg[:total]], _SYM_CODE[rd[:total]]


def ref_mismatch_lists(text, lq_ascii, n_ascii, read_len: int, rev_compl_pair_file: bool = False, seed: int = 38,
                       min_chars_per_mismatch: int = 3, mode: str = "d", pre_seed: int = 0, pre_mode: str = "d", rev_compl: bool = True):
    """The reference's own export step (updateEntry through the harness): returns (MatchResult, offsets, off, pg_sym, read_sym)."""
    buf = np.zeros(np.asarray(text).size + 1, np.uint8)
    buf[:-1] = np.asarray(text, dtype=np.uint8)
    lq = _ascii2d(lq_ascii, read_len)
    nn = _ascii2d(n_ascii, read_len) if n_ascii is not None and len(n_ascii) else np.zeros((0, read_len), np.uint8)
    n = lq.shape[0] + nn.shape[0]
    pos = np.empty(n, np.uint64); rc = np.empty(n, np.uint8); mm = np.empty(n, np.uint8)
    off = np.empty(n + 1, np.uint64)
    cap = n * 256
    o, pg, rd = (np.empty(cap, np.uint8) for _ in range(3))
    secs = ctypes.c_double(0.0)
    r = _ref().pgref_mismatch_lists(buf.ctypes.data, buf.size - 1, lq.ctypes.data, lq.shape[0], nn.ctypes.data, nn.shape[0],
                                    read_len, pre_seed, seed, min_chars_per_mismatch, _mode(pre_mode), _mode(mode),
                                    int(rev_compl), int(rev_compl_pair_file), pos.ctypes.data, rc.ctypes.data, mm.ctypes.data,
                                    off.ctypes.data, o.ctypes.data, pg.ctypes.data, rd.ctypes.data, ctypes.byref(secs))
    if r != 0:
        raise RuntimeError(f"pgref_mismatch_lists failed ({r})")
    t = int(off[-1])
    return (MatchResult(pos, rc, mm, int((mm != 255).sum()), 0, 0, np.bincount(mm, minlength=256).astype(np.uint64), seconds=secs.value),
            off, o[:t], _SYM_CODE[pg[:t]], _SYM_CODE[rd[:t]])


# ------------------------------------------------------------------------------------------ stage 7 (exact matches between pseudogenomes)
_COMPLEMENT = np.arange(256, dtype=np.uint8)
for _a, _b in ("AT", "CG", "GC", "TA"):
    _COMPLEMENT[ord(_a)] = ord(_b)


def reverse_complement(text) -> np.ndarray:
    """PgHelpers::reverseComplement (utils/helper.cpp:395-403): symbols outside ACGT stay as they are."""
    return np.ascontiguousarray(_COMPLEMENT[np.asarray(text, dtype=np.uint8)[::-1]])


def _match_texts(fn, src, dest, dest_is_src, rev_compl, target_len, min_len, extra):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    dest = np.ascontiguousarray(dest, dtype=np.uint8)
    cap = max(1024, dest.size // 8)
    while True:
        out = np.empty((cap, 3), np.uint64)
        cnt = ctypes.c_uint64(0)
        r = fn(src.ctypes.data, src.size, dest.ctypes.data, dest.size, int(dest_is_src), int(rev_compl), target_len,
               min(min_len, 0xFFFFFFFF), *extra(out, cap, cnt))
        if r == -3:
            cap = int(cnt.value)
            continue
        if r != 0:
            raise RuntimeError(f"match_texts failed ({r})")
        return out[:int(cnt.value)].copy()


def oracle_match_texts(src, dest, dest_is_src: bool = False, rev_compl: bool = True, target_len: int = 45,
                       min_len: int = 0xFFFFFFFF, params: dict | None = None) -> np.ndarray:
    """pgo_match_texts: resMatches of CopMEMMatcher::matchTexts as an (n, 3) uint64 array {posSrcText, length, posDestText}
    in push order.  `dest` is the text handed to the matcher (already reverse-complemented by the caller when rev_compl)."""
    par = np.zeros(5, np.uint64)
    res = _match_texts(_oracle().pgo_match_texts, src, dest, dest_is_src, rev_compl, target_len, min_len,
                       lambda out, cap, cnt: (out.ctypes.data, cap, ctypes.byref(cnt), par.ctypes.data))
    if params is not None:
        params.update(K=int(par[0]), k1=int(par[1]), k2=int(par[2]), hash_size=int(par[3]), extensions=int(par[4]))
    return res


def ref_match_texts(src, dest, dest_is_src: bool = False, rev_compl: bool = True, target_len: int = 45,
                    min_len: int = 0xFFFFFFFF, threads: int = 1, seconds: list | None = None) -> np.ndarray:
    """The reference's own CopMEMMatcher (constructor + matchTexts) through the harness."""
    secs = (ctypes.c_double * 2)()
    res = _match_texts(_ref().pgref_match_texts, src, dest, dest_is_src, rev_compl, target_len, min_len,
                       lambda out, cap, cnt: (threads, out.ctypes.data, cap, ctypes.byref(cnt), secs))
    if seconds is not None:
        seconds[:] = [secs[0], secs[1]]
    return res


def ref_mark_matches(src, dest, dest_is_src: bool = False, rev_compl: bool = True, target_len: int = 45,
                     min_len: int = 0xFFFFFFFF, threads: int = 1):
    """SimplePgMatcher::markAndRemoveExactMatches through the harness: (mapped destination, resPgMapOff, resPgMapLen) as
    uint8 arrays.  `dest` is NOT reverse-complemented by the caller (the class does it)."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    n2 = src.size if dest_is_src else np.asarray(dest).size
    buf = np.zeros(max(n2, 1), np.uint8)
    if not dest_is_src:
        buf[:n2] = np.asarray(dest, dtype=np.uint8)
    off = np.empty(n2 + 64, np.uint8); ln = np.empty(n2 + 64, np.uint8)
    m, ol, ll = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
    secs = ctypes.c_double(0.0)
    r = _ref().pgref_mark_matches(src.ctypes.data, src.size, buf.ctypes.data, n2, int(dest_is_src), int(rev_compl), target_len,
                                  min(min_len, 0xFFFFFFFF), threads, ctypes.byref(m), off.ctypes.data, off.size, ctypes.byref(ol),
                                  ln.ctypes.data, ln.size, ctypes.byref(ll), ctypes.byref(secs))
    if r != 0:
        raise RuntimeError(f"pgref_mark_matches failed ({r})")
    return buf[:int(m.value)].copy(), off[:int(ol.value)].copy(), ln[:int(ll.value)].copy()
