"""Host-side mirror of the reference's matcher interface for the GPU path.

``map_reads_into_pg`` mirrors ``PgTools::mapReadsIntoPg`` (matching/ReadsMatchers.cpp:693-783):
same parameter names and meaning, same selection of the exact / approximate matcher and of the
optional second phase; it returns what the reference's matcher object holds afterwards
(``readMatchPos``, ``readMatchRC``, ``readMismatchesCount``, ``matchedReadsCount``,
``matchedCountPerMismatches``; ReadsMatchers.h:32-35,115-116).  ``GpuReadsMatcher`` wraps one
``pgm_ctx`` of the C ABI (include/pgrc_gpu_matcher.h) — one GPU, one stream — and exposes the
per-pass steps so that several ranks (one process per GPU) can merge their per-read
accumulators with ``torch.distributed`` between the scan and the decision
(``map_reads_into_pg_sharded``).

Everything here is plumbing; the work happens in the CUDA kernels behind the C ABI.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import (KERNEL_NAMES, PGM_DISABLED_PREFIX_MODE, PGM_ROUTE_CANDIDATES, PGM_ROUTE_PATTERNS, PGM_ROUTE_WINDOWS, PgmAccumulators,
                   PgmError, PgmRouteBuffer, PgmRoutePeer, PgmStats, PgmTimings)

NOT_MATCHED_POSITION = np.uint64(0xFFFFFFFFFFFFFFFF)   # DefaultReadsMatcher::NOT_MATCHED_POSITION
NOT_MATCHED_COUNT = 255                                 # PgTools::NOT_MATCHED_COUNT
DISABLED_PREFIX_MODE = PGM_DISABLED_PREFIX_MODE         # DefaultReadsMatcher::DISABLED_PREFIX_MODE


@dataclass
class MatchResult:
    pos: "np.ndarray"       # readMatchPos   (uint64)
    rc: "np.ndarray"        # readMatchRC    (uint8 0/1)
    mm: "np.ndarray"        # readMismatchesCount (uint8, 255 = unmatched)
    matched: int = 0        # matchedReadsCount
    per_mm: "np.ndarray" = field(default_factory=lambda: np.zeros(256, np.uint64))  # matchedCountPerMismatches
    stats: dict = field(default_factory=dict)

    def matched_reads_bitmap(self, max_mismatches: int = NOT_MATCHED_COUNT - 1) -> np.ndarray:
        """AbstractReadsApproxMatcher::getMatchedReadsBitmap (ReadsMatchers.cpp:685-691)."""
        return np.asarray(self.mm) <= max_mismatches


def _ptr(a):
    """Address of a numpy array or a torch tensor (host or CUDA); keeps no reference."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr()
    raise TypeError(f"unsupported buffer type {type(a)}")


def _rows(a, row_bytes: int) -> int:
    if a is None:
        return 0
    n = a.numel() if hasattr(a, "numel") else a.size
    if n % row_bytes:
        raise ValueError("packed reads buffer is not a multiple of the packed read length")
    return n // row_bytes


class _DevArray:
    """__cuda_array_interface__ view of a raw device pointer (for torch.as_tensor)."""

    def __init__(self, ptr: int, n: int, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}
        self._owner = owner


class GpuReadsMatcher:
    """One matcher context on one GPU (wraps ``pgm_ctx``)."""

    def __init__(self, device: int = 0, use_torch_stream: bool = False):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self._lib.pgm_create(device, ctypes.byref(h))
        if rc != 0:
            raise PgmError(rc, self._lib.pgm_last_error(None).decode())
        self._h = h
        self.device = device
        self._keep = []
        self.n_reads = 0
        self.read_len = 0
        if use_torch_stream:
            import torch
            # torch's default stream is the legacy default stream (handle 0); the C ABI reads NULL as "own stream",
            # so name it explicitly (cudaStreamLegacy = 0x1): kernels then order with torch ops and NCCL collectives
            handle = torch.cuda.current_stream(device).cuda_stream or 1
            self._check(self._lib.pgm_set_stream(self._h, ctypes.c_void_p(handle)))

    # -- life cycle
    def close(self):
        if getattr(self, "_h", None):
            self._lib.pgm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise PgmError(rc, self._lib.pgm_last_error(self._h).decode())

    # -- inputs
    def set_tuning(self, filter_log2_bits: int = -1, slots_per_pattern: int = 3, ctas_per_sm: int = 4, l2_hints: int = 1):
        self._check(self._lib.pgm_set_tuning(self._h, filter_log2_bits, slots_per_pattern, ctas_per_sm, int(l2_hints)))

    def set_text(self, text):
        """text: ASCII pseudogenome (numpy uint8 / torch uint8, host or device)."""
        n = text.numel() if hasattr(text, "numel") else text.size
        self._keep = [k for k in self._keep if k[0] != "text"] + [("text", text)]
        self._check(self._lib.pgm_set_text(self._h, _ptr(text), n))
        self.pg_len = n

    def set_text_shard(self, text_slice, slice_begin: int, pg_len: int, own_begin: int, own_end: int):
        n = text_slice.numel() if hasattr(text_slice, "numel") else text_slice.size
        self._keep = [k for k in self._keep if k[0] != "text"] + [("text", text_slice)]
        self._check(self._lib.pgm_set_text_shard(self._h, _ptr(text_slice), slice_begin, n, pg_len, own_begin, own_end))
        self.pg_len = pg_len

    def set_reads(self, lq_packed, n_packed, read_len: int):
        """Packed reads exactly as PackedConstantLengthReadsSet stores them (LQ: 4 bases/byte,
        N set: 3 bases/byte); global read index = LQ first, then N."""
        n_lq = _rows(lq_packed, (read_len + 3) // 4)
        n_n = _rows(n_packed, (read_len + 2) // 3)
        self._keep = [k for k in self._keep if k[0] != "reads"] + [("reads", lq_packed, n_packed)]
        self._check(self._lib.pgm_set_reads(self._h, _ptr(lq_packed) if n_lq else None, n_lq,
                                            _ptr(n_packed) if n_n else None, n_n, read_len))
        self.n_reads = n_lq + n_n
        self.read_len = read_len

    # -- per-pass steps (initMatching / executeMatching split at the cross-GPU merge point)
    def match_begin(self, seed_len: int, parts: int, max_mm: int, min_mm: int, continuation: bool = False,
                    interleaved: bool = False):
        """initMatching / initMatchingContinuation of the default (contiguous seeds) or the interleaved matcher."""
        fn = self._lib.pgm_match_begin_interleaved if interleaved else self._lib.pgm_match_begin
        self._check(fn(self._h, seed_len, parts, max_mm, min_mm, int(continuation)))

    def scan_pass(self, rev_mode: bool):
        self._check(self._lib.pgm_scan_pass(self._h, int(rev_mode)))

    def resolve_pass(self, rev_mode: bool):
        self._check(self._lib.pgm_resolve_pass(self._h, int(rev_mode)))

    def copmem_begin(self, part_len: int, max_mm: int, min_mm: int, continuation: bool = False):
        """CopMEMReadsApproxMatcher::initMatching / initMatchingContinuation (mode 'c')."""
        self._check(self._lib.pgm_copmem_begin(self._h, part_len, max_mm, min_mm, int(continuation)))

    def copmem_pass(self, rev_mode: bool):
        """CopMEMReadsApproxMatcher::executeMatching: index of the (RC) text + the query of every read."""
        self._check(self._lib.pgm_copmem_pass(self._h, int(rev_mode)))

    def accumulators(self):
        """torch views (no copy) of the per-read accumulators of the current pass."""
        import torch
        acc = PgmAccumulators()
        self._check(self._lib.pgm_get_accumulators(self._h, ctypes.byref(acc)))
        n = int(acc.n_reads)
        dev = f"cuda:{self.device}"
        mk = lambda p, cnt, ts: torch.as_tensor(_DevArray(p, cnt, ts, self), device=dev)
        return {"best_key": mk(acc.best_key, n, "<i8"), "first_other_order": mk(acc.first_other_order, n, "<i8"),
                "same_pos_mask": mk(acc.same_pos_mask, n, "<i4"), "same_pos_mm": mk(acc.same_pos_mm, n, "|u1"),
                "touched": mk(acc.touched, 1, "<i4")}

    def put_accumulators(self):
        """Copies the (merged) contiguous keys back into the read records (pgm_put_accumulators)."""
        self._check(self._lib.pgm_put_accumulators(self._h))

    def synchronize(self):
        self._check(self._lib.pgm_synchronize(self._h))

    def upload(self):
        """Completes the lazy upload of host inputs: afterwards the caller's buffers are no longer read (pgm_upload)."""
        self._check(self._lib.pgm_upload(self._h))

    # -- routed multi-GPU scheme (pgm_route_*): one context per GPU, exchanges done by the caller
    def _segments(self, buf: PgmRouteBuffer):
        """torch uint8 views of the per-destination segments of a send buffer (None where a segment is empty)."""
        import torch
        dev = f"cuda:{self.device}"
        out = []
        for d in range(buf.world):
            nb = int(buf.count[d]) * buf.entry_bytes
            out.append(torch.as_tensor(_DevArray(buf.base + d * buf.stride_bytes, nb, "|u1", self), device=dev) if nb else None)
        return out

    def route_config(self, rank: int, world: int, read_begin, round_windows: int = 0):
        arr = (ctypes.c_uint64 * (world + 1))(*[int(x) for x in read_begin])
        self._check(self._lib.pgm_route_config(self._h, rank, world, arr, round_windows))
        self._route_world, self._route_rank = world, rank

    def route_slot(self, slot: int):
        self._check(self._lib.pgm_route_slot(self._h, slot))

    def route_rounds(self) -> int:
        r = ctypes.c_uint32()
        self._check(self._lib.pgm_route_rounds(self._h, ctypes.byref(r)))
        return int(r.value)

    def route_begin(self, seed_len, parts, max_mm, min_mm, continuation=False):
        buf = PgmRouteBuffer()
        self._check(self._lib.pgm_route_begin(self._h, seed_len, parts, max_mm, min_mm, int(continuation), ctypes.byref(buf)))
        return _Emit(self, buf)

    def route_recv(self, kind: int, in_counts, entry_bytes: int):
        """Receive buffer for the given per-sender counts: list of uint8 views, one per sender (None where empty)."""
        import torch
        total = int(sum(in_counts))
        ptr = ctypes.c_void_p()
        self._check(self._lib.pgm_route_recv(self._h, kind, total, ctypes.byref(ptr)))
        dev = f"cuda:{self.device}"
        views, off = [], 0
        for c in in_counts:
            nb = int(c) * entry_bytes
            views.append(torch.as_tensor(_DevArray(ptr.value + off, nb, "|u1", self), device=dev) if nb else None)
            off += nb
        return views

    def route_build(self, n_in: int):
        self._check(self._lib.pgm_route_build(self._h, n_in))

    def route_scan(self, rev_mode: bool, rnd: int):
        buf = PgmRouteBuffer()
        self._check(self._lib.pgm_route_scan(self._h, int(rev_mode), rnd, ctypes.byref(buf)))
        return _Emit(self, buf)

    def route_probe(self, rev_mode: bool, rnd: int, in_counts):
        buf = PgmRouteBuffer()
        arr = (ctypes.c_uint64 * len(in_counts))(*[int(x) for x in in_counts])
        self._check(self._lib.pgm_route_probe(self._h, int(rev_mode), rnd, arr, ctypes.byref(buf)))
        return _Emit(self, buf)

    def route_scan_launch(self, rev_mode: bool, rnd: int):
        self._check(self._lib.pgm_route_scan_launch(self._h, int(rev_mode), rnd))

    def route_probe_launch(self, rev_mode: bool, rnd: int, in_counts):
        arr = (ctypes.c_uint64 * len(in_counts))(*[int(x) for x in in_counts])
        self._check(self._lib.pgm_route_probe_launch(self._h, int(rev_mode), rnd, arr))

    def route_fetch(self, kind: int):
        """Waits for the emit step of `kind` in the current slot (not for work queued behind it) and returns its send buffer."""
        buf = PgmRouteBuffer()
        self._check(self._lib.pgm_route_fetch(self._h, kind, ctypes.byref(buf)))
        return _Emit(self, buf)

    def route_export(self, kind: int, buf: PgmRouteBuffer) -> bytes:
        """What the peers need to pull their segments of this rank's send buffer: the buffer descriptor (counts, stride)
        and where it lives (pgm_route_export), as one blob for an all-gather."""
        peer = PgmRoutePeer()
        self._check(self._lib.pgm_route_export(self._h, kind, ctypes.byref(peer)))
        return bytes(buf) + bytes(peer)

    def route_pull(self, kind: int, blobs) -> list:
        """Starts the copies of this rank's segments out of every peer's send buffer (pgm_route_pull; asynchronous: the
        consuming call waits on the device).  blobs[s] = route_export of rank s.  Returns the entries received per sender."""
        n, nb = len(blobs), ctypes.sizeof(PgmRouteBuffer)
        sends = (PgmRouteBuffer * n)(*[PgmRouteBuffer.from_buffer_copy(b[:nb]) for b in blobs])
        peers = (PgmRoutePeer * n)(*[PgmRoutePeer.from_buffer_copy(b[nb:]) for b in blobs])
        self._check(self._lib.pgm_route_pull(self._h, kind, peers, sends))
        return [int(sends[s].count[self._route_rank]) for s in range(n)]

    def route_verify(self, rev_mode: bool, n_in: int):
        self._check(self._lib.pgm_route_verify(self._h, int(rev_mode), n_in))

    def kernel_launches(self) -> int:
        return int(self._lib.pgm_kernel_launches(self._h))

    def set_profiling(self, on: bool = True):
        self._check(self._lib.pgm_set_profiling(self._h, int(on)))

    def timings(self) -> dict:
        """Per-kernel device time (ms) and launch count since the last call: {name: (ms, launches)}."""
        t = PgmTimings()
        self._check(self._lib.pgm_get_timings(self._h, ctypes.byref(t)))
        return {k: (float(t.ms[i]), int(t.launches[i])) for i, k in enumerate(KERNEL_NAMES)}

    def _alloc_out(self, out):
        if out is not None:
            return out
        n = self.n_reads
        return (np.empty(n, np.uint64), np.empty(n, np.uint8), np.empty(n, np.uint8))

    @staticmethod
    def _result(out, st: PgmStats) -> MatchResult:
        per = np.frombuffer(bytes(st.per_mm), np.uint64).copy()
        stats = {k: int(getattr(st, k)) for k in ("patterns_inserted", "table_slots", "candidates", "verified",
                                                  "accepted", "filter_positives")}
        return MatchResult(out[0], out[1], out[2], int(st.matched), per, stats)

    def get_results(self, out=None) -> MatchResult:
        out = self._alloc_out(out)
        st = PgmStats()
        self._check(self._lib.pgm_get_results(self._h, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), ctypes.byref(st)))
        return self._result(out, st)

    def get_mismatches(self):
        """Mismatch lists of the matched reads (pgm_get_mismatches; what the reference's export recomputes per read,
        ReadsMatchers.cpp:555-566): (offsets uint64[n+1], pos uint8[total], pg_sym uint8[total], read_sym uint8[total]),
        symbols as codes A C G T N = 0..4, the read taken reverse-complemented when readMatchRC."""
        total = ctypes.c_uint64()
        self._check(self._lib.pgm_get_mismatches(self._h, None, None, None, 0, ctypes.byref(total)))
        n = int(total.value)
        off = np.empty(self.n_reads + 1, np.uint64)
        pos, syms = np.empty(max(n, 1), np.uint8), np.empty(max(n, 1), np.uint8)
        self._check(self._lib.pgm_get_mismatches(self._h, _ptr(off), _ptr(pos), _ptr(syms), n, ctypes.byref(total)))
        pos, syms = pos[:n], syms[:n]
        return off, pos, syms & 3, syms >> 2

    # -- the whole stage on this GPU
    def map_reads(self, seed: int = 38, min_chars_per_mismatch: int = 3, mode: str = "d", pre_seed: int = 0,
                  pre_mode: str = "d", rev_compl: bool = True, match_prefix_length: int = DISABLED_PREFIX_MODE,
                  out=None) -> MatchResult:
        out = self._alloc_out(out)
        st = PgmStats()
        self._check(self._lib.pgm_map_reads(self._h, match_prefix_length, pre_seed, seed, min_chars_per_mismatch,
                                            pre_mode.encode()[:1], mode.encode()[:1], int(rev_compl),
                                            _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), ctypes.byref(st)))
        return self._result(out, st)


class GpuTextMatcher:
    """Stage 7: the reference's TextMatcher interface (matching/TextMatchers.h:54-61) as
    CopMEMMatcher implements it for SimplePgMatcher (copmem/CopMEMMatcher.cpp:571-624), on the GPU.

        GpuTextMatcher(src_text, target_match_length[, min_match_length])   <->  CopMEMMatcher(srcText, srcLength, ...)
        .match_texts(dest_text, dest_is_src, rev_compl_matching, min_match_length)  <->  matchTexts(resMatches, ...)

    match_texts returns resMatches as an (n, 3) uint64 array {posSrcText, length, posDestText} in the reference's push
    order.  `dest_text` is what the matcher is handed (SimplePgMatcher::exactMatchPg reverse-complements it beforehand when
    rev_compl_matching); dest_text=None with dest_is_src takes the source (or its reverse complement) already on the GPU."""

    def __init__(self, src_text, target_match_length: int, min_match_length: int = 0xFFFFFFFF, device: int = 0, matcher=None):
        self._own = matcher is None
        self.m = matcher if matcher is not None else GpuReadsMatcher(device)
        if src_text is not None:
            self.m.set_text(src_text)
        par = (ctypes.c_uint32 * 4)()
        self.m._check(self.m._lib.pgm_mem_index(self.m._h, target_match_length, min(min_match_length, 0xFFFFFFFF), par))
        self.K, self.k1, self.k2, self.hash_size = (int(x) for x in par)
        self.target_match_length = target_match_length

    def match_texts(self, dest_text, dest_is_src: bool = False, rev_compl_matching: bool = True,
                    min_match_length: int = 0xFFFFFFFF) -> np.ndarray:
        n2 = 0 if dest_text is None else (dest_text.numel() if hasattr(dest_text, "numel") else dest_text.size)
        cnt = ctypes.c_uint64(0)
        self.m._check(self.m._lib.pgm_mem_match(self.m._h, _ptr(dest_text), n2, int(dest_is_src), int(rev_compl_matching),
                                                min(min_match_length, 0xFFFFFFFF), ctypes.byref(cnt)))
        out = np.empty((int(cnt.value), 3), np.uint64)
        self.m._check(self.m._lib.pgm_mem_get_matches(self.m._h, out.ctypes.data if out.size else None, int(cnt.value)))
        return out

    def match_texts_share(self, dest_text, dest_is_src: bool, rev_compl_matching: bool, part: int, n_parts: int,
                          min_match_length: int = 0xFFFFFFFF):
        """One rank of several (one process per GPU): the part-th share of the groups of 256 query positions — (matches (n, 3),
        query positions (n,)) in push order before the "covered by the previous match" test; merge_text_match_shares joins
        the shares."""
        n2 = 0 if dest_text is None else (dest_text.numel() if hasattr(dest_text, "numel") else dest_text.size)
        cnt = ctypes.c_uint64(0)
        self.m._check(self.m._lib.pgm_mem_match_share(self.m._h, _ptr(dest_text), n2, int(dest_is_src), int(rev_compl_matching),
                                                      min(min_match_length, 0xFFFFFFFF), part, n_parts, ctypes.byref(cnt)))
        n = int(cnt.value)
        out, qpos = np.empty((n, 3), np.uint64), np.empty(n, np.uint64)
        self.m._check(self.m._lib.pgm_mem_get_share(self.m._h, out.ctypes.data if n else None, qpos.ctypes.data if n else None, n))
        return out, qpos

    def close(self):
        if self._own and self.m is not None:
            self.m.close()
        self.m = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def merge_text_match_shares(shares, K: int) -> np.ndarray:
    """resMatches from the shares of all ranks, in rank order: an element is dropped when it is its predecessor's match —
    same diagonal, and its K-mer ends inside the predecessor (CopMEMMatcher.cpp:388-393).  shares: [(matches (n, 3), query
    positions (n,)), ...] as GpuTextMatcher.match_texts_share returns them."""
    m = np.concatenate([np.asarray(s[0], np.uint64).reshape(-1, 3) for s in shares]) if shares else np.zeros((0, 3), np.uint64)
    q = np.concatenate([np.asarray(s[1], np.uint64).reshape(-1) for s in shares]) if shares else np.zeros(0, np.uint64)
    if len(m) < 2:
        return m
    diag = m[:, 2] - m[:, 0]                                    # (unsigned wrap-around on both sides, as in the reference)
    same = (diag[1:] == diag[:-1]) & (q[1:] + np.uint64(K) < m[:-1, 2] + m[:-1, 1])
    return m[np.concatenate([[True], ~same])]


def match_texts_distributed(tm, comm_world: int, rank: int, all_gather_arrays, dest_text, dest_is_src: bool = False,
                            rev_compl_matching: bool = True, min_match_length: int = 0xFFFFFFFF) -> np.ndarray:
    """Stage 7 with one process per GPU: this rank's share of the query groups, an all-gather of the (small) shares, the merge.
    tm: a text matcher with match_texts_share and K (GpuTextMatcher); all_gather_arrays(list of numpy arrays) -> per rank
    lists (TorchComm.all_gather_arrays).  Every rank returns the whole resMatches vector."""
    mine = tm.match_texts_share(dest_text, dest_is_src, rev_compl_matching, rank, comm_world, min_match_length)
    shares = all_gather_arrays([mine[0].reshape(-1), mine[1]])
    return merge_text_match_shares([(s[0].reshape(-1, 3), s[1]) for s in shares], tm.K)


class GpuMatcherGroup:
    """Several GPUs behind one handle, inside one process (wraps ``pgm_group``; what the C++ host side of PgRC uses:
    pgrc_b200/host/GpuReadsMatchers.cpp).  Modes 'd'/'D' run the routed scheme with peer-to-peer copies over NVLink,
    'i' and 'c' give every GPU the whole text and a read range.  A device may be listed more than once."""

    def __init__(self, devices):
        self._lib = _lib.load()
        devs = [int(d) for d in devices]
        arr = (ctypes.c_int * len(devs))(*devs)
        h = ctypes.c_void_p()
        rc = self._lib.pgm_group_create(len(devs), arr, ctypes.byref(h))
        if rc != 0:
            raise PgmError(rc, self._lib.pgm_group_last_error(None).decode())
        self._h, self.devices, self._keep, self.n_reads = h, devs, [], 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pgm_group_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise PgmError(rc, self._lib.pgm_group_last_error(self._h).decode())

    def set_text(self, text):
        n = text.numel() if hasattr(text, "numel") else text.size
        self._keep = [k for k in self._keep if k[0] != "text"] + [("text", text)]
        self._check(self._lib.pgm_group_set_text(self._h, _ptr(text), n))

    def set_reads(self, lq_packed, n_packed, read_len: int):
        n_lq, n_n = _rows(lq_packed, (read_len + 3) // 4), _rows(n_packed, (read_len + 2) // 3)
        self._keep = [k for k in self._keep if k[0] != "reads"] + [("reads", lq_packed, n_packed)]
        self._check(self._lib.pgm_group_set_reads(self._h, _ptr(lq_packed) if n_lq else None, n_lq, _ptr(n_packed) if n_n else None, n_n, read_len))
        self.n_reads, self.read_len = n_lq + n_n, read_len

    # -- stage 7 (pgm_group_mem_*): the text set with set_text is the source; every GPU indexes it, the query groups are shared out
    def mem_index(self, target_match_length: int, min_match_length: int = 0xFFFFFFFF):
        par = (ctypes.c_uint32 * 4)()
        self._check(self._lib.pgm_group_mem_index(self._h, target_match_length, min(min_match_length, 0xFFFFFFFF), par))
        return tuple(int(x) for x in par)

    def match_texts(self, dest_text, dest_is_src: bool = False, rev_compl_matching: bool = True,
                    min_match_length: int = 0xFFFFFFFF) -> np.ndarray:
        n2 = 0 if dest_text is None else (dest_text.numel() if hasattr(dest_text, "numel") else dest_text.size)
        cnt = ctypes.c_uint64(0)
        self._check(self._lib.pgm_group_mem_match(self._h, _ptr(dest_text), n2, int(dest_is_src), int(rev_compl_matching),
                                                  min(min_match_length, 0xFFFFFFFF), ctypes.byref(cnt)))
        out = np.empty((int(cnt.value), 3), np.uint64)
        self._check(self._lib.pgm_group_mem_get_matches(self._h, out.ctypes.data if out.size else None, int(cnt.value)))
        return out

    def run_plan(self, plan: "MatchPlan", rev_compl_pg: bool = True):
        """The phases of mapReadsIntoPg through the step-wise group calls (as the C++ matcher classes issue them)."""
        for seed_len, parts, max_mm, min_mm, cont, ilv in plan.phases:
            if ilv == "c":
                self._check(self._lib.pgm_group_copmem_begin(self._h, seed_len, max_mm, min_mm, int(cont)))
                for rev in ((0, 1) if rev_compl_pg else (0,)):
                    self._check(self._lib.pgm_group_copmem_pass(self._h, rev))
            else:
                self._check(self._lib.pgm_group_match_begin(self._h, seed_len, parts, max_mm, min_mm, int(cont), int(bool(ilv))))
                for rev in ((0, 1) if rev_compl_pg else (0,)):
                    self._check(self._lib.pgm_group_pass(self._h, rev))

    def get_results(self, out=None) -> MatchResult:
        n = self.n_reads
        out = out if out is not None else (np.empty(n, np.uint64), np.empty(n, np.uint8), np.empty(n, np.uint8))
        st = PgmStats()
        self._check(self._lib.pgm_group_get_results(self._h, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), ctypes.byref(st)))
        return GpuReadsMatcher._result(out, st)

    def get_mismatches(self):
        total = ctypes.c_uint64()
        self._check(self._lib.pgm_group_get_mismatches(self._h, None, None, None, 0, ctypes.byref(total)))
        n = int(total.value)
        off = np.empty(self.n_reads + 1, np.uint64)
        pos, syms = np.empty(max(n, 1), np.uint8), np.empty(max(n, 1), np.uint8)
        self._check(self._lib.pgm_group_get_mismatches(self._h, _ptr(off), _ptr(pos), _ptr(syms), n, ctypes.byref(total)))
        pos, syms = pos[:n], syms[:n]
        return off, pos, syms & 3, syms >> 2


@dataclass
class MatchPlan:
    """Parameter derivation of mapReadsIntoPg (ReadsMatchers.cpp:699-713,749-756): a list of
    matcher phases, each (seed_len, parts, max_mm, min_mm, continuation, interleaved)."""
    phases: list

    @staticmethod
    def derive(read_len: int, seed: int, min_chars_per_mismatch: int, mode: str, pre_seed: int = 0,
               pre_mode: str = "d") -> "MatchPlan":
        if len(mode) != 1 or len(pre_mode) != 1 or mode.lower() not in "dic" or (pre_seed and pre_mode.lower() not in "dic"):
            raise PgmError(-6, f"unknown matching mode '{mode}' ('d'/'D', 'i'/'I', 'c'/'C')")
        if seed <= 0 or min_chars_per_mismatch <= 0:
            raise PgmError(-1, "seed and min_chars_per_mismatch must be > 0")
        L = read_len
        max_mm = L // min_chars_per_mismatch
        reads_exact, pre_exact = min(seed, L), min(pre_seed, L)
        cur_exact, cur_mode = (pre_exact, pre_mode) if pre_exact > 0 else (reads_exact, mode)
        cur_min = max_mm if cur_mode.isupper() else 0
        target_mm = L // cur_exact - 1
        ilv = lambda c, parts: c.lower() == "i" and parts > 1   # InterleavedReadsApproxMatcher (:728-731, :760-763)
        # mode 'c' (CopMEMReadsApproxMatcher, also when readLength == seed: :717-720) has no seed table: its phases carry
        # the marker "c" in place of the interleaved flag and run through pgm_copmem_begin / pgm_copmem_pass
        if cur_mode.lower() == "c":
            phases = [(cur_exact, target_mm + 1, max_mm, cur_min, False, "c")]
        else:
            phases = ([(L, 1, 0, 0, False, False)] if L == cur_exact
                      else [(cur_exact, target_mm + 1, max_mm, cur_min, False, ilv(cur_mode, target_mm + 1))])
        if pre_exact > 0:
            min2 = max_mm if mode.isupper() else target_mm + 1
            phases.append((reads_exact, L // reads_exact, max_mm, min2, True, "c" if mode.lower() == "c" else ilv(mode, L // reads_exact)))
        return MatchPlan(phases)


def map_reads_into_pg(text, lq_packed, n_packed, read_len: int, *, rev_compl_pg: bool = True,
                      match_prefix_length: int = DISABLED_PREFIX_MODE, pre_reads_exact_matching_chars: int = 0,
                      reads_exact_matching_chars: int = 38, min_chars_per_mismatch: int = 3,
                      pre_matching_mode: str = "d", matching_mode: str = "d", device: int = 0,
                      matcher: GpuReadsMatcher | None = None) -> MatchResult:
    """Single-GPU mirror of ``PgTools::mapReadsIntoPg`` (matching part only; the archive export,
    ReadsMatchers.cpp:785-792, stays on the host side of PgRC)."""
    own = matcher is None
    m = matcher or GpuReadsMatcher(device)
    try:
        m.set_text(text)
        m.set_reads(lq_packed, n_packed, read_len)
        return m.map_reads(reads_exact_matching_chars, min_chars_per_mismatch, matching_mode,
                           pre_reads_exact_matching_chars, pre_matching_mode, rev_compl_pg, match_prefix_length)
    finally:
        if own:
            m.close()


def shard_plan(pg_len: int, rank: int, world: int):
    """(slice_begin, slice_len, own_begin, own_end) of one rank: contiguous ranges of seed-window
    starts with PGM_SHARD_HALO bases of text on either side."""
    lib = _lib.load()
    v = [ctypes.c_uint64() for _ in range(4)]
    rc = lib.pgm_shard_plan(pg_len, rank, world, *[ctypes.byref(x) for x in v])
    if rc != 0:
        raise PgmError(rc, "pgm_shard_plan")
    return tuple(int(x.value) for x in v)


def text_share(pg_len: int, rank: int, world: int):
    """(begin, end, per): the part of the pseudogenome rank `rank` uploads in ``all_gather_text``; `per` is the
    16-byte-aligned share size every rank contributes (the last one is padded)."""
    per = ((pg_len + world - 1) // world + 15) // 16 * 16
    return min(pg_len, rank * per), min(pg_len, (rank + 1) * per), per


def all_gather_text(share_host, pg_len: int, rank: int, world: int, device, bufs: dict | None = None, group=None):
    """Read-sharded multi-GPU runs need the whole pseudogenome on every GPU.  Instead of `world` copies of it
    crossing the host's PCIe links, every rank uploads its ``text_share`` (host tensor, pinned for speed) and one
    NCCL all-gather over NVLink / NVSwitch replicates it (replication from the GPU that received the
    H2D copy).  Returns the ASCII text as a device tensor of pg_len bytes; `bufs` caches the device buffers between calls."""
    import torch
    import torch.distributed as dist
    b, e, per = text_share(pg_len, rank, world)
    bufs = bufs if bufs is not None else {}
    if bufs.get("per") != per or bufs.get("world") != world:
        bufs.update(per=per, world=world, full=torch.empty(per * world, dtype=torch.uint8, device=device))
    full = bufs["full"]
    mine = full[rank * per:(rank + 1) * per]
    mine[:e - b].copy_(share_host[:e - b], non_blocking=True)
    if e - b < per:
        mine[e - b:].fill_(ord("A"))
    if world > 1:
        dist.all_gather_into_tensor(full, mine, group=group)
    return full[:pg_len]


class _Emit:
    """Result of an emit step (route_begin / route_scan / route_probe): the send buffer's descriptor."""

    def __init__(self, m, buf):
        self.m, self.buf, self.eb = m, buf, buf.entry_bytes
        self.counts = [int(buf.count[d]) for d in range(buf.world)]

    @property
    def segs(self):
        return self.m._segments(self.buf)


class _Arrived:
    def wait(self):
        pass


class TorchComm:
    """The collectives the sharded schemes need, over torch.distributed (NCCL across GPUs: NVLink / NVSwitch; gloo in the
    CPU tests).  Tests substitute an object with the same methods that connects several contexts inside one process."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._peers = [dist.get_global_rank(group, r) for r in range(self.world)] if group is not None else list(range(self.world))

    def all_reduce(self, t, op: str):
        import torch.distributed as dist
        dist.all_reduce(t, op={"min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM}[op], group=self.group)

    def exchange_counts(self, counts, device):
        """counts[d] = entries this rank sends to d  ->  list of entries this rank receives from every sender."""
        import torch
        import torch.distributed as dist
        mine = torch.tensor(counts, dtype=torch.int64, device=device)
        allc = torch.empty(self.world * self.world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(allc, mine, group=self.group)
        return allc.view(self.world, self.world)[:, self.rank].tolist()

    def all_gather_bytes(self, blob: bytes, device):
        """Every rank's blob (equal sizes), in rank order."""
        import torch
        import torch.distributed as dist
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
        allb = torch.empty(self.world * len(blob), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allb, mine, group=self.group)
        raw = allb.cpu().numpy().tobytes()
        return [raw[r * len(blob):(r + 1) * len(blob)] for r in range(self.world)]

    def all_gather_arrays(self, arrays, device="cpu"):
        """Every rank's list of 1-D uint64 numpy arrays (sizes differ between ranks), in rank order: [[a0, a1, ...] of rank 0, ...]."""
        import torch
        import torch.distributed as dist
        sizes = torch.tensor([len(a) for a in arrays], dtype=torch.int64, device=device)
        all_sizes = [torch.zeros_like(sizes) for _ in range(self.world)]
        dist.all_gather(all_sizes, sizes, group=self.group)
        out = [[] for _ in range(self.world)]
        for k, a in enumerate(arrays):
            mx = max(int(sz[k].item()) for sz in all_sizes)
            mine = torch.zeros(max(mx, 1), dtype=torch.int64, device=device)
            if len(a):
                mine[:len(a)] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(device)
            bufs = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(bufs, mine, group=self.group)
            for r in range(self.world):
                out[r].append(bufs[r][:int(all_sizes[r][k].item())].cpu().numpy().view(np.uint64).copy())
        return out

    def all_to_all_async(self, recv, send):
        """send[d] -> rank d, recv[s] <- rank s (uint8 tensors or None for empty segments): NCCL send/recv pairs on the
        communicator's own stream; returns an object whose wait() orders the current stream behind the transfer."""
        import torch.distributed as dist
        ops = []
        for k in range(1, self.world):
            d, s_ = (self.rank + k) % self.world, (self.rank - k) % self.world
            if send[d] is not None:
                ops.append(dist.P2POp(dist.isend, send[d], self._peers[d], group=self.group))
            if recv[s_] is not None:
                ops.append(dist.P2POp(dist.irecv, recv[s_], self._peers[s_], group=self.group))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        if recv[self.rank] is not None:
            recv[self.rank].copy_(send[self.rank])
        return _Pending(reqs, (recv, send))

    def all_to_all(self, recv, send):
        self.all_to_all_async(recv, send).wait()

    def sibling(self):
        """A second communicator over the same ranks (its own NCCL stream): transfers on it do not queue behind this one's.
        Collective: every rank of the group calls it."""
        import torch.distributed as dist
        return TorchComm(dist.new_group(self._peers))


class _Pending:
    def __init__(self, reqs, keep):
        self._reqs, self._keep = reqs, keep

    def wait(self):
        for r in self._reqs:
            r.wait()
        self._keep = None


def _as_comm(comm_or_group):
    return comm_or_group if hasattr(comm_or_group, "all_reduce") else TorchComm(comm_or_group)


def merge_accumulators(m, comm=None):
    """The one exchange step of the text-sharded path: per-read MIN / SUM all-reduce of the pass accumulators over NCCL
    (NVLink / NVSwitch).  `touched` is MAX-reduced IN PLACE first — resolve_kernel gates on it, so every rank must see the
    merged flag — and the three rarely used accumulators are merged only when some rank touched them."""
    comm = _as_comm(comm)
    acc = m.accumulators()
    comm.all_reduce(acc["best_key"], "min")
    comm.all_reduce(acc["touched"], "max")
    if int(acc["touched"].item()):
        comm.all_reduce(acc["first_other_order"], "min")
        comm.all_reduce(acc["same_pos_mask"], "sum")
        comm.all_reduce(acc["same_pos_mm"], "min")
    m.put_accumulators()
    return acc


def run_plan_sharded(m, plan: MatchPlan, rev_compl_pg: bool = True, comm=None, merge=merge_accumulators):
    """Runs the matcher phases on a context that holds one text shard; after every scan the
    per-read accumulators are merged across ranks, then every rank applies the decision, so the
    per-read state stays replicated."""
    comm = _as_comm(comm)
    for seed_len, parts, max_mm, min_mm, cont, ilv in plan.phases:
        if ilv == "c":
            raise PgmError(-6, "matching mode 'c' indexes the whole text: it shards by reads only (no text shards)")
        m.match_begin(seed_len, parts, max_mm, min_mm, cont, ilv)
        for rev in ((False, True) if rev_compl_pg else (False,)):
            m.scan_pass(rev)
            merge(m, comm)
            m.resolve_pass(rev)


def read_ranges(n_reads: int, world: int):
    """Even read ranges of the routed / read-sharded schemes: rank g owns [b[g], b[g+1])."""
    return [(n_reads * g) // world for g in range(world + 1)]


def run_plan_routed(m, plan: MatchPlan, rev_compl_pg: bool, comm, n_reads_total: int, round_windows: int = 0, comm2=None,
                    exchange: str = "nccl", deep: bool = False) -> dict:
    """The routed scheme (include/pgrc_gpu_matcher.h, pgm_route_*): this rank's context holds the whole text and its own
    read range (`read_ranges`).  Per phase the seeds are exchanged once (all-to-all by hash owner); per pass and round the
    windows of this rank's text range go to the hash owners and the candidates they find go to the read owners.
    exchange = "nccl": NCCL send/recv pairs of the send segments; "pull": every rank publishes its send buffer (CUDA IPC) and
    pulls its segments out of the peers' buffers with copy engines over NVLink (pgm_route_pull) — no SM time, and the copy
    overlaps whatever kernels follow.  With a second communicator (`comm2`, e.g. comm.sibling(); used for the small
    all-gathers of counts, so that they do not queue behind a transfer) the (pass, round) steps are software-pipelined: the
    windows of step k + 1 are hashed and shipped while step k is probed and verified — the two sets of exchange buffers of the
    context (pgm_route_slot) make that safe.  `deep` = the three-stage variant of the pipeline (scan k+2 / probe k+1 / verify k
    queued together, the host waits per emit step through pgm_route_fetch); measured slower at N = 8 (103.7 vs 94.4 ms per
    step at C5: the exchange, not the host, is what the probe waits for), kept selectable.  Returns the bytes sent per kind."""
    comm = _as_comm(comm)
    dev = f"cuda:{m.device}" if isinstance(m.device, int) else m.device
    m.route_config(comm.rank, comm.world, read_ranges(n_reads_total, comm.world), round_windows)
    sent = {"patterns": 0, "windows": 0, "candidates": 0}
    small = comm2 if comm2 is not None else comm          # counts and candidates
    names = {PGM_ROUTE_PATTERNS: "patterns", PGM_ROUTE_WINDOWS: "windows", PGM_ROUTE_CANDIDATES: "candidates"}

    def ship(kind, em, data_comm):
        """Starts the all-to-all of an emit step; returns (entries received per sender, pending transfer)."""
        sent[names[kind]] += sum(em.counts) * em.eb
        if exchange == "pull":
            return m.route_pull(kind, small.all_gather_bytes(m.route_export(kind, em.buf), dev)), _Arrived()
        in_counts = small.exchange_counts(em.counts, dev)
        recv = m.route_recv(kind, in_counts, em.eb)
        return in_counts, data_comm.all_to_all_async(recv, em.segs)

    def emit_slot(rev, rnd, slot):
        m.route_slot(slot)
        return ship(PGM_ROUTE_WINDOWS, m.route_scan(rev, rnd), comm)

    def consume_slot(rev, rnd, slot, win_in, pending):
        pending.wait()
        m.route_slot(slot)
        cand_in, p2 = ship(PGM_ROUTE_CANDIDATES, m.route_probe(rev, rnd, win_in), small)
        p2.wait()
        m.route_verify(rev, sum(cand_in))

    rounds = 1
    passes = (False, True) if rev_compl_pg else (False,)
    for seed_len, parts, max_mm, min_mm, cont, ilv in plan.phases:
        if ilv:
            raise PgmError(-6, "the routed scheme covers matching modes 'd'/'D' (contiguous seeds); use read ranges for 'i' and 'c'")
        m.route_slot(0)
        pat_in, pat_pending = ship(PGM_ROUTE_PATTERNS, m.route_begin(seed_len, parts, max_mm, min_mm, cont), comm)
        rounds = m.route_rounds()
        seq = [(rev, rnd) for rev in passes for rnd in range(rounds)]
        if comm2 is None:
            pat_pending.wait()
            m.route_build(sum(pat_in))
            for k, (rev, rnd) in enumerate(seq):
                consume_slot(rev, rnd, k & 1, *emit_slot(rev, rnd, k & 1))
                if rnd == rounds - 1:
                    m.resolve_pass(rev)
            continue
        if not deep:
            # pipelined: the (pass, round) pairs form one sequence; the windows of step k + 1 are hashed and shipped while step k is
            # probed and verified — across the pass boundary too (hashing and probing the RC text need no forward result; the
            # verification does: the forward decision is applied before the first RC round is consumed), and the first emit runs
            # while the patterns travel and the table is built (the scan only reads the text)
            cur = emit_slot(*seq[0], 0)
            pat_pending.wait()
            m.route_build(sum(pat_in))
            for k, (rev, rnd) in enumerate(seq):
                nxt = emit_slot(*seq[k + 1], (k + 1) & 1) if k + 1 < len(seq) else None
                consume_slot(rev, rnd, k & 1, *cur)
                if rnd == rounds - 1:
                    m.resolve_pass(rev)
                cur = nxt
            continue
        # pipelined: the (pass, round) pairs form one sequence k = 0, 1, ...; the GPU queue always holds work while the host
        # exchanges counts.  Stream order: ... scan(k+2), probe(k+1), verify(k), scan(k+3), probe(k+2), verify(k+1) ...
        #   * the windows of step k + 2 are hashed and shipped while step k + 1 is probed and step k verified — across the pass
        #     boundary too: hashing and probing the RC text need no forward result; the verification does, and the forward
        #     decision (resolve) is queued behind the last forward verify, ahead of the first RC verify;
        #   * the first two scans run while the patterns travel and the table is built (a scan only reads the text);
        #   * buffer reuse: scan(k+2) overwrites the send slot of windows(k), probe(k+1) the one of candidates(k-1): both follow
        #     the stream synchronize + the all-gather of step (a), after which every rank has probed step k and verified k-1,
        #     i.e. has finished pulling both.
        n = len(seq)
        win = [None] * n
        m.route_slot(0)
        m.route_scan_launch(*seq[0])
        win[0] = ship(PGM_ROUTE_WINDOWS, m.route_fetch(PGM_ROUTE_WINDOWS), comm)
        pat_pending.wait()
        m.route_build(sum(pat_in))
        if n > 1:
            m.route_slot(1)
            m.route_scan_launch(*seq[1])
            win[1] = ship(PGM_ROUTE_WINDOWS, m.route_fetch(PGM_ROUTE_WINDOWS), comm)
        m.route_slot(0)
        win[0][1].wait()
        m.route_probe_launch(*seq[0], win[0][0])
        for k, (rev, rnd) in enumerate(seq):
            m.synchronize()                                               # (a) probe(k) and verify(k-1) are done
            m.route_slot(k & 1)
            cand_in, cand_pending = ship(PGM_ROUTE_CANDIDATES, m.route_fetch(PGM_ROUTE_CANDIDATES), small)
            if k + 2 < n:                                                 # (b)
                m.route_scan_launch(*seq[k + 2])
            if k + 1 < n:                                                 # (c)
                m.route_slot((k + 1) & 1)
                win[k + 1][1].wait()
                m.route_probe_launch(*seq[k + 1], win[k + 1][0])
            m.route_slot(k & 1)                                           # (d)
            cand_pending.wait()
            m.route_verify(rev, sum(cand_in))
            if rnd == rounds - 1:
                m.resolve_pass(rev)
            if k + 2 < n:                                                 # (e) waits for scan(k+2) only
                win[k + 2] = ship(PGM_ROUTE_WINDOWS, m.route_fetch(PGM_ROUTE_WINDOWS), comm)
            win[k] = None
    m.route_slot(0)
    return {"sent_bytes": sent, "rounds_per_pass": rounds, "pipelined": ("deep" if deep else True) if comm2 is not None else False, "exchange": exchange}


def map_reads_into_pg_sharded(text, lq_packed, n_packed, read_len: int, *, rank: int, world: int, device: int,
                              rev_compl_pg: bool = True, pre_reads_exact_matching_chars: int = 0,
                              reads_exact_matching_chars: int = 38, min_chars_per_mismatch: int = 3,
                              pre_matching_mode: str = "d", matching_mode: str = "d", group=None,
                              matcher: GpuReadsMatcher | None = None) -> MatchResult:
    """Multi-GPU mirror of mapReadsIntoPg: the pseudogenome is sharded into contiguous ranges
    (one per rank, halos of PGM_SHARD_HALO bases), the reads and the seed table are replicated,
    and per-read best matches are merged with an NCCL all-reduce after every pass."""
    pg_len = text.numel() if hasattr(text, "numel") else text.size
    sb, sl, ob, oe = shard_plan(pg_len, rank, world)
    own = matcher is None
    m = matcher or GpuReadsMatcher(device, use_torch_stream=True)
    try:
        m.set_text_shard(text[sb:sb + sl], sb, pg_len, ob, oe)
        m.set_reads(lq_packed, n_packed, read_len)
        plan = MatchPlan.derive(read_len, reads_exact_matching_chars, min_chars_per_mismatch, matching_mode,
                                pre_reads_exact_matching_chars, pre_matching_mode)
        run_plan_sharded(m, plan, rev_compl_pg, TorchComm(group))
        return m.get_results()
    finally:
        if own:
            m.close()
