"""CPU model of one rank's matcher context, for the world_size-2 gloo tests of the sharded host logic
(pgrc_b200/matcher.py: shard_plan, run_plan_sharded, merge_accumulators).  It has the interface of
GpuReadsMatcher and restates, in plain Python over small inputs, what the kernels accumulate per pass
(verify_pattern) and how a pass is decided (resolve_kernel) — see pgrc_b200/csrc/pgm_kernels.cuh.  It is
test infrastructure: the product never imports it."""
import numpy as np
import torch

from pgrc_b200 import matcher as M

KEY_INF = 0x7FFFFFFFFFFFFFFF
_COMP = {65: 84, 67: 71, 71: 67, 84: 65}


def _canon(sym, n):
    """Buzhash-equivalence class of a seed: per rotation class (n-1-k) mod 32, the parity mask of the symbols."""
    k = [0] * 32
    for i, c in enumerate(sym):
        k[(n - 1 - i) % 32] ^= {65: 1, 67: 2, 71: 4, 84: 8, 78: 16}[int(c)]
    return tuple(k)


class CpuShardMatcher:
    def __init__(self):
        self.device = "cpu"

    def set_text_shard(self, text_slice, slice_begin, pg_len, own_begin, own_end):
        self.slice = np.asarray(text_slice, np.uint8)
        self.slice_begin, self.pg_len, self.own = slice_begin, pg_len, (own_begin, own_end)

    def set_reads(self, lq_ascii, n_ascii, read_len):
        """ASCII reads here (the model does not need the packed layout)."""
        self.reads = [np.asarray(r, np.uint8) for r in lq_ascii] + [np.asarray(r, np.uint8) for r in (n_ascii if n_ascii is not None else [])]
        self.n_reads, self.read_len = len(self.reads), read_len
        n = self.n_reads
        self.state = [(255, 0, None)] * n   # (mm, rc, pos)

    def match_begin(self, seed_len, parts, max_mm, min_mm, continuation=False, interleaved=False):
        self.seed_len, self.parts, self.max_mm, self.min_mm = seed_len, parts, max_mm, min_mm
        # mode 'i': seed j = read bases j, j+parts, ...; window at g = text[g], text[g+parts], ...; alignment g - j
        self.stride = parts if interleaved and parts > 1 else 1
        self.shift = 1 if self.stride > 1 else seed_len
        self.span = seed_len * self.stride
        n = self.n_reads
        if not continuation:
            self.state = [(255, 0, None)] * n
        self.table = {}
        for r in range(n):
            if continuation and self.state[r][0] <= min_mm:
                continue
            for j in range(parts):
                seed = self.reads[r][j:j + self.span:self.stride] if self.stride > 1 else self.reads[r][j * seed_len:(j + 1) * seed_len]
                self.table.setdefault(_canon(seed, seed_len), []).append(r * parts + j)
        self._reset_acc()

    def _reset_acc(self):
        n = self.n_reads
        self.acc = {"best_key": torch.full((n,), KEY_INF, dtype=torch.int64), "first_other_order": torch.full((n,), KEY_INF, dtype=torch.int64),
                    "same_pos_mask": torch.zeros(n, dtype=torch.int32), "same_pos_mm": torch.full((n,), 255, dtype=torch.uint8),
                    "touched": torch.zeros(1, dtype=torch.int32)}

    def accumulators(self):
        return self.acc

    def put_accumulators(self):
        """The model keeps its keys in the contiguous arrays themselves (the CUDA path copies them back into the read records)."""

    def scan_pass(self, rev):
        n, L, pg = self.seed_len, self.read_len, self.pg_len
        span, sh = self.span, self.shift
        if pg < span or not self.n_reads:
            return
        fb, fe = self.own[0], min(self.own[1], pg - span + 1)
        if fb >= fe:
            return
        sl = self.slice
        if rev:   # this rank's slice of the reverse-complemented text, and its owned window starts there
            sl = np.array([_COMP[int(c)] for c in sl[::-1]], np.uint8)
            origin = pg - (self.slice_begin + len(self.slice))
            ob, oe = pg - span - (fe - 1), pg - span - fb + 1
        else:
            origin, ob, oe = self.slice_begin, fb, fe
        for g in range(ob, oe):
            pats = self.table.get(_canon(sl[g - origin:g - origin + span:self.stride], n))
            if not pats:
                continue
            for pat in pats:
                self._event(pat // self.parts, pat % self.parts, g, sl, origin, rev)

    def _event(self, r, j, g, sl, origin, rev):
        """One (window start g, read r, seed j) hit: what verify / apply_event do in the kernels."""
        L, pg, sh, a = self.read_len, self.pg_len, self.shift, self.acc
        c_in, _, X = self.state[r]
        if c_in <= self.min_mm or j * sh > g:
            return
        al = g - j * sh
        if al + L > pg:
            return
        assert al - origin >= 0 and al - origin + L <= len(sl), "halo too small"
        rep = pg - (al + L) if rev else al
        has_pos = c_in != 255
        limit = c_in - 1 if has_pos else self.max_mm
        c = int(np.count_nonzero(self.reads[r] != sl[al - origin:al - origin + L]))
        if c > limit:
            return
        order = (g << 8) | (self.parts - 1 - j)
        if has_pos and X == rep:
            a["same_pos_mask"][r] |= 1 << j
            a["same_pos_mm"][r] = c
            a["touched"][0] = 1
        else:
            cls = 0 if c <= self.min_mm else c
            a["best_key"][r] = min(int(a["best_key"][r]), (cls << 56) | (order << 8) | c)
            if has_pos:
                a["first_other_order"][r] = min(int(a["first_other_order"][r]), order)
                a["touched"][0] = 1

    def resolve_pass(self, rev):
        n, L, pg = self.shift, self.read_len, self.pg_len
        a = self.acc
        touched = int(a["touched"][0])      # like resolve_kernel: the rare accumulators are only read when the (merged) flag is set
        for r in range(self.n_reads):
            c_in, _, X = self.state[r]
            best, o1 = int(a["best_key"][r]), (int(a["first_other_order"][r]) if touched else KEY_INF)
            mask, cx = (int(a["same_pos_mask"][r]), int(a["same_pos_mm"][r])) if touched else (0, 255)
            if c_in <= self.min_mm:
                continue
            limit = c_in - 1 if c_in != 255 else self.max_mm
            if mask and cx <= limit and o1 != KEY_INF:
                aX = pg - X - L if rev else X
                for j in range(self.parts):
                    if (mask >> j) & 1:
                        order = ((aX + j * n) << 8) | (self.parts - 1 - j)
                        if order > o1:
                            cls = 0 if cx <= self.min_mm else cx
                            best = min(best, (cls << 56) | (order << 8) | cx)
                            break
            if best == KEY_INF:
                continue
            c, jj, g = best & 0xFF, (best >> 8) & 0xFF, (best >> 16) & 0xFFFFFFFFFF
            al = g - (self.parts - 1 - jj) * n
            self.state[r] = (c, 1 if rev else 0, pg - (al + L) if rev else al)
        self._reset_acc()

    def get_results(self, out=None):
        n = self.n_reads
        pos = np.array([0xFFFFFFFFFFFFFFFF if s[0] == 255 else s[2] for s in self.state], np.uint64)
        rc = np.array([s[1] for s in self.state], np.uint8)
        mm = np.array([s[0] for s in self.state], np.uint8)
        return M.MatchResult(pos, rc, mm, int((mm != 255).sum()), np.bincount(mm, minlength=256).astype(np.uint64))

    def close(self):
        pass
