"""CPU model of one rank's context in the ROUTED multi-GPU scheme, for the world_size > 1 gloo tests of the host logic
(pgrc_b200/matcher.py: run_plan_routed, TorchComm.exchange_counts / all_to_all, read_ranges).  Same interface as
GpuReadsMatcher's route_* methods; plain Python over small inputs; the exchange entries are 16 bytes here (the product
treats them as opaque bytes).  Test infrastructure: the product never imports it."""
import hashlib

import numpy as np
import torch

from cpu_shard_model import _COMP, CpuShardMatcher, _canon


def _key(canon) -> int:
    return int.from_bytes(hashlib.blake2b(bytes(canon), digest_size=8).digest(), "little") >> 1


class CpuRouteMatcher(CpuShardMatcher):
    def set_text(self, text):
        self.text = np.asarray(text, np.uint8)
        self.pg_len = len(self.text)
        self.rc_text = np.array([_COMP[int(c)] for c in self.text[::-1]], np.uint8)

    def route_slot(self, slot):
        """Two sets of exchange buffers in the product; the model keeps the receive buffers per slot."""
        self.slot = slot

    def route_config(self, rank, world, read_begin, round_windows=0):
        self.slot, self.recvs = 0, {}
        self.rank, self.world, self.read_begin = rank, world, [int(x) for x in read_begin]
        self.round_windows = round_windows or (1 << 30)

    # -- plan (same on every rank)
    def _cut(self, k):
        nw = max(0, self.pg_len - self.span + 1)
        return nw if k >= self.world else (nw * k // self.world) // 128 * 128

    def _range(self, g, rnd):
        lo, hi = self._cut(g), self._cut(g + 1)
        b = min(hi, lo + rnd * self.round_windows)
        return b, min(hi, b + self.round_windows)

    def route_rounds(self):
        longest = max(self._cut(g + 1) - self._cut(g) for g in range(self.world))
        return max(1, -(-longest // self.round_windows))

    @staticmethod
    def _segs(rows_by_dest, world):
        counts, segs = [], []
        for d in range(world):
            rows = rows_by_dest.get(d, [])
            counts.append(len(rows))
            segs.append(torch.from_numpy(np.array(rows, np.int64).reshape(-1, 2).copy()).view(torch.uint8).reshape(-1) if rows else None)

        class _E:      # what GpuReadsMatcher's emit steps return (matcher._Emit)
            pass
        e = _E()
        e.counts, e.segs, e.eb, e.buf = counts, segs, 16, None
        return e

    def route_begin(self, seed_len, parts, max_mm, min_mm, continuation=False):
        self.seed_len, self.parts, self.max_mm, self.min_mm = seed_len, parts, max_mm, min_mm
        self.stride, self.shift, self.span = 1, seed_len, seed_len
        if not continuation:
            self.state = [(255, 0, None)] * self.n_reads
        self._reset_acc()
        out = {}
        base = self.read_begin[self.rank]
        for r in range(self.n_reads):
            if continuation and self.state[r][0] <= min_mm:
                continue
            for j in range(parts):
                k = _key(_canon(self.reads[r][j * seed_len:(j + 1) * seed_len], seed_len))
                out.setdefault(k % self.world, []).append((k, (base + r) * parts + j))
        return self._segs(out, self.world)

    def route_recv(self, kind, in_counts, entry_bytes):
        self.recvs[(kind, self.slot)] = [torch.empty(int(c) * entry_bytes, dtype=torch.uint8) if c else None for c in in_counts]
        return self.recvs[(kind, self.slot)]

    def _rows(self, kind):
        return [(v.view(torch.int64).reshape(-1, 2).tolist() if v is not None else []) for v in self.recvs[(kind, self.slot)]]

    def route_build(self, n_in):
        self.table = {}
        for rows in self._rows(0):
            for k, pat in rows:
                self.table.setdefault(k, []).append(pat)
        assert sum(len(v) for v in self.table.values()) == n_in

    def route_scan(self, rev, rnd):
        b, e = self._range(self.rank, rnd)
        t = self.rc_text if rev else self.text
        out = {}
        for g in range(b, e):
            k = _key(_canon(t[g:g + self.span], self.seed_len))
            out.setdefault(k % self.world, []).append((k, g - b))
        return self._segs(out, self.world)

    def route_probe(self, rev, rnd, in_counts):
        out = {}
        for s, rows in enumerate(self._rows(1)):
            assert len(rows) == in_counts[s]
            base = self._range(s, rnd)[0]
            for k, rel in rows:
                for pat in self.table.get(k, ()):
                    read = pat // self.parts
                    owner = max(d for d in range(self.world) if self.read_begin[d] <= read)
                    out.setdefault(owner, []).append((base + rel, pat))
        return self._segs(out, self.world)

    # launch / fetch halves (the product overlaps them with exchanges; the model computes at launch and hands over at fetch)
    def route_scan_launch(self, rev, rnd):
        self._emitted = getattr(self, "_emitted", {})
        self._emitted[(1, self.slot)] = self.route_scan(rev, rnd)

    def route_probe_launch(self, rev, rnd, in_counts):
        self._emitted = getattr(self, "_emitted", {})
        self._emitted[(2, self.slot)] = self.route_probe(rev, rnd, in_counts)

    def route_fetch(self, kind):
        return self._emitted[(kind, self.slot)]

    def synchronize(self):
        pass

    def route_verify(self, rev, n_in):
        t = self.rc_text if rev else self.text
        seen = 0
        for rows in self._rows(2):
            for g, pat in rows:
                seen += 1
                self._event(pat // self.parts - self.read_begin[self.rank], pat % self.parts, g, t, 0, rev)
        assert seen == n_in
