"""CPU model of the ORDER-FREE form of stage 7's exact-match query that the CUDA kernels use (pgrc_b200/csrc/pgm_mem.cuh),
in plain numpy / Python over small inputs.  Test infrastructure: the product never imports it.

The reference's query (CopMEMMatcher::processExactMatchQueryTight, copmem/CopMEMMatcher.cpp:332-481) walks the destination
text sequentially and consults resMatches.back().  What it computes is nevertheless local:

  fv(q)      the first entry s of q's bucket (text order) that passes the destIsSrc filter (:384-386) and extends to at
             least minMatchLength characters (:399-407); the 4-byte guards (:396-398) never reject such an entry;
  M(q)       the match fv(q) extends to (left extension stops BEFORE the first character of either text: one character
             is dropped there, :405);
  visited    inside a group of 256 query positions (:364-419) position t + 1 follows t when t has no fv, t + 1 + skip
             otherwise — a push and a "covered by the previous match" jump (:388-393) advance alike, and the jump never
             leaves the group; the tail (:422-473) is one more group without an end;
  pushed     the visited positions with an fv, except those whose M lies on the diagonal of the previous visited-with-fv
             position's match and ends inside it (then it IS that match: the :388-393 test).
"""
import numpy as np

LIMIT = 12      # HASH_COLLISIONS_PER_POSITION_LIMIT (CopMEMMatcher.h:11)
MULTI = 256


def derive(L, min_len, N):
    """initParams + calcCoprimes (CopMEMMatcher.cpp:69-137)."""
    min_len = min(min_len, L)
    if L > 110: K = 56
    elif L > 62: K = 44
    elif L > 53: K = 40
    elif L > 46: K = 36
    elif L > 42: K = 32
    elif L > 32: K = 28
    else: K = (L // 4 - 1) * 4
    assert min_len >= 24
    K = min(K, (min_len // 4 - 1) * 4)
    t = L - K + 1
    assert t > 0
    if t >= 20:
        k1 = int(t ** 0.5) + 1; k2 = k1 - 1
        if k1 * k2 > t:
            k1 -= 1; k2 -= 1
    elif t >= 15: k1, k2 = 5, 3
    elif t >= 12: k1, k2 = 4, 3
    elif t >= 10: k1, k2 = 5, 2
    elif t >= 6: k1, k2 = 3, 2
    else: k1, k2 = t, 1
    order = 24
    hs = 1 << order
    while order < 31 and hs < N // k1:
        order += 1; hs = 1 << order
    return K, k1, k2, hs


def hashes(text, starts, K, hash_size):
    """maRushPrime1HashSparsified<K> (Hashes.h:54-76) at many positions."""
    h = np.full(len(starts), K, np.uint64)
    for j in range(K // 4):
        k = np.zeros(len(starts), np.uint64)
        for b in range(3 if j < 3 else 2):
            k |= text[starts + 4 * j + b].astype(np.uint64) << np.uint64(8 * b)
        h ^= (k + np.uint64(j))
        h *= np.uint64(171717)
    return (h & np.uint64(0xFFFFFFFF) & np.uint64(hash_size - 1)).astype(np.int64)


class CpuTextMatcher:
    """Same interface as pgrc_b200.matcher.GpuTextMatcher's share call (for the world_size > 1 gloo tests of the host logic)."""

    def __init__(self, src, target_len):
        self.src, self.target_len = np.asarray(src, np.uint8), target_len
        self.K, self.k1, self.k2, self.hash_size = derive(target_len, 0xFFFFFFFF, len(self.src))

    def match_texts_share(self, dest, dest_is_src, rev_compl, part, n_parts, min_len=0xFFFFFFFF):
        return match_texts(self.src, dest, dest_is_src, rev_compl, self.target_len, min_len, share=(part, n_parts))


def match_texts(src, dest, dest_is_src, rev_compl, target_len, min_len=0xFFFFFFFF, n_parts=1, share=None):
    """n_parts > 1: the groups of 256 query positions shared out over that many contexts as pgm_group_mem_match does it
    (context r: groups [G r / n, G (r + 1) / n)), the shares concatenated and the suppression test run across the seams."""
    src = np.asarray(src, np.uint8); dest = np.asarray(dest, np.uint8)
    N, N2 = len(src), len(dest)
    K, k1, k2, hs = derive(target_len, min_len, N)
    min_len = min(min_len, target_len)
    skip = K // k1 - 1
    # index: first LIMIT + 1 sampled positions per hash value, ascending
    pos = np.arange(0, N - K + 1, k1, dtype=np.int64)
    buckets = {}
    for p, h in zip(pos.tolist(), hashes(src, pos, K, hs).tolist()):
        b = buckets.setdefault(h, [])
        if len(b) <= LIMIT:
            b.append(p)
    nq = (N2 - K) // k2 + 1 if N2 >= K else 0
    qs = np.arange(nq, dtype=np.int64) * k2
    qh = hashes(dest, qs, K, hs).tolist() if nq else []

    def extend(s, q):
        """(src, len, dest) of the candidate, or None when the K-mers differ"""
        if not np.array_equal(src[s:s + K], dest[q:q + K]):
            return None
        b = 0
        while s + K + b < N and q + K + b < N2 and src[s + K + b] == dest[q + K + b]:
            b += 1
        a, amax = 0, min(s, q)
        while a < amax and src[s - a - 1] == dest[q - a - 1]:
            a += 1
        adj = 1 if a == amax else 0
        return s - a + adj, K + a + b - adj, q - a + adj

    fvm = [None] * nq
    for i in range(nq):
        q = i * k2
        for s in buckets.get(qh[i], ()):
            if dest_is_src and ((N2 - s < q) if rev_compl else (q >= s)):
                continue
            m = extend(s, q)
            if m is not None and m[1] >= min_len:
                fvm[i] = m
                break
    # visited positions, group by group
    n_groups = 0
    while n_groups * MULTI * k2 + K + MULTI * k2 < N2 + 1:
        n_groups += 1
    assert nq == 0 or n_groups + 1 == (nq + MULTI - 1) // MULTI      # the tail (:422-473) has 1 .. 256 positions: ceil(nq / 256) groups in all
    groups = (nq + MULTI - 1) // MULTI
    visited = []
    parts = [share[0]] if share is not None else range(n_parts)
    if share is not None:
        n_parts = share[1]
    for r in parts:                                                    # (a context's share; the walk of a group never looks outside it)
        for g in range(groups * r // n_parts, groups * (r + 1) // n_parts):
            t, end = g * MULTI, min((g + 1) * MULTI, nq)
            while t < end:
                if fvm[t] is not None:
                    visited.append(t); t += skip + 1
                else:
                    t += 1
    if share is not None:                                              # one rank's share: raw matches + their query positions
        return (np.array([fvm[t] for t in visited], np.uint64).reshape(-1, 3), np.array([t * k2 for t in visited], np.uint64))
    out, prev = [], None
    for t in visited:
        m, q = fvm[t], t * k2
        if not (prev is not None and m[2] - m[0] == prev[2] - prev[0] and q + K < prev[2] + prev[1]):
            out.append(m)
        prev = m
    return np.array(out, np.uint64).reshape(-1, 3)
