// pgm_mem.cuh — PgRC stage 7: exact matches between pseudogenomes.
//
// Replaces CopMEMMatcher as SimplePgMatcher uses it (matching/SimplePgMatcher.cpp:11-55): the constructor's index of the
// source text (copmem/CopMEMMatcher.cpp:571-593 -> processRef :176-231, the serial build = the reference at -t 1; the
// kernels of pgm_copmem.cuh) and matchTexts (:605-624 -> processExactMatchQueryTight :332-481).
//
// The reference's query walks the destination text sequentially and consults resMatches.back(); what it computes is local
// (tests/cpu_mem_model.py states the argument and checks it against the sequential oracle):
//   fv(q)     the first entry s of q's bucket, in text order, that passes the destIsSrc filter (:384-386), whose K-mer
//             equals the one at q and which extends to >= minMatchLength characters (:399-407).  The 4-byte guards
//             l1/l2/r1/r2 (:396-398) never reject such an entry: of its >= L - K extra characters at least (L - K) / 2
//             lie on one side, inside both texts, and that side's guard is loaded and equal.
//   M(q)      the match fv(q) extends to.  The left extension stops BEFORE comparing when it reaches the first character
//             of either text, and the match then begins one character late (:405) — reproduced.
//   visited   inside a group of 256 query positions (:364-419) position t + 1 follows t when t has no fv and
//             t + 1 + skip otherwise: a push and a "covered by the previous match" jump (:388-393) advance alike, and the
//             jump is cut at the end of the group; the tail loop (:422-473) is one more group.
//   pushed    the visited positions with an fv, except those whose match lies on the diagonal of the previous such
//             position's match and whose K-mer ends inside it (then it IS that match: the :388-393 test).
// Kernels: mem_pack_kernel (destination text -> planes + mask of the symbols outside ACGT), mem_query_kernel (fv of every
// query position, bounded extension), mem_walk_kernel (thread per group: visited & fv), mem_extend_kernel (full extension
// of the emitted positions, in push order), mem_flag_kernel / mem_compact_kernel (the :388-393 test, final vector).
#pragma once
#include "pgm_copmem.cuh"

#define PGM_MEM_GROUP 256u          // MULTI of processExactMatchQueryTight (CopMEMMatcher.cpp:336)
#define PGM_MEM_THREADS 256

namespace pgm {

struct MemMatch { unsigned long long src, len, dest; };          // = PgTools::TextMatch (matching/TextMatchers.h:11-14)

struct MemParams {
    const uint32_t *slo, *shi;      // source text planes (origin at base 0; zero words in front and behind)
    const uint32_t *dlo, *dhi, *dinv;   // destination text planes; dinv = symbols outside ACGT (nullptr: none)
    uint64_t N, N2;
    uint32_t K, k1, k2, hash_mask, min_len, skip;
    int dest_is_src, rev_compl;
    const uint32_t *nib, *coarse, *entries; // the index (pgm_copmem.cuh: compact bucket directory + kept sample indices)
    uint64_t q0, nq;                // this launch's query positions: indices q0 .. q0 + nq - 1 of 0, k2, 2 k2, ... <= N2 - K (q0: a multiple
                                    // of 256; a context of several takes a range of groups, pgm_group_mem_match)
    uint64_t n_groups;              // groups of 256 query positions in that range, minus one: the main loop's full groups (:364) and the
                                    // tail (:422), which has 1 .. 256 positions — ceil(nq / 256) groups, all walked alike
    uint32_t *fv;                   // [nq] sample index of fv(q), 0xFFFFFFFF = none
    uint32_t *has_fv;               // [ceil(nq / 32) + 8] bit per query position
    uint32_t *emit;                 // same shape: visited & fv
    uint32_t *group_count;          // [n_groups + 1] emitted positions per group
    const uint32_t *group_start;    // exclusive prefix of group_count
    MemMatch *raw;                  // emitted matches in push order, before the :388-393 test
    uint64_t *raw_q;                // their query positions
    uint32_t *keep;                 // [n_raw] 1 = pushed
    const uint32_t *keep_start;     // exclusive prefix of keep
    MemMatch *out;
    uint64_t n_raw;
};

// destination text: ASCII -> two bit planes + the mask of symbols outside ACGT (an 'N', or the '%' a previous call left);
// one thread per 32-base word
__global__ void __launch_bounds__(PGM_MEM_THREADS) mem_pack_kernel(const uint8_t *__restrict__ ascii, uint64_t n, uint32_t *__restrict__ lo,
                                                                   uint32_t *__restrict__ hi, uint32_t *__restrict__ inv) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t base = w * 32;
    if (base >= n) return;
    const uint32_t cnt = (uint32_t)min((uint64_t)32, n - base);
    uint32_t bytes[8];
    if (cnt == 32 && ((reinterpret_cast<uintptr_t>(ascii + base) & 15) == 0)) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(ascii + base));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(ascii + base) + 1);
        bytes[0] = a.x; bytes[1] = a.y; bytes[2] = a.z; bytes[3] = a.w;
        bytes[4] = b.x; bytes[5] = b.y; bytes[6] = b.z; bytes[7] = b.w;
    } else {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t i = q * 4 + b;
                v |= (i < cnt ? (uint32_t)ascii[base + i] : (uint32_t)'A') << (8 * b);
            }
            bytes[q] = v;
        }
    }
    uint32_t l = 0, h = 0, bad = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const uint32_t c = (bytes[q] >> (8 * b)) & 0xFFu;
            const uint32_t x = (c >> 1) & 3u, code = x ^ (x >> 1);                  // A0 C1 G2 T3
            const uint32_t ok = ((0x54474341u >> (8 * code)) & 0xFFu) == c;
            const uint32_t bit = q * 4 + b;
            l |= (ok ? (code & 1u) : 0u) << bit;
            h |= (ok ? (code >> 1) : 0u) << bit;
            bad |= (ok ? 0u : 1u) << bit;
        }
    }
    lo[w] = l; hi[w] = h; inv[w] = bad;
}

// bits [x, x + 64) of a plane, x may be negative (down to the zero words in front of the plane)
__device__ __forceinline__ uint64_t mem_bits64(const uint32_t *plane, long long x) {
    const long long w = x >> 5;                         // floor
    const uint32_t s = (uint32_t)(x & 31);
    const uint32_t a = __ldg(plane + w), b = __ldg(plane + w + 1), c = __ldg(plane + w + 2);
    return (uint64_t)__funnelshift_r(a, b, s) | ((uint64_t)__funnelshift_r(b, c, s) << 32);
}

// differences between source characters [x1, x1 + 64) and destination characters [x2, x2 + 64), one bit per character
__device__ __forceinline__ uint64_t mem_diff64(const MemParams &p, long long x1, long long x2) {
    uint64_t d = (mem_bits64(p.slo, x1) ^ mem_bits64(p.dlo, x2)) | (mem_bits64(p.shi, x1) ^ mem_bits64(p.dhi, x2));
    if (p.dinv) d |= mem_bits64(p.dinv, x2);
    return d;
}

// equal characters from (x1, x2) to the right, at most `limit`
__device__ __forceinline__ uint64_t mem_right(const MemParams &p, uint64_t x1, uint64_t x2, uint64_t limit) {
    uint64_t n = 0;
    while (n < limit) {
        const uint64_t d = mem_diff64(p, (long long)(x1 + n), (long long)(x2 + n));
        if (d) { n += (uint64_t)(__ffsll((long long)d) - 1); break; }
        n += 64;
    }
    return min(n, limit);
}

// equal characters from (x1 - 1, x2 - 1) to the left, at most `limit`
__device__ __forceinline__ uint64_t mem_left(const MemParams &p, uint64_t x1, uint64_t x2, uint64_t limit) {
    uint64_t n = 0;
    while (n < limit) {
        const uint64_t d = mem_diff64(p, (long long)(x1 - n) - 64, (long long)(x2 - n) - 64);    // bit 63 = the character just left of the run
        if (d) { n += (uint64_t)__clzll((long long)d); break; }
        n += 64;
    }
    return min(n, limit);
}

// the destIsSrc filter (CopMEMMatcher.cpp:384-386)
__device__ __forceinline__ bool mem_filtered(const MemParams &p, uint64_t s, uint64_t q) {
    return p.dest_is_src && (p.rev_compl ? p.N2 - s < q : q >= s);
}

__global__ void __launch_bounds__(PGM_MEM_THREADS) mem_query_kernel(const __grid_constant__ MemParams p) {
    __shared__ uint32_t lut[256];
    cm_build_lut(lut);
    const uint64_t i = (uint64_t)blockIdx.x * PGM_MEM_THREADS + threadIdx.x;
    const uint64_t pol_keep = policy_evict_last();
    uint32_t found = 0xFFFFFFFFu;
    if (i < p.nq) {
        const uint64_t q = (p.q0 + i) * p.k2;
        const uint64_t maskK = p.K >= 64 ? ~0ull : ((1ull << p.K) - 1ull);
        const uint64_t lo = mem_bits64(p.dlo, (long long)q), hi = mem_bits64(p.dhi, (long long)q);
        const uint64_t bad = p.dinv ? (mem_bits64(p.dinv, (long long)q) & maskK) : 0ull;
        if (!bad) {                                                         // a K-mer with an N equals no source K-mer
            const uint32_t h = cm_hash(p.K, lo, hi, 0ull, p.hash_mask, lut);
            uint32_t b0, b1;
            cm_bucket(p.nib, p.coarse, h, pol_keep, b0, b1);
            const uint64_t need = p.min_len - p.K;                          // extra characters a match needs (min_len >= K)
            for (uint32_t j = b0; j < b1; j++) {
                const uint32_t e = __ldg(p.entries + j);
                const uint64_t s = (uint64_t)e * p.k1;
                if (mem_filtered(p, s, q)) continue;
                if (((mem_bits64(p.slo, (long long)s) ^ lo) | (mem_bits64(p.shi, (long long)s) ^ hi)) & maskK) continue;   // memcmp(curr1, curr2, K), :407
                // right - p1 > minMatchLength  <=>  K + a + b - adj >= minMatchLength; extensions counted up to what decides it
                const uint64_t b = mem_right(p, s + p.K, q + p.K, min(need + 1, min(p.N - s - p.K, p.N2 - q - p.K)));
                const uint64_t amax = min(s, q);
                const uint64_t a = mem_left(p, s, q, min(need + 2, amax));
                const uint64_t adj = a == amax ? 1ull : 0ull;
                if (p.K + a + b - adj >= p.min_len) { found = e; break; }
            }
        }
        p.fv[i] = found;
    }
    const uint32_t m = __ballot_sync(PGM_FULL, found != 0xFFFFFFFFu);
    if ((threadIdx.x & 31u) == 0 && i < p.nq) p.has_fv[i >> 5] = m;
}

// thread per group of 256 query positions: the positions the sequential query visits, of those the ones with an fv
__global__ void __launch_bounds__(PGM_MEM_THREADS) mem_walk_kernel(const __grid_constant__ MemParams p) {
    const uint64_t g = (uint64_t)blockIdx.x * PGM_MEM_THREADS + threadIdx.x;
    if (g > p.n_groups) return;
    const uint64_t first = g * PGM_MEM_GROUP;
    const uint32_t len = first >= p.nq ? 0u : (uint32_t)min((uint64_t)PGM_MEM_GROUP, p.nq - first);      // (the tail has at most 256 positions too)
    uint32_t w[8], out[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { w[k] = (uint32_t)(32 * k) < len ? p.has_fv[(first >> 5) + k] : 0u; out[k] = 0; }
    uint32_t t = 0, count = 0;
    while (t < len) {
        // next position >= t with an fv
        uint32_t k = t >> 5;
        uint32_t cur = 0;
#pragma unroll
        for (int kk = 0; kk < 8; kk++) if ((uint32_t)kk == k) cur = w[kk];
        cur &= 0xFFFFFFFFu << (t & 31u);
        if (!cur) { t = (k + 1) << 5; continue; }
        t = (k << 5) + (uint32_t)(__ffs((int)cur) - 1);
        if (t >= len) break;
#pragma unroll
        for (int kk = 0; kk < 8; kk++) if ((uint32_t)kk == k) out[kk] |= 1u << (t & 31u);
        count++;
        t += p.skip + 1;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) if ((uint32_t)(32 * k) < len) p.emit[(first >> 5) + k] = out[k];
    p.group_count[g] = count;
}

// thread per 32 query positions: full extension of the emitted ones, written in push order
__global__ void __launch_bounds__(PGM_MEM_THREADS) mem_extend_kernel(const __grid_constant__ MemParams p) {
    const uint64_t wi = (uint64_t)blockIdx.x * PGM_MEM_THREADS + threadIdx.x;
    if (wi * 32 >= p.nq) return;
    uint32_t m = p.emit[wi];
    if (!m) return;
    const uint64_t g = wi >> 3;
    uint64_t rank = p.group_start[g];
    for (uint64_t k = g << 3; k < wi; k++) rank += __popc(p.emit[k]);
    while (m) {
        const uint32_t bit = (uint32_t)(__ffs((int)m) - 1);
        m &= m - 1;
        const uint64_t i = wi * 32 + bit, q = (p.q0 + i) * p.k2;
        const uint64_t s = (uint64_t)p.fv[i] * p.k1;
        const uint64_t b = mem_right(p, s + p.K, q + p.K, min(p.N - s - p.K, p.N2 - q - p.K));
        const uint64_t amax = min(s, q);
        const uint64_t a = mem_left(p, s, q, amax);
        const uint64_t adj = a == amax ? 1ull : 0ull;
        MemMatch mm;
        mm.src = s - a + adj; mm.len = p.K + a + b - adj; mm.dest = q - a + adj;
        p.raw[rank] = mm;
        p.raw_q[rank] = q;
        rank++;
    }
}

// the "covered by the previous match" test (CopMEMMatcher.cpp:388-393) against the previous emitted match
__global__ void __launch_bounds__(PGM_MEM_THREADS) mem_flag_kernel(const __grid_constant__ MemParams p) {
    const uint64_t i = (uint64_t)blockIdx.x * PGM_MEM_THREADS + threadIdx.x;
    if (i >= p.n_raw) return;
    uint32_t keep = 1;
    if (i > 0) {
        const MemMatch cur = p.raw[i], prev = p.raw[i - 1];
        if (cur.dest - cur.src == prev.dest - prev.src && p.raw_q[i] + p.K < prev.dest + prev.len) keep = 0;
    }
    p.keep[i] = keep;
}

__global__ void __launch_bounds__(PGM_MEM_THREADS) mem_compact_kernel(const __grid_constant__ MemParams p) {
    const uint64_t i = (uint64_t)blockIdx.x * PGM_MEM_THREADS + threadIdx.x;
    if (i >= p.n_raw) return;
    if (p.keep[i]) p.out[p.keep_start[i]] = p.raw[i];
}

} // namespace pgm
