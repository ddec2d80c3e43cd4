// pgm_copmem.cuh — matching mode 'c': CopMEMReadsApproxMatcher (ReadsMatchers.cpp:411-451) over CopMEMMatcher
// (copmem/CopMEMMatcher.cpp), the matcher PgRC's release CLI runs by default.  SURVEY.md §8(f) rank 1.
//
// The reference indexes the TEXT (every k1-th position, hash of K characters, at most 13 positions per hash value, first
// come first kept: genCumm / processRef, CopMEMMatcher.cpp:139-233) and queries every read on its own
// (processApproxMatchQueryTight, :483-566): every k2-th read offset, bucket entries in text order, verification with an
// early exit, strict improvement, a false-match budget after which buckets are cut to 4 entries.  Reads are independent,
// so a thread replays one read's query exactly; the index is rebuilt per pass on the (reverse-complement) planes:
//   cm_hash_kernel     sample s <-> text position s * k1: hash of the K characters there, count per hash value
//   (scan)             exclusive prefix sums of the counts (all entries) and of min(count, 13) (kept entries)
//   cm_scatter_kernel  every sample into the (unordered) full bucket of its hash value
//   cm_select_kernel   per hash value: its 13 smallest positions in ascending order = the first 13 in text order
//   cm_query_kernel    thread per read
// This is the reference's SERIAL index build, i.e. its behaviour at -t 1 (its multithreaded build orders the buckets
// differently and races); the hash is maRushPrime1HashSparsified<K> (copmem/Hashes.h:54-76) over the ASCII characters.
#pragma once
#include "pgm_kernels.cuh"

#define PGM_CM_COLLISIONS_LIMIT 12      // HASH_COLLISIONS_PER_POSITION_LIMIT (CopMEMMatcher.h:11): a bucket keeps LIMIT + 1 positions
#define PGM_CM_TRUNCATED_BUCKET 4       // UNLIMITED_NUMBER_OF_HASH_COLLISIONS_PER_POSITION (:13)
#define PGM_CM_THREADS 256
#define PGM_CM_SCAN_BLOCK 1024

namespace pgm {

struct CopmemParams {
    const uint32_t *tlo, *thi;      // planes of this pass's WHOLE text, origin at word 0
    uint64_t pg_len;
    uint32_t K, k1, k2, hash_mask;
    uint32_t n_sampled;             // text positions 0, k1, 2 k1, ... <= pg_len - K
    uint32_t *count;                // [hash_size + 1] entries per hash value (all of them)
    uint32_t *start_all;            // [hash_size + 1] exclusive prefix of count
    uint32_t *cumm;                 // [hash_size + 1] exclusive prefix of min(count, 13): bucket h = [cumm[h], cumm[h+1])
    uint32_t *sample_rank;          // [n_sampled] arrival rank of the sample among those with its hash value
    uint32_t *sample_hash;          // [n_sampled]
    uint32_t *all_entries;          // [n_sampled] sample indices, full buckets, unordered
    uint32_t *entries;              // kept sample indices (position = index * k1), ascending inside a bucket
    // compact bucket directory for stage 7's queries (cm_bucket, pgm_mem.cuh): cumm[] is 4 bytes per hash value (256 MB for the
    // config-2 text: every lookup a random DRAM line); kept entries per hash value fit a nibble (<= 13), so 4 bits per hash
    // value + one 32-bit prefix per 64 hash values (36 MB) can stay in the L2: a lookup is one 256-bit load + nibble sums
    uint32_t *nib;                  // [hash_size / 8] nibble h & 7 of word h >> 3 = min(count[h], 13)
    uint32_t *coarse;               // [hash_size / 64] cumm[64 g]
    // query
    ReadsView reads;
    uint32_t n_reads, max_mm, min_mm;
    int rev_mode;
    // (pgm_copmem_warp.cuh) cm_query_kernel for the marked reads only: only_marked[r - marked_base] == 0xFF; nullptr = every read
    const uint8_t *only_marked;
    uint32_t marked_base, marked_count;
};

// four ASCII characters of four 2-bit codes: index = lo nibble | hi nibble << 4 (character t at bits t / t + 4)
__device__ __forceinline__ uint32_t cm_ascii4(uint32_t lo4, uint32_t hi4) {
    uint32_t w = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const uint32_t code = ((lo4 >> t) & 1u) | (((hi4 >> t) & 1u) << 1);
        w |= ((0x54474341u >> (8 * code)) & 0xFFu) << (8 * t);          // "ACGT"
    }
    return w;
}

// maRushPrime1HashSparsified<K> (Hashes.h:54-76) of K characters given as bit strings (character t at bit t):
// K / 4 little-endian words, the first three masked to 3 characters, the others to 2; `n` marks 'N' characters (reads only).
// lut[lo nibble | hi nibble << 4] = the four ASCII characters of four codes (256 entries, shared memory).
__device__ __forceinline__ uint32_t cm_hash(uint32_t K, uint64_t lo, uint64_t hi, uint64_t n, uint32_t hash_mask, const uint32_t *lut) {
    unsigned long long hash = K;
    const uint32_t words = K >> 2;
    for (uint32_t j = 0; j < words; j++) {
        const uint32_t sh = 4 * j;
        uint32_t k = lut[((uint32_t)(lo >> sh) & 0xFu) | (((uint32_t)(hi >> sh) & 0xFu) << 4)];
        const uint32_t nn = (uint32_t)(n >> sh) & 0xFu;
        if (nn) {
#pragma unroll
            for (int t = 0; t < 4; t++)
                if ((nn >> t) & 1u) k = (k & ~(0xFFu << (8 * t))) | (0x4Eu << (8 * t));       // 'N'
        }
        k &= j < 3 ? 0x00FFFFFFu : 0x0000FFFFu;
        k += j;
        hash ^= k;
        hash *= 171717ull;
    }
    return (uint32_t)hash & hash_mask;
}

__device__ __forceinline__ void cm_build_lut(uint32_t *lut) {      // by a block of >= 256 threads, followed by a barrier
    if (threadIdx.x < 256) lut[threadIdx.x] = cm_ascii4(threadIdx.x & 15u, threadIdx.x >> 4);
    __syncthreads();
}

// bits [x, x + 64) of a plane (words beyond the data are zero pad / tail)
__device__ __forceinline__ uint64_t cm_bits64(const uint32_t *plane, uint64_t x) {
    const uint64_t w = x >> 5;
    const uint32_t s = (uint32_t)(x & 31);
    const uint32_t a = __ldg(plane + w), b = __ldg(plane + w + 1), c = __ldg(plane + w + 2);
    return (uint64_t)__funnelshift_r(a, b, s) | ((uint64_t)__funnelshift_r(b, c, s) << 32);
}

__global__ void __launch_bounds__(PGM_CM_THREADS) cm_hash_kernel(const __grid_constant__ CopmemParams p) {
    __shared__ uint32_t lut[256];
    cm_build_lut(lut);
    const uint32_t s = blockIdx.x * PGM_CM_THREADS + threadIdx.x;
    if (s >= p.n_sampled) return;
    const uint64_t x = (uint64_t)s * p.k1;
    const uint32_t h = cm_hash(p.K, cm_bits64(p.tlo, x), cm_bits64(p.thi, x), 0ull, p.hash_mask, lut);
    p.sample_hash[s] = h;
    p.sample_rank[s] = atomicAdd(p.count + h, 1u);
}

// exclusive prefix sums of a uint32 array in three steps (sums of 1024-element blocks, scan of those — mismatch_scan_kernel —,
// then the blocks themselves); CAP > 0 sums min(v, CAP) instead of v
template <uint32_t CAP>
__global__ void __launch_bounds__(PGM_CM_SCAN_BLOCK) cm_block_sums_kernel(const uint32_t *__restrict__ v, uint32_t n, unsigned long long *block_sums) {
    __shared__ unsigned int s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * PGM_CM_SCAN_BLOCK + threadIdx.x;
    uint32_t x = i < n ? v[i] : 0u;
    if (CAP) x = min(x, CAP);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(PGM_FULL, x, o);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(&s_sum, x);
    __syncthreads();
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s_sum;
}

template <uint32_t CAP>
__global__ void __launch_bounds__(PGM_CM_SCAN_BLOCK) cm_block_scan_kernel(const uint32_t *__restrict__ v, uint32_t n, const unsigned long long *block_base,
                                                                          uint32_t *__restrict__ out /*[n + 1]*/) {
    __shared__ uint32_t s_warp[PGM_CM_SCAN_BLOCK / 32];
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t i = blockIdx.x * PGM_CM_SCAN_BLOCK + t;
    uint32_t c = i < n ? v[i] : 0u;
    if (CAP) c = min(c, CAP);
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(PGM_FULL, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t w = 0; w < warp; w++) wbase += s_warp[w];
    const uint32_t before = (uint32_t)block_base[blockIdx.x] + wbase + (x - c);
    if (i < n) out[i] = before;
    if (i == n - 1) out[n] = before + c;
}

__global__ void __launch_bounds__(PGM_CM_THREADS) cm_scatter_kernel(const __grid_constant__ CopmemParams p) {
    const uint32_t s = blockIdx.x * PGM_CM_THREADS + threadIdx.x;
    if (s >= p.n_sampled) return;
    const uint32_t h = p.sample_hash[s];
    p.all_entries[p.start_all[h] + p.sample_rank[s]] = s;
}

// per hash value: the min(count, 13) smallest sample indices of its full bucket, ascending (= the first in text order, what
// the serial reference keeps: genCumm, CopMEMMatcher.cpp:139-169)
__global__ void __launch_bounds__(PGM_CM_THREADS) cm_select_kernel(const __grid_constant__ CopmemParams p) {
    const uint32_t h = blockIdx.x * PGM_CM_THREADS + threadIdx.x;
    if (h > p.hash_mask) return;
    const uint32_t c = p.count[h];
    if (c == 0) return;
    const uint32_t *src = p.all_entries + p.start_all[h];
    uint32_t *dst = p.entries + p.cumm[h];
    const uint32_t keep = min(c, (uint32_t)PGM_CM_COLLISIONS_LIMIT + 1u);
    // one pass with the 13 smallest so far kept sorted in registers: an element below the largest of them bubbles in
    // (hot hash values — homopolymers, satellites — have millions of samples; almost all of them fail the first compare)
    uint32_t best[PGM_CM_COLLISIONS_LIMIT + 1];
#pragma unroll
    for (int t = 0; t <= PGM_CM_COLLISIONS_LIMIT; t++) best[t] = 0xFFFFFFFFu;
    for (uint32_t k = 0; k < c; k++) {
        uint32_t x = src[k];
        if (x < best[PGM_CM_COLLISIONS_LIMIT]) {
#pragma unroll
            for (int t = 0; t <= PGM_CM_COLLISIONS_LIMIT; t++) {
                const uint32_t lo = min(best[t], x), hi = max(best[t], x);
                best[t] = lo; x = hi;
            }
        }
    }
#pragma unroll
    for (int t = 0; t <= PGM_CM_COLLISIONS_LIMIT; t++)
        if ((uint32_t)t < keep) dst[t] = best[t];
}

// the compact directory from count[] and cumm[] (see CopmemParams): thread per 8 hash values
__global__ void __launch_bounds__(PGM_CM_THREADS) cm_compact_kernel(const __grid_constant__ CopmemParams p) {
    const uint32_t w = blockIdx.x * PGM_CM_THREADS + threadIdx.x;
    if (w > (p.hash_mask >> 3)) return;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) v |= min(p.count[8 * w + k], (uint32_t)PGM_CM_COLLISIONS_LIMIT + 1u) << (4 * k);
    p.nib[w] = v;
    if ((w & 7u) == 0) p.coarse[w >> 3] = p.cumm[8 * w];
}

// sum of the eight nibbles of a word (each <= 13)
__device__ __forceinline__ uint32_t cm_nibsum(uint32_t w) {
    return (((w & 0x0F0F0F0Fu) + ((w >> 4) & 0x0F0F0F0Fu)) * 0x01010101u) >> 24;
}

// bucket [b0, b1) of hash value h from the compact directory (= [cumm[h], cumm[h + 1])); the directory's lines carry an L2
// evict_last policy: it is the one structure of a query that can stay in the L2 next to the entries and the text windows
__device__ __forceinline__ void cm_bucket(const uint32_t *nib, const uint32_t *coarse, uint32_t h, uint64_t pol_keep, uint32_t &b0, uint32_t &b1) {
    const uint32_t g = h >> 6, i = h & 63u, wi = i >> 3;
    const u32x8 w = ld256_stream_hint(nib + 8 * (size_t)g, pol_keep);
    uint32_t before = ld_u32_hint(coarse + g, pol_keep), mine = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if ((uint32_t)k < wi) before += cm_nibsum(w.v[k]);
        else if ((uint32_t)k == wi) {
            const uint32_t sh = 4u * (i & 7u);
            before += cm_nibsum(w.v[k] & ((1u << sh) - 1u));
            mine = (w.v[k] >> sh) & 15u;
        }
    }
    b0 = before; b1 = before + mine;
}

// processApproxMatchQueryTight (CopMEMMatcher.cpp:483-566) + the per-read part of CopMEMReadsApproxMatcher::executeMatching
// (ReadsMatchers.cpp:427-448), one thread per read; the read's state in its record is updated in place.
__global__ void __launch_bounds__(PGM_CM_THREADS, 4) cm_query_kernel(const __grid_constant__ CopmemParams p, unsigned long long *counters) {
    __shared__ uint32_t lut[256];
    cm_build_lut(lut);
    const uint32_t r = blockIdx.x * PGM_CM_THREADS + threadIdx.x;
    if (r >= p.n_reads) return;
    if (p.only_marked && (r < p.marked_base || r - p.marked_base >= p.marked_count || p.only_marked[r - p.marked_base] != 0xFFu)) return;
    uint32_t stride16; bool is_n;
    uint4 *rec = record_of(p.reads, r, stride16, is_n);
    const uint4 hd = __ldcg(rec);
    const unsigned long long st = ((unsigned long long)hd.y << 32) | hd.x;
    const uint32_t mm0 = (uint32_t)(st >> 56);
    if (mm0 <= p.min_mm) return;                                        // ReadsMatchers.cpp:429
    const uint32_t N2 = p.reads.read_len, W = p.reads.W, K = p.K;
    // the read's planes in registers (8 words each at most: read length <= 255; every index below is static); zero beyond the read
    uint32_t rl[10], rh[10], rn[10];
#pragma unroll
    for (int g = 0; g < 10; g++) {
        rl[g] = 0; rh[g] = 0; rn[g] = 0;
        if (g < 8 && (uint32_t)g < W) {
            if (is_n) { const uint4 v = __ldcg(rec + 1 + g); rl[g] = v.x; rh[g] = v.y; rn[g] = v.z; }
            else { const uint4 v = __ldcg(rec + 1 + (g >> 1)); rl[g] = (g & 1) ? v.z : v.x; rh[g] = (g & 1) ? v.w : v.y; }
        }
    }
    uint32_t max_mm = p.max_mm;
    if (mm0 < max_mm) max_mm = mm0 - 1u;                                 // :488-489
    const uint32_t trim8 = (N2 >> 3) << 3;
    const unsigned long long limit = (unsigned long long)((N2 + 1 - K) / p.k2);   // x AVERAGE_HASH_COLLISIONS_PER_POSITION_LIMIT (1)
    unsigned long long cur_false = 0, n_ver = 0;
    uint64_t match_pos = 0;
    uint32_t cur = mm0;
    bool found = false, done = false;
    uint32_t i1 = 0;                                                     // read offsets 0, k2, 2 k2, ... while i1 + K <= N2
#pragma unroll
    for (int w = 0; w < 8; w++) {                                        // offsets inside word w of the planes: rl[w .. w+2] are registers
        for (; (i1 >> 5) == (uint32_t)w && i1 + K < N2 + 1 && !done; i1 += p.k2) {
            const uint32_t s = i1 & 31u;
            const uint64_t lo = (uint64_t)__funnelshift_r(rl[w], rl[w + 1], s) | ((uint64_t)__funnelshift_r(rl[w + 1], rl[w + 2], s) << 32);
            const uint64_t hi = (uint64_t)__funnelshift_r(rh[w], rh[w + 1], s) | ((uint64_t)__funnelshift_r(rh[w + 1], rh[w + 2], s) << 32);
            const uint64_t nn = is_n ? (uint64_t)__funnelshift_r(rn[w], rn[w + 1], s) | ((uint64_t)__funnelshift_r(rn[w + 1], rn[w + 2], s) << 32) : 0ull;
            const uint32_t h = cm_hash(K, lo, hi, nn, p.hash_mask, lut);
            // (the compact directory of cm_bucket was measured here too: 98 - 111 ms instead of 70 ms per config-2 step — this
            // kernel is bound by its 135 candidate verifications per read, not by the directory lines; it serves stage 7 only)
            const uint32_t b0 = __ldg(p.cumm + h);
            uint32_t b1 = __ldg(p.cumm + h + 1);
            if (b0 == b1) continue;
            if (limit < cur_false && b1 > b0 + PGM_CM_TRUNCATED_BUCKET) b1 = b0 + PGM_CM_TRUNCATED_BUCKET;   // :505-509
            for (uint32_t j = b0; j < b1; j++) {
                const uint64_t sp = (uint64_t)__ldg(p.entries + j) * p.k1;
                if (i1 > sp) continue;                                  // :512
                const uint64_t a = sp - i1;
                if (a + N2 > p.pg_len) continue;                        // :514
                n_ver++;
                // mismatches in the first trim8 characters (compared 8 at a time in the reference, exit between blocks) and in the tail
                uint32_t d_blocks = 0, d_tail = 0;
                const uint64_t tw = a >> 5;
                const uint32_t ts = (uint32_t)(a & 31);
                uint32_t la = __ldg(p.tlo + tw), ha = __ldg(p.thi + tw);
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    // (once the count inside the blocks exceeds the limit the candidate is rejected whatever follows: the rest
                    // of its text window is not fetched — most candidates are chance hits of the sparsified hash)
                    if ((uint32_t)g < W && d_blocks <= max_mm) {
                        const uint32_t lb = __ldg(p.tlo + tw + g + 1), hb = __ldg(p.thi + tw + g + 1);
                        uint32_t diff = (rl[g] ^ __funnelshift_r(la, lb, ts)) | (rh[g] ^ __funnelshift_r(ha, hb, ts)) | rn[g];
                        la = lb; ha = hb;
                        const uint32_t base = 32 * g;
                        if (N2 - base < 32) diff &= (1u << (N2 - base)) - 1u;
                        uint32_t in_blocks = 0xFFFFFFFFu;              // bits of this group below trim8
                        if (trim8 <= base) in_blocks = 0;
                        else if (trim8 - base < 32) in_blocks = (1u << (trim8 - base)) - 1u;
                        d_blocks += __popc(diff & in_blocks);
                        d_tail += __popc(diff & ~in_blocks);
                    }
                }
                if (d_blocks > max_mm) { cur_false++; continue; }       // :528-531
                if (d_blocks + d_tail > max_mm) { cur_false += 2; continue; }   // :532-543: counted in the tail loop and once more after it
                cur = d_blocks + d_tail;                                 // :546-548
                match_pos = a;
                found = true;
                if (cur <= p.min_mm) { done = true; break; }             // :549-552
                max_mm = cur - 1u;                                       // :553
            }
        }
    }
    if (found && cur < mm0) {                                            // ReadsMatchers.cpp:437-446
        const uint64_t rep = p.rev_mode ? p.pg_len - (match_pos + N2) : match_pos;
        const unsigned long long ns = ((unsigned long long)cur << 56) | ((unsigned long long)(p.rev_mode ? 1 : 0) << 55) | rep;
        rec[0] = make_uint4((uint32_t)ns, (uint32_t)(ns >> 32), hd.z, hd.w);
    }
    // work counters (diagnostic): [0] candidates [1] verified
    if (n_ver) { atomicAdd(counters + 0, n_ver); atomicAdd(counters + 1, n_ver); }
}

} // namespace pgm
