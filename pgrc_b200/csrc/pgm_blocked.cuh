// pgm_blocked.cuh — stages 2 and 3 of the L2-blocked scan pipeline (stage 1 is scan_kernel<.., MODE 1> in
// pgm_kernels.cuh).
//
// Why.  The fused scan kernel probes the seed table and fetches read records in text order, i.e. at random
// addresses of structures far larger than the L2: every probe and every verification moves a 128-byte DRAM line
// for 32 / 64 useful bytes, and the pass runs at the DRAM's random-transaction rate (profiles/kernels_metrics_c2_r01h.csv:
// 8.06 GB of DRAM traffic for 1.55 GB of algorithmic bytes).  The pipeline re-orders the work twice instead, with
// streaming queues, so that each random access hits the L2:
//   stage 1  text order   hash + filter every window; positives -> queue of their TABLE REGION (a range of home buckets)
//   stage 2  region order all CTAs work on one L2-sized piece of the bucket array at a time; tag hits -> queue of their
//                         READ RANGE (a range of read indices whose records fit the L2)
//   stage 3  range order  XOR/popcount verification: records of one range (L2), text windows gathered from the 2-bit
//                         planes (L2-resident for texts up to a few hundred Mbases), accept test, atomicMin on the key
//                         inside the record (L2).
// The decision logic of stage 3 is the body of DefaultReadsApproxMatcher::executeMatching (ReadsMatchers.cpp:301-331),
// identical to the fused kernel's; only the order in which events arrive differs, and the per-read accumulators are
// order-free.
#pragma once
#include "pgm_kernels.cuh"

#define PGM_PROBE_THREADS 256
#define PGM_PROBE_CHUNK 1024               // queue entries a CTA takes at a time
#define PGM_PROBE_STAGE 2560               // tag hits of one chunk staged in shared memory before they are queued
#define PGM_VERIFY_THREADS 256
#define PGM_VERIFY_CHUNK 2048

namespace pgm {

struct VerifyParams {
    const uint32_t *tlo, *thi;      // planes of this pass's text, local origin at word 0
    uint64_t pos_origin;            // global coordinate (this pass's) of queue position 0
    uint64_t bit_origin;            // local bit offset of queue position 0 in the planes
    uint64_t pg_len;
    uint32_t seed_len, parts, max_mm, min_mm;
    int rev_mode;
    TableView tab;
    ReadsView reads;
    PerRead pr;
    StageQueues sq;
};

// ------------------------------------------------------------------------------------------ stage 2: probe
// Region after region (CTAs move on together, like build_insert_kernel): a CTA takes a chunk of the region's queue,
// every thread probes its entries (one 256-bit load per bucket, default caching: the region is meant to stay in the
// L2), tag hits are staged in shared memory with their rank within their read range, then one global reservation per
// range and chunk, and the hits are written to the range queues.
__global__ void __launch_bounds__(PGM_PROBE_THREADS) probe_kernel(const __grid_constant__ VerifyParams p) {
    __shared__ uint2 s_cand[PGM_PROBE_STAGE];
    __shared__ uint16_t s_rank[PGM_PROBE_STAGE];
    __shared__ unsigned int s_cnt[PGM_SQ_MAX], s_base[PGM_SQ_MAX];
    __shared__ unsigned int s_n, s_first[2];
    const uint32_t t = threadIdx.x;
    const StageQueues &q = p.sq;
    const uint32_t n_regions = 1u << q.region_bits;
    const uint32_t read_shift = p.reads.part_bits + q.range_shift;
    for (uint32_t k = t; k < PGM_SQ_MAX; k += PGM_PROBE_THREADS) s_cnt[k] = 0;
    if (t == 0) s_n = 0;
    bool over = false;
    uint32_t flip = 0;
    for (uint32_t k = 0; k < n_regions; k++) {
        const uint32_t n = min(__ldg(q.pos_count + k), q.pos_cap);
        const uint4 *src = q.pos_entries + (size_t)k * q.pos_cap;
        for (;;) {
            if (t == 0) s_first[flip] = atomicAdd(q.pos_cursor + k, (unsigned int)PGM_PROBE_CHUNK);
            __syncthreads();
            const uint32_t first = s_first[flip];
            flip ^= 1;
            if (first >= n) break;
            const uint32_t last = min(first + PGM_PROBE_CHUNK, n);
            for (uint32_t i = first + t; i < last; i += PGM_PROBE_THREADS) {
                const uint4 e = __ldcs(src + i);
                const uint32_t tag = seed_tag(e.y);
                uint32_t b = __umulhi(e.x, p.tab.n_buckets);
                const uint32_t step = 1u + __umulhi(e.y * 0x9E3779B1u, p.tab.n_buckets - 1u);
                for (;;) {
                    const u32x8 s = ld256_cg(p.tab.buckets + b);
                    bool em = false;
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        em |= s.v[2 * c + 1] == 0xFFFFFFFFu;
                        if ((s.v[2 * c + 1] & 0x7FFFFFFFu) == tag) {
                            const uint32_t at = atomicAdd(&s_n, 1u);
                            if (at < PGM_PROBE_STAGE) {
                                const uint32_t pat = s.v[2 * c];
                                s_cand[at] = make_uint2(e.z | (s.v[2 * c + 1] & 0x80000000u), pat);
                                s_rank[at] = (uint16_t)atomicAdd(&s_cnt[pat >> read_shift], 1u);
                            } else over = true;
                        }
                    }
                    if (em) break;                       // a bucket with an empty slot ends the probe sequence
                    b += step;
                    if (b >= p.tab.n_buckets) b -= p.tab.n_buckets;
                }
            }
            __syncthreads();
            const uint32_t staged = min(s_n, (unsigned int)PGM_PROBE_STAGE);
            for (uint32_t r = t; r < q.n_ranges; r += PGM_PROBE_THREADS) {
                const unsigned int c = s_cnt[r];
                s_base[r] = c ? atomicAdd(q.cand_count + r, c) : 0u;
                s_cnt[r] = 0;
            }
            __syncthreads();
            if (t == 0) s_n = 0;                         // (the next chunk's staging starts behind the next barrier)
            for (uint32_t i = t; i < staged; i += PGM_PROBE_THREADS) {
                const uint2 c = s_cand[i];
                const uint32_t r = c.y >> read_shift, idx = s_base[r] + s_rank[i];
                if (idx < q.cand_cap) __stcs(q.cand_entries + (size_t)r * q.cand_cap + idx, c);
                else over = true;
            }
        }
    }
    if (over) *q.overflow = 1u;
}

// ------------------------------------------------------------------------------------------ stage 3: verify
// W + 1 (<= 7) consecutive words of a text plane starting at word index wi (may reach into the zero pad in front of
// the plane): two aligned 256-bit loads and a 3-stage word barrel shifter (static register indices only).
__device__ __forceinline__ void gather7(const uint32_t *plane, int64_t wi, uint32_t (&y)[7]) {
    const int64_t w0 = wi & ~(int64_t)7;
    const uint32_t o = (uint32_t)(wi - w0);
    const u32x8 a = ld256_cg(plane + w0), b = ld256_cg(plane + w0 + 8);
    uint32_t z[10], u[8];
#pragma unroll
    for (int k = 0; k < 10; k++) {
        const uint32_t x0 = k < 8 ? a.v[k] : b.v[k - 8];
        const uint32_t x4 = k + 4 < 8 ? a.v[k + 4] : b.v[k + 4 - 8];
        z[k] = (o & 4u) ? x4 : x0;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) u[k] = (o & 2u) ? z[k + 2] : z[k];
#pragma unroll
    for (int k = 0; k < 7; k++) y[k] = (o & 1u) ? u[k + 1] : u[k];
}

// Mismatches of a whole read record against the text at local bit offset lbit, one lane on its own with scalar loads
// (ACGNT records, records longer than 64 bytes, chain walks): the rare path.
__device__ __forceinline__ int count_alone(const VerifyParams &p, const uint4 *rec, bool is_n, int64_t lbit) {
    const int64_t wi = lbit >> 5;
    const uint32_t ts = (uint32_t)(lbit & 31);
    const uint32_t W = p.reads.W, L = p.reads.read_len;
    int c = 0;
    uint32_t la = __ldg(p.tlo + wi), ha = __ldg(p.thi + wi);
    for (uint32_t g = 0; g < W; g++) {
        const uint32_t lb = __ldg(p.tlo + wi + g + 1), hb = __ldg(p.thi + wi + g + 1);
        const uint32_t tl = __funnelshift_r(la, lb, ts), th = __funnelshift_r(ha, hb, ts);
        la = lb; ha = hb;
        uint32_t rl, rh, nm = 0;
        if (is_n) {
            const uint4 x = __ldcg(rec + 1 + g);
            rl = x.x; rh = x.y; nm = x.z;                   // an N never equals a text symbol
        } else {
            const uint4 x = __ldcg(rec + 1 + (g >> 1));
            rl = (g & 1) ? x.z : x.x; rh = (g & 1) ? x.w : x.y;
        }
        uint32_t diff = (rl ^ tl) | (rh ^ th) | nm;
        const uint32_t rem = L - 32 * g;
        if (rem < 32) diff &= (1u << rem) - 1u;
        c += __popc(diff);
    }
    return c;
}

// Body of DefaultReadsApproxMatcher::executeMatching (ReadsMatchers.cpp:301-331) up to the decision, for one event:
// read cr, seed cj, hit at global window start gpos, c mismatches over the whole read, {st_lo, st_hi} the read's stored
// state and `seen` a value of its key that is not below the live one.  Same statements as in scan_kernel.
__device__ __forceinline__ void apply_event(const VerifyParams &p, uint32_t cr, uint32_t cj, int c, uint32_t st_lo, uint32_t st_hi,
                                            long long seen, uint64_t gpos, long long *keyp, unsigned long long &n_ver,
                                            unsigned long long &n_acc) {
    const uint32_t L = p.reads.read_len;
    const uint32_t shift = cj * p.seed_len;
    const uint32_t c_in = st_hi >> 24;
    if (c_in > p.min_mm && (uint64_t)shift <= gpos) {                              // :304, :308
        const uint64_t a = gpos - shift;
        if (a + L <= p.pg_len) {                                                   // :311
            n_ver++;
            const bool has_pos = c_in != 255u;
            const int limit = has_pos ? (int)c_in - 1 : (int)p.max_mm;             // :315
            if (c <= limit) {
                n_acc++;
                const uint64_t rep = p.rev_mode ? p.pg_len - (a + L) : a;         // :313,:326 (matchingLength == readLength)
                const uint64_t st_pos = (((uint64_t)st_hi << 32) | st_lo) & PGM_POS_MASK;
                const unsigned long long order = (gpos << 8) | (unsigned long long)(p.parts - 1 - cj);
                if (!(has_pos && st_pos == rep)) {                                 // coordinate-only compare, :313
                    const unsigned long long cls = (uint32_t)c <= p.min_mm ? 0ull : (unsigned long long)c;
                    const long long key = (long long)((cls << 56) | (order << 8) | (unsigned long long)c);
                    if (key < seen) atomicMin(keyp, key);
                    if (has_pos) {
                        atomicMin(p.pr.first_other_order + cr, (long long)order);
                        *p.pr.touched = 1;
                    }
                } else {
                    atomicOr(p.pr.same_pos_mask + cr, 1 << cj);
                    p.pr.same_pos_mm[cr] = (uint8_t)c;
                    *p.pr.touched = 1;
                }
            }
        }
    }
}

// Range after range: a CTA takes a chunk of the range's queue, every lane owns one candidate.  LQ64: the ACGT set's
// records are exactly 64 bytes (read length <= 192) — the two lanes of a pair fetch each other's record halves with
// one 256-bit load each (one instruction = one 64-byte request, as in the fused kernel) and the text window comes
// through gather7; every other candidate (ACGNT read, longer record, chain element) is verified by its lane alone.
template <bool LQ64>
__global__ void __launch_bounds__(PGM_VERIFY_THREADS) verify_kernel(const __grid_constant__ VerifyParams p) {
    __shared__ unsigned int s_first[2];
    const uint32_t t = threadIdx.x, lane = t & 31u, half = lane & 1u;
    const StageQueues &q = p.sq;
    const uint32_t pmask = (1u << p.reads.part_bits) - 1u;
    unsigned long long n_cand = 0, n_ver = 0, n_acc = 0;
    uint32_t flip = 0;
    for (uint32_t k = 0; k < q.n_ranges; k++) {
        const uint32_t n = min(__ldg(q.cand_count + k), q.cand_cap);
        const uint2 *src = q.cand_entries + (size_t)k * q.cand_cap;
        for (;;) {
            if (t == 0) s_first[flip] = atomicAdd(q.cand_cursor + k, (unsigned int)PGM_VERIFY_CHUNK);
            __syncthreads();
            const uint32_t first = s_first[flip];
            flip ^= 1;
            if (first >= n) break;
            const uint32_t last = min(first + PGM_VERIFY_CHUNK, n);
            for (uint32_t i0 = first; i0 < last; i0 += PGM_VERIFY_THREADS) {        // same trip count for the whole warp (shuffles inside)
                const uint32_t i = i0 + t;
                const bool on = i < last;
                const uint2 e = on ? __ldcs(src + i) : make_uint2(0, 0);
                const uint32_t cpos = e.x & 0x7FFFFFFFu;
                const bool chain = (e.x >> 31) != 0;
                uint32_t cpat = e.y;
                uint32_t cr = cpat >> p.reads.part_bits, cj = cpat & pmask;
                const uint64_t gpos = p.pos_origin + cpos;
                bool fast = false;
                if (LQ64) {
                    fast = on && cr < p.reads.n_lq;
                    const uint32_t cr_o = __shfl_xor_sync(PGM_FULL, cr, 1);
                    const bool fast_o = __shfl_xor_sync(PGM_FULL, (int)fast, 1) != 0;
                    const uint32_t rA = half ? cr_o : cr, rB = half ? cr : cr_o;
                    const bool onA = half ? fast_o : fast, onB = half ? fast : fast_o;
                    u32x8 xA, xB;
#pragma unroll
                    for (int w = 0; w < 8; w++) { xA.v[w] = 0; xB.v[w] = 0; }
                    if (onA) xA = ld256_cg(reinterpret_cast<const u32x8 *>(p.reads.lq) + (size_t)rA * 2 + half);
                    if (onB) xB = ld256_cg(reinterpret_cast<const u32x8 *>(p.reads.lq) + (size_t)rB * 2 + half);
                    uint32_t s0[8], s1[8];
#pragma unroll
                    for (int w = 0; w < 8; w++) {
                        const uint32_t got = __shfl_xor_sync(PGM_FULL, half ? xA.v[w] : xB.v[w], 1);
                        s0[w] = half ? got : xA.v[w];      // first 32 bytes of this lane's record: header, groups 0-1
                        s1[w] = half ? xB.v[w] : got;      // second 32 bytes: groups 2-5
                    }
                    if (fast) {
                        n_cand++;
                        const int64_t lbit = (int64_t)(p.bit_origin + cpos) - (int64_t)(cj * p.seed_len);
                        const uint32_t ts = (uint32_t)(lbit & 31);
                        const uint32_t L = p.reads.read_len;
                        uint32_t yl[7], yh[7];
                        gather7(p.tlo, lbit >> 5, yl);
                        gather7(p.thi, lbit >> 5, yh);
                        int c = 0;
#pragma unroll
                        for (int g = 0; g < 6; g++) {
                            if ((uint32_t)g < p.reads.W) {
                                const uint32_t rl = g < 2 ? s0[4 + 2 * g] : s1[2 * (g - 2)], rh = g < 2 ? s0[5 + 2 * g] : s1[2 * (g - 2) + 1];
                                const uint32_t tl = __funnelshift_r(yl[g], yl[g + 1], ts), th = __funnelshift_r(yh[g], yh[g + 1], ts);
                                uint32_t diff = (rl ^ tl) | (rh ^ th);
                                const uint32_t rem = L - 32 * g;
                                if (rem < 32) diff &= (1u << rem) - 1u;
                                c += __popc(diff);
                            }
                        }
                        apply_event(p, cr, cj, c, s0[0], s0[1], (long long)(((uint64_t)s0[3] << 32) | s0[2]), gpos,
                                    reinterpret_cast<long long *>(p.reads.lq + (size_t)cr * 4) + 1, n_ver, n_acc);
                    }
                }
                // everything else, and the chains of hot keys behind a slot: this lane alone
                bool more = on && (!fast || chain);
                bool first_el = !fast;
                while (more) {
                    if (!first_el) {
                        cpat = __ldg(p.tab.next + cpat);
                        if (cpat == PGM_NIL) break;
                        cr = cpat >> p.reads.part_bits; cj = cpat & pmask;
                    }
                    first_el = false;
                    n_cand++;
                    uint32_t stride16; bool is_n;
                    uint4 *rec = record_of(p.reads, cr, stride16, is_n);
                    const uint4 h = __ldcg(rec);
                    const uint32_t shift = cj * p.seed_len;
                    // (an alignment that would start in front of the text is dropped by apply_event, ReadsMatchers.cpp:308; the
                    // planes have PGM_PAD_WORDS zero words in front, so counting it is harmless)
                    const int c = count_alone(p, rec, is_n, (int64_t)(p.bit_origin + cpos) - (int64_t)shift);
                    apply_event(p, cr, cj, c, h.x, h.y, (long long)(((uint64_t)h.w << 32) | h.z), gpos,
                                reinterpret_cast<long long *>(rec) + 1, n_ver, n_acc);
                    more = chain;
                }
            }
        }
    }
    unsigned long long cv[3] = {n_cand, n_ver, n_acc};
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cv[k] += __shfl_xor_sync(PGM_FULL, cv[k], o);
        if (lane == 0 && cv[k]) atomicAdd(q.counters + k, cv[k]);
    }
}

} // namespace pgm
