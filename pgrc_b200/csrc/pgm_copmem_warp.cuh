// pgm_copmem_warp.cuh — matching mode 'c', the per-read query as TWO kernels (opt-in: PGM_CM_WARP=1; the default is the
// thread-per-read cm_query_kernel of pgm_copmem.cuh).
//
// cm_query_kernel replays processApproxMatchQueryTight (CopMEMMatcher.cpp:483-566) with one thread per read: about 62 read
// offsets x 2.2 bucket entries = 135 candidate verifications per read at config 2, strictly one after the other, and a warp
// whose 32 reads are at different points of that chain runs with 11.7 of 32 lanes active (profiles/cm_query_full_r01y.json).
// The staged form (the test oracle carries a CPU model of it, copmem_query_staged, equal to its sequential transcription of the
// reference incl. the log-only counters):
//   cmw_stage1_kernel  a WARP per read.  Lane <-> read offset: hash, bucket bounds (compact L2 directory, cm_bucket), then the
//                      bucket entries of all 32 offsets side by side; an entry becomes its alignment start a = sp - i1 or
//                      "skipped" (:512, :514).  The DISTINCT alignments of a read (the true ones come back at most offsets) are
//                      verified ONCE each, by one lane, with both mismatch counts in full, and kept in a
//                      32-entry table; every candidate is written out as a one-byte index into that table.
//   cmw_stage2_kernel  a thread per read replays the sequential loop over those bytes: bucket truncation by the false-match
//                      budget (:505-509), +1 / +2 false matches (:528-543), strict improvement, stop at min_mm — 135 table
//                      look-ups instead of 135 text windows.
//   cm_query_kernel    once more, for the reads with more than 32 distinct alignments only (CopmemParams::only_marked).
// Exact because a verification's outcome under ANY limit follows from the two full counts (blocks > limit: +1 false match;
// else blocks + tail > limit: +2; else accept).
// STATUS: opt-in, bit-exact (tests/test_gpu_parity.py::test_copmem_staged_query_forced, config 2 at full size against the
// sampled oracle), and SLOWER than the default as built: 165 ms per config-2 step against 77 (profiles/bench_c2_mode_c_warp_r02am.json).
// Two measured reasons.  (1) A looked-up bucket holds 0.84 unrelated text positions on average (28 M samples in 2^25 hash
// values) next to the true ones, so a config-2 read has about 55 DISTINCT alignments, not 3 - 4: the 32-entry table overflows
// for most reads and the thread-per-read kernel redoes them (68 ms).  (2) Stage 1 walks the buckets entry by entry, one
// dependent load and several warp-wide exchanges per step: 89 ms although it stops at the overflow.  What it needs next: a
// 128-entry hashed table, the bucket entries of a batch loaded in bulk before they are looked at, and early rejection of the
// chance hits (one 32-base group decides almost all of them) before the full counts.
#pragma once
#include "pgm_copmem.cuh"

#define PGM_CMW_VT 32                 // distinct alignments of a read kept in its table
#define PGM_CMW_WARPS 8               // reads per CTA of stage 1
#define PGM_CMW_SKIP 0xFEu            // n_vt marker: the read needs no query (already matched within min_mm)
#define PGM_CMW_FALLBACK 0xFFu        // n_vt marker: more than PGM_CMW_VT distinct alignments — the thread-per-read kernel takes it

namespace pgm {

struct CmwParams {
    CopmemParams c;
    uint32_t n_off;                   // read offsets 0, k2, 2 k2, ... <= N2 - K
    uint32_t cap;                     // candidate bytes per read = n_off * 13
    uint32_t r_begin, r_count;        // this batch of reads
    uint8_t *lens;                    // [r_count * n_off] bucket size per offset
    uint8_t *cand;                    // [r_count * cap] table index per candidate, 0xFF = skipped
    unsigned long long *vt;           // [r_count * PGM_CMW_VT] a (40 bits) | blocks << 40 | tail << 48
    uint8_t *n_vt;                    // [r_count] entries of the table, or a marker
};

__global__ void __launch_bounds__(PGM_CMW_WARPS * 32, 4) cmw_stage1_kernel(const __grid_constant__ CmwParams q) {
    __shared__ uint32_t lut[256];
    __shared__ uint32_t s_rl[PGM_CMW_WARPS][10], s_rh[PGM_CMW_WARPS][10], s_rn[PGM_CMW_WARPS][10];
    __shared__ unsigned long long s_vt[PGM_CMW_WARPS][PGM_CMW_VT];
    cm_build_lut(lut);
    const CopmemParams &p = q.c;
    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    const uint32_t ri = blockIdx.x * PGM_CMW_WARPS + w;
    if (ri >= q.r_count) return;                                        // (whole warps leave; no block barrier below)
    const uint32_t r = q.r_begin + ri;
    const uint64_t pol_keep = policy_evict_last();
    uint32_t stride16; bool is_n;
    const uint4 *rec = record_of(p.reads, r, stride16, is_n);
    const uint4 hd = __ldcg(rec);
    const uint32_t mm0 = hd.y >> 24;
    if (mm0 <= p.min_mm) { if (lane == 0) q.n_vt[ri] = (uint8_t)PGM_CMW_SKIP; return; }     // ReadsMatchers.cpp:429
    const uint32_t N2 = p.reads.read_len, W = p.reads.W, K = p.K;
    if (lane < 10) {
        uint32_t l = 0, h = 0, n = 0;
        if (lane < 8 && lane < W) {
            if (is_n) { const uint4 v = __ldcg(rec + 1 + lane); l = v.x; h = v.y; n = v.z; }
            else { const uint4 v = __ldcg(rec + 1 + (lane >> 1)); l = (lane & 1) ? v.z : v.x; h = (lane & 1) ? v.w : v.y; }
        }
        s_rl[w][lane] = l; s_rh[w][lane] = h; s_rn[w][lane] = n;
    }
    __syncwarp();
    const uint32_t trim8 = (N2 >> 3) << 3;
    uint32_t n_vt = 0, cand_base = 0;
    bool overflow = false;
    uint8_t *lens = q.lens + (size_t)ri * q.n_off, *cand = q.cand + (size_t)ri * q.cap;
    for (uint32_t ob = 0; ob < q.n_off && !overflow; ob += 32) {
        const uint32_t o = ob + lane;
        const bool active = o < q.n_off;
        const uint32_t i1 = o * p.k2;
        uint32_t b0 = 0, len = 0;
        if (active) {
            const uint32_t wi = i1 >> 5, s = i1 & 31u;
            const uint64_t lo = (uint64_t)__funnelshift_r(s_rl[w][wi], s_rl[w][wi + 1], s) | ((uint64_t)__funnelshift_r(s_rl[w][wi + 1], s_rl[w][wi + 2], s) << 32);
            const uint64_t hi = (uint64_t)__funnelshift_r(s_rh[w][wi], s_rh[w][wi + 1], s) | ((uint64_t)__funnelshift_r(s_rh[w][wi + 1], s_rh[w][wi + 2], s) << 32);
            const uint64_t nn = is_n ? (uint64_t)__funnelshift_r(s_rn[w][wi], s_rn[w][wi + 1], s) | ((uint64_t)__funnelshift_r(s_rn[w][wi + 1], s_rn[w][wi + 2], s) << 32) : 0ull;
            const uint32_t h = cm_hash(K, lo, hi, nn, p.hash_mask, lut);
            uint32_t b1;
            cm_bucket(p.nib, p.coarse, h, pol_keep, b0, b1);
            len = b1 - b0;
            lens[o] = (uint8_t)len;
        }
        // exclusive prefix of the bucket sizes over the lanes, their sum and maximum
        uint32_t incl = len, mx = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(PGM_FULL, incl, d);
            if (lane >= (uint32_t)d) incl += y;
            mx = max(mx, __shfl_xor_sync(PGM_FULL, mx, d));
        }
        const uint32_t start = incl - len, total = __shfl_sync(PGM_FULL, incl, 31);
        for (uint32_t t = 0; t < mx && !overflow; t++) {
            const bool has = t < len;
            unsigned long long a = ~0ull;
            if (has) {
                const uint64_t sp = (uint64_t)__ldg(p.entries + b0 + t) * p.k1;
                if (i1 <= sp && sp - i1 + N2 <= p.pg_len) a = sp - i1;                 // :512, :514
            }
            const bool valid = a != ~0ull;
            uint32_t slot = 0xFFu;
            if (valid)
                for (uint32_t k = 0; k < n_vt; k++)
                    if ((s_vt[w][k] & PGM_POS_MASK) == a) { slot = k; break; }
            const bool miss = valid && slot == 0xFFu;
            const uint32_t mmask = __ballot_sync(PGM_FULL, miss);
            if (mmask) {                                                               // (warp-uniform)
                uint32_t peers = 0;
                if (miss) peers = __match_any_sync(mmask, a);
                const uint32_t lead_lane = miss ? (uint32_t)__ffs((int)peers) - 1u : 0u;
                const bool leader = miss && lead_lane == lane;
                const uint32_t lmask = __ballot_sync(PGM_FULL, leader);
                const uint32_t n_new = (uint32_t)__popc(lmask);
                if (n_vt + n_new > PGM_CMW_VT) { overflow = true; }
                else {
                    uint32_t my_slot = 0;
                    if (leader) {
                        my_slot = n_vt + (uint32_t)__popc(lmask & lt_mask);
                        // both counts in full: mismatches in the first trim8 characters and in the tail (:516-543)
                        uint32_t d_blocks = 0, d_tail = 0;
                        const uint64_t tw = a >> 5;
                        const uint32_t ts = (uint32_t)(a & 31);
                        uint32_t la = __ldg(p.tlo + tw), ha = __ldg(p.thi + tw);
                        for (uint32_t g = 0; g < W; g++) {
                            const uint32_t lb = __ldg(p.tlo + tw + g + 1), hb = __ldg(p.thi + tw + g + 1);
                            uint32_t diff = (s_rl[w][g] ^ __funnelshift_r(la, lb, ts)) | (s_rh[w][g] ^ __funnelshift_r(ha, hb, ts)) | s_rn[w][g];
                            la = lb; ha = hb;
                            const uint32_t base = 32 * g;
                            if (N2 - base < 32) diff &= (1u << (N2 - base)) - 1u;
                            uint32_t in_blocks = 0xFFFFFFFFu;
                            if (trim8 <= base) in_blocks = 0;
                            else if (trim8 - base < 32) in_blocks = (1u << (trim8 - base)) - 1u;
                            d_blocks += __popc(diff & in_blocks);
                            d_tail += __popc(diff & ~in_blocks);
                        }
                        s_vt[w][my_slot] = a | ((unsigned long long)d_blocks << 40) | ((unsigned long long)d_tail << 48);
                    }
                    __syncwarp();
                    if (miss) slot = __shfl_sync(mmask, my_slot, lead_lane);
                    n_vt += n_new;
                }
            }
            if (has && !overflow) cand[cand_base + start + t] = (uint8_t)(valid ? slot : 0xFFu);
        }
        cand_base += total;
    }
    __syncwarp();
    if (lane == 0) q.n_vt[ri] = (uint8_t)(overflow ? PGM_CMW_FALLBACK : n_vt);
    if (!overflow && lane < n_vt) q.vt[(size_t)ri * PGM_CMW_VT + lane] = s_vt[w][lane];
}

__global__ void __launch_bounds__(PGM_CM_THREADS) cmw_stage2_kernel(const __grid_constant__ CmwParams q, unsigned long long *counters) {
    const CopmemParams &p = q.c;
    const uint32_t ri = blockIdx.x * PGM_CM_THREADS + threadIdx.x;
    if (ri >= q.r_count) return;
    const uint32_t nv = q.n_vt[ri];
    if (nv == PGM_CMW_SKIP || nv == PGM_CMW_FALLBACK) return;
    const uint32_t r = q.r_begin + ri;
    uint32_t stride16; bool is_n;
    uint4 *rec = record_of(p.reads, r, stride16, is_n);
    const uint4 hd = __ldcg(rec);
    const unsigned long long st = ((unsigned long long)hd.y << 32) | hd.x;
    const uint32_t mm0 = (uint32_t)(st >> 56);
    const uint32_t N2 = p.reads.read_len;
    uint32_t max_mm = p.max_mm;
    if (mm0 < max_mm) max_mm = mm0 - 1u;                                 // :488-489
    const unsigned long long limit = (unsigned long long)((N2 + 1 - p.K) / p.k2);
    unsigned long long cur_false = 0, n_ver = 0;
    uint64_t match_pos = 0;
    uint32_t cur = mm0, base = 0;
    bool found = false, done = false;
    const uint8_t *lens = q.lens + (size_t)ri * q.n_off, *cand = q.cand + (size_t)ri * q.cap;
    const unsigned long long *vt = q.vt + (size_t)ri * PGM_CMW_VT;
    for (uint32_t o = 0; o < q.n_off && !done; o++) {
        const uint32_t len = lens[o];
        uint32_t lim = len;
        if (limit < cur_false && len > PGM_CM_TRUNCATED_BUCKET) lim = PGM_CM_TRUNCATED_BUCKET;     // :505-509
        for (uint32_t t = 0; t < lim; t++) {
            const uint32_t v = cand[base + t];
            if (v == 0xFFu) continue;                                    // :512, :514
            n_ver++;
            const unsigned long long e = vt[v];
            const uint32_t d_blocks = (uint32_t)(e >> 40) & 0xFFu, d_tail = (uint32_t)(e >> 48) & 0xFFu;
            if (d_blocks > max_mm) { cur_false++; continue; }            // :528-531
            if (d_blocks + d_tail > max_mm) { cur_false += 2; continue; }   // :532-543
            cur = d_blocks + d_tail;                                     // :546-548
            match_pos = e & PGM_POS_MASK;
            found = true;
            if (cur <= p.min_mm) { done = true; break; }                 // :549-552
            max_mm = cur - 1u;                                           // :553
        }
        base += len;
    }
    if (found && cur < mm0) {                                            // ReadsMatchers.cpp:437-446
        const uint64_t rep = p.rev_mode ? p.pg_len - (match_pos + N2) : match_pos;
        const unsigned long long ns = ((unsigned long long)cur << 56) | ((unsigned long long)(p.rev_mode ? 1 : 0) << 55) | rep;
        rec[0] = make_uint4((uint32_t)ns, (uint32_t)(ns >> 32), hd.z, hd.w);
    }
    if (n_ver) { atomicAdd(counters + 0, n_ver); atomicAdd(counters + 1, n_ver); }
}

} // namespace pgm
