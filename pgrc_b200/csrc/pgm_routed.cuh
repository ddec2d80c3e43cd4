// pgm_routed.cuh — kernels of the ROUTED multi-GPU scheme: every stage of a matcher call divides by
// the number of GPUs.
//
//   reads      GPU g owns the read range [read_begin[g], read_begin[g+1]): records, per-read keys, decision, results.
//   seed table hash-partitioned: the pattern with hash h lives on GPU umulhi(h1, world) (its own L2-resident filter).
//   text       the 2-bit planes (tiny: pg_len / 4 bytes per strand) are on every GPU; the WINDOWS are range-partitioned:
//              GPU g hashes the window starts of its range only.
//
//   build   route_build (build_table_kernel with BuildQueues::route_world): seeds of my reads -> {h1', h2, pattern}
//           into the send segment of the hash owner;  exchange;  route_insert_kernel: received patterns -> my table.
//   pass    route_scan_kernel: every window of my range -> {h1', h2, position} into the send segment of the hash owner
//           exchange #1 (all-to-all over NVLink)
//           route_probe_kernel: received windows -> my filter, my table -> candidates {position, pattern} into the send
//           segment of the READ owner
//           exchange #2
//           route_verify_kernel: received candidates -> XOR/popcount of my read record against the text window,
//           accept test, atomicMin on my key.   Then resolve_kernel on my reads: no cross-GPU merge of keys at all.
//
// Event order never matters: the accumulators are order-free, so the result is the single-GPU
// result bit for bit.  The decision logic is apply_event (pgm_blocked.cuh) = the body of
// DefaultReadsApproxMatcher::executeMatching (ReadsMatchers.cpp:301-331).
#pragma once
#include "pgm_kernels.cuh"
#include "pgm_blocked.cuh"

#define PGM_ROUTE_MAX_WORLD 16
#define PGM_ROUTE_THREADS 256
#define PGM_ROUTE_PROBE_STAGE 1024          // candidates of one chunk staged in shared memory (about 120 expected; hot keys overflow to the slow path)
#define PGM_ROUTE_PROBE_CTAS 8               // resident CTAs per SM: a chunk is one DRAM latency (the probe) + a flush; other CTAs fill the gaps

namespace pgm {

// filter word / bits of a routed pattern, from the 64-bit seed hash (the per-GPU filter only ever sees hashes)
__host__ __device__ __forceinline__ uint32_t route_filter_hash(uint32_t h1p, uint32_t h2) {
    uint32_t x = h1p ^ (h2 * 0x9E3779B1u);
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

struct RouteQueues {
    uint32_t *entries;              // destination d owns 3-word entries [d * cap, d * cap + count[d])
    unsigned int *count;            // PGM_ROUTE_MAX_WORLD
    unsigned int *overflow;
    uint32_t cap;                   // entries per destination
    uint32_t world;
};

// ------------------------------------------------------------------------------------------ build: insert
// Received patterns {h1', h2, pattern, -} -> filter bits + table (home bucket from h1' = the part of h1 the routing did not use).
__global__ void __launch_bounds__(PGM_INSERT_THREADS, 8) route_insert_kernel(TableView tab, const uint4 *__restrict__ src, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * PGM_INSERT_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * PGM_INSERT_THREADS) {
        const uint4 e = __ldcs(src + i);
        if (tab.filter) {
            const uint32_t f = route_filter_hash(e.x, e.y);
            atomicOr(tab.filter + (f & tab.filter_mask), filter_bits(f, tab.filter_k));
        }
        table_insert(tab, e.x, e.y, e.z);
    }
}

// ------------------------------------------------------------------------------------------ pass: scan (emit)
struct RouteScanParams {
    const uint32_t *tlo, *thi;      // planes of this pass's WHOLE text, origin at word 0
    uint64_t begin, end;            // window starts [begin, end) of this round (pass coordinates); begin is a multiple of 128
    uint32_t first_word;            // first word of tile 0 (= begin / 32)
    uint32_t n_tiles;
    uint32_t tail_mask;
    unsigned int *tile_counter;
    RouteQueues q;
};

struct RouteScanShared {
    uint32_t lo[2][PGM_BUF_WORDS];
    uint32_t hi[2][PGM_BUF_WORDS];
    uint32_t h1[PGM_TILE_POS], h2[PGM_TILE_POS];
    uint32_t rk[PGM_TILE_POS];      // destination << 16 | rank among the tile's entries for that destination; PGM_NIL = not a window of this round
    uint64_t bar[2];
    unsigned int cnt[PGM_ROUTE_MAX_WORLD], base[PGM_ROUTE_MAX_WORLD], tile[2];
};

// Persistent CTAs pull 4096-position tiles (TMA-staged planes, as in scan_kernel).  Every window start of the tile that
// lies in [begin, end) is hashed (canonical form -> 64-bit key); its destination is umulhi(h1, world).  Ranks among the
// tile's entries per destination come from warp match + one shared-memory atomic per (warp, destination); one global
// reservation per destination and tile; then the entries {h1 * world, h2, position - begin} go out.
template <int NCH>
__global__ void __launch_bounds__(PGM_ROUTE_THREADS, 4) route_scan_kernel(const __grid_constant__ RouteScanParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RouteScanShared &sm = *reinterpret_cast<RouteScanShared *>(smem_raw);
    const uint32_t t = threadIdx.x, lane = t & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t world = p.q.world;

    auto issue_tile = [&](unsigned int tile, int b) {
        const int64_t w0 = (int64_t)p.first_word + (int64_t)tile * PGM_TILE_WORDS - PGM_HALO_L;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&sm.bar[b], 2 * PGM_BUF_WORDS * 4);
        bulk_g2s(sm.lo[b], p.tlo + w0, PGM_BUF_WORDS * 4, &sm.bar[b]);
        bulk_g2s(sm.hi[b], p.thi + w0, PGM_BUF_WORDS * 4, &sm.bar[b]);
    };
    if (t == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned int tile = atomicAdd(p.tile_counter, 1u);
        sm.tile[0] = tile;
        if (tile < p.n_tiles) issue_tile(tile, 0);
    }
    if (t < PGM_ROUTE_MAX_WORLD) sm.cnt[t] = 0;
    __syncthreads();
    uint32_t parity[2] = {0, 0};
    int buf = 0;
    bool over = false;
    for (;;) {
        const unsigned int tile = sm.tile[buf];
        if (tile >= p.n_tiles) break;
        if (t == 0) {
            const unsigned int nxt = atomicAdd(p.tile_counter, 1u);
            sm.tile[buf ^ 1] = nxt;
            if (nxt < p.n_tiles) issue_tile(nxt, buf ^ 1);
        }
        mbar_wait(&sm.bar[buf], parity[buf]);
        parity[buf] ^= 1;
        const uint32_t *slo = sm.lo[buf] + PGM_HALO_L, *shi = sm.hi[buf] + PGM_HALO_L;
        const uint64_t tile_g0 = ((uint64_t)p.first_word + (uint64_t)tile * PGM_TILE_WORDS) * 32;
        const uint32_t vb = (uint32_t)(p.begin > tile_g0 ? min((uint64_t)PGM_TILE_POS, p.begin - tile_g0) : 0);
        const uint32_t ve = (uint32_t)(p.end > tile_g0 ? min((uint64_t)PGM_TILE_POS, p.end - tile_g0) : 0);
        // phase 1: hash, destination, rank
#pragma unroll 4
        for (uint32_t pos = t; pos < PGM_TILE_POS; pos += PGM_ROUTE_THREADS) {
            const uint64_t hv = window_hash<NCH>(slo, shi, pos, p.tail_mask);
            const uint32_t h1 = (uint32_t)hv, h2 = (uint32_t)(hv >> 32);
            const bool on = pos >= vb && pos < ve;
            const uint32_t dest = on ? __umulhi(h1, world) : 0xFFu;
            const uint32_t peers = __match_any_sync(PGM_FULL, dest);
            uint32_t wbase = 0;
            const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
            if (on && lane == leader) wbase = atomicAdd(&sm.cnt[dest], (unsigned int)__popc(peers));
            wbase = __shfl_sync(PGM_FULL, wbase, leader);
            sm.h1[pos] = h1 * world;                       // the fractional part: uniform again, picks the home bucket
            sm.h2[pos] = h2;
            sm.rk[pos] = on ? ((dest << 16) | (wbase + __popc(peers & lt_mask))) : PGM_NIL;
        }
        __syncthreads();
        if (t < world) {
            const unsigned int c = sm.cnt[t];
            sm.base[t] = c ? atomicAdd(p.q.count + t, c) : 0u;
            sm.cnt[t] = 0;
        }
        __syncthreads();
        // phase 2: write
        const uint64_t rel0 = tile_g0 - p.begin;           // (wraps for positions in front of `begin`: those are off)
        for (uint32_t pos = t; pos < PGM_TILE_POS; pos += PGM_ROUTE_THREADS) {
            const uint32_t r = sm.rk[pos];
            if (r != PGM_NIL) {
                const uint32_t dest = r >> 16, idx = sm.base[dest] + (r & 0xFFFFu);
                if (idx < p.q.cap) {
                    uint32_t *e = p.q.entries + ((size_t)dest * p.q.cap + idx) * 3;
                    e[0] = sm.h1[pos]; e[1] = sm.h2[pos]; e[2] = (uint32_t)(rel0 + pos);
                } else over = true;
            }
        }
        __syncthreads();
        buf ^= 1;
    }
    if (over) *p.q.overflow = 1u;
}

// ------------------------------------------------------------------------------------------ pass: probe
struct RouteProbeParams {
    const uint32_t *src;            // windows of ONE sender that passed the filter (route_filter_kernel), 3 words each
    uint64_t n;                     // their number, or
    const unsigned int *n_ptr;      // ... where the filter kernel counted them (device memory)
    uint64_t pos_base;              // pass coordinate of the sender's position 0
    TableView tab;
    uint32_t part_bits;
    uint32_t world;
    uint64_t read_begin[PGM_ROUTE_MAX_WORLD + 1];
    RouteQueues q;                  // candidate queues by read owner: {position lo, position hi, pattern}
    unsigned long long *counters;   // [3] filter positives
};

__device__ __forceinline__ uint32_t route_owner(const RouteProbeParams &p, uint32_t read) {
    uint32_t d = 0;
#pragma unroll 1
    for (uint32_t k = 1; k < p.world; k++) d += (uint64_t)read >= p.read_begin[k] ? 1u : 0u;
    return d;
}

// The probe of the received windows is two kernels with one kind of latency each (a single kernel with a load, a filter,
// a probe and a flush phase per chunk spent most of its time at barriers: 38 ms per step at the per-rank sizes of an 8-GPU
// config-5 run, and the same 38-44 ms whether 0.45 G or 1.2 G windows went on to the table):
//   route_filter_kernel  streams the received windows (12 bytes each, coalesced into shared memory), four L2-resident filter
//                        gathers in flight per thread, and writes the survivors — a fifth to a quarter of the windows once
//                        the table is spread over 8 GPUs — compacted into a scratch array (one global reservation per chunk);
//   route_probe_kernel   one survivor per thread: one 256-bit load per probed bucket, hot-key chains walked here (next[]
//                        belongs to the table owner), candidates staged in shared memory with their rank per read owner,
//                        one global reservation per owner and chunk (beyond the stage: one global atomic each).
#define PGM_ROUTE_FILTER_PER_THREAD 4
#define PGM_ROUTE_FILTER_CHUNK (PGM_ROUTE_THREADS * PGM_ROUTE_FILTER_PER_THREAD)
struct RouteFilterParams {
    const uint32_t *src;            // received window entries of ONE sender, 3 words each
    uint64_t n;
    TableView tab;
    uint32_t *live;                 // survivors, 3 words each (capacity n)
    unsigned int *n_live;           // their number (device counter, zeroed by the host)
};

__global__ void __launch_bounds__(PGM_ROUTE_THREADS, 8) route_filter_kernel(const __grid_constant__ RouteFilterParams p) {
    __shared__ uint32_t s_in[3 * PGM_ROUTE_FILTER_CHUNK];
    __shared__ uint16_t s_idx[PGM_ROUTE_FILTER_CHUNK];
    __shared__ unsigned int s_nlive, s_base;
    const uint32_t t = threadIdx.x, lane = t & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    if (t == 0) s_nlive = 0;
    __syncthreads();
    const uint64_t n_chunks = (p.n + PGM_ROUTE_FILTER_CHUNK - 1) / PGM_ROUTE_FILTER_CHUNK;
    const uint64_t pol_keep = policy_evict_last();
    for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uint64_t e0 = c * PGM_ROUTE_FILTER_CHUNK;
        const uint32_t n_here = (uint32_t)min((uint64_t)PGM_ROUTE_FILTER_CHUNK, p.n - e0);
        for (uint32_t k = t; k < 3 * n_here; k += PGM_ROUTE_THREADS) s_in[k] = __ldcs(p.src + 3 * e0 + k);
        __syncthreads();
        uint32_t fw[PGM_ROUTE_FILTER_PER_THREAD], fm[PGM_ROUTE_FILTER_PER_THREAD];
#pragma unroll
        for (int k = 0; k < PGM_ROUTE_FILTER_PER_THREAD; k++) {
            const uint32_t i = t + k * PGM_ROUTE_THREADS;
            fw[k] = 0xFFFFFFFFu; fm[k] = 0;
            if (i < n_here) {
                const uint32_t f = route_filter_hash(s_in[3 * i], s_in[3 * i + 1]);
                fm[k] = filter_bits(f, p.tab.filter_k);
                if (p.tab.filter) fw[k] = ld_u32_hint(p.tab.filter + (f & p.tab.filter_mask), pol_keep);
            }
        }
        uint32_t bal[PGM_ROUTE_FILTER_PER_THREAD], tot = 0;
#pragma unroll
        for (int k = 0; k < PGM_ROUTE_FILTER_PER_THREAD; k++) {
            bal[k] = __ballot_sync(PGM_FULL, t + k * PGM_ROUTE_THREADS < n_here && (fw[k] & fm[k]) == fm[k]);
            tot += __popc(bal[k]);
        }
        if (tot) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&s_nlive, tot);
            base = __shfl_sync(PGM_FULL, base, 0);
#pragma unroll
            for (int k = 0; k < PGM_ROUTE_FILTER_PER_THREAD; k++) {
                if ((bal[k] >> lane) & 1u) s_idx[base + __popc(bal[k] & lt_mask)] = (uint16_t)(t + k * PGM_ROUTE_THREADS);
                base += __popc(bal[k]);
            }
        }
        __syncthreads();
        const uint32_t n_live = s_nlive;
        if (t == 0) s_base = n_live ? atomicAdd(p.n_live, n_live) : 0u;
        __syncthreads();
        if (t == 0) s_nlive = 0;
        uint32_t *out = p.live + 3 * (size_t)s_base;
        for (uint32_t k = t; k < 3 * n_live; k += PGM_ROUTE_THREADS) {
            const uint32_t i = k / 3u, w = k - 3u * i;
            out[k] = s_in[3 * s_idx[i] + w];
        }
        __syncthreads();
    }
}

#define PGM_ROUTE_PROBE_CHUNK2 (PGM_ROUTE_THREADS * 2)
__global__ void __launch_bounds__(PGM_ROUTE_THREADS, PGM_ROUTE_PROBE_CTAS) route_probe_kernel(const __grid_constant__ RouteProbeParams p) {
    __shared__ uint32_t s_pos[PGM_ROUTE_PROBE_STAGE], s_pat[PGM_ROUTE_PROBE_STAGE];
    __shared__ uint16_t s_rank[PGM_ROUTE_PROBE_STAGE];
    __shared__ uint8_t s_dest[PGM_ROUTE_PROBE_STAGE];
    __shared__ unsigned int s_cnt[PGM_ROUTE_MAX_WORLD], s_base[PGM_ROUTE_MAX_WORLD];
    __shared__ unsigned int s_n;
    const uint32_t t = threadIdx.x;
    if (t < PGM_ROUTE_MAX_WORLD) s_cnt[t] = 0;
    if (t == 0) s_n = 0;
    __syncthreads();
    bool over = false;
    const uint64_t pol_stream = policy_evict_first();      // bucket lines are one-shot traffic: keep them from evicting the filter
    const uint64_t n = p.n_ptr ? (uint64_t)*p.n_ptr : p.n;
    const uint64_t n_chunks = (n + PGM_ROUTE_PROBE_CHUNK2 - 1) / PGM_ROUTE_PROBE_CHUNK2;
    auto emit_slow = [&](uint32_t rel, uint32_t pat) {
        const uint64_t gpos = p.pos_base + rel;
        const uint32_t dest = route_owner(p, pat >> p.part_bits);
        const uint32_t idx = atomicAdd(p.q.count + dest, 1u);
        if (idx < p.q.cap) {
            uint32_t *e = p.q.entries + ((size_t)dest * p.q.cap + idx) * 3;
            e[0] = (uint32_t)gpos; e[1] = (uint32_t)(gpos >> 32); e[2] = pat;
        } else over = true;
    };
    auto emit = [&](uint32_t rel, uint32_t pat) {
        const uint32_t at = atomicAdd(&s_n, 1u);
        if (at < PGM_ROUTE_PROBE_STAGE) {
            const uint32_t dest = route_owner(p, pat >> p.part_bits);
            s_pos[at] = rel; s_pat[at] = pat; s_dest[at] = (uint8_t)dest;
            s_rank[at] = (uint16_t)atomicAdd(&s_cnt[dest], 1u);
        } else emit_slow(rel, pat);
    };
    for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uint64_t first = c * PGM_ROUTE_PROBE_CHUNK2, last = min(first + PGM_ROUTE_PROBE_CHUNK2, n);
        for (uint64_t i = first + t; i < last; i += PGM_ROUTE_THREADS) {
            const uint32_t h1p = __ldcs(p.src + 3 * i), h2 = __ldcs(p.src + 3 * i + 1), rel = __ldcs(p.src + 3 * i + 2);
            const uint32_t tag = seed_tag(h2);
            uint32_t b = __umulhi(h1p, p.tab.n_buckets);
            const uint32_t step = 1u + __umulhi(h2 * 0x9E3779B1u, p.tab.n_buckets - 1u);
            for (;;) {
                const u32x8 sl = ld256_stream_hint(p.tab.buckets + b, pol_stream);
                bool em = false;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    em |= sl.v[2 * k + 1] == 0xFFFFFFFFu;
                    if ((sl.v[2 * k + 1] & 0x7FFFFFFFu) == tag) {
                        uint32_t pat = sl.v[2 * k];
                        emit(rel, pat);
                        if (sl.v[2 * k + 1] & 0x80000000u)         // hot key: the patterns chained behind this slot
                            for (pat = __ldg(p.tab.next + pat); pat != PGM_NIL; pat = __ldg(p.tab.next + pat)) emit(rel, pat);
                    }
                }
                if (em) break;                                     // a bucket with an empty slot ends the probe sequence
                b += step;
                if (b >= p.tab.n_buckets) b -= p.tab.n_buckets;
            }
        }
        __syncthreads();
        const uint32_t staged = min(s_n, (unsigned int)PGM_ROUTE_PROBE_STAGE);
        if (t < p.world) {
            const unsigned int cc = s_cnt[t];
            s_base[t] = cc ? atomicAdd(p.q.count + t, cc) : 0u;
            s_cnt[t] = 0;
        }
        __syncthreads();
        if (t == 0) s_n = 0;
        for (uint32_t i = t; i < staged; i += PGM_ROUTE_THREADS) {
            const uint32_t dest = s_dest[i], idx = s_base[dest] + s_rank[i];
            if (idx < p.q.cap) {
                const uint64_t gpos = p.pos_base + s_pos[i];
                uint32_t *e = p.q.entries + ((size_t)dest * p.q.cap + idx) * 3;
                e[0] = (uint32_t)gpos; e[1] = (uint32_t)(gpos >> 32); e[2] = s_pat[i];
            } else over = true;
        }
        __syncthreads();
    }
    if (over) *p.q.overflow = 1u;
    if (blockIdx.x == 0 && t == 0 && p.n_ptr) atomicAdd(p.counters + 3, (unsigned long long)n);   // windows that passed the filter
}

// ------------------------------------------------------------------------------------------ pass: verify
struct RouteVerifyParams {
    VerifyParams v;                 // planes of the WHOLE text of this pass (pos_origin = bit_origin = 0), reads = MY reads
    const uint32_t *src;            // received candidates {position lo, position hi, pattern}
    uint64_t n;
    uint32_t read_base;             // global index of my first read
    unsigned long long *counters;   // [0] candidates [1] verified [2] accepted
};

// One candidate per lane.  LQ64: 64-byte ACGT records are fetched by lane pairs (one instruction = one 64-byte request)
// and the text window comes through gather7; every other candidate is verified by its lane alone (count_alone).
template <bool LQ64>
__global__ void __launch_bounds__(PGM_VERIFY_THREADS) route_verify_kernel(const __grid_constant__ RouteVerifyParams rp) {
    const VerifyParams &p = rp.v;
    const uint32_t t = threadIdx.x, lane = t & 31u, half = lane & 1u;
    const uint32_t pmask = (1u << p.reads.part_bits) - 1u;
    unsigned long long n_cand = 0, n_ver = 0, n_acc = 0;
    const uint64_t span = (uint64_t)gridDim.x * PGM_VERIFY_THREADS;
    const uint64_t rounds = (rp.n + span - 1) / span;              // same trip count for every warp (shuffles inside)
    for (uint64_t it = 0; it < rounds; it++) {
        const uint64_t i = it * span + (uint64_t)blockIdx.x * PGM_VERIFY_THREADS + t;
        const bool on = i < rp.n;
        uint32_t e0 = 0, e1 = 0, cpat = 0;
        if (on) { e0 = __ldcs(rp.src + 3 * i); e1 = __ldcs(rp.src + 3 * i + 1); cpat = __ldcs(rp.src + 3 * i + 2); }
        const uint64_t gpos = ((uint64_t)e1 << 32) | e0;
        const uint32_t cr = on ? (cpat >> p.reads.part_bits) - rp.read_base : 0u, cj = cpat & pmask;
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
            // the text window's address is known from the candidate alone: its loads go out BEFORE the shuffles below wait
            // for the record (a warp issues in order: behind the shuffles they would cost a second DRAM latency)
            const int64_t lbit = (int64_t)gpos - (int64_t)(cj * p.seed_len);
            uint32_t yl[7], yh[7];
#pragma unroll
            for (int w = 0; w < 7; w++) { yl[w] = 0; yh[w] = 0; }
            if (fast) {
                gather7(p.tlo, lbit >> 5, yl);
                gather7(p.thi, lbit >> 5, yh);
            }
            uint32_t s0[8], s1[8];
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint32_t got = __shfl_xor_sync(PGM_FULL, half ? xA.v[w] : xB.v[w], 1);
                s0[w] = half ? got : xA.v[w];
                s1[w] = half ? xB.v[w] : got;
            }
            if (fast) {
                n_cand++;
                const uint32_t ts = (uint32_t)(lbit & 31);
                const uint32_t L = p.reads.read_len;
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
        if (on && !fast) {
            n_cand++;
            uint32_t stride16; bool is_n;
            uint4 *rec = record_of(p.reads, cr, stride16, is_n);
            const uint4 h = __ldcg(rec);
            const int c = count_alone(p, rec, is_n, (int64_t)gpos - (int64_t)(cj * p.seed_len));
            apply_event(p, cr, cj, c, h.x, h.y, (long long)(((uint64_t)h.w << 32) | h.z), gpos,
                        reinterpret_cast<long long *>(rec) + 1, n_ver, n_acc);
        }
    }
    unsigned long long cv[3] = {n_cand, n_ver, n_acc};
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cv[k] += __shfl_xor_sync(PGM_FULL, cv[k], o);
        if (lane == 0 && cv[k]) atomicAdd(rp.counters + k, cv[k]);
    }
}

} // namespace pgm
