// pgm_part.cuh — the PARTITIONED scan: an exact pre-filter for pattern sets far beyond what an L2-resident Bloom filter
// can tell apart (configs 4 and 5: 0.25 - 1 G patterns).
//
// Why.  With 990 M patterns behind 2^29 - 2^30 filter bits, 63 - 85 % of the text windows pass the filter and every one of
// them costs a random 128-byte DRAM line of the 24 GB bucket array: the fused scan runs at the DRAM's random-line rate on
// false positives.  A hash join of that size wants its probes PARTITIONED: the windows of a round are written
// into one queue per table partition (a contiguous range of home buckets of about 48 MB — L2-sized), and then the partitions
// are probed one after the other, every probe an L2 hit.  Both steps stream 12 bytes per window; nothing is random.
//
//   part_scan_kernel    text order (TMA-staged 4096-position tiles, as in scan_kernel): every window start is hashed
//                       (canonical form -> 64-bit key, the fused kernel's own window_hash) and its entry {h1, h2, position}
//                       appended to the queue of partition h1 >> (32 - part_bits); ranks inside a tile through shared-memory
//                       counters, one global reservation per partition and tile.
//   part_probe_kernel   partition order, all CTAs on the same partition: one 256-bit load per probed bucket (default caching:
//                       the partition stays in the L2), the fused kernel's probe sequence; a tag hit sets the window's bit in
//                       a bitmap over the round's text positions.
//   scan_kernel<MODE 2> the fused kernel with that bitmap in place of stage A1 (hash + filter gather): the set bits ARE the
//                       windows with a table hit — 11.6 % of the positions at config 5 instead of the 63 % a Bloom filter passes — and stages A2 (probe)
//                       and B (verification against the staged text, accept test, atomicMin) run unchanged on them.
// STATUS: opt-in (PGM_PART_SCAN=1 auto / 2 always / 3 always with tiny queues), off by default.  Measured at config 5 on B200
// (profiles/bench_c5_n1_part_r02x.json): bit-exact, but 1312 ms per step against 450 for the sliced Bloom filter — emit 917 ms
// (a lane per window and 512 queues: every warp store touches 32 sectors), probe 177 ms (each partition sweep walks the round's
// whole 234 MB bitmap), fused kernel on the bitmap 107 ms.  It also showed that 11.6 % of config 5's windows are TRUE table
// hits: even a perfect pre-filter leaves 53 ms per pass of bucket + record lines, so a well-engineered version of this pipeline
// would gain about 20 % per pass at best.
// The bitmap may be conservative (a window whose queue was full is marked without probing; stage A2 probes it again) but
// never misses a hit: part_probe_kernel walks the same probe sequence with the same stop rule as stage A2.  The result is the
// fused kernel's, bit for bit (tests: PGM_PART_SCAN=2 / 3 force the pipeline, 3 with tiny queues).
#pragma once
#include "pgm_kernels.cuh"

#define PGM_PART_MAX 1024                   // most partitions
#define PGM_PART_THREADS 256
#define PGM_PART_PROBE_CHUNK 2048           // queue entries a CTA takes at a time

namespace pgm {

struct PartQueues {
    uint32_t *entries;              // partition k owns the 3-word entries [k * cap, k * cap + min(count[k], cap))
    unsigned int *count;            // PGM_PART_MAX
    unsigned int *cursor;           // PGM_PART_MAX (part_probe_kernel's chunk dispenser)
    uint32_t cap;
    uint32_t part_bits;             // partitions = 1 << part_bits
    uint32_t *hit_bits;             // bit per launch-relative position (tile * 4096 + position in tile)
};

struct PartScanParams {
    const uint32_t *tlo, *thi;      // planes of this pass's text, local origin at word 0 (as ScanParams)
    uint64_t slice_origin, own_begin, own_end;
    uint32_t first_word, n_tiles, tail_mask;
    unsigned int *tile_counter;
    PartQueues q;
};

struct PartScanShared {
    uint32_t lo[2][PGM_BUF_WORDS];
    uint32_t hi[2][PGM_BUF_WORDS];
    uint32_t h1[PGM_TILE_POS], h2[PGM_TILE_POS];
    uint32_t rk[PGM_TILE_POS];      // partition << 16 | rank among the tile's entries of that partition; PGM_NIL = not owned
    uint64_t bar[2];
    unsigned int cnt[PGM_PART_MAX], base[PGM_PART_MAX], tile[2];
};

template <int NCH>
__global__ void __launch_bounds__(PGM_PART_THREADS, 3) part_scan_kernel(const __grid_constant__ PartScanParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PartScanShared &sm = *reinterpret_cast<PartScanShared *>(smem_raw);
    const uint32_t t = threadIdx.x;
    const uint32_t n_parts = 1u << p.q.part_bits, pshift = 32u - p.q.part_bits;

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
    for (uint32_t k = t; k < n_parts; k += PGM_PART_THREADS) sm.cnt[k] = 0;
    __syncthreads();
    uint32_t parity[2] = {0, 0};
    int buf = 0;
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
        const int64_t tile_word0 = (int64_t)p.first_word + (int64_t)tile * PGM_TILE_WORDS;
        const uint64_t tile_g0 = p.slice_origin + (uint64_t)tile_word0 * 32;
        const int64_t vb64 = (int64_t)p.own_begin - (int64_t)tile_g0, ve64 = (int64_t)p.own_end - (int64_t)tile_g0;
        const uint32_t vb = (uint32_t)max((int64_t)0, min(vb64, (int64_t)PGM_TILE_POS));
        const uint32_t ve = (uint32_t)max((int64_t)0, min(ve64, (int64_t)PGM_TILE_POS));
        // phase 1: hash, partition, rank
#pragma unroll 4
        for (uint32_t pos = t; pos < PGM_TILE_POS; pos += PGM_PART_THREADS) {
            uint32_t r = PGM_NIL;
            if (pos >= vb && pos < ve) {
                const uint64_t hv = window_hash<NCH>(slo, shi, pos, p.tail_mask);
                const uint32_t h1 = (uint32_t)hv, part = h1 >> pshift;
                sm.h1[pos] = h1;
                sm.h2[pos] = (uint32_t)(hv >> 32);
                r = (part << 16) | atomicAdd(&sm.cnt[part], 1u);
            }
            sm.rk[pos] = r;
        }
        __syncthreads();
        for (uint32_t k = t; k < n_parts; k += PGM_PART_THREADS) {
            const unsigned int c = sm.cnt[k];
            sm.base[k] = c ? atomicAdd(p.q.count + k, c) : 0u;
            sm.cnt[k] = 0;
        }
        __syncthreads();
        // phase 2: write (a full queue: the window is marked without a probe, the fused kernel's stage A2 looks at it)
        const uint32_t rel0 = tile * PGM_TILE_POS;
        for (uint32_t pos = t; pos < PGM_TILE_POS; pos += PGM_PART_THREADS) {
            const uint32_t r = sm.rk[pos];
            if (r != PGM_NIL) {
                const uint32_t part = r >> 16, idx = sm.base[part] + (r & 0xFFFFu);
                if (idx < p.q.cap) {
                    uint32_t *e = p.q.entries + ((size_t)part * p.q.cap + idx) * 3;
                    __stcs(e, sm.h1[pos]); __stcs(e + 1, sm.h2[pos]); __stcs(e + 2, rel0 + pos);
                } else atomicOr(p.q.hit_bits + ((rel0 + pos) >> 5), 1u << (pos & 31u));
            }
        }
        __syncthreads();
        buf ^= 1;
    }
}

struct PartProbeParams {
    PartQueues q;
    TableView tab;
};

// Partition after partition (the CTAs move on together, a chunk dispenser per partition): every thread probes its entries.
__global__ void __launch_bounds__(PGM_PART_THREADS, 6) part_probe_kernel(const __grid_constant__ PartProbeParams p) {
    __shared__ unsigned int s_first[2];
    const uint32_t t = threadIdx.x;
    const uint32_t n_parts = 1u << p.q.part_bits;
    uint32_t flip = 0;
    for (uint32_t k = 0; k < n_parts; k++) {
        const uint32_t n = min(__ldg(p.q.count + k), p.q.cap);
        const uint32_t *src = p.q.entries + (size_t)k * p.q.cap * 3;
        for (;;) {
            if (t == 0) s_first[flip] = atomicAdd(p.q.cursor + k, (unsigned int)PGM_PART_PROBE_CHUNK);
            __syncthreads();
            const uint32_t first = s_first[flip];
            flip ^= 1;
            if (first >= n) break;
            const uint32_t last = min(first + PGM_PART_PROBE_CHUNK, n);
            for (uint32_t i = first + t; i < last; i += PGM_PART_THREADS) {
                const uint32_t h1 = __ldcs(src + 3 * (size_t)i), h2 = __ldcs(src + 3 * (size_t)i + 1), rel = __ldcs(src + 3 * (size_t)i + 2);
                const uint32_t tag = seed_tag(h2);
                uint32_t b = __umulhi(h1, p.tab.n_buckets);
                const uint32_t step = 1u + __umulhi(h2 * 0x9E3779B1u, p.tab.n_buckets - 1u);
                bool hit = false;
                for (;;) {
                    const u32x8 s = ld256_cg(p.tab.buckets + b);
                    bool em = false;
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        em |= s.v[2 * c + 1] == 0xFFFFFFFFu;
                        hit |= (s.v[2 * c + 1] & 0x7FFFFFFFu) == tag;
                    }
                    if (em || hit) break;                   // a bucket with an empty slot ends the probe sequence
                    b += step;
                    if (b >= p.tab.n_buckets) b -= p.tab.n_buckets;
                }
                if (hit) atomicOr(p.q.hit_bits + (rel >> 5), 1u << (rel & 31u));
            }
        }
    }
}

} // namespace pgm
