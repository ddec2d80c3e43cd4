// pgm_kernels.cuh — sm_100a kernels of the read-vs-pseudogenome matcher.
//
// Data layout in HBM
//   text        two bit planes per strand (lo = code&1, hi = code>>1; A=0 C=1 G=2 T=3), 32 bases
//               per uint32 word, base p at bit (p & 31) of word (p >> 5); PGM_PAD_WORDS zero
//               words in front of the origin, a zero tail behind it.  The reverse-complement
//               strand is materialised once (rc = ~bitreverse) so both passes run one kernel.
//   read record one per read, 64-byte multiples, fetched as ONE DRAM request by a lane pair (2 x 32 bytes):
//               uint4 #0 = {state64, best_key64}; ACGT set: uint4 #(1+g/2) = {lo,hi} of the 32-base
//               groups g, g+1; ACGNT set: uint4 #(1+g) = {lo, hi, nmask, 0} of group g (lo = hi = 0
//               under an N).  state64 = mm:8 | rc:1 | pos:40 ; best_key64 = cls:8 | txtPos:40 |
//               (parts-1-j):8 | mm:8 (MIN-mergeable accumulator of the current pass).
//   seed table  multimap, 32-byte buckets of four 8-byte slots {tag:31 chain:1 | pattern:32} (one 256-bit load
//               per probe), double hashing over a prime number of buckets; pattern = read << part_bits | seed.
//               Every pattern owns a slot of the first
//               bucket of its probe sequence that has room (so a lookup stops at the first bucket with
//               an empty slot); only when PGM_WALK_CAP buckets in a row are full of the same key (hot
//               seeds: poly-A, satellites) a pattern is chained behind a slot through next[].
//   filter      2^f-bit blocked Bloom filter (1..4 bits in one word, by load), sized to stay L2-resident.
//   per read    first_other_order (int64 MIN), same_pos_mask (int32 OR), same_pos_mm (uint8): only
//               touched when a read that already has a match meets a better candidate (rule a-R).
//
// Seed key.  The reference hashes a seed with CyclicHash<uint32>(n, 32)
// (rollinghash/cyclichash.h:29-35,100-123): symbol k is rotated by (n-1-k) mod 32, so two
// seeds collide for EVERY random table iff, per residue class of k mod 32, they contain each
// symbol with the same parity.  That equivalence is reproduced exactly by
// XOR-folding the window's bit planes into 32-bit words: P = fold(lo), Q = fold(hi),
// R = fold(lo & hi) determine the four parity vectors.  key = mix(P, Q, R).
//
// Memory-system facts the kernels are shaped by (tools/ubench.cu, profiles/ubench_r01.txt): a fully
// divergent 4-byte gather costs one L1 wavefront per lane (285 G lookups/s chip-wide); DRAM serves
// about 42 G random requests/s whether they carry 32 or 64 bytes, provided ONE instruction asks for
// the whole item (one lane x 32 bytes, or adjacent lanes covering 64 bytes); one lane issuing 2 x 32 or
// 4 x 16 bytes gets about half of that.  So a table probe is one 256-bit load by one lane and a read
// record is fetched by a lane pair.  A first version with lane quads everywhere was instruction-issue
// bound (profiles/scan_lines_r01d.txt): control flow replicated on four lanes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PGM_TILE_WORDS 128                 // text words per tile
#define PGM_TILE_POS (PGM_TILE_WORDS * 32) // 4096 text positions per tile
#define PGM_HALO_L 12                      // words staged left of a tile  (alignments reach back (parts-1)*n <= 255 bases)
#define PGM_HALO_R 20                      // words staged right of a tile (window + read tail <= 255 + 255 bases)
#define PGM_BUF_WORDS (PGM_HALO_L + PGM_TILE_WORDS + PGM_HALO_R)
#define PGM_PAD_WORDS 64                   // zero words in front of every plane
#define PGM_TAIL_WORDS (PGM_TILE_WORDS + 64)
#define PGM_SCAN_THREADS 256
#define PGM_SCAN_WARPS (PGM_SCAN_THREADS / 32)
#ifndef PGM_A1_UNROLL
#define PGM_A1_UNROLL 4                     // text words (x 32 positions) per A1 iteration of a warp: filter gathers in flight per lane
#endif
#ifndef PGM_SCAN_MIN_CTAS
#define PGM_SCAN_MIN_CTAS 4                 // resident CTAs per SM the scan kernel is compiled for (register budget)
#endif
#define PGM_WORDS_PER_WARP (PGM_TILE_WORDS / PGM_SCAN_WARPS)
#define PGM_WQ_CAP 320                     // per-warp candidate queue entries
#define PGM_WQ_ROUND 128                   // most entries one probe round can add (32 lanes x 4 slots)
#define PGM_ILV_WORDS 256                  // mode 'i': words of the de-interleaved tile per plane (stride <= PGM_ILV_MAX_PARTS)
#define PGM_ILV_MAX_PARTS 31
#define PGM_WALK_CAP 6                     // full buckets walked before a duplicate key is chained

#define PGM_EMPTY64 0xFFFFFFFFFFFFFFFFull
#define PGM_NIL 0xFFFFFFFFu
#define PGM_KEY_INF 0x7FFFFFFFFFFFFFFFll
#define PGM_POS_MASK 0xFFFFFFFFFFull       // 40-bit positions
#define PGM_STATE_UNMATCHED ((255ull << 56) | PGM_POS_MASK)
#define PGM_FULL 0xFFFFFFFFu

namespace pgm {

// ------------------------------------------------------------------------------------------ hashing
// 96-bit canonical seed form -> 64 well-mixed bits (splitmix64 finaliser).
__host__ __device__ __forceinline__ uint64_t seed_hash64(uint32_t P, uint32_t Q, uint32_t R) {
    uint64_t v = (((uint64_t)Q << 32) | P) ^ ((uint64_t)R * 0x9E3779B97F4A7C15ull);
    v ^= v >> 30; v *= 0xBF58476D1CE4E5B9ull;
    v ^= v >> 27; v *= 0x94D049BB133111EBull;
    v ^= v >> 31;
    return v;
}
__host__ __device__ __forceinline__ uint32_t seed_tag(uint32_t h2) {
    const uint32_t t = h2 & 0x7FFFFFFFu;
    return t == 0x7FFFFFFFu ? 0x7FFFFFFEu : t;   // 0x7FFFFFFF is what an empty slot shows
}
// The pre-filter has its own, cheaper 32-bit hash of the canonical form (it is evaluated for EVERY text window; the
// 64-bit key only for the windows that pass).  A weaker hash here can only cost extra table probes, never a result.
__host__ __device__ __forceinline__ uint32_t filter_hash(uint32_t P, uint32_t Q, uint32_t R) {
    const uint32_t q = Q * 0x85EBCA77u, r = R * 0xC2B2AE3Du;
    uint32_t x = (P * 0x9E3779B1u) ^ ((q << 13) | (q >> 19)) ^ ((r << 7) | (r >> 25));
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}
// k = 1 or 2 bits of one filter word, chosen on the host from the filter's bits per entry (about 0.69 * bits / entries
// minimises the false-positive rate; 3 and 4 bits were measured: they cost more instructions per text window than they
// save probes; a table far beyond the L2-resident filter's capacity wants k = 1)
__host__ __device__ __forceinline__ uint32_t filter_bits(uint32_t f, uint32_t k) {
    const uint32_t g = f * 0x9E3779B1u;
    return (1u << (f >> 27)) | (k > 1u ? 1u << (g >> 27) : 0u);
}

// Paired lookups.  The windows starting at text positions x and x+1 share all but one base; offsets [lo, hi) of a
// window (lo = max(0, n-32), hi = min(n, 32)) are the ones the 32-bit rotate-xor hash sees exactly (one symbol per
// rotation class), so bits [lo, hi) of P and Q ARE those bases.  The filter word of a window is chosen by a hash of the
// core it shares with its pair partner — offsets [lo+1, hi) when it is the first of the pair (role 0), [lo, hi-1) when
// it is the second (role 1): the same text bases, hence the same word, and ONE gather answers both windows (the bits
// inside the word still come from the window's full form).  A pattern cannot know its role: it is entered in both
// words, which doubles the filter's load — the host turns pairing on only while that is harmless (few patterns per
// filter bit: multi-GPU read shards, small inputs), where the per-window gather is what bounds the scan.
__host__ __device__ __forceinline__ uint32_t pair_word(uint32_t P, uint32_t Q, uint32_t role, uint32_t lo, uint32_t cmask) {
    const uint32_t sh = lo + 1u - role;
    const uint32_t a = ((P >> sh) & cmask) * 0x9E3779B1u, b = ((Q >> sh) & cmask) * 0x85EBCA77u;
    uint32_t x = a ^ ((b << 15) | (b >> 17));
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15;
    return x;
}

// ------------------------------------------------------------------------------------------ parameters
struct __align__(32) u32x8 { uint32_t v[8]; };

struct TableView {
    u32x8 *buckets;             // 4 slots per bucket; slot = {pattern, chain:1 | tag:31}
    uint32_t *next;             // chains of hot keys
    uint32_t *filter;           // may be null
    uint32_t n_buckets;         // prime
    uint32_t filter_mask;       // #filter words - 1
    uint32_t filter_k;          // bits per pattern in its filter word (1 or 2)
    uint32_t pair;              // 1: paired lookups — the filter word is chosen by the bases two adjacent windows share (below)
    uint32_t pair_lo, pair_mask;   // first exactly-hashed window offset, mask of the shared core's bits
    // Hash-sliced filter (pattern sets far beyond what an L2-resident filter can hold apart): the filter has 2^slice_bits
    // slices of 64 MB, slice = word index >> slice_shift; a scan launch looks at the windows of ONE slice only, so the part
    // of the filter it gathers from stays L2-resident while the whole filter has 4+ bits per pattern.  The text (2 bits per
    // base) is re-hashed once per slice: ALU work instead of one random DRAM line per false positive.
    uint32_t slice_shift;       // 31 when the filter has one slice
};

struct ReadsView {
    uint4 *lq;                  // n_lq records of lq_stride16 uint4
    uint4 *nn;                  // n_n records of n_stride16 uint4
    uint32_t n_lq, n_n;
    uint32_t lq_stride16, n_stride16;
    uint32_t read_len, W;
    uint32_t part_bits;         // pattern id = read << part_bits | seed index
};

struct PerRead {
    long long *first_other_order;   // txtPos:40 | (parts-1-j):8  (earliest accepted event reporting a position != stored)
    int *same_pos_mask;             // bit j: seed j hit the alignment that reports the stored pos
    uint8_t *same_pos_mm;           // its mismatch count
    int *touched;                   // set when any of the three above was written in this pass
};

// Queues of the L2-blocked scan pipeline (pgm_blocked.cuh): the filter stage appends every filter-positive window
// to the queue of its table region (= range of home buckets = range of h1), the probe stage appends every tag hit
// to the queue of its read range.  Entries beyond a queue's capacity raise *overflow: the pass is then redone by the
// fused scan kernel (all per-read accumulators are idempotent MIN / OR updates).
#define PGM_SQ_MAX 256                      // most table regions / read ranges
struct StageQueues {
    uint4 *pos_entries;             // region k owns [k * pos_cap, (k + 1) * pos_cap): {h1, h2, launch-relative position, 0}
    uint2 *cand_entries;            // range k owns [k * cand_cap, (k + 1) * cand_cap): {position | chain << 31, pattern}
    unsigned int *pos_count, *pos_cursor, *cand_count, *cand_cursor;   // PGM_SQ_MAX each
    unsigned int *overflow;
    unsigned long long *counters;   // staged work counters [0] candidates [1] verified [2] accepted [3] filter positives
    uint32_t pos_cap, cand_cap;
    uint32_t region_bits;           // table regions = 1 << region_bits (1..8), region = h1 >> (32 - region_bits)
    uint32_t range_shift, n_ranges; // read range = read >> range_shift
};

struct ScanParams {
    const uint32_t *tlo, *thi;      // planes of this pass's text, local origin at word 0
    uint64_t slice_origin;          // global coordinate of local position 0 (this pass's coordinates)
    uint64_t own_begin, own_end;    // owned seed-window starts, global, already clipped to <= pg_len - n + 1
    uint64_t pg_len;
    uint32_t first_word;            // first local word of tile 0 (multiple of 4)
    uint32_t n_tiles;
    uint32_t seed_len, parts, max_mm, min_mm;
    uint32_t shift_unit;            // alignment start = window start - j * shift_unit: seed_len (mode 'd'), 1 (mode 'i')
    uint32_t ilv;                   // 0: seed j = read bases [j*n, (j+1)*n) (mode 'd'); else the stride (= parts) of mode 'i'
    uint32_t tail_mask;             // valid bits of the last 32-base chunk of a seed
    uint32_t slice;                 // filter slice this launch handles (TableView::slice_shift)
    int rev_mode;
    int l2_hints;                   // (unused: filter loads always carry an L2 evict_last policy)
    int stream_hints;               // 1: bucket / record loads carry an L2 evict_first policy
    TableView tab;
    ReadsView reads;
    PerRead pr;
    unsigned int *tile_counter;
    unsigned long long *counters;   // [0] candidates [1] verified [2] accepted [3] filter positives
    StageQueues sq;                 // MODE 1 (filter stage of the L2-blocked pipeline) only
    const unsigned int *only_if;    // fused kernel as the pipeline's fallback: runs only when *only_if != 0, else commits sq.counters
    const uint32_t *hit_bits;       // MODE 2 (behind the partitioned pre-filter, pgm_part.cuh): bit per launch-relative position
};

// ------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// TMA bulk copy global -> shared (UBLKCP), completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// L2 eviction policies (createpolicy): the filter is the one structure worth keeping in L2; table buckets and
// read records are one-shot random traffic.
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint32_t ld_u32_hint(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint4 ld_u4_hint(const uint4 *p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}

// 256-bit loads (LDG.256, new on sm_100): one request for a whole 32-byte sector
__device__ __forceinline__ u32x8 ld256_stream(const void *p) {
    u32x8 r;
    asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ u32x8 ld256_stream_hint(const void *p, uint64_t pol) {
    u32x8 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ u32x8 ld256_cg(const void *p) {
    u32x8 r;
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t rotr32(uint32_t x, uint32_t r) { return __funnelshift_r(x, x, r); }


// record of global read r: pointer, stride (uint4), ACGNT?
__device__ __forceinline__ uint4 *record_of(const ReadsView &rv, uint32_t r, uint32_t &stride16, bool &is_n) {
    is_n = r >= rv.n_lq;
    stride16 = is_n ? rv.n_stride16 : rv.lq_stride16;
    return is_n ? rv.nn + (size_t)(r - rv.n_lq) * rv.n_stride16 : rv.lq + (size_t)r * rv.lq_stride16;
}

// ------------------------------------------------------------------------------------------ text packing
// ASCII (1 byte/base) -> bit planes.  One thread per 32-base word.  Sets *err when a symbol is
// outside ACGT (the reference's HQ pseudogenome alphabet, DividedPCLReadsSets.cpp:12-13).
__global__ void pack_text_kernel(const uint8_t *__restrict__ ascii, uint64_t n_bases, uint32_t *__restrict__ lo,
                                 uint32_t *__restrict__ hi, uint64_t first_word, int *err) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t base = w * 32;
    if (base >= n_bases) return;
    const uint32_t cnt = (uint32_t)min((uint64_t)32, n_bases - base);
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
                const uint32_t c = i < cnt ? ascii[base + i] : (uint32_t)'A';
                v |= c << (8 * b);
            }
            bytes[q] = v;
        }
    }
    uint32_t l = 0, h = 0, bad = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const uint32_t v = bytes[q];
        // per byte: x = (c >> 1) & 3 gives A0 C1 G3 T2; code = x ^ (x >> 1) gives A0 C1 G2 T3
        const uint32_t x = (v >> 1) & 0x03030303u;
        const uint32_t code = x ^ ((x >> 1) & 0x01010101u);
        // re-encode and compare to validate the symbol: "ACGT" as bytes 0x41 0x43 0x47 0x54
        const uint32_t c0 = (0x54474341u >> (8 * (code & 3))) & 0xFF;
        const uint32_t c1 = (0x54474341u >> (8 * ((code >> 8) & 3))) & 0xFF;
        const uint32_t c2 = (0x54474341u >> (8 * ((code >> 16) & 3))) & 0xFF;
        const uint32_t c3 = (0x54474341u >> (8 * ((code >> 24) & 3))) & 0xFF;
        bad |= (c0 | (c1 << 8) | (c2 << 16) | (c3 << 24)) ^ v;
        // gather bit 0 (lo) and bit 1 (hi) of the four bytes into 4 consecutive bits
        const uint32_t lb = ((code & 0x01010101u) * 0x01020408u) >> 24;
        const uint32_t hb = (((code >> 1) & 0x01010101u) * 0x01020408u) >> 24;
        l |= (lb & 0xF) << (4 * q);
        h |= (hb & 0xF) << (4 * q);
    }
    if (cnt < 32) {  // missing bytes were synthesised as 'A' (valid, code 0): just mask them off
        const uint32_t m = (1u << cnt) - 1u;
        l &= m; h &= m;
    }
    if (bad) atomicExch(err, 1);
    lo[first_word + w] = l;
    hi[first_word + w] = h;
}

// Reverse-complement planes of a slice of `len` bases: rc[i] = complement(fwd[len-1-i]).
__global__ void rc_text_kernel(const uint32_t *__restrict__ flo, const uint32_t *__restrict__ fhi, uint64_t len,
                               uint32_t *__restrict__ rlo, uint32_t *__restrict__ rhi) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w * 32 >= len) return;
    // rc positions [32w, 32w+32) <-> forward positions (len-32w-32 .. len-32w-1], possibly reaching below 0
    const int64_t sb = (int64_t)len - (int64_t)w * 32 - 32; // forward bit position of rc bit 31
    const int64_t wi = sb >> 5;                             // floor; >= -1, the pad words in front are zero
    const uint32_t sh = (uint32_t)(sb & 31);
    const uint32_t vl = __funnelshift_r(flo[wi], flo[wi + 1], sh);
    const uint32_t vh = __funnelshift_r(fhi[wi], fhi[wi + 1], sh);
    uint32_t valid = 0xFFFFFFFFu;
    const uint64_t left = len - w * 32;
    if (left < 32) valid = (1u << left) - 1u;
    rlo[w] = ~__brev(vl) & valid;
    rhi[w] = ~__brev(vh) & valid;
}

// ------------------------------------------------------------------------------------------ reads
// bits 0, 2, 4, ... 30 of x packed into the low 16 bits
__device__ __forceinline__ uint32_t compress_even(uint32_t x) {
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}

// Packed reads (reference layout, SymbolsPackingFacility.cpp:147-185) -> read records (layout above) with a
// fresh header {unmatched, no key}.  A block stages its reads' packed bytes in shared memory with coalesced
// loads, one thread builds one record in shared memory, and the block writes the records out coalesced.
__global__ void unpack_reads_kernel(const uint8_t *__restrict__ packed, uint32_t n_reads, uint32_t read_len,
                                    uint32_t packed_len, int with_n, uint4 *__restrict__ recs, uint32_t stride16,
                                    uint32_t W) {
    extern __shared__ __align__(16) uint8_t sbuf[];
    const uint32_t r0 = blockIdx.x * blockDim.x;
    const uint32_t nb = min(blockDim.x, n_reads - r0);
    const size_t g0 = (size_t)r0 * packed_len;
    const uint32_t bytes = nb * packed_len;
    uint4 *sout = reinterpret_cast<uint4 *>(sbuf + ((blockDim.x * packed_len + 15) & ~15u));
    {   // head bytes up to 16-byte alignment, 16-byte body, tail bytes
        const uint8_t *src = packed + g0;
        const uint32_t head = min(bytes, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15));
        const uint32_t body = (bytes - head) & ~15u;
        for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) sbuf[i] = src[i];
        if (head == 0) {
            for (uint32_t i = threadIdx.x * 16; i < body; i += blockDim.x * 16)
                *reinterpret_cast<uint4 *>(sbuf + i) = __ldg(reinterpret_cast<const uint4 *>(src + i));
        } else {
            for (uint32_t i = threadIdx.x * 16; i < body; i += blockDim.x * 16) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + head + i));
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 16; k++) sbuf[head + i + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
            }
        }
        for (uint32_t i = head + body + threadIdx.x; i < bytes; i += blockDim.x) sbuf[i] = src[i];
    }
    __syncthreads();
    if (threadIdx.x < nb) {
        const uint8_t *src = sbuf + threadIdx.x * packed_len;
        uint4 *dst = sout + (size_t)threadIdx.x * stride16;
        dst[0] = make_uint4(0xFFFFFFFFu, 0xFF0000FFu, 0xFFFFFFFFu, 0x7FFFFFFFu);   // PGM_STATE_UNMATCHED, PGM_KEY_INF
        if (!with_n) {
            // 16 bases per 32-bit word of packed bytes (first base in the two most significant bits of its byte, the
            // tail of the last byte is 'A' = 0): reversing the bits inside every byte (brev + byte swap) puts base i's
            // hi bit at bit 2i and its lo bit at 2i+1; two even-bit compressions then give 16 bits of each plane.
            const uint32_t boff = threadIdx.x * packed_len;
            const uint32_t *sw = reinterpret_cast<const uint32_t *>(sbuf) + (boff >> 2);
            const uint32_t sel = 0x3210u + 0x1111u * (boff & 3u);
            const uint32_t nwords = (packed_len + 3) >> 2;
            for (uint32_t u = 1; u < stride16; u++) {
                uint32_t pl[4] = {0, 0, 0, 0};                 // lo, hi of group 2(u-1); lo, hi of group 2(u-1)+1
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t k = 4 * (u - 1) + q;
                    if (k < nwords) {
                        const uint32_t v = __byte_perm(sw[k], sw[k + 1], sel);
                        const uint32_t w = __byte_perm(__brev(v), 0u, 0x0123u);
                        uint32_t h = compress_even(w), l = compress_even(w >> 1);
                        const int rem = (int)read_len - 16 * (int)k;   // bases of the read in this word (bytes past the read are masked off)
                        const uint32_t m = rem >= 16 ? 0xFFFFu : rem <= 0 ? 0u : (1u << rem) - 1u;
                        l &= m; h &= m;
                        pl[2 * (q >> 1)] |= l << (16 * (q & 1));
                        pl[2 * (q >> 1) + 1] |= h << (16 * (q & 1));
                    }
                }
                dst[u] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
        } else {
            uint32_t lo = 0, hi = 0, nm = 0, g = 0, bitpos = 0, p = 0;
            for (uint32_t b = 0; b < packed_len; b++) {
                const uint32_t v = src[b];
                const uint32_t sy[3] = {v / 25u, (v / 5u) % 5u, v % 5u};
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    if (p < read_len) {
                        const uint32_t sym = sy[k];
                        const uint32_t isn = sym == 3u ? 1u : 0u;
                        const uint32_t code = sym == 4u ? 3u : (isn ? 0u : sym);
                        lo |= (code & 1u) << bitpos; hi |= (code >> 1) << bitpos; nm |= isn << bitpos;
                        p++; bitpos++;
                        if (bitpos == 32) { dst[1 + g] = make_uint4(lo, hi, nm, 0); g++; lo = hi = nm = 0; bitpos = 0; }
                    }
                }
            }
            if (bitpos) { dst[1 + g] = make_uint4(lo, hi, nm, 0); g++; }
            for (uint32_t u = 1 + g; u < stride16; u++) dst[u] = make_uint4(0, 0, 0, 0);
        }
    }
    __syncthreads();
    uint4 *out = recs + (size_t)r0 * stride16;
    for (uint32_t i = threadIdx.x; i < nb * stride16; i += blockDim.x) out[i] = sout[i];
}

// ------------------------------------------------------------------------------------------ per-read state
__global__ void reset_state_kernel(ReadsView reads, PerRead pr, uint32_t n_reads, int reset_state, int records) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    if (records) {
        uint32_t stride16; bool is_n;
        uint4 *rec = record_of(reads, r, stride16, is_n);
        uint4 h = rec[0];
        if (reset_state) { h.x = 0xFFFFFFFFu; h.y = 0xFF0000FFu; }
        h.z = 0xFFFFFFFFu; h.w = 0x7FFFFFFFu;
        rec[0] = h;
    }
    pr.first_other_order[r] = PGM_KEY_INF;
    pr.same_pos_mask[r] = 0;
    pr.same_pos_mm[r] = 255;
}

// ------------------------------------------------------------------------------------------ table build
// 32 bits of a record's plane starting at read position `bit` (plane words sit `stride` words apart)
__device__ __forceinline__ uint32_t extract32(const uint32_t *w, uint32_t nwords, uint32_t stride, uint32_t bit) {
    const uint32_t k = bit >> 5;
    const uint32_t a = k < nwords ? __ldg(w + k * stride) : 0u;
    const uint32_t b = k + 1 < nwords ? __ldg(w + (k + 1) * stride) : 0u;
    return __funnelshift_r(a, b, bit & 31);
}

// Insert one pattern: first bucket of its double-hashing sequence with an empty slot (256-bit bucket load + CAS); when
// PGM_WALK_CAP buckets in a row are full and one of them already holds this key, chain behind that slot instead.
__device__ __forceinline__ void table_insert(const TableView &tab, uint32_t h1, uint32_t h2, uint32_t pat) {
    const uint32_t tag = seed_tag(h2);
    uint32_t b = __umulhi(h1, tab.n_buckets);
    const uint32_t step = 1u + __umulhi(h2 * 0x9E3779B1u, tab.n_buckets - 1u);
    const unsigned long long mine = ((unsigned long long)tag << 32) | pat;
    unsigned long long *same_slot = nullptr;
    uint32_t walked = 0;
    for (;;) {
        unsigned long long *bp = reinterpret_cast<unsigned long long *>(tab.buckets + b);
        const u32x8 s = ld256_cg(bp);
        bool done = false, saw_empty = false;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (done) break;
            if (s.v[2 * k + 1] == 0xFFFFFFFFu) {
                saw_empty = true;
                done = atomicCAS(bp + k, PGM_EMPTY64, mine) == PGM_EMPTY64;   // lost the race: try the next empty slot
            } else if (same_slot == nullptr && (s.v[2 * k + 1] & 0x7FFFFFFFu) == tag) {
                same_slot = bp + k;   // a slot that already holds this key
            }
        }
        if (done) return;
        if (saw_empty) continue;      // every empty slot seen was taken meanwhile: look at the bucket again
        if (++walked >= PGM_WALK_CAP && same_slot != nullptr) {
            // hot key: chain this pattern behind a slot that already holds the key
            unsigned long long old = *reinterpret_cast<volatile unsigned long long *>(same_slot);
            for (;;) {
                // next[] is not initialised: whoever extends a slot that is not chained yet marks its owner as the chain's end
                if (!(old >> 63)) tab.next[(uint32_t)old] = PGM_NIL;
                tab.next[pat] = (uint32_t)old;
                const unsigned long long nw = (((old >> 32) | 0x80000000ull) << 32) | pat;
                const unsigned long long prev = atomicCAS(same_slot, old, nw);
                if (prev == old) return;
                old = prev;
            }
        }
        b += step;
        if (b >= tab.n_buckets) b -= tab.n_buckets;
    }
}

// Mode 'i': the part of seed j (read bases j, j + s, j + 2s, ...; n of them) that lies in the 32-base group g of a read,
// XOR-folded into the canonical form (base j + k*s lands on bit k mod 32).
__device__ __forceinline__ void ilv_fold_group(uint32_t wl, uint32_t wh, uint32_t wn, uint32_t g, uint32_t j, uint32_t s, uint32_t n,
                                               uint32_t &P, uint32_t &Q, uint32_t &R, uint32_t &FN) {
    const uint32_t first = 32u * g, end = min(first + 32u, j + n * s);
    uint32_t k = first > j ? (first - j + s - 1u) / s : 0u;
    for (uint32_t b = j + k * s; b < end; b += s, k++) {
        const uint32_t sh = b - first, kk = k & 31u;
        const uint32_t l = (wl >> sh) & 1u, h = (wh >> sh) & 1u;
        P ^= l << kk; Q ^= h << kk; R ^= (l & h) << kk; FN ^= ((wn >> sh) & 1u) << kk;
    }
}

// Region queues of the two-step table build: the bucket array is cut into 2^region_bits equal ranges of the home
// bucket index (= ranges of h1); step 1 appends {h1, h2, pattern} to the queue of the pattern's region, step 2 inserts
// region after region, so that all CTAs work on one L2-sized piece of the table at a time and every table line is
// fetched from and written back to DRAM once (a direct random insert costs a 128-byte line read and a sector write
// per pattern: the one-step build ran at the DRAM's random-transaction limit, profiles/kernels_metrics_c2_r01h.csv).
struct BuildQueues {
    uint4 *entries;             // region k owns entries [k * cap, (k + 1) * cap)
    unsigned int *count;        // appended entries per region (may exceed cap: the excess was inserted directly)
    unsigned int *cursor;       // step 2: next entry to take per region
    uint32_t cap;
    uint32_t region_bits;       // 0 = queues off (small tables): step 1 inserts directly
    // routed multi-GPU build (pgm_routed.cuh): the "regions" are the `route_world` GPUs, region = umulhi(h1, route_world) = the
    // owner of the hash, the entry carries h1 * route_world (the part of h1 the routing did not use) and the GLOBAL pattern id
    // (read_base = global index of this GPU's first read); a full queue raises *overflow (nothing is inserted locally)
    uint32_t route_world;
    uint32_t read_base;
    unsigned int *overflow;
};

// Step 1.  One thread per read: for each of its `parts` seeds, fold read bases [j*n, (j+1)*n) into the canonical form
// (the record's plane words come through L1: adjacent lanes read adjacent records), set the filter bits and append
// the pattern to its region queue (ranks within the block through shared-memory counters, one global reservation per
// region and block).  Restates addReadsSetOfPatterns (ConstantLengthPatternsOnTextHashMatcher.cpp:23-42); the
// reference's pattern index r * parts + j (:39) is kept as (r << part_bits) | j.  Reads already matched with
// <= min_mm mismatches are left out when `continuation` (the matchedReadsBitmap argument, ReadsMatchers.cpp:290-291).
#define PGM_BUILD_THREADS 256
#define PGM_MAX_REGIONS 64
template <bool FAST>
__global__ void __launch_bounds__(PGM_BUILD_THREADS) build_table_kernel(ReadsView reads, TableView tab, BuildQueues q, uint32_t r_begin,
                                                                        uint32_t r_end, uint32_t seed_len, uint32_t parts, uint32_t min_mm,
                                                                        int continuation, uint32_t tail_mask, uint32_t ilv,
                                                                        unsigned long long *inserted) {
    extern __shared__ uint4 s_ent[];                                   // parts x blockDim staged entries {h1, h2, pattern, region | rank << 8}
    __shared__ unsigned int s_count[PGM_MAX_REGIONS], s_base[PGM_MAX_REGIONS];
    const uint32_t n_regions = q.route_world ? q.route_world : (q.region_bits ? 1u << q.region_bits : 0u);
    unsigned int n_ins = 0;
    const uint32_t span = gridDim.x * blockDim.x;
    const uint32_t rounds = (r_end - r_begin + span - 1) / span;       // same trip count for the whole block (barriers inside)
    for (uint32_t it = 0; it < rounds; it++) {
        const uint32_t r = r_begin + it * span + blockIdx.x * blockDim.x + threadIdx.x;
        bool active = r < r_end;
        uint32_t stride16 = 4; bool is_n = false;
        const uint4 *rec = reads.lq;
        u32x8 w0, w1;                                                  // FAST (ACGT set, 64-byte records): the whole record in registers
        if (FAST) {
#pragma unroll
            for (int k = 0; k < 8; k++) { w0.v[k] = 0; w1.v[k] = 0; }
            if (active) {
                rec = reads.lq + (size_t)r * 4;
                w0 = ld256_stream(rec); w1 = ld256_stream(rec + 2);
                if (continuation && (w0.v[1] >> 24) <= min_mm) active = false;
            }
        } else if (active) {
            rec = record_of(reads, r, stride16, is_n);
            if (continuation && (__ldg(reinterpret_cast<const uint32_t *>(rec) + 1) >> 24) <= min_mm) active = false;
        }
        const uint32_t *pl = reinterpret_cast<const uint32_t *>(rec) + 4;   // ACGT: lo,hi pairs; ACGNT: lo,hi,nm,0 quadruples
        const uint32_t il = is_n ? 4u : 2u;
        const uint32_t nch = (seed_len + 31) >> 5;
        if (n_regions) {
            for (uint32_t k = threadIdx.x; k < n_regions; k += blockDim.x) s_count[k] = 0;
            __syncthreads();
        }
        for (uint32_t j = 0; j < parts; j++) {
            uint4 ent = make_uint4(0, 0, 0, 0xFFFFFFFFu);
            if (active) {
                uint32_t P = 0, Q = 0, R = 0, FN = 0;
                if (FAST && ilv) {
#pragma unroll
                    for (int g = 0; g < 6; g++)
                        ilv_fold_group(g < 2 ? w0.v[4 + 2 * g] : w1.v[2 * (g - 2)], g < 2 ? w0.v[5 + 2 * g] : w1.v[2 * (g - 2) + 1], 0u,
                                       (uint32_t)g, j, ilv, seed_len, P, Q, R, FN);
                } else if (FAST) {
                    // a base at read position x lands on bit (x - j*n) mod 32: every 32-base group contributes
                    // rotr(group & [seed range], (j*n) mod 32) — no cross-register funnel shifts, static register indices
                    const int b0 = (int)(j * seed_len), b1 = b0 + (int)seed_len;
                    const uint32_t sh = (uint32_t)b0 & 31u;
#pragma unroll
                    for (int g = 0; g < 6; g++) {
                        const int lo_bit = max(b0 - 32 * g, 0), hi_bit = min(b1 - 32 * g, 32);
                        if (hi_bit > lo_bit) {
                            uint32_t m = hi_bit == 32 ? 0xFFFFFFFFu : (1u << hi_bit) - 1u;
                            m &= ~((1u << lo_bit) - 1u);
                            const uint32_t l = (g < 2 ? w0.v[4 + 2 * g] : w1.v[2 * (g - 2)]) & m;
                            const uint32_t h = (g < 2 ? w0.v[5 + 2 * g] : w1.v[2 * (g - 2) + 1]) & m;
                            P ^= rotr32(l, sh); Q ^= rotr32(h, sh); R ^= rotr32(l & h, sh);
                        }
                    }
                } else if (ilv) {
                    // mode 'i': seed j = read bases j, j + s, j + 2s, ... (addPackedPatterns, HashMatcher.cpp:82-96)
                    for (uint32_t g = 0; g < reads.W; g++)
                        ilv_fold_group(__ldg(pl + g * il), __ldg(pl + g * il + 1), is_n ? __ldg(pl + g * il + 2) : 0u, g, j, ilv, seed_len,
                                       P, Q, R, FN);
                } else {
                    for (uint32_t i = 0; i < nch; i++) {
                        const uint32_t bit = j * seed_len + 32 * i;
                        const uint32_t m = (i == nch - 1) ? tail_mask : 0xFFFFFFFFu;
                        const uint32_t l = extract32(pl, reads.W, il, bit) & m;
                        const uint32_t h = extract32(pl + 1, reads.W, il, bit) & m;
                        P ^= l; Q ^= h; R ^= (l & h);
                        if (is_n) FN ^= extract32(pl + 2, reads.W, il, bit) & m;
                    }
                }
                // a seed whose N parity is odd in some rotation class can never collide with an ACGT window
                if (FN == 0) {
                    const uint64_t hv = seed_hash64(P, Q, R);
                    const uint32_t h1 = (uint32_t)hv, h2 = (uint32_t)(hv >> 32);
                    const uint32_t pat = ((r + q.read_base) << reads.part_bits) | j;
                    if (tab.filter && !q.route_world) {
                        const uint32_t f = filter_hash(P, Q, R);
                        if (tab.pair) {
                            atomicOr(tab.filter + (pair_word(P, Q, 0, tab.pair_lo, tab.pair_mask) & tab.filter_mask), filter_bits(f, tab.filter_k));
                            atomicOr(tab.filter + (pair_word(P, Q, 1, tab.pair_lo, tab.pair_mask) & tab.filter_mask), filter_bits(f, tab.filter_k));
                        } else atomicOr(tab.filter + (f & tab.filter_mask), filter_bits(f, tab.filter_k));
                    }
                    n_ins++;
                    if (n_regions) {
                        const uint32_t region = q.route_world ? __umulhi(h1, q.route_world) : h1 >> (32 - q.region_bits);
                        ent = make_uint4(q.route_world ? h1 * q.route_world : h1, h2, pat, region | (atomicAdd(&s_count[region], 1u) << 8));
                    } else {
                        table_insert(tab, h1, h2, pat);
                    }
                }
            }
            if (n_regions) s_ent[j * blockDim.x + threadIdx.x] = ent;
        }
        if (n_regions) {
            // one reservation per region and block, then every thread writes its own entries
            __syncthreads();
            for (uint32_t k = threadIdx.x; k < n_regions; k += blockDim.x)
                s_base[k] = s_count[k] ? atomicAdd(q.count + k, s_count[k]) : 0u;
            __syncthreads();
            for (uint32_t j = 0; j < parts; j++) {
                const uint4 e = s_ent[j * blockDim.x + threadIdx.x];
                if (e.w != 0xFFFFFFFFu) {
                    const uint32_t region = e.w & 0xFFu, slot = s_base[region] + (e.w >> 8);
                    if (slot < q.cap) q.entries[(size_t)region * q.cap + slot] = make_uint4(e.x, e.y, e.z, 0);
                    else if (q.route_world) *q.overflow = 1u;
                    else table_insert(tab, e.x, e.y, e.z);             // queue full (skewed hashes): insert directly
                }
            }
        }
    }
    // one counter update per block
    __shared__ unsigned int blk;
    if (threadIdx.x == 0) blk = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_ins += __shfl_xor_sync(PGM_FULL, n_ins, o);
    if ((threadIdx.x & 31) == 0 && n_ins) atomicAdd(&blk, n_ins);
    __syncthreads();
    if (threadIdx.x == 0 && blk) atomicAdd(inserted, (unsigned long long)blk);
}

// Step 2.  Region after region: CTAs take chunks of the region's queue through a shared cursor and insert; a CTA moves
// on when the region is used up, so at any time the running CTAs touch one or two L2-sized pieces of the bucket array
// (a plain grid-stride loop lets the CTAs drift apart over many regions: twice the DRAM traffic, measured).
#define PGM_INSERT_CHUNK 2048
#define PGM_INSERT_THREADS 256
__global__ void __launch_bounds__(PGM_INSERT_THREADS, 8) build_insert_kernel(TableView tab, BuildQueues q, int prefetch) {
    __shared__ unsigned int s_first[2];
    const uint32_t n_regions = 1u << q.region_bits;
    uint32_t flip = 0;
    for (uint32_t k = 0; k < n_regions; k++) {
        const uint32_t n = min(__ldg(q.count + k), q.cap);
        const uint4 *src = q.entries + (size_t)k * q.cap;
        if (prefetch && k + 1 < n_regions) {
            // pull this CTA's share of the next region's bucket lines into the L2 while this region is being filled
            const uint64_t b0 = ((uint64_t)(k + 1) * tab.n_buckets) >> q.region_bits, b1 = ((uint64_t)(k + 2) * tab.n_buckets) >> q.region_bits;
            const uint64_t lines = (b1 - b0 + 3) / 4;
            for (uint64_t l = (uint64_t)blockIdx.x * PGM_INSERT_THREADS + threadIdx.x; l < lines; l += (uint64_t)gridDim.x * PGM_INSERT_THREADS)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(tab.buckets + b0 + 4 * l));
        }
        for (;;) {
            if (threadIdx.x == 0) s_first[flip] = atomicAdd(q.cursor + k, (unsigned int)PGM_INSERT_CHUNK);
            __syncthreads();                                           // one barrier per chunk: the slot alternates
            const uint32_t first = s_first[flip];
            flip ^= 1;
            if (first >= n) break;
            const uint32_t last = min(first + PGM_INSERT_CHUNK, n);
            for (uint32_t i = first + threadIdx.x; i < last; i += PGM_INSERT_THREADS) {
                const uint4 e = __ldcs(src + i);
                table_insert(tab, e.x, e.y, e.z);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ the scan
template <bool ILV>
struct ScanShared {
    uint32_t lo[2][PGM_BUF_WORDS];
    uint32_t hi[2][PGM_BUF_WORDS];
    uint2 wq[PGM_SCAN_WARPS][PGM_WQ_CAP];   // per-warp candidate queues {pos_in_tile | chain << 31, pattern}
    uint16_t q1[PGM_TILE_POS];              // filter-positive positions of the tile
    uint32_t dl[ILV ? PGM_ILV_WORDS : 1], dh[ILV ? PGM_ILV_WORDS : 1];   // mode 'i': the tile's planes de-interleaved by position residue
    uint64_t bar[2];
    unsigned int q1_count[2], q1_cursor[2], tile[2];
};

// canonical form of the seed window starting at tile position `pos` (text staged in shared memory)
template <int NCH>
__device__ __forceinline__ void window_form(const uint32_t *slo, const uint32_t *shi, uint32_t pos, uint32_t tail_mask,
                                            uint32_t &P, uint32_t &Q, uint32_t &R) {
    const uint32_t wi = pos >> 5, s = pos & 31u;
    P = 0; Q = 0; R = 0;
    uint32_t la = slo[wi], ha = shi[wi];
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        const uint32_t lb = slo[wi + i + 1], hb = shi[wi + i + 1];
        uint32_t l = __funnelshift_r(la, lb, s), h = __funnelshift_r(ha, hb, s);
        if (i == NCH - 1) { l &= tail_mask; h &= tail_mask; }
        P ^= l; Q ^= h; R ^= (l & h);
        la = lb; ha = hb;
    }
}
template <int NCH>
__device__ __forceinline__ uint64_t window_hash(const uint32_t *slo, const uint32_t *shi, uint32_t pos, uint32_t tail_mask) {
    uint32_t P, Q, R;
    window_form<NCH>(slo, shi, pos, tail_mask, P, Q, R);
    return seed_hash64(P, Q, R);
}

// Mismatches of this lane's share of a read record (uint4 #u) against the staged text at bit offset boff.
__device__ __forceinline__ int count_groups(uint32_t u, bool is_n, uint4 v, const uint32_t *blo, const uint32_t *bhi,
                                            uint32_t boff, uint32_t W, uint32_t L) {
    if (u == 0) return 0;
    const uint32_t tw = boff >> 5, ts = boff & 31u;
    int c = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        if (is_n && half) break;
        const uint32_t g = is_n ? u - 1 : 2 * (u - 1) + half;
        if (g >= W) break;
        const uint32_t tl = __funnelshift_r(blo[tw + g], blo[tw + g + 1], ts);
        const uint32_t th = __funnelshift_r(bhi[tw + g], bhi[tw + g + 1], ts);
        uint32_t diff = ((half ? v.z : v.x) ^ tl) | ((half ? v.w : v.y) ^ th);
        if (is_n) diff |= v.z;                              // an N never equals a text symbol
        const uint32_t rem = L - 32 * g;
        if (rem < 32) diff &= (1u << rem) - 1u;
        c += __popc(diff);
    }
    return c;
}

// Persistent CTAs pull 4096-position tiles of the 2-bit text; the next tile's planes (+halos) are staged into the
// other shared-memory buffer with TMA bulk copies while the current one is processed.  Per tile:
//  A1  lane <-> text position: every warp derives, for 32 consecutive window starts at a time, the folded
//      canonical seed form (funnel shifts of broadcast shared-memory words), hashes it and tests the L2-resident
//      filter (one 4-byte gather per position: the L1 wavefront rate bounds this stage); positives are
//      ballot-compacted into a shared list.
//  A2  every lane pulls one positive from the list, rehashes it and fetches its 32-byte table bucket with one
//      256-bit load; tag hits are ballot-compacted into the warp's candidate queue; a bucket without an empty
//      slot sends the lane on along its double-hashing sequence.
//  B   when the warp's queue runs full (and at the end) lane pairs verify one candidate each: the two lanes fetch
//      the two 32-byte halves of the read record (one 64-byte DRAM request), XOR/popcount their 32-base groups
//      against the staged text, add up with one shuffle, and the even lane applies the accept test and
//      atomicMin's the key inside the record.  Restates iterateOver/moveNext (HashMatcher.h:42-68) +
//      executeMatching (ReadsMatchers.cpp:297-341) without their sequential order; the decision is deferred to
//      resolve_kernel.
// MODE 1 is the first stage of the L2-blocked pipeline (pgm_blocked.cuh): A1 as above, then every filter positive is
// hashed and appended to the queue of its table region instead of being probed (ranks within the tile through
// shared-memory counters, one global reservation per region and tile); the per-warp candidate queues' shared memory
// holds the ranks.
// ILV (mode 'i', InterleavedReadsApproxMatcher, ReadsMatchers.cpp:343-409): the seed window at text position x is
// text[x], text[x+s], text[x+2s], ... (s = parts; InterleavedConstantLengthPatternsOnTextHashMatcher, HashMatcher.h:108-135),
// and seed j aligns the read at x - j.  Per tile the staged planes are de-interleaved by position residue into `s` phase
// strings in shared memory (phase r, index q <-> tile position r + q*s): in phase space the window is contiguous again and
// stage A1 / the rehash of A2 run unchanged on it; verification (stage B) uses the original planes.
template <int NCH, bool FAST, int MODE, bool ILV, bool PAIR>
__global__ void __launch_bounds__(PGM_SCAN_THREADS, PGM_SCAN_MIN_CTAS) scan_kernel(const __grid_constant__ ScanParams p) {
    __shared__ __align__(128) ScanShared<ILV> sm;
    if (MODE == 0 && p.only_if != nullptr && *p.only_if == 0) {
        // the pipeline finished without a queue overflow: nothing to redo, its staged counters become final
        if (blockIdx.x == 0 && threadIdx.x < 4) p.counters[threadIdx.x] += p.sq.counters[threadIdx.x];
        return;
    }
    // the warp index through a shuffle: lets the compiler treat it (and every branch on it) as warp-uniform
    const uint32_t t = threadIdx.x, warp = __shfl_sync(PGM_FULL, t >> 5, 0), lane = t & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    unsigned long long n_cand = 0, n_ver = 0, n_acc = 0, n_pos = 0;
    const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
    const bool hints = p.stream_hints != 0;

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
        sm.q1_count[0] = 0; sm.q1_cursor[0] = 0;
        if (tile < p.n_tiles) issue_tile(tile, 0);
    }
    if (MODE == 1)   // per-region counters of the tile being queued (behind the rank storage, see below)
        for (uint32_t k = t; k < PGM_SQ_MAX; k += PGM_SCAN_THREADS) reinterpret_cast<unsigned int *>(&sm.wq[0][0])[PGM_TILE_POS + k] = 0;
    __syncthreads();
    uint32_t parity[2] = {0, 0};
    int buf = 0;
    uint2 *wq = sm.wq[warp];

    for (;;) {
        const unsigned int tile = sm.tile[buf];
        if (tile >= p.n_tiles) break;
        if (t == 0) {
            const unsigned int nxt = atomicAdd(p.tile_counter, 1u);
            sm.tile[buf ^ 1] = nxt;
            sm.q1_count[buf ^ 1] = 0; sm.q1_cursor[buf ^ 1] = 0;
            if (nxt < p.n_tiles) issue_tile(nxt, buf ^ 1);
        }
        mbar_wait(&sm.bar[buf], parity[buf]);
        parity[buf] ^= 1;

        const int64_t tile_word0 = (int64_t)p.first_word + (int64_t)tile * PGM_TILE_WORDS;
        const uint64_t tile_g0 = p.slice_origin + (uint64_t)tile_word0 * 32;   // global position of the tile's first base
        const uint32_t *blo = sm.lo[buf], *bhi = sm.hi[buf];
        const uint32_t *slo = blo + PGM_HALO_L, *shi = bhi + PGM_HALO_L;       // tile position 0

        // ---- A1: hash + filter, lane <-> position
        const int64_t vb64 = (int64_t)p.own_begin - (int64_t)tile_g0, ve64 = (int64_t)p.own_end - (int64_t)tile_g0;
        const uint32_t vb = (uint32_t)max((int64_t)0, min(vb64, (int64_t)PGM_TILE_POS));
        const uint32_t ve = (uint32_t)max((int64_t)0, min(ve64, (int64_t)PGM_TILE_POS));
        // mode 'i': phase r of the tile = bits r, r+s, r+2s, ... of the staged planes, MW words per phase
        const uint32_t ilv_s = ILV ? p.ilv : 1u;
        const uint32_t ilv_qc = (((PGM_TILE_POS + ilv_s - 1) / ilv_s) + 31) >> 5;   // 32-position chunks per phase
        const uint32_t ilv_mw = ilv_qc + NCH + 1;
        if constexpr (ILV) {
            constexpr uint32_t SRC_BITS = (PGM_BUF_WORDS - PGM_HALO_L) * 32;
            const uint32_t per_plane = ilv_s * ilv_mw;
            for (uint32_t w = t; w < 2 * per_plane; w += PGM_SCAN_THREADS) {
                const bool hi_plane = w >= per_plane;
                const uint32_t v = hi_plane ? w - per_plane : w, r = v / ilv_mw, m = v - r * ilv_mw;
                const uint32_t *src = hi_plane ? shi : slo;
                uint32_t word = 0, bit = r + 32 * m * ilv_s;
#pragma unroll 8
                for (int k = 0; k < 32; k++, bit += ilv_s)
                    if (bit < SRC_BITS) word |= ((src[bit >> 5] >> (bit & 31u)) & 1u) << k;
                (hi_plane ? sm.dh : sm.dl)[v] = word;
            }
            __syncthreads();
            for (uint32_t c = warp; c < ilv_s * ilv_qc; c += PGM_SCAN_WARPS) {
                const uint32_t r = c / ilv_qc, q = (c - r * ilv_qc) * 32 + lane;
                const uint32_t pos = r + q * ilv_s;
                uint32_t P, Q, R;
                window_form<NCH>(sm.dl + r * ilv_mw, sm.dh + r * ilv_mw, q, p.tail_mask, P, Q, R);
                const uint32_t f = filter_hash(P, Q, R);
                const uint32_t fm = filter_bits(f, p.tab.filter_k);
                const uint32_t fidx = f & p.tab.filter_mask;
                const bool mine = (fidx >> p.tab.slice_shift) == p.slice;
                const uint32_t fw = !p.tab.filter ? (p.slice == 0 ? 0xFFFFFFFFu : 0u) : (mine ? ld_u32_hint(p.tab.filter + fidx, pol_keep) : 0u);
                const bool hit = pos < PGM_TILE_POS && pos >= vb && pos < ve && (fw & fm) == fm;
                const uint32_t bal = __ballot_sync(PGM_FULL, hit);
                if (bal) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&sm.q1_count[buf], __popc(bal));
                    base = __shfl_sync(PGM_FULL, base, 0);
                    if (hit) sm.q1[base + __popc(bal & lt_mask)] = (uint16_t)pos;
                    if (lane == 0) n_pos += __popc(bal);
                }
            }
        } else if constexpr (MODE == 2) {
            // the windows with a table hit are already known (pgm_part.cuh): A1 is a read of this warp's 16 words of the bitmap
            const uint32_t w_first = warp * PGM_WORDS_PER_WARP;
            const uint32_t mine = lane < PGM_WORDS_PER_WARP ? __ldg(p.hit_bits + (size_t)tile * PGM_TILE_WORDS + w_first + lane) : 0u;
            if (__ballot_sync(PGM_FULL, mine != 0u)) {
#pragma unroll 1
                for (uint32_t it = 0; it < PGM_WORDS_PER_WARP; it++) {
                    const uint32_t bits = __shfl_sync(PGM_FULL, mine, it);
                    if (!bits) continue;                                        // (warp-uniform)
                    const uint32_t pos = (w_first + it) * 32 + lane;
                    const bool hit = pos >= vb && pos < ve && ((bits >> lane) & 1u);
                    const uint32_t bal = __ballot_sync(PGM_FULL, hit);
                    if (bal) {
                        uint32_t base = 0;
                        if (lane == 0) base = atomicAdd(&sm.q1_count[buf], __popc(bal));
                        base = __shfl_sync(PGM_FULL, base, 0);
                        if (hit) sm.q1[base + __popc(bal & lt_mask)] = (uint16_t)pos;
                        if (lane == 0) n_pos += __popc(bal);
                    }
                }
            }
        } else {
            constexpr int U = PGM_A1_UNROLL;
            const uint32_t w_first = warp * PGM_WORDS_PER_WARP;
#pragma unroll 1
            for (uint32_t it0 = 0; it0 < PGM_WORDS_PER_WARP; it0 += U) {
                const uint32_t pos0 = (w_first + it0) * 32;
                // slot u of a lane: position pos0 + 32u + lane, or — paired lookups — slots 2v, 2v+1 are the two windows
                // pos0 + 64v + 2 lane (+1), which share one filter word and one gather
                constexpr bool pair = PAIR;
                auto pos_of = [&](int u) -> uint32_t { return pair ? pos0 + 64 * (u >> 1) + 2 * lane + (u & 1) : pos0 + 32 * u + lane; };
                uint32_t fm[U], fi[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    uint32_t P, Q, R;
                    window_form<NCH>(slo, shi, pos_of(u), p.tail_mask, P, Q, R);
                    const uint32_t f = filter_hash(P, Q, R);
                    fi[u] = (pair ? pair_word(P, Q, 0, p.tab.pair_lo, p.tab.pair_mask) : f) & p.tab.filter_mask;
                    fm[u] = filter_bits(f, p.tab.filter_k);
                }
                uint32_t fw[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (pair && (u & 1)) fw[u] = fw[u - 1];
                    else if (!p.tab.filter) fw[u] = 0xFFFFFFFFu;
                    else fw[u] = (fi[u] >> p.tab.slice_shift) == p.slice ? ld_u32_hint(p.tab.filter + fi[u], pol_keep) : 0u;   // evict_last: the filter('s slice) is the one L2-resident structure
                }
                uint32_t bal[U], tot = 0;
                bool hit[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint32_t pos = pos_of(u);
                    hit[u] = pos >= vb && pos < ve && (fw[u] & fm[u]) == fm[u];
                    bal[u] = __ballot_sync(PGM_FULL, hit[u]);
                    tot += __popc(bal[u]);
                }
                if (tot) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&sm.q1_count[buf], tot);
                    base = __shfl_sync(PGM_FULL, base, 0);
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        if (hit[u]) sm.q1[base + __popc(bal[u] & lt_mask)] = (uint16_t)pos_of(u);
                        base += __popc(bal[u]);
                    }
                    if (lane == 0) n_pos += tot;
                }
            }
        }
        __syncthreads();

        if constexpr (MODE == 1) {
            static_assert(sizeof(sm.wq) >= PGM_TILE_POS * 4 + 2 * PGM_SQ_MAX * 4, "rank storage aliases the candidate queues");
            uint32_t *slot = reinterpret_cast<uint32_t *>(&sm.wq[0][0]);            // per positive: region | rank << 8
            unsigned int *cnt = reinterpret_cast<unsigned int *>(slot + PGM_TILE_POS), *rbase = cnt + PGM_SQ_MAX;
            const uint32_t q1n = sm.q1_count[buf];
            const uint32_t n_regions = 1u << p.sq.region_bits, rshift = 32u - p.sq.region_bits;
            for (uint32_t i = t; i < q1n; i += PGM_SCAN_THREADS) {
                const uint32_t h1 = (uint32_t)window_hash<NCH>(slo, shi, sm.q1[i], p.tail_mask);
                const uint32_t region = h1 >> rshift;
                slot[i] = region | (atomicAdd(&cnt[region], 1u) << 8);
            }
            __syncthreads();
            for (uint32_t k = t; k < n_regions; k += PGM_SCAN_THREADS) {
                const unsigned int c = cnt[k];
                rbase[k] = c ? atomicAdd(p.sq.pos_count + k, c) : 0u;
                cnt[k] = 0;
            }
            __syncthreads();
            bool over = false;
            for (uint32_t i = t; i < q1n; i += PGM_SCAN_THREADS) {
                const uint32_t ppos = sm.q1[i];
                const uint64_t hv = window_hash<NCH>(slo, shi, ppos, p.tail_mask);
                const uint32_t s = slot[i], region = s & 0xFFu, idx = rbase[region] + (s >> 8);
                if (idx < p.sq.pos_cap)
                    __stcs(p.sq.pos_entries + (size_t)region * p.sq.pos_cap + idx,
                           make_uint4((uint32_t)hv, (uint32_t)(hv >> 32), tile * PGM_TILE_POS + ppos, 0u));
                else over = true;
            }
            if (over) *p.sq.overflow = 1u;
            __syncthreads();
            buf ^= 1;
            continue;
        }

        // ---- A2 + B, warp-autonomous
        const uint32_t q1n = sm.q1_count[buf];
        const uint32_t pmask = (1u << p.reads.part_bits) - 1u;
        uint32_t wcount = 0;
        bool exhausted = false, act = false;
        uint32_t ppos = 0, ptag = 0, pb = 0, pstep = 0;
        for (;;) {
            // produce: one lane per filter-positive position, one 256-bit load per probed bucket
            while (wcount + PGM_WQ_ROUND <= PGM_WQ_CAP) {
                // idle lanes (probe sequence finished) take the next positives from the list
                if (!exhausted) {
                    const uint32_t idle = __ballot_sync(PGM_FULL, !act);
                    if (idle) {
                        const uint32_t n_idle = __popc(idle);
                        uint32_t base = 0;
                        if (lane == 0) base = atomicAdd(&sm.q1_cursor[buf], n_idle);
                        base = __shfl_sync(PGM_FULL, base, 0);
                        const uint32_t mine = base + __popc(idle & lt_mask);
                        if (!act && mine < q1n) {
                            act = true;
                            ppos = sm.q1[mine];
                            uint64_t hv;
                            if constexpr (ILV) {
                                const uint32_t q = ppos / ilv_s, r = ppos - q * ilv_s;
                                hv = window_hash<NCH>(sm.dl + r * ilv_mw, sm.dh + r * ilv_mw, q, p.tail_mask);
                            } else hv = window_hash<NCH>(slo, shi, ppos, p.tail_mask);
                            const uint32_t h1 = (uint32_t)hv, h2 = (uint32_t)(hv >> 32);
                            ptag = seed_tag(h2);
                            pb = __umulhi(h1, p.tab.n_buckets);
                            pstep = 1u + __umulhi(h2 * 0x9E3779B1u, p.tab.n_buckets - 1u);
                        }
                        if (base + n_idle >= q1n) exhausted = true;
                    }
                }
                if (!__ballot_sync(PGM_FULL, act)) break;
                u32x8 s;
#pragma unroll
                for (int k = 0; k < 8; k++) s.v[k] = 0;
                if (act) s = hints ? ld256_stream_hint(p.tab.buckets + pb, pol_stream) : ld256_stream(p.tab.buckets + pb);
                bool m[4], em = false;
                uint32_t bal[4], mine = 0, tot = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    m[k] = act && (s.v[2 * k + 1] & 0x7FFFFFFFu) == ptag;
                    em |= s.v[2 * k + 1] == 0xFFFFFFFFu;
                    bal[k] = __ballot_sync(PGM_FULL, m[k]);
                    mine += __popc(bal[k] & lt_mask);
                    tot += __popc(bal[k]);
                }
                if (tot) {
                    uint32_t off = wcount + mine;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (m[k]) wq[off++] = make_uint2(ppos | (s.v[2 * k + 1] & 0x80000000u), s.v[2 * k]);
                    wcount += tot;
                }
                if (em) act = false;                     // a bucket with an empty slot ends the probe sequence
                else if (act) { pb += pstep; if (pb >= p.tab.n_buckets) pb -= p.tab.n_buckets; }
            }
            __syncwarp();
            const uint32_t half = lane & 1u;
            if constexpr (FAST) {
                // consume (ACGT set, 64-byte records): every lane owns one candidate; the two lanes of a pair fetch the
                // even lane's record together (one instruction = one 64-byte request), then the odd lane's, and swap
                // the halves they fetched for each other.
                for (uint32_t i0 = 0; i0 < wcount; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    bool on = i < wcount;
                    const uint2 e = on ? wq[i] : make_uint2(0, 0);
                    const uint32_t cpos = e.x & 0x7FFFFFFFu;
                    const bool chain = (e.x >> 31) != 0;
                    uint32_t cpat = e.y;
                    const uint32_t pat_o = __shfl_xor_sync(PGM_FULL, cpat, 1);
                    const bool on_o = __shfl_xor_sync(PGM_FULL, (int)on, 1) != 0;
                    const uint32_t patA = half ? pat_o : cpat, patB = half ? cpat : pat_o;
                    const bool onA = half ? on_o : on, onB = half ? on : on_o;
                    const u32x8 *recA = reinterpret_cast<const u32x8 *>(p.reads.lq) + (size_t)(patA >> p.reads.part_bits) * 2 + half;
                    const u32x8 *recB = reinterpret_cast<const u32x8 *>(p.reads.lq) + (size_t)(patB >> p.reads.part_bits) * 2 + half;
                    u32x8 rA, rB;
#pragma unroll
                    for (int k = 0; k < 8; k++) { rA.v[k] = 0; rB.v[k] = 0; }
                    if (onA) rA = hints ? ld256_stream_hint(recA, pol_stream) : ld256_cg(recA);
                    if (onB) rB = hints ? ld256_stream_hint(recB, pol_stream) : ld256_cg(recB);
                    uint32_t s0[8], s1[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const uint32_t got = __shfl_xor_sync(PGM_FULL, half ? rA.v[k] : rB.v[k], 1);
                        s0[k] = half ? got : rA.v[k];      // first 32 bytes of this lane's record: header, groups 0-1
                        s1[k] = half ? rB.v[k] : got;      // second 32 bytes: groups 2-5
                    }
                    uint32_t cr = cpat >> p.reads.part_bits, cj = cpat & pmask;
                    for (;;) {
                        const uint32_t L = p.reads.read_len;
                        const uint32_t shift = on ? cj * p.shift_unit : 0u;
                        // tile-relative alignment: always inside the staged buffer, so the count needs no validity checks
                        const uint32_t boff = PGM_HALO_L * 32 + cpos - shift;
                        const uint32_t tw = boff >> 5, ts = boff & 31u;
                        int c = 0;
#pragma unroll
                        for (int g = 0; g < 6; g++) {
                            if ((uint32_t)g < p.reads.W) {
                                const uint32_t rl = g < 2 ? s0[4 + 2 * g] : s1[2 * (g - 2)], rh = g < 2 ? s0[5 + 2 * g] : s1[2 * (g - 2) + 1];
                                const uint32_t tl = __funnelshift_r(blo[tw + g], blo[tw + g + 1], ts);
                                const uint32_t th = __funnelshift_r(bhi[tw + g], bhi[tw + g + 1], ts);
                                uint32_t diff = (rl ^ tl) | (rh ^ th);
                                const uint32_t rem = L - 32 * g;
                                if (rem < 32) diff &= (1u << rem) - 1u;
                                c += __popc(diff);
                            }
                        }
                        if (on) {
                            // body of DefaultReadsApproxMatcher::executeMatching (ReadsMatchers.cpp:301-331) up to the decision
                            n_cand++;
                            const uint32_t st_lo = s0[0], st_hi = s0[1];
                            const uint32_t c_in = st_hi >> 24;
                            const uint64_t gpos = tile_g0 + cpos;
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
                                            const long long seen = (long long)(((uint64_t)s0[3] << 32) | s0[2]);   // never below the live value
                                            if (key < seen) atomicMin(reinterpret_cast<long long *>(p.reads.lq + (size_t)cr * 4) + 1, key);
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
                        // hot keys: walk the chain behind the slot (this lane alone fetches the next record)
                        if (on && chain) {
                            cpat = __ldg(p.tab.next + cpat);
                            on = cpat != PGM_NIL;
                        } else on = false;
                        if (!__ballot_sync(PGM_FULL, on)) break;
                        cr = cpat >> p.reads.part_bits; cj = cpat & pmask;
                        if (on) {
                            const u32x8 *rec = reinterpret_cast<const u32x8 *>(p.reads.lq) + (size_t)cr * 2;
                            const u32x8 x0 = ld256_cg(rec), x1 = ld256_cg(rec + 1);
#pragma unroll
                            for (int k = 0; k < 8; k++) { s0[k] = x0.v[k]; s1[k] = x1.v[k]; }
                        }
                    }
                }
            } else {
                // consume: one lane pair per candidate; each lane fetches one 32-byte half of the 64-byte record
                    for (uint32_t i0 = 0; i0 < wcount; i0 += 16) {
                    const uint32_t i = i0 + (lane >> 1);
                    bool on = i < wcount;
                    const uint2 e = on ? wq[i] : make_uint2(0, 0);
                    const uint32_t cpos = e.x & 0x7FFFFFFFu;
                    const bool chain = (e.x >> 31) != 0;
                    uint32_t cpat = e.y;
                    bool again;
                    do {
                        const uint32_t cr = cpat >> p.reads.part_bits, cj = cpat & pmask;
                        uint32_t cs16; bool isn;
                        uint4 *crec = record_of(p.reads, cr, cs16, isn);
                        u32x8 v;
    #pragma unroll
                        for (int k = 0; k < 8; k++) v.v[k] = 0;
                        if (on) v = hints ? ld256_stream_hint(reinterpret_cast<const u32x8 *>(crec) + half, pol_stream)
                                          : ld256_cg(reinterpret_cast<const u32x8 *>(crec) + half);
                        // tile-relative alignment: always inside the staged buffer, so the count needs no validity checks
                        const uint32_t shift = cj * p.shift_unit;
                        const uint32_t boff = PGM_HALO_L * 32 + cpos - (on ? shift : 0u);
                        const uint32_t L = p.reads.read_len, W = p.reads.W;
                        int c = count_groups(2 * half, isn, make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]), blo, bhi, boff, W, L)
                              + count_groups(2 * half + 1, isn, make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]), blo, bhi, boff, W, L);
                        for (uint32_t u = 4 + 2 * half; u < cs16; u += 4) {      // records longer than 64 bytes
                            uint4 x0 = make_uint4(0, 0, 0, 0), x1 = x0;
                            if (on) { x0 = __ldcg(crec + u); x1 = __ldcg(crec + u + 1); }
                            c += count_groups(u, isn, x0, blo, bhi, boff, W, L) + count_groups(u + 1, isn, x1, blo, bhi, boff, W, L);
                        }
                        c += __shfl_xor_sync(PGM_FULL, c, 1);
                        if (half == 0 && on) {
                            // body of DefaultReadsApproxMatcher::executeMatching (ReadsMatchers.cpp:301-331) up to the decision
                            n_cand++;
                            const uint32_t st_lo = v.v[0], st_hi = v.v[1];
                            const uint32_t c_in = st_hi >> 24;
                            const uint64_t gpos = tile_g0 + cpos;
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
                                            const long long seen = (long long)(((uint64_t)v.v[3] << 32) | v.v[2]);   // never below the live value
                                            if (key < seen) atomicMin(reinterpret_cast<long long *>(crec) + 1, key);
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
                        // hot keys: walk the chain behind the slot
                        if (on && chain) {
                            cpat = __ldg(p.tab.next + cpat);
                            on = cpat != PGM_NIL;
                        } else on = false;
                        again = __ballot_sync(PGM_FULL, on) != 0;
                    } while (again);
                }
            }
            __syncwarp();
            wcount = 0;
            if (exhausted && !__ballot_sync(PGM_FULL, act)) break;
        }
        __syncthreads();
        buf ^= 1;
    }

    // counters: warp reduce, one atomic per warp
    unsigned long long cv[4] = {n_cand, n_ver, n_acc, n_pos};
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cv[k] += __shfl_xor_sync(PGM_FULL, cv[k], o);
        if (lane == 0 && cv[k]) atomicAdd((MODE == 1 ? p.sq.counters : p.counters) + k, cv[k]);
    }
}

// ------------------------------------------------------------------------------------------ per-pass decision
// Applies the reference's sequential accept rule of one pass to the order-free accumulators (SURVEY.md
// §8(a)-R, validated against the reference's classes):  with stored state (X, c_in) the winner is the
// smallest (class, scan order) among the events with c < c_in that report a position != X, joined —
// only after the earliest of those — by the events of the alignment that reports X itself
// (ReadsMatchers.cpp:313 skips them while X is still stored).  class = 0 for c <= minMismatches (the
// reference stops updating a read there, :304), else c.
// FIN: the last pass of a matcher call also writes the three archive-visible arrays and the histogram
// (what finalize_kernel does), saving one sweep over the records.
template <bool FIN>
__global__ void resolve_kernel(ReadsView reads, PerRead pr, uint32_t n_reads, uint64_t pg_len, uint32_t seed_len,
                               uint32_t parts, uint32_t max_mm, uint32_t min_mm, int rev_mode,
                               unsigned long long *__restrict__ out_pos, uint8_t *__restrict__ out_rc,
                               uint8_t *__restrict__ out_mm, unsigned long long *hist /*[257]*/) {
    __shared__ unsigned int sh[256];
    if (FIN) {
        sh[threadIdx.x] = 0;
        __syncthreads();
    }
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_reads) {
        const uint32_t read_len = reads.read_len;
        uint32_t stride16; bool is_n;
        uint4 *rec = record_of(reads, r, stride16, is_n);
        const uint4 h = __ldcg(rec);
        const unsigned long long st = ((unsigned long long)h.y << 32) | h.x;
        long long best = (long long)(((unsigned long long)h.w << 32) | h.z);
        const int touched = *pr.touched;       // reset by the host after this kernel
        long long o1 = PGM_KEY_INF;
        int mask = 0;
        uint32_t cx = 255;
        if (touched) {
            o1 = pr.first_other_order[r]; mask = pr.same_pos_mask[r]; cx = pr.same_pos_mm[r];
            if (o1 != PGM_KEY_INF) pr.first_other_order[r] = PGM_KEY_INF;
            if (mask) { pr.same_pos_mask[r] = 0; pr.same_pos_mm[r] = 255; }
        }
        unsigned long long new_st = st;
        if (best != PGM_KEY_INF || mask != 0) {
            const uint32_t c_in = (uint32_t)(st >> 56);
            if (c_in > min_mm) {
                const int limit = c_in != 255u ? (int)c_in - 1 : (int)max_mm;
                if (mask != 0 && (int)cx <= limit && o1 != PGM_KEY_INF) {
                    const uint64_t X = st & PGM_POS_MASK;
                    const uint64_t aX = rev_mode ? pg_len - X - read_len : X;   // that alignment in this pass's coordinates
                    for (uint32_t j = 0; j < parts; j++) {
                        if (!((mask >> j) & 1)) continue;
                        const unsigned long long order = ((aX + (uint64_t)j * seed_len) << 8) | (unsigned long long)(parts - 1 - j);
                        if ((long long)order > o1) {
                            const unsigned long long cls = cx <= min_mm ? 0ull : (unsigned long long)cx;
                            const long long key = (long long)((cls << 56) | (order << 8) | cx);
                            best = min(best, key);
                            break;
                        }
                    }
                }
                if (best != PGM_KEY_INF) {
                    const uint32_t c = (uint32_t)(best & 0xFF);
                    const uint32_t jj = (uint32_t)((best >> 8) & 0xFF);
                    const uint64_t g = ((unsigned long long)best >> 16) & PGM_POS_MASK;
                    const uint64_t a = g - (uint64_t)(parts - 1 - jj) * seed_len;
                    const uint64_t rep = rev_mode ? pg_len - (a + read_len) : a;
                    new_st = ((unsigned long long)c << 56) | ((unsigned long long)(rev_mode ? 1 : 0) << 55) | rep;
                }
            }
            rec[0] = make_uint4((uint32_t)new_st, (uint32_t)(new_st >> 32), 0xFFFFFFFFu, 0x7FFFFFFFu);
        }
        if (FIN) {
            const uint32_t c = (uint32_t)(new_st >> 56);
            out_pos[r] = c == 255u ? 0xFFFFFFFFFFFFFFFFull : (new_st & PGM_POS_MASK);
            out_rc[r] = (uint8_t)((new_st >> 55) & 1);
            out_mm[r] = (uint8_t)c;
            atomicAdd(&sh[c], 1u);
        }
    }
    if (FIN) {
        __syncthreads();
        if (sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
    }
}

// state -> the three archive-visible arrays (+ matched count and per-mismatch histogram)
__global__ void finalize_kernel(ReadsView reads, uint32_t n_reads, unsigned long long *__restrict__ out_pos,
                                uint8_t *__restrict__ out_rc, uint8_t *__restrict__ out_mm, unsigned long long *hist /*[257]*/) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_reads) {
        uint32_t stride16; bool is_n;
        const uint2 h = __ldcg(reinterpret_cast<const uint2 *>(record_of(reads, r, stride16, is_n)));
        const unsigned long long st = ((unsigned long long)h.y << 32) | h.x;
        const uint32_t c = (uint32_t)(st >> 56);
        out_pos[r] = c == 255u ? 0xFFFFFFFFFFFFFFFFull : (st & PGM_POS_MASK);
        out_rc[r] = (uint8_t)((st >> 55) & 1);
        out_mm[r] = (uint8_t)c;
        atomicAdd(&sh[c], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------ mismatch lists (export)
// What AbstractReadsApproxMatcher::updateEntry recomputes on the host for every matched read (ReadsMatchers.cpp:555-566:
// getRead + reverseComplementInPlace + fillEntryWithMismatches, :40-52): the read's mismatches against the pseudogenome at
// readMatchPos, the read taken reverse-complemented when readMatchRC, in ascending offset.  Three small kernels: per-block
// sums of the mismatch counts, an exclusive scan of those sums, and the emit pass (thread per read: XOR of the record's
// planes with the text window; an RC match is compared on the reverse-complement planes, where the read lies as it is
// stored, and its offsets / symbols are mapped back — offset L-1-q, complemented symbols).
#define PGM_MIS_THREADS 256
__global__ void __launch_bounds__(PGM_MIS_THREADS) mismatch_count_kernel(ReadsView reads, uint32_t n_reads, unsigned long long *block_sums) {
    __shared__ unsigned int s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const uint32_t r = blockIdx.x * PGM_MIS_THREADS + threadIdx.x;
    unsigned int c = 0;
    if (r < n_reads) {
        uint32_t stride16; bool is_n;
        const uint2 h = __ldcg(reinterpret_cast<const uint2 *>(record_of(reads, r, stride16, is_n)));
        c = h.y >> 24;
        if (c == 255u) c = 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(PGM_FULL, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_sum, c);
    __syncthreads();
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s_sum;
}

// in-place exclusive scan of block_sums by one block; *total = the sum
__global__ void __launch_bounds__(1024) mismatch_scan_kernel(unsigned long long *block_sums, uint32_t n_blocks, unsigned long long *total) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    if (t == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += 1024) {
        const uint32_t i = b0 + t;
        const unsigned long long v = i < n_blocks ? block_sums[i] : 0ull;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(PGM_FULL, x, o);
            if (lane >= (uint32_t)o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(PGM_FULL, w, o);
                if (lane >= (uint32_t)o) w += y;
            }
            s_warp[lane] = w;                                   // inclusive over warps
        }
        __syncthreads();
        const unsigned long long before = s_carry + (warp ? s_warp[warp - 1] : 0ull) + (x - v);
        if (i < n_blocks) block_sums[i] = before;
        __syncthreads();
        if (t == 1023) s_carry = before + v;
        __syncthreads();
    }
    if (t == 0) *total = s_carry;
}

struct MismatchParams {
    const uint32_t *flo, *fhi, *rlo, *rhi;   // forward / reverse-complement planes of the WHOLE text, origin at word 0
    uint64_t pg_len;
    ReadsView reads;
    uint32_t n_reads;
    const unsigned long long *block_base;     // exclusive prefix of the per-block counts
    unsigned long long *out_offsets;          // n_reads + 1
    uint8_t *out_pos, *out_syms;              // out_syms = pgSymbol:2 | readSymbol:3 << 2 (A C G T = 0..3, N = 4)
    unsigned long long capacity;
    int *err;                                 // set when a list does not have readMismatchesCount entries (internal error)
};

__global__ void __launch_bounds__(PGM_MIS_THREADS) mismatch_emit_kernel(const __grid_constant__ MismatchParams p) {
    __shared__ unsigned int s_warp[PGM_MIS_THREADS / 32];
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t r = blockIdx.x * PGM_MIS_THREADS + t;
    const uint32_t L = p.reads.read_len, W = p.reads.W;
    uint32_t c = 0, stride16 = 4;
    bool is_n = false, rc = false;
    uint64_t pos = 0;
    const uint4 *rec = nullptr;
    if (r < p.n_reads) {
        rec = record_of(p.reads, r, stride16, is_n);
        const uint2 h = __ldcg(reinterpret_cast<const uint2 *>(rec));
        const unsigned long long st = ((unsigned long long)h.y << 32) | h.x;
        c = (uint32_t)(st >> 56);
        rc = ((st >> 55) & 1) != 0;
        pos = st & PGM_POS_MASK;
        if (c == 255u) c = 0;
    }
    // exclusive scan of c within the block
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
    unsigned long long at = p.block_base[blockIdx.x] + wbase + (x - c);
    if (r >= p.n_reads) return;
    p.out_offsets[r] = at;
    if (r == p.n_reads - 1) p.out_offsets[p.n_reads] = at + c;
    if (c == 0) return;
    const uint64_t a = rc ? p.pg_len - pos - L : pos;          // where the stored read lies on this strand's planes
    const uint32_t *tlo = rc ? p.rlo : p.flo, *thi = rc ? p.rhi : p.fhi;
    const uint64_t wi = a >> 5;
    const uint32_t ts = (uint32_t)(a & 31);
    uint32_t emitted = 0;
    for (uint32_t gi = 0; gi < W; gi++) {
        const uint32_t g = rc ? W - 1 - gi : gi;               // ascending offsets of the reverse-complemented read = descending q
        const uint32_t tl = __funnelshift_r(__ldg(tlo + wi + g), __ldg(tlo + wi + g + 1), ts);
        const uint32_t th = __funnelshift_r(__ldg(thi + wi + g), __ldg(thi + wi + g + 1), ts);
        uint32_t rl, rh, nm = 0;
        if (is_n) {
            const uint4 v = __ldcg(rec + 1 + g);
            rl = v.x; rh = v.y; nm = v.z;
        } else {
            const uint4 v = __ldcg(rec + 1 + (g >> 1));
            rl = (g & 1) ? v.z : v.x; rh = (g & 1) ? v.w : v.y;
        }
        uint32_t diff = (rl ^ tl) | (rh ^ th) | nm;
        const uint32_t rem = L - 32 * g;
        if (rem < 32) diff &= (1u << rem) - 1u;
        while (diff) {
            const uint32_t b = rc ? 31u - (uint32_t)__clz(diff) : (uint32_t)__ffs(diff) - 1u;
            diff &= ~(1u << b);
            const uint32_t q = 32 * g + b;
            uint32_t pg_sym = ((tl >> b) & 1u) | (((th >> b) & 1u) << 1);
            uint32_t rd_sym = ((nm >> b) & 1u) ? 4u : (((rl >> b) & 1u) | (((rh >> b) & 1u) << 1));
            uint32_t off = q;
            if (rc) {                                           // back to the reverse-complemented read on the forward text
                off = L - 1 - q;
                pg_sym = 3u - pg_sym;
                if (rd_sym != 4u) rd_sym = 3u - rd_sym;
            }
            if (at + emitted < p.capacity) {
                p.out_pos[at + emitted] = (uint8_t)off;
                p.out_syms[at + emitted] = (uint8_t)(pg_sym | (rd_sym << 2));
            }
            emitted++;
        }
    }
    if (emitted != c) atomicExch(p.err, 2);
}

// Multi-GPU exchange: the per-read keys live inside the records; these two kernels copy them to / from a
// contiguous array that the caller all-reduces (MIN) across the text shards.
__global__ void export_keys_kernel(ReadsView reads, uint32_t n_reads, long long *__restrict__ keys) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    uint32_t stride16; bool is_n;
    const uint4 h = __ldcg(record_of(reads, r, stride16, is_n));
    keys[r] = (long long)(((unsigned long long)h.w << 32) | h.z);
}
__global__ void import_keys_kernel(ReadsView reads, uint32_t n_reads, const long long *__restrict__ keys) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    uint32_t stride16; bool is_n;
    long long *k = reinterpret_cast<long long *>(record_of(reads, r, stride16, is_n)) + 1;
    const long long v = keys[r];
    if (*k != v) *k = v;
}

} // namespace pgm
