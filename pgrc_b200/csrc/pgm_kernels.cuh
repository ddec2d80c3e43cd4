// pgm_kernels.cuh — sm_100a kernels of the read-vs-pseudogenome matcher.
//
// Data layout in HBM
//   text        two bit planes per strand (lo = code&1, hi = code>>1; A=0 C=1 G=2 T=3), 32 bases
//               per uint32 word, base p at bit (p & 31) of word (p >> 5); PGM_PAD_WORDS zero
//               words in front of the origin, a zero tail behind it.  The reverse-complement
//               strand is materialised once (rc = ~bitreverse) so both passes run one kernel.
//   reads       per read, interleaved per 32 bases: {lo, hi} (ACGT set) or {lo, hi, nmask} (ACGNT
//               set), W = ceil(L/32) groups, stride rounded to 16 bytes: fetched with uint4 loads.
//   seed table  open addressing, 32-byte buckets of four 8-byte slots {head:32 | flag:1 tag:31};
//               one slot per distinct seed key, duplicates chained through next[pattern].
//   filter      2^f-bit blocked Bloom filter (2 bits in one word), sized to stay L2-resident.
//   per read    state64 = mm:8 | rc:1 | pos:40 ; accumulators best_key / first_other_order
//               (int64, MIN-mergeable), same_pos_mask (int32), same_pos_mm (uint8).
//
// Seed key.  The reference hashes a seed with CyclicHash<uint32>(n, 32)
// (rollinghash/cyclichash.h:29-35,100-123): symbol k is rotated by (n-1-k) mod 32, so two
// seeds collide for EVERY random table iff, per residue class of k mod 32, they contain each
// symbol with the same parity.  That equivalence is reproduced exactly by
// XOR-folding the window's bit planes into 32-bit words: P = fold(lo), Q = fold(hi),
// R = fold(lo & hi) determine the four parity vectors.  key = mix(P, Q, R).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PGM_TILE_WORDS 256                 // text words per tile == threads per scan CTA
#define PGM_TILE_POS (PGM_TILE_WORDS * 32) // 8192 text positions per tile
#define PGM_HALO_L 12                      // words staged left of a tile  (alignments reach back (parts-1)*n <= 255 bases)
#define PGM_HALO_R 20                      // words staged right of a tile (window + read tail <= 255 + 255 bases)
#define PGM_BUF_WORDS (PGM_HALO_L + PGM_TILE_WORDS + PGM_HALO_R)
#define PGM_PAD_WORDS 64                   // zero words in front of every plane
#define PGM_TAIL_WORDS (PGM_TILE_WORDS + 64)
#define PGM_QCAP 2048                      // per-CTA candidate queue entries
#define PGM_SCAN_THREADS PGM_TILE_WORDS

#define PGM_EMPTY64 0xFFFFFFFFFFFFFFFFull
#define PGM_NIL 0xFFFFFFFFu
#define PGM_KEY_INF 0x7FFFFFFFFFFFFFFFll
#define PGM_POS_MASK 0xFFFFFFFFFFull       // 40-bit positions
#define PGM_STATE_UNMATCHED ((255ull << 56) | PGM_POS_MASK)

namespace pgm {

// ------------------------------------------------------------------------------------------ hashing
__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

// 96-bit canonical seed form -> two independent 32-bit hashes (murmur3-style, shared scrambles).
__host__ __device__ __forceinline__ void seed_hash(uint32_t P, uint32_t Q, uint32_t R, uint32_t &h1, uint32_t &h2) {
    const uint32_t k1 = rotl32(P * 0xcc9e2d51u, 15) * 0x1b873593u;
    const uint32_t k2 = rotl32(Q * 0xcc9e2d51u, 15) * 0x1b873593u;
    const uint32_t k3 = rotl32(R * 0xcc9e2d51u, 15) * 0x1b873593u;
    uint32_t a = 0x9747b28cu, b = 0x3c6ef372u;
    a ^= k1; a = rotl32(a, 13) * 5u + 0xe6546b64u;  b ^= k2; b = rotl32(b, 13) * 5u + 0xe6546b64u;
    a ^= k2; a = rotl32(a, 13) * 5u + 0xe6546b64u;  b ^= k3; b = rotl32(b, 13) * 5u + 0xe6546b64u;
    a ^= k3; a = rotl32(a, 13) * 5u + 0xe6546b64u;  b ^= k1; b = rotl32(b, 13) * 5u + 0xe6546b64u;
    h1 = fmix32(a ^ 12u);
    h2 = fmix32(b ^ 12u);
}

__host__ __device__ __forceinline__ uint32_t seed_tag(uint32_t h2) {
    uint32_t t = h2 & 0x7FFFFFFFu;
    return t == 0x7FFFFFFFu ? 0x7FFFFFFEu : t; // 0x7FFFFFFF is what an empty slot shows
}

__host__ __device__ __forceinline__ uint32_t filter_word(uint32_t h1, uint32_t h2, int word_bits) {
    return ((h1 ^ rotl32(h2, 15)) * 0x9E3779B1u) >> (32 - word_bits);
}
__host__ __device__ __forceinline__ uint32_t filter_mask(uint32_t h1, uint32_t h2) {
    return (1u << (h1 >> 27)) | (1u << (h2 >> 27));
}

// ------------------------------------------------------------------------------------------ parameters
struct TableView {
    unsigned long long *slots;  // 4 slots per bucket
    uint32_t *next;             // chain of patterns sharing a key
    uint32_t *filter;           // may be null
    uint32_t bucket_mask;
    int filter_word_bits;       // log2(#filter words); 0 = no filter
};

struct ReadsView {
    const uint32_t *lq_planes;  // n_lq * lq_stride words
    const uint32_t *n_planes;   // n_n * n_stride words
    uint32_t n_lq, n_n;
    uint32_t lq_stride, n_stride; // words per read
    uint32_t read_len, W;
};

struct PerRead {
    unsigned long long *state;      // mm:8 | rc:1 | pos:40
    long long *best_key;            // cls:8 | txtPos:40 | (parts-1-j):8 | mm:8   (events with rep != stored pos)
    long long *first_other_order;   // txtPos:40 | (parts-1-j):8                 (earliest such event)
    int *same_pos_mask;             // bit j: seed j hit the alignment that reports the stored pos
    uint8_t *same_pos_mm;           // its mismatch count
    int *touched;
};

struct ScanParams {
    const uint32_t *tlo, *thi;      // planes of this pass's text, local origin at word 0
    uint64_t slice_origin;          // global coordinate of local position 0 (this pass's coordinates)
    uint64_t own_begin, own_end;    // owned seed-window starts, global, already clipped to <= pg_len - n + 1
    uint64_t pg_len;
    uint32_t first_word;            // first local word of tile 0 (multiple of 4)
    uint32_t n_tiles;
    uint32_t seed_len, parts, max_mm, min_mm;
    uint32_t tail_mask;             // valid bits of the last 32-base chunk of a seed
    int rev_mode;
    TableView tab;
    ReadsView reads;
    PerRead pr;
    unsigned int *tile_counter;
    unsigned long long *counters;   // [0] candidates [1] verified [2] accepted [3] queue overflows
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

// 32 bits starting at bit position `bit` of a plane whose word k lives at w[k * stride]
__device__ __forceinline__ uint32_t extract32(const uint32_t *w, uint32_t nwords, uint32_t stride, uint32_t bit) {
    const uint32_t k = bit >> 5;
    const uint32_t a = k < nwords ? w[k * stride] : 0u;
    const uint32_t b = k + 1 < nwords ? w[(k + 1) * stride] : 0u;
    return __funnelshift_r(a, b, bit & 31);
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
// Packed reads (reference layout, SymbolsPackingFacility.cpp:147-185) -> interleaved bit planes:
// ACGT set: word 2i = lo bits of bases 32i..32i+31, word 2i+1 = hi bits; ACGNT set: words 3i, 3i+1,
// 3i+2 = lo, hi, N mask (lo = hi = 0 under an N).  One thread per read.
__global__ void unpack_reads_kernel(const uint8_t *__restrict__ packed, uint32_t n_reads, uint32_t read_len,
                                    uint32_t packed_len, int with_n, uint32_t *__restrict__ planes, uint32_t stride,
                                    uint32_t W) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint8_t *src = packed + (size_t)r * packed_len;
    uint32_t *dst = planes + (size_t)r * stride;
    uint32_t lo = 0, hi = 0, nm = 0, wi = 0, bitpos = 0;
    if (!with_n) {
        for (uint32_t b = 0; b < packed_len; b++) {
            const uint32_t v = src[b];
            // bases 4b..4b+3, first base in the two most significant bits; the tail of the last byte is 'A' = 0
            const uint32_t l4 = ((v >> 6) & 1) | (((v >> 4) & 1) << 1) | (((v >> 2) & 1) << 2) | ((v & 1) << 3);
            const uint32_t h4 = ((v >> 7) & 1) | (((v >> 5) & 1) << 1) | (((v >> 3) & 1) << 2) | (((v >> 1) & 1) << 3);
            lo |= l4 << bitpos; hi |= h4 << bitpos;
            bitpos += 4;
            if (bitpos == 32) { dst[2 * wi] = lo; dst[2 * wi + 1] = hi; wi++; lo = hi = 0; bitpos = 0; }
        }
        if (bitpos && wi < W) { dst[2 * wi] = lo; dst[2 * wi + 1] = hi; wi++; }
        for (uint32_t k = 2 * wi; k < stride; k++) dst[k] = 0;
    } else {
        uint32_t p = 0;
        for (uint32_t b = 0; b < packed_len; b++) {
            const uint32_t v = src[b];
            const uint32_t s[3] = {v / 25u, (v / 5u) % 5u, v % 5u};
#pragma unroll
            for (int j = 0; j < 3; j++) {
                if (p < read_len) {
                    const uint32_t sym = s[j];
                    const uint32_t isn = sym == 3u ? 1u : 0u;
                    const uint32_t code = sym == 4u ? 3u : (isn ? 0u : sym);
                    lo |= (code & 1u) << bitpos; hi |= (code >> 1) << bitpos; nm |= isn << bitpos;
                    p++; bitpos++;
                    if (bitpos == 32) {
                        dst[3 * wi] = lo; dst[3 * wi + 1] = hi; dst[3 * wi + 2] = nm;
                        wi++; lo = hi = nm = 0; bitpos = 0;
                    }
                }
            }
        }
        if (bitpos && wi < W) { dst[3 * wi] = lo; dst[3 * wi + 1] = hi; dst[3 * wi + 2] = nm; wi++; }
        for (uint32_t k = 3 * wi; k < stride; k++) dst[k] = 0;
    }
}

// ------------------------------------------------------------------------------------------ per-read state
__global__ void init_state_kernel(PerRead pr, uint32_t n_reads, int reset_state) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    if (reset_state) pr.state[r] = PGM_STATE_UNMATCHED;
    pr.best_key[r] = PGM_KEY_INF;
    pr.first_other_order[r] = PGM_KEY_INF;
    pr.same_pos_mask[r] = 0;
    pr.same_pos_mm[r] = 255;
    if (r == 0) *pr.touched = 0;
}

// ------------------------------------------------------------------------------------------ table build
// One thread per pattern (read r, seed j): canonical key of read bases [j*n, (j+1)*n), insert into the
// open-addressing table.  Restates addReadsSetOfPatterns (ConstantLengthPatternsOnTextHashMatcher.cpp:23-42);
// pattern index = r * parts + j (:39).  Reads already matched with <= min_mm mismatches are left out when
// `continuation` (the matchedReadsBitmap argument, ReadsMatchers.cpp:290-291).
__global__ void build_table_kernel(ReadsView reads, const unsigned long long *__restrict__ state, TableView tab,
                                   uint32_t seed_len, uint32_t parts, uint32_t min_mm, int continuation,
                                   uint32_t tail_mask, unsigned long long *inserted) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n_patterns = (uint64_t)(reads.n_lq + reads.n_n) * parts;
    bool active = p < n_patterns;
    uint32_t r = 0, j = 0;
    if (active) {
        r = (uint32_t)(p / parts);
        j = (uint32_t)(p - (uint64_t)r * parts);
        if (continuation && (uint32_t)(state[r] >> 56) <= min_mm) active = false;
    }
    if (active) {
        const bool is_n = r >= reads.n_lq;
        const uint32_t *pl = is_n ? reads.n_planes + (size_t)(r - reads.n_lq) * reads.n_stride
                                  : reads.lq_planes + (size_t)r * reads.lq_stride;
        const uint32_t il = is_n ? 3u : 2u; // interleave factor
        const uint32_t W = reads.W;
        const uint32_t nch = (seed_len + 31) >> 5;
        uint32_t P = 0, Q = 0, R = 0, FN = 0;
        for (uint32_t i = 0; i < nch; i++) {
            const uint32_t bit = j * seed_len + 32 * i;
            const uint32_t m = (i == nch - 1) ? tail_mask : 0xFFFFFFFFu;
            const uint32_t l = extract32(pl, W, il, bit) & m;
            const uint32_t h = extract32(pl + 1, W, il, bit) & m;
            P ^= l; Q ^= h; R ^= (l & h);
            if (is_n) FN ^= extract32(pl + 2, W, il, bit) & m;
        }
        // a seed whose N parity is odd in some rotation class can never collide with an ACGT window
        if (FN != 0) active = false;
        if (active) {
            uint32_t h1, h2;
            seed_hash(P, Q, R, h1, h2);
            const uint32_t tag = seed_tag(h2);
            const uint32_t pat = (uint32_t)p;
            tab.next[pat] = PGM_NIL;
            uint32_t b = h1 & tab.bucket_mask, s = 0;
            const unsigned long long mine = ((unsigned long long)tag << 32) | pat;
            for (;;) {
                unsigned long long *sl = tab.slots + (size_t)b * 4 + s;
                unsigned long long old = atomicCAS(sl, PGM_EMPTY64, mine);
                if (old == PGM_EMPTY64) break;
                if ((uint32_t)((old >> 32) & 0x7FFFFFFFu) == tag) {
                    // same key already present: become the head of its chain and set the chain flag
                    for (;;) {
                        const unsigned long long nw = (((old >> 32) | 0x80000000ull) << 32) | pat;
                        tab.next[pat] = (uint32_t)old;
                        const unsigned long long prev = atomicCAS(sl, old, nw);
                        if (prev == old) break;
                        old = prev;
                    }
                    break;
                }
                if (++s == 4) { s = 0; b = (b + 1) & tab.bucket_mask; }
            }
            if (tab.filter) atomicOr(tab.filter + filter_word(h1, h2, tab.filter_word_bits), filter_mask(h1, h2));
        }
    }
    const unsigned int cnt = __popc(__ballot_sync(0xFFFFFFFFu, active));
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(inserted, (unsigned long long)cnt);
}

// ------------------------------------------------------------------------------------------ verification
struct ScanShared {
    uint32_t lo[PGM_BUF_WORDS];
    uint32_t hi[PGM_BUF_WORDS];
    uint2 queue[PGM_QCAP];       // {pos_in_tile | chain flag << 31, head pattern}
    uint64_t bar;
    unsigned int q_count;
    unsigned int tile;
};

// Hamming distance between a read (interleaved planes in registers) and the staged text at bit offset boff.
template <int IL>
__device__ __forceinline__ int count_mismatches(const uint32_t *rw, const ScanShared &sm, uint32_t boff, uint32_t W,
                                                uint32_t L) {
    const uint32_t tw = boff >> 5, ts = boff & 31;
    int c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if ((uint32_t)i < W) {
            const uint32_t tl = __funnelshift_r(sm.lo[tw + i], sm.lo[tw + i + 1], ts);
            const uint32_t th = __funnelshift_r(sm.hi[tw + i], sm.hi[tw + i + 1], ts);
            uint32_t diff = (rw[IL * i] ^ tl) | (rw[IL * i + 1] ^ th);
            if (IL == 3) diff |= rw[IL * i + 2];   // an N never equals a text symbol
            const uint32_t rem = L - 32 * i;
            if (rem < 32) diff &= (1u << rem) - 1u;
            c += __popc(diff);
        }
    }
    return c;
}

// Verifies one pattern hit at global text position g (this pass's coordinates).  Restates the body of
// DefaultReadsApproxMatcher::executeMatching (ReadsMatchers.cpp:301-331) up to the decision, which is
// deferred to resolve_kernel so that it does not depend on scan order.
// Returns verified | accepted << 1.
__device__ __forceinline__ uint32_t verify_pattern(const ScanParams &p, const ScanShared &sm, int64_t tile_bit0,
                                                   uint64_t g, uint32_t pat) {
    const uint32_t r = pat / p.parts;
    const uint32_t j = pat - r * p.parts;
    const unsigned long long st = __ldg(p.pr.state + r);
    const uint32_t c_in = (uint32_t)(st >> 56);
    if (c_in <= p.min_mm) return 0;                                 // :304
    const uint32_t shift = j * p.seed_len;
    if (shift > g) return 0;                                        // :308
    const uint64_t a = g - shift;
    const uint32_t L = p.reads.read_len;
    if (a + L > p.pg_len) return 0;                                 // :311
    const uint64_t rep = p.rev_mode ? p.pg_len - (a + L) : a;       // :313,:326 (matchingLength == readLength)
    const bool has_pos = c_in != 255u;
    const bool same_pos = has_pos && ((st & PGM_POS_MASK) == rep);  // coordinate-only compare, :313
    int limit = has_pos ? (int)c_in - 1 : (int)p.max_mm;            // :315
    const unsigned long long order = (g << 8) | (unsigned long long)(p.parts - 1 - j);
    if (!has_pos) {
        // nothing stored yet: only the minimum key matters, so an event that cannot beat the current
        // minimum need not be verified (a stale read is safe: the key only ever decreases)
        const long long k = *(volatile long long *)(p.pr.best_key + r);
        if (k != PGM_KEY_INF) {
            const int kcls = (int)(k >> 56);
            const unsigned long long korder = ((unsigned long long)k >> 8) & 0xFFFFFFFFFFFFull;
            const int T = kcls - (order >= korder ? 1 : 0);
            if (T < 0) return 0;
            limit = min(limit, max(T, (int)p.min_mm));
        }
    }
    if (limit < 0) return 0;
    const uint32_t boff = (uint32_t)((int64_t)(a - p.slice_origin) - tile_bit0);
    const uint32_t W = p.reads.W;
    int c;
    if (r < p.reads.n_lq) {
        const uint4 *pl4 = reinterpret_cast<const uint4 *>(p.reads.lq_planes + (size_t)r * p.reads.lq_stride);
        const uint32_t nvec = p.reads.lq_stride >> 2;
        uint32_t rw[16];
#pragma unroll
        for (int v = 0; v < 4; v++) {
            uint4 q = make_uint4(0, 0, 0, 0);
            if ((uint32_t)v < nvec) q = __ldg(pl4 + v);
            rw[4 * v] = q.x; rw[4 * v + 1] = q.y; rw[4 * v + 2] = q.z; rw[4 * v + 3] = q.w;
        }
        c = count_mismatches<2>(rw, sm, boff, W, L);
    } else {
        const uint4 *pl4 = reinterpret_cast<const uint4 *>(p.reads.n_planes + (size_t)(r - p.reads.n_lq) * p.reads.n_stride);
        const uint32_t nvec = p.reads.n_stride >> 2;
        uint32_t rw[24];
#pragma unroll
        for (int v = 0; v < 6; v++) {
            uint4 q = make_uint4(0, 0, 0, 0);
            if ((uint32_t)v < nvec) q = __ldg(pl4 + v);
            rw[4 * v] = q.x; rw[4 * v + 1] = q.y; rw[4 * v + 2] = q.z; rw[4 * v + 3] = q.w;
        }
        c = count_mismatches<3>(rw, sm, boff, W, L);
    }
    if (c > limit) return 1;
    if (!same_pos) {
        const unsigned long long cls = (uint32_t)c <= p.min_mm ? 0ull : (unsigned long long)c;
        const long long key = (long long)((cls << 56) | (order << 8) | (unsigned long long)c);
        atomicMin(p.pr.best_key + r, key);
        if (has_pos) {
            atomicMin(p.pr.first_other_order + r, (long long)order);
            *p.pr.touched = 1;
        }
    } else {
        atomicOr(p.pr.same_pos_mask + r, 1 << j);
        p.pr.same_pos_mm[r] = (uint8_t)c;
        *p.pr.touched = 1;
    }
    return 3;
}

// One queue entry = one table slot hit; walks the chain of patterns sharing the key.
// Returns candidates | verified << 10 | accepted << 20 (saturating is irrelevant: summed in 64 bits by the caller).
__device__ __noinline__ uint3 verify_entry(const ScanParams &p, const ScanShared &sm, int64_t tile_bit0,
                                           uint64_t tile_g0, uint2 e) {
    const uint64_t g = tile_g0 + (e.x & 0x7FFFFFFFu);
    uint32_t pat = e.y;
    const bool chained = (e.x >> 31) != 0;
    uint3 cnt = make_uint3(0, 0, 0);
    for (;;) {
        const uint32_t v = verify_pattern(p, sm, tile_bit0, g, pat);
        cnt.x++; cnt.y += v & 1; cnt.z += v >> 1;
        if (!chained) break;
        pat = __ldg(p.tab.next + pat);
        if (pat == PGM_NIL) break;
    }
    return cnt;
}

// ------------------------------------------------------------------------------------------ the scan
// Persistent CTAs pull 8192-position tiles of the 2-bit text.  Per tile: (1) one elected thread stages
// the two planes (+halos) into shared memory with TMA bulk copies; (2) every thread owns one 32-base word
// and derives, for its 32 window starts, the folded canonical seed form with funnel shifts, hashes it,
// tests the L2-resident filter and probes one 32-byte table bucket; hits go to a shared queue;
// (3) the queue is drained one candidate per thread: XOR/popcount of the read planes against the staged
// text, then atomicMin on the read's key.  Restates iterateOver/moveNext (HashMatcher.h:42-68) +
// executeMatching (ReadsMatchers.cpp:297-341) without their sequential order.
template <int NCH>
__global__ void __launch_bounds__(PGM_SCAN_THREADS, 3) scan_kernel(const __grid_constant__ ScanParams p) {
    __shared__ __align__(128) ScanShared sm;
    const uint32_t t = threadIdx.x;
    unsigned long long n_cand = 0, n_ver = 0, n_acc = 0, n_ovf = 0;
    if (t == 0) {
        mbar_init(&sm.bar, 1);
        sm.q_count = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    const uint4 *slots4 = reinterpret_cast<const uint4 *>(p.tab.slots);

    for (;;) {
        if (t == 0) {
            const unsigned int tile = atomicAdd(p.tile_counter, 1u);
            sm.tile = tile;
            if (tile < p.n_tiles) {
                const int64_t w0 = (int64_t)p.first_word + (int64_t)tile * PGM_TILE_WORDS - PGM_HALO_L;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&sm.bar, 2 * PGM_BUF_WORDS * 4);
                bulk_g2s(sm.lo, p.tlo + w0, PGM_BUF_WORDS * 4, &sm.bar);
                bulk_g2s(sm.hi, p.thi + w0, PGM_BUF_WORDS * 4, &sm.bar);
            }
        }
        __syncthreads();
        const unsigned int tile = sm.tile;
        if (tile >= p.n_tiles) break;
        mbar_wait(&sm.bar, parity);
        parity ^= 1;

        const int64_t tile_word0 = (int64_t)p.first_word + (int64_t)tile * PGM_TILE_WORDS;
        const uint64_t tile_g0 = p.slice_origin + (uint64_t)tile_word0 * 32;   // global position of the tile's first base
        const int64_t tile_bit0 = (tile_word0 - PGM_HALO_L) * 32;              // local position of sm.lo[0] bit 0

        // ---- stage A: hash + filter + probe, 32 window starts per thread
        const uint64_t base = tile_g0 + (uint64_t)t * 32;
        uint32_t vmask = 0;
        if (base + 32 > p.own_begin && base < p.own_end) {
            vmask = 0xFFFFFFFFu;
            if (base < p.own_begin) vmask &= 0xFFFFFFFFu << (uint32_t)(p.own_begin - base);
            if (base + 32 > p.own_end) vmask &= 0xFFFFFFFFu >> (uint32_t)(base + 32 - p.own_end);
        }
        if (vmask) {
            uint32_t wl[NCH + 1], wh[NCH + 1];
#pragma unroll
            for (int i = 0; i <= NCH; i++) { wl[i] = sm.lo[PGM_HALO_L + t + i]; wh[i] = sm.hi[PGM_HALO_L + t + i]; }
            constexpr int U = 4;
#pragma unroll 1
            for (int s0 = 0; s0 < 32; s0 += U) {
                uint32_t h1[U], h2[U], fw[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int s = s0 + u;
                    uint32_t P = 0, Q = 0, R = 0;
#pragma unroll
                    for (int i = 0; i < NCH; i++) {
                        uint32_t l = __funnelshift_r(wl[i], wl[i + 1], s);
                        uint32_t h = __funnelshift_r(wh[i], wh[i + 1], s);
                        if (i == NCH - 1) { l &= p.tail_mask; h &= p.tail_mask; }
                        P ^= l; Q ^= h; R ^= (l & h);
                    }
                    seed_hash(P, Q, R, h1[u], h2[u]);
                }
#pragma unroll
                for (int u = 0; u < U; u++)
                    fw[u] = p.tab.filter ? __ldg(p.tab.filter + filter_word(h1[u], h2[u], p.tab.filter_word_bits)) : 0xFFFFFFFFu;
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int s = s0 + u;
                    const uint32_t fm = filter_mask(h1[u], h2[u]);
                    if (((vmask >> s) & 1u) && (fw[u] & fm) == fm) {
                        const uint32_t tag = seed_tag(h2[u]);
                        uint32_t b = h1[u] & p.tab.bucket_mask;
                        for (;;) {
                            const uint4 q0 = __ldg(slots4 + (size_t)b * 2);
                            const uint4 q1 = __ldg(slots4 + (size_t)b * 2 + 1);
                            const uint32_t heads[4] = {q0.x, q0.z, q1.x, q1.z};
                            const uint32_t tags[4] = {q0.y, q0.w, q1.y, q1.w};
                            bool empty = false;
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                if ((tags[k] & 0x7FFFFFFFu) == tag) {
                                    const uint2 e = make_uint2((t * 32 + s) | (tags[k] & 0x80000000u), heads[k]);
                                    const unsigned int qi = atomicAdd(&sm.q_count, 1u);
                                    if (qi < PGM_QCAP) sm.queue[qi] = e;
                                    else {
                                        const uint3 cn = verify_entry(p, sm, tile_bit0, tile_g0, e);
                                        n_cand += cn.x; n_ver += cn.y; n_acc += cn.z; n_ovf++;
                                    }
                                }
                                empty |= tags[k] == 0xFFFFFFFFu;
                            }
                            if (empty) break;
                            b = (b + 1) & p.tab.bucket_mask;
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- stage B: drain the candidate queue, one candidate per thread
        const unsigned int total = min(sm.q_count, (unsigned int)PGM_QCAP);
        for (unsigned int i = t; i < total; i += PGM_SCAN_THREADS) {
            const uint3 cn = verify_entry(p, sm, tile_bit0, tile_g0, sm.queue[i]);
            n_cand += cn.x; n_ver += cn.y; n_acc += cn.z;
        }
        __syncthreads();
        if (t == 0) sm.q_count = 0;
    }

    // counters: warp reduce, one atomic per warp
    unsigned long long v[4] = {n_cand, n_ver, n_acc, n_ovf};
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xFFFFFFFFu, v[k], o);
        if ((t & 31) == 0 && v[k]) atomicAdd(p.counters + k, v[k]);
    }
}

// ------------------------------------------------------------------------------------------ per-pass decision
// Applies the reference's sequential accept rule of one pass to the order-free accumulators (SURVEY.md
// §8(a)-R, validated against the reference's classes):  with stored state (X, c_in) the winner is the
// smallest (class, scan order) among the events with c < c_in that report a position != X, joined —
// only after the earliest of those — by the events of the alignment that reports X itself
// (ReadsMatchers.cpp:313 skips them while X is still stored).  class = 0 for c <= minMismatches (the
// reference stops updating a read there, :304), else c.
__global__ void resolve_kernel(PerRead pr, uint32_t n_reads, uint64_t pg_len, uint32_t read_len, uint32_t seed_len,
                               uint32_t parts, uint32_t max_mm, uint32_t min_mm, int rev_mode) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const unsigned long long st = pr.state[r];
    long long best = pr.best_key[r];
    const long long o1 = pr.first_other_order[r];
    const int mask = pr.same_pos_mask[r];
    const uint32_t cx = pr.same_pos_mm[r];
    pr.best_key[r] = PGM_KEY_INF;
    pr.first_other_order[r] = PGM_KEY_INF;
    pr.same_pos_mask[r] = 0;
    pr.same_pos_mm[r] = 255;
    if (r == 0) *pr.touched = 0;
    const uint32_t c_in = (uint32_t)(st >> 56);
    if (c_in <= min_mm) return;
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
    if (best == PGM_KEY_INF) return;
    const uint32_t c = (uint32_t)(best & 0xFF);
    const uint32_t jj = (uint32_t)((best >> 8) & 0xFF);
    const uint64_t g = ((unsigned long long)best >> 16) & PGM_POS_MASK;
    const uint64_t a = g - (uint64_t)(parts - 1 - jj) * seed_len;
    const uint64_t rep = rev_mode ? pg_len - (a + read_len) : a;
    pr.state[r] = ((unsigned long long)c << 56) | ((unsigned long long)(rev_mode ? 1 : 0) << 55) | rep;
}

// state -> the three archive-visible arrays (+ matched count and per-mismatch histogram)
__global__ void finalize_kernel(const unsigned long long *__restrict__ state, uint32_t n_reads,
                                unsigned long long *__restrict__ out_pos, uint8_t *__restrict__ out_rc,
                                uint8_t *__restrict__ out_mm, unsigned long long *hist /*[257]*/) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_reads) {
        const unsigned long long st = state[r];
        const uint32_t c = (uint32_t)(st >> 56);
        out_pos[r] = c == 255u ? 0xFFFFFFFFFFFFFFFFull : (st & PGM_POS_MASK);
        out_rc[r] = (uint8_t)((st >> 55) & 1);
        out_mm[r] = (uint8_t)c;
        atomicAdd(&sh[c], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
}

} // namespace pgm
