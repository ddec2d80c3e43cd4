/* TEST INFRASTRUCTURE — CPU oracle (see pgrc_oracle.h for scope and parity status).
 *
 * Sequential restatement of PgRC's stage-4 read-vs-pseudogenome matching:
 *   pattern table     ConstantLengthPatternsOnTextHashMatcher.cpp:23-42  (addReadsSetOfPatterns)
 *   text iteration    ConstantLengthPatternsOnTextHashMatcher.h:42-68    (iterateOver / moveNext)
 *   rolling hash      rollinghash/cyclichash.h:29-35,100-123             (CyclicHash<uint32>(n, 32))
 *   exact matcher     ReadsMatchers.cpp:190-230
 *   approx matcher    ReadsMatchers.cpp:276-341
 *   interleaved mode  ReadsMatchers.cpp:343-409 (InterleavedReadsApproxMatcher),
 *                     ConstantLengthPatternsOnTextHashMatcher.h:70-137, .cpp:52-101 (strided patterns: pattern j of a
 *                     read = its symbols j, j+parts, j+2*parts, ...; one rolling hash per text position residue)
 *   pass structure    ReadsMatchers.cpp:162-184
 *   CopMEM mode       ReadsMatchers.cpp:411-451 (CopMEMReadsApproxMatcher), copmem/CopMEMMatcher.cpp:71-137 (parameters),
 *                     :139-233 (index of every k1-th text position, serial build = -t 1), :483-566 (per-read query),
 *                     copmem/Hashes.h:54-76 (maRushPrime1HashSparsified)
 *   driver            ReadsMatchers.cpp:693-783 (mapReadsIntoPg, modes 'd'/'D', 'i'/'I' and 'c'/'C')
 *   read layout       SymbolsPackingFacility.cpp:147-185, PackedConstantLengthReadsSet.h:40-46
 *   mismatch count    SymbolsPackingFacility.cpp:344-374 (contract: exact count if <= limit, else 255)
 *   reverse compl.    utils/helper.cpp:383-393
 *
 * Two deliberate differences from a literal transcription, both result-neutral:
 *  (1) The reference's CharacterHash table is seeded from time()/clock()
 *      (mersennetwister.cpp:151-175), so only the table-INDEPENDENT collisions are
 *      reproducible: with word size 32 the symbol at window offset k is rotated by
 *      (n-1-k) mod 32, hence for n > 32 offsets k and k+32 share a rotation and the hash
 *      only sees the per-rotation-class symbol parity.  The oracle hashes
 *      with the same rotate-xor construction on two fixed tables (64-bit key) and then
 *      confirms every hit by comparing that canonical form explicitly, so its candidate set
 *      is exactly "Buzhash-equivalent seeds" and does not depend on the oracle's tables.
 *  (2) std::unordered_multimap is replaced by a sorted array; equal_range order is
 *      reproduced as reverse insertion order (larger pattern index first), which is what
 *      libstdc++ does.
 */
#include "pgrc_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ read layout (a11) */

static int sym_code4(char c) { /* ACGT order, DividedPCLReadsSets.cpp:6 */
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; }
    return -1;
}
static int sym_code5(char c) { /* ACGNT order, DividedPCLReadsSets.cpp:8 */
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'N': return 3; case 'T': return 4; }
    return -1;
}
static const char SYMS4[4] = {'A', 'C', 'G', 'T'};
static const char SYMS5[5] = {'A', 'C', 'G', 'N', 'T'};

int pgo_pack_reads(const char *ascii, uint32_t n, uint32_t read_len, int with_n, uint8_t *out) {
    const uint32_t spe = with_n ? 3 : 4, sigma = with_n ? 5 : 4;
    const uint32_t packed_len = (read_len + spe - 1) / spe;
    for (uint32_t i = 0; i < n; i++) {
        const char *r = ascii + (size_t)i * read_len;
        uint8_t *o = out + (size_t)i * packed_len;
        for (uint32_t b = 0; b < packed_len; b++) {
            /* packSymbols / packSuffixSymbols: value = ((s0*σ + s1)*σ + ...), missing tail = 0 */
            uint32_t v = 0;
            for (uint32_t j = 0; j < spe; j++) {
                uint32_t p = b * spe + j;
                v *= sigma;
                if (p < read_len) {
                    int c = with_n ? sym_code5(r[p]) : sym_code4(r[p]);
                    if (c < 0) return -1;
                    v += (uint32_t)c;
                }
            }
            o[b] = (uint8_t)v;
        }
    }
    return (int)packed_len;
}

/* reverseValue(sequence, pos): symbol at position pos of a packed read */
static char packed_symbol(const uint8_t *packed, uint32_t pos, int with_n) {
    if (!with_n) {
        uint8_t v = packed[pos >> 2];
        return SYMS4[(v >> (2 * (3 - (pos & 3)))) & 3];
    } else {
        uint8_t v = packed[pos / 3];
        uint32_t j = pos % 3;
        uint32_t d = j == 0 ? v / 25 : (j == 1 ? (v / 5) % 5 : v % 5);
        return SYMS5[d];
    }
}

void pgo_unpack_read(const uint8_t *packed, uint32_t read_len, int with_n, char *out) {
    for (uint32_t p = 0; p < read_len; p++) out[p] = packed_symbol(packed, p, with_n);
}

/* SumOfConstantLengthReadsSets view: LQ set (ACGT) then N set (ACGNT) */
typedef struct {
    const uint8_t *lq; uint32_t n_lq, lq_stride;
    const uint8_t *nn; uint32_t n_n, nn_stride;
    uint32_t read_len;
} reads_view;

static char read_symbol(const reads_view *rs, uint32_t i, uint32_t pos) {
    if (i < rs->n_lq) return packed_symbol(rs->lq + (size_t)i * rs->lq_stride, pos, 0);
    return packed_symbol(rs->nn + (size_t)(i - rs->n_lq) * rs->nn_stride, pos, 1);
}

static void get_read(const reads_view *rs, uint32_t i, char *out) {
    for (uint32_t p = 0; p < rs->read_len; p++) out[p] = read_symbol(rs, i, p);
}

/* countSequenceMismatchesVsUnpacked contract on an unpacked read: Hamming distance with
 * early exit; returns the exact count if <= limit, else 255. */
static uint8_t count_mismatches(const char *read, const char *txt, uint32_t len, uint8_t limit) {
    uint8_t res = 0;
    for (uint32_t k = 0; k < len; k++)
        if (read[k] != txt[k])
            if (res++ >= limit) return PGO_NOT_MATCHED_COUNT;
    return res;
}

/* ------------------------------------------------------------------ rolling hash (a1, a2) */

static uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

typedef struct {
    uint32_t t1[256], t2[256]; /* two independent CharacterHash tables */
    uint32_t n, myr;           /* window length, n % 32 (cyclichash.h:33) */
    uint32_t h1, h2;
} buzhash;

static uint32_t rotl32(uint32_t x, uint32_t r) { r &= 31; return r ? (x << r) | (x >> (32 - r)) : x; }

static void buz_init(buzhash *b, uint32_t n) {
    uint64_t s = 0x5047524342323030ULL; /* fixed seed: results must not depend on it */
    for (int k = 0; k < 256; k++) { uint64_t z = splitmix64(&s); b->t1[k] = (uint32_t)z; b->t2[k] = (uint32_t)(z >> 32); }
    b->n = n; b->myr = n % 32; b->h1 = b->h2 = 0;
}
static void buz_reset(buzhash *b) { b->h1 = b->h2 = 0; }
static void buz_eat(buzhash *b, unsigned char c) { /* cyclichash.h:120-123 */
    b->h1 = rotl32(b->h1, 1) ^ b->t1[c];
    b->h2 = rotl32(b->h2, 1) ^ b->t2[c];
}
static void buz_update(buzhash *b, unsigned char out, unsigned char in) { /* cyclichash.h:100-107 */
    b->h1 = rotl32(b->h1, 1) ^ rotl32(b->t1[out], b->myr) ^ b->t1[in];
    b->h2 = rotl32(b->h2, 1) ^ rotl32(b->t2[out], b->myr) ^ b->t2[in];
}
static uint64_t buz_value(const buzhash *b) { return ((uint64_t)b->h2 << 32) | b->h1; }

/* Canonical form of a window under CyclicHash(n, 32): for every rotation class
 * (n-1-k) mod 32 the parity mask of the symbols occurring in it. */
static void canonical_form(const char *w, uint32_t n, uint8_t form[32]) {
    memset(form, 0, 32);
    for (uint32_t k = 0; k < n; k++) {
        uint8_t bit;
        switch (w[k]) { case 'A': bit = 1; break; case 'C': bit = 2; break; case 'G': bit = 4; break;
                        case 'T': bit = 8; break; case 'N': bit = 16; break; default: bit = 32; }
        form[(n - 1 - k) & 31] ^= bit;
    }
}

/* ------------------------------------------------------------------ pattern table (a3, a4) */

typedef struct { uint64_t key; uint32_t idx; } pat_entry;

typedef struct {
    pat_entry *e; uint64_t n;      /* sorted by (key asc, idx desc) */
    uint32_t *dir; uint64_t dir_mask; /* open addressing: first index of each key run, or UINT32_MAX */
    uint32_t pattern_len, parts;
    buzhash hf;
} pat_table;

static int cmp_pat(const void *a, const void *b) {
    const pat_entry *x = (const pat_entry *)a, *y = (const pat_entry *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    if (x->idx != y->idx) return x->idx > y->idx ? -1 : 1; /* LIFO: larger pattern index first */
    return 0;
}

static uint64_t dir_slot(uint64_t key, uint64_t mask) {
    key ^= key >> 33; key *= 0xff51afd7ed558ccdULL; key ^= key >> 33;
    return key & mask;
}

static void table_free(pat_table *t) { free(t->e); free(t->dir); t->e = NULL; t->dir = NULL; }

/* addReadsSetOfPatterns(readsSet, partsCount, matchedReadsBitmap) */
static int table_build(pat_table *t, const reads_view *rs, uint32_t n_reads, uint32_t pattern_len,
                       uint32_t parts, const uint8_t *skip_bitmap, int interleaved) {
    memset(t, 0, sizeof(*t));
    t->pattern_len = pattern_len; t->parts = parts;
    buz_init(&t->hf, pattern_len);
    uint64_t cap = (uint64_t)n_reads * parts;
    t->e = (pat_entry *)malloc((cap ? cap : 1) * sizeof(pat_entry));
    if (!t->e) return -2;
    uint64_t m = 0;
    for (uint32_t i = 0; i < n_reads; i++) {
        if (skip_bitmap && skip_bitmap[i]) continue;
        uint32_t offset = 0;
        for (uint32_t j = 0; j < parts; j++, offset += pattern_len) {
            buz_reset(&t->hf);
            for (uint32_t k = 0; k < pattern_len; k++)   /* addPackedPatterns (HashMatcher.cpp:89-91): symbols j + k*parts */
                buz_eat(&t->hf, (unsigned char)read_symbol(rs, i, interleaved ? j + k * parts : offset + k));
            t->e[m].key = buz_value(&t->hf);
            t->e[m].idx = i * parts + j;
            m++;
        }
    }
    t->n = m;
    qsort(t->e, m, sizeof(pat_entry), cmp_pat);
    uint64_t dsz = 16;
    while (dsz < 2 * m + 1) dsz <<= 1;
    t->dir = (uint32_t *)malloc(dsz * sizeof(uint32_t));
    if (!t->dir) { table_free(t); return -2; }
    memset(t->dir, 0xFF, dsz * sizeof(uint32_t));
    t->dir_mask = dsz - 1;
    for (uint64_t k = 0; k < m; k++) {
        if (k > 0 && t->e[k].key == t->e[k - 1].key) continue;
        uint64_t s = dir_slot(t->e[k].key, t->dir_mask);
        while (t->dir[s] != UINT32_MAX) s = (s + 1) & t->dir_mask;
        t->dir[s] = (uint32_t)k;
    }
    return 0;
}

/* equal_range(hash): returns first index and count of the run with this key */
static uint64_t table_lookup(const pat_table *t, uint64_t key, uint64_t *first) {
    if (t->n == 0) return 0;
    uint64_t s = dir_slot(key, t->dir_mask);
    while (t->dir[s] != UINT32_MAX) {
        uint64_t k = t->dir[s];
        if (t->e[k].key == key) {
            uint64_t c = 1;
            while (k + c < t->n && t->e[k + c].key == key) c++;
            *first = k;
            return c;
        }
        s = (s + 1) & t->dir_mask;
    }
    return 0;
}

/* ------------------------------------------------------------------ matcher state (a5) */

typedef struct {
    const char *pg; uint64_t pg_len; int rev_compl;
    const reads_view *rs; uint32_t n_reads, read_len, matching_len;
    uint64_t *pos; uint8_t *rc; uint8_t *mm; /* readMatchPos, readMatchRC, readMismatchesCount */
    pgo_stats *st;
    /* approx parameters */
    uint32_t part_len, parts; uint8_t max_mm, min_mm;
    int interleaved, copmem;
    char *cur_read, *seed_buf, *win_buf;
} matcher;

/* interleaved mode: the same test on the strided seed (read symbols j + k*parts) and the strided window */
static int seed_equivalent_strided(const matcher *m, uint32_t read_idx, uint32_t j, uint32_t n, const char *txt_at_pos) {
    uint8_t f1[32], f2[32];
    for (uint32_t k = 0; k < n; k++) {
        m->seed_buf[k] = read_symbol(m->rs, read_idx, j + k * m->parts);
        m->win_buf[k] = txt_at_pos[(size_t)k * m->parts];
    }
    canonical_form(m->seed_buf, n, f1);
    canonical_form(m->win_buf, n, f2);
    return memcmp(f1, f2, 32) == 0;
}

/* true iff the hash hit is a table-independent (structural) one */
static int seed_equivalent(const matcher *m, uint32_t read_idx, uint32_t offset, uint32_t n, const char *window) {
    uint8_t f1[32], f2[32];
    for (uint32_t k = 0; k < n; k++) m->seed_buf[k] = read_symbol(m->rs, read_idx, offset + k);
    canonical_form(m->seed_buf, n, f1);
    canonical_form(window, n, f2);
    return memcmp(f1, f2, 32) == 0;
}

/* DefaultReadsExactMatcher::executeMatching (ReadsMatchers.cpp:198-230) */
static void exact_pass(matcher *m, pat_table *t, const char *txt, int rev_mode) {
    const uint64_t L = t->pattern_len;
    if (m->pg_len < L) return;
    buzhash *hf = &t->hf;
    buz_reset(hf);
    for (uint64_t i = 0; i < L; i++) buz_eat(hf, (unsigned char)txt[i]);
    for (uint64_t p = 0; p + L <= m->pg_len; p++) {
        uint64_t first, cnt = table_lookup(t, buz_value(hf), &first);
        unsigned char in = p + L < m->pg_len ? (unsigned char)txt[p + L] : 0; /* txt[len] is the NUL */
        buz_update(hf, (unsigned char)txt[p], in);
        for (uint64_t q = 0; q < cnt; q++) {
            uint32_t r = t->e[first + q].idx;
            m->st->n_events++;
            if (!seed_equivalent(m, r, 0, (uint32_t)L, txt + p)) continue; /* accidental: not reproducible */
            get_read(m->rs, r, m->cur_read);
            m->st->n_verified++;
            int equal = memcmp(m->cur_read, txt + p, m->read_len) == 0; /* compareReadWithPattern == 0 */
            if (equal) {
                if (m->pos[r] == PGO_NOT_MATCHED_POSITION) {
                    m->pos[r] = rev_mode ? m->pg_len - (p + m->matching_len) : p;
                    if (rev_mode) m->rc[r] = 1;
                    m->st->matched++;
                } else
                    m->st->better++;
            } else
                m->st->false_matches++;
        }
    }
}

/* DefaultReadsApproxMatcher::executeMatching (ReadsMatchers.cpp:297-341) */
static void approx_pass(matcher *m, pat_table *t, const char *txt, int rev_mode) {
    const uint64_t L = t->pattern_len;
    if (m->pg_len < L) return;
    buzhash *hf = &t->hf;
    buz_reset(hf);
    for (uint64_t i = 0; i < L; i++) buz_eat(hf, (unsigned char)txt[i]);
    for (uint64_t p = 0; p + L <= m->pg_len; p++) {
        uint64_t first, cnt = table_lookup(t, buz_value(hf), &first);
        unsigned char in = p + L < m->pg_len ? (unsigned char)txt[p + L] : 0;
        buz_update(hf, (unsigned char)txt[p], in);
        for (uint64_t q = 0; q < cnt; q++) {
            const uint32_t pat = t->e[first + q].idx;
            const uint32_t r = pat / m->parts;
            m->st->n_events++;
            if (!seed_equivalent(m, r, (pat % m->parts) * m->part_len, (uint32_t)L, txt + p)) continue;
            if (m->mm[r] <= m->min_mm) continue;
            uint64_t match_pos = p;
            const uint32_t shift = (pat % m->parts) * m->part_len;
            if (shift > match_pos) continue;
            match_pos -= shift;
            if (match_pos + m->read_len > m->pg_len) continue;
            const uint64_t rep = rev_mode ? m->pg_len - (match_pos + m->matching_len) : match_pos;
            if (m->pos[r] == rep) { /* coordinate-only compare */
                if (m->rc[r] != (uint8_t)(rev_mode ? 1 : 0)) m->st->n_cross_strand_skips++;
                continue;
            }
            const uint8_t limit = m->mm[r] == PGO_NOT_MATCHED_COUNT ? m->max_mm : (uint8_t)(m->mm[r] - 1);
            get_read(m->rs, r, m->cur_read);
            m->st->n_verified++;
            const uint8_t c = count_mismatches(m->cur_read, txt + match_pos, m->matching_len, limit);
            if (c < m->mm[r]) {
                if (m->mm[r] == PGO_NOT_MATCHED_COUNT) m->st->matched++;
                else m->st->better++;
                m->st->per_mm[m->mm[r]]--;
                m->st->per_mm[c]++;
                m->pos[r] = rep;
                m->rc[r] = (uint8_t)(rev_mode ? 1 : 0);
                m->mm[r] = c;
            } else
                m->st->false_matches++;
        }
    }
}

/* InterleavedReadsApproxMatcher::executeMatching (ReadsMatchers.cpp:365-409) over
 * InterleavedConstantLengthPatternsOnTextHashMatcher::iterateOver / moveNext (HashMatcher.h:108-135): `parts` rolling
 * hashes, the one of residue txtPos mod parts covers txt[txtPos + k*parts], k < patternLength; a hit of pattern
 * i*parts + j at txtPos aligns the read at txtPos - j.  (A text shorter than patternSpan makes the reference's
 * unsigned loop bound wrap around: undefined there, no events here.) */
static int interleaved_pass(matcher *m, pat_table *t, const char *txt, int rev_mode) {
    const uint64_t n = t->pattern_len, parts = m->parts, span = n * parts;
    if (m->pg_len < span) return 0;
    buzhash *hf = (buzhash *)malloc(parts * sizeof(buzhash));
    if (!hf) return -2;
    for (uint64_t h = 0; h < parts; h++) {
        hf[h] = t->hf;
        buz_reset(&hf[h]);
        for (uint64_t i = h; i < m->pg_len && i < span + h; i += parts) buz_eat(&hf[h], (unsigned char)txt[i]);
    }
    uint64_t cur = 0;
    for (uint64_t p = 0; p + span <= m->pg_len; p++) {
        uint64_t first, cnt = table_lookup(t, buz_value(&hf[cur]), &first);
        unsigned char in = p + span < m->pg_len ? (unsigned char)txt[p + span] : 0; /* txt[len] is the NUL */
        buz_update(&hf[cur], (unsigned char)txt[p], in);
        if (++cur == parts) cur = 0;
        for (uint64_t q = 0; q < cnt; q++) {
            const uint32_t pat = t->e[first + q].idx;
            const uint32_t r = pat / m->parts, j = pat % m->parts;
            m->st->n_events++;
            if (!seed_equivalent_strided(m, r, j, (uint32_t)n, txt + p)) continue;
            if (m->mm[r] <= m->min_mm) continue;
            uint64_t match_pos = p;
            if (j > match_pos) continue;                       /* positionShift = j (:373-376) */
            match_pos -= j;
            if (match_pos + m->read_len > m->pg_len) continue;
            const uint64_t rep = rev_mode ? m->pg_len - (match_pos + m->matching_len) : match_pos;
            if (m->pos[r] == rep) {
                if (m->rc[r] != (uint8_t)(rev_mode ? 1 : 0)) m->st->n_cross_strand_skips++;
                continue;
            }
            const uint8_t limit = m->mm[r] == PGO_NOT_MATCHED_COUNT ? m->max_mm : (uint8_t)(m->mm[r] - 1);
            get_read(m->rs, r, m->cur_read);
            m->st->n_verified++;
            const uint8_t c = count_mismatches(m->cur_read, txt + match_pos, m->matching_len, limit);
            if (c < m->mm[r]) {
                if (m->mm[r] == PGO_NOT_MATCHED_COUNT) m->st->matched++;
                else m->st->better++;
                m->st->per_mm[m->mm[r]]--;
                m->st->per_mm[c]++;
                m->pos[r] = rep;
                m->rc[r] = (uint8_t)(rev_mode ? 1 : 0);
                m->mm[r] = c;
            } else
                m->st->false_matches++;
        }
    }
    free(hf);
    return 0;
}

/* ------------------------------------------------------------------ CopMEM mode ('c')
 * CopMEMReadsApproxMatcher (ReadsMatchers.cpp:411-451): per pass an index of the (reverse-complemented) text — every
 * k1-th position, hashed over K characters, at most 13 positions per hash value, first come first kept — and per read a
 * sequential query: every k2-th read offset, bucket entries in text order, verification with an early exit, strict
 * improvement, a budget of false matches after which buckets are cut to 4 entries.  This restates the SERIAL index
 * build (processRef, CopMEMMatcher.cpp:171-233), i.e. the reference at -t 1; its multithreaded build
 * (processRefMultithreaded, :277-324) orders buckets differently and races. */
#define CM_COLLISIONS_LIMIT 12      /* HASH_COLLISIONS_PER_POSITION_LIMIT            CopMEMMatcher.h:11 */
#define CM_AVG_COLLISIONS_LIMIT 1   /* AVERAGE_HASH_COLLISIONS_PER_POSITION_LIMIT    :12 */
#define CM_TRUNCATED_BUCKET 4       /* UNLIMITED_NUMBER_OF_HASH_COLLISIONS_PER_POSITION :13 */
#define CM_HASH_MIN_ORDER 24
#define CM_HASH_MAX_ORDER 31

typedef struct {
    uint32_t L, K, k1, k2, hash_size;
    uint64_t N;
    const char *text;
    uint32_t *cumm;      /* hash_size + 2 */
    uint64_t *pos;       /* sampled positions, bucket h = [cumm[h], cumm[h+1]) */
} copmem_index;

/* initParams + calcCoprimes (CopMEMMatcher.cpp:71-137); minMatchLength = min(UINT32_MAX, L) = L.  Returns -1 where the
 * reference exits ("Minimal matching length too short", "L and K mismatch"). */
static int copmem_params(copmem_index *x, uint32_t L, uint64_t N) {
    int K;
    if (L > 110) K = 56; else if (L > 62) K = 44; else if (L > 53) K = 40; else if (L > 46) K = 36;
    else if (L > 42) K = 32; else if (L > 32) K = 28; else K = ((int)L / 4 - 1) * 4;
    if (L < 24) return -1;
    const int kmml = ((int)L / 4 - 1) * 4;
    if (kmml < K) K = kmml;
    const int t = (int)L - K + 1;
    if (t <= 0) return -1;
    int k1, k2;
    if (t >= 20) {
        k1 = 1; while ((k1 + 1) * (k1 + 1) <= t) k1++;      /* (int) pow(t, 0.5) */
        k1 += 1; k2 = k1 - 1;
        if (k1 * k2 > t) { --k2; --k1; }
    } else if (t >= 15) { k1 = 5; k2 = 3; } else if (t >= 12) { k1 = 4; k2 = 3; } else if (t >= 10) { k1 = 5; k2 = 2; }
    else if (t >= 6) { k1 = 3; k2 = 2; } else { k1 = t; k2 = 1; }
    x->L = L; x->K = (uint32_t)K; x->k1 = (uint32_t)k1; x->k2 = (uint32_t)k2; x->N = N;
    int i = CM_HASH_MIN_ORDER;
    do { x->hash_size = 1u << (i++); } while (i <= CM_HASH_MAX_ORDER && x->hash_size < N / (uint64_t)k1);
    return 0;
}

/* maRushPrime1HashSparsified<K> (Hashes.h:54-76): K/4 little-endian 32-bit words of the text, the first three masked to
 * their low 3 bytes, the others to their low 2 bytes */
static uint32_t copmem_hash(const copmem_index *x, const char *str) {
    uint64_t hash = x->K;
    for (uint32_t j = 0; j < x->K / 4; j++, str += 4) {
        uint32_t k = (uint32_t)(unsigned char)str[0] | ((uint32_t)(unsigned char)str[1] << 8) |
                     ((uint32_t)(unsigned char)str[2] << 16) | ((uint32_t)(unsigned char)str[3] << 24);
        k &= j < 3 ? 0x00FFFFFFu : 0x0000FFFFu;
        k += j;
        hash ^= k;
        hash *= 171717;
    }
    return (uint32_t)hash & (x->hash_size - 1);
}

static void copmem_free(copmem_index *x) { free(x->cumm); free(x->pos); x->cumm = NULL; x->pos = NULL; }

/* genCumm + processRef (CopMEMMatcher.cpp:139-233): positions 0, k1, 2 k1, ... <= N - K in ascending order; a hash value
 * keeps its first CM_COLLISIONS_LIMIT + 1 positions */
static int copmem_build(copmem_index *x, const char *text) {
    x->text = text;
    x->cumm = (uint32_t *)calloc((size_t)x->hash_size + 2, sizeof(uint32_t));
    if (!x->cumm) return -2;
    uint32_t *cnt = x->cumm + 1;          /* cnt[h] = entries of bucket h; prefix sums below turn cumm[h] into its start */
    if (x->N >= x->K)
        for (uint64_t i = 0; i + x->K <= x->N; i += x->k1) {
            const uint32_t h = copmem_hash(x, text + i);
            if (cnt[h] <= CM_COLLISIONS_LIMIT) cnt[h]++;
        }
    uint64_t total = 0;
    for (uint64_t h = 0; h <= x->hash_size; h++) { const uint32_t c = cnt[h]; x->cumm[h] = (uint32_t)total; total += c; }
    x->cumm[x->hash_size + 1] = (uint32_t)total;
    /* (cumm[h] is now the start of bucket h, cumm[h+1] its end; cnt aliases cumm + 1, rewritten in the loop above) */
    x->pos = (uint64_t *)malloc((total + 2) * sizeof(uint64_t));
    uint32_t *fill = (uint32_t *)calloc((size_t)x->hash_size + 1, sizeof(uint32_t));
    if (!x->pos || !fill) { free(fill); return -2; }
    if (x->N >= x->K)
        for (uint64_t i = 0; i + x->K <= x->N; i += x->k1) {
            const uint32_t h = copmem_hash(x, text + i);
            if (fill[h] <= CM_COLLISIONS_LIMIT) { x->pos[x->cumm[h] + fill[h]] = i; fill[h]++; }
        }
    free(fill);
    return 0;
}

/* processApproxMatchQueryTight (CopMEMMatcher.cpp:483-566) */
static uint64_t copmem_query(const copmem_index *x, const char *read, uint32_t N2, uint8_t max_mm, uint8_t min_mm,
                             uint8_t *mismatches, uint64_t *better, uint64_t *false_matches) {
    if (*mismatches < max_mm) max_mm = (uint8_t)(*mismatches - 1);
    const uint32_t n2trim8 = (N2 / 8) * 8;
    const uint64_t limit = (uint64_t)((N2 + 1 - x->K) / x->k2) * CM_AVG_COLLISIONS_LIMIT;
    uint64_t cur_false = 0, match_pos = PGO_NOT_MATCHED_POSITION;
    for (uint32_t i1 = 0; i1 + x->K < N2 + 1; i1 += x->k2) {
        const uint32_t h = copmem_hash(x, read + i1);
        uint32_t b0 = x->cumm[h], b1 = x->cumm[h + 1];
        if (b0 == b1) continue;
        if (limit < cur_false && b1 > b0 + CM_TRUNCATED_BUCKET) b1 = b0 + CM_TRUNCATED_BUCKET;
        for (uint32_t j = b0; j < b1; j++) {
            const uint64_t sp = x->pos[j];
            if (i1 > sp) continue;
            if (sp - i1 + N2 > x->N) continue;
            const char *txt = x->text + (sp - i1);
            uint8_t res = 0;
            uint32_t p = 0;
            while (res <= max_mm && p != n2trim8) {            /* 8 characters at a time, exit only between blocks */
                for (int q = 0; q < 8; q++) res = (uint8_t)(res + (read[p + q] != txt[p + q]));
                p += 8;
            }
            if (res > max_mm) { cur_false++; continue; }
            while (p != N2) {
                if (read[p] != txt[p]) {
                    p++;
                    if (res++ >= max_mm) { cur_false++; break; }   /* (counted once more just below, as in the reference) */
                } else p++;
            }
            if (res > max_mm) { cur_false++; continue; }
            if (*mismatches != 255) (*better)++;
            *mismatches = res;
            match_pos = sp - i1;
            if (res <= min_mm) { *false_matches += cur_false; return match_pos; }
            max_mm = (uint8_t)(res - 1);
        }
    }
    *false_matches += cur_false;
    return match_pos;
}

/* The STAGED form of copmem_query that the warp-per-read CUDA kernels use (pgrc_b200/csrc/pgm_copmem_warp.cuh), as a CPU model
 * (pgo_set_copmem_staged(1) makes mode 'c' run through it; tests compare it with the sequential transcription above):
 *   stage 1  every read offset: hash, bucket; every bucket entry: its alignment start a = sp - i1, or "skipped" (:512, :514);
 *            the DISTINCT alignments of the read are verified once each, with both mismatch counts in full (first trim8
 *            characters / tail) and kept in a table of CMS_VT entries; a candidate becomes a one-byte index into it;
 *   stage 2  the sequential loop replayed over those bytes: bucket truncation by the false-match budget, strict improvement,
 *            the +1 / +2 false-match accounting, the early stop at min_mm.
 * Exact because a verification's outcome under any limit follows from the two full counts: blocks > limit -> +1 false match;
 * else blocks + tail > limit -> +2; else accept.  More than CMS_VT distinct alignments: the read takes the sequential path. */
#ifndef CMS_VT
#define CMS_VT 32
#endif
static int g_copmem_staged = 0;
void pgo_set_copmem_staged(int on) { g_copmem_staged = on; }

static uint64_t copmem_query_staged(const copmem_index *x, const char *read, uint32_t N2, uint8_t max_mm, uint8_t min_mm,
                                    uint8_t *mismatches, uint64_t *better, uint64_t *false_matches) {
    const uint32_t n_off = N2 >= x->K ? (N2 - x->K) / x->k2 + 1 : 0;
    const uint32_t n2trim8 = (N2 / 8) * 8;
    uint8_t *lens = (uint8_t *)malloc(n_off + 1), *cand = (uint8_t *)malloc((size_t)n_off * (CM_COLLISIONS_LIMIT + 1) + 1);
    uint64_t vt_a[CMS_VT];
    uint8_t vt_db[CMS_VT], vt_dt[CMS_VT];
    uint32_t n_vt = 0, n_cand = 0;
    int overflow = 0;
    for (uint32_t o = 0; o < n_off && !overflow; o++) {                        /* stage 1 */
        const uint32_t i1 = o * x->k2;
        const uint32_t h = copmem_hash(x, read + i1);
        const uint32_t b0 = x->cumm[h], b1 = x->cumm[h + 1];
        lens[o] = (uint8_t)(b1 - b0);
        for (uint32_t j = b0; j < b1 && !overflow; j++) {
            const uint64_t sp = x->pos[j];
            uint8_t v = 0xFF;
            if (i1 <= sp && sp - i1 + N2 <= x->N) {
                const uint64_t a = sp - i1;
                uint32_t k = 0;
                while (k < n_vt && vt_a[k] != a) k++;
                if (k == n_vt) {
                    if (n_vt == CMS_VT) { overflow = 1; break; }
                    const char *txt = x->text + a;
                    uint32_t db = 0, dt = 0;
                    for (uint32_t p = 0; p < N2; p++) if (read[p] != txt[p]) { if (p < n2trim8) db++; else dt++; }
                    vt_a[k] = a; vt_db[k] = (uint8_t)db; vt_dt[k] = (uint8_t)dt; n_vt++;
                }
                v = (uint8_t)k;
            }
            cand[n_cand++] = v;
        }
    }
    if (overflow) { free(lens); free(cand); return copmem_query(x, read, N2, max_mm, min_mm, mismatches, better, false_matches); }
    if (*mismatches < max_mm) max_mm = (uint8_t)(*mismatches - 1);              /* stage 2 */
    const uint64_t limit = (uint64_t)((N2 + 1 - x->K) / x->k2) * CM_AVG_COLLISIONS_LIMIT;
    uint64_t cur_false = 0, match_pos = PGO_NOT_MATCHED_POSITION;
    uint32_t base = 0;
    int done = 0;
    for (uint32_t o = 0; o < n_off && !done; o++) {
        const uint32_t len = lens[o];
        uint32_t lim = len;
        if (limit < cur_false && len > CM_TRUNCATED_BUCKET) lim = CM_TRUNCATED_BUCKET;
        for (uint32_t t = 0; t < lim; t++) {
            const uint8_t v = cand[base + t];
            if (v == 0xFF) continue;
            const uint32_t db = vt_db[v], dt = vt_dt[v];
            if (db > max_mm) { cur_false++; continue; }
            if (db + dt > max_mm) { cur_false += 2; continue; }
            if (*mismatches != 255) (*better)++;
            *mismatches = (uint8_t)(db + dt);
            match_pos = vt_a[v];
            if (db + dt <= min_mm) { done = 1; break; }
            max_mm = (uint8_t)(db + dt - 1);
        }
        base += len;
    }
    *false_matches += cur_false;
    free(lens); free(cand);
    return match_pos;
}

/* CopMEMReadsApproxMatcher::executeMatching (ReadsMatchers.cpp:421-451), one thread */
static int copmem_pass(matcher *m, const char *txt, int rev_mode) {
    copmem_index x;
    memset(&x, 0, sizeof x);
    if (copmem_params(&x, m->part_len, m->pg_len) != 0) return -1;
    int rcode = copmem_build(&x, txt);
    if (rcode != 0) { copmem_free(&x); return rcode; }
    for (uint32_t r = 0; r < m->n_reads; r++) {
        if (m->mm[r] <= m->min_mm) continue;
        get_read(m->rs, r, m->cur_read);
        uint8_t c = m->mm[r];
        const uint64_t p = (g_copmem_staged ? copmem_query_staged : copmem_query)(&x, m->cur_read, m->matching_len, m->max_mm, m->min_mm, &c,
                                                                                   &m->st->better, &m->st->false_matches);
        if (p == PGO_NOT_MATCHED_POSITION) continue;
        if (c < m->mm[r]) {
            if (m->mm[r] == PGO_NOT_MATCHED_COUNT) m->st->matched++;
            m->st->per_mm[m->mm[r]]--;
            m->st->per_mm[c]++;
            m->pos[r] = rev_mode ? m->pg_len - (p + m->matching_len) : p;
            m->rc[r] = (uint8_t)(rev_mode ? 1 : 0);
            m->mm[r] = c;
        }
    }
    copmem_free(&x);
    return 0;
}

/* reverseComplementInPlace semantics on a copy (helper.cpp:383-393) */
static char *reverse_complement(const char *s, uint64_t n) {
    char *o = (char *)malloc(n + 1);
    if (!o) return NULL;
    for (uint64_t i = 0; i < n; i++) {
        char c = s[n - 1 - i];
        o[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
    }
    o[n] = 0;
    return o;
}

/* matchConstantLengthReads / continueMatchingConstantLengthReads pass structure */
static int one_pass(matcher *m, pat_table *t, int exact, const char *txt, int rev_mode) {
    if (m->copmem) return copmem_pass(m, txt, rev_mode);
    if (exact) exact_pass(m, t, txt, rev_mode);
    else if (m->interleaved) return interleaved_pass(m, t, txt, rev_mode);
    else approx_pass(m, t, txt, rev_mode);
    return 0;
}

static int run_passes(matcher *m, pat_table *t, int exact) {
    int rcode = one_pass(m, t, exact, m->pg, 0);
    if (rcode == 0 && m->rev_compl) {
        char *rcpg = reverse_complement(m->pg, m->pg_len);
        if (!rcpg) return -2;
        rcode = one_pass(m, t, exact, rcpg, 1);
        free(rcpg);
    }
    return rcode;
}

/* ------------------------------------------------------------------ driver (a12, a13) */

static int is_upper_mode(char c) { return c >= 'A' && c <= 'Z'; }
static char lower_mode(char c) { return is_upper_mode(c) ? (char)(c - 'A' + 'a') : c; }

int pgo_map_reads(const char *text, uint64_t text_len,
                  const uint8_t *lq_packed, uint32_t n_lq,
                  const uint8_t *n_packed, uint32_t n_n,
                  uint32_t read_len, uint32_t pre_seed, uint32_t seed,
                  uint32_t min_chars_per_mismatch, char pre_mode, char mode, int rev_compl,
                  uint64_t *out_pos, uint8_t *out_rc, uint8_t *out_mm, pgo_stats *stats) {
    if (!text || read_len == 0 || read_len > 255 || seed == 0 || min_chars_per_mismatch == 0) return -1;
    if ((lower_mode(mode) != 'd' && lower_mode(mode) != 'i' && lower_mode(mode) != 'c') ||
        (pre_seed && lower_mode(pre_mode) != 'd' && lower_mode(pre_mode) != 'i' && lower_mode(pre_mode) != 'c')) return -1;
    pgo_stats local;
    if (!stats) stats = &local;
    memset(stats, 0, sizeof(*stats));

    reads_view rs = { lq_packed, n_lq, (read_len + 3) / 4, n_packed, n_n, (read_len + 2) / 3, read_len };
    const uint32_t n = n_lq + n_n;

    /* ReadsMatchers.cpp:699-713 */
    const uint8_t max_mm = (uint8_t)(read_len / min_chars_per_mismatch);
    uint32_t reads_exact = seed > read_len ? read_len : seed;
    uint32_t pre_exact = pre_seed > read_len ? read_len : pre_seed;
    uint32_t cur_exact = reads_exact;
    char cur_mode = mode;
    if (pre_exact > 0) { cur_exact = pre_exact; cur_mode = pre_mode; }
    const uint8_t cur_min_mm = is_upper_mode(cur_mode) ? max_mm : 0;
    const uint8_t target_mm = (uint8_t)(read_len / cur_exact - 1);

    matcher m;
    memset(&m, 0, sizeof(m));
    m.pg = text; m.pg_len = text_len; m.rev_compl = rev_compl;
    m.rs = &rs; m.n_reads = n; m.read_len = read_len; m.matching_len = read_len; /* DISABLED_PREFIX_MODE */
    m.pos = out_pos; m.rc = out_rc; m.mm = out_mm; m.st = stats;
    m.cur_read = (char *)malloc(read_len + 1);
    m.seed_buf = (char *)malloc(read_len + 1);
    m.win_buf = (char *)malloc(read_len + 1);
    if (!m.cur_read || !m.seed_buf || !m.win_buf) { free(m.cur_read); free(m.seed_buf); free(m.win_buf); return -2; }

    /* DefaultReadsMatcher::initMatching (ReadsMatchers.cpp:97-105) */
    for (uint32_t i = 0; i < n; i++) { out_pos[i] = PGO_NOT_MATCHED_POSITION; out_rc[i] = 0; out_mm[i] = PGO_NOT_MATCHED_COUNT; }

    int rcode = 0;
    pat_table t;
    const int first_copmem = lower_mode(cur_mode) == 'c';            /* :717-720, :732-735: CopMEM also when readLength == seed */
    const int first_exact = (read_len == cur_exact) && !first_copmem;
    if (first_copmem) {
        m.copmem = 1;
        m.part_len = cur_exact; m.parts = (uint32_t)target_mm + 1; m.max_mm = max_mm; m.min_mm = cur_min_mm;
        stats->per_mm[PGO_NOT_MATCHED_COUNT] = n;
        rcode = run_passes(&m, NULL, 0);
    } else if (first_exact) {
        /* DefaultReadsExactMatcher::initMatching: whole reads as patterns (parts = 1) */
        rcode = table_build(&t, &rs, n, m.matching_len, 1, NULL, 0);
        if (rcode == 0) { stats->n_patterns = t.n; rcode = run_passes(&m, &t, 1); table_free(&t); }
        /* DefaultReadsExactMatcher::transferMatchingResults (ReadsMatchers.cpp:127-133) */
        for (uint32_t i = 0; i < n; i++) out_mm[i] = out_pos[i] == PGO_NOT_MATCHED_POSITION ? PGO_NOT_MATCHED_COUNT : 0;
        stats->per_mm[0] = stats->matched;
        stats->per_mm[PGO_NOT_MATCHED_COUNT] = n - stats->matched;
    } else {
        /* AbstractReadsApproxMatcher ctor + DefaultReadsApproxMatcher::initMatching */
        m.part_len = cur_exact; m.parts = (uint32_t)target_mm + 1; m.max_mm = max_mm; m.min_mm = cur_min_mm;
        m.interleaved = lower_mode(cur_mode) == 'i';
        stats->per_mm[PGO_NOT_MATCHED_COUNT] = n;
        rcode = table_build(&t, &rs, n, m.part_len, m.parts, NULL, m.interleaved);
        if (rcode == 0) { stats->n_patterns = t.n; rcode = run_passes(&m, &t, 0); table_free(&t); }
    }

    if (rcode == 0 && pre_exact > 0) {
        /* 2nd phase (ReadsMatchers.cpp:749-779): minMismatches uses the FIRST phase's targetMismatches */
        const uint8_t min_mm2 = is_upper_mode(mode) ? max_mm : (uint8_t)(target_mm + 1);
        m.part_len = reads_exact;
        m.parts = read_len / reads_exact; /* targetMismatches + 1 of the new matcher */
        m.max_mm = max_mm; m.min_mm = min_mm2;
        m.interleaved = lower_mode(mode) == 'i';
        m.copmem = lower_mode(mode) == 'c';
        /* initMatchingContinuation: getMatchedReadsBitmap(minMismatches) of the previous matcher:
         * exact matcher ignores the argument (:677-683), approx matcher uses mm <= arg (:685-691) */
        uint8_t *skip = (uint8_t *)malloc(n ? n : 1);
        if (!skip) rcode = -2;
        else if (m.copmem) {
            /* CopMEMReadsApproxMatcher::initMatchingContinuation only takes the results over; its pass skips mm <= minMismatches */
            free(skip);
            rcode = run_passes(&m, NULL, 0);
        } else {
            for (uint32_t i = 0; i < n; i++)
                skip[i] = first_exact ? (out_pos[i] != PGO_NOT_MATCHED_POSITION) : (out_mm[i] <= min_mm2);
            rcode = table_build(&t, &rs, n, m.part_len, m.parts, skip, m.interleaved);
            free(skip);
            if (rcode == 0) { stats->n_patterns = t.n; rcode = run_passes(&m, &t, 0); table_free(&t); }
        }
    }
    free(m.cur_read); free(m.seed_buf); free(m.win_buf);
    return rcode;
}

/* ------------------------------------------------------------------ mismatch lists of the export (a-next, §8(f) rank 2)
 * AbstractReadsApproxMatcher::updateEntry (ReadsMatchers.cpp:548-558) with fillEntryWithMismatches (:40-52) and
 * fillEntryWithReversedMismatches (:54-66): per matched read, the (offset, pseudogenome symbol, read symbol) triples the
 * export hands to mismatch2CxtCode / addMismatch.
 *   variant 0: the forward fill for every read (the read reverse-complemented when rc is set) — the contract of
 *              pgm_get_mismatches, from which the shim derives the other two on the host;
 *   variant 1: updateEntry(entry, i, revComplPairFile = false): reversed fill iff rc;
 *   variant 2: updateEntry(entry, i, revComplPairFile = true):  reversed fill iff rc != (idx % 2), idx = i. */
static char comp_sym(char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c; }

int pgo_mismatch_lists(const char *text, uint64_t text_len, const uint8_t *lq_packed, uint32_t n_lq,
                       const uint8_t *n_packed, uint32_t n_n, uint32_t read_len,
                       const uint64_t *pos, const uint8_t *rc, const uint8_t *mm, int variant,
                       uint64_t *out_offsets, uint8_t *out_off, char *out_pg, char *out_read) {
    reads_view rs = { lq_packed, n_lq, (read_len + 3) / 4, n_packed, n_n, (read_len + 2) / 3, read_len };
    const uint32_t n = n_lq + n_n;
    char *cur = (char *)malloc(read_len + 1);
    if (!cur) return -2;
    uint64_t at = 0;
    for (uint32_t i = 0; i < n; i++) {
        out_offsets[i] = at;
        if (mm[i] == PGO_NOT_MATCHED_COUNT || pos[i] == PGO_NOT_MATCHED_POSITION) continue;
        if (pos[i] + read_len > text_len) { free(cur); return -1; }
        get_read(&rs, i, cur);                                              /* readsSet->getRead */
        if (rc[i])                                                          /* reverseComplementInPlace(currentRead) */
            for (uint32_t a = 0, b = read_len - 1; a <= b && b < read_len; a++, b--) {
                char x = comp_sym(cur[a]), y = comp_sym(cur[b]);
                cur[a] = y; cur[b] = x;
            }
        const char *pg = text + pos[i];
        const int reversed = variant == 0 ? 0 : variant == 1 ? (rc[i] != 0) : ((rc[i] != 0) != (int)(i % 2));
        uint8_t count = 0;
        if (!reversed) {
            for (uint64_t p = 0; count < mm[i] && p < read_len; p++)
                if (cur[p] != pg[p]) { out_off[at] = (uint8_t)p; out_pg[at] = pg[p]; out_read[at] = cur[p]; at++; count++; }
        } else {
            for (uint64_t p = read_len; count < mm[i] && p-- > 0;)
                if (cur[p] != pg[p]) {
                    out_off[at] = (uint8_t)(read_len - p - 1); out_pg[at] = comp_sym(pg[p]); out_read[at] = comp_sym(cur[p]);
                    at++; count++;
                }
        }
        if (count != mm[i]) { free(cur); return -3; }                       /* mm is not the Hamming distance at pos */
    }
    out_offsets[n] = at;
    free(cur);
    return 0;
}

/* ------------------------------------------------------------------ stage 7: exact matches between pseudogenomes (§8(f) rank 4)
 * CopMEMMatcher::matchTexts -> processExactMatchQueryTight (copmem/CopMEMMatcher.cpp:332-481) over the serial index
 * (:139-233, the reference at -t 1), as SimplePgMatcher::exactMatchPg calls it (SimplePgMatcher.cpp:26-55): the source
 * text is indexed once (constructor, :571-593), every k2-th K-mer of the destination text is looked up, the bucket
 * entries are tried in text order, and the first one that extends to at least minMatchLength characters is pushed; the
 * query then jumps `skip` positions ahead — inside the current group of 256 query positions only.
 * A literal, sequential transcription, including
 *   - the "already covered by the previous match" test against resMatches.back() (:388-393),
 *   - the 4-byte guards l1/l2/r1/r2 that keep a stale value when a load would leave the text (:384-385, :396-397),
 *   - the left extension, which stops BEFORE comparing when it reaches the start of either text: the match then begins
 *     one character late (:405).
 * One deliberate difference: the tail loop of the reference loads l1 / r1 without the bounds check of the main loop
 * (:447-448, reading outside the source text for the first / last sampled positions); the oracle keeps the check there too.
 * Result-neutral: the guards only pre-filter, a candidate that extends to minMatchLength has at least (L-K)/2 equal
 * characters on one side, all of them inside both texts, so the guard of that side is loaded and equal. */
typedef struct { uint64_t src, len, dest; } pgo_text_match;

static int copmem_params2(copmem_index *x, uint32_t L, uint32_t min_len, uint64_t N) {
    /* initParams(minMatchLength) with minMatchLength = min(min_len, L) (constructor :574-576) */
    if (min_len > L) min_len = L;
    int K;
    if (L > 110) K = 56; else if (L > 62) K = 44; else if (L > 53) K = 40; else if (L > 46) K = 36;
    else if (L > 42) K = 32; else if (L > 32) K = 28; else K = ((int)L / 4 - 1) * 4;
    if (min_len < 24) return -1;
    const int kmml = ((int)min_len / 4 - 1) * 4;
    if (kmml < K) K = kmml;
    const int t = (int)L - K + 1;
    if (t <= 0) return -1;
    int k1, k2;
    if (t >= 20) {
        k1 = 1; while ((k1 + 1) * (k1 + 1) <= t) k1++;
        k1 += 1; k2 = k1 - 1;
        if (k1 * k2 > t) { --k2; --k1; }
    } else if (t >= 15) { k1 = 5; k2 = 3; } else if (t >= 12) { k1 = 4; k2 = 3; } else if (t >= 10) { k1 = 5; k2 = 2; }
    else if (t >= 6) { k1 = 3; k2 = 2; } else { k1 = t; k2 = 1; }
    x->L = L; x->K = (uint32_t)K; x->k1 = (uint32_t)k1; x->k2 = (uint32_t)k2; x->N = N;
    int i = CM_HASH_MIN_ORDER;
    do { x->hash_size = 1u << (i++); } while (i <= CM_HASH_MAX_ORDER && x->hash_size < N / (uint64_t)k1);
    return 0;
}

static uint32_t load_u32(const char *p) { uint32_t v; memcpy(&v, p, 4); return v; }

typedef struct { pgo_text_match *v; uint64_t n, cap; } match_vec;

static int match_push(match_vec *mv, uint64_t src, uint64_t len, uint64_t dest) {
    if (mv->n == mv->cap) {
        const uint64_t nc = mv->cap ? mv->cap * 2 : 1024;
        pgo_text_match *nv = (pgo_text_match *)realloc(mv->v, nc * sizeof *nv);
        if (!nv) return -2;
        mv->v = nv; mv->cap = nc;
    }
    mv->v[mv->n].src = src; mv->v[mv->n].len = len; mv->v[mv->n].dest = dest; mv->n++;
    return 0;
}

/* One query position (the body shared by the grouped loop :375-418 and the tail loop :424-472).  Returns 1 when the
 * query has to jump ahead (a push, or a position covered by the previous match), 0 otherwise, < 0 on error. */
static int mem_query_position(const copmem_index *x, const char *start2, uint64_t N2, uint64_t q, uint32_t b0, uint32_t b1,
                              int dest_is_src, int rev_compl, uint32_t min_len, uint32_t *l1, uint32_t *l2, uint32_t *r1,
                              uint32_t *r2, match_vec *res, uint64_t *ext) {
    const uint32_t K = x->K;
    const int LK2 = ((int)x->L - (int)K) / 2, K_PLUS_LK24 = (int)K + LK2 - 4;
    const char *start1 = x->text, *end1 = x->text + x->N, *end2 = start2 + N2;
    const char *curr2 = start2 + q;
    if (curr2 - LK2 >= start2) *l2 = load_u32(curr2 - LK2);
    if (curr2 + K_PLUS_LK24 + 4 <= end2) *r2 = load_u32(curr2 + K_PLUS_LK24);
    for (uint32_t j = b0; j < b1; j++) {
        (*ext)++;
        const uint64_t s = x->pos[j];
        const char *curr1 = start1 + s;
        if (dest_is_src && (rev_compl ? N2 - s < q : q >= s)) continue;                                        /* :384-386 */
        if (res->n > 0 && q - s == res->v[res->n - 1].dest - res->v[res->n - 1].src &&
            q + K < res->v[res->n - 1].dest + res->v[res->n - 1].len) return 1;                                /* :388-393 */
        if (curr1 - LK2 >= start1) *l1 = load_u32(curr1 - LK2);
        if (curr1 + K_PLUS_LK24 + 4 <= end1) *r1 = load_u32(curr1 + K_PLUS_LK24);
        if (*r1 == *r2 || *l1 == *l2) {
            const char *p1 = curr1 + K - 1, *p2 = curr2 + K - 1;
            while (++p1 != end1 && ++p2 != end2 && *p1 == *p2) {}
            const char *right = p1;
            p1 = curr1; p2 = curr2;
            while (p1 != start1 && p2 != start2 && *--p1 == *--p2) {}
            if (right - p1 > (long)min_len && memcmp(curr1, curr2, K) == 0) {                                  /* :407 */
                int rcode = match_push(res, (uint64_t)(p1 + 1 - start1), (uint64_t)(right - p1 - 1), (uint64_t)(p2 + 1 - start2));
                return rcode ? rcode : 1;
            }
        }
    }
    return 0;
}

int pgo_match_texts(const char *src, uint64_t n, const char *dest, uint64_t n2, int dest_is_src, int rev_compl,
                    uint32_t target_len, uint32_t min_len, uint64_t *out, uint64_t cap, uint64_t *count, uint64_t *params) {
    copmem_index x;
    memset(&x, 0, sizeof x);
    if (!src || !dest || !count) return -1;
    if (copmem_params2(&x, target_len, min_len, n) != 0 || n < x.K) return -1;
    if (min_len > target_len) min_len = target_len;
    if (min_len < x.K) return -1;                                             /* matchTexts :607-610 */
    int rcode = copmem_build(&x, src);
    if (rcode) { copmem_free(&x); return rcode; }
    if (params) { params[0] = x.K; params[1] = x.k1; params[2] = x.k2; params[3] = x.hash_size; }
    const uint32_t K = x.K, k2 = x.k2, MULTI = 256;
    const uint64_t k2MULTI = (uint64_t)k2 * MULTI;
    const int skip = (int)(K / x.k1) - 1;
    uint32_t l1 = 0, l2 = 0, r1 = 0, r2 = 0;
    uint64_t ext = 0;
    match_vec res = {NULL, 0, 0};
    uint32_t *h_arr = (uint32_t *)malloc(MULTI * sizeof(uint32_t));
    uint64_t i1 = 0;
    for (i1 = 0; i1 + K + k2MULTI < n2 + 1; i1 += k2MULTI) {                  /* :364-419 */
        for (uint32_t t = 0; t < MULTI; t++) h_arr[t] = copmem_hash(&x, dest + i1 + (uint64_t)t * k2);
        for (int64_t t = 0; t < (int64_t)MULTI; t++) {
            const uint32_t b0 = x.cumm[h_arr[t]], b1 = x.cumm[h_arr[t] + 1];
            if (b0 == b1) continue;
            int r = mem_query_position(&x, dest, n2, i1 + (uint64_t)t * k2, b0, b1, dest_is_src, rev_compl, min_len, &l1, &l2, &r1, &r2, &res, &ext);
            if (r < 0) { rcode = r; goto done; }
            if (r) t += skip;                                                 /* curr2 += skipK2; i2 += skip */
        }
    }
    for (; i1 + K < n2 + 1; i1 += k2) {                                       /* :422-473 */
        const uint32_t h = copmem_hash(&x, dest + i1);
        const uint32_t b0 = x.cumm[h], b1 = x.cumm[h + 1];
        if (b0 == b1) continue;
        int r = mem_query_position(&x, dest, n2, i1, b0, b1, dest_is_src, rev_compl, min_len, &l1, &l2, &r1, &r2, &res, &ext);
        if (r < 0) { rcode = r; goto done; }
        if (r) i1 += (uint64_t)skip * k2;
    }
    *count = res.n;
    if (params) params[4] = ext;
    if (res.n > cap) rcode = -3;
    else for (uint64_t i = 0; i < res.n; i++) { out[3 * i] = res.v[i].src; out[3 * i + 1] = res.v[i].len; out[3 * i + 2] = res.v[i].dest; }
done:
    free(h_arr); free(res.v); copmem_free(&x);
    return rcode;
}
