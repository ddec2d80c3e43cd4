/* pgrc_gpu_matcher.h — C ABI of the B200-native read-vs-pseudogenome matcher for PgRC.
 *
 * Drop-in boundary for PgRC's stage 4 (reads -> HQ pseudogenome mapping).  Every entry point
 * names the reference interface it replaces (file:line relative to kowallus/PgRC).  The
 * reference has no FFI: the boundary is the virtual-class interface of
 * PgTools::DefaultReadsMatcher (matching/ReadsMatchers.h:24-83).  INTEGRATION.md shows the
 * C++ subclass a maintainer adds to ReadsMatchers.cpp to call these functions.
 *
 * Conventions
 *  - plain C types only; all buffers are owned by the caller.  Input and output pointers may
 *    be host memory (pageable or pinned) or device memory of the context's GPU; the library
 *    detects which and never writes to an input.
 *  - LIFETIME of host inputs: pgm_set_text / pgm_set_text_shard / pgm_set_reads given a HOST pointer only
 *    record it; the bytes are copied later, chunk by chunk, overlapped with the kernels that consume
 *    them (pgm_match_begin, pgm_scan_pass).  The caller keeps such a buffer alive and unchanged until
 *    pgm_upload, pgm_get_results or pgm_map_reads has returned.  Device inputs are consumed by the
 *    call itself (stream-ordered).
 *  - every function returns PGM_OK (0) or a negative pgm_status; pgm_last_error() gives the
 *    message.  There is NO CPU fallback: without a usable sm_100 device pgm_create fails.
 *  - a context is bound to one GPU and one stream and is not thread-safe (the reference
 *    calls the matcher once, from the main thread: pgrc-encoder.cpp:359).
 *  - multi-GPU: one context per GPU (one process per GPU).  Each context scans its own text
 *    range (pgm_set_text_shard); between pgm_scan_pass and pgm_resolve_pass the caller merges
 *    the per-read accumulators across GPUs (pgm_get_accumulators: MIN / SUM reductions over
 *    NCCL).  With a single GPU, pgm_map_reads does everything.
 */
#ifndef PGRC_GPU_MATCHER_H
#define PGRC_GPU_MATCHER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGM_ABI_VERSION 1

typedef enum pgm_status {
    PGM_OK = 0,
    PGM_ERR_INVALID_ARG = -1,  /* bad pointer / size / parameter combination            */
    PGM_ERR_NO_DEVICE = -2,    /* no CUDA device, or not compute capability 10.x        */
    PGM_ERR_CUDA = -3,         /* a CUDA runtime call failed (message has the detail)   */
    PGM_ERR_OOM = -4,          /* device memory exhausted                               */
    PGM_ERR_BAD_SYMBOL = -5,   /* text contains a symbol outside ACGT                   */
    PGM_ERR_UNSUPPORTED = -6,  /* feature of the reference not covered (see message)    */
    PGM_ERR_STATE = -7         /* call order violated (e.g. scan before match_begin)    */
} pgm_status;

#define PGM_NOT_MATCHED_POSITION UINT64_MAX /* DefaultReadsMatcher::NOT_MATCHED_POSITION, ReadsMatchers.cpp:69 */
#define PGM_NOT_MATCHED_COUNT 255           /* PgTools::NOT_MATCHED_COUNT, ReadsMatchers.h:17                  */
#define PGM_DISABLED_PREFIX_MODE 0xFFFFu    /* DefaultReadsMatcher::DISABLED_PREFIX_MODE, ReadsMatchers.cpp:68 */

typedef struct pgm_ctx pgm_ctx;

/* Result counters.  matched / per_mm reproduce the reference's matchedReadsCount
 * (ReadsMatchers.h:35) and matchedCountPerMismatches (ReadsMatchers.h:116) exactly.
 * betterMatchCount / falseMatchCount (ReadsMatchers.h:42-43) are printed by the reference
 * but never reach the archive and depend on its sequential event order; they are NOT
 * reproduced.  The remaining fields describe the work done (used for the roofline). */
typedef struct pgm_stats {
    uint64_t matched;            /* reads with a match                                   */
    uint64_t per_mm[256];        /* per_mm[k] = reads matched with k mismatches; [255] = unmatched */
    uint64_t patterns_inserted;  /* seeds inserted into the last pattern table           */
    uint64_t table_slots;        /* slots of the last pattern table                      */
    uint64_t candidates;         /* seed hits forwarded to verification (all passes)     */
    uint64_t verified;           /* candidates whose read was actually compared          */
    uint64_t accepted;           /* verified candidates within the mismatch limit        */
    uint64_t filter_positives;   /* text windows that passed the L2-resident pre-filter (all passes) */
} pgm_stats;

/* Device pointers of the per-read accumulators of the current pass, for the cross-GPU
 * merge between pgm_scan_pass and pgm_resolve_pass (n_reads elements each):
 *   best_key  int64  MIN   first_other_order int64 MIN
 *   same_pos_mask int32 SUM (bits are disjoint across text shards)   same_pos_mm uint8 MIN
 * `touched` points to one int32 that is non-zero iff this GPU wrote first_other_order /
 * same_pos_* in this pass (MAX-reduce it first; if 0 everywhere only best_key needs merging). */
typedef struct pgm_accumulators {
    void *best_key;
    void *first_other_order;
    void *same_pos_mask;
    void *same_pos_mm;
    void *touched;
    uint64_t n_reads;
} pgm_accumulators;

/* ---- life cycle ------------------------------------------------------------------------
 * replaces: new/delete of the matcher object in mapReadsIntoPg (ReadsMatchers.cpp:714-794)
 * and the hash matcher it owns (ReadsMatchers.cpp:83-95). */
int pgm_abi_version(void);
int pgm_create(int device, pgm_ctx **out);
void pgm_destroy(pgm_ctx *ctx);
const char *pgm_last_error(const pgm_ctx *ctx); /* ctx may be NULL: last pgm_create error */
/* Use an existing CUDA stream (cudaStream_t) for all work of this context; NULL = own stream
 * (pass cudaStreamLegacy / cudaStreamPerThread to name a default stream). */
int pgm_set_stream(pgm_ctx *ctx, void *cuda_stream);
/* Blocks until all work queued by this context has finished. */
int pgm_synchronize(pgm_ctx *ctx);
/* Completes the lazy upload of host inputs (see LIFETIME above): when it returns the text is packed and the
 * reads are unpacked on the device and the library no longer reads the caller's buffers. */
int pgm_upload(pgm_ctx *ctx);

/* ---- inputs ----------------------------------------------------------------------------
 * pgm_set_text replaces the (char* pgPtr, uint_pg_len_max pgLength) constructor arguments
 * (ReadsMatchers.h:59-60; text = SeparatedPseudoGenome::getPgSequence(), 1 byte per base,
 * alphabet ACGT).  The text is packed on the device into two bit planes per strand; the
 * caller's buffer is only read (the reference reverse-complements it in place and restores
 * it, ReadsMatchers.cpp:167-171 — not needed here). */
int pgm_set_text(pgm_ctx *ctx, const char *text, uint64_t pg_len);
/* Text shard for one GPU of several: `slice` holds text[slice_begin, slice_begin+slice_len);
 * this context owns the seed-window start positions [own_begin, own_end).  The slice must
 * extend PGM_SHARD_HALO bases beyond the owned range on both sides (clipped to the text). */
#define PGM_SHARD_HALO 512
int pgm_set_text_shard(pgm_ctx *ctx, const char *slice, uint64_t slice_begin, uint64_t slice_len,
                       uint64_t pg_len, uint64_t own_begin, uint64_t own_end);
/* Even split of [0, pg_len) over `world` GPUs, with halos; outputs feed pgm_set_text_shard. */
int pgm_shard_plan(uint64_t pg_len, int rank, int world, uint64_t *slice_begin, uint64_t *slice_len,
                   uint64_t *own_begin, uint64_t *own_end);

/* pgm_set_reads replaces the ConstantLengthReadsSetInterface* constructor argument
 * (ReadsMatchers.h:59-60).  lq_packed: n_lq reads of the ACGT set, ceil(read_len/4) bytes
 * each, exactly PackedConstantLengthReadsSet::getPackedRead's layout
 * (PackedConstantLengthReadsSet.h:40, SymbolsPackingFacility.cpp:147-185); n_packed: n_n
 * reads of the ACGNT set, ceil(read_len/3) bytes each.  Global read index = LQ reads first,
 * then N reads (SumOfConstantLengthReadsSets, ReadsSetInterface.h:45-89).  read_len <= 255. */
int pgm_set_reads(pgm_ctx *ctx, const uint8_t *lq_packed, uint32_t n_lq,
                  const uint8_t *n_packed, uint32_t n_n, uint32_t read_len);

/* ---- one matcher object's work, step by step ----------------------------------------------
 * pgm_match_begin replaces initMatching() / initMatchingContinuation()
 * (ReadsMatchers.cpp:97-105,190-196,276-295): builds the seed table in HBM for
 * `parts` seeds of `seed_len` per read.  seed_len == read_len && parts == 1 && max_mm == 0 is
 * the exact path (DefaultReadsExactMatcher).  continuation = 0 resets the per-read state;
 * continuation = 1 keeps it and leaves out reads already matched with <= min_mm mismatches
 * (getMatchedReadsBitmap(minMismatches), ReadsMatchers.cpp:287-295,677-691).
 * max_mm, min_mm <= 127: PgRC's CLI refuses minCharsPerMismatch < 2 (pgrc-params.h:30,254-257), so
 * maxMismatches = readLength / minCharsPerMismatch <= 255 / 2. */
int pgm_match_begin(pgm_ctx *ctx, uint32_t seed_len, uint32_t parts, uint32_t max_mm,
                    uint32_t min_mm, int continuation);
/* pgm_match_begin_interleaved replaces InterleavedReadsApproxMatcher::initMatching() /
 * initMatchingContinuation() (ReadsMatchers.cpp:343-362) and the pattern set of
 * InterleavedConstantLengthPatternsOnTextHashMatcher::addPackedPatterns
 * (ConstantLengthPatternsOnTextHashMatcher.cpp:82-96): seed j of a read = its bases j, j+parts,
 * j+2*parts, ... (seed_len of them), matched against text[x], text[x+parts], ...; a hit aligns
 * the read at x - j (ReadsMatchers.cpp:373-376).  Same arguments as pgm_match_begin; the passes,
 * the accumulators and the decision are the same calls.  parts <= 31. */
int pgm_match_begin_interleaved(pgm_ctx *ctx, uint32_t seed_len, uint32_t parts, uint32_t max_mm,
                                uint32_t min_mm, int continuation);
/* pgm_scan_pass replaces executeMatching(revCompMode) (ReadsMatchers.cpp:198-230,297-341,365-409)
 * up to, but excluding, the per-read decision: scan of the (reverse-complemented, if
 * rev_mode) text, table probes, XOR/popcount verification, per-read accumulators. */
int pgm_scan_pass(pgm_ctx *ctx, int rev_mode);
int pgm_get_accumulators(pgm_ctx *ctx, pgm_accumulators *out);
/* The per-read keys live inside the read records in HBM: pgm_get_accumulators copies them into a
 * contiguous array (the best_key pointer it returns); after the cross-GPU reduction
 * pgm_put_accumulators copies the merged keys back, then pgm_resolve_pass decides. */
int pgm_put_accumulators(pgm_ctx *ctx);
/* pgm_resolve_pass applies the reference's accept/tie-break rule of that pass to the
 * (merged) accumulators: strict improvement, scan order, LIFO among equal-hash patterns,
 * the coordinate-only "already stored" skip of ReadsMatchers.cpp:313, minMismatches stop. */
int pgm_resolve_pass(pgm_ctx *ctx, int rev_mode);
/* pgm_get_results replaces reading readMatchPos / readMatchRC / readMismatchesCount
 * (ReadsMatchers.h:32-33,115) and getMatchedReadsBitmap.  n_reads entries each;
 * stats may be NULL.  Synchronizes the context's stream. */
int pgm_get_results(pgm_ctx *ctx, uint64_t *out_pos, uint8_t *out_rc, uint8_t *out_mm, pgm_stats *stats);

/* ---- matching mode 'c' ---------------------------------------------------------------------
 * pgm_copmem_begin replaces CopMEMReadsApproxMatcher::initMatching / initMatchingContinuation
 * (ReadsMatchers.cpp:411-419): part_len = readsExactMatchingChars (the CopMEMMatcher's target match
 * length, from which K, k1, k2 and the hash size follow: copmem/CopMEMMatcher.cpp:71-137);
 * continuation = 0 resets the per-read state.  pgm_copmem_pass replaces
 * CopMEMReadsApproxMatcher::executeMatching(revCompMode) (ReadsMatchers.cpp:421-451): the index of the
 * (reverse-complemented) text — new CopMEMMatcher(pgPtr, pgLength, partLength), SERIAL build, i.e. the
 * reference at -t 1 (CopMEMMatcher.cpp:139-233) — and the query of every read with more than min_mm
 * mismatches (processApproxMatchQueryTight, :483-566).  No accumulators and no resolve step: reads are
 * independent, the per-read state is final after the pass.  Needs the whole text on this GPU. */
int pgm_copmem_begin(pgm_ctx *ctx, uint32_t part_len, uint32_t max_mm, uint32_t min_mm, int continuation);
int pgm_copmem_pass(pgm_ctx *ctx, int rev_mode);

/* pgm_get_mismatches replaces the per-read work of AbstractReadsApproxMatcher::updateEntry
 * (ReadsMatchers.cpp:555-566: getRead + reverseComplementInPlace + fillEntryWithMismatches,
 * :40-52), i.e. what the reference's export recomputes on the host for every matched read: the
 * list of its mismatches against the pseudogenome at readMatchPos, the read taken
 * reverse-complemented when readMatchRC is set, in ascending read offset.
 *   out_offsets[n_reads + 1]  prefix sums of readMismatchesCount (unmatched reads count 0):
 *                             the entries of read i are [out_offsets[i], out_offsets[i+1])
 *   out_pos[k]                offset in the (possibly reverse-complemented) read
 *   out_syms[k]               pseudogenome symbol : 2 bits | read symbol : 3 bits << 2, codes
 *                             A C G T = 0..3, N = 4 (the caller maps them to mismatch2CxtCode,
 *                             helper.cpp:358; fillEntryWithReversedMismatches, :54-66, is this list
 *                             walked backwards with complemented symbols and offsets L-1-off)
 * `capacity` = entries out_pos / out_syms can hold; the number needed is the sum of
 * k * per_mm[k] (pgm_stats) and is returned in *total (call with null arrays to query it).
 * Host or device pointers.  Needs the whole text on this GPU; call after pgm_get_results. */
int pgm_get_mismatches(pgm_ctx *ctx, uint64_t *out_offsets, uint8_t *out_pos, uint8_t *out_syms,
                       uint64_t capacity, uint64_t *total);

/* ---- routed multi-GPU scheme: every stage divides by the number of GPUs -----------------------------
 * One context per GPU (one process per GPU, or one process driving all of them: pgm_group_* below).  GPU g
 *   - owns the reads [read_begin[g], read_begin[g+1]) (pgm_set_reads is given exactly these): their records, keys,
 *     decision and results are local — no cross-GPU merge of keys;
 *   - owns the part of the seed table whose hashes map to g (its own L2-resident pre-filter);
 *   - holds the whole text (pgm_set_text; the 2-bit planes are pg_len / 4 bytes per strand) and hashes the window
 *     starts of ITS range of the text only.
 * Per phase:  pgm_route_begin -> exchange `send` (patterns, 16-byte entries) -> pgm_route_build.
 * Per pass, for round = 0 .. pgm_route_rounds - 1:
 *     pgm_route_scan  -> exchange (windows, 12-byte entries)    -> pgm_route_probe
 *                     -> exchange (candidates, 12-byte entries) -> pgm_route_verify;
 * then pgm_resolve_pass (local).  An exchange is an all-to-all: segment d of `send` (count[d] entries at
 * base + d * stride_bytes) goes to GPU d, which receives the segments of all senders back to back, in sender
 * order, into pgm_route_recv's buffer (NCCL send/recv over NVLink, or peer copies inside one process).
 * Results (pgm_get_results) are those of the context's own reads.  Replaces, across GPUs, what one
 * DefaultReadsApproxMatcher / DefaultReadsExactMatcher object does (ReadsMatchers.cpp:190-341); modes 'd'/'D'. */
#define PGM_ROUTE_MAX_WORLD 16
enum { PGM_ROUTE_PATTERNS = 0, PGM_ROUTE_WINDOWS = 1, PGM_ROUTE_CANDIDATES = 2 };
typedef struct pgm_route_buffer {
    void *base;                              /* device memory */
    uint64_t stride_bytes;                   /* segment d starts at base + d * stride_bytes */
    uint32_t entry_bytes;
    uint32_t world;
    uint64_t count[PGM_ROUTE_MAX_WORLD];     /* entries for destination d */
} pgm_route_buffer;
/* read_begin[world + 1]: global read ranges of the GPUs; round_windows = window starts a GPU emits per round
 * (0 = default; bounds the exchange buffers). */
int pgm_route_config(pgm_ctx *ctx, int rank, int world, const uint64_t *read_begin, uint64_t round_windows);
int pgm_route_rounds(pgm_ctx *ctx, uint32_t *rounds);
/* The exchange buffers exist twice: pgm_route_scan / _recv / _probe / _verify use the set selected here (0 or 1), so a
 * caller can emit and ship the windows of round r + 1 (other set) while round r is being probed and verified. */
int pgm_route_slot(pgm_ctx *ctx, int slot);
int pgm_route_begin(pgm_ctx *ctx, uint32_t seed_len, uint32_t parts, uint32_t max_mm, uint32_t min_mm,
                    int continuation, pgm_route_buffer *send);
/* The exchange as peer-to-peer copies (copy engines over NVLink, no SM time): after an emit step (pgm_route_begin / _scan /
 * _probe) a rank publishes where its send buffer of that kind lives (pgm_route_export: a CUDA IPC handle for other
 * processes, the pointer itself for contexts of the same process); once every rank's pgm_route_buffer (counts) and
 * pgm_route_peer are known, pgm_route_pull copies this rank's segment of every peer's send buffer into the receive buffer
 * of the current slot, asynchronously on the context's pull stream — the consuming call (pgm_route_build / _probe /
 * _verify) waits for it on the device.  The caller guarantees what any all-to-all needs: a pull starts after every
 * sender's emit step has completed (exchanging the counts implies it: they are read after a stream synchronize), and a
 * sender re-emits into a send buffer only after every receiver has consumed the previous contents (two slots + the count
 * exchanges of the following steps give that order in matcher.run_plan_routed). */
typedef struct pgm_route_peer {
    uint64_t pid;                /* process that owns the buffer */
    void *ptr;                   /* its address there */
    int32_t device;
    int32_t reserved;
    unsigned char ipc_handle[64];
} pgm_route_peer;
int pgm_route_export(pgm_ctx *ctx, int kind, pgm_route_peer *out);
int pgm_route_pull(pgm_ctx *ctx, int kind, const pgm_route_peer *peers /*[world]*/, const pgm_route_buffer *peer_sends /*[world]*/);
/* Device buffer for `n_entries` incoming entries of `kind` (PGM_ROUTE_*). */
int pgm_route_recv(pgm_ctx *ctx, int kind, uint64_t n_entries, void **ptr);
int pgm_route_build(pgm_ctx *ctx, uint64_t n_patterns_in);
int pgm_route_scan(pgm_ctx *ctx, int rev_mode, uint32_t round, pgm_route_buffer *send);
/* in_counts[world]: window entries received from each sender (in sender order in the receive buffer). */
int pgm_route_probe(pgm_ctx *ctx, int rev_mode, uint32_t round, const uint64_t *in_counts, pgm_route_buffer *send);
int pgm_route_verify(pgm_ctx *ctx, int rev_mode, uint64_t n_candidates_in);
/* The emit steps in two halves, for callers that keep the GPU queue filled while they exchange counts: _launch queues the
 * kernels (and the copy of the per-destination counts to pinned host memory); pgm_route_fetch waits for THAT step only —
 * not for work queued behind it — and describes the send buffer.  pgm_route_scan / pgm_route_probe = launch + fetch. */
int pgm_route_scan_launch(pgm_ctx *ctx, int rev_mode, uint32_t round);
int pgm_route_probe_launch(pgm_ctx *ctx, int rev_mode, uint32_t round, const uint64_t *in_counts);
int pgm_route_fetch(pgm_ctx *ctx, int kind, pgm_route_buffer *send);

/* ---- several GPUs behind one handle (one process, one host thread per GPU inside the library) ----------------
 * What the C++ host side of PgRC uses (pgrc_b200/host/GpuReadsMatchers.cpp; PGRC_GPU_DEVICES=0,1,...): the same
 * step-wise calls as a single context, on a group of contexts.  With one device it IS the single-context path
 * (fused scan kernel).  With several: matching modes 'd'/'D' run the routed scheme above, the exchanges being
 * peer-to-peer copies over NVLink (cudaMemcpyPeerAsync between the GPUs' buffers); modes 'i'/'I' and 'c'/'C' give
 * every GPU the whole text and a read range (no exchange).  A device may be listed more than once (several
 * contexts on one GPU: how the single-GPU test box exercises this code).  Inputs are host or device pointers as
 * for a context (device pointers must be accessible from every listed GPU); outputs cover all reads, in the
 * global read order (LQ reads, then N reads).  Errors: pgm_group_last_error. */
typedef struct pgm_group pgm_group;
int pgm_group_create(int n_devices, const int *devices /* NULL = 0 .. n_devices-1 */, pgm_group **out);
void pgm_group_destroy(pgm_group *g);
const char *pgm_group_last_error(const pgm_group *g);   /* g may be NULL: last pgm_group_create error */
int pgm_group_size(const pgm_group *g);
int pgm_group_set_text(pgm_group *g, const char *text, uint64_t pg_len);
int pgm_group_set_reads(pgm_group *g, const uint8_t *lq_packed, uint32_t n_lq, const uint8_t *n_packed, uint32_t n_n,
                        uint32_t read_len);
int pgm_group_upload(pgm_group *g);
/* initMatching / initMatchingContinuation of the exact, default (interleaved = 0) or interleaved matcher */
int pgm_group_match_begin(pgm_group *g, uint32_t seed_len, uint32_t parts, uint32_t max_mm, uint32_t min_mm,
                          int continuation, int interleaved);
/* executeMatching(revCompMode): scan (+ exchanges) + the per-read decision */
int pgm_group_pass(pgm_group *g, int rev_mode);
int pgm_group_copmem_begin(pgm_group *g, uint32_t part_len, uint32_t max_mm, uint32_t min_mm, int continuation);
int pgm_group_copmem_pass(pgm_group *g, int rev_mode);
int pgm_group_get_results(pgm_group *g, uint64_t *out_pos, uint8_t *out_rc, uint8_t *out_mm, pgm_stats *stats);
int pgm_group_get_mismatches(pgm_group *g, uint64_t *out_offsets, uint8_t *out_pos, uint8_t *out_syms,
                             uint64_t capacity, uint64_t *total);

/* ---- the whole stage on one GPU ---------------------------------------------------------
 * pgm_map_reads replaces the matching part of PgTools::mapReadsIntoPg
 * (ReadsMatchers.cpp:693-783) for matching modes 'd'/'D' (DefaultReadsApproxMatcher), 'i'/'I'
 * (InterleavedReadsApproxMatcher) and 'c'/'C' (CopMEMReadsApproxMatcher, results of the reference at
 * -t 1): parameter derivation
 * (:699-713), first matcher (:714-747), optional second phase (:749-779), same argument
 * meaning as the reference (pre_seed = preReadsExactMatchingChars, seed =
 * readsExactMatchingChars, min_chars_per_mismatch = minCharsPerMismatch, upper-case mode
 * letter = shortcut mode).  match_prefix_length must be PGM_DISABLED_PREFIX_MODE (what
 * pgrc-encoder.cpp:360 passes) or >= read_len.  Requires pgm_set_text and pgm_set_reads. */
int pgm_map_reads(pgm_ctx *ctx, uint32_t match_prefix_length, uint32_t pre_seed, uint32_t seed,
                  uint32_t min_chars_per_mismatch, char pre_mode, char mode, int rev_compl,
                  uint64_t *out_pos, uint8_t *out_rc, uint8_t *out_mm, pgm_stats *stats);

/* ---- stage 7: exact matches between pseudogenomes ------------------------------------
 * Replaces CopMEMMatcher as SimplePgMatcher uses it (matching/SimplePgMatcher.cpp:11-55; the TextMatcher interface,
 * matching/TextMatchers.h:54-61).  The source text is the context's text (pgm_set_text; whole text, alphabet ACGT).
 *   pgm_mem_index  = the constructor CopMEMMatcher(srcText, srcLength, targetMatchLength, minMatchLength)
 *                    (copmem/CopMEMMatcher.cpp:571-593): parameters K, k1, k2, hash size (initParams / calcCoprimes,
 *                    :69-137) and the index of every k1-th source position (processRef, :176-231 — the SERIAL build,
 *                    i.e. the reference at -t 1; its multithreaded build orders the buckets differently and races).
 *                    PGM_ERR_UNSUPPORTED where the reference exits (minimal matching length below 24, K outside its
 *                    hash-function table).  params (optional): K, k1, k2, hash size.
 *   pgm_mem_match  = matchTexts(resMatches, destText, destIsSrc, revComplMatching, minMatchLength) (:605-624 ->
 *                    processExactMatchQueryTight, :332-481).  `dest` is the text as the matcher receives it — already
 *                    reverse-complemented by the caller when rev_compl (SimplePgMatcher::exactMatchPg, :33-43) — host
 *                    or device memory, any byte values (symbols outside ACGT match nothing).  dest = NULL with
 *                    dest_is_src: the library takes the source text itself (rev_compl = 0) or its reverse complement
 *                    (rev_compl != 0) from the planes it already holds.  *count = resMatches.size().
 *   pgm_mem_get_matches = resMatches in the reference's push order, `capacity` >= count elements, host or device.
 * min_match_length is clamped to target_match_length as in the reference (:574-575; PgRC passes UINT32_MAX).  A minimal
 * length BELOW the target is PGM_ERR_UNSUPPORTED: the reference's 4-byte guards (:396-398) then reject candidates
 * depending on values left over from earlier candidates, which only a sequential replay reproduces; it is unreachable
 * from pgrc-encoder.cpp:240-245. */
typedef struct pgm_text_match {      /* = PgTools::TextMatch, matching/TextMatchers.h:11-14 */
    uint64_t pos_src_text;
    uint64_t length;
    uint64_t pos_dest_text;
} pgm_text_match;
int pgm_mem_index(pgm_ctx *ctx, uint32_t target_match_length, uint32_t min_match_length, uint32_t *params /*[4] or NULL*/);
int pgm_mem_match(pgm_ctx *ctx, const char *dest, uint64_t dest_len, int dest_is_src, int rev_compl,
                  uint32_t min_match_length, uint64_t *count);
int pgm_mem_get_matches(pgm_ctx *ctx, pgm_text_match *out, uint64_t capacity);

/* One rank of several (one process per GPU, the source text and its index on every GPU): the matches of the part-th of
 * n_parts shares of the groups of 256 query positions, in push order but BEFORE the "covered by the previous match" test
 * (:388-393), with the query position of each — the caller concatenates the shares in rank order and drops an element whose
 * diagonal (pos_dest - pos_src) equals its predecessor's and whose query position + K lies below the predecessor's
 * pos_dest + length (pgrc_b200/matcher.py: merge_text_match_shares; pgm_group_mem_match does the same inside one process). */
int pgm_mem_match_share(pgm_ctx *ctx, const char *dest, uint64_t dest_len, int dest_is_src, int rev_compl,
                        uint32_t min_match_length, int part, int n_parts, uint64_t *count);
int pgm_mem_get_share(pgm_ctx *ctx, pgm_text_match *out_matches, uint64_t *out_query_pos, uint64_t capacity);

/* The same on a group of contexts (pgm_group_set_text gave every GPU the source text): every GPU builds the index; the groups
 * of 256 query positions of a destination text are independent, GPU r takes the r-th share of them, and the shares are
 * concatenated (the "covered by the previous match" test, :388-393, runs across the seams).  Same results as one context. */
int pgm_group_mem_index(pgm_group *g, uint32_t target_match_length, uint32_t min_match_length, uint32_t *params /*[4] or NULL*/);
int pgm_group_mem_match(pgm_group *g, const char *dest, uint64_t dest_len, int dest_is_src, int rev_compl,
                        uint32_t min_match_length, uint64_t *count);
int pgm_group_mem_get_matches(pgm_group *g, pgm_text_match *out, uint64_t capacity);

/* ---- introspection (bench / tests) -------------------------------------------------------*/
/* Number of kernels this context has launched so far. */
uint64_t pgm_kernel_launches(const pgm_ctx *ctx);
/* Per-kernel device times, for the roofline in bench.py.  With profiling on, every kernel
 * launch of this context is bracketed by CUDA events on the context's stream; pgm_get_timings
 * synchronizes, adds up the event durations per kernel since the last call and resets them.
 * launches[] counts the launches behind each sum. */
enum { PGM_K_PACK_TEXT = 0, PGM_K_RC_TEXT, PGM_K_UNPACK_READS, PGM_K_INIT_STATE, PGM_K_BUILD_TABLE,
       PGM_K_SCAN, PGM_K_RESOLVE, PGM_K_FINALIZE, PGM_K_ACCUM,
       PGM_K_SCAN_FILTER, PGM_K_SCAN_PROBE, PGM_K_SCAN_VERIFY, /* the three stages of the L2-blocked scan pipeline */
       PGM_K_MISMATCHES, PGM_K_COPMEM_INDEX, PGM_K_COPMEM_QUERY,
       PGM_K_ROUTE_BUILD, PGM_K_ROUTE_SCAN, PGM_K_ROUTE_PROBE, PGM_K_ROUTE_VERIFY, /* the routed multi-GPU scheme */
       PGM_K_MEM_PACK, PGM_K_MEM_QUERY, PGM_K_MEM_EMIT,                            /* stage 7 (the index is PGM_K_COPMEM_INDEX) */
       PGM_K_COPMEM_STAGE1, PGM_K_COPMEM_STAGE2,                                   /* the staged mode-c query (PGM_CM_WARP=1) */
       PGM_K_COUNT };
typedef struct pgm_timings {
    double ms[PGM_K_COUNT];
    uint64_t launches[PGM_K_COUNT];
} pgm_timings;
int pgm_set_profiling(pgm_ctx *ctx, int on);
int pgm_get_timings(pgm_ctx *ctx, pgm_timings *out);
/* Tuning knobs; call before pgm_match_begin.  filter_log2_bits = 0 disables the L2-resident
 * pre-filter (< 0 = auto); slots_per_pattern sets the table size (>= 2); l2_hints: 0 = none, 1 = the
 * filter is loaded with an L2 evict_last policy and table buckets / read records with evict_first,
 * 2 = evict_first on buckets / records plus a persisting L2 access-policy window over the filter. */
int pgm_set_tuning(pgm_ctx *ctx, int filter_log2_bits, int slots_per_pattern, int ctas_per_sm, int l2_hints);

#ifdef __cplusplus
}
#endif
#endif /* PGRC_GPU_MATCHER_H */
