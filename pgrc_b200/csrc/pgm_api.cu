// pgm_api.cu — C ABI (include/pgrc_gpu_matcher.h) over the kernels in pgm_kernels.cuh.
// Host-side plumbing only: device buffers, stream-ordered launches, parameter derivation
// restated from PgTools::mapReadsIntoPg (matching/ReadsMatchers.cpp:693-783).
#include "../../include/pgrc_gpu_matcher.h"
#include "pgm_kernels.cuh"
#include "pgm_blocked.cuh"
#include "pgm_part.cuh"
#include "pgm_copmem.cuh"
#include "pgm_copmem_warp.cuh"
#include "pgm_mem.cuh"
#include "pgm_routed.cuh"

#include <algorithm>
#include <array>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

constexpr uint64_t TEXT_CHUNK_BASES = 16ull << 20; // H2D / pack chunk of the ASCII text (multiple of the tile size)
constexpr uint32_t READS_CHUNK = 1u << 20;          // reads per H2D / unpack / table-build chunk

} // namespace

struct pgm_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    uint64_t launches = 0;

    // per-kernel event timing (pgm_set_profiling)
    struct EvPair { int kind; cudaEvent_t a, b; };
    bool profiling = false;
    std::vector<EvPair> ev_used;
    std::vector<cudaEvent_t> ev_free;

    // tuning
    int filter_log2_bits = -1; // -1 = auto
    int slots_per_pattern = 3;
    int ctas_per_sm = 4;
    int two_step_build = 1;     // region-queued table build (PGM_TWO_STEP_BUILD=0 turns it off)
    int l2_hints = 1;           // 0 none, 1 per-load eviction hints, 2 hints + persisting access-policy window on the filter
    int insert_prefetch = 1;    // build_insert_kernel prefetches the next region's buckets into the L2 (PGM_INSERT_PREFETCH)
    int blocked_scan = 0;       // L2-blocked scan pipeline: 0 off, 1 auto (by size), 2 always, 3 always with tiny queues (PGM_BLOCKED_SCAN; tests)
    int region_mb = 12, range_mb = 16;   // target sizes of a table region / a read range of the pipeline (PGM_REGION_MB, PGM_RANGE_MB)
    int part_scan = 0;          // partitioned exact pre-filter (pgm_part.cuh), opt-in (measured slower at config 5): 0 off, 1 auto (pattern sets beyond the Bloom filter), 2 always, 3 always with tiny queues (PGM_PART_SCAN; tests)
    int part_mb = 48;           // target size of a table partition (PGM_PART_MB)
    int cm_warp = 0;            // mode 'c': the staged warp-per-read query (pgm_copmem_warp.cuh) instead of the thread-per-read kernel (PGM_CM_WARP)
    size_t persist_max = 0, window_max = 0, persist_set = 0;

    // text
    DevBuf f_lo, f_hi, r_lo, r_hi, ascii_stage;
    uint64_t pg_len = 0, slice_begin = 0, slice_len = 0, own_begin = 0, own_end = 0;
    bool has_text = false;

    // host inputs are uploaded lazily, in chunks on a copy stream, so that the H2D copies overlap the unpack /
    // table-build / forward-scan kernels of the chunks that have already arrived (pgm_match_begin, pgm_scan_pass)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t fence_ev = nullptr;
    std::vector<cudaEvent_t> text_ev, reads_ev, mem_ev;
    const uint8_t *h_text = nullptr;
    bool text_pending = false, text_copies_enqueued = false;
    const uint8_t *h_lq = nullptr, *h_n = nullptr;
    bool reads_pending = false;
    // PAGEABLE host inputs (what PgRC's std::string / std::vector are): cudaMemcpyAsync would stage them inside the driver on
    // one thread at 8 - 10 GB/s and block the caller meanwhile.  Instead each chunk is copied by several host threads into one
    // of four pinned slots and sent from there, while the kernels of the chunks before it run (h2d_chunk).
    struct HostStage { void *buf[4] = {nullptr, nullptr, nullptr, nullptr}; cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
                       bool busy[4] = {false, false, false, false}; size_t bytes = 0; int next = 0; } hstage;
    bool text_pageable = false;         // the pending host text is pageable: its chunks are staged one by one as the scan asks for them
    uint64_t text_chunks_sent = 0;

    // reads: one record per read (header {state, key} + bit planes), see pgm_kernels.cuh
    DevBuf packed_stage, lq_recs, n_recs;
    uint32_t n_lq = 0, n_n = 0, read_len = 0, W = 0, lq_stride16 = 0, n_stride16 = 0;
    bool has_reads = false;
    bool state_fresh = false;   // record headers are {unmatched, no key} (just unpacked)
    bool aux_clean = false;     // first_order / same_mask / same_mm hold their neutral values
    bool outputs_valid = false; // out_pos / out_rc / out_mm / hist reflect the current record states

    // per read (outside the records: rarely touched)
    DevBuf first_order, same_mask, same_mm, touched, keys;

    // table
    DevBuf buckets, next, filter, bq_entries, bq_counters;
    DevBuf sq_pos, sq_cand, sq_counters;   // queues of the L2-blocked scan pipeline (pgm_blocked.cuh)
    DevBuf pq_entries, pq_counters, pq_bits;   // partition queues and hit bitmap of the partitioned pre-filter (pgm_part.cuh)
    DevBuf mis_sums, mis_offsets, mis_pos, mis_syms;   // mismatch lists (pgm_get_mismatches)
    // mode 'c' (pgm_copmem.cuh): per-pass text index + parameters of the current CopMEM phase
    DevBuf cm_count, cm_start, cm_cumm, cm_fill, cm_hash, cm_all, cm_entries, cm_sums, cm_nib, cm_coarse;
    DevBuf cmw_lens, cmw_cand, cmw_vt, cmw_nvt;     // per-read candidate streams of the staged query (pgm_copmem_warp.cuh)
    uint32_t cm_K = 0, cm_k1 = 0, cm_k2 = 0, cm_hash_size = 0;
    bool copmem_active = false;
    // stage 7 (pgm_mem.cuh): index of the context's text in the cm_* buffers + the destination text and the match lists
    DevBuf mem_dlo, mem_dhi, mem_dinv, mem_stage, mem_fv, mem_has, mem_emit, mem_gcount, mem_gstart, mem_raw, mem_rawq, mem_keep, mem_kstart, mem_out;
    bool mem_index_valid = false;
    uint32_t mem_L = 0, mem_K = 0, mem_k1 = 0, mem_k2 = 0, mem_hash_size = 0;
    uint64_t mem_count = 0;
    bool mem_result_valid = false, mem_share_valid = false;
    uint32_t bq_cap = 0, bq_region_bits = 0;
    bool bq_pending = false;    // region queues hold patterns that build_insert_kernel has not inserted yet
    uint64_t n_slots = 0;
    uint32_t n_buckets = 0;
    uint32_t filter_words = 0;  // 0 = no filter
    uint32_t filter_slice_bits = 0;   // log2 of the number of 64 MB slices of the filter (hash-sliced scan, pgm_kernels.cuh TableView)
    int filter_slices_force = -1;     // PGM_FILTER_SLICES: log2 of the slice count (-1 = auto)
    uint32_t filter_k = 2;      // bits per pattern in its filter word
    int filter_k_force = 0;     // PGM_FILTER_K (sweeps)
    int filter_pair_mode = 1;   // paired filter lookups: 0 off, 1 auto (by load), 2 always (PGM_FILTER_PAIR)
    uint32_t filter_pair = 0, pair_lo = 0, pair_mask = 0;

    // phase
    uint32_t seed_len = 0, parts = 0, max_mm = 0, min_mm = 0, part_bits = 0;
    bool interleaved = false;   // mode 'i': seed j = read bases j, j + parts, j + 2 parts, ...
    bool phase_active = false;

    uint32_t ilv() const { return interleaved ? parts : 0u; }                        // stride of the seeds, 0 = contiguous
    uint32_t shift_unit() const { return interleaved ? 1u : seed_len; }              // alignment start = window start - j * shift_unit
    uint64_t seed_span() const { return (uint64_t)seed_len * (interleaved ? parts : 1u); }   // text bases a seed window covers

    // routed multi-GPU scheme (pgm_routed.cuh)
    struct Route {
        int rank = 0, world = 0;            // world = 0: not configured
        uint64_t read_begin[PGM_ROUTE_MAX_WORLD + 1] = {0};
        uint64_t round_windows = 0;         // window starts a GPU emits per round
        uint32_t cap_pat = 0, cap_win = 0, cap_cand = 0;
        uint64_t n_pat_in = 0;
        int slot = 0;                       // which of the two sets of exchange buffers the route_* calls use (pgm_route_slot)
    } route;
    // two sets of exchange buffers: the windows of round r + 1 can be emitted and travel while round r is probed and verified
    DevBuf rt_win_send2[2], rt_win_recv2[2], rt_cand_send2[2], rt_cand_recv2[2], rt_pat_recv, rt_counters, rt_live, rt_live_count;
    // exchange by peer copies (pgm_route_export / pgm_route_pull): copies run on their own stream, one "arrived" event per
    // kind and slot; peer buffers opened through CUDA IPC are cached by handle
    cudaStream_t pull_stream = nullptr;          // carries the "arrived" events; the copies fan out over pull_more[] too (several copy engines)
    cudaStream_t pull_more[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t pull_join[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t pull_ev[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    bool pull_pending[3][2] = {{false, false}, {false, false}, {false, false}};
    cudaEvent_t consumed_ev[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};   // last reader of a receive buffer
    unsigned int *rt_counts_host = nullptr;      // pinned: the counts of the emit steps, per slot and kind
    cudaEvent_t counts_ev[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    std::vector<std::pair<std::string, void *>> ipc_open;
    DevBuf &rt_win_send_() { return rt_win_send2[route.slot]; }
    DevBuf &rt_win_recv_() { return rt_win_recv2[route.slot]; }
    DevBuf &rt_cand_send_() { return rt_cand_send2[route.slot]; }
    DevBuf &rt_cand_recv_() { return rt_cand_recv2[route.slot]; }

    // misc device scalars: counters[0..3] scan, [4] inserted, [5] tile counter (low 32 bits)
    DevBuf counters, hist, err_flag;
    DevBuf out_pos, out_rc, out_mm;

    uint32_t n_reads() const { return n_lq + n_n; }
};

namespace {

int fail(pgm_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

int cuda_fail(pgm_ctx *c, cudaError_t e, const char *what) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    return fail(c, e == cudaErrorMemoryAllocation ? PGM_ERR_OOM : PGM_ERR_CUDA, buf);
}

#define CU(call)                                                         \
    do {                                                                 \
        cudaError_t _e = (call);                                         \
        if (_e != cudaSuccess) return cuda_fail(ctx, _e, #call);         \
    } while (0)

#define LAUNCH_CHECK(name)                                               \
    do {                                                                 \
        ctx->launches++;                                                 \
        cudaError_t _e = cudaGetLastError();                             \
        if (_e != cudaSuccess) return cuda_fail(ctx, _e, name);          \
    } while (0)

cudaEvent_t ev_get(pgm_ctx *c) {
    if (!c->ev_free.empty()) { cudaEvent_t e = c->ev_free.back(); c->ev_free.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// Launch wrapper: counts the launch, checks it, and (profiling on) brackets it with events.
#define KLAUNCH(kind, name, ...)                                         \
    do {                                                                 \
        cudaEvent_t _a = nullptr, _b = nullptr;                          \
        if (ctx->profiling) { _a = ev_get(ctx); _b = ev_get(ctx); cudaEventRecord(_a, ctx->stream); } \
        __VA_ARGS__;                                                     \
        if (ctx->profiling) { cudaEventRecord(_b, ctx->stream); ctx->ev_used.push_back({kind, _a, _b}); } \
        LAUNCH_CHECK(name);                                              \
    } while (0)

int ensure(pgm_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap && b.p) return PGM_OK;
    if (b.p) { CU(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    size_t want = std::max<size_t>(bytes, 256);
    CU(cudaMalloc(&b.p, want));
    b.cap = want;
    return PGM_OK;
}

void release(DevBuf &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
}

bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

inline unsigned int grid_for(uint64_t n, unsigned int block) { return (unsigned int)((n + block - 1) / block); }

uint64_t plane_words(uint64_t slice_len) { return (slice_len + 31) / 32; }
size_t plane_bytes(uint64_t slice_len) { return (PGM_PAD_WORDS + plane_words(slice_len) + PGM_TAIL_WORDS) * sizeof(uint32_t); }

pgm::PerRead per_read(pgm_ctx *c) {
    pgm::PerRead pr;
    pr.first_other_order = c->first_order.as<long long>();
    pr.same_pos_mask = c->same_mask.as<int>();
    pr.same_pos_mm = c->same_mm.as<uint8_t>();
    pr.touched = c->touched.as<int>();
    return pr;
}

pgm::ReadsView reads_view(pgm_ctx *c) {
    pgm::ReadsView rv;
    rv.lq = c->lq_recs.as<uint4>();
    rv.nn = c->n_recs.as<uint4>();
    rv.n_lq = c->n_lq; rv.n_n = c->n_n;
    rv.lq_stride16 = c->lq_stride16; rv.n_stride16 = c->n_stride16;
    rv.read_len = c->read_len; rv.W = c->W;
    rv.part_bits = c->part_bits;
    return rv;
}

int floor_log2_u(uint64_t v) { int b = 0; while ((2ull << b) <= v) b++; return b; }

pgm::TableView table_view(pgm_ctx *c) {
    pgm::TableView tv;
    tv.buckets = c->buckets.as<pgm::u32x8>();
    tv.next = c->next.as<uint32_t>();
    tv.filter = c->filter_words ? c->filter.as<uint32_t>() : nullptr;
    tv.n_buckets = c->n_buckets;
    tv.filter_mask = c->filter_words ? c->filter_words - 1 : 0;
    tv.filter_k = c->filter_k;
    tv.pair = c->filter_words ? c->filter_pair : 0; tv.pair_lo = c->pair_lo; tv.pair_mask = c->pair_mask;
    // slice = top filter_slice_bits bits of the word index
    tv.slice_shift = (c->filter_words && c->filter_slice_bits) ? (uint32_t)floor_log2_u(c->filter_words) - c->filter_slice_bits : 31u;
    return tv;
}

int ceil_log2(uint64_t v) { int b = 0; while ((1ull << b) < v) b++; return b; }


uint32_t next_prime(uint64_t v) {
    if (v < 3) return 2;
    if (!(v & 1)) v++;
    for (;; v += 2) {
        bool prime = true;
        for (uint64_t d = 3; d * d <= v; d += 2)
            if (v % d == 0) { prime = false; break; }
        if (prime) return (uint32_t)v;
    }
}

cudaEvent_t chunk_event(std::vector<cudaEvent_t> &pool, size_t i) {
    while (pool.size() <= i) {
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        pool.push_back(e);
    }
    return pool[i];
}

bool is_pageable_host(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

void parallel_memcpy(void *dst, const void *src, size_t n) {
    static const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t per_min = 2u << 20;
    const unsigned want = (unsigned)std::min<size_t>(std::min<unsigned>(8u, std::max(1u, hw / 2)), std::max<size_t>(1, n / per_min));
    if (want <= 1) { memcpy(dst, src, n); return; }
    std::vector<std::thread> th;
    const size_t per = ((n + want - 1) / want + 63) & ~(size_t)63;
    for (unsigned k = 1; k < want; k++) {
        const size_t b = std::min(n, k * per), e = std::min(n, (k + 1) * per);
        if (e > b) th.emplace_back([=] { memcpy(static_cast<char *>(dst) + b, static_cast<const char *>(src) + b, e - b); });
    }
    memcpy(dst, src, std::min(n, per));
    for (auto &t : th) t.join();
}

// four pinned slots of at least `bytes` each
int ensure_hstage(pgm_ctx *ctx, size_t bytes) {
    pgm_ctx::HostStage &hs = ctx->hstage;
    if (hs.bytes >= bytes) return PGM_OK;
    for (int k = 0; k < 4; k++) {
        if (hs.busy[k]) { CU(cudaEventSynchronize(hs.ev[k])); hs.busy[k] = false; }
        if (hs.buf[k]) { CU(cudaFreeHost(hs.buf[k])); hs.buf[k] = nullptr; }
    }
    hs.bytes = (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
    for (int k = 0; k < 4; k++) {
        CU(cudaHostAlloc(&hs.buf[k], hs.bytes, cudaHostAllocDefault));
        if (!hs.ev[k]) CU(cudaEventCreateWithFlags(&hs.ev[k], cudaEventDisableTiming));
    }
    return PGM_OK;
}

// A result array to the caller, stream-ordered behind the kernels that produce it.  Device and pinned destinations: one
// asynchronous copy.  Pageable destinations: the copy engine fills the pinned slots (up to four copies in flight) and host
// threads move each slot on into the caller's memory — the driver's own staging does this on one thread.
int d2h_out(pgm_ctx *ctx, void *dst, const void *src, size_t n) {
    if (n < (4u << 20) || is_device_ptr(dst) || !is_pageable_host(dst)) {
        CU(cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, ctx->stream));
        return PGM_OK;
    }
    pgm_ctx::HostStage &hs = ctx->hstage;
    { int rc = ensure_hstage(ctx, std::min<size_t>(n, 32u << 20)); if (rc) return rc; }
    for (int k = 0; k < 4; k++) if (hs.busy[k]) { CU(cudaEventSynchronize(hs.ev[k])); hs.busy[k] = false; }
    const size_t chunk = hs.bytes, n_chunks = (n + chunk - 1) / chunk;
    size_t issued = 0, done = 0;
    while (done < n_chunks) {
        while (issued < n_chunks && issued - done < 4) {
            const size_t off = issued * chunk, len = std::min(chunk, n - off);
            CU(cudaMemcpyAsync(hs.buf[issued & 3], static_cast<const char *>(src) + off, len, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaEventRecord(hs.ev[issued & 3], ctx->stream));
            issued++;
        }
        const size_t off = done * chunk, len = std::min(chunk, n - off);
        CU(cudaEventSynchronize(hs.ev[done & 3]));
        parallel_memcpy(static_cast<char *>(dst) + off, hs.buf[done & 3], len);
        done++;
    }
    return PGM_OK;
}

// One chunk of a host input to the device on the copy stream; `done` is recorded behind it.  Pinned / registered memory goes
// straight to the copy engine; pageable memory through the context's pinned slots (see pgm_ctx::hstage).
int h2d_chunk(pgm_ctx *ctx, void *dst, const void *src, size_t n, cudaEvent_t done, bool pageable) {
    if (!pageable || n < (1u << 20)) {                 // (small pageable copies: the driver's own staging is as good)
        CU(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, ctx->copy_stream));
        CU(cudaEventRecord(done, ctx->copy_stream));
        return PGM_OK;
    }
    pgm_ctx::HostStage &hs = ctx->hstage;
    { int rc = ensure_hstage(ctx, n); if (rc) return rc; }
    const int k = hs.next;
    hs.next = (hs.next + 1) & 3;
    if (hs.busy[k]) { CU(cudaEventSynchronize(hs.ev[k])); hs.busy[k] = false; }      // the DMA that last read this slot
    parallel_memcpy(hs.buf[k], src, n);
    CU(cudaMemcpyAsync(dst, hs.buf[k], n, cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(cudaEventRecord(hs.ev[k], ctx->copy_stream));
    hs.busy[k] = true;
    CU(cudaEventRecord(done, ctx->copy_stream));
    return PGM_OK;
}

// The copy stream may overwrite the staging buffers only after everything queued so far on the main stream
// (the kernels of the previous call that read them) has finished.
int fence_copy_stream(pgm_ctx *ctx) {
    CU(cudaEventRecord(ctx->fence_ev, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->fence_ev, 0));
    return PGM_OK;
}

uint64_t text_chunks(const pgm_ctx *ctx) { return (ctx->slice_len + TEXT_CHUNK_BASES - 1) / TEXT_CHUNK_BASES; }

// Queue the H2D copies of a pending host text, chunk by chunk, on the copy stream.
int enqueue_text_copies(pgm_ctx *ctx, bool fence) {
    if (!ctx->text_pending || ctx->text_copies_enqueued) return PGM_OK;
    if (fence) { int rc = fence_copy_stream(ctx); if (rc) return rc; }
    ctx->text_pageable = ctx->slice_len != 0 && is_pageable_host(ctx->h_text);
    ctx->text_chunks_sent = 0;
    ctx->text_copies_enqueued = true;
    if (ctx->text_pageable) return PGM_OK;               // staged chunk by chunk when the consumer asks (text_chunk_ready)
    for (uint64_t c = 0, off = 0; off < ctx->slice_len; c++, off += TEXT_CHUNK_BASES) {
        const uint64_t n = std::min<uint64_t>(TEXT_CHUNK_BASES, ctx->slice_len - off);
        int rc = h2d_chunk(ctx, ctx->ascii_stage.as<uint8_t>() + off, ctx->h_text + off, n, chunk_event(ctx->text_ev, c), false);
        if (rc) return rc;
    }
    ctx->text_chunks_sent = text_chunks(ctx);
    return PGM_OK;
}

// Chunk c of a pending host text is on its way (its event recorded): a pageable text is staged here, in order, so that the
// host copies of chunk c + 1 run while the kernels of chunk c do.
int text_chunk_ready(pgm_ctx *ctx, uint64_t c) {
    while (ctx->text_chunks_sent <= c) {
        const uint64_t k = ctx->text_chunks_sent, off = k * TEXT_CHUNK_BASES;
        const uint64_t n = std::min<uint64_t>(TEXT_CHUNK_BASES, ctx->slice_len - off);
        int rc = h2d_chunk(ctx, ctx->ascii_stage.as<uint8_t>() + off, ctx->h_text + off, n, chunk_event(ctx->text_ev, k), true);
        if (rc) return rc;
        ctx->text_chunks_sent++;
    }
    return PGM_OK;
}

// Pack chunk c of the text (device-resident source, or the staged copy of a pending host text) into the planes.
int pack_text_chunk(pgm_ctx *ctx, const uint8_t *src_base, uint64_t c) {
    const uint64_t off = c * TEXT_CHUNK_BASES;
    const uint64_t n = std::min<uint64_t>(TEXT_CHUNK_BASES, ctx->slice_len - off);
    uint32_t *flo = ctx->f_lo.as<uint32_t>() + PGM_PAD_WORDS, *fhi = ctx->f_hi.as<uint32_t>() + PGM_PAD_WORDS;
    KLAUNCH(PGM_K_PACK_TEXT, "pack_text_kernel", pgm::pack_text_kernel<<<grid_for((n + 31) / 32, 256), 256, 0, ctx->stream>>>(
        src_base + off, n, flo, fhi, off / 32, ctx->err_flag.as<int>()));
    return PGM_OK;
}

int rc_text(pgm_ctx *ctx) {
    if (!ctx->slice_len) return PGM_OK;
    uint32_t *flo = ctx->f_lo.as<uint32_t>() + PGM_PAD_WORDS, *fhi = ctx->f_hi.as<uint32_t>() + PGM_PAD_WORDS;
    uint32_t *rlo = ctx->r_lo.as<uint32_t>() + PGM_PAD_WORDS, *rhi = ctx->r_hi.as<uint32_t>() + PGM_PAD_WORDS;
    KLAUNCH(PGM_K_RC_TEXT, "rc_text_kernel", pgm::rc_text_kernel<<<grid_for(plane_words(ctx->slice_len), 256), 256, 0, ctx->stream>>>(
        flo, fhi, ctx->slice_len, rlo, rhi));
    return PGM_OK;
}

// Complete a pending host-text upload without overlapping it with a scan.
int finish_text_upload(pgm_ctx *ctx) {
    if (!ctx->text_pending) return PGM_OK;
    int rc = enqueue_text_copies(ctx, true);
    if (rc) return rc;
    for (uint64_t c = 0; c < text_chunks(ctx); c++) {
        if ((rc = text_chunk_ready(ctx, c))) return rc;
        CU(cudaStreamWaitEvent(ctx->stream, ctx->text_ev[c], 0));
        if ((rc = pack_text_chunk(ctx, ctx->ascii_stage.as<uint8_t>(), c))) return rc;
    }
    ctx->text_pending = false;
    return rc_text(ctx);
}

struct ReadsPart { const uint8_t *src; uint32_t cnt, plen, stride16; int with_n; uint4 *dst; uint32_t first_read; size_t stage_off; bool host; };

void reads_parts(pgm_ctx *ctx, const uint8_t *lq, const uint8_t *nn, ReadsPart out[2]) {
    const uint32_t lq_plen = (ctx->read_len + 3) / 4, n_plen = (ctx->read_len + 2) / 3;
    const size_t lq_bytes = (size_t)ctx->n_lq * lq_plen;
    out[0] = {lq, ctx->n_lq, lq_plen, ctx->lq_stride16, 0, ctx->lq_recs.as<uint4>(), 0, 0, ctx->n_lq && !is_device_ptr(lq)};
    out[1] = {nn, ctx->n_n, n_plen, ctx->n_stride16, 1, ctx->n_recs.as<uint4>(), ctx->n_lq, (lq_bytes + 15) & ~(size_t)15,
              ctx->n_n && !is_device_ptr(nn)};
}

int unpack_range(pgm_ctx *ctx, const ReadsPart &pt, const uint8_t *src, uint32_t first, uint32_t cnt) {
    const unsigned int threads = 128;
    KLAUNCH(PGM_K_UNPACK_READS, "unpack_reads_kernel",
            pgm::unpack_reads_kernel<<<grid_for(cnt, threads), threads, ((threads * pt.plen + 15) & ~15u) + threads * pt.stride16 * 16, ctx->stream>>>(
                src + (size_t)first * pt.plen, cnt, ctx->read_len, pt.plen, pt.with_n, pt.dst + (size_t)first * pt.stride16, pt.stride16, ctx->W));
    return PGM_OK;
}

pgm::BuildQueues build_queues(pgm_ctx *c) {
    pgm::BuildQueues q;
    q.entries = c->bq_entries.as<uint4>();
    q.count = c->bq_counters.as<unsigned int>();
    q.cursor = c->bq_counters.as<unsigned int>() + PGM_MAX_REGIONS;
    q.cap = c->bq_cap;
    q.region_bits = c->bq_region_bits;
    q.route_world = 0; q.read_base = 0; q.overflow = nullptr;
    return q;
}

// Step 1 of the table build for reads [r_begin, r_end): seeds -> region queues (or straight into a small table).
int build_range(pgm_ctx *ctx, uint32_t r_begin, uint32_t r_end, int continuation) {
    if (r_end <= r_begin) return PGM_OK;
    const uint32_t tail = ctx->seed_len % 32 ? (1u << (ctx->seed_len % 32)) - 1u : 0xFFFFFFFFu;
    const unsigned int grid = (unsigned int)std::min<uint64_t>(grid_for(r_end - r_begin, PGM_BUILD_THREADS), (uint64_t)ctx->sm_count * 8);
    const size_t smem = ctx->bq_region_bits ? (size_t)ctx->parts * PGM_BUILD_THREADS * sizeof(uint4) : 0;
    const bool fast = ctx->n_n == 0 && ctx->lq_stride16 == 4;         // only ACGT reads, 64-byte records: record held in registers
    KLAUNCH(PGM_K_BUILD_TABLE, "build_table_kernel",
            if (fast) pgm::build_table_kernel<true><<<grid, PGM_BUILD_THREADS, smem, ctx->stream>>>(
                reads_view(ctx), table_view(ctx), build_queues(ctx), r_begin, r_end, ctx->seed_len, ctx->parts, ctx->min_mm, continuation, tail, ctx->ilv(),
                ctx->counters.as<unsigned long long>() + 4);
            else pgm::build_table_kernel<false><<<grid, PGM_BUILD_THREADS, smem, ctx->stream>>>(
                reads_view(ctx), table_view(ctx), build_queues(ctx), r_begin, r_end, ctx->seed_len, ctx->parts, ctx->min_mm, continuation, tail, ctx->ilv(),
                ctx->counters.as<unsigned long long>() + 4));
    if (ctx->bq_region_bits) ctx->bq_pending = true;
    return PGM_OK;
}

// Step 2: insert the queued patterns region by region.
int build_flush(pgm_ctx *ctx) {
    if (!ctx->bq_pending) return PGM_OK;
    KLAUNCH(PGM_K_BUILD_TABLE, "build_insert_kernel", pgm::build_insert_kernel<<<ctx->sm_count * 8, PGM_INSERT_THREADS, 0, ctx->stream>>>(
        table_view(ctx), build_queues(ctx), ctx->insert_prefetch));
    ctx->bq_pending = false;
    return PGM_OK;
}

// Upload + unpack pending host reads; when `build` the table inserts of a chunk follow its unpack, so both overlap
// the H2D copy of the next chunk.  The pending text's copies are queued right behind the reads'.
int upload_reads(pgm_ctx *ctx, bool build) {
    if (!ctx->reads_pending) return PGM_OK;
    ReadsPart parts[2];
    reads_parts(ctx, ctx->h_lq, ctx->h_n, parts);
    int rc = fence_copy_stream(ctx);
    if (rc) return rc;
    // chunk by chunk: the copy of chunk c is queued (pageable memory: staged by host threads first), then its unpack and table
    // inserts behind the copy's event — they run while the host stages / the copy engine moves chunk c + 1
    size_t ev = 0;
    const bool pageable[2] = {parts[0].host && is_pageable_host(parts[0].src), parts[1].host && is_pageable_host(parts[1].src)};
    bool text_queued = false;
    for (int k = 0; k < 2; k++) {
        const ReadsPart &pt = parts[k];
        for (uint32_t first = 0; first < pt.cnt; first += READS_CHUNK, ev++) {
            const uint32_t cnt = std::min<uint32_t>(READS_CHUNK, pt.cnt - first);
            const uint8_t *src = pt.src;
            if (pt.host) {
                if ((rc = h2d_chunk(ctx, ctx->packed_stage.as<uint8_t>() + pt.stage_off + (size_t)first * pt.plen, pt.src + (size_t)first * pt.plen,
                                    (size_t)cnt * pt.plen, chunk_event(ctx->reads_ev, ev), pageable[k]))) return rc;
                CU(cudaStreamWaitEvent(ctx->stream, ctx->reads_ev[ev], 0));
                src = ctx->packed_stage.as<uint8_t>() + pt.stage_off;
            }
            // (pinned inputs: the copies are asynchronous, so the text's copies are queued behind the LAST read chunk's copy
            // before that chunk's kernels are launched — the copy engine never waits for the host)
            const bool last = (k == 1 || parts[1].cnt == 0) && first + READS_CHUNK >= pt.cnt;
            if (last && !text_queued) { if ((rc = enqueue_text_copies(ctx, false))) return rc; text_queued = true; }
            if ((rc = unpack_range(ctx, pt, src, first, cnt))) return rc;
            if (build && (rc = build_range(ctx, pt.first_read + first, pt.first_read + first + cnt, 0))) return rc;
        }
    }
    if (!text_queued && (rc = enqueue_text_copies(ctx, false))) return rc;
    ctx->reads_pending = false;
    ctx->state_fresh = true;
    return PGM_OK;
}

// L2 persisting window over the pre-filter for the kernels launched next on the context's stream
int filter_window(pgm_ctx *ctx, bool on, bool force = false) {
    if ((ctx->l2_hints != 2 && !force) || !ctx->filter_words || !ctx->persist_max || !ctx->window_max) return PGM_OK;
    const size_t bytes = std::min((size_t)ctx->filter_words * 4, ctx->window_max);
    if (on && ctx->persist_set < std::min(bytes, ctx->persist_max)) {
        CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min(bytes, ctx->persist_max)));
        ctx->persist_set = std::min(bytes, ctx->persist_max);
    }
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    attr.accessPolicyWindow.base_ptr = ctx->filter.p;
    attr.accessPolicyWindow.num_bytes = on ? bytes : 0;
    attr.accessPolicyWindow.hitRatio = on ? (float)std::min(1.0, (double)ctx->persist_set / (double)bytes) : 0.f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CU(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    return PGM_OK;
}

template <int NCH>
void launch_scan(const pgm::ScanParams &sp, unsigned int grid, cudaStream_t s, int mode) {
    // FAST: only ACGT reads, records of exactly 64 bytes (read length <= 192)
    const bool fast = sp.reads.n_n == 0 && sp.reads.lq_stride16 == 4;
    const bool pair = sp.tab.pair != 0;          // (never set together with ilv)
    if (mode == 1) {                             // filter stage of the L2-blocked pipeline
        if (pair) pgm::scan_kernel<NCH, false, 1, false, true><<<grid, PGM_SCAN_THREADS, 0, s>>>(sp);
        else pgm::scan_kernel<NCH, false, 1, false, false><<<grid, PGM_SCAN_THREADS, 0, s>>>(sp);
    } else if (mode == 2) {                      // behind the partitioned pre-filter: stage A1 reads the hit bitmap
        if (fast) pgm::scan_kernel<NCH, true, 2, false, false><<<grid, PGM_SCAN_THREADS, 0, s>>>(sp);
        else pgm::scan_kernel<NCH, false, 2, false, false><<<grid, PGM_SCAN_THREADS, 0, s>>>(sp);
    } else if (sp.ilv) {
        if (fast) pgm::scan_kernel<NCH, true, 0, true, false><<<grid, PGM_SCAN_THREADS, 0, s>>>(sp);
        else pgm::scan_kernel<NCH, false, 0, true, false><<<grid, PGM_SCAN_THREADS, 0, s>>>(sp);
    } else if (pair) {
        if (fast) pgm::scan_kernel<NCH, true, 0, false, true><<<grid, PGM_SCAN_THREADS, 0, s>>>(sp);
        else pgm::scan_kernel<NCH, false, 0, false, true><<<grid, PGM_SCAN_THREADS, 0, s>>>(sp);
    } else if (fast) pgm::scan_kernel<NCH, true, 0, false, false><<<grid, PGM_SCAN_THREADS, 0, s>>>(sp);
    else pgm::scan_kernel<NCH, false, 0, false, false><<<grid, PGM_SCAN_THREADS, 0, s>>>(sp);
}

void launch_scan_nch(int nch, const pgm::ScanParams &sp, unsigned int grid, cudaStream_t s, int mode) {
    switch (nch) {
        case 1: launch_scan<1>(sp, grid, s, mode); break;
        case 2: launch_scan<2>(sp, grid, s, mode); break;
        case 3: launch_scan<3>(sp, grid, s, mode); break;
        case 4: launch_scan<4>(sp, grid, s, mode); break;
        case 5: launch_scan<5>(sp, grid, s, mode); break;
        case 6: launch_scan<6>(sp, grid, s, mode); break;
        case 7: launch_scan<7>(sp, grid, s, mode); break;
        default: launch_scan<8>(sp, grid, s, mode); break;
    }
}

template <int NCH>
cudaError_t launch_part_scan(const pgm::PartScanParams &ps, unsigned int grid, cudaStream_t s) {
    const size_t smem = sizeof(pgm::PartScanShared);
    cudaError_t e = cudaFuncSetAttribute(pgm::part_scan_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pgm::part_scan_kernel<NCH><<<grid, PGM_PART_THREADS, smem, s>>>(ps);
    return cudaSuccess;
}

cudaError_t launch_part_scan_nch(int nch, const pgm::PartScanParams &ps, unsigned int grid, cudaStream_t s) {
    switch (nch) {
        case 1: return launch_part_scan<1>(ps, grid, s);
        case 2: return launch_part_scan<2>(ps, grid, s);
        case 3: return launch_part_scan<3>(ps, grid, s);
        case 4: return launch_part_scan<4>(ps, grid, s);
        case 5: return launch_part_scan<5>(ps, grid, s);
        case 6: return launch_part_scan<6>(ps, grid, s);
        case 7: return launch_part_scan<7>(ps, grid, s);
        default: return launch_part_scan<8>(ps, grid, s);
    }
}

// The partitioned pre-filter (pgm_part.cuh) for one scan launch: window starts per round, partitions, queue capacity.
// Auto mode: the pattern set is beyond what the Bloom filter tells apart (match_begin gave it more than one slice) and the
// launch is long enough to pay for streaming the bucket array once.
constexpr uint64_t PART_ROUND_MAX = (1ull << 31) - (1ull << 20);     // launch-relative positions are 32 bits; queues of a round <= 26 GB
bool part_wanted(const pgm_ctx *ctx) {
    if (ctx->part_scan <= 0 || ctx->n_reads() == 0 || ctx->interleaved) return false;
    return ctx->part_scan >= 2 || ctx->filter_slice_bits > 0;
}
bool plan_part(const pgm_ctx *ctx, uint64_t n_pos, pgm::PartQueues &q) {
    memset(&q, 0, sizeof q);
    if (!part_wanted(ctx) || n_pos > PART_ROUND_MAX) return false;
    const uint64_t table_bytes = (uint64_t)ctx->n_buckets * 32;
    if (ctx->part_scan == 1 && n_pos * 12 < table_bytes / 8) return false;
    int pb = ceil_log2((table_bytes + ((uint64_t)ctx->part_mb << 20) - 1) / ((uint64_t)ctx->part_mb << 20));
    pb = std::min(10, std::max(ctx->part_scan >= 2 ? 3 : 1, pb));
    q.part_bits = (uint32_t)pb;
    const uint64_t cap = ctx->part_scan == 3 ? 64 : (n_pos >> pb) + (n_pos >> (pb + 6)) + 8192;       // mean + 1.5 % + slack
    q.cap = (uint32_t)std::min<uint64_t>(cap, 0xFFFFFFF0ull);
    return true;
}

int floor_log2(uint64_t v) { int b = 0; while ((2ull << b) <= v) b++; return b; }

// Geometry of the L2-blocked pipeline for one scan launch over `n_pos` window starts; false = use the fused kernel.
bool plan_blocked(pgm_ctx *ctx, uint64_t n_pos, uint32_t n_tiles, pgm::StageQueues &q) {
    memset(&q, 0, sizeof q);
    const int mode = ctx->blocked_scan;
    if (mode <= 0 || ctx->n_reads() == 0 || ctx->interleaved) return false;
    const uint64_t table_bytes = (uint64_t)ctx->n_buckets * 32;
    if ((uint64_t)n_tiles * PGM_TILE_POS >= (1ull << 31)) return false;              // queue positions are 31 bits
    if (mode == 1) {
        // pays when (i) the table is well beyond the L2, (ii) a region sees many more probes than it has lines
        // (a long enough launch), (iii) the planes of the slice stay L2-resident next to a read range
        if (table_bytes < (192ull << 20) || n_pos < (32ull << 20) || ctx->slice_len > (288ull << 20)) return false;
    }
    const bool forced = mode >= 2;
    int rb = ceil_log2((table_bytes + ((uint64_t)ctx->region_mb << 20) - 1) / ((uint64_t)ctx->region_mb << 20));
    rb = std::min(8, std::max(forced ? 3 : 1, rb));
    const uint64_t rec_bytes = (uint64_t)ctx->lq_stride16 * 16;
    int rs = floor_log2(std::max<uint64_t>(1, ((uint64_t)ctx->range_mb << 20) / rec_bytes));
    const uint32_t n = ctx->n_reads();
    if (forced) rs = std::max(0, ceil_log2(n) - 3);
    while (((uint64_t)(n - 1) >> rs) >= PGM_SQ_MAX) rs++;
    q.region_bits = (uint32_t)rb;
    q.range_shift = (uint32_t)rs;
    q.n_ranges = (uint32_t)(((uint64_t)(n - 1) >> rs) + 1);
    // capacity: 60 % of the windows positive / verified, evenly spread, + slack; beyond that the pass falls back
    const uint64_t pos_cap = mode == 3 ? 64 : (n_pos * 6 / 10 >> rb) + 4096;
    const uint64_t cand_cap = mode == 3 ? 64 : n_pos * 6 / 10 / q.n_ranges + 4096;
    if (pos_cap >= 0xFFFFFFF0ull || cand_cap >= 0xFFFFFFF0ull) return false;
    q.pos_cap = (uint32_t)pos_cap;
    q.cand_cap = (uint32_t)cand_cap;
    return true;
}

} // namespace

extern "C" {

int pgm_abi_version(void) { return PGM_ABI_VERSION; }

const char *pgm_last_error(const pgm_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int pgm_create(int device, pgm_ctx **out) {
    pgm_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, PGM_ERR_INVALID_ARG, "pgm_create: out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, PGM_ERR_NO_DEVICE,
                    std::string("pgm_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail(nullptr, PGM_ERR_INVALID_ARG, "pgm_create: device index out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(nullptr, PGM_ERR_NO_DEVICE,
                    std::string("pgm_create: device '") + prop.name + "' is not compute capability 10.x; kernels are built for sm_100a only");
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    ctx = new pgm_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (const char *t = getenv("PGM_TWO_STEP_BUILD")) ctx->two_step_build = atoi(t);   // 0 off, 1 auto, 2 always (tests)
    if (const char *t = getenv("PGM_FILTER_K")) ctx->filter_k_force = atoi(t);
    if (const char *t = getenv("PGM_FILTER_PAIR")) ctx->filter_pair_mode = atoi(t);
    if (const char *t = getenv("PGM_FILTER_SLICES")) ctx->filter_slices_force = std::min(3, std::max(0, atoi(t)));
    if (const char *t = getenv("PGM_INSERT_PREFETCH")) ctx->insert_prefetch = atoi(t);
    if (const char *t = getenv("PGM_BLOCKED_SCAN")) ctx->blocked_scan = atoi(t);
    if (const char *t = getenv("PGM_PART_SCAN")) ctx->part_scan = atoi(t);
    if (const char *t = getenv("PGM_PART_MB")) ctx->part_mb = std::max(1, atoi(t));
    if (const char *t = getenv("PGM_CM_WARP")) ctx->cm_warp = atoi(t);
    if (const char *t = getenv("PGM_REGION_MB")) ctx->region_mb = std::max(1, atoi(t));
    if (const char *t = getenv("PGM_RANGE_MB")) ctx->range_mb = std::max(1, atoi(t));
    if (const char *g = getenv("PGM_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));   // experiment knob
    ctx->persist_max = (size_t)std::max(0, prop.persistingL2CacheMaxSize);
    ctx->window_max = (size_t)std::max(0, prop.accessPolicyMaxWindowSize);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return cuda_fail(nullptr, e, "cudaStreamCreate");
    }
    ctx->own_stream = true;
    if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->fence_ev, cudaEventDisableTiming)) != cudaSuccess) {
        cuda_fail(nullptr, e, "cudaStreamCreate (copy stream)");
        pgm_destroy(ctx);
        return PGM_ERR_CUDA;
    }
    int rc;
    if ((rc = ensure(ctx, ctx->counters, 16 * sizeof(unsigned long long))) != PGM_OK ||
        (rc = ensure(ctx, ctx->hist, 257 * sizeof(unsigned long long))) != PGM_OK ||
        (rc = ensure(ctx, ctx->err_flag, sizeof(int))) != PGM_OK ||
        (rc = ensure(ctx, ctx->touched, sizeof(int))) != PGM_OK) {
        g_create_error = ctx->err;
        pgm_destroy(ctx);
        return rc;
    }
    cudaMemsetAsync(ctx->counters.p, 0, 16 * sizeof(unsigned long long), ctx->stream);
    cudaMemsetAsync(ctx->err_flag.p, 0, sizeof(int), ctx->stream);
    cudaMemsetAsync(ctx->touched.p, 0, sizeof(int), ctx->stream);
    *out = ctx;
    return PGM_OK;
}

void pgm_destroy(pgm_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->f_lo, &ctx->f_hi, &ctx->r_lo, &ctx->r_hi, &ctx->ascii_stage, &ctx->packed_stage,
                      &ctx->lq_recs, &ctx->n_recs, &ctx->keys, &ctx->first_order,
                      &ctx->same_mask, &ctx->same_mm, &ctx->touched, &ctx->buckets, &ctx->next, &ctx->filter, &ctx->bq_entries, &ctx->bq_counters,
                      &ctx->sq_pos, &ctx->sq_cand, &ctx->sq_counters, &ctx->pq_entries, &ctx->pq_counters, &ctx->pq_bits, &ctx->mis_sums, &ctx->mis_offsets, &ctx->mis_pos, &ctx->mis_syms,
                      &ctx->rt_win_send2[0], &ctx->rt_win_send2[1], &ctx->rt_win_recv2[0], &ctx->rt_win_recv2[1], &ctx->rt_cand_send2[0],
                      &ctx->rt_cand_send2[1], &ctx->rt_cand_recv2[0], &ctx->rt_cand_recv2[1], &ctx->rt_pat_recv, &ctx->rt_counters, &ctx->rt_live, &ctx->rt_live_count,
                      &ctx->mem_dlo, &ctx->mem_dhi, &ctx->mem_dinv, &ctx->mem_stage, &ctx->mem_fv, &ctx->mem_has, &ctx->mem_emit, &ctx->mem_gcount,
                      &ctx->mem_gstart, &ctx->mem_raw, &ctx->mem_rawq, &ctx->mem_keep, &ctx->mem_kstart, &ctx->mem_out,
                      &ctx->cm_count, &ctx->cm_start, &ctx->cm_cumm, &ctx->cm_fill, &ctx->cm_hash, &ctx->cm_all, &ctx->cm_entries, &ctx->cm_sums, &ctx->cm_nib, &ctx->cm_coarse, &ctx->cmw_lens, &ctx->cmw_cand, &ctx->cmw_vt, &ctx->cmw_nvt,
                      &ctx->counters, &ctx->hist, &ctx->err_flag, &ctx->out_pos, &ctx->out_rc, &ctx->out_mm};
    for (DevBuf *b : bufs) release(*b);
    for (const pgm_ctx::EvPair &e : ctx->ev_used) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (cudaEvent_t e : ctx->ev_free) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->text_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->reads_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->mem_ev) cudaEventDestroy(e);
    for (auto &kv : ctx->ipc_open) cudaIpcCloseMemHandle(kv.second);
    if (ctx->rt_counts_host) cudaFreeHost(ctx->rt_counts_host);
    for (int k = 0; k < 4; k++) {
        if (ctx->hstage.buf[k]) cudaFreeHost(ctx->hstage.buf[k]);
        if (ctx->hstage.ev[k]) cudaEventDestroy(ctx->hstage.ev[k]);
    }
    for (int k = 0; k < 3; k++) for (int sl = 0; sl < 2; sl++) if (ctx->counts_ev[k][sl]) cudaEventDestroy(ctx->counts_ev[k][sl]);
    for (int k = 0; k < 3; k++) for (int sl = 0; sl < 2; sl++) if (ctx->consumed_ev[k][sl]) cudaEventDestroy(ctx->consumed_ev[k][sl]);
    for (int k = 0; k < 3; k++) for (int sl = 0; sl < 2; sl++) if (ctx->pull_ev[k][sl]) cudaEventDestroy(ctx->pull_ev[k][sl]);
    for (int k = 0; k < 3; k++) {
        if (ctx->pull_more[k]) { cudaStreamSynchronize(ctx->pull_more[k]); cudaStreamDestroy(ctx->pull_more[k]); }
        if (ctx->pull_join[k]) cudaEventDestroy(ctx->pull_join[k]);
    }
    if (ctx->pull_stream) { cudaStreamSynchronize(ctx->pull_stream); cudaStreamDestroy(ctx->pull_stream); }
    if (ctx->fence_ev) cudaEventDestroy(ctx->fence_ev);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int pgm_set_stream(pgm_ctx *ctx, void *cuda_stream) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    if (cuda_stream) ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    else { CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    return PGM_OK;
}

int pgm_synchronize(pgm_ctx *ctx) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return PGM_OK;
}

int pgm_upload(pgm_ctx *ctx);   // (defined behind upload_reads / finish_text_upload's users below)

int pgm_set_tuning(pgm_ctx *ctx, int filter_log2_bits, int slots_per_pattern, int ctas_per_sm, int l2_hints) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (filter_log2_bits > 32 || (filter_log2_bits > 0 && filter_log2_bits < 10))
        return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_set_tuning: filter_log2_bits must be 0 (off), <0 (auto) or 10..32");
    if (slots_per_pattern < 2 || slots_per_pattern > 64) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_set_tuning: slots_per_pattern must be 2..64");
    if (ctas_per_sm < 1 || ctas_per_sm > 8) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_set_tuning: ctas_per_sm must be 1..8");
    ctx->filter_log2_bits = filter_log2_bits;
    ctx->slots_per_pattern = slots_per_pattern;
    ctx->ctas_per_sm = ctas_per_sm;
    ctx->l2_hints = l2_hints < 0 ? 0 : (l2_hints > 2 ? 2 : l2_hints);
    return PGM_OK;
}

uint64_t pgm_kernel_launches(const pgm_ctx *ctx) { return ctx ? ctx->launches : 0; }

int pgm_set_profiling(pgm_ctx *ctx, int on) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    ctx->profiling = on != 0;
    return PGM_OK;
}

int pgm_get_timings(pgm_ctx *ctx, pgm_timings *out) {
    if (!ctx || !out) return PGM_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    memset(out, 0, sizeof *out);
    for (const pgm_ctx::EvPair &e : ctx->ev_used) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e.a, e.b));
        out->ms[e.kind] += ms;
        out->launches[e.kind]++;
        ctx->ev_free.push_back(e.a);
        ctx->ev_free.push_back(e.b);
    }
    ctx->ev_used.clear();
    return PGM_OK;
}

int pgm_shard_plan(uint64_t pg_len, int rank, int world, uint64_t *slice_begin, uint64_t *slice_len,
                   uint64_t *own_begin, uint64_t *own_end) {
    if (world < 1 || rank < 0 || rank >= world || !slice_begin || !slice_len || !own_begin || !own_end) return PGM_ERR_INVALID_ARG;
    auto cut = [&](int k) -> uint64_t {
        if (k >= world) return pg_len;
        unsigned __int128 v = (unsigned __int128)pg_len * (unsigned)k / (unsigned)world;
        return ((uint64_t)v / 128) * 128; // keep tile copies 16-byte aligned relative to the slice
    };
    const uint64_t ob = cut(rank), oe = cut(rank + 1);
    uint64_t sb = ob > PGM_SHARD_HALO ? ob - PGM_SHARD_HALO : 0;
    sb = (sb / 128) * 128;
    const uint64_t se = std::min(pg_len, oe + PGM_SHARD_HALO);
    *own_begin = ob; *own_end = oe; *slice_begin = sb; *slice_len = se - sb;
    return PGM_OK;
}

int pgm_set_text_shard(pgm_ctx *ctx, const char *slice, uint64_t slice_begin, uint64_t slice_len,
                       uint64_t pg_len, uint64_t own_begin, uint64_t own_end) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (!slice && slice_len) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_set_text: text is null");
    if (pg_len >= PGM_POS_MASK) return fail(ctx, PGM_ERR_UNSUPPORTED, "pgm_set_text: pseudogenome longer than 2^40-2 bases");
    if (slice_begin + slice_len > pg_len || own_begin > own_end || own_end > pg_len ||
        own_begin < slice_begin || own_end > slice_begin + slice_len)
        return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_set_text_shard: inconsistent slice / owned range");
    if (slice_begin % 32 != 0) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_set_text_shard: slice_begin must be a multiple of 32");
    if ((own_begin != 0 && own_begin - slice_begin < PGM_SHARD_HALO && slice_begin != 0) ||
        (own_end != pg_len && slice_begin + slice_len - own_end < PGM_SHARD_HALO && slice_begin + slice_len != pg_len))
        return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_set_text_shard: slice must extend PGM_SHARD_HALO bases beyond the owned range");
    CU(cudaSetDevice(ctx->device));
    int rc;
    const size_t pb = plane_bytes(slice_len);
    if ((rc = ensure(ctx, ctx->f_lo, pb)) || (rc = ensure(ctx, ctx->f_hi, pb)) ||
        (rc = ensure(ctx, ctx->r_lo, pb)) || (rc = ensure(ctx, ctx->r_hi, pb))) return rc;
    CU(cudaMemsetAsync(ctx->f_lo.p, 0, pb, ctx->stream));
    CU(cudaMemsetAsync(ctx->f_hi.p, 0, pb, ctx->stream));
    CU(cudaMemsetAsync(ctx->r_lo.p, 0, pb, ctx->stream));
    CU(cudaMemsetAsync(ctx->r_hi.p, 0, pb, ctx->stream));
    ctx->pg_len = pg_len; ctx->slice_begin = slice_begin; ctx->slice_len = slice_len;
    ctx->own_begin = own_begin; ctx->own_end = own_end;
    ctx->has_text = true;
    ctx->mem_index_valid = false; ctx->mem_result_valid = false;
    ctx->text_pending = false; ctx->text_copies_enqueued = false; ctx->h_text = nullptr;
    CU(cudaMemsetAsync(ctx->err_flag.p, 0, sizeof(int), ctx->stream));   // a new text: the symbol check starts over
    if (!slice_len) return PGM_OK;
    if (is_device_ptr(slice)) {
        // device-resident text: one launch over the whole slice (the chunks exist for the pipelined host upload)
        uint32_t *flo = ctx->f_lo.as<uint32_t>() + PGM_PAD_WORDS, *fhi = ctx->f_hi.as<uint32_t>() + PGM_PAD_WORDS;
        KLAUNCH(PGM_K_PACK_TEXT, "pack_text_kernel", pgm::pack_text_kernel<<<grid_for((slice_len + 31) / 32, 256), 256, 0, ctx->stream>>>(
            reinterpret_cast<const uint8_t *>(slice), slice_len, flo, fhi, 0, ctx->err_flag.as<int>()));
        return rc_text(ctx);
    }
    // host text: uploaded lazily (see pgm_match_begin / pgm_scan_pass); the caller keeps the buffer alive and
    // unchanged until the call that returns the results has completed
    if ((rc = ensure(ctx, ctx->ascii_stage, (size_t)slice_len))) return rc;
    ctx->h_text = reinterpret_cast<const uint8_t *>(slice);
    ctx->text_pending = true;
    return PGM_OK;
}

int pgm_set_text(pgm_ctx *ctx, const char *text, uint64_t pg_len) {
    return pgm_set_text_shard(ctx, text, 0, pg_len, pg_len, 0, pg_len);
}

int pgm_set_reads(pgm_ctx *ctx, const uint8_t *lq_packed, uint32_t n_lq, const uint8_t *n_packed, uint32_t n_n,
                  uint32_t read_len) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (read_len == 0 || read_len > 255) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_set_reads: read_len must be 1..255 (PgRC limit)");
    if ((n_lq && !lq_packed) || (n_n && !n_packed)) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_set_reads: null reads pointer");
    if ((uint64_t)n_lq + n_n >= 0xFFFFFFFFull) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_set_reads: too many reads");
    CU(cudaSetDevice(ctx->device));
    const uint32_t W = (read_len + 31) / 32;
    // record = header uint4 + planes, rounded up to whole 64-byte requests
    const uint32_t lq_stride16 = (1 + (W + 1) / 2 + 3) & ~3u, n_stride16 = (1 + W + 3) & ~3u;
    const uint32_t n_plen = (read_len + 2) / 3;
    int rc;
    if ((rc = ensure(ctx, ctx->lq_recs, (size_t)std::max<uint32_t>(n_lq, 1) * lq_stride16 * 16))) return rc;
    if ((rc = ensure(ctx, ctx->n_recs, (size_t)std::max<uint32_t>(n_n, 1) * n_stride16 * 16))) return rc;
    const uint32_t n = n_lq + n_n;
    const size_t n1 = std::max<uint32_t>(n, 1);
    if (ctx->first_order.cap < n1 * 8 || ctx->same_mask.cap < n1 * 4 || ctx->same_mm.cap < n1) ctx->aux_clean = false;
    if ((rc = ensure(ctx, ctx->first_order, n1 * 8)) || (rc = ensure(ctx, ctx->same_mask, n1 * 4)) ||
        (rc = ensure(ctx, ctx->same_mm, n1)) || (rc = ensure(ctx, ctx->out_pos, n1 * 8)) ||
        (rc = ensure(ctx, ctx->out_rc, n1)) || (rc = ensure(ctx, ctx->out_mm, n1))) return rc;
    ctx->n_lq = n_lq; ctx->n_n = n_n; ctx->read_len = read_len; ctx->W = W;
    ctx->lq_stride16 = lq_stride16; ctx->n_stride16 = n_stride16;
    ctx->has_reads = true;
    ctx->state_fresh = false;
    ctx->outputs_valid = false;
    ctx->phase_active = false;
    ctx->h_lq = lq_packed; ctx->h_n = n_packed;
    ReadsPart parts[2];
    reads_parts(ctx, lq_packed, n_packed, parts);
    if (parts[0].host || parts[1].host) {
        // host reads: uploaded lazily, chunk by chunk, overlapped with unpacking and the table build
        // (pgm_match_begin); both sets share one staging buffer
        if ((rc = ensure(ctx, ctx->packed_stage, parts[1].stage_off + (size_t)n_n * n_plen + 16))) return rc;
        ctx->reads_pending = true;
        return PGM_OK;
    }
    ctx->reads_pending = true;      // device-resident reads: unpack now
    return upload_reads(ctx, false);
}

} // extern "C"

namespace {
int match_begin_impl(pgm_ctx *ctx, uint32_t seed_len, uint32_t parts, uint32_t max_mm, uint32_t min_mm, int continuation, bool interleaved);
}

extern "C" {

int pgm_match_begin(pgm_ctx *ctx, uint32_t seed_len, uint32_t parts, uint32_t max_mm, uint32_t min_mm, int continuation) {
    return match_begin_impl(ctx, seed_len, parts, max_mm, min_mm, continuation, false);
}

int pgm_match_begin_interleaved(pgm_ctx *ctx, uint32_t seed_len, uint32_t parts, uint32_t max_mm, uint32_t min_mm, int continuation) {
    return match_begin_impl(ctx, seed_len, parts, max_mm, min_mm, continuation, true);
}

} // extern "C"

namespace {
int match_begin_impl(pgm_ctx *ctx, uint32_t seed_len, uint32_t parts, uint32_t max_mm, uint32_t min_mm, int continuation, bool interleaved) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (interleaved && parts > PGM_ILV_MAX_PARTS)
        return fail(ctx, PGM_ERR_UNSUPPORTED, "pgm_match_begin_interleaved: more than 31 seeds per read");
    if (parts == 1) interleaved = false;            // one seed: stride 1, the contiguous case
    if (!ctx->has_reads) return fail(ctx, PGM_ERR_STATE, "pgm_match_begin: pgm_set_reads has not been called");
    if (seed_len == 0 || parts == 0 || (uint64_t)seed_len * parts > ctx->read_len)
        return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_match_begin: need seed_len >= 1 and seed_len * parts <= read_len");
    if (max_mm > 127 || min_mm > 127) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_match_begin: mismatch limits must be <= 127");
    const uint64_t n_patterns = (uint64_t)ctx->n_reads() * parts;
    // pattern id = read << part_bits | seed; the reference's limit is n_reads * parts < 2^32 (HashMatcher.cpp:39)
    const uint32_t part_bits = (uint32_t)ceil_log2(parts);
    const uint64_t n_ids = (uint64_t)ctx->n_reads() << part_bits;
    if (n_ids >= 0xFFFFFFFFull) return fail(ctx, PGM_ERR_UNSUPPORTED, "pgm_match_begin: pattern index exceeds 32 bits (reads << ceil(log2(parts)))");
    CU(cudaSetDevice(ctx->device));
    const uint32_t n = ctx->n_reads();
    // table geometry: a prime number of 4-slot buckets, slots_per_pattern slots per pattern
    const uint64_t want_slots = std::max<uint64_t>(512, n_patterns * (uint64_t)ctx->slots_per_pattern);
    const uint64_t nb64 = next_prime((want_slots + 3) / 4);
    if (nb64 >= 0x7FFFFFFFull) return fail(ctx, PGM_ERR_UNSUPPORTED, "pgm_match_begin: table too large");
    int rc;
    if ((rc = ensure(ctx, ctx->buckets, nb64 * 32)) || (rc = ensure(ctx, ctx->next, std::max<uint64_t>(n_ids, 1) * 4))) return rc;
    ctx->n_buckets = (uint32_t)nb64;
    ctx->n_slots = nb64 * 4;
    ctx->part_bits = part_bits;
    CU(cudaMemsetAsync(ctx->buckets.p, 0xFF, nb64 * 32, ctx->stream));
    int fbits = ctx->filter_log2_bits;
    // auto: 8 bits per pattern up to 2^28 bits (32 MB stays L2-resident next to the streaming traffic); pattern sets far
    // beyond that get 2^29 bits (64 MB: still mostly resident) — a saturated filter sends every text window to the table
    if (fbits < 0) fbits = std::min(n_patterns > (100ull << 20) ? 29 : 28, std::max(15, ceil_log2(std::max<uint64_t>(n_patterns, 1) * 8)));
    const bool auto_bits = ctx->filter_log2_bits < 0;
    // hash-sliced filter: beyond ~130 M patterns a 2^29-bit filter has fewer than 4 bits per pattern and lets most windows
    // through; a second 64 MB slice (one more scan launch) pays — measured at C5 (990 M patterns, profiles/bench_c5_slices*_r02e.json):
    // scan 389 -> 334 ms per step with 2 slices (63 % of the windows still pass), 340 with 4, 521 with 8: re-hashing the text
    // costs 22 ms per launch, about what the next doubling of the filter saves in table probes
    ctx->filter_slice_bits = 0;
    if (auto_bits && !interleaved && ctx->blocked_scan <= 0) {
        if (ctx->filter_slices_force >= 0) {
            // (tests / sweeps: the filter as sized above, cut into 2^force slices; at full size it grows like the auto rule)
            ctx->filter_slice_bits = (uint32_t)std::min(ctx->filter_slices_force, std::max(0, fbits - 8));
            if (fbits == 29) fbits += (int)ctx->filter_slice_bits;
        } else if (fbits == 29) {
            ctx->filter_slice_bits = (uint32_t)std::min(1, std::max(0, ceil_log2(n_patterns * 4) - 29));
            fbits += (int)ctx->filter_slice_bits;
        }
    }
    if (fbits > 0) {
        // paired lookups (pgm_kernels.cuh, pair_word): one gather for two adjacent windows, every pattern entered in two
        // words; only for contiguous seeds with >= 12 exactly-hashed shared bases, and only while the doubled load stays low
        const uint32_t lo_off = seed_len > 32 ? seed_len - 32 : 0, hi_off = std::min<uint32_t>(seed_len, 32);
        const int core = (int)hi_off - (int)lo_off - 1;
        if (auto_bits && ctx->filter_pair_mode == 1 && !interleaved && core >= 12 && n_patterns * 24 <= (1ull << 28))
            fbits = std::max(fbits, std::min(28, ceil_log2(std::max<uint64_t>(n_patterns, 1) * 24)));   // room for the second entry
        ctx->filter_pair = !interleaved && core >= 12 && !ctx->filter_slice_bits &&
                           (ctx->filter_pair_mode == 2 || (ctx->filter_pair_mode == 1 && n_patterns * 12 <= (1ull << fbits)));
        ctx->pair_lo = lo_off;
        ctx->pair_mask = core >= 12 ? (uint32_t)((1ull << core) - 1) : 0;
        // bits per pattern: k = 0.69 * filter bits / entries minimises the false-positive rate; 1 or 2, in one word
        const double per = (double)(1ull << fbits) / (double)(std::max<uint64_t>(n_patterns, 1) * (ctx->filter_pair ? 2 : 1));
        ctx->filter_k = ctx->filter_k_force ? (uint32_t)std::min(2, std::max(1, ctx->filter_k_force))
                                            : (uint32_t)std::min(2.0, std::max(1.0, 0.69 * per + 0.5));   // (3 or 4 bits: more ALU per window than probes saved, measured)
        ctx->filter_words = 1u << (fbits - 5);
        const size_t fbytes = (size_t)1 << (fbits - 3);
        if ((rc = ensure(ctx, ctx->filter, fbytes))) return rc;
        CU(cudaMemsetAsync(ctx->filter.p, 0, fbytes, ctx->stream));
    } else {
        ctx->filter_words = 0;
    }
    if (!continuation) CU(cudaMemsetAsync(ctx->counters.p, 0, 16 * sizeof(unsigned long long), ctx->stream));
    else CU(cudaMemsetAsync(ctx->counters.as<unsigned long long>() + 4, 0, sizeof(unsigned long long), ctx->stream));
    // region queues of the two-step build: regions of about 16 MB of buckets, only for tables well beyond the L2
    {
        const uint64_t table_bytes = nb64 * 32;
        uint32_t rb = 0;
        // (host reads arriving chunk by chunk: the one-step inserts hide behind the H2D copies, a queue flush would not)
        if ((ctx->two_step_build == 1 && table_bytes >= (256ull << 20) && !ctx->reads_pending) || ctx->two_step_build == 2) {
            rb = (uint32_t)ceil_log2((table_bytes + (16ull << 20) - 1) / (16ull << 20));
            rb = std::min<uint32_t>(std::max<uint32_t>(rb, ctx->two_step_build == 2 ? 3u : 1u), 6);   // PGM_MAX_REGIONS = 64
        }
        ctx->bq_region_bits = rb;
        ctx->bq_pending = false;
        if ((rc = ensure(ctx, ctx->bq_counters, 2 * PGM_MAX_REGIONS * sizeof(unsigned int)))) return rc;
        CU(cudaMemsetAsync(ctx->bq_counters.p, 0, 2 * PGM_MAX_REGIONS * sizeof(unsigned int), ctx->stream));
        if (rb) {
            const uint64_t per_region = (n_patterns >> rb) + (n_patterns >> (rb + 4)) + 4096;   // mean + 6 % + slack
            ctx->bq_cap = (uint32_t)std::min<uint64_t>(per_region, 0xFFFFFFF0ull);
            if ((rc = ensure(ctx, ctx->bq_entries, ((size_t)ctx->bq_cap << rb) * sizeof(uint4)))) return rc;
        } else {
            ctx->bq_cap = 0;
        }
    }
    ctx->seed_len = seed_len; ctx->parts = parts; ctx->max_mm = max_mm; ctx->min_mm = min_mm;
    ctx->interleaved = interleaved;
    ctx->outputs_valid = false;
    if (ctx->reads_pending && continuation && (rc = upload_reads(ctx, false))) return rc;   // (not a sensible call order)
    const bool pipelined = ctx->reads_pending;
    if (n && ((!continuation && !ctx->state_fresh && !pipelined) || !ctx->aux_clean)) {
        // record headers are rewritten by the unpack kernels when the reads are still to be uploaded
        KLAUNCH(PGM_K_INIT_STATE, "reset_state_kernel", pgm::reset_state_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(
            reads_view(ctx), per_read(ctx), n, continuation ? 0 : 1, pipelined ? 0 : 1));
        CU(cudaMemsetAsync(ctx->touched.p, 0, sizeof(int), ctx->stream));
        ctx->aux_clean = true;
    }
    if (pipelined) {
        if ((rc = upload_reads(ctx, n_patterns != 0))) return rc;
    } else if (n_patterns) {
        if ((rc = build_range(ctx, 0, n, continuation ? 1 : 0))) return rc;
        if ((rc = enqueue_text_copies(ctx, true))) return rc;      // start a pending text upload behind the build
    }
    if ((rc = build_flush(ctx))) return rc;
    ctx->phase_active = true;
    return PGM_OK;
}

// One scan launch over the seed-window starts [fb, fe) (FORWARD global coordinates) of the pass.
int scan_range(pgm_ctx *ctx, int rev_mode, uint64_t fb, uint64_t fe) {
    if (fb >= fe) return PGM_OK;
    if (part_wanted(ctx) && fe - fb > PART_ROUND_MAX) {
        // the partitioned pre-filter works in rounds of at most 2^31 window starts (equal rounds, tile-aligned)
        const uint64_t k = (fe - fb + PART_ROUND_MAX - 1) / PART_ROUND_MAX;
        const uint64_t per = (((fe - fb + k - 1) / k) + PGM_TILE_POS - 1) / PGM_TILE_POS * PGM_TILE_POS;
        for (uint64_t b = fb; b < fe; b += per) {
            const int rc = scan_range(ctx, rev_mode, b, std::min(fe, b + per));
            if (rc) return rc;
        }
        return PGM_OK;
    }
    const uint64_t n = ctx->seed_span(), pg = ctx->pg_len;     // n: text bases a seed window covers
    pgm::ScanParams sp;
    memset(&sp, 0, sizeof sp);
    if (!rev_mode) {
        sp.tlo = ctx->f_lo.as<uint32_t>() + PGM_PAD_WORDS; sp.thi = ctx->f_hi.as<uint32_t>() + PGM_PAD_WORDS;
        sp.slice_origin = ctx->slice_begin;
        sp.own_begin = fb; sp.own_end = fe;
    } else {
        // the window starting at forward p is the window starting at q = pg - n - p of the RC text
        sp.tlo = ctx->r_lo.as<uint32_t>() + PGM_PAD_WORDS; sp.thi = ctx->r_hi.as<uint32_t>() + PGM_PAD_WORDS;
        sp.slice_origin = pg - (ctx->slice_begin + ctx->slice_len);
        sp.own_begin = pg - n - (fe - 1); sp.own_end = pg - n - fb + 1;
    }
    sp.pg_len = pg;
    const uint64_t lb = sp.own_begin - sp.slice_origin, le = sp.own_end - sp.slice_origin;
    sp.first_word = (uint32_t)((lb / 32) & ~3ull);
    sp.n_tiles = (uint32_t)((le - (uint64_t)sp.first_word * 32 + PGM_TILE_POS - 1) / PGM_TILE_POS);
    sp.seed_len = ctx->seed_len; sp.parts = ctx->parts; sp.max_mm = ctx->max_mm; sp.min_mm = ctx->min_mm;
    sp.shift_unit = ctx->shift_unit(); sp.ilv = ctx->ilv();
    sp.tail_mask = ctx->seed_len % 32 ? (1u << (ctx->seed_len % 32)) - 1u : 0xFFFFFFFFu;
    sp.rev_mode = rev_mode ? 1 : 0;
    sp.l2_hints = ctx->l2_hints == 1 ? 1 : 0;      // window mode: plain filter loads, the window carries the policy
    sp.stream_hints = ctx->l2_hints != 0;
    sp.tab = table_view(ctx);
    sp.reads = reads_view(ctx);
    sp.pr = per_read(ctx);
    sp.tile_counter = reinterpret_cast<unsigned int *>(ctx->counters.as<unsigned long long>() + 5);
    sp.counters = ctx->counters.as<unsigned long long>();
    CU(cudaMemsetAsync(sp.tile_counter, 0, sizeof(unsigned int), ctx->stream));
    const unsigned int grid = (unsigned int)std::min<uint64_t>(sp.n_tiles, (uint64_t)ctx->sm_count * ctx->ctas_per_sm);
    const int nch = (int)((ctx->seed_len + 31) / 32);
    ctx->state_fresh = false;
    ctx->aux_clean = false;
    ctx->outputs_valid = false;
    pgm::PartQueues pq;
    if (plan_part(ctx, le - lb, pq)) {
        // partitioned pre-filter: every window -> queue of its table partition -> probed partition by partition (L2) -> bitmap
        // of the windows with a table hit; then the fused kernel with the bitmap in place of its hash + filter stage
        int rc;
        const size_t bit_words = (size_t)sp.n_tiles * PGM_TILE_WORDS + 32;
        if ((rc = ensure(ctx, ctx->pq_entries, ((size_t)pq.cap << pq.part_bits) * 12)) ||
            (rc = ensure(ctx, ctx->pq_counters, 2 * PGM_PART_MAX * sizeof(unsigned int))) ||
            (rc = ensure(ctx, ctx->pq_bits, bit_words * 4))) return rc;
        pq.entries = ctx->pq_entries.as<uint32_t>();
        pq.count = ctx->pq_counters.as<unsigned int>(); pq.cursor = pq.count + PGM_PART_MAX;
        pq.hit_bits = ctx->pq_bits.as<uint32_t>();
        CU(cudaMemsetAsync(pq.count, 0, 2 * PGM_PART_MAX * sizeof(unsigned int), ctx->stream));
        CU(cudaMemsetAsync(pq.hit_bits, 0, bit_words * 4, ctx->stream));
        pgm::PartScanParams ps;
        memset(&ps, 0, sizeof ps);
        ps.tlo = sp.tlo; ps.thi = sp.thi; ps.slice_origin = sp.slice_origin; ps.own_begin = sp.own_begin; ps.own_end = sp.own_end;
        ps.first_word = sp.first_word; ps.n_tiles = sp.n_tiles; ps.tail_mask = sp.tail_mask; ps.tile_counter = sp.tile_counter;
        ps.q = pq;
        KLAUNCH(PGM_K_SCAN_FILTER, "part_scan_kernel", {
            cudaError_t le_ = launch_part_scan_nch(nch, ps, (unsigned int)std::min<uint64_t>(sp.n_tiles, (uint64_t)ctx->sm_count * 3), ctx->stream);
            if (le_ != cudaSuccess) return cuda_fail(ctx, le_, "part_scan_kernel (attributes)"); });
        pgm::PartProbeParams pp;
        memset(&pp, 0, sizeof pp);
        pp.q = pq; pp.tab = sp.tab;
        KLAUNCH(PGM_K_SCAN_PROBE, "part_probe_kernel", pgm::part_probe_kernel<<<ctx->sm_count * 6, PGM_PART_THREADS, 0, ctx->stream>>>(pp));
        CU(cudaMemsetAsync(sp.tile_counter, 0, sizeof(unsigned int), ctx->stream));
        sp.hit_bits = pq.hit_bits;
        KLAUNCH(PGM_K_SCAN, "scan_kernel (behind the partitioned pre-filter)", launch_scan_nch(nch, sp, grid, ctx->stream, 2));
        return PGM_OK;
    }
    { int wrc = filter_window(ctx, true); if (wrc) return wrc; }
    if (plan_blocked(ctx, le - lb, sp.n_tiles, sp.sq)) {
        // L2-blocked pipeline: filter stage (text order) -> probe stage (table-region order) -> verify stage (read-range
        // order), then the fused kernel as a fallback that only runs when a queue overflowed
        pgm::StageQueues &q = sp.sq;
        int rc;
        if ((rc = ensure(ctx, ctx->sq_pos, ((size_t)q.pos_cap << q.region_bits) * sizeof(uint4))) ||
            (rc = ensure(ctx, ctx->sq_cand, (size_t)q.cand_cap * q.n_ranges * sizeof(uint2))) ||
            (rc = ensure(ctx, ctx->sq_counters, (4 * PGM_SQ_MAX + 2 + 8) * sizeof(unsigned int)))) return rc;
        unsigned int *cbase = ctx->sq_counters.as<unsigned int>();
        q.pos_entries = ctx->sq_pos.as<uint4>(); q.cand_entries = ctx->sq_cand.as<uint2>();
        q.pos_count = cbase; q.pos_cursor = cbase + PGM_SQ_MAX; q.cand_count = cbase + 2 * PGM_SQ_MAX; q.cand_cursor = cbase + 3 * PGM_SQ_MAX;
        q.overflow = cbase + 4 * PGM_SQ_MAX;
        q.counters = reinterpret_cast<unsigned long long *>(cbase + 4 * PGM_SQ_MAX + 2);
        CU(cudaMemsetAsync(cbase, 0, (4 * PGM_SQ_MAX + 2 + 8) * sizeof(unsigned int), ctx->stream));
        KLAUNCH(PGM_K_SCAN_FILTER, "scan_kernel (filter stage)", launch_scan_nch(nch, sp, grid, ctx->stream, 1));
        pgm::VerifyParams vp;
        memset(&vp, 0, sizeof vp);
        vp.tlo = sp.tlo; vp.thi = sp.thi;
        vp.pos_origin = sp.slice_origin + (uint64_t)sp.first_word * 32;
        vp.bit_origin = (uint64_t)sp.first_word * 32;
        vp.pg_len = pg;
        vp.seed_len = sp.seed_len; vp.parts = sp.parts; vp.max_mm = sp.max_mm; vp.min_mm = sp.min_mm; vp.rev_mode = sp.rev_mode;
        vp.tab = sp.tab; vp.reads = sp.reads; vp.pr = sp.pr; vp.sq = q;
        KLAUNCH(PGM_K_SCAN_PROBE, "probe_kernel", pgm::probe_kernel<<<ctx->sm_count * 6, PGM_PROBE_THREADS, 0, ctx->stream>>>(vp));
        KLAUNCH(PGM_K_SCAN_VERIFY, "verify_kernel",
                if (ctx->lq_stride16 == 4) pgm::verify_kernel<true><<<ctx->sm_count * 4, PGM_VERIFY_THREADS, 0, ctx->stream>>>(vp);
                else pgm::verify_kernel<false><<<ctx->sm_count * 4, PGM_VERIFY_THREADS, 0, ctx->stream>>>(vp));
        sp.only_if = q.overflow;
        CU(cudaMemsetAsync(sp.tile_counter, 0, sizeof(unsigned int), ctx->stream));
    }
    // one launch per filter slice (one, unless the pattern set is far beyond an L2-resident filter): every launch streams the
    // whole text range again (2 bits per base) but gathers from its own 64 MB of the filter only
    const uint32_t n_slices = (ctx->filter_words && !sp.only_if) ? 1u << ctx->filter_slice_bits : 1u;
    for (uint32_t h = 0; h < n_slices; h++) {
        sp.slice = h;
        if (h) CU(cudaMemsetAsync(sp.tile_counter, 0, sizeof(unsigned int), ctx->stream));
        KLAUNCH(PGM_K_SCAN, "scan_kernel", launch_scan_nch(nch, sp, grid, ctx->stream, 0));
    }
    return filter_window(ctx, false);
}
} // namespace

extern "C" {

int pgm_scan_pass(pgm_ctx *ctx, int rev_mode) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (!ctx->phase_active) return fail(ctx, PGM_ERR_STATE, "pgm_scan_pass: pgm_match_begin has not been called");
    if (!ctx->has_text) return fail(ctx, PGM_ERR_STATE, "pgm_scan_pass: pgm_set_text has not been called");
    CU(cudaSetDevice(ctx->device));
    const uint64_t n = ctx->seed_span(), pg = ctx->pg_len;
    int rc;
    if (pg < n || ctx->n_reads() == 0) return finish_text_upload(ctx);
    // owned window starts of this pass (forward global coordinates)
    const uint64_t fb = ctx->own_begin, fe = std::min<uint64_t>(ctx->own_end, pg - n + 1);
    if (rev_mode || !ctx->text_pending) {
        if ((rc = finish_text_upload(ctx))) return rc;
        return scan_range(ctx, rev_mode, fb, fe);
    }
    // forward pass over a host text that is still arriving: pack each chunk as it lands and scan the window
    // starts whose windows and read alignments (<= 255 bases to the right) lie entirely in the chunks seen so far
    if ((rc = enqueue_text_copies(ctx, true))) return rc;
    uint64_t done = fb;
    const uint64_t nchunks = text_chunks(ctx);
    for (uint64_t c = 0; c < nchunks; c++) {
        if ((rc = text_chunk_ready(ctx, c))) return rc;
        CU(cudaStreamWaitEvent(ctx->stream, ctx->text_ev[c], 0));
        if ((rc = pack_text_chunk(ctx, ctx->ascii_stage.as<uint8_t>(), c))) return rc;
        const uint64_t have = ctx->slice_begin + std::min<uint64_t>(ctx->slice_len, (c + 1) * TEXT_CHUNK_BASES);   // bases [slice_begin, have) are packed
        uint64_t upto = c + 1 == nchunks ? fe : (have > 2 * 256 ? std::min<uint64_t>(fe, have - 2 * 256) : 0);
        // (the partitioned pre-filter streams the whole bucket array once per launch: launches of a quarter of the text)
        if (c + 1 != nchunks && part_wanted(ctx) && ctx->part_scan == 1 && (upto <= done || upto - done < std::max<uint64_t>((fe - fb) / 4, 256ull << 20))) continue;
        if (upto > done) {
            if ((rc = scan_range(ctx, 0, done, upto))) return rc;
            done = upto;
        }
    }
    ctx->text_pending = false;
    return rc_text(ctx);
}

int pgm_get_accumulators(pgm_ctx *ctx, pgm_accumulators *out) {
    if (!ctx || !out) return PGM_ERR_INVALID_ARG;
    if (!ctx->has_reads) return fail(ctx, PGM_ERR_STATE, "pgm_get_accumulators: no reads");
    CU(cudaSetDevice(ctx->device));
    const uint32_t n = ctx->n_reads();
    int rc;
    if ((rc = ensure(ctx, ctx->keys, (size_t)std::max<uint32_t>(n, 1) * 8))) return rc;
    if (n) {
        KLAUNCH(PGM_K_ACCUM, "export_keys_kernel", pgm::export_keys_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(
            reads_view(ctx), n, ctx->keys.as<long long>()));
    }
    out->best_key = ctx->keys.p; out->first_other_order = ctx->first_order.p;
    out->same_pos_mask = ctx->same_mask.p; out->same_pos_mm = ctx->same_mm.p;
    out->touched = ctx->touched.p; out->n_reads = n;
    return PGM_OK;
}

int pgm_put_accumulators(pgm_ctx *ctx) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (!ctx->has_reads || !ctx->keys.p) return fail(ctx, PGM_ERR_STATE, "pgm_put_accumulators: pgm_get_accumulators has not been called");
    CU(cudaSetDevice(ctx->device));
    const uint32_t n = ctx->n_reads();
    if (n) {
        KLAUNCH(PGM_K_ACCUM, "import_keys_kernel", pgm::import_keys_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(
            reads_view(ctx), n, ctx->keys.as<long long>()));
    }
    return PGM_OK;
}

} // extern "C"

namespace {
// resolve (+ outputs and histogram when `fin`: the last pass of pgm_map_reads)
int resolve_impl(pgm_ctx *ctx, int rev_mode, bool fin) {
    if (!ctx->phase_active) return fail(ctx, PGM_ERR_STATE, "pgm_resolve_pass: pgm_match_begin has not been called");
    CU(cudaSetDevice(ctx->device));
    const uint32_t n = ctx->n_reads();
    if (!n) return PGM_OK;
    if (fin) {
        CU(cudaMemsetAsync(ctx->hist.p, 0, 257 * sizeof(unsigned long long), ctx->stream));
        KLAUNCH(PGM_K_RESOLVE, "resolve_kernel", pgm::resolve_kernel<true><<<grid_for(n, 256), 256, 0, ctx->stream>>>(
            reads_view(ctx), per_read(ctx), n, ctx->pg_len, ctx->shift_unit(), ctx->parts, ctx->max_mm, ctx->min_mm, rev_mode ? 1 : 0,
            ctx->out_pos.as<unsigned long long>(), ctx->out_rc.as<uint8_t>(), ctx->out_mm.as<uint8_t>(), ctx->hist.as<unsigned long long>()));
        ctx->outputs_valid = true;
    } else {
        KLAUNCH(PGM_K_RESOLVE, "resolve_kernel", pgm::resolve_kernel<false><<<grid_for(n, 256), 256, 0, ctx->stream>>>(
            reads_view(ctx), per_read(ctx), n, ctx->pg_len, ctx->shift_unit(), ctx->parts, ctx->max_mm, ctx->min_mm, rev_mode ? 1 : 0,
            nullptr, nullptr, nullptr, nullptr));
    }
    CU(cudaMemsetAsync(ctx->touched.p, 0, sizeof(int), ctx->stream));
    ctx->aux_clean = true;
    return PGM_OK;
}
} // namespace

extern "C" {

int pgm_resolve_pass(pgm_ctx *ctx, int rev_mode) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    return resolve_impl(ctx, rev_mode, false);
}

int pgm_upload(pgm_ctx *ctx) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((ctx->has_reads && (rc = upload_reads(ctx, false))) || (ctx->has_text && (rc = finish_text_upload(ctx)))) return rc;
    CU(cudaStreamSynchronize(ctx->copy_stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PGM_OK;
}

int pgm_get_results(pgm_ctx *ctx, uint64_t *out_pos, uint8_t *out_rc, uint8_t *out_mm, pgm_stats *stats) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (!ctx->has_reads) return fail(ctx, PGM_ERR_STATE, "pgm_get_results: no reads");
    CU(cudaSetDevice(ctx->device));
    const uint32_t n = ctx->n_reads();
    {   // inputs that were set but never matched against: bring them in, so that the state and the symbol check are real
        int rc;
        if ((rc = upload_reads(ctx, false)) || (rc = finish_text_upload(ctx))) return rc;
    }
    if (n) {
        if (!ctx->outputs_valid) {
            CU(cudaMemsetAsync(ctx->hist.p, 0, 257 * sizeof(unsigned long long), ctx->stream));
            KLAUNCH(PGM_K_FINALIZE, "finalize_kernel", pgm::finalize_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(
                reads_view(ctx), n, ctx->out_pos.as<unsigned long long>(), ctx->out_rc.as<uint8_t>(), ctx->out_mm.as<uint8_t>(),
                ctx->hist.as<unsigned long long>()));
            ctx->outputs_valid = true;
        }
        int rc2;
        if (out_pos && (rc2 = d2h_out(ctx, out_pos, ctx->out_pos.p, (size_t)n * 8))) return rc2;
        if (out_rc && (rc2 = d2h_out(ctx, out_rc, ctx->out_rc.p, n))) return rc2;
        if (out_mm && (rc2 = d2h_out(ctx, out_mm, ctx->out_mm.p, n))) return rc2;
    } else {
        CU(cudaMemsetAsync(ctx->hist.p, 0, 257 * sizeof(unsigned long long), ctx->stream));
    }
    unsigned long long h[257], cnt[16];
    int bad = 0;
    CU(cudaMemcpyAsync(h, ctx->hist.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(cnt, ctx->counters.p, sizeof cnt, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&bad, ctx->err_flag.p, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (bad) return fail(ctx, PGM_ERR_BAD_SYMBOL, "pseudogenome text contains a symbol outside ACGT (2-bit text planes cannot hold it)");
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (int k = 0; k < 256; k++) stats->per_mm[k] = h[k];
        stats->matched = (uint64_t)n - h[255];
        stats->patterns_inserted = cnt[4];
        stats->table_slots = ctx->n_slots;
        stats->candidates = cnt[0]; stats->verified = cnt[1]; stats->accepted = cnt[2]; stats->filter_positives = cnt[3];
    }
    return PGM_OK;
}

} // extern "C"

namespace {
// initParams + calcCoprimes of CopMEMMatcher (copmem/CopMEMMatcher.cpp:71-137) for targetMatchLength = L and
// minMatchLength = L; false where the reference exits ("Minimal matching length too short", "L and K mismatch")
bool copmem_derive(uint32_t L, uint64_t N, uint32_t &K, uint32_t &k1, uint32_t &k2, uint32_t &hash_size) {
    int k;
    if (L > 110) k = 56; else if (L > 62) k = 44; else if (L > 53) k = 40; else if (L > 46) k = 36;
    else if (L > 42) k = 32; else if (L > 32) k = 28; else k = ((int)L / 4 - 1) * 4;
    if (L < 24) return false;
    k = std::min(k, ((int)L / 4 - 1) * 4);
    const int t = (int)L - k + 1;
    if (t <= 0) return false;
    int a, b;
    if (t >= 20) {
        a = 1; while ((a + 1) * (a + 1) <= t) a++;          // (int) pow(t, 0.5)
        a += 1; b = a - 1;
        if (a * b > t) { --b; --a; }
    } else if (t >= 15) { a = 5; b = 3; } else if (t >= 12) { a = 4; b = 3; } else if (t >= 10) { a = 5; b = 2; }
    else if (t >= 6) { a = 3; b = 2; } else { a = t; b = 1; }
    K = (uint32_t)k; k1 = (uint32_t)a; k2 = (uint32_t)b;
    int i = 24;                                              // HASH_SIZE_MIN_ORDER .. HASH_SIZE_MAX_ORDER
    do { hash_size = 1u << (i++); } while (i <= 31 && hash_size < N / (uint64_t)a);
    return true;
}

// exclusive prefix sums out[0..n] of v[0..n) (of min(v, 13) when capped) on the context's stream
template <uint32_t CAP>
int device_scan(pgm_ctx *ctx, const uint32_t *v, uint32_t n, uint32_t *out) {
    const uint32_t n_blocks = grid_for(n, PGM_CM_SCAN_BLOCK);
    int rc;
    if ((rc = ensure(ctx, ctx->cm_sums, ((size_t)n_blocks + 1) * 8))) return rc;
    unsigned long long *sums = ctx->cm_sums.as<unsigned long long>();
    KLAUNCH(PGM_K_COPMEM_INDEX, "cm_block_sums_kernel", pgm::cm_block_sums_kernel<CAP><<<n_blocks, PGM_CM_SCAN_BLOCK, 0, ctx->stream>>>(v, n, sums));
    KLAUNCH(PGM_K_COPMEM_INDEX, "mismatch_scan_kernel", pgm::mismatch_scan_kernel<<<1, 1024, 0, ctx->stream>>>(sums, n_blocks, sums + n_blocks));
    KLAUNCH(PGM_K_COPMEM_INDEX, "cm_block_scan_kernel", pgm::cm_block_scan_kernel<CAP><<<n_blocks, PGM_CM_SCAN_BLOCK, 0, ctx->stream>>>(v, n, sums, out));
    return PGM_OK;
}
} // namespace

extern "C" {

int pgm_copmem_begin(pgm_ctx *ctx, uint32_t part_len, uint32_t max_mm, uint32_t min_mm, int continuation) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (!ctx->has_reads || !ctx->has_text) return fail(ctx, PGM_ERR_STATE, "pgm_copmem_begin: set the text and the reads first");
    if (ctx->slice_begin != 0 || ctx->slice_len != ctx->pg_len)
        return fail(ctx, PGM_ERR_STATE, "pgm_copmem_begin: mode 'c' needs the whole text on this GPU (not a text shard)");
    if (max_mm > 127 || min_mm > 127) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_copmem_begin: mismatch limits must be <= 127");
    if (part_len == 0 || part_len > ctx->read_len) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_copmem_begin: need 1 <= part_len <= read_len");
    uint32_t K, k1, k2, hs;
    if (!copmem_derive(part_len, ctx->pg_len, K, k1, k2, hs))
        return fail(ctx, PGM_ERR_UNSUPPORTED, "pgm_copmem_begin: seed length below 24 (the reference exits: minimal matching length too short)");
    if (ctx->pg_len / k1 >= 0xFFFFFFF0ull) return fail(ctx, PGM_ERR_UNSUPPORTED, "pgm_copmem_begin: text too long for 32-bit sample indices");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = upload_reads(ctx, false))) return rc;
    ctx->cm_K = K; ctx->cm_k1 = k1; ctx->cm_k2 = k2; ctx->cm_hash_size = hs;
    ctx->max_mm = max_mm; ctx->min_mm = min_mm;
    ctx->outputs_valid = false;
    ctx->phase_active = false;
    ctx->copmem_active = true;
    const uint32_t n = ctx->n_reads();
    if (!continuation) {
        CU(cudaMemsetAsync(ctx->counters.p, 0, 16 * sizeof(unsigned long long), ctx->stream));
        if (n && !ctx->state_fresh) {
            // DefaultReadsMatcher::initMatching + readMismatchesCount = NOT_MATCHED_COUNT (ReadsMatchers.cpp:411-415)
            KLAUNCH(PGM_K_INIT_STATE, "reset_state_kernel", pgm::reset_state_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(
                reads_view(ctx), per_read(ctx), n, 1, 1));
            CU(cudaMemsetAsync(ctx->touched.p, 0, sizeof(int), ctx->stream));
            ctx->aux_clean = true;
        }
    }
    return PGM_OK;
}

int pgm_copmem_pass(pgm_ctx *ctx, int rev_mode) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (!ctx->copmem_active) return fail(ctx, PGM_ERR_STATE, "pgm_copmem_pass: pgm_copmem_begin has not been called");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = finish_text_upload(ctx))) return rc;
    const uint32_t n = ctx->n_reads();
    const uint64_t N = ctx->pg_len;
    if (!n) return PGM_OK;
    pgm::CopmemParams cp;
    memset(&cp, 0, sizeof cp);
    cp.tlo = (rev_mode ? ctx->r_lo : ctx->f_lo).as<uint32_t>() + PGM_PAD_WORDS;
    cp.thi = (rev_mode ? ctx->r_hi : ctx->f_hi).as<uint32_t>() + PGM_PAD_WORDS;
    cp.pg_len = N;
    cp.K = ctx->cm_K; cp.k1 = ctx->cm_k1; cp.k2 = ctx->cm_k2; cp.hash_mask = ctx->cm_hash_size - 1;
    cp.n_sampled = N >= cp.K ? (uint32_t)((N - cp.K) / cp.k1 + 1) : 0u;
    const size_t hs = ctx->cm_hash_size;
    if ((rc = ensure(ctx, ctx->cm_count, (hs + 1) * 4)) || (rc = ensure(ctx, ctx->cm_start, (hs + 1) * 4)) ||
        (rc = ensure(ctx, ctx->cm_cumm, (hs + 2) * 4)) || (rc = ensure(ctx, ctx->cm_fill, (size_t)std::max<uint32_t>(cp.n_sampled, 1) * 4)) ||
        (rc = ensure(ctx, ctx->cm_hash, (size_t)std::max<uint32_t>(cp.n_sampled, 1) * 4)) ||
        (rc = ensure(ctx, ctx->cm_all, (size_t)std::max<uint32_t>(cp.n_sampled, 1) * 4)) ||
        (rc = ensure(ctx, ctx->cm_entries, (size_t)std::max<uint32_t>(cp.n_sampled, 1) * 4))) return rc;
    cp.count = ctx->cm_count.as<uint32_t>(); cp.start_all = ctx->cm_start.as<uint32_t>(); cp.cumm = ctx->cm_cumm.as<uint32_t>();
    cp.sample_rank = ctx->cm_fill.as<uint32_t>(); cp.sample_hash = ctx->cm_hash.as<uint32_t>(); cp.all_entries = ctx->cm_all.as<uint32_t>();
    cp.entries = ctx->cm_entries.as<uint32_t>();
    cp.reads = reads_view(ctx); cp.n_reads = n; cp.max_mm = ctx->max_mm; cp.min_mm = ctx->min_mm; cp.rev_mode = rev_mode ? 1 : 0;
    // the index of this pass's text: new CopMEMMatcher(pgPtr, pgLength, partLength) (ReadsMatchers.cpp:424)
    ctx->mem_index_valid = false;      // (stage 7 keeps its index in the same buffers)
    CU(cudaMemsetAsync(cp.count, 0, (hs + 1) * 4, ctx->stream));
    if (cp.n_sampled)
        KLAUNCH(PGM_K_COPMEM_INDEX, "cm_hash_kernel", pgm::cm_hash_kernel<<<grid_for(cp.n_sampled, PGM_CM_THREADS), PGM_CM_THREADS, 0, ctx->stream>>>(cp));
    if ((rc = device_scan<0>(ctx, cp.count, (uint32_t)hs, cp.start_all)) ||
        (rc = device_scan<PGM_CM_COLLISIONS_LIMIT + 1>(ctx, cp.count, (uint32_t)hs, cp.cumm))) return rc;
    if (cp.n_sampled) {
        KLAUNCH(PGM_K_COPMEM_INDEX, "cm_scatter_kernel", pgm::cm_scatter_kernel<<<grid_for(cp.n_sampled, PGM_CM_THREADS), PGM_CM_THREADS, 0, ctx->stream>>>(cp));
        KLAUNCH(PGM_K_COPMEM_INDEX, "cm_select_kernel", pgm::cm_select_kernel<<<grid_for(hs, PGM_CM_THREADS), PGM_CM_THREADS, 0, ctx->stream>>>(cp));
    }
    // every read on its own (ReadsMatchers.cpp:425-448)
    ctx->state_fresh = false;
    ctx->outputs_valid = false;
    const uint32_t n_off = ctx->read_len >= cp.K ? (ctx->read_len - cp.K) / cp.k2 + 1 : 0;
    if (ctx->cm_warp && n_off) {
        // staged query (pgm_copmem_warp.cuh): compact bucket directory, then per batch of reads: stage 1 (warp per read), stage 2
        // (replay, thread per read), and the thread-per-read kernel for the reads whose table overflowed
        if ((rc = ensure(ctx, ctx->cm_nib, hs / 8 * 4 + 64)) || (rc = ensure(ctx, ctx->cm_coarse, hs / 64 * 4 + 64))) return rc;
        cp.nib = ctx->cm_nib.as<uint32_t>(); cp.coarse = ctx->cm_coarse.as<uint32_t>();
        KLAUNCH(PGM_K_COPMEM_INDEX, "cm_compact_kernel", pgm::cm_compact_kernel<<<grid_for(hs / 8, PGM_CM_THREADS), PGM_CM_THREADS, 0, ctx->stream>>>(cp));
        const uint32_t cap = n_off * (PGM_CM_COLLISIONS_LIMIT + 1);
        const size_t per_read = (size_t)n_off + cap + PGM_CMW_VT * 8 + 1;
        const uint32_t batch = (uint32_t)std::min<uint64_t>(n, std::max<uint64_t>(1024, (6ull << 30) / per_read));
        if ((rc = ensure(ctx, ctx->cmw_lens, (size_t)batch * n_off)) || (rc = ensure(ctx, ctx->cmw_cand, (size_t)batch * cap)) ||
            (rc = ensure(ctx, ctx->cmw_vt, (size_t)batch * PGM_CMW_VT * 8)) || (rc = ensure(ctx, ctx->cmw_nvt, batch))) return rc;
        for (uint32_t rb = 0; rb < n; rb += batch) {
            pgm::CmwParams q;
            memset(&q, 0, sizeof q);
            q.c = cp; q.n_off = n_off; q.cap = cap; q.r_begin = rb; q.r_count = std::min(batch, n - rb);
            q.lens = ctx->cmw_lens.as<uint8_t>(); q.cand = ctx->cmw_cand.as<uint8_t>();
            q.vt = ctx->cmw_vt.as<unsigned long long>(); q.n_vt = ctx->cmw_nvt.as<uint8_t>();
            KLAUNCH(PGM_K_COPMEM_STAGE1, "cmw_stage1_kernel", pgm::cmw_stage1_kernel<<<grid_for(q.r_count, PGM_CMW_WARPS), PGM_CMW_WARPS * 32, 0, ctx->stream>>>(q));
            KLAUNCH(PGM_K_COPMEM_STAGE2, "cmw_stage2_kernel", pgm::cmw_stage2_kernel<<<grid_for(q.r_count, PGM_CM_THREADS), PGM_CM_THREADS, 0, ctx->stream>>>(
                q, ctx->counters.as<unsigned long long>()));
            pgm::CopmemParams cf = cp;
            cf.only_marked = q.n_vt; cf.marked_base = rb; cf.marked_count = q.r_count;
            KLAUNCH(PGM_K_COPMEM_QUERY, "cm_query_kernel (table overflow)", pgm::cm_query_kernel<<<grid_for(n, PGM_CM_THREADS), PGM_CM_THREADS, 0, ctx->stream>>>(
                cf, ctx->counters.as<unsigned long long>()));
        }
        return PGM_OK;
    }
    KLAUNCH(PGM_K_COPMEM_QUERY, "cm_query_kernel", pgm::cm_query_kernel<<<grid_for(n, PGM_CM_THREADS), PGM_CM_THREADS, 0, ctx->stream>>>(
        cp, ctx->counters.as<unsigned long long>()));
    return PGM_OK;
}

int pgm_get_mismatches(pgm_ctx *ctx, uint64_t *out_offsets, uint8_t *out_pos, uint8_t *out_syms, uint64_t capacity, uint64_t *total) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (!ctx->has_reads || !ctx->has_text) return fail(ctx, PGM_ERR_STATE, "pgm_get_mismatches: set the text and the reads first");
    if (ctx->slice_begin != 0 || ctx->slice_len != ctx->pg_len)
        return fail(ctx, PGM_ERR_STATE, "pgm_get_mismatches: needs the whole text on this GPU (not a text shard)");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = upload_reads(ctx, false)) || (rc = finish_text_upload(ctx))) return rc;
    const uint32_t n = ctx->n_reads();
    const uint32_t n_blocks = grid_for(std::max<uint32_t>(n, 1), PGM_MIS_THREADS);
    if ((rc = ensure(ctx, ctx->mis_sums, ((size_t)n_blocks + 1) * 8)) || (rc = ensure(ctx, ctx->mis_offsets, ((size_t)n + 1) * 8))) return rc;
    unsigned long long *sums = ctx->mis_sums.as<unsigned long long>();
    unsigned long long tot = 0;
    if (n) {
        KLAUNCH(PGM_K_MISMATCHES, "mismatch_count_kernel", pgm::mismatch_count_kernel<<<n_blocks, PGM_MIS_THREADS, 0, ctx->stream>>>(reads_view(ctx), n, sums));
        KLAUNCH(PGM_K_MISMATCHES, "mismatch_scan_kernel", pgm::mismatch_scan_kernel<<<1, 1024, 0, ctx->stream>>>(sums, n_blocks, sums + n_blocks));
        CU(cudaMemcpyAsync(&tot, sums + n_blocks, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    } else {
        CU(cudaMemsetAsync(ctx->mis_offsets.p, 0, 8, ctx->stream));
    }
    if (total) *total = tot;
    if (n && (out_pos || out_syms || out_offsets)) {
        if ((out_pos || out_syms) && capacity < tot)
            return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_get_mismatches: capacity is smaller than the number of mismatches (sum of k * per_mm[k])");
        if ((rc = ensure(ctx, ctx->mis_pos, std::max<size_t>(tot, 1))) || (rc = ensure(ctx, ctx->mis_syms, std::max<size_t>(tot, 1)))) return rc;
        pgm::MismatchParams mp;
        memset(&mp, 0, sizeof mp);
        mp.flo = ctx->f_lo.as<uint32_t>() + PGM_PAD_WORDS; mp.fhi = ctx->f_hi.as<uint32_t>() + PGM_PAD_WORDS;
        mp.rlo = ctx->r_lo.as<uint32_t>() + PGM_PAD_WORDS; mp.rhi = ctx->r_hi.as<uint32_t>() + PGM_PAD_WORDS;
        mp.pg_len = ctx->pg_len; mp.reads = reads_view(ctx); mp.n_reads = n;
        mp.block_base = sums; mp.out_offsets = ctx->mis_offsets.as<unsigned long long>();
        mp.out_pos = ctx->mis_pos.as<uint8_t>(); mp.out_syms = ctx->mis_syms.as<uint8_t>();
        mp.capacity = tot; mp.err = ctx->err_flag.as<int>();
        KLAUNCH(PGM_K_MISMATCHES, "mismatch_emit_kernel", pgm::mismatch_emit_kernel<<<n_blocks, PGM_MIS_THREADS, 0, ctx->stream>>>(mp));
    }
    if (out_offsets) CU(cudaMemcpyAsync(out_offsets, ctx->mis_offsets.p, ((size_t)n + 1) * 8, cudaMemcpyDefault, ctx->stream));
    if (out_pos && tot) CU(cudaMemcpyAsync(out_pos, ctx->mis_pos.p, tot, cudaMemcpyDefault, ctx->stream));
    if (out_syms && tot) CU(cudaMemcpyAsync(out_syms, ctx->mis_syms.p, tot, cudaMemcpyDefault, ctx->stream));
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, ctx->err_flag.p, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (bad == 1) return fail(ctx, PGM_ERR_BAD_SYMBOL, "pseudogenome text contains a symbol outside ACGT (2-bit text planes cannot hold it)");
    if (bad) return fail(ctx, PGM_ERR_CUDA, "pgm_get_mismatches: a mismatch list does not have readMismatchesCount entries (internal error)");
    return PGM_OK;
}

// ------------------------------------------------------------------------------------------ routed multi-GPU scheme
} // extern "C"

namespace {

// window starts of one pass: [0, pg_len - span]; GPU g owns [cut(g), cut(g + 1)), cuts are multiples of 128 (tile copies stay
// 16-byte aligned); round r of GPU g covers [cut(g) + r * round_windows, ...)
uint64_t route_windows(const pgm_ctx *ctx) {
    const uint64_t span = ctx->seed_span();
    return ctx->pg_len >= span ? ctx->pg_len - span + 1 : 0;
}
uint64_t route_cut(const pgm_ctx *ctx, int k) {
    const uint64_t nw = route_windows(ctx);
    if (k >= ctx->route.world) return nw;
    return (uint64_t)((unsigned __int128)nw * (unsigned)k / (unsigned)ctx->route.world) / 128 * 128;
}
void route_range(const pgm_ctx *ctx, int g, uint32_t round, uint64_t &b, uint64_t &e) {
    const uint64_t lo = route_cut(ctx, g), hi = route_cut(ctx, g + 1);
    b = std::min(hi, lo + (uint64_t)round * ctx->route.round_windows);
    e = std::min(hi, b + ctx->route.round_windows);
}
uint64_t route_longest(const pgm_ctx *ctx) {
    uint64_t longest = 0;
    for (int g = 0; g < ctx->route.world; g++) longest = std::max(longest, route_cut(ctx, g + 1) - route_cut(ctx, g));
    return longest;
}
uint32_t route_rounds(const pgm_ctx *ctx) {
    return (uint32_t)std::max<uint64_t>(1, (route_longest(ctx) + ctx->route.round_windows - 1) / ctx->route.round_windows);
}
// window starts a GPU emits in one round at most (what the exchange buffers are sized for)
uint64_t route_round_size(const pgm_ctx *ctx) { return std::max<uint64_t>(PGM_TILE_POS, std::min(ctx->route.round_windows, route_longest(ctx))); }

uint32_t route_cap(uint64_t total, int world) {
    // a destination's share + 25 % + slack, never more than everything (hot seeds / low-complexity text skew the shares)
    return (uint32_t)std::min<uint64_t>(std::min<uint64_t>(total, total / world + total / (4 * world) + (1u << 20)) + 16, 0xFFFFFFF0ull);
}

// device counters of the exchanges: per buffer slot and per kind {counts[16], overflow, tile counter, ..} = 32 words
constexpr int RT_KIND_WORDS = 32, RT_SLOT_WORDS = 3 * RT_KIND_WORDS;
unsigned int *route_counts(pgm_ctx *ctx, int kind) { return ctx->rt_counters.as<unsigned int>() + ctx->route.slot * RT_SLOT_WORDS + kind * RT_KIND_WORDS; }
unsigned int *route_overflow(pgm_ctx *ctx, int kind) { return route_counts(ctx, kind) + PGM_ROUTE_MAX_WORLD; }
unsigned int *route_tile_counter(pgm_ctx *ctx, int kind) { return route_counts(ctx, kind) + PGM_ROUTE_MAX_WORLD + 1; }

// copies the per-destination counts of `kind` (and the overflow flag) to the host; synchronizes the stream
// queues the copy of the per-destination counts of (kind, current slot) to pinned host memory behind the emit kernels, and an event
int route_post_counts(pgm_ctx *ctx, int kind) {
    const int sl = ctx->route.slot;
    if (!ctx->rt_counts_host) CU(cudaMallocHost(reinterpret_cast<void **>(&ctx->rt_counts_host), 2 * RT_SLOT_WORDS * sizeof(unsigned int)));
    if (!ctx->counts_ev[kind][sl]) CU(cudaEventCreateWithFlags(&ctx->counts_ev[kind][sl], cudaEventDisableTiming));
    CU(cudaMemcpyAsync(ctx->rt_counts_host + sl * RT_SLOT_WORDS + kind * RT_KIND_WORDS, route_counts(ctx, kind), RT_KIND_WORDS * sizeof(unsigned int),
                       cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->counts_ev[kind][sl], ctx->stream));
    return PGM_OK;
}

// waits for the emit step of (kind, current slot) — not for anything queued behind it — and describes its send buffer
int route_fetch_counts(pgm_ctx *ctx, int kind, pgm_route_buffer *send, void *base, uint64_t stride, uint32_t entry_bytes, const char *what) {
    const int sl = ctx->route.slot;
    if (!ctx->rt_counts_host || !ctx->counts_ev[kind][sl]) return fail(ctx, PGM_ERR_STATE, std::string(what) + ": nothing was emitted");
    CU(cudaEventSynchronize(ctx->counts_ev[kind][sl]));
    const unsigned int *h = ctx->rt_counts_host + sl * RT_SLOT_WORDS + kind * RT_KIND_WORDS;
    if (h[PGM_ROUTE_MAX_WORLD])
        return fail(ctx, PGM_ERR_STATE, std::string(what) + ": an exchange queue overflowed (a GPU's share of the hashes / candidates is far above "
                                        "the average: extremely skewed input); rerun with the read-sharded scheme");
    memset(send, 0, sizeof *send);
    send->base = base; send->stride_bytes = stride; send->entry_bytes = entry_bytes; send->world = (uint32_t)ctx->route.world;
    for (int d = 0; d < ctx->route.world; d++) send->count[d] = h[d];
    return PGM_OK;
}

template <int NCH>
cudaError_t launch_route_scan(const pgm::RouteScanParams &sp, unsigned int grid, cudaStream_t s) {
    const size_t smem = sizeof(pgm::RouteScanShared);
    cudaError_t e = cudaFuncSetAttribute(pgm::route_scan_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pgm::route_scan_kernel<NCH><<<grid, PGM_ROUTE_THREADS, smem, s>>>(sp);
    return cudaSuccess;
}

} // namespace

extern "C" {

int pgm_route_config(pgm_ctx *ctx, int rank, int world, const uint64_t *read_begin, uint64_t round_windows) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (world < 1 || world > PGM_ROUTE_MAX_WORLD || rank < 0 || rank >= world || !read_begin)
        return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_route_config: need 1 <= world <= 16, 0 <= rank < world and the read ranges");
    for (int k = 0; k < world; k++)
        if (read_begin[k] > read_begin[k + 1]) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_route_config: read ranges must ascend");
    if (read_begin[world] >= 0xFFFFFFFFull) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_route_config: too many reads");
    CU(cudaSetDevice(ctx->device));
    ctx->route.rank = rank; ctx->route.world = world;
    for (int k = 0; k <= world; k++) ctx->route.read_begin[k] = read_begin[k];
    ctx->route.round_windows = round_windows ? std::max<uint64_t>(PGM_TILE_POS, (round_windows + PGM_TILE_POS - 1) / PGM_TILE_POS * PGM_TILE_POS)
                                             : (512ull << 20);
    int rc;
    if ((rc = ensure(ctx, ctx->rt_counters, 2 * RT_SLOT_WORDS * sizeof(unsigned int)))) return rc;
    ctx->route.slot = 0;
    return PGM_OK;
}

int pgm_route_slot(pgm_ctx *ctx, int slot) {
    if (!ctx || slot < 0 || slot > 1) return PGM_ERR_INVALID_ARG;
    ctx->route.slot = slot;
    return PGM_OK;
}

int pgm_route_rounds(pgm_ctx *ctx, uint32_t *rounds) {
    if (!ctx || !rounds) return PGM_ERR_INVALID_ARG;
    if (!ctx->route.world || !ctx->phase_active || !ctx->has_text) return fail(ctx, PGM_ERR_STATE, "pgm_route_rounds: call pgm_route_config, pgm_set_text and pgm_route_begin first");
    *rounds = route_rounds(ctx);
    return PGM_OK;
}

int pgm_route_begin(pgm_ctx *ctx, uint32_t seed_len, uint32_t parts, uint32_t max_mm, uint32_t min_mm, int continuation, pgm_route_buffer *send) {
    if (!ctx || !send) return PGM_ERR_INVALID_ARG;
    const pgm_ctx::Route &rt = ctx->route;
    if (!rt.world) return fail(ctx, PGM_ERR_STATE, "pgm_route_begin: pgm_route_config has not been called");
    if (!ctx->has_reads) return fail(ctx, PGM_ERR_STATE, "pgm_route_begin: pgm_set_reads has not been called");
    if (ctx->n_reads() != rt.read_begin[rt.rank + 1] - rt.read_begin[rt.rank])
        return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_route_begin: pgm_set_reads must be given exactly this GPU's read range");
    if (seed_len == 0 || parts == 0 || (uint64_t)seed_len * parts > ctx->read_len)
        return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_route_begin: need seed_len >= 1 and seed_len * parts <= read_len");
    if (max_mm > 127 || min_mm > 127) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_route_begin: mismatch limits must be <= 127");
    const uint32_t part_bits = (uint32_t)ceil_log2(parts);
    if ((rt.read_begin[rt.world] << part_bits) >= 0xFFFFFFFFull)
        return fail(ctx, PGM_ERR_UNSUPPORTED, "pgm_route_begin: pattern index exceeds 32 bits (reads << ceil(log2(parts)))");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = upload_reads(ctx, false))) return rc;
    const uint32_t n = ctx->n_reads();
    ctx->seed_len = seed_len; ctx->parts = parts; ctx->max_mm = max_mm; ctx->min_mm = min_mm; ctx->part_bits = part_bits;
    ctx->interleaved = false;
    ctx->outputs_valid = false;
    if (!continuation) CU(cudaMemsetAsync(ctx->counters.p, 0, 16 * sizeof(unsigned long long), ctx->stream));
    else CU(cudaMemsetAsync(ctx->counters.as<unsigned long long>() + 4, 0, sizeof(unsigned long long), ctx->stream));
    if (n && ((!continuation && !ctx->state_fresh) || !ctx->aux_clean)) {
        KLAUNCH(PGM_K_INIT_STATE, "reset_state_kernel", pgm::reset_state_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(
            reads_view(ctx), per_read(ctx), n, continuation ? 0 : 1, 1));
        CU(cudaMemsetAsync(ctx->touched.p, 0, sizeof(int), ctx->stream));
        ctx->aux_clean = true;
    }
    // my reads' seeds -> send segments by hash owner
    const uint64_t n_local = (uint64_t)n * parts;
    ctx->route.cap_pat = route_cap(n_local, rt.world);
    if ((rc = ensure(ctx, ctx->bq_entries, (size_t)ctx->route.cap_pat * rt.world * sizeof(uint4)))) return rc;
    CU(cudaMemsetAsync(route_counts(ctx, PGM_ROUTE_PATTERNS), 0, RT_KIND_WORDS * sizeof(unsigned int), ctx->stream));
    if (n) {
        pgm::BuildQueues q;
        q.entries = ctx->bq_entries.as<uint4>();
        q.count = route_counts(ctx, PGM_ROUTE_PATTERNS); q.cursor = nullptr;
        q.cap = ctx->route.cap_pat; q.region_bits = 0;
        q.route_world = (uint32_t)rt.world; q.read_base = (uint32_t)rt.read_begin[rt.rank]; q.overflow = route_overflow(ctx, PGM_ROUTE_PATTERNS);
        const uint32_t tail = seed_len % 32 ? (1u << (seed_len % 32)) - 1u : 0xFFFFFFFFu;
        const unsigned int grid = (unsigned int)std::min<uint64_t>(grid_for(n, PGM_BUILD_THREADS), (uint64_t)ctx->sm_count * 8);
        const size_t smem = (size_t)parts * PGM_BUILD_THREADS * sizeof(uint4);
        const bool fast = ctx->n_n == 0 && ctx->lq_stride16 == 4;
        pgm::TableView none;
        memset(&none, 0, sizeof none);
        KLAUNCH(PGM_K_ROUTE_BUILD, "build_table_kernel (routed)",
                if (fast) pgm::build_table_kernel<true><<<grid, PGM_BUILD_THREADS, smem, ctx->stream>>>(
                    reads_view(ctx), none, q, 0, n, seed_len, parts, min_mm, continuation, tail, 0, ctx->counters.as<unsigned long long>() + 4);
                else pgm::build_table_kernel<false><<<grid, PGM_BUILD_THREADS, smem, ctx->stream>>>(
                    reads_view(ctx), none, q, 0, n, seed_len, parts, min_mm, continuation, tail, 0, ctx->counters.as<unsigned long long>() + 4));
    }
    ctx->phase_active = true;
    ctx->copmem_active = false;
    if ((rc = route_post_counts(ctx, PGM_ROUTE_PATTERNS))) return rc;
    return route_fetch_counts(ctx, PGM_ROUTE_PATTERNS, send, ctx->bq_entries.p, (uint64_t)ctx->route.cap_pat * sizeof(uint4), sizeof(uint4), "pgm_route_begin");
}

int pgm_route_recv(pgm_ctx *ctx, int kind, uint64_t n_entries, void **ptr) {
    if (!ctx || !ptr || kind < 0 || kind > 2) return PGM_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    DevBuf &b = kind == PGM_ROUTE_PATTERNS ? ctx->rt_pat_recv : kind == PGM_ROUTE_WINDOWS ? ctx->rt_win_recv_() : ctx->rt_cand_recv_();
    const size_t eb = kind == PGM_ROUTE_PATTERNS ? 16 : 12;
    if (b.cap < std::max<size_t>(n_entries, 1) * eb) CU(cudaStreamSynchronize(ctx->stream));   // (a kernel may still read the old buffer)
    int rc;
    if ((rc = ensure(ctx, b, std::max<size_t>(n_entries, 1) * eb))) return rc;
    *ptr = b.p;
    return PGM_OK;
}

} // extern "C"

namespace {
// marks the receive buffer of (kind, current slot) as read by what has been queued so far
int route_mark_consumed(pgm_ctx *ctx, int kind) {
    const int sl = kind == PGM_ROUTE_PATTERNS ? 0 : ctx->route.slot;
    if (!ctx->consumed_ev[kind][sl]) CU(cudaEventCreateWithFlags(&ctx->consumed_ev[kind][sl], cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->consumed_ev[kind][sl], ctx->stream));
    return PGM_OK;
}

// orders the context's stream behind a pending peer pull of (kind, current slot)
int route_wait_pull(pgm_ctx *ctx, int kind) {
    const int sl = kind == PGM_ROUTE_PATTERNS ? 0 : ctx->route.slot;
    if (ctx->pull_pending[kind][sl]) {
        CU(cudaStreamWaitEvent(ctx->stream, ctx->pull_ev[kind][sl], 0));
        ctx->pull_pending[kind][sl] = false;
    }
    return PGM_OK;
}
} // namespace

extern "C" {

int pgm_route_export(pgm_ctx *ctx, int kind, pgm_route_peer *out) {
    if (!ctx || !out || kind < 0 || kind > 2) return PGM_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    DevBuf &b = kind == PGM_ROUTE_PATTERNS ? ctx->bq_entries : kind == PGM_ROUTE_WINDOWS ? ctx->rt_win_send_() : ctx->rt_cand_send_();
    memset(out, 0, sizeof *out);
    if (!b.p) return fail(ctx, PGM_ERR_STATE, "pgm_route_export: nothing has been emitted for this kind yet");
    out->pid = (uint64_t)getpid();
    out->device = ctx->device;
    out->ptr = b.p;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, b.p));
    static_assert(sizeof h == sizeof out->ipc_handle, "CUDA IPC handle size");
    memcpy(out->ipc_handle, &h, sizeof h);
    return PGM_OK;
}

int pgm_route_pull(pgm_ctx *ctx, int kind, const pgm_route_peer *peers, const pgm_route_buffer *peer_sends) {
    if (!ctx || !peers || !peer_sends || kind < 0 || kind > 2) return PGM_ERR_INVALID_ARG;
    const pgm_ctx::Route &rt = ctx->route;
    if (!rt.world) return fail(ctx, PGM_ERR_STATE, "pgm_route_pull: pgm_route_config has not been called");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->pull_stream) {
        CU(cudaStreamCreateWithFlags(&ctx->pull_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 3; k++) {
            CU(cudaStreamCreateWithFlags(&ctx->pull_more[k], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&ctx->pull_join[k], cudaEventDisableTiming));
        }
    }
    const int sl = kind == PGM_ROUTE_PATTERNS ? 0 : rt.slot;
    if (!ctx->pull_ev[kind][sl]) CU(cudaEventCreateWithFlags(&ctx->pull_ev[kind][sl], cudaEventDisableTiming));
    uint64_t total = 0;
    for (int s = 0; s < rt.world; s++) total += peer_sends[s].count[rt.rank];
    void *dst = nullptr;
    int rc;
    if ((rc = pgm_route_recv(ctx, kind, total, &dst))) return rc;
    // the copies may start once the previous reader of this receive buffer is done (NOT everything queued on the stream: the
    // caller keeps the queue filled with other steps while this transfer runs)
    if (ctx->consumed_ev[kind][sl]) {
        CU(cudaStreamWaitEvent(ctx->pull_stream, ctx->consumed_ev[kind][sl], 0));
        for (int k = 0; k < 3; k++) CU(cudaStreamWaitEvent(ctx->pull_more[k], ctx->consumed_ev[kind][sl], 0));
    }
    for (int k = 0; k < rt.world; k++) {
        const int s = (rt.rank + k) % rt.world;                    // own segment first, then round-robin over the peers
        cudaStream_t cs = (k & 3) == 0 ? ctx->pull_stream : ctx->pull_more[(k & 3) - 1];
        uint64_t off = 0;
        for (int q = 0; q < s; q++) off += peer_sends[q].count[rt.rank] * peer_sends[q].entry_bytes;
        const uint64_t bytes = peer_sends[s].count[rt.rank] * peer_sends[s].entry_bytes;
        if (!bytes) continue;
        const char *base = nullptr;
        if (peers[s].pid == (uint64_t)getpid()) base = static_cast<const char *>(peers[s].ptr);      // same process: the pointer itself
        else {
            const std::string key(reinterpret_cast<const char *>(peers[s].ipc_handle), sizeof peers[s].ipc_handle);
            for (auto &kv : ctx->ipc_open) if (kv.first == key) base = static_cast<const char *>(kv.second);
            if (!base) {
                cudaIpcMemHandle_t h;
                memcpy(&h, peers[s].ipc_handle, sizeof h);
                void *mapped = nullptr;
                CU(cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
                ctx->ipc_open.emplace_back(key, mapped);
                base = static_cast<const char *>(mapped);
            }
        }
        const char *src = base + (uint64_t)rt.rank * peer_sends[s].stride_bytes;
        CU(cudaMemcpyAsync(static_cast<char *>(dst) + off, src, bytes, cudaMemcpyDefault, cs));
    }
    for (int k = 0; k < 3; k++) {
        CU(cudaEventRecord(ctx->pull_join[k], ctx->pull_more[k]));
        CU(cudaStreamWaitEvent(ctx->pull_stream, ctx->pull_join[k], 0));
    }
    CU(cudaEventRecord(ctx->pull_ev[kind][sl], ctx->pull_stream));
    ctx->pull_pending[kind][sl] = true;
    return PGM_OK;
}

int pgm_route_build(pgm_ctx *ctx, uint64_t n_in) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (!ctx->route.world || !ctx->phase_active) return fail(ctx, PGM_ERR_STATE, "pgm_route_build: pgm_route_begin has not been called");
    if (n_in && ctx->rt_pat_recv.cap < n_in * 16) return fail(ctx, PGM_ERR_STATE, "pgm_route_build: the patterns have not been received (pgm_route_recv)");
    CU(cudaSetDevice(ctx->device));
    { int wrc = route_wait_pull(ctx, PGM_ROUTE_PATTERNS); if (wrc) return wrc; }
    const pgm_ctx::Route &rt = ctx->route;
    // table geometry for the patterns that actually arrived (hash skew included)
    const uint64_t want_slots = std::max<uint64_t>(512, n_in * (uint64_t)ctx->slots_per_pattern);
    const uint64_t nb64 = next_prime((want_slots + 3) / 4);
    if (nb64 >= 0x7FFFFFFFull) return fail(ctx, PGM_ERR_UNSUPPORTED, "pgm_route_build: table too large");
    const uint64_t n_ids = rt.read_begin[rt.world] << ctx->part_bits;         // next[] is indexed by the GLOBAL pattern id
    int rc;
    if ((rc = ensure(ctx, ctx->buckets, nb64 * 32)) || (rc = ensure(ctx, ctx->next, std::max<uint64_t>(n_ids, 1) * 4))) return rc;
    ctx->n_buckets = (uint32_t)nb64;
    ctx->n_slots = nb64 * 4;
    CU(cudaMemsetAsync(ctx->buckets.p, 0xFF, nb64 * 32, ctx->stream));
    int fbits = ctx->filter_log2_bits;
    if (fbits < 0) fbits = std::min(n_in > (100ull << 20) ? 29 : 28, std::max(15, ceil_log2(std::max<uint64_t>(n_in, 1) * 8)));
    ctx->filter_pair = 0;
    if (fbits > 0) {
        const double per = (double)(1ull << fbits) / (double)std::max<uint64_t>(n_in, 1);
        ctx->filter_k = ctx->filter_k_force ? (uint32_t)std::min(2, std::max(1, ctx->filter_k_force))
                                            : (uint32_t)std::min(2.0, std::max(1.0, 0.69 * per + 0.5));
        ctx->filter_words = 1u << (fbits - 5);
        const size_t fbytes = (size_t)1 << (fbits - 3);
        if ((rc = ensure(ctx, ctx->filter, fbytes))) return rc;
        CU(cudaMemsetAsync(ctx->filter.p, 0, fbytes, ctx->stream));
    } else ctx->filter_words = 0;
    ctx->bq_region_bits = 0; ctx->bq_pending = false;
    ctx->route.n_pat_in = n_in;
    if (n_in)
        KLAUNCH(PGM_K_ROUTE_BUILD, "route_insert_kernel", pgm::route_insert_kernel<<<ctx->sm_count * 8, PGM_INSERT_THREADS, 0, ctx->stream>>>(
            table_view(ctx), ctx->rt_pat_recv.as<uint4>(), n_in));
    return route_mark_consumed(ctx, PGM_ROUTE_PATTERNS);
}

int pgm_route_scan(pgm_ctx *ctx, int rev_mode, uint32_t round, pgm_route_buffer *send) {
    if (!ctx || !send) return PGM_ERR_INVALID_ARG;
    const int rc = pgm_route_scan_launch(ctx, rev_mode, round);
    return rc ? rc : pgm_route_fetch(ctx, PGM_ROUTE_WINDOWS, send);
}

int pgm_route_fetch(pgm_ctx *ctx, int kind, pgm_route_buffer *send) {
    if (!ctx || !send || kind < 0 || kind > 2) return PGM_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    if (kind == PGM_ROUTE_PATTERNS)
        return route_fetch_counts(ctx, kind, send, ctx->bq_entries.p, (uint64_t)ctx->route.cap_pat * sizeof(uint4), sizeof(uint4), "pgm_route_fetch");
    if (kind == PGM_ROUTE_WINDOWS)
        return route_fetch_counts(ctx, kind, send, ctx->rt_win_send_().p, (uint64_t)ctx->route.cap_win * 12, 12, "pgm_route_fetch");
    return route_fetch_counts(ctx, kind, send, ctx->rt_cand_send_().p, (uint64_t)ctx->route.cap_cand * 12, 12, "pgm_route_fetch");
}

int pgm_route_scan_launch(pgm_ctx *ctx, int rev_mode, uint32_t round) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    const pgm_ctx::Route &rt = ctx->route;
    if (!rt.world || !ctx->phase_active) return fail(ctx, PGM_ERR_STATE, "pgm_route_scan: pgm_route_begin has not been called");
    if (!ctx->has_text) return fail(ctx, PGM_ERR_STATE, "pgm_route_scan: pgm_set_text has not been called");
    if (ctx->slice_begin != 0 || ctx->slice_len != ctx->pg_len)
        return fail(ctx, PGM_ERR_STATE, "pgm_route_scan: the routed scheme needs the whole text on every GPU (pgm_set_text, not a text shard)");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = finish_text_upload(ctx))) return rc;
    uint64_t b, e;
    route_range(ctx, rt.rank, round, b, e);
    ctx->route.cap_win = route_cap(route_round_size(ctx), rt.world);
    if ((rc = ensure(ctx, ctx->rt_win_send_(), (size_t)ctx->route.cap_win * rt.world * 12))) return rc;
    CU(cudaMemsetAsync(route_counts(ctx, PGM_ROUTE_WINDOWS), 0, RT_KIND_WORDS * sizeof(unsigned int), ctx->stream));
    if (e > b) {                                   // (reads the text only: may run before the table of this phase exists)
        pgm::RouteScanParams sp;
        memset(&sp, 0, sizeof sp);
        sp.tlo = (rev_mode ? ctx->r_lo : ctx->f_lo).as<uint32_t>() + PGM_PAD_WORDS;
        sp.thi = (rev_mode ? ctx->r_hi : ctx->f_hi).as<uint32_t>() + PGM_PAD_WORDS;
        sp.begin = b; sp.end = e;
        sp.first_word = (uint32_t)(b / 32);
        sp.n_tiles = (uint32_t)((e - b + PGM_TILE_POS - 1) / PGM_TILE_POS);
        sp.tail_mask = ctx->seed_len % 32 ? (1u << (ctx->seed_len % 32)) - 1u : 0xFFFFFFFFu;
        sp.tile_counter = route_tile_counter(ctx, PGM_ROUTE_WINDOWS);
        sp.q.entries = ctx->rt_win_send_().as<uint32_t>(); sp.q.count = route_counts(ctx, PGM_ROUTE_WINDOWS);
        sp.q.overflow = route_overflow(ctx, PGM_ROUTE_WINDOWS); sp.q.cap = ctx->route.cap_win; sp.q.world = (uint32_t)rt.world;
        const unsigned int grid = (unsigned int)std::min<uint64_t>(sp.n_tiles, (uint64_t)ctx->sm_count * 4);
        cudaError_t le = cudaSuccess;
        KLAUNCH(PGM_K_ROUTE_SCAN, "route_scan_kernel",
                switch ((ctx->seed_len + 31) / 32) {
                    case 1: le = launch_route_scan<1>(sp, grid, ctx->stream); break;
                    case 2: le = launch_route_scan<2>(sp, grid, ctx->stream); break;
                    case 3: le = launch_route_scan<3>(sp, grid, ctx->stream); break;
                    case 4: le = launch_route_scan<4>(sp, grid, ctx->stream); break;
                    case 5: le = launch_route_scan<5>(sp, grid, ctx->stream); break;
                    case 6: le = launch_route_scan<6>(sp, grid, ctx->stream); break;
                    case 7: le = launch_route_scan<7>(sp, grid, ctx->stream); break;
                    default: le = launch_route_scan<8>(sp, grid, ctx->stream); break;
                });
        if (le != cudaSuccess) return cuda_fail(ctx, le, "route_scan_kernel (shared memory attribute)");
    }
    ctx->state_fresh = false;
    return route_post_counts(ctx, PGM_ROUTE_WINDOWS);
}

int pgm_route_probe(pgm_ctx *ctx, int rev_mode, uint32_t round, const uint64_t *in_counts, pgm_route_buffer *send) {
    if (!ctx || !send || !in_counts) return PGM_ERR_INVALID_ARG;
    const int rc = pgm_route_probe_launch(ctx, rev_mode, round, in_counts);
    return rc ? rc : pgm_route_fetch(ctx, PGM_ROUTE_CANDIDATES, send);
}

int pgm_route_probe_launch(pgm_ctx *ctx, int rev_mode, uint32_t round, const uint64_t *in_counts) {
    if (!ctx || !in_counts) return PGM_ERR_INVALID_ARG;
    const pgm_ctx::Route &rt = ctx->route;
    if (!rt.world || !ctx->phase_active) return fail(ctx, PGM_ERR_STATE, "pgm_route_probe: pgm_route_begin has not been called");
    (void)rev_mode;
    CU(cudaSetDevice(ctx->device));
    { int wrc = route_wait_pull(ctx, PGM_ROUTE_WINDOWS); if (wrc) return wrc; }
    uint64_t n_in = 0;
    for (int s = 0; s < rt.world; s++) n_in += in_counts[s];
    if (n_in && ctx->rt_win_recv_().cap < n_in * 12) return fail(ctx, PGM_ERR_STATE, "pgm_route_probe: the windows have not been received (pgm_route_recv)");
    // on average well under one candidate per window; hot keys are covered by the slack
    // (a fixed capacity: the buffer, hence its address — peers may hold it open over IPC — never changes between rounds)
    ctx->route.cap_cand = (uint32_t)std::min<uint64_t>((std::max<uint64_t>(n_in, route_round_size(ctx)) + route_round_size(ctx) / 8) / rt.world + (4u << 20), 0xFFFFFFF0ull);
    int rc;
    if ((rc = ensure(ctx, ctx->rt_cand_send_(), (size_t)ctx->route.cap_cand * rt.world * 12))) return rc;
    CU(cudaMemsetAsync(route_counts(ctx, PGM_ROUTE_CANDIDATES), 0, RT_KIND_WORDS * sizeof(unsigned int), ctx->stream));
    // the received windows stream through the L2 (12 bytes per window, 48 x the text they came from) next to the filter;
    // PGM_ROUTE_PERSIST=1 puts a persisting access-policy window over the filter for this kernel (experiment knob: the L2
    // carve-out it needs stays set for the device and takes the same 64 MB away from every other kernel)
    static const bool persist = getenv("PGM_ROUTE_PERSIST") && atoi(getenv("PGM_ROUTE_PERSIST")) != 0;
    if (persist && (rc = filter_window(ctx, true, true))) return rc;
    // scratch for the windows that pass the filter (at most the largest sender segment) + their device counter per sender
    uint64_t n_max = 0;
    for (int s = 0; s < rt.world; s++) n_max = std::max(n_max, in_counts[s]);
    if ((rc = ensure(ctx, ctx->rt_live, std::max<uint64_t>(n_max, 1) * 12)) || (rc = ensure(ctx, ctx->rt_live_count, PGM_ROUTE_MAX_WORLD * sizeof(unsigned int)))) return rc;
    CU(cudaMemsetAsync(ctx->rt_live_count.p, 0, PGM_ROUTE_MAX_WORLD * sizeof(unsigned int), ctx->stream));
    uint64_t off = 0;
    for (int s = 0; s < rt.world; s++) {
        if (!in_counts[s]) continue;
        uint64_t b, e;
        route_range(ctx, s, round, b, e);
        pgm::RouteFilterParams fp;
        memset(&fp, 0, sizeof fp);
        fp.src = ctx->rt_win_recv_().as<uint32_t>() + off * 3;
        fp.n = in_counts[s];
        fp.tab = table_view(ctx);
        fp.live = ctx->rt_live.as<uint32_t>();
        fp.n_live = ctx->rt_live_count.as<unsigned int>() + s;
        {
            const uint64_t chunks = (fp.n + PGM_ROUTE_FILTER_CHUNK - 1) / PGM_ROUTE_FILTER_CHUNK;
            const unsigned int grid = (unsigned int)std::min<uint64_t>(chunks, (uint64_t)ctx->sm_count * 8);
            KLAUNCH(PGM_K_ROUTE_PROBE, "route_filter_kernel", pgm::route_filter_kernel<<<grid, PGM_ROUTE_THREADS, 0, ctx->stream>>>(fp));
        }
        pgm::RouteProbeParams pp;
        memset(&pp, 0, sizeof pp);
        pp.src = fp.live;
        pp.n = in_counts[s];
        pp.n_ptr = fp.n_live;
        pp.pos_base = b;
        pp.tab = table_view(ctx);
        pp.part_bits = ctx->part_bits; pp.world = (uint32_t)rt.world;
        for (int k = 0; k <= rt.world; k++) pp.read_begin[k] = rt.read_begin[k];
        pp.q.entries = ctx->rt_cand_send_().as<uint32_t>(); pp.q.count = route_counts(ctx, PGM_ROUTE_CANDIDATES);
        pp.q.overflow = route_overflow(ctx, PGM_ROUTE_CANDIDATES); pp.q.cap = ctx->route.cap_cand; pp.q.world = (uint32_t)rt.world;
        pp.counters = ctx->counters.as<unsigned long long>();
        {   // (the survivors' number is only known on the device: size the grid for a third of the segment, the kernel strides)
            const uint64_t chunks = (pp.n / 3 + PGM_ROUTE_PROBE_CHUNK2) / PGM_ROUTE_PROBE_CHUNK2;
            const unsigned int grid = (unsigned int)std::min<uint64_t>(chunks, (uint64_t)ctx->sm_count * PGM_ROUTE_PROBE_CTAS);
            KLAUNCH(PGM_K_ROUTE_PROBE, "route_probe_kernel", pgm::route_probe_kernel<<<grid, PGM_ROUTE_THREADS, 0, ctx->stream>>>(pp));
        }
        off += in_counts[s];
    }
    if (persist && (rc = filter_window(ctx, false, true))) return rc;
    if ((rc = route_mark_consumed(ctx, PGM_ROUTE_WINDOWS))) return rc;
    return route_post_counts(ctx, PGM_ROUTE_CANDIDATES);
}

int pgm_route_verify(pgm_ctx *ctx, int rev_mode, uint64_t n_in) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    const pgm_ctx::Route &rt = ctx->route;
    if (!rt.world || !ctx->phase_active) return fail(ctx, PGM_ERR_STATE, "pgm_route_verify: pgm_route_begin has not been called");
    if (n_in && ctx->rt_cand_recv_().cap < n_in * 12) return fail(ctx, PGM_ERR_STATE, "pgm_route_verify: the candidates have not been received (pgm_route_recv)");
    CU(cudaSetDevice(ctx->device));
    { int wrc = route_wait_pull(ctx, PGM_ROUTE_CANDIDATES); if (wrc) return wrc; }
    if (!n_in || !ctx->n_reads()) return PGM_OK;
    pgm::RouteVerifyParams vp;
    memset(&vp, 0, sizeof vp);
    vp.v.tlo = (rev_mode ? ctx->r_lo : ctx->f_lo).as<uint32_t>() + PGM_PAD_WORDS;
    vp.v.thi = (rev_mode ? ctx->r_hi : ctx->f_hi).as<uint32_t>() + PGM_PAD_WORDS;
    vp.v.pos_origin = 0; vp.v.bit_origin = 0; vp.v.pg_len = ctx->pg_len;
    vp.v.seed_len = ctx->shift_unit(); vp.v.parts = ctx->parts; vp.v.max_mm = ctx->max_mm; vp.v.min_mm = ctx->min_mm;
    vp.v.rev_mode = rev_mode ? 1 : 0;
    vp.v.tab = table_view(ctx); vp.v.reads = reads_view(ctx); vp.v.pr = per_read(ctx);
    vp.src = ctx->rt_cand_recv_().as<uint32_t>(); vp.n = n_in;
    vp.read_base = (uint32_t)rt.read_begin[rt.rank];
    vp.counters = ctx->counters.as<unsigned long long>();
    ctx->state_fresh = false; ctx->aux_clean = false; ctx->outputs_valid = false;
    const unsigned int grid = (unsigned int)std::min<uint64_t>(grid_for(n_in, PGM_VERIFY_THREADS), (uint64_t)ctx->sm_count * 8);
    KLAUNCH(PGM_K_ROUTE_VERIFY, "route_verify_kernel",
            if (ctx->lq_stride16 == 4) pgm::route_verify_kernel<true><<<grid, PGM_VERIFY_THREADS, 0, ctx->stream>>>(vp);
            else pgm::route_verify_kernel<false><<<grid, PGM_VERIFY_THREADS, 0, ctx->stream>>>(vp));
    return route_mark_consumed(ctx, PGM_ROUTE_CANDIDATES);
}

int pgm_map_reads(pgm_ctx *ctx, uint32_t match_prefix_length, uint32_t pre_seed, uint32_t seed,
                  uint32_t min_chars_per_mismatch, char pre_mode, char mode, int rev_compl,
                  uint64_t *out_pos, uint8_t *out_rc, uint8_t *out_mm, pgm_stats *stats) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (!ctx->has_text || !ctx->has_reads) return fail(ctx, PGM_ERR_STATE, "pgm_map_reads: set the text and the reads first");
    const uint32_t L = ctx->read_len;
    if (match_prefix_length != PGM_DISABLED_PREFIX_MODE && match_prefix_length < L)
        return fail(ctx, PGM_ERR_UNSUPPORTED, "pgm_map_reads: prefix matching (matchPrefixLength < readLength) is not used by pgrc-encoder and not supported");
    if (seed == 0 || min_chars_per_mismatch == 0) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_map_reads: seed and min_chars_per_mismatch must be > 0");
    auto known_mode = [](char c) { return std::tolower(c) == 'd' || std::tolower(c) == 'i' || std::tolower(c) == 'c'; };
    if (!known_mode(mode) || (pre_seed && !known_mode(pre_mode))) {
        // error convention of the reference: "Unknown matching mode" + exit (ReadsMatchers.cpp:737-739); here a status
        return fail(ctx, PGM_ERR_UNSUPPORTED, std::string("pgm_map_reads: unknown matching mode '") + mode + "' ('d'/'D', 'i'/'I', 'c'/'C')");
    }
    // ReadsMatchers.cpp:699-713
    const uint32_t max_mm = L / min_chars_per_mismatch;
    if (max_mm > 127) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_map_reads: min_chars_per_mismatch must be >= 2 (PgRC limit)");
    const uint32_t reads_exact = std::min(seed, L), pre_exact = std::min(pre_seed, L);
    uint32_t cur_exact = reads_exact;
    char cur_mode = mode;
    if (pre_exact > 0) { cur_exact = pre_exact; cur_mode = pre_mode; }
    const uint32_t cur_min = std::isupper((unsigned char)cur_mode) ? max_mm : 0;
    const uint32_t target_mm = L / cur_exact - 1;
    int rc;
    auto run_passes = [&](bool last_phase) -> int {
        int r;
        if ((r = pgm_scan_pass(ctx, 0)) || (r = resolve_impl(ctx, 0, last_phase && !rev_compl))) return r;
        if (rev_compl && ((r = pgm_scan_pass(ctx, 1)) || (r = resolve_impl(ctx, 1, last_phase)))) return r;
        return PGM_OK;
    };
    auto copmem_passes = [&]() -> int {                                               // CopMEMReadsApproxMatcher::executeMatching x 2
        int r;
        if ((r = pgm_copmem_pass(ctx, 0))) return r;
        if (rev_compl && (r = pgm_copmem_pass(ctx, 1))) return r;
        return PGM_OK;
    };
    if (std::tolower(cur_mode) == 'c') {                                              // :717-720 / :732-735 (also when readLength == seed)
        if ((rc = pgm_copmem_begin(ctx, cur_exact, max_mm, cur_min, 0)) || (rc = copmem_passes())) return rc;
    } else {
    if (L == cur_exact) rc = pgm_match_begin(ctx, L, 1, 0, 0, 0);                    // DefaultReadsExactMatcher (:718-722)
    else rc = match_begin_impl(ctx, cur_exact, target_mm + 1, max_mm, cur_min, 0,    // DefaultReadsApproxMatcher (:724-727) /
                               std::tolower(cur_mode) == 'i');                        // InterleavedReadsApproxMatcher (:728-731)
    if (rc || (rc = run_passes(pre_exact == 0))) return rc;
    }
    if (pre_exact > 0) {
        // second phase (:749-779); minMismatches comes from the FIRST phase's targetMismatches (:755)
        const uint32_t min2 = std::isupper((unsigned char)mode) ? max_mm : target_mm + 1;
        if (std::tolower(mode) == 'c') {
            if ((rc = pgm_copmem_begin(ctx, reads_exact, max_mm, min2, 1)) || (rc = copmem_passes())) return rc;
        } else
        if ((rc = match_begin_impl(ctx, reads_exact, L / reads_exact, max_mm, min2, 1, std::tolower(mode) == 'i')) || (rc = run_passes(true))) return rc;
    }
    return pgm_get_results(ctx, out_pos, out_rc, out_mm, stats);
}

} // extern "C"

#include "pgm_group.inl"


// ------------------------------------------------------------------------------------------------------------------
// Stage 7: exact matches between pseudogenomes (pgm_mem.cuh)
namespace {

// initParams + calcCoprimes (copmem/CopMEMMatcher.cpp:69-137) for targetMatchLength = L and a minimal matching length;
// false where the reference exits or would call a null entry of its hash-function table (:54-63)
bool mem_derive(uint32_t L, uint32_t min_len, uint64_t N, uint32_t &K, uint32_t &k1, uint32_t &k2, uint32_t &hash_size) {
    if (min_len > L) min_len = L;                            // constructor, :574-575
    if (min_len < 24 || L > 0xFFFF) return false;
    int k;
    if (L > 110) k = 56; else if (L > 62) k = 44; else if (L > 53) k = 40; else if (L > 46) k = 36;
    else if (L > 42) k = 32; else if (L > 32) k = 28; else k = ((int)L / 4 - 1) * 4;
    k = std::min(k, ((int)min_len / 4 - 1) * 4);
    if (!(k == 20 || k == 24 || k == 28 || k == 32 || k == 36 || k == 40 || k == 44 || k == 56)) return false;
    const int t = (int)L - k + 1;
    if (t <= 0) return false;
    int a, b;
    if (t >= 20) {
        a = 1; while ((a + 1) * (a + 1) <= t) a++;
        a += 1; b = a - 1;
        if (a * b > t) { --b; --a; }
    } else if (t >= 15) { a = 5; b = 3; } else if (t >= 12) { a = 4; b = 3; } else if (t >= 10) { a = 5; b = 2; }
    else if (t >= 6) { a = 3; b = 2; } else { a = t; b = 1; }
    K = (uint32_t)k; k1 = (uint32_t)a; k2 = (uint32_t)b;
    int i = 24;
    do { hash_size = 1u << (i++); } while (i <= 31 && hash_size < N / (uint64_t)a);
    return true;
}

} // namespace

extern "C" {

int pgm_mem_index(pgm_ctx *ctx, uint32_t target_match_length, uint32_t min_match_length, uint32_t *params) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (!ctx->has_text) return fail(ctx, PGM_ERR_STATE, "pgm_mem_index: set the text first");
    if (ctx->slice_begin != 0 || ctx->slice_len != ctx->pg_len)
        return fail(ctx, PGM_ERR_STATE, "pgm_mem_index: needs the whole text on this GPU (not a text shard)");
    const uint64_t N = ctx->pg_len;
    uint32_t K, k1, k2, hs;
    if (!mem_derive(target_match_length, min_match_length, N, K, k1, k2, hs))
        return fail(ctx, PGM_ERR_UNSUPPORTED, "pgm_mem_index: the reference exits for these lengths (minimal matching length below 24, "
                                              "L and K mismatch, or no hash function for K)");
    // With minMatchLength < targetMatchLength the reference's 4-byte guards (l1/l2/r1/r2, CopMEMMatcher.cpp:396-398) can reject
    // candidates that would extend far enough, and near the text ends they hold values left over from earlier candidates:
    // sequential state the order-free kernels do not carry.  PgRC never asks for it (pgrc-encoder.cpp:240-245 passes the target only).
    if (min_match_length < target_match_length)
        return fail(ctx, PGM_ERR_UNSUPPORTED, "pgm_mem_index: minimal matching length below the target match length is not covered "
                                              "(unreachable from pgrc-encoder.cpp:240-245)");
    if (N < target_match_length) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_mem_index: source text shorter than the target match length "
                                                                        "(SimplePgMatcher creates no matcher then, SimplePgMatcher.cpp:15)");
    if (N / k1 >= 0xFFFFFFF0ull) return fail(ctx, PGM_ERR_UNSUPPORTED, "pgm_mem_index: text too long for 32-bit sample indices");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = finish_text_upload(ctx))) return rc;
    pgm::CopmemParams cp;
    memset(&cp, 0, sizeof cp);
    cp.tlo = ctx->f_lo.as<uint32_t>() + PGM_PAD_WORDS;
    cp.thi = ctx->f_hi.as<uint32_t>() + PGM_PAD_WORDS;
    cp.pg_len = N;
    cp.K = K; cp.k1 = k1; cp.k2 = k2; cp.hash_mask = hs - 1;
    cp.n_sampled = (uint32_t)((N - K) / k1 + 1);
    const size_t hsz = hs;
    if ((rc = ensure(ctx, ctx->cm_count, (hsz + 1) * 4)) || (rc = ensure(ctx, ctx->cm_start, (hsz + 1) * 4)) ||
        (rc = ensure(ctx, ctx->cm_cumm, (hsz + 2) * 4)) || (rc = ensure(ctx, ctx->cm_fill, (size_t)cp.n_sampled * 4)) ||
        (rc = ensure(ctx, ctx->cm_hash, (size_t)cp.n_sampled * 4)) || (rc = ensure(ctx, ctx->cm_all, (size_t)cp.n_sampled * 4)) ||
        (rc = ensure(ctx, ctx->cm_entries, (size_t)cp.n_sampled * 4)) ||
        (rc = ensure(ctx, ctx->cm_nib, hsz / 8 * 4 + 64)) || (rc = ensure(ctx, ctx->cm_coarse, hsz / 64 * 4 + 64))) return rc;
    cp.nib = ctx->cm_nib.as<uint32_t>(); cp.coarse = ctx->cm_coarse.as<uint32_t>();
    cp.count = ctx->cm_count.as<uint32_t>(); cp.start_all = ctx->cm_start.as<uint32_t>(); cp.cumm = ctx->cm_cumm.as<uint32_t>();
    cp.sample_rank = ctx->cm_fill.as<uint32_t>(); cp.sample_hash = ctx->cm_hash.as<uint32_t>(); cp.all_entries = ctx->cm_all.as<uint32_t>();
    cp.entries = ctx->cm_entries.as<uint32_t>();
    ctx->copmem_active = false;
    CU(cudaMemsetAsync(cp.count, 0, (hsz + 1) * 4, ctx->stream));
    KLAUNCH(PGM_K_COPMEM_INDEX, "cm_hash_kernel", pgm::cm_hash_kernel<<<grid_for(cp.n_sampled, PGM_CM_THREADS), PGM_CM_THREADS, 0, ctx->stream>>>(cp));
    if ((rc = device_scan<0>(ctx, cp.count, (uint32_t)hsz, cp.start_all)) ||
        (rc = device_scan<PGM_CM_COLLISIONS_LIMIT + 1>(ctx, cp.count, (uint32_t)hsz, cp.cumm))) return rc;
    KLAUNCH(PGM_K_COPMEM_INDEX, "cm_compact_kernel", pgm::cm_compact_kernel<<<grid_for(hsz / 8, PGM_CM_THREADS), PGM_CM_THREADS, 0, ctx->stream>>>(cp));
    KLAUNCH(PGM_K_COPMEM_INDEX, "cm_scatter_kernel", pgm::cm_scatter_kernel<<<grid_for(cp.n_sampled, PGM_CM_THREADS), PGM_CM_THREADS, 0, ctx->stream>>>(cp));
    KLAUNCH(PGM_K_COPMEM_INDEX, "cm_select_kernel", pgm::cm_select_kernel<<<grid_for(hsz, PGM_CM_THREADS), PGM_CM_THREADS, 0, ctx->stream>>>(cp));
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, ctx->err_flag.p, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (bad) return fail(ctx, PGM_ERR_BAD_SYMBOL, "pgm_mem_index: the source text contains a symbol outside ACGT");
    ctx->mem_L = target_match_length; ctx->mem_K = K; ctx->mem_k1 = k1; ctx->mem_k2 = k2; ctx->mem_hash_size = hs;
    ctx->mem_index_valid = true;
    ctx->mem_result_valid = false;
    if (params) { params[0] = K; params[1] = k1; params[2] = k2; params[3] = hs; }
    return PGM_OK;
}

} // extern "C"

namespace {
// pgm_mem_match for the groups of 256 query positions [G * part / n_parts, G * (part + 1) / n_parts) of the G groups; raw: stop
// before the "covered by the previous match" test (a context of several: the caller applies it across the seams) and leave
// ctx->mem_raw / mem_rawq with ctx->mem_count entries
int mem_match_impl(pgm_ctx *ctx, const char *dest, uint64_t dest_len, int dest_is_src, int rev_compl,
                   uint32_t min_match_length, uint64_t *count, int part, int n_parts, bool raw) {
    if (!ctx || !count) return PGM_ERR_INVALID_ARG;
    if (!ctx->mem_index_valid) return fail(ctx, PGM_ERR_STATE, "pgm_mem_match: pgm_mem_index has not been called for the current text");
    if (!dest && !dest_is_src) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_mem_match: destination text is null");
    const uint64_t N = ctx->pg_len;
    if (!dest) dest_len = N;
    if (dest_is_src && dest_len != N) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_mem_match: dest_is_src with a destination of another length");
    if (min_match_length > ctx->mem_L) min_match_length = ctx->mem_L;
    if (min_match_length < ctx->mem_L)
        return fail(ctx, PGM_ERR_UNSUPPORTED, "pgm_mem_match: minimal matching length below the target match length is not covered (see pgm_mem_index)");
    CU(cudaSetDevice(ctx->device));
    int rc;
    ctx->mem_result_valid = false;
    pgm::MemParams mp;
    memset(&mp, 0, sizeof mp);
    mp.slo = ctx->f_lo.as<uint32_t>() + PGM_PAD_WORDS; mp.shi = ctx->f_hi.as<uint32_t>() + PGM_PAD_WORDS;
    mp.N = N; mp.N2 = dest_len;
    mp.K = ctx->mem_K; mp.k1 = ctx->mem_k1; mp.k2 = ctx->mem_k2; mp.hash_mask = ctx->mem_hash_size - 1;
    mp.min_len = min_match_length; mp.skip = mp.K / mp.k1 - 1;
    mp.dest_is_src = dest_is_src ? 1 : 0; mp.rev_compl = rev_compl ? 1 : 0;
    mp.nib = ctx->cm_nib.as<uint32_t>(); mp.coarse = ctx->cm_coarse.as<uint32_t>(); mp.entries = ctx->cm_entries.as<uint32_t>();
    if (!dest) {
        mp.dlo = (rev_compl ? ctx->r_lo : ctx->f_lo).as<uint32_t>() + PGM_PAD_WORDS;
        mp.dhi = (rev_compl ? ctx->r_hi : ctx->f_hi).as<uint32_t>() + PGM_PAD_WORDS;
        mp.dinv = nullptr;
    } else {
        const size_t pb = plane_bytes(dest_len);
        if ((rc = ensure(ctx, ctx->mem_dlo, pb)) || (rc = ensure(ctx, ctx->mem_dhi, pb)) || (rc = ensure(ctx, ctx->mem_dinv, pb))) return rc;
        CU(cudaMemsetAsync(ctx->mem_dlo.p, 0, pb, ctx->stream));
        CU(cudaMemsetAsync(ctx->mem_dhi.p, 0, pb, ctx->stream));
        CU(cudaMemsetAsync(ctx->mem_dinv.p, 0, pb, ctx->stream));
        const uint8_t *src = reinterpret_cast<const uint8_t *>(dest);
        if (dest_len && !is_device_ptr(dest)) {
            if ((rc = ensure(ctx, ctx->mem_stage, dest_len))) return rc;
            if (dest_len < (8u << 20) || !is_pageable_host(dest)) {
                CU(cudaMemcpyAsync(ctx->mem_stage.p, dest, dest_len, cudaMemcpyHostToDevice, ctx->stream));
            } else {
                // a pageable destination text (PgRC's std::string): staged by host threads through the pinned slots (h2d_chunk)
                if ((rc = fence_copy_stream(ctx))) return rc;
                uint64_t c = 0;
                for (uint64_t off = 0; off < dest_len; off += TEXT_CHUNK_BASES, c++)
                    if ((rc = h2d_chunk(ctx, ctx->mem_stage.as<uint8_t>() + off, dest + off, std::min<uint64_t>(TEXT_CHUNK_BASES, dest_len - off),
                                        chunk_event(ctx->mem_ev, c), true))) return rc;
                CU(cudaStreamWaitEvent(ctx->stream, ctx->mem_ev[c - 1], 0));
            }
            src = ctx->mem_stage.as<uint8_t>();
        }
        uint32_t *dlo = ctx->mem_dlo.as<uint32_t>() + PGM_PAD_WORDS, *dhi = ctx->mem_dhi.as<uint32_t>() + PGM_PAD_WORDS,
                 *dinv = ctx->mem_dinv.as<uint32_t>() + PGM_PAD_WORDS;
        if (dest_len)
            KLAUNCH(PGM_K_MEM_PACK, "mem_pack_kernel", pgm::mem_pack_kernel<<<grid_for((dest_len + 31) / 32, PGM_MEM_THREADS), PGM_MEM_THREADS, 0, ctx->stream>>>(
                src, dest_len, dlo, dhi, dinv));
        mp.dlo = dlo; mp.dhi = dhi; mp.dinv = dinv;
    }
    ctx->mem_count = 0;
    *count = 0;
    if (dest_len < mp.K) { ctx->mem_result_valid = true; return PGM_OK; }
    {
        // groups of 256 query positions: the main loop's (i1 = g * k2 * 256 while i1 + K + k2 * 256 < N2 + 1, CopMEMMatcher.cpp:364)
        // and the tail with the 1 .. 256 positions left
        const uint64_t nq_total = (dest_len - mp.K) / mp.k2 + 1, groups = (nq_total + PGM_MEM_GROUP - 1) / PGM_MEM_GROUP;
        const uint64_t g0 = groups * (uint64_t)part / (uint64_t)n_parts, g1 = groups * (uint64_t)(part + 1) / (uint64_t)n_parts;
        mp.q0 = g0 * PGM_MEM_GROUP;
        mp.nq = std::min(nq_total, g1 * PGM_MEM_GROUP) - mp.q0;
        if (g1 == g0) { ctx->mem_result_valid = true; return PGM_OK; }
        mp.n_groups = g1 - g0 - 1;
    }
    const size_t mask_words = (size_t)((mp.nq + 31) / 32 + 8);
    if ((rc = ensure(ctx, ctx->mem_fv, (size_t)mp.nq * 4)) || (rc = ensure(ctx, ctx->mem_has, mask_words * 4)) ||
        (rc = ensure(ctx, ctx->mem_emit, mask_words * 4)) || (rc = ensure(ctx, ctx->mem_gcount, (size_t)(mp.n_groups + 2) * 4)) ||
        (rc = ensure(ctx, ctx->mem_gstart, (size_t)(mp.n_groups + 2) * 4))) return rc;
    mp.fv = ctx->mem_fv.as<uint32_t>(); mp.has_fv = ctx->mem_has.as<uint32_t>(); mp.emit = ctx->mem_emit.as<uint32_t>();
    mp.group_count = ctx->mem_gcount.as<uint32_t>(); mp.group_start = ctx->mem_gstart.as<uint32_t>();
    CU(cudaMemsetAsync(mp.has_fv, 0, mask_words * 4, ctx->stream));
    CU(cudaMemsetAsync(mp.emit, 0, mask_words * 4, ctx->stream));
    KLAUNCH(PGM_K_MEM_QUERY, "mem_query_kernel", pgm::mem_query_kernel<<<grid_for(mp.nq, PGM_MEM_THREADS), PGM_MEM_THREADS, 0, ctx->stream>>>(mp));
    KLAUNCH(PGM_K_MEM_EMIT, "mem_walk_kernel", pgm::mem_walk_kernel<<<grid_for(mp.n_groups + 1, PGM_MEM_THREADS), PGM_MEM_THREADS, 0, ctx->stream>>>(mp));
    if ((rc = device_scan<0>(ctx, mp.group_count, (uint32_t)(mp.n_groups + 1), ctx->mem_gstart.as<uint32_t>()))) return rc;
    uint32_t n_raw = 0;
    CU(cudaMemcpyAsync(&n_raw, ctx->mem_gstart.as<uint32_t>() + mp.n_groups + 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    mp.n_raw = n_raw;
    if (n_raw) {
        if ((rc = ensure(ctx, ctx->mem_raw, (size_t)n_raw * sizeof(pgm::MemMatch))) || (rc = ensure(ctx, ctx->mem_rawq, (size_t)n_raw * 8)) ||
            (rc = ensure(ctx, ctx->mem_keep, (size_t)n_raw * 4)) || (rc = ensure(ctx, ctx->mem_kstart, ((size_t)n_raw + 1) * 4)) ||
            (rc = ensure(ctx, ctx->mem_out, (size_t)n_raw * sizeof(pgm::MemMatch)))) return rc;
        mp.raw = ctx->mem_raw.as<pgm::MemMatch>(); mp.raw_q = ctx->mem_rawq.as<uint64_t>(); mp.keep = ctx->mem_keep.as<uint32_t>();
        mp.keep_start = ctx->mem_kstart.as<uint32_t>(); mp.out = ctx->mem_out.as<pgm::MemMatch>();
        KLAUNCH(PGM_K_MEM_EMIT, "mem_extend_kernel", pgm::mem_extend_kernel<<<grid_for((mp.nq + 31) / 32, PGM_MEM_THREADS), PGM_MEM_THREADS, 0, ctx->stream>>>(mp));
        if (raw) {
            CU(cudaStreamSynchronize(ctx->stream));
            ctx->mem_count = n_raw;
            *count = n_raw;
            ctx->mem_result_valid = true;
            return PGM_OK;
        }
        KLAUNCH(PGM_K_MEM_EMIT, "mem_flag_kernel", pgm::mem_flag_kernel<<<grid_for(n_raw, PGM_MEM_THREADS), PGM_MEM_THREADS, 0, ctx->stream>>>(mp));
        if ((rc = device_scan<0>(ctx, mp.keep, n_raw, ctx->mem_kstart.as<uint32_t>()))) return rc;
        KLAUNCH(PGM_K_MEM_EMIT, "mem_compact_kernel", pgm::mem_compact_kernel<<<grid_for(n_raw, PGM_MEM_THREADS), PGM_MEM_THREADS, 0, ctx->stream>>>(mp));
        uint32_t n_out = 0;
        CU(cudaMemcpyAsync(&n_out, ctx->mem_kstart.as<uint32_t>() + n_raw, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->mem_count = n_out;
    }
    *count = ctx->mem_count;
    ctx->mem_result_valid = true;
    return PGM_OK;
}
} // namespace

extern "C" {

int pgm_mem_match(pgm_ctx *ctx, const char *dest, uint64_t dest_len, int dest_is_src, int rev_compl,
                  uint32_t min_match_length, uint64_t *count) {
    if (ctx) ctx->mem_share_valid = false;
    return mem_match_impl(ctx, dest, dest_len, dest_is_src, rev_compl, min_match_length, count, 0, 1, false);
}

int pgm_mem_match_share(pgm_ctx *ctx, const char *dest, uint64_t dest_len, int dest_is_src, int rev_compl,
                        uint32_t min_match_length, int part, int n_parts, uint64_t *count) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (n_parts < 1 || part < 0 || part >= n_parts) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_mem_match_share: need 0 <= part < n_parts");
    ctx->mem_share_valid = false;
    const int rc = mem_match_impl(ctx, dest, dest_len, dest_is_src, rev_compl, min_match_length, count, part, n_parts, true);
    if (rc == PGM_OK) { ctx->mem_share_valid = true; ctx->mem_result_valid = false; }
    return rc;
}

int pgm_mem_get_share(pgm_ctx *ctx, pgm_text_match *out_matches, uint64_t *out_query_pos, uint64_t capacity) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (!ctx->mem_share_valid) return fail(ctx, PGM_ERR_STATE, "pgm_mem_get_share: no result (call pgm_mem_match_share)");
    if (capacity < ctx->mem_count || (ctx->mem_count && (!out_matches || !out_query_pos)))
        return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_mem_get_share: capacity below the share's size");
    if (!ctx->mem_count) return PGM_OK;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(out_matches, ctx->mem_raw.p, (size_t)ctx->mem_count * sizeof(pgm_text_match), cudaMemcpyDefault, ctx->stream));
    CU(cudaMemcpyAsync(out_query_pos, ctx->mem_rawq.p, (size_t)ctx->mem_count * 8, cudaMemcpyDefault, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PGM_OK;
}

int pgm_mem_get_matches(pgm_ctx *ctx, pgm_text_match *out, uint64_t capacity) {
    if (!ctx) return PGM_ERR_INVALID_ARG;
    if (!ctx->mem_result_valid) return fail(ctx, PGM_ERR_STATE, "pgm_mem_get_matches: no result (call pgm_mem_match)");
    if (capacity < ctx->mem_count || (ctx->mem_count && !out)) return fail(ctx, PGM_ERR_INVALID_ARG, "pgm_mem_get_matches: capacity below the match count");
    if (!ctx->mem_count) return PGM_OK;
    CU(cudaSetDevice(ctx->device));
    static_assert(sizeof(pgm_text_match) == sizeof(pgm::MemMatch), "layout");
    CU(cudaMemcpyAsync(out, ctx->mem_out.p, (size_t)ctx->mem_count * sizeof(pgm_text_match), cudaMemcpyDefault, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PGM_OK;
}

} // extern "C"


// ------------------------------------------------------------------------------------------------------------------
// Stage 7 on a group of contexts: every GPU holds the source text and its index; the groups of 256 query positions of a
// destination text are independent (pgm_mem.cuh), so GPU r takes the r-th share of them and the shares are concatenated; the
// "covered by the previous match" test (CopMEMMatcher.cpp:388-393) runs over the concatenation, across the seams, on the host.
extern "C" {

int pgm_group_mem_index(pgm_group *g, uint32_t target_match_length, uint32_t min_match_length, uint32_t *params) {
    if (!g) return PGM_ERR_INVALID_ARG;
    g->err.clear();
    g->mem_valid = false;
    std::vector<std::array<uint32_t, 4>> par(g->size());
    const int rc = group_run(g, [&](int r) -> int { return pgm_mem_index(g->ctx[r], target_match_length, min_match_length, par[r].data()); });
    if (rc == PGM_OK && params) for (int k = 0; k < 4; k++) params[k] = par[0][k];
    return rc;
}

int pgm_group_mem_match(pgm_group *g, const char *dest, uint64_t dest_len, int dest_is_src, int rev_compl,
                        uint32_t min_match_length, uint64_t *count) {
    if (!g || !count) return PGM_ERR_INVALID_ARG;
    g->err.clear();
    g->mem_valid = false;
    const int n = g->size();
    if (n == 1) {
        const int rc = pgm_mem_match(g->ctx[0], dest, dest_len, dest_is_src, rev_compl, min_match_length, count);
        if (rc != PGM_OK) return gfail(g, rc, pgm_last_error(g->ctx[0]));
        g->mem_out.assign(*count, pgm_text_match{0, 0, 0});
        const int rc2 = pgm_mem_get_matches(g->ctx[0], g->mem_out.data(), *count);
        if (rc2 != PGM_OK) return gfail(g, rc2, pgm_last_error(g->ctx[0]));
        g->mem_valid = true;
        return PGM_OK;
    }
    std::vector<uint64_t> cnt(n, 0);
    std::vector<std::vector<pgm::MemMatch>> raw(n);
    std::vector<std::vector<uint64_t>> raw_q(n);
    const int rc = group_run(g, [&](int r) -> int {
        pgm_ctx *c = g->ctx[r];
        int rr = mem_match_impl(c, dest, dest_len, dest_is_src, rev_compl, min_match_length, &cnt[r], r, n, true);
        if (rr != PGM_OK || !cnt[r]) return rr;
        raw[r].resize(cnt[r]); raw_q[r].resize(cnt[r]);
        cudaSetDevice(c->device);
        cudaError_t e = cudaMemcpyAsync(raw[r].data(), c->mem_raw.p, cnt[r] * sizeof(pgm::MemMatch), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(raw_q[r].data(), c->mem_rawq.p, cnt[r] * 8, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        return e == cudaSuccess ? PGM_OK : cuda_fail(c, e, "pgm_group_mem_match (copy of the raw matches)");
    });
    if (rc != PGM_OK) return rc;
    // the shares in query order; a position is pushed unless its match is the previous visited position's match
    g->mem_out.clear();
    const uint32_t K = g->ctx[0]->mem_K;
    bool have_prev = false;
    pgm::MemMatch prev{0, 0, 0};
    for (int r = 0; r < n; r++)
        for (uint64_t i = 0; i < cnt[r]; i++) {
            const pgm::MemMatch &m = raw[r][i];
            if (!(have_prev && m.dest - m.src == prev.dest - prev.src && raw_q[r][i] + K < prev.dest + prev.len))
                g->mem_out.push_back(pgm_text_match{m.src, m.len, m.dest});
            prev = m; have_prev = true;
        }
    *count = g->mem_out.size();
    g->mem_valid = true;
    return PGM_OK;
}

int pgm_group_mem_get_matches(pgm_group *g, pgm_text_match *out, uint64_t capacity) {
    if (!g) return PGM_ERR_INVALID_ARG;
    if (!g->mem_valid) return gfail(g, PGM_ERR_STATE, "pgm_group_mem_get_matches: no result (call pgm_group_mem_match)");
    if (capacity < g->mem_out.size() || (!g->mem_out.empty() && !out)) return gfail(g, PGM_ERR_INVALID_ARG, "pgm_group_mem_get_matches: capacity below the match count");
    if (g->mem_out.empty()) return PGM_OK;
    const cudaError_t e = cudaMemcpy(out, g->mem_out.data(), g->mem_out.size() * sizeof(pgm_text_match), cudaMemcpyDefault);
    if (e != cudaSuccess) return gfail(g, PGM_ERR_CUDA, std::string("pgm_group_mem_get_matches: ") + cudaGetErrorString(e));
    return PGM_OK;
}

} // extern "C"
