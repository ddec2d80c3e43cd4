// pgm_group.inl — several GPUs behind one handle (included at the end of pgm_api.cu; see include/pgrc_gpu_matcher.h).
//
// One host thread per context for the duration of a call; the threads run the same step sequence as
// pgrc_b200/matcher.py:run_plan_routed and meet at a barrier around every exchange.  An exchange is pulled by the
// receiver: once every sender's counts are known, rank r asks its context for a receive buffer and copies segment r of
// every peer's send buffer into it with cudaMemcpyPeerAsync on its own stream (NVLink when peer access is enabled).
#include <condition_variable>
#include <mutex>
#include <thread>

struct pgm_group {
    std::vector<pgm_ctx *> ctx;
    std::vector<int> devices;
    std::string err;
    // inputs (global read order: LQ reads, then N reads)
    uint32_t n_lq = 0, n_n = 0, read_len = 0;
    std::vector<uint64_t> read_begin;           // n + 1
    // current phase
    bool routed = false;                        // several contexts and contiguous seeds: the routed scheme
    bool copmem = false;
    uint32_t rounds = 1;
    // barrier + error flag of a threaded call
    std::mutex mu;
    std::condition_variable cv;
    int waiting = 0;
    uint64_t generation = 0;
    bool failed = false;
    std::vector<pgm_route_buffer> send;         // per rank: the buffer of the exchange in flight
    // stage 7 (pgm_group_mem_*): the matches of the last pgm_group_mem_match, in the reference's push order
    std::vector<pgm_text_match> mem_out;
    bool mem_valid = false;

    int size() const { return (int)ctx.size(); }
};

namespace {

thread_local std::string g_group_create_error;

int gfail(pgm_group *g, int code, const std::string &msg) {
    if (g) { std::lock_guard<std::mutex> lk(g->mu); if (g->err.empty() || !g->failed) g->err = msg; g->failed = true; }
    else g_group_create_error = msg;
    return code;
}

// all ranks arrive; returns false when some rank has failed (every rank then leaves the call at the same point)
bool group_sync(pgm_group *g) {
    std::unique_lock<std::mutex> lk(g->mu);
    const uint64_t gen = g->generation;
    if (++g->waiting == g->size()) {
        g->waiting = 0;
        g->generation++;
        g->cv.notify_all();
    } else {
        g->cv.wait(lk, [&] { return g->generation != gen; });
    }
    return !g->failed;
}

// runs body(rank) on one thread per context; body returns a pgm_status
template <class F>
int group_run(pgm_group *g, F body) {
    { std::lock_guard<std::mutex> lk(g->mu); g->failed = false; g->waiting = 0; }
    std::vector<int> rcs(g->size(), PGM_OK);
    if (g->size() == 1) {
        rcs[0] = body(0);
    } else {
        std::vector<std::thread> th;
        for (int r = 0; r < g->size(); r++) th.emplace_back([&, r] { rcs[r] = body(r); });
        for (auto &t : th) t.join();
    }
    for (int r = 0; r < g->size(); r++)
        if (rcs[r] != PGM_OK) {
            std::lock_guard<std::mutex> lk(g->mu);
            if (g->err.empty()) g->err = std::string("GPU ") + std::to_string(g->devices[r]) + ": " + pgm_last_error(g->ctx[r]);
            return rcs[r];
        }
    return PGM_OK;
}

// a context's own failure inside a threaded call: record it, keep walking to the next barrier
int rank_fail(pgm_group *g, int r, int rc) {
    std::lock_guard<std::mutex> lk(g->mu);
    if (!g->failed) g->err = std::string("GPU ") + std::to_string(g->devices[r]) + ": " + pgm_last_error(g->ctx[r]);
    g->failed = true;
    return rc;
}

// One exchange, seen from rank r: `mine` is what r's emit step produced.  Returns the entries received per sender.
int group_exchange(pgm_group *g, int r, int kind, const pgm_route_buffer &mine, int rc_in, std::vector<uint64_t> &in_counts) {
    const int n = g->size();
    g->send[r] = mine;
    if (rc_in != PGM_OK) rank_fail(g, r, rc_in);
    if (!group_sync(g)) return rc_in != PGM_OK ? rc_in : PGM_ERR_STATE;         // every sender's counts are visible
    in_counts.assign(n, 0);
    uint64_t total = 0;
    for (int s = 0; s < n; s++) { in_counts[s] = g->send[s].count[r]; total += in_counts[s]; }
    pgm_ctx *c = g->ctx[r];
    void *dst = nullptr;
    int rc = pgm_route_recv(c, kind, total, &dst);
    if (rc == PGM_OK) {
        cudaSetDevice(c->device);
        for (int k = 0; k < n && rc == PGM_OK; k++) {
            const int s = (r + k) % n;                                          // start with the own segment, then round-robin
            const uint64_t bytes = in_counts[s] * g->send[s].entry_bytes;
            uint64_t at = 0;
            for (int q = 0; q < s; q++) at += in_counts[q] * g->send[q].entry_bytes;
            if (!bytes) continue;
            const char *src = static_cast<const char *>(g->send[s].base) + (uint64_t)r * g->send[s].stride_bytes;
            const cudaError_t e = cudaMemcpyPeerAsync(static_cast<char *>(dst) + at, c->device, src, g->ctx[s]->device, bytes, c->stream);
            if (e != cudaSuccess) rc = cuda_fail(c, e, "cudaMemcpyPeerAsync (exchange)");
        }
        if (rc == PGM_OK) {
            const cudaError_t e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) rc = cuda_fail(c, e, "cudaStreamSynchronize (exchange)");
        }
    }
    if (rc != PGM_OK) rank_fail(g, r, rc);
    if (!group_sync(g)) return rc != PGM_OK ? rc : PGM_ERR_STATE;               // every receiver has pulled: send buffers are free again
    return PGM_OK;
}

} // namespace

extern "C" {

int pgm_group_create(int n_devices, const int *devices, pgm_group **out) {
    if (!out) return gfail(nullptr, PGM_ERR_INVALID_ARG, "pgm_group_create: out is null");
    *out = nullptr;
    if (n_devices < 1 || n_devices > PGM_ROUTE_MAX_WORLD) return gfail(nullptr, PGM_ERR_INVALID_ARG, "pgm_group_create: need 1..16 devices");
    pgm_group *g = new pgm_group();
    for (int k = 0; k < n_devices; k++) {
        const int dev = devices ? devices[k] : k;
        pgm_ctx *c = nullptr;
        const int rc = pgm_create(dev, &c);
        if (rc != PGM_OK) {
            g_group_create_error = std::string("pgm_group_create: device ") + std::to_string(dev) + ": " + pgm_last_error(nullptr);
            pgm_group_destroy(g);
            return rc;
        }
        g->ctx.push_back(c);
        g->devices.push_back(dev);
    }
    // direct NVLink copies between the GPUs of the group (ignored where peers are the same device or already enabled)
    for (int a = 0; a < n_devices; a++)
        for (int b = 0; b < n_devices; b++) {
            if (g->devices[a] == g->devices[b]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, g->devices[a], g->devices[b]) == cudaSuccess && can) {
                cudaSetDevice(g->devices[a]);
                cudaDeviceEnablePeerAccess(g->devices[b], 0);
                cudaGetLastError();
            }
        }
    g->send.resize(n_devices);
    *out = g;
    return PGM_OK;
}

void pgm_group_destroy(pgm_group *g) {
    if (!g) return;
    for (pgm_ctx *c : g->ctx) pgm_destroy(c);
    delete g;
}

const char *pgm_group_last_error(const pgm_group *g) { return g ? g->err.c_str() : g_group_create_error.c_str(); }

int pgm_group_size(const pgm_group *g) { return g ? g->size() : 0; }

int pgm_group_set_text(pgm_group *g, const char *text, uint64_t pg_len) {
    if (!g) return PGM_ERR_INVALID_ARG;
    g->err.clear();
    return group_run(g, [&](int r) -> int { return pgm_set_text(g->ctx[r], text, pg_len); });
}

int pgm_group_set_reads(pgm_group *g, const uint8_t *lq, uint32_t n_lq, const uint8_t *nn, uint32_t n_n, uint32_t read_len) {
    if (!g) return PGM_ERR_INVALID_ARG;
    g->err.clear();
    if (read_len == 0 || read_len > 255) return gfail(g, PGM_ERR_INVALID_ARG, "pgm_group_set_reads: read_len must be 1..255 (PgRC limit)");
    const int n = g->size();
    const uint64_t total = (uint64_t)n_lq + n_n;
    g->n_lq = n_lq; g->n_n = n_n; g->read_len = read_len;
    g->read_begin.assign(n + 1, 0);
    for (int k = 0; k <= n; k++) g->read_begin[k] = total * (uint64_t)k / (uint64_t)n;
    const uint32_t lq_plen = (read_len + 3) / 4, n_plen = (read_len + 2) / 3;
    return group_run(g, [&](int r) -> int {
        // rank r's range of the global read order, as an LQ slice and an N slice
        const uint64_t lo = g->read_begin[r], hi = g->read_begin[r + 1];
        const uint64_t l0 = std::min<uint64_t>(lo, n_lq), l1 = std::min<uint64_t>(hi, n_lq);
        const uint64_t m0 = std::max<uint64_t>(lo, n_lq) - n_lq, m1 = std::max<uint64_t>(hi, n_lq) - n_lq;
        int rc = pgm_set_reads(g->ctx[r], l1 > l0 ? lq + l0 * lq_plen : nullptr, (uint32_t)(l1 - l0), m1 > m0 ? nn + m0 * n_plen : nullptr,
                               (uint32_t)(m1 - m0), read_len);
        if (rc == PGM_OK && n > 1) rc = pgm_route_config(g->ctx[r], r, n, g->read_begin.data(), 0);
        return rc;
    });
}

int pgm_group_upload(pgm_group *g) {
    if (!g) return PGM_ERR_INVALID_ARG;
    g->err.clear();
    return group_run(g, [&](int r) -> int { return pgm_upload(g->ctx[r]); });
}

int pgm_group_match_begin(pgm_group *g, uint32_t seed_len, uint32_t parts, uint32_t max_mm, uint32_t min_mm, int continuation, int interleaved) {
    if (!g) return PGM_ERR_INVALID_ARG;
    g->err.clear();
    g->copmem = false;
    if (parts == 1) interleaved = 0;
    g->routed = g->size() > 1 && !interleaved;
    if (!g->routed)
        return group_run(g, [&](int r) -> int {
            return interleaved ? pgm_match_begin_interleaved(g->ctx[r], seed_len, parts, max_mm, min_mm, continuation)
                               : pgm_match_begin(g->ctx[r], seed_len, parts, max_mm, min_mm, continuation);
        });
    return group_run(g, [&](int r) -> int {
        pgm_route_buffer mine;
        memset(&mine, 0, sizeof mine);
        int rc = pgm_route_begin(g->ctx[r], seed_len, parts, max_mm, min_mm, continuation, &mine);
        std::vector<uint64_t> in_counts;
        if ((rc = group_exchange(g, r, PGM_ROUTE_PATTERNS, mine, rc, in_counts)) != PGM_OK) return rc;
        uint64_t n_in = 0;
        for (uint64_t c : in_counts) n_in += c;
        if ((rc = pgm_route_build(g->ctx[r], n_in)) != PGM_OK) return rc;
        if (r == 0) {
            uint32_t rounds = 1;
            if (g->ctx[r]->has_text && (rc = pgm_route_rounds(g->ctx[r], &rounds)) != PGM_OK) return rc;
            g->rounds = rounds;
        }
        return PGM_OK;
    });
}

int pgm_group_pass(pgm_group *g, int rev_mode) {
    if (!g) return PGM_ERR_INVALID_ARG;
    g->err.clear();
    if (!g->routed)
        return group_run(g, [&](int r) -> int {
            int rc = pgm_scan_pass(g->ctx[r], rev_mode);
            return rc != PGM_OK ? rc : pgm_resolve_pass(g->ctx[r], rev_mode);
        });
    uint32_t rounds = 1;
    {   // (the text may have been set after pgm_group_match_begin)
        const int rc = pgm_route_rounds(g->ctx[0], &rounds);
        if (rc != PGM_OK) return gfail(g, rc, pgm_last_error(g->ctx[0]));
    }
    return group_run(g, [&](int r) -> int {
        pgm_ctx *c = g->ctx[r];
        for (uint32_t rnd = 0; rnd < rounds; rnd++) {
            pgm_route_buffer mine;
            memset(&mine, 0, sizeof mine);
            std::vector<uint64_t> win_in, cand_in;
            int rc = pgm_route_scan(c, rev_mode, rnd, &mine);
            if ((rc = group_exchange(g, r, PGM_ROUTE_WINDOWS, mine, rc, win_in)) != PGM_OK) return rc;
            memset(&mine, 0, sizeof mine);
            rc = pgm_route_probe(c, rev_mode, rnd, win_in.data(), &mine);
            if ((rc = group_exchange(g, r, PGM_ROUTE_CANDIDATES, mine, rc, cand_in)) != PGM_OK) return rc;
            uint64_t n_in = 0;
            for (uint64_t x : cand_in) n_in += x;
            rc = pgm_route_verify(c, rev_mode, n_in);
            if (rc != PGM_OK) rank_fail(g, r, rc);
            if (!group_sync(g)) return rc != PGM_OK ? rc : PGM_ERR_STATE;
        }
        return pgm_resolve_pass(c, rev_mode);
    });
}

int pgm_group_copmem_begin(pgm_group *g, uint32_t part_len, uint32_t max_mm, uint32_t min_mm, int continuation) {
    if (!g) return PGM_ERR_INVALID_ARG;
    g->err.clear();
    g->copmem = true; g->routed = false;
    return group_run(g, [&](int r) -> int { return pgm_copmem_begin(g->ctx[r], part_len, max_mm, min_mm, continuation); });
}

int pgm_group_copmem_pass(pgm_group *g, int rev_mode) {
    if (!g) return PGM_ERR_INVALID_ARG;
    g->err.clear();
    return group_run(g, [&](int r) -> int { return pgm_copmem_pass(g->ctx[r], rev_mode); });
}

int pgm_group_get_results(pgm_group *g, uint64_t *out_pos, uint8_t *out_rc, uint8_t *out_mm, pgm_stats *stats) {
    if (!g) return PGM_ERR_INVALID_ARG;
    g->err.clear();
    std::vector<pgm_stats> st(g->size());
    const int rc = group_run(g, [&](int r) -> int {
        const uint64_t lo = g->read_begin.empty() ? 0 : g->read_begin[r];
        return pgm_get_results(g->ctx[r], out_pos ? out_pos + lo : nullptr, out_rc ? out_rc + lo : nullptr, out_mm ? out_mm + lo : nullptr, &st[r]);
    });
    if (rc != PGM_OK) return rc;
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (const pgm_stats &s : st) {
            stats->matched += s.matched;
            for (int k = 0; k < 256; k++) stats->per_mm[k] += s.per_mm[k];
            stats->patterns_inserted += s.patterns_inserted; stats->table_slots += s.table_slots;
            stats->candidates += s.candidates; stats->verified += s.verified; stats->accepted += s.accepted;
            stats->filter_positives += s.filter_positives;
        }
    }
    return PGM_OK;
}

int pgm_group_get_mismatches(pgm_group *g, uint64_t *out_offsets, uint8_t *out_pos, uint8_t *out_syms, uint64_t capacity, uint64_t *total) {
    if (!g) return PGM_ERR_INVALID_ARG;
    g->err.clear();
    const int n = g->size();
    std::vector<uint64_t> tot(n, 0);
    int rc = group_run(g, [&](int r) -> int { return pgm_get_mismatches(g->ctx[r], nullptr, nullptr, nullptr, 0, &tot[r]); });
    if (rc != PGM_OK) return rc;
    std::vector<uint64_t> base(n + 1, 0);
    for (int r = 0; r < n; r++) base[r + 1] = base[r] + tot[r];
    if (total) *total = base[n];
    if (!out_offsets && !out_pos && !out_syms) return PGM_OK;
    if ((out_pos || out_syms) && capacity < base[n])
        return gfail(g, PGM_ERR_INVALID_ARG, "pgm_group_get_mismatches: capacity is smaller than the number of mismatches");
    if (is_device_ptr(out_offsets) && n > 1) return gfail(g, PGM_ERR_UNSUPPORTED, "pgm_group_get_mismatches: host output arrays only");
    rc = group_run(g, [&](int r) -> int {
        const uint64_t lo = g->read_begin[r], cnt = g->read_begin[r + 1] - lo;
        uint64_t t = 0;
        // every rank writes the offsets of its reads (cnt + 1 entries: the last one is the next rank's first, same value after the shift)
        std::vector<uint64_t> local(out_offsets ? cnt + 1 : 0);
        int rr = pgm_get_mismatches(g->ctx[r], out_offsets ? local.data() : nullptr, out_pos ? out_pos + base[r] : nullptr,
                                    out_syms ? out_syms + base[r] : nullptr, tot[r], &t);
        if (rr == PGM_OK && out_offsets) {
            for (uint64_t i = 0; i < cnt; i++) out_offsets[lo + i] = local[i] + base[r];
            if (r == n - 1) out_offsets[lo + cnt] = local[cnt] + base[r];
        }
        return rr;
    });
    return rc;
}

} // extern "C"
