// TEST INFRASTRUCTURE — not part of the product path.
//
// Matcher-boundary harness around the UNMODIFIED reference classes:
// it is compiled by oracle/Makefile against the reference's own objects (built from
// /root/reference where they lie) into oracle/_ref/libpgrc_ref.so.  It is used to
//   (1) pin the plain-C restatement in pgrc_oracle.c (tests/, tests/golden/make_golden.py),
//   (2) serve as the CPU arm of bench.py (`--impl reference`, `cpu_baseline.kind="reference"`).
// Nothing under pgrc_b200/ may link or load it.
//
// What it drives (reference file:line):
//   PgTools::DefaultReadsApproxMatcher / DefaultReadsExactMatcher   matching/ReadsMatchers.h:85-172
//   ... ::matchConstantLengthReads()                                matching/ReadsMatchers.cpp:162-172
//   AbstractReadsApproxMatcher::continueMatchingConstantLengthReads matching/ReadsMatchers.cpp:174-184
//   parameter derivation and matcher selection restated from
//   PgTools::mapReadsIntoPg                                         matching/ReadsMatchers.cpp:693-779
// (mapReadsIntoPg itself cannot be called: it needs a SeparatedPseudoGenome and exports
// into an archive stream; the matchers' result vectors are protected, hence the Probe<>.)

#include <chrono>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <omp.h>

#include "matching/ReadsMatchers.h"
#include "readsset/PackedConstantLengthReadsSet.h"

using namespace PgTools;
using namespace PgReadsSet;

namespace {

// Exposes the protected result members of any approx matcher class.
template <class M>
struct Probe : M {
    using M::M;
    std::vector<uint64_t>& pos() { return this->readMatchPos; }
    std::vector<bool>& rc() { return this->readMatchRC; }
    std::vector<uint8_t>& mm() { return this->readMismatchesCount; }
    uint64_t matched() { return this->matchedReadsCount; }
    uint64_t better() { return this->betterMatchCount; }
    uint64_t falses() { return this->falseMatchCount; }
    uint_reads_cnt_max* hist() { return this->matchedCountPerMismatches; }
    void fill(DefaultReadsListEntry& e, uint_reads_cnt_max idx, bool revComplPairFile) { this->updateEntry(e, idx, revComplPairFile); }
};

struct ExactProbe : DefaultReadsExactMatcher {
    using DefaultReadsExactMatcher::DefaultReadsExactMatcher;
    std::vector<uint64_t>& pos() { return this->readMatchPos; }
    std::vector<bool>& rc() { return this->readMatchRC; }
    uint64_t matched() { return this->matchedReadsCount; }
    uint64_t better() { return this->betterMatchCount; }
    uint64_t falses() { return this->falseMatchCount; }
};

// LQ set followed by N set, as pgrc-encoder.cpp:349-352 builds it.  The reference's
// SumOfConstantLengthReadsSets leaves its ReadsSetBase properties default-constructed
// (readsCount = 0), which makes addReadsSetOfPatterns see zero patterns.
// The harness works around that WITHOUT touching reference sources: the subclass fills
// the inherited `properties` object.
struct SumWithProperties : SumOfConstantLengthReadsSets {
    SumWithProperties(ConstantLengthReadsSetInterface* a, ConstantLengthReadsSetInterface* b)
        : SumOfConstantLengthReadsSets(a, b) {
        properties->readsCount = readsCount();
        properties->maxReadLength = maxReadLength();
        properties->minReadLength = maxReadLength();
        properties->constantReadLength = true;
    }
};

struct CoutSilencer {
    std::streambuf* old;
    std::ostream* oldLog;
    NullBuffer nb;
    CoutSilencer() : old(std::cout.rdbuf(&nb)), oldLog(PgHelpers::logout) { PgHelpers::logout = &null_stream; }
    ~CoutSilencer() { std::cout.rdbuf(old); PgHelpers::logout = oldLog; }
};

template <class P>
void dumpApprox(P* m, uint32_t n, uint64_t* out_pos, uint8_t* out_rc, uint8_t* out_mm, uint64_t* st) {
    for (uint32_t i = 0; i < n; i++) {
        out_pos[i] = m->pos()[i];
        out_rc[i] = m->rc()[i] ? 1 : 0;
        out_mm[i] = m->mm()[i];
    }
    st[0] = m->matched(); st[1] = m->better(); st[2] = m->falses();
    for (int k = 0; k < 256; k++) st[3 + k] = m->hist()[k];
}

AbstractReadsApproxMatcher* newApprox(char mode, char* pg, uint64_t pgLen, bool rc,
                                      ConstantLengthReadsSetInterface* rs, uint32_t prefix,
                                      uint16_t seed, uint8_t maxMM, uint8_t minMM) {
    switch (tolower(mode)) {
        case 'd': return new Probe<DefaultReadsApproxMatcher>(pg, pgLen, rc, rs, prefix, seed, maxMM, minMM);
        case 'i': return new Probe<InterleavedReadsApproxMatcher>(pg, pgLen, rc, rs, prefix, seed, maxMM, minMM);
        case 'c': return new Probe<CopMEMReadsApproxMatcher>(pg, pgLen, rc, rs, prefix, seed, maxMM, minMM);
    }
    return nullptr;
}

// The reference's own export step for one read (exportMatchesInPgOrder, ReadsMatchers.cpp:583-588): entry at the match
// position, then updateEntry (:548-558) fills the mismatches.  Dumps (offset, actual symbol, mismatch symbol).
struct MisOut { uint64_t* offsets; uint8_t* off; char* pg; char* read; int rev_pair; double* seconds; };

template <class P>
void dumpMismatches(P* m, uint32_t n, const MisOut& o) {
    uint64_t at = 0;
    for (uint32_t i = 0; i < n; i++) {
        o.offsets[i] = at;
        if (m->pos()[i] == DefaultReadsMatcher::NOT_MATCHED_POSITION) continue;
        DefaultReadsListEntry entry(0);
        entry.advanceEntryByPosition(m->pos()[i], i, m->rc()[i]);
        m->fill(entry, i, o.rev_pair != 0);
        for (uint8_t k = 0; k < entry.mismatchesCount; k++, at++) {
            o.off[at] = (uint8_t)entry.mismatchOffset[k];
            o.pg[at] = PgHelpers::value2symbol(PgHelpers::cxtCode2ActualValue(entry.mismatchCode[k]));
            o.read[at] = PgHelpers::value2symbol(PgHelpers::cxtCode2MismatchValue(entry.mismatchCode[k]));
        }
    }
    o.offsets[n] = at;
}

void dumpMismatchesAny(AbstractReadsApproxMatcher* m, char mode, uint32_t n, const MisOut& o) {
    switch (tolower(mode)) {
        case 'd': dumpMismatches(static_cast<Probe<DefaultReadsApproxMatcher>*>(m), n, o); break;
        case 'i': dumpMismatches(static_cast<Probe<InterleavedReadsApproxMatcher>*>(m), n, o); break;
        case 'c': dumpMismatches(static_cast<Probe<CopMEMReadsApproxMatcher>*>(m), n, o); break;
    }
}

void dumpAny(AbstractReadsApproxMatcher* m, char mode, uint32_t n, uint64_t* p, uint8_t* r, uint8_t* c, uint64_t* st) {
    switch (tolower(mode)) {
        case 'd': dumpApprox(static_cast<Probe<DefaultReadsApproxMatcher>*>(m), n, p, r, c, st); break;
        case 'i': dumpApprox(static_cast<Probe<InterleavedReadsApproxMatcher>*>(m), n, p, r, c, st); break;
        case 'c': dumpApprox(static_cast<Probe<CopMEMReadsApproxMatcher>*>(m), n, p, r, c, st); break;
    }
}

int runReference(char* text, uint64_t text_len, const char* lq_reads, uint32_t n_lq, const char* n_reads, uint32_t n_n,
                 uint32_t read_len, uint32_t pre_seed, uint32_t seed, uint32_t min_chars_per_mismatch, char pre_mode, char mode,
                 int rev_compl, int threads, uint64_t* out_pos, uint8_t* out_rc, uint8_t* out_mm, uint64_t* out_stats,
                 double* out_seconds, const MisOut* mis);

}  // namespace

extern "C" {

int pgref_map_reads(char* text, uint64_t text_len, const char* lq_reads, uint32_t n_lq, const char* n_reads, uint32_t n_n,
                    uint32_t read_len, uint32_t pre_seed, uint32_t seed, uint32_t min_chars_per_mismatch, char pre_mode, char mode,
                    int rev_compl, int threads, uint64_t* out_pos, uint8_t* out_rc, uint8_t* out_mm, uint64_t* out_stats,
                    double* out_seconds) {
    return runReference(text, text_len, lq_reads, n_lq, n_reads, n_n, read_len, pre_seed, seed, min_chars_per_mismatch, pre_mode,
                        mode, rev_compl, threads, out_pos, out_rc, out_mm, out_stats, out_seconds, nullptr);
}

// pgref_map_reads + the mismatch lists the reference's export step builds for every matched read (updateEntry,
// ReadsMatchers.cpp:548-558, called as exportMatchesInPgOrder does, :583-588, with entry.idx = the read index):
// out_offsets[n+1], then per mismatch its offset, the actual (pseudogenome) symbol and the mismatch (read) symbol as
// decoded from the reference's context code.  The last matcher must be an approximate one.  *out_fill_seconds = wall
// time of that per-read loop alone (getRead + reverse complement + compare + addMismatch, and the dump).
int pgref_mismatch_lists(char* text, uint64_t text_len, const char* lq_reads, uint32_t n_lq, const char* n_reads, uint32_t n_n,
                         uint32_t read_len, uint32_t pre_seed, uint32_t seed, uint32_t min_chars_per_mismatch, char pre_mode,
                         char mode, int rev_compl, int rev_compl_pair_file, uint64_t* out_pos, uint8_t* out_rc, uint8_t* out_mm,
                         uint64_t* out_offsets, uint8_t* out_off, char* out_pg, char* out_read, double* out_fill_seconds) {
    uint64_t stats[259];
    MisOut mis{out_offsets, out_off, out_pg, out_read, rev_compl_pair_file, out_fill_seconds};
    return runReference(text, text_len, lq_reads, n_lq, n_reads, n_n, read_len, pre_seed, seed, min_chars_per_mismatch, pre_mode,
                        mode, rev_compl, 1, out_pos, out_rc, out_mm, stats, nullptr, &mis);
}

}  // extern "C"

namespace {

// Runs the reference's stage-4 matching exactly as mapReadsIntoPg would configure it
// (ReadsMatchers.cpp:699-779) and returns the three archive-visible per-read arrays.
//   text       : pseudogenome, ASCII ACGT, length text_len (+1 readable byte: the
//                reference reads txt[len] once, HashMatcher.h:62); restored on return
//   lq_reads   : n_lq * read_len ASCII bases over ACGT     (PackedConstantLengthReadsSet "ACGT")
//   n_reads    : n_n  * read_len ASCII bases over ACGNT    (… "ACGNT", 3 symbols/byte)
//   pre_seed   : 0 = single phase; otherwise seed of the pre-matching phase (dev -l)
//   mode/pre_mode : 'd','i','c' (upper case = shortcut mode, ReadsMatchers.cpp:711-712)
//   out_stats  : [0]=matchedReadsCount [1]=betterMatchCount [2]=falseMatchCount
//                [3..258]=matchedCountPerMismatches
//   out_seconds: wall time of matchConstantLengthReads (+ continuation), i.e. table build
//                + forward pass + RC pass, export excluded
// Returns 0, or -1 on bad arguments.
int runReference(char* text, uint64_t text_len,
                 const char* lq_reads, uint32_t n_lq,
                 const char* n_reads, uint32_t n_n,
                 uint32_t read_len, uint32_t pre_seed, uint32_t seed,
                 uint32_t min_chars_per_mismatch, char pre_mode, char mode,
                 int rev_compl, int threads,
                 uint64_t* out_pos, uint8_t* out_rc, uint8_t* out_mm,
                 uint64_t* out_stats, double* out_seconds, const MisOut* mis) {
    if (!text || read_len == 0 || read_len > 255 || seed == 0 || min_chars_per_mismatch == 0) return -1;
    CoutSilencer quiet;
    if (threads > 0) { omp_set_num_threads(threads); PgHelpers::numberOfThreads = threads; }

    PackedConstantLengthReadsSet lq(read_len, "ACGT", 4);
    lq.reserve(n_lq);
    for (uint32_t i = 0; i < n_lq; i++) lq.addRead(lq_reads + (size_t)i * read_len, read_len);
    PackedConstantLengthReadsSet nset(read_len, "ACGNT", 5);
    nset.reserve(n_n);
    for (uint32_t i = 0; i < n_n; i++) nset.addRead(n_reads + (size_t)i * read_len, read_len);
    SumWithProperties sum(&lq, &nset);
    ConstantLengthReadsSetInterface* rs = &sum;
    const uint32_t n = n_lq + n_n;

    // --- restated from mapReadsIntoPg (ReadsMatchers.cpp:699-714) ---
    const uint32_t prefix = DefaultReadsMatcher::DISABLED_PREFIX_MODE;
    uint16_t readLength = rs->maxReadLength();
    uint8_t maxMismatches = readLength / min_chars_per_mismatch;
    uint16_t readsExact = seed > readLength ? readLength : seed;
    uint16_t preExact = pre_seed > readLength ? readLength : pre_seed;
    uint16_t curExact = readsExact;
    char curMode = mode;
    if (preExact > 0) { curExact = preExact; curMode = pre_mode; }
    bool shortcut = toupper(curMode) == curMode;
    uint8_t curMinMM = shortcut ? maxMismatches : 0;
    memset(out_stats, 0, sizeof(uint64_t) * 259);

    auto t0 = std::chrono::steady_clock::now();
    DefaultReadsMatcher* matcher = nullptr;
    bool firstIsExact = false;
    if (readLength == curExact && tolower(curMode) != 'c') {
        matcher = new ExactProbe(text, text_len, rev_compl != 0, rs, prefix);
        firstIsExact = true;
    } else {
        matcher = newApprox(curMode, text, text_len, rev_compl != 0, rs, prefix, curExact, maxMismatches, curMinMM);
        if (!matcher) return -1;
    }
    matcher->matchConstantLengthReads();

    char lastMode = curMode;
    if (preExact > 0) {
        bool shortcut2 = toupper(mode) == mode;
        uint8_t targetMismatches = readLength / curExact - 1;   // as at :713 (from the pre-phase seed)
        uint8_t minMM2 = shortcut2 ? maxMismatches : targetMismatches + 1;   // :755
        AbstractReadsApproxMatcher* approx =
            newApprox(mode, text, text_len, rev_compl != 0, rs, prefix, readsExact, maxMismatches, minMM2);
        if (!approx) { delete matcher; return -1; }
        approx->continueMatchingConstantLengthReads(matcher);
        delete matcher;
        matcher = approx;
        firstIsExact = false;
        lastMode = mode;
    }
    auto t1 = std::chrono::steady_clock::now();
    if (out_seconds) *out_seconds = std::chrono::duration<double>(t1 - t0).count();

    if (firstIsExact) {
        ExactProbe* e = static_cast<ExactProbe*>(matcher);
        for (uint32_t i = 0; i < n; i++) {
            out_pos[i] = e->pos()[i];
            out_rc[i] = e->rc()[i] ? 1 : 0;
            out_mm[i] = e->pos()[i] == DefaultReadsMatcher::NOT_MATCHED_POSITION ? 255 : 0;
        }
        out_stats[0] = e->matched(); out_stats[1] = e->better(); out_stats[2] = e->falses();
    } else {
        dumpAny(static_cast<AbstractReadsApproxMatcher*>(matcher), lastMode, n, out_pos, out_rc, out_mm, out_stats);
        if (mis) {
            auto m0 = std::chrono::steady_clock::now();
            dumpMismatchesAny(static_cast<AbstractReadsApproxMatcher*>(matcher), lastMode, n, *mis);
            if (mis->seconds) *mis->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - m0).count();
        }
    }
    if (mis && firstIsExact) { delete matcher; return -1; }
    delete matcher;
    return 0;
}

}  // namespace

extern "C" {

// Packs ASCII reads with the reference's own SymbolsPackingFacility (layout pin for a11).
// symbols = "ACGT" (4/byte) or "ACGNT" (3/byte). Returns bytes per read.
int pgref_pack_reads(const char* reads, uint32_t n, uint32_t read_len, int with_n, uint8_t* out) {
    PackedConstantLengthReadsSet rs(read_len, with_n ? "ACGNT" : "ACGT", with_n ? 5 : 4);
    int spe = with_n ? 3 : 4;
    int packed = (read_len + spe - 1) / spe;
    for (uint32_t i = 0; i < n; i++) {
        rs.addRead(reads + (size_t)i * read_len, read_len);
        memcpy(out + (size_t)i * packed, rs.getPackedRead(i), packed);
    }
    return packed;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Stage 7: exact matches between pseudogenomes.
//   CopMEMMatcher(src, n, targetMatchLength, minMatchLength)   matching/copmem/CopMEMMatcher.cpp:571-593  (text index)
//   CopMEMMatcher::matchTexts                                  :605-624 -> processExactMatchQueryTight :332-481
//   SimplePgMatcher::markAndRemoveExactMatches                 matching/SimplePgMatcher.cpp:69-155
#include "matching/SimplePgMatcher.h"
#include "matching/copmem/CopMEMMatcher.h"

extern "C" {

// The raw vector matchTexts leaves in resMatches, in push order: out[3 i .. 3 i + 2] = {posSrcText, length, posDestText}.
// `dest` is the text exactly as the matcher is handed it (SimplePgMatcher::exactMatchPg reverse-complements it first
// when revComplMatching).  seconds[0] = index build (constructor), seconds[1] = matchTexts.  Returns 0; -1 bad
// arguments (the reference would exit()); -3 capacity too small (*count is still set).
int pgref_match_texts(const char* src, uint64_t n, const char* dest, uint64_t n2, int dest_is_src, int rev_compl,
                      uint32_t target_len, uint32_t min_len, int threads, uint64_t* out, uint64_t cap, uint64_t* count,
                      double* seconds) {
    uint32_t mml = min_len > target_len ? target_len : min_len;
    if (!src || !dest || mml < 24 || n < target_len) return -1;
    CoutSilencer quiet;
    if (threads > 0) { omp_set_num_threads(threads); PgHelpers::numberOfThreads = threads; }
    auto t0 = std::chrono::steady_clock::now();
    CopMEMMatcher m(src, n, target_len, min_len);
    auto t1 = std::chrono::steady_clock::now();
    std::vector<TextMatch> res;
    std::string d(dest, n2);
    auto t2 = std::chrono::steady_clock::now();
    m.matchTexts(res, d, dest_is_src != 0, rev_compl != 0, mml);
    auto t3 = std::chrono::steady_clock::now();
    if (seconds) {
        seconds[0] = std::chrono::duration<double>(t1 - t0).count();
        seconds[1] = std::chrono::duration<double>(t3 - t2).count();
    }
    *count = res.size();
    if (res.size() > cap) return -3;
    for (size_t i = 0; i < res.size(); i++) {
        out[3 * i] = res[i].posSrcText; out[3 * i + 1] = res[i].length; out[3 * i + 2] = res[i].posDestText;
    }
    return 0;
}

// The whole of SimplePgMatcher::markAndRemoveExactMatches for one destination pseudogenome: `dest` (n2 bytes, not
// reverse-complemented: the class does that itself) is replaced by the mapped sequence (matches cut out, one '%' each),
// *mapped_len = its new length; map_off / map_len = the two side streams (resPgMapOff, resPgMapLen).  dest_is_src: the
// destination IS the source text (`dest` is ignored on input and receives the mapped source).
int pgref_mark_matches(const char* src, uint64_t n, char* dest, uint64_t n2, int dest_is_src, int rev_compl,
                       uint32_t target_len, uint32_t min_len, int threads, uint64_t* mapped_len,
                       uint8_t* map_off, uint64_t off_cap, uint64_t* off_len, uint8_t* map_len, uint64_t len_cap, uint64_t* len_len,
                       double* seconds) {
    uint32_t mml = min_len > target_len ? target_len : min_len;
    if (!src || !dest || mml < 24) return -1;
    CoutSilencer quiet;
    if (threads > 0) { omp_set_num_threads(threads); PgHelpers::numberOfThreads = threads; }
    std::string s(src, n);
    std::string d = dest_is_src ? std::string() : std::string(dest, n2);
    std::string off, len;
    auto t0 = std::chrono::steady_clock::now();
    SimplePgMatcher matcher(s, target_len, min_len);
    matcher.markAndRemoveExactMatches(dest_is_src != 0, dest_is_src ? s : d, off, len, rev_compl != 0, min_len);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const std::string& mapped = dest_is_src ? s : d;
    *mapped_len = mapped.size(); *off_len = off.size(); *len_len = len.size();
    if (mapped.size() > (dest_is_src ? n : n2) || off.size() > off_cap || len.size() > len_cap) return -3;
    memcpy(dest, mapped.data(), mapped.size());
    memcpy(map_off, off.data(), off.size());
    memcpy(map_len, len.data(), len.size());
    return 0;
}

}  // extern "C"
