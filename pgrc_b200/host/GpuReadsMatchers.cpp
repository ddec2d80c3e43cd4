// GpuReadsMatchers.cpp — see GpuReadsMatchers.h.  Host C++ over the C ABI; no CUDA in this file.
//
// How it gets into the PgRC binary without touching a reference source file (oracle/Makefile, target `cli`):
// matching/ReadsMatchers.cpp is compiled with -DmapReadsIntoPg=mapReadsIntoPg_reference, and this file provides
// PgTools::mapReadsIntoPg, which either runs the GPU matchers or calls the renamed reference function.  A maintainer
// would instead add the two classes to ReadsMatchers.{h,cpp} and one `case` to the switches at
// ReadsMatchers.cpp:716-740 / :753-769.
//
// The packed read buffers are taken exactly as PackedConstantLengthReadsSet holds them
// (PackedConstantLengthReadsSet.h:17-18,40).  SumOfConstantLengthReadsSets (ReadsSetInterface.h:45-89) keeps its two
// sets private and has no accessor; until the two one-line accessors of INTEGRATION.md exist upstream, this
// translation unit reaches them through pointers to members obtained by explicit template instantiation (which the
// language exempts from access checking: [temp.spec]/6) — no macro games, no assumption about the object layout.  If
// a reads set is of an unknown class, the reads are re-packed through the public getRead() interface.
#include "GpuReadsMatchers.h"

#include <thread>
#include "readsset/PackedConstantLengthReadsSet.h"

#include <cstdlib>
#include <cstring>

namespace {
    // pointer-to-private-member access (see the header comment)
    template <class Tag, typename Tag::type Member> struct MemberThief { friend typename Tag::type stolen(Tag) { return Member; } };
    struct SumFirst { typedef ConstantLengthReadsSetInterface *SumOfConstantLengthReadsSets::*type; friend type stolen(SumFirst); };
    struct SumSecond { typedef ConstantLengthReadsSetInterface *SumOfConstantLengthReadsSets::*type; friend type stolen(SumSecond); };
    template struct MemberThief<SumFirst, &SumOfConstantLengthReadsSets::clrs1>;
    template struct MemberThief<SumSecond, &SumOfConstantLengthReadsSets::clrs2>;
}

namespace PgTools {

    // the reference's own function, renamed at compile time (see above)
    const vector<bool> mapReadsIntoPg_reference(SeparatedPseudoGenome *sPg, bool revComplPg, bool preserveOrderMode,
                        ConstantLengthReadsSetInterface *readsSet, bool pairFileMode, bool revComplPairFile,
                        uint_read_len_max matchPrefixLength, uint16_t preReadsExactMatchingChars,
                        uint16_t readsExactMatchingChars, uint16_t minCharsPerMismatch, char preMatchingMode,
                        char matchingMode, bool dumpInfo, ostream &pgrcOut, uint8_t compressionLevel,
                        const string &pgDestFilePrefix, IndexesMapping *orgIndexesMapping);

    // ------------------------------------------------------------------------------------------ session
    void GpuMatcherSession::check(int rc, pgm_group *g, const char *what) {
        if (rc == PGM_OK) return;
        // error convention of the reference: message on stderr + exit (e.g. ReadsMatchers.cpp:737-739)
        fprintf(stderr, "GPU matcher: %s failed: %s\n", what, pgm_group_last_error(g));
        exit(EXIT_FAILURE);
    }

    // PGRC_GPU_DEVICES = "0,1,2,3" (device ordinals; an ordinal may repeat) or "<count>" (devices 0 .. count-1);
    // PGRC_GPU_DEVICE = one ordinal; default: device 0
    vector<int> GpuMatcherSession::devicesFromEnvironment() {
        vector<int> devs;
        if (const char *list = getenv("PGRC_GPU_DEVICES")) {
            const string s(list);
            if (s.find(',') == string::npos) {
                for (int k = 0; k < atoi(s.c_str()); k++) devs.push_back(k);
            } else {
                size_t at = 0;
                while (at <= s.size()) {
                    const size_t end = s.find(',', at) == string::npos ? s.size() : s.find(',', at);
                    if (end > at) devs.push_back(atoi(s.substr(at, end - at).c_str()));
                    at = end + 1;
                }
            }
            if (devs.empty()) {
                fprintf(stderr, "GPU matcher: PGRC_GPU_DEVICES=%s names no device.\n", list);
                exit(EXIT_FAILURE);
            }
        } else {
            const char *dev = getenv("PGRC_GPU_DEVICE");
            devs.push_back(dev ? atoi(dev) : 0);
        }
        return devs;
    }

    namespace {
        struct PackedView { const uint8_t *data = nullptr; uint32_t count = 0; bool withN = false; vector<uint8_t> owned; };

        // re-pack through the public interface (SymbolsPackingFacility.cpp:168-185 layout)
        void repack(ConstantLengthReadsSetInterface *rs, uint_reads_cnt_max first, uint_reads_cnt_max count, bool withN, PackedView &v) {
            const uint32_t L = rs->maxReadLength(), spe = withN ? 3 : 4, sigma = withN ? 5 : 4, plen = (L + spe - 1) / spe;
            v.owned.assign((size_t)count * plen, 0);
            string read(L, 'A');
            for (uint_reads_cnt_max i = 0; i < count; i++) {
                rs->getRead(first + i, (char *)read.data());
                for (uint32_t b = 0; b < plen; b++) {
                    uint32_t val = 0;
                    for (uint32_t k = 0; k < spe; k++) {
                        const uint32_t p = b * spe + k;
                        const char c = p < L ? read[p] : 'A';
                        const uint32_t code = withN ? (c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'N' ? 3 : 4)
                                                    : (c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3);
                        val = val * sigma + code;
                    }
                    v.owned[(size_t)i * plen + b] = (uint8_t)val;
                }
            }
            v.data = v.owned.data(); v.count = count; v.withN = withN;
        }

        PackedView viewOf(ConstantLengthReadsSetInterface *rs) {
            PackedView v;
            v.count = rs->readsCount();
            v.withN = rs->getReadsSetProperties()->symbolsCount > 4;
            if (auto *packed = dynamic_cast<PgReadsSet::PackedConstantLengthReadsSet *>(rs)) {
                v.data = v.count ? packed->getPackedRead(0) : nullptr;
            } else {
                bool hasN = false;
                string read(rs->maxReadLength(), 'A');
                for (uint_reads_cnt_max i = 0; i < v.count && !hasN; i++) {
                    rs->getRead(i, (char *)read.data());
                    hasN = read.find('N') != string::npos;
                }
                repack(rs, 0, v.count, hasN, v);
            }
            return v;
        }
    }

    namespace {
        // Creating the CUDA context and loading the library's kernels takes about a second on a B200 — as long as the whole
        // matching stage of a small input.  With PGRC_GPU_MATCHER set, a thread started at program load does it while
        // PgRC reads the FASTQ and builds the pseudogenome (stages 1-3); the session joins it before its first call.
        struct GpuWarmup {
            std::thread th;
            GpuWarmup() {
                const char *env = getenv("PGRC_GPU_MATCHER");
                if (env && strcmp(env, "1") == 0)
                    th = std::thread([] {
                        for (int dev : GpuMatcherSession::devicesFromEnvironment()) {
                            pgm_ctx *c = nullptr;
                            if (pgm_create(dev, &c) == PGM_OK) pgm_destroy(c);
                        }
                    });
            }
            void wait() { if (th.joinable()) th.join(); }
            ~GpuWarmup() { wait(); }
        } gpuWarmup;
    }

    GpuMatcherSession::GpuMatcherSession(const char *pgPtr, uint64_t pgLength, ConstantLengthReadsSetInterface *readsSet) {
        gpuWarmup.wait();
        const vector<int> devs = devicesFromEnvironment();
        check(pgm_group_create((int) devs.size(), devs.data(), &grp), nullptr, "pgm_group_create");
        if (devs.size() > 1) cout << "GPU matcher: stage sharded over " << devs.size() << " device contexts." << endl;
        check(pgm_group_set_text(grp, pgPtr, pgLength), grp, "pgm_group_set_text");   // read-only: no in-place reverse complement
        readsCount = readsSet->readsCount();
        PackedView a, b;
        if (auto *sum = dynamic_cast<SumOfConstantLengthReadsSets *>(readsSet)) {
            a = viewOf(sum->*stolen(SumFirst()));                                   // LQ set, then N set: the global read index
            b = viewOf(sum->*stolen(SumSecond()));                                  // of SumOfConstantLengthReadsSets
        } else {
            a = viewOf(readsSet);
        }
        // C ABI: ACGT-packed reads first, ACGNT-packed reads second
        if (b.count == 0 && a.withN) { b = std::move(a); a = PackedView(); }
        if (a.withN || (b.count && !b.withN)) {
            // (never produced by pgrc-encoder: LQ is ACGT and N is ACGNT, or one single set) — re-pack both as ACGNT
            PackedView all;
            repack(readsSet, 0, readsCount, true, all);
            check(pgm_group_set_reads(grp, nullptr, 0, all.data, all.count, readsSet->maxReadLength()), grp, "pgm_group_set_reads");
            check(pgm_group_upload(grp), grp, "pgm_group_upload");                  // the re-packed copy dies with this constructor
            return;
        }
        check(pgm_group_set_reads(grp, a.data, a.count, b.data, b.count, readsSet->maxReadLength()), grp, "pgm_group_set_reads");
        if (!a.owned.empty() || !b.owned.empty())
            check(pgm_group_upload(grp), grp, "pgm_group_upload");                  // re-packed copies die with this constructor
    }

    GpuMatcherSession::~GpuMatcherSession() { pgm_group_destroy(grp); }

    void GpuMatcherSession::begin(uint32_t seedLength, uint32_t parts, uint32_t maxMismatches, uint32_t minMismatches, bool continuation,
                                  bool interleaved) {
        check(pgm_group_match_begin(grp, seedLength, parts, maxMismatches, minMismatches, continuation ? 1 : 0, interleaved ? 1 : 0), grp,
              "pgm_group_match_begin");
    }

    void GpuMatcherSession::beginCopmem(uint32_t partLength, uint32_t maxMismatches, uint32_t minMismatches, bool continuation) {
        check(pgm_group_copmem_begin(grp, partLength, maxMismatches, minMismatches, continuation ? 1 : 0), grp, "pgm_group_copmem_begin");
    }

    void GpuMatcherSession::passCopmem(bool revCompMode) {
        check(pgm_group_copmem_pass(grp, revCompMode ? 1 : 0), grp, "pgm_group_copmem_pass");
    }

    void GpuMatcherSession::pass(bool revCompMode) {
        check(pgm_group_pass(grp, revCompMode ? 1 : 0), grp, "pgm_group_pass");
    }

    void GpuMatcherSession::fetch(vector<uint64_t> &readMatchPos, vector<bool> &readMatchRC, vector<uint8_t> *readMismatchesCount,
                                  uint_reads_cnt_max &matchedReadsCount, uint_reads_cnt_max *matchedCountPerMismatches) {
        vector<uint8_t> rc(readsCount), mm(readsCount);
        readMatchPos.resize(readsCount);
        pgm_stats st;
        check(pgm_group_get_results(grp, readMatchPos.data(), rc.data(), mm.data(), &st), grp, "pgm_group_get_results");
        readMatchRC.resize(readsCount);
        for (uint_reads_cnt_max i = 0; i < readsCount; i++) readMatchRC[i] = rc[i] != 0;
        matchedReadsCount = (uint_reads_cnt_max) st.matched;
        if (readMismatchesCount) *readMismatchesCount = std::move(mm);
        if (matchedCountPerMismatches)
            for (int k = 0; k <= NOT_MATCHED_COUNT; k++) matchedCountPerMismatches[k] = (uint_reads_cnt_max) st.per_mm[k];
    }

    void GpuMatcherSession::fetchMismatches(vector<uint64_t> &offsets, vector<uint8_t> &pos, vector<uint8_t> &syms) {
        uint64_t total = 0;
        check(pgm_group_get_mismatches(grp, nullptr, nullptr, nullptr, 0, &total), grp, "pgm_group_get_mismatches");
        offsets.resize((size_t) readsCount + 1);
        pos.resize(total + 1);
        syms.resize(total + 1);
        check(pgm_group_get_mismatches(grp, offsets.data(), pos.data(), syms.data(), total, &total), grp, "pgm_group_get_mismatches");
    }

    // ------------------------------------------------------------------------------------------ exact path
    GpuReadsExactMatcher::GpuReadsExactMatcher(GpuMatcherSession *session, char *pgPtr, const uint_pg_len_max pgLength, bool revComplPg,
                                               ConstantLengthReadsSetInterface *readsSet, uint32_t matchPrefixLength)
            : DefaultReadsExactMatcher(pgPtr, pgLength, revComplPg, readsSet, matchPrefixLength), session(session) {}

    void GpuReadsExactMatcher::initMatching() {                                    // replaces ReadsMatchers.cpp:190-196
        DefaultReadsMatcher::initMatching();
        session->begin(matchingLength, 1, 0, 0, false);
    }

    void GpuReadsExactMatcher::executeMatching(bool revCompMode) {                 // replaces ReadsMatchers.cpp:198-230
        time_checkpoint();
        session->pass(revCompMode);
        if (revCompMode == revComplPg) {                                           // last pass of this matcher
            session->fetch(readMatchPos, readMatchRC, nullptr, matchedReadsCount, nullptr);
            cout << "... exact matched " << matchedReadsCount << " reads (" << (readsCount - matchedReadsCount)
                 << " left) on the GPU in " << time_millis() << " msec (last pass incl. results)." << endl;
        }
    }

    // ------------------------------------------------------------------------------------------ k-mismatch path
    GpuReadsApproxMatcher::GpuReadsApproxMatcher(GpuMatcherSession *session, char *pgPtr, const uint_pg_len_max pgLength, bool revComplPg,
                                                 ConstantLengthReadsSetInterface *readsSet, uint32_t matchPrefixLength,
                                                 uint16_t readsExactMatchingChars, uint8_t maxMismatches, uint8_t minMismatches,
                                                 bool interleaved, bool copmem)
            : AbstractReadsApproxMatcher(pgPtr, pgLength, revComplPg, readsSet, matchPrefixLength, readsExactMatchingChars,
                                         maxMismatches, minMismatches), session(session), partLength(readsExactMatchingChars),
              interleaved(interleaved), copmem(copmem) {}

    void GpuReadsApproxMatcher::initMatching() {                                   // replaces ReadsMatchers.cpp:276-285
        DefaultReadsMatcher::initMatching();
        readMismatchesCount.clear();
        readMismatchesCount.insert(readMismatchesCount.end(), readsCount, NOT_MATCHED_COUNT);
        if (copmem) session->beginCopmem(partLength, maxMismatches, minMismatches, false);   // replaces ReadsMatchers.cpp:411-415
        else session->begin(partLength, targetMismatches + 1, maxMismatches, minMismatches, false, interleaved);
    }

    void GpuReadsApproxMatcher::initMatchingContinuation(DefaultReadsMatcher *pMatcher) {   // replaces :287-295
        // the device still holds the first phase's per-read state; reads matched with <= minMismatches are left
        // out of the new table there (getMatchedReadsBitmap(minMismatches), ReadsMatchers.cpp:290-291)
        AbstractReadsApproxMatcher::initMatchingContinuation(pMatcher);
        if (copmem) session->beginCopmem(partLength, maxMismatches, minMismatches, true);    // replaces ReadsMatchers.cpp:417-419
        else session->begin(partLength, targetMismatches + 1, maxMismatches, minMismatches, true, interleaved);
    }

    void GpuReadsApproxMatcher::executeMatching(bool revCompMode) {                // replaces ReadsMatchers.cpp:297-341
        time_checkpoint();
        cout << "Matching" << (revCompMode ? " in Pg reverse" : "") << " (GPU)...\n" << endl;
        if (copmem) session->passCopmem(revCompMode);                             // replaces ReadsMatchers.cpp:421-451
        else session->pass(revCompMode);
        if (revCompMode == revComplPg)                                             // last pass of this matcher
            session->fetch(readMatchPos, readMatchRC, &readMismatchesCount, matchedReadsCount, matchedCountPerMismatches);
        printApproxMatchingStats();
    }

    void GpuReadsApproxMatcher::initEntryUpdating() {                              // replaces ReadsMatchers.cpp:546
        const char *env = getenv("PGRC_GPU_EXPORT");
        deviceLists = !(env && strcmp(env, "0") == 0);
        if (deviceLists) {
            time_checkpoint();
            session->fetchMismatches(misOffsets, misPos, misSyms);
            *logout << "... mismatch lists of " << matchedReadsCount << " reads (" << misOffsets.back() << " mismatches) from the GPU in "
                    << time_millis() << " msec. " << endl;
        }
    }

    void GpuReadsApproxMatcher::updateEntry(DefaultReadsListEntry &entry, uint_reads_cnt_max matchIdx, bool revComplPairFile) {
        if (!deviceLists) {                                                        // the reference's host code, unchanged
            AbstractReadsApproxMatcher::updateEntry(entry, matchIdx, revComplPairFile);
            return;
        }
        // replaces ReadsMatchers.cpp:548-558: the device list is the forward fill (fillEntryWithMismatches, :40-52) of the
        // read as it lies on the pseudogenome; fillEntryWithReversedMismatches (:54-66) is the same list walked backwards
        // with complemented symbols and mirrored offsets
        static const char SYM[5] = {'A', 'C', 'G', 'T', 'N'};
        const uint64_t b = misOffsets[matchIdx], e = misOffsets[matchIdx + 1];
        const bool reversed = revComplPairFile ? (readMatchRC[matchIdx] != (bool) (entry.idx % 2)) : (bool) readMatchRC[matchIdx];
        if (!reversed) {
            for (uint64_t k = b; k < e; k++)
                entry.addMismatch(mismatch2CxtCode(SYM[misSyms[k] & 3], SYM[misSyms[k] >> 2]), misPos[k]);
        } else {
            for (uint64_t k = e; k-- > b;)
                entry.addMismatch(mismatch2CxtCode(reverseComplement(SYM[misSyms[k] & 3]), reverseComplement(SYM[misSyms[k] >> 2])),
                                  readLength - misPos[k] - 1);
        }
    }

    void GpuReadsApproxMatcher::closeEntryUpdating() {
        vector<uint64_t>().swap(misOffsets);
        vector<uint8_t>().swap(misPos);
        vector<uint8_t>().swap(misSyms);
    }

    // ------------------------------------------------------------------------------------------ mapReadsIntoPg
    const vector<bool> mapReadsIntoPgOnGpu(SeparatedPseudoGenome *sPg, bool revComplPg, bool preserveOrderMode,
                        ConstantLengthReadsSetInterface *readsSet, bool pairFileMode, bool revComplPairFile,
                        uint_read_len_max matchPrefixLength, uint16_t preReadsExactMatchingChars,
                        uint16_t readsExactMatchingChars, uint16_t minCharsPerMismatch, char preMatchingMode,
                        char matchingMode, bool dumpInfo, ostream &pgrcOut, uint8_t compressionLevel,
                        const string &pgDestFilePrefix, IndexesMapping *orgIndexesMapping) {
        // parameter derivation and phase structure of the reference, ReadsMatchers.cpp:699-779
        const uint_read_len_max readLength = readsSet->maxReadLength();
        const uint8_t maxMismatches = readLength / minCharsPerMismatch;
        readsExactMatchingChars = std::min<uint16_t>(readsExactMatchingChars, readLength);
        preReadsExactMatchingChars = std::min<uint16_t>(preReadsExactMatchingChars, readLength);
        const bool twoPhases = preReadsExactMatchingChars > 0;
        const uint16_t firstSeed = twoPhases ? preReadsExactMatchingChars : readsExactMatchingChars;
        const char firstMode = twoPhases ? preMatchingMode : matchingMode;
        const uint8_t firstMinMismatches = isupper((unsigned char) firstMode) ? maxMismatches : 0;
        uint8_t targetMismatches = readLength / firstSeed - 1;
        char *pgPtr = (char *) sPg->getPgSequence().data();
        const uint_pg_len_max pgLength = sPg->getPgSequence().length();

        GpuMatcherSession session(pgPtr, pgLength, readsSet);
        DefaultReadsMatcher *matcher;
        if (readLength == firstSeed && tolower(firstMode) != 'c')                  // (mode 'c' uses CopMEM here too, :717-720)
            matcher = new GpuReadsExactMatcher(&session, pgPtr, pgLength, revComplPg, readsSet, matchPrefixLength);
        else
            matcher = new GpuReadsApproxMatcher(&session, pgPtr, pgLength, revComplPg, readsSet, matchPrefixLength,
                                                firstSeed, maxMismatches, firstMinMismatches, tolower(firstMode) == 'i',
                                                tolower(firstMode) == 'c');
        cout << "Target pseudogenome length: " << pgLength << endl;
        *logout << endl;
        cout << "readsAlignmentSeedLength (minCharsPerMismatch, matchingMode): " << (int) firstSeed <<
             " (" << (int) minCharsPerMismatch << ", " << firstMode << ")" << endl;
        *logout << "targetMismatches (maxMismatches, minMismatches): " << (int) targetMismatches <<
                " (" << (int) maxMismatches << ", " << (int) firstMinMismatches << ")" << endl;
        matcher->matchConstantLengthReads();

        if (twoPhases) {
            const uint8_t secondMinMismatches = isupper((unsigned char) matchingMode) ? maxMismatches : targetMismatches + 1;
            AbstractReadsApproxMatcher *approxMatcher = new GpuReadsApproxMatcher(&session, pgPtr, pgLength, revComplPg, readsSet,
                    matchPrefixLength, readsExactMatchingChars, maxMismatches, secondMinMismatches, tolower(matchingMode) == 'i',
                    tolower(matchingMode) == 'c');
            targetMismatches = readLength / readsExactMatchingChars - 1;
            cout << endl << "Reads matching 2nd PHASE." << endl;
            cout << "readsExactMatchingChars (minCharsPerMismatch, matchingMode): " << (int) readsExactMatchingChars <<
                 " (" << (int) minCharsPerMismatch << ", " << matchingMode << ")" << endl;
            cout << "targetMismatches (maxMismatches, minMismatches): " << (int) targetMismatches <<
                 " (" << (int) maxMismatches << ", " << (int) secondMinMismatches << ")" << endl;
            approxMatcher->continueMatchingConstantLengthReads(matcher);
            delete matcher;
            matcher = approxMatcher;
        }
        if (dumpInfo) matcher->writeMatchesInfo(pgDestFilePrefix);
        const vector<bool> res = matcher->getMatchedReadsBitmap();
        // the reference's own export code, unchanged (ReadsMatchers.cpp:785-792)
        if (matchPrefixLength == DefaultReadsMatcher::DISABLED_PREFIX_MODE) {
            if (preserveOrderMode)
                matcher->exportMatchesInOriginalOrder(sPg, pgrcOut, compressionLevel, pgDestFilePrefix, orgIndexesMapping, pairFileMode, revComplPairFile);
            else
                matcher->exportMatchesInPgOrder(sPg, pgrcOut, compressionLevel, pgDestFilePrefix, orgIndexesMapping, pairFileMode, revComplPairFile);
        }
        delete matcher;
        return res;
    }

    // What pgrc-encoder.cpp:359 calls.
    const vector<bool> mapReadsIntoPg(SeparatedPseudoGenome *sPg, bool revComplPg, bool preserveOrderMode,
                        ConstantLengthReadsSetInterface *readsSet, bool pairFileMode, bool revComplPairFile,
                        uint_read_len_max matchPrefixLength, uint16_t preReadsExactMatchingChars,
                        uint16_t readsExactMatchingChars, uint16_t minCharsPerMismatch, char preMatchingMode,
                        char matchingMode, bool dumpInfo, ostream &pgrcOut, uint8_t compressionLevel,
                        const string &pgDestFilePrefix, IndexesMapping *orgIndexesMapping) {
        // ConstantLengthPatternsOnTextHashMatcher.cpp:30 takes the pattern count from
        // getReadsSetProperties()->readsCount, which SumOfConstantLengthReadsSets never fills — mode 'd' then sees 0
        // patterns.  Filling the (public, otherwise unused) field makes the reference's mode 'd' the oracle it is meant
        // to be; the reference sources stay untouched.
        if (readsSet->getReadsSetProperties()->readsCount == 0)
            readsSet->getReadsSetProperties()->readsCount = readsSet->readsCount();
        // Selection.  Mode letter 'g' / 'G' (dev build: -s g38): the GPU matchers with the semantics of 'd' / 'D'.
        // PGRC_GPU_MATCHER=1: the GPU matchers for the letters the CPU build knows ('d', 'i', 'c': same letter in the archive
        // header, byte-identical archives).  Unset, empty or 0: the reference's own function.  Anything else is an error —
        // a mistyped value must not become a silent CPU run.
        const char *env = getenv("PGRC_GPU_MATCHER");
        bool wantGpu = false;
        if (env && *env && strcmp(env, "0") != 0) {
            if (strcmp(env, "1") != 0) {
                fprintf(stderr, "GPU matcher: PGRC_GPU_MATCHER=%s is not understood (use 1 or 0).\n", env);
                exit(EXIT_FAILURE);
            }
            wantGpu = true;
        }
        auto gpuLetter = [](char c) { return tolower(c) == 'g'; };
        if (gpuLetter(matchingMode) || (preReadsExactMatchingChars > 0 && gpuLetter(preMatchingMode))) {
            wantGpu = true;
            if (gpuLetter(matchingMode)) matchingMode = isupper((unsigned char) matchingMode) ? 'D' : 'd';
            if (gpuLetter(preMatchingMode)) preMatchingMode = isupper((unsigned char) preMatchingMode) ? 'D' : 'd';
        }
        if (wantGpu) {
            auto hashMode = [](char c) { return tolower(c) == 'd' || tolower(c) == 'i' || tolower(c) == 'c'; };   // (c: the -t 1 results)
            if (!hashMode(matchingMode) || (preReadsExactMatchingChars > 0 && !hashMode(preMatchingMode))) {
                fprintf(stderr, "GPU matcher: unknown matching mode %c (GPU matchers exist for d, i, c and g).\n", matchingMode);
                exit(EXIT_FAILURE);
            }
            if (matchPrefixLength != DefaultReadsMatcher::DISABLED_PREFIX_MODE) {
                fprintf(stderr, "GPU matcher: prefix matching is not supported on the GPU (unset PGRC_GPU_MATCHER to run the CPU matchers).\n");
                exit(EXIT_FAILURE);
            }
            // the device holds the pseudogenome in 2 bits per base: with the dev option -N (N reads not separated) the HQ
            // pseudogenome may contain N (DividedPCLReadsSets.cpp:12-13) — refuse up front, not after the matching has run
            const string &pg = sPg->getPgSequence();
            for (size_t i = 0; i < pg.size(); i++) {
                const char c = pg[i];
                if (c != 'A' && c != 'C' && c != 'G' && c != 'T') {
                    fprintf(stderr, "GPU matcher: the pseudogenome contains the symbol %c at %zu: only ACGT pseudogenomes (N reads separated, the default) "
                                    "can be matched on the GPU (unset PGRC_GPU_MATCHER to run the CPU matchers).\n", c, i);
                    exit(EXIT_FAILURE);
                }
            }
            return mapReadsIntoPgOnGpu(sPg, revComplPg, preserveOrderMode, readsSet, pairFileMode, revComplPairFile, matchPrefixLength,
                                       preReadsExactMatchingChars, readsExactMatchingChars, minCharsPerMismatch, preMatchingMode,
                                       matchingMode, dumpInfo, pgrcOut, compressionLevel, pgDestFilePrefix, orgIndexesMapping);
        }
        return mapReadsIntoPg_reference(sPg, revComplPg, preserveOrderMode, readsSet, pairFileMode, revComplPairFile, matchPrefixLength,
                                        preReadsExactMatchingChars, readsExactMatchingChars, minCharsPerMismatch, preMatchingMode,
                                        matchingMode, dumpInfo, pgrcOut, compressionLevel, pgDestFilePrefix, orgIndexesMapping);
    }
}
