// GpuReadsMatchers.h — the reference-side binding of the B200 matcher: C++ host code that plugs the C ABI
// (include/pgrc_gpu_matcher.h) in behind PgRC's own matcher class interface.
//
// Two classes sit behind the reference's bases and override exactly the protected virtuals the reference's
// hash-matcher classes override (matching/ReadsMatchers.h:45-46,128):
//     GpuReadsExactMatcher  : DefaultReadsExactMatcher     (ReadsMatchers.h:85-107,  .cpp:190-230)
//     GpuReadsApproxMatcher : AbstractReadsApproxMatcher   (ReadsMatchers.h:109-144, .cpp:276-341; mode 'i': .cpp:343-409)
// They fill the inherited result members (readMatchPos, readMatchRC, readMismatchesCount, matchedReadsCount,
// matchedCountPerMismatches), so everything downstream — getMatchedReadsBitmap, exportMatchesInPgOrder /
// exportMatchesInOriginalOrder, the archive writer — is the reference's unmodified code.
//
// This header needs the reference's headers on the include path (-I<PgRC source root>); it is compiled only where
// the reference sources are present (oracle/Makefile, target `cli`), never on the GPU box.
#ifndef PGRC_B200_GPU_READS_MATCHERS_H
#define PGRC_B200_GPU_READS_MATCHERS_H

#include "matching/ReadsMatchers.h"
#include "pgrc_gpu_matcher.h"

namespace PgTools {

    // One group of device contexts per mapReadsIntoPg call (one GPU by default; PGRC_GPU_DEVICES=0,1,... or =<count>
    // shards the stage over several): text + reads are uploaded once and shared by both phases.
    class GpuMatcherSession {
        pgm_group *grp = nullptr;
        uint_reads_cnt_max readsCount = 0;
    public:
        GpuMatcherSession(const char *pgPtr, uint64_t pgLength, ConstantLengthReadsSetInterface *readsSet);
        ~GpuMatcherSession();
        static void check(int rc, pgm_group *g, const char *what);   // prints pgm_group_last_error, exit(EXIT_FAILURE)
        static vector<int> devicesFromEnvironment();                 // PGRC_GPU_DEVICES / PGRC_GPU_DEVICE, default {0}
        void begin(uint32_t seedLength, uint32_t parts, uint32_t maxMismatches, uint32_t minMismatches, bool continuation,
                   bool interleaved = false);
        void pass(bool revCompMode);
        // mode 'c' (CopMEMReadsApproxMatcher): no seed table; a pass = text index + per-read queries
        void beginCopmem(uint32_t partLength, uint32_t maxMismatches, uint32_t minMismatches, bool continuation);
        void passCopmem(bool revCompMode);
        // copies the per-read results into the reference's member vectors
        void fetch(vector<uint64_t> &readMatchPos, vector<bool> &readMatchRC, vector<uint8_t> *readMismatchesCount,
                   uint_reads_cnt_max &matchedReadsCount, uint_reads_cnt_max *matchedCountPerMismatches);
        // mismatch lists of all matched reads (pgm_get_mismatches): offsets[readsCount + 1], read offsets, symbol codes
        void fetchMismatches(vector<uint64_t> &offsets, vector<uint8_t> &pos, vector<uint8_t> &syms);
    };

    class GpuReadsExactMatcher : public DefaultReadsExactMatcher {
        GpuMatcherSession *session;
    protected:
        void initMatching() override;
        void executeMatching(bool revCompMode = false) override;
    public:
        GpuReadsExactMatcher(GpuMatcherSession *session, char *pgPtr, const uint_pg_len_max pgLength, bool revComplPg,
                             ConstantLengthReadsSetInterface *readsSet, uint32_t matchPrefixLength);
    };

    // interleaved = true stands in for InterleavedReadsApproxMatcher (mode 'i', ReadsMatchers.h:146-165, .cpp:343-409),
    // false for DefaultReadsApproxMatcher (mode 'd'): same base class, same members, only the seed geometry differs.
    class GpuReadsApproxMatcher : public AbstractReadsApproxMatcher {
        GpuMatcherSession *session;
        uint_read_len_max partLength;
        bool interleaved;
        bool copmem;       // stands in for CopMEMReadsApproxMatcher (mode 'c', ReadsMatchers.h:167-185, .cpp:411-451; results of -t 1)
    protected:
        void initMatching() override;
        void initMatchingContinuation(DefaultReadsMatcher *pMatcher) override;
        void executeMatching(bool revCompMode = false) override;
        // export hooks (ReadsMatchers.h:52-54): the reference unpacks every matched read, reverse-complements it and
        // compares it with the pseudogenome again (ReadsMatchers.cpp:548-558); here the lists come from the device
        // (PGRC_GPU_EXPORT=0 keeps the reference's host code)
        void initEntryUpdating() override;
        void updateEntry(DefaultReadsListEntry &entry, uint_reads_cnt_max matchIdx, bool revComplPairFile) override;
        void closeEntryUpdating() override;
        bool deviceLists = false;
        vector<uint64_t> misOffsets;
        vector<uint8_t> misPos, misSyms;
    public:
        GpuReadsApproxMatcher(GpuMatcherSession *session, char *pgPtr, const uint_pg_len_max pgLength, bool revComplPg,
                              ConstantLengthReadsSetInterface *readsSet, uint32_t matchPrefixLength,
                              uint16_t readsExactMatchingChars, uint8_t maxMismatches, uint8_t minMismatches = 0,
                              bool interleaved = false, bool copmem = false);
    };

    // Same signature as PgTools::mapReadsIntoPg (ReadsMatchers.h:202-207).  The GPU matchers run when the matching mode letter
    // is 'g' / 'G' (dev build: -s g38; the hash-matcher semantics of 'd' / 'D'), or when PGRC_GPU_MATCHER=1 and the mode is
    // 'd'/'D', 'i'/'I' or 'c'/'C' (same letter as the CPU run: byte-identical archives).  There is no silent CPU run: a GPU
    // request that cannot be served (no device, prefix matching, a pseudogenome with N, an unknown PGRC_GPU_MATCHER value)
    // ends the program with a message, as the reference does for its own errors.
    const vector<bool> mapReadsIntoPgOnGpu(SeparatedPseudoGenome *sPg, bool revComplPg, bool preserveOrderMode,
                        ConstantLengthReadsSetInterface *readsSet, bool pairFileMode, bool revComplPairFile,
                        uint_read_len_max matchPrefixLength, uint16_t preReadsExactMatchingChars,
                        uint16_t readsExactMatchingChars, uint16_t minCharsPerMismatch, char preMatchingMode,
                        char matchingMode, bool dumpInfo, ostream &pgrcOut, uint8_t compressionLevel,
                        const string &pgDestFilePrefix, IndexesMapping *orgIndexesMapping);
}

#endif
