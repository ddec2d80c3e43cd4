// GpuTextMatcher.h — reference-side binding of stage 7 (exact matches between pseudogenomes): a PgTools::TextMatcher
// (matching/TextMatchers.h:54-61) over the C ABI's pgm_mem_* calls (include/pgrc_gpu_matcher.h), standing in for
// CopMEMMatcher (matching/copmem/CopMEMMatcher.h:22-89) where SimplePgMatcher creates it (SimplePgMatcher.cpp:16).
// Everything downstream of matchTexts — correctDestPositionDueToRevComplMatching, resolveMappingCollisionsInTheSameText,
// the sort / unique / overlap pass and the three output streams of markAndRemoveExactMatches (SimplePgMatcher.cpp:57-155)
// — is the reference's unmodified code.
//
// Needs the reference's headers on the include path; compiled only where the reference sources are present
// (oracle/Makefile, target `cli`).
#ifndef PGRC_B200_GPU_TEXT_MATCHER_H
#define PGRC_B200_GPU_TEXT_MATCHER_H

#include "matching/TextMatchers.h"
#include "pgrc_gpu_matcher.h"

namespace PgTools {

    class GpuTextMatcher : public TextMatcher {
        pgm_group *grp = nullptr;     // one device context by default; PGRC_GPU_DEVICES=0,1,... shares the query out over several GPUs
        size_t srcLength;
        uint32_t targetMatchLength;
        static void check(int rc, pgm_group *g, const char *what);   // message on stderr + exit, the reference's convention
    public:
        // = CopMEMMatcher(srcText, srcLength, targetMatchLength, minMatchLength): uploads the source text and builds its index
        GpuTextMatcher(const char *srcText, const size_t srcLength, const uint32_t targetMatchLength, uint32_t minMatchLength = UINT32_MAX);
        ~GpuTextMatcher() override;
        void matchTexts(vector<TextMatch> &resMatches, const string &destText, bool destIsSrc, bool revComplMatching,
                        uint32_t minMatchLength) override;
    };

    // What SimplePgMatcher's constructor gets instead of `new CopMEMMatcher(...)` (oracle/Makefile compiles
    // SimplePgMatcher.cpp with GpuTextMatcherHook.h force-included): the GPU matcher when PGRC_GPU_MATCHER=1 (and
    // PGRC_GPU_PGMATCH is not 0), the reference's CopMEMMatcher otherwise.  A GPU request that cannot be served ends the
    // program; it never becomes a silent CPU run.
    TextMatcher *newPgTextMatcher(const char *srcText, const size_t srcLength, const uint32_t targetMatchLength, uint32_t minMatchLength);

    // `new PgTextMatcherProxy(a, b, c, d)` is what the hook turns `new CopMEMMatcher(a, b, c, d)` into
    class PgTextMatcherProxy : public TextMatcher {
        TextMatcher *impl;
    public:
        PgTextMatcherProxy(const char *srcText, const size_t srcLength, const uint32_t targetMatchLength, uint32_t minMatchLength)
            : impl(newPgTextMatcher(srcText, srcLength, targetMatchLength, minMatchLength)) {}
        ~PgTextMatcherProxy() override { delete impl; }
        void matchTexts(vector<TextMatch> &resMatches, const string &destText, bool destIsSrc, bool revComplMatching,
                        uint32_t minMatchLength) override {
            impl->matchTexts(resMatches, destText, destIsSrc, revComplMatching, minMatchLength);
        }
    };
}

#endif
