// GpuTextMatcher.cpp — see GpuTextMatcher.h.  Host C++ over the C ABI; no CUDA in this file.
#include "GpuTextMatcher.h"
#include "GpuReadsMatchers.h"

#include "matching/copmem/CopMEMMatcher.h"

#include <cstdlib>
#include <cstring>

namespace PgTools {

    void GpuTextMatcher::check(int rc, pgm_group *g, const char *what) {
        if (rc == PGM_OK) return;
        fprintf(stderr, "GPU text matcher: %s failed: %s\n", what, pgm_group_last_error(g));
        exit(EXIT_FAILURE);
    }

    GpuTextMatcher::GpuTextMatcher(const char *srcText, const size_t srcLength, const uint32_t targetMatchLength, uint32_t minMatchLength)
            : srcLength(srcLength), targetMatchLength(targetMatchLength) {
        const vector<int> devices = GpuMatcherSession::devicesFromEnvironment();
        check(pgm_group_create((int) devices.size(), devices.data(), &grp), nullptr, "pgm_group_create");
        check(pgm_group_set_text(grp, srcText, srcLength), grp, "pgm_group_set_text");
        uint32_t par[4];
        check(pgm_group_mem_index(grp, targetMatchLength, minMatchLength, par), grp, "pgm_group_mem_index");
        // CopMEMMatcher::displayParams (copmem/CopMEMMatcher.cpp:98-108)
        cout << "copMEM PARAMETERS: l = " << targetMatchLength << "; K = " << par[0] << "; HASH_SIZE = " << par[3] << "; k1 = " << par[1]
             << "; k2 = " << par[2] << std::endl;
        cout << "Pseudogenome text index on the GPU";
        if (devices.size() > 1) cout << ", queries shared out over " << devices.size() << " device contexts";
        cout << endl;
    }

    GpuTextMatcher::~GpuTextMatcher() {
        pgm_group_destroy(grp);
    }

    void GpuTextMatcher::matchTexts(vector<TextMatch> &resMatches, const string &destText, bool destIsSrc, bool revComplMatching,
                                    uint32_t minMatchLength) {
        resMatches.clear();
        if (destIsSrc && destText.length() != srcLength) {
            fprintf(stderr, "GPU text matcher: destIsSrc with a destination of another length.\n");
            exit(EXIT_FAILURE);
        }
        // destIsSrc: the destination is the source (or its reverse complement, SimplePgMatcher.cpp:35-36), which the
        // devices already hold in both orientations — nothing to upload
        uint64_t count = 0;
        check(pgm_group_mem_match(grp, destIsSrc ? nullptr : destText.data(), destText.length(), destIsSrc, revComplMatching, minMatchLength, &count),
              grp, "pgm_group_mem_match");
        static_assert(sizeof(TextMatch) == sizeof(pgm_text_match), "TextMatch is three uint64 (matching/TextMatchers.h:11-14)");
        resMatches.resize(count, TextMatch(0, 0, 0));
        check(pgm_group_mem_get_matches(grp, reinterpret_cast<pgm_text_match *>(resMatches.data()), count), grp, "pgm_group_mem_get_matches");
        *logout << "Exact matches on the GPU: " << count << endl;
    }

    TextMatcher *newPgTextMatcher(const char *srcText, const size_t srcLength, const uint32_t targetMatchLength, uint32_t minMatchLength) {
        const char *env = getenv("PGRC_GPU_MATCHER");
        bool wantGpu = false;
        if (env && *env && strcmp(env, "0") != 0) {
            if (strcmp(env, "1") != 0) {
                fprintf(stderr, "GPU matcher: PGRC_GPU_MATCHER=%s is not understood (use 1 or 0).\n", env);
                exit(EXIT_FAILURE);
            }
            wantGpu = true;
        }
        if (const char *pm = getenv("PGRC_GPU_PGMATCH")) {      // 0 keeps stage 7 on the CPU while stage 4 runs on the GPU
            if (strcmp(pm, "0") == 0) wantGpu = false;
            else if (strcmp(pm, "1") != 0) {
                fprintf(stderr, "GPU matcher: PGRC_GPU_PGMATCH=%s is not understood (use 1 or 0).\n", pm);
                exit(EXIT_FAILURE);
            }
        }
        if (!wantGpu)
            return new ::CopMEMMatcher(srcText, srcLength, targetMatchLength, minMatchLength);
        for (size_t i = 0; i < srcLength; i++) {
            const char c = srcText[i];
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T') {
                fprintf(stderr, "GPU text matcher: the source pseudogenome contains the symbol %c at %zu: only ACGT pseudogenomes (N reads separated, "
                                "the default) can be indexed on the GPU (PGRC_GPU_PGMATCH=0 keeps this stage on the CPU).\n", c, i);
                exit(EXIT_FAILURE);
            }
        }
        return new GpuTextMatcher(srcText, srcLength, targetMatchLength, minMatchLength);
    }
}
