// GpuTextMatcherHook.h — force-included (after the reference's own matching/copmem/CopMEMMatcher.h) when oracle/Makefile
// compiles the reference's matching/SimplePgMatcher.cpp for the `cli` target: the one place that file creates its text
// matcher, `new CopMEMMatcher(srcPg.data(), srcPg.length(), targetMatchLength, minMatchLength)` (SimplePgMatcher.cpp:16),
// then creates a PgTools::PgTextMatcherProxy, which picks the GPU matcher or the reference's CopMEMMatcher at run time.
// No reference source is modified.  A maintainer would write `matcher = newPgTextMatcher(...)` there instead.
#ifndef PGRC_B200_GPU_TEXT_MATCHER_HOOK_H
#define PGRC_B200_GPU_TEXT_MATCHER_HOOK_H
#include "GpuTextMatcher.h"
#define CopMEMMatcher(a, b, c, d) PgTools::PgTextMatcherProxy(a, b, c, d)
#endif
