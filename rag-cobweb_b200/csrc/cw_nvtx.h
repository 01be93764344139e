// cw_nvtx.h -- NVTX ranges around the C-ABI entry points and the stages of the fused predict (SURVEY.md section 5:
// tracing).  Header-only NVTX v3: a no-op unless a profiler (ncu --nvtx, nsys) injects its library.
#pragma once
#include <nvtx3/nvToolsExt.h>

struct CwRange {
    explicit CwRange(const char *name) { nvtxRangePushA(name); }
    ~CwRange() { nvtxRangePop(); }
    CwRange(const CwRange &) = delete;
    CwRange &operator=(const CwRange &) = delete;
};
