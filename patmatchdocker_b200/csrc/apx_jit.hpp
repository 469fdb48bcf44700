// apx_jit.hpp -- per-request specialisation of the approximate SPLIT scan (see apx_jit.cpp, jit_apx_kernel.inc)
#pragma once
#include <string>
#include <vector>
#include "apx.hpp"

// The request-specific part of the CUDA source (constants + straight-line dense_<p> functions).  It doubles as the
// cache key of the compiled kernel: two requests with the same text here run the same code.
std::string apx_generate_prefix(const ApxPat *pats, int npat);
// prefix + the constant kernel body
std::string apx_full_source(const std::string &prefix);
// NVRTC (loaded with dlopen on first use) -> cubin for sm_100a.  Returns 0 on success; `log` carries the reason otherwise.
int apx_jit_compile(const std::string &source, std::vector<char> &cubin, std::string &log);

// Argument block of k_scan_apx_jit (must match jit_apx_kernel.inc)
struct JitArgs {
    const unsigned *hi, *lo, *xx;
    long long nwords, n, tile0, ntiles;
    unsigned long long *keys, *count;
    long long cap;
    long long a0[2], a1[2];
    unsigned long long keytag[2];
};
