// apx_jit.hpp -- per-request specialisation of the approximate SPLIT scan (see apx_jit.cpp, jit_apx_kernel.inc)
#pragma once
#include <string>
#include <vector>
#include "apx.hpp"

// The request-specific part of the CUDA source (constants + straight-line dense_<p> functions).  It doubles as the
// cache key of the compiled kernel: two requests with the same text here run the same code.
// resident CTAs per SM the kernel is compiled for (register budget); the launch grid is sms x this
int apx_jit_ctas();
// words of 32 pattern starts a lane holds (8: 4 warps per CTA, 4: 8 warps per CTA with half the registers each)
int apx_jit_wpl();
// launch geometry of the kernel apx_generate_prefix writes (environment knobs PM_JIT_* are for experiments)
struct ApxJitShape {
    int stream;          // 1: streaming form (jit_stream_kernel.inc), 0: unrolled tile form (jit_apx_kernel.inc)
    int w;               // words of 32 bases per lane (segment length / words per lane of a warp tile)
    int warps, stages, ctas;
    int tile_words;      // words per plane of a block tile
    size_t smem;         // dynamic shared memory per CTA
};
ApxJitShape apx_jit_shape(bool force_stream = false);
// exact = true: k = 0 patterns described as ApxPat with one piece evaluated position by position (dn / dshift / dcls);
// the kernel then writes hit keys directly instead of running the Landau-Vishkin check
std::string apx_generate_prefix(const ApxPat *pats, int npat, bool exact = false);
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
