// engine.cu -- CUDA kernels (sm_100a) and the C ABI of the B200-native PatMatch engine.
//
// Replaces the reference's search engine process (nrgrep_coords, spawned by
// www/FlaskApp/FlaskApp/patmatch.py:733-743).  Pipeline of one search:
//
//   scan    k_scan_bytes    bit-parallel Shift-And over the resident text; emits one
//                           candidate (anchor position, piece) per exact piece hit
//                           [esimpleScan @4136d0 type 1 / simpleScan @416600]
//           k_scan_dense    BWD / FWD plans: every anchor whose verification succeeds
//   sort    cub radix sort  candidate keys -> the order the reference meets them
//   verify  k_verify        anchored k-error NFA left and right of the anchor
//                           [esimple checkMatch @4151d0, checkMatch1 @414190]
//   chain   k_chain         the reference restarts its scan at the end of every reported
//                           hit (recSearchFile @402250); resolved per dependency cluster
//   select  cub select      compacts the chosen hits in output order
//
// No CPU fallback anywhere: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/device/device_scan.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <cmath>
#include <memory>
#include <mutex>
#include <thread>
#include <condition_variable>
#include "../../include/patmatch_b200.h"
#include "plan.hpp"
#include "apx_jit.hpp"
#include <map>

#define PM_MAXK 15

static thread_local std::string g_err;
const char *pm_last_error(void) { return g_err.c_str(); }
const char *pm_version(void) { return "patmatch_b200 0.1 (sm_100a)"; }

#define CK(call)                                                                       \
    do {                                                                               \
        cudaError_t _e = (call);                                                       \
        if (_e != cudaSuccess) {                                                       \
            g_err = std::string(#call) + " (engine.cu:" + std::to_string(__LINE__) + "): " + cudaGetErrorString(_e); \
            return PM_ERR_CUDA;                                                        \
        }                                                                              \
    } while (0)

// ---------------------------------------------------------------------------------------
// device-side plan description (passed by value)
struct DevPlan {
    int type, m, k, L, npieces;
    int ins, del, subs;
    int start_line, end_line;     // leading '^' / trailing '$' (recCheckLeftContext @402170 / recCheckRightContext @4021e0)
    int V[PM_MAX_PIECES];
    unsigned long long trig[PM_MAX_PIECES];
    unsigned long long init, fin;
    // EXTENDED plans: keys carry the start of the scanned window of plain positions; the anchor sits ext_off bytes
    // further and splits the pattern at position ext_anchor.  Closure masks per walk (extendedLoadVerif @412c60).
    int ext_off, ext_anchor, ext_repeats;
    int maxleft;                  // bytes left of its anchor a verification can examine at most (bounded plans)
    unsigned long long IL, FL, AL, initL, IR, FR, AR, initR;
    int ext_lead_opt;             // end-anchored plan, pattern begins with an optional position: forward-scan quirk
    unsigned long long IS, FS, AS;
};

// candidate key = pattern id << PM_PID_SHIFT | text position << 4 | piece  (pattern id 0 for single-pattern searches)
#define PM_POS_BITS 36
#define PM_PID_SHIFT 40
__host__ __device__ __forceinline__ long long key_pos(unsigned long long key) { return (long long)((key >> 4) & ((1ULL << PM_POS_BITS) - 1ULL)); }
__host__ __device__ __forceinline__ int key_pid(unsigned long long key) { return (int)(key >> PM_PID_SHIFT); }

__device__ __forceinline__ bool plan_is_ext(const DevPlan &pl) { return pl.type == PM_PLAN_EXT_BEG || pl.type == PM_PLAN_EXT_END; }
// text position the verification is anchored at, from a candidate key
__device__ __forceinline__ long long anchor_of(const DevPlan &pl, long long key) { return key_pos((unsigned long long)key) + (plan_is_ext(pl) ? pl.ext_off : 0); }
// the byte that decides which record / fill a candidate belongs to
__device__ __forceinline__ long long locus_of(const DevPlan &pl, long long anchor)
{
    return (pl.type == PM_PLAN_FWD || pl.type == PM_PLAN_EXT_END) ? anchor - 1 : anchor;
}

struct Cand { long long key, beg, end, reach; };   // == pm_candidate
struct H16 { long long a, b; };                     // == pm_hit

#include "packed.cuh"

// ---------------------------------------------------------------------------------------
// Anchored k-error NFA on one side of the anchor (checkMatch1 @414190, one direction).
// dir < 0: consumes text[pos-1], text[pos-2], ... down to `lim` (record start / scan start);
// dir > 0: consumes text[pos], text[pos+1], ... up to `lim` (record end).  The record
// delimiter '\n' also stops the walk (recGetRecord @402030).
// Returns 1 with *ext = bytes consumed by the chosen match, *err = its error row and
// *steps = bytes examined (for the dependency test of the chain stage).
// ctx: the pattern is anchored on this side ('^' for dir < 0, '$' for dir > 0): a final state only counts
// when the match boundary sits at the record / scan-range limit (recCheckLeftContext @402170,
// recCheckRightContext @4021e0).
__device__ __forceinline__ int nfa_side(const unsigned char *__restrict__ text, const unsigned long long *__restrict__ T,
                                        int dir, int plen, int kmax, int ins, int del, int subs, int ctx,
                                        long long pos, long long lim, long long *ext, int *err, long long *steps)
{
    auto edge_ok = [&](long long step) -> bool {
        if (!ctx) return true;
        if (dir < 0) { const long long e = pos - step; return e <= lim || text[e - 1] == '\n'; }
        const long long e = pos + step;
        return e >= lim || text[e] == '\n';
    };
    const unsigned long long fin = 1ULL << (plen - 1);
    const unsigned long long live = (fin << 1) - 1ULL;      // plen == 64 -> all ones
    unsigned long long R[PM_MAXK + 1];
    int kb = kmax;
    long long best_ext = -1;
    int best_err = kmax;
    *steps = 0;
    for (int e = 0; e <= kb; e++) {
        R[e] = del ? (e >= 64 ? ~0ULL : ((1ULL << e) - 1ULL)) : 0ULL;
        if ((R[e] & fin) && edge_ok(0)) { best_err = e; kb = e - 1; best_ext = 0; }
    }
    unsigned long long first = 1;
    long long step = 0;
    for (;;) {
        long long tp = dir < 0 ? pos - step - 1 : pos + step;
        if (dir < 0 ? (tp < lim) : (tp >= lim)) break;
        unsigned c = text[tp];
        if (c == '\n') break;
        step++;
        *steps = step;
        const unsigned long long Tc = T[c];
        unsigned long long oldp = R[0];
        R[0] = ((R[0] << 1) | first) & Tc;
        unsigned long long newp = R[0];
        if ((R[0] & fin) && edge_ok(step)) { *ext = step; *err = 0; return 1; }
        for (int e = 1; e <= kb; e++) {
            unsigned long long x = 0;
            if (del) x = newp << 1;
            if (ins) x |= oldp;
            if (subs) x |= (oldp << 1) | first;
            const unsigned long long nr = (((R[e] << 1) | first) & Tc) | x;
            oldp = R[e];
            R[e] = nr;
            newp = nr;
            if ((nr & fin) && edge_ok(step)) {
                int ec = e, ed;
                for (;;) {
                    ed = ec - 1;
                    if (ed < 0) break;
                    if (!(R[ed] & fin)) break;
                    ec = ed;
                }
                if (ed < 0) { *ext = step; *err = 0; return 1; }
                kb = ed; best_err = ec; best_ext = step;
                break;
            }
        }
        if (!(R[kb] & live)) break;
        first = 0;
    }
    if (best_ext < 0) return 0;
    *ext = best_ext;
    *err = best_err;
    return 1;
}

// Buffer fills of the reference (bufLoad @41bbf0, recSearchFile @402298): the file is scanned one
// fill [S_k, E_k) at a time; fills are computed on the host (compute_fills) and looked up here.
struct Fills { const long long *S, *E; int n; };

__device__ __forceinline__ int fill_of(const Fills &f, long long x) { return fill_lookup(f.S, f.n, x); }     // last k with S[k] <= x

// ---------------------------------------------------------------------------------------
// Patterns of 65..255 positions: the same anchored k-error NFA with PM_MW-word state vectors (the reference keeps
// multi-word masks too, createMask @41b430).  Tables: T[byte][PM_MW].  Only verification needs this: the scan
// kernels look at pieces / sub-patterns of at most 64 positions.
#define PM_MW 4
struct MWord { unsigned long long w[PM_MW]; };

__device__ __forceinline__ MWord mw_shl1(const MWord &a, unsigned long long in)
{
    MWord r;
    unsigned long long carry = in;
#pragma unroll
    for (int w = 0; w < PM_MW; w++) { r.w[w] = (a.w[w] << 1) | carry; carry = a.w[w] >> 63; }
    return r;
}

__device__ int nfa_side_mw(const unsigned char *__restrict__ text, const unsigned long long *__restrict__ T,
                           int dir, int plen, int kmax, int ins, int del, int subs, int ctx,
                           long long pos, long long lim, long long *ext, int *err, long long *steps)
{
    auto edge_ok = [&](long long step) -> bool {
        if (!ctx) return true;
        if (dir < 0) { const long long e = pos - step; return e <= lim || text[e - 1] == '\n'; }
        const long long e = pos + step;
        return e >= lim || text[e] == '\n';
    };
    const int fw = (plen - 1) >> 6;
    const unsigned long long fin = 1ULL << ((plen - 1) & 63);
    auto has_fin = [&](const MWord &a) -> bool { return (a.w[fw] & fin) != 0; };
    auto alive = [&](const MWord &a) -> bool {
        for (int w = 0; w < fw; w++) if (a.w[w]) return true;
        return (a.w[fw] & ((fin << 1) - 1ULL)) != 0;
    };
    MWord R[PM_MAXK + 1];
    int kb = kmax;
    long long best_ext = -1;
    int best_err = kmax;
    *steps = 0;
    for (int e = 0; e <= kb; e++) {
#pragma unroll
        for (int w = 0; w < PM_MW; w++) R[e].w[w] = 0;
        if (del) R[e].w[0] = (1ULL << e) - 1ULL;                       // e <= 15
        if (has_fin(R[e]) && edge_ok(0)) { best_err = e; kb = e - 1; best_ext = 0; }
    }
    unsigned long long first = 1;
    long long step = 0;
    for (;;) {
        const long long tp = dir < 0 ? pos - step - 1 : pos + step;
        if (dir < 0 ? (tp < lim) : (tp >= lim)) break;
        const unsigned c = text[tp];
        if (c == '\n') break;
        step++;
        *steps = step;
        const unsigned long long *Tc = T + (size_t)c * PM_MW;
        MWord oldp = R[0];
        MWord sh = mw_shl1(R[0], first);
#pragma unroll
        for (int w = 0; w < PM_MW; w++) R[0].w[w] = sh.w[w] & Tc[w];
        MWord newp = R[0];
        if (has_fin(R[0]) && edge_ok(step)) { *ext = step; *err = 0; return 1; }
        bool lowered = false;
        for (int e = 1; e <= kb; e++) {
            MWord x;
#pragma unroll
            for (int w = 0; w < PM_MW; w++) x.w[w] = 0;
            if (del) x = mw_shl1(newp, 0);
            if (ins) {
#pragma unroll
                for (int w = 0; w < PM_MW; w++) x.w[w] |= oldp.w[w];
            }
            if (subs) {
                const MWord s = mw_shl1(oldp, first);
#pragma unroll
                for (int w = 0; w < PM_MW; w++) x.w[w] |= s.w[w];
            }
            const MWord y = mw_shl1(R[e], first);
            MWord nr;
#pragma unroll
            for (int w = 0; w < PM_MW; w++) nr.w[w] = (y.w[w] & Tc[w]) | x.w[w];
            oldp = R[e];
            R[e] = nr;
            newp = nr;
            if (has_fin(nr) && edge_ok(step)) {
                int ec = e, ed;
                for (;;) {
                    ed = ec - 1;
                    if (ed < 0) break;
                    if (!has_fin(R[ed])) break;
                    ec = ed;
                }
                if (ed < 0) { *ext = step; *err = 0; return 1; }
                kb = ed; best_err = ec; best_ext = step;
                lowered = true;
                break;
            }
        }
        (void)lowered;
        if (!alive(R[kb])) break;
        first = 0;
    }
    if (best_ext < 0) return 0;
    *ext = best_ext;
    *err = best_err;
    return 1;
}

// esimple checkMatch @4151d0 for candidate (piece i, anchor pos) with scan range [tbeg, n).
__device__ int check_match(const DevPlan &pl, const unsigned char *__restrict__ text, long long n,
                           const unsigned long long *__restrict__ TL, const unsigned long long *__restrict__ TR,
                           int i, long long pos, long long tbeg, long long *beg, long long *end, long long *reach)
{
    const long long p = pl.type == PM_PLAN_FWD ? pos - 1 : pos;
    *reach = pos;
    if (p < tbeg || p >= n) return 0;
    if (text[p] == '\n') return 0;
    const int lb = pl.V[i], rl = pl.m - lb;
    long long bext = 0, fext = 0, steps = 0;
    int berr = 0, ferr = 0;
    if (lb > 0) {
        int ok = pl.m > 64 ? nfa_side_mw(text, TL + (size_t)i * 256 * PM_MW, -1, lb, pl.k, pl.ins, pl.del, pl.subs, pl.start_line, pos, tbeg, &bext, &berr, &steps)
                           : nfa_side(text, TL + (size_t)i * 256, -1, lb, pl.k, pl.ins, pl.del, pl.subs, pl.start_line, pos, tbeg, &bext, &berr, &steps);
        *reach = pos - steps - (pl.start_line ? 1 : 0);    // '^' also reads the byte left of the boundary
        if (!ok) return 0;
    } else if (pl.start_line) {
        // 4141ef-414236: only inserted text bytes may separate an empty left part from the record start
        long long ptr = pos;
        int e = 0;
        for (;;) {
            if (ptr <= tbeg || text[ptr - 1] == '\n') break;
            if (!pl.ins) { *reach = ptr - 1; return 0; }
            ptr--; e++;
            if (e > pl.k) { *reach = ptr - 1; return 0; }
        }
        *reach = ptr - 1;
        bext = pos - ptr; berr = e;
    }
    if (rl > 0) {
        if (!(pl.m > 64 ? nfa_side_mw(text, TR + (size_t)i * 256 * PM_MW, +1, rl, pl.k - berr, pl.ins, pl.del, pl.subs, pl.end_line, pos, n, &fext, &ferr, &steps)
                        : nfa_side(text, TR + (size_t)i * 256, +1, rl, pl.k - berr, pl.ins, pl.del, pl.subs, pl.end_line, pos, n, &fext, &ferr, &steps))) return 0;
    } else if (pl.end_line) {
        // 414eae-414f19: the same on the right of an empty right part
        const int kf = pl.k - berr;
        long long q = pos;
        int e = 0;
        for (;;) {
            if (q >= n || text[q] == '\n') break;
            if (!pl.ins) return 0;
            q++; e++;
            if (e > kf) return 0;
        }
        fext = q - pos;
    }
    *beg = pos - bext;
    *end = pos + fext;
    return 1;
}

// ---------------------------------------------------------------------------------------
// EXTENDED patterns (positions with '?'), k = 0: one walk of checkMatch @411aa0.  State bit u = element u (counted
// away from the anchor) has consumed a byte or was skipped; T[c] = elements accepting byte c; I / F / A are the
// closure masks of the runs of optional elements.  The initial state only pre-skips the FIRST element when it is
// optional -- a run of two or more optional elements next to the anchor cannot be skipped as a whole on the first
// byte (the reference's observable behaviour: (GAT.?.?.?AAGTCC) does not match GATAAGTCC).
// Returns the bytes consumed by the shortest accepted extension, -1 if none; *steps = bytes examined.
__device__ __forceinline__ long long ext_side(const unsigned char *__restrict__ text, const unsigned long long *__restrict__ T, int repeats,
                                              int len, unsigned long long I, unsigned long long F, unsigned long long A,
                                              unsigned long long init, int dir, int ctx, long long pos, long long lim, long long *steps)
{
    const unsigned long long fin = 1ULL << (len - 1);
    unsigned long long D = init, carry = 1;
    long long n = 0;
    *steps = 0;
    for (;;) {
        if (D & fin) {
            bool ok = true;
            if (ctx) {
                if (dir < 0) { const long long e = pos - n; ok = e <= lim || text[e - 1] == '\n'; }
                else { const long long e = pos + n; ok = e >= lim || text[e] == '\n'; }
            }
            if (ok) return n;
        }
        const long long tp = dir > 0 ? pos + n : pos - n - 1;
        if (dir > 0 ? tp >= lim : tp < lim) return -1;
        const unsigned c = text[tp];
        if (c == '\n') return -1;                                      // record delimiter (recGetRecord @402030)
        n++;
        *steps = n;
        D = (((D << 1) | carry) & T[c]) | (repeats ? D & T[256 + c] : 0ULL);     // advance | stay ('*' / '+')
        carry = 0;
        if (!D) return -1;
        const unsigned long long x = D | F;
        D = (((~(x - I)) ^ x) & A) | D;
    }
}

// checkMatch @411aa0 at `anchor` with scan range [tbeg, n)
__device__ int check_match_ext(const DevPlan &pl, const unsigned char *__restrict__ text, long long n,
                               const unsigned long long *__restrict__ TL, const unsigned long long *__restrict__ TR,
                               long long anchor, long long tbeg, long long *beg, long long *end, long long *reach)
{
    const long long p = locus_of(pl, anchor);
    *reach = anchor;
    if (p < tbeg || p >= n) return 0;
    if (text[p] == '\n') return 0;
    const int a = pl.ext_anchor, m = pl.m;
    long long bext = 0, fext = 0, steps = 0;
    if (a == 0) {
        *reach = anchor - (pl.start_line ? 1 : 0);
        if (pl.start_line && !(anchor <= tbeg || text[anchor - 1] == '\n')) return 0;
    } else {
        bext = ext_side(text, TL, pl.ext_repeats, a, pl.IL, pl.FL, pl.AL, pl.initL, -1, pl.start_line, anchor, tbeg, &steps);
        *reach = anchor - steps - (pl.start_line ? 1 : 0);
        if (bext < 0) return 0;
        if (pl.ext_lead_opt) {
            // extendedScan @4116f0 (forward mode) proposes this anchor only if ITS automaton is in a final state here.
            // It starts empty at the scan start and after every '\n', and the run of optional positions at the start
            // of the pattern becomes enterable only through the closure that follows each byte.  An occurrence that
            // starts later than such a byte is always seen; one that starts ON it is seen only if that automaton,
            // re-run from there, ends final at the anchor.
            const long long b = anchor - bext;
            if (b - 1 < *reach) *reach = b - 1;                               // the outcome depends on the byte before b
            if (b <= tbeg || text[b - 1] == '\n') {
                unsigned long long D = 0;
                for (long long q = b; q < anchor; q++) {
                    const unsigned c = text[q];
                    const unsigned long long Bc = __brevll(TL[c]) >> (64 - a);      // forward order of P[0, anchor)
                    const unsigned long long Sc = pl.ext_repeats ? __brevll(TL[256 + c]) >> (64 - a) : 0ULL;
                    D = (((D << 1) | 1ULL) & Bc) | (D & Sc);
                    const unsigned long long x = D | pl.FS;
                    D = (((~(x - pl.IS)) ^ x) & pl.AS) | D;
                }
                if (!(D & (1ULL << (a - 1)))) return 0;
            }
        }
    }
    if (a == m) {
        if (pl.end_line && !(anchor >= n || text[anchor] == '\n')) return 0;
    } else {
        fext = ext_side(text, TR, pl.ext_repeats, m - a, pl.IR, pl.FR, pl.AR, pl.initR, +1, pl.end_line, anchor, n, &steps);
        if (fext < 0) return 0;
    }
    *beg = anchor - bext;
    *end = anchor + fext;
    return 1;
}

// ---------------------------------------------------------------------------------------
// Scan kernel: byte-level Shift-And over superimposed pieces.
//   D = ((D << 1) | init) & B[text[p]] ;  piece i ends at p  <=>  D bit i*L+L-1
// One block stages TILE bytes (+64 bytes of left halo) into shared memory with coalesced
// 128-bit loads; every thread then walks its own SEG-byte run.  Rows are padded by 16
// bytes so that the per-thread 128-bit shared loads of a quarter-warp hit distinct banks.
// 5-bit residue codes for peptide datasets: letters (either case) -> 0..25, every other byte -> PEP_OTHER; six codes
// per 32-bit word, 0.667 B per residue instead of 1.  Replaces the bytes nrgrep_coords reads for proteome searches.
#define PEP_OTHER 31
__global__ void __launch_bounds__(256) k_pack5(const unsigned char *__restrict__ text, long long n, long long nwords, unsigned *__restrict__ codes)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += stride) {
        unsigned v = 0;
#pragma unroll
        for (int r = 0; r < 6; r++) {
            const long long i = w * 6 + r;
            unsigned code = PEP_OTHER;
            if (i < n) {
                const unsigned c = text[i] | 0x20u;
                if (c >= 'a' && c <= 'z') code = c - 'a';
            }
            v |= code << (5 * r);
        }
        codes[w] = v;
    }
}

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_SEG = 128;                              // bytes per thread per tile
constexpr int SCAN_ROW = SCAN_SEG + 16;                    // padded row
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_SEG;         // 32 KiB
constexpr int SCAN_HALO = 64;

template <typename T>
struct ScanArgs {
    const unsigned char *text;
    const unsigned *codes;       // non-null: scan the 5-bit residue codes (6 per 32-bit word, k_pack5) instead of the raw bytes
    long long n;
    long long p0, p1;            // end positions p handled: p0 <= p < p1
    long long tile0;             // first tile index (tile = p / SCAN_TILE)
    long long ntiles;
    const unsigned long long *B; // 256 masks
    T init, fin;
    T trig[PM_MAX_PIECES];
    int L, npieces;
    unsigned long long *keys;
    unsigned long long *count;   // [0] = candidates produced
    long long cap;
    unsigned long long keytag;   // pid << PM_PID_SHIFT (multi-pattern requests)
};

template <typename T>
__device__ __forceinline__ void scan_emit(const ScanArgs<T> &a, T hit, long long p)
{
    if (p < a.p0 || p >= a.p1 || p >= a.n) return;
    const long long w = p - a.L + 1;
    for (int i = 0; i < a.npieces; i++) {
        if (hit & a.trig[i]) {
            unsigned long long idx = atomicAdd(a.count, 1ULL);
            if ((long long)idx < a.cap) a.keys[idx] = a.keytag | ((unsigned long long)w << 4) | (unsigned)i;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_bytes(const ScanArgs<T> a)
{
    __shared__ T sB[256];
    __shared__ __align__(16) unsigned char srow[SCAN_THREADS * SCAN_ROW];
    __shared__ __align__(16) unsigned char shalo[SCAN_HALO];
    const int tid = threadIdx.x;
    if (a.codes) {
        // table over the 32 residue codes: a letter code collects both cases, code 31 every other byte (an
        // over-approximation that k_verify's re-check on the raw bytes removes)
        unsigned long long m = 0;
        if (tid < 26) m = a.B['A' + tid] | a.B['a' + tid];
        else if (tid == PEP_OTHER) {
            for (int c = 0; c < 256; c++)
                if (!((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'))) m |= a.B[c];
        }
        sB[tid] = (T)m;
    } else sB[tid] = (T)a.B[tid];
    const bool aligned = ((size_t)a.text & 15) == 0;

    for (long long t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        const long long tstart = (a.tile0 + t) * SCAN_TILE;
        __syncthreads();
        // ---- stage tile ----
        if (a.codes) {
            // 5-bit codes: every word carries 6 residues; they are unpacked into the byte rows the run below reads
            const long long w0 = (tstart - SCAN_HALO) / 6 - (tstart < SCAN_HALO ? 1 : 0);
            const long long w1 = (tstart + SCAN_TILE + 5) / 6;
            for (long long w = w0 + tid; w < w1; w += SCAN_THREADS) {
                if (w < 0) continue;
                const unsigned v = (w * 6 < a.n) ? __ldg(a.codes + w) : 0x3fffffffu;
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    const long long o = w * 6 + r - tstart;              // offset of this residue relative to the tile start
                    const unsigned char code = (unsigned char)((v >> (5 * r)) & 31u);
                    if (o >= 0 && o < SCAN_TILE) srow[(o / SCAN_SEG) * SCAN_ROW + (o % SCAN_SEG)] = code;
                    else if (o < 0 && o >= -SCAN_HALO) shalo[SCAN_HALO + o] = code;
                }
            }
        } else {
#pragma unroll
        for (int it = 0; it < SCAN_TILE / 16 / SCAN_THREADS; it++) {
            const int c = it * SCAN_THREADS + tid;              // 16-byte chunk in tile
            const long long g = tstart + (long long)c * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (aligned && g + 16 <= a.n) v = __ldg(reinterpret_cast<const uint4 *>(a.text + g));
            else if (g < a.n) {
                unsigned char tmp[16];
#pragma unroll
                for (int b = 0; b < 16; b++) tmp[b] = (g + b < a.n) ? a.text[g + b] : 0;
                v = *reinterpret_cast<uint4 *>(tmp);
            }
            const int row = c / (SCAN_SEG / 16), col = c % (SCAN_SEG / 16);
            *reinterpret_cast<uint4 *>(srow + row * SCAN_ROW + col * 16) = v;
        }
        if (tid < SCAN_HALO) {
            const long long g = tstart - SCAN_HALO + tid;
            shalo[tid] = (g >= 0 && g < a.n) ? a.text[g] : 0;
        }
        }
        __syncthreads();
        // ---- per-thread run ----
        const long long seg = tstart + (long long)tid * SCAN_SEG;
        if (seg >= a.p1 || seg >= a.n) continue;
        if (seg + SCAN_SEG <= a.p0) continue;
        T D = 0;
        // warm-up over the L-1 bytes before the run (no emission)
        {
            int o = tid * SCAN_SEG - (a.L - 1);                 // offset relative to tile start
            if (tstart + o < 0) o = (int)(-tstart);
            for (; o < tid * SCAN_SEG; o++) {
                unsigned c = o < 0 ? shalo[SCAN_HALO + o] : srow[(o / SCAN_SEG) * SCAN_ROW + (o % SCAN_SEG)];
                D = ((D << 1) | a.init) & sB[c];
            }
        }
        const unsigned char *my = srow + tid * SCAN_ROW;
#pragma unroll 2
        for (int v4 = 0; v4 < SCAN_SEG / 16; v4++) {
            const uint4 q = *reinterpret_cast<const uint4 *>(my + v4 * 16);
            const unsigned ww[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int wi = 0; wi < 4; wi++) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const unsigned c = (ww[wi] >> (8 * b)) & 0xffu;
                    D = ((D << 1) | a.init) & sB[c];
                    const T hit = D & a.fin;
                    if (hit) scan_emit(a, hit, seg + v4 * 16 + wi * 4 + b);
                }
            }
        }
    }
}

// BWD / FWD plans: the reference's approximate filters only propose anchors; the hit list is
// decided by checkMatch alone (validated against the reference), so every anchor is verified.
struct DenseArgs {
    DevPlan pl;
    const unsigned char *text;
    long long n, a0, a1;         // anchors a0 <= a < a1
    const unsigned long long *TL, *TR;
    unsigned long long *keys, *count;
    long long cap;
    Fills fills;
    unsigned long long keytag;
};

__global__ void __launch_bounds__(256) k_scan_dense(const DenseArgs a)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long pos = a.a0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; pos < a.a1; pos += stride) {
        long long b, e, r;
        const long long p = a.pl.type == PM_PLAN_FWD ? pos - 1 : pos;
        const int f = fill_of(a.fills, p);
        const long long S = a.fills.S[f], E = a.fills.E[f];
        if (a.pl.type == PM_PLAN_BWD && pos + (a.pl.L - a.pl.k) > E) continue;
        if (check_match(a.pl, a.text, E, a.TL, a.TR, 0, pos, S, &b, &e, &r)) {
            unsigned long long idx = atomicAdd(a.count, 1ULL);
            if ((long long)idx < a.cap) a.keys[idx] = a.keytag | ((unsigned long long)pos << 4);
        }
    }
}

// Exact re-check of a packed-scan candidate on the raw bytes: which pieces match the window
// at w, and does that set fire piece i (same rule as scan_emit).
__device__ __forceinline__ bool raw_trigger(const DevPlan &pl, const unsigned long long *__restrict__ B,
                                            const unsigned char *__restrict__ text, long long n, long long w, int i)
{
    if (w < 0 || w + pl.L > n) return false;
    unsigned long long D = ~0ULL;
    for (int j = 0; j < pl.L; j++) D &= B[text[w + j]] >> j;      // bit i*L set <=> piece i accepts all L bytes
    unsigned long long top = 0;
    for (int q = 0; q < pl.npieces; q++)
        if ((D >> (q * pl.L)) & 1ULL) top |= 1ULL << (q * pl.L + pl.L - 1);
    return (top & pl.trig[i]) != 0;
}

// verification of sorted candidates, unclipped (scan start = 0)
__global__ void __launch_bounds__(128) k_verify(const DevPlan pl, const unsigned char *__restrict__ text, long long n,
                                                const unsigned long long *__restrict__ B,
                                                const unsigned long long *__restrict__ TL, const unsigned long long *__restrict__ TR,
                                                const unsigned long long *__restrict__ keys, long long ncand, Cand *__restrict__ out,
                                                int recheck, const Fills fills)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncand) return;
    const unsigned long long key = keys[j];
    const long long pos = (long long)(key >> 4);
    const int i = (int)(key & 15);
    Cand c;
    c.key = (long long)key;
    const long long anchor = anchor_of(pl, (long long)key);
    const long long p = locus_of(pl, anchor);
    const int f = fill_of(fills, p);
    const long long S = fills.S[f], E = fills.E[f];
    // the scanned window of the candidate must lie inside its fill
    const long long wlen = pl.type == PM_PLAN_SIMPLE ? pl.m : (pl.type == PM_PLAN_SPLIT || plan_is_ext(pl)) ? pl.L
                         : pl.type == PM_PLAN_BWD ? pl.L - pl.k : 0;
    if (pos + wlen > E || (plan_is_ext(pl) && pos < S) || (recheck && !raw_trigger(pl, B, text, n, pos, i))) {
        c.beg = -1; c.end = -1; c.reach = anchor;
    } else if (plan_is_ext(pl)) {
        long long b = -1, e = -1, r = anchor;
        if (!check_match_ext(pl, text, E, TL, TR, anchor, S, &b, &e, &r)) { b = -1; e = -1; }
        c.beg = b; c.end = e; c.reach = r;
    } else if (pl.type == PM_PLAN_SIMPLE) {
        bool ok = true;
        for (int jj = 64; jj < pl.m && ok; jj++) {                                  // positions beyond the scanned window
            const unsigned ch = text[pos + jj];
            ok = (TL[(size_t)(jj - 64) * 4 + (ch >> 6)] >> (ch & 63)) & 1ULL;
        }
        if (ok) { c.beg = pos; c.end = pos + pl.m; c.reach = pos - (pl.start_line ? 1 : 0); }   // the chain stage applies '^' / '$'
        else { c.beg = -1; c.end = -1; c.reach = pos; }
    } else {
        long long b = -1, e = -1, r = pos;
        if (!check_match(pl, text, E, TL, TR, i, pos, S, &b, &e, &r)) { b = -1; e = -1; }
        c.beg = b; c.end = e; c.reach = r;
    }
    out[j] = c;
}

// Chain stage.  The reference reports the first verified candidate at or after its scan
// start, then restarts the scan at the end of that hit; verification never looks left of
// the scan start.  Candidate j opens an independent cluster when no earlier candidate can
// produce a hit that ends right of everything j depends on.
__device__ __forceinline__ long long dep_lo(const DevPlan &pl, const Cand &c)
{
    const long long a = locus_of(pl, anchor_of(pl, c.key));
    return c.reach < a ? c.reach : a;
}

// ---------------------------------------------------------------------------------------
// Chain stage: the reference restarts its scan at the end of every reported hit (recSearchFile @402250), so whether a
// verified candidate is reported depends on where the previous hit ended.  Candidates are cut into dependency clusters
// (cand_opens_cluster) and a cluster is resolved by ONE WARP: it loads 32 candidates at a time, coalesced, and walks them
// in order with ballots -- the first candidate that is still eligible under the current scan start is examined by its
// lane (re-verified with the scan start clipped when its unclipped verification had looked left of it), everything in
// front of it is settled in the same step.  A step is taken per candidate that gets examined, not per candidate, so
// dense searches (a short motif with many errors: hundreds of thousands of overlapping candidates per buffer fill) cost
// a few instructions per skipped candidate instead of a dependent global load each (3.1 Gb, (GATAAGC) -k 3ids, 836 M
// candidates: the one-thread-per-cluster walk took 1.02 s).
//
// Candidate j opens a cluster when nothing before it can influence it OR ANY LATER candidate: every hit of an
// earlier candidate ends at or before the leftmost byte that j and its successors can examine.  Bounded plans:
// a hit ends within `span` of its anchor and a verification looks at most maxleft (+ the '^' context byte, + the
// one-byte locus offset of end-anchored plans) to the left.  '*' / '+' patterns: prefix maxima of the hit ends
// against suffix minima of the examined ranges.  Another pattern of the request or another buffer fill always does.
__device__ __forceinline__ bool cand_opens_cluster(const DevPlan &pl, const Cand *__restrict__ cands, long long ncand, long long j, const Fills &fills,
                                                   const long long *__restrict__ maxend, const long long *__restrict__ mindep_rev)
{
    if (j == 0) return true;
    const unsigned long long ka = (unsigned long long)cands[j - 1].key, kb = (unsigned long long)cands[j].key;
    if (key_pid(ka) != key_pid(kb)) return true;
    if ((maxend && mindep_rev && pl.ext_repeats) ? maxend[j - 1] <= mindep_rev[ncand - 1 - j]
                                                 : anchor_of(pl, (long long)ka) + (pl.m + pl.k) + pl.maxleft + 2 <= anchor_of(pl, (long long)kb)) return true;
    return fill_of(fills, locus_of(pl, anchor_of(pl, (long long)ka))) != fill_of(fills, locus_of(pl, anchor_of(pl, (long long)kb)));
}

// Resolves the cluster that starts at candidate j0; called by all 32 lanes of a warp with the same arguments.
// use_scan: decide cluster borders with the maxend / mindep_rev arrays (patterns with '*' / '+').  Returns the hits chosen.
__device__ unsigned long long chain_cluster_warp(const DevPlan &pl, const unsigned char *__restrict__ text,
                                                 const unsigned long long *__restrict__ TL, const unsigned long long *__restrict__ TR,
                                                 const Cand *__restrict__ cands, long long ncand, long long j0,
                                                 pm_hit *__restrict__ hits, unsigned char *__restrict__ sel, const Fills &fills,
                                                 const long long *__restrict__ maxend, const long long *__restrict__ mindep_rev)
{
    const int lane = threadIdx.x & 31;
    const unsigned below = (1u << lane) - 1u;
    const int cur = fill_of(fills, locus_of(pl, anchor_of(pl, cands[j0].key)));
    const long long n_fill = fills.E[cur];
    long long pos = fills.S[cur];
    unsigned long long nsel = 0;
    bool stop = false;
    for (long long t0 = j0; t0 < ncand && !stop; t0 += 32) {
        const long long t = t0 + lane;
        const bool exists = t < ncand;
        // the cluster ends in front of the first later candidate that opens a new one
        const bool opens = exists && t > j0 && cand_opens_cluster(pl, cands, ncand, t, fills, maxend, mindep_rev);
        const unsigned endm = __ballot_sync(0xffffffffu, opens || !exists);
        const int nin = endm ? __ffs(endm) - 1 : 32;            // lanes 0 .. nin-1 belong to the cluster
        const bool incl = lane < nin;
        Cand c;
        c.key = 0; c.beg = -1; c.end = -1; c.reach = 0;
        long long anchor = 0, p = 0, dl = 0;
        if (incl) {
            c = cands[t];
            sel[t] = 0;
            anchor = anchor_of(pl, c.key);
            p = locus_of(pl, anchor);
            dl = dep_lo(pl, c);
        }
        unsigned todo = incl ? 0xffffffffu : 0u;                 // per lane: all ones while the candidate is not settled
        for (;;) {
            // a failed verification is final unless '^' is in play: then a scan start inside the examined bytes can
            // turn it into a match (the left context is satisfied AT the scan start), so it is redone clipped below
            const bool elig = todo && p >= pos && (c.beg >= 0 || (pl.start_line && pl.type != PM_PLAN_SIMPLE && dl < pos));
            const unsigned em = __ballot_sync(0xffffffffu, elig);
            if (!em) break;
            const int L = __ffs(em) - 1;
            if (lane <= L) todo = 0;                              // everything in front of L is settled under this scan start
            long long b = c.beg, e = c.end;
            int ok = 0;
            if (lane == L) {
                ok = 1;
                if (pl.type == PM_PLAN_SIMPLE) {
                    // simple checkMatch @416790: '^' / '$' look at the byte next to the match unless it touches the scan range
                    if (pl.start_line && anchor > pos && text[anchor - 1] != '\n') ok = 0;
                    if (ok && pl.end_line && e < n_fill && text[e] != '\n') ok = 0;
                } else if (dl < pos) {
                    // the unclipped verification looked left of the new scan start: redo it clipped
                    long long r;
                    if (plan_is_ext(pl)) ok = check_match_ext(pl, text, n_fill, TL, TR, anchor, pos, &b, &e, &r) ? 1 : 0;
                    else ok = check_match(pl, text, n_fill, TL, TR, (int)(c.key & 15), anchor, pos, &b, &e, &r) ? 1 : 0;
                }
                if (ok) { hits[t].beg = b; hits[t].end = e; sel[t] = 1; }
            }
            ok = __shfl_sync(0xffffffffu, ok, L);
            if (ok) {
                const long long ne = __shfl_sync(0xffffffffu, e, L), nb = __shfl_sync(0xffffffffu, b, L);
                nsel++;
                if (ne <= pos && nb == ne) { stop = true; break; }   // zero-length hit: the reference would not advance either
                pos = ne;
            }
        }
        if (nin < 32) break;
    }
    (void)below;
    return nsel;
}

// Short clusters (the usual case of a selective search: the few pieces that fire at one site) are walked by the lane of
// their first candidate, up to CHAIN_SHORT candidates; returns the hits chosen, or -1 when the cluster is longer than
// that (the warp then resolves it from the start, chain_cluster_warp).
#define CHAIN_SHORT 24
__device__ __forceinline__ int chain_cluster_thread(const DevPlan &pl, const unsigned char *__restrict__ text,
                                                    const unsigned long long *__restrict__ TL, const unsigned long long *__restrict__ TR,
                                                    const Cand *__restrict__ cands, long long ncand, long long j0,
                                                    pm_hit *__restrict__ hits, unsigned char *__restrict__ sel, const Fills &fills,
                                                    const long long *__restrict__ maxend, const long long *__restrict__ mindep_rev)
{
    const int cur = fill_of(fills, locus_of(pl, anchor_of(pl, cands[j0].key)));
    const long long n_fill = fills.E[cur];
    long long pos = fills.S[cur];
    int nsel = 0;
    for (long long t = j0; t < ncand; t++) {
        if (t > j0 && cand_opens_cluster(pl, cands, ncand, t, fills, maxend, mindep_rev)) break;
        if (t - j0 >= CHAIN_SHORT) return -1;
        sel[t] = 0;
        const Cand c = cands[t];
        const long long dl = dep_lo(pl, c);
        if (c.beg < 0 && !(pl.start_line && pl.type != PM_PLAN_SIMPLE && dl < pos)) continue;
        const long long anchor = anchor_of(pl, c.key);
        const long long p = locus_of(pl, anchor);
        if (p < pos) continue;
        long long b = c.beg, e = c.end;
        if (pl.type == PM_PLAN_SIMPLE) {
            if (pl.start_line && anchor > pos && text[anchor - 1] != '\n') continue;
            if (pl.end_line && e < n_fill && text[e] != '\n') continue;
        } else if (dl < pos) {
            long long r;
            if (plan_is_ext(pl)) { if (!check_match_ext(pl, text, n_fill, TL, TR, anchor, pos, &b, &e, &r)) continue; }
            else if (!check_match(pl, text, n_fill, TL, TR, (int)(c.key & 15), anchor, pos, &b, &e, &r)) continue;
        }
        hits[t].beg = b;
        hits[t].end = e;
        sel[t] = 1;
        nsel++;
        if (e <= pos && b == e) break;                      // zero-length hit: the reference would not advance either
        pos = e;
    }
    return nsel;
}

// one warp per 32 consecutive candidates: the lanes find the cluster heads among them; short clusters are walked by
// the lane of their head (all in parallel), clusters longer than CHAIN_SHORT by the whole warp, one after the other
__global__ void __launch_bounds__(128) k_chain(const DevPlan pl, const unsigned char *__restrict__ text, long long n,
                                               const unsigned long long *__restrict__ TL, const unsigned long long *__restrict__ TR,
                                               const Cand *__restrict__ cands, long long ncand,
                                               pm_hit *__restrict__ hits, unsigned char *__restrict__ sel, const Fills fills,
                                               const long long *__restrict__ maxend,        // prefix maxima of Cand.end, or null
                                               const long long *__restrict__ mindep_rev)    // prefix minima of dep_lo over the REVERSED list
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool head = j < ncand && cand_opens_cluster(pl, cands, ncand, j, fills, maxend, mindep_rev);
    bool longc = false;
    if (head) longc = chain_cluster_thread(pl, text, TL, TR, cands, ncand, j, hits, sel, fills, maxend, mindep_rev) < 0;
    unsigned heads = __ballot_sync(0xffffffffu, longc);
    const long long wbase = j - lane;
    while (heads) {
        const int h = __ffs(heads) - 1;
        heads &= heads - 1;
        chain_cluster_warp(pl, text, TL, TR, cands, ncand, wbase + h, hits, sel, fills, maxend, mindep_rev);
    }
    (void)n;
}

struct MaxLL { __host__ __device__ long long operator()(long long a, long long b) const { return a > b ? a : b; } };

struct MinLL { __host__ __device__ long long operator()(long long a, long long b) const { return a < b ? a : b; } };

// hit ends in list order and leftmost examined bytes in reversed order (inputs of the two scans of the chain stage)
__global__ void k_cand_ends(const DevPlan pl, const Cand *__restrict__ cands, long long n, long long *__restrict__ ends,
                            long long *__restrict__ deps_rev)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    // with '^' a failed verification can still succeed for a scan start inside the bytes it examined, and where that
    // hit would end is not known: it keeps every later candidate of the fill in the same cluster
    ends[j] = (cands[j].beg < 0 && pl.start_line && cands[j].reach < anchor_of(pl, cands[j].key)) ? (1LL << 60) : cands[j].end;
    deps_rev[n - 1 - j] = dep_lo(pl, cands[j]);
}

// ---------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return PM_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = std::max(bytes, (size_t)1 << 16);
        CK(cudaMalloc(&p, want));
        cap = want;
        return PM_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct pm_engine {
    int device = 0;
    int sms = 148;
    cudaStream_t own = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf keys, keys2, cands, hits, hits2, sel, tables, counters, cubtmp, scanbuf;
    // A pm_engine owns mutable scratch (the buffers above, pinned staging, stats): every entry point that takes an
    // engine holds this lock, so concurrent callers (e.g. mod_wsgi request threads) are serialised per engine.
    std::recursive_mutex mu;
    std::vector<unsigned char> h_blob;              // host staging of the per-request tables
    long long req_cap_hint = 1 << 16;               // candidate capacity of the next request (adapted from the last one)
    long long req_hits_hint = 4096;                 // hits copied speculatively with the header
    const pm_hit *last_hits = nullptr;              // device hit list of the last search (pm_last_hits)
    bool stats_pending = false;                     // device timings not yet read back (done lazily by pm_get_stats)
    unsigned long long *h_count = nullptr;          // pinned
    void *h_stage = nullptr; size_t h_stage_cap = 0; // pinned staging for small result copies
    int scan_mode = 0;                              // 0 auto, 1 byte Shift-And, 2 packed bit-sliced
    long long bufsize = 1600000;                    // patmatch.py:37 MAX_BUFFER_SIZE (-b, in bytes)
    int fused_filter = 1;                           // packed scan drops candidates whose verification surely fails
    int qgram_filter = 1;                           // bit-sliced q-gram pre-filter in front of the Myers filter
    int split_kernel = 1;                           // 1: k_scan_apx (chunk-built pieces + Landau-Vishkin check), 2: k_scan_split (first generation), 0: block-tile kernel
    int pep5 = 1;                                   // non-DNA datasets: scan the 5-bit residue codes (0: raw bytes)
    int jit_mode = 1;                               // 0: never, 1: auto (genome-scale requests), 2: always -- specialised scan kernels (apx_jit.cpp)
    struct BatchCache;                               // host-side index of the last motif batch (search_batch_fused)
    std::shared_ptr<BatchCache> bcache;
    DevBuf batch_tab;                               // its lookup tables on the device
    std::string batch_tab_key;
    int batch_hash = 1;                             // batches of >= 64 exact motifs: q-gram lookup kernel (0: always the dense multi-pattern kernel)
    bool attr_hash = false;
    bool attr_exact = false, attr_split = false, attr_apx = false;   // cudaFuncSetAttribute is per device: kept per engine
    // one spare text buffer and one spare plane buffer, so that re-creating a dataset of the same size
    // (a request that uploads its file every time) does not pay cudaMalloc/cudaFree of gigabytes
    void *pool_text = nullptr; size_t pool_text_cap = 0;
    void *pool_planes = nullptr; size_t pool_planes_cap = 0;
    pm_stats stats{};
};

// words of padding in front of the hi plane: the streaming scan kernels stage tiles together with the words before them
#define PM_PLANE_FRONT 128
struct pm_dataset {
    pm_engine *e = nullptr;
    const unsigned char *d_text = nullptr;
    void *owned = nullptr;
    size_t owned_cap = 0, planes_cap = 0;
    long long n = 0;
    // 2-bit packed planes (packed.cuh)
    unsigned *hi = nullptr, *lo = nullptr, *xx = nullptr;
    unsigned *codes = nullptr;     // 5-bit residue codes (k_pack5): datasets that are not DNA-like
    void *planes_base = nullptr;   // allocation that holds the planes: PM_PLANE_FRONT words of padding, then hi | lo | xx
    long long nwords = 0;
    long long nexc = 0;            // bytes that are not ACGTacgt
    bool dna_like = false;         // few enough exceptions for the packed scan to pay off
    // buffer fills of the reference for the engine's buffer size
    std::vector<long long> newlines;   // sorted positions of '\n'
    long long fills_bufsize = -1;
    std::vector<long long> fill_starts; // host copy of S[] (pm_search_fills_device snaps its range to fills)
    std::vector<long long> fill_ends;   // host copy of E[]
    bool fills_complete = true;         // false while pm_search_stream is still extending the table
    long long *d_fills = nullptr;      // S[0..nfills) then E[0..nfills), then the forced cuts
    size_t fills_cap = 0;              // entries allocated at d_fills
    int nfills = 0;
    int ncuts = 0;                     // fill boundaries that do not fall on a '\n'
    // windowed datasets (pm_dataset_create_window): only the bytes / plane words of [win_lo, win_hi) exist on the device
    bool windowed = false;
    long long win_lo = 0, win_hi = 0;
    unsigned long long *d_wcount = nullptr;   // [0] non-ACGT bytes, [1] newlines of the window (device)

};

static int pack_dataset(pm_engine *e, pm_dataset *d)
{
    const long long n = d->n;
    long long nw = (n + 31) / 32;
    nw = (nw + 1023) / 1024 * 1024 + 1024;          // whole block tiles of the TMA-staged scan + halo
    d->nwords = nw;
    void *p = nullptr;
    const size_t need = ((size_t)nw * 3 + PM_PLANE_FRONT) * 4;
    if (e->pool_planes && e->pool_planes_cap >= need) {
        p = e->pool_planes; d->planes_cap = e->pool_planes_cap;
        e->pool_planes = nullptr; e->pool_planes_cap = 0;
    } else {
        CK(cudaMalloc(&p, need));
        d->planes_cap = need;
    }
    d->planes_base = p;
    d->hi = (unsigned *)p + PM_PLANE_FRONT; d->lo = d->hi + nw; d->xx = d->lo + nw;
    CK(cudaMemsetAsync(p, 0, PM_PLANE_FRONT * 4, e->stream));
    int rc;
    if ((rc = e->counters.reserve(64))) return rc;
    unsigned long long *d_exc = (unsigned long long *)((char *)e->counters.p + 32);
    CK(cudaMemsetAsync(d_exc, 0, 24, e->stream));
    // newline positions are captured by the pack kernel itself (a .seq file has two per sequence);
    // a file with more lines than the buffer holds falls back to a second pass
    const long long nl_cap = 1 << 20;
    if ((rc = e->keys2.reserve((size_t)nl_cap * 8))) return rc;
    unsigned long long *d_nl = (unsigned long long *)e->keys2.p;
    const int grid = (int)std::min<long long>((nw + 255) / 256, (long long)e->sms * 16);
    k_pack<<<std::max(grid, 1), 256, 0, e->stream>>>(d->d_text, n, nw, d->hi, d->lo, d->xx, d_exc, d_nl, nl_cap);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(e->h_count + 4, d_exc, 16, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    d->nexc = (long long)e->h_count[4];
    d->dna_like = n > 0 && d->nexc * 8 <= n;
    if (!d->dna_like && n > 0) {
        // proteomes: 5-bit residue codes for the Shift-And scan
        const long long ncw = (n + 5) / 6 + 16;
        CK(cudaMalloc((void **)&d->codes, (size_t)ncw * 4));
        const int g5 = (int)std::min<long long>((ncw + 255) / 256, (long long)e->sms * 16);
        k_pack5<<<std::max(g5, 1), 256, 0, e->stream>>>(d->d_text, n, ncw, d->codes);
        CK(cudaGetLastError());
    }
    const long long nnl = (long long)e->h_count[5];
    d->newlines.resize((size_t)nnl);
    if (nnl > 0 && nnl <= nl_cap) {
        CK(cudaMemcpyAsync(d->newlines.data(), d_nl, (size_t)nnl * 8, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        std::sort(d->newlines.begin(), d->newlines.end());
    } else if (nnl > nl_cap) {
        void *dn = nullptr;
        CK(cudaMalloc(&dn, (size_t)nnl * 8));
        const int g2 = (int)std::min<long long>((n / 16 + 255) / 256 + 1, (long long)e->sms * 16);
        k_newlines<<<g2, 256, 0, e->stream>>>(d->d_text, n, (unsigned long long *)dn, d_exc + 2);
        cudaError_t rc2 = cudaGetLastError();
        if (rc2 == cudaSuccess) rc2 = cudaMemcpyAsync(d->newlines.data(), dn, (size_t)nnl * 8, cudaMemcpyDeviceToHost, e->stream);
        if (rc2 == cudaSuccess) rc2 = cudaStreamSynchronize(e->stream);
        cudaFree(dn);
        if (rc2 != cudaSuccess) { g_err = std::string("newline index: ") + cudaGetErrorString(rc2); return PM_ERR_CUDA; }
        std::sort(d->newlines.begin(), d->newlines.end());
    }
    return PM_OK;
}

// Buffer fills [S_k, E_k) exactly as the reference produces them (bufSetFile @41bbc0, bufLoad @41bbf0,
// recSearchFile @402298-4024f0): a fill of `bufsize` bytes that does not reach EOF is scanned up to and
// including its last '\n' and the next fill starts AT that '\n'; without a usable '\n' the fill is scanned
// whole ("Record longer than buffer size ... has been split") and the next one starts right after it.
// `avail` = number of file bytes whose newlines are known; with complete = false only the fills that are already
// determined (s0 + bufsize <= avail) are produced -- the streaming upload searches them while the rest arrives.
static void compute_fills(const std::vector<long long> &newlines, long long n, long long bufsize, long long avail, bool complete,
                          std::vector<long long> &S, std::vector<long long> &E, std::vector<long long> &cuts)
{
    const long long bs = bufsize > 0 ? bufsize : n + 1;
    long long s0 = 0;
    while (n - s0 > 0) {
        const long long dsize = std::min(bs, n - s0);
        if (!complete && s0 + dsize > avail) break;
        long long en, next;
        if (dsize < bs) { en = s0 + dsize; next = n; }
        else {
            // last newline p with s0 < p <= s0 + dsize - 1
            auto it = std::upper_bound(newlines.begin(), newlines.end(), s0 + dsize - 1);
            long long p = -1;
            if (it != newlines.begin()) p = *(it - 1);
            if (p > s0) { en = p + 1; next = p; }
            else { en = s0 + dsize; next = s0 + dsize; if (next < n) cuts.push_back(next); }
        }
        S.push_back(s0); E.push_back(en);
        s0 = next;
    }
    if (complete && S.empty()) { S.push_back(0); E.push_back(0); }
}

static int upload_fills(pm_engine *e, pm_dataset *d, const std::vector<long long> &S, const std::vector<long long> &E,
                        const std::vector<long long> &cuts)
{
    // S | E | cuts, packed; the buffer is only reallocated when it has to grow (cudaFree synchronises the whole
    // device, which would serialise the streaming upload with its own copies)
    const size_t need = S.size() * 2 + cuts.size() + 1;
    if (need > d->fills_cap) {
        if (d->d_fills) { cudaFree(d->d_fills); d->d_fills = nullptr; d->fills_cap = 0; }
        const size_t want = std::max(need * 2, (size_t)4096);
        CK(cudaMalloc((void **)&d->d_fills, want * 8));
        d->fills_cap = want;
    }
    if (!S.empty()) {
        CK(cudaMemcpyAsync(d->d_fills, S.data(), S.size() * 8, cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemcpyAsync(d->d_fills + S.size(), E.data(), S.size() * 8, cudaMemcpyHostToDevice, e->stream));
    }
    if (!cuts.empty())
        CK(cudaMemcpyAsync(d->d_fills + 2 * S.size(), cuts.data(), cuts.size() * 8, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    d->nfills = (int)S.size();
    d->ncuts = (int)cuts.size();
    d->fill_starts = S;
    d->fill_ends = E;
    return PM_OK;
}

static int ensure_fills(pm_engine *e, pm_dataset *d, Fills *out)
{
    if (d->fills_bufsize != e->bufsize) {
        std::vector<long long> S, E, cuts;
        compute_fills(d->newlines, d->n, e->bufsize, d->n, true, S, E, cuts);
        int rc = upload_fills(e, d, S, E, cuts);
        if (rc) return rc;
        d->fills_bufsize = e->bufsize;
    }
    out->S = d->d_fills; out->E = d->d_fills + d->nfills; out->n = d->nfills;
    return PM_OK;
}

int pm_engine_create(int device, pm_engine **out)
{
    if (!out) { g_err = "out is NULL"; return PM_ERR_ARG; }
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { g_err = "no such CUDA device (this engine has no CPU fallback)"; return PM_ERR_CUDA; }
    CK(cudaSetDevice(device));
    pm_engine *e = new pm_engine();
    e->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    e->sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&e->own, cudaStreamNonBlocking));
    e->stream = e->own;
    for (auto &ev : e->ev) CK(cudaEventCreate(&ev));
    CK(cudaMallocHost((void **)&e->h_count, 64));
    if (const char *j = getenv("PM_JIT")) { const int m = atoi(j); if (m >= 0 && m <= 2) e->jit_mode = m; }   // test hook: force / forbid the specialised kernels
    *out = e;
    return PM_OK;
}

void pm_engine_destroy(pm_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    if (e->pool_text) cudaFree(e->pool_text);
    if (e->pool_planes) cudaFree(e->pool_planes);
    for (DevBuf *b : {&e->keys, &e->keys2, &e->cands, &e->hits, &e->hits2, &e->sel, &e->tables, &e->counters, &e->cubtmp, &e->scanbuf, &e->batch_tab}) b->release();
    for (auto &ev : e->ev) if (ev) cudaEventDestroy(ev);
    if (e->h_count) cudaFreeHost(e->h_count);
    if (e->h_stage) cudaFreeHost(e->h_stage);
    if (e->own) cudaStreamDestroy(e->own);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    delete e;
}

int pm_engine_set_stream(pm_engine *e, void *s)
{
    if (!e) { g_err = "engine is NULL"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    e->stream = s ? (cudaStream_t)s : e->own;
    return PM_OK;
}

int pm_engine_set_scan_mode(pm_engine *e, int mode)
{
    if (!e || mode < 0 || mode > 2) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    e->scan_mode = mode;
    return PM_OK;
}

int pm_engine_set_fused_filter(pm_engine *e, int on)
{
    if (!e) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    e->fused_filter = on ? 1 : 0;
    e->qgram_filter = (on == 1 || on == 4) ? 1 : 0;
    e->split_kernel = on == 3 ? 0 : (on == 4 || on == 5) ? 2 : 1;   // 4 / 5: first-generation k_scan_split with / without the q-gram count
    return PM_OK;
}

int pm_engine_set_batch_lookup(pm_engine *e, int on)
{
    if (!e) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    e->batch_hash = on ? 1 : 0;
    return PM_OK;
}

int pm_engine_set_peptide_codes(pm_engine *e, int on)
{
    if (!e) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    e->pep5 = on ? 1 : 0;
    return PM_OK;
}

int pm_engine_set_jit(pm_engine *e, int mode)
{
    if (!e || mode < 0 || mode > 2) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    e->jit_mode = mode;
    return PM_OK;
}

int pm_engine_set_buffer_size(pm_engine *e, int64_t bytes)
{
    if (!e || bytes < 0) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    e->bufsize = bytes;
    return PM_OK;
}

int pm_engine_synchronize(pm_engine *e)
{
    if (!e) { g_err = "engine is NULL"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    return PM_OK;
}

int pm_dataset_create(pm_engine *e, const uint8_t *host, int64_t n, pm_dataset **out)
{
    if (!e || !out || n < 0 || (!host && n > 0)) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    CK(cudaSetDevice(e->device));
    if (n >= (1LL << PM_POS_BITS) - 4096) { g_err = "dataset too large (candidate keys hold 36-bit positions)"; return PM_ERR_UNSUPPORTED; }
    pm_dataset *d = new pm_dataset();
    d->e = e; d->n = n;
    void *p = nullptr;
    cudaError_t rc = cudaSuccess;
    if (e->pool_text && e->pool_text_cap >= (size_t)n + 256) {
        p = e->pool_text; d->owned_cap = e->pool_text_cap;
        e->pool_text = nullptr; e->pool_text_cap = 0;
    } else {
        rc = cudaMalloc(&p, (size_t)n + 256);
        d->owned_cap = (size_t)n + 256;
    }
    if (rc != cudaSuccess) { delete d; g_err = std::string("cudaMalloc dataset: ") + cudaGetErrorString(rc); return PM_ERR_CUDA; }
    d->owned = p; d->d_text = (const unsigned char *)p;
    if (n > 0) {
        rc = cudaMemcpyAsync(p, host, (size_t)n, cudaMemcpyHostToDevice, e->stream);
        if (rc == cudaSuccess) rc = cudaMemsetAsync((char *)p + n, 0, 256, e->stream);
        if (rc == cudaSuccess) rc = cudaStreamSynchronize(e->stream);
        if (rc != cudaSuccess) { cudaFree(p); delete d; g_err = std::string("dataset upload: ") + cudaGetErrorString(rc); return PM_ERR_CUDA; }
    }
    int prc = pack_dataset(e, d);
    if (prc) { pm_dataset_destroy(d); return prc; }
    *out = d;
    return PM_OK;
}

int pm_dataset_wrap_device(pm_engine *e, const uint8_t *dev, int64_t n, pm_dataset **out)
{
    if (!e || !out || n < 0 || (!dev && n > 0)) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    if (n >= (1LL << PM_POS_BITS) - 4096) { g_err = "dataset too large (candidate keys hold 36-bit positions)"; return PM_ERR_UNSUPPORTED; }
    pm_dataset *d = new pm_dataset();
    d->e = e; d->n = n; d->d_text = dev; d->owned = nullptr;
    CK(cudaSetDevice(e->device));
    int prc = pack_dataset(e, d);
    if (prc) { pm_dataset_destroy(d); return prc; }
    *out = d;
    return PM_OK;
}

// Windowed dataset for multi-GPU cold requests: text and planes are allocated for the whole file, so positions stay
// absolute everywhere, but only the window is copied and packed.  No host synchronisation.
int pm_dataset_create_window(pm_engine *e, const uint8_t *host, int64_t n, int64_t win_lo, int64_t win_hi, void *dev_newlines, int64_t nl_rows,
                             pm_dataset **out)
{
    if (!e || !out || !host || n <= 0 || win_lo < 0 || win_hi > n || win_lo >= win_hi || !dev_newlines || nl_rows < 2) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    if (n >= (1LL << PM_POS_BITS) - 4096) { g_err = "dataset too large (candidate keys hold 36-bit positions)"; return PM_ERR_UNSUPPORTED; }
    CK(cudaSetDevice(e->device));
    pm_dataset *d = new pm_dataset();
    d->e = e; d->n = n;
    auto fail = [&](int rc) { pm_dataset_destroy(d); return rc; };
#define CKW(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { g_err = std::string(#x) + " (engine.cu:" + std::to_string(__LINE__) + "): " + cudaGetErrorString(e_); return fail(PM_ERR_CUDA); } } while (0)
    {
        void *p = nullptr;
        if (e->pool_text && e->pool_text_cap >= (size_t)n + 256) {
            p = e->pool_text; d->owned_cap = e->pool_text_cap;
            e->pool_text = nullptr; e->pool_text_cap = 0;
        } else {
            CKW(cudaMalloc(&p, (size_t)n + 256));
            d->owned_cap = (size_t)n + 256;
        }
        d->owned = p; d->d_text = (const unsigned char *)p;
    }
    long long nw = (n + 31) / 32;
    nw = (nw + 1023) / 1024 * 1024 + 1024;
    d->nwords = nw;
    {
        void *p = nullptr;
        const size_t need = ((size_t)nw * 3 + PM_PLANE_FRONT) * 4;
        if (e->pool_planes && e->pool_planes_cap >= need) {
            p = e->pool_planes; d->planes_cap = e->pool_planes_cap;
            e->pool_planes = nullptr; e->pool_planes_cap = 0;
        } else {
            CKW(cudaMalloc(&p, need));
            d->planes_cap = need;
        }
        d->planes_base = p;
        d->hi = (unsigned *)p + PM_PLANE_FRONT; d->lo = d->hi + nw; d->xx = d->lo + nw;
        CKW(cudaMemsetAsync(p, 0, PM_PLANE_FRONT * 4, e->stream));
    }
    // whole words: [lo32, hi32)
    const long long lo32 = win_lo & ~31LL, hi32 = std::min<long long>((win_hi + 31) & ~31LL, n);
    d->windowed = true; d->win_lo = lo32; d->win_hi = hi32;
    CKW(cudaMemcpyAsync((char *)d->owned + lo32, host + lo32, (size_t)(hi32 - lo32), cudaMemcpyHostToDevice, e->stream));
    if (hi32 == n) CKW(cudaMemsetAsync((char *)d->owned + n, 0, 256, e->stream));
    CKW(cudaMalloc((void **)&d->d_wcount, 32));
    CKW(cudaMemsetAsync(d->d_wcount, 0, 32, e->stream));
    {
        const long long q0 = lo32 / 32, q1 = hi32 == n ? std::min<long long>((n + 31) / 32 + 8, nw) : hi32 / 32;
        const int grid = (int)std::min<long long>((q1 - q0 + 255) / 256, (long long)e->sms * 16);
        k_pack<<<std::max(grid, 1), 256, 0, e->stream>>>(d->d_text + lo32, n - lo32, q1 - q0, d->hi + q0, d->lo + q0, d->xx + q0, d->d_wcount,
                                                         (unsigned long long *)dev_newlines + 1, nl_rows - 1, lo32);
        CKW(cudaGetLastError());
        CKW(cudaMemcpyAsync(dev_newlines, d->d_wcount + 1, 8, cudaMemcpyDeviceToDevice, e->stream));
    }
#undef CKW
    d->fills_complete = false;
    *out = d;
    return PM_OK;
}

// installs the newline index of the WHOLE file (sorted positions), e.g. gathered from the ranks' windows; the buffer
// fills follow from it.  Synchronises the engine's stream once to learn how DNA-like the window is.
int pm_dataset_set_newlines(pm_engine *e, pm_dataset *d, const int64_t *pos, int64_t count)
{
    if (!e || !d || d->e != e || count < 0 || (count > 0 && !pos)) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    CK(cudaSetDevice(e->device));
    d->newlines.assign(pos, pos + count);
    if (!std::is_sorted(d->newlines.begin(), d->newlines.end())) std::sort(d->newlines.begin(), d->newlines.end());
    d->fills_bufsize = -1;
    d->fills_complete = true;
    if (d->windowed && d->d_wcount) {
        CK(cudaMemcpyAsync(e->h_count + 4, d->d_wcount, 16, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        d->nexc = (long long)e->h_count[4];
        d->dna_like = d->nexc * 8 <= (d->win_hi - d->win_lo);
    }
    return PM_OK;
}

void pm_dataset_destroy(pm_dataset *d)
{
    if (!d) return;
    pm_engine *e = d->e;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    if (d->owned) {
        if (!e->pool_text || e->pool_text_cap < d->owned_cap) {
            if (e->pool_text) cudaFree(e->pool_text);
            e->pool_text = d->owned; e->pool_text_cap = d->owned_cap;
        } else cudaFree(d->owned);
    }
    if (d->hi) {
        if (!e->pool_planes || e->pool_planes_cap < d->planes_cap) {
            if (e->pool_planes) cudaFree(e->pool_planes);
            e->pool_planes = d->planes_base; e->pool_planes_cap = d->planes_cap;
        } else cudaFree(d->planes_base);
    }
    if (d->d_fills) cudaFree(d->d_fills);
    if (d->codes) cudaFree(d->codes);
    if (d->d_wcount) cudaFree(d->d_wcount);
    delete d;
}

int64_t pm_dataset_size(const pm_dataset *d) { return d ? d->n : 0; }

// page-locked host memory for large result arrays (device-to-host copies into pageable memory run at a
// fraction of the PCIe rate)
void *pm_host_alloc(int64_t bytes)
{
    void *p = nullptr;
    if (bytes <= 0) return nullptr;
    if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) { g_err = "cudaMallocHost failed"; (void)cudaGetLastError(); return nullptr; }
    return p;
}

void pm_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

static void finish_stats(pm_engine *e);

int pm_get_stats(pm_engine *e, pm_stats *out)
{
    if (!e || !out) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    if (e->stats_pending) {                              // asynchronous request: the events are read here
        CK(cudaSetDevice(e->device));
        CK(cudaStreamSynchronize(e->stream));
        finish_stats(e);
    }
    *out = e->stats;
    return PM_OK;
}

// ---------------------------------------------------------------------------------------
struct Compiled {
    pm::Pattern P;
    pm::Options o;
    pm::Plan plan;
    DevPlan dp;
    pm::FilterTables ft;
    pm::VerifyTables vt;
    // EXTENDED plans: the exact scan runs on the window of plain positions around the anchor, compiled as a SIMPLE pattern
    std::shared_ptr<Compiled> scan;
    std::string key;                  // (pattern, -k) text this was compiled from: key of the per-process caches
};

static int compile_uncached(const char *pattern, const char *kopt, Compiled &c, bool need_tables);

// Parsing, the reference's cost model (esimplePreproc @415540) and the verification tables are pure functions of
// (pattern, -k): the results are kept per process, so that a repeated request costs a map lookup and a copy.
static std::mutex g_compile_mu;
static std::map<std::string, std::shared_ptr<const Compiled>> g_compile_cache;

static int compile(const char *pattern, const char *kopt, Compiled &c, bool need_tables)
{
    std::string key = std::string(need_tables ? "T" : "P") + (pm::compat_deployed() ? "D" : "Z") + kopt + '\x01' + pattern;
    {
        std::lock_guard<std::mutex> lock(g_compile_mu);
        auto it = g_compile_cache.find(key);
        if (it != g_compile_cache.end()) { c = *it->second; return PM_OK; }
    }
    int rc = compile_uncached(pattern, kopt, c, need_tables);
    if (rc) return rc;
    c.key = key;
    std::lock_guard<std::mutex> lock(g_compile_mu);
    if (g_compile_cache.size() >= 4096) g_compile_cache.clear();
    g_compile_cache[key] = std::make_shared<const Compiled>(c);
    return PM_OK;
}

static int compile_uncached(const char *pattern, const char *kopt, Compiled &c, bool need_tables)
{
    std::string err;
    int rc = pm::parse_kopt(kopt, c.o, err);
    if (rc) { g_err = err; return rc; }
    rc = pm::parse_pattern(pattern, /*icase: patmatch.py always passes -i*/ true, c.P, err);
    if (rc) { g_err = err; return rc; }
    rc = pm::make_plan(c.P, c.o, c.plan, err);
    if (rc) { g_err = err; return rc; }
    if (!need_tables) return PM_OK;
    if (c.plan.type == pm::SIMPLE && c.P.m() == 1 && c.P.pos[0].has('\n')) {
        // a buffer fill that ends at a '\n' shares that byte with the next one, and the reference reports a one-byte
        // hit on it once per fill; candidates here are keyed by position and belong to one fill
        g_err = "single-position pattern that matches the record delimiter: not supported";
        return PM_ERR_UNSUPPORTED;
    }
    const bool long_simple = c.plan.type == pm::SIMPLE && c.P.m() > 64 && c.P.m() <= 255;
    if (c.P.m() > 255 || (c.P.m() > 64 && c.P.extended())) {
        g_err = "patterns longer than 255 positions (64 with ? * +) are not supported";
        return PM_ERR_UNSUPPORTED;
    }
    DevPlan &d = c.dp;
    memset(&d, 0, sizeof d);
    d.type = c.plan.type; d.m = c.plan.m; d.k = c.plan.k; d.L = c.plan.L; d.npieces = c.plan.npieces;
    d.ins = c.plan.ins; d.del = c.plan.del; d.subs = c.plan.subs;
    d.start_line = c.P.start_line ? 1 : 0; d.end_line = c.P.end_line ? 1 : 0;
    for (int i = 0; i < PM_MAX_PIECES; i++) { d.V[i] = c.plan.V[i]; d.trig[i] = c.plan.trig[i]; }
    if (long_simple) {
        // exact pattern of 65..255 positions: the scan kernels look for its first 64 positions, k_verify compares the
        // rest on the raw bytes (the classes of positions 64.. travel in the TL table slot, 4 words per position)
        auto s = std::make_shared<Compiled>();
        s->o = c.o;
        for (int j = 0; j < 64; j++) { s->P.pos.push_back(c.P.pos[j]); s->P.op.push_back(pm::OP_NONE); }
        s->plan.type = pm::SIMPLE; s->plan.m = s->plan.L = 64; s->plan.npieces = 1;
        DevPlan &sd = s->dp;
        memset(&sd, 0, sizeof sd);
        sd.type = PM_PLAN_SIMPLE; sd.m = sd.L = 64; sd.npieces = 1; sd.ins = sd.del = sd.subs = 1;
        pm::build_filter(s->P, s->plan, s->ft);
        sd.init = s->ft.init; sd.fin = s->ft.fin; sd.trig[0] = s->ft.fin;
        c.ft = s->ft;
        d.L = 64; d.init = c.ft.init; d.fin = c.ft.fin; d.trig[0] = c.ft.fin;
        const int tail = c.P.m() - 64;
        c.vt.TL.assign((size_t)tail * 4, 0);
        c.vt.TR.assign((size_t)tail * 4, 0);
        for (int j = 0; j < tail; j++)
            for (int w = 0; w < 4; w++) c.vt.TL[(size_t)j * 4 + w] = c.P.pos[64 + j].w[w];
        c.scan = s;
        return PM_OK;
    }
    if (c.plan.type == pm::EXT_BEG || c.plan.type == pm::EXT_END) {
        auto s = std::make_shared<Compiled>();
        s->o = c.o;
        for (int j = c.plan.win_lo; j < c.plan.win_hi; j++) { s->P.pos.push_back(c.P.pos[j]); s->P.op.push_back(pm::OP_NONE); }
        s->plan.type = pm::SIMPLE; s->plan.m = s->plan.L = s->P.m(); s->plan.npieces = 1;
        DevPlan &sd = s->dp;
        memset(&sd, 0, sizeof sd);
        sd.type = PM_PLAN_SIMPLE; sd.m = sd.L = s->P.m(); sd.npieces = 1; sd.ins = sd.del = sd.subs = 1;
        pm::build_filter(s->P, s->plan, s->ft);
        sd.init = s->ft.init; sd.fin = s->ft.fin; sd.trig[0] = s->ft.fin;
        c.ft = s->ft;                                       // k_verify re-checks the window on the raw bytes with it
        d.init = c.ft.init; d.fin = c.ft.fin; d.trig[0] = c.ft.fin;
        d.ext_off = c.plan.anchor - c.plan.win_lo; d.ext_anchor = c.plan.anchor; d.ext_repeats = c.plan.ext_repeats;
        d.maxleft = c.plan.anchor;
        d.IL = c.plan.IL; d.FL = c.plan.FL; d.AL = c.plan.AL; d.initL = c.plan.initL;
        d.IR = c.plan.IR; d.FR = c.plan.FR; d.AR = c.plan.AR; d.initR = c.plan.initR;
        d.ext_lead_opt = c.plan.ext_lead_opt; d.IS = c.plan.IS; d.FS = c.plan.FS; d.AS = c.plan.AS;
        pm::build_verify(c.P, c.plan, c.vt);               // TL: positions anchor-1, anchor-2, ... ; TR: anchor, anchor+1, ...
        c.scan = s;
        return PM_OK;
    }
    if (c.plan.type == pm::SIMPLE || c.plan.type == pm::SPLIT) {
        pm::build_filter(c.P, c.plan, c.ft);
        d.init = c.ft.init; d.fin = c.ft.fin;
        if (c.plan.type == pm::SIMPLE) d.trig[0] = c.ft.fin;
    }
    if (c.plan.type != pm::SIMPLE) {
        int vmax = 0;
        for (int i = 0; i < c.plan.npieces; i++) vmax = std::max(vmax, c.plan.V[i]);
        d.maxleft = vmax + c.plan.k;                        // the left walk reads at most V + k bytes
    }
    if (c.plan.type != pm::SIMPLE) pm::build_verify(c.P, c.plan, c.vt);
    return PM_OK;
}

int pm_set_compat_deployed_glibc(int on)
{
    pm::set_compat_deployed(on);
    return PM_OK;
}

int pm_plan(const char *pattern, const char *kopt, pm_plan_info *info)
{
    if (!pattern || !kopt || !info) { g_err = "bad argument"; return PM_ERR_ARG; }
    Compiled c;
    int rc = compile(pattern, kopt, c, false);
    if (rc) return rc;
    memset(info, 0, sizeof *info);
    info->m = c.plan.m; info->k = c.plan.k; info->ins = c.plan.ins; info->del = c.plan.del; info->subs = c.plan.subs;
    info->type = c.plan.type; info->L = c.plan.L; info->npieces = c.plan.npieces;
    for (int i = 0; i < PM_MAX_PIECES; i++) info->V[i] = c.plan.V[i];
    info->split_cost = c.plan.split_cost; info->fb_cost = c.plan.fb_cost;
    return PM_OK;
}

// upload B / TL / TR into e->tables; returns device pointers
static int upload_tables(pm_engine *e, const Compiled &c, const unsigned long long **dB, const unsigned long long **dTL,
                         const unsigned long long **dTR)
{
    const size_t nb = 256, nv = c.vt.TL.size();
    int rc = e->tables.reserve((nb + 2 * nv) * 8 + 64);
    if (rc) return rc;
    unsigned long long *base = (unsigned long long *)e->tables.p;
    CK(cudaMemcpyAsync(base, c.ft.B, nb * 8, cudaMemcpyHostToDevice, e->stream));
    if (nv) {
        CK(cudaMemcpyAsync(base + nb, c.vt.TL.data(), nv * 8, cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemcpyAsync(base + nb + nv, c.vt.TR.data(), nv * 8, cudaMemcpyHostToDevice, e->stream));
    }
    *dB = base; *dTL = base + nb; *dTR = base + nb + nv;
    return PM_OK;
}

static unsigned packed_class_of(const pm::ByteSet &bs, bool *mixed);
static inline int plane_of_class(unsigned cls) { return cls == 1 ? 0 : cls == 2 ? 1 : cls == 4 ? 2 : cls == 8 ? 3 : cls == 16 ? 4 : 5; }

// where the scan kernels append their candidate keys
struct ScanTarget { unsigned long long *keys, *count; long long cap; };

// does the scan of this compiled pattern run on the 2-bit planes?
static bool scan_uses_packed(const pm_engine *e, const pm_dataset *d, const Compiled &c_full)
{
    const Compiled &c = c_full.scan ? *c_full.scan : c_full;
    const bool packable = (c.dp.type == PM_PLAN_SIMPLE || c.dp.type == PM_PLAN_SPLIT) && c.dp.npieces <= 4 && d->hi != nullptr;
    return packable && (e->scan_mode == 2 || (e->scan_mode == 0 && d->dna_like));
}

// does the byte Shift-And scan of this pattern read the 5-bit residue codes?  (its candidates are re-checked on the raw bytes)
static bool scan_uses_pep5(const pm_engine *e, const pm_dataset *d, const Compiled &c_full)
{
    const Compiled &c = c_full.scan ? *c_full.scan : c_full;
    return d->codes != nullptr && e->pep5 && e->scan_mode != 2 && !scan_uses_packed(e, d, c_full) &&
           (c.dp.type == PM_PLAN_SIMPLE || c.dp.type == PM_PLAN_SPLIT);
}

// ---- exact scan (k_scan_packed_exact): up to EX_MAXPAT patterns per pass over the planes ----
static void fill_exact_pat(const Compiled &c, long long a0, long long a1, long long n, unsigned long long tag, ExactPat &pt)
{
    memset(&pt, 0, sizeof pt);
    const DevPlan &dp = c.dp;
    pt.a0 = a0; pt.a1 = std::min(a1, n - dp.L + 1); pt.keytag = tag; pt.L = dp.L;
    for (int j = 0; j < dp.L; j++) {
        const unsigned cls = packed_class_of(c.P.pos[j], nullptr);
        if (cls == 31u) continue;                 // accepts every byte: no constraint (the window bound is checked per hit)
        const int s = plane_of_class(cls);
        if (s == 5) pt.cls[pt.npos[5]] = (unsigned char)cls;
        pt.shift[s][pt.npos[s]++] = (unsigned char)j;
    }
}

static bool jit_wanted(const pm_engine *e, long long bases);
static int launch_apx_jit(pm_engine *e, const ApxArgs &a, long long lo, long long hi, bool exact);

static int launch_exact(pm_engine *e, pm_dataset *d, const ExactPat *pats, int npat, unsigned long long bad, const ScanTarget &t)
{
    ExactArgs a;
    memset(&a, 0, sizeof a);
    long long lo = -1, hi = -1;
    bool longp = false;
    for (int p = 0; p < npat; p++) {
        if (pats[p].a1 <= pats[p].a0) continue;
        a.pat[a.npat++] = pats[p];
        lo = lo < 0 ? pats[p].a0 : std::min(lo, pats[p].a0);
        hi = std::max(hi, pats[p].a1);
        longp = longp || pats[p].L > 32;
    }
    if (a.npat == 0) return PM_OK;
    a.hi = d->hi; a.lo = d->lo; a.xx = d->xx; a.nwords = d->nwords; a.n = d->n;
    a.tile0 = (lo / 32) / 128;
    a.ntiles = ((hi - 1) / 32) / 128 + 1 - a.tile0;
    a.bad = bad;
    a.keys = t.keys; a.count = t.count; a.cap = t.cap;
    // genome-scale requests: the streaming kernel compiled for exactly these patterns (apx_jit.cpp, exact mode)
    if (jit_wanted(e, hi - lo)) {
        ApxArgs ja;
        memset(&ja, 0, sizeof ja);
        ja.hi = d->hi; ja.lo = d->lo; ja.xx = d->xx; ja.nwords = d->nwords; ja.n = d->n;
        ja.keys = t.keys; ja.count = t.count; ja.cap = t.cap;
        ja.npat = a.npat;
        bool fits = true;
        for (int p = 0; p < a.npat; p++) {
            const ExactPat &xp = a.pat[p];
            ApxPat &ap = ja.pat[p];
            ap.a0 = xp.a0; ap.a1 = xp.a1; ap.keytag = xp.keytag;
            ap.m = xp.L; ap.k = 0; ap.L = xp.L; ap.npieces = 1; ap.indel = 0; ap.win = 1;
            int nd = 0;
            static const unsigned char plane_cls[5] = {1, 2, 4, 8, 16};
            for (int sidx = 0; sidx < 6; sidx++)
                for (int q = 0; q < xp.npos[sidx]; q++) {
                    if (nd >= AX_MAXDENSE) { fits = false; break; }
                    ap.dshift[nd] = xp.shift[sidx][q];                       // k + j with k = 0
                    ap.dcls[nd] = sidx < 5 ? plane_cls[sidx] : xp.cls[q];
                    nd++;
                }
            ap.dn[0] = (unsigned char)nd;
            if (nd == 0) ap.dwild[0] = 1;
        }
        if (fits) {
            int rc = launch_apx_jit(e, ja, lo, hi, true);
            if (rc == PM_OK) {
                e->stats.scan_bases += (hi - lo) * a.npat;
                e->stats.packed = 1;
                return PM_OK;
            }
            if (e->jit_mode == 2 || rc != PM_ERR_UNSUPPORTED) return rc;   // auto mode: no NVRTC on this machine -> generic kernel
        }
    }
    const size_t smem = EX_STAGES * EX_STAGE_BYTES + 2 * EX_STAGES * 8;
    if (!e->attr_exact) {
        CK(cudaFuncSetAttribute(k_scan_packed_exact<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_scan_packed_exact<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        e->attr_exact = true;
    }
    const long long nbt = (a.ntiles + 7) / 8;
    const int grid_ex = std::max((int)std::min<long long>(nbt, (long long)e->sms * 5), 1);
    if (longp) k_scan_packed_exact<true><<<grid_ex, (EX_WARPS + 1) * 32, smem, e->stream>>>(a);
    else k_scan_packed_exact<false><<<grid_ex, (EX_WARPS + 1) * 32, smem, e->stream>>>(a);
    CK(cudaGetLastError());
    e->stats.launches++;
    e->stats.scan_bytes += a.ntiles * 128 * 4 * 3;
    e->stats.scan_bases += (hi - lo) * a.npat;
    e->stats.packed = 1;
    return PM_OK;
}

// ---- approximate scan, second generation (k_scan_apx) ----
static bool plain_triggers(const DevPlan &dp)
{
    for (int i = 0; i < dp.npieces; i++)
        if (dp.trig[i] != (1ULL << (i * dp.L + dp.L - 1))) return false;
    return true;
}

static bool apx_eligible(const pm_engine *e, const pm_dataset *d, const Compiled &c)
{
    const DevPlan &dp = c.dp;
    return !c.scan && dp.type == PM_PLAN_SPLIT && scan_uses_packed(e, d, c) && e->fused_filter && e->split_kernel == 1 &&
           dp.k >= 1 && dp.k <= 3 && dp.m + 2 * dp.k <= 64 && plain_triggers(dp);
}

// Chunk selection for k_scan_apx.  Candidate chunk sets: (A) every piece cut into 1..3 sub-chunks of about equal
// information (then piece planes are ANDs of chunk planes and every pattern position is evaluated once), (B) uniform
// tilings and equal-information segmentations that ignore the piece borders (pieces evaluated position by position),
// (C) no q-gram count at all.  Cost in warp instructions per 8192-base warp tile: dense work + expected pattern starts
// that reach the per-anchor check.
// Cost constants of the two kernels that execute an ApxPat, in warp instructions per 8192-base warp tile.
struct ApxCost {
    double pos;            // one constrained position ANDed into its chunk plane (shift + AND)
    double cls;            // one class plane (per run of equal classes inside a chunk)
    double derive;         // chunk plane shifted to the nominal diagonal and ANDed into its piece
    double dense_pos;      // one position of a piece evaluated directly at the nominal diagonal
    double dense_piece;    // per piece evaluated directly
    double dil_step;       // one dilation step (generic kernel: doubling; specialised: see dil[])
    double dil[4];         // specialised kernel: whole dilation for win = 1, 3, 5, 7
    double count_fix, count_row;
    double survivor;       // one pattern start that reaches the sparse stage (extraction + 1/32 Landau-Vishkin round)
};
static const ApxCost kCostGeneric = {45.0, 0.0, 18.0, 43.0, 12.0, 20.0, {0, 0, 0, 0}, 26.0, 16.0, 0.0};
static const ApxCost kCostJit = {13.5, 9.0, 12.0, 13.5, 4.0, 0.0, {8.0, 24.0, 43.0, 51.0}, 0.0, 8.0, 28.0};

static void build_apx_pat_uncached(bool qgram, bool jit, const Compiled &c, ApxPat &out);

static std::mutex g_apx_mu;
static std::map<std::string, ApxPat> g_apx_cache;

static void build_apx_pat(bool qgram, bool jit, const Compiled &c, long long a0, long long a1, unsigned long long tag, ApxPat &out)
{
    const std::string key = std::string(qgram ? "Q" : "N") + (jit ? "J" : "G") + c.key;
    bool hit = false;
    if (!c.key.empty()) {
        std::lock_guard<std::mutex> lock(g_apx_mu);
        auto it = g_apx_cache.find(key);
        if (it != g_apx_cache.end()) { out = it->second; hit = true; }
    }
    if (!hit) {
        build_apx_pat_uncached(qgram, jit, c, out);
        if (!c.key.empty()) {
            std::lock_guard<std::mutex> lock(g_apx_mu);
            if (g_apx_cache.size() >= 4096) g_apx_cache.clear();
            g_apx_cache[key] = out;
        }
    }
    out.a0 = a0; out.a1 = a1; out.keytag = tag;
}

// Chunk selection.  Candidate chunk sets: (A) every piece cut into 1..3 sub-chunks of about equal information and the
// stretches between / around the pieces cut into 1..2 chunks (then piece planes are ANDs of chunk planes and every
// pattern position is evaluated once), (B) uniform tilings and equal-information segmentations that ignore the piece
// borders (pieces evaluated position by position), (C) no q-gram count at all.  Cost: dense work + expected pattern
// starts that reach the per-anchor check, with the constants of the kernel that will run the plan.
static void build_apx_pat_uncached(bool qgram, bool jit, const Compiled &c, ApxPat &out)
{
    const ApxCost &K = jit ? kCostJit : kCostGeneric;
    const DevPlan &dp = c.dp;
    const int m = dp.m, k = dp.k, L = dp.L, np = dp.npieces;
    const int win = (dp.ins || dp.del) ? 2 * k + 1 : 1;
    const int rows = k + 1;
    memset(&out, 0, sizeof out);
    out.m = m; out.k = k; out.L = L; out.npieces = np; out.indel = (dp.ins || dp.del) ? 1 : 0;
    for (int i = 0; i < np; i++) out.V[i] = dp.V[i];
    out.win = win;
    std::vector<unsigned> pcls((size_t)m);
    std::vector<double> pprob((size_t)m), info((size_t)m);
    static const unsigned code_bit[4] = {1u, 2u, 8u, 4u};              // code = hi<<1|lo : A, C, T, G
    for (int j = 0; j < m; j++) {
        pcls[j] = packed_class_of(c.P.pos[j], nullptr);
        pprob[j] = pcls[j] == 31u ? 1.0 : std::min(__builtin_popcount(pcls[j] & 15u) / 4.0 + 0.001, 1.0);
        info[j] = pcls[j] == 31u ? 0.0 : -std::log2(pprob[j]);
        for (int q = 0; q < 4; q++) if (pcls[j] & code_bit[q]) out.posmask[q] |= 1ULL << j;
        if (pcls[j] & 16u) out.posmask[4] |= 1ULL << j;
    }
    double piece_rate[PM_MAX_PIECES] = {0};
    for (int i = 0; i < np; i++) {
        double pr = 1;
        for (int j = 0; j < L; j++) pr *= pprob[dp.V[i] + j];
        piece_rate[i] = pr;
    }
    int steps = 0;
    for (int cw = 1; cw < win; cw += std::min(cw, win - cw)) steps++;
    const double dil_cost = jit ? K.dil[std::min(win / 2, 3)] : K.dil_step * steps;
    const double count_cost = dil_cost + K.count_fix + K.count_row * rows;
    // per pattern start that reaches the sparse stage
    const double anchor_ops = jit ? K.survivor : (m + 2 * k <= 32 ? 1.0 : 1.6) * (4.0 + 1.2 * (2 * k + 1) * (k + 1)) + 3.0;
    struct Seg { int b, e, piece; };
    ApxPat best = out;
    double best_cost = -1, best_ops = 0, best_surv = 0;
    const char *fam_env = getenv("PM_APX_FAMILY");                     // experiments: restrict the candidate families (A, B, C)
    const bool dbg = getenv("PM_APX_DEBUG") != nullptr;
    auto fam_ok = [&](char f) { return !fam_env || strchr(fam_env, f) != nullptr; };
    auto evaluate = [&](const std::vector<Seg> &segs, bool derive) {
        ApxPat cur = out;
        double ops = 0;
        std::vector<std::pair<int, int>> span;                         // constrained positions of each COUNTED chunk: [first, last]
        std::vector<int> covered((size_t)m, 0);
        std::vector<double> chunk_ops;                                 // evaluation cost of each kept chunk (to undo it)
        for (const Seg &sg : segs) {
            if (cur.nch >= AX_MAXCH) { if (derive) return; break; }
            ApxChunk ch;
            memset(&ch, 0, sizeof ch);
            ch.piece = 0xff;
            double pr = 1;
            int first = -1, last = -1;
            bool cut = false;
            unsigned seen_cls = 0;
            int ncls = 0;
            for (int j = sg.b; j < sg.e; j++) {
                if (pcls[j] == 31u) continue;
                if (first < 0) first = j;
                if (ch.npos >= AX_MAXLEN || j - first >= 32) { cut = true; break; }
                ch.t[ch.npos] = (unsigned char)(j - first);
                ch.cls[ch.npos] = (unsigned char)pcls[j];
                ch.npos++;
                if (!(seen_cls >> pcls[j] & 1u)) { seen_cls |= 1u << pcls[j]; ncls++; }
                pr *= pprob[j];
                last = j;
            }
            if (first < 0) continue;
            if (cut && derive && sg.piece >= 0) return;                // a derived piece needs all its positions in chunks
            pr = std::min(pr * win, 1.0);
            ch.counted = pr <= 0.6 ? 1 : 0;                            // hardly ever missing: not worth counting
            const bool builds = derive && sg.piece >= 0;
            if (!ch.counted && !builds) continue;
            ch.off = (unsigned char)(first + (win > 1 ? 0 : k));
            ch.poff = (unsigned char)(first + k);
            if (builds) ch.piece = (unsigned char)sg.piece;
            const double ev = K.pos * ch.npos + K.cls * ncls;
            ops += ev + (builds ? (ch.counted || !jit ? K.derive : K.dense_piece) : 0.0);
            if (ch.counted) { ops += count_cost; span.push_back({first, last}); }
            for (int t = 0; t < ch.npos; t++) covered[first + ch.t[t]] = 1;
            chunk_ops.push_back(ev);
            cur.ch[cur.nch++] = ch;
        }
        int ncounted = (int)span.size();
        if (ncounted <= k) {                                           // k missing chunks are always allowed: no count
            if (!derive) { cur.nch = 0; ops = 0; chunk_ops.clear(); }
            else for (int g = 0; g < cur.nch; g++) {
                if (cur.ch[g].counted) ops -= count_cost;
                cur.ch[g].counted = 0;
            }
            ncounted = 0;
            span.clear();
        }
        cur.ncounted = ncounted;
        // pieces: derived from their chunks, or position by position
        int dp_n = 0;
        for (int i = 0; i < np; i++) {
            int ncons = 0, ncov = 0;
            for (int j = dp.V[i]; j < dp.V[i] + L; j++) if (pcls[j] != 31u) { ncons++; ncov += covered[j]; }
            if (ncons == 0) { cur.dwild[i] = 1; continue; }
            bool derived = derive && ncov == ncons;
            if (derived) {
                int gf = -1, gl = -1;
                for (int g = 0; g < cur.nch; g++) if (cur.ch[g].piece == i) { if (gf < 0) gf = g; gl = g; }
                if (gf < 0) derived = false;
                else { cur.ch[gf].first = 1; cur.ch[gl].last = 1; }
            }
            if (!derived) {
                for (int g = 0; g < cur.nch; g++) if (cur.ch[g].piece == i) { cur.ch[g].piece = 0xff; ops -= K.derive; }
                unsigned seen_cls = 0;
                for (int j = dp.V[i]; j < dp.V[i] + L; j++) {
                    if (pcls[j] == 31u) continue;
                    if (dp_n >= AX_MAXDENSE) return;
                    cur.dshift[dp_n] = (unsigned char)(k + j);
                    cur.dcls[dp_n] = (unsigned char)pcls[j];
                    dp_n++;
                    cur.dn[i]++;
                    ops += K.dense_pos;
                    if (!(seen_cls >> pcls[j] & 1u)) { seen_cls |= 1u << pcls[j]; ops += K.cls; }
                }
                ops += K.dense_piece;
            }
        }
        // uncounted chunks that build no piece are useless
        {
            int w = 0;
            for (int g = 0; g < cur.nch; g++) {
                if (!cur.ch[g].counted && cur.ch[g].piece == 0xff) { ops -= chunk_ops[(size_t)g]; continue; }
                cur.ch[w++] = cur.ch[g];
            }
            cur.nch = w;
        }
        // Expected pattern starts that survive.  An anchor of piece i matched that piece exactly, so the chunk
        // positions inside the piece are present for free; the others are taken as independent.
        double survivors = 0;
        for (int i = 0; i < np; i++) {
            double pass = 1;
            if (ncounted > 0) {
                std::vector<double> dist((size_t)ncounted + 1, 0.0);
                dist[0] = 1;
                for (int g = 0; g < ncounted; g++) {
                    double pout = 1, pall = 1;
                    for (int j = span[g].first; j <= span[g].second; j++) {
                        if (pcls[j] == 31u) continue;
                        pall *= pprob[j];
                        if (j < dp.V[i] || j >= dp.V[i] + L) pout *= pprob[j];
                    }
                    const double pgi = std::min(pout + (win - 1) * pall, 1.0);
                    for (int r = g + 1; r >= 0; r--)
                        dist[r] = dist[r] * pgi + (r > 0 ? dist[r - 1] * (1 - pgi) : 0.0);
                }
                pass = 0;
                for (int r = 0; r <= k && r <= ncounted; r++) pass += dist[r];
            }
            survivors += piece_rate[i] * pass;
        }
        const double cost = ops + std::min(survivors, 1.0) * 8192.0 * anchor_ops;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = cur; best_ops = ops; best_surv = survivors; }
    };
    // (C) no count, pieces position by position
    evaluate({}, false);
    if (qgram) {
        if (fam_ok('A'))
        // (A) pieces cut into 1..3 sub-chunks of about equal information; the stretches outside the pieces into 0..2
        {
            // equal-information split of [pb, pe) into `parts` segments
            auto split = [&](int pb, int pe, int parts, int piece, std::vector<Seg> &segs) {
                double tot = 0;
                for (int j = pb; j < pe; j++) tot += info[j];
                int jb = pb;
                double acc = 0;
                for (int r = 1; r <= parts && jb < pe; r++) {
                    const double target = tot * r / parts;
                    int je = jb;
                    double a2 = acc;
                    while (je < pe && (je == jb || std::fabs(a2 + info[je] - target) <= std::fabs(a2 - target))) { a2 += info[je]; je++; }
                    if (r == parts) je = pe;
                    segs.push_back({jb, je, piece});
                    acc = 0;
                    for (int j = pb; j < je; j++) acc += info[j];
                    jb = je;
                }
            };
            // stretches not covered by pieces (pieces are sorted by V and disjoint)
            std::vector<std::pair<int, int>> gaps;
            {
                int pos = 0;
                for (int i = 0; i < np; i++) {
                    if (dp.V[i] > pos) gaps.push_back({pos, dp.V[i]});
                    pos = std::max(pos, dp.V[i] + L);
                }
                if (pos < m) gaps.push_back({pos, m});
                std::vector<std::pair<int, int>> keep;
                for (auto &g : gaps) {
                    double t = 0;
                    for (int j = g.first; j < g.second; j++) t += info[j];
                    if (t > 0) keep.push_back(g);
                }
                gaps.swap(keep);
            }
            const int ng = std::min((int)gaps.size(), 4);
            std::vector<int> parts((size_t)np, 1), gparts((size_t)ng, 0);
            for (;;) {
                std::vector<Seg> segs;
                {
                    // in pattern order, so that the chunks of a piece stay consecutive
                    size_t gi = 0;
                    for (int i = 0; i < np; i++) {
                        while (gi < (size_t)ng && gaps[gi].first < dp.V[i]) { if (gparts[gi]) split(gaps[gi].first, gaps[gi].second, gparts[gi], -1, segs); gi++; }
                        split(dp.V[i], dp.V[i] + L, parts[i], i, segs);
                    }
                    while (gi < (size_t)ng) { if (gparts[gi]) split(gaps[gi].first, gaps[gi].second, gparts[gi], -1, segs); gi++; }
                }
                evaluate(segs, true);
                int q = 0;
                while (q < np && ++parts[q] > 3) parts[q++] = 1;
                if (q == np) {
                    int g = 0;
                    const int gmax = np + ng > 6 ? 1 : 2;              // keeps the enumeration below ~2000 evaluations
                    while (g < ng && ++gparts[g] > gmax) gparts[g++] = 0;
                    if (g == ng) break;
                }
            }
        }
        // (B1) uniform tilings of every length and phase
        if (fam_ok('B'))
        for (int q = 2; q <= AX_MAXLEN; q++)
            for (int a_off = 0; a_off < q; a_off++) {
                std::vector<Seg> segs;
                for (int j0 = a_off - q; j0 < m; j0 += q) {
                    const int jb = std::max(j0, 0), je = std::min(j0 + q, m);
                    if (je > jb) segs.push_back({jb, je, -1});
                }
                evaluate(segs, false);
            }
        // (B2) T segments of (nearly) equal information: wildcards carry none, so chunks stretch over them
        if (fam_ok('B')) {
            std::vector<double> cum((size_t)m + 1, 0.0);
            for (int j = 0; j < m; j++) cum[j + 1] = cum[j] + info[j];
            for (int T = k + 1; T <= std::min(2 * k + 3, AX_MAXCH); T++) {
                std::vector<Seg> segs;
                int jb = 0;
                for (int r = 1; r <= T && jb < m; r++) {
                    const double target = cum[m] * r / T;
                    int je = jb + 1;
                    while (je < m && std::fabs(cum[je + 1] - target) <= std::fabs(cum[je] - target)) je++;
                    if (r == T) je = m;
                    segs.push_back({jb, je, -1});
                    jb = je;
                }
                evaluate(segs, false);
            }
        }
    }
    out = best;
    if (dbg) fprintf(stderr, "[apx plan] jit=%d nch=%d counted=%d derived-pieces=%d dense ops=%.0f survivors/start=%.5f cost=%.0f\n", (int)jit, best.nch,
                     best.ncounted, (int)std::count_if(best.ch, best.ch + best.nch, [](const ApxChunk &ch) { return ch.piece != 0xff && ch.first; }),
                     best_ops, best_surv, best_cost);
}

int pm_debug_reset_caches(void)
{
    { std::lock_guard<std::mutex> lock(g_compile_mu); g_compile_cache.clear(); }
    { std::lock_guard<std::mutex> lock(g_apx_mu); g_apx_cache.clear(); }
    return PM_OK;
}

// ---- specialised kernels (NVRTC): compiled once per request text, shared by all engines of the process ----
#define PM_JIT_MIN_WORK (1LL << 27)     // bases per pattern from which a ~0.3 s compilation pays off within a few requests
static bool jit_wanted(const pm_engine *e, long long bases) { return e->jit_mode == 2 || (e->jit_mode == 1 && bases >= PM_JIT_MIN_WORK); }
struct JitEntry {
    enum State { COMPILING, COMPILED, READY, FAILED } state = COMPILING;
    std::vector<char> cubin;               // COMPILED: the image, not yet loaded (loading needs the caller's CUDA context)
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kern = nullptr;
    int rc = PM_OK;
    std::string err;
    unsigned long long attr_devices = 0;   // devices on which the shared-memory attribute has been set
};
// never destroyed: a background compilation may still be running when the process exits
static std::mutex &g_jit_mu = *new std::mutex;
static std::condition_variable &g_jit_cv = *new std::condition_variable;
static std::map<std::string, JitEntry> &g_jit_cache = *new std::map<std::string, JitEntry>;
static int g_jit_pending = 0;              // compilations in flight (g_jit_mu)

// NVRTC only (host work): runs on the caller's thread in mode 2, on its own thread in auto mode
static void jit_compile_entry(const std::string &key, const std::string &source)
{
    std::vector<char> cubin;
    std::string log;
    const int bad = apx_jit_compile(source, cubin, log);
    std::lock_guard<std::mutex> lock(g_jit_mu);
    auto it = g_jit_cache.find(key);
    if (it != g_jit_cache.end() && it->second.state == JitEntry::COMPILING) {     // (the caches may have been reset meanwhile)
        JitEntry &en = it->second;
        if (bad) {
            en.state = JitEntry::FAILED;
            en.rc = log.find("libnvrtc") != std::string::npos ? PM_ERR_UNSUPPORTED : PM_ERR_CUDA;
            en.err = "specialised scan kernel: " + log;
        } else {
            en.cubin.swap(cubin);
            en.state = JitEntry::COMPILED;
        }
    }
    g_jit_pending--;
    g_jit_cv.notify_all();
}

int pm_jit_wait(void)
{
    std::unique_lock<std::mutex> lock(g_jit_mu);
    g_jit_cv.wait(lock, [] { return g_jit_pending == 0; });
    return PM_OK;
}

static int launch_apx_jit(pm_engine *e, const ApxArgs &a, long long lo, long long hi, bool exact)
{
    const ApxJitShape shp = apx_jit_shape(exact);
    // key: the pattern descriptions without what stays a run-time argument (ranges, key tags) + the launch geometry
    std::string key((size_t)a.npat * sizeof(ApxPat) + sizeof shp, '\0');
    for (int p = 0; p < a.npat; p++) {
        ApxPat t = a.pat[p];
        t.a0 = t.a1 = 0; t.keytag = 0;
        memcpy(&key[(size_t)p * sizeof(ApxPat)], &t, sizeof t);
    }
    {
        ApxJitShape t;
        memset(&t, 0, sizeof t);
        t.stream = shp.stream + (exact ? 2 : 0); t.w = shp.w; t.warps = shp.warps; t.stages = shp.stages; t.ctas = shp.ctas; t.tile_words = shp.tile_words; t.smem = shp.smem;
        memcpy(&key[(size_t)a.npat * sizeof(ApxPat)], &t, sizeof t);
    }
    cudaKernel_t kern = nullptr;
    {
        std::unique_lock<std::mutex> lock(g_jit_mu);
        auto it = g_jit_cache.find(key);
        if (it == g_jit_cache.end()) {
            it = g_jit_cache.emplace(key, JitEntry()).first;
            g_jit_pending++;
            std::string source = apx_full_source(apx_generate_prefix(a.pat, a.npat, exact));
            if (e->jit_mode == 2) {
                lock.unlock();
                jit_compile_entry(key, source);
                lock.lock();
            } else {
                // auto mode: the compilation (about 0.2 s) runs beside the requests; until it is there the generic
                // kernel answers (the caller falls back on PM_ERR_UNSUPPORTED)
                // a process that exits while NVRTC is still working must not unload it under the compiler's feet
                static std::once_flag at_exit_once;
                std::call_once(at_exit_once, [] { atexit([] { (void)pm_jit_wait(); }); });
                std::thread(jit_compile_entry, key, std::move(source)).detach();
            }
            it = g_jit_cache.find(key);
            if (it == g_jit_cache.end()) { g_err = "specialised scan kernel: caches were reset during the compilation"; return PM_ERR_UNSUPPORTED; }
        }
        if (it->second.state == JitEntry::COMPILING) {
            if (e->jit_mode != 2) { g_err = "specialised scan kernel: still compiling"; return PM_ERR_UNSUPPORTED; }
            g_jit_cv.wait(lock, [&] { auto i2 = g_jit_cache.find(key); return i2 == g_jit_cache.end() || i2->second.state != JitEntry::COMPILING; });
            it = g_jit_cache.find(key);
            if (it == g_jit_cache.end()) { g_err = "specialised scan kernel: caches were reset during the compilation"; return PM_ERR_CUDA; }
        }
        JitEntry &en = it->second;
        if (en.state == JitEntry::COMPILED) {
            cudaError_t ce = cudaLibraryLoadData(&en.lib, en.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
            if (ce == cudaSuccess) ce = cudaLibraryGetKernel(&en.kern, en.lib, "k_scan_apx_jit");
            if (ce != cudaSuccess) {
                en.state = JitEntry::FAILED; en.rc = PM_ERR_CUDA;
                en.err = std::string("loading the specialised scan kernel: ") + cudaGetErrorString(ce);
                (void)cudaGetLastError();
            } else en.state = JitEntry::READY;
            std::vector<char>().swap(en.cubin);
        }
        if (en.state == JitEntry::FAILED) { g_err = en.err; return en.rc; }
        if (!(en.attr_devices >> (e->device & 63) & 1ULL)) {
            CK(cudaFuncSetAttribute((const void *)en.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shp.smem));
            en.attr_devices |= 1ULL << (e->device & 63);
        }
        kern = en.kern;
    }
    JitArgs ja;
    memset(&ja, 0, sizeof ja);
    ja.hi = a.hi; ja.lo = a.lo; ja.xx = a.xx; ja.nwords = a.nwords; ja.n = a.n;
    ja.keys = a.keys; ja.count = a.count; ja.cap = a.cap;
    for (int p = 0; p < a.npat; p++) { ja.a0[p] = a.pat[p].a0; ja.a1[p] = a.pat[p].a1; ja.keytag[p] = a.pat[p].keytag; }
    long long nblk;
    if (shp.stream) {
        // pattern starts lo <= b < hi; the kernel works in end coordinates e = b + span <= b + 63
        const long long w0 = (lo / 32) & ~3LL;
        const long long w1 = std::min((hi + 63) / 32 + 1, a.nwords);
        nblk = std::max<long long>((w1 - w0 + shp.tile_words - 1) / shp.tile_words, 1);
        ja.tile0 = w0; ja.ntiles = nblk;
    } else {
        ja.tile0 = a.tile0; ja.ntiles = a.ntiles;
        nblk = (a.ntiles + 7) / 8;
    }
    const int grid = std::max((int)std::min<long long>(nblk, (long long)e->sms * shp.ctas), 1);
    void *params[] = {&ja};
    CK(cudaLaunchKernel((const void *)kern, dim3((unsigned)grid), dim3((unsigned)shp.warps * 32), params, shp.smem, e->stream));
    e->stats.launches++;
    e->stats.jit = 1;
    e->stats.scan_bytes += nblk * (shp.stream ? shp.tile_words : 1024) * 4 * 3;
    return PM_OK;
}

static int launch_apx(pm_engine *e, pm_dataset *d, const ApxPat *pats, int npat, const ScanTarget &t)
{
    ApxArgs a;
    memset(&a, 0, sizeof a);
    long long lo = -1, hi = -1;
    bool narrow = true;
    int k = 0;
    for (int p = 0; p < npat; p++) {
        if (pats[p].a1 <= pats[p].a0) continue;
        a.pat[a.npat++] = pats[p];
        // pattern starts b = anchor - k - V[i] >= 0 must be covered for every anchor in [a0, a1)
        const long long b0 = std::max<long long>(pats[p].a0 - pats[p].k - pats[p].V[pats[p].npieces - 1], 0);
        lo = lo < 0 ? b0 : std::min(lo, b0);
        hi = std::max(hi, pats[p].a1);
        if (pats[p].m + 2 * pats[p].k > 32) narrow = false;
        k = pats[p].k;
    }
    if (a.npat == 0) return PM_OK;
    a.hi = d->hi; a.lo = d->lo; a.xx = d->xx; a.nwords = d->nwords; a.n = d->n;
    a.tile0 = (lo / 32) / 128;
    a.ntiles = ((hi - 1) / 32) / 128 + 1 - a.tile0;
    a.keys = t.keys; a.count = t.count; a.cap = t.cap;
    const size_t smem = SP_STAGES * EX_STAGE_BYTES + 2 * SP_STAGES * 8;
    if (!e->attr_apx) {
#define PM_ATTR(R, W) CK(cudaFuncSetAttribute(k_scan_apx<R, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
        PM_ATTR(2, unsigned); PM_ATTR(3, unsigned); PM_ATTR(4, unsigned);
        PM_ATTR(2, unsigned long long); PM_ATTR(3, unsigned long long); PM_ATTR(4, unsigned long long);
#undef PM_ATTR
        e->attr_apx = true;
    }
    const long long nbt = (a.ntiles + 7) / 8;
    const int grid = std::max((int)std::min<long long>(nbt, (long long)e->sms * SP_CTAS), 1);
    e->stats.scan_bases += (hi - lo) * a.npat;
    e->stats.packed = 1;
    e->stats.qgram_chunks = a.pat[0].ncounted;
    // genome-scale requests: kernel compiled for exactly the request's patterns (apx_jit.cpp)
    if (jit_wanted(e, hi - lo)) {
        int rc = launch_apx_jit(e, a, lo, hi, false);
        if (rc == PM_OK) return PM_OK;
        if (e->jit_mode == 2 || rc != PM_ERR_UNSUPPORTED) return rc;   // auto mode: no NVRTC on this machine -> generic kernel
    }
    e->stats.scan_bytes += a.ntiles * 128 * 4 * 3;
#define PM_LAUNCH(R, W) k_scan_apx<R, W><<<grid, EX_WARPS * 32, smem, e->stream>>>(a)
    if (narrow) { if (k == 1) PM_LAUNCH(2, unsigned); else if (k == 2) PM_LAUNCH(3, unsigned); else PM_LAUNCH(4, unsigned); }
    else { if (k == 1) PM_LAUNCH(2, unsigned long long); else if (k == 2) PM_LAUNCH(3, unsigned long long); else PM_LAUNCH(4, unsigned long long); }
#undef PM_LAUNCH
    CK(cudaGetLastError());
    e->stats.launches++;
    return PM_OK;
}

// One scan launch for one compiled pattern over the anchors [a0, a1): candidate keys (tagged) are appended to t.
static int launch_scan(pm_engine *e, pm_dataset *d, const Compiled &c_full, long long a0, long long a1, const Fills &fills,
                       const unsigned long long *dB, const unsigned long long *dTL, const unsigned long long *dTR,
                       const ScanTarget &t, unsigned long long tag, unsigned long long bad)
{
    const Compiled &c = c_full.scan ? *c_full.scan : c_full;   // what the scan kernels see (EXTENDED: the plain window)
    const DevPlan &dp = c.dp;
    const long long n = d->n;
    const bool use_packed = scan_uses_packed(e, d, c_full);
    {
        if (use_packed) {
            const long long wend = std::min(a1, n - dp.L + 1);
            if (wend > a0) {
                const long long tile0 = (a0 / 32) / 128;
                const long long ntiles = ((wend - 1) / 32) / 128 + 1 - tile0;
                auto packed_class = [&](const pm::ByteSet &bs) -> unsigned { return packed_class_of(bs, nullptr); };
                auto plane_of = [](unsigned cls) -> int { return plane_of_class(cls); };
                if (dp.type == PM_PLAN_SIMPLE) {
                    ExactPat pt;
                    fill_exact_pat(c, a0, a1, n, tag, pt);
                    return launch_exact(e, d, &pt, 1, bad, t);
                } else if (apx_eligible(e, d, c)) {
                    ApxPat ap;
                    build_apx_pat(e->qgram_filter != 0, jit_wanted(e, wend - a0), c, a0, wend, tag, ap);
                    return launch_apx(e, d, &ap, 1, t);
                } else {
                    PackedArgs<4> a;
                    memset(&a, 0, sizeof a);
                    a.hi = d->hi; a.lo = d->lo; a.xx = d->xx; a.nwords = d->nwords; a.n = n; a.a0 = a0; a.a1 = wend;
                    a.tile0 = tile0; a.ntiles = ntiles;
                    a.L = dp.L; a.npieces = dp.npieces;
                    a.keytag = tag;
                    for (int i = 0; i < dp.npieces; i++) {
                        for (int j = 0; j < dp.npieces; j++)
                            if (dp.trig[i] & (1ULL << (j * dp.L + dp.L - 1))) a.trigsets[i] |= 1u << j;
                        for (int j = 0; j < dp.L; j++) {
                            const unsigned cls = packed_class(c.P.pos[dp.V[i] + j]);
                            PackedPos pp;
                            pp.cls = (unsigned char)cls;
                            pp.sel = (unsigned char)plane_of(cls);
                            a.pos[i][j] = pp;
                        }
                    }
                    a.keys = t.keys; a.count = t.count; a.cap = t.cap;
                    PackedVerify<4> pv;
                    memset(&pv, 0, sizeof pv);
                    pv.enabled = (dp.k <= 3 && e->fused_filter && dp.m <= 64) ? 1 : 0;   // the Myers filter keeps 64-bit parts
                    pv.m = dp.m; pv.k = dp.k; pv.ins = dp.ins; pv.del = dp.del; pv.subs = dp.subs;
                    pv.cuts = d->d_fills + 2 * (size_t)d->nfills; pv.ncuts = d->ncuts;
                    bool narrow = true;                                   // every pattern part fits a 32-bit state word
                    static const unsigned char code_byte[4] = {'A', 'C', 'T', 'G'};   // code = hi<<1 | lo
                    for (int i = 0; i < dp.npieces; i++) {
                        pv.V[i] = dp.V[i];
                        if (dp.V[i] > 32 || dp.m - dp.V[i] > 32) narrow = false;
                        const int lbp = dp.V[i], rlp = dp.m - dp.V[i];
                        pv.itmax = std::max(pv.itmax, (lbp > 0 ? lbp + dp.k : 0) + (rlp > 0 ? rlp + dp.k : 0));
                        for (int j = 0; j < lbp; j++) if (packed_class(c.P.pos[lbp - 1 - j]) & 16u) pv.TLX[i] |= 1ULL << j;
                        for (int j = 0; j < rlp; j++) if (packed_class(c.P.pos[lbp + j]) & 16u) pv.TRX[i] |= 1ULL << j;
                        for (int q = 0; q < 4; q++) {
                            pv.TL[i][q] = c.vt.TL[(size_t)i * 256 + code_byte[q]];
                            pv.TR[i][q] = c.vt.TR[(size_t)i * 256 + code_byte[q]];
                        }
                    }
                    const int rows = std::min(std::max(dp.k, 1), 3) + 1;
                    // ---- q-gram pre-filter: pick the disjoint chunks (see packed.cuh) ----
                    QFilter qf;
                    memset(&qf, 0, sizeof qf);
                    bool plain_trig = true;
                    for (int i = 0; i < dp.npieces; i++) if (a.trigsets[i] != (1u << i)) plain_trig = false;
                    long long t0q = tile0, ntq = ntiles;
                    // register-resident kernel in pattern-start coordinates (packed.cuh: k_scan_split)
                    const bool bcoords = pv.enabled && e->split_kernel == 2 && plain_trig && dp.k >= 1 && dp.k <= 3 && dp.m + 2 * dp.k <= 64;
                    if (bcoords && e->qgram_filter) {
                        const int win = (dp.ins || dp.del) ? 2 * dp.k + 1 : 1;
                        std::vector<unsigned> pcls((size_t)dp.m);
                        std::vector<double> pprob((size_t)dp.m);
                        for (int j = 0; j < dp.m; j++) {
                            pcls[j] = packed_class(c.P.pos[j]);
                            pprob[j] = pcls[j] == 31u ? 1.0 : std::max(__builtin_popcount(pcls[j] & 15u), 0) / 4.0 + 0.001;
                        }
                        double cand_rate = 0;                         // piece hits per base
                        double piece_rate[PM_MAX_PIECES] = {0};
                        for (int i = 0; i < dp.npieces; i++) {
                            double pr = 1;
                            for (int j = 0; j < dp.L; j++) pr *= std::min(pprob[dp.V[i] + j], 1.0);
                            piece_rate[i] = pr;
                            cand_rate += pr;
                        }
                        // cost model in thread operations per base: a pattern position costs ~42 warp instructions per
                        // 8192-base warp tile; a surviving anchor ~80 (latency-bound Myers round + extraction), the value
                        // that ranked the candidate chunk sets best in a sweep over 11 motifs (tools/qf_bench.py)
                        const double anchor_cost = 2500.0;
                        double best_cost = cand_rate * anchor_cost * 0.7;     // worth it only with a clear margin
                        int steps = 0;
                        for (int cw = 1; cw < win; cw += std::min(cw, win - cw)) steps++;
                        auto evaluate = [&](const std::vector<std::pair<int, int>> &segs) {
                            QFilter cur;
                            memset(&cur, 0, sizeof cur);
                            cur.win = win;
                            double ops = 0;
                            std::vector<std::pair<int, int>> span;            // constrained positions of each kept chunk: [first, last]
                            for (const auto &sg : segs) {
                                if (cur.nch >= QF_MAXCH) break;
                                QChunk ch;
                                memset(&ch, 0, sizeof ch);
                                double pr = 1;
                                int first = -1, last = -1;
                                for (int j = sg.first; j < sg.second; j++) {
                                    if (pcls[j] == 31u) continue;
                                    if (first < 0) first = j;
                                    if (ch.npos >= QF_MAXLEN || j - first >= 32) break;
                                    ch.t[ch.npos] = (unsigned char)(j - first);
                                    ch.pos[ch.npos].cls = (unsigned char)pcls[j];
                                    ch.pos[ch.npos].sel = (unsigned char)plane_of(pcls[j]);
                                    ch.npos++;
                                    pr *= pprob[j];
                                    last = j;
                                }
                                if (first < 0) continue;
                                pr = std::min(pr * win, 1.0);
                                if (pr > 0.6) continue;                       // hardly ever missing: not worth its operations
                                ops += 21.0 * ch.npos + 8.0 * steps + 8 + 5 * rows;
                                ch.off = (unsigned char)(first + (win > 1 ? 0 : dp.k));
                                cur.ch[cur.nch++] = ch;
                                span.push_back({first, last});
                            }
                            if (cur.nch <= dp.k) return;                      // k missing chunks are always allowed
                            // Expected anchors that survive.  An anchor of piece i matched that piece exactly, so the chunk
                            // positions inside the piece are present for free; the others are taken as independent.
                            double survivors = 0;
                            for (int i = 0; i < dp.npieces; i++) {
                                std::vector<double> dist((size_t)cur.nch + 1, 0.0);
                                dist[0] = 1;
                                for (int g = 0; g < cur.nch; g++) {
                                    double pout = 1, pall = 1;
                                    for (int j = span[g].first; j <= span[g].second; j++) {
                                        if (pcls[j] == 31u) continue;
                                        pall *= pprob[j];
                                        if (j < dp.V[i] || j >= dp.V[i] + dp.L) pout *= pprob[j];
                                    }
                                    const double pgi = std::min(pout + (win - 1) * pall, 1.0);
                                    for (int r = g + 1; r >= 0; r--)
                                        dist[r] = dist[r] * pgi + (r > 0 ? dist[r - 1] * (1 - pgi) : 0.0);
                                }
                                double pass = 0;
                                for (int r = 0; r <= dp.k && r <= cur.nch; r++) pass += dist[r];
                                survivors += piece_rate[i] * pass;
                            }
                            const double cost = ops / 128.0 + survivors * anchor_cost;
                            if (cost < best_cost) { best_cost = cost; qf = cur; }
                        };
                        // (1) uniform tilings of every length and phase
                        for (int q = 2; q <= QF_MAXLEN; q++)
                            for (int a_off = 0; a_off < q; a_off++) {
                                std::vector<std::pair<int, int>> segs;
                                for (int j0 = a_off - q; j0 < dp.m; j0 += q) {
                                    const int jb = std::max(j0, 0), je = std::min(j0 + q, dp.m);
                                    if (je > jb) segs.push_back({jb, je});
                                }
                                evaluate(segs);
                            }
                        // (2) T segments of (nearly) equal information: wildcards carry none, so chunks stretch over them
                        {
                            std::vector<double> cum((size_t)dp.m + 1, 0.0);
                            for (int j = 0; j < dp.m; j++) cum[j + 1] = cum[j] + (pcls[j] == 31u ? 0.0 : -std::log2(std::min(pprob[j], 1.0)));
                            for (int T = dp.k + 1; T <= std::min(2 * dp.k + 3, QF_MAXCH); T++) {
                                std::vector<std::pair<int, int>> segs;
                                int jb = 0;
                                for (int r = 1; r <= T && jb < dp.m; r++) {
                                    const double target = cum[dp.m] * r / T;
                                    int je = jb + 1;
                                    while (je < dp.m && std::fabs(cum[je + 1] - target) <= std::fabs(cum[je] - target)) je++;
                                    if (r == T) je = dp.m;
                                    segs.push_back({jb, je});
                                    jb = je;
                                }
                                evaluate(segs);
                            }
                        }
                    }
                    if (bcoords) {
                        // pattern starts b = anchor - k - V[i] >= 0 must be covered for every anchor in [a0, wend)
                        const long long b0 = std::max<long long>(a0 - dp.k - dp.V[dp.npieces - 1], 0);
                        t0q = (b0 / 32) / 128;
                        ntq = ((wend - 1) / 32) / 128 + 1 - t0q;
                        a.tile0 = t0q; a.ntiles = ntq;
                    }
                    const int grid_bk = std::max((int)std::min<long long>((ntq + 7) / 8, (long long)e->sms * 6), 1);
                    const int grid_sp = std::max((int)std::min<long long>((ntq + 7) / 8, (long long)e->sms * SP_CTAS), 1);
                    const size_t smem_sp = SP_STAGES * EX_STAGE_BYTES + 2 * SP_STAGES * 8;
                    if (!e->attr_split) {                    // static queues + dynamic ring exceed the 48 KB default
#define PM_ATTR(W, R) CK(cudaFuncSetAttribute(k_scan_split<4, W, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sp))
                        PM_ATTR(unsigned, 2); PM_ATTR(unsigned, 3); PM_ATTR(unsigned, 4);
                        PM_ATTR(unsigned long long, 2); PM_ATTR(unsigned long long, 3); PM_ATTR(unsigned long long, 4);
#undef PM_ATTR
                        e->attr_split = true;
                    }
#define PM_LAUNCH(W, R) do { if (bcoords) k_scan_split<4, W, R><<<grid_sp, EX_WARPS * 32, smem_sp, e->stream>>>(a, pv, qf); \
                             else k_scan_packed<4, W, R><<<grid_bk, 256, 0, e->stream>>>(a, pv); } while (0)
                    if (narrow) { if (rows == 2) PM_LAUNCH(unsigned, 2); else if (rows == 3) PM_LAUNCH(unsigned, 3); else PM_LAUNCH(unsigned, 4); }
                    else { if (rows == 2) PM_LAUNCH(unsigned long long, 2); else if (rows == 3) PM_LAUNCH(unsigned long long, 3); else PM_LAUNCH(unsigned long long, 4); }
                    e->stats.qgram_chunks = qf.nch;
#undef PM_LAUNCH
                }
                e->stats.launches++;
                e->stats.scan_bytes += ntiles * 128 * 4 * 3;
                e->stats.scan_bases += wend - a0;
                e->stats.packed = 1;
            }
        } else if (dp.type == PM_PLAN_SIMPLE || dp.type == PM_PLAN_SPLIT) {
            // window starts w in [a0, a1) <=> end positions p = w + L - 1
            long long p0 = a0 + dp.L - 1, p1 = std::min(a1 + dp.L - 1, n);
            if (p1 > p0) {
                const long long tile0 = p0 / SCAN_TILE, tile1 = (p1 - 1) / SCAN_TILE + 1;
                const long long ntiles = tile1 - tile0;
                const int grid = (int)std::min<long long>(ntiles, (long long)e->sms * 8);
                if (dp.npieces * dp.L <= 32) {
                    ScanArgs<unsigned> a;
                    a.text = d->d_text; a.codes = scan_uses_pep5(e, d, c_full) ? d->codes : nullptr; a.n = n; a.p0 = p0; a.p1 = p1; a.tile0 = tile0; a.ntiles = ntiles; a.B = dB;
                    a.init = (unsigned)dp.init; a.fin = (unsigned)dp.fin;
                    for (int i = 0; i < PM_MAX_PIECES; i++) a.trig[i] = (unsigned)dp.trig[i];
                    a.L = dp.L; a.npieces = dp.npieces; a.keys = t.keys; a.count = t.count; a.cap = t.cap; a.keytag = tag;
                    k_scan_bytes<unsigned><<<grid, SCAN_THREADS, 0, e->stream>>>(a);
                } else {
                    ScanArgs<unsigned long long> a;
                    a.text = d->d_text; a.codes = scan_uses_pep5(e, d, c_full) ? d->codes : nullptr; a.n = n; a.p0 = p0; a.p1 = p1; a.tile0 = tile0; a.ntiles = ntiles; a.B = dB;
                    a.init = dp.init; a.fin = dp.fin;
                    for (int i = 0; i < PM_MAX_PIECES; i++) a.trig[i] = dp.trig[i];
                    a.L = dp.L; a.npieces = dp.npieces; a.keys = t.keys; a.count = t.count; a.cap = t.cap; a.keytag = tag;
                    k_scan_bytes<unsigned long long><<<grid, SCAN_THREADS, 0, e->stream>>>(a);
                }
                e->stats.launches++;
                e->stats.scan_bytes += scan_uses_pep5(e, d, c_full) ? (p1 - p0 + 5) / 6 * 4 : p1 - p0;
                e->stats.scan_bases += p1 - p0;
                e->stats.packed = scan_uses_pep5(e, d, c_full) ? 2 : 0;
            }
        } else {
            // BWD: anchors w with w + (L - k) <= n ; FWD: anchors pos in [1, n]
            long long lo = a0, hi = a1;
            if (dp.type == PM_PLAN_BWD) { hi = std::min(hi, n - (dp.L - dp.k) + 1); }
            else { lo = std::max<long long>(lo, 1); hi = std::min(hi, n + 1); }
            if (hi > lo) {
                DenseArgs a;
                a.pl = dp; a.text = d->d_text; a.n = n; a.a0 = lo; a.a1 = hi; a.TL = dTL; a.TR = dTR;
                a.pl.start_line = 0;          // '^' is not monotone in the scan start: k_verify / k_chain decide it
                a.keys = t.keys; a.count = t.count; a.cap = t.cap; a.fills = fills; a.keytag = tag;
                const long long want = (hi - lo + 255) / 256;
                const int grid = (int)std::min<long long>(want, (long long)e->sms * 16);
                k_scan_dense<<<grid, 256, 0, e->stream>>>(a);
                e->stats.launches++;
                e->stats.scan_bytes += hi - lo;
                e->stats.scan_bases += hi - lo;
            }
        }

    }
    CK(cudaGetLastError());
    return PM_OK;
}

// scan + sort + verify: leaves ncand verified candidates (sorted) in e->cands
static int produce_candidates(pm_engine *e, pm_dataset *d, const Compiled &c_full, long long a0, long long a1,
                              const unsigned long long *dB, const unsigned long long *dTL, const unsigned long long *dTR,
                              long long *ncand_out)
{
    const DevPlan &vdp = c_full.dp;                             // what verification sees
    const long long n = d->n;
    int rc;
    Fills fills;
    if ((rc = ensure_fills(e, d, &fills))) return rc;
    const bool use_packed = scan_uses_packed(e, d, c_full);
    if ((rc = e->counters.reserve(64))) return rc;
    unsigned long long *d_count = (unsigned long long *)e->counters.p;
    long long cap = std::max<long long>((long long)(e->keys.cap / 8), 1 << 16);
    long long ncand = 0;
    for (int attempt = 0; attempt < 3; attempt++) {
        if ((rc = e->keys.reserve((size_t)cap * 8))) return rc;
        CK(cudaMemsetAsync(d_count, 0, 16, e->stream));
        CK(cudaEventRecord(e->ev[0], e->stream));
        e->stats.scan_bytes = 0; e->stats.scan_bases = 0;
        const ScanTarget tgt{(unsigned long long *)e->keys.p, d_count, cap};
        if ((rc = launch_scan(e, d, c_full, a0, a1, fills, dB, dTL, dTR, tgt, 0ULL, ((unsigned long long)(n + 1) << 4) | 15ULL))) return rc;
        CK(cudaGetLastError());
        CK(cudaEventRecord(e->ev[1], e->stream));
        CK(cudaMemcpyAsync(e->h_count, d_count, 16, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        ncand = (long long)e->h_count[0];
        if (ncand <= cap) break;
        cap = ncand + 1024;                                    // grow and rescan
    }
    if (ncand > cap) { g_err = "candidate buffer keeps overflowing"; return PM_ERR_CUDA; }
    if (ncand >= (1LL << 31) - 1) { g_err = "more than 2^31 candidates: not supported"; return PM_ERR_UNSUPPORTED; }
    const long long nplaceholders = (long long)e->h_count[1];  // out-of-range slots written by k_scan_packed_exact
    e->stats.candidates = ncand - nplaceholders;
    // ---- sort ----
    unsigned long long *keys = (unsigned long long *)e->keys.p;
    if (ncand > 1) {
        if ((rc = e->keys2.reserve((size_t)ncand * 8))) return rc;
        size_t tmp = 0;
        int end_bit = 64;
        {
            unsigned long long maxkey = ((unsigned long long)(n + 2) << 4) | 15ULL;
            end_bit = 1;
            while (end_bit < 64 && (maxkey >> end_bit)) end_bit++;
        }
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, keys, (unsigned long long *)e->keys2.p, (int)ncand, 0, end_bit, e->stream));
        if ((rc = e->cubtmp.reserve(tmp))) return rc;
        CK(cub::DeviceRadixSort::SortKeys(e->cubtmp.p, tmp, keys, (unsigned long long *)e->keys2.p, (int)ncand, 0, end_bit, e->stream));
        keys = (unsigned long long *)e->keys2.p;
        e->stats.launches += 3;
    }
    ncand -= nplaceholders;                                   // they carry the largest key and sit at the end
    CK(cudaEventRecord(e->ev[2], e->stream));
    // ---- verify ----
    if (ncand > 0) {
        if ((rc = e->cands.reserve((size_t)ncand * sizeof(Cand)))) return rc;
        k_verify<<<(unsigned)((ncand + 127) / 128), 128, 0, e->stream>>>(vdp, d->d_text, n, dB, dTL, dTR, keys, ncand, (Cand *)e->cands.p,
                                                                        (use_packed || scan_uses_pep5(e, d, c_full)) ? 1 : 0, fills);
        CK(cudaGetLastError());
        e->stats.launches++;
    }
    CK(cudaEventRecord(e->ev[3], e->stream));
    *ncand_out = ncand;
    return PM_OK;
}

// device -> caller's host buffer.  Copies into pageable memory run far below the PCIe rate and block the
// host; results up to 64 MiB go through a page-locked staging buffer instead.
static int copy_to_host(pm_engine *e, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return PM_OK;
    if (bytes > ((size_t)64 << 20)) {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        return PM_OK;
    }
    if (e->h_stage_cap < bytes) {
        if (e->h_stage) cudaFreeHost(e->h_stage);
        e->h_stage = nullptr; e->h_stage_cap = 0;
        size_t want = std::max(bytes * 2, (size_t)1 << 20);
        CK(cudaMallocHost(&e->h_stage, want));
        e->h_stage_cap = want;
    }
    CK(cudaMemcpyAsync(e->h_stage, src, bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    memcpy(dst, e->h_stage, bytes);
    return PM_OK;
}

// chain + select over ncand candidates in d_cands; hits copied to the host
static int resolve_candidates(pm_engine *e, pm_dataset *d, const Compiled &c, const Cand *d_cands, long long ncand,
                              const unsigned long long *dTL, const unsigned long long *dTR,
                              pm_hit *hits, int64_t cap, int64_t *nhits, bool hits_on_device = false)
{
    int rc;
    long long nh = 0;
    Fills fills;
    if ((rc = ensure_fills(e, d, &fills))) return rc;
    if (ncand > 0) {
        if ((rc = e->hits.reserve((size_t)ncand * sizeof(pm_hit)))) return rc;
        if ((rc = e->hits2.reserve((size_t)ncand * sizeof(pm_hit)))) return rc;
        if ((rc = e->sel.reserve((size_t)ncand))) return rc;
        if ((rc = e->counters.reserve(64))) return rc;
        CK(cudaMemsetAsync(e->sel.p, 0, (size_t)ncand, e->stream));
        const long long *d_maxend = nullptr, *d_mindep = nullptr;
        if (c.dp.ext_repeats) {
            // prefix maxima of the hit ends, suffix minima of the examined ranges (keys2 is free at this point)
            if ((rc = e->keys2.reserve((size_t)ncand * 32))) return rc;
            long long *ends = (long long *)e->keys2.p, *mx = ends + ncand, *deps = mx + ncand, *mn = deps + ncand;
            k_cand_ends<<<(unsigned)((ncand + 255) / 256), 256, 0, e->stream>>>(c.dp, d_cands, ncand, ends, deps);
            size_t tmpb = 0, tmpc = 0;
            CK(cub::DeviceScan::InclusiveScan(nullptr, tmpb, ends, mx, MaxLL(), (int)ncand, e->stream));
            CK(cub::DeviceScan::InclusiveScan(nullptr, tmpc, deps, mn, MinLL(), (int)ncand, e->stream));
            tmpb = std::max(tmpb, tmpc);
            if ((rc = e->cubtmp.reserve(tmpb))) return rc;
            CK(cub::DeviceScan::InclusiveScan(e->cubtmp.p, tmpb, ends, mx, MaxLL(), (int)ncand, e->stream));
            CK(cub::DeviceScan::InclusiveScan(e->cubtmp.p, tmpb, deps, mn, MinLL(), (int)ncand, e->stream));
            d_maxend = mx; d_mindep = mn;
            e->stats.launches += 3;
        }
        k_chain<<<(unsigned)((ncand + 127) / 128), 128, 0, e->stream>>>(c.dp, d->d_text, d->n, dTL, dTR, d_cands, ncand,
                                                                       (pm_hit *)e->hits.p, (unsigned char *)e->sel.p, fills, d_maxend, d_mindep);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e->ev[4], e->stream));
        size_t tmp = 0;
        long long *d_nsel = (long long *)((char *)e->counters.p + 16);
        // select on 16-byte hit records
        CK(cub::DeviceSelect::Flagged(nullptr, tmp, (H16 *)e->hits.p, (unsigned char *)e->sel.p, (H16 *)e->hits2.p, d_nsel, (int)ncand, e->stream));
        if ((rc = e->cubtmp.reserve(tmp))) return rc;
        CK(cub::DeviceSelect::Flagged(e->cubtmp.p, tmp, (H16 *)e->hits.p, (unsigned char *)e->sel.p, (H16 *)e->hits2.p, d_nsel, (int)ncand, e->stream));
        CK(cudaMemcpyAsync(e->h_count + 2, d_nsel, 8, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->stats.launches += 3;
        nh = (long long)e->h_count[2];
    } else {
        CK(cudaEventRecord(e->ev[4], e->stream));
    }
    *nhits = nh;
    e->stats.hits = nh;
    e->last_hits = (const pm_hit *)e->hits2.p;
    bool overflow = false;
    CK(cudaEventRecord(e->ev[5], e->stream));
    if (hits && nh > 0) {
        if (nh > cap) overflow = true;                   // the list stays on the device for pm_last_hits
        else if (hits_on_device) CK(cudaMemcpyAsync(hits, e->hits2.p, (size_t)nh * sizeof(pm_hit), cudaMemcpyDeviceToDevice, e->stream));
        else if ((rc = copy_to_host(e, hits, e->hits2.p, (size_t)nh * sizeof(pm_hit)))) return rc;
    }
    CK(cudaStreamSynchronize(e->stream));
    if (overflow) { g_err = "hit buffer too small"; return PM_ERR_OVERFLOW; }
    return PM_OK;
}

static void finish_stats(pm_engine *e)
{
    cudaEventElapsedTime(&e->stats.scan_ms, e->ev[0], e->ev[1]);
    cudaEventElapsedTime(&e->stats.sort_ms, e->ev[1], e->ev[2]);
    cudaEventElapsedTime(&e->stats.verify_ms, e->ev[2], e->ev[3]);
    cudaEventElapsedTime(&e->stats.chain_ms, e->ev[3], e->ev[5]);
    cudaEventElapsedTime(&e->stats.total_ms, e->ev[0], e->ev[5]);
    (void)cudaGetLastError();                            // an unrecorded event must not poison later calls
    e->stats_pending = false;
}

int pm_search(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt, pm_hit *hits, int64_t cap, int64_t *nhits)
{
    if (!e || !d || !pattern || !kopt || !nhits || d->e != e) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    CK(cudaSetDevice(e->device));
    (void)cudaGetLastError();
    Compiled c;
    int rc = compile(pattern, kopt, c, true);
    if (rc) return rc;
    e->stats = pm_stats{};
    *nhits = 0;
    const unsigned long long *dB, *dTL, *dTR;
    if ((rc = upload_tables(e, c, &dB, &dTL, &dTR))) return rc;
    long long ncand = 0;
    if ((rc = produce_candidates(e, d, c, 0, d->n + 1, dB, dTL, dTR, &ncand))) return rc;
    e->stats.verified = ncand;
    rc = resolve_candidates(e, d, c, (const Cand *)e->cands.p, ncand, dTL, dTR, hits, cap, nhits);
    finish_stats(e);
    return rc;
}

// scan + verification + chain stage of the fills f0 .. f1-1 (whole fills are independent units); the fill table of
// the dataset must hold at least f1 entries
static int search_fill_range(pm_engine *e, pm_dataset *d, const Compiled &c, const unsigned long long *dB,
                             const unsigned long long *dTL, const unsigned long long *dTR, long long f0, long long f1,
                             pm_hit *hits, int64_t cap, int64_t *nhits, bool hits_on_device)
{
    const std::vector<long long> &S = d->fill_starts;
    long long ncand = 0;
    int rc;
    if (f0 < f1) {
        // anchors whose fill (fill_of: the last fill starting at or before the anchor; FWD plans anchor one byte later)
        // is one of f0 .. f1-1
        // (EXTENDED plans: keys are window starts, the deciding byte lies ext_off further, one less for anchors at an end)
        const long long shift = c.dp.type == PM_PLAN_FWD ? 1
                              : (c.dp.type == PM_PLAN_EXT_BEG || c.dp.type == PM_PLAN_EXT_END) ? (c.dp.type == PM_PLAN_EXT_END ? 1 : 0) - c.dp.ext_off : 0;
        const long long a0 = std::max<long long>(S[f0] + shift, 0);
        const long long a1 = f1 < (long long)S.size() ? S[f1] + shift : d->n + 1;
        if ((rc = produce_candidates(e, d, c, a0, a1, dB, dTL, dTR, &ncand))) return rc;
    } else {
        for (int i = 0; i < 4; i++) CK(cudaEventRecord(e->ev[i], e->stream));
    }
    e->stats.verified = ncand;
    return resolve_candidates(e, d, c, (const Cand *)e->cands.p, ncand, dTL, dTR, hits, cap, nhits, hits_on_device);
}

// Fill-sharded search for multi-GPU runs.  recSearchFile @402298 restarts its scan at every buffer fill and no
// hit crosses a fill, so whole fills are independent units: this call searches the fills that START in
// [pos_beg, pos_end) completely (scan, verification, chain stage) and leaves their hits, in output order, in
// device memory.  Ranks that cover [0, n] with contiguous ranges produce the complete hit list by
// concatenation, with no candidate exchange and no serial stage on one rank.
int pm_search_fills_device(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt, int64_t pos_beg, int64_t pos_end,
                           pm_hit *dev_hits, int64_t cap, int64_t *nhits, int64_t *dev_count)
{
    if (!e || !d || !pattern || !kopt || !nhits || d->e != e) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    CK(cudaSetDevice(e->device));
    (void)cudaGetLastError();
    Compiled c;
    int rc = compile(pattern, kopt, c, true);
    if (rc) return rc;
    e->stats = pm_stats{};
    *nhits = 0;
    const unsigned long long *dB, *dTL, *dTR;
    if ((rc = upload_tables(e, c, &dB, &dTL, &dTR))) return rc;
    Fills fills;
    if ((rc = ensure_fills(e, d, &fills))) return rc;
    const std::vector<long long> &S = d->fill_starts;
    const long long f0 = std::lower_bound(S.begin(), S.end(), (long long)pos_beg) - S.begin();
    const long long f1 = std::lower_bound(S.begin(), S.end(), (long long)pos_end) - S.begin();
    rc = search_fill_range(e, d, c, dB, dTL, dTR, f0, f1, dev_hits, cap, nhits, true);
    if (dev_count && (rc == PM_OK || rc == PM_ERR_OVERFLOW)) {
        // the count travels with the hits (e.g. as the header row of an all-gather) without a host round trip
        e->h_count[3] = (unsigned long long)*nhits;
        CK(cudaMemcpyAsync(dev_count, e->h_count + 3, 8, cudaMemcpyHostToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    finish_stats(e);
    return rc;
}

int pm_last_hits(pm_engine *e, pm_hit *hits, int64_t cap, int64_t *nhits);

#include "request.cuh"

// Cold request: the file is still in host memory.  Upload it in chunks on a copy stream and, while the next chunk
// is on the PCIe bus, pack the chunk that has arrived into the 2-bit planes and search the buffer fills that it
// completes (fills are independent, see pm_search_fills_device).  What remains after the last byte has landed is the
// last chunk's packing and the search of the last few fills.
int pm_search_stream(pm_engine *e, const uint8_t *host, int64_t n, int npat, const char *const *patterns, const char *kopt,
                     int64_t chunk_bytes, pm_hit *hits, int64_t cap, int64_t *offsets, pm_dataset **out)
{
    if (!e || !out || n < 0 || (!host && n > 0) || npat < 1 || !patterns || !kopt || !offsets) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    CK(cudaSetDevice(e->device));
    (void)cudaGetLastError();
    Request rq;
    {
        int rc = compile_request(npat, patterns, kopt, rq);
        if (rc) return rc;
    }
    std::vector<int64_t> offs((size_t)npat + 1, 0);
    const long long gran = 32LL * 1024;                                    // chunk edges on block-tile boundaries of the planes
    long long chunk = chunk_bytes > 0 ? chunk_bytes : (256LL << 20);
    chunk = std::max(gran, chunk / gran * gran);
    const int nchunks = (int)std::max<long long>(1, (n + chunk - 1) / chunk);
    pm_dataset *d = new pm_dataset();
    d->e = e; d->n = n;
    {
        void *p = nullptr;
        if (e->pool_text && e->pool_text_cap >= (size_t)n + 256) {
            p = e->pool_text; d->owned_cap = e->pool_text_cap;
            e->pool_text = nullptr; e->pool_text_cap = 0;
        } else {
            cudaError_t rc = cudaMalloc(&p, (size_t)n + 256);
            if (rc != cudaSuccess) { delete d; g_err = std::string("cudaMalloc dataset: ") + cudaGetErrorString(rc); return PM_ERR_CUDA; }
            d->owned_cap = (size_t)n + 256;
        }
        d->owned = p; d->d_text = (const unsigned char *)p;
    }
    auto fail = [&](int rc) { pm_dataset_destroy(d); return rc; };
#define CKD(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { g_err = std::string(#x) + " (engine.cu:" + std::to_string(__LINE__) + "): " + cudaGetErrorString(e_); return fail(PM_ERR_CUDA); } } while (0)
    long long nw = (n + 31) / 32;
    nw = (nw + 1023) / 1024 * 1024 + 1024;
    d->nwords = nw;
    {
        void *p = nullptr;
        const size_t need = ((size_t)nw * 3 + PM_PLANE_FRONT) * 4;
        if (e->pool_planes && e->pool_planes_cap >= need) {
            p = e->pool_planes; d->planes_cap = e->pool_planes_cap;
            e->pool_planes = nullptr; e->pool_planes_cap = 0;
        } else {
            CKD(cudaMalloc(&p, need));
            d->planes_cap = need;
        }
        d->planes_base = p;
        d->hi = (unsigned *)p + PM_PLANE_FRONT; d->lo = d->hi + nw; d->xx = d->lo + nw;
        CKD(cudaMemsetAsync(p, 0, PM_PLANE_FRONT * 4, e->stream));
    }
    // all copies are queued at once on their own stream; one event per chunk
    if (!e->copy_stream) CKD(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> arrived((size_t)nchunks, nullptr);
    auto drop_events = [&]() { for (auto ev : arrived) if (ev) cudaEventDestroy(ev); };
    CKD(cudaEventRecord(e->ev[0], e->stream));
    CKD(cudaStreamWaitEvent(e->copy_stream, e->ev[0], 0));                 // the buffers may still be in use by earlier work
    for (int c = 0; c < nchunks; c++) {
        const long long off = (long long)c * chunk, len = std::min<long long>(chunk, n - off);
        if (len > 0) CKD(cudaMemcpyAsync((char *)d->owned + off, host + off, (size_t)len, cudaMemcpyHostToDevice, e->copy_stream));
        if (c == nchunks - 1) CKD(cudaMemsetAsync((char *)d->owned + n, 0, 256, e->copy_stream));
        CKD(cudaEventCreateWithFlags(&arrived[c], cudaEventDisableTiming));
        CKD(cudaEventRecord(arrived[c], e->copy_stream));
    }
    int rc = PM_OK;
    if ((rc = e->counters.reserve(64))) { drop_events(); return fail(rc); }
    unsigned long long *d_exc = (unsigned long long *)((char *)e->counters.p + 32);
    const long long nl_cap = 1 << 20;
    std::vector<std::vector<pm_hit>> found((size_t)npat);
    std::vector<pm_hit> tmp;
    pm_stats total{};
    long long ndone = 0;                                                   // fills already searched
    d->fills_bufsize = e->bufsize;                                         // the fill table is maintained here, chunk by chunk
    {
        const long long bs = e->bufsize > 0 ? e->bufsize : n + 1;
        const size_t want = (size_t)(n / bs + 2) * 3 + 8192;               // room for S, E and cuts of every fill, plus record ends
        cudaError_t ce = cudaMalloc((void **)&d->d_fills, want * 8);
        if (ce != cudaSuccess) { drop_events(); g_err = std::string("cudaMalloc fills: ") + cudaGetErrorString(ce); return fail(PM_ERR_CUDA); }
        d->fills_cap = want;
    }
    for (int c = 0; c < nchunks && rc == PM_OK; c++) {
        const long long off = (long long)c * chunk, end = std::min<long long>(n, off + chunk);
        const bool last = c == nchunks - 1;
        const long long q0 = off / 32, q1 = last ? nw : end / 32;          // plane words of this chunk (the last one also pads)
        cudaError_t ce = cudaStreamWaitEvent(e->stream, arrived[c], 0);
        if (ce == cudaSuccess) ce = cudaMemsetAsync(d_exc, 0, 24, e->stream);
        if (ce != cudaSuccess) { g_err = std::string("stream upload: ") + cudaGetErrorString(ce); rc = PM_ERR_CUDA; break; }
        if ((rc = e->keys2.reserve((size_t)nl_cap * 8))) break;
        unsigned long long *d_nl = (unsigned long long *)e->keys2.p;
        const int grid = (int)std::min<long long>((q1 - q0 + 255) / 256, (long long)e->sms * 16);
        k_pack<<<std::max(grid, 1), 256, 0, e->stream>>>(d->d_text + off, n - off, q1 - q0, d->hi + q0, d->lo + q0, d->xx + q0, d_exc, d_nl, nl_cap, off);
        ce = cudaGetLastError();
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->h_count + 4, d_exc, 16, cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        if (ce != cudaSuccess) { g_err = std::string("stream pack: ") + cudaGetErrorString(ce); rc = PM_ERR_CUDA; break; }
        d->nexc += (long long)e->h_count[4];
        if (c == 0) d->dna_like = end > 0 && d->nexc * 8 <= end;          // decided on the first chunk, like the whole-file rule
        const long long nnl = (long long)e->h_count[5];
        if (nnl > nl_cap) { g_err = "stream upload: more than 2^20 lines in one chunk (use pm_dataset_create)"; rc = PM_ERR_UNSUPPORTED; break; }
        if (nnl > 0) {
            const size_t base = d->newlines.size();
            d->newlines.resize(base + (size_t)nnl);
            ce = cudaMemcpyAsync(d->newlines.data() + base, d_nl, (size_t)nnl * 8, cudaMemcpyDeviceToHost, e->stream);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
            if (ce != cudaSuccess) { g_err = std::string("stream newlines: ") + cudaGetErrorString(ce); rc = PM_ERR_CUDA; break; }
            std::sort(d->newlines.begin() + (long)base, d->newlines.end());
        }
        // fills that this chunk completes
        std::vector<long long> S, E, cuts;
        compute_fills(d->newlines, n, e->bufsize, end, last, S, E, cuts);
        const long long nfinal = (long long)S.size();
        if (nfinal <= ndone && !last) continue;
        if ((rc = upload_fills(e, d, S, E, cuts))) break;
        d->fills_complete = last;
        {
            // the fills this chunk completed, every pattern of the request in one pass (request.cuh)
            for (int p = 0; p < npat; p++) fill_anchor_range(d, rq.comp[p], ndone, nfinal, &rq.a0[p], &rq.a1[p]);
            tmp.resize(std::max<size_t>(tmp.size(), 1 << 16));
            rc = run_request_host(e, d, rq, tmp.data(), (int64_t)tmp.size(), offs.data());
            if (rc == PM_ERR_OVERFLOW) {                                   // the list is still on the device
                int64_t nh = offs[(size_t)npat];
                tmp.resize((size_t)nh);
                rc = pm_last_hits(e, tmp.data(), nh, &nh);
            }
            if (rc) break;
            for (int p = 0; p < npat; p++) found[p].insert(found[p].end(), tmp.begin() + offs[(size_t)p], tmp.begin() + offs[(size_t)p + 1]);
            finish_stats(e);
            total.scan_ms += e->stats.scan_ms; total.sort_ms += e->stats.sort_ms; total.verify_ms += e->stats.verify_ms;
            total.chain_ms += e->stats.chain_ms; total.total_ms += e->stats.total_ms;
            total.candidates += e->stats.candidates; total.verified += e->stats.verified; total.hits += e->stats.hits;
            total.scan_bytes += e->stats.scan_bytes; total.scan_bases += e->stats.scan_bases; total.launches += e->stats.launches;
            total.packed = e->stats.packed; total.qgram_chunks = e->stats.qgram_chunks; total.syncs += e->stats.syncs;
        }
        ndone = nfinal;
    }
    cudaStreamSynchronize(e->copy_stream);
    drop_events();
#undef CKD
    if (rc) return fail(rc);
    d->fills_complete = true;
    e->stats = total;
    *out = d;
    int64_t off = 0;
    bool overflow = false;
    for (int p = 0; p < npat; p++) {
        offsets[p] = off;
        const int64_t k = (int64_t)found[p].size();
        if (hits && off + k <= cap) memcpy(hits + off, found[p].data(), (size_t)k * sizeof(pm_hit));
        else if (hits) overflow = true;
        off += k;
    }
    offsets[npat] = off;
    if (overflow) { g_err = "hit buffer too small (the dataset is resident: search it again with pm_search)"; return PM_ERR_OVERFLOW; }
    return PM_OK;
}

int pm_last_hits(pm_engine *e, pm_hit *hits, int64_t cap, int64_t *nhits)
{
    if (!e || !nhits) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    CK(cudaSetDevice(e->device));
    const long long nh = e->stats.hits;
    *nhits = nh;
    if (hits && nh > 0) {
        if (nh > cap) { g_err = "hit buffer too small"; return PM_ERR_OVERFLOW; }
        CK(cudaMemcpyAsync(hits, e->last_hits ? (const void *)e->last_hits : e->hits2.p, (size_t)nh * sizeof(pm_hit), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    return PM_OK;
}

// class of one pattern position over the packed alphabet: bits A,C,G,T and X (= non-ACGT bytes);
// *mixed is set when the class accepts some but not all non-ACGT bytes (then X is an over-approximation)
static unsigned packed_class_of(const pm::ByteSet &bs, bool *mixed)
{
    unsigned cls = (bs.has('A') ? 1u : 0u) | (bs.has('C') ? 2u : 0u) | (bs.has('G') ? 4u : 0u) | (bs.has('T') ? 8u : 0u);
    int other = __builtin_popcountll(bs.w[0]) + __builtin_popcountll(bs.w[1]) + __builtin_popcountll(bs.w[2]) + __builtin_popcountll(bs.w[3]);
    for (const char *q = "ACGTacgt"; *q; q++) other -= bs.has((unsigned char)*q) ? 1 : 0;      // bytes outside ACGTacgt the class accepts
    if (other) cls |= 16u;
    if (mixed) *mixed = other != 0 && other != 248;
    return cls;
}

// Batched exact motifs on a packed dataset: one scan launch evaluates every pattern on each staged tile.
// Returns 1 when the fast path does not apply (caller falls back to one search per pattern).
// everything search_batch_fused derives from the motif list alone: kept for the next call with the same list (a motif
// library searched against many datasets), device tables included
struct pm_engine::BatchCache {
    std::string key;
    std::vector<MultiPat> mp;            // dense-kernel descriptors (all motifs, or the ones without a lookup window)
    std::vector<unsigned short> mlen;
    std::vector<HashPat> hpat;
    std::vector<unsigned> hoffs, hents, dense_map;
    int ndense = 0;
    bool use_hash = false;
};

// parses the motifs and builds the dense descriptors and the lookup index; 1 = the fused path does not apply
static int build_batch_cache(const pm_engine *e, int npat, const char *const *patterns, pm_engine::BatchCache &bc)
{
    std::string err;
    std::vector<MultiPat> mp((size_t)npat);
    std::vector<unsigned short> mlen((size_t)npat);
    std::vector<unsigned char> allcls((size_t)npat * 32, 0);       // packed class of every position of every motif
    for (int b = 0; b < npat; b++) {
        pm::Pattern P;
        if (pm::parse_pattern(patterns[b], true, P, err)) return 1;
        if (P.start_line || P.end_line || P.extended() || P.m() > 32 || P.m() < 1) return 1;
        MultiPat &m = mp[b];
        memset(&m, 0, sizeof m);
        m.m = (unsigned short)P.m();
        mlen[b] = m.m;
        unsigned char mixedv[32];
        for (int j = 0; j < P.m(); j++) {
            bool mixed = false;
            allcls[(size_t)b * 32 + j] = (unsigned char)packed_class_of(P.pos[j], &mixed);
            mixedv[j] = mixed ? 1 : 0;
        }
        // entries grouped by plane so that consecutive entries tend to take the same branch
        for (int s = 0; s < 6; s++)
            for (int j = 0; j < P.m(); j++) {
                const unsigned cls = allcls[(size_t)b * 32 + j];
                if (mixedv[j]) return 1;
                const int sel = cls == 1 ? 0 : cls == 2 ? 1 : cls == 4 ? 2 : cls == 8 ? 3 : cls == 16 ? 4 : 5;
                if (sel != s) continue;
                if (cls == 31) continue;                     // accepts everything: no constraint (window validity is checked separately)
                m.ent[m.nent++] = (unsigned short)(sel | (j << 3) | (cls << 8));
            }
    }
    // ---- large batches: q-gram lookup (packed.cuh: k_scan_multi_hash).  Every motif with an 8-position window whose
    // classes stay inside ACGT is indexed by the 8-mers that window accepts; the others keep the dense kernel.
    std::vector<HashPat> hpat;
    std::vector<unsigned> hoffs, hents, dense_map;
    std::vector<MultiPat> dense_mp;
    const bool use_hash = npat >= 64 && e->batch_hash;
    if (use_hash) {
        std::vector<std::pair<unsigned, unsigned>> pairs;       // (code, entry)
        std::vector<unsigned> clsv;
        for (int b = 0; b < npat; b++) {
            const int m = mlen[(size_t)b];
            clsv.assign((size_t)m, 0);
            for (int j = 0; j < m; j++) clsv[(size_t)j] = allcls[(size_t)b * 32 + j];
            int best_o = -1, best_lev = -1;
            double best_cost = 0;
            for (int lev = 0; lev < MH_NLEV && best_o < 0; lev++) {
                const int Q = mh_q(lev);
                for (int o = 0; o + Q <= m; o++) {
                    double cost = 1;
                    bool ok = true;
                    for (int t = 0; t < Q && ok; t++) {
                        const unsigned c = clsv[(size_t)(o + t)];
                        if ((c & 16u) || !(c & 15u)) ok = false;
                        cost *= __builtin_popcount(c & 15u);
                    }
                    if (ok && cost <= (lev == 0 ? 64 : 16) && (best_o < 0 || cost < best_cost)) { best_o = o; best_cost = cost; best_lev = lev; }
                }
            }
            if (best_o < 0) { dense_mp.push_back(mp[(size_t)b]); dense_map.push_back((unsigned)b); continue; }
            const int Q = mh_q(best_lev);
            HashPat hp;
            memset(&hp, 0, sizeof hp);
            for (int j = 0; j < m; j++) {
                const unsigned c = clsv[(size_t)j];
                if (c & 1u) hp.ma |= 1u << j;
                if (c & 2u) hp.mc |= 1u << j;
                if (c & 4u) hp.mg |= 1u << j;
                if (c & 8u) hp.mt |= 1u << j;
                if (c & 16u) hp.mx |= 1u << j;
            }
            hp.lenmask = m >= 32 ? 0xffffffffu : ((1u << m) - 1u);
            hp.m = (unsigned)m; hp.pid = (unsigned)b;
            const unsigned hidx = (unsigned)hpat.size();
            hpat.push_back(hp);
            // every concrete q-mer of the window: letters as (hi, lo) = A 00, C 01, G 11, T 10
            static const unsigned hi_of[4] = {0, 0, 1, 1}, lo_of[4] = {0, 1, 1, 0};      // by class bit index A, C, G, T
            unsigned idx[MH_Q] = {0};
            for (;;) {
                unsigned code = 0;
                for (int t = 0; t < Q; t++) {
                    const unsigned c = clsv[(size_t)(best_o + t)] & 15u;
                    unsigned cc = c, bit = 0;                        // idx[t]-th set bit of c
                    for (unsigned q = 0; q <= idx[t]; q++) { bit = (unsigned)__builtin_ctz(cc); cc &= cc - 1; }
                    code |= (hi_of[bit] << (Q + t)) | (lo_of[bit] << t);
                }
                pairs.push_back({mh_base(best_lev) + code, ((unsigned)best_o << 20) | hidx});
                int t = 0;
                while (t < Q) {
                    if (++idx[t] < (unsigned)__builtin_popcount(clsv[(size_t)(best_o + t)] & 15u)) break;
                    idx[t] = 0;
                    t++;
                }
                if (t == Q) break;
            }
        }
        hoffs.assign(MH_BUCKETS + 1, 0);
        for (const auto &pr : pairs) hoffs[pr.first + 1]++;
        for (int c = 0; c < MH_BUCKETS; c++) hoffs[(size_t)c + 1] += hoffs[(size_t)c];
        hents.resize(pairs.size() + 1);
        std::vector<unsigned> fillp(hoffs.begin(), hoffs.end() - 1);
        for (const auto &pr : pairs) hents[fillp[pr.first]++] = pr.second;
    }
    bc.use_hash = use_hash;
    bc.ndense = use_hash ? (int)dense_mp.size() : npat;
    bc.mp = use_hash ? dense_mp : mp;
    bc.mlen = mlen;
    bc.hpat = hpat; bc.hoffs = hoffs; bc.hents = hents; bc.dense_map = dense_map;
    return 0;
}

// compact results: hit begins as 32-bit offsets from a0 (every motif of the fused path is an exact SIMPLE pattern, so a
// hit ends m positions behind its begin): 4 instead of 16 bytes per hit cross PCIe
__global__ void k_compact_hits(const H16 *__restrict__ hits, long long n, long long base, unsigned *__restrict__ out)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = (unsigned)(hits[j].a - base);
}

static int search_batch_fused(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns, const char *kopt,
                              long long a0, long long a1, pm_hit *hits, int64_t cap, int64_t *offsets,
                              uint32_t *compact = nullptr, uint16_t *motif_len = nullptr)
{
    pm::Options o;
    std::string err;
    if (pm::parse_kopt(kopt, o, err) || o.k != 0) return 1;
    if (!d->hi || e->scan_mode == 1 || (e->scan_mode == 0 && !d->dna_like) || npat < 2 || npat >= (1 << 20) || d->n >= (1LL << 36)) return 1;
    std::string bkey = std::string(kopt) + (e->batch_hash ? "\x01H" : "\x01D");
    for (int b = 0; b < npat; b++) { if (!patterns[b]) return 1; bkey += '\n'; bkey += patterns[b]; }
    std::shared_ptr<pm_engine::BatchCache> bc = e->bcache;
    if (!bc || bc->key != bkey) {
        bc = std::make_shared<pm_engine::BatchCache>();
        if (build_batch_cache(e, npat, patterns, *bc)) return 1;
        bc->key = bkey;
        e->bcache = bc;
    }
    const std::vector<MultiPat> &mp = bc->mp;
    const std::vector<unsigned short> &mlen = bc->mlen;
    const std::vector<HashPat> &hpat = bc->hpat;
    const std::vector<unsigned> &hoffs = bc->hoffs, &hents = bc->hents, &dense_map = bc->dense_map;
    const bool use_hash = bc->use_hash;
    const int ndense = bc->ndense;
    CK(cudaSetDevice(e->device));
    (void)cudaGetLastError();
    e->stats = pm_stats{};
    int rc;
    Fills fills;
    if ((rc = ensure_fills(e, d, &fills))) return rc;
    if ((rc = e->tables.reserve((size_t)npat * (sizeof(MultiPat) + 2 + 8) + 256))) return rc;
    const unsigned *d_hoffs = nullptr, *d_hents = nullptr, *d_dmap = nullptr;
    const HashPat *d_hpat = nullptr;
    if (use_hash) {
        const size_t b0 = (MH_BUCKETS + 1) * 4, b1 = hents.size() * 4, b2 = (hpat.size() + 1) * sizeof(HashPat), b3 = (dense_map.size() + 1) * 4;
        auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
        if (e->batch_tab.cap < al(b0) + al(b1) + al(b2) + al(b3)) e->batch_tab_key.clear();
        if ((rc = e->batch_tab.reserve(al(b0) + al(b1) + al(b2) + al(b3)))) return rc;
        char *base = (char *)e->batch_tab.p;
        if (e->batch_tab_key != bkey) {
            CK(cudaMemcpyAsync(base, hoffs.data(), b0, cudaMemcpyHostToDevice, e->stream));
            CK(cudaMemcpyAsync(base + al(b0), hents.data(), b1, cudaMemcpyHostToDevice, e->stream));
            if (!hpat.empty()) CK(cudaMemcpyAsync(base + al(b0) + al(b1), hpat.data(), hpat.size() * sizeof(HashPat), cudaMemcpyHostToDevice, e->stream));
            if (!dense_map.empty()) CK(cudaMemcpyAsync(base + al(b0) + al(b1) + al(b2), dense_map.data(), dense_map.size() * 4, cudaMemcpyHostToDevice, e->stream));
            e->batch_tab_key = bkey;
        }
        d_hoffs = (const unsigned *)base; d_hents = (const unsigned *)(base + al(b0));
        d_hpat = (const HashPat *)(base + al(b0) + al(b1)); d_dmap = (const unsigned *)(base + al(b0) + al(b1) + al(b2));
    }
    MultiPat *d_pats = (MultiPat *)e->tables.p;
    unsigned short *d_mlen = (unsigned short *)((char *)e->tables.p + (((size_t)npat * sizeof(MultiPat) + 15) & ~(size_t)15));
    unsigned long long *d_perpat = (unsigned long long *)((char *)d_mlen + (((size_t)npat * 2 + 15) & ~(size_t)15));
    if (ndense) CK(cudaMemcpyAsync(d_pats, mp.data(), (size_t)ndense * sizeof(MultiPat), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(d_mlen, mlen.data(), (size_t)npat * 2, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemsetAsync(d_perpat, 0, (size_t)npat * 8, e->stream));
    if ((rc = e->counters.reserve(64))) return rc;
    unsigned long long *d_count = (unsigned long long *)e->counters.p;
    const long long n = d->n;
    long long capk = std::max<long long>((long long)(e->keys.cap / 8), 1 << 20);
    if (const char *dbg = getenv("PM_BATCH_CAP")) capk = std::max<long long>(atoll(dbg), 1024);     // experiments: initial key capacity
    long long nkeys = 0;
    // motif starts a0 <= w < a1 (whole buffer fills; the whole dataset: 0 .. n + 1)
    const long long last = std::max<long long>(std::min<long long>(a1, n) - 1, a0);
    const long long tile0 = (a0 / 32) / 128;
    const long long ntiles = (last / 32) / 128 + 1 - tile0;
    for (int attempt = 0; attempt < 3 && a1 > a0; attempt++) {
        if ((rc = e->keys.reserve((size_t)capk * 8))) return rc;
        CK(cudaMemsetAsync(d_count, 0, 16, e->stream));
        CK(cudaEventRecord(e->ev[0], e->stream));
        if (use_hash && !hpat.empty()) {
            HashArgs h;
            h.hi = d->hi; h.lo = d->lo; h.xx = d->xx; h.nwords = d->nwords; h.n = n;
            h.tile0 = (a0 / 32) / MH_WORDS;
            h.ntiles = (last / 32) / MH_WORDS + 1 - h.tile0;
            h.a0 = a0; h.a1 = a1;
            h.offs = d_hoffs; h.ents = d_hents; h.pats = d_hpat;
            h.keys = (unsigned long long *)e->keys.p; h.count = d_count; h.cap = capk;
            const size_t smem = MH_STAGES * MH_STAGE_BYTES + 2 * MH_STAGES * 8;
            if (!e->attr_hash) {
                CK(cudaFuncSetAttribute(k_scan_multi_hash, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                e->attr_hash = true;
            }
            const int gridh = std::max((int)std::min<long long>(h.ntiles, (long long)e->sms * 4), 1);
            k_scan_multi_hash<<<gridh, (MH_WARPS + 1) * 32, smem, e->stream>>>(h);
            CK(cudaGetLastError());
            e->stats.launches++;
        }
        if (ndense) {
        MultiArgs a;
        a.hi = d->hi; a.lo = d->lo; a.xx = d->xx; a.nwords = d->nwords; a.n = n; a.a0 = a0; a.a1 = a1; a.tile0 = tile0; a.ntiles = ntiles;
        a.pats = d_pats; a.npat = ndense; a.keys = (unsigned long long *)e->keys.p; a.count = d_count; a.cap = capk;
        a.pid_map = use_hash ? d_dmap : nullptr;
        a.bad = (unsigned long long)npat << 40;
        const int grid = std::max((int)std::min<long long>((ntiles + 7) / 8, (long long)e->sms * 4), 1);
        k_scan_packed_multi<<<grid, 256, 0, e->stream>>>(a);
        CK(cudaGetLastError());
        }
        CK(cudaEventRecord(e->ev[1], e->stream));
        CK(cudaMemcpyAsync(e->h_count, d_count, 16, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->stats.launches++;
        nkeys = (long long)e->h_count[0];
        if (nkeys <= capk) break;
        capk = nkeys + 4096;
    }
    if (nkeys > capk) { g_err = "hit buffer keeps overflowing"; return PM_ERR_CUDA; }
    if (nkeys >= (1LL << 31) - 1) { g_err = "more than 2^31 hits in one batch: split the batch"; return PM_ERR_UNSUPPORTED; }
    const long long nplace = a1 > a0 ? (long long)e->h_count[1] : 0;
    e->stats.scan_bytes = (a1 > a0 ? ntiles : 0) * 128 * 4 * 3 * ((use_hash && !hpat.empty() ? 1 : 0) + (ndense ? 1 : 0));
    e->stats.scan_bases = (std::min<long long>(a1, n) - a0) * (long long)npat;
    e->stats.packed = 1;
    e->stats.qgram_chunks = use_hash ? (int)hpat.size() : 0;      // batches: motifs served by the q-gram lookup kernel
    unsigned long long *keys = (unsigned long long *)e->keys.p;
    if (nkeys > 1) {
        if ((rc = e->keys2.reserve((size_t)nkeys * 8))) return rc;
        size_t tmp = 0;
        int end_bit = 41;
        while (end_bit < 64 && ((unsigned long long)npat >> (end_bit - 40))) end_bit++;
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, keys, (unsigned long long *)e->keys2.p, (int)nkeys, 0, end_bit, e->stream));
        if ((rc = e->cubtmp.reserve(tmp))) return rc;
        CK(cub::DeviceRadixSort::SortKeys(e->cubtmp.p, tmp, keys, (unsigned long long *)e->keys2.p, (int)nkeys, 0, end_bit, e->stream));
        keys = (unsigned long long *)e->keys2.p;
        e->stats.launches += 3;
    }
    nkeys -= nplace;
    CK(cudaEventRecord(e->ev[2], e->stream));
    CK(cudaEventRecord(e->ev[3], e->stream));
    e->stats.candidates = nkeys;
    e->stats.verified = nkeys;
    long long nh = 0;
    std::vector<unsigned long long> perpat((size_t)npat, 0);
    if (nkeys > 0) {
        if ((rc = e->hits.reserve((size_t)nkeys * sizeof(pm_hit)))) return rc;
        if ((rc = e->hits2.reserve((size_t)nkeys * sizeof(pm_hit)))) return rc;
        if ((rc = e->sel.reserve((size_t)nkeys))) return rc;
        CK(cudaMemsetAsync(e->sel.p, 0, (size_t)nkeys, e->stream));
        k_chain_multi<<<(unsigned)((nkeys + 127) / 128), 128, 0, e->stream>>>(keys, nkeys, d_mlen, fills.S, fills.E, fills.n,
                                                                             (long long *)e->hits.p, (unsigned char *)e->sel.p, d_perpat);
        CK(cudaGetLastError());
        size_t tmp = 0;
        long long *d_nsel = (long long *)((char *)e->counters.p + 16);
        CK(cub::DeviceSelect::Flagged(nullptr, tmp, (H16 *)e->hits.p, (unsigned char *)e->sel.p, (H16 *)e->hits2.p, d_nsel, (int)nkeys, e->stream));
        if ((rc = e->cubtmp.reserve(tmp))) return rc;
        CK(cub::DeviceSelect::Flagged(e->cubtmp.p, tmp, (H16 *)e->hits.p, (unsigned char *)e->sel.p, (H16 *)e->hits2.p, d_nsel, (int)nkeys, e->stream));
        CK(cudaMemcpyAsync(e->h_count + 2, d_nsel, 8, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(perpat.data(), d_perpat, (size_t)npat * 8, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->stats.launches += 3;
        nh = (long long)e->h_count[2];
    }
    CK(cudaEventRecord(e->ev[4], e->stream));
    offsets[0] = 0;
    for (int b = 0; b < npat; b++) offsets[b + 1] = offsets[b] + (int64_t)perpat[b];
    e->stats.hits = nh;
    e->last_hits = (const pm_hit *)e->hits2.p;
    bool overflow = false;
    if (motif_len) for (int b = 0; b < npat; b++) motif_len[b] = mlen[(size_t)b];
    if (compact && nh > 0) {
        if (nh > cap) overflow = true;
        else {
            // e->hits (the unselected list) is free again: it receives the 32-bit begins
            k_compact_hits<<<(unsigned)((nh + 255) / 256), 256, 0, e->stream>>>((const H16 *)e->hits2.p, nh, a0, (unsigned *)e->hits.p);
            CK(cudaGetLastError());
            e->stats.launches++;
            CK(cudaMemcpyAsync(compact, e->hits.p, (size_t)nh * 4, cudaMemcpyDeviceToHost, e->stream));
        }
    } else if (hits && nh > 0) {
        if (nh > cap) overflow = true;
        else CK(cudaMemcpyAsync(hits, e->hits2.p, (size_t)nh * sizeof(pm_hit), cudaMemcpyDeviceToHost, e->stream));
    }
    CK(cudaEventRecord(e->ev[5], e->stream));
    CK(cudaStreamSynchronize(e->stream));
    finish_stats(e);
    if (overflow) { g_err = "hit buffer too small"; return PM_ERR_OVERFLOW; }
    return PM_OK;
}

int pm_search_request(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns, const char *kopt,
                      pm_hit *hits, int64_t cap, int64_t *offsets);

static int search_request_range(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns, const char *kopt,
                                long long f0, long long f1, pm_hit *hits, int64_t cap, int64_t *offsets);

// batch over the buffer fills f0 .. f1-1 of the dataset (the whole dataset: 0 .. number of fills)
static int search_batch_range(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns, const char *kopt,
                              long long f0, long long f1, pm_hit *hits, int64_t cap, int64_t *offsets)
{
    {
        const std::vector<long long> &S = d->fill_starts;
        const long long a0 = f0 < f1 ? S[(size_t)f0] : 0;
        const long long a1 = f0 < f1 ? (f1 < (long long)S.size() ? S[(size_t)f1] : d->n + 1) : 0;
        const int frc = search_batch_fused(e, d, npat, patterns, kopt, a0, a1, hits, cap, offsets);
        if (frc != 1) return frc;
    }
    // general batches (errors allowed, any plan type): groups of patterns through the request pipeline
    const int group = 32;
    int64_t total = 0;
    offsets[0] = 0;
    pm_stats acc{};
    bool overflow = false;
    std::vector<int64_t> off((size_t)group + 1);
    for (int i0 = 0; i0 < npat; i0 += group) {
        const int g = std::min(group, npat - i0);
        const bool room = hits && total <= cap && !overflow;
        int rc = search_request_range(e, d, g, patterns + i0, kopt, f0, f1, room ? hits + total : nullptr, room ? cap - total : 0, off.data());
        if (rc == PM_ERR_OVERFLOW) { overflow = true; rc = PM_OK; }
        if (rc) return rc;
        for (int q = 0; q < g; q++) offsets[i0 + q + 1] = total + off[(size_t)q + 1];
        total += off[(size_t)g];
        pm_stats st{};
        pm_get_stats(e, &st);
        acc.scan_ms += st.scan_ms; acc.sort_ms += st.sort_ms; acc.verify_ms += st.verify_ms;
        acc.chain_ms += st.chain_ms; acc.total_ms += st.total_ms;
        acc.candidates += st.candidates; acc.verified += st.verified; acc.hits += st.hits;
        acc.scan_bytes += st.scan_bytes; acc.scan_bases += st.scan_bases; acc.launches += st.launches; acc.syncs += st.syncs;
        acc.packed = st.packed;
    }
    e->stats = acc;
    if (overflow || (hits && total > cap)) {
        if (npat > group) e->stats.hits = -1;            // the device holds only the last group's list: pm_last_hits does not apply
        g_err = "hit buffer too small";
        return PM_ERR_OVERFLOW;
    }
    return PM_OK;
}

static int batch_fill_range(pm_engine *e, pm_dataset *d, int64_t pos_beg, int64_t pos_end, long long *f0, long long *f1)
{
    Fills fills;
    int rc;
    CK(cudaSetDevice(e->device));
    (void)cudaGetLastError();
    if ((rc = ensure_fills(e, d, &fills))) return rc;
    const std::vector<long long> &S = d->fill_starts;
    if (pos_end < 0) { *f0 = 0; *f1 = (long long)S.size(); }
    else {
        *f0 = std::lower_bound(S.begin(), S.end(), (long long)pos_beg) - S.begin();
        *f1 = std::lower_bound(S.begin(), S.end(), (long long)pos_end) - S.begin();
    }
    if (d->windowed && *f1 > *f0 && (S[(size_t)*f0] < d->win_lo || d->fill_ends[(size_t)*f1 - 1] > d->win_hi)) {
        g_err = "the buffer fills of this position range are not inside the dataset's window";
        return PM_ERR_ARG;
    }
    return PM_OK;
}

int pm_search_batch(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns, const char *kopt,
                    pm_hit *hits, int64_t cap, int64_t *offsets)
{
    if (!e || !d || npat < 0 || !patterns || !offsets || !kopt || d->e != e) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    long long f0, f1;
    int rc = batch_fill_range(e, d, 0, -1, &f0, &f1);
    if (rc) return rc;
    return search_batch_range(e, d, npat, patterns, kopt, f0, f1, hits, cap, offsets);
}

int pm_search_batch_fills_compact(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns, const char *kopt,
                                  int64_t pos_beg, int64_t pos_end, uint32_t *begins, int64_t cap, int64_t *offsets,
                                  int64_t *base, uint16_t *motif_len)
{
    if (!e || !d || npat < 0 || !patterns || !offsets || !kopt || d->e != e || pos_beg < 0 || (pos_end >= 0 && pos_end < pos_beg) || !begins || !base) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    long long f0, f1;
    int rc = batch_fill_range(e, d, pos_beg, pos_end, &f0, &f1);
    if (rc) return rc;
    const std::vector<long long> &S = d->fill_starts;
    const long long a0 = f0 < f1 ? S[(size_t)f0] : 0;
    const long long a1 = f0 < f1 ? (f1 < (long long)S.size() ? S[(size_t)f1] : d->n + 1) : 0;
    if (a1 - a0 >= (1LL << 32)) { g_err = "compact hit lists: the position range must span less than 2^32 bytes"; return PM_ERR_UNSUPPORTED; }
    *base = a0;
    rc = search_batch_fused(e, d, npat, patterns, kopt, a0, a1, nullptr, cap, offsets, begins, motif_len);
    if (rc == 1) { g_err = "compact hit lists: only batches of exact motifs on a DNA dataset (the fused batch path)"; return PM_ERR_UNSUPPORTED; }
    return rc;
}

int pm_search_batch_fills(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns, const char *kopt,
                          int64_t pos_beg, int64_t pos_end, pm_hit *hits, int64_t cap, int64_t *offsets)
{
    if (!e || !d || npat < 0 || !patterns || !offsets || !kopt || d->e != e || pos_beg < 0 || pos_end < pos_beg) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    long long f0, f1;
    int rc = batch_fill_range(e, d, pos_beg, pos_end, &f0, &f1);
    if (rc) return rc;
    return search_batch_range(e, d, npat, patterns, kopt, f0, f1, hits, cap, offsets);
}

static int candidates_impl(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt, int64_t pos_beg, int64_t pos_end,
                           pm_candidate *cands, int64_t cap, int64_t *ncands, cudaMemcpyKind kind);

int pm_candidates(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt, int64_t pos_beg, int64_t pos_end,
                  pm_candidate *cands, int64_t cap, int64_t *ncands)
{
    return candidates_impl(e, d, pattern, kopt, pos_beg, pos_end, cands, cap, ncands, cudaMemcpyDeviceToHost);
}

int pm_candidates_device(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt, int64_t pos_beg, int64_t pos_end,
                         pm_candidate *dev_cands, int64_t cap, int64_t *ncands)
{
    return candidates_impl(e, d, pattern, kopt, pos_beg, pos_end, dev_cands, cap, ncands, cudaMemcpyDeviceToDevice);
}

static int candidates_impl(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt, int64_t pos_beg, int64_t pos_end,
                           pm_candidate *cands, int64_t cap, int64_t *ncands, cudaMemcpyKind kind)
{
    if (!e || !d || !pattern || !kopt || !ncands || d->e != e) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    CK(cudaSetDevice(e->device));
    Compiled c;
    int rc = compile(pattern, kopt, c, true);
    if (rc) return rc;
    e->stats = pm_stats{};
    const unsigned long long *dB, *dTL, *dTR;
    if ((rc = upload_tables(e, c, &dB, &dTL, &dTR))) return rc;
    long long ncand = 0;
    if ((rc = produce_candidates(e, d, c, std::max<int64_t>(pos_beg, 0), std::min<int64_t>(pos_end, d->n + 1), dB, dTL, dTR, &ncand))) return rc;
    *ncands = ncand;
    if (cands && ncand > 0) {
        if (ncand > cap) { g_err = "candidate buffer too small"; return PM_ERR_OVERFLOW; }
        CK(cudaMemcpyAsync(cands, e->cands.p, (size_t)ncand * sizeof(Cand), kind, e->stream));
    }
    CK(cudaEventRecord(e->ev[4], e->stream));
    CK(cudaEventRecord(e->ev[5], e->stream));
    CK(cudaStreamSynchronize(e->stream));
    finish_stats(e);
    return PM_OK;
}

static int resolve_impl(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt, const pm_candidate *cands, int64_t ncands,
                        pm_hit *hits, int64_t cap, int64_t *nhits, cudaMemcpyKind kind);

int pm_resolve(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt, const pm_candidate *cands, int64_t ncands,
               pm_hit *hits, int64_t cap, int64_t *nhits)
{
    return resolve_impl(e, d, pattern, kopt, cands, ncands, hits, cap, nhits, cudaMemcpyHostToDevice);
}

int pm_resolve_device(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt, const pm_candidate *dev_cands, int64_t ncands,
                      pm_hit *hits, int64_t cap, int64_t *nhits)
{
    return resolve_impl(e, d, pattern, kopt, dev_cands, ncands, hits, cap, nhits, cudaMemcpyDeviceToDevice);
}

static int resolve_impl(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt, const pm_candidate *cands, int64_t ncands,
                        pm_hit *hits, int64_t cap, int64_t *nhits, cudaMemcpyKind kind)
{
    if (!e || !d || !pattern || !kopt || !nhits || ncands < 0 || (ncands && !cands) || d->e != e) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    CK(cudaSetDevice(e->device));
    Compiled c;
    int rc = compile(pattern, kopt, c, true);
    if (rc) return rc;
    const pm_stats before = e->stats;                    // keep the scan figures of the preceding pm_candidates
    const unsigned long long *dB, *dTL, *dTR;
    if ((rc = upload_tables(e, c, &dB, &dTL, &dTR))) return rc;
    if ((rc = e->cands.reserve((size_t)std::max<int64_t>(ncands, 1) * sizeof(Cand)))) return rc;
    CK(cudaEventRecord(e->ev[3], e->stream));
    if (ncands) CK(cudaMemcpyAsync(e->cands.p, cands, (size_t)ncands * sizeof(Cand), kind, e->stream));
    rc = resolve_candidates(e, d, c, (const Cand *)e->cands.p, ncands, dTL, dTR, hits, cap, nhits);
    float chain_ms = 0;
    cudaEventElapsedTime(&chain_ms, e->ev[3], e->ev[5]);
    (void)cudaGetLastError();
    const long long nh = e->stats.hits;
    const int launches = e->stats.launches;
    e->stats = before;
    e->stats.chain_ms = chain_ms;
    e->stats.hits = nh;
    e->stats.launches = launches;                        // resolve_candidates added its launches to the kept count
    return rc;
}

