// packed.cuh -- 2-bit packed, bit-sliced DNA scan (included by engine.cu).
//
// Dataset side: k_pack turns the resident .seq bytes into three bit planes, one bit per
// byte, 32 bytes per word (bit b of word q <-> byte 32q+b):
//     HI = bit 2 of the ASCII code, LO = bit 1   (A=00 C=01 T=10 G=11, case-insensitive)
//     X  = byte is not one of ACGTacgt           (headers, '\n', N, IUPAC codes ...)
// HI/LO are forced to 0 where X is set.  0.375 bytes per base instead of 1.
//
// Scan side: k_scan_packed evaluates a piece for 32 window starts per instruction.
// With one-hot planes PA,PC,PG,PT,PX the piece matches at the 32 starts of word q iff
//     AND_j  ( E_j >> j )      E_j = OR of the planes position j accepts,
// the shift crossing into the following words (funnel shift).  Single-letter positions use
// a plane directly; IUPAC classes cost five logic ops per word.  Positions whose class
// accepts only SOME non-ACGT bytes are treated as accepting all of them; k_verify then
// re-checks every candidate against the raw bytes, so the candidate set stays exact.
#pragma once
#include "apx.hpp"

struct PackedPos { unsigned char sel; unsigned char cls; };   // sel 0..4 = plane A,C,G,T,X ; 5 = class `cls` (bits A,C,G,T,X)

template <int NP>
struct PackedArgs {
    const unsigned *hi, *lo, *xx;
    long long nwords;            // words per plane (padded, multiple of 128, + 128 spare)
    long long n;                 // text bytes
    long long a0, a1;            // window starts a0 <= w < a1
    long long tile0, ntiles;     // tiles of 128 words (4096 bases)
    int L, npieces;
    unsigned trigsets[NP];       // bit j set: a match of piece j at w makes piece i a candidate
    PackedPos pos[NP][64];
    unsigned long long *keys, *count;
    long long cap;
    unsigned long long keytag;   // pattern id of a multi-pattern request, already shifted (pid << PM_PID_SHIFT)
};

// SWAR helpers on 4 packed bytes
__device__ __forceinline__ unsigned swar_eq(unsigned w, unsigned pat)          // 0x80 in every byte of w equal to the byte of pat
{
    const unsigned z = w ^ pat;
    const unsigned t = (z & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(t | z | 0x7f7f7f7fu);
}
__device__ __forceinline__ unsigned swar_gather(unsigned m)                    // bit 7 of each byte -> bits 0..3
{
    return ((m >> 7) * 0x00204081u >> 21) & 0xfu;                              // bytes 0..3 land on bits 0..3
}

// One thread packs 32 consecutive text bytes into one word of each plane and records the positions of
// '\n' (record delimiter) it meets.  nl_out may overflow its capacity: the count keeps running and the
// host then falls back to k_newlines.
__global__ void __launch_bounds__(256) k_pack(const unsigned char *__restrict__ text, long long n, long long nwords,
                                              unsigned *__restrict__ hi, unsigned *__restrict__ lo, unsigned *__restrict__ xx,
                                              unsigned long long *__restrict__ nexc,    // [0] non-ACGT bytes, [1] newlines
                                              unsigned long long *__restrict__ nl_out, long long nl_cap,
                                              long long pos0 = 0)                       // file offset of text[0] (chunked packing)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool aligned = ((size_t)text & 15) == 0;
    unsigned long long exc = 0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nwords; q += stride) {
        const long long base = q * 32;
        unsigned w[8];
        if (aligned && base + 32 <= n) {
            const uint4 a = __ldg(reinterpret_cast<const uint4 *>(text + base));
            const uint4 b = __ldg(reinterpret_cast<const uint4 *>(text + base + 16));
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                unsigned v = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const long long idx = base + 4 * k + b;
                    v |= (unsigned)(idx < n ? text[idx] : 0xffu) << (8 * b);
                }
                w[k] = v;
            }
        }
        unsigned mh = 0, ml = 0, mx = 0, mn = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const unsigned f = w[k] | 0x20202020u;
            const unsigned acgt = swar_eq(f, 0x61616161u) | swar_eq(f, 0x63636363u) | swar_eq(f, 0x67676767u) | swar_eq(f, 0x74747474u);
            const unsigned h = (w[k] << 5) & acgt;          // ASCII bit 2 -> bit 7
            const unsigned l = (w[k] << 6) & acgt;          // ASCII bit 1 -> bit 7
            mh |= swar_gather(h) << (4 * k);
            ml |= swar_gather(l) << (4 * k);
            mx |= swar_gather(~acgt & 0x80808080u) << (4 * k);
            mn |= swar_gather(swar_eq(w[k], 0x0a0a0a0au)) << (4 * k);
        }
        if (base + 32 > n) {                                 // bytes past the end are X, never newlines
            const int valid = n > base ? (int)(n - base) : 0;
            const unsigned keep = valid >= 32 ? ~0u : ((1u << valid) - 1u);
            mn &= keep;
            exc += __popc(mx & keep);
        } else exc += __popc(mx);
        hi[q] = mh; lo[q] = ml; xx[q] = mx;
        if (mn) {
            const int cnt = __popc(mn);
            unsigned long long slot = atomicAdd(nexc + 1, (unsigned long long)cnt);
            while (mn) {
                const int b = __ffs(mn) - 1;
                mn &= mn - 1;
                if ((long long)slot < nl_cap) nl_out[slot] = (unsigned long long)(pos0 + base + b);
                slot++;
            }
        }
    }
    for (int o = 16; o; o >>= 1) exc += __shfl_xor_sync(0xffffffffu, exc, o);
    if ((threadIdx.x & 31) == 0 && exc) atomicAdd(nexc, exc);
}

// positions of the record delimiter (fallback when k_pack's newline buffer overflowed; unordered)
__global__ void __launch_bounds__(256) k_newlines(const unsigned char *__restrict__ text, long long n,
                                                  unsigned long long *__restrict__ out, unsigned long long *__restrict__ count)
{
    const long long stride = (long long)gridDim.x * blockDim.x * 16;
    for (long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16; base < n; base += stride) {
#pragma unroll
        for (int b = 0; b < 16; b++)
            if (base + b < n && text[base + b] == '\n') out[atomicAdd(count, 1ULL)] = (unsigned long long)(base + b);
    }
}

template <bool LONG = true>
__device__ __forceinline__ void packed_apply(unsigned (&M)[4], const unsigned (&P)[6], int sh)
{
    if (!LONG || sh < 32) {
#pragma unroll
        for (int w = 0; w < 4; w++) M[w] &= __funnelshift_r(P[w], P[w + 1], sh);
    } else {
#pragma unroll
        for (int w = 0; w < 4; w++) M[w] &= __funnelshift_r(P[w + 1], P[w + 2], sh - 32);
    }
}

// ---------------------------------------------------------------------------------------
// Fused candidate filter.  For SPLIT plans every exact piece hit is a candidate the reference
// verifies (checkMatch1 @414190); with short or degenerate pieces that is one candidate per
// 16-100 bases and almost all of them fail.  The scan kernels therefore apply NECESSARY
// conditions of that verification on the packed planes and only pass on what survives:
//   * Myers' bit-vector edit distance of the pattern part left of the anchor against the text
//     read leftwards, and of the right part against the text read rightwards:
//     min_left + min_right <= k.  The reference's NFA accepts a subset of the edit-distance
//     alignments; record ends, fill cuts, line anchors and the scan start only take alignments
//     away.  An X symbol (any non-ACGT byte) matches every position whose class accepts some
//     non-ACGT byte, so the condition stays necessary on headers, N runs and IUPAC letters.
//   * k_scan_split additionally runs the bit-sliced q-gram pre-filter below in front of it.
// Every survivor is decided on the raw bytes by k_verify; dropping is safe for the chain stage as
// well ('^' patterns included: a dropped candidate fails for every scan start).
#define PK_HALO 4                       // words of halo on each side of a 128-word warp tile

template <int NP>
struct PackedVerify {
    unsigned long long TL[NP][4], TR[NP][4];     // match masks per nucleotide code (hi<<1 | lo)
    unsigned long long TLX[NP], TRX[NP];         // positions accepting some non-ACGT byte (X symbols)
    int itmax;                                   // max over pieces of the filter's step count
    int V[NP];
    int m, k, ins, del, subs, enabled;
    const long long *cuts;                       // forced buffer cuts (fill starts not at a '\n'), sorted
    int ncuts;
};

// ---------------------------------------------------------------------------------------
// Exact (k = 0, SIMPLE) scan: one piece, registers only, next tile prefetched while the current
// one is evaluated.  Positions are grouped by the plane they read so that no per-position
// dispatch is needed.
struct ExactPat {
    long long a0, a1;                     // window starts a0 <= w < a1 (already cut to n - L + 1)
    unsigned long long keytag;            // pid << PM_PID_SHIFT
    int L;
    unsigned char npos[6];                // positions reading plane A,C,G,T,X and general classes
    unsigned char shift[6][64];
    unsigned char cls[64];                // general classes: bits A,C,G,T,X (parallel to shift[5])
};
struct ExactArgs {
    const unsigned *hi, *lo, *xx;
    long long nwords, n, tile0, ntiles;
    int npat;
    unsigned long long bad;               // placeholder key of out-of-range hits: sorts after every real key
    ExactPat pat[EX_MAXPAT];
    unsigned long long *keys, *count;
    long long cap;
};

// --- TMA (bulk async copy) + mbarrier helpers -------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_addr(unsigned bar_addr)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
// Release of a staged tile that has just been copied to registers.  The shared-memory loads are only ISSUED at this
// point: measured on B200 (k_scan_multi_hash with the release right behind its loads, lookups slowing the consumers
// down), the producer saw the arrive, refilled the stage by TMA and the refill landed before queued loads of the same
// warp had read the old tile.  `dep` must depend on the destination registers of every load: the arrive then cannot
// issue before their data has come back (dep is stored to a scratch word first: a store ptxas cannot drop).
__device__ __forceinline__ void mbar_arrive_after_loads(unsigned bar_addr, unsigned dep, unsigned sink_addr)
{
    asm volatile("st.volatile.shared.u32 [%2], %1;\n\tmbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr), "r"(dep), "r"(sink_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(unsigned bar_addr, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar_addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Block tile of the exact scan: 8 warp tiles = 1024 words (32768 bases) per plane plus 4 halo
// words, staged by TMA bulk copies into a ring of EX_STAGES shared-memory buffers.
#define EX_WORDS 1024
#define EX_ROW (EX_WORDS + 4)
#define EX_STAGES 3
#define EX_STAGE_BYTES (3 * EX_ROW * 4)

#define EX_HITBUF 128                   // per-warp hit buffer (keys), flushed with one global atomic
#define EX_WPL 8                        // words per lane: a warp tile is 256 words (8192 bases)
#define EX_WARPS 4                      // 128 threads per CTA, block tile = 4 warp tiles = EX_WORDS

template <bool LONG>
__device__ __forceinline__ void exact_apply(unsigned (&M)[EX_WPL], const unsigned (&P)[EX_WPL + 2], int sh)
{
    if (!LONG || sh < 32) {
#pragma unroll
        for (int w = 0; w < EX_WPL; w++) M[w] &= __funnelshift_r(P[w], P[w + 1], sh);
    } else {
#pragma unroll
        for (int w = 0; w < EX_WPL; w++) M[w] &= __funnelshift_r(P[w + 1], P[w + 2], sh - 32);
    }
}

// class planes: one LOP3 per word with the truth table of the class over (hi, lo, x)
template <int LUT>
__device__ __forceinline__ unsigned lop3_const(unsigned a, unsigned b, unsigned c)
{
    unsigned r;
    asm volatile("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));   // volatile: keeps the 32 cases from being speculated
    return r;
}
// truth table over (hi, lo, x) of a class given as bits A,C,G,T,X ; A=000 C=010 G=110 T=100 X=xx1
template <int CLS>
struct SpLut {
    static const int value = ((CLS & 1) ? 0x01 : 0) | ((CLS & 2) ? 0x04 : 0) | ((CLS & 4) ? 0x40 : 0) | ((CLS & 8) ? 0x10 : 0) | ((CLS & 16) ? 0xAA : 0);
};
template <int CLS>
__device__ __forceinline__ void sp_plane(unsigned (&P)[EX_WPL + 2], const unsigned (&H)[EX_WPL + 2], const unsigned (&L)[EX_WPL + 2],
                                         const unsigned (&X)[EX_WPL + 2])
{
#pragma unroll
    for (int w = 0; w < EX_WPL + 2; w++) P[w] = lop3_const<SpLut<CLS>::value>(H[w], L[w], X[w]);
}
__device__ __forceinline__ void sp_plane_dyn(int cls, unsigned (&P)[EX_WPL + 2], const unsigned (&H)[EX_WPL + 2],
                                             const unsigned (&L)[EX_WPL + 2], const unsigned (&X)[EX_WPL + 2])
{
    switch (cls & 31) {
#define SP_CASE(c) case c: sp_plane<c>(P, H, L, X); break;
        SP_CASE(0) SP_CASE(1) SP_CASE(2) SP_CASE(3) SP_CASE(4) SP_CASE(5) SP_CASE(6) SP_CASE(7)
        SP_CASE(8) SP_CASE(9) SP_CASE(10) SP_CASE(11) SP_CASE(12) SP_CASE(13) SP_CASE(14) SP_CASE(15)
        SP_CASE(16) SP_CASE(17) SP_CASE(18) SP_CASE(19) SP_CASE(20) SP_CASE(21) SP_CASE(22) SP_CASE(23)
        SP_CASE(24) SP_CASE(25) SP_CASE(26) SP_CASE(27) SP_CASE(28) SP_CASE(29) SP_CASE(30) SP_CASE(31)
#undef SP_CASE
    }
}

// two positions reading the same plane: one 3-input logic op per word instead of two ANDs
template <bool LONG>
__device__ __forceinline__ void exact_apply2(unsigned (&M)[EX_WPL], const unsigned (&P)[EX_WPL + 2], int s1, int s2)
{
    if (!LONG || (s1 < 32 && s2 < 32)) {
#pragma unroll
        for (int w = 0; w < EX_WPL; w++) M[w] = M[w] & __funnelshift_r(P[w], P[w + 1], s1) & __funnelshift_r(P[w], P[w + 1], s2);
    } else {
        exact_apply<LONG>(M, P, s1);
        exact_apply<LONG>(M, P, s2);
    }
}
template <bool LONG>
__device__ __forceinline__ void exact_apply_group(unsigned (&M)[EX_WPL], const unsigned (&P)[EX_WPL + 2], const unsigned char *sh, int n)
{
    int e = 0;
    for (; e + 1 < n; e += 2) exact_apply2<LONG>(M, P, sh[e], sh[e + 1]);
    if (e < n) exact_apply<LONG>(M, P, sh[e]);
}

template <bool LONG>
__global__ void __launch_bounds__((EX_WARPS + 1) * 32, 5) k_scan_packed_exact(const ExactArgs a)
{
    extern __shared__ __align__(128) unsigned char ex_smem[];
    __shared__ unsigned long long hitbuf_all[EX_WARPS + 1][EX_HITBUF];
    __shared__ unsigned dep_sink[EX_WARPS + 1];
    unsigned long long *hitbuf = hitbuf_all[threadIdx.x >> 5];
    unsigned nbuf = 0;                                      // warp-uniform fill of hitbuf
    auto flush = [&]() {
        if (nbuf == 0) return;
        unsigned long long basei = 0;
        if ((threadIdx.x & 31) == 0) basei = atomicAdd(a.count, (unsigned long long)nbuf);
        basei = __shfl_sync(0xffffffffu, basei, 0);
        for (unsigned e = threadIdx.x & 31; e < nbuf; e += 32)
            if ((long long)(basei + e) < a.cap) a.keys[basei + e] = hitbuf[e];
        __syncwarp();
        nbuf = 0;
    };
    unsigned *stage_base = reinterpret_cast<unsigned *>(ex_smem);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(ex_smem + EX_STAGES * EX_STAGE_BYTES);
    unsigned long long *empty = full + EX_STAGES;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // block tiles handled by this CTA: bt = blockIdx.x, blockIdx.x + gridDim.x, ...
    const long long nbt = (a.ntiles + 7) / 8;
    const long long my = blockIdx.x < nbt ? (nbt - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < EX_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], EX_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](long long it, int s) {                // producer lane only
        const long long bt = blockIdx.x + it * gridDim.x;
        const long long q = (a.tile0 * 128) + bt * EX_WORDS;
        unsigned *dst = stage_base + (size_t)s * (3 * EX_ROW);
        mbar_expect_tx(&full[s], EX_STAGE_BYTES);
        tma_load_1d(dst, a.hi + q, EX_ROW * 4, &full[s]);
        tma_load_1d(dst + EX_ROW, a.lo + q, EX_ROW * 4, &full[s]);
        tma_load_1d(dst + 2 * EX_ROW, a.xx + q, EX_ROW * 4, &full[s]);
    };
    if (wib == EX_WARPS) {
        // producer warp: one lane keeps the ring full, independent of the consumers' progress
        if (lane == 0) {
            int s = 0;
            unsigned ph = 1;                              // parity of the round before the one being filled
            for (long long it = 0; it < my; it++) {
                if (it >= EX_STAGES) mbar_wait(&empty[s], ph);
                issue(it, s);
                if (++s == EX_STAGES) { s = 0; ph ^= 1u; }
            }
        }
        return;
    }
    // ring position kept incrementally: no 64-bit division by EX_STAGES per tile
    int s = 0;
    unsigned ph = 0;
    const unsigned full_base = smem_u32(full), empty_base = smem_u32(empty);
    for (long long it = 0; it < my; it++, s = (s + 1 == EX_STAGES ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
        mbar_wait_addr(full_base + 8u * s, ph);
        const unsigned *sp = stage_base + (size_t)s * (3 * EX_ROW) + wib * (32 * EX_WPL) + EX_WPL * lane;
        unsigned H[EX_WPL + 2], Lw[EX_WPL + 2], X[EX_WPL + 2];
#pragma unroll
        for (int v4 = 0; v4 < EX_WPL / 4; v4++) {
            const uint4 h4 = *reinterpret_cast<const uint4 *>(sp + 4 * v4), l4 = *reinterpret_cast<const uint4 *>(sp + EX_ROW + 4 * v4),
                        x4 = *reinterpret_cast<const uint4 *>(sp + 2 * EX_ROW + 4 * v4);
            H[4 * v4] = h4.x; H[4 * v4 + 1] = h4.y; H[4 * v4 + 2] = h4.z; H[4 * v4 + 3] = h4.w;
            Lw[4 * v4] = l4.x; Lw[4 * v4 + 1] = l4.y; Lw[4 * v4 + 2] = l4.z; Lw[4 * v4 + 3] = l4.w;
            X[4 * v4] = x4.x; X[4 * v4 + 1] = x4.y; X[4 * v4 + 2] = x4.z; X[4 * v4 + 3] = x4.w;
        }
        {
            const uint2 h2 = *reinterpret_cast<const uint2 *>(sp + EX_WPL), l2 = *reinterpret_cast<const uint2 *>(sp + EX_ROW + EX_WPL),
                        x2 = *reinterpret_cast<const uint2 *>(sp + 2 * EX_ROW + EX_WPL);
            H[EX_WPL] = h2.x; H[EX_WPL + 1] = h2.y; Lw[EX_WPL] = l2.x; Lw[EX_WPL + 1] = l2.y; X[EX_WPL] = x2.x; X[EX_WPL + 1] = x2.y;
        }
        __syncwarp();
        {
            // one register of every load instruction above (uint4 / uint2 per plane)
            const unsigned dep = (H[0] ^ H[4] ^ H[EX_WPL]) ^ (Lw[0] ^ Lw[4] ^ Lw[EX_WPL]) ^ (X[0] ^ X[4] ^ X[EX_WPL]);
            if (lane == 0) mbar_arrive_after_loads(empty_base + 8u * s, dep, smem_u32(&dep_sink[wib]));   // this warp's slice is in registers
        }
        const long long bt = blockIdx.x + it * gridDim.x;
        const long long wbase = ((a.tile0 * 128) + bt * EX_WORDS + wib * (32 * EX_WPL)) * 32;   // text position of the warp tile
        const unsigned lrel = (unsigned)(EX_WPL * lane) * 32;
        // every pattern of the request is evaluated on the registers of this tile: the planes are read once
#pragma unroll 1
        for (int pi = 0; pi < a.npat; pi++) {
            const ExactPat &pt = a.pat[pi];
            unsigned M[EX_WPL];
#pragma unroll
            for (int w = 0; w < EX_WPL; w++) M[w] = ~0u;
            unsigned P[EX_WPL + 2];
            if (pt.npos[0]) {
#pragma unroll
                for (int w = 0; w < EX_WPL + 2; w++) P[w] = ~(H[w] | Lw[w] | X[w]);
                exact_apply_group<LONG>(M, P, pt.shift[0], pt.npos[0]);
            }
            if (pt.npos[1]) {
#pragma unroll
                for (int w = 0; w < EX_WPL + 2; w++) P[w] = Lw[w] & ~H[w];
                exact_apply_group<LONG>(M, P, pt.shift[1], pt.npos[1]);
            }
            if (pt.npos[2]) {
#pragma unroll
                for (int w = 0; w < EX_WPL + 2; w++) P[w] = H[w] & Lw[w];
                exact_apply_group<LONG>(M, P, pt.shift[2], pt.npos[2]);
            }
            if (pt.npos[3]) {
#pragma unroll
                for (int w = 0; w < EX_WPL + 2; w++) P[w] = H[w] & ~Lw[w];
                exact_apply_group<LONG>(M, P, pt.shift[3], pt.npos[3]);
            }
            if (pt.npos[4]) {
                exact_apply_group<LONG>(M, X, pt.shift[4], pt.npos[4]);
            }
            for (int e = 0; e < pt.npos[5]; e++) {
                sp_plane_dyn(pt.cls[e], P, H, Lw, X);
                exact_apply<LONG>(M, P, pt.shift[5][e]);
            }
            // ---- hits of this warp tile: buffered in shared memory, no per-hit global atomics ----
            unsigned mine = 0;
#pragma unroll
            for (int w = 0; w < EX_WPL; w++) mine += __popc(M[w]);
            const unsigned have = __ballot_sync(0xffffffffu, mine != 0);
            if (!have) continue;
            // offsets: one ballot when no lane holds more than one hit (the usual case), a shuffle scan otherwise
            unsigned incl, total;
            if (!__any_sync(0xffffffffu, mine > 1)) {
                incl = __popc(have & (0xffffffffu >> (31 - lane)));
                total = __popc(have);
            } else {
                incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                total = __shfl_sync(0xffffffffu, incl, 31);
            }
            // only the first / last tiles of the scanned range need the per-hit range test
            const long long pa0 = pt.a0, pa1 = pt.a1;
            const bool inside = wbase >= pa0 && wbase + 32 * 32 * EX_WPL <= pa1;
            const unsigned long long tag = pt.keytag;
            if (total > EX_HITBUF) {
                // dense tile: straight to global memory
                flush();
                unsigned long long basei = 0;
                if (lane == 0) basei = atomicAdd(a.count, (unsigned long long)total);
                basei = __shfl_sync(0xffffffffu, basei, 0) + (incl - mine);
#pragma unroll
                for (int w = 0; w < EX_WPL; w++) {
                    unsigned c = M[w];
                    while (c) {
                        const int b = __ffs(c) - 1;
                        c &= c - 1;
                        const long long p = wbase + (lrel + w * 32 + b);
                        const bool ok = inside || (p >= pa0 && p < pa1);
                        if ((long long)basei < a.cap) a.keys[basei] = ok ? (tag | ((unsigned long long)p << 4)) : a.bad;
                        if (!ok) atomicAdd(a.count + 1, 1ULL);          // placeholders sort last and are cut off afterwards
                        basei++;
                    }
                }
                continue;
            }
            if (nbuf + total > EX_HITBUF) flush();
            unsigned slot = nbuf + (incl - mine);
#pragma unroll
            for (int w = 0; w < EX_WPL; w++) {
                unsigned c = M[w];
                while (c) {
                    const int b = __ffs(c) - 1;
                    c &= c - 1;
                    const long long p = wbase + (lrel + w * 32 + b);
                    const bool ok = inside || (p >= pa0 && p < pa1);
                    hitbuf[slot++] = ok ? (tag | ((unsigned long long)p << 4)) : a.bad;
                    if (!ok) atomicAdd(a.count + 1, 1ULL);              // placeholders sort last and are cut off afterwards
                }
            }
            nbuf += total;
            __syncwarp();
        }
    }
    flush();
}

// ---------------------------------------------------------------------------------------
// q-gram pre-filter (bit-sliced, 32 pattern starts per instruction).  Take disjoint chunks
// g_1..g_T of the pattern.  An alignment of the whole pattern with at most k errors leaves at
// least T-k chunks untouched, and an untouched chunk sits within +-k positions of where the
// error-free alignment puts it (0 when only substitutions are allowed).  So
//     #{ g : chunk g occurs in its window } >= T - k
// is necessary for checkMatch1 @414190 to accept a candidate, whatever piece triggered it.
// Everything is evaluated in "pattern start" coordinates b = anchor - V[i] - k, so that every
// offset is non-negative and fits the 6 plane words a thread holds (needs m + 2k <= 64):
//     chunk at pattern index j  ->  text bits b + j + [0, 2k]   (b + j + k without indels)
//     piece i                   ->  text bits b + k + V[i] + [0, L)
// Low-selectivity pieces (e.g. ANNNRY: one window in 16) make this worthwhile: the dense
// filter costs a few operations per base and removes >90 % of the Myers evaluations.
#define QF_MAXCH 12
#define QF_MAXLEN 12
struct QChunk {
    unsigned char off;                  // first bit of the window, relative to b
    unsigned char npos;                 // constrained positions of the chunk
    unsigned char t[QF_MAXLEN];         // their offsets inside the chunk
    PackedPos pos[QF_MAXLEN];
};
struct QFilter {
    int nch;                            // 0 = filter off
    int win;                            // window width: 2k+1 with indels, 1 without
    QChunk ch[QF_MAXCH];
};

#define BK_WORDS 1024                   // block tile: 8 warp tiles of 128 words (32768 bases) per plane
#define BK_ROW (BK_WORDS + 2 * PK_HALO)
#define BK_QUEUE 2048                   // per-block, per-piece candidate queue (overflow goes to k_verify unfiltered)

template <int NP, typename W, int ROWS>
__global__ void __launch_bounds__(256) k_scan_packed(const PackedArgs<NP> a, const PackedVerify<NP> v)
{
    __shared__ unsigned sh[3 * BK_ROW];                 // hi | lo | x rows of the block tile, with halos
    __shared__ unsigned queue[NP][BK_QUEUE];            // one queue per piece: a round of candidates shares all parameters
    __shared__ unsigned qcount[NP];
    __shared__ W sT[NP][2][8];                          // match masks by symbol (A,C,T,G,X) for the left / right part of each piece
    const int tid = threadIdx.x;
    if (tid < NP * 2 * 5) {
        const int i = tid / 10, side = (tid / 5) & 1, c = tid % 5;
        sT[i][side][c] = (W)(c < 4 ? (side ? v.TR[i][c] : v.TL[i][c]) : (side ? v.TRX[i] : v.TLX[i]));
    }
    const long long nbt = (a.ntiles + 7) / 8;
    for (long long bt = blockIdx.x; bt < nbt; bt += gridDim.x) {
        const long long qb = a.tile0 * 128 + bt * BK_WORDS;             // first word of the block tile
        const long long q0 = qb + 4 * tid;                              // first of this thread's 4 words
        __syncthreads();
        // ---- stage the block tile (+ halos) of the three planes in shared memory ----
        {
            const uint4 h4 = __ldg(reinterpret_cast<const uint4 *>(a.hi + q0));
            const uint4 l4 = __ldg(reinterpret_cast<const uint4 *>(a.lo + q0));
            const uint4 x4 = __ldg(reinterpret_cast<const uint4 *>(a.xx + q0));
            *reinterpret_cast<uint4 *>(sh + PK_HALO + 4 * tid) = h4;
            *reinterpret_cast<uint4 *>(sh + BK_ROW + PK_HALO + 4 * tid) = l4;
            *reinterpret_cast<uint4 *>(sh + 2 * BK_ROW + PK_HALO + 4 * tid) = x4;
            if (tid < 2 * PK_HALO) {
                const bool left = tid < PK_HALO;
                const long long q = left ? qb - PK_HALO + tid : qb + BK_WORDS + (tid - PK_HALO);
                const int si = left ? tid : PK_HALO + BK_WORDS + (tid - PK_HALO);
                const bool ok = q >= 0 && q < a.nwords;
                sh[si] = ok ? __ldg(a.hi + q) : 0u;
                sh[BK_ROW + si] = ok ? __ldg(a.lo + q) : 0u;
                sh[2 * BK_ROW + si] = ok ? __ldg(a.xx + q) : 0xffffffffu;
            }
            if (tid < NP) qcount[tid] = 0;
        }
        __syncthreads();
        unsigned PA[6], PC[6], PG[6], PT[6], PX[6];
#pragma unroll
        for (int w = 0; w < 6; w++) {
            const unsigned h = sh[PK_HALO + 4 * tid + w], l = sh[BK_ROW + PK_HALO + 4 * tid + w],
                           x = sh[2 * BK_ROW + PK_HALO + 4 * tid + w];
            PX[w] = x;
            PA[w] = ~(h | l | x);
            PC[w] = l & ~h;
            PG[w] = h & l;
            PT[w] = h & ~l;
        }
        // class plane of one pattern position, handed to f as a 6-word array
        auto with_plane = [&](const PackedPos pp, auto f) {
            switch (pp.sel) {
            case 0: f(PA); break;
            case 1: f(PC); break;
            case 2: f(PG); break;
            case 3: f(PT); break;
            case 4: f(PX); break;
            default: {
                const unsigned sA = (pp.cls & 1) ? ~0u : 0u, sC = (pp.cls & 2) ? ~0u : 0u, sG = (pp.cls & 4) ? ~0u : 0u,
                               sT = (pp.cls & 8) ? ~0u : 0u, sX = (pp.cls & 16) ? ~0u : 0u;
                unsigned E[6];
#pragma unroll
                for (int w = 0; w < 6; w++)
                    E[w] = (PA[w] & sA) | (PC[w] & sC) | (PG[w] & sG) | (PT[w] & sT) | (PX[w] & sX);
                f(E);
            }
            }
        };
        // ---- pieces ----
        unsigned M[NP][4];
#pragma unroll
        for (int i = 0; i < NP; i++) {
#pragma unroll
            for (int w = 0; w < 4; w++) M[i][w] = 0xffffffffu;
            if (i < a.npieces) {
                for (int j = 0; j < a.L; j++)
                    with_plane(a.pos[i][j], [&](const unsigned (&P)[6]) { packed_apply<true>(M[i], P, j); });
            } else {
#pragma unroll
                for (int w = 0; w < 4; w++) M[i][w] = 0;
            }
        }
        // The filter ignores record ends and forced buffer cuts: they only take alignments away from the
        // reference's verification, so a necessary condition computed without them stays necessary.
        const bool fused = v.enabled != 0;
        // ---- candidates: piece i fires when any piece of trigsets[i] matched at the same start.
        // They are queued per block and per piece in shared memory, then filtered by all threads.
#pragma unroll
        for (int i = 0; i < NP; i++) {
            if (i >= a.npieces) break;
            unsigned C[4] = {0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < NP; j++)
                if ((a.trigsets[i] >> j) & 1u) {
#pragma unroll
                    for (int w = 0; w < 4; w++) C[w] |= M[j][w];
                }
#pragma unroll
            for (int w = 0; w < 4; w++) {
                unsigned c = C[w];
                while (c) {
                    const int b = __ffs(c) - 1;
                    c &= c - 1;
                    const unsigned rel = (unsigned)((4 * tid + w) * 32 + b);       // window start inside the block tile
                    const long long p = qb * 32 + rel;
                    if (p < a.a0 || p >= a.a1 || p + a.L > a.n) continue;
                    const unsigned slot = fused ? atomicAdd(&qcount[i], 1u) : BK_QUEUE;
                    if (slot < BK_QUEUE) queue[i][slot] = rel;
                    else {                                                          // not filtered: decided by k_verify
                        const unsigned long long idx = atomicAdd(a.count, 1ULL);
                        if ((long long)idx < a.cap) a.keys[idx] = a.keytag | ((unsigned long long)p << 4) | (unsigned)i;
                    }
                }
            }
        }
        __syncthreads();
        // Lock-step filter: Myers' bit-vector edit distance of the pattern part left of the anchor
        // against the text read leftwards, then of the right part against the text read rightwards.
        // The reference's NFA accepts a subset of the edit-distance alignments and spends at most k
        // errors over both sides, so  min_left + min_right <= k  is necessary for its verification to
        // succeed; X symbols count as matching every position that accepts any non-ACGT byte.  All
        // candidates of a round belong to one piece: part lengths, masks and trip counts are uniform.
#pragma unroll
        for (int i = 0; i < NP; i++) {
            if (i >= a.npieces) break;
            const unsigned nq = min(qcount[i], (unsigned)BK_QUEUE);
            const int lb = v.V[i], rl = v.m - lb;
            const W topL = lb > 0 ? (W)((W)1 << (lb - 1)) : (W)0, topR = rl > 0 ? (W)((W)1 << (rl - 1)) : (W)0;
            const int nL = lb > 0 ? lb + v.k : 0, nR = rl > 0 ? rl + v.k : 0;
            const W *tl = sT[i][0], *tr = sT[i][1];
            for (unsigned e = tid; e < nq; e += 256) {
                const int rel0 = (int)queue[i][e] + PK_HALO * 32;                   // bit index in the staged rows
                int minL = 0, minR = 0;
                if (nL > 0) {
                    // symbols rel0-1, rel0-2, ... : 32 at a time, bit-reversed so that bit 0 is the next one
                    W Pv = (W)~(W)0, Mv = 0;
                    int score = lb, best = lb;
                    unsigned hw = 0, lw = 0, xw = 0;
#pragma unroll 1
                    for (int it = 0; it < nL; it++) {
                        if ((it & 31) == 0) {
                            const int hi_bit = rel0 - 1 - it;                       // window = bits [hi_bit-31, hi_bit]
                            const int lo_bit = hi_bit - 31;
                            const int wi = lo_bit >> 5, bi = lo_bit & 31;
                            hw = __brev(__funnelshift_r(sh[wi], sh[wi + 1], bi));
                            lw = __brev(__funnelshift_r(sh[BK_ROW + wi], sh[BK_ROW + wi + 1], bi));
                            xw = __brev(__funnelshift_r(sh[2 * BK_ROW + wi], sh[2 * BK_ROW + wi + 1], bi));
                        }
                        const W Eq = tl[((xw & 1u) << 2) | ((hw & 1u) << 1) | (lw & 1u)];
                        hw >>= 1; lw >>= 1; xw >>= 1;
                        const W Xv = Eq | Mv;
                        const W Xh = (W)((((Eq & Pv) + Pv) ^ Pv) | Eq);
                        W Ph = (W)(Mv | ~(Xh | Pv));
                        W Mh = (W)(Pv & Xh);
                        score += (Ph & topL) ? 1 : 0;
                        score -= (Mh & topL) ? 1 : 0;
                        Ph = (W)((Ph << 1) | 1);
                        Mh = (W)(Mh << 1);
                        Pv = (W)(Mh | ~(Xv | Ph));
                        Mv = (W)(Ph & Xv);
                        best = min(best, score);
                    }
                    minL = best;
                }
                if (nR > 0 && minL <= v.k) {                       // the left side alone can already rule the candidate out
                    W Pv = (W)~(W)0, Mv = 0;
                    int score = rl, best = rl;
                    unsigned hw = 0, lw = 0, xw = 0;
#pragma unroll 1
                    for (int it = 0; it < nR; it++) {
                        if ((it & 31) == 0) {
                            const int lo_bit = rel0 + it;
                            const int wi = lo_bit >> 5, bi = lo_bit & 31;
                            hw = __funnelshift_r(sh[wi], sh[wi + 1], bi);
                            lw = __funnelshift_r(sh[BK_ROW + wi], sh[BK_ROW + wi + 1], bi);
                            xw = __funnelshift_r(sh[2 * BK_ROW + wi], sh[2 * BK_ROW + wi + 1], bi);
                        }
                        const W Eq = tr[((xw & 1u) << 2) | ((hw & 1u) << 1) | (lw & 1u)];
                        hw >>= 1; lw >>= 1; xw >>= 1;
                        const W Xv = Eq | Mv;
                        const W Xh = (W)((((Eq & Pv) + Pv) ^ Pv) | Eq);
                        W Ph = (W)(Mv | ~(Xh | Pv));
                        W Mh = (W)(Pv & Xh);
                        score += (Ph & topR) ? 1 : 0;
                        score -= (Mh & topR) ? 1 : 0;
                        Ph = (W)((Ph << 1) | 1);
                        Mh = (W)(Mh << 1);
                        Pv = (W)(Mh | ~(Xv | Ph));
                        Mv = (W)(Ph & Xv);
                        best = min(best, score);
                    }
                    minR = best;
                }
                if (minL + minR <= v.k) {
                    const long long p = qb * 32 + (rel0 - PK_HALO * 32);
                    const unsigned long long idx = atomicAdd(a.count, 1ULL);
                    if ((long long)idx < a.cap) a.keys[idx] = a.keytag | ((unsigned long long)p << 4) | (unsigned)i;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// SPLIT scan, register-resident (m + 2k <= 64).  Same TMA ring and warp tiles as the exact
// scan: a lane holds 8+2 words of the three planes, evaluates the q-gram pre-filter and the
// k+1 pieces on them in pattern-start coordinates, and queues the surviving anchors per warp
// and per piece.  Whenever a queue holds 32 anchors the warp runs one lock-step round of the
// Myers filter on them (all lanes busy, one piece per round so every lane has the same part
// lengths and tables), reading the few plane words it needs from global memory (L2 hits: the
// tile has just been streamed).  There is no block-wide synchronisation inside the tile loop.
#define SP_STAGES 3
#define SP_CTAS 4                        // CTAs per SM the register budget is cut for (5 x 96 registers measured slower)
#define SP_QUEUE 128                     // per-warp anchor queue (ring of keys)

// One lock-step round of the Myers filter on up to 64 queued anchors, two per lane.  A queue entry is
// the key (anchor << 4 | piece); the 64 symbols starting at the anchor's pattern start b are fetched
// from the planes in global memory (L2 hits: the tile has just been streamed), all loads issued before
// the loop.  Per anchor both sides run interleaved (the left part P[0..V) against the symbols read
// leftwards from the anchor, the right part P[V..m) against the symbols read rightwards), so a lane
// carries four independent dependency chains.  Each side yields the minimum edit distance over all
// prefixes read, and  min_left + min_right <= k  is necessary for checkMatch1 @414190 to succeed.
struct SpParams {                        // per piece, in shared memory
    int base, lb, rl;                    // base = k + V[i] = index of the anchor inside the fetched window
};

__device__ __forceinline__ unsigned long long sp_window(const unsigned *__restrict__ plane, long long bit)
{
    const long long wi = bit >> 5;
    const int bi = (int)(bit & 31);
    const unsigned w0 = __ldg(plane + wi), w1 = __ldg(plane + wi + 1), w2 = __ldg(plane + wi + 2);
    return ((unsigned long long)__funnelshift_r(w1, w2, bi) << 32) | __funnelshift_r(w0, w1, bi);
}

template <typename W>
__device__ __noinline__ void sp_round(const unsigned long long *__restrict__ qkey, unsigned head, unsigned cnt,
                                      const unsigned *__restrict__ hi, const unsigned *__restrict__ lo, const unsigned *__restrict__ xx,
                                      const W *__restrict__ tabs, const SpParams *__restrict__ par, int k,
                                      unsigned long long *__restrict__ keys, unsigned long long *__restrict__ count, long long cap,
                                      unsigned long long keytag)
{
    const int lane = threadIdx.x & 31;
    unsigned long long key[2];
    bool act[2], keep[2];
    unsigned sL[2][3], sR[2][3];                         // next 32 symbols of the (hi, lo, x) streams, left / right
    long long b_[2];                                     // pattern start (symbols beyond 32 are fetched again)
    int base_[2];
    const W *tl[2];
    W topL[2], topR[2], PvL[2], MvL[2], PvR[2], MvR[2];
    int nL[2], nR[2], scL[2], scR[2], bestL[2], bestR[2];
    int nit = 0;
#pragma unroll
    for (int c = 0; c < 2; c++) {
        act[c] = (unsigned)(lane + 32 * c) < cnt;
        key[c] = qkey[(head + lane + 32 * c) & (SP_QUEUE - 1)];
        const int i = act[c] ? (int)(key[c] & 15) : 0;
        const SpParams pp = par[i];
        const long long b = act[c] ? (long long)(key[c] >> 4) - pp.base : 0;
        const unsigned long long h = sp_window(hi, b), l = sp_window(lo, b), x = sp_window(xx, b);
        // left stream: bit t = window bit base-1-t ; right stream: bit t = window bit base+t   (1 <= base <= 63)
        const int shl = 64 - pp.base;
        const unsigned long long hL = __brevll(h) >> shl, lL = __brevll(l) >> shl, xL = __brevll(x) >> shl;
        const unsigned long long hR = h >> pp.base, lR = l >> pp.base, xR = x >> pp.base;
        sL[c][0] = (unsigned)hL; sL[c][1] = (unsigned)lL; sL[c][2] = (unsigned)xL;
        sR[c][0] = (unsigned)hR; sR[c][1] = (unsigned)lR; sR[c][2] = (unsigned)xR;
        b_[c] = b;
        base_[c] = pp.base;
        tl[c] = tabs + i * 16;
        nL[c] = pp.lb > 0 ? pp.base : 0;
        nR[c] = pp.rl + k;
        topL[c] = pp.lb > 0 ? (W)((W)1 << (pp.lb - 1)) : (W)0;
        topR[c] = (W)((W)1 << (pp.rl - 1));
        PvL[c] = (W)~(W)0; MvL[c] = 0; PvR[c] = (W)~(W)0; MvR[c] = 0;
        scL[c] = bestL[c] = pp.lb;
        scR[c] = bestR[c] = pp.rl;
        if (act[c]) nit = max(nit, max(nL[c], nR[c]));
    }
#pragma unroll 1
    for (int t = 0; t < nit; t++) {
        if (t == 32) {
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const unsigned long long wq = sp_window(q == 0 ? hi : q == 1 ? lo : xx, b_[c]);
                    sL[c][q] = (unsigned)((__brevll(wq) >> (64 - base_[c])) >> 32);
                    sR[c][q] = (unsigned)((wq >> base_[c]) >> 32);
                }
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const W EqL = tl[c][((sL[c][2] & 1u) << 2) | ((sL[c][0] & 1u) << 1) | (sL[c][1] & 1u)];
            const W EqR = tl[c][8 + (((sR[c][2] & 1u) << 2) | ((sR[c][0] & 1u) << 1) | (sR[c][1] & 1u))];
#pragma unroll
            for (int q = 0; q < 3; q++) { sL[c][q] >>= 1; sR[c][q] >>= 1; }
            {
                const W Xv = EqL | MvL[c];
                const W Xh = (W)((((EqL & PvL[c]) + PvL[c]) ^ PvL[c]) | EqL);
                W Ph = (W)(MvL[c] | ~(Xh | PvL[c]));
                W Mh = (W)(PvL[c] & Xh);
                scL[c] += (Ph & topL[c]) ? 1 : 0;
                scL[c] -= (Mh & topL[c]) ? 1 : 0;
                Ph = (W)((Ph << 1) | 1);
                Mh = (W)(Mh << 1);
                PvL[c] = (W)(Mh | ~(Xv | Ph));
                MvL[c] = (W)(Ph & Xv);
                if (t < nL[c]) bestL[c] = min(bestL[c], scL[c]);
            }
            {
                const W Xv = EqR | MvR[c];
                const W Xh = (W)((((EqR & PvR[c]) + PvR[c]) ^ PvR[c]) | EqR);
                W Ph = (W)(MvR[c] | ~(Xh | PvR[c]));
                W Mh = (W)(PvR[c] & Xh);
                scR[c] += (Ph & topR[c]) ? 1 : 0;
                scR[c] -= (Mh & topR[c]) ? 1 : 0;
                Ph = (W)((Ph << 1) | 1);
                Mh = (W)(Mh << 1);
                PvR[c] = (W)(Mh | ~(Xv | Ph));
                MvR[c] = (W)(Ph & Xv);
                if (t < nR[c]) bestR[c] = min(bestR[c], scR[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 2; c++) keep[c] = act[c] && bestL[c] + bestR[c] <= k;
    const unsigned bal0 = __ballot_sync(0xffffffffu, keep[0]), bal1 = __ballot_sync(0xffffffffu, keep[1]);
    if (bal0 | bal1) {
        unsigned long long basei = 0;
        if (lane == 0) basei = atomicAdd(count, (unsigned long long)(__popc(bal0) + __popc(bal1)));
        basei = __shfl_sync(0xffffffffu, basei, 0);
        const unsigned below = (1u << lane) - 1u;
        const unsigned long long i0 = basei + __popc(bal0 & below), i1 = basei + __popc(bal0) + __popc(bal1 & below);
        if (keep[0] && (long long)i0 < cap) keys[i0] = key[0] | keytag;
        if (keep[1] && (long long)i1 < cap) keys[i1] = key[1] | keytag;
    }
    __syncwarp();
}

template <int NP, typename W, int ROWS>
__global__ void __launch_bounds__(EX_WARPS * 32, SP_CTAS) k_scan_split(const PackedArgs<NP> a, const PackedVerify<NP> v, const QFilter qf)
{
    extern __shared__ __align__(128) unsigned char ex_smem[];
    __shared__ unsigned long long q_key[EX_WARPS][SP_QUEUE];
    __shared__ unsigned dep_sink[EX_WARPS];
    __shared__ unsigned s_anchor[EX_WARPS][EX_WPL * 32];   // anchor bits of the piece being extracted, [word][lane]
    __shared__ W sT[NP * 16];                           // match masks by symbol (A,C,T,G,X): [piece][left | right][8]
    __shared__ SpParams spar[NP];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    if (tid < NP * 2 * 5) {
        const int i = tid / 10, side = (tid / 5) & 1, c = tid % 5;
        sT[i * 16 + side * 8 + c] = (W)(c < 4 ? (side ? v.TR[i][c] : v.TL[i][c]) : (side ? v.TRX[i] : v.TLX[i]));
    }
    if (tid < NP) { spar[tid].base = v.k + v.V[tid]; spar[tid].lb = v.V[tid]; spar[tid].rl = v.m - v.V[tid]; }
    unsigned *stage_base = reinterpret_cast<unsigned *>(ex_smem);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(ex_smem + SP_STAGES * EX_STAGE_BYTES);
    unsigned long long *empty = full + SP_STAGES;
    const long long nbt = (a.ntiles + 7) / 8;
    const long long my = blockIdx.x < nbt ? (nbt - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
        for (int s = 0; s < SP_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], EX_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // No producer warp here (it would pin a fifth of the register file): thread 0 refills the ring, always
    // SP_STAGES-1 tiles ahead, waiting only for the stage every warp has already copied to registers.
    auto issue = [&](long long it) {
        const int s = (int)(it % SP_STAGES);
        if (it >= SP_STAGES) mbar_wait(&empty[s], (unsigned)(((it / SP_STAGES) - 1) & 1));
        const long long q = (a.tile0 * 128) + (blockIdx.x + it * gridDim.x) * EX_WORDS;
        unsigned *dst = stage_base + (size_t)s * (3 * EX_ROW);
        mbar_expect_tx(&full[s], EX_STAGE_BYTES);
        tma_load_1d(dst, a.hi + q, EX_ROW * 4, &full[s]);
        tma_load_1d(dst + EX_ROW, a.lo + q, EX_ROW * 4, &full[s]);
        tma_load_1d(dst + 2 * EX_ROW, a.xx + q, EX_ROW * 4, &full[s]);
    };
    if (tid == 0)
        for (long long it = 0; it < SP_STAGES - 1 && it < my; it++) issue(it);
    if (tid == 32 && blockIdx.x == 0 && a.tile0 == 0) {
        // anchors closer than k + V[i] to the start of the text have no pattern start b >= 0:
        // they go to k_verify unfiltered (it re-checks the trigger on the raw bytes)
        for (int i = 0; i < a.npieces; i++)
            for (long long p = 0; p < v.k + v.V[i]; p++) {
                if (p < a.a0 || p >= a.a1 || p + a.L > a.n) continue;
                const unsigned long long idx = atomicAdd(a.count, 1ULL);
                if ((long long)idx < a.cap) a.keys[idx] = a.keytag | ((unsigned long long)p << 4) | (unsigned)i;
            }
    }
    unsigned long long *qk = q_key[wib];
    unsigned *sM = s_anchor[wib];
    unsigned qhead = 0, qcnt = 0;                       // warp-uniform ring state
    for (long long it = 0; it < my; it++) {
        const int s = (int)(it % SP_STAGES);
        const unsigned ph = (unsigned)((it / SP_STAGES) & 1);
        if (tid == 0 && it + SP_STAGES - 1 < my) issue(it + SP_STAGES - 1);
        __syncwarp();
        mbar_wait(&full[s], ph);
        const unsigned *sp = stage_base + (size_t)s * (3 * EX_ROW) + wib * (32 * EX_WPL) + EX_WPL * lane;
        unsigned H[EX_WPL + 2], Lw[EX_WPL + 2], X[EX_WPL + 2];
#pragma unroll
        for (int v4 = 0; v4 < EX_WPL / 4; v4++) {
            const uint4 h4 = *reinterpret_cast<const uint4 *>(sp + 4 * v4), l4 = *reinterpret_cast<const uint4 *>(sp + EX_ROW + 4 * v4),
                        x4 = *reinterpret_cast<const uint4 *>(sp + 2 * EX_ROW + 4 * v4);
            H[4 * v4] = h4.x; H[4 * v4 + 1] = h4.y; H[4 * v4 + 2] = h4.z; H[4 * v4 + 3] = h4.w;
            Lw[4 * v4] = l4.x; Lw[4 * v4 + 1] = l4.y; Lw[4 * v4 + 2] = l4.z; Lw[4 * v4 + 3] = l4.w;
            X[4 * v4] = x4.x; X[4 * v4 + 1] = x4.y; X[4 * v4 + 2] = x4.z; X[4 * v4 + 3] = x4.w;
        }
        {
            const uint2 h2 = *reinterpret_cast<const uint2 *>(sp + EX_WPL), l2 = *reinterpret_cast<const uint2 *>(sp + EX_ROW + EX_WPL),
                        x2 = *reinterpret_cast<const uint2 *>(sp + 2 * EX_ROW + EX_WPL);
            H[EX_WPL] = h2.x; H[EX_WPL + 1] = h2.y; Lw[EX_WPL] = l2.x; Lw[EX_WPL + 1] = l2.y; X[EX_WPL] = x2.x; X[EX_WPL + 1] = x2.y;
        }
        __syncwarp();
        {
            const unsigned dep = (H[0] ^ H[4] ^ H[EX_WPL]) ^ (Lw[0] ^ Lw[4] ^ Lw[EX_WPL]) ^ (X[0] ^ X[4] ^ X[EX_WPL]);
            if (lane == 0) mbar_arrive_after_loads(smem_u32(&empty[s]), dep, smem_u32(&dep_sink[wib]));   // this warp's slice is in registers
        }
        const long long bt = blockIdx.x + it * gridDim.x;
        unsigned P[EX_WPL + 2];
        // ---- q-gram pre-filter: pass[w] bit b <=> at most k chunks are missing for pattern start b ----
        unsigned pass[EX_WPL];
#pragma unroll
        for (int w = 0; w < EX_WPL; w++) pass[w] = 0xffffffffu;
        if (qf.nch > 0) {
            unsigned miss[ROWS][EX_WPL];                  // miss[r] : more than r chunks missing
#pragma unroll
            for (int r = 0; r < ROWS; r++)
#pragma unroll
                for (int w = 0; w < EX_WPL; w++) miss[r][w] = 0;
            for (int g = 0; g < qf.nch; g++) {
                unsigned G[EX_WPL + 2];
#pragma unroll
                for (int w = 0; w < EX_WPL + 2; w++) G[w] = 0xffffffffu;
                const int npos = qf.ch[g].npos;
                for (int c = 0; c < npos; c++) {
                    const int t = qf.ch[g].t[c];
                    sp_plane_dyn(qf.ch[g].pos[c].cls, P, H, Lw, X);
#pragma unroll
                    for (int w = 0; w < EX_WPL + 1; w++) G[w] &= __funnelshift_r(P[w], P[w + 1], t);
                    G[EX_WPL + 1] &= P[EX_WPL + 1] >> t;
                }
                // dilate: G[x] |= G[x+1] | ... | G[x+win-1]   (doubling)
                int cw = 1;
                while (cw < qf.win) {
                    const int sft = min(cw, qf.win - cw);
#pragma unroll
                    for (int w = 0; w < EX_WPL + 1; w++) G[w] |= __funnelshift_r(G[w], G[w + 1], sft);
                    G[EX_WPL + 1] |= G[EX_WPL + 1] >> sft;
                    cw += sft;
                }
                unsigned x[EX_WPL];
#pragma unroll
                for (int w = 0; w < EX_WPL; w++) x[w] = 0xffffffffu;
                exact_apply<true>(x, G, qf.ch[g].off);
#pragma unroll
                for (int w = 0; w < EX_WPL; w++) {
                    const unsigned ms = ~x[w];
#pragma unroll
                    for (int r = ROWS - 1; r > 0; r--) miss[r][w] |= miss[r - 1][w] & ms;
                    miss[0][w] |= ms;
                }
            }
#pragma unroll
            for (int w = 0; w < EX_WPL; w++) pass[w] = ~miss[ROWS - 1][w];
        }
        const long long wbase = ((a.tile0 * 128) + bt * EX_WORDS + wib * (32 * EX_WPL)) * 32;   // first pattern start of the warp tile
        const unsigned lrel = (unsigned)(EX_WPL * lane) * 32;
        // every anchor of this warp tile lies in [a0, a1) and far enough from the end of the text (true for all but the first / last tiles)
        const bool inside = wbase + v.k >= a.a0 && wbase + v.m + v.k + 32 * 32 * EX_WPL <= a.a1 && wbase + v.m + v.k + 32 * 32 * EX_WPL + a.L <= a.n;
        // ---- pieces, in pattern-start coordinates: piece i sits k + V[i] bits further ----
#pragma unroll 1
        for (int i = 0; i < a.npieces; i++) {
            unsigned M[EX_WPL];
#pragma unroll
            for (int w = 0; w < EX_WPL; w++) M[w] = pass[w];
            const int base = v.k + v.V[i];
            for (int j = 0; j < a.L; j++) {
                const int cls = a.pos[i][j].cls;
                if (cls == 31) continue;                   // accepts every byte
                sp_plane_dyn(cls, P, H, Lw, X);
                exact_apply<true>(M, P, base + j);
            }
            const long long p0 = wbase + base + lrel;      // anchor of bit 0 of this lane's first word
            unsigned mine = 0;
#pragma unroll
            for (int w = 0; w < EX_WPL; w++) mine += __popc(M[w]);
            if (!__any_sync(0xffffffffu, mine != 0)) continue;
            // the anchors leave the registers here: extraction walks them in shared memory, [word][lane]
#pragma unroll
            for (int w = 0; w < EX_WPL; w++) sM[w * 32 + lane] = M[w];
            if (!inside) {
                // first / last tiles of the scanned range: drop the anchors outside [a0, a1) or too close to the end
                mine = 0;
#pragma unroll 1
                for (int w = 0; w < EX_WPL; w++) {
                    unsigned c = sM[w * 32 + lane], keepm = c;
                    while (c) {
                        const int b = __ffs(c) - 1;
                        c &= c - 1;
                        const long long p = p0 + w * 32 + b;
                        if (!(p >= a.a0 && p < a.a1 && p + a.L <= a.n)) keepm &= ~(1u << b);
                    }
                    sM[w * 32 + lane] = keepm;
                    mine += __popc(keepm);
                }
                if (!__any_sync(0xffffffffu, mine != 0)) continue;
            }
            unsigned incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
            if (qcnt + total <= SP_QUEUE) {
                // sparse: every lane appends its anchors at its own offset, then full rounds are drained
                unsigned slot = qhead + qcnt + (incl - mine);
                if (mine) {
#pragma unroll 1
                    for (int w = 0; w < EX_WPL; w++) {
                        unsigned c = sM[w * 32 + lane];
                        while (c) {
                            const int b = __ffs(c) - 1;
                            c &= c - 1;
                            qk[slot++ & (SP_QUEUE - 1)] = ((unsigned long long)(p0 + w * 32 + b) << 4) | (unsigned)i;
                        }
                    }
                }
                qcnt += total;
                __syncwarp();
                while (qcnt >= 64) {
                    sp_round<W>(qk, qhead, 64, a.hi, a.lo, a.xx, sT, spar, v.k, a.keys, a.count, a.cap, a.keytag);
                    qhead = (qhead + 64) & (SP_QUEUE - 1);
                    qcnt -= 64;
                }
            } else {
                // dense: one anchor per lane and step, a round as soon as 64 are queued
#pragma unroll 1
                for (int w = 0; w < EX_WPL; w++) {
                    unsigned c = sM[w * 32 + lane];
                    while (__any_sync(0xffffffffu, c != 0)) {
                        const bool have = c != 0;
                        const unsigned bal = __ballot_sync(0xffffffffu, have);
                        if (have) {
                            const int b = __ffs(c) - 1;
                            c &= c - 1;
                            qk[(qhead + qcnt + __popc(bal & ((1u << lane) - 1u))) & (SP_QUEUE - 1)] = ((unsigned long long)(p0 + w * 32 + b) << 4) | (unsigned)i;
                        }
                        qcnt += __popc(bal);
                        __syncwarp();
                        if (qcnt >= 64) {
                            sp_round<W>(qk, qhead, 64, a.hi, a.lo, a.xx, sT, spar, v.k, a.keys, a.count, a.cap, a.keytag);
                            qhead = (qhead + 64) & (SP_QUEUE - 1);
                            qcnt -= 64;
                        }
                    }
                }
            }
        }
    }
    if (qcnt) sp_round<W>(qk, qhead, qcnt, a.hi, a.lo, a.xx, sT, spar, v.k, a.keys, a.count, a.cap, a.keytag);
}

// ---------------------------------------------------------------------------------------
// SPLIT scan, second generation (k_scan_apx): one pass over the planes for every pattern of a request (a motif and
// its reverse complement), m + 2k <= 64, k <= 3.  Same TMA ring, warp tiles and pattern-start coordinates
// b = anchor - V[i] - k as k_scan_split, but
//   * the k+1 exact pieces are not evaluated position by position: the host picks q-gram chunks that refine the
//     pieces, and a piece plane is the AND of the (undilated) planes of its chunks -- every pattern position is
//     turned into a class plane once per tile;
//   * the per-anchor filter is not a symbol-by-symbol Myers loop but Landau-Vishkin over the 2k+1 diagonals of the
//     64-symbol window: D_d = pattern positions that match on diagonal d (ten logic operations per diagonal from the
//     one-hot symbol planes and per-symbol position masks), then the furthest-reaching pattern index per (errors,
//     diagonal) with a count-trailing-ones slide.  Straight-line code, no loop over text symbols, one check per
//     pattern start b whatever the number of pieces that fired there.  "P aligns to text[b + d0 ...] with at most k
//     errors for some start diagonal d0 in [0, 2k]" is necessary for checkMatch1 @414190 to accept any anchor whose
//     pattern start is b (its alignment stays within k diagonals of the nominal one, d = k, which it crosses at the
//     exact piece); record ends, fill cuts, line anchors and the scan start only remove alignments.  Which pieces
//     matched exactly comes for free from the nominal diagonal D_k.
// Survivors are decided on the raw bytes by k_verify as before.
struct ApxArgs {
    const unsigned *hi, *lo, *xx;
    long long nwords, n, tile0, ntiles;
    int npat;
    ApxPat pat[EX_MAXPAT];
    unsigned long long *keys, *count;
    long long cap;
};
struct ApxSparse {                       // what the per-anchor check needs, copied to shared memory
    long long a0, a1;
    unsigned long long keytag;
    unsigned long long posmask[5];
    int m, L, npieces, indel;
    int V[4];
};

__device__ __forceinline__ int ax_tz(unsigned x) { return __clz(__brev(x)); }
__device__ __forceinline__ int ax_tz(unsigned long long x) { return __clzll(__brevll(x)); }

template <int K, typename DW>
__device__ __noinline__ void apx_round(const unsigned long long *__restrict__ qk, unsigned head, unsigned cnt,
                                       const unsigned *__restrict__ hi, const unsigned *__restrict__ lo, const unsigned *__restrict__ xx,
                                       const ApxSparse *__restrict__ sp, long long n,
                                       unsigned long long *__restrict__ keys, unsigned long long *__restrict__ count, long long cap)
{
    constexpr int ND = 2 * K + 1;
    const int lane = threadIdx.x & 31;
    const bool act = (unsigned)lane < cnt;
    const unsigned long long ent = qk[(head + lane) & (AX_QUEUE - 1)];
    const ApxSparse &pt = sp[act ? (int)(ent & 1) : 0];
    const long long b = act ? (long long)(ent >> 1) : 0;
    DW A, C, T, G, X;
    if (sizeof(DW) == 8) {
        const unsigned long long h = sp_window(hi, b), l = sp_window(lo, b), x = sp_window(xx, b);
        A = (DW)~(h | l | x); C = (DW)(l & ~h); T = (DW)(h & ~l); G = (DW)(h & l); X = (DW)x;
    } else {
        const long long wi = b >> 5;
        const int bi = (int)(b & 31);
        const unsigned h = __funnelshift_r(__ldg(hi + wi), __ldg(hi + wi + 1), bi), l = __funnelshift_r(__ldg(lo + wi), __ldg(lo + wi + 1), bi),
                       x = __funnelshift_r(__ldg(xx + wi), __ldg(xx + wi + 1), bi);
        A = (DW)~(h | l | x); C = (DW)(l & ~h); T = (DW)(h & ~l); G = (DW)(h & l); X = (DW)x;
    }
    const DW mA = (DW)pt.posmask[0], mC = (DW)pt.posmask[1], mT = (DW)pt.posmask[2], mG = (DW)pt.posmask[3], mX = (DW)pt.posmask[4];
    DW D[ND];
#pragma unroll
    for (int d = 0; d < ND; d++) D[d] = (DW)((mA & (A >> d)) | (mC & (C >> d)) | (mT & (T >> d)) | (mG & (G >> d)) | (mX & (X >> d)));
    const int m = pt.m;
    const bool indel = pt.indel != 0;
    int f[ND];
    bool pass = false;
#pragma unroll
    for (int d = 0; d < ND; d++) {
        f[d] = (indel || d == K) ? ax_tz((DW)~D[d]) : -1000;
        pass |= f[d] >= m;
    }
#pragma unroll
    for (int e = 1; e <= K; e++) {
        int g[ND];
#pragma unroll
        for (int d = 0; d < ND; d++) {
            int v = f[d] + 1;
            if (indel) {
                if (d > 0) v = max(v, f[d - 1]);
                if (d < ND - 1) v = max(v, f[d + 1] + 1);
            }
            if (v >= m) { v = m; pass = true; }
            else if (v >= 0) {
                v += ax_tz((DW)~(DW)(D[d] >> v));
                pass |= v >= m;
            }
            g[d] = v;
        }
#pragma unroll
        for (int d = 0; d < ND; d++) f[d] = g[d];
    }
    // pieces that match exactly at their nominal place (diagonal k), restricted to the anchors asked for
    unsigned pieces = 0;
    if (act && pass) {
        const DW Lm = (DW)(((DW)1 << pt.L) - 1);       // L <= 31: a SPLIT plan has at least two pieces
        for (int i = 0; i < pt.npieces; i++) {
            const long long p = b + K + pt.V[i];
            if ((DW)((D[K] >> pt.V[i]) & Lm) == Lm && p >= pt.a0 && p < pt.a1 && p + pt.L <= n) pieces |= 1u << i;
        }
    }
    const unsigned np = __popc(pieces);
    if (__any_sync(0xffffffffu, np != 0)) {
        unsigned incl = np;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned long long basei = 0;
        if (lane == 0) basei = atomicAdd(count, (unsigned long long)total);
        basei = __shfl_sync(0xffffffffu, basei, 0) + (incl - np);
        unsigned c = pieces;
        while (c) {
            const int i = __ffs(c) - 1;
            c &= c - 1;
            if ((long long)basei < cap) keys[basei] = pt.keytag | ((unsigned long long)(b + K + pt.V[i]) << 4) | (unsigned)i;
            basei++;
        }
    }
    __syncwarp();
}

template <int ROWS, typename DW>
__global__ void __launch_bounds__(EX_WARPS * 32, SP_CTAS) k_scan_apx(const ApxArgs a)
{
    constexpr int K = ROWS - 1;
    extern __shared__ __align__(128) unsigned char ex_smem[];
    __shared__ unsigned long long q_key[EX_WARPS][AX_QUEUE];
    __shared__ unsigned dep_sink[EX_WARPS];
    __shared__ unsigned s_cand[EX_WARPS][EX_WPL * 32];     // dense tiles only: candidate bits, [word][lane]
    __shared__ ApxSparse s_sp[EX_MAXPAT];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    if (tid < a.npat) {
        const ApxPat &pt = a.pat[tid];
        ApxSparse &o = s_sp[tid];
        o.a0 = pt.a0; o.a1 = pt.a1; o.keytag = pt.keytag;
        for (int q = 0; q < 5; q++) o.posmask[q] = pt.posmask[q];
        o.m = pt.m; o.L = pt.L; o.npieces = pt.npieces; o.indel = pt.indel;
        for (int q = 0; q < 4; q++) o.V[q] = pt.V[q];
    }
    unsigned *stage_base = reinterpret_cast<unsigned *>(ex_smem);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(ex_smem + SP_STAGES * EX_STAGE_BYTES);
    unsigned long long *empty = full + SP_STAGES;
    const long long nbt = (a.ntiles + 7) / 8;
    const long long my = blockIdx.x < nbt ? (nbt - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
        for (int s = 0; s < SP_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], EX_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](long long it) {
        const int s = (int)(it % SP_STAGES);
        if (it >= SP_STAGES) mbar_wait(&empty[s], (unsigned)(((it / SP_STAGES) - 1) & 1));
        const long long q = (a.tile0 * 128) + (blockIdx.x + it * gridDim.x) * EX_WORDS;
        unsigned *dst = stage_base + (size_t)s * (3 * EX_ROW);
        mbar_expect_tx(&full[s], EX_STAGE_BYTES);
        tma_load_1d(dst, a.hi + q, EX_ROW * 4, &full[s]);
        tma_load_1d(dst + EX_ROW, a.lo + q, EX_ROW * 4, &full[s]);
        tma_load_1d(dst + 2 * EX_ROW, a.xx + q, EX_ROW * 4, &full[s]);
    };
    if (tid == 0)
        for (long long it = 0; it < SP_STAGES - 1 && it < my; it++) issue(it);
    if (tid == 32 && blockIdx.x == 0 && a.tile0 == 0) {
        // anchors closer than k + V[i] to the start of the text have no pattern start b >= 0:
        // they go to k_verify unfiltered (it re-checks the trigger on the raw bytes)
        for (int pi = 0; pi < a.npat; pi++) {
            const ApxPat &pt = a.pat[pi];
            for (int i = 0; i < pt.npieces; i++)
                for (long long p = 0; p < pt.k + pt.V[i]; p++) {
                    if (p < pt.a0 || p >= pt.a1 || p + pt.L > a.n) continue;
                    const unsigned long long idx = atomicAdd(a.count, 1ULL);
                    if ((long long)idx < a.cap) a.keys[idx] = pt.keytag | ((unsigned long long)p << 4) | (unsigned)i;
                }
        }
    }
    unsigned long long *qk = q_key[wib];
    unsigned *sC = s_cand[wib];
    unsigned qhead = 0, qcnt = 0;                       // warp-uniform ring state
    for (long long it = 0; it < my; it++) {
        const int s = (int)(it % SP_STAGES);
        const unsigned ph = (unsigned)((it / SP_STAGES) & 1);
        if (tid == 0 && it + SP_STAGES - 1 < my) issue(it + SP_STAGES - 1);
        __syncwarp();
        mbar_wait(&full[s], ph);
        const unsigned *sp = stage_base + (size_t)s * (3 * EX_ROW) + wib * (32 * EX_WPL) + EX_WPL * lane;
        unsigned H[EX_WPL + 2], Lw[EX_WPL + 2], X[EX_WPL + 2];
#pragma unroll
        for (int v4 = 0; v4 < EX_WPL / 4; v4++) {
            const uint4 h4 = *reinterpret_cast<const uint4 *>(sp + 4 * v4), l4 = *reinterpret_cast<const uint4 *>(sp + EX_ROW + 4 * v4),
                        x4 = *reinterpret_cast<const uint4 *>(sp + 2 * EX_ROW + 4 * v4);
            H[4 * v4] = h4.x; H[4 * v4 + 1] = h4.y; H[4 * v4 + 2] = h4.z; H[4 * v4 + 3] = h4.w;
            Lw[4 * v4] = l4.x; Lw[4 * v4 + 1] = l4.y; Lw[4 * v4 + 2] = l4.z; Lw[4 * v4 + 3] = l4.w;
            X[4 * v4] = x4.x; X[4 * v4 + 1] = x4.y; X[4 * v4 + 2] = x4.z; X[4 * v4 + 3] = x4.w;
        }
        {
            const uint2 h2 = *reinterpret_cast<const uint2 *>(sp + EX_WPL), l2 = *reinterpret_cast<const uint2 *>(sp + EX_ROW + EX_WPL),
                        x2 = *reinterpret_cast<const uint2 *>(sp + 2 * EX_ROW + EX_WPL);
            H[EX_WPL] = h2.x; H[EX_WPL + 1] = h2.y; Lw[EX_WPL] = l2.x; Lw[EX_WPL + 1] = l2.y; X[EX_WPL] = x2.x; X[EX_WPL + 1] = x2.y;
        }
        __syncwarp();
        {
            const unsigned dep = (H[0] ^ H[4] ^ H[EX_WPL]) ^ (Lw[0] ^ Lw[4] ^ Lw[EX_WPL]) ^ (X[0] ^ X[4] ^ X[EX_WPL]);
            if (lane == 0) mbar_arrive_after_loads(smem_u32(&empty[s]), dep, smem_u32(&dep_sink[wib]));   // this warp's slice is in registers
        }
        const long long bt = blockIdx.x + it * gridDim.x;
        const long long wbase = ((a.tile0 * 128) + bt * EX_WORDS + wib * (32 * EX_WPL)) * 32;   // first pattern start of the warp tile
        const unsigned lrel = (unsigned)(EX_WPL * lane) * 32;
#pragma unroll 1
        for (int pi = 0; pi < a.npat; pi++) {
            const ApxPat &pt = a.pat[pi];
            unsigned P[EX_WPL + 2];
            unsigned U[EX_WPL], M[EX_WPL];
            unsigned miss[ROWS][EX_WPL];                  // miss[r] : more than r counted chunks missing
#pragma unroll
            for (int w = 0; w < EX_WPL; w++) { U[w] = 0; M[w] = 0; }
#pragma unroll
            for (int r = 0; r < ROWS; r++)
#pragma unroll
                for (int w = 0; w < EX_WPL; w++) miss[r][w] = 0;
            const int nch = pt.nch, win = pt.win;
#pragma unroll 1
            for (int g = 0; g < nch; g++) {
                const ApxChunk &ch = pt.ch[g];
                unsigned G[EX_WPL + 2];
#pragma unroll
                for (int w = 0; w < EX_WPL + 2; w++) G[w] = 0xffffffffu;
                const int npos = ch.npos;
                for (int c = 0; c < npos; c++) {
                    const int t = ch.t[c];
                    sp_plane_dyn(ch.cls[c], P, H, Lw, X);
#pragma unroll
                    for (int w = 0; w < EX_WPL + 1; w++) G[w] &= __funnelshift_r(P[w], P[w + 1], t);
                    G[EX_WPL + 1] &= P[EX_WPL + 1] >> t;
                }
                if (ch.piece != 0xff) {                   // the undilated chunk plane is a factor of its piece
                    if (ch.first) {
#pragma unroll
                        for (int w = 0; w < EX_WPL; w++) M[w] = 0xffffffffu;
                    }
                    exact_apply<true>(M, G, ch.poff);
                    if (ch.last) {
#pragma unroll
                        for (int w = 0; w < EX_WPL; w++) U[w] |= M[w];
                    }
                }
                if (ch.counted) {
                    // dilate: G[x] |= G[x+1] | ... | G[x+win-1]   (doubling)
                    int cw = 1;
                    while (cw < win) {
                        const int sft = min(cw, win - cw);
#pragma unroll
                        for (int w = 0; w < EX_WPL + 1; w++) G[w] |= __funnelshift_r(G[w], G[w + 1], sft);
                        G[EX_WPL + 1] |= G[EX_WPL + 1] >> sft;
                        cw += sft;
                    }
                    unsigned x[EX_WPL];
#pragma unroll
                    for (int w = 0; w < EX_WPL; w++) x[w] = 0xffffffffu;
                    exact_apply<true>(x, G, ch.off);
#pragma unroll
                    for (int w = 0; w < EX_WPL; w++) {
                        const unsigned ms = ~x[w];
#pragma unroll
                        for (int r = ROWS - 1; r > 0; r--) miss[r][w] |= miss[r - 1][w] & ms;
                        miss[0][w] |= ms;
                    }
                }
            }
            // pieces that the chunks do not cover: position by position
            {
                int dp = 0;
#pragma unroll 1
                for (int i = 0; i < pt.npieces; i++) {
                    const int nd = pt.dn[i];
                    if (pt.dwild[i]) {
#pragma unroll
                        for (int w = 0; w < EX_WPL; w++) U[w] = 0xffffffffu;
                    }
                    if (nd == 0) continue;
#pragma unroll
                    for (int w = 0; w < EX_WPL; w++) M[w] = 0xffffffffu;
                    for (int j = 0; j < nd; j++) {
                        sp_plane_dyn(pt.dcls[dp + j], P, H, Lw, X);
                        exact_apply<true>(M, P, pt.dshift[dp + j]);
                    }
                    dp += nd;
#pragma unroll
                    for (int w = 0; w < EX_WPL; w++) U[w] |= M[w];
                }
            }
            unsigned mine = 0;
#pragma unroll
            for (int w = 0; w < EX_WPL; w++) { U[w] &= ~miss[ROWS - 1][w]; mine += __popc(U[w]); }
            const unsigned have = __ballot_sync(0xffffffffu, mine != 0);
            if (!have) continue;
            unsigned incl, total;
            if (!__any_sync(0xffffffffu, mine > 1)) {
                incl = __popc(have & (0xffffffffu >> (31 - lane)));
                total = __popc(have);
            } else {
                incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                total = __shfl_sync(0xffffffffu, incl, 31);
            }
            const unsigned long long b0 = (unsigned long long)(wbase + lrel);
            if (qcnt + total <= AX_QUEUE) {
                // sparse: every lane appends its pattern starts at its own offset, then full rounds are drained
                unsigned slot = qhead + qcnt + (incl - mine);
#pragma unroll
                for (int w = 0; w < EX_WPL; w++) {
                    unsigned c = U[w];
                    while (c) {
                        const int bb = __ffs(c) - 1;
                        c &= c - 1;
                        qk[slot++ & (AX_QUEUE - 1)] = ((b0 + (unsigned)(w * 32 + bb)) << 1) | (unsigned)pi;
                    }
                }
                qcnt += total;
                __syncwarp();
                while (qcnt >= 32) {
                    apx_round<K, DW>(qk, qhead, 32, a.hi, a.lo, a.xx, s_sp, a.n, a.keys, a.count, a.cap);
                    qhead = (qhead + 32) & (AX_QUEUE - 1);
                    qcnt -= 32;
                }
            } else {
                // dense: one pattern start per lane and step, a round as soon as 32 are queued
#pragma unroll
                for (int w = 0; w < EX_WPL; w++) sC[w * 32 + lane] = U[w];
#pragma unroll 1
                for (int w = 0; w < EX_WPL; w++) {
                    unsigned c = sC[w * 32 + lane];
                    while (__any_sync(0xffffffffu, c != 0)) {
                        const bool hv = c != 0;
                        const unsigned bal = __ballot_sync(0xffffffffu, hv);
                        if (hv) {
                            const int bb = __ffs(c) - 1;
                            c &= c - 1;
                            qk[(qhead + qcnt + __popc(bal & ((1u << lane) - 1u))) & (AX_QUEUE - 1)] = ((b0 + (unsigned)(w * 32 + bb)) << 1) | (unsigned)pi;
                        }
                        qcnt += __popc(bal);
                        __syncwarp();
                        while (qcnt >= 32) {
                            apx_round<K, DW>(qk, qhead, 32, a.hi, a.lo, a.xx, s_sp, a.n, a.keys, a.count, a.cap);
                            qhead = (qhead + 32) & (AX_QUEUE - 1);
                            qcnt -= 32;
                        }
                    }
                }
            }
        }
    }
    if (qcnt) apx_round<K, DW>(qk, qhead, qcnt, a.hi, a.lo, a.xx, s_sp, a.n, a.keys, a.count, a.cap);
}

// ---------------------------------------------------------------------------------------
// Multi-pattern exact scan (batched motifs): the block tile is staged once and every pattern of
// the batch is evaluated on it from registers, so HBM traffic is paid once per batch and the
// kernel is bound by the integer pipe.  Patterns are k = 0, at most 32 positions, with classes
// that are exact over the packed alphabet (A,C,G,T,X).  Keys carry the pattern id:
//     key = pid << 40 | window start << 4
struct MultiPat {
    unsigned short m, nent;
    unsigned short ent[32];              // sel (bits 0-2) | shift (bits 3-7) | class A,C,G,T,X (bits 8-12)
};

struct MultiArgs {
    const unsigned *hi, *lo, *xx;
    long long nwords, n, a0, a1, tile0, ntiles;
    const MultiPat *pats;
    int npat;
    unsigned long long *keys, *count;
    long long cap;
    const unsigned *pid_map;             // index in the caller's batch of every pattern handed to this launch (nullptr: identity)
    unsigned long long bad;              // placeholder key of out-of-range hits: sorts after every real key
};

#define MP_CHUNK 32                      // pattern descriptors staged in shared memory at a time
#define MP_HITBUF 96

__global__ void __launch_bounds__(256) k_scan_packed_multi(const MultiArgs a)
{
    __shared__ unsigned sh[3 * BK_ROW];
    __shared__ MultiPat spat[MP_CHUNK];
    __shared__ unsigned long long hitbuf_all[8][MP_HITBUF];
    const int tid = threadIdx.x, lane = tid & 31;
    unsigned long long *hitbuf = hitbuf_all[tid >> 5];
    unsigned nbuf = 0;
    auto flush = [&]() {
        if (nbuf == 0) return;
        unsigned long long basei = 0;
        if (lane == 0) basei = atomicAdd(a.count, (unsigned long long)nbuf);
        basei = __shfl_sync(0xffffffffu, basei, 0);
        for (unsigned e = lane; e < nbuf; e += 32)
            if ((long long)(basei + e) < a.cap) a.keys[basei + e] = hitbuf[e];
        __syncwarp();
        nbuf = 0;
    };
    const long long nbt = (a.ntiles + 7) / 8;
    for (long long bt = blockIdx.x; bt < nbt; bt += gridDim.x) {
        const long long qb = a.tile0 * 128 + bt * BK_WORDS;
        const long long q0 = qb + 4 * tid;
        __syncthreads();
        {
            const uint4 h4 = __ldg(reinterpret_cast<const uint4 *>(a.hi + q0));
            const uint4 l4 = __ldg(reinterpret_cast<const uint4 *>(a.lo + q0));
            const uint4 x4 = __ldg(reinterpret_cast<const uint4 *>(a.xx + q0));
            *reinterpret_cast<uint4 *>(sh + PK_HALO + 4 * tid) = h4;
            *reinterpret_cast<uint4 *>(sh + BK_ROW + PK_HALO + 4 * tid) = l4;
            *reinterpret_cast<uint4 *>(sh + 2 * BK_ROW + PK_HALO + 4 * tid) = x4;
            if (tid < PK_HALO) {
                const long long q = qb + BK_WORDS + tid;
                const bool ok = q < a.nwords;
                sh[PK_HALO + BK_WORDS + tid] = ok ? __ldg(a.hi + q) : 0u;
                sh[BK_ROW + PK_HALO + BK_WORDS + tid] = ok ? __ldg(a.lo + q) : 0u;
                sh[2 * BK_ROW + PK_HALO + BK_WORDS + tid] = ok ? __ldg(a.xx + q) : 0xffffffffu;
            }
        }
        __syncthreads();
        unsigned PA[5], PC[5], PG[5], PT[5], PX[5];        // 4 words + 1 halo word: patterns are at most 32 long
#pragma unroll
        for (int w = 0; w < 5; w++) {
            const unsigned h = sh[PK_HALO + 4 * tid + w], l = sh[BK_ROW + PK_HALO + 4 * tid + w],
                           x = sh[2 * BK_ROW + PK_HALO + 4 * tid + w];
            PX[w] = x;
            PA[w] = ~(h | l | x);
            PC[w] = l & ~h;
            PG[w] = h & l;
            PT[w] = h & ~l;
        }
        const long long tbase = (qb + 4 * tid) * 32;        // text position of this thread's first window start
        const long long wend = a.a1 < a.n + 1 ? a.a1 : a.n + 1;
        for (int c0 = 0; c0 < a.npat; c0 += MP_CHUNK) {
            __syncthreads();
            {
                // stage the next chunk of descriptors (68 bytes each) as 32-bit words
                const int nc = min(MP_CHUNK, a.npat - c0);
                const unsigned *src = reinterpret_cast<const unsigned *>(a.pats + c0);
                unsigned *dst = reinterpret_cast<unsigned *>(spat);
                for (int x = tid; x < nc * (int)(sizeof(MultiPat) / 4); x += 256) dst[x] = __ldg(src + x);
            }
            __syncthreads();
            const int nc = min(MP_CHUNK, a.npat - c0);
            for (int pb = 0; pb < nc; pb++) {
                const int nent = spat[pb].nent, m = spat[pb].m;
                unsigned M0 = ~0u, M1 = ~0u, M2 = ~0u, M3 = ~0u;
                for (int e = 0; e < nent; e++) {
                    const unsigned ent = spat[pb].ent[e];
                    const int sel = ent & 7, shf = (ent >> 3) & 31;
                    unsigned E0, E1, E2, E3, E4;
                    switch (sel) {
                    case 0: E0 = PA[0]; E1 = PA[1]; E2 = PA[2]; E3 = PA[3]; E4 = PA[4]; break;
                    case 1: E0 = PC[0]; E1 = PC[1]; E2 = PC[2]; E3 = PC[3]; E4 = PC[4]; break;
                    case 2: E0 = PG[0]; E1 = PG[1]; E2 = PG[2]; E3 = PG[3]; E4 = PG[4]; break;
                    case 3: E0 = PT[0]; E1 = PT[1]; E2 = PT[2]; E3 = PT[3]; E4 = PT[4]; break;
                    case 4: E0 = PX[0]; E1 = PX[1]; E2 = PX[2]; E3 = PX[3]; E4 = PX[4]; break;
                    default: {
                        const unsigned c = ent >> 8;
                        const unsigned sA = (c & 1) ? ~0u : 0u, sC = (c & 2) ? ~0u : 0u, sG = (c & 4) ? ~0u : 0u, sT = (c & 8) ? ~0u : 0u,
                                       sX = (c & 16) ? ~0u : 0u;
                        E0 = (PA[0] & sA) | (PC[0] & sC) | (PG[0] & sG) | (PT[0] & sT) | (PX[0] & sX);
                        E1 = (PA[1] & sA) | (PC[1] & sC) | (PG[1] & sG) | (PT[1] & sT) | (PX[1] & sX);
                        E2 = (PA[2] & sA) | (PC[2] & sC) | (PG[2] & sG) | (PT[2] & sT) | (PX[2] & sX);
                        E3 = (PA[3] & sA) | (PC[3] & sC) | (PG[3] & sG) | (PT[3] & sT) | (PX[3] & sX);
                        E4 = (PA[4] & sA) | (PC[4] & sC) | (PG[4] & sG) | (PT[4] & sT) | (PX[4] & sX);
                    }
                    }
                    M0 &= __funnelshift_r(E0, E1, shf);
                    M1 &= __funnelshift_r(E1, E2, shf);
                    M2 &= __funnelshift_r(E2, E3, shf);
                    M3 &= __funnelshift_r(E3, E4, shf);
                }
                const unsigned mine = __popc(M0) + __popc(M1) + __popc(M2) + __popc(M3);
                if (!__any_sync(0xffffffffu, mine != 0)) continue;
                unsigned incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned vv = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += vv;
                }
                const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
                const unsigned long long pidbits = (unsigned long long)(a.pid_map ? __ldg(a.pid_map + c0 + pb) : (unsigned)(c0 + pb)) << 40;
                const unsigned long long bad = a.bad;      // sorts after every real key, cut off by the host
                const unsigned Mw[4] = {M0, M1, M2, M3};
                if (total > MP_HITBUF) {
                    flush();
                    unsigned long long basei = 0;
                    if (lane == 0) basei = atomicAdd(a.count, (unsigned long long)total);
                    basei = __shfl_sync(0xffffffffu, basei, 0) + (incl - mine);
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                        unsigned c = Mw[w];
                        while (c) {
                            const int b = __ffs(c) - 1;
                            c &= c - 1;
                            const long long p = tbase + w * 32 + b;
                            const bool ok = p >= a.a0 && p < wend && p + m <= a.n;
                            if ((long long)basei < a.cap) a.keys[basei] = ok ? (pidbits | ((unsigned long long)p << 4)) : bad;
                            if (!ok) atomicAdd(a.count + 1, 1ULL);
                            basei++;
                        }
                    }
                    continue;
                }
                if (nbuf + total > MP_HITBUF) flush();
                unsigned slot = nbuf + (incl - mine);
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    unsigned c = Mw[w];
                    while (c) {
                        const int b = __ffs(c) - 1;
                        c &= c - 1;
                        const long long p = tbase + w * 32 + b;
                        const bool ok = p >= a.a0 && p < wend && p + m <= a.n;
                        hitbuf[slot++] = ok ? (pidbits | ((unsigned long long)p << 4)) : bad;
                        if (!ok) atomicAdd(a.count + 1, 1ULL);
                    }
                }
                nbuf += total;
                __syncwarp();
            }
        }
    }
    flush();
}

// ---------------------------------------------------------------------------------------
// Multi-pattern exact scan by q-gram lookup (large batches).  Evaluating every motif at every position costs
// npat x positions operations; here every text position is hashed ONCE: the 8 bases starting at it form a 16-bit code
// (8 bits of the hi plane | 8 bits of the lo plane), a CSR table built on the host maps the code to the motifs whose
// most selective 8-position window (all of its classes inside ACGT) accepts that 8-mer, and only those motifs are
// verified -- on the packed planes, all positions at once:  ((A & ma) | (C & mc) | (G & mg) | (T & mt) | (X & mx))
// must cover the motif's length.  With 10 000 IUPAC motifs that is < 1 candidate per text position instead of
// 10 000 evaluations.  Text windows that contain a non-ACGT symbol cannot match such a window and are skipped.  Motifs
// without a usable window (shorter than 8, wildcards everywhere, too many expansions) stay with k_scan_packed_multi.
// Same keys as there: pid << 40 | window start << 4.   Staging: 1-D bulk TMA copies, 3-stage ring, producer warp.
#define MH_Q 8
#define MH_NLEV 3                       // lookup tables by window length 8, 6, 4 (motifs without a longer wildcard-free window)
#define MH_BUCKETS (65536 + 4096 + 256)
__host__ __device__ __forceinline__ int mh_q(int lev) { return lev == 0 ? 8 : lev == 1 ? 6 : 4; }
__host__ __device__ __forceinline__ unsigned mh_base(int lev) { return lev == 0 ? 0u : lev == 1 ? 65536u : 65536u + 4096u; }
#define MH_WORDS 1024
#define MH_FRONT 4
#define MH_ROW (MH_FRONT + MH_WORDS + 4)
#define MH_STAGES 3
#define MH_STAGE_BYTES (3 * MH_ROW * 4)
#define MH_WPL 8
#define MH_WARPS 4
#define MH_HITBUF 256
struct HashPat { unsigned ma, mc, mg, mt, mx, lenmask, m, pid; };   // position masks per symbol; pid = index in the caller's batch
struct HashArgs {
    const unsigned *hi, *lo, *xx;
    long long nwords, n, ntiles;           // ntiles = block tiles of MH_WORDS words, the first one is tile0
    long long tile0, a0, a1;               // motif starts a0 <= w < a1 only (a range of buffer fills: text-sharded batches)
    const unsigned *offs;                  // MH_BUCKETS + 1 CSR offsets by (level, q-mer code)
    const unsigned *ents;                  // entries: window offset inside the motif << 20 | index into pats
    const HashPat *pats;
    unsigned long long *keys, *count;
    long long cap;
};

__global__ void __launch_bounds__((MH_WARPS + 1) * 32, 4) k_scan_multi_hash(const HashArgs a)
{
    extern __shared__ __align__(128) unsigned char ex_smem[];
    __shared__ unsigned long long hitbuf_all[MH_WARPS + 1][MH_HITBUF];
    unsigned long long *hitbuf = hitbuf_all[threadIdx.x >> 5];
    unsigned nbuf = 0;                                      // warp-uniform fill of hitbuf
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    auto flush = [&]() {
        if (nbuf == 0) return;
        unsigned long long basei = 0;
        if (lane == 0) basei = atomicAdd(a.count, (unsigned long long)nbuf);
        basei = __shfl_sync(0xffffffffu, basei, 0);
        for (unsigned e = lane; e < nbuf; e += 32)
            if ((long long)(basei + e) < a.cap) a.keys[basei + e] = hitbuf[e];
        __syncwarp();
        nbuf = 0;
    };
    unsigned *stage_base = reinterpret_cast<unsigned *>(ex_smem);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(ex_smem + MH_STAGES * MH_STAGE_BYTES);
    unsigned long long *empty = full + MH_STAGES;
    const long long my = blockIdx.x < a.ntiles ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < MH_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], MH_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (wib == MH_WARPS) {
        if (lane == 0) {
            int s = 0;
            unsigned ph = 1;
            for (long long it = 0; it < my; it++) {
                if (it >= MH_STAGES) {
                    // the consumers hold a stage for a whole tile: poll at leisure instead of burning issue slots
                    const unsigned bar = smem_u32(&empty[s]);
                    for (;;) {
                        unsigned done;
                        asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(ph) : "memory");
                        if (done) break;
                        __nanosleep(400);
                    }
                }
                const long long q = (a.tile0 + blockIdx.x + it * gridDim.x) * (long long)MH_WORDS;
                long long words = MH_ROW;
                if (q + MH_WORDS + 4 > a.nwords) words = a.nwords - q + MH_FRONT;
                const unsigned bytes = (unsigned)words * 4u;
                unsigned *dst = stage_base + (size_t)s * (3 * MH_ROW);
                mbar_expect_tx(&full[s], 3u * bytes);
                tma_load_1d(dst, a.hi + q - MH_FRONT, bytes, &full[s]);
                tma_load_1d(dst + MH_ROW, a.lo + q - MH_FRONT, bytes, &full[s]);
                tma_load_1d(dst + 2 * MH_ROW, a.xx + q - MH_FRONT, bytes, &full[s]);
                if (++s == MH_STAGES) { s = 0; ph ^= 1u; }
            }
        }
        return;
    }
    int s = 0;
    unsigned ph = 0;
    for (long long it = 0; it < my; it++) {
        mbar_wait(&full[s], ph);
        // lane segment: words [seg, seg + 8) of the tile; index 0 of the register arrays is the word before it
        const int seg = (wib * 32 + lane) * MH_WPL;
        const unsigned *sp = stage_base + (size_t)s * (3 * MH_ROW) + MH_FRONT + seg - 1;
        unsigned H[MH_WPL + 2], L[MH_WPL + 2], X[MH_WPL + 2];
        const long long q = (a.tile0 + blockIdx.x + it * gridDim.x) * (long long)MH_WORDS;
#pragma unroll
        for (int w = 0; w < MH_WPL + 2; w++) { H[w] = sp[w]; L[w] = sp[MH_ROW + w]; X[w] = sp[2 * MH_ROW + w]; }
        const int s_rel = s;                             // released after the tile has been processed (see mbar_arrive_after_loads)
        if (++s == MH_STAGES) { s = 0; ph ^= 1u; }
        const long long segpos = (q + seg) * 32;                 // text position of bit 0 of H[1]
#pragma unroll
        for (int t = 0; t < MH_WPL; t++) {
            const unsigned h0 = H[t], h1 = H[t + 1], h2 = H[t + 2], l0 = L[t], l1 = L[t + 1], l2 = L[t + 2], x0 = X[t], x1 = X[t + 1], x2 = X[t + 2];
#pragma unroll 1
            for (int r = 0; r < 32; r++) {
                // q-mers that start at bit r of word t + 1
                const unsigned xs8 = __funnelshift_r(x1, x2, r) & 0xffu, hs8 = __funnelshift_r(h1, h2, r) & 0xffu, ls8 = __funnelshift_r(l1, l2, r) & 0xffu;
#pragma unroll 1
                for (int lev = 0; lev < MH_NLEV; lev++) {
                const int qq = mh_q(lev);
                const unsigned qm = (1u << qq) - 1u;
                const unsigned code = mh_base(lev) + (((hs8 & qm) << qq) | (ls8 & qm));
                unsigned e0 = 0, e1 = 0;
                if ((xs8 & qm) == 0) { e0 = __ldg(a.offs + code); e1 = __ldg(a.offs + code + 1); }
                while (__any_sync(0xffffffffu, e0 < e1)) {
                    bool ok = false;
                    unsigned long long key = 0;
                    if (e0 < e1) {
                        const unsigned ent = __ldg(a.ents + e0);
                        e0++;
                        const uint4 pa = __ldg(reinterpret_cast<const uint4 *>(a.pats + (ent & 0xfffffu)));
                        const uint4 pb = __ldg(reinterpret_cast<const uint4 *>(a.pats + (ent & 0xfffffu)) + 1);
                        const int sh = 32 + r - (int)(ent >> 20);            // window start relative to bit 0 of word t: 8 .. 63
                        unsigned hw, lw, xw;
                        if (sh < 32) { hw = __funnelshift_r(h0, h1, sh); lw = __funnelshift_r(l0, l1, sh); xw = __funnelshift_r(x0, x1, sh); }
                        else { hw = __funnelshift_r(h1, h2, sh - 32); lw = __funnelshift_r(l1, l2, sh - 32); xw = __funnelshift_r(x1, x2, sh - 32); }
                        const unsigned A = ~(hw | lw | xw), C = lw & ~hw, G = hw & lw, T = hw & ~lw;
                        const unsigned acc = (A & pa.x) | (C & pa.y) | (G & pa.z) | (T & pa.w) | (xw & pb.x);
                        const long long w = segpos + t * 32 + r - (long long)(ent >> 20);
                        ok = (acc & pb.y) == pb.y && w >= a.a0 && w < a.a1 && w + (long long)pb.z <= a.n;
                        key = ((unsigned long long)pb.w << 40) | ((unsigned long long)w << 4);
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, ok);
                    if (bal) {
                        const unsigned cnt = __popc(bal);
                        if (nbuf + cnt > MH_HITBUF) flush();
                        if (ok) hitbuf[nbuf + __popc(bal & ((1u << lane) - 1u))] = key;
                        nbuf += cnt;
                        __syncwarp();
                    }
                }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s_rel]);
    }
    flush();
}

// last k with S[k] <= x in the table of buffer-fill starts (0 when x lies in front of the first one)
__device__ __forceinline__ int fill_lookup(const long long *__restrict__ S, int n, long long x)
{
    int lo = 0, hi = n - 1;
    if (n > 8) {
        // fills are nearly equally long (each is the buffer size minus one partial line): interpolate, then close in on
        // the answer with a few probes; whatever is left goes to the bisection below
        const long long span = S[hi] - S[0];
        int g = span > 0 ? (int)((double)(x - S[0]) * (double)hi / (double)span) : 0;
        g = g < 0 ? 0 : g > hi ? hi : g;
#pragma unroll 1
        for (int probe = 0; probe < 3; probe++) {
            if (S[g] > x) { hi = g - 1; g = g - 1 - probe; if (g < lo) g = lo; }
            else if (g < hi && S[g + 1] <= x) { lo = g + 1; g = g + 1 + probe; if (g > hi) g = hi; }
            else return g;
            if (lo >= hi) break;
        }
        if (hi < lo) hi = lo;
    }
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (S[mid] <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// chain stage for batched exact patterns: keys sorted by (pattern, position)
__global__ void __launch_bounds__(128) k_chain_multi(const unsigned long long *__restrict__ keys, long long nkeys,
                                                     const unsigned short *__restrict__ mlen, const long long *__restrict__ fillS,
                                                     const long long *__restrict__ fillE, int nfills,
                                                     long long *__restrict__ hits /* beg,end pairs */, unsigned char *__restrict__ sel,
                                                     unsigned long long *__restrict__ per_pattern)
{
    // per-pattern counts: keys are sorted by pattern, so the 128 keys of a block belong to a handful of consecutive
    // motifs -- counted in shared memory relative to the block's first motif, one global atomic per block and motif
    // (one per hit, or per warp, lands on the same few L2 addresses from every SM at once)
    __shared__ unsigned cnt[128];
    cnt[threadIdx.x] = 0;
    const long long jb = (long long)blockIdx.x * blockDim.x;
    const int pid0 = jb < nkeys ? (int)(keys[jb] >> 40) : 0;
    __syncthreads();
    const long long j0 = jb + threadIdx.x;
    const unsigned long long POSMASK = (1ULL << 36) - 1;
    auto fill_idx = [&](long long x) -> int { return fill_lookup(fillS, nfills, x); };
    auto independent = [&](long long j) -> bool {
        if (j == 0) return true;
        const unsigned long long a = keys[j - 1], b = keys[j];
        if ((a >> 40) != (b >> 40)) return true;
        const long long pa = (long long)((a >> 4) & POSMASK), pb = (long long)((b >> 4) & POSMASK);
        if (pa + mlen[b >> 40] <= pb) return true;
        return fill_idx(pa) != fill_idx(pb);
    };
    if (j0 < nkeys && independent(j0)) {
        long long pos = -1;
        unsigned mine = 0;                                   // hits of this cluster (one motif)
        int pid = 0;
        for (long long t = j0; t < nkeys; t++) {
            if (t > j0 && independent(t)) break;
            const unsigned long long key = keys[t];
            pid = (int)(key >> 40);
            const long long p = (long long)((key >> 4) & POSMASK);
            const int m = mlen[pid];
            const int f = fill_idx(p);
            sel[t] = 0;
            if (p + m > fillE[f]) continue;                      // the window must lie inside its buffer fill
            if (p < pos) continue;
            hits[2 * t] = p;
            hits[2 * t + 1] = p + m;
            sel[t] = 1;
            mine++;
            pos = p + m;
        }
        if (mine) {
            const int d = pid - pid0;
            if (d >= 0 && d < 128) atomicAdd(&cnt[d], mine);
            else atomicAdd(&per_pattern[pid], (unsigned long long)mine);
        }
    }
    __syncthreads();
    if (cnt[threadIdx.x]) atomicAdd(&per_pattern[pid0 + threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
}
