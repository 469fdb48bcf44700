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
};

__global__ void __launch_bounds__(256) k_pack(const unsigned char *__restrict__ text, long long n, long long nwords,
                                              unsigned *__restrict__ hi, unsigned *__restrict__ lo, unsigned *__restrict__ xx,
                                              unsigned long long *__restrict__ nexc)   // [0] non-ACGT bytes, [1] newlines
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long exc = 0, nls = 0;
    for (long long g = warp; g * 32 < nwords; g += nwarps) {
        unsigned mh = 0, ml = 0, mx = 0;
        const long long base = g * 1024;
#pragma unroll 4
        for (int i = 0; i < 32; i++) {
            const long long idx = base + 32 * i + lane;
            const unsigned c = idx < n ? text[idx] : 0xffu;
            const unsigned f = c | 0x20u;
            const bool acgt = (f == 'a') | (f == 'c') | (f == 'g') | (f == 't');
            const unsigned wh = __ballot_sync(0xffffffffu, acgt && (c & 4u));
            const unsigned wl = __ballot_sync(0xffffffffu, acgt && (c & 2u));
            const unsigned wx = __ballot_sync(0xffffffffu, !acgt);
            if (lane == i) { mh = wh; ml = wl; mx = wx; }
            if (idx < n && !acgt) exc++;
            if (idx < n && c == '\n') nls++;
        }
        const long long q = g * 32 + lane;
        if (q < nwords) { hi[q] = mh; lo[q] = ml; xx[q] = mx; }
    }
    for (int o = 16; o; o >>= 1) { exc += __shfl_xor_sync(0xffffffffu, exc, o); nls += __shfl_xor_sync(0xffffffffu, nls, o); }
    if (lane == 0 && exc) atomicAdd(nexc, exc);
    if (lane == 0 && nls) atomicAdd(nexc + 1, nls);
}

// positions of the record delimiter, for the host-side buffer-fill table (unordered; sorted on the host)
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

__device__ __forceinline__ void packed_apply(unsigned (&M)[4], const unsigned (&P)[6], int sh)
{
    if (sh < 32) {
#pragma unroll
        for (int w = 0; w < 4; w++) M[w] &= __funnelshift_r(P[w], P[w + 1], sh);
    } else {
#pragma unroll
        for (int w = 0; w < 4; w++) M[w] &= __funnelshift_r(P[w + 1], P[w + 2], sh - 32);
    }
}

template <int NP>
__global__ void __launch_bounds__(256) k_scan_packed(const PackedArgs<NP> a)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = warp; t < a.ntiles; t += nwarps) {
        const long long q0 = (a.tile0 + t) * 128 + 4 * lane;          // first of this lane's 4 words
        // ---- load 4 words + 2 halo words of each plane ----
        const uint4 h4 = __ldg(reinterpret_cast<const uint4 *>(a.hi + q0));
        const uint4 l4 = __ldg(reinterpret_cast<const uint4 *>(a.lo + q0));
        const uint4 x4 = __ldg(reinterpret_cast<const uint4 *>(a.xx + q0));
        const uint2 h2 = __ldg(reinterpret_cast<const uint2 *>(a.hi + q0 + 4));
        const uint2 l2 = __ldg(reinterpret_cast<const uint2 *>(a.lo + q0 + 4));
        const uint2 x2 = __ldg(reinterpret_cast<const uint2 *>(a.xx + q0 + 4));
        const unsigned H[6] = {h4.x, h4.y, h4.z, h4.w, h2.x, h2.y};
        const unsigned Lw[6] = {l4.x, l4.y, l4.z, l4.w, l2.x, l2.y};
        unsigned PA[6], PC[6], PG[6], PT[6], PX[6];
#pragma unroll
        for (int w = 0; w < 6; w++) {
            const unsigned x = w < 4 ? (&x4.x)[w] : (&x2.x)[w - 4];
            PX[w] = x;
            PA[w] = ~(H[w] | Lw[w] | x);
            PC[w] = Lw[w] & ~H[w];
            PG[w] = H[w] & Lw[w];
            PT[w] = H[w] & ~Lw[w];
        }
        // ---- pieces ----
        unsigned M[NP][4];
#pragma unroll
        for (int i = 0; i < NP; i++) {
#pragma unroll
            for (int w = 0; w < 4; w++) M[i][w] = 0xffffffffu;
            if (i < a.npieces) {
                for (int j = 0; j < a.L; j++) {
                    const PackedPos pp = a.pos[i][j];
                    switch (pp.sel) {
                    case 0: packed_apply(M[i], PA, j); break;
                    case 1: packed_apply(M[i], PC, j); break;
                    case 2: packed_apply(M[i], PG, j); break;
                    case 3: packed_apply(M[i], PT, j); break;
                    case 4: packed_apply(M[i], PX, j); break;
                    default: {
                        const unsigned sA = (pp.cls & 1) ? ~0u : 0u, sC = (pp.cls & 2) ? ~0u : 0u, sG = (pp.cls & 4) ? ~0u : 0u,
                                       sT = (pp.cls & 8) ? ~0u : 0u, sX = (pp.cls & 16) ? ~0u : 0u;
                        unsigned E[6];
#pragma unroll
                        for (int w = 0; w < 6; w++)
                            E[w] = (PA[w] & sA) | (PC[w] & sC) | (PG[w] & sG) | (PT[w] & sT) | (PX[w] & sX);
                        packed_apply(M[i], E, j);
                    }
                    }
                }
            } else {
#pragma unroll
                for (int w = 0; w < 4; w++) M[i][w] = 0;
            }
        }
        // ---- candidates: piece i fires when any piece of trigsets[i] matched at the same start ----
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
                    const long long p = (q0 + w) * 32 + b;
                    if (p >= a.a0 && p < a.a1 && p + a.L <= a.n) {
                        const unsigned long long idx = atomicAdd(a.count, 1ULL);
                        if ((long long)idx < a.cap) a.keys[idx] = ((unsigned long long)p << 4) | (unsigned)i;
                    }
                }
            }
        }
    }
}
