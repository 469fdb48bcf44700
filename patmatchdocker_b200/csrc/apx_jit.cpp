// apx_jit.cpp -- code generator + NVRTC driver for the specialised approximate scan kernel.
//
// The generic kernel (packed.cuh: k_scan_apx) interprets the chunk description of a pattern: a 32-way switch per class
// plane, shift amounts and chunk tables read at run time, loops that cannot be unrolled.  ncu on the headline workload
// (18-nt degenerate motif, -k 2ids, both strands, 3.1 Gb): 334 warp instructions per 32-base word and pattern, ALU pipe
// 67 % busy -- the kernel is bound by the number of LOP3/SHF instructions it issues.  For genome-scale requests the
// engine therefore writes the dense part as straight-line CUDA for exactly the request's patterns and compiles it with
// NVRTC for sm_100a (about 0.2 s, cached per request text): immediate shifts and truth tables, class planes shared by
// the positions of a chunk, halo words computed only where a later step reads them, no dispatch.  Search semantics are
// those of k_scan_apx (esimpleScan @4136d0 candidates + necessary conditions of checkMatch1 @414190); the output is
// decided by k_verify on the raw bytes either way.
#include "apx_jit.hpp"
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <mutex>

static const char *kKernelBody =
#include "jit_apx_kernel.inc"
    ;

namespace {

struct Emit {
    std::string s;
    void f(const char *fmt, ...)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        s += buf;
    }
};

// truth table over (hi, lo, x) of a class given as bits A,C,G,T,X ; A=000 C=010 G=110 T=100 X=xx1 (packed.cuh: SpLut)
int lut_of(unsigned cls)
{
    return ((cls & 1) ? 0x01 : 0) | ((cls & 2) ? 0x04 : 0) | ((cls & 4) ? 0x40 : 0) | ((cls & 8) ? 0x10 : 0) | ((cls & 16) ? 0xAA : 0);
}

struct Pos { int s; unsigned cls; };          // pattern index of a constrained position, its class
struct Group {
    std::vector<Pos> pos;
    int piece = -1;                           // piece the group is a factor of
    bool pfirst = false, plast = false;       // first / last factor of that piece
    bool counted = false;                     // takes part in the q-gram count
};

// word `w` of plane `name` shifted right by `sh` bits; `bits` = number of low bits of the result that are read
std::string shifted(const char *name, int w, int sh, int bits)
{
    char buf[96];
    const int lo = w + sh / 32, r = sh % 32;
    if (r == 0) snprintf(buf, sizeof buf, "%s[%d]", name, lo);
    else if (r + bits <= 32) snprintf(buf, sizeof buf, "(%s[%d] >> %d)", name, lo, r);
    else snprintf(buf, sizeof buf, "__funnelshift_r(%s[%d], %s[%d], %d)", name, lo, name, lo + 1, r);
    return buf;
}

void gen_dense(Emit &o, const ApxPat &pt, int pi)
{
    const int K = pt.k, win = pt.win;
    std::vector<Group> groups;
    for (int g = 0; g < pt.nch; g++) {
        const ApxChunk &ch = pt.ch[g];
        Group gr;
        const int first = (int)ch.poff - K;
        for (int c = 0; c < ch.npos; c++) gr.pos.push_back({first + ch.t[c], ch.cls[c]});
        gr.piece = ch.piece == 0xff ? -1 : ch.piece;
        gr.pfirst = ch.first != 0; gr.plast = ch.last != 0;
        gr.counted = ch.counted != 0 && pt.ncounted > K;
        if (gr.piece < 0 && !gr.counted) continue;
        groups.push_back(gr);
    }
    {
        int dp = 0;
        for (int i = 0; i < pt.npieces; i++) {
            const int nd = pt.dn[i];
            if (nd == 0) continue;
            Group gr;
            for (int j = 0; j < nd; j++) gr.pos.push_back({(int)pt.dshift[dp + j] - K, pt.dcls[dp + j]});
            gr.piece = i; gr.pfirst = gr.plast = true; gr.counted = false;
            groups.push_back(gr);
            dp += nd;
        }
    }
    bool wild = false;
    for (int i = 0; i < pt.npieces; i++) wild = wild || pt.dwild[i];
    int ncounted = 0;
    for (const Group &g : groups) ncounted += g.counted ? 1 : 0;

    o.f("__device__ __forceinline__ void dense_%d(const unsigned (&H)[10], const unsigned (&L)[10], const unsigned (&X)[10], unsigned (&U)[8])\n{\n", pi);
    o.f("    unsigned P[10], G[9], T[9], M[8], D[8];\n");
    if (ncounted) o.f("    unsigned c0[8], c1[8], c2[8], c3[8];\n");
    o.f("    (void)P; (void)G; (void)T; (void)M; (void)D;\n");
    int cur_cls = -1, cur_pw = 0;             // class plane held in P and how many of its words are valid
    bool u_set = false;
    int seen = 0;                             // counted chunks so far
    for (size_t gi = 0; gi < groups.size(); gi++) {
        Group &g = groups[gi];
        const bool direct = !g.counted;       // read only at the nominal diagonal: evaluate there, 8 words
        // bits of G beyond 255 that later steps read
        int ext = 0;
        if (!direct) ext = std::max(win > 1 ? win - 1 : K, g.piece >= 0 ? K : 0);
        const int nG = ext > 0 ? 9 : 8;
        std::stable_sort(g.pos.begin(), g.pos.end(), [](const Pos &a, const Pos &b) { return a.cls < b.cls; });
        o.f("    // group %zu: %zu position(s)%s%s\n", gi, g.pos.size(), g.counted ? ", counted" : "", g.piece >= 0 ? ", factor of a piece" : "");
        bool first_term = true;
        for (size_t a = 0; a < g.pos.size();) {
            size_t b = a;
            int maxs = 0;
            while (b < g.pos.size() && g.pos[b].cls == g.pos[a].cls) { maxs = std::max(maxs, g.pos[b].s + (direct ? K : 0)); b++; }
            const unsigned cls = g.pos[a].cls;
            const int pw = (255 + ext + maxs) / 32 + 1;
            if ((int)cls != cur_cls || cur_pw < pw) {
                for (int w = 0; w < pw; w++) {
                    if (cls == 16u) o.f("    P[%d] = X[%d];\n", w, w);
                    else o.f("    P[%d] = lop3_<0x%02x>(H[%d], L[%d], X[%d]);\n", w, lut_of(cls), w, w, w);
                }
                cur_cls = (int)cls; cur_pw = pw;
            }
            for (size_t c = a; c < b; c++) {
                const int sh = g.pos[c].s + (direct ? K : 0);
                for (int w = 0; w < nG; w++) {
                    const std::string t = shifted("P", w, sh, w < 8 ? 32 : ext);
                    if (first_term) o.f("    G[%d] = %s;\n", w, t.c_str());
                    else o.f("    G[%d] &= %s;\n", w, t.c_str());
                }
                first_term = false;
            }
            a = b;
        }
        if (g.piece >= 0) {
            for (int w = 0; w < 8; w++) {
                const std::string t = direct ? std::string("G[") + std::to_string(w) + "]" : shifted("G", w, K, 32);
                if (g.pfirst) o.f("    M[%d] = %s;\n", w, t.c_str());
                else o.f("    M[%d] &= %s;\n", w, t.c_str());
            }
            if (g.plast) {
                for (int w = 0; w < 8; w++) o.f(u_set ? "    U[%d] |= M[%d];\n" : "    U[%d] = M[%d];\n", w, w);
                u_set = true;
            }
        }
        if (g.counted) {
            // D[b] = OR of G[b .. b+win-1] (the chunk occurs within +-K of its nominal place); substitutions only: G[b+K]
            if (win == 1) {
                for (int w = 0; w < 8; w++) o.f("    D[%d] = %s;\n", w, shifted("G", w, K, 32).c_str());
            } else if (win == 3) {
                for (int w = 0; w < 8; w++) o.f("    D[%d] = G[%d] | %s | %s;\n", w, w, shifted("G", w, 1, 32).c_str(), shifted("G", w, 2, 32).c_str());
            } else {
                // T = window of 3, then D = T | T>>2 (5) or T | T>>2 | T>>4 (7)
                const int text = win - 3;             // bits of T beyond 255 that D reads
                for (int w = 0; w < 9; w++)
                    o.f("    T[%d] = G[%d] | %s | %s;\n", w, w, shifted("G", w, 1, w < 8 ? 32 : text).c_str(), shifted("G", w, 2, w < 8 ? 32 : text).c_str());
                for (int w = 0; w < 8; w++) {
                    if (win == 5) o.f("    D[%d] = T[%d] | %s;\n", w, w, shifted("T", w, 2, 32).c_str());
                    else o.f("    D[%d] = T[%d] | %s | %s;\n", w, w, shifted("T", w, 2, 32).c_str(), shifted("T", w, 4, 32).c_str());
                }
            }
            // saturating bit-sliced count of missing chunks: c<r> = more than r counted chunks missing so far
            for (int w = 0; w < 8; w++) {
                for (int r = std::min(seen, K); r >= 1; r--) {
                    if (r == seen) o.f("    c%d[%d] = c%d[%d] & ~D[%d];\n", r, w, r - 1, w, w);
                    else o.f("    c%d[%d] |= c%d[%d] & ~D[%d];\n", r, w, r - 1, w, w);
                }
                if (seen == 0) o.f("    c0[%d] = ~D[%d];\n", w, w);
                else o.f("    c0[%d] |= ~D[%d];\n", w, w);
            }
            seen++;
        }
    }
    if (wild || !u_set) for (int w = 0; w < 8; w++) o.f("    U[%d] = 0xffffffffu;\n", w);
    if (ncounted > K) for (int w = 0; w < 8; w++) o.f("    U[%d] &= ~c%d[%d];\n", w, K, w);
    o.f("}\n\n");
}

}  // namespace

std::string apx_generate_prefix(const ApxPat *pats, int npat)
{
    Emit o;
    bool wide = false;
    for (int p = 0; p < npat; p++) wide = wide || pats[p].m + 2 * pats[p].k > 32;
    o.f("#define AX_K %d\n#define AX_WIDE %d\n#define AX_NPAT %d\n", pats[0].k, wide ? 1 : 0, npat);
    for (int p = 0; p < npat; p++) {
        const ApxPat &pt = pats[p];
        o.f("#define AX%d_M %d\n#define AX%d_L %d\n#define AX%d_NP %d\n#define AX%d_INDEL %d\n", p, pt.m, p, pt.L, p, pt.npieces, p, pt.indel);
        for (int i = 0; i < 4; i++) o.f("#define AX%d_V%d %d\n", p, i, pt.V[i]);
        static const char *nm[5] = {"MA", "MC", "MT", "MG", "MX"};
        for (int q = 0; q < 5; q++) o.f("#define AX%d_%s 0x%llxULL\n", p, nm[q], pt.posmask[q]);
    }
    o.f("template <int LUT> __device__ __forceinline__ unsigned lop3_(unsigned a, unsigned b, unsigned c)\n"
        "{\n    unsigned r;\n    asm(\"lop3.b32 %%0, %%1, %%2, %%3, %%4;\" : \"=r\"(r) : \"r\"(a), \"r\"(b), \"r\"(c), \"n\"(LUT));\n    return r;\n}\n\n");
    for (int p = 0; p < npat; p++) gen_dense(o, pats[p], p);
    return o.s;
}

std::string apx_full_source(const std::string &prefix) { return prefix + kKernelBody; }

// ---------------------------------------------------------------------------------------
// NVRTC through dlopen: the product library must load (and export its symbols) on machines without the toolkit.
namespace {
typedef struct _nvrtcProgram *nvrtcProgram;
struct Nvrtc {
    void *h = nullptr;
    int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    int (*DestroyProgram)(nvrtcProgram *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string why;
};

Nvrtc &nvrtc()
{
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        static const char *names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so.13"};
        for (const char *nm : names) {
            n.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (n.h) break;
        }
        if (!n.h) { n.why = "libnvrtc not found (dlopen)"; return; }
#define PM_SYM(field, name) *(void **)(&n.field) = dlsym(n.h, name); if (!n.field) { n.why = std::string("libnvrtc lacks ") + name; n.h = nullptr; return; }
        PM_SYM(CreateProgram, "nvrtcCreateProgram")
        PM_SYM(CompileProgram, "nvrtcCompileProgram")
        PM_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
        PM_SYM(GetCUBIN, "nvrtcGetCUBIN")
        PM_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
        PM_SYM(GetProgramLog, "nvrtcGetProgramLog")
        PM_SYM(DestroyProgram, "nvrtcDestroyProgram")
        PM_SYM(GetErrorString, "nvrtcGetErrorString")
#undef PM_SYM
    });
    return n;
}
}  // namespace

int apx_jit_compile(const std::string &source, std::vector<char> &cubin, std::string &log)
{
    Nvrtc &n = nvrtc();
    if (!n.h) { log = n.why; return 1; }
    nvrtcProgram prog = nullptr;
    int rc = n.CreateProgram(&prog, source.c_str(), "k_scan_apx_jit.cu", 0, nullptr, nullptr);
    if (rc) { log = std::string("nvrtcCreateProgram: ") + n.GetErrorString(rc); return 1; }
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--extra-device-vectorization"};
    rc = n.CompileProgram(prog, 4, opts);
    if (rc) {
        size_t ls = 0;
        n.GetProgramLogSize(prog, &ls);
        std::string l(ls, '\0');
        if (ls) n.GetProgramLog(prog, &l[0]);
        log = std::string("nvrtcCompileProgram: ") + n.GetErrorString(rc) + "\n" + l;
        n.DestroyProgram(&prog);
        return 1;
    }
    size_t sz = 0;
    rc = n.GetCUBINSize(prog, &sz);
    if (rc || sz == 0) { log = "nvrtcGetCUBINSize failed"; n.DestroyProgram(&prog); return 1; }
    cubin.resize(sz);
    rc = n.GetCUBIN(prog, cubin.data());
    n.DestroyProgram(&prog);
    if (rc) { log = "nvrtcGetCUBIN failed"; return 1; }
    return 0;
}
