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
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>

static const char *kKernelBody =
#include "jit_apx_kernel.inc"
    ;
static const char *kStreamBody =
#include "jit_stream_kernel.inc"
    ;

namespace {

struct Emit {
    std::string s;
    void f(const char *fmt, ...)
    {
        char buf[512];
        va_list ap, ap2;
        va_start(ap, fmt);
        va_copy(ap2, ap);
        const int need = vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        if (need < (int)sizeof buf) s += buf;
        else {                                             // long expressions (many positions in one chunk)
            std::string big((size_t)need + 1, '\0');
            vsnprintf(&big[0], big.size(), fmt, ap2);
            big.resize((size_t)need);
            s += big;
        }
        va_end(ap2);
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

void gen_dense(Emit &o, const ApxPat &pt, int pi, int W)
{
    // W = words (of 32 pattern starts) per lane; planes carry two more words of halo
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

    o.f("__device__ __forceinline__ void dense_%d(const unsigned (&H)[%d], const unsigned (&L)[%d], const unsigned (&X)[%d], unsigned (&U)[%d])\n{\n", pi, W + 2, W + 2, W + 2, W);
    o.f("    unsigned P[%d], G[%d], T[%d], M[%d], D[%d];\n", W + 2, W + 1, W + 1, W, W);
    if (ncounted) o.f("    unsigned c0[%d], c1[%d], c2[%d], c3[%d];\n", W, W, W, W);
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
        const int nG = ext > 0 ? W + 1 : W;
        std::stable_sort(g.pos.begin(), g.pos.end(), [](const Pos &a, const Pos &b) { return a.cls < b.cls; });
        {
            std::string where;
            for (const Pos &q : g.pos) where += " " + std::to_string(q.s) + ":" + std::to_string(q.cls);
            o.f("    // group %zu: positions (index:class)%s%s%s\n", gi, where.c_str(), g.counted ? ", counted" : "", g.piece >= 0 ? ", factor of a piece" : "");
        }
        bool first_term = true;
        for (size_t a = 0; a < g.pos.size();) {
            size_t b = a;
            int maxs = 0;
            while (b < g.pos.size() && g.pos[b].cls == g.pos[a].cls) { maxs = std::max(maxs, g.pos[b].s + (direct ? K : 0)); b++; }
            const unsigned cls = g.pos[a].cls;
            const int pw = (32 * W - 1 + ext + maxs) / 32 + 1;
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
                    const std::string t = shifted("P", w, sh, w < W ? 32 : ext);
                    if (first_term) o.f("    G[%d] = %s;\n", w, t.c_str());
                    else o.f("    G[%d] &= %s;\n", w, t.c_str());
                }
                first_term = false;
            }
            a = b;
        }
        if (g.piece >= 0) {
            for (int w = 0; w < W; w++) {
                const std::string t = direct ? std::string("G[") + std::to_string(w) + "]" : shifted("G", w, K, 32);
                if (g.pfirst) o.f("    M[%d] = %s;\n", w, t.c_str());
                else o.f("    M[%d] &= %s;\n", w, t.c_str());
            }
            if (g.plast) {
                for (int w = 0; w < W; w++) o.f(u_set ? "    U[%d] |= M[%d];\n" : "    U[%d] = M[%d];\n", w, w);
                u_set = true;
            }
        }
        if (g.counted) {
            // D[b] = OR of G[b .. b+win-1] (the chunk occurs within +-K of its nominal place); substitutions only: G[b+K]
            if (win == 1) {
                for (int w = 0; w < W; w++) o.f("    D[%d] = %s;\n", w, shifted("G", w, K, 32).c_str());
            } else if (win == 3) {
                for (int w = 0; w < W; w++) o.f("    D[%d] = G[%d] | %s | %s;\n", w, w, shifted("G", w, 1, 32).c_str(), shifted("G", w, 2, 32).c_str());
            } else {
                // T = window of 3, then D = T | T>>2 (5) or T | T>>2 | T>>4 (7)
                const int text = win - 3;             // bits of T beyond 255 that D reads
                for (int w = 0; w <= W; w++)
                    o.f("    T[%d] = G[%d] | %s | %s;\n", w, w, shifted("G", w, 1, w < W ? 32 : text).c_str(), shifted("G", w, 2, w < W ? 32 : text).c_str());
                for (int w = 0; w < W; w++) {
                    if (win == 5) o.f("    D[%d] = T[%d] | %s;\n", w, w, shifted("T", w, 2, 32).c_str());
                    else o.f("    D[%d] = T[%d] | %s | %s;\n", w, w, shifted("T", w, 2, 32).c_str(), shifted("T", w, 4, 32).c_str());
                }
            }
            // saturating bit-sliced count of missing chunks: c<r> = more than r counted chunks missing so far
            for (int w = 0; w < W; w++) {
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
    if (wild || !u_set) for (int w = 0; w < W; w++) o.f("    U[%d] = 0xffffffffu;\n", w);
    if (ncounted > K) for (int w = 0; w < W; w++) o.f("    U[%d] &= ~c%d[%d];\n", w, K, w);
    o.f("}\n\n");
}


// ---------------------------------------------------------------------------------------
// Streaming form (jit_stream_kernel.inc): dense_step consumes one word of the planes and slides its state.
// End coordinates: e = b + SPAN, SPAN = m - 1 + 2K.  Position j of the pattern on diagonal d (0..2K) sits at text
// e - (m-1-j) - (2K-d); a chunk plane is G[y] = AND_j P_j[y - (m-1-j)] ("the pattern would end at y on diagonal 2K"),
// so the chunk occurs within +-K of its nominal place iff OR_{i=0..2K} G[e-i], and at the nominal place iff G[e-K].
struct SGroup {
    std::vector<Pos> pos;
    int piece = -1;
    bool pfirst = false, plast = false, counted = false;
};

// plane `name` looked back by `sh` bits: current word = name, earlier words = name_1, name_2 (fields of the state)
static std::string back(const std::string &name, int sh)
{
    const int q = sh / 32, r = sh % 32;
    auto word = [&](int k) { return k == 0 ? name : "s." + name + "_" + std::to_string(k); };
    if (r == 0) return word(q);
    return "__funnelshift_l(" + word(q + 1) + ", " + word(q) + ", " + std::to_string(r) + ")";
}

static void collect_groups(const ApxPat &pt, std::vector<SGroup> &groups, bool &wild)
{
    const int K = pt.k;
    for (int g = 0; g < pt.nch; g++) {
        const ApxChunk &ch = pt.ch[g];
        SGroup gr;
        const int first = (int)ch.poff - K;
        for (int c = 0; c < ch.npos; c++) gr.pos.push_back({first + ch.t[c], ch.cls[c]});
        gr.piece = ch.piece == 0xff ? -1 : ch.piece;
        gr.pfirst = ch.first != 0; gr.plast = ch.last != 0;
        gr.counted = ch.counted != 0 && pt.ncounted > K;
        if (gr.piece < 0 && !gr.counted) continue;
        groups.push_back(gr);
    }
    int dp = 0;
    for (int i = 0; i < pt.npieces; i++) {
        const int nd = pt.dn[i];
        if (nd == 0) continue;
        SGroup gr;
        for (int j = 0; j < nd; j++) gr.pos.push_back({(int)pt.dshift[dp + j] - K, pt.dcls[dp + j]});
        gr.piece = i; gr.pfirst = gr.plast = true; gr.counted = false;
        groups.push_back(gr);
        dp += nd;
    }
    wild = false;
    for (int i = 0; i < pt.npieces; i++) wild = wild || pt.dwild[i];
}

static void gen_stream(Emit &o, const ApxPat *pats, int npat)
{
    std::vector<std::vector<SGroup>> groups((size_t)npat);
    std::vector<char> wild((size_t)npat, 0);
    int hist[32];                              // words of history a class plane needs (0 = class not used)
    memset(hist, 0, sizeof hist);
    bool used[32];
    memset(used, 0, sizeof used);
    for (int p = 0; p < npat; p++) {
        bool w = false;
        collect_groups(pats[p], groups[(size_t)p], w);
        wild[(size_t)p] = w;
        const int m = pats[p].m, K = pats[p].k;
        for (const SGroup &g : groups[(size_t)p])
            for (const Pos &q : g.pos) {
                const int sh = (m - 1 - q.s) + (g.counted ? 0 : K);
                used[q.cls & 31] = true;
                hist[q.cls & 31] = std::max(hist[q.cls & 31], sh / 32 + (sh % 32 ? 1 : 0));
            }
    }
    // ---- state ----
    std::vector<std::string> fields;
    for (int c = 0; c < 32; c++)
        for (int k = 1; k <= hist[c]; k++) fields.push_back("pc" + std::to_string(c) + "_" + std::to_string(k));
    for (int p = 0; p < npat; p++) {
        const int win = pats[p].win;
        for (size_t gi = 0; gi < groups[(size_t)p].size(); gi++) {
            const SGroup &g = groups[(size_t)p][gi];
            if (!g.counted) continue;
            fields.push_back("g" + std::to_string(p) + "_" + std::to_string(gi) + "_1");
            if (win >= 5) fields.push_back("w" + std::to_string(p) + "_" + std::to_string(gi) + "_1");
        }
    }
    o.f("struct DenseState {\n    unsigned dummy;\n");
    for (const std::string &f : fields) o.f("    unsigned %s;\n", f.c_str());
    o.f("};\n__device__ __forceinline__ void dense_init(DenseState &s)\n{\n    s.dummy = 0;\n");
    for (const std::string &f : fields) o.f("    s.%s = 0;\n", f.c_str());
    o.f("}\n\n");
    // ---- step ----
    o.f("__device__ __forceinline__ void dense_step(DenseState &s, const unsigned h, const unsigned l, const unsigned x, unsigned &U0%s)\n{\n", npat > 1 ? ", unsigned &U1" : "");
    for (int c = 0; c < 32; c++) {
        if (!used[c]) continue;
        if (c == 16) o.f("    const unsigned pc16 = x;\n");
        else o.f("    const unsigned pc%d = lop3_<0x%02x>(h, l, x);\n", c, lut_of((unsigned)c));
    }
    for (int p = 0; p < npat; p++) {
        const ApxPat &pt = pats[p];
        const int m = pt.m, K = pt.k, win = pt.win;
        int ncounted = 0;
        for (const SGroup &g : groups[(size_t)p]) ncounted += g.counted ? 1 : 0;
        o.f("    // ---- pattern %d: m = %d, K = %d, window %d, %d counted chunk(s)\n", p, m, K, win, ncounted);
        int seen = 0;
        bool u_set = false;
        for (size_t gi = 0; gi < groups[(size_t)p].size(); gi++) {
            const SGroup &g = groups[(size_t)p][gi];
            const std::string gn = "g" + std::to_string(p) + "_" + std::to_string(gi);
            const std::string wn = "w" + std::to_string(p) + "_" + std::to_string(gi);
            const std::string dn = "d" + std::to_string(p) + "_" + std::to_string(gi);
            std::string expr;
            for (const Pos &q : g.pos) {
                const int sh = (m - 1 - q.s) + (g.counted ? 0 : K);
                if (!expr.empty()) expr += " & ";
                expr += back("pc" + std::to_string(q.cls & 31), sh);
            }
            o.f("    const unsigned %s = %s;\n", gn.c_str(), expr.c_str());
            if (g.piece >= 0) {
                const std::string term = g.counted ? back(gn, K) : gn;
                const std::string mn = "m" + std::to_string(p) + "_" + std::to_string(g.piece);
                if (g.pfirst) o.f("    unsigned %s = %s;\n", mn.c_str(), term.c_str());
                else o.f("    %s &= %s;\n", mn.c_str(), term.c_str());
                if (g.plast) {
                    if (!u_set) o.f("    unsigned u%d = %s;\n", p, mn.c_str());
                    else o.f("    u%d |= %s;\n", p, mn.c_str());
                    u_set = true;
                }
            }
            if (g.counted) {
                if (win == 1) o.f("    const unsigned %s = %s;\n", dn.c_str(), back(gn, K).c_str());
                else if (win == 3) o.f("    const unsigned %s = %s | %s | %s;\n", dn.c_str(), gn.c_str(), back(gn, 1).c_str(), back(gn, 2).c_str());
                else {
                    o.f("    const unsigned %s = %s | %s | %s;\n", wn.c_str(), gn.c_str(), back(gn, 1).c_str(), back(gn, 2).c_str());
                    if (win == 5) o.f("    const unsigned %s = %s | %s;\n", dn.c_str(), wn.c_str(), back(wn, 2).c_str());
                    else o.f("    const unsigned %s = %s | %s | %s;\n", dn.c_str(), wn.c_str(), back(wn, 2).c_str(), back(wn, 4).c_str());
                }
                // saturating bit-sliced count of missing chunks: c<r> = more than r counted chunks missing so far
                for (int r = std::min(seen, K); r >= 1; r--) {
                    if (r == seen) o.f("    unsigned c%d_%d = c%d_%d & ~%s;\n", p, r, p, r - 1, dn.c_str());
                    else o.f("    c%d_%d |= c%d_%d & ~%s;\n", p, r, p, r - 1, dn.c_str());
                }
                if (seen == 0) o.f("    unsigned c%d_0 = ~%s;\n", p, dn.c_str());
                else o.f("    c%d_0 |= ~%s;\n", p, dn.c_str());
                seen++;
            }
        }
        if (wild[(size_t)p] || !u_set) o.f("    %su%d = 0xffffffffu;\n", u_set ? "" : "unsigned ", p);
        if (ncounted > K) o.f("    U%d = u%d & ~c%d_%d;\n", p, p, p, K);
        else o.f("    U%d = u%d;\n", p, p);
        // slide the chunk planes / dilation windows
        for (size_t gi = 0; gi < groups[(size_t)p].size(); gi++) {
            const SGroup &g = groups[(size_t)p][gi];
            if (!g.counted) continue;
            o.f("    s.g%d_%zu_1 = g%d_%zu;\n", p, gi, p, gi);
            if (win >= 5) o.f("    s.w%d_%zu_1 = w%d_%zu;\n", p, gi, p, gi);
        }
    }
    for (int c = 0; c < 32; c++) {
        for (int k = hist[c]; k >= 2; k--) o.f("    s.pc%d_%d = s.pc%d_%d;\n", c, k, c, k - 1);
        if (hist[c] >= 1) o.f("    s.pc%d_1 = pc%d;\n", c, c);
    }
    o.f("}\n\n");
}

}  // namespace

int apx_jit_wpl()
{
    if (const char *e = getenv("PM_JIT_WPL")) { const int v = atoi(e); if (v == 4 || v == 8) return v; }      // experiments
    return 8;
}

int apx_jit_ctas()
{
    if (const char *e = getenv("PM_JIT_CTAS")) { const int v = atoi(e); if (v >= 1 && v <= 8) return v; }   // experiments
    return apx_jit_wpl() == 8 ? 4 : 3;
}

static int env_int(const char *name, int dflt, int lo, int hi)
{
    if (const char *e = getenv(name)) { const int v = atoi(e); if (v >= lo && v <= hi) return v; }
    return dflt;
}

ApxJitShape apx_jit_shape(bool force_stream)
{
    ApxJitShape sh;
    sh.stream = force_stream ? 1 : env_int("PM_JIT_STREAM", 1, 0, 1);
    if (sh.stream) {
        sh.w = env_int("PM_JIT_W", 9, 1, 63) | 1;           // odd: conflict-free 32-bit shared-memory reads
        sh.warps = env_int("PM_JIT_WARPS", 8, 1, 16);
        sh.stages = env_int("PM_JIT_STAGES", 2, 2, 4);
        sh.ctas = env_int("PM_JIT_CTAS", 3, 1, 16);
        sh.tile_words = sh.warps * 32 * sh.w;
        sh.smem = (size_t)sh.stages * 3 * (sh.tile_words + 4) * 4 + 2 * sh.stages * 8;
    } else {
        sh.w = apx_jit_wpl();
        sh.warps = 32 / sh.w;
        sh.stages = 3;
        sh.ctas = apx_jit_ctas();
        sh.tile_words = 1024;
        sh.smem = (size_t)3 * 3 * (1024 + 4) * 4 + 2 * 3 * 8;
    }
    return sh;
}

std::string apx_generate_prefix(const ApxPat *pats, int npat, bool exact)
{
    const ApxJitShape shp = apx_jit_shape(exact);
    if (shp.stream || exact) {
        Emit o;
        bool wide = false;
        for (int p = 0; p < npat; p++) wide = wide || pats[p].m + 2 * pats[p].k > 32;
        o.f("#define AX_STREAM 1\n#define AX_EXACT %d\n", exact ? 1 : 0);
        o.f("#define AX_K %d\n#define AX_WIDE %d\n#define AX_NPAT %d\n#define AX_CTAS %d\n#define AX_W %d\n#define AX_WARPS %d\n#define AX_STAGES %d\n",
            pats[0].k, wide ? 1 : 0, npat, shp.ctas, shp.w, shp.warps, shp.stages);
        for (int p = 0; p < 2; p++) {
            const ApxPat &pt = pats[p < npat ? p : 0];
            o.f("#define AX%d_M %d\n#define AX%d_L %d\n#define AX%d_NP %d\n#define AX%d_INDEL %d\n#define AX%d_SPAN %d\n", p, pt.m, p, pt.L, p, pt.npieces, p, pt.indel,
                p, pt.m - 1 + 2 * pt.k);
            for (int i = 0; i < 4; i++) o.f("#define AX%d_V%d %d\n", p, i, pt.V[i]);
            static const char *nm[5] = {"MA", "MC", "MT", "MG", "MX"};
            for (int q = 0; q < 5; q++) o.f("#define AX%d_%s 0x%llxULL\n", p, nm[q], pt.posmask[q]);
        }
        o.f("template <int LUT> __device__ __forceinline__ unsigned lop3_(unsigned a, unsigned b, unsigned c)\n"
            "{\n    unsigned r;\n    asm(\"lop3.b32 %%0, %%1, %%2, %%3, %%4;\" : \"=r\"(r) : \"r\"(a), \"r\"(b), \"r\"(c), \"n\"(LUT));\n    return r;\n}\n\n");
        gen_stream(o, pats, npat);
        return o.s;
    }
    Emit o;
    bool wide = false;
    for (int p = 0; p < npat; p++) wide = wide || pats[p].m + 2 * pats[p].k > 32;
    o.f("#define AX_K %d\n#define AX_WIDE %d\n#define AX_NPAT %d\n#define AX_CTAS %d\n#define AX_WPL %d\n", pats[0].k, wide ? 1 : 0, npat, apx_jit_ctas(), apx_jit_wpl());
    for (int p = 0; p < npat; p++) {
        const ApxPat &pt = pats[p];
        o.f("#define AX%d_M %d\n#define AX%d_L %d\n#define AX%d_NP %d\n#define AX%d_INDEL %d\n", p, pt.m, p, pt.L, p, pt.npieces, p, pt.indel);
        for (int i = 0; i < 4; i++) o.f("#define AX%d_V%d %d\n", p, i, pt.V[i]);
        static const char *nm[5] = {"MA", "MC", "MT", "MG", "MX"};
        for (int q = 0; q < 5; q++) o.f("#define AX%d_%s 0x%llxULL\n", p, nm[q], pt.posmask[q]);
    }
    o.f("template <int LUT> __device__ __forceinline__ unsigned lop3_(unsigned a, unsigned b, unsigned c)\n"
        "{\n    unsigned r;\n    asm(\"lop3.b32 %%0, %%1, %%2, %%3, %%4;\" : \"=r\"(r) : \"r\"(a), \"r\"(b), \"r\"(c), \"n\"(LUT));\n    return r;\n}\n\n");
    for (int p = 0; p < npat; p++) gen_dense(o, pats[p], p, apx_jit_wpl());
    return o.s;
}

std::string apx_full_source(const std::string &prefix) { return prefix + (prefix.compare(0, 17, "#define AX_STREAM") == 0 ? kStreamBody : kKernelBody); }

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
