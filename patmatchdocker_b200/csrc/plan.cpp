// plan.cpp -- see plan.hpp.  Product code (host side of the C ABI); no oracle, no reference.
#include "plan.hpp"
#include "../../include/patmatch_b200.h"
#include <cctype>
#include <cstring>
#include <cstdlib>

namespace pm {

// Text-frequency model the reference's cost functions use (letterProb @621120):
// occurrences per million bytes of English text, indexed by byte value.
static const int kBytePpm[256] = {
         0,      0,      0,      0,      0,      0,      0,      0,      0,    344,  20793,      0,      0,      0,      0,      0,
         0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,
    146588,     43,    460,    398,  11430,   3034,   1013,   1707,   4156,   4162,    506,    998,   8441,   3342,   9616,    903,
      2255,   4002,   2441,   1222,    937,   1102,    874,    828,    970,   1810,    679,    168,    190,   1562,    143,     35,
        86,   2093,   1334,   1530,    818,    981,   1181,    571,    754,   1534,    156,    228,    656,   1308,    922,   1299,
      1202,    261,    689,   1809,   3403,    669,    340,    961,    158,    390,    234,    847,  15840,    846,   1258,   1695,
       715,  53857,  11376,  27900,  21596,  94887,  15707,  13246,  30408,  54368,    933,   3729,  28211,  20693,  48064,  47054,
     18812,   2436,  44806,  48118,  65831,  16154,   6572,   8692,   5656,   7099,   1124,   8146,    445,   8146,   1852,      0,
         0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,
         0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,
         1,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,
         0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,
         0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,
         0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,      0,
         0,     35,      0,      0,      0,      0,      0,      0,      0,     19,      0,      0,      0,     32,      0,      0,
         0,      9,      0,     40,      0,      0,      0,      0,      0,      0,     27,      0,      0,      0,      0,      0,
};

// -k <n>[idst]: trailing letters select the operations, none = all (main @400fe3-4010c6).
int parse_kopt(const char *kopt, Options &o, std::string &err)
{
    std::string s = kopt ? kopt : "";
    size_t nd = 0;
    while (nd < s.size() && isdigit((unsigned char)s[nd])) nd++;
    std::string flags = s.substr(nd);
    if (nd == 0) { err = "bad -k argument '" + s + "'"; return PM_ERR_SYNTAX; }
    o.k = atoi(s.substr(0, nd).c_str());
    if (flags.empty()) { err = "-k without [ids] enables transpositions, which PatMatch never requests"; return PM_ERR_UNSUPPORTED; }
    o.ins = o.del = o.subs = false;
    for (char c : flags) {
        if (c == 'i') o.ins = true;
        else if (c == 'd') o.del = true;
        else if (c == 's') o.subs = true;
        else if (c == 't') { err = "transpositions (-k ...t) are not on the PatMatch path"; return PM_ERR_UNSUPPORTED; }
        else { err = "bad -k argument '" + s + "'"; return PM_ERR_SYNTAX; }
    }
    if (o.k > 15) { err = "k > 15"; return PM_ERR_UNSUPPORTED; }
    return PM_OK;
}

static void add_byte(ByteSet &b, unsigned c, bool icase)
{
    b.add(c);
    if (icase && isalpha((int)c)) b.add(isupper((int)c) ? (unsigned)tolower((int)c) : (unsigned)toupper((int)c));
}

int parse_pattern(const char *pattern, bool icase, Pattern &P, std::string &err)
{
    P = Pattern();
    std::string s = pattern ? pattern : "";
    if (!s.empty() && s[0] == '^') { P.start_line = true; s.erase(0, 1); }
    if (!s.empty() && s.back() == '$') { P.end_line = true; s.pop_back(); }
    int depth = 0;
    const size_t n = s.size();
    for (size_t i = 0; i < n;) {
        unsigned c = (unsigned char)s[i++];
        if (c == '(') { depth++; continue; }
        if (c == ')') { if (--depth < 0) { err = "unbalanced ')'"; return PM_ERR_SYNTAX; } continue; }
        if (c == '?' || c == '*' || c == '+') {
            // EXTENDED pattern: an operator on a single position (what PatMatch's X{m,n} / X{m,} become);
            // operators on groups would be nrgrep's REGULAR engine
            if (P.pos.empty() || P.op.back() != OP_NONE || (i >= 2 && s[i - 2] == ')')) {
                err = std::string("operator '") + (char)c + "' on a group or doubled: nrgrep's REGULAR engine (not this path)";
                return PM_ERR_UNSUPPORTED;
            }
            P.op.back() = c == '?' ? OP_OPT : c == '*' ? OP_STAR : OP_PLUS;
            continue;
        }
        if (c == '|' || c == '\\' || c == '#') {
            err = std::string("operator '") + (char)c + "' selects nrgrep's REGULAR engine (not this path)";
            return PM_ERR_UNSUPPORTED;
        }
        ByteSet b;
        if (c == '.') b.fill();
        else if (c == '[') {
            bool neg = false, first = true;
            if (i < n && s[i] == '^') { neg = true; i++; }
            while (i < n && (s[i] != ']' || first)) {
                unsigned a = (unsigned char)s[i];
                if (a == '\\') { err = "escape inside class"; return PM_ERR_UNSUPPORTED; }
                first = false;
                if (i + 2 < n && s[i + 1] == '-' && s[i + 2] != ']') {
                    unsigned z = (unsigned char)s[i + 2];
                    for (unsigned x = a; x <= z; x++) add_byte(b, x, icase);
                    i += 3;
                } else { add_byte(b, a, icase); i++; }
            }
            if (i >= n) { err = "unterminated class"; return PM_ERR_SYNTAX; }
            i++;
            if (neg) b.invert();
        } else add_byte(b, c, icase);
        P.pos.push_back(b);
        P.op.push_back(OP_NONE);
    }
    if (depth != 0) { err = "unbalanced '('"; return PM_ERR_SYNTAX; }
    if (P.pos.empty()) { err = "empty pattern"; return PM_ERR_SYNTAX; }
    if (P.extended()) {
        P.had_ops = true;
        // The reference's parser rewrites operator positions at the two ends of the pattern (they cannot change
        // whether a line matches; the reported coordinates then follow the rewritten pattern).  Rules observed on the
        // binary and pinned by oracle/ref/difftest_ext -e:
        //   front: ONE rewrite -- an optional first position ('?', '*') is dropped, a '+' one loses its operator;
        //   back : every trailing optional position is dropped; if none was, a trailing '+' loses its operator.
        size_t lo = 0, hi = P.pos.size();
        if (P.op[0] == OP_OPT || P.op[0] == OP_STAR) lo = 1;
        else if (P.op[0] == OP_PLUS) P.op[0] = OP_NONE;
        bool dropped = false;
        while (hi > lo && (P.op[hi - 1] == OP_OPT || P.op[hi - 1] == OP_STAR)) { hi--; dropped = true; }
        if (!dropped && hi > lo && hi - 1 != 0 && P.op[hi - 1] == OP_PLUS) P.op[hi - 1] = OP_NONE;
        if (hi <= lo) { err = "nothing is left of the pattern once its optional ends are dropped"; return PM_ERR_SYNTAX; }
        P.pos.erase(P.pos.begin() + (long)hi, P.pos.end()); P.op.erase(P.op.begin() + (long)hi, P.op.end());
        P.pos.erase(P.pos.begin(), P.pos.begin() + (long)lo); P.op.erase(P.op.begin(), P.op.begin() + (long)lo);
    }
    return PM_OK;
}

static std::vector<double> class_probs(const Pattern &P)
{
    std::vector<double> prob(P.m());
    for (int i = 0; i < P.m(); i++) {
        double p = 0.0;
        for (unsigned c = 0; c < 256; c++) if (P.pos[i].has(c)) p += (double)kBytePpm[c] / 1000000.0;
        prob[i] = p;
    }
    return prob;
}

// prefix-product table pp[i][l] = prob[i] * pp[i+1][l-1] (probability that P[i..i+l) matches a text window)
static std::vector<double> prefix_products(const std::vector<double> &prob, int m, int lmax)
{
    const int S = lmax + 1;
    std::vector<double> pp((size_t)(m + 1) * S, 0.0);
    pp[(size_t)m * S] = 1.0;
    for (int i = m - 1; i >= 0; i--) {
        pp[(size_t)i * S] = 1.0;
        for (int l = 1; l <= lmax; l++) pp[(size_t)i * S + l] = pp[(size_t)(i + 1) * S + (l - 1)] * prob[i];
    }
    return pp;
}

// simpleFindBest @416a10 -- cheapest BNDM window [beg,end) for a k-error search.
static double find_best_window(const Pattern &P, int k, int &flag, int &beg, int &end)
{
    const int m = P.m(), S = m + 1, K1 = k + 1;
    std::vector<double> pp = prefix_products(class_probs(P), m, m);
    std::vector<double> mprob(m);
    std::vector<int> last(m);
    beg = end = 0;
    const double dK1 = (double)K1;
    double best = 0.8;
    for (int i = 0; i < m; i++) {
        for (int x = 0; x < m; x++) { mprob[x] = 0.0; last[x] = i - 1 + x; }
        int j = K1 + i;
        if (m < j) continue;
        int len = j - i;
        if ((unsigned)len > 64u) continue;
        int lk = len - k;
        for (;;) {
            const int lk1 = lk + 1;
            const double dlk1 = (double)lk1;
            double sum = dK1;
            if (len > 0 && dK1 < dlk1) {
                const double dlk = (double)lk;
                const double c0 = dK1 / ((dlk - dK1) + 1.0);
                if (!(c0 >= best)) {
                    for (int l = 1;; ) {
                        int e = last[l - 1] + 1;
                        double v = mprob[l - 1];
                        for (int row = e - l + 1; e <= j; e++, row++) {
                            const double a = 1.0 - pp[(size_t)row * S + l];
                            const double b = 1.0 - v;
                            v = 1.0 - b * a;
                            mprob[l - 1] = v;
                        }
                        last[l - 1] = j;
                        sum += v;
                        if (l + 1 > len || sum >= dlk1) break;
                        const double c = sum / ((dlk - sum) + 1.0);
                        l++;
                        if (!(c < best)) break;
                    }
                }
            }
            if (dlk1 > sum) {
                const double c = sum / (((double)lk - sum) + 1.0);
                if (best > c) { best = c; beg = i; end = j; }
            }
            if (m < ++j) break;
            len++;
            lk = lk1;
            if ((unsigned)(j - i) > 64u) break;
        }
    }
    if (end - beg <= K1) beg = end = 0;
    flag = end != 0;
    if (!flag) end = m >= 65 ? 64 : m;
    return best < 0.8 ? best : 1.0;
}

// extendedFindBest @411fe0 for K errors (the EXTENDED engine only runs with K = 0 here): cost of scanning for
// every sub-pattern [i, pos] from tables of the probability that a text window of l bytes matches a factor ending
// at `a` and starting at `d` (T) and that any such factor start survives (U).  Rows are filled lazily per end
// position, in the same order and with the same doubles as the reference, because ties between sub-patterns are
// decided by strict comparisons.  Returns the best cost; wlen = mandatory positions of the chosen sub-pattern
// (0: no sub-pattern beats 0.7 -- forward scan for the whole pattern).
static double find_best_extended(const Pattern &P, int K, int &beg, int &end, int &wlen)
{
    const int m = P.m(), N1 = m + 1, NN = m * N1;
    std::vector<double> prob((size_t)m), prob2((size_t)m);
    for (int j = 0; j < m; j++) {
        double p = 0.0, q = 0.0;
        for (unsigned c = 0; c < 256; c++)
            if (P.pos[j].has(c)) { p += (double)kBytePpm[c] / 1000000.0; if (P.repeats(j)) q += (double)kBytePpm[c] / 1000000.0; }
        prob[j] = p; prob2[j] = q;
    }
    std::vector<double> T((size_t)NN * N1, 0.0), U((size_t)NN * N1, 0.0);
    std::vector<int> last((size_t)m, 0);
    auto ix = [&](int d, int a, int l) { return (size_t)d * NN + (size_t)a * N1 + (size_t)l; };
    for (int a = 0; a < m; a++) {
        for (int d = 0; d <= a; d++) { U[ix(d, a, 0)] = 1.0; T[ix(d, a, 0)] = 1.0; }
        U[ix(a + 1, a, 0)] = 0.0; T[ix(a + 1, a, 0)] = 0.0;
    }
    double best = 0.7;
    beg = end = wlen = 0;
    const int K2 = 2 * K;
    const double dK1 = (double)K + 1.0;
    for (int i = 0; i < m; i++) {
        int count = 0;                                  // mandatory positions in [i, pos]
        for (int pos = i; pos < m; pos++) {
            if ((unsigned)(pos - i + 1) > 64u) continue;
            if (P.optional(pos)) { if (K2 >= count) continue; }
            else { count++; if (count <= K2) continue; }
            double sum = dK1, dlk1;
            int lk;
            if (count > 0) {
                lk = count - K;
                dlk1 = (double)(lk + 1);
                if (!(dK1 >= dlk1)) {
                    const double dlk = (double)lk;
                    const double c0 = dK1 / ((dlk - dK1) + 1.0);
                    if (!(c0 >= best)) {
                        for (int l = 1;;) {
                            if (last[pos] < l) {
                                U[ix(pos + 1, pos, l)] = 0.0; T[ix(pos + 1, pos, l)] = 0.0;
                                for (int q = pos; q >= 0; q--) {
                                    double s1 = prob[q] * T[ix(q + 1, pos, l - 1)];
                                    const double s0 = prob2[q] * T[ix(q, pos, l - 1)];
                                    s1 = s1 + s0;
                                    const double s = P.optional(q) ? T[ix(q + 1, pos, l)] + s1 : 0.0 + s1;
                                    double one_minus;
                                    if (s > 1.0) { T[ix(q, pos, l)] = 1.0; one_minus = 0.0; }
                                    else { T[ix(q, pos, l)] = s; one_minus = 1.0 - s; }
                                    U[ix(q, pos, l)] = 1.0 - (1.0 - U[ix(q + 1, pos, l)]) * one_minus;
                                }
                                last[pos] = l;
                            }
                            sum += U[ix(i, pos, l)];
                            l++;
                            if (l > count) break;
                            if (sum >= dlk1) break;
                            const double c = sum / ((dlk - sum) + 1.0);
                            if (!(c < best)) break;
                        }
                    }
                }
            } else { lk = -K; dlk1 = (double)(1 - K); }
            if (dlk1 > sum) {
                const double c = sum / (((double)lk - sum) + 1.0);
                if (best > c) { best = c; beg = i; end = pos + 1; wlen = count; }
            }
        }
    }
    if (wlen > 0) {
        while (beg < end && P.optional(beg)) beg++;
        while (beg < end && P.optional(end - 1)) end--;
        if (beg == end) wlen = 0;
    }
    if (wlen == 0) {
        end = m > 64 ? 64 : m;
        while (P.optional(end - 1)) end--;
        best = 1.0;
    }
    return best;
}

// closure masks of one verification walk (extendedLoadVerif @412c60); element u = pattern position from + u*step
static void closure_masks(const Pattern &P, int from, int len, int step, uint64_t &I, uint64_t &F, uint64_t &A, uint64_t &init)
{
    I = F = A = init = 0;
    bool flag = false;
    for (int u = 0; u < len; u++) {
        if (!P.optional(from + u * step)) continue;
        if (u > 0) {
            if ((F >> (u - 1)) & 1ULL) {
                F &= ~(1ULL << (u - 1)); F |= 1ULL << u;
                if (flag) A |= 1ULL << u; else init |= 1ULL << u;
            } else { I |= 1ULL << (u - 1); F |= 1ULL << u; flag = true; A |= 1ULL << u; }
        } else { if (flag) A |= 1ULL << u; else init |= 1ULL << u; }
    }
}

static int make_plan_extended(const Pattern &P, const Options &o, Plan &plan, std::string &err)
{
    const int m = P.m();
    if (o.k != 0) { err = "EXTENDED pattern with errors (nrgrep's eextended engine): not supported yet"; return PM_ERR_UNSUPPORTED; }
    if (m > 64) { err = "EXTENDED pattern longer than 64 positions"; return PM_ERR_UNSUPPORTED; }
    for (int j = 0; j < m; j++)
        if (P.repeats(j)) plan.ext_repeats = 1;
    plan.fb_cost = find_best_extended(P, 0, plan.ext_beg, plan.ext_end, plan.ext_wlen);
    if (plan.ext_wlen > 0) { plan.type = EXT_BEG; plan.anchor = plan.ext_beg; }
    else { plan.type = EXT_END; plan.anchor = plan.ext_end; }
    if (plan.type == EXT_END && P.optional(0)) {
        // Only possible after the parser's rewrite, e.g. (A?A?C) -> A?C.  The binary's forward scan (extendedScan
        // @4116f0, wlen <= 0) applies the closure of the optional runs after each byte and starts from an empty state,
        // so an occurrence that has to skip the leading optional run on the very first byte of a scan range or record
        // is not seen.  The verification re-runs that automaton for candidates whose match starts at such a byte;
        // closure masks of the scan, forward order over P[0, anchor):
        plan.ext_lead_opt = 1;
        for (int u = 0; u < plan.anchor; u++) {
            if (!P.optional(u)) continue;
            if (u > 0 && ((plan.FS >> (u - 1)) & 1ULL)) { plan.FS &= ~(1ULL << (u - 1)); plan.FS |= 1ULL << u; plan.AS |= 1ULL << u; }
            else { if (u > 0) plan.IS |= 1ULL << (u - 1); plan.FS |= 1ULL << u; plan.AS |= 1ULL << u; }
        }
    }
    // plain positions next to the anchor: present in every match at a fixed offset, so an exact scan finds all anchors
    // (a '+' position next to the run still pins one byte -- its first occurrence on the right of the anchor, its last
    // on the left -- but nothing beyond it)
    int lo = plan.anchor, hi = plan.anchor;
    while (lo > 0 && P.op[lo - 1] == OP_NONE) lo--;
    if (lo > 0 && P.op[lo - 1] == OP_PLUS) lo--;
    while (hi < m && P.op[hi] == OP_NONE) hi++;
    if (hi < m && P.op[hi] == OP_PLUS) hi++;
    plan.win_lo = lo; plan.win_hi = hi;
    plan.L = hi - lo; plan.npieces = 1; plan.V[0] = plan.anchor;      // V[0]: where the verification splits the pattern
    if (plan.L < 1) { err = "internal: empty scan window"; return PM_ERR_UNSUPPORTED; }
    closure_masks(P, plan.anchor - 1, plan.anchor, -1, plan.IL, plan.FL, plan.AL, plan.initL);
    closure_masks(P, plan.anchor, m - plan.anchor, +1, plan.IR, plan.FR, plan.AR, plan.initR);
    return PM_OK;
}

static int g_compat_deployed = 0;
void set_compat_deployed(int on) { g_compat_deployed = on ? 1 : 0; }
int compat_deployed() { return g_compat_deployed; }

int make_plan(const Pattern &P, const Options &o, Plan &plan, std::string &err)
{
    plan = Plan();
    const int m = P.m(), k = o.k;
    plan.m = m; plan.k = k; plan.ins = o.ins; plan.del = o.del; plan.subs = o.subs;
    if (P.extended()) return make_plan_extended(P, o, plan, err);
    if (P.had_ops && k != 0) { err = "pattern with ? * + and errors (nrgrep's eextended engine): not supported yet"; return PM_ERR_UNSUPPORTED; }
    if (k == 0) { plan.type = SIMPLE; plan.L = m; plan.npieces = 1; plan.V[0] = 0; return PM_OK; }
    if (k >= m) { err = "k >= pattern length"; return PM_ERR_UNSUPPORTED; }
    const int K1 = k + 1, K2 = k + 2;
    plan.fb_cost = find_best_window(P, k, plan.fb_flag, plan.fb_beg, plan.fb_end);

    const int lmax = (m > 64 ? 64 : m) / K1;
    std::vector<double> pp = prefix_products(class_probs(P), m, lmax);
    const int S = lmax + 1;

    // expected bytes inspected per BNDM window for every piece start i and length l (415a02-415ac5);
    // scratch rows start from zero (the defined behaviour of the reference, see DESIGN.md "UB").
    std::vector<double> work((size_t)m * (lmax ? lmax : 1), 0.0);
    std::vector<double> rows((size_t)(lmax + 1) * (lmax ? lmax : 1), 0.0);
    // The reference never writes the cells rows[l][l] and reads them.  Opt-in (pm_set_compat_deployed_glibc): what they
    // hold in the DEPLOYED process -- simpleFindBest's (m+1)^2 table is too large for glibc's tcache once m >= 11, its
    // freed chunk is split from the front by esimplePreproc's allocations (V: a 32-byte chunk, the (m+1) x (lmax+1)
    // table, then these rows), so the rows see the old table's doubles at that offset; rows whose chunk size equals
    // that of the 8m-byte probability array reuse that array; everything else is fresh zero memory.  Traced with an
    // LD_PRELOAD shim and checked against the stock binary on 2 000 random searches (tools/deployed_gap.py).
    std::vector<double> stale((size_t)(lmax > 0 ? lmax : 1), 0.0);
    if (g_compat_deployed && lmax > 1) {
        auto cs = [](size_t n) { const size_t c = (n + 8 + 15) & ~(size_t)15; return c < 32 ? (size_t)32 : c; };
        const size_t szpp0 = 8 * (size_t)(m + 1) * (size_t)(m + 1), szA = 8 * (size_t)(lmax + 1) * (size_t)lmax;
        if (cs(szA) == cs(8 * (size_t)m)) {                    // the probability array just freed, from the tcache
            const std::vector<double> prob = class_probs(P);
            for (int l = 1; l < lmax; l++) {
                const size_t idx = (size_t)l * lmax + l;
                stale[(size_t)l] = (idx >= 2 && idx < (size_t)m) ? prob[idx] : 0.0;
            }
        } else if (cs(szA) == cs(4 * (size_t)m)) {
            // simpleFindBest's int array: as doubles, denormals that vanish in 1.0 - x
        } else if (szpp0 <= 1032) {
            if (cs(szA) == cs(szpp0)) {                           // the old table itself, from the tcache
                const std::vector<double> old = prefix_products(class_probs(P), m, m);
                for (int l = 1; l < lmax; l++) {
                    const size_t idx = (size_t)l * lmax + l;
                    stale[(size_t)l] = (idx >= 2 && idx < old.size()) ? old[idx] : 0.0;
                }
            }
        } else {
            const std::vector<double> old = prefix_products(class_probs(P), m, m);          // simpleFindBest's table
            const size_t off = (cs(4 * (size_t)K1) + cs(8 * (size_t)(m + 1) * (size_t)S)) / 8;
            for (int l = 1; l < lmax; l++) {
                const size_t idx = off + (size_t)l * lmax + l;
                stale[(size_t)l] = idx < old.size() ? old[idx] : 0.0;
            }
        }
    }
    for (int i = 0; i < m && lmax > 0; i++) {
        std::fill(rows.begin(), rows.end(), 0.0);
        for (int l = 1; l < lmax; l++) rows[(size_t)l * lmax + l] = stale[(size_t)l];
        double *prev = rows.data();
        for (int l = 1; l <= lmax; l++) {
            double *cur = prev + lmax;
            double sum = 1.0;
            for (int r = 0; r < l; r++) {
                const int row = i + l - 1 - r;
                const double pv = row <= m ? pp[(size_t)row * S + (1 + r)] : 0.0;
                const double a = 1.0 - prev[r], b = 1.0 - pv;
                const double v = 1.0 - b * a;
                cur[r] = v;
                sum += v;
            }
            work[(size_t)i * lmax + (l - 1)] = sum;
            prev = cur;
        }
    }

    double best = 0.97;
    int bestL = 0, bestV[16] = {0};
    if (lmax > 1 && !(1.0 / (double)lmax > 0.97)) {
        std::vector<double> cost((size_t)(m + 1) * K2, 0.0);
        std::vector<int> choice((size_t)(m + 1) * K2, 0);
        for (int Lc = lmax;;) {
            for (int i = 0; i <= m; i++) cost[(size_t)i * K2] = 0.0;
            for (int j = 1; j <= K1; j++) cost[(size_t)m * K2 + j] = 1.0;
            const double dL = (double)Lc, dL1 = (double)(Lc + 1);
            for (int j = 1; j <= K1; j++) {
                const int istart = m - j * Lc;
                for (int i = istart; i >= 0; i--) {
                    const double x = work[(size_t)i * lmax + (Lc - 1)];
                    double c = 0.0;
                    if (dL1 > x) { c = x / ((dL - x) + 1.0); c = c > 1.0 ? 0.0 : 1.0 - c; }
                    double val = 1.0 - c * (1.0 - cost[(size_t)(i + Lc) * K2 + (j - 1)]);
                    int ch = i;
                    if (i < istart) {
                        const double nx = cost[(size_t)(i + 1) * K2 + j];
                        if (val > nx) { val = nx; ch = choice[(size_t)(i + 1) * K2 + j]; }
                    }
                    cost[(size_t)i * K2 + j] = val;
                    choice[(size_t)i * K2 + j] = ch;
                }
            }
            const double total = cost[K1];
            if (best > total) {
                for (int j = K1, i = 0, n = 0; j >= 1; j--) { const int s = choice[(size_t)i * K2 + j]; bestV[n++] = s; i = s + Lc; }
                best = total; bestL = Lc;
            }
            if (--Lc <= 1) break;
            if (1.0 / (double)Lc > best) break;
        }
    }
    plan.split_cost = best;
    if (0.97 > best && !(best >= (double)K1 * plan.fb_cost) && bestL != 0) {
        plan.type = SPLIT; plan.L = bestL; plan.npieces = K1;
        for (int i = 0; i < K1; i++) plan.V[i] = bestV[i];
        // esimpleScan @41384b tests piece i with (1 << (L-1+i*L)) computed as a 32-bit int:
        // the shift count wraps mod 32 and bit 31 sign-extends.  trig[i] is that mask.
        for (int i = 0; i < K1; i++)
            plan.trig[i] = (uint64_t)(int64_t)(int32_t)(1u << ((bestL - 1 + i * bestL) & 31));
    } else {
        plan.npieces = 1;
        plan.L = plan.fb_end - plan.fb_beg;
        if (plan.fb_flag) { plan.type = BWD; plan.V[0] = plan.fb_beg; }
        else { plan.type = FWD; plan.V[0] = plan.fb_end; }
    }
    return PM_OK;
}

void build_filter(const Pattern &P, const Plan &plan, FilterTables &ft)
{
    memset(&ft, 0, sizeof ft);
    const int L = plan.L, np = (plan.type == SPLIT) ? plan.npieces : 1;
    ft.bits = np * L;
    for (int i = 0; i < np; i++) {
        const int base = (plan.type == SPLIT) ? plan.V[i] : 0;
        ft.init |= 1ULL << (i * L);
        ft.fin |= 1ULL << (i * L + L - 1);
        for (int j = 0; j < L; j++)
            for (unsigned c = 0; c < 256; c++)
                if (P.pos[base + j].has(c)) ft.B[c] |= 1ULL << (i * L + j);
    }
}

void build_verify(const Pattern &P, const Plan &plan, VerifyTables &vt)
{
    const int np = plan.npieces, m = P.m();
    if (plan.type == EXT_BEG || plan.type == EXT_END) {
        // [0,256): elements accepting the byte (walk order, away from the anchor); [256,512): those that may also stay
        // on it ('*' / '+')
        const int a = plan.anchor;
        vt.TL.assign(512, 0);
        vt.TR.assign(512, 0);
        for (unsigned c = 0; c < 256; c++) {
            for (int u = 0; u < a; u++)
                if (P.pos[a - 1 - u].has(c)) { vt.TL[c] |= 1ULL << u; if (P.repeats(a - 1 - u)) vt.TL[256 + c] |= 1ULL << u; }
            for (int u = 0; u < m - a; u++)
                if (P.pos[a + u].has(c)) { vt.TR[c] |= 1ULL << u; if (P.repeats(a + u)) vt.TR[256 + c] |= 1ULL << u; }
        }
        return;
    }
    if (m > 64) {
        // 65..255 positions: four words per byte, [piece][byte][word]
        const int NW = 4;
        vt.TL.assign((size_t)np * 256 * NW, 0);
        vt.TR.assign((size_t)np * 256 * NW, 0);
        for (int i = 0; i < np; i++) {
            const int lb = plan.V[i], rl = m - lb;
            for (unsigned c = 0; c < 256; c++) {
                uint64_t *l = &vt.TL[((size_t)i * 256 + c) * NW], *r = &vt.TR[((size_t)i * 256 + c) * NW];
                for (int j = 0; j < lb; j++) if (P.pos[lb - 1 - j].has(c)) l[j >> 6] |= 1ULL << (j & 63);
                for (int j = 0; j < rl; j++) if (P.pos[lb + j].has(c)) r[j >> 6] |= 1ULL << (j & 63);
            }
        }
        return;
    }
    vt.TL.assign((size_t)np * 256, 0);
    vt.TR.assign((size_t)np * 256, 0);
    for (int i = 0; i < np; i++) {
        const int lb = plan.V[i], rl = m - lb;
        for (unsigned c = 0; c < 256; c++) {
            uint64_t l = 0, r = 0;
            for (int j = 0; j < lb; j++) if (P.pos[lb - 1 - j].has(c)) l |= 1ULL << j;
            for (int j = 0; j < rl; j++) if (P.pos[lb + j].has(c)) r |= 1ULL << j;
            vt.TL[(size_t)i * 256 + c] = l;
            vt.TR[(size_t)i * 256 + c] = r;
        }
    }
}

}  // namespace pm
