/* nrgrep_oracle.c -- TEST INFRASTRUCTURE ONLY; see nrgrep_oracle.h.
 *
 * Restatement of nrgrep_coords (reference: /root/reference/www/bin/nrgrep_coords,
 * invoked by www/FlaskApp/FlaskApp/patmatch.py:733-743).  The binary has no
 * sources; each routine cites the unstripped symbol it restates.
 */
#include "nrgrep_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

/* ------------------------------------------------------------------------- */
/* letterProb @621120 (.data): per-byte text frequencies, stored in the       */
/* binary as the doubles n/1e6 for the integers below (verified bit-exact).   */
static const int letter_ppm[256] = {
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

static inline int cls_has(const nro_pattern *P, int j, unsigned c)
{
    return (int)((P->cls[j][c >> 6] >> (c & 63)) & 1);
}
static inline void cls_add(nro_pattern *P, int j, unsigned c, int icase)
{
    P->cls[j][c >> 6] |= 1ULL << (c & 63);
    if (icase && isalpha((int)c)) {
        unsigned d = isupper((int)c) ? (unsigned)tolower((int)c) : (unsigned)toupper((int)c);
        P->cls[j][d >> 6] |= 1ULL << (d & 63);
    }
}

/* ------------------------------------------------------------------------- */
/* Pattern syntax: main @400e00 strips a leading '^' / trailing '$'; parse    */
/* @41ae90 (parseOr/parseConc/parseLiteral/getAclass) builds the tree and     */
/* simplify @41a170 flattens plain groups.  Only the SIMPLE sub-language      */
/* (letters, classes, '.', plain grouping) is on this path; operators make    */
/* the reference take its EXTENDED/REGULAR engines and are reported as        */
/* NRO_ERR_UNSUPPORTED.                                                       */
static int parse_class(const char *s, int *ip, int n, nro_pattern *P, int j, int icase)
{
    int i = *ip;                 /* s[i-1] == '[' */
    int neg = 0;
    if (i < n && s[i] == '^') { neg = 1; i++; }
    memset(P->cls[j], 0, sizeof P->cls[j]);
    int first = 1;
    while (i < n && (s[i] != ']' || first)) {
        unsigned c = (unsigned char)s[i];
        if (c == '\\') return NRO_ERR_UNSUPPORTED;
        first = 0;
        if (i + 2 < n && s[i + 1] == '-' && s[i + 2] != ']') {
            unsigned d = (unsigned char)s[i + 2];
            if (d == '\\') return NRO_ERR_UNSUPPORTED;
            for (unsigned x = c; x <= d; x++) cls_add(P, j, x, icase);
            i += 3;
        } else {
            cls_add(P, j, c, icase);
            i++;
        }
    }
    if (i >= n) return NRO_ERR_SYNTAX;
    i++;                         /* ']' */
    if (neg)
        for (int w = 0; w < 4; w++) P->cls[j][w] = ~P->cls[j][w];
    *ip = i;
    return NRO_OK;
}

int nro_parse(const char *pattern, int icase, nro_pattern *P)
{
    memset(P, 0, sizeof *P);
    int n = (int)strlen(pattern);
    const char *s = pattern;
    if (n > 0 && s[0] == '^') { P->start_line = 1; s++; n--; }
    if (n > 0 && s[n - 1] == '$') { P->end_line = 1; n--; }
    int depth = 0, i = 0, m = 0;
    while (i < n) {
        unsigned c = (unsigned char)s[i++];
        switch (c) {
        case '(': depth++; break;
        case ')':
            if (--depth < 0) return NRO_ERR_SYNTAX;
            /* a group followed by an operator is EXTENDED/REGULAR */
            break;
        case '?': case '*': case '+': case '|': case '\\': case '#':
            return NRO_ERR_UNSUPPORTED;
        case '[': {
            if (m >= NRO_MAXM) return NRO_ERR_TOOLONG;
            int rc = parse_class(s, &i, n, P, m, icase);
            if (rc) return rc;
            m++;
            break;
        }
        case '.':
            if (m >= NRO_MAXM) return NRO_ERR_TOOLONG;
            memset(P->cls[m], 0xff, sizeof P->cls[m]);
            m++;
            break;
        default:
            if (m >= NRO_MAXM) return NRO_ERR_TOOLONG;
            memset(P->cls[m], 0, sizeof P->cls[m]);
            cls_add(P, m, c, icase);
            m++;
        }
    }
    if (depth != 0 || m == 0) return NRO_ERR_SYNTAX;
    P->m = m;
    return NRO_OK;
}

/* ------------------------------------------------------------------------- */
/* prob[i] = sum of letterProb over the class of position i, bytes ascending  */
/* (esimplePreproc @415696-4156f5, simpleFindBest @416a5c-416aab).            */
static void class_probs(const nro_pattern *P, double *prob)
{
    for (int i = 0; i < P->m; i++) {
        double p = 0.0;
        for (unsigned c = 0; c < 256; c++)
            if (cls_has(P, i, c)) p += (double)letter_ppm[c] / 1000000.0;
        prob[i] = p;
    }
}

/* simpleFindBest @416a10: best BNDM sub-pattern [beg,end) under the cost     */
/* model, for a search with k errors; returns its cost (1.0 if >= 0.8).       */
/* Deployed-allocator compatibility (see nro_set_compat): the table simpleFindBest frees is what glibc hands back,
 * unzeroed, when esimplePreproc allocates its scratch rows a moment later. */
static int g_compat_deployed = 0;
static double *g_sfb_pp = NULL;          /* copy of simpleFindBest's table as it is when freed, (m+1) x (m+1) */
static int g_sfb_m = 0;
void nro_set_compat(int deployed_glibc) { g_compat_deployed = deployed_glibc; }

static double simple_find_best(const nro_pattern *P, int k, int *flag, int *beg, int *end)
{
    int m = P->m, S = m + 1;
    double *prob = malloc(sizeof(double) * (size_t)(m ? m : 1));
    class_probs(P, prob);
    double *pp = calloc((size_t)S * (size_t)S, sizeof(double));   /* pp[i*S + l] */
    pp[m * S + 0] = 1.0;
    for (int i = m - 1; i >= 0; i--) {
        pp[i * S + 0] = 1.0;
        for (int l = 1; l <= m; l++) pp[i * S + l] = pp[(i + 1) * S + (l - 1)] * prob[i];
    }
    free(prob);
    *end = 0; *beg = 0;
    double *mprob = malloc(sizeof(double) * (size_t)(m ? m : 1));
    int *last = malloc(sizeof(int) * (size_t)(m ? m : 1));
    const int K1 = k + 1;
    const double dK1 = (double)K1;
    double best = 0.8;
    for (int i = 0; i < m; i++) {
        for (int x = 0; x < m; x++) { mprob[x] = 0.0; last[x] = i - 1 + x; }
        int j = K1 + i;
        if (m < j) continue;
        int len = j - i;
        if ((unsigned)len > 64u) continue;
        int lk = len - k;                      /* r14d */
        for (;;) {
            int lk1 = lk + 1;                  /* r15d */
            double dlk1 = (double)lk1;         /* xmm4 */
            double sum;                        /* xmm3 */
            if (len <= 0 || dK1 >= dlk1) {
                sum = dK1;
            } else {
                double dlk = (double)lk;       /* xmm13 */
                double c0 = dK1 / ((dlk - dK1) + 1.0);
                sum = dK1;
                if (!(c0 >= best)) {
                    int l = 1;                 /* rsi == r9d */
                    for (;;) {
                        int e = last[l - 1];
                        double v = mprob[l - 1];
                        e++;
                        if (e <= j) {
                            int row = e - l + 1;
                            do {
                                double a = 1.0 - pp[row * S + l];
                                double b = 1.0 - v;
                                v = 1.0 - b * a;
                                mprob[l - 1] = v;
                                row++;
                                e++;
                            } while (e != j + 1);
                        }
                        last[l - 1] = j;
                        sum += v;
                        if (l + 1 > len) break;
                        if (sum >= dlk1) break;
                        double c = sum / ((dlk - sum) + 1.0);
                        l++;
                        if (!(c < best)) break;
                    }
                }
            }
            if (dlk1 > sum) {
                double c = sum / (((double)lk - sum) + 1.0);
                if (best > c) { best = c; *beg = i; *end = j; }
            }
            j = j + 1;
            if (m < j) break;
            len++;
            lk = lk1;
            if ((unsigned)(j - i) > 64u) break;
        }
    }
    if (g_compat_deployed) {
        free(g_sfb_pp);
        g_sfb_pp = malloc(sizeof(double) * (size_t)S * (size_t)S);
        memcpy(g_sfb_pp, pp, sizeof(double) * (size_t)S * (size_t)S);
        g_sfb_m = m;
    }
    free(pp); free(mprob); free(last);
    if (*end - *beg <= K1) { *end = 0; *beg = 0; }
    *flag = (*end != 0);
    if (!*flag) *end = m >= 65 ? 64 : m;
    return best < 0.8 ? best : 1.0;
}

/* esimplePreproc @415540: choose between k+1 exact pieces (SPLIT), backward  */
/* (BWD) and forward (FWD) filtering, and fix the split points V[].           */
int nro_plan_make(const nro_pattern *P, int k, int ins, int del, int subs, nro_plan *plan)
{
    memset(plan, 0, sizeof *plan);
    plan->k = k; plan->ins = ins; plan->del = del; plan->subs = subs;
    const int m = P->m;
    if (k == 0) { plan->type = NRO_SIMPLE; plan->L = m; plan->npieces = 1; return NRO_OK; }
    if (k > NRO_MAXK) return NRO_ERR_TOOLONG;
    const int K1 = k + 1, K2 = k + 2, transp = 0;

    plan->fb_cost = simple_find_best(P, k, &plan->fb_flag, &plan->fb_beg, &plan->fb_end);

    int mm = m - k * transp;
    if (mm > 64) mm = 64;
    const int Lmax = mm / K1;
    const int S = Lmax + 1;

    double *prob = malloc(sizeof(double) * (size_t)m);
    class_probs(P, prob);
    double *pp = calloc((size_t)(m + 1) * (size_t)S, sizeof(double));    /* pp[i*S + l] */
    pp[m * S + 0] = 1.0;
    for (int i = m - 1; i >= 0; i--) {
        pp[i * S + 0] = 1.0;
        for (int l = 1; l <= Lmax; l++) pp[i * S + l] = pp[(i + 1) * S + (l - 1)] * prob[i];
    }
    free(prob);

    /* Bm[i*Lmax + (l-1)]: expected characters inspected per BNDM window for
     * the piece P[i..i+l) (415a02-415ac5).  The reference reads one never-
     * written cell A[l-1][l-1] of its scratch rows here (fresh heap memory,
     * i.e. +0.0, in the CLI process); calloc reproduces that. */
    double *A = calloc((size_t)(Lmax + 1) * (size_t)(Lmax ? Lmax : 1), sizeof(double));
    double *Bm = calloc((size_t)m * (size_t)(Lmax ? Lmax : 1), sizeof(double));
    if (g_compat_deployed && Lmax > 1 && g_sfb_pp && g_sfb_m == m) {
        /* What the never-written cells A[l][l] hold in the deployed CLI process (glibc 2.39 malloc, traced with an
         * LD_PRELOAD shim, tools/deployed_gap.py): simpleFindBest's (m+1)^2 table is too large for the tcache once
         * m >= 11, so its freed chunk is split from the front by esimplePreproc's allocations -- the V array
         * (4(k+1) bytes, a 32-byte chunk), the (m+1) x (Lmax+1) table, then these scratch rows -- and the rows see
         * the old table's doubles at that offset.  Smaller tables go to the tcache and the rows come from fresh
         * (zero) memory; rows whose chunk size equals that of the 8m-byte probability array reuse that array. */
#define NRO_CS(n) ((((size_t)(n) + 8 + 15) & ~(size_t)15) < 32 ? (size_t)32 : (((size_t)(n) + 8 + 15) & ~(size_t)15))
        const size_t szpp0 = 8 * (size_t)(m + 1) * (size_t)(m + 1);
        const size_t szA = 8 * (size_t)(Lmax + 1) * (size_t)Lmax;
        if (NRO_CS(szA) == NRO_CS(8 * (size_t)m)) {
            /* the probability array esimplePreproc has just freed sits in the tcache bin of this size (first 16
             * bytes clobbered by the tcache) */
            double *prob2 = malloc(sizeof(double) * (size_t)m);
            class_probs(P, prob2);
            for (int l = 1; l < Lmax; l++) {
                const size_t idx = (size_t)l * (size_t)Lmax + (size_t)l;
                A[idx] = (idx >= 2 && idx < (size_t)m) ? prob2[idx] : 0.0;
            }
            free(prob2);
        } else if (NRO_CS(szA) == NRO_CS(4 * (size_t)m)) {
            /* simpleFindBest's int array: as doubles, denormals that vanish in 1.0 - x */
        } else if (szpp0 <= 1032) {
            if (NRO_CS(szA) == NRO_CS(szpp0))                      /* the old table itself, from the tcache */
                for (int l = 1; l < Lmax; l++) {
                    const size_t idx = (size_t)l * (size_t)Lmax + (size_t)l;
                    A[idx] = (idx >= 2 && idx < (size_t)(m + 1) * (size_t)(m + 1)) ? g_sfb_pp[idx] : 0.0;
                }
        } else {
            const size_t off = (NRO_CS(4 * (size_t)K1) + NRO_CS(8 * (size_t)(m + 1) * (size_t)S)) / 8;
            for (int l = 1; l < Lmax; l++) {
                const size_t idx = off + (size_t)l * (size_t)Lmax + (size_t)l;
                A[(size_t)l * (size_t)Lmax + (size_t)l] = idx < (size_t)(m + 1) * (size_t)(m + 1) ? g_sfb_pp[idx] : 0.0;
            }
        }
    }
    for (int i = 0; i < m && Lmax > 0; i++) {
        memset(A, 0, sizeof(double) * (size_t)Lmax);
        double *prev = A;
        for (int l = 1; l <= Lmax; l++) {
            double *cur = prev + Lmax;
            double sum = 1.0;
            for (int r = 0; r < l; r++) {
                int row = i + l - 1 - r;
                double pv = row <= m ? pp[row * S + (1 + r)] : 0.0;
                double a = 1.0 - prev[r];
                double b = 1.0 - pv;
                double v = 1.0 - b * a;
                cur[r] = v;
                sum += v;
            }
            Bm[i * Lmax + (l - 1)] = sum;
            prev = cur;
        }
    }

    double best = 0.97;
    int bestL = 0;
    int V[NRO_MAXK + 1];
    if (Lmax > 1 && !(1.0 / (double)Lmax > 0.97)) {
        double *cost = calloc((size_t)(m + 1) * (size_t)K2, sizeof(double));
        int *choice = calloc((size_t)(m + 1) * (size_t)K2, sizeof(int));
        int Lc = Lmax;
        for (;;) {
            for (int i = 0; i <= m; i++) cost[i * K2 + 0] = 0.0;
            for (int j = 1; j <= K1; j++) cost[m * K2 + j] = 1.0;
            const double dL = (double)Lc, dL1 = (double)(Lc + 1);
            for (int j = 1; j <= K1; j++) {
                int istart = m - j * Lc - (j - 1) * transp;
                for (int i = istart; i >= 0; i--) {
                    double x = Bm[i * Lmax + (Lc - 1)];
                    double c;
                    if (dL1 > x) {
                        c = x / ((dL - x) + 1.0);
                        if (c > 1.0) c = 0.0; else c = 1.0 - c;
                    } else c = 0.0;
                    double val = 1.0 - c * (1.0 - cost[(i + Lc + transp) * K2 + (j - 1)]);
                    choice[i * K2 + j] = i;
                    if (i < istart) {
                        double nx = cost[(i + 1) * K2 + j];
                        if (val > nx) { val = nx; choice[i * K2 + j] = choice[(i + 1) * K2 + j]; }
                    }
                    cost[i * K2 + j] = val;
                }
            }
            double total = cost[0 * K2 + K1];
            if (best > total) {
                int i = 0, n = 0;
                for (int j = K1; j >= 1; j--) {
                    int s = choice[i * K2 + j];
                    V[n++] = s;
                    i = s + Lc + transp;
                }
                best = total;
                bestL = Lc;
            }
            Lc--;
            if (Lc <= 1) break;
            if (1.0 / (double)Lc > best) break;
        }
        free(cost); free(choice);
    }
    free(pp); free(A); free(Bm);

    plan->split_cost = best;
    int split = 0;
    if (0.97 > best && !(best >= (double)K1 * plan->fb_cost) && bestL != 0) split = 1;
    if (split) {
        plan->type = NRO_SPLIT; plan->L = bestL; plan->npieces = K1;
        for (int i = 0; i < K1; i++) plan->V[i] = V[i];
    } else {
        plan->npieces = 1;
        plan->L = plan->fb_end - plan->fb_beg;
        if (plan->fb_flag) { plan->type = NRO_BWD; plan->V[0] = plan->fb_beg; }
        else { plan->type = NRO_FWD; plan->V[0] = plan->fb_end; }
    }
    return NRO_OK;
}

/* ------------------------------------------------------------------------- */
/* recCheckLeftContext @402170 / recCheckRightContext @4021e0 with           */
/* OptWholeWord = OptWholeRecord = 0 (patmatch.py never passes -w/-x).        */
static inline int leftctx(const nro_pattern *P, const uint8_t *t, int64_t ptr, int64_t rbeg)
{
    if (P->start_line && ptr > rbeg && t[ptr - 1] != '\n') return 0;
    return 1;
}
static inline int rightctx(const nro_pattern *P, const uint8_t *t, int64_t ptr, int64_t rend)
{
    if (P->end_line && ptr < rend && t[ptr] != '\n') return 0;
    return 1;
}

/* multi-word state vectors (createMask @41b430) */
typedef struct { uint64_t w[NRO_WORDS]; } mask_t;

static inline void m_lowbits(mask_t *r, int e)          /* bits [0,e) set */
{
    for (int w = 0; w < NRO_WORDS; w++) {
        int lo = w * 64;
        if (e >= lo + 64) r->w[w] = ~0ULL;
        else if (e > lo) r->w[w] = ~(~0ULL << (e - lo));
        else r->w[w] = 0;
    }
}
static inline mask_t m_shl1(const mask_t *a, uint64_t carry, int nw)
{
    mask_t r; memset(&r, 0, sizeof r);
    for (int w = 0; w < nw; w++) { r.w[w] = (a->w[w] << 1) | carry; carry = a->w[w] >> 63; }
    return r;
}

/* One direction of checkMatch1 @414190: anchored k-row NFA over a pattern part
 * of length plen whose position j (0-based, counted from the anchor) accepts
 * the bytes of P->cls[map(j)].  dir = -1 walks text leftwards from pos
 * (first byte read is t[pos-1]) down to lim = rbeg; dir = +1 walks rightwards
 * from pos (first byte t[pos]) up to lim = rend.
 * Returns 1 and (*ext, *err) on success: *ext = number of text bytes consumed,
 * *err = the error row that matched (only meaningful for the caller's budget
 * when dir = -1).                                                            */
static int nfa_side(const nro_pattern *P, const nro_plan *pl, const uint8_t *t,
                    int dir, int pbase, int plen, int kmax,
                    int64_t pos, int64_t lim, int64_t *ext, int *err)
{
    const int nw = (plen + 63) / 64;
    const uint64_t fin = 1ULL << ((plen - 1) & 63);
    const int fw = nw - 1;
    mask_t R[NRO_MAXK + 1];
    int kb = kmax;
    int64_t best_ext = -1; int best_err = kmax;

    if (kmax < 0) {
        /* 415115 / 41515a: no rows are initialised; the reference then runs the
         * scan with an empty row set.  With kmax < 0 nothing can match. */
        return 0;
    }
    for (int e = 0; e <= kb; e++) {
        if (pl->del) m_lowbits(&R[e], e); else memset(&R[e], 0, sizeof R[e]);
        if ((R[e].w[fw] & fin) &&
            (dir < 0 ? leftctx(P, t, pos, lim) : rightctx(P, t, pos, lim))) {
            best_err = e; kb = e - 1; best_ext = 0;
        }
    }
    int64_t avail = dir < 0 ? pos - lim : lim - pos;
    uint64_t first = 1;
    for (int64_t step = 1; step <= avail; step++) {
        int64_t tp = dir < 0 ? pos - step : pos + step - 1;   /* byte consumed */
        int64_t edge = dir < 0 ? tp : tp + 1;                 /* match boundary after it */
        unsigned c = t[tp];
        /* T[c]: bit j set iff part position j accepts c */
        mask_t T; memset(&T, 0, sizeof T);
        for (int j = 0; j < plen; j++) {
            int pj = dir < 0 ? pbase - 1 - j : pbase + j;
            if (cls_has(P, pj, c)) T.w[j >> 6] |= 1ULL << (j & 63);
        }
        mask_t oldp = R[0];
        mask_t sh = m_shl1(&R[0], first, nw);
        for (int w = 0; w < nw; w++) R[0].w[w] = sh.w[w] & T.w[w];
        mask_t newp = R[0];
        if ((R[0].w[fw] & fin) &&
            (dir < 0 ? leftctx(P, t, edge, lim) : rightctx(P, t, edge, lim))) {
            *ext = step; *err = 0;
            return 1;                                /* 415194 / 414bd3->414f86 */
        }
        for (int e = 1; e <= kb; e++) {
            mask_t x; memset(&x, 0, sizeof x);
            if (pl->del) { mask_t d = m_shl1(&newp, 0, nw); x = d; }
            if (pl->ins) for (int w = 0; w < nw; w++) x.w[w] |= oldp.w[w];
            if (pl->subs) { mask_t s = m_shl1(&oldp, first, nw); for (int w = 0; w < nw; w++) x.w[w] |= s.w[w]; }
            mask_t y = m_shl1(&R[e], first, nw);
            mask_t nw_row;  memset(&nw_row, 0, sizeof nw_row);
            for (int w = 0; w < nw; w++) nw_row.w[w] = (y.w[w] & T.w[w]) | x.w[w];
            oldp = R[e];
            R[e] = nw_row;
            newp = nw_row;
            if ((nw_row.w[fw] & fin) &&
                (dir < 0 ? leftctx(P, t, edge, lim) : rightctx(P, t, edge, lim))) {
                /* 4147fa-41483c / 414f51-414f84: walk down while lower rows also accept */
                int ec = e, ed;
                for (;;) {
                    ed = ec - 1;
                    if (ed == -1) break;
                    if (!(R[ed].w[fw] & fin)) break;
                    ec = ed;
                }
                if (ed == -1) {
                    if (dir > 0) { *ext = step; *err = 0; return 1; }     /* 414f86 */
                    *ext = step; *err = 0; return 1;                       /* 41483c, e = 0 */
                }
                kb = ed; best_err = ec; best_ext = step;                   /* 414da8 / 414fae */
                break;                                                     /* e = ed+1 > kb */
            }
        }
        /* 414de3 / 414fe3: is any state of row kb still alive? */
        int alive = 0;
        for (int w = 0; w < nw; w++) {
            uint64_t v = R[kb].w[w];
            if (w == fw) v &= (fin << 1) - 1;
            if (v) { alive = 1; break; }
        }
        if (!alive) break;
        first = 0;
    }
    if (best_ext < 0) return 0;
    *ext = best_ext; *err = best_err;
    return 1;
}

/* checkMatch1 @414190 (OptTransp = 0).                                       */
static int check_match1(const nro_pattern *P, const nro_plan *pl, int i, const uint8_t *t,
                        int64_t pos, int64_t rbeg, int64_t rend, int64_t *beg, int64_t *end)
{
    const int k = pl->k;
    const int lb = pl->V[i], rl = P->m - pl->V[i];
    int64_t bext; int berr;
    if (lb == 0) {
        /* 4141ef-414236: only insertions may precede an empty left part */
        int e = 0; int64_t ptr = pos;
        if (k < 0) return 0;
        for (;;) {
            if (leftctx(P, t, ptr, rbeg)) break;
            if (ptr == rbeg) return 0;
            if (!pl->ins) return 0;
            ptr--; e++;
            if (e > k) return 0;
        }
        bext = pos - ptr; berr = e;
    } else {
        if (!nfa_side(P, pl, t, -1, lb, lb, k, pos, rbeg, &bext, &berr)) return 0;
    }
    const int kf = k - berr;
    int64_t fext; int ferr;
    if (rl == 0) {
        /* 414eae-414f19 */
        if (kf < 0) return 0;
        int n = 0; int64_t p = pos;
        for (;;) {
            if (rightctx(P, t, p, rend)) break;
            if (p - pos == rend - pos) return 0;
            if (!pl->ins) return 0;
            n++; p++;
            if (kf < n) return 0;
        }
        fext = p - pos;
    } else {
        if (!nfa_side(P, pl, t, +1, lb, rl, kf, pos, rend, &fext, &ferr)) return 0;
    }
    *beg = pos - bext;
    *end = pos + fext;
    return 1;
}

/* esimple checkMatch @4151d0 with recGetRecord @402030 ('\n' records clipped  */
/* to the current scan range [tbeg, tend)).                                    */
int nro_check_match(const nro_pattern *P, const nro_plan *pl, int i, const uint8_t *t,
                    int64_t pos, int64_t tbeg, int64_t tend, int64_t *beg, int64_t *end)
{
    int64_t p = pl->type == NRO_FWD ? pos - 1 : pos;
    int64_t rbeg = tbeg, rend = tend;
    for (int64_t q = p - 1; q >= tbeg; q--) if (t[q] == '\n') { rbeg = q + 1; break; }
    for (int64_t q = p; q < tend; q++) if (t[q] == '\n') { rend = q; break; }
    if (p < rbeg || p >= rend) return 0;
    return check_match1(P, pl, i, t, pos, rbeg, rend, beg, end);
}

/* ------------------------------------------------------------------------- */
/* simple checkMatch @416790 in coordinate mode ([P+0x1028] != 0): the whole  */
/* pattern must lie inside the scan range; no record logic.                   */
static int simple_at(const nro_pattern *P, const uint8_t *t, int64_t s, int64_t tbeg, int64_t tend)
{
    if (s < tbeg || s + P->m > tend) return 0;
    for (int j = 0; j < P->m; j++) if (!cls_has(P, j, t[s + j])) return 0;
    if (P->start_line || P->end_line) {
        if (!leftctx(P, t, s, tbeg)) return 0;
        if (!rightctx(P, t, s + P->m, tend)) return 0;
    }
    return 1;
}

/* first match in [tbeg, tend): simpleScan @416600 / esimpleScan @4136d0       */
static int scan_first(const nro_pattern *P, const nro_plan *pl, const uint8_t *t,
                      int64_t tbeg, int64_t tend, int64_t *beg, int64_t *end)
{
    const int m = P->m;
    switch (pl->type) {
    case NRO_SIMPLE:
        for (int64_t s = tbeg; s + m <= tend; s++)
            if (simple_at(P, t, s, tbeg, tend)) { *beg = s; *end = s + m; return 1; }
        return 0;
    case NRO_SPLIT: {
        /* 413725-4138a2: multi-pattern BNDM over the k+1 superimposed pieces.
         * After a full window the state D holds bit b_j = L-1 + j*L for every
         * piece j that matches the window exactly.  Piece i is then tested with
         * the C expression D & (1 << b_i) evaluated in 32-bit int arithmetic
         * (41384b: shl eax,cl ; cdqe): for b_i >= 32 the shift count wraps
         * modulo 32 and bit 31 sign-extends.  The quirk is part of the
         * reference's observable behaviour and is restated here. */
        const int L = pl->L;
        for (int64_t w = tbeg; w + L <= tend; w++) {
            uint64_t D = 0;
            for (int i = 0; i < pl->npieces; i++) {
                int ok = 1;
                for (int j = 0; j < L && ok; j++) ok = cls_has(P, pl->V[i] + j, t[w + j]);
                if (ok) D |= 1ULL << (L - 1 + i * L);
            }
            if (!D) continue;
            for (int i = 0; i < pl->npieces; i++) {
                uint64_t mask = (uint64_t)(int64_t)(int32_t)(1u << ((L - 1 + i * L) & 31));
                if ((D & mask) && nro_check_match(P, pl, i, t, w, tbeg, tend, beg, end)) return 1;
            }
        }
        return 0;
    }
    case NRO_FWD:
        /* 4138a7-4139e6 / 413d44 / 413f8c: candidate = text position just
         * after an approximate occurrence of the scanned prefix */
        for (int64_t pos = tbeg + 1; pos <= tend; pos++)
            if (nro_check_match(P, pl, 0, t, pos, tbeg, tend, beg, end)) return 1;
        return 0;
    case NRO_BWD:
        /* 4139eb-413d0c / 413e58 / 41404c: candidate = start of a window of
         * L-k bytes */
        for (int64_t w = tbeg; w + (pl->L - pl->k) <= tend; w++)
            if (nro_check_match(P, pl, 0, t, w, tbeg, tend, beg, end)) return 1;
        return 0;
    }
    return 0;
}

/* recSearchFile @402250 over the text of one buffer fill [s, e).                */
static int64_t search_range(const nro_pattern *P, const nro_plan *pl, const uint8_t *t, int64_t s, int64_t e,
                            nro_hit *hits, int64_t cap, int64_t count)
{
    int64_t pos = s;
    for (;;) {
        int64_t b, en;
        if (!scan_first(P, pl, t, pos, e, &b, &en)) break;
        if (count < cap) { hits[count].beg = b; hits[count].end = en; }
        count++;
        if (en == e) break;
        if (en <= pos && b == en) break;          /* zero-length hit: the reference would spin here */
        pos = en;
    }
    return count;
}

/* Buffer fills (bufCreate @41bb60, bufSetFile @41bbc0, bufLoad @41bbf0 and the top of
 * recSearchFile @402298-4024f0).  The buffer holds `bufsize` BYTES (main @401150 stores
 * atoi(-b) unscaled; patmatch.py passes 1600000).  A fill that does not reach EOF is
 * scanned up to and including its last '\n'; the next fill starts AT that '\n'.  A fill
 * without a usable '\n' ("Record longer than buffer size ... has been split") is scanned
 * whole and the next fill starts right after it.  Returns the number of fills written to
 * seg (pairs start,end).                                                               */
int64_t nro_buffer_fills(const uint8_t *t, int64_t n, int64_t bufsize, int64_t *seg, int64_t segcap)
{
    int64_t k = 0, S = 0;
    if (bufsize <= 0) bufsize = n + 1;
    while (n - S > 0) {
        int64_t dsize = n - S < bufsize ? n - S : bufsize;
        int64_t E, next;
        if (dsize < bufsize) { E = S + dsize; next = n; }
        else {
            int64_t p = S + dsize - 1;
            while (p > S && t[p] != '\n') p--;
            if (p > S) { E = p + 1; next = p; }
            else { E = S + dsize; next = S + dsize; }
        }
        if (k < segcap) { seg[2 * k] = S; seg[2 * k + 1] = E; }
        k++;
        S = next;
    }
    return k;
}

int64_t nro_search_buffered(const nro_pattern *P, const nro_plan *pl, const uint8_t *t, int64_t n, int64_t bufsize,
                            nro_hit *hits, int64_t cap)
{
    int64_t count = 0, S = 0;
    if (bufsize <= 0) bufsize = n + 1;
    while (n - S > 0) {
        int64_t seg[2];
        /* one fill at a time, same rule as nro_buffer_fills */
        int64_t dsize = n - S < bufsize ? n - S : bufsize, next;
        if (dsize < bufsize) { seg[0] = S; seg[1] = S + dsize; next = n; }
        else {
            int64_t p = S + dsize - 1;
            while (p > S && t[p] != '\n') p--;
            if (p > S) { seg[0] = S; seg[1] = p + 1; next = p; }
            else { seg[0] = S; seg[1] = S + dsize; next = S + dsize; }
        }
        count = search_range(P, pl, t, seg[0], seg[1], hits, cap, count);
        S = next;
    }
    return count;
}

/* whole file in one buffer fill */
int64_t nro_search(const nro_pattern *P, const nro_plan *pl, const uint8_t *t, int64_t n,
                   nro_hit *hits, int64_t cap)
{
    if (n <= 0) return 0;
    return search_range(P, pl, t, 0, n, hits, cap, 0);
}

/* ========================================================================= */
/* EXTENDED patterns, k = 0 (positions followed by '?', '*' or '+').  PatMatch's */
/* X{m,n} repeats reach the engine in this form (patmatch_to_nrgrep.pl:184-211). */
/* Restated: extendedPreproc @413260 (plan), extendedFindBest @411fe0 (which     */
/* sub-pattern is scanned for, bit-exact doubles), and the language-level        */
/* meaning of extendedScan @4116f0 + checkMatch @411aa0: anchors are tried in    */
/* increasing order and the first one whose verification succeeds is reported;   */
/* the verification takes the SHORTEST extension to the left of the anchor that  */
/* matches the pattern part before it, then the shortest one to the right.       */
/* Not covered (NRO_ERR_UNSUPPORTED): groups, alternation, and patterns whose    */
/* first or last position carries an operator (the parser simplifies those).     */

int nrx_parse(const char *pattern, int icase, nrx_pattern *X)
{
    memset(X, 0, sizeof *X);
    nro_pattern *P = &X->P;
    int n = (int)strlen(pattern);
    const char *s = pattern;
    if (n > 0 && s[0] == '^') { P->start_line = 1; s++; n--; }
    if (n > 0 && s[n - 1] == '$') { P->end_line = 1; n--; }
    /* parentheses that no operator applies to only group a concatenation (patmatch_to_nrgrep.pl wraps the whole
     * pattern in one pair, the reverse complement in two): they are skipped */
    int i = 0, m = 0, nops = 0, depth = 0, after_close = 0;
    while (i < n) {
        unsigned c = (unsigned char)s[i++];
        if (c == '(') { depth++; after_close = 0; continue; }
        if (c == ')') { if (--depth < 0) return NRO_ERR_SYNTAX; after_close = 1; continue; }
        if (c == '|' || c == '\\' || c == '#') return NRO_ERR_UNSUPPORTED;
        if (c == '?' || c == '*' || c == '+') {
            if (after_close) return NRO_ERR_UNSUPPORTED;              /* operator on a group: REGULAR engine */
            if (m == 0 || X->op[m - 1] != NRX_NONE) return NRO_ERR_UNSUPPORTED;
            X->op[m - 1] = c == '?' ? NRX_OPT : c == '*' ? NRX_STAR : NRX_PLUS;
            nops++;
            continue;
        }
        after_close = 0;
        if (m >= 64) return NRO_ERR_TOOLONG;
        if (c == '[') {
            int rc = parse_class(s, &i, n, P, m, icase);
            if (rc) return rc;
        } else if (c == '.') {
            memset(P->cls[m], 0xff, sizeof P->cls[m]);
        } else {
            memset(P->cls[m], 0, sizeof P->cls[m]);
            cls_add(P, m, c, icase);
        }
        m++;
    }
    if (m == 0 || depth != 0) return NRO_ERR_SYNTAX;
    P->m = m;
    if (nops == 0) return NRO_ERR_UNSUPPORTED;                 /* a SIMPLE pattern: not this path */
    /* The reference's parser rewrites operator positions at the two ends of the pattern (they cannot change WHETHER a
     * line matches, which is all grep needs; the coordinates then follow the rewritten pattern).  Observed on the
     * binary (NOTES_extended.md) and pinned by difftest_ext -e:
     *   front: ONE rewrite -- an optional first position ('?', '*') is dropped, a '+' one loses its operator;
     *   back : every trailing optional position is dropped; if none was, a trailing '+' loses its operator. */
    {
        int lo = 0, hi = m;
        if (X->op[0] == NRX_OPT || X->op[0] == NRX_STAR) lo = 1;
        else if (X->op[0] == NRX_PLUS) X->op[0] = NRX_NONE;
        int dropped = 0;
        while (hi > lo && (X->op[hi - 1] == NRX_OPT || X->op[hi - 1] == NRX_STAR)) { hi--; dropped = 1; }
        if (!dropped && hi > lo && X->op[hi - 1] == NRX_PLUS && !(hi - 1 == 0)) X->op[hi - 1] = NRX_NONE;
        if (hi <= lo) return NRO_ERR_UNSUPPORTED;              /* nothing left: the binary refuses the pattern */
        if (lo > 0 || hi < m) {
            memmove(P->cls[0], P->cls[lo], sizeof P->cls[0] * (size_t)(hi - lo));
            memmove(X->op, X->op + lo, (size_t)(hi - lo));
            memset(X->op + (hi - lo), 0, (size_t)(m - (hi - lo)));
            m = hi - lo;
            P->m = m;
        }
        int left = 0;
        for (int j = 0; j < m; j++) left += X->op[j] != NRX_NONE;
        if (!left) return NRX_REWRITTEN_SIMPLE;                /* what remains is a SIMPLE pattern (X->P): simpleScan */
        /* e.g. (A?A?C) -> A?C still begins with an optional position: fine for plans anchored at a sub-pattern's start;
         * nrx_plan_make refuses the forward-scan plan for such patterns (see there) */
    }
    return NRO_OK;
}

static inline int nrx_optional(const nrx_pattern *X, int j) { return X->op[j] == NRX_OPT || X->op[j] == NRX_STAR; }
static inline int nrx_repeat(const nrx_pattern *X, int j) { return X->op[j] == NRX_STAR || X->op[j] == NRX_PLUS; }

/* extendedFindBest @411fe0 with K = 0.  T[d][a][l], U[d][a][l]: d = first position, a = last position, l = symbols  */
/* read; rows are filled lazily per last position (last[a] = highest l done), exactly in the binary's order.        */
static double ext_find_best(const nrx_pattern *X, int K, int *beg, int *end, int *wlen)
{
    const nro_pattern *P = &X->P;
    const int m = P->m, N1 = m + 1, NN = m * N1;
    double *prob = malloc(sizeof(double) * (size_t)m), *prob2 = malloc(sizeof(double) * (size_t)m);
    for (int j = 0; j < m; j++) {
        double p = 0.0, q = 0.0;
        for (unsigned c = 0; c < 256; c++)
            if (cls_has(P, j, c)) { p += (double)letter_ppm[c] / 1000000.0; if (nrx_repeat(X, j)) q += (double)letter_ppm[c] / 1000000.0; }
        prob[j] = p; prob2[j] = q;
    }
    double *T = calloc((size_t)NN * (size_t)N1, sizeof(double)), *U = calloc((size_t)NN * (size_t)N1, sizeof(double));
    int *last = malloc(sizeof(int) * (size_t)m);
#define IX(d, a, l) ((size_t)(d) * (size_t)NN + (size_t)(a) * (size_t)N1 + (size_t)(l))
    for (int a = 0; a < m; a++) {
        last[a] = 0;
        for (int d = 0; d <= a; d++) { U[IX(d, a, 0)] = 1.0; T[IX(d, a, 0)] = 1.0; }
        U[IX(a + 1, a, 0)] = 0.0; T[IX(a + 1, a, 0)] = 0.0;
    }
    double best = 0.7;
    *beg = 0; *end = 0; *wlen = 0;
    const int K2 = 2 * K;
    const double dK1 = (double)K + 1.0;
    for (int i = 0; i < m; i++) {
        int count = 0;
        for (int pos = i; pos < m; pos++) {
            if ((unsigned)(pos - i + 1) > 64u) continue;
            double sum = dK1, dlk1;
            int lk;
            if (nrx_optional(X, pos)) {
                if (K2 >= count) continue;
            } else {
                count++;
                if (count <= K2) continue;
            }
            if (count > 0) {
                lk = count - K;
                dlk1 = (double)(lk + 1);
                if (!(dK1 >= dlk1)) {
                    const double dlk = (double)lk;
                    double c0 = dK1 / ((dlk - dK1) + 1.0);
                    if (!(c0 >= best)) {
                        int l = 1;
                        for (;;) {
                            if (last[pos] < l) {
                                U[IX(pos + 1, pos, l)] = 0.0; T[IX(pos + 1, pos, l)] = 0.0;
                                for (int q = pos; q >= 0; q--) {
                                    double s1 = prob[q] * T[IX(q + 1, pos, l - 1)];
                                    double s0 = prob2[q] * T[IX(q, pos, l - 1)];
                                    s1 = s1 + s0;
                                    double s = nrx_optional(X, q) ? T[IX(q + 1, pos, l)] + s1 : 0.0 + s1;
                                    double one_minus;
                                    if (s > 1.0) { T[IX(q, pos, l)] = 1.0; one_minus = 0.0; }
                                    else { T[IX(q, pos, l)] = s; one_minus = 1.0 - s; }
                                    U[IX(q, pos, l)] = 1.0 - (1.0 - U[IX(q + 1, pos, l)]) * one_minus;
                                }
                                last[pos] = l;
                            }
                            sum += U[IX(i, pos, l)];
                            l++;
                            if (l > count) break;
                            if (sum >= dlk1) break;
                            double c = sum / ((dlk - sum) + 1.0);
                            if (!(c < best)) break;
                        }
                    }
                }
            } else {
                lk = -K;
                dlk1 = (double)(1 - K);
            }
            if (dlk1 > sum) {
                double c = sum / (((double)lk - sum) + 1.0);
                if (best > c) { best = c; *beg = i; *end = pos + 1; *wlen = count; }
            }
        }
    }
#undef IX
    free(prob); free(prob2); free(T); free(U); free(last);
    if (*wlen > 0) {
        while (*beg < *end && nrx_optional(X, *beg)) (*beg)++;
        while (*beg < *end && nrx_optional(X, *end - 1)) (*end)--;
        if (*beg == *end) *wlen = 0;
    }
    if (*wlen == 0) {
        *end = m > 64 ? 64 : m;
        while (nrx_optional(X, *end - 1)) (*end)--;
        best = 1.0;
    }
    return best;
}

int nrx_plan_make(const nrx_pattern *X, nrx_plan *pl)
{
    memset(pl, 0, sizeof *pl);
    pl->cost = ext_find_best(X, 0, &pl->beg, &pl->end, &pl->wlen);
    if (pl->wlen > 0) { pl->type = 2; pl->anchor = pl->beg; }       /* 413550: verification split at the sub-pattern's start */
    else { pl->type = 3; pl->anchor = pl->end; }                    /* 413371: ... at its end (forward scan) */
    /* forward scan of a pattern that begins with an optional position (only possible after the parser's rewrite, e.g.
     * (A?A?C) -> A?C): the binary's scan automaton has no initial closure and misses matches at the first byte of a
     * scan range; not restated */
    return NRO_OK;
}

/* One direction of checkMatch @411aa0 with the tables of extendedLoadVerif @412c60, restated literally (single     */
/* 64-bit word: the oracle takes m <= 64).  Elements are numbered in walk order: u = 0 is the pattern position next to */
/* the anchor (`from`), u grows away from it (`step` = +1 / -1).  State bit u = element u has consumed a byte or was  */
/* skipped.  Quirk kept on purpose: the initial state only pre-skips the FIRST element when it is optional; a run of   */
/* two or more optional elements next to the anchor cannot be skipped as a whole on the first byte, so e.g.            */
/* (GAT.?.?.?AAGTCC) does not match GATAAGTCC (the binary prints nothing for it).                                      */
static int64_t ext_side(const nrx_pattern *X, int from, int len, int step, const uint8_t *t, int64_t pos, int64_t lim)
{
    const nro_pattern *P = &X->P;
    uint64_t I = 0, F = 0, A = 0, init = 0;
    int flag = 0;
    for (int u = 0; u < len; u++) {                       /* 412ee0-412fdc */
        const int j = from + u * step;
        if (!nrx_optional(X, j)) continue;
        if (u > 0) {
            if ((F >> (u - 1)) & 1ULL) {
                F &= ~(1ULL << (u - 1)); F |= 1ULL << u;
                if (flag) A |= 1ULL << u; else init |= 1ULL << u;
            } else {
                I |= 1ULL << (u - 1); F |= 1ULL << u; flag = 1; A |= 1ULL << u;
            }
        } else {
            if (flag) A |= 1ULL << u; else init |= 1ULL << u;
        }
    }
    const uint64_t fin = 1ULL << (len - 1);
    uint64_t D = init, carry = 1;
    int64_t n = 0;
    for (;;) {
        if (D & fin) {
            const int64_t edge = step > 0 ? pos + n : pos - n;
            if (step < 0 ? leftctx(P, t, edge, lim) : rightctx(P, t, edge, lim)) return n;
        }
        const int64_t tp = step > 0 ? pos + n : pos - n - 1;
        if (step > 0 ? tp >= lim : tp < lim) return -1;
        const unsigned c = t[tp];
        uint64_t Bc = 0, Sc = 0;
        for (int u = 0; u < len; u++) {
            const int j = from + u * step;
            if (cls_has(P, j, c)) { Bc |= 1ULL << u; if (nrx_repeat(X, j)) Sc |= 1ULL << u; }
        }
        D = (((D << 1) | carry) & Bc) | (D & Sc);
        carry = 0;
        n++;
        if (!D) return -1;
        const uint64_t x = D | F;
        D = (((~(x - I)) ^ x) & A) | D;
    }
}

/* checkMatch @411aa0 at anchor `pos` (type 2: start of the sub-pattern occurrence, type 3: its end) */
static int ext_check(const nrx_pattern *X, const nrx_plan *pl, const uint8_t *t, int64_t pos, int64_t tbeg, int64_t tend,
                     int64_t *beg, int64_t *end)
{
    const int64_t p = pl->type == 3 ? pos - 1 : pos;
    int64_t rbeg = tbeg, rend = tend;
    for (int64_t q = p - 1; q >= tbeg; q--) if (t[q] == '\n') { rbeg = q + 1; break; }
    for (int64_t q = p; q < tend; q++) if (t[q] == '\n') { rend = q; break; }
    if (p < rbeg || p >= rend) return 0;
    const int m = X->P.m, a = pl->anchor;
    int64_t bext = 0, fext = 0;
    if (a == 0) { if (!leftctx(&X->P, t, pos, rbeg)) return 0; }
    else if ((bext = ext_side(X, a - 1, a, -1, t, pos, rbeg)) < 0) return 0;
    if (a == m) { if (!rightctx(&X->P, t, pos, rend)) return 0; }
    else if ((fext = ext_side(X, a, m - a, +1, t, pos, rend)) < 0) return 0;
    *beg = pos - bext;
    *end = pos + fext;
    return 1;
}

/* extendedScan @4116f0, forward mode (wlen <= 0): Shift-And over the sub-pattern [beg, end) with the closure of its
 * optional runs.  The closure is applied AFTER each byte and the state starts empty (and is emptied by the record
 * delimiter), so a run of optional positions at the start of the sub-pattern is "always on" except for the very first
 * byte of a scan range or record: an occurrence that has to skip it on that byte is not seen. */
typedef struct { uint64_t I, F, A, fin; int beg, len; } ext_fwd;

static void ext_fwd_init(const nrx_pattern *X, const nrx_plan *pl, ext_fwd *f)
{
    memset(f, 0, sizeof *f);
    f->beg = pl->beg; f->len = pl->end - pl->beg;
    for (int u = 0; u < f->len; u++) {
        if (!nrx_optional(X, f->beg + u)) continue;
        if (u > 0 && ((f->F >> (u - 1)) & 1ULL)) { f->F &= ~(1ULL << (u - 1)); f->F |= 1ULL << u; f->A |= 1ULL << u; }
        else { if (u > 0) f->I |= 1ULL << (u - 1); f->F |= 1ULL << u; f->A |= 1ULL << u; }
    }
    f->fin = 1ULL << (f->len - 1);
}

static inline uint64_t ext_fwd_step(const nrx_pattern *X, const ext_fwd *f, uint64_t D, unsigned c)
{
    uint64_t Bc = 0, Sc = 0;
    for (int u = 0; u < f->len; u++)
        if (cls_has(&X->P, f->beg + u, c)) { Bc |= 1ULL << u; if (nrx_repeat(X, f->beg + u)) Sc |= 1ULL << u; }
    D = (((D << 1) | 1ULL) & Bc) | (D & Sc);
    const uint64_t x = D | f->F;
    return (((~(x - f->I)) ^ x) & f->A) | D;
}

/* recSearchFile @402250 over one scan range [lo, hi) (a buffer fill) */
static int64_t ext_search_range(const nrx_pattern *X, const nrx_plan *pl, const uint8_t *text, int64_t lo, int64_t hi,
                                nro_hit *hits, int64_t cap, int64_t cnt)
{
    int64_t pos = lo;
    while (pos < hi) {
        int found = 0;
        int64_t b = 0, e = 0;
        if (pl->type == 2) {
            for (int64_t w = pos; w + pl->wlen <= hi; w++)
                if (ext_check(X, pl, text, w, pos, hi, &b, &e)) { found = 1; break; }
        } else {
            ext_fwd f;
            ext_fwd_init(X, pl, &f);
            uint64_t D = 0;
            for (int64_t q = pos + 1; q <= hi; q++) {
                const unsigned c = text[q - 1];
                if (c == '\n') { D = 0; continue; }
                D = ext_fwd_step(X, &f, D, c);
                if ((D & f.fin) && ext_check(X, pl, text, q, pos, hi, &b, &e)) { found = 1; break; }
            }
        }
        if (!found) break;
        if (cnt < cap) { hits[cnt].beg = b; hits[cnt].end = e; }
        cnt++;
        if (e == hi) break;
        if (e <= pos && b == e) break;
        pos = e;
    }
    return cnt;
}

int64_t nrx_search(const nrx_pattern *X, const nrx_plan *pl, const uint8_t *text, int64_t n, nro_hit *hits, int64_t cap)
{
    if (n <= 0) return 0;
    return ext_search_range(X, pl, text, 0, n, hits, cap, 0);
}

/* with the reference's buffer fills (same rule as nro_search_buffered) */
int64_t nrx_search_buffered(const nrx_pattern *X, const nrx_plan *pl, const uint8_t *t, int64_t n, int64_t bufsize,
                            nro_hit *hits, int64_t cap)
{
    int64_t count = 0, S = 0;
    if (bufsize <= 0) bufsize = n + 1;
    while (n - S > 0) {
        int64_t lo, hi, next;
        int64_t dsize = n - S < bufsize ? n - S : bufsize;
        if (dsize < bufsize) { lo = S; hi = S + dsize; next = n; }
        else {
            int64_t p = S + dsize - 1;
            while (p > S && t[p] != '\n') p--;
            if (p > S) { lo = S; hi = p + 1; next = p; }
            else { lo = S; hi = S + dsize; next = S + dsize; }
        }
        count = ext_search_range(X, pl, t, lo, hi, hits, cap, count);
        S = next;
    }
    return count;
}
