/* difftest_ext.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Differential fuzzing of the EXTENDED (k = 0) restatement in oracle/nrgrep_oracle.c (nrx_*) against the reference
 * binary mapped in process (refload.c): the plan (verification type, anchor, window) is compared with the structure
 * extendedPreproc @413260 builds, the hit lists with searchScan @402820 driven like recSearchFile @402250.
 * usage: difftest_ext [-n cases] [-s seed] [-a dna|pep] [-v]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include "refload.h"
#include "../nrgrep_oracle.h"

static unsigned long long rs = 88172645463325252ULL;
static unsigned rint_(unsigned n) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (unsigned)((rs >> 11) % n); }
static int my_puts(const char *s) { (void)s; return 0; }
static void *zmalloc(size_t n) { return calloc(1, n ? n : 1); }

int main(int argc, char **argv)
{
    long cases = 2000; int verbose = 0, edges = 0; const char *alpha = "ACGT";
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-n")) cases = atol(argv[++i]);
        else if (!strcmp(argv[i], "-s")) rs ^= (unsigned long long)atol(argv[++i]) * 0x9E3779B97F4A7C15ULL;
        else if (!strcmp(argv[i], "-a")) alpha = !strcmp(argv[++i], "pep") ? "ACDEFGHIKLMNPQRSTVWY" : "ACGT";
        else if (!strcmp(argv[i], "-v")) verbose++;
        else if (!strcmp(argv[i], "-e")) edges = 1;            /* operators on the first / last position too */
    }
    ref_override("puts", (void *)my_puts);
    ref_override("malloc", (void *)zmalloc);
    if (ref_load("/root/reference/www/bin/nrgrep_coords")) return 2;
    ((ref_void_t)REF_recPreproc)();
    const int na = (int)strlen(alpha);
    static unsigned char buf[1 << 16];
    unsigned char *text = buf + 256;
    long plan_bad = 0, hit_bad = 0, total_hits = 0, t2 = 0, t3 = 0, simple_scan = 0, nsimple = 0;
    for (long cs = 0; cs < cases; cs++) {
        int m = 3 + (int)rint_(rint_(3) ? 10 : 22);
        char pat[1024]; int o = 0;
        /* members[j]: a byte that position j accepts (for planting), ops[j] */
        char member[64]; int ops[64];
        pat[o++] = '(';
        int nops = 0;
        for (int j = 0; j < m; j++) {
            unsigned r = rint_(100);
            if (r < 12) { pat[o++] = '.'; member[j] = alpha[rint_(na)]; }
            else if (r < 27) { char a = alpha[rint_(na)], b = alpha[rint_(na)]; pat[o++] = '['; pat[o++] = a; pat[o++] = b; pat[o++] = ']'; member[j] = rint_(2) ? a : b; }
            else { char a = alpha[rint_(na)]; pat[o++] = a; member[j] = a; }
            ops[j] = 0;
            if ((j > 0 && j < m - 1) || (edges && rint_(2))) {
                unsigned q = rint_(100);
                if (q < 25) { pat[o++] = '?'; ops[j] = 1; nops++; }
                else if (q < 33) { pat[o++] = '*'; ops[j] = 2; nops++; }
                else if (q < 38) { pat[o++] = '+'; ops[j] = 3; nops++; }
            }
        }
        pat[o++] = ')'; pat[o] = 0;
        if (!nops) { cs--; continue; }
        char pat2[1024]; strcpy(pat2, pat);
        nrx_pattern X; nrx_plan pl;
        int rc = nrx_parse(pat, 1, &X);
        if (rc == NRO_ERR_UNSUPPORTED && edges) { cs--; continue; }       /* nothing left after the rewrites */
        if (rc != NRO_OK && rc != NRX_REWRITTEN_SIMPLE) { fprintf(stderr, "oracle parse failed %d on %s\n", rc, pat); return 1; }
        const int simple = rc == NRX_REWRITTEN_SIMPLE;
        nro_plan spl;
        memset(&pl, 0, sizeof pl);
        if (simple) nro_plan_make(&X.P, 0, 1, 1, 1, &spl);
        else if (nrx_plan_make(&X, &pl) == NRO_ERR_UNSUPPORTED) { cs--; continue; }

        REF_OptCaseInsensitive = 1; REF_OptErrors = 0; REF_OptIns = 1; REF_OptDel = 1; REF_OptSubs = 1; REF_OptTransp = 0;
        REF_OptStartLine = 0; REF_OptEndLine = 0;
        long *sd = ((ref_searchPreproc_t)REF_searchPreproc)(pat2);
        if (!sd) { fprintf(stderr, "ref preproc failed on %s\n", pat); return 1; }
        int rtype = (int)sd[0];
        unsigned char *E = (unsigned char *)sd[1];
        int bad = 0, vt = -1, anchor = -1, rwl = -1; unsigned long scanfn = 0;
        if (simple) { if (rtype != 1) bad = 1; nsimple++; }
        else if (rtype != 2) bad = 1;
        else {
            scanfn = *(unsigned long *)E;
            anchor = *(int *)(E + 0x2060); vt = *(int *)(E + 0x2068);
            unsigned char *S = *(unsigned char **)(E + 0x10);
            rwl = scanfn == 0x416600UL ? *(int *)(S + 0x800) : *(int *)(S + 0x1018);
            if (vt != pl.type || anchor != pl.anchor) bad = 1;
            if (scanfn == 0x416600UL) { simple_scan++; if (rwl != pl.end - pl.beg) bad = 1; }
            else if (rwl != pl.wlen) bad = 1;
        }
        if (!simple) { if (pl.type == 2) t2++; else t3++; }
        if (bad || verbose > 1)
            fprintf(stderr, "%s plan: %s  ref: class=%d scan=%lx vtype=%d anchor=%d win=%d   oracle: type=%d anchor=%d beg=%d end=%d wlen=%d cost=%.17g\n",
                    bad ? "BAD" : "ok", pat, rtype, scanfn, vt, anchor, rwl, pl.type, pl.anchor, pl.beg, pl.end, pl.wlen, pl.cost);
        if (bad) { plan_bad++; ((ref_free_t)REF_searchFree)(sd); continue; }

        for (int rep = 0; rep < 3; rep++) {
            int n = 0, want = 30 + (int)rint_(400);
            while (n < want) {
                unsigned r = rint_(100);
                if (r < 35) {          /* plant an instance: optional positions 0-2 times, with an occasional mutation */
                    for (int j = 0; j < m; j++) {
                        int reps = ops[j] == 0 ? 1 : ops[j] == 1 ? (int)rint_(2) : ops[j] == 2 ? (int)rint_(3) : 1 + (int)rint_(2);
                        for (int x = 0; x < reps; x++) text[n++] = (unsigned char)(rint_(25) ? member[j] : alpha[rint_(na)]);
                    }
                } else if (r < 39) text[n++] = '\n';
                else if (r < 41) text[n++] = (unsigned char)"N>x- "[rint_(5)];
                else { int run = 1 + (int)rint_(10); for (int x = 0; x < run; x++) text[n++] = (unsigned char)alpha[rint_(na)]; }
            }
            if (rint_(4) == 0) for (int x = 0; x < n; x++) if (rint_(3) == 0) text[x] = (unsigned char)tolower(text[x]);
            text[n] = '\n'; text[-1] = '\n';
            nro_hit oh[512]; int64_t on = simple ? nro_search(&X.P, &spl, text, n, oh, 512) : nrx_search(&X, &pl, text, n, oh, 512);
            nro_hit rh[512]; int64_t rn = 0;
            unsigned char *pos = text, *top = text + n;
            for (;;) {
                unsigned char *b = pos, *e = top;
                if (!((ref_searchScan_t)REF_searchScan)(&b, &e, sd)) break;
                if (rn < 512) { rh[rn].beg = b - text; rh[rn].end = e - text; }
                rn++;
                if (e == top) break;
                if (e <= pos && b == e) break;
                pos = e;
            }
            total_hits += rn;
            int hb = (on != rn);
            for (int64_t x = 0; !hb && x < rn && x < 512; x++) if (oh[x].beg != rh[x].beg || oh[x].end != rh[x].end) hb = 1;
            if (hb) {
                hit_bad++;
                if (hit_bad <= 8 || verbose) {
                    fprintf(stderr, "BAD hits: %s type=%d anchor=%d wlen=%d\n text(%d)=", pat, pl.type, pl.anchor, pl.wlen, n);
                    for (int x = 0; x < n; x++) fputc(text[x] == '\n' ? '|' : text[x], stderr);
                    fprintf(stderr, "\n  ref:"); for (int64_t x = 0; x < rn && x < 512; x++) fprintf(stderr, " [%ld,%ld)", (long)rh[x].beg, (long)rh[x].end);
                    fprintf(stderr, "\n  orc:"); for (int64_t x = 0; x < on && x < 512; x++) fprintf(stderr, " [%ld,%ld)", (long)oh[x].beg, (long)oh[x].end);
                    fprintf(stderr, "\n");
                }
            }
        }
        ((ref_free_t)REF_searchFree)(sd);
    }
    printf("cases=%ld plan_mismatch=%ld hit_mismatch=%ld  type2=%ld type3=%ld simple_scan=%ld rewritten_to_simple=%ld ref_hits=%ld\n",
           cases, plan_bad, hit_bad, t2, t3, simple_scan, nsimple, total_hits);
    return (plan_bad || hit_bad) ? 1 : 0;
}
