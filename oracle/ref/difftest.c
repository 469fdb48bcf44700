/* difftest.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Differential fuzzer: the reference engine (nrgrep_coords, mapped in-process
 * by refload.c) against the C restatement in oracle/nrgrep_oracle.c.
 * Compares (a) the search plan chosen by esimplePreproc (type, split points,
 * piece length) and (b) the complete hit list of recSearchFile's scan loop on
 * random texts.  Runs only in the build container (needs /root/reference).
 *
 *   usage: difftest [-n cases] [-s seed] [-a dna|pep|mix] [-z] [-v]
 *     -z  bind the reference's malloc to a zero-filling allocator
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "refload.h"
#include "../nrgrep_oracle.h"
#include <ctype.h>

static int my_puts(const char *s) { (void)s; return 0; }
static void *zmalloc(size_t n) { return calloc(1, n ? n : 1); }

static uint64_t rng_state = 88172645463325252ULL;
static uint64_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static int rint_(int n) { return (int)(rnd() % (uint64_t)n); }

static const char *DNA = "ACGT";
static const char *PEP = "ACDEFGHIKLMNPQRSTVWY";

static int gen_pattern(char *out, int m, const char *alpha, int cls_pct, int dot_pct, int neg_pct)
{
    int na = (int)strlen(alpha), o = 0;
    out[o++] = '(';
    for (int j = 0; j < m; j++) {
        int r = rint_(100);
        if (r < dot_pct) out[o++] = '.';
        else if (r < dot_pct + cls_pct) {
            out[o++] = '[';
            if (rint_(100) < neg_pct) out[o++] = '^';
            int n = 2 + rint_(na > 4 ? 4 : 2);
            char used[64] = {0};
            for (int x = 0; x < n; x++) { int c = rint_(na); if (!used[c]) { used[c] = 1; out[o++] = alpha[c]; } }
            out[o++] = ']';
        } else out[o++] = alpha[rint_(na)];
    }
    out[o++] = ')';
    out[o] = 0;
    return o;
}

static void mutate_into(unsigned char *t, int *n, int cap, const nro_pattern *P, const char *alpha, int nerr)
{
    /* plant a (possibly mutated) instance of the pattern */
    int na = (int)strlen(alpha);
    unsigned char inst[NRO_MAXM * 2]; int L = 0;
    for (int j = 0; j < P->m; j++) {
        /* pick a byte from the class, prefer alphabet letters */
        int c = -1;
        for (int tries = 0; tries < 16 && c < 0; tries++) { int x = alpha[rint_(na)]; if ((P->cls[j][x >> 6] >> (x & 63)) & 1) c = x; }
        if (c < 0) c = alpha[rint_(na)];
        inst[L++] = (unsigned char)c;
    }
    for (int e = 0; e < nerr && L > 1; e++) {
        int op = rint_(3), p = rint_(L);
        if (op == 0) inst[p] = (unsigned char)alpha[rint_(na)];
        else if (op == 1) { memmove(inst + p, inst + p + 1, (size_t)(L - p - 1)); L--; }
        else { memmove(inst + p + 1, inst + p, (size_t)(L - p)); inst[p] = (unsigned char)alpha[rint_(na)]; L++; }
    }
    if (*n + L < cap) { memcpy(t + *n, inst, (size_t)L); *n += L; }
}

int main(int argc, char **argv)
{
    long cases = 20000; int verbose = 0, zero = 0, longm = 0; const char *amode = "mix";
    int opt;
    while ((opt = getopt(argc, argv, "n:s:a:zvM:")) != -1) {
        if (opt == 'M') longm = atoi(optarg);            /* patterns of 65 .. M positions (multi-word masks) */
        if (opt == 'n') cases = atol(optarg);
        else if (opt == 's') rng_state ^= (uint64_t)atol(optarg) * 0x9E3779B97F4A7C15ULL;
        else if (opt == 'a') amode = optarg;
        else if (opt == 'z') zero = 1;
        else if (opt == 'v') verbose++;
    }
    ref_override("puts", (void *)my_puts);
    if (zero) ref_override("malloc", (void *)zmalloc);
    if (ref_load("/root/reference/www/bin/nrgrep_coords")) return 2;
    ((ref_void_t)REF_recPreproc)();

    static unsigned char buf[1 << 16];
    unsigned char *text = buf + 256;
    long plan_bad = 0, hit_bad = 0, types[4] = {0, 0, 0, 0}, total_hits = 0;
    for (long cs = 0; cs < cases; cs++) {
        const char *alpha = !strcmp(amode, "dna") ? DNA : !strcmp(amode, "pep") ? PEP : (rint_(2) ? DNA : PEP);
        int m = longm > 64 ? 65 + (int)rint_((unsigned)(longm - 64)) : 3 + rint_(rint_(4) ? 22 : 60);
        int k = rint_(8) == 0 ? 0 : 1 + rint_(3);
        if (k >= m) k = m - 1;
        int ids = 1 + rint_(7);
        int ins = ids & 1, del = (ids >> 1) & 1, subs = (ids >> 2) & 1;
        char pat[1024], pat2[1024];
        gen_pattern(pat, m, alpha, rint_(3) ? 20 : 0, rint_(3) ? 8 : 0, 15);
        strcpy(pat2, pat);

        nro_pattern P; nro_plan pl;
        int rc = nro_parse(pat, 1, &P);
        if (rc) { fprintf(stderr, "oracle parse failed %d on %s\n", rc, pat); return 1; }
        nro_plan_make(&P, k, ins, del, subs, &pl);

        REF_OptCaseInsensitive = 1; REF_OptErrors = k;
        REF_OptIns = ins; REF_OptDel = del; REF_OptSubs = subs; REF_OptTransp = 0;
        REF_OptStartLine = 0; REF_OptEndLine = 0;
        long *sd = ((ref_searchPreproc_t)REF_searchPreproc)(pat2);
        if (!sd) { fprintf(stderr, "ref preproc failed on %s\n", pat); return 1; }
        int rtype = (int)sd[0];       /* 1 SIMPLE, 4 ESIMPLE */
        int bad = 0;
        if (k == 0) { if (rtype != 1) bad = 1; }
        else if (rtype != 4) bad = 1;
        else {
            unsigned char *E = (unsigned char *)sd[1];
            int etype = *(int *)(E + 0x1c);
            int *V1 = *(int **)(E + 0x30);
            unsigned char *S = *(unsigned char **)(E + 0x10);
            int slen = *(int *)(*(unsigned char **)S + 0x800);
            if (etype != pl.type) bad = 1;
            else {
                for (int i = 0; i < pl.npieces; i++) if (V1[i] != pl.V[i]) bad = 1;
                if (slen != pl.L) bad = 1;
            }
            if (bad || verbose > 1) {
                fprintf(stderr, "%s plan: %s k=%d ids=%d%d%d  ref type=%d L=%d V=", bad ? "BAD" : "ok", pat, k, ins, del, subs, etype, slen);
                for (int i = 0; i < (etype == 1 ? k + 1 : 1); i++) fprintf(stderr, "%d,", V1[i]);
                fprintf(stderr, "  oracle type=%d L=%d V=", pl.type, pl.L);
                for (int i = 0; i < pl.npieces; i++) fprintf(stderr, "%d,", pl.V[i]);
                fprintf(stderr, " split=%.17g fb=%.17g\n", pl.split_cost, pl.fb_cost);
            }
        }
        types[pl.type]++;
        if (bad) { plan_bad++; ((ref_free_t)REF_searchFree)(sd); continue; }

        for (int rep = 0; rep < 3; rep++) {
            int n = 0, want = 30 + rint_(400), na = (int)strlen(alpha);
            int lower = rint_(4) == 0;
            while (n < want) {
                int r = rint_(100);
                if (r < 30) mutate_into(text, &n, want + 200, &P, alpha, rint_(k + 2));
                else if (r < 34) text[n++] = '\n';
                else if (r < 36) text[n++] = (unsigned char)"N>x- "[rint_(5)];
                else { int run = 1 + rint_(12); for (int x = 0; x < run; x++) text[n++] = (unsigned char)alpha[rint_(na)]; }
            }
            if (lower) for (int x = 0; x < n; x++) if (rint_(3) == 0) text[x] = (unsigned char)tolower(text[x]);
            text[n] = '\n'; text[-1] = '\n';

            nro_hit oh[512]; int64_t on = nro_search(&P, &pl, text, n, oh, 512);
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
                if (hit_bad <= 10 || verbose) {
                    fprintf(stderr, "BAD hits: %s k=%d ids=%d%d%d type=%d L=%d V=", pat, k, ins, del, subs, pl.type, pl.L);
                    for (int i = 0; i < pl.npieces; i++) fprintf(stderr, "%d,", pl.V[i]);
                    fprintf(stderr, "\n text(%d)=", n);
                    for (int x = 0; x < n; x++) fputc(text[x] == '\n' ? '|' : text[x], stderr);
                    fprintf(stderr, "\n  ref:"); for (int64_t x = 0; x < rn && x < 512; x++) fprintf(stderr, " [%ld,%ld)", (long)rh[x].beg, (long)rh[x].end);
                    fprintf(stderr, "\n  orc:"); for (int64_t x = 0; x < on && x < 512; x++) fprintf(stderr, " [%ld,%ld)", (long)oh[x].beg, (long)oh[x].end);
                    fprintf(stderr, "\n");
                }
            }
        }
        ((ref_free_t)REF_searchFree)(sd);
    }
    printf("cases=%ld plan_mismatch=%ld hit_mismatch=%ld  types: simple=%ld split=%ld bwd=%ld fwd=%ld  ref_hits=%ld\n",
           cases, plan_bad, hit_bad, types[0], types[1], types[2], types[3], total_hits);
    return (plan_bad || hit_bad) ? 1 : 0;
}
