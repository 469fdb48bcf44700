/* nrgrep_oracle.h -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped).
 *
 * Plain-C restatement of the search semantics of the reference's search
 * engine, /root/reference/www/bin/nrgrep_coords (nrgrep 1.1 patched by the
 * PatMatch authors to print match coordinates; shipped as an unstripped
 * x86-64 binary WITHOUT sources).  Because there is no source, every
 * function below cites the symbol + address range of the routine in that
 * binary whose behaviour it restates (see `nm nrgrep_coords`), and the call
 * site in the reference's Python that invokes it
 * (www/FlaskApp/FlaskApp/patmatch.py:733-743).
 *
 * Covered: the SIMPLE (k = 0) and ESIMPLE (k > 0) engines, line anchors, buffer fills, and the EXTENDED engine
 * for k = 0 (nrx_*: positions followed by ? * +).
 *
 * Parity pinning: oracle/ref/difftest.c (and difftest_ext.c for EXTENDED patterns) maps the reference binary into the
 * test process (oracle/ref/refload.c) and compares this restatement against
 * the reference's own searchPreproc/searchScan on millions of random
 * (pattern, options, text) cases; tests/golden/ holds vectors produced by
 * running the reference binary itself (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * use this code.  The product path (patmatchdocker_b200/csrc) never does.
 */
#ifndef NRGREP_ORACLE_H
#define NRGREP_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRO_MAXM 256            /* max pattern positions handled by the oracle */
#define NRO_MAXK 15
#define NRO_WORDS (NRO_MAXM / 64)

enum { NRO_SIMPLE = 0, NRO_SPLIT = 1, NRO_BWD = 2, NRO_FWD = 3 };

enum {
    NRO_OK = 0,
    NRO_ERR_SYNTAX = -1,        /* malformed pattern */
    NRO_ERR_UNSUPPORTED = -2,   /* extended / regular pattern (?, *, +, |): not this path */
    NRO_ERR_TOOLONG = -3
};

typedef struct {
    int m;                       /* number of pattern positions */
    int start_line, end_line;    /* leading '^' / trailing '$' (main @400e00: 40164d, 40162a) */
    uint64_t cls[NRO_MAXM][4];   /* 256-bit byte class per position */
} nro_pattern;

typedef struct {
    int k, ins, del, subs;       /* -k <k>[ids] */
    int type;                    /* NRO_SIMPLE (k == 0) / NRO_SPLIT / NRO_BWD / NRO_FWD */
    int L;                       /* SPLIT: piece length; BWD/FWD: sub-pattern length */
    int npieces;                 /* SPLIT: k+1 ; BWD/FWD: 1 */
    int V[NRO_MAXK + 1];         /* split points: left part = P[0..V[i]), right part = P[V[i]..m) */
    double split_cost, fb_cost;  /* cost-model values (exact doubles of the reference) */
    int fb_flag, fb_beg, fb_end; /* simpleFindBest outputs */
} nro_plan;

typedef struct { int64_t beg, end; } nro_hit;

int nro_parse(const char *pattern, int icase, nro_pattern *P);
int nro_plan_make(const nro_pattern *P, int k, int ins, int del, int subs, nro_plan *plan);
/* 1: nro_plan_make reproduces the piece choice of the DEPLOYED binary (default glibc 2.39 allocator: esimplePreproc's
 * never-written scratch cells hold the table simpleFindBest has just freed); 0 (default): the defined behaviour, the
 * cells read as +0.0 (what the binary does with zero-filled malloc).  Not thread-safe (test infrastructure). */
void nro_set_compat(int deployed_glibc);

/* Whole-file search exactly as recSearchFile @402250 drives searchScan: returns the
 * number of hits (may exceed cap; only the first cap are stored). Offsets are
 * byte offsets into text, [beg, end). */
int64_t nro_search(const nro_pattern *P, const nro_plan *plan,
                   const uint8_t *text, int64_t n, nro_hit *hits, int64_t cap);

/* The same with the reference's buffer fills of `bufsize` bytes (-b; patmatch.py:731 passes
 * 1600000): recSearchFile scans one fill at a time and a hit never crosses a fill. */
int64_t nro_search_buffered(const nro_pattern *P, const nro_plan *plan, const uint8_t *text, int64_t n,
                            int64_t bufsize, nro_hit *hits, int64_t cap);
int64_t nro_buffer_fills(const uint8_t *text, int64_t n, int64_t bufsize, int64_t *seg, int64_t segcap);

/* One verification call: esimple checkMatch @4151d0 (+ checkMatch1 @414190) for
 * candidate (piece i, position pos) with the scan range [tbeg, tend). */
int nro_check_match(const nro_pattern *P, const nro_plan *plan, int i,
                    const uint8_t *text, int64_t pos, int64_t tbeg, int64_t tend,
                    int64_t *beg, int64_t *end);

/* ---- EXTENDED patterns, k = 0 (see the end of nrgrep_oracle.c and NOTES_extended.md) ---- */
enum { NRX_NONE = 0, NRX_OPT = 1, NRX_STAR = 2, NRX_PLUS = 3 };
#define NRX_REWRITTEN_SIMPLE 1   /* nrx_parse: the parser's rewrites left a SIMPLE pattern in X->P */
typedef struct { nro_pattern P; unsigned char op[NRO_MAXM]; } nrx_pattern;
typedef struct { int type, anchor, wlen, beg, end; double cost; } nrx_plan;
int nrx_parse(const char *pattern, int icase, nrx_pattern *X);
int nrx_plan_make(const nrx_pattern *X, nrx_plan *plan);
int64_t nrx_search(const nrx_pattern *X, const nrx_plan *plan, const uint8_t *text, int64_t n, nro_hit *hits, int64_t cap);
int64_t nrx_search_buffered(const nrx_pattern *X, const nrx_plan *plan, const uint8_t *text, int64_t n, int64_t bufsize,
                            nro_hit *hits, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
