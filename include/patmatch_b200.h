/* patmatch_b200.h -- C ABI of the B200-native PatMatch search engine.
 *
 * Drop-in boundary.  The reference Flask app (www/FlaskApp/FlaskApp/patmatch.py)
 * reaches its search engine through two shell-outs:
 *
 *   patmatch.py:733-735   output  = os.popen(nrgrep_coords -i -b 1600000
 *                                       -k <k><ids> '<pattern>' '<datafile>').read()
 *   patmatch.py:739-743   output2 = the same for the reverse-complement pattern
 *
 * and then parses the "[beg, end]: text" lines (patmatch.py:505-516).  Every
 * entry point below replaces one of those interfaces; the ctypes stub a
 * maintainer adds is shown in INTEGRATION.md.  Plain pointers and sizes only.
 *
 * All functions return 0 on success, a negative PM_ERR_* code otherwise;
 * pm_last_error() gives the message for the calling thread.  There is no CPU
 * fallback: without a CUDA device pm_engine_create fails.
 *
 * Threads: a pm_engine owns mutable scratch (candidate buffers, pinned staging, statistics).  Every entry point
 * that takes an engine holds the engine's mutex for its duration, so concurrent callers are serialised per engine;
 * pm_last_hits / pm_get_stats refer to the last call of ANY thread, so callers that need them must hold their own
 * lock around the pair or use one engine per thread.
 */
#ifndef PATMATCH_B200_H
#define PATMATCH_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_OK               0
#define PM_ERR_CUDA        -1   /* CUDA runtime failure (message has the cudaError) */
#define PM_ERR_SYNTAX      -2   /* malformed nrgrep pattern / -k option */
#define PM_ERR_UNSUPPORTED -3   /* pattern needs nrgrep's REGULAR engine (groups with operators, |), is an EXTENDED pattern
                                 * (? * +) with errors, is a single position that accepts '\n', or has more than 255 positions
                                 * (64 for EXTENDED patterns) */
#define PM_ERR_ARG         -4
#define PM_ERR_OVERFLOW    -5   /* caller's hit buffer too small; *nhits holds the required size */

#define PM_MAX_PIECES 16

/* search plan types = nrgrep's esimple scan types (esimpleScan @4136d0) */
#define PM_PLAN_SIMPLE 0        /* k = 0: exact scan (simpleScan @416600) */
#define PM_PLAN_SPLIT  1        /* k+1 exact pieces + verification */
#define PM_PLAN_BWD    2        /* backward approximate filter */
#define PM_PLAN_FWD    3        /* forward approximate filter */
/* nrgrep's EXTENDED engine, k = 0 (positions followed by ? * +: PatMatch's X{m,n} / X{m,}); extendedPreproc @413260 */
#define PM_PLAN_EXT_BEG 4       /* verification anchored at the START of the scanned sub-pattern */
#define PM_PLAN_EXT_END 5       /* ... at its END (forward scan) */

typedef struct pm_engine pm_engine;
typedef struct pm_dataset pm_dataset;

/* one output line of nrgrep_coords: "[beg, end]: <bytes beg..end-1>" (recSearchFile @402250) */
typedef struct { int64_t beg, end; } pm_hit;

/* what esimplePreproc @415540 decides for (pattern, k); host-only introspection */
typedef struct {
    int m, k, ins, del, subs;
    int type, L, npieces;
    int V[PM_MAX_PIECES];
    double split_cost, fb_cost;
} pm_plan_info;

/* per-search device timings (ms, CUDA events on the engine's stream) and counters */
typedef struct {
    float scan_ms, sort_ms, verify_ms, chain_ms, total_ms;
    int64_t candidates, verified, hits;
    int64_t scan_bytes;          /* HBM bytes the scan kernel had to read (its algorithmic bytes) */
    int64_t scan_bases;          /* text positions the scan kernel covered */
    int launches;                /* kernels launched by the last search */
    int packed;                  /* 1: 2-bit packed bit-sliced scan, 2: Shift-And over 5-bit residue codes, 0: byte Shift-And / dense */
    int qgram_chunks;            /* pattern chunks of the bit-sliced q-gram pre-filter (0 = not used) */
    int syncs;                   /* host synchronisations of the last pm_search_request */
    int jit;                     /* 1: the scan ran as a kernel specialised for the request (NVRTC), 0: generic kernels */
} pm_stats;

const char *pm_last_error(void);
const char *pm_version(void);

/* host-only: parse `pattern` (nrgrep syntax as written by patmatch_to_nrgrep.pl) and the
 * -k argument ("2ids", "1s", "0ids", ...) and report the plan the reference would choose. */
int pm_plan(const char *pattern, const char *kopt, pm_plan_info *info);

/* Approximate searches of the reference read never-written scratch cells while choosing their k+1 pieces
 * (esimplePreproc @415a68), so the DEPLOYED binary's piece choice -- and through it some hit boundaries -- depends on
 * what glibc's malloc hands back.  Default (0): the defined behaviour, the cells read as +0.0 (the stock binary under
 * GLIBC_TUNABLES=glibc.malloc.tcache_count=0:glibc.malloc.perturb=255).  1: what the cells hold in the stock CLI
 * process with the default glibc 2.39 allocator (the table simpleFindBest @416a10 has just freed, at the offset its
 * chunk is split at); reproduces the deployed binary on 2 000 of 2 000 random searches (tools/deployed_gap.py).
 * Process-wide; affects patterns of 11 or more positions with k > 0 only. */
int pm_set_compat_deployed_glibc(int on);

int pm_engine_create(int device, pm_engine **out);
void pm_engine_destroy(pm_engine *e);
/* launch on this cudaStream_t instead of the engine's own (non-blocking) stream; 0 = back to own.  To share the
 * legacy default stream (handle 0, e.g. torch's default stream) pass cudaStreamLegacy, (void *)1. */
int pm_engine_set_stream(pm_engine *e, void *cuda_stream);
int pm_engine_synchronize(pm_engine *e);
/* scan kernel selection: 0 = auto (packed for DNA-like datasets), 1 = byte Shift-And, 2 = packed */
int pm_engine_set_scan_mode(pm_engine *e, int mode);
/* packed scan only: filter candidates inside the scan kernel on the packed planes and drop the ones
 * that surely fail.  1 (default) = q-gram pre-filter + Myers filter, 2 = Myers filter only,
 * 0 = off: every exact piece hit goes through k_verify */
int pm_engine_set_fused_filter(pm_engine *e, int on);
/* pm_search_batch with 64 or more exact motifs hashes every text position once (8-mer code -> the motifs whose most
 * selective 8-position window accepts it) and verifies only those motifs, instead of evaluating every motif at every
 * position.  1 (default) / 0 = always the dense multi-pattern kernel.  Results are identical. */
int pm_engine_set_batch_lookup(pm_engine *e, int on);
/* Datasets that are not DNA-like (proteomes) are also kept as 5-bit residue codes (letters of either case -> 0..25,
 * every other byte -> 31; six codes per 32-bit word) and the Shift-And scan reads those instead of the raw bytes:
 * 0.667 B per residue.  Candidates are re-checked on the raw bytes, results do not change.  1 (default) / 0. */
int pm_engine_set_peptide_codes(pm_engine *e, int on);
/* Specialised scan kernels: for approximate (SPLIT) searches the engine can write the dense part of the scan as
 * straight-line CUDA for exactly the request's patterns and compile it with NVRTC for sm_100a (about 0.2 s, cached
 * per process and request text).  0 = never; 1 (default) = for requests that cover at least 2^27 bases per pattern,
 * when libnvrtc is present: the compilation runs on a background thread and the generic kernel answers until the
 * specialised one is there, so no request ever waits for the compiler; 2 = always, compiled synchronously (error when
 * NVRTC is missing).  Results are identical either way. */
int pm_engine_set_jit(pm_engine *e, int mode);
/* Blocks until no background compilation (mode 1 above) is in flight: a service calls it after warming up its common
 * motifs, a benchmark before its timed region. */
int pm_jit_wait(void);
/* host-only: the CUDA source the engine would compile for this request (debugging, SASS inspection).
 * Returns the number of bytes needed (including the terminating 0) or a negative PM_ERR_*. */
int64_t pm_jit_source(int npat, const char *const *patterns, const char *kopt, char *buf, int64_t cap);
/* drops the per-process caches of compiled patterns and chunk plans (experiments that change PM_APX_* environment knobs) */
int pm_debug_reset_caches(void);
/* the reference's -b buffer size in BYTES (patmatch.py:37,733 pass 1600000, the default here):
 * nrgrep_coords scans the file one buffer fill at a time and no hit crosses a fill. 0 = one fill. */
int pm_engine_set_buffer_size(pm_engine *e, int64_t bytes);

/* Replaces '<datafile>': the bytes of a .seq FASTA file (one sequence per line,
 * patmatch.py:700) are copied to HBM once and stay resident across searches. */
int pm_dataset_create(pm_engine *e, const uint8_t *host_bytes, int64_t n, pm_dataset **out);
/* same, bytes already on the device (pointer stays owned by the caller) */
int pm_dataset_wrap_device(pm_engine *e, const uint8_t *device_bytes, int64_t n, pm_dataset **out);
/* Multi-GPU cold request (replaces the part of the file a rank would read; the reference reads the whole file per run,
 * patmatch.py:733-743): a rank that searches only the buffer fills starting in [pos_beg, pos_end) needs only the bytes
 * of those fills.  host_bytes points at the WHOLE file in host memory (pinned for an asynchronous copy); only
 * [win_lo, win_hi) is copied and packed, positions stay absolute.  dev_newlines (device memory, nl_rows x int64)
 * receives [0] = the number of record delimiters in the window and [1..] their positions (unordered, at most
 * nl_rows - 1 stored).  No host synchronisation.  Before searching, the caller installs the newline index of the whole
 * file (gathered from all ranks) with pm_dataset_set_newlines; pm_request_fills_device then works on fills inside
 * the window. */
int pm_dataset_create_window(pm_engine *e, const uint8_t *host_bytes, int64_t n, int64_t win_lo, int64_t win_hi,
                             void *dev_newlines, int64_t nl_rows, pm_dataset **out);
int pm_dataset_set_newlines(pm_engine *e, pm_dataset *d, const int64_t *sorted_positions, int64_t count);
void pm_dataset_destroy(pm_dataset *d);
int64_t pm_dataset_size(const pm_dataset *d);

/* Replaces one nrgrep_coords run (patmatch.py:733-735): all hits, in output order.
 * hits may be NULL to only count.  *nhits receives the number of hits found. */
int pm_search(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt,
              pm_hit *hits, int64_t cap, int64_t *nhits);

/* Cold request (what every nrgrep_coords run is: it reads the file again): the file bytes are still in HOST memory.
 * Uploads them in chunks (chunk_bytes, 0 = 256 MiB; pinned memory gives the overlap) and, while the next chunk
 * crosses PCIe, packs the chunk that arrived and searches the buffer fills it completes, for every pattern.
 * Results as pm_search_batch (offsets has npat+1 entries); *out receives the dataset, resident for later searches. */
int pm_search_stream(pm_engine *e, const uint8_t *host_bytes, int64_t n, int npat, const char *const *patterns,
                     const char *kopt, int64_t chunk_bytes, pm_hit *hits, int64_t cap, int64_t *offsets, pm_dataset **out);

/* After pm_search returned PM_ERR_OVERFLOW (or was called with hits = NULL): copy the hit list of that
 * search, which is still on the device, without searching again. */
int pm_last_hits(pm_engine *e, pm_hit *hits, int64_t cap, int64_t *nhits);

/* Batched motifs (same -k for all): hit lists are concatenated, offsets[i]..offsets[i+1]
 * delimit pattern i.  offsets has npat+1 entries. */
int pm_search_batch(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns,
                    const char *kopt, pm_hit *hits, int64_t cap, int64_t *offsets);
/* The same batch over the buffer fills that START in [pos_beg, pos_end) only (whole or windowed dataset; positions are
 * file offsets).  Buffer fills are independent by the reference's own restart rule (bufLoad @41bbf0 hands whole
 * fills to the scan), so the lists of a partition of the file into position ranges concatenate, motif by motif, to the
 * lists of pm_search_batch: this is how a batch is TEXT-sharded over the GPUs of a box (every rank: all motifs, 1/N of
 * the genomes) -- the lookup kernel hashes every position once whatever the number of motifs, so sharding the text
 * scales where sharding the motif list does not. */
int pm_search_batch_fills(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns, const char *kopt,
                          int64_t pos_beg, int64_t pos_end, pm_hit *hits, int64_t cap, int64_t *offsets);
/* The same with COMPACT hit lists, for batches whose result is large (10 000 motifs x 600 Mb: 2.3e8 hits = 3.7 GB of
 * pm_hit rows, more PCIe time than the search takes): begins[i] = hit begin - *base as 32 bits, motif_len[p] (npat
 * entries, may be NULL) = positions of motif p, so hit i of motif p is [*base + begins[i], *base + begins[i] +
 * motif_len[p]).  Only for what the fused batch path serves -- exact motifs (-k 0) without repeats on a DNA dataset,
 * a position range below 2^32 bytes (pos_end < 0: the whole dataset) -- otherwise PM_ERR_UNSUPPORTED and the caller
 * uses pm_search_batch_fills.  PM_ERR_OVERFLOW: offsets[npat] holds the number of hits, call again with that much room. */
int pm_search_batch_fills_compact(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns, const char *kopt,
                                  int64_t pos_beg, int64_t pos_end, uint32_t *begins, int64_t cap, int64_t *offsets,
                                  int64_t *base, uint16_t *motif_len);

/* One PatMatch request in ONE pass (replaces BOTH nrgrep_coords runs of patmatch.py:733-735 and :739-743: the pattern
 * and its reverse complement; any number of patterns with the same -k works).  All patterns are evaluated on each
 * staged tile of the dataset, their candidates share one sort / verification / chain / select pipeline, and the host
 * synchronises once.  Results as pm_search_batch: hits[offsets[i] .. offsets[i+1]) belong to patterns[i], each list
 * bit-identical to pm_search of that pattern.  A page-locked `hits` buffer (pm_host_alloc) is filled directly by
 * the device-to-host copy.  On PM_ERR_OVERFLOW offsets[] are valid and the list is still on the device (pm_last_hits). */
int pm_search_request(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns, const char *kopt,
                      pm_hit *hits, int64_t cap, int64_t *offsets);

/* The same request restricted to the buffer fills that START in [pos_beg, pos_end) (see pm_search_fills_device), fully
 * asynchronous on the engine's stream: NO host synchronisation.  dev_out (device memory, out_rows rows of 2 x int64)
 * receives a header followed by the hits in output order:
 *   row 0: hits, candidates          row 1: placeholders, 0          then ceil(npat / 2) rows of per-pattern hit counts;
 *   hits start at row PM_REQUEST_HEADER_ROWS(npat).
 * sort_cap = candidate capacity of this call (0: engine default).  The caller inspects the header after its own
 * synchronisation (e.g. after the all-gather + D2H of a multi-GPU merge): candidates > sort_cap or
 * hits > out_rows - header rows means the call has to be repeated with more room. */
#define PM_REQUEST_HEADER_ROWS(npat) ((4 + (npat) + 1) / 2)
int pm_request_fills_device(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns, const char *kopt,
                            int64_t pos_beg, int64_t pos_end, int64_t sort_cap, void *dev_out, int64_t out_rows);

/* Multi-GPU requests, after the all-gather of the per-rank blocks pm_request_fills_device wrote: dev_all = `world` blocks
 * of `rows` rows in rank (= file) order.  Builds, on the device and without the host looking at any count,
 * dev_out = [the world header blocks, verbatim | pattern 0: hits of rank 0, rank 1, ... | pattern 1: ...], i.e. the
 * per-pattern lists of the whole file, ready for one device-to-host copy.  Asynchronous on the engine's stream.  At most
 * 64 patterns and 16 ranks (else PM_ERR_UNSUPPORTED: merge on the host). */
int pm_merge_request_shards(pm_engine *e, const void *dev_all, int world, int64_t rows, int npat, void *dev_out, int64_t out_rows);

/* Sharded search for multi-GPU runs: only candidates whose anchor position lies in
 * [pos_beg, pos_end) are produced; they are verified but NOT chained.  The caller
 * gathers the per-rank lists (already sorted) and calls pm_resolve on one rank. */
typedef struct { int64_t key, beg, end, reach; } pm_candidate;
int pm_candidates(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt,
                  int64_t pos_beg, int64_t pos_end,
                  pm_candidate *cands, int64_t cap, int64_t *ncands);
int pm_resolve(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt,
               const pm_candidate *cands, int64_t ncands,
               pm_hit *hits, int64_t cap, int64_t *nhits);

/* the same two calls with the candidate array in DEVICE memory of e's device (e.g. a torch tensor
 * that NCCL gathers into): nothing but the final hit list crosses PCIe */
int pm_candidates_device(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt,
                         int64_t pos_beg, int64_t pos_end,
                         pm_candidate *dev_cands, int64_t cap, int64_t *ncands);
int pm_resolve_device(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt,
                      const pm_candidate *dev_cands, int64_t ncands,
                      pm_hit *hits, int64_t cap, int64_t *nhits);

/* Fill-sharded search (the multi-GPU path bench.py uses): nrgrep_coords restarts its scan at every buffer fill
 * (recSearchFile @402298) and no hit crosses a fill, so fills are independent.  Searches the fills that START in
 * [pos_beg, pos_end) completely and leaves their hits, in output order, in DEVICE memory; ranks covering the
 * file with contiguous ranges get the whole hit list by concatenation. */
int pm_search_fills_device(pm_engine *e, pm_dataset *d, const char *pattern, const char *kopt,
                           int64_t pos_beg, int64_t pos_end, pm_hit *dev_hits, int64_t cap, int64_t *nhits,
                           int64_t *dev_count /* optional: the count is also stored here, in device memory */);

/* page-locked host buffers for large hit arrays (optional; any host pointer works for `hits`) */
void *pm_host_alloc(int64_t bytes);
void pm_host_free(void *p);

int pm_get_stats(pm_engine *e, pm_stats *out);

#ifdef __cplusplus
}
#endif
#endif
