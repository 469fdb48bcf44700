// request.cuh -- one PatMatch request = ONE pipeline (included at the end of engine.cu).
//
// The reference answers a "Both strands" request with two nrgrep_coords runs, one for the pattern and one for its
// reverse complement (www/FlaskApp/FlaskApp/patmatch.py:733-735 and :739-743), each reading the data file again.
// Here all patterns of a request (same -k) go through the dataset together:
//
//   scan     every pattern is evaluated on the staged tile of the 2-bit planes in ONE launch where the plans allow it
//            (k_scan_packed_exact / k_scan_apx with up to EX_MAXPAT patterns; other plan types add their own launch);
//            candidate keys carry the pattern id:  pid << 40 | position << 4 | piece
//   sort     one radix sort: candidates come out grouped by pattern, in the order the reference meets them
//   verify   one k_verify_req (plan and tables looked up by pattern id)
//   chain    one k_chain_req: the restart rule per dependency cluster; clusters never span patterns
//   select   one cub select; a header row block in front of the hit list carries the counts
//
// Nothing is read back between the stages: the candidate buffers have a capacity `cap` (adapted from the previous
// request), unused slots hold an all-ones key that sorts last, and the later stages read the true candidate count
// from device memory.  The host synchronises ONCE, on the copy of header + hit list; only a request whose
// candidates exceed `cap` is run again with a larger one.  In device mode (pm_request_fills_device, multi-GPU) the
// call does not synchronise at all: header and hits stay in device memory, ready to be all-gathered.

struct ReqPat {                         // device-resident description of one pattern of the request
    DevPlan pl;
    const unsigned long long *B, *TL, *TR;
    int recheck;                        // candidates come from the packed scan: re-check the trigger on the raw bytes
    int pad;
};

#define REQ_HDR_FIXED 4                 // header words: [0] hits, [1] candidates (raw count), [2] placeholders, [3] reserved, [4 + p] hits of pattern p

__device__ __forceinline__ long long req_nvalid(const unsigned long long *__restrict__ hdr, long long cap)
{
    const long long raw = (long long)hdr[1];
    return (raw < cap ? raw : cap) - (long long)hdr[2];
}

__global__ void __launch_bounds__(128) k_verify_req(const ReqPat *__restrict__ pats, const unsigned char *__restrict__ text, long long n,
                                                    const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ hdr,
                                                    long long cap, Cand *__restrict__ out, const Fills fills)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= req_nvalid(hdr, cap)) return;
    const unsigned long long key = keys[j];
    const ReqPat &rp = pats[key_pid(key)];
    const DevPlan &pl = rp.pl;
    const long long pos = key_pos(key);
    const int i = (int)(key & 15);
    Cand c;
    c.key = (long long)key;
    const long long anchor = anchor_of(pl, (long long)key);
    const long long p = locus_of(pl, anchor);
    const int f = fill_of(fills, p);
    const long long S = fills.S[f], E = fills.E[f];
    const long long wlen = pl.type == PM_PLAN_SIMPLE ? pl.m : (pl.type == PM_PLAN_SPLIT || plan_is_ext(pl)) ? pl.L
                         : pl.type == PM_PLAN_BWD ? pl.L - pl.k : 0;
    if (pos + wlen > E || (plan_is_ext(pl) && pos < S) || (rp.recheck && !raw_trigger(pl, rp.B, text, n, pos, i))) {
        c.beg = -1; c.end = -1; c.reach = anchor;
    } else if (plan_is_ext(pl)) {
        long long b = -1, e = -1, r = anchor;
        if (!check_match_ext(pl, text, E, rp.TL, rp.TR, anchor, S, &b, &e, &r)) { b = -1; e = -1; }
        c.beg = b; c.end = e; c.reach = r;
    } else if (pl.type == PM_PLAN_SIMPLE) {
        bool ok = true;
        for (int jj = 64; jj < pl.m && ok; jj++) {                                  // positions beyond the scanned window
            const unsigned ch = text[pos + jj];
            ok = (rp.TL[(size_t)(jj - 64) * 4 + (ch >> 6)] >> (ch & 63)) & 1ULL;
        }
        if (ok) { c.beg = pos; c.end = pos + pl.m; c.reach = pos - (pl.start_line ? 1 : 0); }
        else { c.beg = -1; c.end = -1; c.reach = pos; }
    } else {
        long long b = -1, e = -1, r = pos;
        if (!check_match(pl, text, E, rp.TL, rp.TR, i, pos, S, &b, &e, &r)) { b = -1; e = -1; }
        c.beg = b; c.end = e; c.reach = r;
    }
    out[j] = c;
}

// inputs of the two scans that '*' / '+' patterns need (hits of any length): hit ends in list order, leftmost examined
// bytes in reversed order.  Values are offset by the pattern id so that a scan never carries across patterns.
#define REQ_SEG (1LL << 41)
__global__ void k_cand_ends_req(const ReqPat *__restrict__ pats, const Cand *__restrict__ cands, const unsigned long long *__restrict__ hdr,
                                long long cap, long long *__restrict__ ends, long long *__restrict__ deps_rev)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cap) return;
    const long long nv = req_nvalid(hdr, cap);
    if (j >= nv) {
        ends[j] = -(1LL << 62);
        deps_rev[j] = (1LL << 62);                         // slots nv .. cap-1 of the reversed list: neutral for a minimum
        return;
    }
    const Cand c = cands[j];
    const int pid = key_pid((unsigned long long)c.key);
    const DevPlan &pl = pats[pid].pl;
    long long en = (c.beg < 0 && pl.start_line && c.reach < anchor_of(pl, c.key)) ? (REQ_SEG >> 1) : c.end;
    if (en < 0) en = 0;
    ends[j] = en + pid * REQ_SEG;
    long long dl = dep_lo(pl, c);
    if (dl < 0) dl = 0;
    deps_rev[nv - 1 - j] = dl + pid * REQ_SEG;
}

// one warp per 32 consecutive candidates: the lanes find the cluster heads among them, the warp resolves those clusters
// (engine.cu: chain_cluster_warp); clusters never span patterns
#define REQ_BLOCK_PIDS 32                // patterns whose hit counts a block adds up in shared memory before touching the header
__global__ void __launch_bounds__(128) k_chain_req(const ReqPat *__restrict__ pats, const unsigned char *__restrict__ text, long long n,
                                                   const Cand *__restrict__ cands, unsigned long long *__restrict__ hdr, long long cap,
                                                   pm_hit *__restrict__ hits, unsigned char *__restrict__ sel, const Fills fills,
                                                   const long long *__restrict__ maxend, const long long *__restrict__ mindep_rev)
{
    // per-pattern hit counts: one shared-memory add per cluster, one global atomic per block and pattern.  (One global
    // atomic per cluster -- 7e4 of them on two addresses for the bench request -- serialised in L2 and WAS the kernel:
    // 113 us, 56 % of the stall samples on the header line; profiles/r02_verify_chain.txt)
    __shared__ unsigned blk_cnt[REQ_BLOCK_PIDS];
    if (threadIdx.x < REQ_BLOCK_PIDS) blk_cnt[threadIdx.x] = 0;
    __syncthreads();
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long ncand = req_nvalid(hdr, cap);
    const long long wbase = j - (threadIdx.x & 31);
    auto count = [&](int pid, unsigned long long nsel) {
        if (pid < REQ_BLOCK_PIDS) atomicAdd(&blk_cnt[pid], (unsigned)nsel);
        else atomicAdd(hdr + REQ_HDR_FIXED + pid, nsel);
    };
    if (wbase < ncand) {                                      // warp-uniform
        bool head = false;
        if (j < ncand) head = cand_opens_cluster(pats[key_pid((unsigned long long)cands[j].key)].pl, cands, ncand, j, fills, maxend, mindep_rev);
        bool longc = false;
        {
            // short clusters: walked by the lane of their first candidate
            int pid = 0, nsel = 0;
            if (head) {
                pid = key_pid((unsigned long long)cands[j].key);
                const ReqPat &rp = pats[pid];
                nsel = chain_cluster_thread(rp.pl, text, rp.TL, rp.TR, cands, ncand, j, hits, sel, fills, maxend, mindep_rev);
                longc = nsel < 0;
            }
            if (head && nsel > 0) count(pid, (unsigned long long)nsel);
        }
        unsigned heads = __ballot_sync(0xffffffffu, longc);
        while (heads) {
            const int h = __ffs(heads) - 1;
            heads &= heads - 1;
            const long long j0 = wbase + h;
            const int pid = key_pid((unsigned long long)cands[j0].key);
            const ReqPat &rp = pats[pid];
            const unsigned long long nsel = chain_cluster_warp(rp.pl, text, rp.TL, rp.TR, cands, ncand, j0, hits, sel, fills, maxend, mindep_rev);
            if (nsel && (threadIdx.x & 31) == 0) count(pid, nsel);
        }
    }
    __syncthreads();
    if (threadIdx.x < REQ_BLOCK_PIDS && blk_cnt[threadIdx.x]) atomicAdd(hdr + REQ_HDR_FIXED + threadIdx.x, (unsigned long long)blk_cnt[threadIdx.x]);
    (void)n;
}

struct Request {
    std::vector<Compiled> comp;
    std::vector<long long> a0, a1;       // anchor range per pattern
    int k = 0;
};

static int req_hdr_rows(int npat) { return (REQ_HDR_FIXED + npat + 1) / 2; }

static int compile_request(int npat, const char *const *patterns, const char *kopt, Request &rq)
{
    rq.comp.resize((size_t)npat);
    for (int p = 0; p < npat; p++) {
        if (!patterns[p]) { g_err = "pattern is NULL"; return PM_ERR_ARG; }
        int rc = compile(patterns[p], kopt, rq.comp[p], true);
        if (rc) return rc;
    }
    rq.a0.assign((size_t)npat, 0);
    rq.a1.assign((size_t)npat, 0);
    return PM_OK;
}

// tables of every pattern + the ReqPat array, one blob, one copy
static int upload_request(pm_engine *e, pm_dataset *d, const Request &rq, const ReqPat **d_pats)
{
    const int npat = (int)rq.comp.size();
    size_t words = 0;
    for (const Compiled &c : rq.comp) words += 256 + 2 * c.vt.TL.size();
    const size_t pats_off = words * 8;
    const size_t total = pats_off + (size_t)npat * sizeof(ReqPat);
    int rc = e->tables.reserve(total + 64);
    if (rc) return rc;
    std::vector<unsigned char> &blob = e->h_blob;
    blob.assign(total, 0);
    unsigned long long *base = (unsigned long long *)e->tables.p;
    size_t off = 0;
    for (int p = 0; p < npat; p++) {
        const Compiled &c = rq.comp[p];
        const size_t nv = c.vt.TL.size();
        memcpy(blob.data() + off * 8, c.ft.B, 256 * 8);
        if (nv) {
            memcpy(blob.data() + (off + 256) * 8, c.vt.TL.data(), nv * 8);
            memcpy(blob.data() + (off + 256 + nv) * 8, c.vt.TR.data(), nv * 8);
        }
        ReqPat rp;
        memset(&rp, 0, sizeof rp);
        rp.pl = c.dp;
        rp.B = base + off; rp.TL = base + off + 256; rp.TR = base + off + 256 + nv;
        rp.recheck = (scan_uses_packed(e, d, c) || scan_uses_pep5(e, d, c)) ? 1 : 0;
        memcpy(blob.data() + pats_off + (size_t)p * sizeof(ReqPat), &rp, sizeof rp);
        off += 256 + 2 * nv;
    }
    CK(cudaMemcpyAsync(e->tables.p, blob.data(), total, cudaMemcpyHostToDevice, e->stream));
    *d_pats = (const ReqPat *)((char *)e->tables.p + pats_off);
    return PM_OK;
}

// all scan launches of a request; patterns whose plans allow it share a launch
static int scan_request(pm_engine *e, pm_dataset *d, const Request &rq, const Fills &fills, const ReqPat *d_pats, const ScanTarget &tgt)
{
    const int npat = (int)rq.comp.size();
    const unsigned long long bad = (unsigned long long)npat << PM_PID_SHIFT;
    std::vector<ExactPat> ex;
    std::vector<ApxPat> ax;
    int rc;
    auto flush_exact = [&]() -> int {
        int r = ex.empty() ? PM_OK : launch_exact(e, d, ex.data(), (int)ex.size(), bad, tgt);
        ex.clear();
        return r;
    };
    auto flush_apx = [&]() -> int {
        int r = ax.empty() ? PM_OK : launch_apx(e, d, ax.data(), (int)ax.size(), tgt);
        ax.clear();
        return r;
    };
    for (int p = 0; p < npat; p++) {
        const Compiled &cf = rq.comp[p];
        const Compiled &c = cf.scan ? *cf.scan : cf;
        const unsigned long long tag = (unsigned long long)p << PM_PID_SHIFT;
        const long long a0 = rq.a0[p], a1 = rq.a1[p];
        if (a1 <= a0) continue;
        if (scan_uses_packed(e, d, cf) && c.dp.type == PM_PLAN_SIMPLE) {
            ExactPat pt;
            fill_exact_pat(c, a0, a1, d->n, tag, pt);
            ex.push_back(pt);
            if ((int)ex.size() == EX_MAXPAT && (rc = flush_exact())) return rc;
        } else if (apx_eligible(e, d, c)) {
            ApxPat ap;
            build_apx_pat(e->qgram_filter != 0, jit_wanted(e, a1 - a0), c, a0, std::min(a1, d->n - c.dp.L + 1), tag, ap);
            ax.push_back(ap);
            if ((int)ax.size() == EX_MAXPAT && (rc = flush_apx())) return rc;
        } else {
            // tables of this pattern inside the uploaded blob (same layout as upload_request)
            size_t off = 0;
            for (int q = 0; q < p; q++) off += 256 + 2 * rq.comp[q].vt.TL.size();
            const unsigned long long *base = (const unsigned long long *)e->tables.p;
            const size_t nv = cf.vt.TL.size();
            if ((rc = launch_scan(e, d, cf, a0, a1, fills, base + off, base + off + 256, base + off + 256 + nv, tgt, tag, bad))) return rc;
        }
    }
    if ((rc = flush_exact())) return rc;
    if ((rc = flush_apx())) return rc;
    (void)d_pats;
    return PM_OK;
}

static int sort_end_bit(int npat)
{
    const unsigned long long top = ((unsigned long long)npat << PM_PID_SHIFT);     // the placeholder key; unused slots are all ones
    int end_bit = PM_PID_SHIFT + 1;
    while (end_bit < 64 && (top >> end_bit)) end_bit++;
    return end_bit;
}

// Runs the pipeline on e->stream with candidate capacity `cap`; leaves header + hits in e->hits2 (header first).
// No host synchronisation.
static int enqueue_request(pm_engine *e, pm_dataset *d, const Request &rq, long long cap)
{
    const int npat = (int)rq.comp.size();
    const int hrows = req_hdr_rows(npat);
    int rc;
    if (cap >= (1LL << 31) - 1) { g_err = "more than 2^31 candidates: not supported"; return PM_ERR_UNSUPPORTED; }
    Fills fills;
    if ((rc = ensure_fills(e, d, &fills))) return rc;
    const ReqPat *d_pats = nullptr;
    if ((rc = upload_request(e, d, rq, &d_pats))) return rc;
    if ((rc = e->keys.reserve((size_t)cap * 8))) return rc;
    if ((rc = e->keys2.reserve((size_t)cap * 8))) return rc;
    if ((rc = e->cands.reserve((size_t)cap * sizeof(Cand)))) return rc;
    if ((rc = e->hits.reserve((size_t)cap * sizeof(pm_hit)))) return rc;
    if ((rc = e->hits2.reserve((size_t)(cap + hrows) * sizeof(pm_hit)))) return rc;
    if ((rc = e->sel.reserve((size_t)cap))) return rc;
    unsigned long long *hdr = (unsigned long long *)e->hits2.p;
    pm_hit *out_hits = (pm_hit *)e->hits2.p + hrows;
    CK(cudaMemsetAsync(hdr, 0, (size_t)hrows * sizeof(pm_hit), e->stream));
    CK(cudaMemsetAsync(e->keys.p, 0xff, (size_t)cap * 8, e->stream));
    CK(cudaMemsetAsync(e->sel.p, 0, (size_t)cap, e->stream));
    CK(cudaEventRecord(e->ev[0], e->stream));
    e->stats.scan_bytes = 0; e->stats.scan_bases = 0;
    const ScanTarget tgt{(unsigned long long *)e->keys.p, hdr + 1, cap};
    if ((rc = scan_request(e, d, rq, fills, d_pats, tgt))) return rc;
    CK(cudaEventRecord(e->ev[1], e->stream));
    {
        size_t tmp = 0;
        const int end_bit = sort_end_bit(npat);
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, (unsigned long long *)e->keys.p, (unsigned long long *)e->keys2.p, (int)cap, 0, end_bit, e->stream));
        if ((rc = e->cubtmp.reserve(tmp))) return rc;
        CK(cub::DeviceRadixSort::SortKeys(e->cubtmp.p, tmp, (unsigned long long *)e->keys.p, (unsigned long long *)e->keys2.p, (int)cap, 0, end_bit, e->stream));
        e->stats.launches += 3;
    }
    const unsigned long long *keys = (const unsigned long long *)e->keys2.p;
    CK(cudaEventRecord(e->ev[2], e->stream));
    const unsigned nblk = (unsigned)((cap + 127) / 128);
    k_verify_req<<<nblk, 128, 0, e->stream>>>(d_pats, d->d_text, d->n, keys, hdr, cap, (Cand *)e->cands.p, fills);
    CK(cudaGetLastError());
    e->stats.launches++;
    CK(cudaEventRecord(e->ev[3], e->stream));
    const long long *d_maxend = nullptr, *d_mindep = nullptr;
    bool repeats = false;
    for (const Compiled &c : rq.comp) repeats = repeats || c.dp.ext_repeats;
    if (repeats) {
        // the sorted keys are no longer needed: keys (the unsorted buffer) and a fresh buffer hold the scan arrays
        if ((rc = e->scanbuf.reserve((size_t)cap * 32))) return rc;
        long long *ends = (long long *)e->scanbuf.p, *mx = ends + cap, *deps = mx + cap, *mn = deps + cap;
        k_cand_ends_req<<<(unsigned)((cap + 255) / 256), 256, 0, e->stream>>>(d_pats, (const Cand *)e->cands.p, hdr, cap, ends, deps);
        size_t tmpb = 0, tmpc = 0;
        CK(cub::DeviceScan::InclusiveScan(nullptr, tmpb, ends, mx, MaxLL(), (int)cap, e->stream));
        CK(cub::DeviceScan::InclusiveScan(nullptr, tmpc, deps, mn, MinLL(), (int)cap, e->stream));
        tmpb = std::max(tmpb, tmpc);
        if ((rc = e->cubtmp.reserve(tmpb))) return rc;
        CK(cub::DeviceScan::InclusiveScan(e->cubtmp.p, tmpb, ends, mx, MaxLL(), (int)cap, e->stream));
        CK(cub::DeviceScan::InclusiveScan(e->cubtmp.p, tmpb, deps, mn, MinLL(), (int)cap, e->stream));
        d_maxend = mx; d_mindep = mn;
        e->stats.launches += 3;
    }
    k_chain_req<<<nblk, 128, 0, e->stream>>>(d_pats, d->d_text, d->n, (const Cand *)e->cands.p, hdr, cap, (pm_hit *)e->hits.p,
                                            (unsigned char *)e->sel.p, fills, d_maxend, d_mindep);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e->ev[4], e->stream));
    {
        size_t tmp = 0;
        long long *d_nsel = (long long *)hdr;
        CK(cub::DeviceSelect::Flagged(nullptr, tmp, (H16 *)e->hits.p, (unsigned char *)e->sel.p, (H16 *)out_hits, d_nsel, (int)cap, e->stream));
        if ((rc = e->cubtmp.reserve(tmp))) return rc;
        CK(cub::DeviceSelect::Flagged(e->cubtmp.p, tmp, (H16 *)e->hits.p, (unsigned char *)e->sel.p, (H16 *)out_hits, d_nsel, (int)cap, e->stream));
        e->stats.launches += 3;
    }
    CK(cudaEventRecord(e->ev[5], e->stream));
    e->stats_pending = true;
    e->last_hits = out_hits;
    return PM_OK;
}

static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// host-mode request: hits and per-pattern offsets in host memory, one synchronisation in the usual case
static int run_request_host(pm_engine *e, pm_dataset *d, const Request &rq, pm_hit *hits, int64_t cap_hits, int64_t *offsets)
{
    const int npat = (int)rq.comp.size();
    const int hrows = req_hdr_rows(npat);
    int rc;
    long long cap = std::max<long long>(e->req_cap_hint, 1 << 14);
    const bool direct = hits && cap_hits > 0 && is_pinned_host(hits);
    for (int attempt = 0;; attempt++) {
        e->stats = pm_stats{};
        if ((rc = enqueue_request(e, d, rq, cap))) return rc;
        // speculative copy: header + as many hits as the last request had (with head-room)
        long long spec = std::min<long long>(std::max<long long>(e->req_hits_hint, 4096), cap);
        if (!hits) spec = 0;
        if (direct) spec = std::min<long long>(spec, cap_hits);
        const size_t hbytes = (size_t)hrows * sizeof(pm_hit);
        const size_t need_stage = hbytes + (direct ? 0 : (size_t)spec * sizeof(pm_hit));
        if (e->h_stage_cap < need_stage) {
            if (e->h_stage) cudaFreeHost(e->h_stage);
            e->h_stage = nullptr; e->h_stage_cap = 0;
            const size_t want = std::max(need_stage * 2, (size_t)1 << 20);
            CK(cudaMallocHost(&e->h_stage, want));
            e->h_stage_cap = want;
        }
        CK(cudaMemcpyAsync(e->h_stage, e->hits2.p, hbytes, cudaMemcpyDeviceToHost, e->stream));
        pm_hit *stage_hits = (pm_hit *)((char *)e->h_stage + hbytes);
        if (spec > 0)
            CK(cudaMemcpyAsync(direct ? hits : stage_hits, e->last_hits, (size_t)spec * sizeof(pm_hit), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->stats.syncs++;
        const unsigned long long *hdr = (const unsigned long long *)e->h_stage;
        const long long nsel = (long long)hdr[0], raw = (long long)hdr[1], nplace = (long long)hdr[2];
        if (raw > cap) {
            if (attempt >= 3) { g_err = "candidate buffer keeps overflowing"; return PM_ERR_CUDA; }
            cap = raw + raw / 4 + 1024;
            continue;
        }
        e->req_cap_hint = std::max<long long>(raw + raw / 4 + 1024, 1 << 14);
        e->req_hits_hint = nsel + nsel / 4 + 256;
        e->stats.candidates = raw - nplace;
        e->stats.verified = raw - nplace;
        e->stats.hits = nsel;
        offsets[0] = 0;
        for (int p = 0; p < npat; p++) offsets[p + 1] = offsets[p] + (int64_t)hdr[REQ_HDR_FIXED + p];
        if (!hits || nsel == 0) return PM_OK;
        if (nsel > cap_hits) { g_err = "hit buffer too small"; return PM_ERR_OVERFLOW; }     // the list stays on the device (pm_last_hits)
        if (nsel > spec) {
            // more hits than guessed: fetch the rest
            if (direct) {
                CK(cudaMemcpyAsync(hits + spec, e->last_hits + spec, (size_t)(nsel - spec) * sizeof(pm_hit), cudaMemcpyDeviceToHost, e->stream));
                CK(cudaStreamSynchronize(e->stream));
            } else {
                memcpy(hits, stage_hits, (size_t)spec * sizeof(pm_hit));
                if ((rc = copy_to_host(e, hits + spec, e->last_hits + spec, (size_t)(nsel - spec) * sizeof(pm_hit)))) return rc;
            }
            e->stats.syncs++;
        } else if (!direct) memcpy(hits, stage_hits, (size_t)nsel * sizeof(pm_hit));
        return PM_OK;
    }
}

// anchor range of the fills f0 .. f1-1 for one compiled pattern (see search_fill_range)
static void fill_anchor_range(const pm_dataset *d, const Compiled &c, long long f0, long long f1, long long *a0, long long *a1)
{
    const std::vector<long long> &S = d->fill_starts;
    if (f0 >= f1) { *a0 = 0; *a1 = 0; return; }
    const long long shift = c.dp.type == PM_PLAN_FWD ? 1
                          : (c.dp.type == PM_PLAN_EXT_BEG || c.dp.type == PM_PLAN_EXT_END) ? (c.dp.type == PM_PLAN_EXT_END ? 1 : 0) - c.dp.ext_off : 0;
    *a0 = std::max<long long>(S[f0] + shift, 0);
    // while pm_search_stream is still uploading, the table ends at the last fill that is complete: nothing beyond it
    // has been packed yet
    *a1 = f1 < (long long)S.size() ? S[f1] + shift : d->fills_complete ? d->n + 1 : std::min<long long>(d->fill_ends[(size_t)f1 - 1] + 1 + shift, d->n + 1);
}

// the patterns of one request over the buffer fills f0 .. f1-1 (host-mode results); engine mutex held by the caller
static int search_request_range(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns, const char *kopt,
                                long long f0, long long f1, pm_hit *hits, int64_t cap, int64_t *offsets)
{
    Request rq;
    int rc = compile_request(npat, patterns, kopt, rq);
    if (rc) return rc;
    Fills fills;
    if ((rc = ensure_fills(e, d, &fills))) return rc;
    for (int p = 0; p < npat; p++) fill_anchor_range(d, rq.comp[p], f0, f1, &rq.a0[p], &rq.a1[p]);
    return run_request_host(e, d, rq, hits, cap, offsets);
}

int pm_search_request(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns, const char *kopt,
                      pm_hit *hits, int64_t cap, int64_t *offsets)
{
    if (!e || !d || npat < 1 || npat >= (1 << 20) || !patterns || !offsets || !kopt || d->e != e) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    CK(cudaSetDevice(e->device));
    (void)cudaGetLastError();
    Request rq;
    int rc = compile_request(npat, patterns, kopt, rq);
    if (rc) return rc;
    if (d->windowed) { g_err = "windowed dataset: only pm_request_fills_device can search it"; return PM_ERR_ARG; }
    for (int p = 0; p < npat; p++) { rq.a0[p] = 0; rq.a1[p] = d->n + 1; }
    return run_request_host(e, d, rq, hits, cap, offsets);
}

int pm_request_fills_device(pm_engine *e, pm_dataset *d, int npat, const char *const *patterns, const char *kopt,
                            int64_t pos_beg, int64_t pos_end, int64_t sort_cap, void *dev_out, int64_t out_rows)
{
    if (!e || !d || npat < 1 || npat >= (1 << 20) || !patterns || !kopt || d->e != e || !dev_out || out_rows < req_hdr_rows(npat)) { g_err = "bad argument"; return PM_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    CK(cudaSetDevice(e->device));
    (void)cudaGetLastError();
    Request rq;
    int rc = compile_request(npat, patterns, kopt, rq);
    if (rc) return rc;
    Fills fills;
    if ((rc = ensure_fills(e, d, &fills))) return rc;
    const std::vector<long long> &S = d->fill_starts;
    const long long f0 = std::lower_bound(S.begin(), S.end(), (long long)pos_beg) - S.begin();
    const long long f1 = std::lower_bound(S.begin(), S.end(), (long long)pos_end) - S.begin();
    if (d->windowed && f1 > f0 && (S[(size_t)f0] < d->win_lo || d->fill_ends[(size_t)f1 - 1] > d->win_hi)) {
        g_err = "the buffer fills of this position range are not inside the dataset's window";
        return PM_ERR_ARG;
    }
    for (int p = 0; p < npat; p++) fill_anchor_range(d, rq.comp[p], f0, f1, &rq.a0[p], &rq.a1[p]);
    const long long cap = std::max<long long>(sort_cap > 0 ? sort_cap : e->req_cap_hint, 1 << 12);
    e->stats = pm_stats{};
    if ((rc = enqueue_request(e, d, rq, cap))) return rc;
    const long long rows = std::min<long long>(out_rows, cap + req_hdr_rows(npat));
    CK(cudaMemcpyAsync(dev_out, e->hits2.p, (size_t)rows * sizeof(pm_hit), cudaMemcpyDeviceToDevice, e->stream));
    return PM_OK;
}

// ---- merge of the all-gathered per-rank blocks (multi-GPU requests) --------------------------------------------------
// `all` holds `world` blocks of `rows` rows, each [header rows | hits of pattern 0 | hits of pattern 1 | ...] as written
// by pm_request_fills_device, in rank (= file) order.  The per-pattern lists of the whole file are the ranks' lists one
// after the other: out = [the world headers, verbatim | pattern 0: rank 0, rank 1, ... | pattern 1: ...].  Everything the
// kernel needs is in the headers it reads from device memory, so the host can enqueue the merge and the result copy
// right behind the collective without looking at the counts first.
#define REQ_MERGE_MAXPAT 64
#define REQ_MERGE_MAXWORLD 16
__global__ void __launch_bounds__(256) k_merge_request_shards(const pm_hit *__restrict__ all, int world, long long rows, int npat, int hrows,
                                                              pm_hit *__restrict__ out, long long out_hits_cap)
{
    __shared__ long long s_src[REQ_MERGE_MAXWORLD][REQ_MERGE_MAXPAT + 1];    // first row (inside the rank's block) of pattern p's hits
    __shared__ long long s_dst[REQ_MERGE_MAXWORLD][REQ_MERGE_MAXPAT];        // first output row of the rank's hits of pattern p
    __shared__ long long s_nh[REQ_MERGE_MAXWORLD];
    if (threadIdx.x == 0) {
        long long at = 0;
        for (int r = 0; r < world; r++) {
            const unsigned long long *hdr = (const unsigned long long *)(all + (size_t)r * rows);
            long long nh = (long long)hdr[0];
            if (nh > rows - hrows) nh = rows - hrows;             // a block that did not fit: the host sees it in the header and retries
            s_nh[r] = nh;
            long long a = 0;
            for (int p = 0; p < npat; p++) { s_src[r][p] = a; a += (long long)hdr[REQ_HDR_FIXED + p]; }
            s_src[r][npat] = a;
        }
        for (int p = 0; p < npat; p++)
            for (int r = 0; r < world; r++) { s_dst[r][p] = at; at += s_src[r][p + 1] - s_src[r][p]; }
    }
    __syncthreads();
    const long long per = rows;                                    // thread index space: world x rows
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)world * per; t += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(t / per);
        const long long i = t % per;
        if (i < hrows) { out[(size_t)r * hrows + i] = all[(size_t)r * rows + i]; continue; }      // headers, verbatim
        const long long h = i - hrows;
        if (h >= s_nh[r]) continue;
        int p = 0;
        while (p + 1 < npat && h >= s_src[r][p + 1]) p++;
        const long long d = s_dst[r][p] + (h - s_src[r][p]);
        if (d < out_hits_cap) out[(size_t)world * hrows + d] = all[(size_t)r * rows + i];
    }
}

int pm_merge_request_shards(pm_engine *e, const void *dev_all, int world, int64_t rows, int npat, void *dev_out, int64_t out_rows)
{
    if (!e || !dev_all || !dev_out || world < 1 || npat < 1 || rows < req_hdr_rows(npat) || out_rows < (int64_t)world * req_hdr_rows(npat)) { g_err = "bad argument"; return PM_ERR_ARG; }
    if (npat > REQ_MERGE_MAXPAT || world > REQ_MERGE_MAXWORLD) { g_err = "merge kernel: at most 64 patterns and 16 ranks"; return PM_ERR_UNSUPPORTED; }
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    CK(cudaSetDevice(e->device));
    (void)cudaGetLastError();
    const int hrows = req_hdr_rows(npat);
    const long long total = (long long)world * rows;
    const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)e->sms * 8));
    k_merge_request_shards<<<grid, 256, 0, e->stream>>>((const pm_hit *)dev_all, world, rows, npat, hrows, (pm_hit *)dev_out, out_rows - (long long)world * hrows);
    CK(cudaGetLastError());
    return PM_OK;
}

// host-only: the source apx_jit.cpp would hand to NVRTC for this request
int64_t pm_jit_source(int npat, const char *const *patterns, const char *kopt, char *buf, int64_t cap)
{
    if (npat < 1 || npat > EX_MAXPAT || !patterns || !kopt) { g_err = "bad argument"; return PM_ERR_ARG; }
    Request rq;
    int rc = compile_request(npat, patterns, kopt, rq);
    if (rc) return rc;
    std::vector<ApxPat> ax((size_t)npat);
    for (int p = 0; p < npat; p++) {
        const Compiled &c = rq.comp[p];
        const DevPlan &dp = c.dp;
        if (c.scan || dp.type != PM_PLAN_SPLIT || dp.k < 1 || dp.k > 3 || dp.m + 2 * dp.k > 64 || dp.npieces > 4 || !plain_triggers(dp)) {
            g_err = "no specialised kernel for this plan"; return PM_ERR_UNSUPPORTED;
        }
        build_apx_pat(true, true, c, 0, 1LL << 40, (unsigned long long)p << PM_PID_SHIFT, ax[(size_t)p]);
    }
    const std::string src = apx_full_source(apx_generate_prefix(ax.data(), npat));
    if (buf && cap > 0) {
        const size_t ncopy = std::min<size_t>(src.size(), (size_t)cap - 1);
        memcpy(buf, src.data(), ncopy);
        buf[ncopy] = 0;
    }
    return (int64_t)src.size() + 1;
}
