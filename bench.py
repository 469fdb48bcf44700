#!/usr/bin/env python
"""bench.py -- headline benchmark of the PatMatch hot path on B200.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                   (the reference engine on the host cores)

Workload (BASELINE.json configs[4]): a single 2-mismatch (-k 2ids) degenerate DNA motif searched on
both strands of a synthetic 3.1 Gb, 24-chromosome human-shaped genome; with N ranks every rank scans 1/N
of the file positions (strong scaling) and rank 0 merges the verified candidates.  One step = one PatMatch request = two engine searches (motif and reverse complement,
patmatch.py:733-743) over the whole genome.  metric = pattern.Gbases/s = 2 * genome bases / step time.

  value : dataset resident in HBM, device time by CUDA events on the launching stream
  e2e   : same request through the C ABI from pinned HOST buffers: the genome is copied
          host->device inside the timed region every step and the hit list comes back to the host

Outside the timed region the benched hit list is checked against the CPU oracle in sampled windows (around planted
copies of the motif and around random hits, both strands): `parity_windows_ok`; a disagreement aborts the run.
`secondary` carries the other BASELINE configs (configs[0] on 12 Mb and on 3.1 Gb, [1], [2], [3]), each with its own
kernel time and roofline fraction; with N > 1 ranks only configs[3] is repeated (text-sharded; the motif-sharded split beside it).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MOTIF = "TGASTCANNNRYGATAAG"          # 18-nt degenerate motif (AP-1 site + spacer + GATA site), 2 errors
MISMATCH = 2
# human-shaped chromosome lengths (Mb), scaled so that the total is --bases
CHROM_MB = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]
ZERO_ENV = dict(os.environ, GLIBC_TUNABLES="glibc.malloc.tcache_count=0:glibc.malloc.perturb=255")


def chrom_lengths(total):
    s = float(sum(CHROM_MB))
    return [max(1000, int(total * mb / s)) for mb in CHROM_MB]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML every 5 ms from a thread
    (nvidia-smi -lms as the fallback: its first sample alone can take longer than a short run)."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.samples, self.proc, self.stop_flag, self.thread, self.max_mhz = index, [], None, False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append([str(mhz), str(self.max_mhz)] + ["Active" if mask & bit else "Not Active" for bit, _ in self.REASONS])
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def patterns():
    from patmatchdocker_b200 import patmatch as host
    conv, comp, opt = host.process_pattern(MOTIF, "dna", "Both strands", None, None, None, MISMATCH)
    return [conv, comp], opt


def make_genome_torch(lengths, mine, device):
    """this rank's chromosomes as one .seq byte tensor on the device (one sequence per line)"""
    import torch
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    parts = []
    for i in mine:
        g = torch.Generator(device=device)
        g.manual_seed(1000 + i)
        head = (">chr%d synthetic human-shaped\n" % (i + 1)).encode()
        parts.append(torch.tensor(list(head), dtype=torch.uint8, device=device))
        idx = torch.randint(0, 4, (lengths[i],), generator=g, device=device, dtype=torch.uint8)
        parts.append(lut[idx.long()] if lengths[i] < (1 << 28) else torch.cat([lut[c.long()] for c in idx.split(1 << 27)]))
        parts.append(torch.tensor([10], dtype=torch.uint8, device=device))
        del idx
    return torch.cat(parts)


def python_fills(newlines, n, bufsize=1600000):
    """Buffer fills [S, E) of the reference for a file of n bytes (engine.cu: compute_fills; bufLoad @41bbf0):
    a fill that does not reach EOF ends at its last newline and the next one starts AT that newline; without a usable
    newline the fill is scanned whole and the next starts right after it."""
    import bisect
    S, E = [], []
    s0 = 0
    while n - s0 > 0:
        dsize = min(bufsize, n - s0)
        if dsize < bufsize:
            en, nxt = s0 + dsize, n
        else:
            i = bisect.bisect_right(newlines, s0 + dsize - 1)
            p = newlines[i - 1] if i > 0 else -1
            if p > s0:
                en, nxt = p + 1, p
            else:
                en, nxt = s0 + dsize, s0 + dsize
        S.append(s0); E.append(en)
        s0 = nxt
    return S, E


def genome_layout(lengths):
    """(sequence start, sequence end) of every chromosome line and the newline positions of make_genome_torch's file"""
    pos, lines, newlines = 0, [], []
    for i, ln in enumerate(lengths):
        head = len(">chr%d synthetic human-shaped\n" % (i + 1))
        newlines.append(pos + head - 1)
        lines.append((pos + head, pos + head + ln))
        pos += head + ln + 1
        newlines.append(pos - 1)
    return lines, newlines, pos


def plant_sites(genome, lines, site, count, seed, maxerr=2):
    """Overwrite `count` random places of the device genome with copies of `site` carrying 0..maxerr random edits
    (substitutions, insertions, deletions), half of them reverse-complemented.  Returns the positions."""
    import random
    import torch
    rng = random.Random(seed)
    comp = str.maketrans("ACGT", "TGCA")
    where = []
    for _ in range(count):
        a, b = lines[rng.randrange(len(lines))]
        p = rng.randrange(a + 4000, b - 4000)
        s = list(site)
        for _e in range(rng.randint(0, maxerr)):
            q = rng.randrange(len(s))
            r = rng.randint(0, 2)
            if r == 0:
                s[q] = rng.choice("ACGT")
            elif r == 1:
                del s[q]
            else:
                s.insert(q, rng.choice("ACGT"))
        s = "".join(s)
        if rng.random() < 0.5:
            s = s[::-1].translate(comp)
        genome[p:p + len(s)] = torch.tensor(list(s.encode()), dtype=torch.uint8, device=genome.device)
        where.append(p)
    return where


def parity_windows(genome, lines, newlines, n, hit_lists, pats, kopt, centers, half=3000, margin=200):
    """The benched hit lists against the CPU oracle inside windows of +-half bytes around `centers`: a window stays
    inside one sequence line and one buffer fill; hits are compared away from the window edges (the oracle's scan
    starts at the window start, the product's wherever the previous hit ended).  -> (windows checked, windows equal)"""
    import bisect
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    S, _E = python_fills(newlines, n)
    checked = ok = 0
    for c in centers:
        li = bisect.bisect_right([a for a, _ in lines], c) - 1
        if li < 0:
            continue
        a, b = lines[li]
        lo, hi = max(a, c - half), min(b, c + half)
        k = bisect.bisect_right(S, lo)                         # first fill that starts after lo
        if k < len(S) and S[k] < hi:                           # a fill boundary inside: shrink to the side that holds c
            if S[k] <= c:
                lo = S[k]
            else:
                hi = S[k]
        if hi - lo < 4 * margin:
            continue
        text = genome[lo:hi].cpu().numpy().tobytes()
        good = True
        for pat, hits in zip(pats, hit_lists):
            want = [(x + lo, y + lo) for x, y in oracle_lib.search(pat, text, kopt)]
            i0, i1 = np.searchsorted(hits["beg"], lo), np.searchsorted(hits["beg"], hi)
            mine = [(int(x), int(y)) for x, y in zip(hits["beg"][i0:i1], hits["end"][i0:i1]) if y <= hi]
            inner = lambda hs: [h for h in hs if h[0] >= lo + margin and h[1] <= hi - margin]
            good = good and inner(mine) == inner(want)
        checked += 1
        ok += 1 if good else 0
    return checked, ok


def profile_constants():
    """Per-launch counters of the headline scan kernel from the committed ncu capture (profiles/r02_traffic.json):
    DRAM bytes and executed warp instructions per scanned base.  They cannot be measured without a profiler, so the
    bench line scales them to the launch it timed and says where they come from."""
    tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(tp):
        return None
    return json.load(open(tp)).get("k_scan_apx_jit")


def synth_lines(nlines, total, seed, alphabet=b"ACGT", name="chr"):
    """a .seq file of nlines sequences (one per line, '>name' header lines) as bytes, and its hosts numpy RNG"""
    rng = np.random.default_rng(seed)
    lut = np.frombuffer(alphabet, dtype=np.uint8)
    per = max(total // nlines, 8)
    parts = []
    for i in range(nlines):
        ln = int(per * (0.5 + rng.random())) if nlines > 64 else per
        parts.append((">%s%d\n" % (name, i + 1)).encode())
        parts.append(lut[rng.integers(0, len(lut), size=ln, dtype=np.uint8)].tobytes())
        parts.append(b"\n")
    return b"".join(parts)


def timed_requests(torch, eng, fn, steps, warmup):
    """wall-clock per step of fn() (each call synchronises on its own result copy), scan kernel ms and bytes per step"""
    import patmatchdocker_b200 as pm
    for i in range(max(warmup, 1)):
        fn()
        if i == 0:
            pm.jit_wait()            # the first request started the background compilation of its specialised kernel (if any)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    scan_ms = scan_bytes = 0
    out = None
    for _ in range(steps):
        out = fn()
        st = eng.stats()
        scan_ms += st["scan_ms"]; scan_bytes += st["scan_bytes"]
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    return ms, scan_ms / steps, scan_bytes // steps, out, st


def secondary_configs(args, torch, dist, pm, eng, dev, rank, world, big_ds, big_bases, peak):
    """The other BASELINE configs, each a whole request through the C ABI with the dataset resident, checked against the
    CPU oracle.  Returns the `secondary` object of the bench line."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from patmatchdocker_b200 import patmatch as host
    sec = {}

    def entry(workload, bases, npat, ms, kms, kbytes, kernel, hits, parity, bytes_per_base):
        ach = kbytes / (kms / 1e3) / 1e9 if kms > 0 else 0.0
        return {"workload": workload, "value": round(npat * bases / (ms / 1e3) / 1e9, 3), "unit": "pattern*Gbases/s", "ms_per_step": round(ms, 4),
                "kernel": kernel, "kernel_ms": round(kms, 4), "algorithmic_bytes": int(kbytes), "bytes_per_base": bytes_per_base,
                "achieved_GBps": round(ach, 1), "frac": round(ach / peak, 4), "hits_per_step": int(hits), "parity": parity}

    if rank == 0 and world == 1:
        # configs[0]: exact IUPAC motif, both strands, 12 Mb 16-chromosome genome -- whole hit list against the oracle
        text = synth_lines(16, 12_000_000, 100)
        ds = eng.load_dataset(text)
        conv, comp, opt = host.process_pattern("GATAAG", "dna", "Both strands", None, None, None, 0)
        ms, kms, kb, out, st = timed_requests(torch, eng, lambda: eng.search_request(ds, [conv, comp], opt, cap=1 << 20), args.steps, args.warmup)
        ok = all([(int(x), int(y)) for x, y in h] == oracle_lib.search(p, text, opt, cap=1 << 22) for p, h in zip((conv, comp), out))
        if not ok:
            raise SystemExit("bench: configs[0] hit list differs from the oracle")
        sec["configs[0]"] = entry("GATAAG exact, both strands, synthetic 12 Mb 16-chromosome genome (one request = 2 patterns, one pass)", len(text), 2, ms, kms, kb,
                                  "k_scan_packed_exact (TMA ring, both patterns per staged tile)", sum(len(h) for h in out), "whole hit list == CPU oracle", 0.375)
        ds.close()
        # the same request on the 3.1 Gb genome: the single-pattern-scan roofline the north star names
        from patmatchdocker_b200._native import pinned_empty, HIT_DTYPE
        big_out, _big_keep = pinned_empty(1 << 21, HIT_DTYPE)          # page-locked: the 24 MB hit list crosses PCIe at full speed
        ms, kms, kb, out, st = timed_requests(torch, eng, lambda: eng.search_request(big_ds, [conv, comp], opt, out=big_out), args.steps, args.warmup)
        sec["configs[0]@3.1Gb"] = entry("GATAAG exact, both strands, the 3.1 Gb genome of the headline (one request = 2 patterns, one pass over the planes)", big_bases, 2, ms, kms, kb,
                                        "k_scan_apx_jit in exact mode (NVRTC-specialised streaming kernel, both patterns per word, TMA ring)" if st["jit"] else "k_scan_packed_exact (TMA ring, both patterns per staged tile)",
                                        sum(len(h) for h in out), "sorted, non-overlapping; count == sum over single searches (tests)", 0.375)
        sec["configs[0]@3.1Gb"]["note"] = "frac = plane bytes of ONE pass / kernel time / HBM peak; the pass serves two patterns (per pattern the planes would cost twice that)"
        sec["configs[0]@3.1Gb"]["stage_ms"] = {k: round(st[k], 4) for k in ("scan_ms", "sort_ms", "verify_ms", "chain_ms", "total_ms")}
        # configs[1]: peptide pattern with degenerate classes, 1 substitution, ~6000-ORF proteome
        prot = synth_lines(6000, 2_900_000, 101, alphabet=b"ACDEFGHIKLMNPQRSTVWY", name="YORF")
        ds = eng.load_dataset(prot)
        conv, _, opt = host.process_pattern("CXXC[ILVM]XXHXXXH", "pep", None, None, None, "substitution", 1)
        ms, kms, kb, out, st = timed_requests(torch, eng, lambda: eng.search_request(ds, [conv], opt), args.steps, args.warmup)
        if [(int(x), int(y)) for x, y in out[0]] != oracle_lib.search(conv, prot, opt):
            raise SystemExit("bench: configs[1] hit list differs from the oracle")
        ptype = pm.plan(conv, opt)["type"]
        kname = ("k_scan_bytes over 5-bit residue codes (six per word, k_pack5)" if st["packed"] == 2 else
                 "k_scan_dense (%s plan of the reference: its approximate filter only proposes anchors, every anchor is verified on the raw bytes)" % ptype if ptype in ("BWD", "FWD") else
                 "k_scan_bytes (1 B/residue Shift-And)")
        sec["configs[1]"] = entry("peptide CXXC[ILVM]XXHXXXH, 1 substitution, synthetic 6000-ORF proteome (%.1f M residues), plan %s" % (len(prot) / 1e6, ptype), len(prot), 1, ms, kms, kb,
                                  kname, len(out[0]), "whole hit list == CPU oracle", 0.667 if st["packed"] == 2 else 1.0)
        # the same motif without errors: SIMPLE plan, Shift-And over the 5-bit residue codes
        conv0, _, opt0 = host.process_pattern("CXXC[ILVM]XXHXXXH", "pep", None, None, None, None, 0)
        ms, kms, kb, out, st = timed_requests(torch, eng, lambda: eng.search_request(ds, [conv0], opt0), args.steps, args.warmup)
        if [(int(x), int(y)) for x, y in out[0]] != oracle_lib.search(conv0, prot, opt0):
            raise SystemExit("bench: configs[1] (exact) hit list differs from the oracle")
        sec["configs[1]-exact"] = entry("peptide CXXC[ILVM]XXHXXXH, exact, same proteome, plan %s" % pm.plan(conv0, opt0)["type"], len(prot), 1, ms, kms, kb,
                                        "k_scan_bytes over 5-bit residue codes (six per word, k_pack5)" if st["packed"] == 2 else "k_scan_bytes (1 B/residue Shift-And)",
                                        len(out[0]), "whole hit list == CPU oracle", 0.667 if st["packed"] == 2 else 1.0)
        ds.close()
        # configs[2]: 20-nt pattern, 2 errors with indels, both strands, 12 Mb
        text = bytearray(synth_lines(16, 12_000_000, 102))
        rng = np.random.default_rng(8)
        motif = "TGACGTCAGATAAGCCGATT"
        for _ in range(300):
            q = int(rng.integers(1000, len(text) - 1000))
            if b"\n" in text[q - 40:q + 60] or b">" in text[q - 40:q + 60]:
                continue
            text[q:q + len(motif)] = motif.encode()
        text = bytes(text)
        ds = eng.load_dataset(text)
        conv, comp, opt = host.process_pattern(motif, "dna", "Both strands", None, None, None, 2)
        ms, kms, kb, out, st = timed_requests(torch, eng, lambda: eng.search_request(ds, [conv, comp], opt), args.steps, args.warmup)
        ok = all([(int(x), int(y)) for x, y in h] == oracle_lib.search(p, text, opt) for p, h in zip((conv, comp), out))
        if not ok:
            raise SystemExit("bench: configs[2] hit list differs from the oracle")
        sec["configs[2]"] = entry("20-nt TGACGTCAGATAAGCCGATT, -k 2ids, both strands, synthetic 12 Mb genome", len(text), 2, ms, kms, kb,
                                  "k_scan_apx (generic; the specialised kernel is compiled from 2^27 bases per pattern)", sum(len(h) for h in out), "whole hit list == CPU oracle", 0.375)
        ds.close()

    # configs[3]: 10,000 IUPAC motifs x 50 fungal-sized genomes (600 Mb), no collective on the data path
    import random
    from patmatchdocker_b200.distributed import shard_ranges
    rng = random.Random(5)
    iupac = {"R": "[AG]", "Y": "[CT]", "S": "[GC]", "W": "[AT]", "M": "[AC]", "K": "[GT]", "N": ".", "B": "[CGT]", "D": "[AGT]"}
    npat = args.batch_patterns
    pats = []
    for _ in range(npat):
        m = rng.randint(8, 14)
        pats.append("(" + "".join(rng.choice("ACGT") if rng.random() < 0.75 else iupac[rng.choice(list(iupac))] for _ in range(m)) + ")")
    lengths = [args.batch_bases // 800] * 800
    genome = make_genome_torch(lengths, list(range(len(lengths))), dev)
    ds = eng.wrap_device(genome.data_ptr(), genome.numel())

    def run_batch(motifs, pos_range):
        best, kms, kb, st = 1e18, 0.0, 0, None
        hits = off = None
        for rep in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            begins, off, base, mlen = eng.search_batch_compact(ds, motifs, "0ids", pos_range=pos_range)
            hits = (begins, base, mlen)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if rep > 0 and float(dt) < best:
                best = float(dt)
                st = eng.stats()
                kms, kb = st["scan_ms"], st["scan_bytes"]
        tot = torch.tensor([int(off[-1])], device=dev)
        if world > 1:
            dist.all_reduce(tot)
        return best, kms, kb, st, hits, off, int(tot)

    # N > 1: the batch is TEXT-sharded (every rank: all motifs x the buffer fills of 1/N of the file, pm_search_batch_fills);
    # the motif-sharded split of round 1 is timed beside it (the lookup kernel hashes every position once whatever the
    # number of motifs, so splitting the motif list leaves most of the work on every rank)
    by_motif = None
    if world > 1:
        b_, k_, _kb, _st, _h, _o, t_ = run_batch(pats[rank::world], None)
        by_motif = {"value": round(npat * genome.numel() / b_ / 1e9, 1), "ms_per_step": round(b_ * 1e3, 2), "kernel_ms": round(k_, 3), "hits_per_step": t_}
    mine = pats
    pos_range = shard_ranges(genome.numel(), world)[rank] if world > 1 else None
    best, kms, kb, st, hits, off, tot = run_batch(mine, pos_range)
    if by_motif is not None and by_motif["hits_per_step"] != tot:
        raise SystemExit("bench: configs[3] text-sharded and motif-sharded batches disagree on the number of hits")
    # the same work pipelined: the rank's range in 8 sub-ranges alternating between two engines (own stream and scratch)
    # driven by two host threads, so that the result copy of one sub-range overlaps the scan / sort / chain of the next
    from patmatchdocker_b200.distributed import PipelinedBatch
    eng2 = pm.Engine(dev.index)
    ds2 = eng2.wrap_device(genome.data_ptr(), genome.numel())
    pipe = PipelinedBatch([eng, eng2], [ds, ds2], parts=8)
    pbest, pres = 1e18, None
    for rep in range(4):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pres = pipe.search(mine, "0ids", pos_range=pos_range)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rep > 1:
            pbest = min(pbest, float(dt))
    ptot = torch.tensor([sum(int(o[-1]) for _b, o, _base, _ml in pres)], device=dev)
    if world > 1:
        dist.all_reduce(ptot)
    if int(ptot) != tot:
        raise SystemExit("bench: configs[3] pipelined batch disagrees on the number of hits")
    parity = None
    if rank == 0:
        # three sampled motifs against the oracle on a prefix of the file that ends at a buffer-fill boundary
        lines_, newlines_, n_ = [], [], 0
        pos = 0
        for i, ln in enumerate(lengths):
            head = len(">chr%d synthetic human-shaped\n" % (i + 1))
            newlines_.append(pos + head - 1); pos += head + ln + 1; newlines_.append(pos - 1)
        S, _E = python_fills(newlines_, pos)
        cut = S[min(24, len(S) - 1)]
        prefix = genome[:cut + 1].cpu().numpy().tobytes()
        okc = 0
        for i in (0, len(mine) // 2, len(mine) - 1):
            want = [h for h in oracle_lib.search(mine[i], prefix, "0ids", cap=1 << 22) if h[1] <= cut]
            begins, base, mlen = hits
            got = [(int(x) + base, int(x) + base + int(mlen[i])) for x in begins[off[i]:off[i + 1]] if int(x) + base + int(mlen[i]) <= cut]
            gotp = [(int(x) + pb_, int(x) + pb_ + int(ml_[i])) for b_, o_, pb_, ml_ in pres for x in b_[o_[i]:o_[i + 1]] if int(x) + pb_ + int(ml_[i]) <= cut]
            okc += 1 if got == want and gotp == want else 0
        if okc != 3:
            raise SystemExit("bench: configs[3] hit lists differ from the oracle")
        parity = "3 sampled motifs (single call and pipelined) == CPU oracle on the first %.0f Mb" % (cut / 1e6)
        ach = kb / (kms / 1e3) / 1e9 if kms > 0 else 0.0
        sec["configs[3]"] = {"workload": "%d IUPAC motifs (8-14 nt, 25%% degenerate positions) x synthetic %.0f Mb in 800 chromosomes (50 genomes x 16), exact, %d rank(s)" % (npat, genome.numel() / 1e6, world),
                             "value": round(npat * genome.numel() / min(pbest, best) / 1e9, 1), "unit": "pattern*Gbases/s", "ms_per_step": round(min(pbest, best) * 1e3, 2),
                             "mode": "pipelined" if pbest <= best else "single_call",
                             "pipelined": {"value": round(npat * genome.numel() / pbest / 1e9, 1), "ms_per_step": round(pbest * 1e3, 2)},
                             "pipeline": "each rank's range in 8 sub-ranges alternating between two engines of the GPU (two streams, two host threads): result copies overlap the next sub-range's scan / sort / chain (distributed.PipelinedBatch)",
                             "single_call": {"value": round(npat * genome.numel() / best / 1e9, 1), "ms_per_step": round(best * 1e3, 2), "note": "one pm_search_batch_fills_compact call per rank; kernel_ms / stage_ms / device_ms below are this call's"},
                             "kernel": "k_scan_multi_hash (TMA ring; every text position hashed once: 8/6/4-mer code -> CSR list of motifs -> bit-parallel verification on the planes) for %d of %d motifs on this rank, k_scan_packed_multi for the rest" % (st["qgram_chunks"], len(mine)), "kernel_ms": round(kms, 3),
                             "kernel_pattern_Gbases_per_s_per_gpu": round(len(mine) * genome.numel() / world / (kms / 1e3) / 1e9, 1) if kms > 0 else None,
                             "algorithmic_bytes": int(kb), "hits_per_step": int(tot), "d2h_bytes_per_step": int(tot) * 4, "parity": parity,
                             "result_format": "compact hit lists (pm_search_batch_fills_compact): 32-bit begin per hit + one length per motif; the 16-byte rows would be %.1f GB" % (tot * 16 / 1e9),
                             "device_ms": round(st["total_ms"], 2),
                             "stage_ms": {k: round(st[k], 2) for k in ("scan_ms", "sort_ms", "verify_ms", "chain_ms", "total_ms")},
                             "sharding": "text: every rank all motifs x the buffer fills that start in its 1/%d of the file; hit lists stay on the rank that found them (each crosses its own PCIe link), per-motif totals are all-reduced" % world if world > 1 else "single GPU",
                             "motif_sharded": by_motif,
                             "note": "lookup-bound (no HBM roofline: the planes are read once per batch of patterns); wall time includes sort, chain and the D2H copy of every hit (the index of the motif list is cached by the engine after the first call)"}
    ds2.close()
    eng2.close()
    ds.close()
    del genome
    return sec


def run_ours(args):
    import torch
    import torch.distributed as dist
    import patmatchdocker_b200 as pm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    from patmatchdocker_b200 import distributed as pmd
    lengths = chrom_lengths(args.bases)
    # every rank holds the whole genome (3.1 GB of 180 GB) and scans its share of the positions;
    # only hit lists travel (patmatchdocker_b200/distributed.py)
    mine = list(range(len(lengths)))
    total_bases = sum(lengths)
    pats, kopt = patterns()

    genome = make_genome_torch(lengths, mine, dev)
    lines, newlines, nfile = genome_layout(lengths)
    planted = plant_sites(genome, lines, "TGAGTCATTTACGATAAG", 256, seed=77)      # same on every rank
    nbytes = genome.numel()
    assert nbytes == nfile
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.copy_(genome)
    torch.cuda.synchronize()

    eng = pm.Engine(local)
    stream = torch.cuda.current_stream()
    eng.use_torch_stream(stream)
    ds = eng.wrap_device(genome.data_ptr(), nbytes)
    host_np = host.numpy()                                 # pinned host buffer holding the .seq file bytes

    stats_acc = {"scan_ms": 0.0, "scan_bytes": 0, "launches": 0, "searches": 0, "total_ms": 0.0, "packed": 0, "syncs": 0, "jit": 0, "last": None}

    sharded = pmd.DeviceShardedSearch(eng, rank, world, dev, cap=1 << 19) if world > 1 else None

    from patmatchdocker_b200._native import pinned_empty, HIT_DTYPE
    out_buf, _out_keep = pinned_empty(1 << 19, HIT_DTYPE)  # page-locked result buffer, filled directly by the D2H copy

    def run_request(d):
        # ONE pass for both strands (pm_search_request / pm_request_fills_device): one scan launch evaluates the motif
        # and its reverse complement on every staged tile; one sort / verify / chain / select; one synchronisation
        if world == 1:
            return eng.search_request(d, pats, kopt, out=out_buf)
        return sharded.request_fills(d, pats, kopt)        # fill-sharded: hit lists on rank 0, None elsewhere

    def step_resident():
        hits = run_request(ds)
        s = eng.stats()
        stats_acc["scan_ms"] += s["scan_ms"]
        stats_acc["scan_bytes"] += s["scan_bytes"]
        stats_acc["launches"] += s["launches"]
        stats_acc["total_ms"] += s["total_ms"]
        stats_acc["searches"] += 1
        stats_acc["packed"] = s["packed"]
        stats_acc["syncs"] = s["syncs"]
        stats_acc["jit"] = s["jit"]
        stats_acc["last"] = s
        return sum(len(h) for h in hits) if hits is not None else 0

    def step_e2e():
        # the call a user makes: file bytes in HOST memory -> dataset (H2D copy, 2-bit packing and record
        # index on the device) -> the request -> hit lists back in host memory
        if world == 1:
            # streaming: chunks are packed and searched while the next ones are still crossing PCIe (pm_search_stream)
            d, hits = eng.search_stream(host_np, pats, kopt)
            s = eng.stats()
            stats_acc["e2e_launches"] = s["launches"]
        else:
            # every rank uploads and packs only the bytes its own buffer fills need (windowed dataset)
            d = sharded.load_window(host)
            hits = run_request(d)
        d.close()
        return (sum(len(h) for h in hits), sum(h.nbytes for h in hits)) if hits is not None else (0, 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = None
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, out

    for i in range(max(args.warmup, 1)):
        step_resident()
        if i == 0:
            pm.jit_wait()            # specialised kernel of this request: compiled beside the first (generic) pass, used from here on
    # ---- parity of the benched hit list (outside the timed region) ----
    parity = None
    hit_lists = run_request(ds)
    if rank == 0:
        hit_lists = [np.array(h, copy=True) for h in hit_lists]
        for h in hit_lists:
            assert np.all(h["beg"][1:] >= h["end"][:-1]) and np.all(h["end"] > h["beg"]), "hit list not sorted / overlapping"
        rs = np.random.default_rng(3)
        centers = list(planted[:48])
        for h in hit_lists:
            centers += [int(h["beg"][i]) for i in rs.integers(0, len(h), size=24)]
        checked, okw = parity_windows(genome, lines, newlines, nbytes, hit_lists, pats, kopt, centers)
        covered = 0
        for p_ in planted:
            hit = False
            for h in hit_lists:
                j = np.searchsorted(h["end"], p_, side="right")
                hit = hit or (j < len(h) and h["beg"][j] < p_ + 18)
            covered += 1 if hit else 0
        parity = {"windows_checked": checked, "windows_ok": okw, "planted": len(planted), "planted_covered_by_a_hit": covered}
        if okw != checked or checked < 60 or covered != len(planted):
            print(json.dumps({"error": "benched hit list disagrees with the CPU oracle", "parity": parity}), flush=True)
            raise SystemExit(2)
    for k in stats_acc:
        stats_acc[k] = 0 if k != "last" else None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, nhits = timed(step_resident, args.steps)
    scan_ms, scan_bytes, nsearch = stats_acc["scan_ms"], stats_acc["scan_bytes"], stats_acc["searches"]
    launches = stats_acc["launches"]
    for i in range(max(1, args.warmup // 2)):
        step_e2e()
        if i == 0:
            pm.jit_wait()
    ms_e2e, (nhits2, d2h) = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    value = 2 * total_bases / (ms_step / 1e3) / 1e9
    e2e = 2 * total_bases / (ms_e2e / 1e3) / 1e9
    peak, peak_src = peaks()
    kms = scan_ms / max(nsearch, 1)
    achieved = (scan_bytes / nsearch) / (kms / 1e3) / 1e9 if scan_ms > 0 else 0.0
    prof = profile_constants() if stats_acc["jit"] else None
    traffic = inst = None
    int_pipe_frac = alu_pipe_frac = None
    sms, f_sm = 148, (clocks or {}).get("sm_mhz") or 1965
    if prof:
        share = (total_bases / world) / prof["scan_bases"]
        traffic = int((prof["dram_bytes_read"] + prof["dram_bytes_write"]) * share)
        inst = prof["warp_instructions"] * share
        issue_peak = sms * 4 * f_sm * 1e6                      # warp instructions per second, all schedulers
        int_pipe_frac = round(inst / (kms / 1e3) / issue_peak, 4)
        alu_pipe_frac = round(prof["alu_pipe_warp_instructions"] * share / (kms / 1e3) / (issue_peak / 2), 4)
    secondary = secondary_configs(args, torch, dist, pm, eng, dev, rank, world, ds, total_bases, peak) if args.secondary else None
    line = {
        "metric": "pattern.Gbases/s scanned (2-error degenerate motif, both strands)",
        "value": round(value, 3), "unit": "pattern*Gbases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "configs[4]: single 2-mismatch degenerate motif %s (-k %s), both strands, synthetic %.2f Gb 24-chromosome genome with 256 planted copies, chromosome-sharded" % (MOTIF, kopt, total_bases / 1e9),
                   "bases": total_bases, "patterns_per_step": 2, "plan": pm.plan(pats[0], kopt)["type"],
                   "l2_policy": "inputs (>= %.1f GB per rank) larger than the 126 MB L2" % (nbytes / 1e9),
                   "parallelism": "buffer fills (1.6 MB, independent by the reference's own restart rule) split over %d rank(s) by position; one NCCL all-gather of the per-rank hit lists" % world},
        "e2e": {"value": round(e2e, 3), "unit": "pattern*Gbases/s", "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": round(ms_e2e, 3),
                "path": ("pm_search_stream: file in pinned host memory -> 256 MiB chunks over PCIe, each packed and its completed buffer fills searched (both strands) while the next chunk is in flight -> hit lists in host memory"
                         if world == 1 else
                         "every rank uploads only the bytes of its own buffer fills (1/N of the pinned host file + one fill of overlap) over its own PCIe link, packs them, exchanges newline positions (one small all-gather), fill-sharded request, all-gather of hit lists, one D2H on rank 0; h2d_bytes_per_step is the whole job")},
        "gpu_launches": int(launches),
        "hits_per_step": int(nhits),
        "parity_windows_ok": (parity["windows_ok"] == parity["windows_checked"]) if parity else None, "parity": parity,
        "roofline": {"bound": "hbm", "limiter": "alu_pipe" if stats_acc["jit"] else None, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": traffic, "traffic_source": "profiles/r02_traffic.json (ncu --set full of this kernel on this workload, scaled to the bases of this launch)" if prof else None,
                     "kernel": ("k_scan_apx_jit (NVRTC-specialised for the request: both strands in one launch, 2-bit planes via TMA ring, streaming bit-sliced q-gram chunks + pieces, Landau-Vishkin check per surviving pattern start)"
                                if stats_acc["jit"] else "k_scan_apx (generic)") if stats_acc["packed"] else "k_scan_bytes",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(scan_bytes / max(nsearch, 1)),
                     "kernel_ms": round(kms, 4),
                     "kernel_share_of_step": round(kms / ms_step, 3),
                     "warp_instructions_per_launch": int(inst) if inst else None,
                     "int_pipe_frac": int_pipe_frac, "alu_pipe_frac": alu_pipe_frac,
                     "int_pipe_note": "int_pipe_frac = executed warp instructions / (148 SMs x 4 schedulers x f_SM x kernel time); alu_pipe_frac = LOP3/SHF/IADD-class instructions / the ALU pipe's issue rate (one warp instruction per 2 clocks and scheduler, tools/pipe_bench.cu); instruction counts from the committed ncu capture" if prof else None},
        "host_syncs_per_step": stats_acc["syncs"], "stage_ms": {k: round(stats_acc["last"][k], 4) for k in ("scan_ms", "sort_ms", "verify_ms", "chain_ms", "total_ms")} if stats_acc["last"] else None,
        "candidates_per_step": int(stats_acc["last"]["candidates"]) if stats_acc["last"] else None,
        "clocks": clocks,
        "secondary": secondary,
    }
    if rank == 0:
        line["cpu_baseline"] = cpu_baseline(pats, kopt, sample_bytes=args.cpu_sample, procs=1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "nrgrep_coords")
    return p if os.path.exists(p) else None


def cpu_baseline(pats, kopt, sample_bytes, procs):
    """The reference engine (oracle/_ref/nrgrep_coords, the unmodified binary) on the host cores:
    `procs` processes, each scanning its own sample of the workload with the forward-strand pattern.

    The sample is deliberately tiny: for k > 0 the reference re-locates the record boundaries for
    every candidate (recGetRecord @402030, O(record length) each), so on chromosome-sized lines it
    runs at ~40 kbases/s per core; a 0.8 Mb record costs ~12 s.  Records of the real workload are
    longer than the 1.6 MB buffer, where every candidate pays the full buffer: the sample favours
    the reference by up to 2x."""
    pats = pats[:1]
    rng = np.random.default_rng(5)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    binary = reference_binary()
    with tempfile.TemporaryDirectory() as td:
        paths = []
        for i in range(procs):
            p = os.path.join(td, "s%d.seq" % i)
            with open(p, "wb") as f:
                f.write(b">sample%d\n" % i)
                f.write(lut[rng.integers(0, 4, size=sample_bytes, dtype=np.uint8)].tobytes())
                f.write(b"\n")
            paths.append(p)
        t0 = time.perf_counter()
        if binary:
            kind = "reference"
            ps = []
            for p in paths:
                for pat in pats:
                    ps.append(subprocess.Popen([binary, "-i", "-b", "1600000", "-k", kopt, pat, p],
                                               stdout=subprocess.DEVNULL, env=ZERO_ENV))
                    if len(ps) >= procs:
                        ps.pop(0).wait()
            for q in ps:
                q.wait()
        else:
            kind = "port"
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib
            for p in paths:
                data = open(p, "rb").read()
                for pat in pats:
                    oracle_lib.search(pat, data, kopt)
        dt = time.perf_counter() - t0
    v = len(pats) * sample_bytes * procs / dt / 1e9
    return {"value": round(v, 7), "unit": "pattern*Gbases/s", "cores": procs, "kind": kind, "seconds": round(dt, 2),
            "sample": "%d process(es) x (one %.1f Mb synthetic chromosome record x %d pattern), nrgrep_coords -i -b 1600000 -k %s" % (procs, sample_bytes / 1e6, len(pats), kopt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pats, kopt = patterns()
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    steps = []
    for i in range(args.warmup + args.steps):
        b = cpu_baseline(pats, kopt, sample_bytes=args.cpu_sample, procs=procs)
        if i >= args.warmup:
            steps.append(b)
    v = float(np.mean([b["value"] for b in steps]))
    ms = float(np.mean([b["seconds"] for b in steps])) * 1e3
    total = sum(chrom_lengths(args.bases))
    line = {"impl": "reference", "metric": "pattern.Gbases/s scanned (2-error degenerate motif, both strands)",
            "value": round(v, 7), "unit": "pattern*Gbases/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "configs[4]: single 2-mismatch degenerate motif %s (-k %s), both strands, synthetic %.2f Gb genome; each step a bounded sample" % (MOTIF, kopt, total / 1e9)},
            "cpu_baseline": dict(steps[-1], value=round(v, 7)),
            "e2e": {"value": round(v, 7), "unit": "pattern*Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bases", type=int, default=3_100_000_000)
    ap.add_argument("--cpu-sample", type=int, default=800_000, help="bases per process of the CPU baseline sample")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false", help="skip the other BASELINE configs")
    ap.add_argument("--batch-patterns", type=int, default=10_000, help="configs[3]: motifs in the batch")
    ap.add_argument("--batch-bases", type=int, default=600_000_000, help="configs[3]: bases of the multi-genome dataset")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        import __graft_entry__ as g
        if not os.path.exists(g.LIB):
            g.build()
        run_ours(args)


if __name__ == "__main__":
    main()
