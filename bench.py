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
    # only verified candidates travel (patmatchdocker_b200/distributed.py)
    mine = list(range(len(lengths)))
    total_bases = sum(lengths)
    pats, kopt = patterns()

    genome = make_genome_torch(lengths, mine, dev)
    nbytes = genome.numel()
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.copy_(genome)
    torch.cuda.synchronize()

    eng = pm.Engine(local)
    stream = torch.cuda.current_stream()
    eng.use_torch_stream(stream)
    ds = eng.wrap_device(genome.data_ptr(), nbytes)
    host_np = host.numpy()                                 # pinned host buffer holding the .seq file bytes

    stats_acc = {"scan_ms": 0.0, "scan_bytes": 0, "launches": 0, "searches": 0, "total_ms": 0.0, "packed": 0, "syncs": 0, "last": None}

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
        stats_acc["last"] = s
        return sum(len(h) for h in hits) if hits is not None else 0

    def step_e2e():
        # the call a user makes: file bytes in HOST memory -> dataset (H2D copy, 2-bit packing and record
        # index on the device) -> the two searches of the request -> hit lists back in host memory
        # (N > 1: every rank uploads 1/N of the file over its own PCIe link, NCCL all-gathers the slices over NVLink)
        if world == 1:
            # streaming: chunks are packed and searched while the next ones are still crossing PCIe (pm_search_stream)
            d, hits = eng.search_stream(host_np, pats, kopt)
            s = eng.stats()
            stats_acc["e2e_launches"] = s["launches"]
        else:
            d = sharded.load_dataset(host)
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

    for _ in range(args.warmup):
        step_resident()
    for k in stats_acc:
        stats_acc[k] = 0 if k != "last" else None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, nhits = timed(step_resident, args.steps)
    scan_ms, scan_bytes, nsearch = stats_acc["scan_ms"], stats_acc["scan_bytes"], stats_acc["searches"]
    launches = stats_acc["launches"]
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    ms_e2e, (nhits2, d2h) = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    value = 2 * total_bases / (ms_step / 1e3) / 1e9
    e2e = 2 * total_bases / (ms_e2e / 1e3) / 1e9
    peak, peak_src = peaks()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp) and stats_acc["packed"]:
        tr = json.load(open(tp)).get("k_scan_split")
        if tr:                                             # DRAM bytes of one launch from the committed ncu capture,
            per_base = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["scan_bases"]   # scaled to this launch's bases
            traffic = int(per_base * total_bases / world)
    achieved = (scan_bytes / nsearch) / (scan_ms / nsearch / 1e3) / 1e9 if scan_ms > 0 else 0.0
    line = {
        "metric": "pattern.Gbases/s scanned (2-error degenerate motif, both strands)",
        "value": round(value, 3), "unit": "pattern*Gbases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "configs[4]: single 2-mismatch degenerate motif %s (-k %s), both strands, synthetic %.2f Gb 24-chromosome genome, chromosome-sharded" % (MOTIF, kopt, total_bases / 1e9),
                   "bases": total_bases, "patterns_per_step": 2, "plan": pm.plan(pats[0], kopt)["type"],
                   "l2_policy": "inputs (>= %.1f GB per rank) larger than the 126 MB L2" % (nbytes / 1e9),
                   "parallelism": "buffer fills (1.6 MB, independent by the reference's own restart rule) split over %d rank(s) by position; one NCCL all-gather of the per-rank hit lists" % world},
        "e2e": {"value": round(e2e, 3), "unit": "pattern*Gbases/s", "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": round(ms_e2e, 3),
                "path": ("pm_search_stream: file in pinned host memory -> 256 MiB chunks over PCIe, each packed and its completed buffer fills searched (both strands) while the next chunk is in flight -> hit lists in host memory"
                         if world == 1 else
                         "every rank uploads 1/N of the pinned host file, NCCL all-gather over NVLink, pack, fill-sharded search of both strands, all-gather of hit lists, one D2H on rank 0")},
        "gpu_launches": int(launches),
        "hits_per_step": int(nhits),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": traffic, "kernel": "k_scan_apx<3,u32> (both strands in one launch: 2-bit planes via TMA ring, pieces built from bit-sliced q-gram chunks, Landau-Vishkin check per surviving pattern start; integer-pipe bound)" if stats_acc["packed"] else "k_scan_bytes",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(scan_bytes / max(nsearch, 1)),
                     "kernel_ms": round(scan_ms / max(nsearch, 1), 4),
                     "kernel_share_of_step": round(scan_ms / max(nsearch, 1) / ms_step, 3)},
        "host_syncs_per_step": stats_acc["syncs"], "stage_ms": {k: round(stats_acc["last"][k], 4) for k in ("scan_ms", "sort_ms", "verify_ms", "chain_ms", "total_ms")} if stats_acc["last"] else None,
        "candidates_per_step": int(stats_acc["last"]["candidates"]) if stats_acc["last"] else None,
        "clocks": clocks,
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
