#!/usr/bin/env python
"""Scan-kernel experiments on the bench request (3.1 Gb, both strands): every VARIANT is a comma-separated list of
environment knobs (PM_APX_FAMILY=A, PM_JIT_CTAS=3, ...) applied before the kernels are (re)built.
usage: jit_bench.py [bases] VARIANT [VARIANT ...]      ('-' = defaults)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import patmatchdocker_b200 as pm
import bench
args = sys.argv[1:]
bases = int(float(args.pop(0))) if args and args[0][0].isdigit() else 3_100_000_000
variants = args or ["-"]
pats, kopt = bench.patterns()
dev = torch.device("cuda", 0)
lengths = bench.chrom_lengths(bases)
genome = bench.make_genome_torch(lengths, list(range(len(lengths))), dev)
eng = pm.Engine(0)
ds = eng.wrap_device(genome.data_ptr(), genome.numel())
lib = pm.load()
ref = None
for v in variants:
    knobs = dict(kv.split("=") for kv in v.split(",")) if v != "-" else {}
    for k in list(os.environ):
        if k.startswith("PM_APX_") or k.startswith("PM_JIT_"):
            del os.environ[k]
    os.environ.update(knobs)
    lib.pm_debug_reset_caches()
    eng.set_jit("off" if knobs.get("PM_JIT") == "0" else "always")
    t0 = time.perf_counter()
    h = eng.search_request(ds, pats, kopt)
    t_first = time.perf_counter() - t0
    best, wall = 1e9, 1e9
    for rep in range(6):
        t0 = time.perf_counter()
        h = eng.search_request(ds, pats, kopt)
        wall = min(wall, time.perf_counter() - t0)
        st = eng.stats()
        best = min(best, st["scan_ms"])
    sig = (len(h[0]), len(h[1]), int(h[0]["beg"].sum()), int(h[1]["end"].sum()))
    ref = ref or sig
    print("%-40s scan %.3f ms  total %.3f ms  wall %.3f ms  first %.0f ms  cands %d  jit %d  hits %s %s" %
          (v, best, st["total_ms"], wall * 1e3, t_first * 1e3, st["candidates"], st["jit"], sig[:2], "OK" if sig == ref else "MISMATCH"), flush=True)
