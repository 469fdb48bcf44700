#!/usr/bin/env python
"""Per-stage device times of one search on a synthetic genome (CUDA events inside the library).
usage: stage_times.py [bases] [pattern] [kopt]"""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import patmatchdocker_b200 as pm
import bench

bases = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400_000_000
pats, kopt = bench.patterns()
if len(sys.argv) > 2:
    pats, kopt = [sys.argv[2]], sys.argv[3]
dev = torch.device("cuda", 0)
lengths = bench.chrom_lengths(bases)
genome = bench.make_genome_torch(lengths, list(range(len(lengths))), dev)
eng = pm.Engine(0)
eng.set_jit("always")      # profiling: compile the specialised kernel synchronously (the default compiles in the background)
ds = eng.wrap_device(genome.data_ptr(), genome.numel())
for mode in ("packed", "bytes"):
    eng.set_scan_mode(mode)
    for p in pats[:1]:
        for rep in range(3):
            t0 = time.perf_counter()
            h = eng.search(ds, p, kopt)
            dt = (time.perf_counter() - t0) * 1e3
        s = eng.stats()
        s = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.items()}
        print(json.dumps({"mode": mode, "pattern": p, "kopt": kopt, "plan": pm.plan(p, kopt)["type"], "L": pm.plan(p, kopt)["L"],
                          "bases": genome.numel(), "wall_ms": round(dt, 3), **s}))
