#!/usr/bin/env python
"""Scan-kernel time of exact (k = 0) searches on the 3.1 Gb synthetic genome: best of 7 by CUDA events."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import patmatchdocker_b200 as pm
import bench
bases = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
dev = torch.device("cuda", 0)
lengths = bench.chrom_lengths(bases)
genome = bench.make_genome_torch(lengths, list(range(len(lengths))), dev)
eng = pm.Engine(0)
ds = eng.wrap_device(genome.data_ptr(), genome.numel())
peak = bench.peaks()[0]
for p in ("(GATAAG)", "(GATAAGCCTTAGGCATTGCA)", "(GA[AT]A[AG][CG]N[CT]TA)", "(TATAAA)"):
    best = 1e9
    for rep in range(7):
        n = eng.count(ds, p, "0ids")
        s = eng.stats()
        best = min(best, s["scan_ms"])
    gbs = s["scan_bytes"] / best / 1e6
    print(json.dumps({"pattern": p, "hits": int(n), "scan_ms": round(best, 4), "GB/s": round(gbs, 1), "frac_of_copy_peak": round(gbs / peak, 4)}), flush=True)
