#!/usr/bin/env python
"""Small fixed workload for ncu: 3 searches of one pattern on a 400 Mb synthetic genome.
usage: ncu_target.py [exact|approx|selective] [bases]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import patmatchdocker_b200 as pm
import bench
kind = sys.argv[1] if len(sys.argv) > 1 else "approx"
bases = int(float(sys.argv[2])) if len(sys.argv) > 2 else 400_000_000
pats, kopt = bench.patterns()
if kind == "exact":
    pats, kopt = ["(GATAAG)"], "0ids"
elif kind == "selective":                       # approximate search whose pieces rarely match: the dense part of the scan
    pats, kopt = ["(GATAAGCC[AT]TTACGGA)"], "2ids"
dev = torch.device("cuda", 0)
lengths = bench.chrom_lengths(bases)
genome = bench.make_genome_torch(lengths, list(range(len(lengths))), dev)
eng = pm.Engine(0)
ds = eng.wrap_device(genome.data_ptr(), genome.numel())
for rep in range(3):
    h = eng.search(ds, pats[0], kopt)
print(kind, len(h), eng.stats())
