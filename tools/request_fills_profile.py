#!/usr/bin/env python
"""Host-side cost of DeviceShardedSearch.request_fills (world = 1, no collective): where the time between the device
work and the returned hit lists goes.  usage: request_fills_profile.py [bases]"""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import patmatchdocker_b200 as pm
from patmatchdocker_b200 import distributed as pmd
import bench
bases = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_550_000_000
dev = torch.device("cuda", 0)
lengths = bench.chrom_lengths(bases)
g = bench.make_genome_torch(lengths, list(range(len(lengths))), dev)
eng = pm.Engine(0)
eng.set_jit("always")      # profiling: compile the specialised kernel synchronously (the default compiles in the background)
sh = pmd.DeviceShardedSearch(eng, 0, 1, dev)
ds = eng.wrap_device(g.data_ptr(), g.numel())
pats, kopt = bench.patterns()
for _ in range(4):
    sh.request_fills(ds, pats, kopt)
t0 = time.perf_counter()
for _ in range(20):
    h = sh.request_fills(ds, pats, kopt)
dt = (time.perf_counter() - t0) / 20 * 1e3
st = eng.stats()
print("wall %.3f ms  device total %.3f ms  hits %d" % (dt, st["total_ms"], sum(len(x) for x in h)))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    sh.request_fills(ds, pats, kopt)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
