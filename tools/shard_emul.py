#!/usr/bin/env python
"""What one rank of an N-GPU fill-sharded request costs, emulated on ONE GPU: pm_request_fills_device over 1/N of the
positions (device stages, enqueue wall time, wall time to completion).  usage: shard_emul.py [bases]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import patmatchdocker_b200 as pm
from patmatchdocker_b200 import distributed as pmd
from patmatchdocker_b200._native import request_header_rows
import bench
bases = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
dev = torch.device("cuda", 0)
lengths = bench.chrom_lengths(bases)
g = bench.make_genome_torch(lengths, list(range(len(lengths))), dev)
eng = pm.Engine(0); eng.use_torch_stream()
eng.set_jit("always")      # profiling: compile the specialised kernel synchronously (the default compiles in the background)
ds = eng.wrap_device(g.data_ptr(), g.numel())
pats, kopt = bench.patterns()
rows = 1 << 18
buf = torch.zeros((rows, 2), dtype=torch.int64, device=dev)
for world in (1, 2, 4, 8):
    beg, end = pmd.shard_ranges(len(ds), world)[world // 2]
    cap = 1 << 19
    for _ in range(3):
        eng.request_fills_device(ds, pats, kopt, beg, end, cap, buf.data_ptr(), rows); torch.cuda.synchronize()
        cap = int(int(buf[0, 1]) * 1.25) + 1024
    acc = {}
    reps = 20
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.request_fills_device(ds, pats, kopt, beg, end, cap, buf.data_ptr(), rows)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        st = eng.stats()
        for k in ("scan_ms", "sort_ms", "verify_ms", "chain_ms", "total_ms"): acc[k] = acc.get(k, 0) + st[k]
        acc["enqueue_wall_ms"] = acc.get("enqueue_wall_ms", 0) + (t1 - t0) * 1e3
        acc["done_wall_ms"] = acc.get("done_wall_ms", 0) + (t2 - t0) * 1e3
    print(json.dumps({"world": world, "cap": cap, "launches": st["launches"], "jit": st["jit"], **{k: round(v / reps, 4) for k, v in acc.items()}}), flush=True)
