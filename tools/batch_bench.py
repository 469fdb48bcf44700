#!/usr/bin/env python
"""BASELINE configs[3]-shaped run on one GPU: a batch of IUPAC motifs over a multi-genome dataset,
fused multi-pattern scan.  usage: batch_bench.py [npatterns] [bases]"""
import sys, os, time, json, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import patmatchdocker_b200 as pm
import bench

npat = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
bases = int(float(sys.argv[2])) if len(sys.argv) > 2 else 600_000_000
rng = random.Random(5)
iupac = {"R": "[AG]", "Y": "[CT]", "S": "[GC]", "W": "[AT]", "M": "[AC]", "K": "[GT]", "N": ".", "B": "[CGT]", "D": "[AGT]"}
pats = []
for _ in range(npat):
    m = rng.randint(8, 14)
    pats.append("(" + "".join(rng.choice("ACGT") if rng.random() < 0.75 else iupac[rng.choice(list(iupac))] for _ in range(m)) + ")")
dev = torch.device("cuda", 0)
lengths = [bases // 800] * 800                      # 50 genomes x 16 chromosomes
genome = bench.make_genome_torch(lengths, list(range(len(lengths))), dev)
eng = pm.Engine(0)
ds = eng.wrap_device(genome.data_ptr(), genome.numel())
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hits, off = eng.search_batch(ds, pats, "0ids", cap=1 << 24)
    dt = time.perf_counter() - t0
s = eng.stats()
print(json.dumps({"patterns": npat, "bases": genome.numel(), "hits": int(off[-1]), "wall_s": round(dt, 4),
                  "pattern_Gbases_per_s_wall": round(npat * genome.numel() / dt / 1e9, 1),
                  "scan_ms": round(s["scan_ms"], 3), "pattern_Gbases_per_s_scan": round(npat * genome.numel() / (s["scan_ms"] / 1e3) / 1e9, 1),
                  "sort_ms": round(s["sort_ms"], 3), "chain_ms": round(s["chain_ms"], 3), "launches": s["launches"]}))
