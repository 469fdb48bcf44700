#!/usr/bin/env python
"""Small fixed workloads for ncu (one kernel family each).
usage: ncu_target.py request|exact|batch|pep|generic [bases]
  request : the bench request (motif + reverse complement, -k 2ids) -> k_scan_apx_jit
  generic : the same request with the specialised kernels switched off -> k_scan_apx
  exact   : GATAAG / CTTATC in one pass -> k_scan_packed_exact
  batch   : 2000 IUPAC motifs -> k_scan_multi_hash (+ k_scan_packed_multi for the motifs without a window)
  pep     : exact peptide motif on a proteome -> k_scan_bytes over 5-bit residue codes"""
import sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import numpy as np
import patmatchdocker_b200 as pm
import bench
kind = sys.argv[1] if len(sys.argv) > 1 else "request"
bases = int(float(sys.argv[2])) if len(sys.argv) > 2 else 3_100_000_000
dev = torch.device("cuda", 0)
eng = pm.Engine(0)
eng.set_jit("always")      # profiling: compile the specialised kernel synchronously (the default compiles in the background)
if kind == "pep":
    prot = bench.synth_lines(60000, bases if len(sys.argv) > 2 else 30_000_000, 101, alphabet=b"ACDEFGHIKLMNPQRSTVWY", name="YORF")
    ds = eng.load_dataset(prot)
    for rep in range(3):
        h = eng.search_request(ds, ["(C..C[ILVM]..H...H)"], "0ids")
    print(kind, len(h[0]), eng.stats())
    sys.exit(0)
lengths = bench.chrom_lengths(bases) if kind != "batch" else [bases // 800] * 800
genome = bench.make_genome_torch(lengths, list(range(len(lengths))), dev)
torch.cuda.synchronize()
ds = eng.wrap_device(genome.data_ptr(), genome.numel())
if kind in ("request", "generic"):
    pats, kopt = bench.patterns()
    if kind == "generic":
        eng.set_jit("off")
    for rep in range(3):
        h = eng.search_request(ds, pats, kopt)
elif kind == "exact":
    for rep in range(3):
        h = eng.search_request(ds, ["(GATAAG)", "(CTTATC)"], "0ids", cap=1 << 22)
elif kind == "batch":
    rng = random.Random(5)
    iupac = {"R": "[AG]", "Y": "[CT]", "S": "[GC]", "W": "[AT]", "M": "[AC]", "K": "[GT]", "N": ".", "B": "[CGT]", "D": "[AGT]"}
    pats = []
    for _ in range(2000):
        m = rng.randint(8, 14)
        pats.append("(" + "".join(rng.choice("ACGT") if rng.random() < 0.75 else iupac[rng.choice(list(iupac))] for _ in range(m)) + ")")
    for rep in range(2):
        h = eng.search_batch(ds, pats, "0ids", cap=1 << 24, copy=False)
print(kind, eng.stats())
