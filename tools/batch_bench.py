#!/usr/bin/env python
"""BASELINE configs[3]-shaped run: a batch of IUPAC motifs over a multi-genome dataset (50 genomes x 16
chromosomes), fused multi-pattern scan, patterns sharded over the ranks (no collective on the data path;
rank 0 only sums the hit counts).

  python tools/batch_bench.py [npatterns] [bases]
  python -m torch.distributed.run --nproc-per-node N ... tools/batch_bench.py [npatterns] [bases]
"""
import sys, os, time, json, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import patmatchdocker_b200 as pm
import bench

npat = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
bases = int(float(sys.argv[2])) if len(sys.argv) > 2 else 600_000_000
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
rng = random.Random(5)
iupac = {"R": "[AG]", "Y": "[CT]", "S": "[GC]", "W": "[AT]", "M": "[AC]", "K": "[GT]", "N": ".", "B": "[CGT]", "D": "[AGT]"}
pats = []
for _ in range(npat):
    m = rng.randint(8, 14)
    pats.append("(" + "".join(rng.choice("ACGT") if rng.random() < 0.75 else iupac[rng.choice(list(iupac))] for _ in range(m)) + ")")
mine = pats[rank::world]
dev = torch.device("cuda", local)
lengths = [bases // 800] * 800
genome = bench.make_genome_torch(lengths, list(range(len(lengths))), dev)
eng = pm.Engine(local)
ds = eng.wrap_device(genome.data_ptr(), genome.numel())
times = []
for rep in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hits, off = eng.search_batch(ds, mine, "0ids", cap=1 << 24)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    nh = torch.tensor([int(off[-1])], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(nh)
    times.append(float(dt))
s = eng.stats()
scan = torch.tensor([s["scan_ms"]], device=dev)
if world > 1:
    dist.all_reduce(scan, op=dist.ReduceOp.MAX)
if rank == 0:
    dt = min(times[1:])
    print(json.dumps({"n_gpus": world, "patterns": npat, "bases": genome.numel(), "hits": int(nh), "wall_s": round(dt, 4),
                      "pattern_Gbases_per_s_wall": round(npat * genome.numel() / dt / 1e9, 1),
                      "scan_ms_max_rank": round(float(scan), 3),
                      "pattern_Gbases_per_s_scan": round(npat * genome.numel() / (float(scan) / 1e3) / 1e9, 1),
                      "sort_ms": round(s["sort_ms"], 3), "chain_ms": round(s["chain_ms"], 3), "launches": s["launches"]}))
if world > 1:
    dist.destroy_process_group()
