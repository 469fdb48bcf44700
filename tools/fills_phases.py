"""Device and wall time of a fill-sharded search per phase.  torchrun --nproc-per-node N tools/fills_phases.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import patmatchdocker_b200 as pm
from patmatchdocker_b200 import distributed as pmd
import bench
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
lengths = bench.chrom_lengths(3_100_000_000)
g = bench.make_genome_torch(lengths, list(range(len(lengths))), dev)
eng = pm.Engine(local); eng.set_stream(torch.cuda.current_stream().cuda_stream)
eng.set_jit("always")      # profiling: compile the specialised kernel synchronously (the default compiles in the background)
ds = eng.wrap_device(g.data_ptr(), g.numel())
pats, kopt = bench.patterns()
sh = pmd.DeviceShardedSearch(eng, rank, world, dev, cap=1 << 19)
for _ in range(3): sh.search_fills(ds, pats[0], kopt)
buf = torch.zeros((1 << 17, 2), dtype=torch.int64, device=dev)
beg, end = pmd.shard_ranges(len(ds), world)[rank]
acc = {}
for _ in range(20):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = eng.search_fills_device(ds, pats[0], kopt, beg, end, buf.data_ptr(), buf.shape[0])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    st = eng.stats()
    for k in ("scan_ms", "sort_ms", "verify_ms", "chain_ms", "total_ms"): acc[k] = acc.get(k, 0) + st[k]
    acc["search_fills_device_wall"] = acc.get("search_fills_device_wall", 0) + (t1 - t0) * 1e3
    t0 = time.perf_counter(); h = sh.search_fills(ds, pats[0], kopt); torch.cuda.synchronize(); t1 = time.perf_counter()
    acc["search_fills_total_wall"] = acc.get("search_fills_total_wall", 0) + (t1 - t0) * 1e3
if rank == 0: print(json.dumps({k: round(v / 20, 3) for k, v in acc.items()}), n, st["candidates"])
if world > 1: dist.barrier(); dist.destroy_process_group()
