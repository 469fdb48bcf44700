"""Where a sharded search spends its time (host wall clock per phase, device synchronised between phases).
torchrun --nproc-per-node N tools/shard_phases.py [bases]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import patmatchdocker_b200 as pm
from patmatchdocker_b200 import distributed as pmd
import bench

def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bases = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
    lengths = bench.chrom_lengths(bases)
    g = bench.make_genome_torch(lengths, list(range(len(lengths))), dev)
    eng = pm.Engine(local)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    ds = eng.wrap_device(g.data_ptr(), g.numel())
    pats, kopt = bench.patterns()
    sh = pmd.DeviceShardedSearch(eng, rank, world, dev, cap=1 << 19)
    for _ in range(3):
        sh.search(ds, pats[0], kopt)
    acc = {}
    def mark(name, t0):
        torch.cuda.synchronize()
        t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0) * 1e3; return t1
    reps = 20
    for _ in range(reps):
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        beg, end = pmd.shard_ranges(len(ds), world)[rank]
        n = eng.candidates_device(ds, pats[0], kopt, beg, end, sh.mine[1:].data_ptr(), sh.cap - 1)
        st = eng.stats()
        t = mark("candidates_device", t)
        acc["  scan_ms(dev)"] = acc.get("  scan_ms(dev)", 0) + st["scan_ms"]; acc["  sort_ms(dev)"] = acc.get("  sort_ms(dev)", 0) + st["sort_ms"]
        acc["  verify_ms(dev)"] = acc.get("  verify_ms(dev)", 0) + st["verify_ms"]
        sh.mine[0, 0] = abs(n)
        t = mark("header", t)
        rows = min(sh.rows, sh.cap)
        if world > 1:
            dist.all_gather_into_tensor(sh.allbuf[: world * rows].view(world, rows, 4), sh.mine[:rows])
        else:
            sh.allbuf[:rows] = sh.mine[:rows]
        counts = sh.allbuf[: world * rows].view(world, rows, 4)[:, 0, 0].tolist()
        t = mark("all_gather+counts", t)
        if rank == 0:
            view = sh.allbuf[: world * rows].view(world, rows, 4); off = 0
            for r in range(world):
                c = int(counts[r]); sh.merged[off:off + c] = view[r, 1:1 + c]; off += c
            t = mark("merge", t)
            hits = eng.resolve_device(ds, pats[0], kopt, sh.merged.data_ptr(), off)
            t = mark("resolve_device", t)
    if rank == 0:
        print(json.dumps({"world": world, "per_search_ms": {k: round(v / reps, 4) for k, v in acc.items()}}))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
main()
