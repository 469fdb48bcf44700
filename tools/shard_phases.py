"""Where a multi-GPU request spends its time: host wall clock per phase with the device synchronised between the phases
(so the phases do not overlap as they do in bench.py; the sum is an upper bound of the step).
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/shard_phases.py [bases]
resident: pm_request_fills_device on this rank's fills | all-gather of [header | hits] | D2H + sync on rank 0 | host merge
cold    : pm_dataset_create_window (H2D of this rank's window + pack) | newline all-gather + D2H + fill table | request"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
import patmatchdocker_b200 as pm
from patmatchdocker_b200 import distributed as pmd
from patmatchdocker_b200._native import request_header_rows
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
    host = torch.empty(g.numel(), dtype=torch.uint8, pin_memory=True)
    host.copy_(g)
    torch.cuda.synchronize()
    eng = pm.Engine(local)
    eng.set_jit("always")      # profiling: compile the specialised kernel synchronously (the default compiles in the background)
    sh = pmd.DeviceShardedSearch(eng, rank, world, dev)
    ds = eng.wrap_device(g.data_ptr(), g.numel())
    pats, kopt = bench.patterns()
    for _ in range(4):
        sh.request_fills(ds, pats, kopt)
    acc = {}

    def mark(name, t0):
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        acc[name] = acc.get(name, 0.0) + (t1 - t0) * 1e3
        return t1
    reps = 20
    hr = request_header_rows(len(pats))
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = t = time.perf_counter()
        beg, end = pmd.shard_ranges(len(ds), world)[rank]
        rows = max(sh.rq_rows, hr + 16)
        mine = sh.rq_mine[:rows]
        eng.request_fills_device(ds, pats, kopt, beg, end, sh.rq_cap, mine.data_ptr(), rows)
        acc["resident.enqueue_host"] = acc.get("resident.enqueue_host", 0.0) + (time.perf_counter() - t) * 1e3
        t = mark("resident.request_fills_device", t)
        st = eng.stats()
        for k in ("scan_ms", "sort_ms", "verify_ms", "chain_ms"):
            acc["resident.dev." + k] = acc.get("resident.dev." + k, 0.0) + st[k]
        t = time.perf_counter()
        flat = sh.rq_all[: world * rows]
        if world > 1:
            dist.all_gather_into_tensor(flat, mine)
        else:
            flat.copy_(mine)
        t = mark("resident.all_gather", t)
        if rank == 0:
            hview = sh.rq_host[: world * rows]
            hview.copy_(flat, non_blocking=True)
            t = mark("resident.d2h", t)
            h = hview.numpy().reshape(world, rows, 2)
            counts = h[:, 2:hr].reshape(world, -1)[:, :len(pats)]
            offs = np.zeros((world, len(pats) + 1), dtype=np.int64); offs[:, 1:] = np.cumsum(counts, axis=1)
            out = [np.concatenate([h[r, hr + offs[r, p]: hr + offs[r, p + 1]] for r in range(world)]) for p in range(len(pats))]
            t = mark("resident.host_merge", t)
        acc["resident.sum"] = acc.get("resident.sum", 0.0) + (time.perf_counter() - t0) * 1e3
        # the real thing, phases overlapping
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        sh.request_fills(ds, pats, kopt)
        t = mark("resident.request_fills_call", t)
    for _ in range(2):
        sh.load_window(host).close()
    for _ in range(5):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        n = int(host.numel())
        beg, end = pmd.shard_ranges(n, world)[rank]
        lo, hi = max(beg - 4096, 0), min(n, end + 1600000 + 4096)
        d = eng.load_window(host.numpy(), lo, hi, sh.nl_mine.data_ptr(), sh.nl_mine.shape[0])
        t = mark("cold.h2d_window+pack", t)
        acc["cold.window_bytes"] = acc.get("cold.window_bytes", 0.0) + (hi - lo)
        if world > 1:
            dist.all_gather_into_tensor(sh.nl_all, sh.nl_mine)
        else:
            sh.nl_all.copy_(sh.nl_mine)
        sh.nl_host.copy_(sh.nl_all, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        hh = sh.nl_host.numpy().reshape(world, -1)
        allnl = np.unique(np.concatenate([hh[r, 1:1 + int(hh[r, 0])] for r in range(world)]))
        eng.set_newlines(d, allnl)
        t = mark("cold.newlines_exchange+fills", t)
        sh.request_fills(d, pats, kopt)
        t = mark("cold.request", t)
        d.close()
        t = mark("cold.close", t)
    for k in ("resident.d2h", "resident.host_merge"):          # rank 0 only: every rank must reduce the same keys
        acc.setdefault(k, 0.0)
    res = {k: round(v / (5 if k.startswith("cold") else reps), 4) for k, v in acc.items()}
    tl = torch.tensor([res[k] for k in sorted(res)], device=dev)
    if world > 1:
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"world": world, "bases": bases, "max_over_ranks_ms": dict(zip(sorted(res), [round(float(x), 4) for x in tl]))}), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


main()
