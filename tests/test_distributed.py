"""CPU: the N>1 host logic over gloo with world_size 2 (no GPU needed)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from patmatchdocker_b200 import distributed as D
from patmatchdocker_b200._native import CAND_DTYPE


def test_shard_ranges_cover_every_anchor_once():
    for n in (0, 1, 17, 1000, 12_000_037):
        for world in (1, 2, 3, 4, 8):
            r = D.shard_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == n + 1
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            assert all(e >= b for b, e in r)


def test_merge_batch_shards_is_pattern_major_in_file_order():
    from patmatchdocker_b200._native import HIT_DTYPE
    rng = np.random.default_rng(7)
    npat, world, n = 37, 3, 90000
    whole = []
    for p in range(npat):
        b = np.sort(rng.choice(n, size=int(rng.integers(0, 40)), replace=False))
        whole.append(b)
    edges = [0, 31000, 31000, n + 1]                      # the middle rank owns nothing
    parts = []
    for r in range(world):
        hs, off = [], [0]
        for p in range(npat):
            b = whole[p][(whole[p] >= edges[r]) & (whole[p] < edges[r + 1])]
            h = np.zeros(len(b), dtype=HIT_DTYPE)
            h["beg"], h["end"] = b, b + 9
            hs.append(h)
            off.append(off[-1] + len(b))
        parts.append((np.concatenate(hs) if hs else np.zeros(0, HIT_DTYPE), np.array(off)))
    hits, off = D.merge_batch_shards(parts, npat)
    assert off[-1] == sum(len(w) for w in whole)
    for p in range(npat):
        assert np.array_equal(hits["beg"][off[p]:off[p + 1]], whole[p])
        assert np.array_equal(hits["end"][off[p]:off[p + 1]], whole[p] + 9)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100)
    # one global candidate list, cut at the rank boundary of shard_ranges
    n = 100000
    keys = np.sort(rng.choice(n * 16, size=5000, replace=False)).astype(np.int64)
    allc = np.zeros(len(keys), dtype=CAND_DTYPE)
    allc["key"] = keys
    allc["beg"] = keys >> 4
    allc["end"] = (keys >> 4) + 7
    allc["reach"] = (keys >> 4) - 3
    beg, end = D.shard_ranges(n, world)[rank]
    mine = allc[((allc["key"] >> 4) >= beg) & ((allc["key"] >> 4) < end)]
    merged = D.gather_candidates(mine, rank, world)
    if rank == 0:
        q.put(bool(np.array_equal(merged, allc)))
    else:
        q.put(merged is None)
    # an empty shard must not break the gather
    merged = D.gather_candidates(mine if rank == 0 else mine[:0], rank, world)
    if rank == 0:
        q.put(len(merged) == len(mine))
    # dataset assembly: each rank uploads its slice, the all-gather completes every rank's copy
    import torch
    for nbytes in (1, 255, 1000, 70001):
        host = torch.from_numpy(np.random.default_rng(nbytes).integers(0, 256, nbytes, dtype=np.uint8))
        full = D.assemble_from_host(host, rank, world, "cpu")
        q.put(bool(torch.equal(full[:nbytes], host)) and full.numel() % (256 * world) == 0)
    # text-sharded motif batch: every rank "searches" all motifs over its position range (a stand-in engine that cuts
    # a precomputed whole-file result), the totals are all-reduced and rank 0 can ask for the merged lists
    from patmatchdocker_b200._native import HIT_DTYPE
    npat, nfile = 23, 50000
    r2 = np.random.default_rng(5)
    whole = [np.sort(r2.choice(nfile, size=int(r2.integers(0, 60)), replace=False)) for _ in range(npat)]

    class FakeEngine:
        def search_batch(self, dataset, patterns, kopt, copy=True, pos_range=None):
            hs, off = [], [0]
            for b in whole:
                b = b[(b >= pos_range[0]) & (b < pos_range[1])]
                h = np.zeros(len(b), dtype=HIT_DTYPE)
                h["beg"], h["end"] = b, b + 8
                hs.append(h)
                off.append(off[-1] + len(b))
            return np.concatenate(hs), np.array(off, dtype=np.int64)

    class FakeDataset:
        def __len__(self):
            return nfile
    hits, off, totals = D.search_batch_text_sharded(FakeEngine(), FakeDataset(), ["x"] * npat, "0ids", rank, world, gather=True)
    q.put(bool(np.array_equal(totals, [len(w) for w in whole])))
    if rank == 0:
        q.put(all(np.array_equal(hits["beg"][off[p]:off[p + 1]], whole[p]) for p in range(npat)))
    else:
        q.put(hits is None)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_over_gloo_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(3 + 2 * 4 + 2 * 2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)
