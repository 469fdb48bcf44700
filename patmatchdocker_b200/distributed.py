"""Multi-GPU search: one process per GPU (torch.distributed), no data-path collective.

Requests: DeviceShardedSearch.request_fills (buffer fills split over the ranks by position, one all-gather of the hit
lists).  Motif batches: search_batch_text_sharded (every rank all motifs x 1/N of the file).  The first-generation
candidate-gather path is kept below:

Every rank holds the dataset and scans only its share of the file positions
(`Engine.candidates`, i.e. pm_candidates): the reference's scan is embarrassingly parallel up
to the point where it restarts after each reported hit, and that sequential rule is applied
once, on rank 0, to the gathered list of verified candidates (`Engine.resolve`, pm_resolve).
The only communication is the final gather of the (sparse) candidate records over NCCL.
Because candidates do not depend on where the file is cut, the result is bit-identical to the
single-GPU search for any number of ranks (tests/test_gpu.py::test_sharded_candidates_...).
"""
import numpy as np

from ._native import CAND_DTYPE


def shard_ranges(n, world):
    """Position ranges [beg, end) of the `world` ranks over a file of n bytes (anchors 0..n)."""
    edges = [(n + 1) * r // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


def gather_candidates(cands, rank, world, device=None, group=None):
    """Gather per-rank candidate arrays on rank 0 in rank (= file) order. Returns the merged array on
    rank 0 and None elsewhere.  Works with the nccl (device tensors) and gloo (CPU tensors) backends."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return cands
    flat = torch.from_numpy(np.ascontiguousarray(cands).view(np.int64).reshape(-1, 4))
    if device is not None:
        flat = flat.to(device)
    count = torch.tensor([flat.shape[0]], dtype=torch.int64, device=flat.device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    mx = max(int(c) for c in counts)
    pad = torch.zeros((mx, 4), dtype=torch.int64, device=flat.device)
    pad[: flat.shape[0]] = flat
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if rank != 0:
        return None
    parts = [bufs[r][: int(counts[r])].cpu().numpy() for r in range(world)]
    merged = np.concatenate(parts) if parts else np.zeros((0, 4), np.int64)
    return np.ascontiguousarray(merged).view(CAND_DTYPE).reshape(-1)


def search_sharded(engine, dataset, pattern, kopt, rank, world, device=None, group=None):
    """The hit list of Engine.search, computed by `world` ranks; returned on rank 0 (None elsewhere)."""
    beg, end = shard_ranges(len(dataset), world)[rank]
    cands = engine.candidates(dataset, pattern, kopt, beg, end)
    merged = gather_candidates(cands, rank, world, device=device, group=group)
    if rank != 0:
        return None
    return engine.resolve(dataset, pattern, kopt, merged)


def merge_batch_shards(parts, npat):
    """parts[r] = (hits, offsets) of rank r's position range, ranks in file order -> (hits, offsets) of the whole file:
    motif by motif, the ranks' lists one after the other (buffer fills are independent, pm_search_batch_fills)."""
    counts = np.stack([np.diff(np.asarray(off, dtype=np.int64)) for _, off in parts]) if parts else np.zeros((0, npat), np.int64)
    offsets = np.zeros(npat + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(counts.sum(axis=0))
    dtype = parts[0][0].dtype if parts else np.int64
    out = np.empty(int(offsets[-1]), dtype=dtype)
    at = offsets[:-1].copy()                                  # where the next rank's hits of motif p go
    for r, (hits, off) in enumerate(parts):
        off = np.asarray(off, dtype=np.int64)
        c = counts[r]
        if int(off[-1]) == 0:
            continue
        # destination index of every hit of this rank: its motif's write position + its index inside the motif's list
        pid = np.repeat(np.arange(npat), c)
        dst = at[pid] + (np.arange(int(off[-1])) - off[:-1][pid])
        out[dst] = hits[: int(off[-1])]
        at += c
    return out, offsets


def search_batch_text_sharded(engine, dataset, patterns, kopt, rank, world, group=None, gather=False, device=None):
    """A motif batch over `world` ranks, TEXT-sharded: every rank runs ALL motifs over the buffer fills that start in its
    position range (pm_search_batch_fills), no collective on the data path.  -> (hits, offsets, totals): this rank's hit
    lists (file offsets; a view of the engine's page-locked buffer, valid until its next batch) and the per-motif hit
    counts summed over the ranks (one small all-reduce).  gather=True: the lists are also merged on rank 0
    (merge_batch_shards) and returned there instead of the local ones; elsewhere None."""
    import torch
    import torch.distributed as dist
    beg, end = shard_ranges(len(dataset), world)[rank]
    hits, off = engine.search_batch(dataset, patterns, kopt, copy=False, pos_range=(beg, end))
    counts = torch.from_numpy(np.diff(off))
    if world > 1:
        if device is not None:
            counts = counts.to(device)
        dist.all_reduce(counts, group=group)
        counts = counts.cpu()
    totals = counts.numpy()
    if not gather or world == 1:
        return hits, off, totals
    box = [None] * world if rank == 0 else None
    dist.gather_object((np.array(hits, copy=True), off), box, dst=0, group=group)
    if rank != 0:
        return None, None, totals
    mh, mo = merge_batch_shards(box, len(patterns))
    return mh, mo, totals


class PipelinedBatch:
    """A motif batch over a position range, cut into `parts` sub-ranges that alternate between two engines of the same GPU
    (own stream and scratch each), driven by two host threads: the result copy of one sub-range (4 bytes per hit, but
    10^8 hits) crosses PCIe while the other engine scans, sorts and chains the next one.  Buffer fills are independent,
    so the sub-range lists are the whole-range lists cut by position (pm_search_batch_fills_compact).

    engines / datasets: two Engine objects on one device and the same file as a Dataset of each (Engine.wrap_device on
    one device buffer shares the text; the planes are packed per engine)."""

    def __init__(self, engines, datasets, parts=4):
        from ._native import CompactBuffer
        assert len(engines) == len(datasets) >= 1
        self.engines, self.datasets, self.parts = engines, datasets, max(int(parts), 1)
        self.buffers = [CompactBuffer() for _ in range(self.parts)]

    def search(self, patterns, kopt="0ids", pos_range=None):
        """-> list over the sub-ranges, in file order, of (begins, offsets, base, motif_len) as Engine.search_batch_compact
        returns them; the arrays are views of page-locked buffers this object reuses on its next search."""
        import threading
        n = len(self.datasets[0])
        beg, end = (0, n + 1) if pos_range is None else (int(pos_range[0]), int(pos_range[1]))
        edges = [beg + (end - beg) * i // self.parts for i in range(self.parts + 1)]
        res = [None] * self.parts
        errors = []

        def worker(t):
            try:
                for i in range(t, self.parts, len(self.engines)):
                    res[i] = self.engines[t].search_batch_compact(self.datasets[t], patterns, kopt, pos_range=(edges[i], edges[i + 1]), out=self.buffers[i])
            except Exception as ex:                          # noqa: BLE001 -- re-raised on the caller's thread
                errors.append(ex)
        threads = [threading.Thread(target=worker, args=(t,)) for t in range(1, len(self.engines))]
        for th in threads:
            th.start()
        worker(0)
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
        return res


def assemble_from_host(host, rank, world, device, full=None, group=None):
    """Bring a file that sits in (pinned) host memory into the memory of every rank with ONE trip over PCIe:
    rank r copies only its 1/world slice host->device, then the slices are all-gathered in place over
    NVLink/NVSwitch (NCCL).  `host` is a 1-D uint8 torch tensor, identical on all ranks.  Returns a device
    tensor whose first len(host) bytes are the file (the tail is padding)."""
    import torch
    import torch.distributed as dist
    n = int(host.numel())
    per = ((n + world - 1) // world + 255) // 256 * 256
    if full is None or full.numel() != per * world or full.device != torch.device(device):
        full = torch.empty(per * world, dtype=torch.uint8, device=device)
    lo = min(rank * per, n)
    hi = min(n, lo + per)
    mine = full[rank * per:(rank + 1) * per]
    if hi > lo:
        mine[: hi - lo].copy_(host[lo:hi], non_blocking=True)
    if world > 1:
        dist.all_gather_into_tensor(full, mine, group=group)
    return full


class DeviceShardedSearch:
    """search_sharded with the candidates kept in device memory end to end (nccl backend):
    pm_candidates_device -> dist.gather of device tensors -> pm_resolve_device on rank 0."""

    def __init__(self, engine, rank, world, device, cap=1 << 16, group=None):
        import torch
        self.engine, self.rank, self.world, self.device, self.group = engine, rank, world, device, group
        self.torch = torch
        if str(device) != "cpu":
            # torch (NCCL collectives, copy_) and the engine must work on ONE stream: the engine reads buffers that
            # torch has only enqueued writes for, and vice versa.  Without this the engine would launch on its own
            # non-blocking stream and race with them.
            engine.use_torch_stream(torch.cuda.current_stream(torch.device(device)))
        self._alloc(cap)

    def request_fills(self, dataset, patterns, kopt):
        """One PatMatch request (all patterns, e.g. motif + reverse complement) over `world` ranks with ONE collective:
        every rank runs pm_request_fills_device on the buffer fills that start in its position range -- a single
        pass over its part of the planes, one sort / verify / chain / select, no host synchronisation -- and the
        per-rank [header | hits] blocks are all-gathered over NVLink.  Rank 0 merges them into per-pattern lists on
        the device (pm_merge_request_shards: driven by the gathered headers, no host round trip), copies headers and
        lists to the host in one go and returns the lists (== Engine.search_request); the other ranks read only the
        header rows (to agree on a retry when some rank needed more room) and return None.  The arrays returned on
        rank 0 are views of a page-locked buffer that the next call reuses: copy them to keep them."""
        import torch.distributed as dist
        from ._native import HIT_DTYPE, request_header_rows
        torch = self.torch
        npat = len(patterns)
        hr = request_header_rows(npat)
        on_gpu = str(self.device) != "cpu"
        if not hasattr(self, "rq_rows"):
            self.rq_rows, self.rq_cap, self.rq_alloc, self.rq_spec = 1 << 14, 1 << 16, 0, 1 << 14
            self.rq_merged = self.rq_out = None
        beg, end = shard_ranges(len(dataset), self.world)[self.rank]
        device_merge = on_gpu and self.rank == 0 and npat <= 64 and self.world <= 16
        spec = 0
        while True:
            rows = max(self.rq_rows, hr + 16)
            if rows > self.rq_alloc:
                self.rq_alloc = rows + rows // 2
                self.rq_mine = torch.zeros((self.rq_alloc, 2), dtype=torch.int64, device=self.device)
                self.rq_all = torch.empty((self.rq_alloc * self.world, 2), dtype=torch.int64, device=self.device)
                self.rq_hdr = torch.empty((self.world, hr, 2), dtype=torch.int64, pin_memory=on_gpu)
            if self.rq_hdr.shape[1] != hr:
                self.rq_hdr = torch.empty((self.world, hr, 2), dtype=torch.int64, pin_memory=on_gpu)
            mine = self.rq_mine[:rows]
            self.engine.request_fills_device(dataset, patterns, kopt, beg, end, self.rq_cap, mine.data_ptr(), rows)
            flat = self.rq_all[: self.world * rows]
            if self.world > 1:
                dist.all_gather_into_tensor(flat, mine, group=self.group)
            else:
                flat.copy_(mine)
            view = flat.view(self.world, rows, 2)
            if device_merge:
                # [world headers | per-pattern lists] built on the device; headers + as many hits as the last request
                # had (with head-room) cross PCIe before the host has seen a single count
                need = self.world * rows
                if self.rq_merged is None or self.rq_merged.shape[0] < need:
                    self.rq_merged = torch.empty((need + need // 2, 2), dtype=torch.int64, device=self.device)
                if self.rq_out is None or self.rq_out.shape[0] < need:
                    self.rq_out = torch.empty((need + need // 2, 2), dtype=torch.int64, pin_memory=True)
                self.engine.merge_request_shards(flat.data_ptr(), self.world, rows, npat, self.rq_merged.data_ptr(), self.rq_merged.shape[0])
                spec = min(self.rq_spec, self.world * (rows - hr))
                ncopy = self.world * hr + spec
                self.rq_out[:ncopy].copy_(self.rq_merged[:ncopy], non_blocking=True)
                torch.cuda.current_stream().synchronize()
                hdrs = self.rq_out[: self.world * hr].numpy().reshape(self.world, hr, 2).copy()
            else:
                # every rank reads the header rows (a few hundred bytes): they agree on a retry and on the next sizes
                self.rq_hdr.copy_(view[:, :hr], non_blocking=True)
                if on_gpu:
                    torch.cuda.current_stream().synchronize()
                hdrs = self.rq_hdr.numpy()
            nh = hdrs[:, 0, 0]
            ncand = hdrs[:, 0, 1]
            ok = int(ncand.max()) <= self.rq_cap and int(nh.max()) + hr <= rows
            # agreed by construction: every rank saw the same headers
            self.rq_cap = max(1 << 12, int(int(ncand.max()) * 1.25) + 1024)
            self.rq_rows = max(256, int((int(nh.max()) + hr) * 1.25) + 16)
            if ok:
                break
        if self.rank != 0:
            return None
        counts = hdrs[:, 2:hr].reshape(self.world, -1)[:, :npat]
        total = int(counts.sum())
        if total == 0:
            return [np.zeros(0, dtype=HIT_DTYPE) for _ in range(npat)]
        if device_merge:
            h0 = self.world * hr
            if total > spec:                                  # more hits than guessed: fetch the rest
                self.rq_out[h0 + spec: h0 + total].copy_(self.rq_merged[h0 + spec: h0 + total], non_blocking=True)
                torch.cuda.current_stream().synchronize()
            self.rq_spec = total + total // 4 + 256
            flat_np = self.rq_out[h0: h0 + total].numpy().view(HIT_DTYPE).reshape(-1)
        else:
            # the per-pattern lists are put together with torch (pattern-major, ranks in file order) and cross PCIe
            # once, exactly sized
            offs = np.zeros((self.world, npat + 1), dtype=np.int64)
            offs[:, 1:] = np.cumsum(counts, axis=1)
            parts = [view[r, hr + int(offs[r, p]): hr + int(offs[r, p + 1])] for p in range(npat) for r in range(self.world)]
            merged = torch.cat(parts)
            if self.rq_out is None or self.rq_out.shape[0] < total:
                self.rq_out = torch.empty((total + total // 4 + 256, 2), dtype=torch.int64, pin_memory=on_gpu)
            hout = self.rq_out[:total]
            hout.copy_(merged, non_blocking=True)
            if on_gpu:
                torch.cuda.current_stream().synchronize()
            flat_np = hout.numpy().view(HIT_DTYPE).reshape(-1)
        out, at = [], 0
        for p in range(npat):
            c = int(counts[:, p].sum())
            out.append(flat_np[at:at + c])
            at += c
        return out

    def _alloc(self, cap):
        torch = self.torch
        self.cap = cap
        self.rows = getattr(self, "rows", cap)                # first exchange: full buffers, then adapted
        self.mine = torch.zeros((cap, 4), dtype=torch.int64, device=self.device)          # row 0 = header (count)
        self.allbuf = torch.empty((cap * self.world, 4), dtype=torch.int64, device=self.device)
        self.merged = torch.empty((cap * self.world, 4), dtype=torch.int64, device=self.device) if self.rank == 0 else None

    def load_dataset(self, host):
        """Dataset from file bytes in host memory (1-D uint8 torch tensor, pinned for full speed): every rank
        uploads 1/world of it and NCCL all-gathers the rest; each rank then packs its own copy."""
        self.full = assemble_from_host(host, self.rank, self.world, self.device, getattr(self, "full", None), self.group)
        return self.engine.wrap_device(self.full.data_ptr(), int(host.numel()))

    def load_window(self, host, bufsize=1600000, nl_rows=1 << 16):
        """Cold multi-GPU request: every rank uploads and packs only the bytes its own buffer fills can touch -- its
        position range plus one buffer fill of overlap -- over its own PCIe link (pm_dataset_create_window); nothing
        but the newline positions (a few KB, one all-gather) is exchanged, because the fill table of the reference
        depends on every newline of the file.  `host`: 1-D uint8 torch tensor holding the file, pinned, identical on
        all ranks.  Returns a windowed Dataset that request_fills() can search."""
        import torch.distributed as dist
        torch = self.torch
        n = int(host.numel())
        beg, end = shard_ranges(n, self.world)[self.rank]
        lo, hi = max(beg - 4096, 0), min(n, end + bufsize + 4096)
        if getattr(self, "nl_mine", None) is None or self.nl_mine.shape[0] != nl_rows:
            self.nl_mine = torch.zeros(nl_rows, dtype=torch.int64, device=self.device)
            self.nl_all = torch.zeros(nl_rows * self.world, dtype=torch.int64, device=self.device)
            self.nl_host = torch.zeros(nl_rows * self.world, dtype=torch.int64, pin_memory=str(self.device) != "cpu")
        ds = self.engine.load_window(host.numpy(), lo, hi, self.nl_mine.data_ptr(), nl_rows)
        if self.world > 1:
            dist.all_gather_into_tensor(self.nl_all, self.nl_mine, group=self.group)
        else:
            self.nl_all.copy_(self.nl_mine)
        self.nl_host.copy_(self.nl_all, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        h = self.nl_host.numpy().reshape(self.world, nl_rows)
        if int(h[:, 0].max()) > nl_rows - 1:
            ds.close()
            raise RuntimeError("more than %d newlines in one rank's window: pass a larger nl_rows" % (nl_rows - 1))
        allnl = np.unique(np.concatenate([h[r, 1:1 + int(h[r, 0])] for r in range(self.world)]))     # windows overlap: unique
        self.engine.set_newlines(ds, allnl)
        return ds

    def search_fills(self, dataset, pattern, kopt):
        """Fill-sharded search (pm_search_fills_device): the reference restarts its scan at every buffer fill, so
        every rank searches the fills that start in its position range completely, chain stage included, and
        the only exchange is one all-gather of the per-rank hit lists (16 B per hit; a header row, written on the
        device, carries the count).  Rank 0 returns the hit list of Engine.search as a view of a pinned buffer
        that stays valid until the next call; the other ranks return None."""
        import torch.distributed as dist
        torch = self.torch
        from ._native import HIT_DTYPE
        if not hasattr(self, "hcap"):
            self.hcap = 1 << 16
            self.hrows = self.hcap
            self.hmine = None
        beg, end = shard_ranges(len(dataset), self.world)[self.rank]
        on_gpu = str(self.device) != "cpu"
        while True:
            if self.hmine is None or self.hmine.shape[0] != self.hcap:
                self.hmine = torch.zeros((self.hcap, 2), dtype=torch.int64, device=self.device)
                self.hall = torch.empty((self.hcap * self.world, 2), dtype=torch.int64, device=self.device)
                self.hhost = torch.empty((self.hcap * self.world, 2), dtype=torch.int64, pin_memory=on_gpu) if self.rank == 0 else None
            n = self.engine.search_fills_device(dataset, pattern, kopt, beg, end, self.hmine[1:].data_ptr(), self.hcap - 1,
                                                self.hmine.data_ptr())
            rows = min(self.hrows, self.hcap)
            flat = self.hall[: self.world * rows]
            view = flat.view(self.world, rows, 2)
            if self.world > 1:
                dist.all_gather_into_tensor(view, self.hmine[:rows], group=self.group)
            else:
                view[0] = self.hmine[:rows]
            if self.rank == 0:                                     # one D2H of everything gathered, one synchronisation
                hview = self.hhost[: self.world * rows]
                hview.copy_(flat, non_blocking=True)
                if on_gpu:
                    torch.cuda.current_stream().synchronize()
                counts = hview.view(self.world, rows, 2)[:, 0, 0].tolist()
            else:
                counts = view[:, 0, 0].tolist()
            need = max(counts) + 1
            self.hrows = max(256, int(need * 1.25) + 16)          # agreed by construction: everyone saw the same counts
            if need <= rows:
                break
            if need > self.hcap:                                   # some rank overflowed: grow everywhere and search again
                self.hcap = need + 1024
        if self.rank != 0:
            return None
        h = self.hhost[: self.world * rows].numpy().reshape(self.world, rows, 2)
        if self.world == 1:
            out = h[0, 1:1 + counts[0]]
        else:
            out = np.concatenate([h[r, 1:1 + int(counts[r])] for r in range(self.world)])
        return np.ascontiguousarray(out).view(HIT_DTYPE).reshape(-1)

    def search(self, dataset, pattern, kopt):
        """One collective per search: every rank all-gathers `rows` candidate records plus a header row
        holding its true count, so all ranks learn all counts and agree on the size of the next exchange
        (and on a retry when a rank had more candidates than rows)."""
        import torch.distributed as dist
        torch = self.torch
        beg, end = shard_ranges(len(dataset), self.world)[self.rank]
        n = self.engine.candidates_device(dataset, pattern, kopt, beg, end, self.mine[1:].data_ptr(), self.cap - 1)
        while True:
            self.mine[0, 0] = abs(n)
            rows = min(self.rows, self.cap)
            dist.all_gather_into_tensor(self.allbuf[: self.world * rows].view(self.world, rows, 4), self.mine[:rows], group=self.group)
            counts = self.allbuf[: self.world * rows].view(self.world, rows, 4)[:, 0, 0].tolist()
            need = max(counts) + 1
            self.rows = max(256, int(need * 1.25) + 16)      # agreed by construction: everyone saw the same counts
            if need <= rows:
                break
            if need > self.cap:                                # some rank overflowed its buffer: grow everywhere, rescan
                self._alloc(need + 1024)
                n = self.engine.candidates_device(dataset, pattern, kopt, beg, end, self.mine[1:].data_ptr(), self.cap - 1)
        if self.rank != 0:
            return None
        view = self.allbuf[: self.world * rows].view(self.world, rows, 4)
        off = 0
        for r in range(self.world):
            c = int(counts[r])
            self.merged[off:off + c] = view[r, 1:1 + c]
            off += c
        return self.engine.resolve_device(dataset, pattern, kopt, self.merged.data_ptr(), off)
