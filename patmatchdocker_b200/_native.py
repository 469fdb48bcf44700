"""ctypes binding of lib/libpatmatch_b200.so (C ABI: include/patmatch_b200.h)."""
import ctypes
import os
import threading
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PM_MAX_PIECES = 16
PLAN_NAMES = {0: "SIMPLE", 1: "SPLIT", 2: "BWD", 3: "FWD", 4: "EXT_BEG", 5: "EXT_END"}
PM_ERR_OVERFLOW = -5


class NativeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("patmatch_b200 error %d: %s" % (code, msg))
        self.code = code


class PmHit(ctypes.Structure):
    _fields_ = [("beg", ctypes.c_int64), ("end", ctypes.c_int64)]


class PmCandidate(ctypes.Structure):
    _fields_ = [("key", ctypes.c_int64), ("beg", ctypes.c_int64), ("end", ctypes.c_int64), ("reach", ctypes.c_int64)]


class PmPlanInfo(ctypes.Structure):
    _fields_ = [("m", ctypes.c_int), ("k", ctypes.c_int), ("ins", ctypes.c_int), ("del_", ctypes.c_int),
                ("subs", ctypes.c_int), ("type", ctypes.c_int), ("L", ctypes.c_int), ("npieces", ctypes.c_int),
                ("V", ctypes.c_int * PM_MAX_PIECES), ("split_cost", ctypes.c_double), ("fb_cost", ctypes.c_double)]


class PmStats(ctypes.Structure):
    _fields_ = [("scan_ms", ctypes.c_float), ("sort_ms", ctypes.c_float), ("verify_ms", ctypes.c_float),
                ("chain_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
                ("candidates", ctypes.c_int64), ("verified", ctypes.c_int64), ("hits", ctypes.c_int64),
                ("scan_bytes", ctypes.c_int64), ("scan_bases", ctypes.c_int64), ("launches", ctypes.c_int),
                ("packed", ctypes.c_int), ("qgram_chunks", ctypes.c_int), ("syncs", ctypes.c_int), ("jit", ctypes.c_int)]


HIT_DTYPE = np.dtype([("beg", "<i8"), ("end", "<i8")])
CAND_DTYPE = np.dtype([("key", "<i8"), ("beg", "<i8"), ("end", "<i8"), ("reach", "<i8")])

_lib = None


def lib_path():
    return os.path.join(_HERE, "lib", "libpatmatch_b200.so")


def load():
    """Load the native library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise NativeError(-1, "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    L = ctypes.CDLL(path)
    vp, cp, i64 = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64
    L.pm_last_error.restype = cp
    L.pm_version.restype = cp
    L.pm_plan.argtypes = [cp, cp, ctypes.POINTER(PmPlanInfo)]
    L.pm_engine_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    L.pm_engine_destroy.argtypes = [vp]
    L.pm_engine_destroy.restype = None
    L.pm_engine_set_stream.argtypes = [vp, vp]
    L.pm_engine_synchronize.argtypes = [vp]
    L.pm_engine_set_scan_mode.argtypes = [vp, ctypes.c_int]
    L.pm_engine_set_buffer_size.argtypes = [vp, i64]
    L.pm_engine_set_fused_filter.argtypes = [vp, ctypes.c_int]
    L.pm_engine_set_jit.argtypes = [vp, ctypes.c_int]
    L.pm_jit_wait.argtypes = []
    L.pm_merge_request_shards.argtypes = [vp, vp, ctypes.c_int, i64, ctypes.c_int, vp, i64]
    L.pm_engine_set_peptide_codes.argtypes = [vp, ctypes.c_int]
    L.pm_engine_set_batch_lookup.argtypes = [vp, ctypes.c_int]
    L.pm_dataset_create_window.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, vp, ctypes.c_int64, ctypes.POINTER(vp)]
    L.pm_dataset_set_newlines.argtypes = [vp, vp, vp, ctypes.c_int64]
    L.pm_jit_source.restype = ctypes.c_int64
    L.pm_jit_source.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64]
    L.pm_search_fills_device.argtypes = [vp, vp, ctypes.c_char_p, ctypes.c_char_p, i64, i64, vp, i64, ctypes.POINTER(i64), vp]
    L.pm_search_stream.argtypes = [vp, vp, i64, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.c_char_p, i64, vp, i64,
                                   ctypes.POINTER(i64), ctypes.POINTER(vp)]
    L.pm_dataset_create.argtypes = [vp, vp, i64, ctypes.POINTER(vp)]
    L.pm_dataset_wrap_device.argtypes = [vp, vp, i64, ctypes.POINTER(vp)]
    L.pm_dataset_destroy.argtypes = [vp]
    L.pm_dataset_destroy.restype = None
    L.pm_dataset_size.argtypes = [vp]
    L.pm_dataset_size.restype = i64
    L.pm_search.argtypes = [vp, vp, cp, cp, vp, i64, ctypes.POINTER(i64)]
    L.pm_last_hits.argtypes = [vp, vp, i64, ctypes.POINTER(i64)]
    L.pm_search_batch.argtypes = [vp, vp, ctypes.c_int, ctypes.POINTER(cp), cp, vp, i64, ctypes.POINTER(i64)]
    L.pm_search_batch_fills.argtypes = [vp, vp, ctypes.c_int, ctypes.POINTER(cp), cp, i64, i64, vp, i64, ctypes.POINTER(i64)]
    L.pm_search_batch_fills_compact.argtypes = [vp, vp, ctypes.c_int, ctypes.POINTER(cp), cp, i64, i64, vp, i64, ctypes.POINTER(i64), ctypes.POINTER(i64), vp]
    L.pm_search_request.argtypes = [vp, vp, ctypes.c_int, ctypes.POINTER(cp), cp, vp, i64, ctypes.POINTER(i64)]
    L.pm_request_fills_device.argtypes = [vp, vp, ctypes.c_int, ctypes.POINTER(cp), cp, i64, i64, i64, vp, i64]
    L.pm_candidates.argtypes = [vp, vp, cp, cp, i64, i64, vp, i64, ctypes.POINTER(i64)]
    L.pm_resolve.argtypes = [vp, vp, cp, cp, vp, i64, vp, i64, ctypes.POINTER(i64)]
    L.pm_candidates_device.argtypes = [vp, vp, cp, cp, i64, i64, vp, i64, ctypes.POINTER(i64)]
    L.pm_resolve_device.argtypes = [vp, vp, cp, cp, vp, i64, vp, i64, ctypes.POINTER(i64)]
    L.pm_host_alloc.argtypes = [i64]
    L.pm_host_alloc.restype = vp
    L.pm_host_free.argtypes = [vp]
    L.pm_host_free.restype = None
    L.pm_get_stats.argtypes = [vp, ctypes.POINTER(PmStats)]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise NativeError(rc, load().pm_last_error().decode("utf-8", "replace"))


def _b(s):
    return s if isinstance(s, bytes) else s.encode("latin-1")


def _pattern_array(owner, patterns):
    """char*[] of a motif list; the array of the previous call is reused when the list has not changed (a motif
    library searched again and again: marshalling 10 000 strings costs more than a millisecond)"""
    key = tuple(patterns)
    cached = getattr(owner, "_pat_cache", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    arr = (ctypes.c_char_p * len(key))(*[_b(p) for p in key])
    owner._pat_cache = (key, arr)
    return arr


class _Pinned:
    """page-locked host memory from pm_host_alloc.  `.array` hands out numpy views whose base keeps this object alive, so
    the memory is freed with the last VIEW, not with the last reference to the holder (a result array that outlives the
    engine or buffer object it came from stays valid)."""

    def __init__(self, count, dtype):
        self.count, self.dtype = max(int(count), 1), dtype
        self.nbytes = self.count * dtype.itemsize
        self.ptr = load().pm_host_alloc(self.nbytes)
        if not self.ptr:
            raise NativeError(-1, "pm_host_alloc(%d) failed" % self.nbytes)

    @property
    def array(self):
        buf = (ctypes.c_char * self.nbytes).from_address(self.ptr)
        buf._owner = self                 # view -> base array -> buf -> this object (no cycle: nothing here points back)
        return np.frombuffer(buf, dtype=self.dtype, count=self.count)

    def __del__(self):
        try:
            load().pm_host_free(self.ptr)
        except Exception:
            pass


class CompactBuffer:
    """page-locked result buffer of Engine.search_batch_compact (allocated and grown on demand)"""

    def __init__(self):
        self.pinned = None


def pinned_empty(count, dtype):
    """-> (array, keepalive): array lives in page-locked memory as long as keepalive does"""
    p = _Pinned(count, dtype)
    return p.array, p


def set_compat_deployed_glibc(on):
    """pm_set_compat_deployed_glibc: piece choice of the deployed binary (default glibc allocator) instead of the
    defined zero-scratch behaviour; process-wide."""
    _check(load().pm_set_compat_deployed_glibc(1 if on else 0))


def plan(pattern, kopt):
    """Host-only: the search plan the reference's esimplePreproc picks for (pattern, -k kopt)."""
    info = PmPlanInfo()
    _check(load().pm_plan(_b(pattern), _b(kopt), ctypes.byref(info)))
    return {"m": info.m, "k": info.k, "ins": info.ins, "del": info.del_, "subs": info.subs,
            "type": PLAN_NAMES[info.type], "L": info.L, "V": list(info.V)[:info.npieces],
            "split_cost": info.split_cost, "fb_cost": info.fb_cost}


class Dataset:
    """A .seq file resident in HBM (replaces the '<datafile>' argument of nrgrep_coords)."""

    def __init__(self, engine, handle, host_bytes=None):
        self.engine, self._h, self.host_bytes = engine, handle, host_bytes
        engine._datasets.add(self)             # an engine closes its datasets before it goes away (pm_dataset holds a pointer to it)

    def __len__(self):
        return load().pm_dataset_size(self._h)

    def close(self):
        if self._h:
            if self.engine._h:                  # engine already destroyed: the handle points into freed memory
                load().pm_dataset_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def request_header_rows(npat):
    """rows (2 x int64) of the header in front of the hit list of pm_request_fills_device"""
    return (4 + npat + 1) // 2


def _locked(fn):
    def wrapper(self, *a, **kw):
        with self._lock:
            return fn(self, *a, **kw)
    wrapper.__name__, wrapper.__doc__ = fn.__name__, fn.__doc__
    return wrapper


class Engine:
    """One engine per GPU (one process per GPU in multi-GPU runs).

    An engine owns mutable scratch (candidate buffers, pinned staging, statistics), so calls are serialised: the C
    entry points take the engine's mutex, and the methods here that combine several calls (a search followed by
    pm_last_hits after an overflow, ...) hold `self._lock` across the combination.  Safe to share between the request
    threads of a WSGI process; use one engine per thread for concurrency."""

    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        self._lock = threading.RLock()
        self._datasets = weakref.WeakSet()
        _check(load().pm_engine_create(int(device), ctypes.byref(self._h)))
        self.device = device

    @_locked
    def search_request(self, dataset, patterns, kopt="0ids", out=None, cap=1 << 16):
        """pm_search_request: all patterns of one PatMatch request (the motif and its reverse complement) in ONE pass
        over the dataset and one pipeline.  -> list of hit arrays, one per pattern, each equal to search() of it.
        `out`: optional page-locked result array (pinned_empty(...)[0], HIT_DTYPE) that the device-to-host copy fills
        directly; the returned arrays are then views of it, valid until the caller reuses it."""
        L = load()
        arr = (ctypes.c_char_p * len(patterns))(*[_b(p) for p in patterns])
        offs = (ctypes.c_int64 * (len(patterns) + 1))()
        hits = out if out is not None else np.empty(cap, dtype=HIT_DTYPE)
        rc = L.pm_search_request(self._h, dataset._h, len(patterns), arr, _b(kopt), ctypes.c_void_p(hits.ctypes.data), len(hits), offs)
        if rc == PM_ERR_OVERFLOW:                      # the list is still on the device
            n = ctypes.c_int64()
            hits = np.empty(int(offs[len(patterns)]), dtype=HIT_DTYPE)
            rc = L.pm_last_hits(self._h, ctypes.c_void_p(hits.ctypes.data), len(hits), ctypes.byref(n))
        _check(rc)
        return [hits[offs[i]:offs[i + 1]] for i in range(len(patterns))]

    def request_fills_device(self, dataset, patterns, kopt, pos_beg, pos_end, sort_cap, dev_ptr, out_rows):
        """pm_request_fills_device: asynchronous; header + hits of the fills starting in [pos_beg, pos_end) are written
        to device memory at dev_ptr (out_rows rows of 2 x int64).  Nothing is returned: read the header after your own
        synchronisation (row 0 = hits, candidates; per-pattern counts from row 2)."""
        arr = (ctypes.c_char_p * len(patterns))(*[_b(p) for p in patterns])
        _check(load().pm_request_fills_device(self._h, dataset._h, len(patterns), arr, _b(kopt), int(pos_beg), int(pos_end),
                                              int(sort_cap), ctypes.c_void_p(dev_ptr), int(out_rows)))

    def merge_request_shards(self, all_ptr, world, rows, npat, out_ptr, out_rows):
        """pm_merge_request_shards: asynchronous device-side merge of the all-gathered [header | hits] blocks into
        [world headers | per-pattern lists of the whole file] at out_ptr."""
        _check(load().pm_merge_request_shards(self._h, ctypes.c_void_p(all_ptr), int(world), int(rows), int(npat), ctypes.c_void_p(out_ptr), int(out_rows)))

    def close(self):
        if self._h:
            for d in list(self._datasets):
                d.close()
            load().pm_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_handle):
        """launch on this cudaStream_t (0 / None = back to the engine's own non-blocking stream)"""
        _check(load().pm_engine_set_stream(self._h, ctypes.c_void_p(cuda_stream_handle or 0)))

    def use_torch_stream(self, stream=None):
        """Share torch's stream (default: the current one), so that engine work and torch work (NCCL collectives,
        copy_) are ordered.  torch's default stream has handle 0 = the legacy default stream, which the C ABI
        spells cudaStreamLegacy (1) because 0 means "the engine's own stream" there."""
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(torch.device("cuda", self.device))
        self.set_stream(s.cuda_stream or 1)

    @_locked
    def search_stream(self, data, patterns, kopt="0ids", chunk_bytes=0, cap=1 << 18):
        """pm_search_stream: `data` (bytes or a uint8 numpy array, pinned for full speed) is uploaded in chunks while
        the chunks that have arrived are packed and searched.  -> (Dataset resident for later searches,
        [hit array per pattern])."""
        L = load()
        buf = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.ascontiguousarray(data, dtype=np.uint8)
        arr = (ctypes.c_char_p * len(patterns))(*[_b(p) for p in patterns])
        offs = (ctypes.c_int64 * (len(patterns) + 1))()
        h = ctypes.c_void_p()
        hits = np.empty(cap, dtype=HIT_DTYPE)            # filled by memcpy from host vectors: pageable is fine
        rc = L.pm_search_stream(self._h, ctypes.c_void_p(buf.ctypes.data), buf.size, len(patterns), arr, _b(kopt), int(chunk_bytes),
                                ctypes.c_void_p(hits.ctypes.data), cap, offs, ctypes.byref(h))
        if rc == PM_ERR_OVERFLOW:                      # the dataset is resident: fetch the lists with ordinary searches
            ds = Dataset(self, h)
            return ds, [self.search(ds, p, kopt, cap=max(int(offs[len(patterns)]), 1)) for p in patterns]
        _check(rc)
        ds = Dataset(self, h)
        ds._keep = buf
        return ds, [hits[offs[i]:offs[i + 1]] for i in range(len(patterns))]

    def search_fills_device(self, dataset, pattern, kopt, pos_beg, pos_end, dev_ptr, cap, dev_count_ptr=None):
        """pm_search_fills_device: hits of the fills starting in [pos_beg, pos_end) into device memory (cap records
        of 2 x int64). Returns the number of hits, negated when it exceeds cap (nothing was copied then); the count
        is also stored at dev_count_ptr (device memory) when given."""
        n = ctypes.c_int64()
        rc = load().pm_search_fills_device(self._h, dataset._h, _b(pattern), _b(kopt), int(pos_beg), int(pos_end),
                                           ctypes.c_void_p(dev_ptr), int(cap), ctypes.byref(n),
                                           ctypes.c_void_p(dev_count_ptr) if dev_count_ptr else None)
        if rc == PM_ERR_OVERFLOW:
            return -int(n.value)
        _check(rc)
        return int(n.value)

    def set_scan_mode(self, mode):
        """'auto' | 'bytes' | 'packed' -- which scan kernel SIMPLE/SPLIT plans use."""
        _check(load().pm_engine_set_scan_mode(self._h, {"auto": 0, "bytes": 1, "packed": 2}[mode]))

    def set_fused_filter(self, on):
        """True/1 = q-gram pre-filter + Myers filter (default), 2 = Myers filter only, False/0 = off"""
        _check(load().pm_engine_set_fused_filter(self._h, int(on)))

    def set_batch_lookup(self, on):
        """batches of >= 64 exact motifs: q-gram lookup kernel (default) or the dense multi-pattern kernel"""
        _check(load().pm_engine_set_batch_lookup(self._h, 1 if on else 0))

    def set_peptide_codes(self, on):
        """proteomes: scan the 5-bit residue codes (default) or the raw bytes"""
        _check(load().pm_engine_set_peptide_codes(self._h, 1 if on else 0))

    def set_jit(self, mode):
        """'off' | 'auto' | 'always' -- per-request specialised scan kernels (NVRTC); results are identical"""
        _check(load().pm_engine_set_jit(self._h, {"off": 0, "auto": 1, "always": 2}[mode]))

    def set_buffer_size(self, nbytes):
        """the reference's -b (bytes); patmatch.py uses 1600000, the default"""
        _check(load().pm_engine_set_buffer_size(self._h, int(nbytes)))

    def synchronize(self):
        _check(load().pm_engine_synchronize(self._h))

    def load_window(self, data, win_lo, win_hi, dev_newlines_ptr, nl_rows):
        """pm_dataset_create_window: `data` = the whole file in host memory (numpy uint8, pinned for an asynchronous
        copy); only [win_lo, win_hi) is uploaded and packed.  dev_newlines_ptr: device memory (nl_rows x int64) that
        receives the count and the positions of the window's newlines.  Asynchronous."""
        arr = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
        h = ctypes.c_void_p()
        _check(load().pm_dataset_create_window(self._h, ctypes.c_void_p(arr.ctypes.data), arr.size, int(win_lo), int(win_hi),
                                               ctypes.c_void_p(dev_newlines_ptr), int(nl_rows), ctypes.byref(h)))
        return Dataset(self, h, arr)

    def set_newlines(self, dataset, positions):
        """pm_dataset_set_newlines: sorted newline positions of the whole file (numpy int64)"""
        pos = np.ascontiguousarray(positions, dtype=np.int64)
        _check(load().pm_dataset_set_newlines(self._h, dataset._h, ctypes.c_void_p(pos.ctypes.data), pos.size))

    def load_dataset(self, data):
        """data: bytes / bytearray / numpy uint8 array with the .seq file contents (host memory)."""
        arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
        h = ctypes.c_void_p()
        _check(load().pm_dataset_create(self._h, ctypes.c_void_p(arr.ctypes.data), arr.size, ctypes.byref(h)))
        return Dataset(self, h, arr)

    def wrap_device(self, device_ptr, nbytes):
        h = ctypes.c_void_p()
        _check(load().pm_dataset_wrap_device(self._h, ctypes.c_void_p(device_ptr), nbytes, ctypes.byref(h)))
        return Dataset(self, h)

    @_locked
    def search(self, dataset, pattern, kopt="0ids", cap=1 << 16):
        """-> numpy structured array (beg, end): the '[beg, end]' pairs nrgrep_coords prints, in order."""
        L = load()
        n = ctypes.c_int64()
        hits = np.empty(cap, dtype=HIT_DTYPE)
        rc = L.pm_search(self._h, dataset._h, _b(pattern), _b(kopt), ctypes.c_void_p(hits.ctypes.data), cap, ctypes.byref(n))
        if rc == PM_ERR_OVERFLOW:                      # the hit list is still on the device: fetch it, do not search again
            hits = np.empty(int(n.value), dtype=HIT_DTYPE)
            rc = L.pm_last_hits(self._h, ctypes.c_void_p(hits.ctypes.data), len(hits), ctypes.byref(n))
        _check(rc)
        return hits[: n.value]

    @_locked
    def count(self, dataset, pattern, kopt="0ids"):
        n = ctypes.c_int64()
        _check(load().pm_search(self._h, dataset._h, _b(pattern), _b(kopt), None, 0, ctypes.byref(n)))
        return n.value

    @_locked
    def search_batch(self, dataset, patterns, kopt="0ids", cap=1 << 20, copy=True, pos_range=None):
        """-> (hits, offsets): hits[offsets[i]:offsets[i+1]] is the hit list of patterns[i].  Large results
        cross PCIe into a page-locked staging buffer owned by the engine; the caller receives its own copy unless it
        passes copy=False (then the array is a view of that buffer and only valid until the next search_batch).
        pos_range=(beg, end): pm_search_batch_fills -- only the buffer fills that START in that position range (a
        text-sharded batch: the per-range lists of a partition of the file concatenate to the whole-file lists)."""
        L = load()
        arr = _pattern_array(self, patterns)
        offsets = (ctypes.c_int64 * (len(patterns) + 1))()
        n = ctypes.c_int64()
        hits = np.empty(cap, dtype=HIT_DTYPE)

        def run(buf, room):
            if pos_range is None:
                return L.pm_search_batch(self._h, dataset._h, len(patterns), arr, _b(kopt), ctypes.c_void_p(buf.ctypes.data), room, offsets)
            return L.pm_search_batch_fills(self._h, dataset._h, len(patterns), arr, _b(kopt), int(pos_range[0]), int(pos_range[1]),
                                           ctypes.c_void_p(buf.ctypes.data), room, offsets)
        rc = run(hits, cap)
        from_keep = False
        if rc == PM_ERR_OVERFLOW:
            total = int(offsets[len(patterns)])
            if total > 0 and self.stats()["hits"] == total:      # fused batch: the list is still on the device
                keep = getattr(self, "_keep", None)                # page-locked result buffer, reused across calls
                if keep is None or keep.array.size < total:
                    _, keep = pinned_empty(total + total // 8, HIT_DTYPE)
                    self._keep = keep
                hits = keep.array
                from_keep = True
                rc = L.pm_last_hits(self._h, ctypes.c_void_p(hits.ctypes.data), total, ctypes.byref(n))
            else:                                                 # per-pattern fallback: run again with room
                while rc == PM_ERR_OVERFLOW:
                    cap *= 4
                    hits = np.empty(cap, dtype=HIT_DTYPE)
                    rc = run(hits, cap)
        _check(rc)
        off = np.frombuffer(offsets, dtype=np.int64).copy()
        if copy and from_keep:
            return hits[: off[-1]].copy(), off          # the pinned staging buffer is reused by the next call: hand out a copy
        return hits[: off[-1]], off

    @_locked
    def search_batch_compact(self, dataset, patterns, kopt="0ids", pos_range=None, out=None):
        """pm_search_batch_fills_compact: batches of exact motifs with large results.  -> (begins, offsets, base, motif_len):
        hit i of motif p (offsets[p] <= i < offsets[p+1]) is [base + begins[i], base + begins[i] + motif_len[p]); 4 bytes
        per hit cross PCIe instead of 16.  `begins` is a view of a page-locked buffer the engine object reuses: valid
        until the next call.  expand_compact() gives the HIT_DTYPE rows.  NativeError(PM_ERR_UNSUPPORTED) for batches the
        fused path does not serve (errors, anchors, repeats, proteomes): use search_batch.
        out: a CompactBuffer that receives the begins instead of the engine's own buffer (it grows when too small), for
        callers that keep several results alive (distributed.PipelinedBatch)."""
        L = load()
        npat = len(patterns)
        arr = _pattern_array(self, patterns)
        offsets = (ctypes.c_int64 * (npat + 1))()
        base = ctypes.c_int64()
        mlen = np.zeros(npat, dtype=np.uint16)
        beg, end = (0, -1) if pos_range is None else (int(pos_range[0]), int(pos_range[1]))
        if out is None:
            out = getattr(self, "_keep_c", None)
            if out is None:
                out = self._keep_c = CompactBuffer()
        while True:
            if out.pinned is None:
                out.pinned = pinned_empty(1 << 20, np.dtype(np.uint32))[1]
            a = out.pinned.array
            rc = L.pm_search_batch_fills_compact(self._h, dataset._h, npat, arr, _b(kopt), beg, end, ctypes.c_void_p(a.ctypes.data),
                                                 a.size, offsets, ctypes.byref(base), ctypes.c_void_p(mlen.ctypes.data))
            if rc != PM_ERR_OVERFLOW:
                break
            total = int(offsets[npat])
            out.pinned = pinned_empty(total + total // 8 + 4096, np.dtype(np.uint32))[1]
        _check(rc)
        off = np.frombuffer(offsets, dtype=np.int64).copy()
        return out.pinned.array[: off[-1]], off, int(base.value), mlen

    @_locked
    def candidates(self, dataset, pattern, kopt, pos_beg, pos_end, cap=1 << 16):
        L = load()
        n = ctypes.c_int64()
        while True:
            c = np.empty(cap, dtype=CAND_DTYPE)
            rc = L.pm_candidates(self._h, dataset._h, _b(pattern), _b(kopt), pos_beg, pos_end, ctypes.c_void_p(c.ctypes.data), cap, ctypes.byref(n))
            if rc == PM_ERR_OVERFLOW:
                cap = int(n.value)
                continue
            _check(rc)
            return c[: n.value]

    def candidates_device(self, dataset, pattern, kopt, pos_beg, pos_end, dev_ptr, cap):
        """candidates written to device memory at dev_ptr (cap records of 32 bytes); -> count (may exceed cap: overflow)"""
        n = ctypes.c_int64()
        rc = load().pm_candidates_device(self._h, dataset._h, _b(pattern), _b(kopt), pos_beg, pos_end,
                                         ctypes.c_void_p(dev_ptr), cap, ctypes.byref(n))
        if rc == PM_ERR_OVERFLOW:
            return -int(n.value)
        _check(rc)
        return int(n.value)

    @_locked
    def resolve_device(self, dataset, pattern, kopt, dev_ptr, ncands, cap=1 << 16):
        L = load()
        n = ctypes.c_int64()
        hits = np.empty(cap, dtype=HIT_DTYPE)
        rc = L.pm_resolve_device(self._h, dataset._h, _b(pattern), _b(kopt), ctypes.c_void_p(dev_ptr), ncands,
                                 ctypes.c_void_p(hits.ctypes.data), cap, ctypes.byref(n))
        if rc == PM_ERR_OVERFLOW:
            hits = np.empty(int(n.value), dtype=HIT_DTYPE)
            rc = L.pm_last_hits(self._h, ctypes.c_void_p(hits.ctypes.data), len(hits), ctypes.byref(n))
        _check(rc)
        return hits[: n.value]

    def resolve(self, dataset, pattern, kopt, cands, cap=None):
        L = load()
        cands = np.ascontiguousarray(cands, dtype=CAND_DTYPE)
        cap = max(len(cands), 1) if cap is None else cap
        hits = np.empty(cap, dtype=HIT_DTYPE)
        n = ctypes.c_int64()
        _check(L.pm_resolve(self._h, dataset._h, _b(pattern), _b(kopt), ctypes.c_void_p(cands.ctypes.data), len(cands),
                            ctypes.c_void_p(hits.ctypes.data), cap, ctypes.byref(n)))
        return hits[: n.value]

    def stats(self):
        s = PmStats()
        _check(load().pm_get_stats(self._h, ctypes.byref(s)))
        return {f: getattr(s, f) for f, _ in PmStats._fields_}


def jit_wait():
    """pm_jit_wait: blocks until the background compilations of specialised scan kernels (Engine.set_jit("auto")) are done."""
    _check(load().pm_jit_wait())


def expand_compact(begins, offsets, base, motif_len):
    """(begins, offsets, base, motif_len) of Engine.search_batch_compact -> the HIT_DTYPE rows of Engine.search_batch."""
    out = np.empty(len(begins), dtype=HIT_DTYPE)
    out["beg"] = begins.astype(np.int64) + base
    out["end"] = out["beg"] + np.repeat(motif_len.astype(np.int64), np.diff(offsets))
    return out
