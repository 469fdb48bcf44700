"""CPU: host logic -- pattern conversion, planner, C ABI surface, request mirror."""
import ctypes
import random
import re

import numpy as np
import pytest

import oracle_lib as O
import patmatchdocker_b200 as pm
from patmatchdocker_b200 import patmatch as host
from patmatchdocker_b200 import _native
from synth import DNA, PEP, random_pattern


def test_pattern_conversion_matches_perl(pattern_golden):
    bad = [(c["cls"], c["pattern"]) for c in pattern_golden if pm.convert(c["pattern"], c["cls"]) != c["nrgrep"]]
    assert not bad, bad[:5]


def test_cabi_exports_every_declared_symbol():
    import os
    header = open(os.path.join(os.path.dirname(_native.lib_path()), "..", "..", "include", "patmatch_b200.h")).read()
    names = set(re.findall(r"\b(pm_[a-z_]+)\s*\(", header))
    assert len(names) >= 14
    lib = ctypes.CDLL(_native.lib_path())
    for n in names:
        assert hasattr(lib, n), n
    lib.pm_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.pm_version()


def test_ctypes_binding_matches_the_header_prototypes():
    # every prototype of include/patmatch_b200.h against the argtypes the ctypes binding declares: the same number of
    # parameters, pointers where the header has pointers, 64-bit integers where it has int64_t
    import os
    header = open(os.path.join(os.path.dirname(_native.lib_path()), "..", "..", "include", "patmatch_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = re.findall(r"\b(?:int|int64_t|void|const char \*|void \*)\s*\**\s*(pm_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", header)
    assert len(protos) >= 30, len(protos)
    L = _native.load()
    checked = 0
    for name, params in protos:
        fn = getattr(L, name)
        if fn.argtypes is None:
            continue                                        # functions the Python host never calls
        plist = [] if params.strip() in ("", "void") else [x.strip() for x in params.split(",")]
        assert len(plist) == len(fn.argtypes), (name, plist, fn.argtypes)
        for text, at in zip(plist, fn.argtypes):
            is_ptr = "*" in text
            ptr_like = at in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(at, "contents") or getattr(at, "_type_", None) is not None and issubclass(at, ctypes._Pointer)
            assert is_ptr == bool(ptr_like), (name, text, at)
            if not is_ptr and "int64_t" in text:
                assert ctypes.sizeof(at) == 8, (name, text, at)
            if not is_ptr and re.match(r"(const\s+)?int\b", text):
                assert ctypes.sizeof(at) == 4, (name, text, at)
        checked += 1
    assert checked >= 25, checked


def test_plan_is_bit_exact_with_oracle():
    rng = random.Random(11)
    for it in range(1500):
        alpha = rng.choice([DNA, PEP])
        m = rng.randint(3, 26) if rng.random() < 0.8 else rng.randint(27, 64)
        k = min(rng.choice([1, 1, 2, 2, 3, 4]), m - 1)
        pat, _ = random_pattern(rng, alpha, m)
        kopt = "%d%s" % (k, rng.choice(["ids", "s", "id", "d"]))
        mine = pm.plan(pat, kopt)
        _, pl = O.plan(pat, kopt)
        assert mine["type"] == O.TYPE_NAMES[pl.type], (pat, kopt)
        assert mine["L"] == pl.L and mine["V"] == list(pl.V)[: pl.npieces], (pat, kopt)
        assert mine["split_cost"] == pl.split_cost and mine["fb_cost"] == pl.fb_cost, (pat, kopt)


def test_deployed_compat_plan_is_bit_exact_with_oracle():
    # pm_set_compat_deployed_glibc: the piece choice of the DEPLOYED binary (default glibc allocator).  Product
    # (plan.cpp) and oracle model the stale scratch cells independently; they must agree, and the mode must change
    # some plans of patterns with >= 11 positions and none of the shorter exact-arithmetic cases it does not touch.
    rng = random.Random(12)
    changed = 0
    try:
        for it in range(1200):
            m = rng.randint(5, 30)
            k = min(rng.choice([1, 1, 2, 2, 3]), m - 1)
            pat, _ = random_pattern(rng, DNA, m, neg_pct=0.0)
            kopt = "%d%s" % (k, rng.choice(["ids", "s", "id"]))
            pm.set_compat_deployed_glibc(False); O.set_compat(False)
            zero = pm.plan(pat, kopt)
            pm.set_compat_deployed_glibc(True); O.set_compat(True)
            mine = pm.plan(pat, kopt)
            _, pl = O.plan(pat, kopt)
            assert mine["type"] == O.TYPE_NAMES[pl.type], (pat, kopt)
            assert mine["L"] == pl.L and mine["V"] == list(pl.V)[: pl.npieces], (pat, kopt)
            assert mine["split_cost"] == pl.split_cost, (pat, kopt)
            changed += 1 if (mine["type"], mine["L"], mine["V"]) != (zero["type"], zero["L"], zero["V"]) else 0
        assert changed > 20, changed
    finally:
        pm.set_compat_deployed_glibc(False); O.set_compat(False)


def test_unsupported_patterns_fail_loudly():
    for pat in ("(GA(TA)*AG)", "(GA(TA)?AG)", "(GAT|AAG)", "(G(AT)+AAG)"):
        with pytest.raises(pm.NativeError) as ei:
            pm.plan(pat, "0ids")
        assert ei.value.code == -3
    with pytest.raises(pm.NativeError):
        pm.plan("(GATAAG", "0ids")
    with pytest.raises(pm.NativeError):
        pm.plan("(GATAAG)", "ids")


def test_engine_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pm.NativeError):
        pm.Engine(0)


class _OracleEngine:
    """Stand-in for the GPU engine so that the request mirror can be checked on CPU."""

    def load_dataset(self, raw):
        class DS:
            pass
        d = DS()
        d.raw = bytes(raw)
        return d

    def search(self, ds, pattern, kopt):
        hits = O.search(pattern, ds.raw, kopt)
        return np.array(hits, dtype=_native.HIT_DTYPE)

    def search_request(self, ds, patterns, kopt):
        return [self.search(ds, p, kopt) for p in patterns]


def _locus(text):
    out = {}
    for line in text.splitlines():
        p = line.split("\t")
        out[p[0]] = (p[1], p[2], p[3] if len(p) > 3 else "")
    return out


def check_requests(service_engine, request_golden):
    svc = host.PatMatch(engine=service_engine)
    ds = request_golden["datasets"]
    svc.add_dataset("orf_dna.seq", ds["orf_dna.seq"].encode(), locus=_locus(ds["locus.txt"]))
    svc.add_dataset("orf_pep.seq", ds["orf_pep.seq"].encode(), locus=_locus(ds["locus.txt"]))
    for r in request_golden["requests"]:
        req = dict(r["request"])
        res = svc.run_patmatch(req.pop("pattern"), **req)
        if r["error"] and not r["hits"]:
            assert res.get("error") == r["error"], (r["request"], res)
            continue
        assert res["hits"] == r["hits"], r["request"]
        assert (res["uniqueHits"], res["totalHits"]) == (r["uniqueHits"], r["totalHits"]), r["request"]


def test_request_mirror_matches_reference_python(request_golden):
    check_requests(_OracleEngine(), request_golden)


def test_record_index_matches_reference_layout():
    data = b">a desc\nACGT\n>b,\nGG\n>c\n"
    offs, names = host.get_record_offset(data)
    assert offs == [0, 8, 13, 17, 20, 23]
    assert names == {0: ">a", 8: "a", 13: ">b,", 17: "b,", 20: ">c", 23: "c"}
    assert host.get_name_offset(10, offs) == 8


def test_extended_plan_matches_oracle():
    # EXTENDED patterns (positions with '?'): the product's planner (find_best_extended in csrc/plan.cpp) must choose the
    # same verification type and anchor as the oracle's restatement of extendedFindBest, which is pinned on the binary
    import random
    import oracle_lib as O
    import patmatchdocker_b200 as pm
    rng = random.Random(123)
    seen = set()
    for _ in range(400):
        alpha = rng.choice(["ACGT", "ACDEFGHIKLMNPQRSTVWY"])
        m = rng.randint(3, 24)
        pat, nops = "(", 0
        for j in range(m):
            r = rng.random()
            pat += "." if r < 0.12 else "[" + "".join(rng.sample(alpha, 2)) + "]" if r < 0.27 else rng.choice(alpha)
            if 0 < j < m - 1 and rng.random() < 0.3:
                pat += rng.choice("????*+")
                nops += 1
        pat += ")"
        if not nops:
            continue
        _, xpl = O.plan_ext(pat)
        p = pm.plan(pat, "0ids")
        assert p["type"] == {2: "EXT_BEG", 3: "EXT_END"}[xpl.type], pat
        assert p["V"][0] == xpl.anchor, pat
        seen.add(p["type"])
    assert seen == {"EXT_BEG", "EXT_END"}
    assert pm.plan("(A?CGT)", "0ids")["type"] == "SIMPLE" and pm.plan("(ACGT?G*)", "0ids")["m"] == 3      # parser rewrites
    assert pm.plan("(A+CG?T+)", "0ids")["type"] in ("EXT_BEG", "EXT_END")
    with pytest.raises(pm.NativeError):
        pm.plan("(A?CGT)", "1ids")
    assert pm.plan("(A?C?G)", "0ids")["type"] == "EXT_END"             # C?G: forward scan, leading optional position
    for bad in ("(A(CG)?T)",):
        with pytest.raises(pm.NativeError):
            pm.plan(bad, "0ids")
    with pytest.raises(pm.NativeError):
        pm.plan("(AC?GT)", "1ids")
