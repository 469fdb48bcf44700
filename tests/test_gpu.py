"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the
vectors produced by the reference.  Bar: bit-exact hit lists."""
import random

import numpy as np
import pytest

import oracle_lib as O
import patmatchdocker_b200 as pm
from patmatchdocker_b200 import patmatch as host
from synth import DNA, PEP, genome, random_pattern, random_text
from test_host import check_requests

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["bytes", "packed"])
def scan_mode(request, engine):
    """run the parity tests through both scan kernels (byte Shift-And and 2-bit packed bit-sliced)"""
    engine.set_scan_mode(request.param)
    yield request.param
    engine.set_scan_mode("auto")


def extended_pattern(rng, alpha, m):
    """random EXTENDED pattern (operators on inner positions) and, per position, the accepted letters and the operator"""
    pat, members, ops = "(", [], []
    for j in range(m):
        r = rng.random()
        if r < 0.12:
            pat += "."; members.append(list(alpha))
        elif r < 0.27:
            ch = rng.sample(alpha, 2); pat += "[" + "".join(ch) + "]"; members.append(ch)
        else:
            c = rng.choice(alpha); pat += c; members.append([c])
        op = rng.choice("????*+") if 0 < j < m - 1 and rng.random() < 0.3 else ""
        pat += op
        ops.append(op)
    if not any(ops):
        return extended_pattern(rng, alpha, m)
    return pat + ")", members, ops


def extended_pattern_with_ends(rng, alpha, m):
    """like extended_pattern, operators on the first / last positions too (the reference's parser rewrites those)"""
    toks, members, ops = [], [], []
    for j in range(m):
        r = rng.random()
        if r < 0.12:
            toks.append("."); members.append(list(alpha))
        elif r < 0.27:
            ch = rng.sample(alpha, 2); toks.append("[" + "".join(ch) + "]"); members.append(ch)
        else:
            c = rng.choice(alpha); toks.append(c); members.append([c])
        ops.append(rng.choice("????*+") if rng.random() < (0.5 if j in (0, 1, m - 1) else 0.25) else "")
    if not any(ops):
        ops[0] = "?"
    return "(" + "".join(a + b for a, b in zip(toks, ops)) + ")", members, ops


def extended_text(rng, alpha, members, ops, nrec, lo, hi):
    lines = []
    for r in range(nrec):
        lines.append(">x%d" % r)
        t = ""
        target = rng.randint(lo, hi)
        while len(t) < target:
            if rng.random() < 0.3:
                for cls, op in zip(members, ops):
                    reps = 1 if op == "" else rng.randint(0, 1) if op == "?" else rng.randint(0, 3) if op == "*" else rng.randint(1, 3)
                    t += "".join(rng.choice(cls) for _ in range(reps))
            else:
                t += "".join(rng.choice(alpha) for _ in range(rng.randint(1, 10)))
        lines.append(t)
    return ("\n".join(lines) + "\n").encode()


def gpu_hits(engine, text, pattern, kopt):
    raw = text.encode("latin-1") if isinstance(text, str) else text
    ds = engine.load_dataset(raw)
    try:
        return [(int(b), int(e)) for b, e in engine.search(ds, pattern, kopt)]
    finally:
        ds.close()


def test_reference_golden_vectors(engine, search_golden, scan_mode):
    bad = []
    try:
        for c in search_golden:
            engine.set_buffer_size(c["bufsize"])
            try:
                got = [list(h) for h in gpu_hits(engine, c["text"], c["pattern"], c["kopt"])]
            except pm.NativeError as e:
                raise AssertionError((c["pattern"], str(e)))
            if got != c["hits"]:
                bad.append((c["pattern"], c["kopt"], c["bufsize"], got[:4], c["hits"][:4]))
    finally:
        engine.set_buffer_size(1600000)
    assert not bad, bad[:5]


def test_buffer_fills_against_oracle(engine, scan_mode):
    # the reference cuts the file into -b sized fills; exercise many fills per file
    rng = random.Random(2718)
    try:
        for it in range(150):
            m = rng.randint(3, 14)
            k = min(rng.choice([0, 0, 1, 2, 3]), m - 1)
            kopt = "%d%s" % (k, rng.choice(["ids", "s", "id"]))
            pat, members = random_pattern(rng, DNA, m, dot_pct=0.2 if k == 0 else 0.05, neg_pct=0.3)
            text = random_text(rng, members, DNA, k, nrec=rng.randint(1, 6), lo=20, hi=1500)
            bs = rng.choice([16, 33, 64, 100, 257, 1000, 4096])
            engine.set_buffer_size(bs)
            assert gpu_hits(engine, text, pat, kopt) == O.search(pat, text, kopt, bufsize=bs), (pat, kopt, bs)
    finally:
        engine.set_buffer_size(1600000)


@pytest.mark.parametrize("alpha", [DNA, PEP])
def test_random_cases_against_oracle(engine, alpha, scan_mode):
    rng = random.Random(1234 if alpha == DNA else 4321)
    types = set()
    for it in range(250):
        m = rng.randint(3, 24) if rng.random() < 0.85 else rng.randint(25, 60)
        k = min(rng.choice([0, 1, 1, 2, 2, 3]), m - 1)
        kopt = "%d%s" % (k, rng.choice(["ids", "ids", "s", "id", "is", "ds", "i", "d"]))
        pat, members = random_pattern(rng, alpha, m)
        text = random_text(rng, members, alpha, k, nrec=rng.randint(1, 4), lo=80, hi=3000)
        want = O.search(pat, text, kopt)
        got = gpu_hits(engine, text, pat, kopt)
        assert got == want, (pat, kopt)
        types.add(pm.plan(pat, kopt)["type"])
    assert types == {"SIMPLE", "SPLIT", "BWD", "FWD"}


def test_multi_tile_texts(engine, scan_mode):
    # texts several scan tiles long, planted hits straddling tile and thread-run borders
    rng = random.Random(5)
    for it in range(12):
        alpha = DNA
        m = rng.randint(6, 40)
        k = min(rng.choice([0, 1, 2, 3]), m - 1)
        kopt = "%d%s" % (k, rng.choice(["ids", "s", "id"]))
        pat, members = random_pattern(rng, alpha, m, cls_pct=0.1, dot_pct=0.03)
        text = random_text(rng, members, alpha, k, nrec=3, lo=40000, hi=120000, plant=0.02)
        assert gpu_hits(engine, text, pat, kopt) == O.search(pat, text, kopt), (pat, kopt)


def test_edge_cases(engine, scan_mode):
    cases = [
        ("(GATAAG)", "0ids", ""),
        ("(GATAAG)", "1ids", ""),
        ("(GATAAG)", "0ids", "GAT"),
        ("(GATAAG)", "2ids", "GAT"),
        ("(GATAAG)", "0ids", ">only a header"),
        ("(GATAAG)", "0ids", "GATAAG"),                       # hit ends at EOF, no newline
        ("(GATAAG)", "1ids", ">s\nGATAAG"),
        ("(GATAAG)", "1ids", ">s\n\n\n>t\nGATAG\n\n"),
        ("(AAA)", "0ids", ">s\n" + "A" * 1000 + "\n"),        # dense overlapping occurrences
        ("(AAAA)", "1ids", ">s\n" + "A" * 777 + "\n"),
        ("(A.A)", "0ids", ">s\nA\nA\nAAA\nA\n"),              # '.' over record separators
        ("(...)", "0ids", ">s\nACGTAC\n"),
        ("([^A][^C])", "0ids", ">s\nACGT\nNN\n"),
        ("(gataag)", "0ids", ">s\nGATAAGgataagGaTaAg\n"),     # -i
        ("(GATAAG)", "3ids", ">s\nGTAGGATAGATAAGGGATTAAGTT\n"),  # short pattern, k = 3 -> BWD/FWD plans
        ("(ACGTAC)", "2s", ">s\n" + "ACGTAC" * 50 + "\n"),
    ]
    for pat, kopt, text in cases:
        assert gpu_hits(engine, text, pat, kopt) == O.search(pat, text, kopt), (pat, kopt, text[:30])


def test_small_hit_buffer_is_regrown(engine):
    text = (">s\n" + "GATAAG" * 5000 + "\n").encode()
    ds = engine.load_dataset(text)
    hits = engine.search(ds, "(GATAAG)", "0ids", cap=16)
    assert len(hits) == 5000 and engine.count(ds, "(GATAAG)", "0ids") == 5000
    ds.close()


def test_scan_kernel_selection_and_equivalence(engine):
    # auto mode: packed bit-sliced scan for DNA-like datasets, byte Shift-And for proteomes; both kernels
    # give the same hit list on a genome with lower case, N runs, IUPAC letters and a non-ACGT-accepting class
    rng = random.Random(12)
    g = bytearray(genome(21, 6, 3_000_000))
    for _ in range(200):
        p = rng.randrange(100, len(g) - 3000)
        if b"\n" in g[p - 5:p + 2100] or b">" in g[p - 5:p + 2100]:
            continue
        kind = rng.randint(0, 2)
        ln = rng.randint(1, 2000)
        if kind == 0:
            g[p:p + ln] = b"N" * ln
        elif kind == 1:
            g[p:p + ln] = bytes(g[p:p + ln]).lower()
        else:
            g[p:p + 6] = b"RYKMSW"
    g = bytes(g)
    ds = engine.load_dataset(g)
    for pat, kopt in (("(GAT[AG]AG)", "0ids"), ("(GA.AAG[^C])", "0ids"), ("(NNGATAAG)", "0ids"), ("(GATAAGCC[AT]TT)", "1ids"),
                      ("(TGA[GC]TCA...[AG][CT]GATAAG)", "2ids"), ("(GATAAG[^A]N)", "1s")):
        engine.set_scan_mode("auto")
        a = engine.search(ds, pat, kopt)
        packed = engine.stats()["packed"]
        engine.set_scan_mode("bytes")
        b = engine.search(ds, pat, kopt)
        engine.set_scan_mode("auto")
        assert np.array_equal(a, b), (pat, kopt)
        if pm.plan(pat, kopt)["type"] in ("SIMPLE", "SPLIT"):
            assert packed == 1
        assert [(int(x), int(y)) for x, y in a] == O.search(pat, g, kopt), (pat, kopt)
    ds.close()
    prot = genome(22, 500, 200_000, alphabet=PEP.encode(), name="YORF")
    ds = engine.load_dataset(prot)
    engine.search(ds, "(CAAC[ILVM]QQH)", "1s")
    assert engine.stats()["packed"] == 2                       # proteomes: Shift-And over the 5-bit residue codes
    ds.close()


@pytest.mark.parametrize("filter_mode", [1, 4])
def test_qgram_prefilter_keeps_every_hit(engine, filter_mode):
    # filter_mode 1: k_scan_apx (pieces built from q-gram chunks, Landau-Vishkin check per pattern start);
    # filter_mode 4: the first-generation k_scan_split (q-gram count + Myers filter).
    # low-selectivity pieces (many wildcards / wide classes) switch the bit-sliced q-gram pre-filter on; it
    # may only drop candidates whose verification fails: hit lists equal the oracle's and the unfiltered scan's
    rng = random.Random(77)
    engine.set_scan_mode("packed")
    used = 0
    try:
        for it in range(160):
            k = rng.randint(1, 3)
            m = rng.randint(max(2 * k + 2, 6), 40)
            pat, members = random_pattern(rng, DNA, m, cls_pct=0.3, dot_pct=0.3)
            kopt = "%d%s" % (k, rng.choice(["ids", "ids", "s", "id", "is", "ds", "i", "d"]))
            try:
                plan = pm.plan(pat, kopt)
            except pm.NativeError:
                continue
            if plan["type"] != "SPLIT":
                continue
            text = random_text(rng, members, DNA, k, nrec=rng.randint(1, 3), lo=200, hi=30000 if it % 8 == 0 else 3000, plant=0.1)
            if it % 3 == 0:                                   # a (possibly truncated) occurrence at the very start of the file
                s0 = "".join(rng.choice(c) for c in members)
                text = s0[rng.randint(0, k):] + text[text.index("\n") + 1:]
            raw = text.encode("latin-1")
            ds = engine.load_dataset(raw)
            engine.set_fused_filter(filter_mode)
            a = engine.search(ds, pat, kopt)
            st = engine.stats()
            used += 1 if st["qgram_chunks"] > 0 else 0
            engine.set_fused_filter(0)
            b = engine.search(ds, pat, kopt)
            ds.close()
            assert np.array_equal(a, b), (pat, kopt)
            assert [(int(x), int(y)) for x, y in a] == O.search(pat, raw, kopt), (pat, kopt)
        assert used >= 20, used
    finally:
        engine.set_fused_filter(1)
        engine.set_scan_mode("auto")


def test_line_anchors_against_oracle(engine, scan_mode):
    # '<' / '>' of the PatMatch syntax = leading '^' / trailing '$': the match must start / end at a record boundary,
    # or exactly at the scan start (the end of the previous hit), which makes '^' depend on the chain of hits
    rng = random.Random(31)
    n_hits = 0
    for it in range(220):
        alpha = rng.choice([DNA, PEP])
        k = rng.choice([0, 0, 1, 1, 2, 3])
        m = rng.randint(max(2, 2 * k + 1), 14 if it % 5 else 40)
        pat, members = random_pattern(rng, alpha, m, cls_pct=0.2, dot_pct=0.1)
        mode = rng.randint(1, 3)
        pat = ("^" if mode & 1 else "") + pat + ("$" if mode & 2 else "")
        kopt = "%d%s" % (k, rng.choice(["ids", "s", "id", "is", "ds", "i"]) if k else "ids")
        lines = []
        for r in range(rng.randint(2, 9)):
            lines.append(">s%d" % r)
            s0 = "".join(rng.choice(c) for c in members)
            body = "".join(rng.choice(alpha) for _ in range(rng.randint(0, 12)))
            t = [s0 + body, body + s0, s0, s0 + s0 + body + s0, body + s0 + body, s0 * 4][rng.randint(0, 5)]
            if k and rng.random() < 0.5 and len(t) > 2:
                q = rng.randrange(len(t))
                t = t[:q] + rng.choice(alpha) + t[q + 1:]
            if k and rng.random() < 0.3 and len(t) > 3:
                q = rng.randrange(len(t))
                t = t[:q] + t[q + 1:]
            lines.append(t)
        text = ("\n".join(lines) + ("\n" if rng.random() < 0.8 else "")).encode()
        for bufsize in (1600000, rng.choice([32, 50, 100])):
            engine.set_buffer_size(bufsize)
            try:
                got = gpu_hits(engine, text, pat, kopt)
            finally:
                engine.set_buffer_size(1600000)
            want = O.search(pat, text, kopt, bufsize=bufsize)
            assert got == want, (pat, kopt, bufsize, text)
            n_hits += len(want)
    assert n_hits > 500


def test_extended_patterns_against_oracle(engine, scan_mode):
    # nrgrep's EXTENDED engine (k = 0): PatMatch's X{m,n} repeats arrive as runs of optional positions.  Anchors come
    # from an exact scan of the plain positions around the chosen sub-pattern, verification is the reference's walk
    # (shortest extension, with its initial-state quirk), the chain stage is the common one.
    rng = random.Random(404)
    n_hits = 0
    types = set()
    for it in range(200):
        alpha = rng.choice([DNA, DNA, PEP])
        m = rng.randint(3, 20 if it % 4 else 40)
        pat, members, ops = "(", [], []
        for j in range(m):
            r = rng.random()
            if r < 0.12:
                pat += "."; members.append(list(alpha))
            elif r < 0.27:
                ch = rng.sample(alpha, 2); pat += "[" + "".join(ch) + "]"; members.append(ch)
            else:
                c = rng.choice(alpha); pat += c; members.append([c])
            op = rng.choice("????*+") if 0 < j < m - 1 and rng.random() < 0.3 else ""
            pat += op
            ops.append(op)
        pat += ")"
        if not any(ops):
            continue
        if it % 6 == 0:
            pat = "^" + pat
        if it % 10 == 0:
            pat = pat + "$"
        lines = []
        for r in range(rng.randint(1, 5)):
            lines.append(">x%d" % r)
            t = ""
            target = rng.randint(30, 3000 if it % 5 else 40000)
            while len(t) < target:
                if rng.random() < 0.3:
                    for cls, op in zip(members, ops):
                        reps = 1 if op == "" else rng.randint(0, 1) if op == "?" else rng.randint(0, 3) if op == "*" else rng.randint(1, 3)
                        t += "".join(rng.choice(cls) for _ in range(reps))
                else:
                    t += "".join(rng.choice(alpha) for _ in range(rng.randint(1, 10)))
            if rng.random() < 0.3:
                t = "".join(ch.lower() if rng.random() < 0.3 else ch for ch in t)
            lines.append(t)
        text = ("\n".join(lines) + "\n").encode()
        types.add(pm.plan(pat, "0ids")["type"])
        for bufsize in (1600000, rng.choice([64, 300, 2000])):
            engine.set_buffer_size(bufsize)
            try:
                got = gpu_hits(engine, text, pat, "0ids")
            finally:
                engine.set_buffer_size(1600000)
            want = O.search(pat, text, "0ids", bufsize=bufsize)
            assert got == want, (pat, bufsize, got[:4], want[:4])
            n_hits += len(want)
    assert n_hits > 1000 and types == {"EXT_BEG", "EXT_END"}
    # operators at the pattern ends: parser rewrites, and the forward scan's blind spot at the first byte of a range
    n_hits = 0
    for it in range(200):
        alpha = rng.choice([DNA, DNA, PEP])
        pat, members, ops = extended_pattern_with_ends(rng, alpha, rng.randint(2, 9))
        try:
            O.plan_ext(pat)
        except ValueError:
            continue                                       # nothing left after the rewrites
        text = extended_text(rng, alpha, members, ops, rng.randint(1, 6), 5, 400)
        for bufsize in (1600000, rng.choice([16, 64, 300])):
            engine.set_buffer_size(bufsize)
            try:
                got = gpu_hits(engine, text, pat, "0ids")
            except pm.NativeError as e:
                assert "single-position pattern" in str(e), (pat, str(e))      # e.g. (Y?.) -> '.', refused
                continue
            finally:
                engine.set_buffer_size(1600000)
            want = O.search(pat, text, "0ids", bufsize=bufsize)
            assert got == want, (pat, bufsize, text, got[:6], want[:6])
            n_hits += len(want)
    assert n_hits > 1000
    assert gpu_hits(engine, b"AATAAC\n", "(T?[TG]*A)", "0ids") == [(1, 2), (3, 4)]      # A0 and A4 open a scan range
    # the reference's quirk: zero occurrences of a run of two or more optional positions next to the anchor do not match
    assert gpu_hits(engine, b">q\nCCGATAAGTCCAA\nCCGATCAAGTCCAA\n", "(GAT.?.?.?AAGTCC)", "0ids") == [(19, 29)]


def test_chain_stage_on_tandem_repeats(engine, scan_mode):
    # back-to-back (mutated) occurrences: hits end where the next candidates begin to look, so the chain stage's cluster
    # boundaries are exercised hard -- a boundary is only sound when no LATER candidate of the cluster can reach left of
    # an earlier hit's end
    from synth import planted
    rng = random.Random(2024)
    n_hits = 0
    for it in range(150):
        alpha = rng.choice([DNA, DNA, PEP])
        k = rng.choice([1, 1, 2, 3])
        m = rng.randint(max(4, 2 * k + 2), 16)
        pat, members = random_pattern(rng, alpha, m, cls_pct=0.1, dot_pct=0.05)
        kopt = "%d%s" % (k, rng.choice(["ids", "ids", "s", "id", "is"]))
        lines = []
        for r in range(rng.randint(1, 3)):
            lines.append(">t%d" % r)
            t = ""
            for _ in range(rng.randint(3, 60)):
                t += planted(rng, members, alpha, k)
                if rng.random() < 0.3:
                    t += "".join(rng.choice(alpha) for _ in range(rng.randint(1, m + k + 2)))
            lines.append(t)
        text = ("\n".join(lines) + "\n").encode()
        got = gpu_hits(engine, text, pat, kopt)
        want = O.search(pat, text, kopt)
        assert got == want, (pat, kopt, text)
        n_hits += len(want)
    assert n_hits > 2000


def test_exact_patterns_longer_than_64_positions(engine, scan_mode):
    # oligo-sized exact patterns: the scan looks for the first 64 positions, k_verify compares the rest
    rng = random.Random(65)
    n_hits = 0
    for it in range(40):
        alpha = rng.choice([DNA, DNA, PEP])
        m = rng.randint(65, 255 if it % 3 else 80)
        pat, members = random_pattern(rng, alpha, m, cls_pct=0.1, dot_pct=0.05)
        if it % 5 == 0:
            pat = "^" + pat
        lines = []
        for r in range(rng.randint(1, 4)):
            lines.append(">o%d" % r)
            t = ""
            for _ in range(rng.randint(1, 12)):
                s0 = [rng.choice(c) for c in members]
                if rng.random() < 0.4:                         # spoil it somewhere (often beyond the first 64 positions)
                    s0[rng.randrange(len(s0))] = rng.choice(alpha)
                t += "".join(s0) + "".join(rng.choice(alpha) for _ in range(rng.randint(0, 30)))
            lines.append(t)
        text = ("\n".join(lines) + "\n").encode()
        for bufsize in (1600000, rng.choice([300, 1000])):
            engine.set_buffer_size(bufsize)
            try:
                got = gpu_hits(engine, text, pat, "0ids")
            finally:
                engine.set_buffer_size(1600000)
            want = O.search(pat, text, "0ids", bufsize=bufsize)
            assert got == want, (pat, bufsize)
            n_hits += len(want)
    assert n_hits > 100


def test_approximate_patterns_longer_than_64_positions(engine, scan_mode):
    # multi-word verification (nfa_side_mw): SPLIT pieces / BWD sub-patterns still fit 64 positions, the anchored NFA
    # walks parts of up to 255
    from synth import planted
    rng = random.Random(650)
    n_hits, types = 0, set()
    for it in range(40):
        alpha = rng.choice([DNA, DNA, PEP])
        k = rng.choice([1, 1, 2, 3])
        m = rng.randint(65, 200 if it % 3 else 70)
        pat, members = random_pattern(rng, alpha, m, cls_pct=0.1, dot_pct=0.05)
        kopt = "%d%s" % (k, rng.choice(["ids", "ids", "s", "id"]))
        types.add(pm.plan(pat, kopt)["type"])
        lines = []
        for r in range(rng.randint(1, 3)):
            lines.append(">a%d" % r)
            t = ""
            for _ in range(rng.randint(1, 8)):
                t += planted(rng, members, alpha, k) + "".join(rng.choice(alpha) for _ in range(rng.randint(0, 40)))
            lines.append(t)
        text = ("\n".join(lines) + "\n").encode()
        for bufsize in (1600000, rng.choice([700, 3000])):
            engine.set_buffer_size(bufsize)
            try:
                got = gpu_hits(engine, text, pat, kopt)
            finally:
                engine.set_buffer_size(1600000)
            want = O.search(pat, text, kopt, bufsize=bufsize)
            assert got == want, (pat, kopt, bufsize)
            n_hits += len(want)
    assert n_hits > 60 and {"SPLIT", "BWD"} <= types


def test_hit_list_stays_on_device_after_overflow(engine):
    import ctypes
    from patmatchdocker_b200 import _native
    text = (">s\n" + "GATAAGT" * 3000 + "\n").encode()
    ds = engine.load_dataset(text)
    L = _native.load()
    n = ctypes.c_int64()
    small = np.empty(10, dtype=_native.HIT_DTYPE)
    rc = L.pm_search(engine._h, ds._h, b"(GATAAG)", b"0ids", ctypes.c_void_p(small.ctypes.data), 10, ctypes.byref(n))
    assert rc == -5 and n.value == 3000
    launches = engine.stats()["launches"]
    big = np.empty(3000, dtype=_native.HIT_DTYPE)
    assert L.pm_last_hits(engine._h, ctypes.c_void_p(big.ctypes.data), 3000, ctypes.byref(n)) == 0
    assert engine.stats()["launches"] == launches              # fetched, not searched again
    assert big["beg"][0] == 3 and big["end"][-1] == 3 + 7 * 2999 + 6
    ds.close()


def test_sharded_candidates_then_resolve_equals_search(engine, scan_mode):
    rng = random.Random(77)
    for it in range(20):
        m = rng.randint(5, 22)
        k = min(rng.choice([0, 1, 2, 3]), m - 1)
        kopt = "%dids" % k
        pat, members = random_pattern(rng, DNA, m)
        text = random_text(rng, members, DNA, k, nrec=4, lo=2000, hi=9000, plant=0.1).encode()
        if it % 4 == 3:                                    # EXTENDED plans shard by window start like any other
            pat, members, ops = extended_pattern(rng, DNA, rng.randint(4, 14))
            kopt = "0ids"
            text = extended_text(rng, DNA, members, ops, 4, 2000, 9000)
        ds = engine.load_dataset(text)
        whole = engine.search(ds, pat, kopt)
        n = len(text)
        cuts = sorted({0, n + 1, *(rng.randrange(n) for _ in range(3))})
        parts = [engine.candidates(ds, pat, kopt, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
        merged = np.concatenate(parts)
        assert np.all(np.diff(merged["key"]) > 0)
        got = engine.resolve(ds, pat, kopt, merged)
        assert [tuple(x) for x in got.tolist()] == [tuple(x) for x in whole.tolist()], (pat, kopt)
        ds.close()


def test_fill_sharded_search_equals_search(engine, scan_mode):
    # pm_search_fills_device: whole fills are independent (the reference restarts at every fill), so the hit lists of
    # contiguous position ranges, each snapped to the fills that start inside it, concatenate to the full hit list
    import torch
    rng = random.Random(5)
    buf = torch.zeros((1 << 16, 2), dtype=torch.int64, device="cuda")
    for it in range(40):
        k = rng.choice([0, 1, 2])
        m = rng.randint(max(3, 2 * k + 2), 16)
        pat, members = random_pattern(rng, DNA, m)
        if it % 7 == 0:
            pat = "^" + pat
        kopt = "%d%s" % (k, rng.choice(["ids", "s", "id"]))
        text = random_text(rng, members, DNA, k, nrec=rng.randint(2, 6), lo=50, hi=2500).encode("latin-1")
        if it % 3 == 1:                                    # EXTENDED plans: keys are window starts, fills go by the anchor
            pat, members, ops = extended_pattern(rng, DNA, rng.randint(4, 14))
            kopt = "0ids"
            text = extended_text(rng, DNA, members, ops, rng.randint(2, 6), 50, 2500)
        bufsize = rng.choice([1600000, 64, 200, 1000])
        engine.set_buffer_size(bufsize)
        try:
            ds = engine.load_dataset(text)
            full = [(int(b), int(e)) for b, e in engine.search(ds, pat, kopt)]
            world = rng.randint(2, 5)
            cuts = sorted(rng.randint(0, len(text) + 1) for _ in range(world - 1))
            edges = [0] + cuts + [len(text) + 1]
            got = []
            for r in range(world):
                n = engine.search_fills_device(ds, pat, kopt, edges[r], edges[r + 1], buf.data_ptr(), buf.shape[0])
                assert n >= 0
                got += [(int(b), int(e)) for b, e in buf[:n].cpu().numpy()]
            ds.close()
        finally:
            engine.set_buffer_size(1600000)
        assert got == full, (pat, kopt, bufsize, edges)
        assert full == O.search(pat, text, kopt, bufsize=bufsize)


def test_streaming_upload_search_equals_search(engine, scan_mode):
    # pm_search_stream: chunks are packed and their completed fills searched while later chunks are still in flight;
    # hit lists and the dataset left behind must equal the plain path for any chunking and buffer size
    rng = random.Random(9)
    for it in range(24):
        k = rng.choice([0, 1, 2])
        m = rng.randint(max(3, 2 * k + 2), 14)
        pat, members = random_pattern(rng, DNA, m)
        pat2, _ = random_pattern(rng, DNA, m)
        if it % 5 == 0:
            pat = "^" + pat
        kopt = "%d%s" % (k, rng.choice(["ids", "s", "id"]))
        nrec = rng.randint(1, 5)
        text = random_text(rng, members, DNA, k, nrec=nrec, lo=20000, hi=120000, plant=0.02).encode("latin-1")
        if it % 3 == 1:
            pat, members, ops = extended_pattern(rng, DNA, rng.randint(5, 14))
            pat2, _, _ = extended_pattern(rng, DNA, rng.randint(5, 14))
            kopt = "0ids"
            text = extended_text(rng, DNA, members, ops, nrec, 20000, 120000)
        bufsize = rng.choice([1600000, 5000, 40000, 1000])
        engine.set_buffer_size(bufsize)
        try:
            ds, lists = engine.search_stream(text, [pat, pat2], kopt, chunk_bytes=rng.choice([32768, 65536, 1 << 20]))
            for p, got in zip((pat, pat2), lists):
                want = O.search(p, text, kopt, bufsize=bufsize)
                assert [(int(b), int(e)) for b, e in got] == want, (p, kopt, bufsize)
                again = engine.search(ds, p, kopt)                 # the resident dataset is complete
                assert [(int(b), int(e)) for b, e in again] == want, (p, kopt, bufsize)
            ds.close()
        finally:
            engine.set_buffer_size(1600000)
    ds, lists = engine.search_stream(b"", ["(ACGT)"], "1ids")
    assert len(lists[0]) == 0
    ds.close()


def test_batch_equals_single(engine):
    rng = random.Random(3)
    pats = [random_pattern(rng, DNA, rng.randint(5, 12), cls_pct=0.3)[0] for _ in range(40)]
    text = genome(3, 4, 400000)
    ds = engine.load_dataset(text)
    hits, off = engine.search_batch(ds, pats, "0ids")
    for i, p in enumerate(pats):
        one = engine.search(ds, p, "0ids")
        assert np.array_equal(hits[off[i]:off[i + 1]], one), p
    ds.close()


def test_batch_of_iupac_motifs_against_oracle(engine):
    # BASELINE configs[3] in small: many IUPAC motifs (with N and negated classes) over several
    # genomes in one fused multi-pattern launch, small buffer fills included
    rng = random.Random(31)
    iupac = {"R": "[AG]", "Y": "[CT]", "S": "[GC]", "W": "[AT]", "M": "[AC]", "K": "[GT]", "N": ".", "B": "[CGT]", "V": "[^T]"}
    pats = []
    for _ in range(300):
        m = rng.randint(5, 14)
        pats.append("(" + "".join(rng.choice("ACGT") if rng.random() < 0.7 else iupac[rng.choice(list(iupac))] for _ in range(m)) + ")")
    text = genome(17, 12, 600_000)
    ds = engine.load_dataset(text)
    try:
        for bs in (1600000, 70000):
            engine.set_buffer_size(bs)
            hits, off = engine.search_batch(ds, pats, "0ids")
            assert engine.stats()["launches"] < 20                 # one fused scan, not 300 searches
            assert off[-1] == len(hits)
            for i in rng.sample(range(len(pats)), 40):
                want = O.search(pats[i], text, "0ids", bufsize=bs)
                got = [(int(b), int(e)) for b, e in hits[off[i]:off[i + 1]]]
                assert got == want, (pats[i], bs)
    finally:
        engine.set_buffer_size(1600000)
        ds.close()


def test_batch_lookup_kernel_equals_dense_kernel(engine):
    # batches of >= 64 motifs: k_scan_multi_hash (8-mer lookup, then verification of the listed motifs) against the
    # dense multi-pattern kernel and the oracle, on a genome with N runs, lower case and IUPAC letters; motifs shorter
    # than 8 or without an ACGT-only window stay on the dense kernel inside the same call
    rng = random.Random(99)
    iupac = {"R": "[AG]", "Y": "[CT]", "S": "[GC]", "W": "[AT]", "M": "[AC]", "K": "[GT]", "N": ".", "B": "[CGT]", "V": "[^T]", "H": "[ACT]"}
    pats = []
    for i in range(700):
        m = rng.randint(4, 32) if i % 7 == 0 else rng.randint(8, 14)
        pats.append("(" + "".join(rng.choice("ACGT") if rng.random() < 0.72 else iupac[rng.choice(list(iupac))] for _ in range(m)) + ")")
    pats += ["(ACGTACGT)", "(NNNNNNNNNN)".replace("N", "."), "(ACGTAC.TACGTACGTAC.TACGTACGTACGT)", "(.CGTACGTA)", "(ACGTACGT.)"]
    g = bytearray(genome(41, 9, 900_000))
    for _ in range(120):
        p = rng.randrange(100, len(g) - 3000)
        if b"\n" in g[p - 5:p + 2100] or b">" in g[p - 5:p + 2100]:
            continue
        kind = rng.randint(0, 2)
        ln = rng.randint(1, 2000)
        if kind == 0:
            g[p:p + ln] = b"N" * ln
        elif kind == 1:
            g[p:p + ln] = bytes(g[p:p + ln]).lower()
        else:
            g[p:p + 6] = b"RYKMSW"
    g[0:8] = b"ACGTACGT"                       # a hit at the very start of the file (header line replaced)
    g = bytes(g)
    ds = engine.load_dataset(g)
    try:
        for bs in (1600000, 50000):
            engine.set_buffer_size(bs)
            engine.set_batch_lookup(True)
            hits, off = engine.search_batch(ds, pats, "0ids", cap=1 << 22)
            st = engine.stats()
            assert 500 < st["qgram_chunks"] < len(pats), st          # most motifs indexed, some left to the dense kernel
            engine.set_batch_lookup(False)
            hits2, off2 = engine.search_batch(ds, pats, "0ids", cap=1 << 22)
            assert engine.stats()["qgram_chunks"] == 0
            assert np.array_equal(off, off2) and np.array_equal(hits, hits2)
            for i in rng.sample(range(len(pats)), 30) + list(range(len(pats) - 5, len(pats))):
                want = O.search(pats[i], g, "0ids", bufsize=bs, cap=1 << 22)
                got = [(int(b), int(e)) for b, e in hits[off[i]:off[i + 1]]]
                assert got == want, (pats[i], bs)
    finally:
        engine.set_batch_lookup(True)
        engine.set_buffer_size(1600000)
        ds.close()


def test_request_level_parity_with_reference_python(engine, request_golden):
    check_requests(engine, request_golden)


# ---- BASELINE.json configurations at (or near) full size -------------------------------
def test_config0_exact_motif_both_strands_yeast_size(engine, scan_mode):
    text = genome(0, 16, 12_000_000)
    ds = engine.load_dataset(text)
    conv, comp, opt = host.process_pattern("GATAAG", "dna", "Both strands", None, None, None, 0)
    for p in (conv, comp):
        got = [(int(b), int(e)) for b, e in engine.search(ds, p, opt)]
        assert got == O.search(p, text, opt, cap=1 << 22)
    ds.close()


def test_config1_peptide_one_substitution(engine):
    text = genome(1, 6000, 2_900_000, alphabet=PEP.encode(), name="YORF")
    ds = engine.load_dataset(text)
    conv, _, opt = host.process_pattern("CXXC[ILVM]XXHXXXH", "pep", None, None, None, "substitution", 1)
    got = [(int(b), int(e)) for b, e in engine.search(ds, conv, opt)]
    assert got == O.search(conv, text, opt) and len(got) > 0
    ds.close()


def test_peptide_codes_equal_raw_bytes_and_oracle(engine):
    # proteomes are scanned as 5-bit residue codes (six per word); lower case, 'X', '*', digits and header lines map to
    # codes that over-approximate -- the re-check on the raw bytes must remove every difference
    rng = random.Random(31)
    text = bytearray(genome(23, 800, 400_000, alphabet=PEP.encode(), name="YORF"))
    for _ in range(300):
        p = rng.randrange(50, len(text) - 50)
        if b"\n" in text[p - 2:p + 12] or b">" in text[p - 2:p + 12]:
            continue
        text[p:p + 8] = rng.choice([b"xxxxXXXX", b"cqqc*mqh", b"ACDEFGHI".lower(), b"C12C3L4H", b"BJOUZbjz"])
    text = bytes(text)
    ds = engine.load_dataset(text)
    cases = [("(CAAC[ILVM]QQH)", "1s"), ("(C..C[ILVM]..H...H)", "1s"), ("(C..C)", "0ids"), ("([^C]Q[ILVM].K)", "0ids"), ("(MKV.L)", "1ids"),
             ("(xxXX)", "0ids"), ("(C.?.?C[ILVM])", "0ids"), ("^(MK)", "0ids"), ("(W.W[FY]..[DE])", "2ids"), ("(BJOUZ)", "1s")]
    used = 0
    try:
        for pat, kopt in cases:
            engine.set_peptide_codes(True)
            a = engine.search(ds, pat, kopt)
            used += 1 if engine.stats()["packed"] == 2 else 0
            r = engine.search_request(ds, [pat], kopt)[0]
            engine.set_peptide_codes(False)
            b = engine.search(ds, pat, kopt)
            assert engine.stats()["packed"] == 0
            assert np.array_equal(a, b) and np.array_equal(a, r), (pat, kopt)
            assert [(int(x), int(y)) for x, y in a] == O.search(pat, text, kopt), (pat, kopt)
        assert used >= 7, used
    finally:
        engine.set_peptide_codes(True)
        ds.close()


def test_config2_20nt_two_errors_with_indels(engine, scan_mode):
    text = bytearray(genome(2, 16, 12_000_000))
    motif = b"TGACGTCAGATAAGCCGATT"
    rng = random.Random(8)
    for _ in range(300):                                     # plant mutated copies
        p = rng.randrange(100, len(text) - 100)
        if b"\n" in text[p - 30:p + 60] or b">" in text[p - 30:p + 60]:
            continue
        inst = bytearray(motif)
        for _ in range(rng.randint(0, 3)):
            q = rng.randrange(len(inst))
            op = rng.randint(0, 2)
            if op == 0:
                inst[q] = rng.choice(b"ACGT")
            elif op == 1:
                del inst[q]
            else:
                inst.insert(q, rng.choice(b"ACGT"))
        text[p:p + len(inst)] = inst[:len(inst)]
    text = bytes(text)
    ds = engine.load_dataset(text)
    conv, comp, opt = host.process_pattern(motif.decode(), "dna", "Both strands", None, None, None, 2)
    assert opt == "2ids"
    total = 0
    for p in (conv, comp):
        got = [(int(b), int(e)) for b, e in engine.search(ds, p, opt)]
        assert got == O.search(p, text, opt)
        total += len(got)
    assert total > 100
    ds.close()


def test_large_genome_properties(engine):
    # 256 Mb (the oracle is too slow to run whole): ordering, no overlap, determinism, every
    # planted copy is covered by a hit, and the hit list equals the oracle's inside sampled
    # windows (compared away from the window edges, where the oracle's scan start differs).
    n = 256_000_000
    text = bytearray(genome(9, 8, n))
    motif = b"GATTACAGATTACA"
    rng = random.Random(10)
    planted_at = []
    for _ in range(200):
        p = rng.randrange(5000, len(text) - 5000)
        if b"\n" in text[p - 40:p + 40]:
            continue
        text[p:p + len(motif)] = motif
        planted_at.append(p)
    text = bytes(text)
    ds = engine.load_dataset(text)
    hits = engine.search(ds, "(GATTACAGATTACA)", "2ids")
    assert engine.stats()["packed"] == 1                       # auto mode picks the packed scan for DNA
    # a request of this size has its specialised kernel compiled in the background (pm_engine_set_jit, mode 1): the
    # generic kernel answers meanwhile; once pm_jit_wait returns the specialised one runs -- same hit list
    pm.jit_wait()
    hits_jit = engine.search(ds, "(GATTACAGATTACA)", "2ids")
    assert engine.stats()["jit"] == 1
    assert np.array_equal(hits, hits_jit)
    b, e = hits["beg"], hits["end"]
    assert len(hits) > len(planted_at)
    assert np.all(b[1:] >= e[:-1]) and np.all(e > b)
    engine.set_scan_mode("bytes")
    again = engine.search(ds, "(GATTACAGATTACA)", "2ids")
    engine.set_scan_mode("auto")
    assert engine.stats()["packed"] == 0
    assert np.array_equal(hits, again)                         # both scan kernels, same hit list
    for p in planted_at:
        j = np.searchsorted(e, p, side="right")
        assert j < len(hits) and b[j] < p + len(motif), p      # some hit overlaps the planted copy
    centers = planted_at[:30] + [int(b[i]) for i in rng.sample(range(len(hits)), 30)]
    for c in centers:
        lo = max(text.rfind(b"\n", 0, c) + 1, c - 3000)
        hi = text.find(b"\n", c)
        hi = min(hi if hi >= 0 else len(text), c + 3000)
        local = [(x + lo, y + lo) for x, y in O.search("(GATTACAGATTACA)", text[lo:hi], "2ids")]
        inner = lambda hs: [h for h in hs if h[0] >= lo + 200 and h[1] <= hi - 200]
        mine = [(int(x), int(y)) for x, y in zip(b, e) if x >= lo and y <= hi]
        assert inner(mine) == inner(local), c
    ds.close()


def _request_case(rng, it):
    """(patterns, kopt, text): what one PatMatch request hands to the engine -- a pattern and a second one of the same
    length (for DNA the reference sends the reverse complement), random plan types, line anchors, EXTENDED plans"""
    if it % 5 == 4:
        pat, members, ops = extended_pattern(rng, DNA, rng.randint(4, 14))
        pat2, members2, ops2 = extended_pattern(rng, DNA, rng.randint(4, 14))
        text = extended_text(rng, DNA, members, ops, rng.randint(1, 3), 100, 2500) + extended_text(rng, DNA, members2, ops2, rng.randint(1, 3), 100, 2500)
        return [pat, pat2], "0ids", text
    m = rng.randint(3, 24) if rng.random() < 0.85 else rng.randint(25, 50)
    k = min(rng.choice([0, 0, 1, 1, 2, 2, 3]), m - 1)
    kopt = "%d%s" % (k, rng.choice(["ids", "ids", "s", "id", "is", "d"]))
    pat, members = random_pattern(rng, DNA, m)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    members2 = [[comp[c] for c in cls] for cls in reversed(members)]
    pat2 = "(" + "".join(c[0] if len(c) == 1 else "." if len(c) == 4 else "[" + "".join(c) + "]" for c in members2) + ")"
    if it % 7 == 0:
        pat, pat2 = "^" + pat, pat2 + "$"
    text = (random_text(rng, members, DNA, k, nrec=rng.randint(1, 3), lo=80, hi=3000) +
            random_text(rng, members2, DNA, k, nrec=rng.randint(1, 3), lo=80, hi=3000)).encode("latin-1")
    return [pat, pat2], kopt, text


def test_request_one_pass_equals_two_searches(engine, scan_mode):
    # pm_search_request: both patterns of a request in one pass and one pipeline; each hit list must equal the
    # single-pattern search and the oracle (the reference runs nrgrep_coords once per pattern, patmatch.py:733-743)
    rng = random.Random(20262)
    types = set()
    for it in range(160):
        pats, kopt, text = _request_case(rng, it)
        bufsize = rng.choice([1600000, 1600000, 100, 700])
        engine.set_buffer_size(bufsize)
        try:
            ds = engine.load_dataset(text)
            got = engine.search_request(ds, pats, kopt)
            single = [engine.search(ds, p, kopt) for p in pats]
            ds.close()
        finally:
            engine.set_buffer_size(1600000)
        for p, g, s in zip(pats, got, single):
            want = O.search(p, text, kopt, bufsize=bufsize)
            assert [(int(b), int(e)) for b, e in g] == want, (p, kopt, bufsize)
            assert np.array_equal(g, s), (p, kopt, bufsize)
            types.add(pm.plan(p, kopt)["type"])
    assert {"SIMPLE", "SPLIT", "EXT_BEG"} <= types, types


def test_request_on_genome_both_strands(engine):
    # the bench request on a 6 Mb genome with planted copies: one pass for motif + reverse complement, one host
    # synchronisation; a page-locked result buffer is filled directly
    from patmatchdocker_b200._native import pinned_empty, HIT_DTYPE
    g = bytearray(genome(31, 6, 6_000_000))
    rng = random.Random(9)
    conv, comp, opt = host.process_pattern("TGASTCANNNRYGATAAG", "dna", "Both strands", None, None, None, 2)
    site = "TGAGTCATTTACGATAAG"
    for _ in range(300):
        p = rng.randrange(100, len(g) - 100)
        if b"\n" in g[p - 30:p + 60] or b">" in g[p - 30:p + 60]:
            continue
        s = list(site)
        for _e in range(rng.randint(0, 2)):
            q = rng.randrange(len(s))
            r = rng.randint(0, 2)
            if r == 0:
                s[q] = rng.choice("ACGT")
            elif r == 1:
                del s[q]
            else:
                s.insert(q, rng.choice("ACGT"))
        s = "".join(s)
        if rng.random() < 0.5:
            s = s[::-1].translate(str.maketrans("ACGT", "TGCA"))
        g[p:p + len(s)] = s.encode()
    g = bytes(g)
    ds = engine.load_dataset(g)
    out, keep = pinned_empty(1 << 16, HIT_DTYPE)
    for kopt, pats in ((opt, [conv, comp]), ("0ids", ["(GATAAG)", "(CTTATC)"])):
        lists = engine.search_request(ds, pats, kopt, out=out)
        st = engine.stats()
        assert st["packed"] == 1
        lists2 = engine.search_request(ds, pats, kopt, out=out)        # capacity hints are warm now: one synchronisation
        assert engine.stats()["syncs"] == 1, engine.stats()
        for p, a, b in zip(pats, lists, lists2):
            want = O.search(p, g, kopt)
            assert len(want) > 100
            assert [(int(x), int(y)) for x, y in a] == want, (p, kopt)
            assert [(int(x), int(y)) for x, y in b] == want, (p, kopt)
    ds.close()
    del keep


def test_request_fills_device_equals_search(engine, scan_mode):
    # pm_request_fills_device (the multi-GPU path): asynchronous, header + hits of both patterns in device memory;
    # contiguous position ranges concatenate to the full per-pattern hit lists
    import torch
    from patmatchdocker_b200._native import request_header_rows
    rng = random.Random(515)
    rows = 1 << 15
    buf = torch.zeros((rows, 2), dtype=torch.int64, device="cuda")
    engine.use_torch_stream()
    try:
        for it in range(40):
            pats, kopt, text = _request_case(rng, it)
            bufsize = rng.choice([1600000, 64, 200, 1000])
            engine.set_buffer_size(bufsize)
            ds = engine.load_dataset(text)
            world = rng.randint(2, 5)
            cuts = sorted(rng.randint(0, len(text) + 1) for _ in range(world - 1))
            edges = [0] + cuts + [len(text) + 1]
            hr = request_header_rows(len(pats))
            got = [[] for _ in pats]
            for r in range(world):
                engine.request_fills_device(ds, pats, kopt, edges[r], edges[r + 1], 1 << 13, buf.data_ptr(), rows)
                h = buf.cpu().numpy()
                nh, ncand = int(h[0, 0]), int(h[0, 1])
                assert ncand <= (1 << 13) and nh + hr <= rows
                counts = [int(x) for x in h[2:hr].reshape(-1)[:len(pats)]]
                assert sum(counts) == nh
                off = hr
                for i, c in enumerate(counts):
                    got[i] += [(int(b), int(e)) for b, e in h[off:off + c]]
                    off += c
            ds.close()
            for p, gl in zip(pats, got):
                assert gl == O.search(p, text, kopt, bufsize=bufsize), (p, kopt, bufsize, edges)
    finally:
        engine.set_buffer_size(1600000)
        engine.set_stream(0)


def test_windowed_datasets_equal_whole_dataset(engine, scan_mode):
    # pm_dataset_create_window (the multi-GPU cold path): every "rank" uploads and packs only its position range plus
    # one buffer fill of overlap; the newline index of the whole file is the union of what the windows found.  The
    # fill-sharded request on the windows must give the hit lists of the whole dataset (and of the oracle).
    import torch
    from patmatchdocker_b200._native import request_header_rows
    from patmatchdocker_b200.distributed import shard_ranges
    rng = random.Random(616)
    rows = 1 << 15
    buf = torch.zeros((rows, 2), dtype=torch.int64, device="cuda")
    nl_rows = 4096
    engine.use_torch_stream()
    try:
        for it in range(24):
            pats, kopt, text = _request_case(rng, it)
            if isinstance(text, str):
                text = text.encode("latin-1")
            bufsize = rng.choice([1600000, 90, 300, 1000])
            engine.set_buffer_size(bufsize)
            world = rng.randint(2, 4)
            data = np.frombuffer(text, dtype=np.uint8)
            n = len(text)
            ranges = shard_ranges(n, world)
            wins, nls = [], []
            for r in range(world):
                beg, end = ranges[r]
                lo, hi = max(beg - 64, 0), min(n, end + min(bufsize, n) + 64)
                nl = torch.zeros(nl_rows, dtype=torch.int64, device="cuda")
                wins.append(engine.load_window(data, lo, hi, nl.data_ptr(), nl_rows))
                h = nl.cpu().numpy()
                nls.append(h[1:1 + int(h[0])])
            allnl = np.unique(np.concatenate(nls))
            assert [int(x) for x in allnl] == [i for i, c in enumerate(text) if c == 10]
            hr = request_header_rows(len(pats))
            got = [[] for _ in pats]
            for r in range(world):
                engine.set_newlines(wins[r], allnl)
                engine.request_fills_device(wins[r], pats, kopt, ranges[r][0], ranges[r][1], 1 << 13, buf.data_ptr(), rows)
                h = buf.cpu().numpy()
                nh = int(h[0, 0])
                counts = [int(x) for x in h[2:hr].reshape(-1)[:len(pats)]]
                assert sum(counts) == nh
                off = hr
                for i, c in enumerate(counts):
                    got[i] += [(int(b), int(e)) for b, e in h[off:off + c]]
                    off += c
                wins[r].close()
            for p, gl in zip(pats, got):
                assert gl == O.search(p, text, kopt, bufsize=bufsize), (p, kopt, bufsize, world)
    finally:
        engine.set_buffer_size(1600000)
        engine.set_stream(0)


def test_sharded_request_class_world_1(engine):
    # DeviceShardedSearch.request_fills / load_window with one rank (no process group): the code path bench.py runs
    # per rank -- device-side merge of the per-pattern lists, pinned result views, windowed cold dataset
    import torch
    from patmatchdocker_b200.distributed import DeviceShardedSearch
    rng = random.Random(717)
    sh = DeviceShardedSearch(engine, 0, 1, torch.device("cuda", 0))
    try:
        for it in range(12):
            pats, kopt, text = _request_case(rng, it)
            if isinstance(text, str):
                text = text.encode("latin-1")
            ds = engine.load_dataset(text)
            want = [np.array(h, copy=True) for h in engine.search_request(ds, pats, kopt)]
            got = sh.request_fills(ds, pats, kopt)
            for a, b in zip(got, want):
                assert np.array_equal(a, b), (pats, kopt)
            ds.close()
            host = torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy()).pin_memory()
            dw = sh.load_window(host)
            got = sh.request_fills(dw, pats, kopt)
            for a, b in zip(got, want):
                assert np.array_equal(a, b), (pats, kopt, "window")
            dw.close()
    finally:
        engine.set_stream(0)


def test_batch_with_errors_equals_single(engine):
    # pm_search_batch with k > 0: groups of patterns through the request pipeline
    rng = random.Random(99)
    g = genome(41, 4, 400_000)
    ds = engine.load_dataset(g)
    pats = []
    for _ in range(70):
        p, _m = random_pattern(rng, DNA, rng.randint(8, 16), cls_pct=0.1, dot_pct=0.05, neg_pct=0.0)
        pats.append(p)
    hits, off = engine.search_batch(ds, pats, "1ids")
    for i in (0, 1, 31, 32, 33, 69):
        want = O.search(pats[i], g, "1ids")
        assert [(int(b), int(e)) for b, e in hits[off[i]:off[i + 1]]] == want, pats[i]
    ds.close()


def test_specialised_kernels_equal_generic_and_oracle(engine):
    # pm_engine_set_jit: SPLIT requests as straight-line kernels compiled by NVRTC for exactly the request's patterns
    # (apx_jit.cpp).  Same candidates, same hits as the generic kernel and as the oracle; one and two patterns per
    # launch, 32- and 64-symbol windows, substitutions only and indels, texts with N runs / lower case / IUPAC letters.
    rng = random.Random(4242)
    engine.set_scan_mode("packed")
    used = 0
    try:
        for it in range(36):
            k = rng.randint(1, 3)
            m = rng.randint(max(2 * k + 2, 6), 26 if it % 3 else 56)
            pat, members = random_pattern(rng, DNA, m, cls_pct=0.3, dot_pct=0.25)
            kopt = "%d%s" % (k, rng.choice(["ids", "ids", "s", "id", "is", "ds"]))
            try:
                if pm.plan(pat, kopt)["type"] != "SPLIT":
                    continue
            except pm.NativeError:
                continue
            comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
            members2 = [[comp[c] for c in cls] for cls in reversed(members)]
            pat2 = "(" + "".join(c[0] if len(c) == 1 else "." if len(c) == 4 else "[" + "".join(c) + "]" for c in members2) + ")"
            text = bytearray((random_text(rng, members, DNA, k, nrec=rng.randint(1, 3), lo=2000, hi=60000 if it % 4 == 0 else 9000, plant=0.02) +
                              random_text(rng, members2, DNA, k, nrec=rng.randint(1, 2), lo=2000, hi=9000, plant=0.02)).encode("latin-1"))
            for _ in range(6):
                p = rng.randrange(0, len(text) - 40)
                if b"\n" in text[p:p + 40] or b">" in text[p:p + 40]:
                    continue
                text[p:p + rng.randint(1, 30)] = rng.choice([b"N" * 30, b"RYKMSW" * 5, bytes(text[p:p + 30]).lower()])[:30]
            text = bytes(text)
            if it % 5 == 0:
                engine.set_buffer_size(rng.choice([300, 2000]))
            ds = engine.load_dataset(text)
            pats = [pat, pat2] if it % 2 == 0 else [pat]
            engine.set_jit("always")
            a = engine.search_request(ds, pats, kopt)
            used += engine.stats()["jit"]
            engine.set_jit("off")
            b = engine.search_request(ds, pats, kopt)
            assert engine.stats()["jit"] == 0
            bufsize = 1600000
            ds.close()
            for p_, x, y in zip(pats, a, b):
                assert np.array_equal(x, y), (p_, kopt)
            if it % 5 != 0:
                for p_, x in zip(pats, a):
                    assert [(int(s), int(e)) for s, e in x] == O.search(p_, text, kopt, bufsize=bufsize), (p_, kopt)
            engine.set_buffer_size(1600000)
        assert used >= 10, used
        # exact requests (k = 0) through the same streaming kernel in exact mode: one or two patterns per launch,
        # IUPAC classes, a class that accepts non-ACGT bytes, line anchors, 33..70 positions (two halo words / tail compare)
        g = bytearray(genome(77, 5, 400_000))
        for _ in range(60):
            p = rng.randrange(100, len(g) - 200)
            if b"\n" in g[p - 5:p + 120] or b">" in g[p - 5:p + 120]:
                continue
            g[p:p + rng.randint(1, 60)] = rng.choice([b"N" * 60, b"RYKMSW" * 10, bytes(g[p:p + 60]).lower()])[:rng.randint(1, 60)]
        long_pat = "".join(rng.choice("ACGT") for _ in range(70))
        g[5000:5070] = long_pat.encode(); g[90000:90040] = long_pat[:40].encode()
        g = bytes(g)
        ds = engine.load_dataset(g)
        for pats, kopt in ((["(GATAAG)", "(CTTATC)"], "0ids"), (["(GAT[AG]AG)"], "0ids"), (["(GA.AAG[^C])", "(NNGATAAG)".replace("N", ".")], "0ids"),
                           (["^(>chr)", "(ACGT)$"], "0ids"), (["(" + long_pat + ")", "(" + long_pat[:40] + ")"], "0ids"), (["(" + long_pat[:33] + ")"], "0ids")):
            engine.set_jit("always")
            a = engine.search_request(ds, pats, kopt, cap=1 << 20)
            assert engine.stats()["jit"] == 1, pats
            engine.set_jit("off")
            b = engine.search_request(ds, pats, kopt, cap=1 << 20)
            assert engine.stats()["jit"] == 0
            for p_, x, y in zip(pats, a, b):
                assert np.array_equal(x, y), (p_, kopt)
                assert [(int(s), int(e)) for s, e in x] == O.search(p_, g, kopt, cap=1 << 22), (p_, kopt)
        ds.close()
        # low-complexity text: every start survives the dense filter, the per-lane survivor lists overflow and the
        # kernel falls back to unfiltered candidates for the rest (jx_unfiltered); hit lists must not change
        motif = "ACACACACACACACAC"
        text = (">r1\n" + "AC" * 30000 + "\n>r2\n" + "A" * 40000 + "\n>r3\n" + ("ACACACACTCACACAC" + "GT") * 3000 + "\n").encode()
        ds = engine.load_dataset(text)
        for pats, kopt in (([f"({motif})", "(GTGTGTGTGTGTGTGT)"], "2ids"), (["(AAAAAAAAAAAA)"], "1ids"), (["(ACACAC[AC]CAC.CAC)"], "2s"),
                           (["(ACAC)", "(AAAA)"], "0ids"), (["(A)"], "0ids")):
            engine.set_jit("always")
            a = engine.search_request(ds, pats, kopt)
            assert engine.stats()["jit"] == 1
            engine.set_jit("off")
            b = engine.search_request(ds, pats, kopt)
            for p_, x, y in zip(pats, a, b):
                assert np.array_equal(x, y), (p_, kopt)
                assert [(int(s), int(e)) for s, e in x] == O.search(p_, text, kopt), (p_, kopt)
        ds.close()
    finally:
        engine.set_jit("auto")
        engine.set_buffer_size(1600000)
        engine.set_scan_mode("auto")


def test_deployed_compat_mode_reproduces_the_stock_binary(engine, scan_mode):
    # pm_set_compat_deployed_glibc(1): hit lists of the unmodified nrgrep_coords with its DEFAULT allocator
    # (tests/golden/deployed_golden.json, 60 cases where they differ from the zero-scratch behaviour + 30 where not)
    import json, os
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "deployed_golden.json")))["cases"]
    try:
        for c in cases:
            text = c["text"].encode("latin-1")
            ds = engine.load_dataset(text)
            pm.set_compat_deployed_glibc(False)
            a = [(int(x), int(y)) for x, y in engine.search(ds, c["pattern"], c["kopt"])]
            pm.set_compat_deployed_glibc(True)
            b = [(int(x), int(y)) for x, y in engine.search(ds, c["pattern"], c["kopt"])]
            ds.close()
            assert a == [tuple(h) for h in c["zero_fill"]], (c["pattern"], c["kopt"])
            assert b == [tuple(h) for h in c["deployed"]], (c["pattern"], c["kopt"])
    finally:
        pm.set_compat_deployed_glibc(False)


def test_engine_is_safe_to_share_between_threads(engine):
    # mod_wsgi runs the Flask app with 15 request threads per process: every pm_* entry point holds
    # the engine's mutex, and the Python methods that combine calls hold Engine._lock, so concurrent requests on ONE
    # engine are serialised and each gets its own, correct hit list
    import threading
    rng = random.Random(818)
    cases = []
    for it in range(10):
        pats, kopt, text = _request_case(rng, it)
        if isinstance(text, str):
            text = text.encode("latin-1")
        cases.append((pats, kopt, text, [O.search(p_, text, kopt) for p_ in pats]))
    datasets = [engine.load_dataset(c[2]) for c in cases]
    errors = []

    def worker(tid):
        try:
            r = random.Random(tid)
            for _ in range(25):
                i = r.randrange(len(cases))
                pats, kopt, text, want = cases[i]
                if r.random() < 0.5:
                    got = engine.search_request(datasets[i], pats, kopt)
                else:
                    got = [engine.search(datasets[i], p_, kopt) for p_ in pats]
                for g, w in zip(got, want):
                    if [(int(b), int(e)) for b, e in g] != w:
                        errors.append((tid, pats, kopt))
        except Exception as ex:                              # noqa: BLE001
            errors.append((tid, repr(ex)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for d in datasets:
        d.close()
    assert not errors, errors[:3]


def test_text_sharded_batch_equals_whole_batch(engine):
    # pm_search_batch_fills: a motif batch over the buffer fills that start in a position range.  The per-range lists of
    # a partition of the file, merged motif by motif (distributed.merge_batch_shards), must equal pm_search_batch of the
    # whole file and the oracle -- for the lookup kernel (>= 64 exact motifs), the dense kernel (< 64) and the general
    # path (errors allowed: groups of motifs through the request pipeline)
    from patmatchdocker_b200 import distributed as D
    rng = random.Random(4242)
    iupac = {"R": "[AG]", "Y": "[CT]", "S": "[GC]", "W": "[AT]", "M": "[AC]", "K": "[GT]", "N": ".", "B": "[CGT]"}

    def motifs(count, lo, hi):
        return ["(" + "".join(rng.choice("ACGT") if rng.random() < 0.75 else iupac[rng.choice(list(iupac))] for _ in range(rng.randint(lo, hi))) + ")"
                for _ in range(count)]
    g = bytearray(genome(77, 7, 700_000))
    g[0:8] = b"ACGTACGT"
    g = bytes(g)
    ds = engine.load_dataset(g)
    try:
        for pats, kopt, bs in ((motifs(300, 6, 14) + ["(ACGTACGT)"], "0ids", 30000), (motifs(20, 5, 12), "0ids", 1600000),
                               (motifs(40, 9, 16), "1ids", 50000), (motifs(70, 6, 12), "0ids", 700)):
            engine.set_buffer_size(bs)
            whole_h, whole_o = engine.search_batch(ds, pats, kopt, cap=1 << 22)
            for world in (2, 5):
                parts = []
                for beg, end in D.shard_ranges(len(g), world):
                    h, o = engine.search_batch(ds, pats, kopt, cap=1 << 22, pos_range=(beg, end))
                    parts.append((np.array(h, copy=True), o))
                mh, mo = D.merge_batch_shards(parts, len(pats))
                assert np.array_equal(mo, whole_o), (kopt, bs, world)
                assert np.array_equal(mh, whole_h), (kopt, bs, world)
            h, o = engine.search_batch(ds, pats, kopt, pos_range=(5, 5))          # an empty range
            assert o[-1] == 0
            if kopt == "0ids":
                # compact lists (32-bit begins relative to the first fill of the range) carry the same hits
                from patmatchdocker_b200._native import expand_compact
                b_, o_, base_, ml_ = engine.search_batch_compact(ds, pats, kopt)
                assert np.array_equal(o_, whole_o) and np.array_equal(expand_compact(b_, o_, base_, ml_), whole_h)
                beg, end = D.shard_ranges(len(g), 3)[1]
                b_, o_, base_, ml_ = engine.search_batch_compact(ds, pats, kopt, pos_range=(beg, end))
                h, o = engine.search_batch(ds, pats, kopt, cap=1 << 22, pos_range=(beg, end))
                assert (base_ > 0) == (o[-1] > 0) and np.array_equal(o_, o) and np.array_equal(expand_compact(b_, o_, base_, ml_), h)
            else:
                with pytest.raises(pm.NativeError):
                    engine.search_batch_compact(ds, pats, kopt)
            for i in rng.sample(range(len(pats)), 6):
                want = O.search(pats[i], g, kopt, bufsize=bs, cap=1 << 22)
                assert [(int(b), int(e)) for b, e in whole_h[whole_o[i]:whole_o[i + 1]]] == want, (pats[i], kopt, bs)
    finally:
        engine.set_buffer_size(1600000)
        ds.close()


def test_merge_request_shards_kernel_emulated_ranks(engine):
    # pm_merge_request_shards: the blocks that `world` ranks would all-gather are produced here by ONE engine (one
    # pm_request_fills_device call per position range, laid out rank after rank), merged on the device, and must give
    # the per-pattern lists of pm_search_request; small buffer fills so that every range owns several of them
    import torch
    from patmatchdocker_b200 import distributed as D
    from patmatchdocker_b200._native import request_header_rows, HIT_DTYPE
    rng = random.Random(31337)
    engine.use_torch_stream()
    try:
        for it in range(14):
            pats, kopt, text = _request_case(rng, it)
            if isinstance(text, str):
                text = text.encode("latin-1")
            engine.set_buffer_size(rng.choice([300, 700, 1600000]))
            ds = engine.load_dataset(text)
            want = [np.array(h, copy=True) for h in engine.search_request(ds, pats, kopt)]
            for world in (1, 3, 8):
                npat, hr = len(pats), request_header_rows(len(pats))
                rows = hr + sum(len(w) for w in want) + 64
                allb = torch.zeros((world * rows, 2), dtype=torch.int64, device="cuda")
                for r, (beg, end) in enumerate(D.shard_ranges(len(text), world)):
                    engine.request_fills_device(ds, pats, kopt, beg, end, 1 << 14, allb[r * rows:].data_ptr(), rows)
                out = torch.zeros((world * rows, 2), dtype=torch.int64, device="cuda")
                engine.merge_request_shards(allb.data_ptr(), world, rows, npat, out.data_ptr(), out.shape[0])
                torch.cuda.synchronize()
                o = out.cpu().numpy()
                hdrs = o[: world * hr].reshape(world, hr, 2)
                assert np.array_equal(hdrs, allb.cpu().numpy().reshape(world, rows, 2)[:, :hr])
                counts = hdrs[:, 2:hr].reshape(world, -1)[:, :npat]
                flat = o[world * hr:].copy().view(HIT_DTYPE).reshape(-1)
                at = 0
                for p in range(npat):
                    c = int(counts[:, p].sum())
                    assert np.array_equal(flat[at:at + c], want[p]), (pats, kopt, world, p)
                    at += c
            ds.close()
    finally:
        engine.set_buffer_size(1600000)
        engine.set_stream(0)


def test_pipelined_batch_two_engines_equals_whole_batch(engine):
    # distributed.PipelinedBatch: sub-ranges of the file alternate between two engines of one GPU (two host threads);
    # the sub-range lists, merged motif by motif, are the lists of the whole batch
    import torch
    from patmatchdocker_b200 import distributed as D
    from patmatchdocker_b200._native import expand_compact
    rng = random.Random(5150)
    iupac = {"R": "[AG]", "Y": "[CT]", "S": "[GC]", "W": "[AT]", "N": "."}
    pats = ["(" + "".join(rng.choice("ACGT") if rng.random() < 0.75 else iupac[rng.choice(list(iupac))] for _ in range(rng.randint(6, 13))) + ")"
            for _ in range(200)]
    g = genome(78, 5, 500_000)
    dev = torch.from_numpy(np.frombuffer(g, dtype=np.uint8).copy()).cuda()
    other = pm.Engine(0)
    try:
        for bs in (20000, 1600000):
            engine.set_buffer_size(bs)
            other.set_buffer_size(bs)
            d0 = engine.wrap_device(dev.data_ptr(), dev.numel())
            d1 = other.wrap_device(dev.data_ptr(), dev.numel())
            whole_h, whole_o = engine.search_batch(d0, pats, "0ids", cap=1 << 22)
            for parts in (1, 2, 5):
                pb = D.PipelinedBatch([engine, other], [d0, d1], parts=parts)
                for rep in range(2):                              # the second search reuses the page-locked buffers
                    res = pb.search(pats, "0ids")
                    shards = [(expand_compact(b, o, base, ml), o) for b, o, base, ml in res]
                    mh, mo = D.merge_batch_shards(shards, len(pats))
                    assert np.array_equal(mo, whole_o) and np.array_equal(mh, whole_h), (bs, parts, rep)
            lo, hi = D.shard_ranges(len(g), 3)[1]
            res = D.PipelinedBatch([engine, other], [d0, d1], parts=3).search(pats, "0ids", pos_range=(lo, hi))
            mh, mo = D.merge_batch_shards([(expand_compact(b, o, base, ml), o) for b, o, base, ml in res], len(pats))
            h, o = engine.search_batch(d0, pats, "0ids", cap=1 << 22, pos_range=(lo, hi))
            assert np.array_equal(mo, o) and np.array_equal(mh, h), bs
            d0.close()
            d1.close()
    finally:
        engine.set_buffer_size(1600000)
        other.close()
