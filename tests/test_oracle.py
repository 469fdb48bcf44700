"""CPU: the oracle against vectors produced by the reference itself (no GPU needed)."""
import oracle_lib as O


def test_oracle_matches_reference_golden(search_golden):
    bad = []
    for c in search_golden:
        got = [list(h) for h in O.search(c["pattern"], c["text"], c["kopt"], bufsize=c["bufsize"])]
        if got != c["hits"]:
            bad.append((c["pattern"], c["kopt"], got[:4], c["hits"][:4]))
    assert not bad, bad[:5]


def test_golden_covers_every_plan_type(search_golden):
    seen = set()
    for c in search_golden:
        if O.is_extended(c["pattern"]):
            _, xpl = O.plan_ext(c["pattern"])
            seen.add("EXT%d" % xpl.type if xpl is not None else "SIMPLE")       # None: the parser's rewrites left a SIMPLE pattern
            continue
        _, pl = O.plan(c["pattern"], c["kopt"])
        seen.add(O.TYPE_NAMES[pl.type])
    assert seen == {"SIMPLE", "SPLIT", "BWD", "FWD", "EXT2", "EXT3"}


def test_banner(search_golden):
    for c in search_golden:
        k = O.parse_kopt(c["kopt"])[0]
        ext = O.is_extended(c["pattern"]) and O.plan_ext(c["pattern"])[1] is not None
        want = "EXTENDED search" if ext else "SIMPLE search" if k == 0 else "ESIMPLE search"
        assert c["banner"] == want, c["pattern"]


def test_known_answers():
    # hand-checked behaviours of the reference (see DESIGN.md "semantics")
    assert O.search("(AAA)", ">s1\nAAAAAAA\n", "0ids") == [(4, 7), (7, 10)]                  # restart at hit end
    assert O.search("(AC..)", ">ab AC\nGTNN\n", "0ids") == [(4, 8)]                           # '.' crosses '\n' when k = 0
    assert O.search("(ACGT)", ">s\nAAGT\n", "1ids") == [(4, 7)]                               # shortest left extension
    assert O.search("(ACGA)", ">s\nACGACGA\n", "1ids") == [(3, 7), (7, 10)]                   # never left of the scan start


def test_buffer_fills_cut_hits():
    # nrgrep scans one -b sized fill at a time (bufLoad @41bbf0): with -b 6 the file ">s\nGGA\nCGG\n"
    # is cut after the first '\n' inside each fill, so the k=0 hit "A\nC" that crosses a cut is lost
    assert O.search("(A.C)", ">s\nGGA\nCGG\n", "0ids") == [(5, 8)]
    assert O.search("(A.C)", ">s\nGGA\nCGG\n", "0ids", bufsize=6) == []


def test_32bit_piece_test_quirk():
    # k+1 pieces of length 11 occupy 44 state bits; the reference tests piece i with a 32-bit
    # `1 << bit`, so pieces 2 and 3 can never start a verification (esimpleScan @41384b)
    pat = "([ACT]GCGGCT[^TG]TA[AG]TTCCCCC[AG]A[ATG]GT[^GAT]TA.CGGGAA[AC].TGGTCCA[AC][AT]CC)"
    _, pl = O.plan(pat, "3id")
    assert (pl.type, pl.L, list(pl.V)[:4]) == (1, 11, [1, 12, 23, 35])
    text = ">s\nCTCCCTACCCGTGCGGCTACTAATTCCCCCAAAGGTCTAGCGGGAACTGGTCCAAACCGAGTGCG\n"
    assert O.search(pat, text, "3id") == []


def test_deployed_compat_reproduces_the_stock_binary_with_its_default_allocator():
    # tests/golden/deployed_golden.json (tools/deployed_gap.py --emit): hit lists printed by the unmodified
    # nrgrep_coords with the default glibc allocator ("deployed") and with zero-filled malloc ("zero_fill").  The
    # oracle's default is the zero-fill behaviour; with nro_set_compat(1) it must print the deployed lists, including
    # the cases where the two differ.
    import json, os
    path = os.path.join(os.path.dirname(__file__), "golden", "deployed_golden.json")
    cases = json.load(open(path))["cases"]
    differing = 0
    try:
        for c in cases:
            text = c["text"].encode("latin-1")
            O.set_compat(False)
            assert O.search(c["pattern"], text, c["kopt"]) == [tuple(h) for h in c["zero_fill"]], (c["pattern"], c["kopt"])
            O.set_compat(True)
            assert O.search(c["pattern"], text, c["kopt"]) == [tuple(h) for h in c["deployed"]], (c["pattern"], c["kopt"])
            differing += 1 if c["deployed"] != c["zero_fill"] else 0
    finally:
        O.set_compat(False)
    assert differing >= 40 and len(cases) - differing >= 20, (differing, len(cases))
