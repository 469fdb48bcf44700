#!/usr/bin/env python
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Runs only where /root/reference exists (the build container).  Three fixtures:

  search_golden.json   nrgrep_coords (the reference engine binary) run exactly as
                       patmatch.py:733 runs it, on small synthetic .seq files.
  pattern_golden.json  patmatch_to_nrgrep.pl outputs (patmatch.py:291-297).
  request_golden.json  the reference's own Python (patmatch.run_test, patmatch.py:768-838)
                       end to end on synthetic orf_dna.seq / orf_pep.seq datasets.

The engine reads an uninitialised scratch cell while choosing its k+1 pieces
(esimplePreproc @415a68, see DESIGN.md "Reference UB"), so its approximate-search
output depends on allocator history.  The fixtures pin the DEFINED behaviour by
running the unmodified binary with glibc tunables that make malloc return zeroed
memory (GLIBC_TUNABLES below).  `deployed_agrees` records, per case, whether the
binary run with the default allocator printed the same hit list.
"""
import json
import os
import random
import re
import subprocess
import sys
import tempfile

REF = "/root/reference"
BIN = REF + "/www/bin/nrgrep_coords"
PERL = REF + "/www/bin/patmatch_to_nrgrep.pl"
HERE = os.path.dirname(os.path.abspath(__file__))
ZERO_ENV = dict(os.environ, GLIBC_TUNABLES="glibc.malloc.tcache_count=0:glibc.malloc.perturb=255")
DNA, PEP = "ACGT", "ACDEFGHIKLMNPQRSTVWY"


def run_engine(pattern, kopt, path, env, bufsize=1600000):
    out = subprocess.run([BIN, "-i", "-b", str(bufsize), "-k", kopt, pattern, path], capture_output=True, env=env).stdout
    out = out.decode("latin-1")
    return out.split("\n")[0], [[int(a), int(b)] for a, b in re.findall(r"^\[(\d+), (\d+)\]: ", out, re.M)]


def planted(rng, pattern_chars, alpha, k):
    s = [rng.choice(c) for c in pattern_chars]
    for _ in range(rng.randint(0, k + 1)):
        if len(s) < 2:
            break
        op, p = rng.randint(0, 2), rng.randrange(len(s))
        if op == 0:
            s[p] = rng.choice(alpha)
        elif op == 1:
            del s[p]
        else:
            s.insert(p, rng.choice(alpha))
    return "".join(s)


def random_case(rng, alpha, m, k, ids):
    pat, members = "(", []
    for _ in range(m):
        r = rng.random()
        if r < 0.07:
            pat += "."
            members.append(alpha)
        elif r < 0.25:
            n = rng.randint(2, 3)
            chars = rng.sample(alpha, n)
            if rng.random() < 0.15:
                pat += "[^" + "".join(chars) + "]"
                members.append([c for c in alpha if c not in chars])
            else:
                pat += "[" + "".join(chars) + "]"
                members.append(chars)
        else:
            c = rng.choice(alpha)
            pat += c
            members.append([c])
    pat += ")"
    lines = []
    for r in range(rng.randint(1, 3)):
        lines.append(">seq%d synthetic record" % r)
        t = ""
        target = rng.randint(80, 500)
        while len(t) < target:
            t += planted(rng, members, alpha, k) if rng.random() < 0.3 else "".join(rng.choice(alpha) for _ in range(rng.randint(1, 14)))
        if rng.random() < 0.3:
            t = "".join(ch.lower() if rng.random() < 0.3 else ch for ch in t)
        if rng.random() < 0.2:
            p = rng.randrange(len(t))
            t = t[:p] + "NNNN" + t[p:]
        lines.append(t)
    return pat, "%d%s" % (k, ids), "\n".join(lines) + "\n"


def search_fixture():
    rng = random.Random(20261018)
    cases = []
    fixed = [
        ("(GATAAG)", "0ids", ">chr1 test\nACGTGATAAGCCGATAAGTT\nCTTATCGG\n>chr2\nGGGATAAGG\n"),
        ("(AAA)", "0ids", ">s1\nAAAAAAA\n>s2\nAAATAAAA\n"),
        ("(AC..)", "0ids", ">ab AC\nGTNNACGT\nacgtACGT\n\nAC\nGT\n"),
        ("(..AC)", "0ids", ">ab AC\nGTNNACGT\nacgtACGT\n\nAC\nGT\n"),
        ("([^A]C[GT])", "0ids", ">s\nACGTCGTTCTNCG\n"),
        ("(ACGT)", "1ids", ">s\nAAGTACGTTTACTGTTTACGGTACGACG\n"),
        ("(ACGA)", "1ids", ">s\nACGACGATT\n"),
        ("(AAAA)", "1ids", ">s\nAAATAAAAATAAAAAAAA\n"),
        ("(TAGGAATAATC)", "1id", ">s\nCAATACGGTTCAATTAGGAATAATCGCCCTACTGCATG\n"),
        ("(GATAAG)", "0ids", ""),
        ("(GATAAG)", "1ids", ">only header"),
        ("(GATAAG)", "2ids", ">s\n\n\n>t\nGATAAG"),
        ("(GATAAGGATAAGCCGATTGA)", "2ids", ">s\nTTGATAAGGATAAGCCGATTGATTGATAAGATAAGCCGATTGATTGATAAGGATAACCCGTTTGATT\n"),
        ("(C..C[ILVM]..H...H)", "1s", ">p1\nMKCAACLQQHAAAHLLCAACVQQHAAAHKKCAACLQQHAAHH*\n"),
    ]
    for pat, kopt, text in fixed:
        cases.append((pat, kopt, text))
    for alpha in (DNA, PEP):
        for _ in range(120):
            m = rng.randint(3, 24) if rng.random() < 0.85 else rng.randint(25, 60)
            k = rng.choice([0, 1, 1, 2, 2, 3])
            k = min(k, m - 1)
            ids = rng.choice(["ids", "ids", "s", "id", "is", "ds", "i", "d"])
            cases.append(random_case(rng, alpha, m, k, ids))
    cases = [(p, k, t, 1600000) for p, k, t in cases]
    # small -b values: buffer fills (bufLoad @41bbf0) cut the file; hits never cross a fill
    for _ in range(80):
        m = rng.randint(3, 12)
        k = min(rng.choice([0, 0, 1, 2]), m - 1)
        pat, kopt, text = random_case(rng, DNA, m, k, rng.choice(["ids", "s", "id"]))
        if k == 0:
            pat = pat.replace(rng.choice("ACGT"), ".", 1)
        lines = text.split("\n")
        text = "\n".join(l if rng.random() < 0.7 else l[: rng.choice([0, 1, 7, 40])] for l in lines)
        cases.append((pat, kopt, text, rng.choice([16, 33, 64, 100, 128, 257, 1000])))
    cases.append(("(A.C)", "0ids", ">s\nGGA\nCGG\nA\nC\n", 6))
    # line anchors: patmatch_to_nrgrep.pl turns '<' / '>' into a leading '^' / trailing '$' (recCheckLeftContext
    # @402170, recCheckRightContext @4021e0).  Short records, so that the anchors do fire.  Own generator: the
    # cases above stay byte-identical.
    arng = random.Random(4242)
    for _ in range(120):
        alpha = arng.choice([DNA, PEP])
        k = arng.choice([0, 0, 1, 1, 2])
        m = arng.randint(max(2, 2 * k + 1), 10)
        pat, kopt, _ = random_case(arng, alpha, m, k, arng.choice(["ids", "s", "id", "is"]) if k else "ids")
        members = []
        for tok in re.findall(r"\[\^?[A-Z]+\]|\.|[A-Z]", pat[1:-1]):
            if tok == ".":
                members.append(list(alpha))
            elif tok.startswith("[^"):
                members.append([c for c in alpha if c not in tok[2:-1]])
            elif tok.startswith("["):
                members.append(list(tok[1:-1]))
            else:
                members.append([tok])
        mode = arng.randint(1, 3)
        pat = ("^" if mode & 1 else "") + pat + ("$" if mode & 2 else "")
        lines = []
        for r in range(arng.randint(2, 8)):
            lines.append(">s%d" % r)
            s = "".join(arng.choice(c) for c in members)
            body = "".join(arng.choice(alpha) for _ in range(arng.randint(0, 12)))
            t = [s + body, body + s, s, s + s + body + s, body + s + body][arng.randint(0, 4)]
            if k and arng.random() < 0.5 and len(t) > 2:
                q = arng.randrange(len(t))
                t = t[:q] + arng.choice(alpha) + t[q + 1:]
            lines.append(t)
        cases.append((pat, kopt, "\n".join(lines) + "\n", arng.choice([1600000, 1600000, 1600000, 40, 64])))
    # EXTENDED patterns (k = 0): what PatMatch's X{m,n} / X{m,} repeats become (patmatch_to_nrgrep.pl:184-211).
    # First and last position mandatory (the shapes the parser does not rewrite).  Own generator again.
    xrng = random.Random(777)
    for it in range(140):
        alpha = xrng.choice([DNA, DNA, PEP])
        m = xrng.randint(3, 14)
        pat, members, ops = "(", [], []
        for j in range(m):
            r = xrng.random()
            if r < 0.12:
                pat += "."
                members.append(list(alpha))
            elif r < 0.27:
                chars = xrng.sample(alpha, 2)
                pat += "[" + "".join(chars) + "]"
                members.append(chars)
            else:
                c = xrng.choice(alpha)
                pat += c
                members.append([c])
            op = ""
            if 0 < j < m - 1:
                q = xrng.random()
                op = "?" if q < 0.3 else "*" if q < 0.36 else "+" if q < 0.4 else ""
            pat += op
            ops.append(op)
        pat += ")"
        if not any(ops):
            continue
        if it % 9 == 0:
            pat = "^" + pat
        lines = []
        for r in range(xrng.randint(1, 4)):
            lines.append(">x%d" % r)
            t = ""
            target = xrng.randint(40, 400)
            while len(t) < target:
                if xrng.random() < 0.35:
                    for c, op in zip(members, ops):
                        reps = 1 if op == "" else xrng.randint(0, 1) if op == "?" else xrng.randint(0, 2) if op == "*" else xrng.randint(1, 2)
                        t += "".join(xrng.choice(c) for _ in range(reps))
                else:
                    t += "".join(xrng.choice(alpha) for _ in range(xrng.randint(1, 10)))
            lines.append(t)
        cases.append((pat, "0ids", "\n".join(lines) + "\n", xrng.choice([1600000, 1600000, 1600000, 50, 200])))
    # operators on the first / last position: the reference's parser rewrites them (one optional first position dropped
    # or a leading '+' stripped; every trailing optional position dropped, else a trailing '+' stripped)
    erng = random.Random(888)
    for it in range(70):
        alpha = erng.choice([DNA, DNA, PEP])
        m = erng.randint(3, 10)
        toks, ops = [], []
        for j in range(m):
            r = erng.random()
            toks.append("." if r < 0.1 else "[" + "".join(erng.sample(alpha, 2)) + "]" if r < 0.25 else erng.choice(alpha))
            ops.append(erng.choice(["", "", "?", "?", "*", "+"]) if (j in (0, m - 1) or erng.random() < 0.25) else "")
        if ops[0] in ("?", "*") and ops[1] in ("?", "*"):
            ops[1] = ""                        # would still begin with an optional position after the rewrite
        if not any(ops):
            ops[-1] = "?"
        pat = "(" + "".join(a + b for a, b in zip(toks, ops)) + ")"
        lines = []
        for r in range(erng.randint(1, 3)):
            lines.append(">e%d" % r)
            lines.append("".join(erng.choice(alpha) for _ in range(erng.randint(30, 300))))
        cases.append((pat, "0ids", "\n".join(lines) + "\n", erng.choice([1600000, 1600000, 100])))
    # the reference's quirk: a run of two or more optional positions next to the anchor cannot be skipped as a whole
    cases.append(("(GAT.?.?.?AAGTCC)", "0ids", ">q\nCCGATAAGTCCAA\nCCGATCAAGTCCAA\nCCGATCCCAAGTCCAA\n", 1600000))
    # forward scan (no sub-pattern is cheap enough) of a pattern that still begins with optional positions after the
    # parser's rewrite: occurrences that must skip them on the first byte of a scan range or record are not reported
    for pat, text in (("(T?[TG]*A)", "AATAAC\nATTA\nGA\n"), ("(A?A?C)", "CCACAAC\nC\nAC\n"), ("(.?.?G)", ">g\nGGAGTTG\nG\n"),
                      ("(C*C?[AG]?T)", "TTCTAT\nT\nGT\nCCCGT\n")):
        cases.append((pat, "0ids", text, 1600000))
        cases.append((pat, "0ids", text, 8))
    # exact patterns of more than 64 positions (multi-word masks in the reference)
    lrng = random.Random(6500)
    for it in range(14):
        alpha = lrng.choice([DNA, DNA, PEP])
        m = lrng.randint(65, 200)
        pat, kopt, _ = random_case(lrng, alpha, m, 0, "ids")
        members = []
        for tok in re.findall(r"\[\^?[A-Z]+\]|\.|[A-Z]", pat[1:-1]):
            members.append(list(alpha) if tok == "." else [c for c in alpha if c not in tok[2:-1]] if tok.startswith("[^")
                           else list(tok[1:-1]) if tok.startswith("[") else [tok])
        lines = []
        for r in range(lrng.randint(1, 3)):
            lines.append(">L%d" % r)
            t = ""
            for _ in range(lrng.randint(1, 8)):
                s = [lrng.choice(c) for c in members]
                if lrng.random() < 0.4:
                    s[lrng.randrange(len(s))] = lrng.choice(alpha)
                t += "".join(s) + "".join(lrng.choice(alpha) for _ in range(lrng.randint(0, 25)))
            lines.append(t)
        cases.append((pat, "0ids", "\n".join(lines) + "\n", lrng.choice([1600000, 1600000, 500])))
    # ... and approximate ones (the verification NFA runs on multi-word masks)
    for it in range(12):
        alpha = lrng.choice([DNA, DNA, PEP])
        m, k = lrng.randint(65, 160), lrng.choice([1, 1, 2])
        cases.append(random_case(lrng, alpha, m, k, lrng.choice(["ids", "s", "id"])) + (lrng.choice([1600000, 1600000, 800]),))
    out = []
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "t.seq")
        for pat, kopt, text, bufsize in cases:
            with open(path, "w") as f:
                f.write(text)
            banner, hits = run_engine(pat, kopt, path, ZERO_ENV, bufsize)
            _, dep = run_engine(pat, kopt, path, os.environ, bufsize)
            out.append({"pattern": pat, "kopt": kopt, "text": text, "bufsize": bufsize, "banner": banner, "hits": hits,
                        "deployed_agrees": dep == hits})
    json.dump(out, open(os.path.join(HERE, "search_golden.json"), "w"), indent=0)
    agree = sum(c["deployed_agrees"] for c in out)
    print("search_golden: %d cases, %d hits, default-allocator run agrees on %d" % (len(out), sum(len(c["hits"]) for c in out), agree))


def pattern_fixture():
    rng = random.Random(7)

    def perl(cls, p):
        try:
            return subprocess.run(["perl", PERL, cls, p], capture_output=True, text=True, timeout=5).stdout
        except subprocess.TimeoutExpired:
            return None

    def rnd(alpha, nested=True):
        s = ""
        for _ in range(rng.randint(1, 9)):
            r = rng.random()
            if r < 0.6:
                s += rng.choice(alpha)
            elif r < 0.75:
                s += "[" + ("^" if rng.random() < 0.2 else "") + "".join(rng.choice(alpha) for _ in range(rng.randint(1, 4))) + "]"
            elif r < 0.85 and nested:
                s += "(" + rnd(alpha, False) + ")"
            else:
                s += rng.choice(alpha)
            if rng.random() < 0.2:
                a = rng.randint(0, 3)
                b = a + rng.randint(0, 3)
                s += rng.choice(["{%d}" % a, "{%d,%d}" % (a, b), "{%d,}" % a, "{,%d}" % b])
        return s

    cases = [("-n", "GATAAG"), ("-c", "GATAAG"), ("-n", "GATRNNY{2,4}C"), ("-c", "(GAT[AG]AG)"), ("-p", "CX{2,4}C[ILVM]JOBZ"),
             ("-n", "<ATG(TAG){2}N>"), ("-c", "<ATG(TAG){2}N>"), ("-n", "AT[^CG]W{3,}"), ("-n", "gat aag"), ("-p", "c x x c")]
    for _ in range(300):
        cls = rng.choice(["-n", "-c", "-p"])
        alpha = "ACGTRYSWMKVHDBNX" if cls != "-p" else "ACDEFGHIKLMNPQRSTVWYXJOBZ"
        p = rnd(alpha)
        if rng.random() < 0.1:
            p = "<" + p
        if rng.random() < 0.1:
            p += ">"
        cases.append((cls, p))
    out = []
    for cls, p in cases:
        o = perl(cls, p)
        if o is None:
            continue
        out.append({"cls": cls, "pattern": p, "nrgrep": o})
        if cls == "-n":                       # what process_pattern feeds back for the other strand
            o2 = perl("-c", o)
            if o2 is not None:
                out.append({"cls": "-c", "pattern": o, "nrgrep": o2})
    json.dump(out, open(os.path.join(HERE, "pattern_golden.json"), "w"), indent=0)
    print("pattern_golden: %d cases" % len(out))


def request_fixture():
    """The reference's own patmatch.run_test on synthetic datasets."""
    sys.path.insert(0, REF + "/www/FlaskApp/FlaskApp")
    import patmatch as ref
    rng = random.Random(99)
    os.environ.update(GLIBC_TUNABLES=ZERO_ENV["GLIBC_TUNABLES"])
    cwd = os.getcwd()
    out = {"datasets": {}, "requests": []}
    with tempfile.TemporaryDirectory() as td:
        dna, pep, locus = [], [], []
        for g in range(40):
            name = "Y%s%03d%s" % (rng.choice("ABCD"), g, rng.choice("WC"))
            seq = "".join(rng.choice(DNA) for _ in range(rng.randint(150, 900)))
            if g % 5 == 0:
                p = rng.randrange(len(seq) - 20)
                seq = seq[:p] + "TGACGTCAGATAAG" + seq[p:]
            dna.append(">%s %s SGDID:S%06d, Chr I from 1-100, Verified ORF, \"synthetic\"\n%s\n" % (name, "GEN%d" % g, g, seq))
            prot = "M" + "".join(rng.choice(PEP) for _ in range(rng.randint(60, 300)))
            if g % 4 == 0:
                p = rng.randrange(len(prot) - 20)
                prot = prot[:p] + "CAACLQQHAAAH" + prot[p:]
            pep.append(">%s %s SGDID:S%06d\n%s*\n" % (name, "GEN%d" % g, g, prot))
            locus.append("%s\t%s\tS%06d\tsynthetic gene %d\n" % (name, "GEN%d" % g if g % 3 else name, g, g))
        files = {"orf_dna.seq": "".join(dna), "orf_pep.seq": "".join(pep), "locus.txt": "".join(locus)}
        for fn, content in files.items():
            open(os.path.join(td, fn), "w").write(content)
        out["datasets"] = files
        os.chdir(td)
        reqs = [
            dict(pattern="GATAAG", seqtype="dna", strand="Both strands"),
            dict(pattern="GATAAG", seqtype="dna", strand="Watson strand"),
            dict(pattern="GATAAG", seqtype="dna", strand="Reverse complement strand"),
            dict(pattern="TGACGTCAGATAAG", seqtype="dna", strand="Both strands", mismatch=2),
            dict(pattern="TGACGTCAGATAAG", seqtype="dna", strand="Both strands", mismatch=1, substitution="substitution"),
            dict(pattern="TGASGTCANATWAG", seqtype="dna", strand=None, mismatch=2, insertion="insertion", deletion="deletion"),
            dict(pattern="<ATG", seqtype="dna", strand="Watson strand"),
            dict(pattern="RYRYRY", seqtype="dna", strand="Both strands", max_hits="no limit"),
            dict(pattern="CXXC[ILVM]XXHXXXH", seqtype="pep", mismatch=1, substitution="substitution"),
            dict(pattern="CAACLQQHAAAH", seqtype="pep", mismatch=2),
            dict(pattern="<MK", seqtype="pep"),
            dict(pattern="KK>", seqtype="pep"),
            dict(pattern="J[^P]OB", seqtype="pep", max_hits=25),
            dict(pattern="AC", seqtype="dna"),
            dict(pattern="GATEAG", seqtype="dna"),
            # X{m,n} repeats: nrgrep's EXTENDED engine (appended: the requests above keep their fixtures)
            dict(pattern="TGACN{2,4}CAGA", seqtype="dna", strand="Both strands"),
            dict(pattern="GAN{0,3}TAAG", seqtype="dna", strand="Watson strand"),
            dict(pattern="CX{2,4}C[ILVM]", seqtype="pep"),
            dict(pattern="MX{0,2}K", seqtype="pep", max_hits=50),
            dict(pattern="N{0,2}GATAAG", seqtype="dna", strand="Both strands"),       # repeats at the ends: parser rewrites
            dict(pattern="CAACX{1,}", seqtype="pep"),
        ]
        for r in reqs:
            kw = dict(r)
            res = ref.run_test(kw.pop("pattern"), root_dir=REF, root_data_dir=td + "/", **kw)
            data, uniq, total, err = res
            out["requests"].append({"request": r, "hits": data, "uniqueHits": uniq, "totalHits": total, "error": err})
        os.chdir(cwd)
    json.dump(out, open(os.path.join(HERE, "request_golden.json"), "w"), indent=0)
    print("request_golden: %d requests, %d rows" % (len(out["requests"]), sum(len(r["hits"]) for r in out["requests"])))


if __name__ == "__main__":
    if not os.path.exists(BIN):
        sys.exit("reference not present: fixtures can only be regenerated in the build container")
    search_fixture()
    pattern_fixture()
    request_fixture()
