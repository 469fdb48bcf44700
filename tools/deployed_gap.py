#!/usr/bin/env python
"""How far is the product (= the oracle's defined behaviour: esimplePreproc's never-written scratch cells read as +0.0)
from the DEPLOYED reference binary, whose piece choice depends on what malloc returns ("Reference UB")?

Runs the stock nrgrep_coords twice per case -- default allocator, and zero-filled malloc (GLIBC_TUNABLES) -- on random
approximate searches and on the BASELINE motifs, and compares hit lists and search plans.  CPU only; needs
/root/reference (container).  usage: deployed_gap.py [ncases]"""
import json, os, random, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from synth import DNA, random_pattern, random_text
BIN = "/root/reference/www/bin/nrgrep_coords"
ZERO = dict(os.environ, GLIBC_TUNABLES="glibc.malloc.tcache_count=0:glibc.malloc.perturb=255")


def run(pat, kopt, path, env):
    out = subprocess.run([BIN, "-i", "-b", "1600000", "-k", kopt, pat, path], capture_output=True, text=True, env=env).stdout
    hits = []
    for line in out.splitlines():
        if line.startswith("["):
            a, b = line[1:line.index("]")].split(",")
            hits.append((int(a), int(b)))
    return hits


def main():
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    emit = sys.argv[sys.argv.index("--emit") + 1] if "--emit" in sys.argv else None      # write golden vectors (deployed hit lists)
    vectors, nsame = [], 0
    rng = random.Random(2026)
    stats = {"cases": 0, "split_plans": 0, "deployed_differs_from_zero_fill": 0, "oracle_differs_from_zero_fill": 0,
             "same_hit_count_when_different": 0, "by_m": {}}
    examples = []
    fixed = [("(TGA[GC]TCA...[AG][CT]GATAAG)", "2ids"), ("(TGACGTCAGATAAGCCGATT)", "2ids"), ("(GATTACAGATTACA)", "2ids")]
    with tempfile.TemporaryDirectory() as td:
        for it in range(ncases + len(fixed)):
            if it < len(fixed):
                pat, kopt = fixed[it]
                members = None
                text = ">r\n" + "".join(rng.choice("ACGT") for _ in range(60000)) + "\n"
            else:
                k = rng.randint(1, 3)
                m = rng.randint(max(2 * k + 2, 6), 30)
                pat, members = random_pattern(rng, DNA, m, cls_pct=0.2, dot_pct=0.15, neg_pct=0.0)
                kopt = "%d%s" % (k, rng.choice(["ids", "ids", "s", "id", "is"]))
                text = random_text(rng, members, DNA, k, nrec=2, lo=1500, hi=4000, plant=0.05)
            try:
                plan = O.plan(pat, kopt) if hasattr(O, "plan") else None
            except Exception:
                plan = None
            path = os.path.join(td, "t.seq")
            open(path, "w").write(text)
            dep = run(pat, kopt, path, os.environ)
            zero = run(pat, kopt, path, ZERO)
            try:
                orc = O.search(pat, text.encode("latin-1"), kopt)
                O.set_compat(True)
                orc_dep = O.search(pat, text.encode("latin-1"), kopt)
                O.set_compat(False)
            except Exception:
                O.set_compat(False)
                continue
            stats["cases"] += 1
            if emit and members and len(text) < 9000 and ((dep != zero and len(vectors) - nsame < 60) or (dep == zero and len(members) >= 11 and nsame < 30)):
                nsame += 1 if dep == zero else 0
                vectors.append({"pattern": pat, "kopt": kopt, "text": text, "deployed": dep, "zero_fill": zero})
            m_ = len(members) if members else 18
            bm = stats["by_m"].setdefault(str(m_), [0, 0])
            bm[0] += 1
            if orc != zero:
                stats["oracle_differs_from_zero_fill"] += 1
            if orc_dep != dep:
                stats["oracle_compat_differs_from_deployed"] = stats.get("oracle_compat_differs_from_deployed", 0) + 1
                if len(stats.setdefault("compat_misses", [])) < 6:
                    stats["compat_misses"].append({"pattern": pat, "k": kopt, "m": len(members) if members else 18})
            if dep != zero:
                stats["deployed_differs_from_zero_fill"] += 1
                bm[1] += 1
                if len(dep) == len(zero):
                    stats["same_hit_count_when_different"] += 1
                if len(examples) < 8:
                    d = [(a, b) for a, b in zip(dep, zero) if a != b][:2]
                    examples.append({"pattern": pat, "k": kopt, "hits_deployed": len(dep), "hits_zero_fill": len(zero), "first_differences(deployed, zero_fill)": d})
    stats["examples"] = examples
    if emit:
        json.dump({"how": "tools/deployed_gap.py --emit: stock nrgrep_coords -i -b 1600000 with the default glibc allocator (deployed) and with zero-filled malloc (zero_fill)",
                   "cases": vectors}, open(emit, "w"))
    print(json.dumps(stats, indent=1))


main()
