"""Seeded synthetic inputs shared by the tests (and sized-down versions of BASELINE.json's configs)."""
import random

import numpy as np

DNA, PEP = "ACGT", "ACDEFGHIKLMNPQRSTVWY"


def random_pattern(rng, alpha, m, cls_pct=0.18, dot_pct=0.07, neg_pct=0.15):
    pat, members = "(", []
    for _ in range(m):
        r = rng.random()
        if r < dot_pct:
            pat += "."
            members.append(list(alpha))
        elif r < dot_pct + cls_pct:
            chars = rng.sample(alpha, rng.randint(2, 3))
            if rng.random() < neg_pct:
                pat += "[^" + "".join(chars) + "]"
                members.append([c for c in alpha if c not in chars])
            else:
                pat += "[" + "".join(chars) + "]"
                members.append(chars)
        else:
            c = rng.choice(alpha)
            pat += c
            members.append([c])
    return pat + ")", members


def planted(rng, members, alpha, k):
    s = [rng.choice(c) for c in members]
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


def random_text(rng, members, alpha, k, nrec=3, lo=80, hi=600, plant=0.3):
    lines = []
    for r in range(nrec):
        lines.append(">seq%d synthetic record" % r)
        t, target = "", rng.randint(lo, hi)
        while len(t) < target:
            t += planted(rng, members, alpha, k) if rng.random() < plant else "".join(
                rng.choice(alpha) for _ in range(rng.randint(1, 14)))
        if rng.random() < 0.3:
            t = "".join(ch.lower() if rng.random() < 0.3 else ch for ch in t)
        if rng.random() < 0.2:
            p = rng.randrange(len(t))
            t = t[:p] + "NNNN" + t[p:]
        lines.append(t)
    return "\n".join(lines) + "\n"


def genome(seed, nchrom, total, alphabet=b"ACGT", name="chr"):
    """FASTA bytes with one sequence per line (the reference's .seq layout), numpy-generated."""
    rng = np.random.default_rng(seed)
    parts = []
    per = total // nchrom
    lut = np.frombuffer(alphabet, dtype=np.uint8)
    for c in range(nchrom):
        parts.append((">%s%d synthetic\n" % (name, c + 1)).encode())
        parts.append(lut[rng.integers(0, len(lut), size=per, dtype=np.uint8)].tobytes())
        parts.append(b"\n")
    return b"".join(parts)
