"""Stage times of the headline approximate search with the three candidate-filter modes."""
import sys, json, os
import torch
sys.path.insert(0, ".")
import patmatchdocker_b200 as pm
from bench import make_genome_torch, chrom_lengths, patterns

MODES = tuple(int(x) for x in os.environ.get('QF_MODES', '1,2').split(','))

def main():
    nb = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
    eng = pm.Engine(0)
    lengths = chrom_lengths(nb)
    g = make_genome_torch(lengths, list(range(len(lengths))), torch.device("cuda", 0))
    ds = eng.wrap_device(g.data_ptr(), g.numel())
    bp, bk = patterns()
    for pat, kopt in ((bp[0], bk), (bp[1], bk),
                      ("(GATAAGCC[AT]TTACGGA)", "2ids"), ("(TGA[GC]TCA...[AG][CT]GATAAG)", "2s"), ("(GAT..G[AC]CC[AT]TT)", "1ids"),
                      ("(CC[AT].....[AT]GG)", "1ids"), ("(GGTCA...TGACC)", "2ids"), ("([AG]GGTCA...TGACC[CT]GATTACA)", "3ids"),
                      ("(TATA[AT]A[AT][AG]....GC)", "1s"), ("(AC.TG.CA.GT.AC.TG.CA)", "2ids"), ("(TTGACA.................TATAAT)", "3ids")):
        for mode in MODES:
            eng.set_fused_filter(mode)
            for _ in range(3):
                n = eng.count(ds, pat, kopt)
            st = eng.stats()
            print(json.dumps({"pattern": pat, "k": kopt, "filter": mode, "hits": int(n), "scan_ms": round(st["scan_ms"], 3),
                              "total_ms": round(st["total_ms"], 3), "candidates": st["candidates"], "chunks": st["qgram_chunks"]}), flush=True)

main()
