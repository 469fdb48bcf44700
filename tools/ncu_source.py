#!/usr/bin/env python
"""Per-region instruction and stall-sample totals from `ncu -i X.ncu-rep --page source --csv` (SASS view).
Regions are split at instructions executed a markedly different number of times (loop nests)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iex]) for r in rows[2:] if len(r) > iex)
tots = sum(int(r[ismp]) for r in rows[2:] if len(r) > iex)
print("total inst", tot, "samples", tots)
ops = collections.Counter(); smp = collections.Counter()
for r in rows[2:]:
    if len(r) <= iex: continue
    op = r[isrc].split()
    op = op[1] if op[0].startswith("@") else op[0]
    op = op.split(".")[0]
    ops[op] += int(r[iex]); smp[op] += int(r[ismp])
for op, v in ops.most_common(18):
    print("%-10s %6.2f%% inst  %6.2f%% samples" % (op, 100.0 * v / tot, 100.0 * smp[op] / max(tots, 1)))
# regions by execution count
print("--- regions (consecutive instructions with similar execution counts)")
cur = None; acc = 0; n = 0; first = 0; sm = 0
def flush():
    if n: print("  insts %5d..%5d  n=%5d  exec/inst=%10d  share=%5.1f%%  samples=%5.1f%%" % (first, first + n - 1, n, acc // n, 100.0 * acc / tot, 100.0 * sm / max(tots, 1)))
for i, r in enumerate(rows[2:]):
    if len(r) <= iex: continue
    e = int(r[iex])
    if cur is None or not (0.7 * cur <= e <= 1.4 * cur):
        flush(); cur = max(e, 1); acc = 0; n = 0; first = i; sm = 0
    acc += e; n += 1; sm += int(r[ismp])
flush()
