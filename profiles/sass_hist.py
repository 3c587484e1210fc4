"""Aggregate an `ncu --page source --csv` dump: executed warp-instructions by SASS opcode."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r][0]
h = rows[hi]
ci, cs, cn = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
agg = collections.Counter()
smp = collections.Counter()
tot = 0
for r in rows[hi + 1:]:
    if len(r) <= ci or not r[ci].isdigit():
        continue
    toks = r[cs].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    agg[op] += int(r[ci])
    smp[op] += int(r[cn]) if r[cn].isdigit() else 0
    tot += int(r[ci])
print("total warp-instructions", tot)
ts = sum(smp.values())
for op, c in agg.most_common(28):
    print(f"{op:10s} {c:14d} {100 * c / tot:6.2f}%   samples {100 * smp[op] / max(ts, 1):6.2f}%")
