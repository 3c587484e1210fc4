"""Per-SASS-instruction executed counts from `ncu --page source --csv`: prints the instructions sorted by address with
executed count (in millions) and stall samples, only those above a threshold, plus the total."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
h = rows[1]
ia, isrc, iex, ismp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
tot = 0
out = []
for r in rows[2:]:
    try:
        ex = int(r[iex])
    except ValueError:
        continue
    tot += ex
    out.append((r[ia], ex, int(r[ismp] or 0), r[isrc]))
print("total executed (M):", tot / 1e6)
for a, ex, s, src in out:
    if ex / 1e6 >= thr:
        print(f"{a[-5:]} {ex/1e6:9.2f} {s:6d}  {src[:110]}")
