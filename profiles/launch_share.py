"""Per-kernel share of the step from an ncu launch list (`--metrics gpu__time_duration.sum --csv`)."""
import csv
import collections
import re
import sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    name = re.sub(r"^void ", "", name)[:70]
    v = float(r[-1].replace(",", ""))
    unit = r[-2]
    ns = v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
    tot[name] += ns
    cnt[name] += 1
all_ns = sum(tot.values())
print(f"{len(rows)} launches, {all_ns/1e6:.3f} ms total (cold-cache, serialised: shares, not absolutes)")
for name, ns in sorted(tot.items(), key=lambda kv: -kv[1])[:25]:
    print(f"{100*ns/all_ns:6.2f} %  {ns/1e6:9.3f} ms  x{cnt[name]:<4d} {name}")
