"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    name = r[ki].split("(")[0][-60:]
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    a = agg.setdefault(name, [0, 0.0, 1e30, 0.0])
    a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
tot = sum(v[1] for v in agg.values())
print(f"{'total us':>10} {'share':>6} {'n':>5} {'avg us':>9} {'min':>8} {'max':>8}  kernel")
for n, (c, t, mn, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} {100 * t / tot:5.1f}% {c:5d} {t / c:9.2f} {mn:8.2f} {mx:8.2f}  {n}")
print(f"{tot:10.1f} total (cold-cache, serialised launches: compare shares, not absolutes)")
