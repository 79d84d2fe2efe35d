"""Aggregate an ncu gpu__time_duration launch list (csv) into per-kernel totals / shares."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ix = {n: i for i, n in enumerate(h)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    name = r[ix["Kernel Name"]].split("(")[0]
    try:
        v = float(r[ix["Metric Value"]])
    except ValueError:
        continue
    u = r[ix["Metric Unit"]]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# per-kernel totals from {sys.argv[1]} (cold-cache, serialised: compare SHARES); total {tot / 1e3:.2f} ms")
print(f'{"kernel":58s} {"launches":>8s} {"total_us":>12s} {"share":>7s}')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:58s} {n:8d} {t:12.1f} {100 * t / tot:6.1f}%")
