"""ncu `--metrics gpu__time_duration.sum --csv` launch list -> per-kernel share table of the SECOND half of the launches
(the first half is the warm-up step).  usage: launch_shares.py in.csv out.txt "description" [--all]"""
import collections
import csv
import sys

src, dst, what = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src, errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
rows = rows[hdr + 1:]
if "--all" not in sys.argv:
    rows = rows[len(rows) // 2:]
acc, cnt = collections.Counter(), collections.Counter()
for r in rows:
    try:
        v = float(r[mv].replace(",", ""))
    except Exception:
        continue
    if r[mu] == "us":
        v *= 1e3
    elif r[mu] == "ms":
        v *= 1e6
    name = r[kn].split("(")[0][:80]
    acc[name] += v
    cnt[name] += 1
tot = sum(acc.values())
with open(dst, "w") as fh:
    fh.write(f"total {tot / 1e6:.2f} ms over {sum(cnt.values())} launches ({what}; ncu serialises the launches)\n")
    for k, v in acc.most_common(40):
        fh.write(f"{v / tot * 100:6.2f}%  {v / 1e6:9.3f} ms  {cnt[k]:5d}x  {k}\n")
print(open(dst).read())
