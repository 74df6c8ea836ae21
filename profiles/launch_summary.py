"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, total and SHARE."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
i = next(k for k, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[i]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[i + 1:]:
    if len(r) < len(hdr):
        continue
    v = float(r[mv].replace(",", ""))
    v = v / 1000 if r[mu] == "ns" else v * 1000 if r[mu] == "ms" else v
    agg.setdefault(r[kn].split("(")[0][:48], []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':48s} {'n':>4s} {'mean_us':>9s} {'total_us':>10s} {'share':>6s}")
for k, v in agg.items():
    print(f"{k:48s} {len(v):4d} {sum(v)/len(v):9.1f} {sum(v):10.1f} {100*sum(v)/tot:5.1f}%")
