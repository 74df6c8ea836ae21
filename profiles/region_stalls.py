"""Stall-reason breakdown of SASS line ranges of an `ncu --page source --csv` dump:  region_stalls.py dump.csv 0-100 101-200 ..."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
ia, ie = hdr.index("Source"), hdr.index("Instructions Executed")
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows if len(r) == len(hdr) and r[ie].isdigit()]
first = data[0][ia]
for k in range(1, len(data)):
    if data[k][ia] == first and k > 50:
        data = data[:k]
        break
tot_all = sum(int(r[c] or 0) for r in data for c in cols)
for rg in sys.argv[2:]:
    lo, hi = map(int, rg.split("-"))
    seg = data[lo:hi + 1]
    ex = sum(int(r[ie]) for r in seg)
    s = {hdr[c]: sum(int(r[c] or 0) for r in seg) for c in cols}
    tot = sum(s.values())
    top = sorted(s.items(), key=lambda kv: -kv[1])[:7]
    print(f"{rg:11s} instr {ex/1e6:8.2f}M samples {100*tot/max(tot_all,1):5.1f}%: " +
          " ".join(f"{k[6:]}={100*v/max(tot,1):.0f}%" for k, v in top))
