"""Summarise an `ncu --page source --csv` dump: hottest SASS lines by executed warp instructions / stall samples."""
import csv
import sys

path, thresh = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
ia, ie, it, iss = (hdr.index(k) for k in ("Source", "Instructions Executed", "Avg. Threads Executed",
                                          "Warp Stall Sampling (All Samples)"))
data = [r for r in rows if len(r) == len(hdr) and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in data)
stot = sum(int(r[iss]) for r in data)
print(f"total warp instr {tot}  stall samples {stot}  sass lines {len(data)}")
for k, r in enumerate(data):
    e, s = int(r[ie]), int(r[iss])
    if e > thresh * tot or s > thresh * stot:
        print(f"{k:4d} {r[ia][:64]:64s} exec {100*e/tot:5.2f}%  thr {float(r[it]):5.1f}  stall {100*s/max(stot,1):5.2f}%")
