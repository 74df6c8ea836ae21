"""Group an `ncu --page source --csv` dump into contiguous SASS regions with similar execution counts."""
import csv
import sys

path = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
ia, ie, it, iss = (hdr.index(k) for k in ("Source", "Instructions Executed", "Avg. Threads Executed",
                                          "Warp Stall Sampling (All Samples)"))
data = [r for r in rows if len(r) == len(hdr) and r[ie].isdigit()]
# the dump repeats the kernel once per captured launch: keep the first copy
first = data[0][ia]
for k in range(1, len(data)):
    if data[k][ia] == first and k > 50:
        data = data[:k]
        break
tot = sum(int(r[ie]) for r in data)
stot = sum(int(r[iss]) for r in data)
print(f"warp instr {tot}  stall samples {stot}  sass lines {len(data)}")
start, prev = 0, None
for k, r in enumerate(data + [None]):
    e = int(r[ie]) if r else -1
    if prev is not None and (e > prev * 1.3 or e < prev / 1.3):
        seg = data[start:k]
        se = sum(int(x[ie]) for x in seg)
        ss = sum(int(x[iss]) for x in seg)
        if 100 * se / tot >= minpct or 100 * ss / max(stot, 1) >= minpct:
            thr = sum(float(x[it]) * int(x[ie]) for x in seg) / max(se, 1)
            print(f"{start:4d}-{k-1:4d} n={k-start:3d} exec {100*se/tot:5.1f}% stall {100*ss/max(stot,1):5.1f}% "
                  f"x{int(seg[0][ie])/1e6:6.2f}M thr {thr:4.1f}  {seg[0][ia][:56]}")
        start = k
    prev = e
