"""Per CUDA-source-line profile of one kernel from an .ncu-rep (needs -lineinfo and --import-source on):
    python profiles/line_profile.py rep.ncu-rep kernel_regex [min_pct]
Prints, for every source line above min_pct of the stall samples or executed instructions: file:line, share of warp
stall samples (~ share of warp time), share of executed warp instructions, the top stall reasons and the source text."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg, order, text = {}, [], {}
cur_file, cur_line, hdr, seen, launches = None, None, None, set(), 0


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        if hdr is None:
            hdr = r
            i_addr, i_s, i_e = hdr.index("Address"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
            stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":
        cur_line = (cur_file, int(r[0]))
        text.setdefault(cur_line, r[1])
        continue
    key = (cur_line, r[i_addr])
    if key in seen:
        continue
    seen.add(key)
    a = agg.setdefault(cur_line, [0, 0, [0] * len(stall_cols)])
    if cur_line not in order:
        order.append(cur_line)
    a[0] += num(r[i_s])
    a[1] += num(r[i_e])
    for k, c in enumerate(stall_cols):
        a[2][k] += num(r[c])
ts = sum(a[0] for a in agg.values()) or 1
te = sum(a[1] for a in agg.values()) or 1
print(f"kernel {kern}: stall samples {ts}, warp instructions {te}")
for ln in sorted(order):
    s, e, st = agg[ln]
    if 100 * s / ts < minpct and 100 * e / te < minpct:
        continue
    top = sorted(zip(st, (hdr[c][6:] for c in stall_cols)), reverse=True)[:3]
    tt = sum(st) or 1
    print(f"{ln[0]:20s}:{ln[1]:4d} time {100*s/ts:5.1f}% instr {100*e/te:5.1f}%  "
          + " ".join(f"{n}={100*v/tt:.0f}%" for v, n in top) + f"  | {text.get(ln, '').strip()[:70]}")
