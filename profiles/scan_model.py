"""Offline model of the density scan's lane utilisation (CPU, numpy; no GPU needed).

For the dam-break workload it computes, per particle, how many candidates of its cell's walk-order sequence are examined
before the 32nd hit (the scan length), then counts warp-steps of the lane-per-particle scan under different assignments of
particles to lanes.  A warp-step = one candidate tested by every still-active lane of the warp; lanes of one cell walk
the same sequence, lanes of different cells pay max(segment length) per segment.

    python profiles/scan_model.py [n_particles] > profiles/r1c/scan_model.txt
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
from cuda_sph_b200 import workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
params, st = workloads.dam_break(n, 2.5, seed=0)
pos = st.position.astype(np.float32).astype(np.float64)
vox = 2.0
dims = np.ceil(np.asarray(params.space_size) / vox).astype(np.int64)
c = (pos / vox).astype(np.int64)
key = c[:, 0] + c[:, 1] * dims[0] + c[:, 2] * dims[0] * dims[1]
order = np.lexsort((np.arange(n), key))
sk, sp = key[order], pos[order]
ncells = int(np.prod(dims))
begin = np.searchsorted(sk, np.arange(ncells + 1))
offs = [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)]   # walk order

scan = np.zeros(n, np.int64)        # candidates examined (sorted order index)
seg_stop = np.zeros(n, np.int64)    # segment in which the particle stops
cell_segs = {}                      # cell -> 27 segment lengths
for k in np.unique(sk):
    cz, rem = divmod(int(k), int(dims[0] * dims[1]))
    cy, cx = divmod(rem, int(dims[0]))
    segs, cand = [], []
    for dx, dy, dz in offs:
        x, y, z = cx + dx, cy + dy, cz + dz
        if 0 <= x < dims[0] and 0 <= y < dims[1] and 0 <= z < dims[2]:
            nk = x + y * dims[0] + z * dims[0] * dims[1]
            b, e = begin[nk], begin[nk + 1]
        else:
            b = e = 0
        segs.append(e - b)
        cand.append(np.arange(b, e))
    cell_segs[int(k)] = np.asarray(segs)
    cand = np.concatenate(cand)
    mine = np.arange(begin[k], begin[k + 1])
    d2 = ((sp[mine][:, None, :] - sp[cand][None, :, :]) ** 2).sum(-1)
    hit = np.cumsum(d2 <= 4.0, axis=1)
    full = hit[:, -1] >= 32
    first32 = np.where(full, (hit >= 32).argmax(axis=1) + 1, len(cand))
    scan[mine] = first32
    seg_stop[mine] = np.searchsorted(np.cumsum(segs), first32, side="left")


def warp_steps(lanes):
    """lanes: sorted indices handled by one warp.  Per segment the warp pays the max over its active lanes."""
    total = 0
    cells = sk[lanes]
    for s in range(27):
        m = 0
        for t, k in zip(lanes, cells):
            segs = cell_segs[int(k)]
            before = int(segs[:s].sum())
            if scan[t] <= before:
                continue
            m = max(m, min(int(segs[s]), int(scan[t]) - before))
        total += m
    return total


rng = np.random.default_rng(0)
tiles = rng.choice(n // 128, size=min(400, n // 128), replace=False)
res = {"sorted order (current)": 0, "tile sorted by scan length": 0, "per-cell fast/slow halves": 0,
       "flat iterator, sorted order": 0, "flat iterator, tile sorted by scan length": 0,
       "flat iterator, 2 particles per lane (long + short)": 0, "flat iterator, tile sorted by fractional x": 0,
       "flat iterator, 2 per lane paired by fractional x": 0, "ideal (mean)": 0}
for tb in tiles:
    t0 = tb * 128
    idx = np.arange(t0, t0 + 128)
    res["ideal (mean)"] += scan[idx].sum() / 32.0
    res["sorted order (current)"] += sum(warp_steps(idx[w * 32:(w + 1) * 32]) for w in range(4))
    by_len = idx[np.argsort(scan[idx], kind="stable")]
    res["tile sorted by scan length"] += sum(warp_steps(by_len[w * 32:(w + 1) * 32]) for w in range(4))
    fast, slow = [], []
    for k in np.unique(sk[idx]):
        m = idx[sk[idx] == k]
        m = m[np.argsort(scan[m], kind="stable")]
        fast += list(m[:len(m) // 2])
        slow += list(m[len(m) // 2:])
    # flat iterator: every lane walks its own candidate stream, a warp pays max over its lanes (no per-segment max)
    res["flat iterator, sorted order"] += sum(scan[idx[w * 32:(w + 1) * 32]].max() for w in range(4))
    res["flat iterator, tile sorted by scan length"] += sum(scan[by_len[w * 32:(w + 1) * 32]].max() for w in range(4))
    pair = scan[by_len][:64] + scan[by_len][::-1][:64]      # shortest with longest on one lane, two warps of 32 lanes
    res["flat iterator, 2 particles per lane (long + short)"] += pair[:32].max() + pair[32:].max()
    fx = sp[idx, 0] / vox - np.floor(sp[idx, 0] / vox)       # the cheap predictor: dx is the outermost walk dimension
    by_fx = idx[np.argsort(fx, kind="stable")]
    res["flat iterator, tile sorted by fractional x"] += sum(scan[by_fx[w * 32:(w + 1) * 32]].max() for w in range(4))
    pair = scan[by_fx][:64] + scan[by_fx][::-1][:64]
    res["flat iterator, 2 per lane paired by fractional x"] += pair[:32].max() + pair[32:].max()
    halves = np.asarray(fast + slow)
    res["per-cell fast/slow halves"] += sum(warp_steps(halves[w * 32:(w + 1) * 32]) for w in range(4))
print(f"dam-break, N={n}: scan length mean {scan.mean():.0f}, median {np.median(scan):.0f}, p90 "
      f"{np.percentile(scan, 90):.0f}, max {scan.max()}; capped lists {np.mean(scan < 10**9):.0%}")
base = res["sorted order (current)"]
for k_, v in res.items():
    print(f"{k_:32s} warp-steps per tile {v / len(tiles):8.0f}   vs current {v / base:5.2f}   lane utilisation "
          f"{res['ideal (mean)'] / v:5.2f}")
