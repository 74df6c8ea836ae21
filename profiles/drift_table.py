"""Short-horizon trajectory drift of the fp32 engine against the fp64 oracle (BASELINE north_star: "short-horizon
trajectory drift reported"): per step max / median |dx| / h over particles finite in both, and the non-finite counts of
both sides.  One JSON object per workload on stdout.

    python profiles/drift_table.py > profiles/r2/drift.jsonl
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cuda_sph_b200 import B200SPHStrategy, SphConstants, workloads  # noqa: E402
from oracle import oracle as orc  # noqa: E402

orc.set_exact_pow(False)
for name, (params, st) in {"dam20k": workloads.dam_break(20000, 2.5, seed=7),
                           "box8_30k": workloads.uniform_box(30000, 8.0, seed=8)}.items():
    n = len(st.position)
    s = B200SPHStrategy(params, SphConstants(mode="BOX"))
    P = orc.OracleParams(n=n, space=tuple(params.space_size), dt=1 / params.fps)
    s.upload(st)
    pos, vel = st.position, st.velocity
    rows = []
    for step in range(1, 101):
        s.step(1)
        r = orc.step(P, pos, vel, light=True)
        pos, vel = r.position, r.velocity
        if step <= 10 or step % 10 == 0:
            got = s.download()
            fo, fe = np.isfinite(pos).all(axis=1), np.isfinite(got.position).all(axis=1)
            both = fo & fe
            d = np.linalg.norm(got.position[both] - pos[both], axis=1) / 2.0
            rows.append({"step": step, "max_dx_over_h": float(d.max()) if len(d) else None,
                         "median_dx_over_h": float(np.median(d)) if len(d) else None,
                         "frac_below_1e-3": float((d < 1e-3).mean()) if len(d) else None,
                         "nonfinite_engine": int((~fe).sum()), "nonfinite_oracle": int((~fo).sum())})
    print(json.dumps({"workload": name, "n": n, "rows": rows}), flush=True)
    s.close()
