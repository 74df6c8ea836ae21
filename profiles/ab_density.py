"""A/B of the density sweeps (SPH_DENSITY=flat|rows): per-stage CUDA-event times of steps 1..3 from the start state and
a bitwise comparison of the results.  Not a bench line -- the table goes to profiles/.

    python profiles/ab_density.py [--cases dam1m,box8_4m,box2.5_1m,box25_1m]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cuda_sph_b200 import B200SPHStrategy, SphConstants, workloads  # noqa: E402

CASES = {
    "dam1m": lambda: workloads.dam_break(1 << 20, 2.5, seed=0),
    "box8_4m": lambda: workloads.uniform_box(1 << 22, 8.0, seed=0),
    "box2.5_1m": lambda: workloads.uniform_box(1 << 20, 2.5, seed=0),
    "box25_1m": lambda: workloads.uniform_box(1 << 20, 25.0, seed=0),
    "dam64k": lambda: workloads.dam_break(1 << 16, 2.5, seed=0),
}
ap = argparse.ArgumentParser()
ap.add_argument("--cases", default="dam1m,box8_4m,box2.5_1m,box25_1m")
ap.add_argument("--variants", default="flat,rows")
a = ap.parse_args()
for name in a.cases.split(","):
    params, st = CASES[name]()
    res = {}
    for var in a.variants.split(","):
        os.environ["SPH_DENSITY"] = var
        s = B200SPHStrategy(params, SphConstants(mode="BOX"), record_neighbour_counts=True)
        s.upload(st)
        s.save_state()
        nsteps = 4
        best = [None] * nsteps
        for rep in range(4):
            s.restore_state()
            for k in range(nsteps):
                t = s.step_timed(1)
                if rep and (best[k] is None or t["density_ms"] < best[k]["density_ms"]):
                    best[k] = t
        for k in range(nsteps):
            b = best[k]
            print(f"{name:10s} {var:5s} step {k + 1}: density {b['density_ms']:7.4f} force {b['force_ms']:7.4f} "
                  f"sort {b['sort_ms']:7.4f} reorder {b['reorder_ms']:7.4f} hash {b['hash_ms']:7.4f} "
                  f"total {b['total_ms']:7.4f} ms", flush=True)
        out = s.download(np.float32)
        res[var] = (out, s.neighbour_counts())
        s.close()
    if len(res) == 2:
        (oa, ca), (ob, cb) = res.values()
        same = all(np.array_equal(x.view(np.uint32), y.view(np.uint32)) for x, y in
                   ((oa.position, ob.position), (oa.velocity, ob.velocity), (oa.density, ob.density)))
        print(f"{name:10s} bitwise equal after 4 steps: {same}; counts equal: {np.array_equal(ca, cb)}", flush=True)
