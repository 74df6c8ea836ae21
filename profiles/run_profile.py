"""Small driver for ncu captures: a few device-resident steps of a bench workload (no timing claims made here)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402
from cuda_sph_b200 import B200SPHStrategy, SphConstants  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="dam1m")
ap.add_argument("--particles", type=int, default=None)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--ppc", type=float, default=None, help="uniform box at this many particles per cell instead of --workload")
a = ap.parse_args()
if a.ppc:
    from cuda_sph_b200 import workloads
    n = a.particles or (1 << 20)
    params, st = workloads.uniform_box(n, a.ppc, seed=0)
    mode, desc = "BOX", f"uniform box {a.ppc}/cell N={n}"
else:
    params, st, mode, desc = bench.make_workload(a.workload, a.particles)
s = B200SPHStrategy(params, SphConstants(mode=mode), use_graph=a.graph)
s.upload(st)
s.step(a.steps)
s.synchronize()
print(desc, s.stats())
