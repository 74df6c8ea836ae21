"""Small cases that drive every path of the sweeps (whole-tile plans, 32-particle passes, one-thread walk, PIPE epilogue)
for compute-sanitizer:  compute-sanitizer --tool memcheck|racecheck python profiles/sanitize_cases.py"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
from cuda_sph_b200 import B200SPHStrategy, SphConstants, workloads  # noqa: E402

cases = [("dam-break 25/cell", "BOX", workloads.dam_break(8192, 2.5, seed=0)),
         ("uniform 60/cell (32-particle passes)", "BOX", workloads.uniform_box(12000, 60.0, seed=1)),
         ("uniform 140/cell (walk)", "BOX", workloads.uniform_box(8000, 140.0, seed=2)),
         ("uniform 2.5/cell", "BOX", workloads.uniform_box(6000, 2.5, seed=3)),
         ("uniform 8/cell, several sort tiles", "BOX", workloads.uniform_box(20000, 8.0, seed=5)),
         ("pipe", "PIPE", workloads.pipe_flow(5000, seed=4))]
for name, mode, (params, st) in cases:
    s = B200SPHStrategy(params, SphConstants(mode=mode))
    s.upload(st)
    s.step(2)
    out = s.download()
    print(name, "ok, finite positions:", int(np.isfinite(out.position).all(axis=1).sum()), "of", len(out.position))
    s.close()
