"""Per-step stage times and state statistics of the pipe4m workload (steps 1..K from the start state).

    python profiles/pipe_steps.py [--n 4194304] [--steps 4]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cuda_sph_b200 import B200SPHStrategy, SphConstants, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 22)
ap.add_argument("--steps", type=int, default=4)
a = ap.parse_args()
params, st = workloads.pipe_flow(a.n, seed=0)
s = B200SPHStrategy(params, SphConstants(mode="PIPE"), record_neighbour_counts=True)
s.upload(st)
s.save_state()
for rep in range(2):
    s.restore_state()
    for k in range(a.steps):
        t = s.step_timed(1)
        if rep:
            stt = s.stats()
            cnt = s.neighbour_counts()
            print(f"step {k + 1}: density {t['density_ms']:9.3f} force {t['force_ms']:9.3f} sort {t['sort_ms']:7.3f} "
                  f"reorder {t['reorder_ms']:7.3f} total {t['total_ms']:9.3f} ms | dead {stt['n_dead']} nonfinite "
                  f"{stt['n_nonfinite']} max_rho {stt['max_density']:.3g} | mean count {cnt.mean():.2f} capped "
                  f"{(cnt >= 32).mean():.3f}", flush=True)
s.close()
