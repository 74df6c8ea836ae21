"""BASELINE configs[4]: single-step density + force microbench, uniform random particles, N = 2^16 .. 2^26 at 2.5 and 25
particles per cell.  Per-stage CUDA-event times of the first step from the start state
(eager launches, best of 3, state restored in between); not a bench line -- the table goes to profiles/.

    python profiles/microbench_sweep.py [--max-log2 26] > profiles/r1c/microbench_sweep.txt
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cuda_sph_b200 import B200SPHStrategy, SphConstants, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--max-log2", type=int, default=24)
a = ap.parse_args()
print(f"{'N':>10s} {'ppc':>5s} {'hash':>7s} {'sort':>7s} {'reord':>7s} {'density':>8s} {'force':>8s} {'step_ms':>8s} "
      f"{'Mupd/s':>9s} {'dens GB/s':>9s} {'force GB/s':>10s}")
for ppc in (2.5, 25.0):
    for lg in range(16, a.max_log2 + 1, 2):
        n = 1 << lg
        params, st = workloads.uniform_box(n, ppc, seed=lg)
        s = B200SPHStrategy(params, SphConstants(mode="BOX"))
        s.upload(st)
        s.save_state()
        best = None
        for rep in range(4):
            s.restore_state()
            t = s.step_timed(1)
            if rep and (best is None or t["total_ms"] < best["total_ms"]):
                best = t
        s.close()
        row = {"n": n, "ppc": ppc, **{k: round(v, 4) for k, v in best.items() if k.endswith("_ms")}}
        print(f"{n:10d} {ppc:5.1f} {best['hash_ms']:7.3f} {best['sort_ms']:7.3f} {best['reorder_ms']:7.3f} "
              f"{best['density_ms']:8.3f} {best['force_ms']:8.3f} {best['total_ms']:8.3f} "
              f"{n / best['total_ms'] / 1e3:9.1f} {20 * n / best['density_ms'] / 1e6:9.1f} "
              f"{84 * n / best['force_ms'] / 1e6:10.1f}", flush=True)
        print("#", json.dumps(row), flush=True)
