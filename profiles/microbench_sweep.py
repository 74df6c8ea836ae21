"""BASELINE configs[4]: single-step density + force microbench, uniform random particles, N = 2^16 .. 2^26 at 2.5 and 25
particles per cell.  Per-stage CUDA-event times of the first step from the start state
(eager launches, best of 3, state restored in between), and beside every size up to --cpu-max-log2 the fp64 CPU port
(oracle/sph_oracle.c, all host cores) on the same state; not a bench line -- the table goes to profiles/.

The reference's own numba kernels cannot be timed on the B200: /root/reference does not travel to the GPU box and its
sources may not be copied into this repository; in this container (no GPU) they run under numba's CUDA simulator at
~44 particle-updates/s.

    python profiles/microbench_sweep.py [--max-log2 26] > profiles/r2/microbench_sweep.txt
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cuda_sph_b200 import B200SPHStrategy, SphConstants, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--max-log2", type=int, default=26)
ap.add_argument("--cpu-max-log2", type=int, default=22)
a = ap.parse_args()
import time  # noqa: E402
from oracle import oracle as orc  # noqa: E402
orc.set_exact_pow(False)
orc.set_num_threads(os.cpu_count() or 1)
print(f"{'N':>10s} {'ppc':>5s} {'hash':>7s} {'sort':>7s} {'reord':>7s} {'density':>8s} {'force':>8s} {'step_ms':>8s} "
      f"{'Mupd/s':>9s} {'dens GB/s':>9s} {'force GB/s':>10s} {'cpu_ms':>9s} {'cpu Mupd/s':>10s} {'cores':>5s}")
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
        cpu_ms = None
        if lg <= a.cpu_max_log2:
            P = orc.OracleParams(n=n, space=tuple(params.space_size), dt=1 / params.fps)
            t0 = time.perf_counter()
            orc.step(P, st.position, st.velocity, light=True)
            cpu_ms = (time.perf_counter() - t0) * 1e3
        row = {"n": n, "ppc": ppc, **{k: round(v, 4) for k, v in best.items() if k.endswith("_ms")}, "cpu_ms": cpu_ms,
               "cpu_cores": orc.num_threads()}
        print(f"{n:10d} {ppc:5.1f} {best['hash_ms']:7.3f} {best['sort_ms']:7.3f} {best['reorder_ms']:7.3f} "
              f"{best['density_ms']:8.3f} {best['force_ms']:8.3f} {best['total_ms']:8.3f} "
              f"{n / best['total_ms'] / 1e3:9.1f} {20 * n / best['density_ms'] / 1e6:9.1f} "
              f"{84 * n / best['force_ms'] / 1e6:10.1f} "
              + (f"{cpu_ms:9.1f} {n / cpu_ms / 1e3:10.2f} {orc.num_threads():5d}" if cpu_ms else f"{'-':>9s} {'-':>10s} {'-':>5s}"),
              flush=True)
        print("#", json.dumps(row), flush=True)
