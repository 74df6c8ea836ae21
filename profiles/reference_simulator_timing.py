"""the REFERENCE'S OWN Voxel step (unmodified kernels of /root/reference, loaded through
oracle/ref_shim.py) timed under numba's CUDA simulator, next to the C oracle port on the same input and host.

Only runnable where /root/reference exists (the build container: no GPU, so the simulator is the only way the reference's
kernels execute at all); the result is committed as profiles/r2/reference_simulator_timing.json.  The simulator runs one
Python thread per CUDA thread, so the figure is a statement about the reference's CPU-runnable path (BASELINE configs[0]),
not about its numba-CUDA kernels on a GPU.

    python profiles/reference_simulator_timing.py [--n 4096] [--steps 2]
"""
import argparse
import json
import os
import platform
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import oracle as orc  # noqa: E402
from oracle.ref_shim import load_reference, reference_available  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
if not reference_available():
    raise SystemExit("reference tree not mounted")
warnings.filterwarnings("ignore")
n = a.n
# an interior blob at the reference's own start density (config.py:84-87: N particles in the first 10 % of the box):
# every particle keeps its 27 cells inside the domain, which the unmodified neighbour kernel needs (ref_shim.py item 4)
ref = load_reference(n, "BOX")
space = np.asarray(ref.config.params.space_size, np.float64)
rng = np.random.default_rng(0)
side = (n / 25.0) ** (1.0 / 3.0) * 2.0          # ~25 particles per cell of edge 2
lo = np.maximum(space / 2 - side / 2, 4.0)
pos = (lo + rng.random((n, 3)) * side).astype(np.float32).astype(np.float64)
vel = (np.array([1.5, -5.0, -5.0]) + rng.uniform(-0.5, 0.5, (n, 3))).astype(np.float32).astype(np.float64)
strat = ref.VoxelStrategy(ref.config.params)
st = ref.data_classes.SimulationState(pos.copy(), vel.copy(), np.zeros(n))
t_ref = []
for k in range(a.steps):
    t0 = time.perf_counter()
    st = strat.compute_next_state(st)
    t_ref.append(time.perf_counter() - t0)
# the C oracle port, same input, all host threads (what bench.py --impl reference times on the GPU box)
orc.set_num_threads(os.cpu_count())
p, v = pos.copy(), vel.copy()
t_orc = []
for k in range(a.steps):
    t0 = time.perf_counter()
    r = orc.step(orc.OracleParams(n=n), p, v)
    t_orc.append(time.perf_counter() - t0)
    p, v = r.position, r.velocity
same_bits = bool(np.array_equal(np.asarray(st.position).view(np.uint64), np.asarray(p).view(np.uint64)))
out = {"what": "reference VoxelSPHStrategy.compute_next_state under numba's CUDA simulator (unmodified kernels)",
       "n": n, "steps": a.steps, "seconds_per_step": t_ref, "particle_updates_per_s": n / float(np.mean(t_ref)),
       "configs0_extrapolated_s": 100 * float(np.mean(t_ref)) * 4096 / n,
       "oracle_port_seconds_per_step": t_orc, "oracle_port_particle_updates_per_s": n / float(np.min(t_orc)),
       "positions_bitwise_equal_to_oracle_after_steps": same_bits,
       "host": {"cpu": platform.processor() or platform.machine(), "cores": os.cpu_count()},
       "numba": __import__("numba").__version__}
print(json.dumps(out))
