import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
from cuda_sph_b200 import workloads
from cuda_sph_b200.data_classes import SimulationState
from oracle import oracle as orc
from tests.test_gpu_parity import _strategy
n = 30000
params, st = workloads.pipe_flow(n, seed=4)
rng = np.random.default_rng(5)
vel = rng.uniform(-30, 30, (n, 3)).astype(np.float32).astype(np.float64)
vel[: n // 20, 0] = 2000.0
vel[n // 20: n // 10, 0] = -2000.0
st = SimulationState(st.position, vel, st.density)
table = params.pipe.to_numpy()
s = _strategy(n, "PIPE", params.space_size, params.voxel_size, params.external_force, params.fps, table)
s.compute_next_state(st)
P = orc.OracleParams(n=n, mode="PIPE", space=tuple(params.space_size), ext=tuple(params.external_force), dt=1 / params.fps, pipe=table)
orc.set_exact_pow(False)
r = orc.step(P, st.position, st.velocity, rng=orc.rng_init(n), want_neighbours=True)
pr, vi = s.terms()
for name, mine, ref in [("visc", vi, r.viscosity), ("press", pr, r.pressure), ("force", s.result_force, r.force), ("vel", s.new_state.velocity, r.velocity), ("pos", s.new_state.position, r.position)]:
    fin = np.isfinite(ref).all(1)
    rel = np.linalg.norm(mine[fin] - ref[fin], axis=1) / np.maximum(np.linalg.norm(ref[fin], axis=1), 1e-300)
    idx = np.flatnonzero(fin)[np.argsort(rel)[-3:]]
    print(name, "max", rel.max(), "p99.9", np.quantile(rel, 0.999), "median", np.median(rel))
    for i in idx:
        print("   i", i, "mine", mine[i], "ref", ref[i], "rho", r.density[i], "cnt", r.neigh_count[i])
# scale of viscosity sums for worst particle
i = idx[-1]
