"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN KERNELS (numba CUDA simulator, oracle/ref_shim.py).

Run in the build container only (needs /root/reference):  python tests/golden/generate_golden.py
The outputs are committed; tests compare the C oracle (and, on the GPU box, the CUDA engine) against them.

Inputs are fp32-representable (generated in fp32, up-cast to fp64) so the fp32 engine and the fp64 reference see
identical values.  All BOX cases keep every particle >= 1 cell away from the domain boundary
because the reference reads uninitialised neighbour-cell ids there (voxel_kernels.py:44,64 -- SURVEY Q3).
"""
from __future__ import annotations

import os
import sys
import time
import warnings

import numpy as np

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle.ref_shim import load_reference  # noqa: E402


os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
from numba import cuda  # noqa: E402  (module-level name: the simulator swaps it inside kernels)

_get_neighbours = None
_get_index = None


def _count_kernel_py(cnt, lists, position, voxel_begin, vmap, voxel_size, space_dim):
    i = _get_index()
    if i >= position.shape[0]:
        return
    neighbours = cuda.local.array(32, np.int32)
    c = _get_neighbours(neighbours, i, position, voxel_size, space_dim, voxel_begin, vmap)
    cnt[i] = c
    for q in range(c):
        lists[i][q] = neighbours[q]


def _make_count_kernel(ref):
    global _get_neighbours, _get_index
    _get_neighbours = ref.voxel_kernels.get_neighbours
    _get_index = ref.util_kernels.get_index
    return cuda.jit(_count_kernel_py)


def _draw_py(out, st):
    import numba.cuda.random as nrandom
    i = cuda.grid(1)
    if i < out.shape[0]:
        for q in range(out.shape[1]):
            out[i][q] = nrandom.xoroshiro128p_uniform_float64(st, i)


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def run_voxel_step(ref, pos, vel):
    n = len(pos)
    st = ref.data_classes.SimulationState(pos.copy(), vel.copy(), np.zeros(n))
    s = ref.VoxelStrategy(ref.config.params)
    rng0 = np.array(s.rng_states.copy_to_host() if hasattr(s.rng_states, "copy_to_host") else s.rng_states)
    t = time.time()
    out = s.compute_next_state(st)
    took = time.time() - t
    # stage outputs, captured from the strategy's device arrays (they are plain numpy under the simulator)
    pressure = s.d_new_pressure_term.copy_to_host()
    viscosity = s.d_new_viscosity_term.copy_to_host()
    # neighbour counts through the reference's own get_neighbours (debug kernel defined in this file)
    cuda = ref.cuda
    count_kernel = _make_count_kernel(ref)
    cnt = cuda.to_device(np.zeros(n, np.int32))
    lists = cuda.to_device(np.full((n, 32), -1, np.int32))
    count_kernel[s.grid_size, s.block_size](cnt, lists, cuda.to_device(pos), s.d_voxel_begin,
                                            s.d_voxel_particle_map, s.d_voxel_size, s.d_space_dim)
    rng1 = np.array(s.rng_states.copy_to_host() if hasattr(s.rng_states, "copy_to_host") else s.rng_states)
    return dict(
        pos_in=pos, vel_in=vel, keys=np.asarray(s.voxels), map_ids=np.asarray(s.voxel_particle_map["particle_id"]),
        map_keys=np.asarray(s.voxel_particle_map["voxel_id"]), voxel_begin=np.asarray(s.voxel_begin),
        density=out.density, pressure=pressure, viscosity=viscosity, force=np.asarray(s.result_force),
        pos_out=out.position, vel_out=out.velocity, neigh_count=cnt.copy_to_host(), neighbours=lists.copy_to_host(),
        rng_in=rng0.view(np.uint64).reshape(-1, 2)[:n], rng_out=rng1.view(np.uint64).reshape(-1, 2)[:n],
        ref_seconds=np.float64(took), space=np.asarray(ref.config.params.space_size, np.float64),
        voxel=np.asarray(ref.config.params.voxel_size, np.float64),
        ext=np.asarray(ref.config.params.external_force, np.float64),
        pipe=np.asarray(ref.config.params.pipe.to_numpy(), np.float64), fps=np.float64(ref.config.params.fps),
    )


def case_kat4():
    """SURVEY.md Appendix A."""
    ref = load_reference(4, "BOX")
    pos = np.array([[11, 11, 11], [12, 11, 11], [11, 12.5, 11], [12.25, 12.5, 11.5]], dtype=np.float64)
    vel = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, -0.5, 0.25]], dtype=np.float64)
    return run_voxel_step(ref, pos, vel)


def case_box_dense(n=256, seed=1):
    """8 cells at ~32 particles/cell: nearly every list hits the 32 cap (traversal order decides the physics)."""
    rng = np.random.default_rng(seed)
    ref = load_reference(n, "BOX")
    pos = f32(rng.uniform(18.0, 22.0, (n, 3)))
    vel = f32(np.array([1.5, -5.0, -5.0]) + rng.uniform(-0.5, 0.5, (n, 3)))
    return run_voxel_step(ref, pos, vel)


def case_box_sparse(n=400, seed=2):
    """~1.9 particles/cell over 6^3 cells: uncapped lists, a few isolated particles (rho = 0 -> inf/NaN)."""
    rng = np.random.default_rng(seed)
    ref = load_reference(n, "BOX")
    pos = f32(rng.uniform(14.0, 26.0, (n, 3)))
    vel = f32(np.array([1.5, -5.0, -5.0]) + rng.uniform(-0.5, 0.5, (n, 3)))
    return run_voxel_step(ref, pos, vel)


def case_box_medium(n=600, seed=3):
    """4^3 cells at ~9 particles/cell: mixture of capped and uncapped lists."""
    rng = np.random.default_rng(seed)
    ref = load_reference(n, "BOX")
    pos = f32(rng.uniform(16.0, 24.0, (n, 3)))
    vel = f32(np.array([1.5, -5.0, -5.0]) + rng.uniform(-0.5, 0.5, (n, 3)))
    return run_voxel_step(ref, pos, vel)


def case_pipe_step(n=300, seed=4):
    """Full PIPE-mode step (config.py:68-76 pipe, 20x3x3 cells): interior cell row only; wall bounces on cylinder and
    cone segments, x<0 mirrors and x>=x_end recycles through xoroshiro128+."""
    rng = np.random.default_rng(seed)
    ref = load_reference(n, "PIPE")
    pos = np.empty((n, 3))
    pos[:, 0] = rng.uniform(2.0, 38.0, n)
    pos[:, 1:] = rng.uniform(2.0, 4.0, (n, 2))
    vel = rng.uniform(-25.0, 25.0, (n, 3))
    vel[: n // 10, 0] = rng.uniform(40.0, 900.0, n // 10)        # some leave through the outlet -> recycle
    vel[n // 10: n // 5, 0] = -rng.uniform(40.0, 900.0, n // 5 - n // 10)   # some leave through the inlet -> mirror
    return run_voxel_step(ref, f32(pos), f32(vel))


def case_pipe_collide(n=512, seed=5):
    """collision_kernel alone (base_kernels.py:56-72) on random points in and around the config pipe."""
    rng = np.random.default_rng(seed)
    ref = load_reference(n, "PIPE")
    cuda = ref.cuda
    pipe = np.asarray(ref.config.params.pipe.to_numpy(), np.float64)
    pos = np.empty((n, 3))
    pos[:, 0] = rng.uniform(-3.0, 43.0, n)
    pos[:, 1:] = rng.uniform(-0.5, 6.5, (n, 2))
    vel = rng.uniform(-10.0, 10.0, (n, 3))
    pos, vel = f32(pos), f32(vel)
    import numba.cuda.random as nrandom
    states = nrandom.create_xoroshiro128p_states(n, seed=16435234)
    rng_in = np.array(states.copy_to_host()).view(np.uint64).reshape(-1, 2).copy()
    d_pos, d_vel = cuda.to_device(pos.copy()), cuda.to_device(vel.copy())
    ref.base_kernels.collision_kernel[(n + 63) // 64, 64](d_pos, d_vel, cuda.to_device(pipe), states)
    rng_out = np.array(states.copy_to_host()).view(np.uint64).reshape(-1, 2).copy()
    return dict(pos_in=pos, vel_in=vel, pos_out=d_pos.copy_to_host(), vel_out=d_vel.copy_to_host(), pipe=pipe,
                rng_in=rng_in, rng_out=rng_out)


def case_rng(n=64):
    """xoroshiro128+ states and the first uniforms, from numba itself."""
    load_reference(4, "BOX")
    import numba.cuda.random as nrandom

    states = nrandom.create_xoroshiro128p_states(n, seed=16435234)
    st0 = np.array(states.copy_to_host()).view(np.uint64).reshape(-1, 2).copy()

    out = cuda.to_device(np.zeros((n, 4)))
    cuda.jit(_draw_py)[1, n](out, states)
    return dict(states=st0, uniforms=out.copy_to_host(), seed=np.uint64(16435234))


CASES = dict(kat4=case_kat4, box_dense=case_box_dense, box_sparse=case_box_sparse, box_medium=case_box_medium,
             pipe_step=case_pipe_step, pipe_collide=case_pipe_collide, rng=case_rng)

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for name in names:
        t = time.time()
        data = CASES[name]()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **data)
        print(f"{name}: {time.time() - t:.1f}s  ->  {name}.npz", flush=True)
