"""Run configuration: the constants and factories of the reference's config.py:13-132, as functions instead of
import-time side effects, with SEEDED, vectorised start-state generators (the reference fills its start state with
unseeded per-particle Python loops, config.py:79-120)."""
from __future__ import annotations

import numpy as np

from .data_classes import Pipe, Segment, SimulationParameters, SimulationState
from .pipe_builder import PipeBuilder
from .strategy import SphConstants

SIM_MODE = "BOX"          # 'BOX' or 'PIPE'
SIM_STRATEGY = "B200"     # the reference knows 'NAIVE' and 'VOXEL'
DURATION = 15
FPS = 20
PARTICLE_COUNT = 20_000
MASS = 1.0
RHO_0 = 1.0
INF_R = 2.0
VISC = 0.5
K = 10.0
DAMP = 0.7
INF_R_2 = INF_R ** 2
INF_R_6 = INF_R ** 6
INF_R_9 = INF_R ** 9
W_CONST = 315.0 / (64.0 * np.pi * INF_R_9)
GRAD_W_CONST = -45.0 / (np.pi * INF_R_6)
LAP_W_CONST = 45.0 / (np.pi * INF_R_6)
MAX_NEIGHBOURS = 32
NEIGHBOURING_VOXELS_COUNT = 27
VOXEL_SIZE = [INF_R, INF_R, INF_R]
BOX_SPACE_SIZE = [20 * INF_R, 20 * INF_R, 20 * INF_R]
PIPE_SPACE_SIZE = [20 * INF_R, 3 * INF_R, 3 * INF_R]
GRAVITY = [0.0, -2.0, 0.0]
HORIZONTAL_FORCE = [2.0, 0.0, 0.0]
PARAMS_FILENAME = "params.json"
OUT_DIRNAME = "out"


def constants(mode: str = SIM_MODE) -> SphConstants:
    return SphConstants(mode=mode, h=INF_R, mass=MASS, rho0=RHO_0, visc=VISC, k=K, damp=DAMP,
                        max_neighbours=MAX_NEIGHBOURS)


def build_pipe(space_size=PIPE_SPACE_SIZE) -> Pipe:
    """The six-segment pipe of config.py:68-76 scaled into `space_size`."""
    return (PipeBuilder().with_starting_radius(1)
            .add_roller_segment(1).add_increasing_segment(1, 1.2).add_roller_segment(1)
            .add_lessening_segment(1, 1.2).add_roller_segment(1)
            .transform(space_size[0], space_size[1]).get_result())


def box_params(particle_count=PARTICLE_COUNT, space_size=BOX_SPACE_SIZE, duration=DURATION, fps=FPS):
    """config.py:56-65.  Box mode saves an empty Pipe so the viewer draws no pipe wireframe (viewport_layer.py:24)."""
    return SimulationParameters(particle_count=int(particle_count), external_force=np.asarray(GRAVITY),
                                duration=duration, fps=fps, pipe=Pipe(), space_size=np.asarray(space_size, float),
                                voxel_size=np.asarray(VOXEL_SIZE))


def pipe_params(particle_count=PARTICLE_COUNT, space_size=PIPE_SPACE_SIZE, duration=DURATION, fps=FPS, pipe=None):
    """config.py:44-53."""
    return SimulationParameters(particle_count=int(particle_count), external_force=np.asarray(HORIZONTAL_FORCE),
                                duration=duration, fps=fps, pipe=pipe or build_pipe(space_size),
                                space_size=np.asarray(space_size, float), voxel_size=np.asarray(VOXEL_SIZE))


def start_state_box_wall(particle_count, space_size=BOX_SPACE_SIZE, seed=0, dtype=np.float64) -> SimulationState:
    """Dam-break column: uniform in the first 10 % of x, full y and z; velocity [1.5,-5,-5] +- 0.5 (config.py:79-97).
    Values are generated in fp32 so the fp32 engine and an fp64 checker see identical inputs."""
    rng = np.random.default_rng(seed)
    n = int(particle_count)
    pos = rng.random((n, 3), dtype=np.float32)
    pos *= np.asarray([space_size[0] * 0.1, space_size[1], space_size[2]], dtype=np.float32)
    vel = (rng.random((n, 3), dtype=np.float32) - np.float32(0.5)) + np.asarray([1.5, -5.0, -5.0], dtype=np.float32)
    return SimulationState(pos.astype(dtype), vel.astype(dtype), np.zeros(n, dtype))


def start_state_inside_pipe(particle_count, pipe: Pipe, seed=0, dtype=np.float64) -> SimulationState:
    """Uniform in x along the pipe, uniform over 98 % of the local disc, zero velocity (config.py:100-120)."""
    rng = np.random.default_rng(seed)
    n = int(particle_count)
    x = rng.random(n) * pipe.get_length()
    table = pipe.to_numpy()
    starts = table[:-1, 0]
    seg = np.clip(np.searchsorted(starts, x, side="right") - 1, 0, len(starts) - 1)
    r0 = table[seg, 3]
    r1 = np.append(table[1:-1, 3], table[-1, 3])[seg]
    r_max = r0 + (r1 - r0) * (x - starts[seg]) / table[seg, 4]
    r = np.sqrt(rng.random(n)) * r_max * 0.98
    theta = rng.random(n) * 2.0 * np.pi
    pos = np.stack([x, table[0, 1] + r * np.cos(theta), table[0, 2] + r * np.sin(theta)], axis=1)
    pos = pos.astype(np.float32)
    return SimulationState(pos.astype(dtype), np.zeros((n, 3), dtype), np.zeros(n, dtype))


# ---- host mirror of the device-side start-state generators (csrc/sph_kernels.cuh: generate_kernel) ------------------
def _u24(seed: int, ids: np.ndarray, stream: int) -> np.ndarray:
    """Draw `stream` of every particle: SplitMix64 finaliser of seed + golden * (16 id + stream + 1), top 24 bits."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (ids.astype(np.uint64) * np.uint64(16)
                                                                + np.uint64(stream + 1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)


def hashed_start_state(kind: str, params: SimulationParameters, seed: int = 0) -> SimulationState:
    """The state sph_generate_state(kind, seed) writes on the device, bit for bit (section 8(f)2): a pure function of
    (kind, seed, particle id), so any rank of a multi-GPU run can produce its own particles."""
    n = int(params.particle_count)
    ids = np.arange(n, dtype=np.uint64)
    space = np.asarray(params.space_size, np.float64)
    if kind in ("box_wall", "uniform"):
        ext = space.copy()
        if kind == "box_wall":
            ext[0] = space[0] * 0.1
        ext32 = ext.astype(np.float32)
        top = np.nextafter(space.astype(np.float32), np.float32(0))
        pos = np.stack([np.minimum(_u24(seed, ids, d) * ext32[d], top[d]) for d in range(3)], axis=1)
        base = np.asarray([1.5, -5.0, -5.0], np.float32)
        vel = np.stack([(_u24(seed, ids, 3 + d) + np.float32(-0.5)) + base[d] for d in range(3)], axis=1)
        return SimulationState(pos.astype(np.float64), vel.astype(np.float64), np.zeros(n))
    if kind != "pipe":
        raise ValueError(kind)
    t = np.ascontiguousarray(params.pipe.to_numpy(), np.float64)
    rows = len(t)
    length = t[rows - 1, 0] - t[0, 0]
    xd = _u24(seed, ids, 0).astype(np.float64) * length
    starts = t[1:rows - 1, 0] - t[0, 0]                       # a particle is in segment s while x < start of s + 1
    s_ = np.searchsorted(starts, xd, side="right")
    r0, r1 = t[s_, 3], t[s_ + 1, 3]
    frac = (xd - (t[s_, 0] - t[0, 0])) / t[s_, 4]
    rmax = (r0 + (r1 - r0) * frac) * 0.98
    a, b, done = np.zeros(n), np.zeros(n), np.zeros(n, bool)
    for tr in range(7):
        ta = _u24(seed, ids, 1 + 2 * tr).astype(np.float64) * 2.0 - 1.0
        tb = _u24(seed, ids, 2 + 2 * tr).astype(np.float64) * 2.0 - 1.0
        ok = ~done & (ta * ta + tb * tb <= 1.0)
        a[ok], b[ok] = ta[ok], tb[ok]
        done |= ok
    pos = np.stack([xd + t[0, 0], t[0, 1] + a * rmax, t[0, 2] + b * rmax], axis=1).astype(np.float32)
    return SimulationState(pos.astype(np.float64), np.zeros((n, 3)), np.zeros(n))
