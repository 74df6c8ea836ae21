"""Seeded synthetic workloads of SURVEY.md section 8d (S1 uniform, S2 dam-break column, S3 pipe), fp32-representable."""
from __future__ import annotations

import math

import numpy as np

from . import config
from .data_classes import SimulationParameters, SimulationState


def cubic_dims(n: int, ppc: float) -> int:
    """Cells per side so that n particles fill a cube at `ppc` particles per cell."""
    return max(3, math.ceil((n / ppc) ** (1.0 / 3.0)))


def uniform_box(n: int, ppc: float = 2.5, seed: int = 0, fps: int = config.FPS):
    """S1: positions U[0, space)^3, velocities [1.5,-5,-5] + U(-0.5, 0.5)."""
    d = cubic_dims(n, ppc)
    space = [d * config.INF_R] * 3
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3), dtype=np.float32) * np.float32(space[0])
    pos = np.minimum(pos, np.nextafter(np.float32(space[0]), np.float32(0)))
    vel = (rng.random((n, 3), dtype=np.float32) - np.float32(0.5)) + np.asarray([1.5, -5.0, -5.0], np.float32)
    params = config.box_params(n, space, fps=fps)
    return params, SimulationState(pos.astype(np.float64), vel.astype(np.float64), np.zeros(n))


def dam_break(n: int, ppc_global: float = 2.5, seed: int = 0, fps: int = config.FPS):
    """S2: box sized for `ppc_global`, particles uniform in the first 10 % of x (config.py:84-87) -> ~25 per cell."""
    d = cubic_dims(n, ppc_global)
    space = [d * config.INF_R] * 3
    params = config.box_params(n, space, fps=fps)
    return params, config.start_state_box_wall(n, space, seed)


def pipe_flow(n: int, seed: int = 0, fps: int = config.FPS):
    """S3: the config.py:68-76 pipe scaled so n / volume matches the reference (20 000 particles in 40 x 6 x 6)."""
    s = (n / 20000.0) ** (1.0 / 3.0)
    cells = [max(3, round(20 * s)), max(3, round(3 * s)), max(3, round(3 * s))]
    space = [cells[0] * config.INF_R, cells[1] * config.INF_R, cells[1] * config.INF_R]
    pipe = config.build_pipe(space)
    params = config.pipe_params(n, space, fps=fps, pipe=pipe)
    return params, config.start_state_inside_pipe(n, pipe, seed)
