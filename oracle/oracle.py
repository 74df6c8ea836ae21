"""TEST INFRASTRUCTURE -- ctypes front-end of oracle/sph_oracle.c (fp64 CPU restatement of the reference step).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")
MAX_NEIGHBOURS = 32


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sph_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liborc.so"])
    return _LIB_PATH


class _Params(C.Structure):
    _fields_ = [("n", C.c_int32), ("mode", C.c_int32), ("h", C.c_double), ("mass", C.c_double),
                ("rho0", C.c_double), ("k", C.c_double), ("visc", C.c_double), ("damp", C.c_double),
                ("dt", C.c_double), ("ext", C.c_double * 3), ("space", C.c_double * 3),
                ("voxel", C.c_double * 3), ("pipe_rows", C.c_int32), ("pad_", C.c_int32),
                ("pipe", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_n_cells.restype = C.c_int64
        _lib.orc_x_at_segment_beginning.restype = C.c_double
        _lib.orc_vector_length.restype = C.c_double
        _lib.orc_distance_between_points.restype = C.c_double
        _lib.orc_rng_uniform.restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class OracleParams:
    """Run parameters: config.py:13-36 constants + SimulationParameters (common/data_classes.py:70-78)."""
    n: int
    mode: str = "BOX"
    h: float = 2.0
    mass: float = 1.0
    rho0: float = 1.0
    k: float = 10.0
    visc: float = 0.5
    damp: float = 0.7
    dt: float = 1 / 20
    ext: tuple = (0.0, -2.0, 0.0)
    space: tuple = (40.0, 40.0, 40.0)
    voxel: tuple = (2.0, 2.0, 2.0)
    pipe: np.ndarray | None = None  # (S+1, 5) table from Pipe.to_numpy()
    _keep: list = field(default_factory=list, repr=False)

    def c(self) -> _Params:
        s = _Params()
        s.n, s.mode = int(self.n), 1 if self.mode.upper() == "PIPE" else 0
        s.h, s.mass, s.rho0, s.k, s.visc, s.damp, s.dt = (float(self.h), float(self.mass), float(self.rho0),
                                                          float(self.k), float(self.visc), float(self.damp),
                                                          float(self.dt))
        for d in range(3):
            s.ext[d], s.space[d], s.voxel[d] = float(self.ext[d]), float(self.space[d]), float(self.voxel[d])
        if self.pipe is not None and np.size(self.pipe) > 0:
            t = np.ascontiguousarray(self.pipe, dtype=np.float64)
            self._keep.append(t)
            s.pipe_rows, s.pipe = t.shape[0], t.ctypes.data
        else:
            s.pipe_rows, s.pipe = 0, None
        return s


def set_exact_pow(on: bool):
    lib().orc_set_exact_pow(int(bool(on)))


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(t: int):
    lib().orc_set_num_threads(int(t))


def constants(h: float):
    out = np.zeros(3)
    lib().orc_constants(C.c_double(h), _p(out))
    return tuple(out)


def dims(P: OracleParams):
    c3, t3 = np.zeros(3, np.int32), np.zeros(3, np.int32)
    cp = P.c()
    lib().orc_dims(C.byref(cp), _p(c3), _p(t3))
    return c3, t3


def n_cells(P: OracleParams) -> int:
    cp = P.c()
    return int(lib().orc_n_cells(C.byref(cp)))


def cell_keys(P: OracleParams, pos: np.ndarray) -> np.ndarray:
    pos = np.ascontiguousarray(pos, np.float64)
    keys = np.empty(P.n, np.int32)
    cp = P.c()
    lib().orc_cell_keys(C.byref(cp), _p(pos), _p(keys))
    return keys


def sort_cells(P: OracleParams, keys: np.ndarray):
    """-> (map_ids, map_keys, voxel_begin, n_dead)  (voxel_sph_strategy.py:81-107)"""
    keys = np.ascontiguousarray(keys, np.int32)
    ids, mk = np.empty(P.n, np.int32), np.empty(P.n, np.int32)
    vb = np.empty(n_cells(P), np.int32)
    cp = P.c()
    nd = lib().orc_sort_cells(C.byref(cp), _p(keys), _p(ids), _p(mk), _p(vb))
    return ids, mk, vb, int(nd)


def rng_init(n: int, seed: int = 16435234) -> np.ndarray:
    st = np.zeros((n, 2), np.uint64)
    lib().orc_rng_init(_p(st), C.c_int64(n), C.c_uint64(seed))
    return st


@dataclass
class StepResult:
    position: np.ndarray
    velocity: np.ndarray
    density: np.ndarray
    force: np.ndarray
    pressure: np.ndarray
    viscosity: np.ndarray
    keys: np.ndarray
    map_ids: np.ndarray
    voxel_begin: np.ndarray
    neigh_count: np.ndarray
    neighbours: np.ndarray | None
    n_dead: int


def step(P: OracleParams, pos, vel, rng=None, want_neighbours: bool = False, light: bool = False) -> StepResult:
    """One compute_next_state.  `rng` (N x 2 uint64) is advanced in place in PIPE mode.
    light=True skips the debug outputs (used when timing the CPU baseline)."""
    pos = np.array(pos, dtype=np.float64, order="C", copy=True)
    vel = np.array(vel, dtype=np.float64, order="C", copy=True)
    n = P.n
    assert pos.shape == (n, 3) and vel.shape == (n, 3)
    if P.mode.upper() == "PIPE" and rng is None:
        raise ValueError("PIPE mode needs rng states (oracle.rng_init)")
    rho = np.empty(n)
    if light:
        force = pr = vi = keys = ids = vb = cnt = neigh = None
    else:
        force, pr, vi = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 3))
        keys, ids, cnt = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.int32)
        vb = np.empty(n_cells(P) + 1, np.int32)
        neigh = np.full((n, MAX_NEIGHBOURS), -1, np.int32) if want_neighbours else None
    cp = P.c()
    nd = lib().orc_step(C.byref(cp), _p(pos), _p(vel), _p(rng), _p(rho), _p(force), _p(pr), _p(vi), _p(keys),
                        _p(ids), _p(vb), _p(cnt), _p(neigh))
    return StepResult(pos, vel, rho, force, pr, vi, keys, ids, None if vb is None else vb[:-1], cnt, neigh, int(nd))


# ---- pipe-geometry hooks for the reference's own known-answer tests (sim/tests/test_collisions.py) ----
def find_segment(pipe, x):
    t = np.ascontiguousarray(pipe, np.float64)
    return int(lib().orc_find_segment(_p(t), C.c_int(t.shape[0]), C.c_double(x)))


def x_at_segment_beginning(pipe, s):
    t = np.ascontiguousarray(pipe, np.float64)
    return float(lib().orc_x_at_segment_beginning(_p(t), C.c_int(s)))


def vector_length(v):
    return float(lib().orc_vector_length(_p(np.ascontiguousarray(v, np.float64))))


def distance_between_points(a, b):
    return float(lib().orc_distance_between_points(_p(np.ascontiguousarray(a, np.float64)),
                                                   _p(np.ascontiguousarray(b, np.float64))))


def is_out_of_pipe(pos, pipe, s):
    t = np.ascontiguousarray(pipe, np.float64)
    return bool(lib().orc_is_out_of_pipe(_p(np.ascontiguousarray(pos, np.float64)), _p(t), C.c_int(s)))


def solve_collision(pos, vel, pipe, s):
    t = np.ascontiguousarray(pipe, np.float64)
    p, v = np.array(pos, np.float64), np.array(vel, np.float64)
    lib().orc_solve_collision(_p(p), _p(v), _p(t), C.c_int(s))
    return p, v


def collide_pipe(P: OracleParams, pos, vel, rng):
    pos = np.array(pos, dtype=np.float64, order="C", copy=True)
    vel = np.array(vel, dtype=np.float64, order="C", copy=True)
    cp = P.c()
    lib().orc_collide_pipe(C.byref(cp), _p(pos), _p(vel), _p(rng))
    return pos, vel
