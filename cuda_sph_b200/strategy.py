"""Host-side mirror of the reference's strategy interface for the Voxel step, backed by libsph_b200.so.

Replaces sim/src/sph/strategies/{abstract,voxel}_sph_strategy.py: same constructor argument, same
``compute_next_state(old_state) -> SimulationState`` contract (fp64, C-contiguous, particle-id-ordered fresh arrays;
the input's density is ignored; inputs are never mutated), same public attributes (``params``, ``dt``, ``grid_size``,
``block_size``, ``result_force``, ``old_state``, ``new_state``).  It is duck-typed rather than a subclass of the
reference's AbstractSPHStrategy, whose constructor raises KeyError on compute capability 10.0
(sim/src/sph/thread_layout.py:10-29).

On top of that it offers the device-resident fast path the reference lacks: ``upload`` / ``step(n)`` / ``download``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from .data_classes import SimulationParameters, SimulationState

SWEEP_BLOCK = 128  # threads per CTA of the neighbour sweeps (csrc/sph_kernels.cuh)


@dataclass
class SphConstants:
    """The module constants the reference freezes into its kernels (config.py:13-36)."""
    mode: str = "BOX"          # config.SIM_MODE
    h: float = 2.0             # INF_R
    mass: float = 1.0          # MASS
    rho0: float = 1.0          # RHO_0
    visc: float = 0.5          # VISC
    k: float = 10.0            # K
    damp: float = 0.7          # DAMP
    max_neighbours: int = 32   # MAX_NEIGHBOURS
    rng_seed: int = 16435234   # abstract_sph_strategy.py:27


def _f64(a, shape):
    out = np.ascontiguousarray(a, dtype=np.float64)
    if out.shape != shape:
        raise ValueError(f"expected array of shape {shape}, got {out.shape}")
    return out


class B200SPHStrategy:
    def __init__(self, params: SimulationParameters, constants: Optional[SphConstants] = None, *, device: int = 0,
                 record_neighbour_counts: bool = False, record_terms: bool = False, use_graph: bool = True,
                 cuda_stream: Optional[int] = None):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.params = params
        self.constants = constants or SphConstants()
        self.dt = 1 / self.params.fps                      # abstract_sph_strategy.py:20
        self.old_state: Optional[SimulationState] = None
        self.new_state: Optional[SimulationState] = None
        n = int(params.particle_count)
        self.n = n
        self.block_size = SWEEP_BLOCK                      # read by sim/src/main.py:21-22
        self.grid_size = (n + SWEEP_BLOCK - 1) // SWEEP_BLOCK
        self._result_force = np.zeros((n, 3), dtype=np.float64)   # abstract_sph_strategy.py:26
        self._force_stale = False

        cst = self.constants
        p = _lib.SphParams()
        p.particle_count = n
        p.mode = _lib.MODE_PIPE if cst.mode.upper() == "PIPE" else _lib.MODE_BOX
        p.h, p.mass, p.rho0, p.k, p.visc, p.damp = cst.h, cst.mass, cst.rho0, cst.k, cst.visc, cst.damp
        p.dt = self.dt
        ext = np.asarray(params.external_force, dtype=np.float64).reshape(3)
        space = np.asarray(params.space_size, dtype=np.float64).reshape(3)
        voxel = np.asarray(params.voxel_size, dtype=np.float64).reshape(3)
        for d in range(3):
            p.external_force[d], p.space_size[d], p.voxel_size[d] = ext[d], space[d], voxel[d]
        p.max_neighbours = cst.max_neighbours
        p.flags = ((_lib.FLAG_RECORD_NEIGHBOUR_COUNTS if record_neighbour_counts else 0)
                   | (_lib.FLAG_RECORD_TERMS if record_terms else 0) | (0 if use_graph else _lib.FLAG_NO_GRAPH))
        p.rng_seed = cst.rng_seed
        _lib.check(self._lib.sph_create(C.byref(p), int(device), C.byref(self._h)))
        if cuda_stream is not None:
            _lib.check(self._lib.sph_set_stream(self._h, C.c_void_p(int(cuda_stream))))
        if p.mode == _lib.MODE_PIPE:
            table = np.ascontiguousarray(params.pipe.to_numpy(), dtype=np.float64)
            if table.ndim != 2 or table.shape[1] != 5:
                raise ValueError("PIPE mode needs params.pipe with at least one segment")
            _lib.check(self._lib.sph_set_pipe(self._h, table.ctypes.data, table.shape[0]))

    # ------------------------------------------------------------------ reference-facing call
    def compute_next_state(self, old_state: SimulationState) -> SimulationState:
        """abstract_sph_strategy.py:31-46: one step, host arrays in, fresh host arrays out.

        Precision: the engine computes in fp32 and rounds the fp64 input once, on the way in, so the result is the
        reference's result on float32(old_state) -- bit-exact grid / sort / neighbour lists, physics within 1e-4
        (INTEGRATION.md "Input precision"; a state that is not fp32-representable may differ from the reference run on
        the fp64 values where the rounding moves a particle across a cell face or a pair across r = h)."""
        self.old_state = old_state
        n = self.n
        pos = _f64(old_state.position, (n, 3))
        vel = _f64(old_state.velocity, (n, 3))
        out_p, out_v, out_r = np.empty((n, 3)), np.empty((n, 3)), np.empty(n)
        self.compute_next_state_into(pos, vel, out_p, out_v, out_r)
        self.new_state = SimulationState(out_p, out_v, out_r)
        return self.new_state

    def compute_next_state_into(self, pos, vel, out_pos, out_vel, out_rho):
        """Same, into caller-owned fp64 buffers (e.g. pinned memory); nothing is allocated.  The C ABI takes raw
        pointers, so dtype, shape and contiguity are checked here."""
        n = self.n
        for name, arr, shape in (("pos", pos, (n, 3)), ("vel", vel, (n, 3)), ("out_pos", out_pos, (n, 3)),
                                 ("out_vel", out_vel, (n, 3)), ("out_rho", out_rho, (n,))):
            if not (isinstance(arr, np.ndarray) and arr.dtype == np.float64 and arr.shape == shape
                    and arr.flags.c_contiguous):
                raise ValueError(f"{name}: expected a C-contiguous float64 array of shape {shape}")
        _lib.check(self._lib.sph_compute_next_state(self._h, pos.ctypes.data, vel.ctypes.data, out_pos.ctypes.data,
                                                    out_vel.ctypes.data, out_rho.ctypes.data))
        self._force_stale = True

    @property
    def result_force(self) -> np.ndarray:
        """abstract_sph_strategy.py:83 -- fetched lazily from the device."""
        if self._force_stale:
            _lib.check(self._lib.sph_get_forces(self._h, self._result_force.ctypes.data))
            self._force_stale = False
        return self._result_force

    # ------------------------------------------------------------------ device-resident fast path
    def upload(self, state: SimulationState):
        n = self.n
        pos, vel = np.asarray(state.position), np.asarray(state.velocity)
        if pos.dtype == np.float32 and vel.dtype == np.float32:
            pos, vel = np.ascontiguousarray(pos), np.ascontiguousarray(vel)
            if pos.shape != (n, 3) or vel.shape != (n, 3):
                raise ValueError(f"position / velocity must have shape ({n}, 3)")
            _lib.check(self._lib.sph_upload_f32(self._h, pos.ctypes.data, vel.ctypes.data))
        else:
            pos, vel = _f64(pos, (n, 3)), _f64(vel, (n, 3))
            _lib.check(self._lib.sph_upload(self._h, pos.ctypes.data, vel.ctypes.data))
        self.old_state = state

    def step(self, n_steps: int = 1):
        _lib.check(self._lib.sph_step(self._h, int(n_steps)))
        self._force_stale = True

    def step_timed(self, n_steps: int = 1) -> dict:
        t = _lib.SphTimings()
        _lib.check(self._lib.sph_step_timed(self._h, int(n_steps), C.byref(t)))
        self._force_stale = True
        return {name: getattr(t, name) for name, _ in t._fields_}

    def download(self, dtype=np.float64) -> SimulationState:
        n = self.n
        dtype = np.dtype(dtype)
        if dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("download() supports float64 and float32 only")
        pos, vel, rho = np.empty((n, 3), dtype), np.empty((n, 3), dtype), np.empty(n, dtype)
        fn = self._lib.sph_download if dtype == np.float64 else self._lib.sph_download_f32
        _lib.check(fn(self._h, pos.ctypes.data, vel.ctypes.data, rho.ctypes.data))
        self.new_state = SimulationState(pos, vel, rho)
        return self.new_state

    # ------------------------------------------------------------------ frame export pipeline (section 8(f)1)
    def export_async(self, slot: int, stride: int = 1) -> None:
        """Snapshot the state (fp64, id order, every `stride`-th particle) into pinned host buffer `slot` (0..2) on a
        second stream; returns at once -- the copy runs under the steps enqueued next."""
        _lib.check(self._lib.sph_export_begin(self._h, int(slot), int(stride)))

    def export_wait(self, slot: int, copy: bool = True) -> SimulationState:
        """The frame of `slot` once its copy has landed.  copy=False returns views of the pinned buffer, valid until the
        slot is exported into again."""
        p, v, r, m = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(self._lib.sph_export_wait(self._h, int(slot), C.byref(p), C.byref(v), C.byref(r), C.byref(m)))
        k = int(m.value)

        def view(ptr, shape):
            buf = (C.c_double * int(np.prod(shape))).from_address(ptr.value)
            a = np.frombuffer(buf, dtype=np.float64).reshape(shape)
            return a.copy() if copy else a
        return SimulationState(view(p, (k, 3)), view(v, (k, 3)), view(r, (k,)))

    def generate_state(self, kind: str, seed: int = 0) -> None:
        """Seeded start state generated on the device (section 8(f)2; config.hashed_start_state is the host mirror):
        kind = 'box_wall' (config.py:84-95), 'uniform', 'pipe' (config.py:105-115)."""
        code = {"box_wall": _lib.GEN_BOX_WALL, "uniform": _lib.GEN_UNIFORM, "pipe": _lib.GEN_PIPE}[kind]
        _lib.check(self._lib.sph_generate_state(self._h, code, C.c_uint64(int(seed))))

    @staticmethod
    def _stats_dict(st) -> dict:
        out = {name: getattr(st, name) for name, _ in st._fields_ if name not in ("neighbour_hist", "reserved")}
        out["neighbour_hist"] = [int(x) for x in st.neighbour_hist]
        return out

    def frame_stats(self) -> dict:
        """On-device reductions over the current state (analize.py:9-14 + neighbour-count histogram)."""
        st = _lib.SphFrameStats()
        _lib.check(self._lib.sph_get_frame_stats(self._h, C.byref(st)))
        return self._stats_dict(st)

    def export_stats(self, slot: int) -> dict:
        """The same reductions for the frame exported into `slot` (they travelled with it)."""
        st = _lib.SphFrameStats()
        _lib.check(self._lib.sph_export_stats(self._h, int(slot), C.byref(st)))
        return self._stats_dict(st)

    def save_state(self):
        """Device-side snapshot of the particle state (checkpoint)."""
        _lib.check(self._lib.sph_save_state(self._h))

    def restore_state(self):
        _lib.check(self._lib.sph_restore_state(self._h))

    def synchronize(self):
        _lib.check(self._lib.sph_sync(self._h))

    # ------------------------------------------------------------------ parity taps
    def _i32(self, fn, count):
        out = np.empty(count, np.int32)
        _lib.check(fn(self._h, out.ctypes.data))
        return out

    def keys(self):
        return self._i32(self._lib.sph_get_keys, self.n)

    def sorted_ids(self):
        return self._i32(self._lib.sph_get_sorted_ids, self.n)

    def sorted_keys(self):
        return self._i32(self._lib.sph_get_sorted_keys, self.n)

    def n_cells(self) -> int:
        return int(self._lib.sph_n_cells(self._h))

    def cell_dims(self):
        c3, t3 = np.zeros(3, np.int32), np.zeros(3, np.int32)
        _lib.check(self._lib.sph_cell_dims(self._h, c3.ctypes.data, t3.ctypes.data))
        return c3, t3

    def voxel_begin(self):
        nc = self.n_cells()
        out = np.empty(nc, np.int32)
        _lib.check(self._lib.sph_get_voxel_begin(self._h, out.ctypes.data, nc))
        return out

    def neighbour_counts(self):
        return self._i32(self._lib.sph_get_neighbour_counts, self.n)

    def neighbour_lists(self) -> np.ndarray:
        """(N, 32) int32: the neighbour lists of the most recent step as particle ids in list order, -1 padded
        (`neighbours` of get_neighbours, voxel_kernels.py:29-85)."""
        out = np.empty((self.n, 32), np.int32)
        _lib.check(self._lib.sph_get_neighbour_lists(self._h, out.ctypes.data))
        return out

    def terms(self):
        pr, vi = np.empty((self.n, 3)), np.empty((self.n, 3))
        _lib.check(self._lib.sph_get_terms(self._h, pr.ctypes.data, vi.ctypes.data))
        return pr, vi

    def rng_states(self):
        out = np.empty((self.n, 2), np.uint64)
        _lib.check(self._lib.sph_get_rng_states(self._h, out.ctypes.data))
        return out

    def set_rng_states(self, states):
        st = np.ascontiguousarray(states, np.uint64)
        if st.shape != (self.n, 2):
            raise ValueError(f"rng states must have shape ({self.n}, 2)")
        _lib.check(self._lib.sph_set_rng_states(self._h, st.ctypes.data))

    def stats(self) -> dict:
        s = _lib.SphStats()
        _lib.check(self._lib.sph_get_stats(self._h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in s._fields_}

    def path_counters(self) -> dict:
        """Work-item counts of the most recent step: which share of the tiles ran on which sweep path."""
        out = np.zeros(4, np.int32)
        _lib.check(self._lib.sph_path_counters(self._h, out.ctypes.data))
        return dict(passes=int(out[0]), flat_refused=int(out[1]), dense_tiles=int(out[2]), tiles=int(out[3]))

    def launch_count(self) -> int:
        return int(self._lib.sph_launch_count(self._h))

    def device_ptr(self, which: int):
        ptr, cnt = C.c_void_p(), C.c_int64()
        _lib.check(self._lib.sph_device_ptr(self._h, which, C.byref(ptr), C.byref(cnt)))
        return ptr.value, cnt.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.sph_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
