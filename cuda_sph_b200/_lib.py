"""ctypes binding of libsph_b200.so (include/sph_b200.h).  There is no CPU fallback: a missing library or a missing
GPU raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPH_B200_LIB: developer knob for A/B runs of variant builds (build.build_variant); the product library otherwise
LIB_PATH = os.environ.get("SPH_B200_LIB") or os.path.join(_HERE, "libsph_b200.so")

MODE_BOX, MODE_PIPE = 0, 1
FLAG_RECORD_NEIGHBOUR_COUNTS, FLAG_RECORD_TERMS, FLAG_NO_GRAPH, FLAG_SLAB = 1, 2, 4, 8


class SphParams(C.Structure):
    _fields_ = [("particle_count", C.c_int32), ("mode", C.c_int32), ("h", C.c_double), ("mass", C.c_double),
                ("rho0", C.c_double), ("k", C.c_double), ("visc", C.c_double), ("damp", C.c_double),
                ("dt", C.c_double), ("external_force", C.c_double * 3), ("space_size", C.c_double * 3),
                ("voxel_size", C.c_double * 3), ("max_neighbours", C.c_int32), ("flags", C.c_uint32),
                ("rng_seed", C.c_uint64)]


class SphTimings(C.Structure):
    _fields_ = [("hash_ms", C.c_float), ("sort_ms", C.c_float), ("reorder_ms", C.c_float),
                ("density_ms", C.c_float), ("force_ms", C.c_float), ("total_ms", C.c_float), ("steps", C.c_int32),
                ("launches_per_step", C.c_int32), ("sort_passes", C.c_int32)]


class SphStats(C.Structure):
    _fields_ = [("n_particles", C.c_int32), ("n_dead", C.c_int32), ("n_nonfinite", C.c_int32),
                ("n_cells", C.c_int32), ("max_density", C.c_float), ("max_speed", C.c_float),
                ("steps_done", C.c_int64)]


class SphFrameStats(C.Structure):
    _fields_ = [("steps_done", C.c_int64), ("n_particles", C.c_int32), ("n_dead", C.c_int32),
                ("n_nonfinite", C.c_int32), ("reserved", C.c_int32), ("max_position", C.c_float),
                ("min_position", C.c_float), ("max_velocity", C.c_float), ("max_speed", C.c_float),
                ("max_density", C.c_float), ("neighbour_hist", C.c_int32 * 33)]


GEN_BOX_WALL, GEN_UNIFORM, GEN_PIPE = 0, 1, 2

# every symbol include/sph_b200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
_P = C.c_void_p
EXPORTS = {
    "sph_last_error": (C.c_char_p, []),
    "sph_version": (C.c_int, []),
    "sph_create": (C.c_int, [C.POINTER(SphParams), C.c_int, C.POINTER(_H)]),
    "sph_destroy": (C.c_int, [_H]),
    "sph_set_pipe": (C.c_int, [_H, _P, C.c_int32]),
    "sph_set_stream": (C.c_int, [_H, _P]),
    "sph_upload": (C.c_int, [_H, _P, _P]),
    "sph_upload_f32": (C.c_int, [_H, _P, _P]),
    "sph_step": (C.c_int, [_H, C.c_int32]),
    "sph_step_timed": (C.c_int, [_H, C.c_int32, C.POINTER(SphTimings)]),
    "sph_download": (C.c_int, [_H, _P, _P, _P]),
    "sph_download_f32": (C.c_int, [_H, _P, _P, _P]),
    "sph_compute_next_state": (C.c_int, [_H, _P, _P, _P, _P, _P]),
    "sph_sync": (C.c_int, [_H]),
    "sph_save_state": (C.c_int, [_H]),
    "sph_restore_state": (C.c_int, [_H]),
    "sph_slab_configure": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int64]),
    "sph_slab_step": (C.c_int, [_H, C.c_int32, C.c_int32]),
    "sph_slab_exchange_init": (C.c_int, [_H, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P, C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_void_p), _P]),
    "sph_slab_route": (C.c_int, [_H]),
    "sph_slab_unpack": (C.c_int, [_H]),
    "sph_slab_step_all": (C.c_int, [_H]),
    "sph_slab_step_all_timed": (C.c_int, [_H, C.POINTER(SphTimings)]),
    "sph_slab_compact": (C.c_int, [_H]),
    "sph_slab_counters": (C.c_int, [_H, _P]),
    "sph_slab_parity": (C.c_int, [_H, _P]),
    "sph_slab_ipc_handle": (C.c_int, [_H, _P]),
    "sph_slab_open_peers": (C.c_int, [_H, _P, _P]),
    "sph_slab_exchange_p2p": (C.c_int, [_H]),
    "sph_slab_exchange_p2p_timed": (C.c_int, [_H, _P]),
    "sph_get_keys": (C.c_int, [_H, _P]),
    "sph_get_sorted_ids": (C.c_int, [_H, _P]),
    "sph_get_sorted_keys": (C.c_int, [_H, _P]),
    "sph_get_voxel_begin": (C.c_int, [_H, _P, C.c_int64]),
    "sph_get_neighbour_counts": (C.c_int, [_H, _P]),
    "sph_get_neighbour_lists": (C.c_int, [_H, _P]),
    "sph_get_forces": (C.c_int, [_H, _P]),
    "sph_get_terms": (C.c_int, [_H, _P, _P]),
    "sph_get_rng_states": (C.c_int, [_H, _P]),
    "sph_set_rng_states": (C.c_int, [_H, _P]),
    "sph_get_stats": (C.c_int, [_H, C.POINTER(SphStats)]),
    "sph_export_begin": (C.c_int, [_H, C.c_int32, C.c_int32]),
    "sph_export_wait": (C.c_int, [_H, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_int64)]),
    "sph_generate_state": (C.c_int, [_H, C.c_int32, C.c_uint64]),
    "sph_get_frame_stats": (C.c_int, [_H, C.POINTER(SphFrameStats)]),
    "sph_export_stats": (C.c_int, [_H, C.c_int32, C.POINTER(SphFrameStats)]),
    "sph_n_cells": (C.c_int64, [_H]),
    "sph_path_counters": (C.c_int, [_H, _P]),
    "sph_cell_dims": (C.c_int, [_H, _P, _P]),
    "sph_device_ptr": (C.c_int, [_H, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "sph_launch_count": (C.c_int64, [_H]),
}

_lib = None


class SphError(RuntimeError):
    pass


def load():
    """Load libsph_b200.so; raises if it has not been built (python -m cuda_sph_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SphError(f"{LIB_PATH} is missing: build it with `python -m cuda_sph_b200.build` "
                           "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name, None)
            if fn is None:
                if os.environ.get("SPH_B200_LIB"):   # an A/B variant library built from an older tree
                    continue
                raise SphError(f"{LIB_PATH} lacks the symbol {name}: rebuild it (python -m cuda_sph_b200.build)")
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        raise SphError(load().sph_last_error().decode())
