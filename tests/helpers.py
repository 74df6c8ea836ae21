"""Shared helpers for the parity tests."""
from __future__ import annotations

import os

import numpy as np

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def params_from_golden(g, mode):
    n = len(g["pos_in"])
    return orc.OracleParams(n=n, mode=mode, dt=1.0 / float(g["fps"]), ext=tuple(g["ext"]), space=tuple(g["space"]),
                            voxel=tuple(g["voxel"]), pipe=g["pipe"] if mode == "PIPE" else None)


def same(a, b):
    """Bitwise equality with NaN == NaN."""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def max_rel(a, b):
    """max |a-b| / max(|b|, tiny) over finite entries; non-finite entries must match exactly."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), "finite masks differ"
    nf = ~fin
    assert same(a[nf], b[nf]), "non-finite entries differ"
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300)))


def vec_rel(a, b, floor=None):
    """max over particles of ||a_i - b_i|| / max(||b_i||, floor_i).  Rows whose reference is non-finite must be
    non-finite in the same places.  `floor` (per particle) is used for the pressure / viscosity TERMS only: they act
    through the net force, so their error is judged against max(|term|, |force|) -- a viscosity component that is a
    cancelling sum of 1e4-sized pair terms next to two components clamped to 1 is not a 1e-4-relative quantity."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    fin = np.isfinite(b).all(axis=1)
    assert np.array_equal(np.isfinite(a[~fin]), np.isfinite(b[~fin])), "non-finite patterns differ"
    if not fin.any():
        return 0.0
    num = np.linalg.norm(a[fin] - b[fin], axis=1)
    den = np.maximum(np.linalg.norm(b[fin], axis=1), 1e-300)
    if floor is not None:
        fl = np.asarray(floor, np.float64)[fin]
        den = np.maximum(den, np.where(np.isfinite(fl), fl, 0.0))
    return float(np.max(num / den))
