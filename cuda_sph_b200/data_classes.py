"""Data model that crosses the strategy boundary -- same names, fields and array layouts as the reference's
common/data_classes.py:7-85, written so that it imports on Python >= 3.11 (the reference's mutable ndarray defaults do
not).  Objects of the reference's own classes are accepted everywhere by duck typing."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np


@dataclass
class Segment:
    """One pipe segment: a cylinder (start_radius == end_radius) or a truncated cone."""
    start_point: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    start_radius: float = 1.0
    end_radius: float = 1.0
    length: float = 1.0

    def to_numpy(self) -> np.ndarray:
        # [x, y, z, start_radius, length]  (reference common/data_classes.py:14-17)
        return np.array([*self.start_point, self.start_radius, self.length], dtype=np.float64)

    def radius_at(self, t: float) -> float:
        if self.start_radius == self.end_radius:
            return self.start_radius
        frac = (t - self.start_point[0]) / self.length
        return self.start_radius + (self.end_radius - self.start_radius) * frac


@dataclass(frozen=True)
class Pipe:
    segments: List[Segment] = field(default_factory=list)

    def to_numpy(self) -> np.ndarray:
        """(S+1, 5) table; the extra last row is [x_end, y, z, end_radius_of_last, length_of_last]
        (reference common/data_classes.py:34-44).  Empty pipe -> empty array."""
        if not self.segments:
            return np.asarray([])
        rows = [s.to_numpy() for s in self.segments]
        tail = self.segments[-1].to_numpy()
        tail[0] += tail[4]
        tail[3] = self.segments[-1].end_radius
        rows.append(tail)
        return np.stack(rows)

    def get_length(self) -> float:
        return float(sum(s.length for s in self.segments))

    def find_segment(self, t: float) -> int:
        for i, s in enumerate(self.segments):
            if s.start_point[0] <= t <= s.start_point[0] + s.length:
                return i
        return -1

    def radius_at(self, t: float) -> float:
        return self.segments[self.find_segment(t)].radius_at(t)


@dataclass(frozen=True)
class SimulationParameters:
    particle_count: int = 100
    external_force: np.ndarray = field(default_factory=lambda: np.asarray([0, 0, 0]))
    duration: int = 10
    fps: int = 20
    pipe: Pipe = field(default_factory=lambda: Pipe([Segment()]))
    space_size: np.ndarray = field(default_factory=lambda: np.asarray([1, 1, 1]))
    voxel_size: np.ndarray = field(default_factory=lambda: np.asarray([1, 1, 1]))


@dataclass(frozen=True)
class SimulationState:
    position: np.ndarray = field(default_factory=lambda: np.asarray([[1, 1, 1]]))  # (n, 3)
    velocity: np.ndarray = field(default_factory=lambda: np.asarray([[1, 1, 1]]))  # (n, 3)
    density: np.ndarray = field(default_factory=lambda: np.asarray([1]))           # (n,)
