"""Fluent pipe builder with the API of the reference's common/pipe_builder.py:6-161 (same method names, argument
meaning, assertion messages and arithmetic), so configs written for the reference keep working."""
from __future__ import annotations

from typing import Optional, Tuple

from .data_classes import Pipe, Segment


class PipeBuilder:
    _FIRST_SEGMENT_MESSAGE = "First segment can't be change after adding new segments"
    _NEGATIVE_RADIUS_MESSAGE = "Radius must be positive"
    _NEGATIVE_CHANGE_MESSAGE = "Change must be positive"
    _NEGATIVE_LENGTH_MESSAGE = "Length must be positive"

    def __init__(self) -> None:
        self._segments = [Segment()]
        self._head_open = True      # the implicit first segment may only be edited before anything is appended

    # -- first segment ---------------------------------------------------------------------------------------------
    def _edit_head(self, **changes) -> "PipeBuilder":
        assert self._head_open, self._FIRST_SEGMENT_MESSAGE
        for name, value in changes.items():
            setattr(self._segments[0], name, value)
        return self

    def with_starting_position(self, position: Tuple[float, float, float]) -> "PipeBuilder":
        return self._edit_head(start_point=position)

    def with_starting_radius(self, radius: float) -> "PipeBuilder":
        assert self._head_open, self._FIRST_SEGMENT_MESSAGE
        assert radius > 0, self._NEGATIVE_RADIUS_MESSAGE
        return self._edit_head(start_radius=radius)

    def with_ending_radius(self, radius: float) -> "PipeBuilder":
        assert self._head_open, self._FIRST_SEGMENT_MESSAGE
        assert radius > 0, self._NEGATIVE_RADIUS_MESSAGE
        return self._edit_head(end_radius=radius)

    def with_starting_length(self, length: float) -> "PipeBuilder":
        assert self._head_open, self._FIRST_SEGMENT_MESSAGE
        assert length > 0, self._NEGATIVE_LENGTH_MESSAGE
        return self._edit_head(length=length)

    # -- appended segments -----------------------------------------------------------------------------------------
    def add_roller_segment(self, length) -> "PipeBuilder":
        """Cylinder continuing from the end of the previous segment."""
        self._head_open = False
        assert length > 0, self._NEGATIVE_LENGTH_MESSAGE
        prev = self._segments[-1]
        x, y, z = prev.start_point
        self._segments.append(Segment(start_point=(x + prev.length, y, z), start_radius=prev.end_radius,
                                      end_radius=prev.end_radius, length=length))
        return self

    def add_lessening_segment(self, length, change) -> "PipeBuilder":
        """Truncated cone whose end radius is smaller than its start radius by `change`."""
        self.add_roller_segment(length)
        assert change > 0, self._NEGATIVE_CHANGE_MESSAGE
        seg = self._segments[-1]
        assert change < seg.end_radius, "After change radius must be positive"
        seg.end_radius = seg.end_radius - change
        return self

    def add_increasing_segment(self, length, change) -> "PipeBuilder":
        """Truncated cone whose end radius is larger than its start radius by `change`."""
        self.add_roller_segment(length)
        assert change > 0, self._NEGATIVE_CHANGE_MESSAGE
        seg = self._segments[-1]
        seg.end_radius = seg.end_radius + change
        return self

    def transform(self, space_size_x: float, space_size_yz: float, max_radius: Optional[float] = None):
        """Rescale so the pipe spans [0, space_size_x] in x, is centred at space_size_yz/2 in y and z, and its
        largest radius becomes max_radius (default space_size_yz/2)."""
        total = 0.0
        for seg in self._segments:
            total += seg.length
        stretch = space_size_x / total
        x = 0.0
        for seg in self._segments:
            seg.start_point = (x, space_size_yz / 2.0, space_size_yz / 2.0)
            seg.length = seg.length * stretch
            x += seg.length
        if not max_radius:
            max_radius = space_size_yz / 2.0
        widest = 0.0
        for seg in self._segments:
            widest = max(widest, seg.start_radius, seg.end_radius)
        scale = max_radius / widest
        for seg in self._segments:
            seg.start_radius = seg.start_radius * scale
            seg.end_radius = seg.end_radius * scale
        return self

    def get_result(self) -> Pipe:
        return Pipe(segments=self._segments)
