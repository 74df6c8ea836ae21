"""Frame iterator with the interface of the reference's sim/src/state_generator.py:14-40 (one SimulationState per
frame, duration * fps frames), selecting the B200 strategy (the reference hard-codes NAIVE / VOXEL at :21-23).

Unlike the reference loop, the particle state stays on the GPU between frames and the frames leave through the export
pipeline of the engine (sph_export_begin / sph_export_wait): while the caller consumes frame k
the GPU already computes frame k + 1, and the device -> pinned-host copy of frame k ran under those steps.  The fp64
cast happens on the device.  `export_stride` > 1 down-samples the frames (every stride-th particle id), e.g. for the
viewer's 100 000-point cap (vis/src/opengl/scene_components/gl_point_field.py:11).

Resume (section 8(f)3): `first_frame` / `rng_states` continue a run from a frame written earlier (Loader + the saved
xoroshiro states, see serializer.Saver)."""
from __future__ import annotations

import logging
from timeit import default_timer as timer
from typing import Optional

from .data_classes import SimulationParameters, SimulationState
from .strategy import B200SPHStrategy, SphConstants

logger = logging.getLogger(__name__)


class StateGenerator:
    def __init__(self, start_state: Optional[SimulationState], params: SimulationParameters,
                 constants: Optional[SphConstants] = None, steps_per_frame: int = 1, device: int = 0, *,
                 export_stride: int = 1, zero_copy: bool = False, first_frame: int = 0, rng_states=None,
                 generate: Optional[tuple] = None, checkpoint_every: int = 0) -> None:
        """start_state: host state to upload, or None with generate=(kind, seed) to create it on the device
        (B200SPHStrategy.generate_state).  zero_copy=True hands out views of the pinned export buffers, valid until the
        generator has been advanced twice more.  checkpoint_every=k (PIPE mode): every k-th frame carries the xoroshiro
        states after that frame as `state.rng_states` (the pipeline drains there), which Saver writes next to the frame
        so that a run can be resumed from it."""
        self.current_state = start_state
        self.current_frame_idx = int(first_frame)
        self.n_frames = params.duration * params.fps
        self.steps_per_frame = int(steps_per_frame)
        self.export_stride = int(export_stride)
        self.zero_copy = bool(zero_copy)
        self.checkpoint_every = int(checkpoint_every)
        self._pipe_mode = (constants.mode if constants else "BOX").upper() == "PIPE"
        self.sph_strategy = B200SPHStrategy(params, constants, device=device)
        if start_state is not None:
            self.sph_strategy.upload(start_state)
        elif generate is not None:
            self.sph_strategy.generate_state(*generate)
        else:
            raise ValueError("StateGenerator needs a start state or generate=(kind, seed)")
        if rng_states is not None:
            self.sph_strategy.set_rng_states(rng_states)
        self._issued = self.current_frame_idx     # frames whose steps + export have been enqueued
        logger.info("Simulation parameters: %s", params)
        logger.info("Algorithm used: %s", self.sph_strategy.__class__)

    def __iter__(self) -> "StateGenerator":
        return self

    def _issue(self) -> None:
        """Enqueue the steps of the next frame and its export (asynchronous)."""
        self.sph_strategy.step(self.steps_per_frame)
        self.sph_strategy.export_async(self._issued % 3, self.export_stride)
        self._issued += 1

    def __next__(self) -> SimulationState:
        if self.current_frame_idx >= self.n_frames:
            raise StopIteration
        start = timer()
        idx = self.current_frame_idx
        if self._issued == idx:
            self._issue()
        ckpt = self._pipe_mode and self.checkpoint_every > 0 and (idx + 1) % self.checkpoint_every == 0
        if not ckpt and self._issued < self.n_frames:   # run one frame ahead: its steps overlap this frame's copy + consumer
            self._issue()
        self.current_state = self.sph_strategy.export_wait(idx % 3, copy=not self.zero_copy)
        self.current_stats = self.sph_strategy.export_stats(idx % 3)   # reductions of exactly this frame
        if ckpt:   # nothing beyond this frame has been enqueued: the device holds the RNG states after frame idx
            # (SimulationState is a frozen dataclass like the reference's; the extra attribute is picked up by Saver)
            object.__setattr__(self.current_state, "rng_states", self.sph_strategy.rng_states())
            if self._issued < self.n_frames:
                self._issue()
        logger.info("frame %d / %d ready after %.4f seconds", idx, self.n_frames, timer() - start)
        self.current_frame_idx += 1
        return self.current_state
