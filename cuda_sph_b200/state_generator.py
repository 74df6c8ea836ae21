"""Frame iterator with the interface of the reference's sim/src/state_generator.py:14-40 (one SimulationState per
frame, duration * fps frames), selecting the B200 strategy (the reference hard-codes NAIVE / VOXEL at :21-23).

Unlike the reference loop, the particle state stays on the GPU between frames: the start state is uploaded once, each
frame advances `steps_per_frame` device-resident steps and downloads the result."""
from __future__ import annotations

import logging
from timeit import default_timer as timer
from typing import Optional

from .data_classes import SimulationParameters, SimulationState
from .strategy import B200SPHStrategy, SphConstants

logger = logging.getLogger(__name__)


class StateGenerator:
    def __init__(self, start_state: SimulationState, params: SimulationParameters,
                 constants: Optional[SphConstants] = None, steps_per_frame: int = 1, device: int = 0) -> None:
        self.current_state = start_state
        self.current_frame_idx = 0
        self.n_frames = params.duration * params.fps
        self.steps_per_frame = int(steps_per_frame)
        self.sph_strategy = B200SPHStrategy(params, constants, device=device)
        self.sph_strategy.upload(start_state)
        logger.info("Simulation parameters: %s", params)
        logger.info("Algorithm used: %s", self.sph_strategy.__class__)

    def __iter__(self) -> "StateGenerator":
        return self

    def __next__(self) -> SimulationState:
        if self.current_frame_idx >= self.n_frames:
            raise StopIteration
        start = timer()
        self.sph_strategy.step(self.steps_per_frame)
        self.current_state = self.sph_strategy.download()
        logger.info("frame %d / %d computed in %.4f seconds", self.current_frame_idx, self.n_frames, timer() - start)
        self.current_frame_idx += 1
        return self.current_state
