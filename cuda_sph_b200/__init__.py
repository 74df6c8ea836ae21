"""B200-native SPH step engine behind the simulator-strategy interface of iwoplaza/cuda-sph.

Product code.  The hot path is hand-written sm_100a CUDA in csrc/ behind the C ABI of include/sph_b200.h; this
package is the host-side mirror of the reference's strategy interface (sim/src/sph/strategies).
"""
from .data_classes import Pipe, Segment, SimulationParameters, SimulationState  # noqa: F401
from .pipe_builder import PipeBuilder  # noqa: F401
from .strategy import B200SPHStrategy, SphConstants  # noqa: F401

__all__ = ["B200SPHStrategy", "SphConstants", "Pipe", "Segment", "SimulationParameters", "SimulationState",
           "PipeBuilder"]
