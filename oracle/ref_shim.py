"""TEST INFRASTRUCTURE — loader shim that runs the *unmodified reference kernels* under numba's CUDA simulator.

Only usable where /root/reference exists (the build container), never on the GPU box.  It is used by
``tests/golden/generate_golden.py`` to produce the committed golden vectors and by
``tests/test_oracle_vs_reference.py`` (skipped when the reference tree is absent) to pin the C restatement
in ``oracle/sph_oracle.c`` against the reference's own code.  Nothing in the product path imports this.

Why a shim is needed (nothing under /root/reference is modified or copied to disk):
  1. common/data_classes.py:70-85 declares numpy arrays / Pipe objects as dataclass defaults, which Python >= 3.11
     rejects.  We read the source text, rewrite those seven defaults to ``field(default_factory=...)`` *in memory*
     and register the result as ``sys.modules['common.data_classes']``.
  2. sim/src/sph/thread_layout.py:28-29 asks ``cuda.get_current_device().compute_capability`` which the simulator
     of numba 0.65 lacks; we provide an object reporting CC (8, 0) -> block size 64.
  3. config.py builds a PARTICLE_COUNT-sized start state with Python loops at import and has no override hook;
     we exec its source with PARTICLE_COUNT / SIM_MODE textually replaced.
  4. voxel_kernels.py:70 reads voxel_begin[n_voxels] before its bounds check (IndexError in the simulator); a
     one-element sentinel appended to the device copy of voxel_begin defeats that without changing any result
     (either way ``end = len(map)``).  voxel_kernels.py:44 leaves neigh_voxels uninitialised, so the reference is
     only runnable for particles whose 27 cells are all inside the domain.
"""
from __future__ import annotations

import importlib
import os
import re
import sys
import types

REFERENCE_ROOT = os.environ.get("SPH_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "sim/src/sph/kernels/voxel_kernels.py"))


def _purge_modules():
    for name in list(sys.modules):
        root = name.split(".")[0]
        if root in ("config", "common", "sim"):
            del sys.modules[name]


def load_reference(particle_count: int, sim_mode: str = "BOX", space_scale=None):
    """Import the reference with config.PARTICLE_COUNT / SIM_MODE overridden.

    Returns a namespace with ``config``, ``data_classes``, ``VoxelStrategy`` (sentinel subclass),
    ``NaiveSPHStrategy``, ``base_kernels``, ``voxel_kernels``, ``util_kernels``, ``PipeBuilder``, ``cuda``.
    Each call re-imports everything (kernels freeze ``config`` constants at import time).
    """
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
    from numba import cuda  # noqa: E402  (must come after the env var)

    if not hasattr(cuda, "get_current_device"):
        cuda.get_current_device = lambda: types.SimpleNamespace(compute_capability=(8, 0))

    _purge_modules()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    # (1) data classes with hashable defaults
    src = open(os.path.join(REFERENCE_ROOT, "common/data_classes.py")).read()
    src = re.sub(r"= (np\.asarray\([^\n]*?\))(\s*(#[^\n]*)?)\n", r"= field(default_factory=lambda: \1)\2\n", src)
    src = src.replace("pipe:           Pipe = Pipe([Segment()])",
                      "pipe:           Pipe = field(default_factory=lambda: Pipe([Segment()]))")
    importlib.import_module("common")
    dc = types.ModuleType("common.data_classes")
    dc.__file__ = os.path.join(REFERENCE_ROOT, "common/data_classes.py")
    sys.modules["common.data_classes"] = dc
    exec(compile(src, dc.__file__, "exec"), dc.__dict__)

    # (3) config with overridden size/mode
    csrc = open(os.path.join(REFERENCE_ROOT, "config.py")).read()
    csrc = re.sub(r"^PARTICLE_COUNT = .*$", f"PARTICLE_COUNT = {int(particle_count)}", csrc, flags=re.M)
    csrc = re.sub(r"^SIM_MODE = .*$", f"SIM_MODE = '{sim_mode}'", csrc, flags=re.M)
    csrc = re.sub(r"^SIM_STRATEGY = .*$", "SIM_STRATEGY = 'VOXEL'", csrc, flags=re.M)
    cfg = types.ModuleType("config")
    cfg.__file__ = os.path.join(REFERENCE_ROOT, "config.py")
    sys.modules["config"] = cfg
    exec(compile(csrc, cfg.__file__, "exec"), cfg.__dict__)

    import logging
    logging.disable(logging.CRITICAL)  # the reference logs at DEBUG from every stage
    voxel_mod = importlib.import_module("sim.src.sph.strategies.voxel_sph_strategy")
    naive_mod = importlib.import_module("sim.src.sph.strategies.naive_sph_strategy")
    base_kernels = importlib.import_module("sim.src.sph.kernels.base_kernels")
    voxel_kernels = importlib.import_module("sim.src.sph.kernels.voxel_kernels")
    util_kernels = importlib.import_module("sim.src.sph.kernels.util_kernels")
    pipe_builder = importlib.import_module("common.pipe_builder")
    import numpy as np

    class SentinelVoxelStrategy(voxel_mod.VoxelSPHStrategy):
        """Reference VoxelSPHStrategy + the one-element voxel_begin sentinel (item 4 above)."""

        def _initialize_computation(self):
            super()._initialize_computation()
            self.d_voxel_begin = cuda.to_device(
                np.append(self.voxel_begin, np.int32(self.params.particle_count)).astype(np.int32))

    ns = types.SimpleNamespace(
        config=cfg, data_classes=dc, VoxelStrategy=SentinelVoxelStrategy,
        NaiveSPHStrategy=naive_mod.NaiveSPHStrategy, base_kernels=base_kernels,
        voxel_kernels=voxel_kernels, util_kernels=util_kernels, PipeBuilder=pipe_builder.PipeBuilder,
        cuda=cuda, np=np)
    return ns
