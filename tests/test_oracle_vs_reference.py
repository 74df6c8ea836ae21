"""Runs the REFERENCE ITSELF (numba CUDA simulator, oracle/ref_shim.py) next to the C oracle on fresh random inputs and
demands bitwise equality.  Only possible where /root/reference exists (the build container); skipped elsewhere."""
import warnings

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.ref_shim import load_reference, reference_available
from tests.helpers import same

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")


@pytest.mark.parametrize("seed,n,lo,hi", [(11, 96, 18.0, 22.0), (12, 150, 16.0, 24.0)])
def test_live_reference_box_step_bitwise(seed, n, lo, hi):
    warnings.filterwarnings("ignore")
    orc.set_exact_pow(True)
    rng = np.random.default_rng(seed)
    pos = rng.uniform(lo, hi, (n, 3)).astype(np.float32).astype(np.float64)
    vel = (np.array([1.5, -5.0, -5.0]) + rng.uniform(-0.5, 0.5, (n, 3))).astype(np.float32).astype(np.float64)
    ref = load_reference(n, "BOX")
    strat = ref.VoxelStrategy(ref.config.params)
    out = strat.compute_next_state(ref.data_classes.SimulationState(pos.copy(), vel.copy(), np.zeros(n)))
    r = orc.step(orc.OracleParams(n=n), pos, vel)
    assert np.array_equal(r.keys, strat.voxels)
    assert np.array_equal(r.map_ids, strat.voxel_particle_map["particle_id"])
    assert np.array_equal(r.voxel_begin, strat.voxel_begin)
    assert same(r.density, out.density)
    assert same(r.force, strat.result_force)
    assert same(r.velocity, out.velocity)
    assert same(r.position, out.position)
