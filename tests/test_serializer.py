"""Serializer format tests (restated from the reference's common/tests/test_saving_loading.py:11-56) + byte
compatibility with what the reference's own Saver writes (when the reference tree is mounted)."""
import json
import os

import numpy as np
import pytest

from cuda_sph_b200 import Pipe, Segment, SimulationParameters, SimulationState
from cuda_sph_b200.serializer import Loader, Saver


def test_params_round_trip(tmp_path):
    params = SimulationParameters(pipe=Pipe([Segment(), Segment((1.0, 0.0, 0.0), 1, 4, 5)]))
    Saver("data", params, root=str(tmp_path))
    loaded = Loader("data", root=str(tmp_path)).load_simulation_parameters()
    for name in vars(params):
        a, b = getattr(params, name), getattr(loaded, name)
        assert type(a) is type(b)
        if isinstance(a, np.ndarray):
            assert np.all(a == b)
        else:
            assert a == b


@pytest.mark.parametrize("asynchronous", [False, True])
def test_state_round_trip(tmp_path, asynchronous):
    params = SimulationParameters()
    s0 = SimulationState(position=np.asarray([0, 1, 0]))
    s1 = SimulationState(position=np.asarray([0, 2, 0]))
    saver = Saver("data", params, root=str(tmp_path), asynchronous=asynchronous)
    saver.save_next_state(s0)
    saver.save_next_state(s1)
    saver.close()
    loader = Loader("data", root=str(tmp_path))
    for true, got in zip([s0, s1], [loader.load_simulation_state(0), loader.load_simulation_state(1)]):
        for name in vars(true):
            assert np.all(getattr(true, name) == getattr(got, name))
    assert sorted(os.listdir(tmp_path / "data")) == ["density_0.npy", "density_1.npy", "params.json",
                                                     "position_0.npy", "position_1.npy", "velocity_0.npy",
                                                     "velocity_1.npy"]


def test_params_json_layout(tmp_path):
    from cuda_sph_b200 import config
    Saver("out", config.pipe_params(100), root=str(tmp_path))
    text = open(tmp_path / "out" / "params.json").read()
    d = json.loads(text)
    assert list(d) == sorted(d) == ["duration", "external_force", "fps", "particle_count", "pipe", "space_size",
                                    "voxel_size"]
    assert list(d["pipe"]["segments"][0]) == ["end_radius", "length", "start_point", "start_radius"]
    assert text.startswith('{\n    "duration"')
    Saver("box", config.box_params(100), root=str(tmp_path))
    assert json.load(open(tmp_path / "box" / "params.json"))["pipe"] == {"segments": []}


def test_bytes_equal_reference_saver(tmp_path):
    """The reference's own Saver, run live, writes byte-identical files."""
    from oracle.ref_shim import load_reference, reference_available
    if not reference_available():
        pytest.skip("reference tree not mounted")
    from cuda_sph_b200 import config
    ref = load_reference(16, "PIPE")
    ref.config.ROOT_PROJ_DIRNAME = str(tmp_path)
    import importlib
    ref_saver = importlib.import_module("common.serializer.saver")
    rng = np.random.default_rng(0)
    state = dict(position=rng.random((16, 3)), velocity=rng.random((16, 3)), density=rng.random(16))
    rs = ref_saver.Saver("ref_out", ref.config.params)
    rs.save_next_state(ref.data_classes.SimulationState(**state))
    ours = Saver("our_out", config.pipe_params(16), root=str(tmp_path))
    ours.save_next_state(SimulationState(**state))
    for f in ["params.json", "position_0.npy", "velocity_0.npy", "density_0.npy"]:
        assert open(tmp_path / "ref_out" / f, "rb").read() == open(tmp_path / "our_out" / f, "rb").read(), f
    # and the reference's Loader reads ours
    ref_loader = importlib.import_module("common.serializer.loader")
    got = ref_loader.Loader("our_out").load_simulation_state(0)
    assert np.array_equal(got.position, state["position"])
    assert ref_loader.Loader("our_out").load_simulation_parameters().particle_count == 16
