"""CPU-only checks: the C-ABI library loads and exports every symbol include/sph_b200.h declares; the host-side mirror
of the reference's interface (data classes, PipeBuilder, config) behaves like the reference's (expectations restated
from common/tests/test_pipe.py, common/tests/test_pipe_builder.py); the product fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.helpers import load_golden

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def built_lib():
    from cuda_sph_b200 import build
    return build.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sph_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sph_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    from cuda_sph_b200 import _lib
    lib = ctypes.CDLL(built_lib)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sph_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared          # the ctypes binding covers exactly the header


def test_no_cpu_fallback(built_lib):
    """Without a GPU the product path must fail loudly (it never routes through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cuda_sph_b200 import B200SPHStrategy, _lib, config
    with pytest.raises(_lib.SphError, match="no CUDA device"):
        B200SPHStrategy(config.box_params(128))


def test_product_package_does_not_import_the_oracle():
    import importlib
    import sys
    for name in list(sys.modules):
        if name.startswith("cuda_sph_b200"):
            del sys.modules[name]
    before = {m for m in sys.modules if m.startswith("oracle")}
    importlib.import_module("cuda_sph_b200")
    importlib.import_module("cuda_sph_b200.state_generator")
    importlib.import_module("cuda_sph_b200.serializer")
    after = {m for m in sys.modules if m.startswith("oracle")}
    assert after == before
    importlib.import_module("cuda_sph_b200.slab")
    assert {m for m in sys.modules if m.startswith("oracle")} == before
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|#include.*oracle|liborc", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cuda_sph_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_params_struct_layout_matches_header(built_lib):
    from cuda_sph_b200 import _lib
    # int32 x2, double x7, double[3] x3, int32, uint32, uint64
    assert ctypes.sizeof(_lib.SphParams) == 8 + 7 * 8 + 9 * 8 + 8 + 8
    assert _lib.SphParams.h.offset == 8 and _lib.SphParams.rng_seed.offset == 8 + 56 + 72 + 8


# ---- data classes / pipe table: common/tests/test_pipe.py:12-34 ------------------------------------------------------
def test_segment_and_pipe_to_numpy():
    from cuda_sph_b200 import Pipe, PipeBuilder, Segment
    assert np.all(Segment(start_radius=4.5).to_numpy() == np.array([0, 0, 0, 4.5, 1]))
    assert np.all(Pipe([Segment()]).to_numpy() == np.array([[0, 0, 0, 1, 1], [1, 0, 0, 1, 1]]))
    pipe = PipeBuilder().with_starting_radius(3).add_roller_segment(1).add_increasing_segment(1, 2).get_result()
    assert np.all(pipe.to_numpy() == np.array([[0, 0, 0, 3, 1], [1, 0, 0, 1, 1], [2, 0, 0, 1, 1], [3, 0, 0, 3, 1]]))
    assert Pipe().to_numpy().size == 0


# ---- PipeBuilder: common/tests/test_pipe_builder.py:10-133 -----------------------------------------------------------
def test_builder_semantics():
    from cuda_sph_b200 import Pipe, PipeBuilder, Segment
    assert PipeBuilder().get_result() == Pipe([Segment()])
    p = (PipeBuilder().with_starting_position((1, 2, 3)).with_starting_length(20).with_starting_radius(4)
         .with_ending_radius(5).get_result())
    assert p == Pipe([Segment(start_point=(1, 2, 3), start_radius=4, end_radius=5, length=20)])
    assert PipeBuilder().with_ending_radius(4).add_roller_segment(30).get_result() == \
        Pipe([Segment(end_radius=4), Segment((1, 0, 0), 4, 4, 30)])
    assert PipeBuilder().with_ending_radius(7).add_lessening_segment(length=1, change=6).get_result() == \
        Pipe([Segment(end_radius=7), Segment((1, 0, 0), 7, 1, 1)])
    assert PipeBuilder().with_ending_radius(5).add_increasing_segment(length=1, change=5).get_result() == \
        Pipe([Segment(end_radius=5), Segment((1, 0, 0), 5, 10, 1)])
    many = (PipeBuilder().with_starting_radius(10).with_ending_radius(5).add_roller_segment(length=1)
            .add_increasing_segment(1, 5).get_result())
    assert many == Pipe([Segment(start_radius=10, end_radius=5), Segment((1, 0, 0), 5, 5, 1),
                         Segment((2, 0, 0), 5, 10, 1)])


def test_builder_errors():
    from cuda_sph_b200 import PipeBuilder
    with pytest.raises(AssertionError, match=PipeBuilder._NEGATIVE_LENGTH_MESSAGE):
        PipeBuilder().with_starting_length(-10)
    with pytest.raises(AssertionError, match=PipeBuilder._NEGATIVE_RADIUS_MESSAGE):
        PipeBuilder().with_starting_radius(-1)
    with pytest.raises(AssertionError, match=PipeBuilder._NEGATIVE_RADIUS_MESSAGE):
        PipeBuilder().with_ending_radius(-1)
    with pytest.raises(AssertionError, match="First segment"):
        PipeBuilder().add_roller_segment(20).with_starting_radius(1)
    with pytest.raises(AssertionError, match=PipeBuilder._NEGATIVE_CHANGE_MESSAGE):
        PipeBuilder().add_lessening_segment(1, -10)
    with pytest.raises(AssertionError):
        PipeBuilder().add_lessening_segment(1, 5)       # radius would become negative


def test_builder_transform():
    from cuda_sph_b200 import PipeBuilder
    pipe = (PipeBuilder().add_increasing_segment(2., 3.).add_roller_segment(1.).add_increasing_segment(2., 3.)
            .transform(600, 600, 70).get_result())
    expected = (PipeBuilder().with_starting_position((0., 300., 300.)).with_starting_radius(10.)
                .with_ending_radius(10.).with_starting_length(100.).add_increasing_segment(200., 30.)
                .add_roller_segment(100.).add_increasing_segment(200., 30.).get_result().to_numpy())
    assert np.all(pipe.to_numpy() == expected)


def test_config_pipe_equals_reference_config_pipe():
    """Our config.build_pipe() table is bitwise the table the reference's config.py:68-76 produced (golden)."""
    from cuda_sph_b200 import config
    g = load_golden("pipe_step")
    assert np.array_equal(config.build_pipe().to_numpy(), g["pipe"])
    assert np.array_equal(np.asarray(config.PIPE_SPACE_SIZE), g["space"])
    assert np.array_equal(np.asarray(config.HORIZONTAL_FORCE), g["ext"])
    b = load_golden("box_dense")
    assert np.array_equal(np.asarray(config.BOX_SPACE_SIZE), b["space"])
    assert np.array_equal(np.asarray(config.GRAVITY), b["ext"])
    assert config.FPS == int(b["fps"])


def test_config_constants():
    from cuda_sph_b200 import config
    assert config.W_CONST == 0.0030599247481657124
    assert config.GRAD_W_CONST == -0.22381163872297782
    assert config.LAP_W_CONST == 0.22381163872297782
    c = config.constants("PIPE")
    assert (c.h, c.mass, c.rho0, c.visc, c.k, c.damp, c.max_neighbours) == (2.0, 1.0, 1.0, 0.5, 10.0, 0.7, 32)


def test_start_states_are_seeded_and_in_domain():
    from cuda_sph_b200 import config, workloads
    a = config.start_state_box_wall(5000, seed=3)
    b = config.start_state_box_wall(5000, seed=3)
    assert np.array_equal(a.position, b.position) and np.array_equal(a.velocity, b.velocity)
    assert a.position[:, 0].max() <= 4.0 and a.position.min() >= 0 and a.position[:, 1:].max() <= 40.0
    assert np.array_equal(a.position, a.position.astype(np.float32).astype(np.float64))   # fp32-representable
    assert np.abs(a.velocity - np.array([1.5, -5, -5])).max() <= 0.5
    pipe = config.build_pipe()
    st = config.start_state_inside_pipe(4000, pipe, seed=1)
    r = np.hypot(st.position[:, 1] - 3.0, st.position[:, 2] - 3.0)
    rmax = np.array([pipe.radius_at(x) for x in st.position[:, 0]])
    assert (r <= rmax).all() and (st.velocity == 0).all()
    params, s1 = workloads.uniform_box(10000, 2.5, seed=0)
    assert params.space_size[0] == 2.0 * workloads.cubic_dims(10000, 2.5)
    assert (s1.position >= 0).all() and (s1.position < params.space_size[0]).all()
    params, s3 = workloads.pipe_flow(1 << 15, seed=0)
    assert len(params.pipe.segments) == 6 and params.space_size[1] == params.space_size[2]


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path = the oracle port, all host threads) needs no GPU and prints
    ONE JSON line with the contract keys; under torchrun only rank 0 works."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--particles", "20000", "--steps", "2",
           "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="0"))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-updates/sec" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # any other rank exits 0 without output
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1"))
    assert out.returncode == 0 and not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
