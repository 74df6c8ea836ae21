"""GPU tests of the rows next to the hot path: frame export pipeline + down-sampling, seeded
device-side start states, resume from saved frames (+ RNG states), per-frame on-device reductions.  Each is checked
against something other than itself: the plain download path, the host mirror of the generators, an uninterrupted run,
numpy reductions over the downloaded state / the reference-format files read back by Loader."""
import json

import numpy as np
import pytest

from tests.helpers import same

pytestmark = pytest.mark.gpu


def _gen(params, start, mode, **kw):
    from cuda_sph_b200 import config
    from cuda_sph_b200.state_generator import StateGenerator
    return StateGenerator(start, params, config.constants(mode), **kw)


@pytest.mark.parametrize("stride", [1, 7])
def test_export_pipeline_equals_download(stride):
    """Frames through sph_export_begin / sph_export_wait (pinned triple buffer, one frame of look-ahead) == the state a
    plain step + download loop produces, every frame, also down-sampled (every 7th id)."""
    from cuda_sph_b200 import B200SPHStrategy, config
    n = 30000
    params = config.box_params(n, duration=1, fps=8)
    start = config.start_state_box_wall(n, params.space_size, seed=91)
    ref = B200SPHStrategy(params, config.constants("BOX"))
    ref.upload(start)
    frames = 0
    for state in _gen(params, start, "BOX", steps_per_frame=2, export_stride=stride):
        ref.step(2)
        want = ref.download()
        assert state.position.dtype == np.float64 and state.position.flags.c_contiguous
        assert same(state.position, want.position[::stride]) and same(state.velocity, want.velocity[::stride])
        assert same(state.density, want.density[::stride])
        frames += 1
    assert frames == 8
    ref.close()


@pytest.mark.parametrize("kind,mode", [("box_wall", "BOX"), ("uniform", "BOX"), ("pipe", "PIPE")])
def test_device_generators_equal_host_mirror(kind, mode):
    """sph_generate_state (counter-based draws on the GPU) == config.hashed_start_state (numpy) bit for bit; the states
    respect config.py:84-95 / :105-115 (column in the first 10 % of x; inside 98 % of the local pipe radius)."""
    from cuda_sph_b200 import B200SPHStrategy, config
    n = 50000
    params = config.pipe_params(n) if mode == "PIPE" else config.box_params(n)
    s = B200SPHStrategy(params, config.constants(mode))
    for seed in (0, 12345):
        s.generate_state(kind, seed)
        got = s.download()
        want = config.hashed_start_state(kind, params, seed)
        assert np.array_equal(got.position, want.position) and np.array_equal(got.velocity, want.velocity)
    space = np.asarray(params.space_size)
    assert (got.position >= 0).all() and (got.position < space).all()
    if kind == "box_wall":
        assert got.position[:, 0].max() <= 0.1 * space[0]
        assert np.abs(got.velocity - np.array([1.5, -5.0, -5.0])).max() <= 0.5
    if kind == "pipe":
        t = params.pipe.to_numpy()
        r = np.hypot(got.position[:, 1] - t[0, 1], got.position[:, 2] - t[0, 2])
        seg = np.clip(np.searchsorted(t[:-1, 0], got.position[:, 0], side="right") - 1, 0, len(t) - 2)
        r_loc = t[seg, 3] + (t[seg + 1, 3] - t[seg, 3]) * (got.position[:, 0] - t[seg, 0]) / t[seg, 4]
        assert (r <= 0.98 * r_loc + 1e-5).all() and not np.any(got.velocity)
    s.step(1)   # the generated state is a valid start state
    assert s.stats()["n_nonfinite"] == 0
    s.close()


@pytest.mark.parametrize("mode", ["BOX", "PIPE"])
def test_resume_from_saved_frames(tmp_path, mode):
    """A run interrupted after frame K and resumed from the files Saver wrote (Loader.load_simulation_state(K) + the RNG
    states of the checkpoint frame) writes the same remaining frames, bit for bit, as the uninterrupted run."""
    from cuda_sph_b200 import main
    from cuda_sph_b200.serializer import Loader
    common = ["--mode", mode, "-n", "6000", "--fps", "6", "--root", str(tmp_path), "--checkpoint-every", "3",
              "--preview", "1000", "--stats", "--device-start", "--seed", "5"]
    main.main(common + ["--duration", "1", "--out", "full"])
    main.main(common + ["--duration", "1", "--out", "part"])              # writes frames 0..5 as well ...
    main.main(common + ["--duration", "1", "--out", "part", "--resume", "2"])   # ... and 3..5 again from frame 2
    full, part = Loader("full", root=str(tmp_path)), Loader("part", root=str(tmp_path))
    for k in range(6):
        a, b = full.load_simulation_state(k), part.load_simulation_state(k)
        assert same(a.position, b.position) and same(a.velocity, b.velocity) and same(a.density, b.density), k
    if mode == "PIPE":
        assert full.load_rng_states(2) is not None and full.load_rng_states(1) is None
        assert np.array_equal(full.load_rng_states(5), part.load_rng_states(5))
    # down-sampled copy for the viewer: own params.json, every 6th particle
    prev = Loader("full_preview", root=str(tmp_path))
    assert prev.load_simulation_parameters().particle_count == 1000
    assert same(prev.load_simulation_state(4).position, full.load_simulation_state(4).position[::6])
    # per-frame statistics match numpy over the saved frames (what analize.py prints)
    lines = [json.loads(ln) for ln in open(tmp_path / "full" / "stats.jsonl")]
    assert [ln["epoch"] for ln in lines] == list(range(6))
    for ln in lines:
        st = full.load_simulation_state(ln["epoch"])
        fin = np.isfinite(st.position).all(axis=1) & np.isfinite(st.velocity).all(axis=1)
        assert ln["n_nonfinite"] == int((~fin).sum())
        assert ln["max_position"] == np.float32(st.position[fin].max())
        assert ln["max_velocity"] == np.float32(st.velocity[fin].max())
        assert ln["max_density"] == np.float32(st.density[np.isfinite(st.density)].max())
        assert sum(ln["neighbour_hist"]) == 6000


def test_frame_stats_match_numpy_and_neighbour_counts():
    from cuda_sph_b200 import B200SPHStrategy, SphConstants, workloads
    n = 40000
    params, st = workloads.dam_break(n, 2.5, seed=93)
    s = B200SPHStrategy(params, SphConstants(mode="BOX"), record_neighbour_counts=True)
    s.upload(st)
    s.step(6)            # by now the reference's physics has produced some non-finite particles
    fs, out, cnt = s.frame_stats(), s.download(np.float32), s.neighbour_counts()
    fin = np.isfinite(out.position).all(axis=1) & np.isfinite(out.velocity).all(axis=1)
    assert fs["n_nonfinite"] == int((~fin).sum()) and fs["n_particles"] == n and fs["steps_done"] == 6
    assert fs["max_position"] == out.position[fin].max() and fs["min_position"] == out.position[fin].min()
    assert fs["max_velocity"] == out.velocity[fin].max()
    assert np.isclose(fs["max_speed"], np.linalg.norm(out.velocity[fin].astype(np.float64), axis=1).max(), rtol=1e-6)
    assert fs["max_density"] == out.density[np.isfinite(out.density)].max()
    assert fs["neighbour_hist"] == np.bincount(cnt, minlength=33).tolist()
    assert fs["n_dead"] == s.stats()["n_dead"]
    s.close()
