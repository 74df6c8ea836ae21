"""Pins the C oracle (oracle/sph_oracle.c) against golden vectors produced by executing the reference's own numba
kernels (tests/golden/generate_golden.py), SURVEY.md Appendix A, and the known-answer values of the reference's own
sim/tests/test_collisions.py:90-228.  CPU only.

Everything is compared BITWISE (NaN == NaN): integer outputs (keys, map order, voxel_begin, neighbour lists) and
fp64 outputs alike.  The latter works because the oracle calls the same libm pow() the simulator's `**` does (these
CPU tests run on the image the goldens were generated on).
"""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import load_golden, max_rel, params_from_golden, same

RTOL = 1e-12


@pytest.fixture(autouse=True)
def _exact():
    orc.set_exact_pow(True)
    yield
    orc.set_exact_pow(True)


def test_constants_match_config():
    # config.py:27-29 via SURVEY Appendix A
    w, g, l = orc.constants(2.0)
    assert w == 0.0030599247481657124
    assert g == -0.22381163872297782
    assert l == 0.22381163872297782


def test_appendix_a_known_answer():
    """SURVEY.md Appendix A (values printed by the reference's own code)."""
    pos = np.array([[11, 11, 11], [12, 11, 11], [11, 12.5, 11], [12.25, 12.5, 11.5]], dtype=np.float64)
    vel = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, -0.5, 0.25]], dtype=np.float64)
    r = orc.step(orc.OracleParams(n=4), pos, vel)
    assert r.keys.tolist() == [2105, 2106, 2125, 2126]
    assert r.map_ids.tolist() == [0, 1, 2, 3]
    vb = np.full(8000, -1, np.int32)
    vb[[2105, 2106, 2125, 2126]] = [0, 1, 2, 3]
    assert np.array_equal(r.voxel_begin, vb)
    assert r.density.tolist() == [0.09901725239767485, 0.09299825491330689, 0.04972004189799047, 0.0411192329073577]
    np.testing.assert_allclose(r.pressure[0], [-440.38824262003044, -266.5038333879263, 0.0], rtol=1e-15)
    assert r.viscosity[3].tolist() == [1.0, 1.0, 1.0]            # the min(1, .) clamp
    np.testing.assert_allclose(r.force[2], [844.1827096539996, -317.06262801827, 295.01396396889095], rtol=1e-15)
    np.testing.assert_allclose(r.position[2], [39.999, 0.001, 25.883754795207366], rtol=1e-15)
    assert r.position[3].tolist() == [0.001, 0.001, 0.001]
    np.testing.assert_allclose(r.velocity[3], [750.9384527628847, 192.3849176947932, 350.3920003889847], rtol=1e-15)


@pytest.mark.parametrize("name,mode", [("kat4", "BOX"), ("box_dense", "BOX"), ("box_sparse", "BOX"),
                                       ("box_medium", "BOX"), ("pipe_step", "PIPE")])
def test_full_step_matches_reference_run(name, mode):
    g = load_golden(name)
    P = params_from_golden(g, mode)
    rng = g["rng_in"].copy() if mode == "PIPE" else None
    r = orc.step(P, g["pos_in"], g["vel_in"], rng=rng, want_neighbours=True)
    # integer / index outputs: identical
    assert np.array_equal(r.keys, g["keys"])
    assert np.array_equal(r.map_ids, g["map_ids"])
    assert np.array_equal(r.voxel_begin, g["voxel_begin"])
    assert np.array_equal(r.neigh_count, g["neigh_count"])
    assert np.array_equal(r.neighbours, g["neighbours"])
    # floating outputs
    for mine, ref in [(r.density, g["density"]), (r.pressure, g["pressure"]), (r.viscosity, g["viscosity"]),
                      (r.force, g["force"]), (r.velocity, g["vel_out"]), (r.position, g["pos_out"])]:
        assert max_rel(mine, ref) <= RTOL
        assert same(mine, ref)
    if mode == "PIPE":
        assert np.array_equal(rng, g["rng_out"])


def test_golden_cases_exercise_the_quirks():
    d = load_golden("box_dense")
    assert (d["neigh_count"] == 32).mean() > 0.9          # the 32 cap decides the physics
    s = load_golden("box_sparse")
    assert (s["neigh_count"] == 1).any()                   # isolated particles: rho = 0
    assert (~np.isfinite(s["pos_out"])).any()               # -> inf / NaN, as in the reference
    m = load_golden("box_medium")
    assert (m["viscosity"] == 1.0).any()                    # min(1, .) clamp
    p = load_golden("pipe_step")
    assert (p["pos_out"][:, 0] == 0.0).any()                # outlet recycle (xoroshiro draw)
    assert not np.array_equal(p["rng_in"], p["rng_out"])


def test_pipe_collision_kernel_matches_reference_run():
    g = load_golden("pipe_collide")
    n = len(g["pos_in"])
    P = orc.OracleParams(n=n, mode="PIPE", pipe=g["pipe"])
    rng = g["rng_in"].copy()
    pos, vel = orc.collide_pipe(P, g["pos_in"], g["vel_in"], rng)
    assert same(pos, g["pos_out"])
    assert same(vel, g["vel_out"])
    assert np.array_equal(rng, g["rng_out"])
    moved = (g["pos_out"] != g["pos_in"]).any(axis=1)
    assert moved.sum() > 100                                # bounces, mirrors and recycles all occur
    assert (g["pos_out"][:, 0] == 0.0).sum() > 5


def test_xoroshiro_matches_numba():
    g = load_golden("rng")
    n = len(g["states"])
    st = orc.rng_init(n, int(g["seed"]))
    assert np.array_equal(st, g["states"])
    from ctypes import c_void_p
    out = np.zeros((n, 4))
    for i in range(n):
        for q in range(4):
            out[i, q] = orc.lib().orc_rng_uniform(c_void_p(st[i].ctypes.data))
    assert np.array_equal(out, g["uniforms"])


# ---- the reference's own known-answer tests for the pipe geometry (sim/tests/test_collisions.py) -----------------
def _pipe(rows):
    return np.asarray(rows, np.float64)


def test_ref_kat_find_segment():
    # test_collisions.py:91-106: PipeBuilder().add_roller_segment(1).add_increasing_segment(1, 3)
    pipe = _pipe([[0, 0, 0, 1, 1], [1, 0, 0, 1, 1], [2, 0, 0, 1, 1], [3, 0, 0, 4, 1]])
    xs = [-0.1, 0.5, 1, 2.5, 5]
    assert [orc.find_segment(pipe, x) for x in xs] == [-1, 0, 1, 2, -1]


def test_ref_kat_vector_length_and_distance():
    # test_collisions.py:108-135
    assert orc.vector_length([0, 2, 2]) == 8 ** 0.5
    assert orc.vector_length([1, 1, 1]) == 3 ** 0.5
    assert orc.distance_between_points([1, 1, 1], [2, 3, 5]) == 21 ** 0.5
    assert orc.distance_between_points([4, 5, 6], [-2, 5, 1]) == 61 ** 0.5


def test_ref_kat_x_at_segment_beginning():
    # test_collisions.py:137-155: default segment + increasing(4, 1) + lessening(6, 1.5)
    pipe = _pipe([[0, 0, 0, 1, 1], [1, 0, 0, 1, 4], [5, 0, 0, 2, 6], [11, 0, 0, 0.5, 6]])
    assert [orc.x_at_segment_beginning(pipe, s) for s in (0, 1, 2)] == [0, 1, 5]


def test_ref_kat_is_out_of_pipe():
    # test_collisions.py:157-182: default segment + increasing(1, 1) + lessening(1, 1.5)
    pipe = _pipe([[0, 0, 0, 1, 1], [1, 0, 0, 1, 1], [2, 0, 0, 2, 1], [3, 0, 0, 0.5, 1]])
    positions = [[0.5, 0.5, 0.5], [0.76, 0.9, 0.9], [0.01, 1, 0.5], [1.25, 0.8, 0.8], [1.5, 1.25, 1.40],
                 [1.99, 1.3, 1.6], [2.2, 1, 1], [2.6, 1, 1], [2.95, 0.2, 0.2]]
    segments = [0, 0, 0, 1, 1, 1, 2, 2, 2]
    expected = [False, True, True, False, True, True, False, True, False]
    assert [orc.is_out_of_pipe(p, pipe, s) for p, s in zip(positions, segments)] == expected


def test_ref_kat_collision_resolution():
    # test_collisions.py:184-209: the single sample ends up inside the pipe
    pipe = _pipe([[0, 0, 0, 1, 1], [1, 0, 0, 1, 1], [2, 0, 0, 2, 2], [4, 0, 0, 0.5, 2]])
    pos, vel = np.array([1.3, 1.5, 0.0]), np.array([-0.5, -0.0, 0.0])
    s = orc.find_segment(pipe, pos[0])
    assert s == 1 and orc.is_out_of_pipe(pos, pipe, s)
    pos2, vel2 = orc.solve_collision(pos, vel, pipe, s)
    assert not orc.is_out_of_pipe(pos2, pipe, orc.find_segment(pipe, pos2[0]))


def test_dead_cell_policy_d1():
    """Non-finite / out-of-table positions get key n_cells, sort to the tail, have no neighbours."""
    pos = np.array([[11, 11, 11], [np.nan, 1, 1], [12, 11, 11], [1e30, 0, 0], [-9, 1, 1], [39.9, 39.9, 39.9]])
    vel = np.zeros_like(pos)
    P = orc.OracleParams(n=len(pos))
    r = orc.step(P, pos, vel)
    nc = orc.n_cells(P)
    assert nc == 8000
    assert r.keys.tolist() == [2105, nc, 2106, nc, nc, 7999]
    assert r.map_ids.tolist() == [0, 2, 5, 1, 3, 4]
    assert r.n_dead == 3
    assert r.neigh_count.tolist() == [2, 0, 2, 0, 0, 1]
    assert r.density[[1, 3, 4, 5]].tolist() == [0, 0, 0, 0]
