"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI (libsph_b200.so).

Gates (BASELINE.json north_star / SURVEY.md section 8d): bit-exact cell keys, sorted (key, id) order, voxel_begin and
neighbour counts; density relative error <= 1e-4; force / velocity / position per-particle ||delta|| / ||ref|| <= 1e-4
for one step (fp32 engine vs the fp64 reference), NaN == NaN.  Typical measured errors are 1e-7..1e-6; the worst case
for density (1.3e-5) is a particle whose only neighbour sits just inside the cut-off, where (h^2 - r^2)^3 cancels.
"""
import numpy as np
import pytest

from tests.helpers import load_golden, max_rel, params_from_golden, same, vec_rel

pytestmark = pytest.mark.gpu

RHO_RTOL = 1e-4
VEC_RTOL = 1e-4


def _strategy(n, mode, space, voxel, ext, fps, pipe_table=None, **kw):
    from cuda_sph_b200 import B200SPHStrategy, SphConstants
    from cuda_sph_b200.data_classes import Pipe, SimulationParameters

    class _RawPipe(Pipe):
        def to_numpy(self):
            return pipe_table

    pipe = _RawPipe() if pipe_table is not None else Pipe()
    params = SimulationParameters(particle_count=n, external_force=np.asarray(ext, float), duration=1, fps=int(fps),
                                  pipe=pipe, space_size=np.asarray(space, float), voxel_size=np.asarray(voxel, float))
    return B200SPHStrategy(params, SphConstants(mode=mode), record_neighbour_counts=True, record_terms=True, **kw)


def _check_against(s, ref, *, rho=RHO_RTOL, vec=VEC_RTOL):
    """ref: dict with keys/map_ids/voxel_begin/neigh_count/density/pressure/viscosity/force/vel_out/pos_out."""
    out = s.new_state
    assert np.array_equal(s.keys(), ref["keys"])
    assert np.array_equal(s.sorted_ids(), ref["map_ids"])
    assert np.array_equal(s.voxel_begin(), ref["voxel_begin"])
    assert np.array_equal(s.neighbour_counts(), ref["neigh_count"])
    assert max_rel(out.density, ref["density"]) <= rho
    pr, vi = s.terms()
    fnorm = np.linalg.norm(ref["force"], axis=1)
    assert vec_rel(pr, ref["pressure"], floor=fnorm) <= vec
    assert vec_rel(vi, ref["viscosity"], floor=fnorm) <= vec
    assert vec_rel(s.result_force, ref["force"]) <= vec
    assert vec_rel(out.velocity, ref["vel_out"]) <= vec
    assert vec_rel(out.position, ref["pos_out"]) <= vec


@pytest.mark.parametrize("name,mode", [("kat4", "BOX"), ("box_dense", "BOX"), ("box_sparse", "BOX"),
                                       ("box_medium", "BOX"), ("pipe_step", "PIPE")])
def test_one_step_matches_reference_golden(name, mode):
    """CUDA engine vs outputs of the reference's own kernels (tests/golden)."""
    from cuda_sph_b200.data_classes import SimulationState
    g = load_golden(name)
    n = len(g["pos_in"])
    s = _strategy(n, mode, g["space"], g["voxel"], g["ext"], g["fps"], g["pipe"] if mode == "PIPE" else None)
    if mode == "PIPE":
        assert np.array_equal(s.rng_states(), g["rng_in"])          # xoroshiro init == numba's
    s.compute_next_state(SimulationState(g["pos_in"], g["vel_in"], np.zeros(n)))
    _check_against(s, g)
    if mode == "PIPE":
        assert np.array_equal(s.rng_states(), g["rng_out"])
    s.close()


def _oracle_ref(P, pos, vel, rng=None):
    from oracle import oracle as orc
    orc.set_exact_pow(False)
    r = orc.step(P, pos, vel, rng=rng)
    return dict(keys=r.keys, map_ids=r.map_ids, voxel_begin=r.voxel_begin, neigh_count=r.neigh_count,
                density=r.density, pressure=r.pressure, viscosity=r.viscosity, force=r.force, vel_out=r.velocity,
                pos_out=r.position)


@pytest.mark.parametrize("n,ppc,seed", [(20000, 2.5, 0), (30000, 8.0, 1), (40000, 25.0, 2)])
def test_uniform_box_vs_oracle(n, ppc, seed):
    """S1 incl. boundary cells (where the reference itself cannot run: SURVEY Q3) vs the C oracle."""
    from cuda_sph_b200 import workloads
    from oracle import oracle as orc
    params, st = workloads.uniform_box(n, ppc, seed)
    s = _strategy(n, "BOX", params.space_size, params.voxel_size, params.external_force, params.fps)
    s.compute_next_state(st)
    P = orc.OracleParams(n=n, space=tuple(params.space_size), dt=1 / params.fps)
    _check_against(s, _oracle_ref(P, st.position, st.velocity))
    s.close()


def test_dam_break_column_vs_oracle():
    """S2: ~25 particles per cell, nearly all lists capped at 32 -> traversal order decides the result."""
    from cuda_sph_b200 import workloads
    from oracle import oracle as orc
    n = 60000
    params, st = workloads.dam_break(n, 2.5, seed=3)
    s = _strategy(n, "BOX", params.space_size, params.voxel_size, params.external_force, params.fps)
    s.compute_next_state(st)
    P = orc.OracleParams(n=n, space=tuple(params.space_size), dt=1 / params.fps)
    ref = _oracle_ref(P, st.position, st.velocity)
    assert (ref["neigh_count"] == 32).mean() > 0.9
    _check_against(s, ref)
    s.close()


def test_pipe_flow_vs_oracle():
    """S3: multi-segment pipe, wall bounces, inlet mirror and outlet recycle (xoroshiro) vs the C oracle."""
    from cuda_sph_b200 import workloads
    from cuda_sph_b200.data_classes import SimulationState
    from oracle import oracle as orc
    n = 30000
    params, st = workloads.pipe_flow(n, seed=4)
    rng = np.random.default_rng(5)
    vel = rng.uniform(-30, 30, (n, 3)).astype(np.float32).astype(np.float64)
    vel[: n // 20, 0] = 2000.0            # leave through the outlet
    vel[n // 20: n // 10, 0] = -2000.0    # leave through the inlet
    st = SimulationState(st.position, vel, st.density)
    table = params.pipe.to_numpy()
    s = _strategy(n, "PIPE", params.space_size, params.voxel_size, params.external_force, params.fps, table)
    s.compute_next_state(st)
    P = orc.OracleParams(n=n, mode="PIPE", space=tuple(params.space_size), ext=tuple(params.external_force),
                         dt=1 / params.fps, pipe=table)
    orng = orc.rng_init(n)
    ref = _oracle_ref(P, st.position, st.velocity, rng=orng)
    assert (ref["pos_out"][:, 0] == 0.0).sum() > 100
    _check_against(s, ref)
    assert np.array_equal(s.rng_states(), orng)
    s.close()


def test_dead_cell_policy_and_nan_semantics():
    """DESIGN.md D1 + reference NaN semantics (isolated particle: rho = 0 -> v = +-inf / NaN)."""
    from cuda_sph_b200.data_classes import SimulationState
    from oracle import oracle as orc
    pos = np.array([[11, 11, 11], [np.nan, 1, 1], [12, 11, 11], [1e30, 0, 0], [-9, 1, 1], [39.9, 39.9, 39.9]])
    vel = np.zeros_like(pos)
    n = len(pos)
    s = _strategy(n, "BOX", [40, 40, 40], [2, 2, 2], [0, -2, 0], 20)
    s.compute_next_state(SimulationState(pos, vel, np.zeros(n)))
    ref = _oracle_ref(orc.OracleParams(n=n), pos, vel)
    _check_against(s, ref)
    assert s.stats()["n_dead"] == 3
    assert s.keys().tolist() == [2105, 8000, 2106, 8000, 8000, 7999]
    s.close()


def test_device_resident_steps_equal_host_round_trips():
    """step(n) without host round trips == n x compute_next_state (the reference's loop)."""
    from cuda_sph_b200 import workloads
    n = 20000
    params, st = workloads.dam_break(n, 2.5, seed=6)
    a = _strategy(n, "BOX", params.space_size, params.voxel_size, params.external_force, params.fps)
    b = _strategy(n, "BOX", params.space_size, params.voxel_size, params.external_force, params.fps, use_graph=False)
    a.upload(st)
    a.step(5)
    out_a = a.download()
    cur = st
    for _ in range(5):
        cur = b.compute_next_state(cur)
    assert same(out_a.position, cur.position)
    assert same(out_a.velocity, cur.velocity)
    assert same(out_a.density, cur.density)
    a.close()
    b.close()


def test_multi_step_drift_vs_oracle():
    """Short-horizon drift (reported, loosely bounded): 10 steps of the dam-break column, fp32 engine vs fp64 oracle."""
    from cuda_sph_b200 import workloads
    from oracle import oracle as orc
    orc.set_exact_pow(False)
    n = 20000
    params, st = workloads.dam_break(n, 2.5, seed=7)
    s = _strategy(n, "BOX", params.space_size, params.voxel_size, params.external_force, params.fps)
    P = orc.OracleParams(n=n, space=tuple(params.space_size), dt=1 / params.fps)
    s.upload(st)
    pos, vel = st.position, st.velocity
    report = []
    for step in range(1, 11):
        s.step(1)
        r = orc.step(P, pos, vel, light=True)
        pos, vel = r.position, r.velocity
        got = s.download()
        fin = np.isfinite(pos).all(axis=1) & np.isfinite(got.position).all(axis=1)
        d = np.linalg.norm(got.position[fin] - pos[fin], axis=1) / 2.0
        report.append((step, float(d.max()), float(np.median(d)), int((~fin).sum())))
    print("drift (step, max |dx|/h, median |dx|/h, non-finite):", report)
    assert report[0][1] < 1e-4
    s.close()


def test_full_size_properties_1m():
    """BASELINE config 2 size (1M dam-break): size-independent invariants of the grid / sort output."""
    from cuda_sph_b200 import workloads
    n = 1 << 20
    params, st = workloads.dam_break(n, 2.5, seed=8)
    s = _strategy(n, "BOX", params.space_size, params.voxel_size, params.external_force, params.fps)
    s.upload(st)
    s.step(1)
    keys, ids, skeys, vb, cnt = s.keys(), s.sorted_ids(), s.sorted_keys(), s.voxel_begin(), s.neighbour_counts()
    # the map is a permutation, sorted by (key, id)
    assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.int32))
    assert np.array_equal(skeys, keys[ids])
    order = np.lexsort((np.arange(n), keys))
    assert np.array_equal(ids, order.astype(np.int32))
    # voxel_begin == first occurrence of each key
    first = np.full(s.n_cells() + 1, -1, np.int64)
    uk, ui = np.unique(skeys, return_index=True)
    first[uk] = ui
    assert np.array_equal(vb, first[:-1].astype(np.int32))
    assert cnt.min() >= 1 and cnt.max() <= 32
    out = s.download()
    assert np.isfinite(out.position).all() and np.isfinite(out.density).all()
    assert (out.position >= 0).all() and (out.position <= np.asarray(params.space_size)).all()
    s.close()
