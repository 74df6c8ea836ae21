"""x-slab decomposition logic on CPU: world_size 2 and 3 over gloo, with the fp64 oracle as the local step.

The decomposition (two-column halos, ownership, migration to arbitrary ranks, in-cell order by global id, xoroshiro
state travelling with a particle) must reproduce the single-domain oracle BITWISE after several steps -- the physics
is violent enough (speeds of 1e2..1e5 cells per step) that particles cross several slabs per step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cuda_sph_b200 import config, workloads
from cuda_sph_b200.slab import HALO, SingleExchangeSlabRunner, SlabRunner, equal_count_bounds
from oracle import oracle as orc
from tests.helpers import same


class _OracleStep:
    """Local step = the CPU oracle on (owned + ghost) particles; mixed into both protocol runners."""

    def __init__(self, P: orc.OracleParams, n_global, capacity, n_cols, bounds, pipe_mode):
        super().__init__(n_cols, P.voxel[0], bounds)
        self.oparams = P
        self.P = torch.zeros((capacity, 4), dtype=torch.float64)
        self.V = torch.zeros((capacity, 4), dtype=torch.float64)
        self.G = torch.zeros(capacity, dtype=torch.int32)
        self.R = torch.from_numpy(orc.rng_init(n_global).view(np.int64).copy()) if pipe_mode else None

    def _local_step(self, n_own, n_local):
        if n_local == 0:
            return
        gid = self.G[:n_local].numpy()
        order = np.argsort(gid, kind="stable")      # local row order == global id order, like the single domain
        inv = np.empty_like(order)
        inv[order] = np.arange(n_local)
        pos = self.P[:n_local, :3].numpy()[order]
        vel = self.V[:n_local, :3].numpy()[order]
        P = orc.OracleParams(**{**self.oparams.__dict__, "n": n_local, "_keep": []})
        rng = None
        if self.R is not None:
            rng = self.R.numpy()[gid[order]].view(np.uint64).copy()
        r = orc.step(P, pos, vel, rng=rng, light=True)
        own = inv[:n_own]
        self.P[:n_own, :3] = torch.from_numpy(r.position[own])
        self.P[:n_own, 3] = torch.from_numpy(r.density[own])
        self.V[:n_own, :3] = torch.from_numpy(r.velocity[own])
        if self.R is not None:
            self.R[torch.from_numpy(gid[:n_own].astype(np.int64))] = torch.from_numpy(rng[own].view(np.int64))


class OracleSlabRunner(_OracleStep, SlabRunner):
    """Two exchanges per step: halo, local step, migration."""


class OracleSingleExchangeRunner(_OracleStep, SingleExchangeSlabRunner):
    """One exchange per step (the protocol of the native CUDA path): route + exchange, local step."""


def _case(mode):
    if mode == "BOX":
        n = 3000
        params, st = workloads.uniform_box(n, 8.0, seed=11)
        P = orc.OracleParams(n=n, mode="BOX", space=tuple(params.space_size), dt=1 / params.fps)
    else:
        n = 2500
        params = config.pipe_params(n)
        st = config.start_state_inside_pipe(n, params.pipe, seed=12)
        rng = np.random.default_rng(13)
        vel = rng.uniform(-20, 20, (n, 3))
        vel[:, 0] += 60.0                                   # strong flow towards the outlet -> recycles
        st = type(st)(st.position, vel.astype(np.float32).astype(np.float64), st.density)
        P = orc.OracleParams(n=n, mode="PIPE", space=tuple(params.space_size), ext=tuple(params.external_force),
                             dt=1 / params.fps, pipe=params.pipe.to_numpy())
    return n, params, st, P


def _worker(rank, world, port, mode, steps, out_path, single_exchange=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc.set_exact_pow(False)
        orc.set_num_threads(1)
        n, params, st, P = _case(mode)
        n_cols = int(np.ceil(params.space_size[0] / params.voxel_size[0]))
        cols = np.clip((st.position[:, 0] / params.voxel_size[0]).astype(np.int64), 0, n_cols - 1)
        bounds = equal_count_bounds(np.bincount(cols, minlength=n_cols), world)
        cls = OracleSingleExchangeRunner if single_exchange else OracleSlabRunner
        run = cls(P, n, capacity=2 * n, n_cols=n_cols, bounds=bounds, pipe_mode=(mode == "PIPE"))
        run.load_global(st.position, st.velocity)
        assert run.count_global() == n
        run.step(steps)
        assert run.count_global() == n                       # nothing lost, nothing duplicated
        pos, vel, rho = run.gather_global(n)
        stats = torch.tensor([run.stats["halo_sent"], run.stats["migrated"]])
        dist.all_reduce(stats)
        if rank == 0:
            np.savez(out_path, pos=pos, vel=vel, rho=rho, halo=int(stats[0]), migrated=int(stats[1]),
                     bounds=np.asarray(bounds))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("single_exchange", [False, True], ids=["two-exchanges", "single-exchange"])
@pytest.mark.parametrize("world,mode", [(2, "BOX"), (3, "BOX"), (2, "PIPE")])
def test_slab_decomposition_matches_single_domain_bitwise(tmp_path, world, mode, single_exchange):
    steps = 3
    out = str(tmp_path / "slab.npz")
    mp.spawn(_worker, args=(world, _free_port(), mode, steps, out, single_exchange), nprocs=world, join=True)
    got = np.load(out)
    # single-domain oracle
    orc.set_exact_pow(False)
    n, params, st, P = _case(mode)
    rng = orc.rng_init(n) if mode == "PIPE" else None
    pos, vel, rho = st.position, st.velocity, None
    for _ in range(steps):
        r = orc.step(P, pos, vel, rng=rng, light=True)
        pos, vel, rho = r.position, r.velocity, r.density
    assert same(got["pos"], pos)
    assert same(got["vel"], vel)
    assert same(got["rho"], rho)
    assert got["halo"] > 0 and got["migrated"] > 0           # the exchanges really happened
    assert all(b - a >= HALO for a, b in zip(got["bounds"][:-1], got["bounds"][1:]))


def _rebalance_worker(rank, world, port, out_path, single_exchange):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc.set_exact_pow(False)
        orc.set_num_threads(1)
        n, params, st, P = _dam_case()
        n_cols = int(np.ceil(params.space_size[0] / params.voxel_size[0]))
        cols = np.clip((st.position[:, 0] / params.voxel_size[0]).astype(np.int64), 0, n_cols - 1)
        bounds = equal_count_bounds(np.bincount(cols, minlength=n_cols), world)
        cls = OracleSingleExchangeRunner if single_exchange else OracleSlabRunner
        run = cls(P, n, capacity=2 * n, n_cols=n_cols, bounds=bounds, pipe_mode=False)
        run.load_global(st.position, st.velocity)
        run.step(2)
        moved = run.rebalance()                              # the column has spread: the boundaries move
        new_bounds = list(run.bounds)
        run.step(2)
        assert run.count_global() == n
        pos, vel, rho = run.gather_global(n)
        if rank == 0:
            np.savez(out_path, pos=pos, vel=vel, rho=rho, moved=moved, old=np.asarray(bounds), new=np.asarray(new_bounds))
    finally:
        dist.destroy_process_group()


def _dam_case():
    n = 6000
    params, st = workloads.dam_break(n, 0.5, seed=7)
    P = orc.OracleParams(n=n, mode="BOX", space=tuple(params.space_size), dt=1 / params.fps)
    return n, params, st, P


@pytest.mark.parametrize("single_exchange", [False, True], ids=["two-exchanges", "single-exchange"])
def test_rebalance_between_steps_continues_bitwise(tmp_path, single_exchange):
    """SURVEY section 7 "Load balance": boundaries re-evaluated between steps of a dam break (world 2, gloo).  Two steps,
    rebalance, two more steps == four steps of the single-domain oracle, bit for bit."""
    out = str(tmp_path / "rebalance.npz")
    mp.spawn(_rebalance_worker, args=(2, _free_port(), out, single_exchange), nprocs=2, join=True)
    got = np.load(out)
    assert bool(got["moved"]) and list(got["old"]) != list(got["new"])
    orc.set_exact_pow(False)
    n, params, st, P = _dam_case()
    pos, vel, rho = st.position, st.velocity, None
    for _ in range(4):
        r = orc.step(P, pos, vel, light=True)
        pos, vel, rho = r.position, r.velocity, r.density
    assert same(got["pos"], pos) and same(got["vel"], vel) and same(got["rho"], rho)


def test_equal_count_bounds():
    hist = np.zeros(75, np.int64)
    hist[:8] = 1000                                          # dam-break column: everything in the first columns
    b = equal_count_bounds(hist, 4)
    assert b == [0, 2, 4, 6, 75]
    b = equal_count_bounds(np.full(40, 10), 8)
    assert b == [0, 5, 10, 15, 20, 25, 30, 35, 40]
    with pytest.raises(ValueError):
        equal_count_bounds(np.ones(6), 4)


def test_column_and_owner_mapping():
    run = SlabRunner.__new__(SlabRunner)
    run.voxel_x, run.n_cols, run.bounds = 2.0, 10, [0, 3, 7, 10]
    x = torch.tensor([0.0, 1.99, 2.0, -0.5, 5.99, 6.0, 13.99, 14.0, 19.99, float("nan"), float("inf"), 25.0])
    col = run.columns(x)
    assert col.tolist() == [0, 0, 1, 0, 2, 3, 6, 7, 9, -1, -1, 12]
    assert run.owner_of(col[:9]).tolist() == [0, 0, 0, 0, 0, 1, 1, 2, 2]
    assert run.owner_of(torch.tensor([12])).tolist() == [2]   # clamped


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("pipe_mode", [False, True])
def test_native_exchange_capacities_are_symmetric(world, pipe_mode):
    """The native exchange is ONE all_to_all with static split sizes: the block a -> b must be as large as b -> a, and
    every rank must derive that from the global histogram alone."""
    from cuda_sph_b200.slab import plan_capacities
    rng = np.random.default_rng(world)
    hist = rng.integers(50, 4000, size=64)
    hist[:6] *= 9                                   # a dam-break-like pile at low x
    bounds = equal_count_bounds(hist, world)
    n = int(hist.sum())
    plans = [plan_capacities(hist, bounds, r, n, pipe_mode) for r in range(world)]
    for a in range(world):
        assert plans[a]["cap_m"][a] == 0            # nobody migrates to itself
        own = int(hist[bounds[a]:bounds[a + 1]].sum())
        assert plans[a]["own_cap"] >= own and plans[a]["capacity"] > plans[a]["own_cap"]
        for b in range(world):
            assert plans[a]["cap_m"][b] == plans[b]["cap_m"][a]
            assert plans[a]["cap_g"][b] == plans[b]["cap_g"][a]
        for b in (a - 1, a + 1):                    # a neighbour's two boundary columns fit its ghost block
            if 0 <= b < world:
                lo, hi = bounds[b], bounds[b + 1]
                band = hist[hi - HALO:hi].sum() if b < a else hist[lo:lo + HALO].sum()
                assert plans[a]["cap_g"][b] >= band
    if pipe_mode and world > 2:
        assert plans[0]["cap_m"][world - 1] >= n // world // 4      # outlet -> inlet recycle


def test_balanced_bounds_puts_wider_slabs_at_the_ends():
    """Work-balanced boundaries: cost = owned + 0.6 x ghost columns; edge ranks have one halo and get more columns."""
    from cuda_sph_b200.slab import balanced_bounds
    hist = np.full(162, 1000)
    b = balanced_bounds(hist, 8)
    widths = np.diff(b)
    assert b[0] == 0 and b[-1] == 162 and widths.min() >= HALO
    assert widths[0] >= widths[1:-1].max() and widths[-1] >= widths[1:-1].max()

    def cost(lo, hi):
        return hist[lo:hi].sum() + 0.6 * (hist[max(lo - HALO, 0):lo].sum() + hist[hi:hi + HALO].sum())
    worst = max(cost(b[r], b[r + 1]) for r in range(8))
    e = equal_count_bounds(hist, 8)
    assert worst <= max(cost(e[r], e[r + 1]) for r in range(8))
    # a dam-break pile: still a valid partition, and never worse than equal counts
    hist2 = np.zeros(75, np.int64)
    hist2[:8] = 131072
    b2 = balanced_bounds(hist2, 4)
    assert b2[0] == 0 and b2[-1] == 75 and np.diff(b2).min() >= HALO


def test_slab_partition_properties():
    """Random column histograms: both boundary planners give a monotone cover with every slab >= HALO wide, and the
    native-exchange capacities stay symmetric."""
    from hypothesis import given, settings
    from hypothesis import strategies as hst
    from cuda_sph_b200.slab import balanced_bounds, plan_capacities

    @settings(max_examples=60, deadline=None)
    @given(hst.integers(2, 8), hst.integers(0, 2 ** 31 - 1), hst.booleans())
    def check(world, seed, pile):
        rng = np.random.default_rng(seed)
        ncol = int(rng.integers(world * HALO, 200))
        hist = rng.integers(0, 5000, size=ncol)
        if pile:
            hist[ncol // 8:] = 0                     # everything in the first columns (dam-break)
        for planner in (equal_count_bounds, balanced_bounds):
            b = planner(hist, world)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == ncol
            assert all(y - x >= HALO for x, y in zip(b[:-1], b[1:]))
        b = balanced_bounds(hist, world)
        n = int(hist.sum())
        plans = [plan_capacities(hist, b, r, n, False) for r in range(world)]
        for a in range(world):
            for c in range(world):
                assert plans[a]["cap_m"][c] == plans[c]["cap_m"][a] and plans[a]["cap_g"][c] == plans[c]["cap_g"][a]
            assert plans[a]["own_cap"] >= int(hist[b[a]:b[a + 1]].sum())

    check()


def test_single_exchange_runner_without_a_process_group():
    """world == 1: the single-exchange runner degenerates to the plain step (no halo, no migration)."""
    orc.set_exact_pow(False)
    n, params, st, P = _case("BOX")
    n_cols = int(np.ceil(params.space_size[0] / params.voxel_size[0]))
    run = OracleSingleExchangeRunner(P, n, capacity=2 * n, n_cols=n_cols, bounds=[0, n_cols], pipe_mode=False)
    run.load_global(st.position, st.velocity)
    run.step(2)
    pos, vel, rho = run.gather_global(n)
    p, v = st.position, st.velocity
    for _ in range(2):
        r = orc.step(P, p, v, light=True)
        p, v = r.position, r.velocity
    assert same(pos, p) and same(vel, v) and same(rho, r.density)
