"""Multi-GPU x-slab run vs the single-GPU engine (needs >= 2 GPUs; NCCL).  The slab run must reproduce the single-GPU
engine BITWISE (same neighbour lists, same summation order), and therefore the reference within the fp32 tolerance.

BOX: 3 steps (the walls keep every particle inside the domain).  PIPE: 1 step bitwise -- it already covers halos,
migration and the outlet -> inlet recycle with the xoroshiro state travelling between ranks; later steps contain
particles that bounced beyond x_end, which one GPU handles with the reference's key aliasing (quirk Q5) and a slab
declares dead, so from then on only the particle count is checked."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, n, steps, extra_steps, out_path, kind="native"):
    import faulthandler
    import torch.distributed as dist
    faulthandler.dump_traceback_later(150, exit=True)   # a rank stuck in a collective must not hang the suite
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cuda_sph_b200 import SphConstants
        from cuda_sph_b200.slab import GpuSlabRunner, NativeSlabRunner, equal_count_bounds
        from tests.test_gpu_slab import _case
        params, st = _case(mode, n)
        n_cols = int(np.ceil(params.space_size[0] / params.voxel_size[0]))
        cols = np.clip((st.position[:, 0] / params.voxel_size[0]).astype(np.int64), 0, n_cols - 1)
        hist = np.bincount(cols, minlength=n_cols)
        bounds = equal_count_bounds(hist, world)
        if kind in ("native", "native_a2a"):
            # the pipe case moves two thirds of a rank's particles to the other rank in ONE step (outlet -> inlet
            # recycle at |v| ~ 1e2): size the migrant blocks for that
            run = NativeSlabRunner(params, SphConstants(mode=mode), col_hist=hist, bounds=bounds, device=rank,
                                   compact_every=2, migrant_frac=1.0 if mode == "PIPE" else 0.05, own_slack=2.0,
                                   p2p=(kind == "native"))
            assert run.p2p == (kind == "native")
        else:
            run = GpuSlabRunner(params, SphConstants(mode=mode), capacity=2 * n, bounds=bounds, device=rank)
        run.load_global(st.position, st.velocity)
        run.step(steps)
        assert run.count_global() == n
        pos, vel, rho = run.gather_global(n)
        if rank == 0:
            if kind != "torch":
                halo, migrated = run.status()["ghosts"], 1
            else:
                halo, migrated = run.stats["halo_sent"], run.stats["migrated"]
            np.savez(out_path, pos=pos, vel=vel, rho=rho, halo=halo, migrated=migrated)
        run.step(extra_steps)
        assert run.count_global() == n
        run.close()
    finally:
        faulthandler.cancel_dump_traceback_later()
        dist.destroy_process_group()


def _rebalance_worker(rank, world, port, n, out_path):
    import faulthandler
    import torch.distributed as dist
    faulthandler.dump_traceback_later(150, exit=True)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cuda_sph_b200 import SphConstants, workloads
        from cuda_sph_b200.slab import NativeSlabRunner, equal_count_bounds
        params, st = workloads.dam_break(n, 0.5, seed=7)
        n_cols = int(np.ceil(params.space_size[0] / params.voxel_size[0]))
        cols = np.clip((st.position[:, 0] / params.voxel_size[0]).astype(np.int64), 0, n_cols - 1)
        hist = np.bincount(cols, minlength=n_cols)
        bounds = equal_count_bounds(hist, world)
        run = NativeSlabRunner(params, SphConstants(mode="BOX"), col_hist=hist, bounds=bounds, device=rank,
                               compact_every=2, migrant_frac=0.5, own_slack=2.0)
        run.load_global(st.position, st.velocity)
        run.step(2)
        old_bounds = list(run.bounds)
        run2 = run.rebalanced(min_gain=0.0)          # the column has spread: new boundaries, new runner, same particles
        changed = run2 is not run
        run2.step(2)
        assert run2.count_global() == n
        pos, vel, rho = run2.gather_global(n)
        if rank == 0:
            np.savez(out_path, pos=pos, vel=vel, rho=rho, changed=changed, old=old_bounds, new=list(run2.bounds))
        run2.close()
    finally:
        faulthandler.cancel_dump_traceback_later()
        dist.destroy_process_group()


def test_two_gpu_rebalance_continues_bitwise(tmp_path):
    """SURVEY section 7 "Load balance": slab boundaries re-evaluated between steps of a dam break.  The run with new
    boundaries (new runner, particles re-owned by global id) must continue bit for bit like the single-GPU engine."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from cuda_sph_b200 import B200SPHStrategy, SphConstants, workloads
    n, out = 150000, str(tmp_path / "rebalance.npz")
    mp.spawn(_rebalance_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    got = np.load(out)
    assert bool(got["changed"]) and list(got["old"]) != list(got["new"])
    params, st = workloads.dam_break(n, 0.5, seed=7)
    s = B200SPHStrategy(params, SphConstants(mode="BOX"))
    s.upload(st)
    s.step(4)
    ref = s.download()
    s.close()
    eq = lambda a, b: bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))  # noqa: E731
    assert eq(got["pos"], ref.position) and eq(got["vel"], ref.velocity) and eq(got["rho"], ref.density)


def _overflow_worker(rank, world, port, n, out_path):
    import faulthandler
    import torch.distributed as dist
    faulthandler.dump_traceback_later(150, exit=True)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cuda_sph_b200 import SphConstants, workloads
        from cuda_sph_b200.slab import NativeSlabRunner, equal_count_bounds
        params, st = workloads.uniform_box(n, 8.0, seed=21)
        n_cols = int(np.ceil(params.space_size[0] / params.voxel_size[0]))
        cols = np.clip((st.position[:, 0] / params.voxel_size[0]).astype(np.int64), 0, n_cols - 1)
        hist = np.bincount(cols, minlength=n_cols)
        # migrant blocks of 8 records: the first step already has hundreds of migrants per direction
        run = NativeSlabRunner(params, SphConstants(mode="BOX"), col_hist=hist, bounds=equal_count_bounds(hist, world),
                               device=rank, migrant_frac=0.0, migrant_floor=8, own_slack=2.0, poll_every=0)
        run.load_global(st.position, st.velocity)
        run.step(2)
        stt = run.status()
        tot = torch.tensor([stt["live"], stt["overflow"] & 1], dtype=torch.int64, device=f"cuda:{rank}")
        dist.all_reduce(tot)
        raised = False
        try:
            run.check()
        except RuntimeError:
            raised = True
        if rank == 0:
            np.savez(out_path, live=int(tot[0]), overflowed=int(tot[1]), raised=raised)
        run.close()
    finally:
        faulthandler.cancel_dump_traceback_later()
        dist.destroy_process_group()


def test_two_gpu_migrant_overflow_keeps_the_particles(tmp_path):
    """a migrant whose record does not fit its send block must not vanish.  With 8-record migrant blocks the
    exchange overflows at once: the flag is raised (sticky, check() raises on every rank) and every particle is still
    owned by somebody -- the sender keeps what it could not send and retries in the next step."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    n, out = 200000, str(tmp_path / "overflow.npz")
    mp.spawn(_overflow_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    got = np.load(out)
    assert int(got["overflowed"]) > 0 and bool(got["raised"])
    assert int(got["live"]) == n


def _case(mode, n):
    from cuda_sph_b200 import config, workloads
    if mode == "BOX":
        return workloads.uniform_box(n, 8.0, seed=21)
    params = config.pipe_params(n)
    st = config.start_state_inside_pipe(n, params.pipe, seed=22)
    vel = np.random.default_rng(23).uniform(-20, 20, (n, 3))
    vel[:, 0] += 60.0
    return params, type(st)(st.position, vel.astype(np.float32).astype(np.float64), st.density)


@pytest.mark.parametrize("kind", ["native", "native_a2a", "torch"])
@pytest.mark.parametrize("mode,n,steps,extra", [("BOX", 200000, 3, 0), ("PIPE", 20000, 1, 2)])
def test_two_gpu_slabs_equal_single_gpu(tmp_path, mode, n, steps, extra, kind):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from cuda_sph_b200 import B200SPHStrategy, SphConstants
    out = str(tmp_path / "slab.npz")
    mp.spawn(_worker, args=(2, _free_port(), mode, n, steps, extra, out, kind), nprocs=2, join=True)
    got = np.load(out)
    params, st = _case(mode, n)
    s = B200SPHStrategy(params, SphConstants(mode=mode))
    s.upload(st)
    s.step(steps)
    ref = s.download()
    s.close()
    eq = lambda a, b: bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))  # noqa: E731
    assert eq(got["pos"], ref.position)
    assert eq(got["vel"], ref.velocity)
    assert eq(got["rho"], ref.density)
    assert got["halo"] > 0 and got["migrated"] > 0


@pytest.mark.parametrize("p2p", [True, False])
@pytest.mark.parametrize("mode,n,steps", [("BOX", 120000, 3), ("PIPE", 20000, 1)])
def test_single_rank_native_slab_equals_plain_engine(mode, n, steps, p2p):
    """World size 1 (runs on the driver's 1-GPU box): the slot layout with holes, the routing (p2p: fused into the force
    sweep's epilogue, records stored straight into the receive buffer + flag barrier; else slab_route + self-copy),
    slab_unpack, the in-cell order repair by global id (fix_order_kernel) and the compaction must reproduce the plain
    engine bit for bit -- BOX three steps (compaction after step 2), PIPE one step incl. the outlet -> inlet recycle."""
    from cuda_sph_b200 import B200SPHStrategy, SphConstants
    from cuda_sph_b200.slab import NativeSlabRunner
    params, st = _case(mode, n)
    n_cols = int(np.ceil(params.space_size[0] / params.voxel_size[0]))
    cols = np.clip((st.position[:, 0] / params.voxel_size[0]).astype(np.int64), 0, n_cols - 1)
    hist = np.bincount(cols, minlength=n_cols)
    run = NativeSlabRunner(params, SphConstants(mode=mode), col_hist=hist, bounds=[0, n_cols], device=0,
                           compact_every=2, migrant_frac=1.0 if mode == "PIPE" else 0.05, own_slack=2.0, p2p=p2p)
    # shuffle the slot order: arrival order on a slab is arbitrary, only the global ids define the in-cell order
    perm = np.random.default_rng(5).permutation(n)
    run.load_global(st.position, st.velocity)
    k = int(run.counters[0].item())
    idx = torch.as_tensor(perm[:k], device=run.P.device)
    run.P[:k], run.V[:k], run.G[:k] = run.P[:k][idx].clone(), run.V[:k][idx].clone(), run.G[:k][idx].clone()
    snap = run.snapshot()
    run.step(steps)
    assert run.count_global() == n
    pos, vel, rho = run.gather_global(n)
    s = B200SPHStrategy(params, SphConstants(mode=mode))
    s.upload(st)
    s.step(steps)
    ref = s.download()
    eq = lambda a, b: bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))  # noqa: E731
    assert eq(pos, ref.position) and eq(vel, ref.velocity) and eq(rho, ref.density)
    if mode == "PIPE":   # the xoroshiro states travelled with the recycled particles
        assert np.array_equal(run.R.cpu().numpy().view(np.uint64), s.rng_states())
    # snapshot / restore (bench windows) incl. the RNG states: the same steps again give the same bits
    run.restore(snap)
    run.step(steps)
    pos2, vel2, rho2 = run.gather_global(n)
    assert eq(pos2, pos) and eq(vel2, vel) and eq(rho2, rho)
    s.close()
    run.close()
