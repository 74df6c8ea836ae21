"""Multi-GPU arm of bench.py: strong scaling of one box workload over x-slabs (one rank per GPU, NCCL).

value = N_global x steps / (sum over windows of the max-over-ranks device time).  Every window starts from the same
start state (see bench.py "Window"): each rank snapshots its slab after loading and restores it between windows,
outside the timed events."""
from __future__ import annotations

import json
import time

import numpy as np
import torch
import torch.distributed as dist


def run(args, rank, world, local_rank):
    import bench
    from .slab import NativeSlabRunner, balanced_bounds
    from .strategy import B200SPHStrategy, SphConstants

    name = args.workload or bench.DEFAULT_WORKLOAD
    params, st, mode, desc = bench.make_workload(name, args.particles)
    n = int(params.particle_count)
    voxel_x = float(params.voxel_size[0])
    n_cols = int(np.ceil(params.space_size[0] / voxel_x))
    cols = np.clip((st.position[:, 0] / voxel_x).astype(np.int64), 0, n_cols - 1)
    hist = np.bincount(cols, minlength=n_cols)
    bounds = balanced_bounds(hist, world)
    own = int(hist[bounds[rank]:bounds[rank + 1]].sum())
    capacity = int(1.35 * max(own, n // world)) + 6 * int(hist.max()) + 4096
    window = args.window
    K, W = args.steps, max(args.warmup, 3)
    dev = torch.device("cuda", local_rank)

    # ---- single-GPU reference point on the same workload (rank 0 only, short): its throughput, and its state after one
    #      window for the bitwise parity check of the slab run below ----
    single, ref_state, ref_state1 = None, None, None
    pw = window or 4
    if rank == 0 and not args.no_single:
        s1 = B200SPHStrategy(params, SphConstants(mode=mode), device=local_rank)
        s1.upload(st)
        s1.save_state()
        s1.step(2)
        s1.synchronize()
        tot = 0.0
        for _ in range(2):
            s1.restore_state()
            s1.synchronize()
            t0 = time.perf_counter()
            s1.step(pw)
            s1.synchronize()
            tot += time.perf_counter() - t0
        single = n * 2 * pw / tot
        s1.restore_state()
        s1.step(1)
        ref_state1 = s1.download(np.float32)
        s1.step(pw - 1)
        ref_state = s1.download(np.float32)
        s1.close()
        del s1
        torch.cuda.empty_cache()
    dist.barrier()

    run_ = NativeSlabRunner(params, SphConstants(mode=mode), col_hist=hist, bounds=bounds, device=local_rank)
    capacity = run_.capacity
    run_.load_global(st.position, st.velocity)
    snap = run_.snapshot()

    def restore():
        run_.restore(snap)

    def timed_steps(k):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_.step(k)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    done = 0
    while done < W:
        g = min(window or W, W - done)
        if window:
            restore()
        run_.step(g)
        done += g
    # ---- parity: the state gathered by global id vs the 1-GPU engine, bit for bit: after ONE step (must be equal) and
    #      after one window (equal unless a particle left the domain in between: one GPU aliases its cell key like the
    #      reference, quirk Q5, a slab declares it dead, DESIGN.md D4 -- the differing particles are counted) ----
    parity = None
    if not args.no_single:
        def differing(got, ref):
            bad = np.zeros(n, bool)
            for a_, b_ in zip(got, (ref.position, ref.velocity, ref.density)):
                a_, b_ = np.asarray(a_, np.float32).reshape(n, -1), np.asarray(b_, np.float32).reshape(n, -1)
                bad |= ~np.all((a_ == b_) | (np.isnan(a_) & np.isnan(b_)), axis=1)
            return int(bad.sum())
        restore()
        run_.step(1)
        got = run_.gather_global(n)
        if rank == 0:
            parity = {"after_1_step": {"differing_particles": differing(got, ref_state1)}}
        run_.step(pw - 1)
        got = run_.gather_global(n)
        if rank == 0:
            parity[f"after_{pw}_steps"] = {"differing_particles": differing(got, ref_state)}
            parity["bitwise_equal_after_1_step"] = parity["after_1_step"]["differing_particles"] == 0
            parity["bitwise_equal"] = all(v_["differing_particles"] == 0 for v_ in parity.values() if isinstance(v_, dict))
        del got
        ref_state = ref_state1 = None
    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = run_.launch_count()
    ms_total, done = 0.0, 0
    t_wall0 = time.perf_counter()
    while done < K:
        g = min(window or K, K - done)
        if window:
            restore()
        ms_total += timed_steps(g)
        done += g
    t_wall = time.perf_counter() - t_wall0
    launches = torch.tensor([run_.launch_count() - launches0], device=dev)
    dist.all_reduce(launches)
    cnt = run_.count_global()
    # per-phase breakdown (one extra window, CUDA events around every phase, max over ranks)
    restore()
    acc = {}
    for _ in range(window or 4):
        for k_, v_ in run_.step_timed().items():
            acc[k_] = acc.get(k_, 0.0) + v_ / (window or 4)
    keys_ = sorted(acc)
    bt = torch.tensor([acc[k_] for k_ in keys_], device=dev, dtype=torch.float64)
    # per rank: the work of its own step (everything except the wait at the flag barrier / inside the all_to_all), to see
    # how well the slab boundaries balance it
    own_work = torch.tensor([sum(v_ for k_, v_ in acc.items() if k_ not in ("flag_barrier_ms", "all_to_all_ms"))],
                            device=dev, dtype=torch.float64)
    work_all = [torch.empty_like(own_work) for _ in range(world)]
    dist.all_gather(work_all, own_work)
    dist.all_reduce(bt, op=dist.ReduceOp.MAX)
    breakdown = {k_: float(v_) for k_, v_ in zip(keys_, bt.tolist())}
    stt = run_.check()
    # ---- e2e: every step starts from HOST buffers (fp64, pinned) and ends in them: H2D of the rank's owned particles,
    #      exchange + step, D2H of the owned region; wall clock, max over ranks ----
    k0 = int(snap[3][0].item())
    hp = snap[0][:k0, :3].double().cpu().pin_memory()
    hv = snap[1][:k0, :3].double().cpu().pin_memory()
    own_cap = run_.own_cap
    op_ = torch.empty((own_cap, 4), dtype=torch.float64).pin_memory()
    ov_ = torch.empty((own_cap, 3), dtype=torch.float64).pin_memory()
    og_ = torch.empty((own_cap,), dtype=torch.int32).pin_memory()
    KE = min(K, 8)
    e2e_s, d2h_bytes = 0.0, 0
    for i in range(-2, KE):
        restore()                                   # slot layout of the start state (ids, counters); untimed
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        run_.P[:k0, :3] = hp.to(dev, non_blocking=True).float()
        run_.V[:k0, :3] = hv.to(dev, non_blocking=True).float()
        run_.step(1)
        hw = int(run_.counters[0].item())           # owned slots in use after the step (device -> host read)
        op_[:hw].copy_(run_.P[:hw].double(), non_blocking=True)
        ov_[:hw].copy_(run_.V[:hw, :3].double(), non_blocking=True)
        og_[:hw].copy_(run_.G[:hw], non_blocking=True)
        torch.cuda.synchronize()
        if i >= 0:
            e2e_s += time.perf_counter() - t0
            d2h_bytes = hw * (32 + 24 + 4)
    et = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e2e = {"value": n * KE / float(et.item()), "unit": bench.UNIT, "h2d_bytes_per_step": 48 * k0,
           "d2h_bytes_per_step": d2h_bytes, "steps": KE, "ms_per_step": float(et.item()) / KE * 1e3,
           "api": "per rank: pinned fp64 (N_own,3) position/velocity -> device, NativeSlabRunner.step(1), owned region -> "
                  "pinned fp64 host (bytes are rank 0's)"}
    restore()
    run_.check()
    stats = torch.tensor([stt["ghosts"], stt["hwm"], stt["live"]], device=dev, dtype=torch.float64)
    allstats = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(allstats, stats)
    clocks = sampler.stop() if rank == 0 else None
    assert cnt == n, f"particles lost: {cnt} != {n}"
    if rank == 0:
        hbm_peak, peak_src = bench.peaks()
        own_max = max(int(s_[2]) for s_ in allstats)
        dom = "density" if breakdown["density_ms"] >= breakdown["force_ms"] else "force"
        gbs = bench.BYTES[dom] * own_max / (breakdown[dom + "_ms"] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom + "_rows_kernel", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": gbs / hbm_peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_particle": bench.BYTES[dom], "kernel_ms": breakdown[dom + "_ms"],
                    "note": "slowest rank: owned particles of that rank x algorithmic bytes / its kernel time; the "
                            "sweeps are FP32-issue / latency bound, not HBM bound (DESIGN.md section 4)"}
        value = n * K / (ms_total * 1e-3)
        line = {"metric": bench.METRIC, "value": value, "unit": bench.UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": bench.shared_config(name, desc, n, window),
                "run": {"parallelism": (f"x-slabs x{world}, 2-column ghost halos + migration; "
                                        + ("records stored into the receivers' buffers over NVLink peer memory by the "
                                           "force sweep's epilogue, one flag-barrier kernel per step, no collective on "
                                           "the step path" if run_.p2p else
                                           "one fixed-size all_to_all per step (NCCL), device-side routing")
                                        + ", no host sync in the step loop"),
                        "slab_bounds": bounds, "capacity_per_rank": capacity,
                        "l2": "working set per GPU larger than L2: no flush",
                        "timing": "CUDA events per window, max over ranks, barrier + synchronize on both sides"},
                "parity_vs_1gpu": parity,
                "wall_s_timed_region": t_wall, "clocks": clocks, "gpu_launches": int(launches.item()),
                "e2e": e2e, "roofline": roofline, "cpu_baseline": None,
                "value_1gpu_same_workload": single,
                "ghosts_per_rank": [int(s[0]) for s in allstats],
                "owned_high_water_per_rank": [int(s[1]) for s in allstats],
                "owned_per_rank": [int(s[2]) for s in allstats],
                "exchange_bytes_per_step_per_rank": int(sum(run_.block_bytes)),
                "phase_ms_max_over_ranks": breakdown,
                "work_ms_per_rank": [round(float(w_.item()), 4) for w_ in work_all]}
        print(json.dumps(line), flush=True)
    run_.close()
    dist.barrier()
    dist.destroy_process_group()
