#!/usr/bin/env python
"""bench.py -- particle-updates/s of the SPH step hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # CUDA engine (libsph_b200.so)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port, all host threads)

A "step" is one full Voxel SPH step over the resident particle set: hash -> sort -> cell table + reorder -> density
sweep -> fused pressure/viscosity/integrate/collide sweep (reference: compute_next_state,
sim/src/sph/strategies/abstract_sph_strategy.py:31-46).

Workloads:
    box32m  BASELINE configs[3]: uniform box, 2^25 particles, 8 per cell -- the default at EVERY N (it fits one GPU), so
            the 1 -> 8 GPU scaling curve is one workload
    dam1m   BASELINE configs[1]: box dam-break column, 2^20 particles -- at N = 1 a second record of the same line
            (`secondary.dam1m`: value, roofline, e2e, and the 1000-step straight-through run with dead / non-finite
            counts next to the oracle's)
    pipe4m  BASELINE configs[2]: six-segment pipe, 2^22 particles, inflow/outflow recycle (--workload pipe4m)

Window: the reference's Voxel physics diverges within ~10 steps on every workload (fp64 oracle: |v| ~ 1e21 by step
10, most particles NaN by step 40 -- DESIGN.md "Workloads"), after which a step is mostly dead particles.  So every
arm keeps the state inside steps 1..WINDOW of the workload: after WINDOW steps the start state is restored (a
device-to-device copy outside the timed events).  `--window 0` disables that and runs the steps straight through.

`value`  : N x steps / device time, inputs resident in HBM, L2 flushed between steps (per-step CUDA events summed).
`e2e`    : the same through the reference-facing call compute_next_state (fp64 host buffers in pinned memory,
           H2D + step + D2H every step; copies interleaved with the step on two streams), wall clock around the
           synchronous calls.
`roofline`: dominant kernel (the density sweep), algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json, its
           measured DRAM traffic (ncu, profiles/traffic.json) and the two roofs that actually bind it: shared-memory
           wavefronts (the LSU data pipe) and issue slots.
`cpu_baseline`: the oracle port timed on this box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "particle-updates/sec"
UNIT = "particle-updates/s"

# algorithmic bytes per particle (DESIGN.md "Kernels")
BYTES = {"hash": 20, "reorder": 72, "density": 20, "force": 84}   # B / particle


def sort_bytes(passes: int) -> int:
    # onesweep: every digit pass reads key + value (8) and writes key + value (8); pass 0 has no value read (values are
    # iota); the digit histograms come from hash_hist_kernel (counted under "hash": it reads the positions, not the keys)
    return passes * 16 - 4


def make_workload(name: str, n_override: int | None):
    from cuda_sph_b200 import workloads
    if name == "dam1m":
        n = n_override or (1 << 20)
        params, st = workloads.dam_break(n, 2.5, seed=0)
        desc = f"box dam-break column (first 10% of x, ~25/cell), N={n}, box {params.space_size[0]:.0f}^3"
        return params, st, "BOX", desc
    if name == "box32m":
        n = n_override or (1 << 25)
        params, st = workloads.uniform_box(n, 8.0, seed=0)
        desc = f"uniform box 8 particles/cell, N={n}, box {params.space_size[0]:.0f}^3"
        return params, st, "BOX", desc
    if name == "pipe4m":
        n = n_override or (1 << 22)
        params, st = workloads.pipe_flow(n, seed=0)
        desc = f"six-segment pipe, N={n}, space {list(map(float, params.space_size))}"
        return params, st, "PIPE", desc
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


WINDOW = 4
DEFAULT_WORKLOAD = "box32m"   # BASELINE configs[3]: the one workload every N runs (it fits one GPU), so the driver's
                              # scaling efficiency compares like with like; dam1m (configs[1]) rides along at N = 1


def shared_config(name: str, desc: str, n: int, window: int) -> dict:
    """The `config` object of the JSON line: identical in both arms (b200 / reference) and at every N."""
    return {"workload": desc, "name": name, "particles": n,
            "window": f"steps 1..{window} of the workload, start state restored in between (untimed)" if window
            else "none: steps run straight through"}


def _oracle(params, mode):
    from oracle import oracle as orc
    orc.set_exact_pow(False)
    orc.set_num_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every host core
    n = int(params.particle_count)
    pipe = params.pipe.to_numpy() if mode == "PIPE" else None
    P = orc.OracleParams(n=n, mode=mode, space=tuple(params.space_size), ext=tuple(params.external_force),
                         dt=1 / params.fps, pipe=pipe)
    rng = orc.rng_init(n) if mode == "PIPE" else None
    return orc, P, rng


def cpu_baseline(params, st, mode, window, budget_s=12.0):
    """Oracle port (fp64, OpenMP) on the host cores, bounded sample of the same workload: the first steps of it."""
    orc, P, rng = _oracle(params, mode)
    n = int(params.particle_count)
    pos, vel = st.position, st.velocity
    el, steps = 0.0, 0
    while True:
        if window and steps % window == 0:
            pos, vel = st.position, st.velocity
        t0 = time.perf_counter()
        r = orc.step(P, pos, vel, rng=rng, light=True)
        el += time.perf_counter() - t0
        pos, vel = r.position, r.velocity
        steps += 1
        if el > budget_s or steps >= 48 or el / steps * (steps + 1) > 2.5 * budget_s:
            break
    return {"value": n * steps / el, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
            "sample": f"{steps} full steps of the same workload (N={n}; steps 1..{window or steps} from its start "
                      f"state), fp64 C/OpenMP restatement oracle/sph_oracle.c, {el:.1f} s"}


def run_reference_arm(args):
    """--impl reference: the reference's CPU path.  The reference is pure Python/numba and not installable on the box,
    so this is the oracle port (oracle/sph_oracle.c) with all host threads.  Every one of the K steps is a BOUNDED
    SAMPLE of the workload: the same generator at the same particles-per-cell with fewer particles, sized so that
    K + W steps take about two minutes (throughput per particle does not depend on the box size)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload or DEFAULT_WORKLOAD
    full_n = args.particles or {"dam1m": 1 << 20, "box32m": 1 << 25, "pipe4m": 1 << 22}[name]
    K, W = args.steps, max(args.warmup, 1)
    # calibrate on a small sample, then size the sample for ~120 s in total
    cal_n = min(full_n, 1 << 18)
    params, st, mode, _ = make_workload(name, cal_n)
    orc, P, rng = _oracle(params, mode)
    orc.step(P, st.position, st.velocity, rng=rng, light=True)
    t0 = time.perf_counter()
    orc.step(P, st.position, st.velocity, rng=rng, light=True)
    rate = cal_n / (time.perf_counter() - t0)
    n = full_n
    while n > cal_n and n * (K + W) / rate > 120.0:
        n //= 2
    params, st, mode, _ = make_workload(name, n)
    _, _, _, desc = make_workload_desc(name, full_n)
    orc, P, rng = _oracle(params, mode)
    pos, vel = st.position, st.velocity
    for i in range(W):
        if args.window and i % args.window == 0:
            pos, vel = st.position, st.velocity
        r = orc.step(P, pos, vel, rng=rng, light=True)
        pos, vel = r.position, r.velocity
    el = 0.0
    for i in range(K):
        if args.window and i % args.window == 0:
            pos, vel = st.position, st.velocity
        t0 = time.perf_counter()
        r = orc.step(P, pos, vel, rng=rng, light=True)
        el += time.perf_counter() - t0
        pos, vel = r.position, r.velocity
    value = n * K / el
    sample = (f"each of the {K} timed steps (after {W} warm-up) is one full step over {n} particles of the same generator "
              f"and density ({n}/{full_n} of the workload's particles), steps 1..{args.window or K} from its start "
              f"state; fp64 C/OpenMP restatement oracle/sph_oracle.c")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": el / K * 1e3, "sample_particles": n, "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": shared_config(name, desc, full_n, args.window),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def make_workload_desc(name: str, n: int):
    """Description string of a workload without generating it (the reference arm names the FULL workload)."""
    from cuda_sph_b200 import config, workloads
    if name == "dam1m":
        d = workloads.cubic_dims(n, 2.5)
        return None, None, "BOX", f"box dam-break column (first 10% of x, ~25/cell), N={n}, box {d * config.INF_R:.0f}^3"
    if name == "box32m":
        d = workloads.cubic_dims(n, 8.0)
        return None, None, "BOX", f"uniform box 8 particles/cell, N={n}, box {d * config.INF_R:.0f}^3"
    params, st, mode, desc = make_workload(name, n)
    return None, None, mode, desc


def measure_single(name, args, local_rank, stream, flush, *, K, W, e2e_steps, with_cpu):
    """One workload on one GPU: value (device-resident, L2 flushed between steps), stage times + roofline of the
    dominant kernel, e2e through compute_next_state, CPU baseline."""
    import torch
    from cuda_sph_b200 import B200SPHStrategy, SphConstants
    params, st, mode, desc = make_workload(name, args.particles if name == (args.workload or DEFAULT_WORKLOAD) else None)
    n = int(params.particle_count)
    s = B200SPHStrategy(params, SphConstants(mode=mode), device=local_rank, cuda_stream=stream.cuda_stream)
    s.upload(st)
    s.save_state()
    s.synchronize()
    window = args.window

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.zero_()

    for i in range(W):
        if window and i % window == 0:
            s.restore_state()
        s.step(1)
    s.synchronize()

    # ---- value: device-resident steps, L2 flushed between steps, per-step events summed -------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = s.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for i in range(K):
        if window and i % window == 0:
            s.restore_state()
        flush_l2()
        starts[i].record(stream)
        s.step(1)
        ends[i].record(stream)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    launches = s.launch_count() - launches0
    ms_total = float(sum(a.elapsed_time(b) for a, b in zip(starts, ends)))
    value = n * K / (ms_total * 1e-3)

    # back-to-back (no flush) for context: groups of `window` steps, events around each group
    grp = window or K
    hot_ms, done = 0.0, 0
    while done < K:
        g_ = min(grp, K - done)
        if window:
            s.restore_state()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        s.step(g_)
        e1.record(stream)
        torch.cuda.synchronize()
        hot_ms += e0.elapsed_time(e1)
        done += g_
    clocks = sampler.stop()
    stats = s.stats()

    # ---- per-kernel CUDA-event timings (eager launches, L2 flushed between steps) --------------------------------
    KT = min(K, 20)
    acc = {}
    for i in range(KT):
        if window and i % window == 0:
            s.restore_state()
        flush_l2()
        t = s.step_timed(1)
        for k_, v_ in t.items():
            acc[k_] = acc.get(k_, 0.0) + v_
    passes = t["sort_passes"]
    hbm_peak, peak_src = peaks()
    stage_ms = {k_: acc[k_ + "_ms"] / KT for k_ in ("hash", "sort", "reorder", "density", "force")}
    stage_bytes = dict(BYTES, sort=sort_bytes(passes))
    stage_gbs = {k_: stage_bytes[k_] * n / (stage_ms[k_] * 1e-3) / 1e9 for k_ in stage_ms}
    dom = max(stage_ms, key=stage_ms.get)
    traffic, issue, lsu = None, None, None
    try:   # measured DRAM bytes per launch of the dominant kernel, from the committed ncu capture of this workload
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
        if tj and tj["particles"] == n and dom in tj:
            traffic = tj[dom]["bytes"]
            if "warp_inst" in tj[dom]:
                # second roof, the one that binds: warp instructions per launch (ncu) / live kernel time vs the issue
                # slots of the chip (148 SMs x 4 schedulers x SM clock)
                peak_issue = 148 * 4 * 1.965e9
                rate = tj[dom]["warp_inst"] / (stage_ms[dom] * 1e-3)
                issue = {"warp_inst_per_launch": tj[dom]["warp_inst"], "achieved_ginst_s": rate / 1e9,
                         "peak_ginst_s": peak_issue / 1e9, "frac": rate / peak_issue,
                         "peak_source": "148 SMs x 4 issue slots x 1.965 GHz"}
            if "smem_wavefronts" in tj[dom]:
                # third roof, the one that binds the sweeps first: shared-memory wavefronts per launch (ncu) / live kernel
                # time vs one wavefront per SM and cycle (the LSU data pipe)
                peak_wf = 148 * 1.965e9
                wrate = tj[dom]["smem_wavefronts"] / (stage_ms[dom] * 1e-3)
                lsu = {"smem_wavefronts_per_launch": tj[dom]["smem_wavefronts"], "achieved_gwavefronts_s": wrate / 1e9,
                       "peak_gwavefronts_s": peak_wf / 1e9, "frac": wrate / peak_wf,
                       "peak_source": "148 SMs x 1 wavefront per cycle x 1.965 GHz"}
    except Exception:
        pass
    kname = {"density": "density_flat_kernel", "force": "force_rows_kernel"}.get(dom, dom + "_kernel")
    roofline = {"bound": "hbm", "kernel": kname, "achieved": stage_gbs[dom], "peak": hbm_peak,
                "unit": "GB/s", "frac": stage_gbs[dom] / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": stage_bytes[dom] * n, "issue_roofline": issue,
                "shared_memory_roofline": lsu,
                "algorithmic_bytes_per_particle": stage_bytes[dom], "kernel_ms": stage_ms[dom],
                "stages_ms": stage_ms, "stages_gbs": stage_gbs,
                "step_bytes_per_particle": sum(stage_bytes.values()),
                "step_frac": sum(stage_bytes.values()) * value / 1e9 / hbm_peak,
                "note": "the neighbour sweeps are bound by shared-memory bandwidth and instruction issue, not HBM (ncu: "
                        "LSU data pipe 60-80 % busy, 4-20 % DRAM throughput; DESIGN.md section 4); the HBM fraction is "
                        "reported as the contract asks; kernel_ms of the density stage includes its row-plan kernel; "
                        "traffic (ncu, dram read + write of the named kernel alone) exceeds the algorithmic bytes because "
                        "the neighbour lists (64 B/particle) and the pair factors (two 4-byte lanes of 16-byte records = "
                        "32 B of sectors per particle) are engine-internal"}

    # ---- e2e: compute_next_state through the C ABI with pinned fp64 host buffers ----------------------------------
    e2e = None
    if e2e_steps > 0:
        def pinned(shape):
            return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()

        hp, hv = pinned((n, 3)), pinned((n, 3))
        op, ov, orho = pinned((n, 3)), pinned((n, 3)), pinned((n,))
        KE = min(K, e2e_steps)
        nwarm = 3 if n <= (1 << 22) else 1
        e2e_s = 0.0
        for i in range(-nwarm, KE):          # untimed warm-up calls
            if i <= 0 or (window and i % window == 0):
                hp[:], hv[:] = st.position, st.velocity     # back to the start state (host side, untimed)
            t0 = time.perf_counter()
            s.compute_next_state_into(hp, hv, op, ov, orho)
            if i >= 0:
                e2e_s += time.perf_counter() - t0
            hp, op = op, hp
            hv, ov = ov, hv
        e2e = {"value": n * KE / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 48 * n, "d2h_bytes_per_step": 56 * n,
               "steps": KE, "ms_per_step": e2e_s / KE * 1e3,
               "api": "B200SPHStrategy.compute_next_state_into -> sph_compute_next_state (fp64 pinned host buffers)"}
        del hp, hv, op, ov, orho

    cpu = cpu_baseline(params, st, mode, window) if with_cpu else None
    out = {"name": name, "desc": desc, "n": n, "mode": mode, "value": value, "ms_per_step": ms_total / K,
           "value_no_flush": n * K / (hot_ms * 1e-3), "wall_s": t_wall, "clocks": clocks, "launches": int(launches),
           "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
           "state": {"n_dead": stats["n_dead"], "n_nonfinite": stats["n_nonfinite"],
                     "max_density": stats["max_density"], "steps_done": stats["steps_done"]}}
    return out, s, (params, st, mode)


def straight_through(s, params, st, mode, steps, oracle_steps=5):
    """BASELINE configs[1] as written: `steps` steps straight through (no window), throughput by CUDA events around the
    whole run, dead / non-finite particle counts along the way and -- for the first steps -- next to the fp64 oracle's."""
    import torch
    n = int(params.particle_count)
    s.upload(st)
    s.synchronize()
    counts = []
    orc, P, rng = _oracle(params, mode)
    pos, vel = st.position, st.velocity
    for k in range(oracle_steps):
        s.step(1)
        r = orc.step(P, pos, vel, rng=rng, light=True)
        pos, vel = r.position, r.velocity
        stt = s.stats()
        counts.append({"step": k + 1, "engine_nonfinite": int(stt["n_nonfinite"]), "engine_dead": int(stt["n_dead"]),
                       "oracle_nonfinite": int((~(np.isfinite(pos).all(axis=1) & np.isfinite(vel).all(axis=1))).sum())})
    s.upload(st)
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms, done, trace = 0.0, 0, []
    while done < steps:
        g_ = min(100, steps - done)
        torch.cuda.synchronize()
        e0.record()
        s.step(g_)
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
        done += g_
        stt = s.stats()
        trace.append({"step": done, "nonfinite": int(stt["n_nonfinite"]), "dead": int(stt["n_dead"])})
    return {"steps": steps, "value": n * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
            "first_steps_vs_oracle": counts, "every_100_steps": trace,
            "note": "no window: the reference's physics diverges (DESIGN.md section 6), later steps are increasingly dead "
                    "or wall-piled particles; reported beside the windowed headline, not instead of it"}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from cuda_sph_b200 import bench_multi
        return bench_multi.run(args, rank, world, local_rank)

    name = args.workload or DEFAULT_WORKLOAD
    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    K, W = args.steps, max(args.warmup, 3)
    main, s, _ = measure_single(name, args, local_rank, stream, flush, K=K, W=W,
                                e2e_steps=48 if name != "box32m" else 8, with_cpu=not args.no_cpu_baseline)
    s.close()
    del s
    torch.cuda.empty_cache()

    secondary = None
    if name == DEFAULT_WORKLOAD and not args.no_secondary:
        # BASELINE configs[1] (dam1m) beside the headline: same measurement, its own roofline, and the 1000-step run
        sec, s2, (p2, st2, m2) = measure_single("dam1m", args, local_rank, stream, flush, K=max(K, 20), W=W,
                                                e2e_steps=20, with_cpu=False)
        sec["straight_through"] = straight_through(s2, p2, st2, m2, args.straight_steps)
        s2.close()
        secondary = {"dam1m": {"config": shared_config("dam1m", sec["desc"], sec["n"], args.window),
                               "value": sec["value"], "unit": UNIT, "ms_per_step": sec["ms_per_step"],
                               "value_no_flush": sec["value_no_flush"], "e2e": sec["e2e"], "roofline": sec["roofline"],
                               "gpu_launches": sec["launches"], "state": sec["state"],
                               "straight_through": sec["straight_through"]}}

    line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": shared_config(name, main["desc"], main["n"], args.window),
            "run": {"parallelism": "1 GPU",
                    "l2": "flushed between steps (256 MiB memset), per-step CUDA events summed",
                    "timing": "CUDA events on the engine stream"},
            "value_no_flush": main["value_no_flush"], "wall_s_timed_region": main["wall_s"],
            "clocks": main["clocks"], "e2e": main["e2e"], "gpu_launches": main["launches"],
            "roofline": main["roofline"], "cpu_baseline": main["cpu_baseline"], "state": main["state"],
            "secondary": secondary}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "dam1m", "box32m", "pipe4m"])
    ap.add_argument("--particles", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="N = 1: skip the dam1m record beside the headline")
    ap.add_argument("--straight-steps", type=int, default=1000,
                    help="steps of the straight-through dam1m run in the secondary record (BASELINE configs[1])")
    ap.add_argument("--no-single", action="store_true", help="multi-GPU: skip the 1-GPU run of the same workload")
    ap.add_argument("--window", type=int, default=WINDOW,
                    help="restore the start state every WINDOW steps (0 = never; see module docstring)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
