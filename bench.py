#!/usr/bin/env python
"""bench.py -- particle-updates/s of the SPH step hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # CUDA engine (libsph_b200.so)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port, all host threads)

A "step" is one full Voxel SPH step over the resident particle set: hash -> sort -> cell table + reorder -> density
sweep -> fused pressure/viscosity/integrate/collide sweep (reference: compute_next_state,
sim/src/sph/strategies/abstract_sph_strategy.py:31-46).

Workloads:
    dam1m   BASELINE configs[1]: box dam-break column, 2^20 particles, fp32 engine          (default at N = 1)
    box32m  BASELINE configs[3]: uniform box, 2^25 particles                               (default at N > 1)
    pipe4m  BASELINE configs[2]: six-segment pipe, 2^22 particles, inflow/outflow recycle

Window: the reference's Voxel physics diverges within ~10 steps on every workload (fp64 oracle: |v| ~ 1e21 by step
10, most particles NaN by step 40 -- DESIGN.md "Workloads"), after which a step is mostly dead particles.  So every
arm keeps the state inside steps 1..WINDOW of the workload: after WINDOW steps the start state is restored (a
device-to-device copy outside the timed events).  `--window 0` disables that and runs the steps straight through.

`value`  : N x steps / device time, inputs resident in HBM, L2 flushed between steps (per-step CUDA events summed).
`e2e`    : the same through the reference-facing call compute_next_state (fp64 host buffers in pinned memory,
           H2D + step + D2H every step; copies interleaved with the step on two streams), wall clock around the
           synchronous calls.
`roofline`: dominant kernel (the density sweep), algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json, its
           measured DRAM traffic (ncu, profiles/traffic.json) and the issue-slot roofline that actually binds it.
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
    # onesweep: one histogram pass reads the keys (4); every digit pass reads key + value (8) and writes key + value
    # (8); pass 0 has no value read (values are iota)
    return 4 + passes * 16 - 4


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


def cpu_baseline(params, st, mode, window, budget_s=12.0):
    """Oracle port (fp64, OpenMP) on the host cores, bounded sample of the same workload: the first steps of it."""
    from oracle import oracle as orc
    orc.set_exact_pow(False)
    n = int(params.particle_count)
    pipe = params.pipe.to_numpy() if mode == "PIPE" else None
    P = orc.OracleParams(n=n, mode=mode, space=tuple(params.space_size), ext=tuple(params.external_force),
                         dt=1 / params.fps, pipe=pipe)
    rng = orc.rng_init(n) if mode == "PIPE" else None
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
    so this is the oracle port (oracle/sph_oracle.c) with all host threads, on the same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.set_exact_pow(False)
    name = args.workload or ("dam1m" if args.gpus == 1 else "box32m")
    params, st, mode, desc = make_workload(name, args.particles)
    n = int(params.particle_count)
    pipe = params.pipe.to_numpy() if mode == "PIPE" else None
    P = orc.OracleParams(n=n, mode=mode, space=tuple(params.space_size), ext=tuple(params.external_force),
                         dt=1 / params.fps, pipe=pipe)
    rng = orc.rng_init(n) if mode == "PIPE" else None
    pos, vel = st.position, st.velocity
    # bounded: cap the total CPU time at ~150 s by shrinking the number of timed steps if one step is slow
    t0 = time.perf_counter()
    r = orc.step(P, pos, vel, rng=rng, light=True)
    pos, vel = r.position, r.velocity
    one = time.perf_counter() - t0
    warm = max(0, min(args.warmup, int(20.0 / max(one, 1e-9))) - 1)
    steps = max(1, min(args.steps, int(150.0 / max(one, 1e-9))))
    for _ in range(warm):
        r = orc.step(P, pos, vel, rng=rng, light=True)
        pos, vel = r.position, r.velocity
    el = 0.0
    for i in range(steps):
        if args.window and i % args.window == 0:
            pos, vel = st.position, st.velocity
        t0 = time.perf_counter()
        r = orc.step(P, pos, vel, rng=rng, light=True)
        el += time.perf_counter() - t0
        pos, vel = r.position, r.velocity
    value = n * steps / el
    sample = (f"{steps} timed full steps (of {args.steps} requested; bounded to ~150 s) after {warm + 1} warm-up, "
              f"same workload N={n}, steps 1..{args.window or steps} from its start state")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm + 1, "ms_per_step": el / steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "name": name, "particles": n, "window": args.window},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from cuda_sph_b200 import B200SPHStrategy, SphConstants

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if world > 1:
        from cuda_sph_b200 import bench_multi
        return bench_multi.run(args, rank, world, local_rank)

    name = args.workload or "dam1m"
    params, st, mode, desc = make_workload(name, args.particles)
    n = int(params.particle_count)
    stream = torch.cuda.Stream()
    s = B200SPHStrategy(params, SphConstants(mode=mode), device=local_rank, cuda_stream=stream.cuda_stream)
    s.upload(st)
    s.save_state()
    s.synchronize()
    window = args.window

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    K, W = args.steps, max(args.warmup, 3)

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
    ms_steps = [a.elapsed_time(b) for a, b in zip(starts, ends)]
    ms_total = float(sum(ms_steps))
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
    value_hot = n * K / (hot_ms * 1e-3)
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
    traffic, issue = None, None
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
    except Exception:
        pass
    kname = {"density": "density_rows_kernel", "force": "force_rows_kernel"}.get(dom, dom + "_kernel")
    roofline = {"bound": "hbm", "kernel": kname, "achieved": stage_gbs[dom], "peak": hbm_peak,
                "unit": "GB/s", "frac": stage_gbs[dom] / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": stage_bytes[dom] * n, "issue_roofline": issue,
                "algorithmic_bytes_per_particle": stage_bytes[dom], "kernel_ms": stage_ms[dom],
                "stages_ms": stage_ms, "stages_gbs": stage_gbs,
                "step_bytes_per_particle": sum(stage_bytes.values()),
                "step_frac": sum(stage_bytes.values()) * value / 1e9 / hbm_peak,
                "note": "neighbour sweeps are FP32-issue / shared-memory / latency bound at >2 particles/cell "
                        "(SURVEY 8d; ncu: 1-8 % DRAM throughput), the HBM fraction is reported as the contract asks; "
                        "traffic exceeds the algorithmic bytes because the neighbour lists (64 B/particle) and pair "
                        "factors are engine-internal"}

    # ---- e2e: compute_next_state through the C ABI with pinned fp64 host buffers ----------------------------------
    def pinned(shape):
        return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()

    hp, hv = pinned((n, 3)), pinned((n, 3))
    op, ov, orho = pinned((n, 3)), pinned((n, 3)), pinned((n,))
    KE = min(K, 48)
    e2e_s = 0.0
    for i in range(-3, KE):          # 3 untimed warm-up calls
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

    # ---- CPU baseline beside it ---------------------------------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_baseline(params, st, mode, window)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "name": name, "particles": n, "parallelism": "1 GPU",
                       "window": f"state restored to the start state every {window} steps (untimed D2D copy)"
                       if window else "none: steps run straight through",
                       "l2": "flushed between steps (256 MiB memset), per-step CUDA events summed",
                       "timing": "CUDA events on the engine stream"},
            "value_no_flush": value_hot, "wall_s_timed_region": t_wall,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu,
            "state": {"n_dead": stats["n_dead"], "n_nonfinite": stats["n_nonfinite"],
                      "max_density": stats["max_density"], "steps_done": stats["steps_done"]}}
    print(json.dumps(line), flush=True)
    s.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "dam1m", "box32m", "pipe4m"])
    ap.add_argument("--particles", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
