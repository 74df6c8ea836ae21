"""Where the x-slab step loses time against the plain engine, measured on ONE GPU:
  (a) plain engine on a cube                         -- the reference point
  (b) world-1 NativeSlabRunner on the same cube      -- cost of the slab machinery (slots with holes, global ids, in-cell
                                                        order repair, routing) without any halo
  (c) plain engine on a NARROW box (24 x 162 x 162 cells, the geometry of one rank of an 8-way split of box32m)
                                                     -- cost of short grid rows (tiles wrap rows, fewer shared blocks)
Per-stage CUDA-event times, best of a few repetitions of step 1 from the start state."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cuda_sph_b200 import B200SPHStrategy, SphConstants, config, workloads  # noqa: E402
from cuda_sph_b200.data_classes import SimulationState  # noqa: E402
from cuda_sph_b200.slab import NativeSlabRunner  # noqa: E402

KEYS = ("hash_ms", "sort_ms", "reorder_ms", "density_ms", "force_ms", "total_ms")


def plain(params, st, tag):
    s = B200SPHStrategy(params, SphConstants(mode="BOX"))
    s.upload(st)
    s.save_state()
    best = None
    for rep in range(4):
        s.restore_state()
        t = s.step_timed(1)
        if rep and (best is None or t["total_ms"] < best["total_ms"]):
            best = t
    n = int(params.particle_count)
    print(f"{tag:34s} N={n:9d} " + " ".join(f"{k[:-3]} {best[k]:7.3f}" for k in KEYS)
          + f"  ns/particle {1e6 * best['total_ms'] / n:6.3f}  paths {s.path_counters()}", flush=True)
    s.close()


def slab1(params, st, tag):
    n = int(params.particle_count)
    n_cols = int(np.ceil(params.space_size[0] / params.voxel_size[0]))
    cols = np.clip((st.position[:, 0] / params.voxel_size[0]).astype(np.int64), 0, n_cols - 1)
    hist = np.bincount(cols, minlength=n_cols)
    run = NativeSlabRunner(params, SphConstants(mode="BOX"), col_hist=hist, bounds=[0, n_cols], device=0)
    run.load_global(st.position, st.velocity)
    snap = run.snapshot()
    best = None
    for rep in range(4):
        run.restore(snap)
        t = run.step_timed()
        t["total_ms"] = sum(v for k, v in t.items() if k.endswith("_ms"))
        if rep and (best is None or t["total_ms"] < best["total_ms"]):
            best = t
    print(f"{tag:34s} N={n:9d} " + " ".join(f"{k[:-3]} {v:7.3f}" for k, v in sorted(best.items()))
          + f"  ns/particle {1e6 * best['total_ms'] / n:6.3f}", flush=True)
    run.close()


if __name__ == "__main__":
    n = 1 << 24
    params, st = workloads.uniform_box(n, 8.0, seed=0)
    plain(params, st, "(a) plain, cube")
    slab1(params, st, "(b) world-1 slab runner, cube")
    del st
    torch.cuda.empty_cache()
    cells = (24, 162, 162)
    n2 = cells[0] * cells[1] * cells[2] * 8
    space = [c * config.INF_R for c in cells]
    rng = np.random.default_rng(0)
    pos = (rng.random((n2, 3), dtype=np.float32) * np.asarray(space, np.float32)).astype(np.float32)
    pos = np.minimum(pos, np.nextafter(np.asarray(space, np.float32), np.float32(0)))
    vel = (rng.random((n2, 3), dtype=np.float32) - np.float32(0.5)) + np.asarray([1.5, -5.0, -5.0], np.float32)
    p2 = config.box_params(n2, space)
    plain(p2, SimulationState(pos.astype(np.float64), vel.astype(np.float64), np.zeros(n2)), "(c) plain, narrow box 24x162x162")
