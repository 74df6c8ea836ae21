"""A/B of compile-time variants of libsph_b200.so (cuda_sph_b200.build.build_variant): per-stage CUDA-event times of
steps 1..2 from the start state for every variant library found, each in its own process (SPH_B200_LIB), and a checksum
of the result so that variants can be compared bitwise.  Not a bench line -- the table goes to profiles/.

    python profiles/ab_variants.py --libs base,fl6,fc4 --cases box8_4m,dam1m
    python profiles/ab_variants.py --libs base,base:SPH_SORT=count,base:SPH_SORT=lookback2     # env knobs after ':'
"""
import argparse
import hashlib
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

CASES = {
    "dam1m": ("dam_break", 1 << 20, 2.5),
    "box8_4m": ("uniform_box", 1 << 22, 8.0),
    "box8_16m": ("uniform_box", 1 << 24, 8.0),
    "box2.5_1m": ("uniform_box", 1 << 20, 2.5),
    "box25_1m": ("uniform_box", 1 << 20, 25.0),
}


def child(cases, env_note):
    import numpy as np
    from cuda_sph_b200 import B200SPHStrategy, SphConstants, workloads
    for name in cases:
        kind, n, ppc = CASES[name]
        params, st = getattr(workloads, kind)(n, ppc, seed=0)
        s = B200SPHStrategy(params, SphConstants(mode="BOX"), record_neighbour_counts=True)
        s.upload(st)
        s.save_state()
        nsteps = 2
        best = [None] * nsteps
        for rep in range(5):
            s.restore_state()
            for k in range(nsteps):
                t = s.step_timed(1)
                if rep and (best[k] is None or t["total_ms"] < best[k]["total_ms"]):
                    best[k] = t
        out = s.download(np.float32)
        h = hashlib.sha1(out.position.tobytes() + out.velocity.tobytes() + out.density.tobytes()
                         + s.neighbour_counts().tobytes()).hexdigest()[:12]
        for k in range(nsteps):
            b = best[k]
            print(f"{env_note:16s} {name:10s} step {k + 1}: density {b['density_ms']:7.4f} force {b['force_ms']:7.4f} "
                  f"sort {b['sort_ms']:7.4f} reorder {b['reorder_ms']:7.4f} hash {b['hash_ms']:7.4f} "
                  f"total {b['total_ms']:7.4f} ms  sha {h}", flush=True)
        s.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", default="base")
    ap.add_argument("--cases", default="box8_4m,dam1m")
    ap.add_argument("--child", default=None)
    a = ap.parse_args()
    if a.child is not None:
        child(a.cases.split(","), a.child)
        sys.exit(0)
    for spec in a.libs.split(","):
        lib, _, knobs = spec.partition(":")
        env = dict(os.environ)
        if lib != "base":
            env["SPH_B200_LIB"] = os.path.join(ROOT, "cuda_sph_b200", f"libsph_b200_{lib}.so")
        note = lib
        for kv in filter(None, knobs.split("+")):
            k, _, v = kv.partition("=")
            env[k] = v
            note += ":" + v
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", note, "--cases", a.cases], env=env)
