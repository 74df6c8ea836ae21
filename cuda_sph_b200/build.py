"""Builds libsph_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsph_b200.so")
SOURCES = ["sph_engine.cu"]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    # every header of csrc/ (all are reachable from sph_engine.cu) + the public header + this script
    deps = [os.path.join(CSRC, s) for s in os.listdir(CSRC) if s.endswith((".cu", ".cuh", ".h"))]
    deps += [os.path.join(HERE, "..", "include", "sph_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name: str, defines: list[str]) -> str:
    """A/B build with extra -D flags (tile sizes, staging capacities ...) next to the product library."""
    return build(force=True, out=os.path.join(HERE, f"libsph_b200_{name}.so"), defines=defines)


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines: list[str] | None = None) -> str:
    if not force and not needs_build():
        return LIB
    LIB_OUT = out or LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "--shared", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
           "-I", os.path.join(HERE, "..", "include"), "-o", LIB_OUT]
    cmd += [f"-D{d}" for d in (defines or [])]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_OUT


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
