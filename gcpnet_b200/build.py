"""Build gcpnet_b200/libgcpnet_b200.so (the C-ABI library of include/gcpnet_b200.h) with nvcc for
sm_100a.  In-tree, so the built library travels with the repo snapshot to the GPU box.

    python -m gcpnet_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libgcpnet_b200.so")
HEADER = os.path.join(os.path.dirname(PKG), "include", "gcpnet_b200.h")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--use_fast_math=false", "-Xcompiler", "-fPIC", "--shared", "-Wno-deprecated-gpu-targets",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    return [HEADER] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]


def _deps_of(src: str) -> list:
    """Transitive closure of the quoted #includes of one source file (in-tree headers only)."""
    import re
    seen, todo = set(), [src]
    while todo:
        f = todo.pop()
        if f in seen or not os.path.exists(f):
            continue
        seen.add(f)
        for inc in re.findall(r'^\s*#include\s+"([^"]+)"', open(f).read(), flags=re.M):
            todo.append(os.path.normpath(os.path.join(os.path.dirname(f), inc)))
    return sorted(seen)


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libgcpnet_b200.so")
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    flags += os.environ.get("GCPNET_NVCC_FLAGS", "").split()  # e.g. -DGCP_STAMPS=1 for the stage-timing scripts
    objs, jobs = [], []
    for src in sources():
        obj = os.path.join(PKG, "build", os.path.basename(src) + ".o")
        os.makedirs(os.path.dirname(obj), exist_ok=True)
        if force or not os.path.exists(obj) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in _deps_of(src)):
            cmd = [nvcc] + [f for f in flags if f != "--shared"] + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
            jobs.append((cmd, subprocess.Popen(cmd)))  # translation units compile concurrently
        objs.append(obj)
    for cmd, proc in jobs:
        if proc.wait() != 0:
            raise subprocess.CalledProcessError(proc.returncode, cmd)
    subprocess.check_call([nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
