"""In-tree build of the CUDA extension and host shims (explicit nvcc / g++; no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd: list[str]) -> None:
    print("+", " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)


def build_all(force: bool = False) -> None:
    hdrs = sorted(os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".cuh", ".h"))) + [os.path.join(ROOT, "include", "skgpu_batch.h")]
    # 1. the product: libskgpu.so (CUDA kernels + batch C ABI)
    tgt = os.path.join(CSRC, "libskgpu.so")
    src = os.path.join(CSRC, "skgpu.cu")
    if force or _stale(tgt, [src] + hdrs):
        _run([_nvcc()] + NVCC_FLAGS + ["-o", tgt, src])
    # 2. host-only test shim of the phase-run generator (same header the kernels compile)
    tgt = os.path.join(CSRC, "libsk_phase_host.so")
    src = os.path.join(CSRC, "phase_runs_host.cpp")
    if force or _stale(tgt, [src, os.path.join(CSRC, "phase_runs.h")]):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", tgt, src])
    # 3. optional extra targets (plugins, host mirror) register themselves here
    extra = os.path.join(CSRC, "Makefile")
    if os.path.exists(extra):
        _run(["make", "-s", "-C", CSRC] + (["-B"] if force else []))


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
