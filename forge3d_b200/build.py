"""Builds libforge3d_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

The numerics contract requires IEEE division/sqrt and NO FMA contraction,
hence -fmad=false -prec-div=true -prec-sqrt=true -ftz=false.  cudart is linked statically so the
library loads through ctypes without torch.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libforge3d_b200.so"
SOURCES = ["f3d_backend.cu", "f3d_smoke.cu", "f3d_viewshed.cu", "f3d_wavefront.cu"]
HEADERS = ["f3d_host.h", "f3d_math.cuh", "f3d_aether.cuh", "f3d_smoke.cuh", "f3d_viewshed.cuh", "f3d_lbvh.cuh", "f3d_wavefront.cuh", "f3d_trace.cuh", "f3d_trace_fast.cuh", "f3d_kernels.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build libforge3d_b200.so")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    deps = [CSRC / n for n in SOURCES + HEADERS] + [PKG.parent / "include" / "forge3d_b200.h"]
    return LIB.stat().st_mtime < max(p.stat().st_mtime for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    # F3D_B200_DEFINES="F3D_ANYHIT_SIGN_ORDER=1 ..." builds a compile-time variant of the kernels (A/B runs; see
    # csrc/f3d_trace_fast.cuh and tests/test_traversal_emulation.py); unset = the validated default
    defines = [f"-D{d}" for d in os.environ.get("F3D_B200_DEFINES", "").split()]
    cmd = [_nvcc(), *NVCC_FLAGS, *defines, "-o", str(LIB)] + [str(CSRC / s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}")
    if verbose:
        print(res.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
