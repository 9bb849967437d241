"""Builds libforge3d_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

The numerics contract requires IEEE division/sqrt and NO FMA contraction,
hence -fmad=false -prec-div=true -prec-sqrt=true -ftz=false.  cudart is linked statically so the
library loads through ctypes without torch.
"""
from __future__ import annotations

import hashlib
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


def _deps():
    return [CSRC / n for n in SOURCES + HEADERS] + [PKG.parent / "include" / "forge3d_b200.h"]


def source_hash() -> str:
    """Content hash of every source the library is made from (mtimes lie after a checkout)."""
    h = hashlib.sha1()
    for p in _deps():
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()[:16]


def built_info(lib: Path = LIB) -> str:
    """The `f3d_build_info()` string of a built library without loading it: "src=<hash>;defines=<...>" ("" if absent)."""
    try:
        blob = lib.read_bytes()
    except OSError:
        return ""
    i = blob.find(b"F3D_BUILD_INFO:")
    if i < 0:
        return ""
    j = blob.find(b"\0", i)
    return blob[i + len(b"F3D_BUILD_INFO:"):j].decode(errors="replace")


def needs_build(defines: str = "") -> bool:
    """True unless LIB was built from exactly these sources with exactly these defines (a variant build left behind by an
    A/B run is rebuilt, not silently reused as the default)."""
    return built_info() != f"src={source_hash()};defines={','.join(defines.split())}"


def build(force: bool = False, verbose: bool = False, out: Path | None = None, defines: str | None = None) -> Path:
    # defines / F3D_B200_DEFINES="F3D_CULL_FAST=0 ..." builds a compile-time variant of the kernels (A/B runs; see
    # csrc/f3d_trace_fast.cuh and tests/test_traversal_emulation.py); unset = the validated default
    if defines is None:
        defines = os.environ.get("F3D_B200_DEFINES", "")
    defines = " ".join(defines.split())
    out = Path(out) if out else LIB
    if not force and out == LIB and not needs_build(defines):
        return LIB
    info = f"src={source_hash()};defines={','.join(defines.split())}"
    dflags = [f"-D{d}" for d in defines.split()] + [f'-DF3D_BUILD_INFO_STR="{info}"']
    cmd = [_nvcc(), *NVCC_FLAGS, *dflags, "-o", str(out)] + [str(CSRC / s) for s in SOURCES]
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
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
