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

LIB_FAST = PKG / "libforge3d_b200_fast.so"     # the throughput-numerics variant (csrc/f3d_math.cuh F3D_FAST_NUMERICS)

_COMMON_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"]
_LINK_FLAGS = ["-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-ldl"]
_NUMERICS_FLAGS = {
    "exact": ["-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false"],
    "fast": ["-fmad=true", "-prec-div=false", "-prec-sqrt=false", "-ftz=true", "-DF3D_FAST_NUMERICS=1"],
}
NVCC_FLAGS = [*_COMMON_FLAGS, *_NUMERICS_FLAGS["exact"], *_LINK_FLAGS]


def lib_path(numerics: str = "exact") -> Path:
    if numerics not in _NUMERICS_FLAGS:
        raise ValueError(f"numerics must be 'exact' or 'fast', got {numerics!r}")
    return LIB if numerics == "exact" else LIB_FAST


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


HOT_SOURCES = ["f3d_backend.cu", "f3d_host.h", "f3d_math.cuh", "f3d_trace.cuh", "f3d_trace_fast.cuh", "f3d_kernels.cuh", "f3d_aether.cuh", "f3d_lbvh.cuh"]


def hot_hash() -> str:
    """Content hash of the sources the TERRAIN path's kernels are made from (what an ncu capture of the frame kernels depends
    on); a change to the smoke / viewshed / wavefront rows does not invalidate profiles/traffic.json."""
    h = hashlib.sha1()
    for n in HOT_SOURCES:
        h.update(n.encode())
        h.update((CSRC / n).read_bytes())
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


def _info(defines: str, numerics: str) -> str:
    return f"src={source_hash()};hot={hot_hash()};defines={','.join(defines.split())}" + (";numerics=fast" if numerics == "fast" else "")


def needs_build(defines: str = "", numerics: str = "exact") -> bool:
    """True unless the library was built from exactly these sources with exactly these defines (a variant build left behind
    by an A/B run is rebuilt, not silently reused as the default)."""
    return built_info(lib_path(numerics)) != _info(defines, numerics)


def build(force: bool = False, verbose: bool = False, out: Path | None = None, defines: str | None = None,
          numerics: str = "exact") -> Path:
    # defines / F3D_B200_DEFINES="F3D_CULL_FAST=0 ..." builds a compile-time variant of the kernels (A/B runs; see
    # csrc/f3d_trace_fast.cuh and tests/test_traversal_emulation.py); unset = the validated default
    if defines is None:
        defines = os.environ.get("F3D_B200_DEFINES", "")
    defines = " ".join(defines.split())
    default_out = lib_path(numerics)
    out = Path(out) if out else default_out
    if not force and out == default_out and not needs_build(defines, numerics):
        return out
    info = _info(defines, numerics)
    dflags = [f"-D{d}" for d in defines.split()] + [f'-DF3D_BUILD_INFO_STR="{info}"']
    cmd = [_nvcc(), *_COMMON_FLAGS, *_NUMERICS_FLAGS[numerics], *_LINK_FLAGS, *dflags, "-o", str(out)] + [str(CSRC / s) for s in SOURCES]
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


VARIANT_DIR = PKG.parent / "variants"
# Compile-time variants that ship beside the default library: measured alternatives kept buildable and bit-exact
# (tests/test_gpu_parity.py::test_compile_time_variants_are_bit_identical), named by what they switch.
VARIANTS = {
    "tma": "F3D_TMA_STAGE=1",            # TMA (cp.async.bulk + mbarrier) staging of the top pyramid levels: measured slower
}


def variant_path(name: str) -> Path:
    return VARIANT_DIR / f"lib_{name}.so"


def build_variant(name: str, force: bool = False) -> Path:
    """Builds variants/lib_<name>.so from the current sources with VARIANTS[name] (no-op when it is up to date)."""
    defines = " ".join(VARIANTS[name].split())
    out = variant_path(name)
    if not force and built_info(out) == _info(defines, "exact"):
        return out
    VARIANT_DIR.mkdir(exist_ok=True)
    return build(force=True, out=out, defines=defines)


def build_all(force: bool = False) -> list:
    """Everything the GPU box needs: the exact library, the throughput-numerics library, the shipped variants."""
    outs = [build(force=force, defines=""), build(force=force, defines="", numerics="fast")]
    outs += [build_variant(n, force=force) for n in VARIANTS]
    return outs


if __name__ == "__main__":
    import sys

    print(build(force=True, verbose=True, numerics="fast" if "--fast" in sys.argv else "exact"))
