"""Builds and loads tests/c/emu/trace_emu.cpp: the product's DEVICE traversal headers compiled with g++ and run on the
CPU as a one-lane warp (see tests/c/emu/cuda_runtime.h).  Test infrastructure."""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
EMU = ROOT / "tests" / "c" / "emu"
CSRC = ROOT / "forge3d_b200" / "csrc"
BUILD = EMU / "_build"


def _tag(defines, deps) -> str:
    """Build products are keyed by the CONTENT of everything they are made from (mtimes lie after a checkout)."""
    h = hashlib.sha1(" ".join(sorted(defines)).encode())
    for p in deps:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()[:12]


def build(defines=()) -> Path:
    deps = [EMU / "trace_emu.cpp", EMU / "cuda_runtime.h"] + sorted(CSRC.glob("f3d_*.cuh"))
    out = BUILD / f"libtrace_emu_{_tag(defines, deps)}.so"
    if out.exists():
        return out
    BUILD.mkdir(exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", f"-I{EMU}", f"-I{CSRC}",
           *[f"-D{d}" for d in defines], "-o", str(out), str(EMU / "trace_emu.cpp")]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError(f"g++ failed:\n{' '.join(cmd)}\n{res.stdout}")
    return out


def trace_rays(heights, spacing, origin_xz, exaggeration, rays, *, any_hit, apply_curvature, inv_two_r_prime=0.0,
               curvature_enabled=False, defines=()):
    """Same contract as oracle.trace_rays / _native.trace_rays; returns (hit, t, normal, nodes_popped)."""
    L = C.CDLL(str(build(defines)))
    fp = C.POINTER(C.c_float)
    L.emu_trace_rays.argtypes = [fp, C.c_uint32, C.c_uint32, fp, fp, C.c_float, C.c_float, C.c_int32, fp, C.c_uint64,
                                 C.c_int32, C.c_int32, C.POINTER(C.c_uint8), fp, fp, C.POINTER(C.c_uint64)]
    dem = np.ascontiguousarray(heights, np.float32)
    r = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    n = r.shape[0]
    hit = np.zeros(n, np.uint8)
    t = np.zeros(n, np.float32)
    nrm = np.zeros((n, 3), np.float32)
    nodes = C.c_uint64()
    f = lambda a: a.ctypes.data_as(fp)
    sp = (C.c_float * 2)(*map(float, spacing))
    og = (C.c_float * 2)(*map(float, origin_xz))
    rc = L.emu_trace_rays(f(dem), dem.shape[1], dem.shape[0], sp, og, float(exaggeration), float(inv_two_r_prime),
                          int(bool(curvature_enabled)), f(r), n, int(bool(any_hit)), int(bool(apply_curvature)),
                          hit.ctypes.data_as(C.POINTER(C.c_uint8)), f(t), f(nrm), C.byref(nodes))
    if rc != 0:
        raise RuntimeError(f"emu_trace_rays failed ({rc})")
    return hit.astype(bool), t, nrm, int(nodes.value)


def trace_rays_bottom_up(heights, spacing, origin_xz, exaggeration, rays, *, apply_curvature, inv_two_r_prime=0.0,
                         curvature_enabled=False, defines=()):
    """Any-hit rays through the bottom-up start (k_ascent + k_trace's seed traversal, one lane); returns (hit, nodes)."""
    L = C.CDLL(str(build(defines)))
    fp = C.POINTER(C.c_float)
    L.emu_trace_rays_bottom_up.argtypes = [fp, C.c_uint32, C.c_uint32, fp, fp, C.c_float, C.c_float, C.c_int32, fp, C.c_uint64,
                                           C.c_int32, C.POINTER(C.c_uint8), fp, C.POINTER(C.c_uint64)]
    dem = np.ascontiguousarray(heights, np.float32)
    r = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    n = r.shape[0]
    hit = np.zeros(n, np.uint8)
    t = np.zeros(n, np.float32)
    nodes = C.c_uint64()
    f = lambda a: a.ctypes.data_as(fp)
    sp = (C.c_float * 2)(*map(float, spacing))
    og = (C.c_float * 2)(*map(float, origin_xz))
    rc = L.emu_trace_rays_bottom_up(f(dem), dem.shape[1], dem.shape[0], sp, og, float(exaggeration), float(inv_two_r_prime),
                                    int(bool(curvature_enabled)), f(r), n, int(bool(apply_curvature)),
                                    hit.ctypes.data_as(C.POINTER(C.c_uint8)), f(t), C.byref(nodes))
    if rc != 0:
        raise RuntimeError(f"emu_trace_rays_bottom_up failed ({rc})")
    return hit.astype(bool), int(nodes.value)


# ---------------------------------------------------------------------------------------------------------------
# Whole backend under the SIMT interpreter: the product's host driver (f3d_backend.cu, launches rewritten by
# tests/c/emu/gen_backend.py) and ALL its kernels compiled by g++; every CUDA thread is a fiber (tests/c/emu/simt.h).
# ---------------------------------------------------------------------------------------------------------------
def build_backend(defines=()) -> Path:
    deps = [EMU / n for n in ("gen_backend.py", "simt.cpp", "simt.h", "cuda_runtime.h", "cuda_fake_runtime.h", "cuda_fp16.h")]
    deps += sorted(CSRC.glob("f3d_*.cu*")) + [CSRC / "f3d_host.h", ROOT / "include" / "forge3d_b200.h"]
    tag = _tag(defines, deps)
    out = BUILD / f"libforge3d_b200_emu_{tag}.so"
    if out.exists():
        return out
    BUILD.mkdir(exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    import sys
    from forge3d_b200.build import SOURCES
    gens = []
    for name in SOURCES:          # every translation unit of the product library, launches rewritten for the interpreter
        gen = BUILD / f"{Path(name).stem}_emu_{tag}.cpp"
        subprocess.run([sys.executable, str(EMU / "gen_backend.py"), str(CSRC / name), str(gen)], check=True, env=env)
        gens.append(str(gen))
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-DEMU_SIMT", f"-I{EMU}",
           f"-I{CSRC}", f"-I{ROOT / 'include'}", *[f"-D{d}" for d in defines], "-o", str(out), *gens, str(EMU / "simt.cpp")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError(f"g++ failed:\n{' '.join(cmd)}\n{res.stdout}")
    return out


class emulated_backend:
    """Context manager: forge3d_b200._native talks to the emulated library instead of libforge3d_b200.so.  Only tests
    do this (the product has no such switch): it is how the CPU suite runs the real host driver + kernels."""

    def __init__(self, defines=()):
        self.defines = tuple(defines)

    def __enter__(self):
        from forge3d_b200 import _native

        self._native = _native
        self._saved = (_native.LIB_PATH, dict(_native._libs))
        _native.LIB_PATH = build_backend(self.defines)
        _native._libs.clear()
        _native.lib()
        return _native

    def __exit__(self, *exc):
        self._native.LIB_PATH = self._saved[0]
        self._native._libs.clear()
        self._native._libs.update(self._saved[1])
        return False
