#!/usr/bin/env python3
"""SIMT-efficiency model of the persistent ray scheduler (k_trace), evaluated on the CPU.

Builds the product's CUDA source with F3D_SCHED_STATS for the SIMT interpreter (tests/c/emu: real 32-lane warps, so the
per-warp event counts are the ones a GPU would see), renders a reduced-size frame of the bench workload for each compile-time
variant given on the command line, checks the accumulation buffer is bit-identical to the default build, and prints

    expansion steps x lanes/step | leaf phases x lanes/phase | refill steps | modeled warp-instruction cost

with cost = steps * (C_expand + C_iter) + phases * C_leaf + refills * C_refill (static SASS counts per site, see
profiles/r01_static_kernels.txt; C_refill calibrated so that the GPU-tuned F3D_REFILL_BELOW=24 beats 28).  The emulator reproduces
the lane counts ncu measured on the B200 for the default build (expansion ~24 vs 22 lanes, leaf ~8.1 vs 8 lanes), so the model
ranks scheduling policies before GPU time is spent on them.

usage: python tools/sched_model.py ["DEF=1 DEF2=3" ...]      (each argument is one variant; the default build is always run first)
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

C_EXPAND, C_LEAF, C_REFILL, C_ITER_DEFER = 150.0, 330.0, 300.0, 14.0


def main():
    import _emu
    import bench

    dem, cam, kw = bench.workload()
    W, H, frames = 192, 108, 2
    base = None
    for variant in [""] + sys.argv[1:]:
        defs = tuple(variant.split()) + ("F3D_SCHED_STATS",)
        with _emu.emulated_backend(defs) as native:
            g = native.hybrid_render_terrain_reference(dem, W, H, cam, max_frames=frames, min_frames=frames, variance_threshold=1e30,
                                                       want_accum=True, **kw)
            out = (C.c_ulonglong * 40)()
            native.lib().f3d_debug_sched_stats(out, 1)
        es, el, ls, ll, rs, _ = list(out)[:6]
        for name, v in (("sun", out[6]), ("ibl", out[7])):
            if v >> 32:
                print(f"    bottom-up start, {name} rays: {(v & 0xFFFFFFFF) / (v >> 32):.2f} seeds per ray ({v >> 32} rays)")
        if sum(out[8:24]):
            print("    sun horizon: cleared from column k on -> rays (seeds/ray): " + "  ".join(
                f"{'none' if i == 15 else [1, 2, 3, 4, 6, 8, 12, 16][i]}: {out[8 + i]} ({out[24 + i] / max(out[8 + i], 1):.2f})" for i in range(16) if out[8 + i]))
        if base is None:
            base = g["accum"].copy()
        exact = np.array_equal(g["accum"].view(np.uint32), base.view(np.uint32))
        it = C_ITER_DEFER if "F3D_TRACE_DEFER_LEAVES" in variant else 0.0
        cost = es * (C_EXPAND + it) + ls * C_LEAF + rs * C_REFILL
        print(f"{variant or '(default)':78s} exact={exact}  expand {es:6d} x {el / es:5.2f}  leaf {ls:5d} x {ll / max(ls, 1):5.2f}  "
              f"refill {rs:5d}  cost {cost / 1e6:6.2f} M warp-instr", flush=True)


if __name__ == "__main__":
    main()
