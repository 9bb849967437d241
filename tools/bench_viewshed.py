#!/usr/bin/env python
"""Side bench for the HELIOS viewshed and solar shadow mask: one observer / one sun position over an
n x n DEM (every cell is one ray), kernel time from CUDA events inside the library, end to end around the public call with host
arrays, and the CPU oracle on a bounded sample.  NOT bench.py.  usage: python tools/bench_viewshed.py [--n 1024] [--cpu-n 256]"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--cpu-n", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import _helpers as H
    from forge3d_b200 import viewshed as V

    def inputs(n):
        dem = H.rainier_dem(n)
        b = (7.0, 45.6, 8.0, 46.4)
        return (V.viewshed_inputs(dem, (46.0, 7.5), bounds=b, height_system="ellipsoidal", observer_height=30.0),
                V.shadow_mask_inputs(dem, 245.0, 14.0, bounds=b, height_system="ellipsoidal"))

    (h, pos, opts), (sh, sinp, sopts) = inputs(args.n)
    V.compute_viewshed(h, pos, opts)                                   # warm-up: context, buffers
    best = {"viewshed": 1e30, "shadow_mask": 1e30}
    e2e = dict(best)
    for _ in range(args.steps):
        t0 = time.perf_counter()
        out = V.compute_viewshed(h, pos, opts)
        e2e["viewshed"] = min(e2e["viewshed"], (time.perf_counter() - t0) * 1e3)
        best["viewshed"] = min(best["viewshed"], max(out["kernel_ms"], 1e-9))
        t0 = time.perf_counter()
        V.compute_shadow_mask(sh, sinp, sopts)
        e2e["shadow_mask"] = min(e2e["shadow_mask"], (time.perf_counter() - t0) * 1e3)
    cells = args.n * args.n
    line = {"metric": "viewshed Mcells/s (one any-hit ray per DEM cell)", "value": cells / best["viewshed"] / 1e3, "unit": "Mcells/s",
            "ms_per_step": best["viewshed"], "n_gpus": 1, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"rainier-shaped {args.n}x{args.n} DEM, WGS84 ellipsoid + standard refraction", "visible_fraction": float(np.mean(out["visibility"]))},
            "e2e": {"viewshed_ms": e2e["viewshed"], "shadow_mask_ms": e2e["shadow_mask"], "viewshed_mcells_per_s": cells / e2e["viewshed"] / 1e3,
                    "shadow_mask_mcells_per_s": cells / e2e["shadow_mask"] / 1e3, "h2d_bytes_per_step": 12 * cells, "d2h_bytes_per_step": 13 * cells}}
    # issue-slot roofline: warp-instructions of this exact workload (ncu smsp__inst_executed.sum of k_viewshed on the B200, the count
    # is deterministic for a fixed DEM / observer; profiles/r02_rows_instr.json) over the SM issue rate 148 SMs x 4 schedulers x 1.965 GHz
    try:
        rows = json.loads((ROOT / "profiles" / "r02_rows_instr.json").read_text())
        wi = rows.get(f"k_viewshed_{args.n}")
        if wi:
            peak = 148 * 4 * 1.965e9
            line["roofline"] = {"bound": "issue", "achieved": wi / (best["viewshed"] * 1e-3), "peak": peak, "unit": "warp-instr/s",
                                "frac": wi / (best["viewshed"] * 1e-3) / peak, "traffic": None,
                                "model": f"{wi:.4g} warp-instructions per launch (ncu, 27.1 of 32 lanes active) / kernel time / (148 x 4 x 1.965e9)"}
    except Exception:
        pass
    try:
        from oracle import oracle

        (ch, cpos, copts), _ = inputs(args.cpu_n)
        t0 = time.perf_counter()
        oracle.viewshed(ch, cpos, copts)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": args.cpu_n ** 2 / dt / 1e6, "unit": "Mcells/s", "cores": oracle.get_threads(), "kind": "port",
                                "sample": f"{args.cpu_n}x{args.cpu_n} DEM of the same shape"}
    except Exception as exc:
        line["cpu_baseline"] = {"unavailable": str(exc)[:200]}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
