#!/usr/bin/env python3
"""Throughput-numerics build (libforge3d_b200_fast.so) against the exact build on the same scene and seed (run under gpurun).
Prints one JSON line: RGBA RMSE (0..1 scale, north-star tolerance 1e-3), share of differing bytes, max |diff|, depth AOV
relative error, hit-mask mismatches (pixels that are terrain in one build and sky in the other), ray-count differences."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def compare(exact, fast):
    d = exact["rgba"][..., :3].astype(np.float64) - fast["rgba"][..., :3].astype(np.float64)
    de, df = exact["depth"].astype(np.float64), fast["depth"].astype(np.float64)
    hit_e, hit_f = np.isfinite(de) & (de > 0) & (de < 1e29), np.isfinite(df) & (df > 0) & (df < 1e29)
    both = hit_e & hit_f
    rel = np.abs(de[both] - df[both]) / np.maximum(np.abs(de[both]), 1e-6)
    return {
        "rgb_rmse": float(np.sqrt(np.mean((d / 255.0) ** 2))),
        "bytes_differing": float((d != 0).mean()),
        "max_abs_diff_of_255": int(np.abs(d).max()),
        "hit_mask_mismatches": int((hit_e != hit_f).sum()),
        "depth_rel_err_max": float(rel.max()) if rel.size else 0.0,
        "depth_rel_err_mean": float(rel.mean()) if rel.size else 0.0,
        "normal_max_abs_diff": float(np.abs(exact["normal"].astype(np.float64) - fast["normal"].astype(np.float64)).max()),
        "albedo_identical": bool(np.array_equal(exact["albedo"], fast["albedo"])),
    }


def render(numerics, dem, W, Hh, cam, kw, frames):
    from forge3d_b200.session import Session

    s = Session(dem, W, Hh, cam, numerics=numerics, **kw, max_frames=frames, min_frames=frames, variance_threshold=1e30)
    s.render_frames(frames)
    s.sync()
    ms = s.last_frames_ms()
    out = {k: np.array(v) for k, v in s.resolve_host().items() if isinstance(v, np.ndarray)}
    st = s.stats()
    s.close()
    return out, st, ms / frames


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="c2", choices=["c2", "golden"])
    ap.add_argument("--frames", type=int, default=256)
    args = ap.parse_args()
    import _helpers as H
    import bench

    if args.scene == "c2":
        dem, cam, kw = bench.workload()
        W, Hh = 1920, 1080
    else:
        dem = H.golden_dem()
        cam, kw, W, Hh = H.CAM, H.scene_kwargs(dem), 256, 256
        for k in ("max_frames", "min_frames", "variance_threshold"):
            kw.pop(k, None)
    e, se, ms_e = render("exact", dem, W, Hh, cam, kw, args.frames)
    f, sf, ms_f = render("fast", dem, W, Hh, cam, kw, args.frames)
    res = compare(e, f)
    res.update(scene=args.scene, frames=args.frames, ms_per_frame_exact=ms_e, ms_per_frame_fast=ms_f,
               rays_exact={k: int(se[k]) for k in ("rays_primary", "rays_shadow", "rays_ibl")},
               rays_fast={k: int(sf[k]) for k in ("rays_primary", "rays_shadow", "rays_ibl")})
    print("NUMERICS " + json.dumps(res))


if __name__ == "__main__":
    main()
