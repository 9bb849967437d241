#!/usr/bin/env python3
"""Gather-only partition against the one-GPU image: every rank of a `world`-way partition is rendered on this
GPU in turn (the mode has no communication, so that is the real computation), rows are assembled, and the RGBA RMSE (0..1 scale),
the share of differing bytes and a seam metric (mean |diff| on rows next to a block border vs elsewhere) are printed."""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="c2", choices=["c2", "golden"])
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--block-rows", type=int, default=16)
    args = ap.parse_args()
    import _helpers as H
    import bench
    from forge3d_b200 import distributed as D
    from forge3d_b200.session import Session

    if args.scene == "c2":
        dem, cam, kw = bench.workload()
        W, Hh = 1920, 1080
    else:
        dem = H.golden_dem()
        cam, kw, W, Hh = H.CAM, H.scene_kwargs(dem), 256, 256
    fixed = dict(max_frames=args.frames, min_frames=args.frames, variance_threshold=1e30)
    s = Session(dem, W, Hh, cam, **kw, **fixed)
    s.render_frames(args.frames)
    ref = s.resolve_host()["rgba"].copy()
    s.close()
    got = np.zeros_like(ref)
    for r in range(args.world):
        s = Session(dem, W, Hh, cam, part_rank=r, part_world=args.world, part_block_rows=args.block_rows, part_mode=1, **kw, **fixed)
        s.render_frames(args.frames)
        rows = D.owned_rows(Hh, args.world, r, args.block_rows)
        got[rows] = s.resolve_host()["rgba"][rows]
        s.close()
    d = got[..., :3].astype(np.float64) - ref[..., :3].astype(np.float64)
    rmse = float(np.sqrt(np.mean((d / 255.0) ** 2)))
    br = D.effective_block_rows(args.block_rows, Hh, args.world)
    y = np.arange(Hh)
    near = ((y % br) < 4) | ((y % br) >= br - 4)
    print(f"gather-only world={args.world} block_rows={br} frames={args.frames} scene={args.scene}: RGB RMSE {rmse:.3e} (tolerance 1e-3), "
          f"bytes differing {float((d != 0).mean()):.4f}, max |diff| {int(np.abs(d).max())}/255, "
          f"mean |diff| near borders {float(np.abs(d[near]).mean()):.4f} vs interior {float(np.abs(d[~near]).mean()):.4f} (of 255)")


if __name__ == "__main__":
    main()
