#!/usr/bin/env python3
"""Developer A/B helper (run under gpurun): times the resident C2 frame loop for one build of the library, without torch.

usage: F3D_B200_LIB=variants/lib_x.so python tools/ab_bench.py [--frames 128] [--warmup 8] [--width 1920 --height 1080] [--part R/W]
Prints one line: tag, ms/frame (CUDA events inside the library), Mrays/s, nodes/ray, checksum of the accumulation image.
The checksum must be identical across variants (every variant is required to be bit-exact)."""
import argparse
import hashlib
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=1)
    ap.add_argument("--part", default="0/1")
    ap.add_argument("--repeat", type=int, default=2)
    args = ap.parse_args()
    import bench
    from forge3d_b200 import _native
    from forge3d_b200.session import Session

    dem, cam, kw = bench.workload()
    kw["spp"] = args.spp
    r, w = map(int, args.part.split("/"))
    total = args.warmup + args.frames * args.repeat
    s = Session(dem, args.width, args.height, cam, part_rank=r, part_world=w, **kw, max_frames=total, min_frames=total,
                variance_threshold=1e30)
    s.render_frames(args.warmup)
    s.sync()
    best = None
    for _ in range(args.repeat):
        s0 = s.stats()
        s.render_frames(args.frames)
        s.sync()
        ms = s.last_frames_ms()
        s1 = s.stats()
        if best is None or ms < best[0]:
            best = (ms, s0, s1)
    ms, s0, s1 = best
    rays = sum(s1[k] - s0[k] for k in ("rays_primary", "rays_shadow", "rays_ibl"))
    nodes = s1["nodes_popped"] - s0["nodes_popped"]
    out = s.resolve_host(want_accum=True)
    s.close()
    tag = os.environ.get("F3D_B200_LIB", "default")
    print(f"AB {Path(tag).stem:14s} ms/frame {ms / args.frames:.4f}  Mrays/s {rays / ms / 1e3:9.1f}  nodes/ray {nodes / rays:5.2f}  "
          f"accum sha {hashlib.sha1(out['accum'].tobytes()).hexdigest()[:12]}  build {_native.lib().f3d_build_info().decode()}", flush=True)


if __name__ == "__main__":
    main()
