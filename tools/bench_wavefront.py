"""Side bench for the wavefront multi-bounce tracer: the reference gate's configuration
(adjudication scene, 512 x 512, spp frames) through the public call, rays/s on the device (CUDA events inside the library) and end to end
(wall clock around the call, host buffers in and out), next to the CPU oracle on a bounded sample of frames.  Not bench.py: the
headline metric stays the terrain path tracer's.  Usage: python tools/bench_wavefront.py [--size 512] [--spp 1024] [--oracle-spp 32]"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT)]
from forge3d_b200 import wavefront as wf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--oracle-spp", type=int, default=32)
    ap.add_argument("--repeats", type=int, default=3)
    a = ap.parse_args()
    scene = wf.scene_from_desc(wf.adjudication_scene())
    import hashlib

    import numpy as np

    small = wf.render_pt_reference(scene, 64, 48, 6)          # warm-up + parity: the oracle's radiance for this call is pinned by checksum
    pin = json.loads((ROOT / "tests" / "golden" / "wavefront_pin.json").read_text())
    parity = hashlib.sha256(np.ascontiguousarray(small).view(np.uint8).tobytes()).hexdigest() == pin["oracle_64x48x6_hdr_sha256"]
    wf.render_pt_reference(scene, a.size, a.size, 4)          # warm-up at the timed size: allocations
    best = None
    for _ in range(a.repeats):
        t0 = time.perf_counter()
        _, _, st = wf.render_pt_reference(scene, a.size, a.size, a.spp, return_rgba8=True, return_stats=True)
        wall = time.perf_counter() - t0
        if best is None or wall < best[0]:
            best = (wall, st)
    wall, st = best
    out = {"metric": "wavefront path tracer Mrays/s", "bit_identical_to_pinned_oracle_64x48x6": bool(parity),
           "workload": f"adjudication scene {a.size}x{a.size}x{a.spp}spp", "rays": st.rays, "launches": st.launches,
           "kernel_ms": round(st.kernel_ms, 3), "e2e_ms": round(wall * 1e3, 3),
           "device_mrays_per_s": round(st.rays / max(st.kernel_ms, 1e-9) / 1e3, 1), "e2e_mrays_per_s": round(st.rays / wall / 1e6, 1),
           "us_per_frame": round(st.kernel_ms * 1e3 / a.spp, 2)}
    # issue-slot roofline: warp-instructions per 512 x 512 frame of this scene (ncu smsp__inst_executed.sum summed over the k_wf_*
    # launches of a 16-frame render on the B200, profiles/r02_rows_instr.json) over the SM issue rate 148 x 4 x 1.965 GHz
    try:
        rows = json.loads((ROOT / "profiles" / "r02_rows_instr.json").read_text())
        if a.size == 512 and rows.get("k_wf_per_frame_512"):
            wi, peak = rows["k_wf_per_frame_512"] * a.spp, 148 * 4 * 1.965e9
            out["roofline"] = {"bound": "issue", "achieved": wi / (st.kernel_ms * 1e-3), "peak": peak, "unit": "warp-instr/s",
                               "frac": wi / (st.kernel_ms * 1e-3) / peak, "traffic": None,
                               "model": f"{rows['k_wf_per_frame_512']:.4g} warp-instructions per frame (ncu, {rows.get('k_wf_lanes', 0):.1f} of 32 lanes active) x spp / kernel time / (148 x 4 x 1.965e9)"}
    except Exception:
        pass
    if a.oracle_spp > 0:
        from oracle import oracle
        t0 = time.perf_counter()
        r = oracle.wavefront_render(scene, a.size, a.size, a.oracle_spp)
        dt = time.perf_counter() - t0
        out["cpu_oracle"] = {"mrays_per_s": round(r["rays"] / dt / 1e6, 2), "cores": oracle.get_threads() if hasattr(oracle, "get_threads") else None,
                             "sample": f"{a.oracle_spp} of {a.spp} frames"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
