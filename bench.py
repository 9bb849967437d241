#!/usr/bin/env python
"""bench.py -- headline benchmark of the path-traced DEM snapshot path (BASELINE.json metric).

Workload (BASELINE.json configs[1]): the Rainier-shaped closed-form 2048x2048
DEM, 1920x1080 image, spp=1 per accumulation frame; 256 frames = the "256 spp" snapshot.
A STEP is one accumulation frame = one pass of the hot path (k_primary -> k_trace -> k_accum: spatial reuse,
primary / sun-shadow / IBL rays, temporal reuse, accumulate, Welford) over the whole image.
`--steps 256` (the default) therefore times exactly the 1920x1080x256spp render.

  value      Mrays/s with the scene resident in HBM (CUDA events around the K timed frames)
  e2e        the same metric through the reference-facing call `hybrid_render_terrain_reference`
             with HOST numpy buffers: DEM upload, pyramid build, K frames, resolve, D2H of
             RGBA + AOVs all inside the timed region
  roofline   SURVEY section 8d algorithmic bytes per frame / measured frame time vs the measured HBM peak
  cpu_baseline  the CPU oracle (port of the reference algorithm) on a bounded sample, host cores

`--impl reference` times the reference algorithm's CPU implementation (the oracle port: the Rust+wgpu
reference cannot be built in this image) on the same workload at a bounded sample size.

N > 1 (torchrun): the image is dealt to ranks in interleaved 16-row blocks; halo rows and the per-frame
cross-GPU barrier travel over NVLink peer stores inside k_primary; one NCCL all-gather assembles the frame.
scaling = "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

WIDTH, HEIGHT, DEM_N, DEM_SPACING = 1920, 1080, 2048, 10.0
CPU_SAMPLE = (480, 270, 32)   # width, height, frames of the bounded CPU sample


def workload():
    import _helpers as H

    dem = H.rainier_dem(DEM_N)
    cam = H.rainier_camera(DEM_N, DEM_SPACING, dem)
    kw = dict(spacing=(DEM_SPACING, DEM_SPACING), exaggeration=1.0, albedo=H.ALBEDO, sun_azimuth_deg=302.0,
              sun_elevation_deg=24.0, sun_intensity=2.5, env_intensity=0.35, seed=7, spp=1)
    return dem, cam, kw


def algorithmic_bytes_per_frame(width, height, dem_n):
    """B_frame = 648*W*H + (4 + 8*4/3) * Wd_pad*Hd_pad."""
    pad = 1
    while pad < dem_n - 1:
        pad *= 2
    return 648.0 * width * height + (4.0 + 8.0 * 4.0 / 3.0) * pad * pad


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc, self.stamps = index, [], None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])
            self.stamps.append(time.time())

    def stop(self, t_begin=0.0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.time()
        time.sleep(0.06)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for r, ts in zip(self.rows, self.stamps) if t_begin - 0.02 <= ts <= t_end + 0.05] or self.rows[-3:]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def calibrate_oracle_threads(dem, cam, kw):
    """Pick the OpenMP thread count that is actually fastest on this host (cgroup CPU quotas and SMT make
    `all logical CPUs` a bad default on some boxes): 6-frame probes at n, n/2, n/4 threads."""
    from oracle import oracle

    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    w, h, _ = CPU_SAMPLE
    best, best_t = n, None
    oracle.set_threads(n)
    oracle.render(dem, w, h, cam, **kw, max_frames=3, min_frames=3, variance_threshold=1e30)   # warm threads + page cache
    for cand in sorted({max(n, 1), max(n // 2, 1), max(n // 4, 1)}, reverse=True):
        oracle.set_threads(cand)
        out = oracle.render(dem, w, h, cam, **kw, max_frames=6, min_frames=6, variance_threshold=1e30)
        if best_t is None or out["frames_seconds"] < 0.9 * best_t:     # fewer threads only if clearly faster
            best, best_t = cand, out["frames_seconds"]
    oracle.set_threads(best)
    return best


def cpu_baseline(dem, cam, kw, threads=0):
    """The oracle (port of the reference algorithm) on a bounded sample of the same workload."""
    from oracle import oracle

    if threads:
        oracle.set_threads(threads)
    else:
        calibrate_oracle_threads(dem, cam, kw)
    w, h, frames = CPU_SAMPLE
    out = oracle.render(dem, w, h, cam, **kw, max_frames=frames, min_frames=frames, variance_threshold=1e30)
    dt = out["frames_seconds"]
    rays = out["rays_primary"] + out["rays_shadow"] + out["rays_ibl"]
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": oracle.get_threads(), "kind": "port",
            "sample": f"{w}x{h} px x {frames} frames spp=1 of the same scene ({rays} rays in {dt:.2f} s of frame loop; "
                      f"setup {out['setup_seconds']:.2f} s excluded, as for the GPU value)",
            "ms_per_frame_sample": dt * 1e3 / frames}


def widened_rows():
    """Side benches of the SURVEY section 8f rows (tools/bench_*.py), each in its own process under a 100 s timeout so that nothing they do
    can disturb the headline line; their JSON is attached under "widened_rows".  N = 1 only; F3D_BENCH_ROWS=0 skips them."""
    rows = {}
    for name, cmd in (("wavefront", ["tools/bench_wavefront.py", "--spp", "512", "--oracle-spp", "8", "--repeats", "2"]),
                      ("smoke", ["tools/bench_smoke.py", "--steps", "3"]),
                      ("viewshed", ["tools/bench_viewshed.py"])):
        t0 = time.time()
        try:
            res = subprocess.run([sys.executable, *cmd], cwd=str(ROOT), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=100)
            last = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
            rows[name] = json.loads(last[-1]) if res.returncode == 0 and last else {"error": (res.stderr or res.stdout)[-300:]}
        except Exception as exc:   # timeout, missing file, bad JSON: the headline stands without the row
            rows[name] = {"error": repr(exc)[:300]}
        rows[name]["wall_s"] = round(time.time() - t0, 1)
    return rows


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.pop("OMP_NUM_THREADS", None)          # torchrun pins it to 1; the CPU arm uses the host's cores
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
    from oracle import oracle

    dem, cam, kw = workload()
    w, h, _ = CPU_SAMPLE
    common = dict(variance_threshold=1e30)
    calibrate_oracle_threads(dem, cam, kw)
    if args.warmup:
        oracle.render(dem, w, h, cam, **kw, max_frames=max(args.warmup, 2), min_frames=max(args.warmup, 2), **common)
    k = max(args.steps, 2)
    out = oracle.render(dem, w, h, cam, **kw, max_frames=k, min_frames=k, **common)
    dt = out["frames_seconds"]   # the K-frame loop; DEM pyramid build (setup) excluded as for the GPU value
    rays = out["rays_primary"] + out["rays_shadow"] + out["rays_ibl"]
    val = rays / dt / 1e6
    sample = f"{w}x{h} px (1/16 of the 1920x1080 frame) per step, {k} steps, same scene and DEM"
    line = {"impl": "reference", "metric": "Mrays/s, path-traced DEM snapshot (primary+shadow+IBL rays)", "value": val,
            "unit": "Mrays/s", "n_gpus": args.gpus, "steps": k, "warmup": args.warmup, "ms_per_step": dt * 1e3 / k,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "rainier-shaped 2048x2048 DEM, 1920x1080, spp=1/frame (BASELINE configs[1])",
                       "reference_impl": "CPU oracle port of the WGSL/Rust path (reference needs cargo+wgpu: unbuildable here)"},
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": oracle.get_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the spp=8 x 32 frames run")
    ap.add_argument("--no-rows", action="store_true", help="skip the side benches of the widened rows")
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    os.environ["NCCL_DEBUG"] = os.environ.get("F3D_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
    import torch
    import torch.distributed as dist

    from forge3d_b200 import _native
    from forge3d_b200.distributed import PartitionedRender

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (forge3d_b200 has no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if distributed:
            dist.barrier()

    def all_reduce(t, op=None):
        if distributed:
            dist.all_reduce(t, op=op if op is not None else dist.ReduceOp.SUM)

    W, Hh = args.width, args.height
    dem, cam, kw = workload()
    K, Wm = max(args.steps, 1), max(args.warmup, 3)
    fixed = dict(max_frames=K + Wm, min_frames=K + Wm, variance_threshold=1e30)

    # ---------------- resident-scene timing (value) ----------------
    pr = PartitionedRender(dem, W, Hh, cam, **kw, **fixed)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)            # nvidia-smi start-up; samples before the timed region are dropped below
    pr.render_frames(Wm)
    torch.cuda.synchronize()
    s0 = pr.session.stats()
    barrier()
    t_region0 = time.time()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    pr.render_frames(K)
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop(t_region0) if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    all_reduce(ms, dist.ReduceOp.MAX)
    s1 = pr.session.stats()
    rays_local = [s1[k] - s0[k] for k in ("rays_primary", "rays_shadow", "rays_ibl")]
    rays = torch.tensor(rays_local + [s1["nodes_popped"] - s0["nodes_popped"]], dtype=torch.float64, device="cuda")
    all_reduce(rays)
    total_ms = float(ms.item())
    n_primary, n_shadow, n_ibl, n_nodes = [float(v) for v in rays.tolist()]
    total_rays = n_primary + n_shadow + n_ibl
    value = total_rays / (total_ms * 1e-3) / 1e6
    images = pr.resolve(aovs=False)
    pr.close()

    # ---------------- end to end through the public call with host buffers (e2e) ----------------
    e2e = None
    if not args.no_e2e:
        if world == 1:
            pinned = torch.from_numpy(dem).pin_memory()
            dem_host = pinned.numpy()
            kw_e2e = dict(kw)
            kw_e2e.update(max_frames=K, min_frames=K, variance_threshold=1e30)
            _native.hybrid_render_terrain_reference(dem_host, W, Hh, cam, **{**kw_e2e, "max_frames": 3, "min_frames": 3})
            t0 = time.perf_counter()
            out = _native.hybrid_render_terrain_reference(dem_host, W, Hh, cam, **kw_e2e)
            dt = time.perf_counter() - t0
            r = out["rays_primary"] + out["rays_shadow"] + out["rays_ibl"]
            d2h = out["rgba"].nbytes + out["albedo"].nbytes + out["normal"].nbytes + out["depth"].nbytes
            e2e = {"value": r / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": dem_host.nbytes / K,
                   "d2h_bytes_per_step": d2h / K, "call_ms": dt * 1e3, "frames": K,
                   "setup_ms": out["setup_ms"], "frames_ms": out["frames_ms"], "readback_ms": out["readback_ms"],
                   "note": "one hybrid_render_terrain_reference call = K steps; bytes are per call / K"}
        else:
            # partitioned call: per-rank DEM upload + pyramid build + K frames + resolve + NCCL gather + D2H
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pr2 = PartitionedRender(dem, W, Hh, cam, **kw, max_frames=K, min_frames=K, variance_threshold=1e30)
            pr2.render_frames(K)
            imgs = pr2.resolve(aovs=True)
            torch.cuda.synchronize()
            dist.barrier()
            dt = time.perf_counter() - t0
            st = pr2.session.stats()
            r = torch.tensor([st["rays_primary"] + st["rays_shadow"] + st["rays_ibl"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(r)
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            d2h = sum(v.nbytes for v in imgs.values())
            e2e = {"value": float(r.item()) / float(tt.item()) / 1e6, "unit": "Mrays/s",
                   "h2d_bytes_per_step": dem.nbytes * world / K, "d2h_bytes_per_step": d2h / K,
                   "call_ms": float(tt.item()) * 1e3, "frames": K,
                   "note": "partitioned render incl. per-rank DEM upload, NCCL row gather and D2H on every rank"}
            pr2.close()

    # ---------------- SURVEY 8d secondary run: spp = 8 x 32 frames (same 256 samples per pixel, 1/8 of the per-frame state traffic) ----
    secondary = None
    if world == 1 and not args.no_secondary:
        try:
            kw8 = dict(kw)
            kw8["spp"] = 8
            pr8 = PartitionedRender(dem, W, Hh, cam, **kw8, max_frames=35, min_frames=35, variance_threshold=1e30)
            pr8.render_frames(3)
            torch.cuda.synchronize()
            q0 = pr8.session.stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pr8.render_frames(32)
            e1.record()
            torch.cuda.synchronize()
            q1 = pr8.session.stats()
            pr8.close()
            ms8 = float(e0.elapsed_time(e1))
            rays8 = float(sum(q1[k] - q0[k] for k in ("rays_primary", "rays_shadow", "rays_ibl")))
            b8 = algorithmic_bytes_per_frame(W, Hh, DEM_N)
            peak8, _ = measured_hbm_peak()
            secondary = {"workload": "same scene, spp=8 per frame x 32 frames", "value": rays8 / (ms8 * 1e-3) / 1e6, "unit": "Mrays/s",
                         "ms_per_frame": ms8 / 32, "bytes_per_ray": b8 / (rays8 / 32),
                         "roofline_frac": b8 / (ms8 / 32 * 1e-3) / 1e9 / peak8,
                         "note": "issue-bound regime: 8 samples share one frame's state traffic, so the HBM fraction is small by construction"}
        except Exception as exc:   # the headline stands without it
            secondary = {"error": repr(exc)[:300]}

    if rank == 0:
        ms_per_step = total_ms / K
        b_frame = algorithmic_bytes_per_frame(W, Hh, DEM_N)
        peak, peak_src = measured_hbm_peak()
        achieved = b_frame / (ms_per_step * 1e-3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("frame_dram_bytes")
            except Exception:
                traffic = None
        line = {
            "metric": "Mrays/s, path-traced DEM snapshot (primary+shadow+IBL rays)", "value": value, "unit": "Mrays/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"rainier-shaped {DEM_N}x{DEM_N} DEM (SURVEY 8d C2), {W}x{Hh}, spp=1 per frame, "
                                   f"{K} frames timed (256 = the 256-spp snapshot)",
                       "step": "one accumulation frame = k_primary + k_trace (sun list, then IBL list) + k_accum over the image",
                       "l2": "per-frame working set (state 118 MB + DEM cells/pyramid 108 MB) exceeds the 126 MB L2; no flush",
                       "partition": f"interleaved 16-row blocks over {world} GPU(s)", "ms_per_frame": ms_per_step,
                       "rays_per_frame": total_rays / K, "f_shadow": n_shadow / max(n_primary, 1),
                       "f_ibl": n_ibl / max(n_primary, 1), "nodes_per_ray": n_nodes / max(total_rays, 1)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": b_frame,
                         "bytes_per_ray": b_frame / (total_rays / K),
                         "kernel": "frame = k_primary + k_trace + k_accum (dominant: k_trace)"},
            "clocks": clocks, "gpu_launches": 3 * K * kw["spp"], "e2e": e2e, "secondary": secondary,
            "image_mean_rgb": float(images["rgba"][..., :3].mean()),
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(dem, cam, kw)
            if os.environ.get("F3D_BENCH_ROWS", "1") != "0" and not args.no_rows:
                line["widened_rows"] = widened_rows()
        elif not args.no_cpu_baseline:
            line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "port",
                                    "sample": "reported at N=1 only"}
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
