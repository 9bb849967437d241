#!/bin/bash
# developer helper (under gpurun, from the repo root): bash tools/gpu_rows.sh <tag>
# Widened rows on the device: their -m gpu parity tests, the smoke side bench, and one ncu launch list per row
# (per-kernel gpu__time_duration of one representative call each).  Outputs land in gpurun_out/.
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests/test_aether.py tests/test_smoke.py tests/test_viewshed.py tests/test_lbvh.py tests/test_wavefront.py -m gpu -q 2>&1 | tail -4
python tools/bench_smoke.py > gpurun_out/bench_smoke_$TAG.json 2> gpurun_out/bench_smoke_$TAG.err; tail -c 600 gpurun_out/bench_smoke_$TAG.json
python tools/bench_wavefront.py --spp 1024 > gpurun_out/bench_wavefront_$TAG.json 2> gpurun_out/bench_wavefront_$TAG.err; tail -c 600 gpurun_out/bench_wavefront_$TAG.json
F3D_B200_WF_BATCH=1 python tools/bench_wavefront.py --spp 256 --oracle-spp 0 > gpurun_out/bench_wavefront_batch1_$TAG.json 2>&1   # A/B: one frame per batch
for wd in 2 3 5 8; do F3D_B200_WF_WIDE_DEPTH=$wd python tools/bench_wavefront.py --spp 256 --oracle-spp 0 > gpurun_out/bench_wavefront_wide${wd}_$TAG.json 2>&1; done
cat > /tmp/f3d_rows_once.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "tests")
import _helpers as H
from forge3d_b200 import hybrid_render_terrain_reference, _native
from forge3d_b200 import viewshed as V
n = 1024
dem, spacing = H.rainier_dem(n), 10.0 * 2048 / n
cam = H.rainier_camera(n, spacing, dem)
kw = dict(spacing=(spacing, spacing), exaggeration=1.0, albedo=H.ALBEDO, sun_azimuth_deg=302.0, sun_elevation_deg=24.0, max_frames=2, min_frames=2, variance_threshold=1e30)
hybrid_render_terrain_reference(dem, 1920, 1080, cam, atmosphere={"turbidity": 3.0}, **kw)                 # k_aether at 1080p
h, pos, opts = V.viewshed_inputs(dem, (46.0, 7.5), bounds=(7.0, 45.6, 8.0, 46.4), height_system="ellipsoidal", observer_height=30.0)
print("viewshed kernel_ms", V.compute_viewshed(h, pos, opts)["kernel_ms"])                                   # k_viewshed on 1024^2 cells
sh, sinp, sopts = V.shadow_mask_inputs(dem, 245.0, 14.0, bounds=(7.0, 45.6, 8.0, 46.4), height_system="ellipsoidal")
print("lit fraction", V.compute_shadow_mask(sh, sinp, sopts).mean())                                          # k_shadow_mask
rng = np.random.default_rng(0); m = 1 << 20
c = rng.uniform(-100, 100, (m, 1, 3)).astype(np.float32) + rng.uniform(-0.5, 0.5, (m, 3, 3)).astype(np.float32)
_native.lbvh_build(c.reshape(-1, 3), np.arange(3 * m, dtype=np.uint32).reshape(m, 3))                         # LBVH of 1 M triangles
from forge3d_b200 import wavefront as wf
wf.render_pt_reference(wf.adjudication_scene(), 512, 512, 16)                                                  # one batch of 8 + 8 frames
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_aether|k_viewshed|k_shadow_mask|k_lbvh|k_smoke|k_wf_" --csv \
    --log-file gpurun_out/launches_rows_$TAG.csv python /tmp/f3d_rows_once.py > gpurun_out/rows_$TAG.log 2>&1
tail -3 gpurun_out/rows_$TAG.log
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/launches_rows_$TAG.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
c = rows[h]; ki = c.index("Kernel Name"); vi = c.index("Metric Value"); ui = c.index("Metric Unit")
agg = {}
for r in rows[h + 1:]:
    if len(r) > vi: agg.setdefault(r[ki][:32], []).append((float(r[vi].replace(",", "")), r[ui]))
for k, v in agg.items():
    print("NCU", k, "launches", len(v), "total", round(sum(x for x, _ in v), 1), v[0][1])
PY
