"""Renders the adjudication scene with the CPU oracle at the reference gate's size (512 x 512 x 4096 spp,
tests/test_adjudication_gate.py:41-43) and scores it against the reference's own golden pt_reference.png with the
reference's drift gate (SSIM >= 0.995, mean |diff| <= 2.0; :48-49,136-153).  Writes tests/golden/wavefront_pin.json and the
oracle's own 512 x 512 render (tests/golden/wavefront_oracle_512.png) so the GPU box can be checked against it without
/root/reference.  Needs /root/reference; run from the repo root:  python tools/wavefront_golden_pin.py [spp]"""
import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "oracle"), str(ROOT / "tests")]
import oracle  # noqa: E402
from _png import read_png, write_png  # noqa: E402
from _ssim_gate import ssim  # noqa: E402
from forge3d_b200 import wavefront as wf  # noqa: E402

def adjudication_metrics(pt_rgba):
    """The reference's own adjudication gate (tests/test_adjudication_gate.py:156-200: dE2000 < 2.0 on >= 95 % of lit pixels, SSIM > 0.96
    on the shadow-boundary band) with the reference's own helpers (tests/_deltae.py, tests/_ssim.py, imported from the checkout -- this
    tool only runs where /root/reference exists) and the reference's committed raster golden standing in for the raster render."""
    sys.path.insert(0, "/root/reference/tests")
    from _deltae import band_bbox, delta_e_2000, lit_mask, shadow_boundary_band, srgb_to_lab
    from _ssim import ssim as ref_ssim

    raster = read_png("/root/reference/tests/golden/adjudication/raster_reference.png")
    lit = lit_mask(pt_rgba)
    de = delta_e_2000(srgb_to_lab(pt_rgba), srgb_to_lab(raster))
    ys, xs = band_bbox(shadow_boundary_band(pt_rgba))
    return {"lit_pixels": int(lit.sum()), "delta_e2000_below_2_fraction_of_lit": float((de[lit] < 2.0).mean()),
            "mean_delta_e2000_lit": float(de[lit].mean()),
            "shadow_band_ssim": float(ref_ssim(pt_rgba[ys, xs, :3], raster[ys, xs, :3], data_range=255.0)),
            "gate": {"lit_fraction_min": 0.95, "band_ssim_min": 0.96}}


golden = read_png("/root/reference/tests/golden/adjudication/pt_reference.png")
if len(sys.argv) > 1 and sys.argv[1] == "--metrics-only":   # re-score the committed oracle render without rendering again
    pin_path = ROOT / "tests/golden/wavefront_pin.json"
    pin = json.loads(pin_path.read_text())
    pin["adjudication_vs_reference_raster_golden"] = {"oracle_pt": adjudication_metrics(read_png(ROOT / "tests/golden/wavefront_oracle_512.png")),
                                                      "reference_golden_pt": adjudication_metrics(golden)}
    pin_path.write_text(json.dumps(pin, indent=1) + "\n")
    print(json.dumps(pin["adjudication_vs_reference_raster_golden"], indent=1))
    sys.exit(0)
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
scene = wf.scene_from_desc(wf.adjudication_scene())
t0 = time.time()
r = oracle.wavefront_render(scene, 512, 512, spp)
secs = time.time() - t0
a, e = r["rgba8"][..., :3], golden[..., :3]
diff = a.astype(np.float32) - e.astype(np.float32)
out = {
    "width": 512, "height": 512, "spp": spp, "oracle_seconds": round(secs, 1),
    "ssim_vs_reference_golden": float(ssim(a, e, data_range=255.0)),
    "mean_abs_vs_reference_golden": float(np.abs(diff).mean()),
    "mean_signed_vs_reference_golden": float(diff.mean()),
    "max_abs_vs_reference_golden": float(np.abs(diff).max()),
    "gate": {"ssim_min": 0.995, "mean_abs_max": 2.0},
    "rays": r["rays"], "max_rays_per_frame": r["max_rays_per_frame"], "min_iterations": r["min_iterations"],
    "adjudication_vs_reference_raster_golden": {"oracle_pt": adjudication_metrics(r["rgba8"]), "reference_golden_pt": adjudication_metrics(golden)},
}
small = oracle.wavefront_render(scene, 64, 48, 6)   # arithmetic regression pin used by tests/test_wavefront.py
out["oracle_64x48x6_hdr_sha256"] = hashlib.sha256(np.ascontiguousarray(small["hdr"]).view(np.uint8).tobytes()).hexdigest()
print(json.dumps(out, indent=1))
if spp == 4096:
    old = ROOT / "tests/golden/wavefront_pin.json"
    if old.exists():   # keep the record of the full-size rehearsal of the CUDA build only while the oracle's image is unchanged
        prev = json.loads(old.read_text())
        if prev.get("ssim_vs_reference_golden") == out["ssim_vs_reference_golden"] and "cuda_source_on_simt_interpreter_byte_equal_at_gate_size" in prev:
            out["cuda_source_on_simt_interpreter_byte_equal_at_gate_size"] = prev["cuda_source_on_simt_interpreter_byte_equal_at_gate_size"]
    old.write_text(json.dumps(out, indent=1) + "\n")
    write_png(ROOT / "tests/golden/wavefront_oracle_512.png", r["rgba8"])
