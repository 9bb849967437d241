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

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
golden = read_png("/root/reference/tests/golden/adjudication/pt_reference.png")
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
