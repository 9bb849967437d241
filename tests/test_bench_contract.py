"""bench.py contract checks that run without a GPU: the reference arm (CPU oracle) prints exactly one JSON line
with the required keys, non-zero ranks of a torchrun launch stay silent, and the roofline model constants match
SURVEY section 8d."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = _run(["--impl", "reference", "--gpus", "1", "--steps", "3", "--warmup", "1"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d, key
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_nonzero_ranks():
    res = _run(["--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "0"], {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_roofline_model_matches_survey_worked_example():
    sys.path.insert(0, str(ROOT))
    import bench

    b = bench.algorithmic_bytes_per_frame(1920, 1080, 2048)
    assert abs(b - (648 * 1920 * 1080 + (4 + 8 * 4 / 3) * 2048 * 2048)) < 1.0
    assert 1.40e9 < b < 1.41e9                      # B_frame = 1.405 GB
    peak, src = bench.measured_hbm_peak()
    assert peak > 1000 and ("measured" in src or "fallback" in src)
    assert (bench.WIDTH, bench.HEIGHT, bench.DEM_N) == (1920, 1080, 2048)


def test_widened_rows_never_raise_and_report_errors_in_place():
    """The side benches run in their own processes; without a GPU each must come back as an 'error' entry, not an exception."""
    sys.path.insert(0, str(ROOT))
    import bench

    rows = bench.widened_rows()
    assert set(rows) == {"wavefront", "smoke", "viewshed"}
    for name, row in rows.items():
        assert "wall_s" in row and ("error" in row or "metric" in row), (name, row)
