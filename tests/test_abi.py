"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a
GPU, exports every symbol include/forge3d_b200.h declares, the ctypes mirrors have the C layout,
and the public Python facade keeps the reference's signature and validation contracts
(tests/test_hybrid_terrain_pt.py:411-458,461-633 of the reference).  No compute calls."""
import ctypes as C
import inspect
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import _helpers as H
from forge3d_b200 import _native, build as fbuild, hybrid_render_terrain_reference

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "forge3d_b200.h"


def test_library_builds_loads_and_exports_every_declared_symbol():
    fbuild.build()
    L = _native.lib()
    text = HEADER.read_text()
    declared = sorted(set(re.findall(r"\b(f3d_[a-z0-9_]+)\s*\(", text)))
    assert declared, "no declarations parsed from the header"
    assert sorted(_native.EXPORTS) == declared
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert L.f3d_abi_version() == 2
    assert int(re.search(r"#define F3D_ABI_VERSION (\d+)", text).group(1)) == 2


def test_struct_layout_matches_c():
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "forge3d_b200.h"
    int main(void) {
      printf("%zu %zu %zu ", sizeof(f3d_viewshed_options), offsetof(f3d_viewshed_options, earth_latitude_deg), offsetof(f3d_viewshed_options, device));
      printf("%zu %zu %zu %zu %zu %zu ", sizeof(f3d_wavefront_scene), offsetof(f3d_wavefront_scene, spheres), offsetof(f3d_wavefront_scene, environment),
             offsetof(f3d_wavefront_scene, ninstances), sizeof(f3d_wavefront_stats), offsetof(f3d_wavefront_stats, kernel_ms));
      printf("%zu %zu %zu %zu ", sizeof(f3d_smoke_volume), offsetof(f3d_smoke_volume, frame_index), sizeof(f3d_smoke_settings),
             offsetof(f3d_smoke_settings, soot_absorption));
      printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(f3d_terrain_desc), offsetof(f3d_terrain_desc, observer_lat_deg),
             offsetof(f3d_terrain_desc, env_rgb), offsetof(f3d_terrain_desc, width), offsetof(f3d_terrain_desc, part_block_rows),
             sizeof(f3d_terrain_out), offsetof(f3d_terrain_out, kernel_launches), offsetof(f3d_terrain_desc, atmosphere),
             sizeof(f3d_atmosphere), offsetof(f3d_atmosphere, transmittance_mu), offsetof(f3d_atmosphere, ground_albedo));
      return 0; }'''
    tmp = Path("/tmp/f3d_layout.c")
    tmp.write_text(src)
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(tmp), "-o", "/tmp/f3d_layout"], check=True)
    vals = list(map(int, subprocess.run(["/tmp/f3d_layout"], check=True, capture_output=True, text=True).stdout.split()))
    D, O, A = _native.TerrainDesc, _native.TerrainOut, _native.Atmosphere
    VO = _native.ViewshedOptions
    assert vals[:3] == [C.sizeof(VO), VO.earth_latitude_deg.offset, VO.device.offset]
    vals = vals[3:]
    WS, WT = _native.WavefrontSceneC, _native.WavefrontStats
    assert vals[:6] == [C.sizeof(WS), WS.spheres.offset, WS.environment.offset, WS.ninstances.offset, C.sizeof(WT), WT.kernel_ms.offset]
    vals = vals[6:]
    SV, SS = _native.SmokeVolume, _native.SmokeSettings
    assert vals[:4] == [C.sizeof(SV), SV.frame_index.offset, C.sizeof(SS), SS.soot_absorption.offset]
    vals = vals[4:]
    assert vals == [C.sizeof(D), D.observer_lat_deg.offset, D.env_rgb.offset, D.width.offset, D.part_block_rows.offset,
                    C.sizeof(O), O.kernel_launches.offset, D.atmosphere.offset, C.sizeof(A), A.transmittance_mu.offset,
                    A.ground_albedo.offset]


def _build_c_consumer():
    exe = Path("/tmp/f3d_abi_smoke")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I", str(ROOT / "include"), str(ROOT / "tests" / "c" / "abi_smoke.c"),
                    "-o", str(exe), "-L", str(ROOT / "forge3d_b200"), "-lforge3d_b200",
                    f"-Wl,-rpath,{ROOT / 'forge3d_b200'}"], check=True)
    return exe


def test_plain_c_consumer_links_and_fails_loudly_without_gpu():
    """The boundary is a real C ABI: a C99 program links libforge3d_b200.so with nothing but the header."""
    fbuild.build()
    exe = _build_c_consumer()
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    if _native.lib().f3d_device_count() > 0:
        assert res.returncode == 0 and "frames=4" in res.stdout, res.stdout + res.stderr
    else:
        assert res.returncode == 3, (res.returncode, res.stdout, res.stderr)       # F3D_ERR_DEVICE
        assert "no CUDA device" in res.stdout and "no CPU fallback" in res.stdout


def test_no_gpu_fails_loudly():
    L = _native.lib()
    if L.f3d_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hybrid_render_terrain_reference(np.zeros((4, 4), np.float32), 8, 8, H.CAM, max_frames=4, min_frames=2)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        _native.trace_rays(np.zeros((4, 4), np.float32), (1, 1), (0, 0), 1.0, np.zeros((1, 8), np.float32),
                           any_hit=True, apply_curvature=False)


def test_native_validation_runs_before_any_device_work():
    # the trust-boundary checks of validate_desc (render_terrain.rs:474-557) need no GPU, so they are testable here;
    # the 2^31-pixel limit is this backend's own (32-bit pixel indices in the kernels)
    dem = np.zeros((4, 4), np.float32)
    with pytest.raises(ValueError, match="2\\^31-pixel addressing limit"):
        _native.hybrid_render_terrain_reference(dem, 65536, 65536, {}, max_frames=2, min_frames=2)
    with pytest.raises(RuntimeError, match="spp must be in 1..=64"):
        _native.hybrid_render_terrain_reference(dem, 8, 8, {}, spp=65, max_frames=2, min_frames=2)
    with pytest.raises(RuntimeError, match="min_frames .* must be <= max_frames"):
        _native.hybrid_render_terrain_reference(dem, 8, 8, {}, max_frames=2, min_frames=3)
    with pytest.raises(RuntimeError, match="camera look_at must differ from origin"):
        _native.hybrid_render_terrain_reference(dem, 8, 8, {"origin": (0, 1, 0), "look_at": (0, 1, 0)}, max_frames=2, min_frames=2)


def test_product_never_touches_the_oracle():
    for path in (ROOT / "forge3d_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".cpp", ".h"):
            assert "oracle" not in path.read_text().replace("CPU oracle", "").replace("the oracle", "").replace(
                "oracle/f3d_oracle.c", ""), f"{path} references the oracle"


def test_signature_matches_reference_facade():
    # python/forge3d/path_tracing.py:893-926 (order, kinds, defaults)
    sig = inspect.signature(hybrid_render_terrain_reference)
    expected = [
        ("heightmap", inspect.Parameter.POSITIONAL_OR_KEYWORD, inspect._empty), ("width", 1, inspect._empty),
        ("height", 1, inspect._empty), ("camera", 1, None),
        ("spacing", 3, (1.0, 1.0)), ("exaggeration", 3, 1.0), ("albedo", 3, (0.6, 0.6, 0.6)),
        ("sun_azimuth_deg", 3, None), ("sun_elevation_deg", 3, None), ("solar_time", 3, None),
        ("sun_intensity", 3, 2.5), ("sun_color", 3, (1.0, 0.97, 0.92)), ("env_map", 3, None),
        ("env_intensity", 3, 0.35), ("mesh_vertices", 3, None), ("mesh_indices", 3, None), ("spp", 3, 1),
        ("max_frames", 3, 512), ("min_frames", 3, 32), ("variance_threshold", 3, 1e-3), ("seed", 3, 7),
        ("certificate", 3, False), ("cache", 3, None), ("observer_latitude_deg", 3, None),
        ("observer_longitude_deg", 3, None), ("earth_model", 3, "ellipsoid"), ("sphere_radius_m", 3, 6_371_008.8),
        ("refraction_model", 3, "bennett"), ("refraction_k", 3, 0.13), ("pressure_mbar", 3, None),
        ("temperature_c", 3, None), ("atmosphere", 3, None),
    ]
    got = [(p.name, int(p.kind), p.default) for p in sig.parameters.values()]
    assert got == [(n, int(k), d) for n, k, d in expected]
    # native seam positional order (terrain_reference.rs:224-256)
    nat = list(inspect.signature(_native.hybrid_render_terrain_reference).parameters)
    assert nat[:31] == ["heightmap", "width", "height", "cam", "spacing", "exaggeration", "albedo", "sun_azimuth_deg",
                        "sun_elevation_deg", "sun_intensity", "env_map", "env_intensity", "mesh_vertices",
                        "mesh_indices", "spp", "max_frames", "min_frames", "variance_threshold", "seed", "certificate",
                        "sun_color", "cache", "observer_latitude_deg", "observer_longitude_deg", "earth_model",
                        "sphere_radius_m", "refraction_model", "refraction_k", "pressure_mbar", "temperature_c",
                        "atmosphere"]


def test_python_side_validation_contracts():
    dem = H.golden_dem()
    kw = H.scene_kwargs(dem)
    f = hybrid_render_terrain_reference
    with pytest.raises(ValueError, match="non-finite"):
        f(np.full((16, 16), np.nan, np.float32), 64, 64, H.CAM, max_frames=8)
    with pytest.raises(ValueError, match="at least 2x2"):
        f(np.zeros((1, 1), np.float32), 64, 64, H.CAM, max_frames=8)
    with pytest.raises(ValueError, match="min_frames"):
        f(dem, 64, 64, H.CAM, **{**kw, "max_frames": 4, "min_frames": 8})
    with pytest.raises(ValueError, match="spacing"):
        f(dem, 64, 64, H.CAM, **{**kw, "spacing": (0.0, 1.0)})
    with pytest.raises(ValueError, match="spp"):
        f(dem, 64, 64, H.CAM, **{**kw, "spp": 0})
    with pytest.raises(ValueError, match="together"):
        f(dem, 64, 64, H.CAM, **{**kw, "mesh_vertices": np.zeros((3, 3), np.float32)})
    with pytest.raises(ValueError, match="2D"):
        f(np.zeros((2, 2, 2), np.float32), 8, 8)
    for bad in [(1.0, float("nan"), 1.0), (1.0, float("inf"), 1.0), (1.0, -0.1, 1.0), (1.0, 1.0), (1.0, 1.0, 1.0, 1.0),
                0.5, "abc", ("0.5", "0.9", "0.8"), bytearray([1, 1, 1]), memoryview(bytes([1, 1, 1]))]:
        with pytest.raises(ValueError):
            f(np.zeros((4, 4), np.float32), 8, 8, H.CAM, sun_color=bad, max_frames=4, min_frames=2)
    for bad in [(1.0, -1.0, 1.0), (1.0, float("nan"), 1.0), (1.0, 1.0), 0.5, "abc", ("0.5", "0.9", "0.8"),
                np.float64(0.5), np.array([1.0, 1.0]), bytearray([1, 1, 1])]:
        with pytest.raises(ValueError):
            _native.hybrid_render_terrain_reference(np.zeros((4, 4), np.float32), 8, 8, H.CAM, sun_color=bad,
                                                    max_frames=4, min_frames=2, variance_threshold=1e30)
    with pytest.raises(ValueError, match="unsupported earth_model"):
        f(dem, 8, 8, H.CAM, earth_model="mean-earth")
    with pytest.raises(ValueError, match="unsupported refraction_model"):
        f(dem, 8, 8, H.CAM, refraction_model="standard")


def test_solar_time_resolution(monkeypatch):
    # tests/test_hybrid_terrain_pt.py:860-926 with a duck-typed SolarTime
    import forge3d_b200.path_tracing as pt

    captured = {}

    class Native:
        @staticmethod
        def hybrid_render_terrain_reference(*args, **kwargs):
            captured.update(kwargs)
            return {}

    class When:
        observer_lat, observer_lon, pressure_mbar, temperature_c = 39.742476, -105.1786, 820.0, 11.0

        def position(self):
            return {"azimuth_deg": 194.34, "true_elevation_deg": 39.88, "apparent_elevation_deg": 39.89}

    monkeypatch.setattr(pt, "_NATIVE", Native())
    res = pt.hybrid_render_terrain_reference(np.zeros((2, 2), np.float32), 2, 2, solar_time=When(), min_frames=1, max_frames=1)
    assert captured["sun_azimuth_deg"] == pytest.approx(194.34) and captured["sun_elevation_deg"] == pytest.approx(39.89)
    assert captured["pressure_mbar"] == 820.0 and captured["observer_latitude_deg"] == pytest.approx(39.742476)
    assert "sun_source" not in captured and res["sun_source"] == "solar_time"
    pt.hybrid_render_terrain_reference(np.zeros((2, 2), np.float32), 2, 2, solar_time=When(), refraction_model="none",
                                       min_frames=1, max_frames=1)
    assert captured["sun_elevation_deg"] == pytest.approx(39.88)
    with pytest.raises(ValueError, match="cannot be combined"):
        pt.hybrid_render_terrain_reference(np.zeros((2, 2), np.float32), 2, 2, solar_time=When(), sun_azimuth_deg=123.0)
