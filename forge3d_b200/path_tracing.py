"""Public entry point of the B200 backend: `hybrid_render_terrain_reference`.

Mirror of the reference's Python facade `forge3d.path_tracing.hybrid_render_terrain_reference`
(/root/reference/python/forge3d/path_tracing.py:893-1095): same parameter order, defaults, validation
order, exception types, message substrings and result keys, so the reference's own tests
(tests/test_hybrid_terrain_pt.py) read the same against this module.  The native call goes to
forge3d_b200._native (ctypes -> libforge3d_b200.so -> CUDA); there is no CPU fallback.
"""
from __future__ import annotations

from typing import Any, Mapping, Sequence

import numpy as np

from . import _native as _NATIVE


def _resolve_solar_time(solar_time, refraction_model):
    """path_tracing.py:1018-1032.  The reference coerces through forge3d.geo.SolarTime (ephemeris code,
    out of scope); any object exposing the same duck type is accepted here:
    .position() -> {'azimuth_deg', 'true_elevation_deg', 'apparent_elevation_deg'} and the attributes
    observer_lat, observer_lon, pressure_mbar, temperature_c."""
    if not hasattr(solar_time, "position"):
        raise TypeError("solar_time must provide position() and observer_lat/observer_lon/pressure_mbar/temperature_c")
    solar = solar_time.position()
    az = solar["azimuth_deg"]
    el = solar["true_elevation_deg" if refraction_model == "none" else "apparent_elevation_deg"]
    return az, el, solar_time.observer_lat, solar_time.observer_lon, solar_time.pressure_mbar, solar_time.temperature_c


def hybrid_render_terrain_reference(
    heightmap: "np.ndarray",
    width: int,
    height: int,
    camera: "dict | None" = None,
    *,
    spacing: "tuple[float, float]" = (1.0, 1.0),
    exaggeration: float = 1.0,
    albedo: "tuple[float, float, float]" = (0.6, 0.6, 0.6),
    sun_azimuth_deg: float | None = None,
    sun_elevation_deg: float | None = None,
    solar_time: "object | None" = None,
    sun_intensity: float = 2.5,
    sun_color: "Sequence[float] | np.ndarray" = (1.0, 0.97, 0.92),
    env_map: "np.ndarray | None" = None,
    env_intensity: float = 0.35,
    mesh_vertices: "np.ndarray | None" = None,
    mesh_indices: "np.ndarray | None" = None,
    spp: int = 1,
    max_frames: int = 512,
    min_frames: int = 32,
    variance_threshold: float = 1e-3,
    seed: int = 7,
    certificate: bool | str = False,
    cache: str | None = None,
    observer_latitude_deg: float | None = None,
    observer_longitude_deg: float | None = None,
    earth_model: str = "ellipsoid",
    sphere_radius_m: float = 6_371_008.8,
    refraction_model: str = "bennett",
    refraction_k: float = 0.13,
    pressure_mbar: float | None = None,
    temperature_c: float | None = None,
    atmosphere: "Mapping[str, Any] | Any | None" = None,
) -> dict:
    """Converged GPU path-traced reference of a DEM under sun + IBL on a B200.

    Returns a dict with ``rgba`` (H,W,4) uint8, ``albedo``/``normal`` (H,W,3) float32, ``depth``
    (H,W) float32 ray distance (NaN on miss), ``frames``, ``variance``, ``converged``,
    ``peak_host_visible_bytes``, ``minmax_pyramid_bytes``, ``gpu_resource_bytes``, ``sun_source``,
    ``solar_azimuth_deg``, ``solar_elevation_deg`` (reference keys, terrain_reference.rs:437-450 and
    path_tracing.py:1092-1094) plus ray counters and timings of this backend.
    """
    _ = cache
    dem = np.ascontiguousarray(heightmap, dtype=np.float32)
    if dem.ndim != 2:
        raise ValueError(f"heightmap must be 2D (H, W), got shape {dem.shape}")
    if dem.shape[0] < 2 or dem.shape[1] < 2:
        raise ValueError(
            f"terrain heightfield must be at least 2x2 texels, got {dem.shape[1]}x{dem.shape[0]}"
        )
    if not np.isfinite(dem).all():
        raise ValueError("heightmap contains non-finite samples")
    if int(min_frames) > int(max_frames):
        raise ValueError(f"min_frames ({min_frames}) must be <= max_frames ({max_frames})")
    if not (1 <= int(spp) <= 64):
        raise ValueError(f"spp must be in 1..=64, got {spp}")
    if not (float(spacing[0]) > 0.0 and float(spacing[1]) > 0.0):
        raise ValueError(f"spacing must be > 0, got {spacing}")
    if isinstance(sun_color, (str, bytes, bytearray, memoryview)):
        raise ValueError(f"sun_color must be three numbers, got {sun_color!r}")
    try:
        if len(sun_color) != 3:
            raise ValueError(f"sun_color must have exactly three components, got {len(sun_color)}")
        sun_rgb = (float(sun_color[0]), float(sun_color[1]), float(sun_color[2]))
    except (TypeError, ValueError) as exc:
        raise ValueError(f"sun_color must be a sequence of three numbers: {exc}")
    if any(isinstance(c, (str, bytes, bytearray, memoryview)) for c in sun_color):
        raise ValueError(f"sun_color components must be numbers, got {sun_color!r}")
    if not all(bool(np.isfinite(c)) for c in sun_rgb):
        raise ValueError(f"sun_color components must be finite, got {sun_color!r}")
    if any(c < 0.0 for c in sun_rgb):
        raise ValueError(f"sun_color components must be non-negative, got {sun_color!r}")
    sun_source = "manual_angles"
    if solar_time is not None:
        if any(v is not None for v in (sun_azimuth_deg, sun_elevation_deg, observer_latitude_deg,
                                       observer_longitude_deg, pressure_mbar, temperature_c)):
            raise ValueError(
                "solar_time cannot be combined with manual sun, observer, pressure, or temperature values"
            )
        (sun_azimuth_deg, sun_elevation_deg, observer_latitude_deg, observer_longitude_deg,
         pressure_mbar, temperature_c) = _resolve_solar_time(solar_time, refraction_model)
        sun_source = "solar_time"
    else:
        sun_azimuth_deg = 315.0 if sun_azimuth_deg is None else sun_azimuth_deg
        sun_elevation_deg = 45.0 if sun_elevation_deg is None else sun_elevation_deg
        observer_latitude_deg = 0.0 if observer_latitude_deg is None else observer_latitude_deg
        observer_longitude_deg = 0.0 if observer_longitude_deg is None else observer_longitude_deg
        pressure_mbar = 1013.25 if pressure_mbar is None else pressure_mbar
        temperature_c = 15.0 if temperature_c is None else temperature_c
    cam = dict(camera or {})
    env = None
    if env_map is not None:
        env = np.ascontiguousarray(env_map, dtype=np.float32)
        if env.ndim != 3 or env.shape[2] != 3:
            raise ValueError(f"env_map must be (H, W, 3) float32, got {env.shape}")
    if (mesh_vertices is None) != (mesh_indices is None):
        raise ValueError("mesh_vertices and mesh_indices must be provided together")
    mv = mi = None
    if mesh_vertices is not None:
        mv = np.ascontiguousarray(mesh_vertices, dtype=np.float32)
        mi = np.ascontiguousarray(mesh_indices, dtype=np.uint32)
        if mv.ndim != 2 or mv.shape[1] != 3:
            raise ValueError(f"mesh_vertices must be (N, 3), got {mv.shape}")
        if mi.ndim != 2 or mi.shape[1] != 3:
            raise ValueError(f"mesh_indices must be (M, 3), got {mi.shape}")
    result = _NATIVE.hybrid_render_terrain_reference(
        dem,
        int(width),
        int(height),
        cam,
        spacing=(float(spacing[0]), float(spacing[1])),
        exaggeration=float(exaggeration),
        albedo=(float(albedo[0]), float(albedo[1]), float(albedo[2])),
        sun_azimuth_deg=float(sun_azimuth_deg),
        sun_elevation_deg=float(sun_elevation_deg),
        sun_intensity=float(sun_intensity),
        sun_color=sun_rgb,
        env_map=env,
        env_intensity=float(env_intensity),
        mesh_vertices=mv,
        mesh_indices=mi,
        spp=int(spp),
        max_frames=int(max_frames),
        min_frames=int(min_frames),
        variance_threshold=float(variance_threshold),
        seed=int(seed),
        certificate=certificate,
        observer_latitude_deg=float(observer_latitude_deg),
        observer_longitude_deg=float(observer_longitude_deg),
        earth_model=earth_model,
        sphere_radius_m=float(sphere_radius_m),
        refraction_model=refraction_model,
        refraction_k=float(refraction_k),
        pressure_mbar=float(pressure_mbar),
        temperature_c=float(temperature_c),
        atmosphere=atmosphere,
    )
    result["sun_source"] = sun_source
    result["solar_azimuth_deg"] = float(sun_azimuth_deg)
    result["solar_elevation_deg"] = float(sun_elevation_deg)
    return result
