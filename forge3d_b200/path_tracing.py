"""Public entry point of the B200 backend: `hybrid_render_terrain_reference`.

Stand-in for the reference's Python facade `forge3d.path_tracing.hybrid_render_terrain_reference`
(/root/reference/python/forge3d/path_tracing.py:893-1095): same parameter order and defaults (pinned by the
reference's tests/test_hybrid_terrain_pt.py:461-590), same validation ORDER, exception types and message
substrings (:411-458, :593-633), same result keys.  The native call goes to forge3d_b200._native
(ctypes -> libforge3d_b200.so -> CUDA); there is no CPU fallback.
"""
from __future__ import annotations

from typing import Any, Mapping, Sequence

import numpy as np

from . import _native as _NATIVE

_TEXT_LIKE = (str, bytes, bytearray, memoryview)


def _checked_dem(heightmap) -> np.ndarray:
    """Trust boundary for the DEM (reference facade :970-978)."""
    dem = np.ascontiguousarray(heightmap, dtype=np.float32)
    if dem.ndim != 2:
        raise ValueError(f"heightmap must be 2D (H, W), got shape {dem.shape}")
    rows, cols = dem.shape
    if rows < 2 or cols < 2:
        raise ValueError(f"terrain heightfield must be at least 2x2 texels, got {cols}x{rows}")
    if not np.isfinite(dem).all():
        raise ValueError("heightmap contains non-finite samples")
    return dem


def _checked_sun_color(sun_color) -> tuple:
    """Three finite, non-negative numbers; anything else is a ValueError before any GPU work (:987-1000)."""
    if isinstance(sun_color, _TEXT_LIKE):
        raise ValueError(f"sun_color must be three numbers, got {sun_color!r}")
    try:
        count = len(sun_color)
        if count != 3:
            raise ValueError(f"sun_color must have exactly three components, got {count}")
        rgb = tuple(float(sun_color[i]) for i in range(3))
    except (TypeError, ValueError) as exc:
        raise ValueError(f"sun_color must be a sequence of three numbers: {exc}")
    if any(isinstance(c, _TEXT_LIKE) for c in sun_color):
        raise ValueError(f"sun_color components must be numbers, got {sun_color!r}")
    if not all(bool(np.isfinite(c)) for c in rgb):
        raise ValueError(f"sun_color components must be finite, got {sun_color!r}")
    if min(rgb) < 0.0:
        raise ValueError(f"sun_color components must be non-negative, got {sun_color!r}")
    return rgb


def _resolve_sun(solar_time, manual: dict, refraction_model: str):
    """Either the six manual values (with the reference's defaults, :1034-1043) or everything from a
    SolarTime-like object (:1002-1032).  The reference coerces through forge3d.geo.SolarTime (ephemeris code,
    out of scope: SURVEY section 2 row 24); any object with the same duck type is accepted here:
    position() -> {'azimuth_deg', 'true_elevation_deg', 'apparent_elevation_deg'} and the attributes
    observer_lat, observer_lon, pressure_mbar, temperature_c."""
    if solar_time is None:
        defaults = dict(sun_azimuth_deg=315.0, sun_elevation_deg=45.0, observer_latitude_deg=0.0,
                        observer_longitude_deg=0.0, pressure_mbar=1013.25, temperature_c=15.0)
        return {k: (defaults[k] if v is None else v) for k, v in manual.items()}, "manual_angles"
    if any(v is not None for v in manual.values()):
        raise ValueError("solar_time cannot be combined with manual sun, observer, pressure, or temperature values")
    if not hasattr(solar_time, "position"):
        raise TypeError("solar_time must provide position() and observer_lat/observer_lon/pressure_mbar/temperature_c")
    pos = solar_time.position()
    elevation_key = "true_elevation_deg" if refraction_model == "none" else "apparent_elevation_deg"
    resolved = dict(sun_azimuth_deg=pos["azimuth_deg"], sun_elevation_deg=pos[elevation_key],
                    observer_latitude_deg=solar_time.observer_lat, observer_longitude_deg=solar_time.observer_lon,
                    pressure_mbar=solar_time.pressure_mbar, temperature_c=solar_time.temperature_c)
    return resolved, "solar_time"


def _checked_mesh(mesh_vertices, mesh_indices):
    """Both or neither; (N,3) float32 + (M,3) uint32 (:1050-1059)."""
    if (mesh_vertices is None) != (mesh_indices is None):
        raise ValueError("mesh_vertices and mesh_indices must be provided together")
    if mesh_vertices is None:
        return None, None
    verts = np.ascontiguousarray(mesh_vertices, dtype=np.float32)
    tris = np.ascontiguousarray(mesh_indices, dtype=np.uint32)
    if verts.ndim != 2 or verts.shape[1] != 3:
        raise ValueError(f"mesh_vertices must be (N, 3), got {verts.shape}")
    if tris.ndim != 2 or tris.shape[1] != 3:
        raise ValueError(f"mesh_indices must be (M, 3), got {tris.shape}")
    return verts, tris


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

    Accumulates frames until the per-pixel luminance variance of the running mean over the last 32-frame
    window drops below ``variance_threshold`` (raises after ``max_frames``: no silent fake convergence).
    Returns ``rgba`` (H,W,4) uint8, ``albedo`` / ``normal`` (H,W,3) float32, ``depth`` (H,W) float32 ray
    distance (NaN on miss), ``frames``, ``variance``, ``converged``, ``peak_host_visible_bytes``,
    ``minmax_pyramid_bytes``, ``gpu_resource_bytes``, ``sun_source``, ``solar_azimuth_deg``,
    ``solar_elevation_deg`` (the reference's keys, terrain_reference.rs:437-450 and path_tracing.py:1092-1094)
    plus this backend's ray counters and timings.  ``cache`` is accepted and ignored, as in the reference.
    """
    del cache
    dem = _checked_dem(heightmap)
    if int(min_frames) > int(max_frames):
        raise ValueError(f"min_frames ({min_frames}) must be <= max_frames ({max_frames})")
    if not 1 <= int(spp) <= 64:
        raise ValueError(f"spp must be in 1..=64, got {spp}")
    if not (float(spacing[0]) > 0.0 and float(spacing[1]) > 0.0):
        raise ValueError(f"spacing must be > 0, got {spacing}")
    sun_rgb = _checked_sun_color(sun_color)
    sun, sun_source = _resolve_sun(
        solar_time,
        dict(sun_azimuth_deg=sun_azimuth_deg, sun_elevation_deg=sun_elevation_deg,
             observer_latitude_deg=observer_latitude_deg, observer_longitude_deg=observer_longitude_deg,
             pressure_mbar=pressure_mbar, temperature_c=temperature_c),
        refraction_model)
    env = None
    if env_map is not None:
        env = np.ascontiguousarray(env_map, dtype=np.float32)
        if env.ndim != 3 or env.shape[2] != 3:
            raise ValueError(f"env_map must be (H, W, 3) float32, got {env.shape}")
    verts, tris = _checked_mesh(mesh_vertices, mesh_indices)

    result = _NATIVE.hybrid_render_terrain_reference(
        dem, int(width), int(height), dict(camera or {}),
        spacing=(float(spacing[0]), float(spacing[1])),
        exaggeration=float(exaggeration),
        albedo=tuple(float(c) for c in albedo[:3]),
        sun_azimuth_deg=float(sun["sun_azimuth_deg"]),
        sun_elevation_deg=float(sun["sun_elevation_deg"]),
        sun_intensity=float(sun_intensity),
        sun_color=sun_rgb,
        env_map=env,
        env_intensity=float(env_intensity),
        mesh_vertices=verts,
        mesh_indices=tris,
        spp=int(spp),
        max_frames=int(max_frames),
        min_frames=int(min_frames),
        variance_threshold=float(variance_threshold),
        seed=int(seed),
        certificate=certificate,
        observer_latitude_deg=float(sun["observer_latitude_deg"]),
        observer_longitude_deg=float(sun["observer_longitude_deg"]),
        earth_model=earth_model,
        sphere_radius_m=float(sphere_radius_m),
        refraction_model=refraction_model,
        refraction_k=float(refraction_k),
        pressure_mbar=float(sun["pressure_mbar"]),
        temperature_c=float(sun["temperature_c"]),
        atmosphere=atmosphere,
    )
    result["sun_source"] = sun_source
    result["solar_azimuth_deg"] = float(sun["sun_azimuth_deg"])
    result["solar_elevation_deg"] = float(sun["sun_elevation_deg"])
    return result
