"""HELIOS viewshed and solar shadow mask on the GPU: `viewshed`, `shadow_mask_from_angles`, `compute_viewshed`, `compute_shadow_mask`.

Host-side stand-in (Python, where the reference is Rust + PyO3) for
  * `forge3d.viewshed`                     python/forge3d/terrain.py:32-80
  * `_forge3d.terrain_viewshed`            src/py_functions/geodesy.rs:184-392 (bounds / observer validation, per-cell geodesic
                                           offsets, ViewshedOptions)
  * `compute_viewshed / compute_shadow_mask`  src/terrain/analysis/viewshed.rs:341-347,396-570 - the native seam: these two go
                                           through the C ABI (`f3d_viewshed`, `f3d_shadow_mask`) to the CUDA kernels.
Two things the reference computes above the seam are outside this package and are stated as such: the Karney geodesic INVERSE
(src/geo/geodesic.rs) is replaced by Vincenty's inverse in float64 (agrees to < 0.1 mm, i.e. identically after the cast to f32
except at rounding ties; the sphere model uses the reference's own closed form), and `shadow_mask`'s per-cell SPA solar ephemeris
(src/geo/solar.rs) is not provided - `shadow_mask_from_angles` takes the sun's azimuth / launch elevation (scalars or per-cell
arrays) instead.  The EGM96 geoid conversion ('orthometric_egm96') is likewise not provided.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _native

_FOOTPRINT_ERROR = "viewshed geodesic leaves the DEM footprint"


def _wrap180(lon):
    return (lon + 180.0) % 360.0 - 180.0            # f64::rem_euclid


def vincenty_inverse(lat1_deg, lon1_deg, lat2_deg, lon2_deg):
    """WGS84 geodesic inverse (distance m, forward azimuth rad at point 1), float64, vectorised over point 2."""
    a, f = 6378137.0, 1.0 / 298.257223563
    b = a * (1.0 - f)
    u1 = math.atan((1.0 - f) * math.tan(math.radians(lat1_deg)))
    u2 = np.arctan((1.0 - f) * np.tan(np.radians(np.asarray(lat2_deg, np.float64))))
    L = np.radians(_wrap180(np.asarray(lon2_deg, np.float64) - lon1_deg))
    su1, cu1, su2, cu2 = math.sin(u1), math.cos(u1), np.sin(u2), np.cos(u2)
    lam = L.copy()
    for _ in range(200):
        sl, cl = np.sin(lam), np.cos(lam)
        ss = np.sqrt((cu2 * sl) ** 2 + (cu1 * su2 - su1 * cu2 * cl) ** 2)
        cs = su1 * su2 + cu1 * cu2 * cl
        sigma = np.arctan2(ss, cs)
        with np.errstate(invalid="ignore", divide="ignore"):
            sa = np.where(ss > 0, cu1 * cu2 * sl / ss, 0.0)
            c2a = 1.0 - sa * sa
            c2sm = np.where(c2a > 0, cs - 2.0 * su1 * su2 / c2a, 0.0)
        cc = f / 16.0 * c2a * (4.0 + f * (4.0 - 3.0 * c2a))
        new = L + (1.0 - cc) * f * sa * (sigma + cc * ss * (c2sm + cc * cs * (-1.0 + 2.0 * c2sm * c2sm)))
        done = np.max(np.abs(new - lam)) < 1e-14
        lam = new
        if done:
            break
    sl, cl = np.sin(lam), np.cos(lam)
    ss = np.sqrt((cu2 * sl) ** 2 + (cu1 * su2 - su1 * cu2 * cl) ** 2)
    cs = su1 * su2 + cu1 * cu2 * cl
    sigma = np.arctan2(ss, cs)
    with np.errstate(invalid="ignore", divide="ignore"):
        sa = np.where(ss > 0, cu1 * cu2 * sl / ss, 0.0)
        c2a = 1.0 - sa * sa
        c2sm = np.where(c2a > 0, cs - 2.0 * su1 * su2 / c2a, 0.0)
    usq = c2a * (a * a - b * b) / (b * b)
    A = 1.0 + usq / 16384.0 * (4096.0 + usq * (-768.0 + usq * (320.0 - 175.0 * usq)))
    B = usq / 1024.0 * (256.0 + usq * (-128.0 + usq * (74.0 - 47.0 * usq)))
    ds = B * ss * (c2sm + B / 4.0 * (cs * (-1.0 + 2.0 * c2sm ** 2) - B / 6.0 * c2sm * (-3.0 + 4.0 * ss ** 2) * (-3.0 + 4.0 * c2sm ** 2)))
    s = b * A * (sigma - ds)
    az = np.arctan2(cu2 * sl, cu1 * su2 - su1 * cu2 * cl)
    return np.where(ss > 0, s, 0.0), np.where(ss > 0, az, 0.0)


def make_options(width, height, **kw) -> dict:
    """ViewshedOptions as a dict (the keyword set of compute_viewshed / compute_shadow_mask)."""
    opts = dict(width=int(width), height=int(height), observer_x=0.0, observer_y=0.0, observer_height_m=0.0, target_height_m=0.0,
                max_distance_m=1.0, observer_latitude_rad=0.0, observer_longitude_rad=0.0, left_unwrapped_deg=0.0, top_deg=0.0,
                longitude_step_deg=1.0, latitude_step_deg=1.0, geodesic_sphere_radius_m=0.0, earth_model="ellipsoid",
                earth_latitude_deg=0.0, sphere_radius_m=6_371_008.8, refraction_model="bennett", refraction_k=0.13,
                pressure_mbar=1013.25, temperature_c=15.0)
    unknown = set(kw) - set(opts)
    if unknown:
        raise TypeError(f"unexpected viewshed options: {sorted(unknown)}")
    opts.update(kw)
    return opts


def _native_options(opts: dict, device: int):
    if opts["earth_model"] not in _native.EARTH_MODELS:
        raise ValueError(f"unsupported earth_model {opts['earth_model']!r}")
    if opts["refraction_model"] not in _native.REFRACTION_MODELS:
        raise ValueError(f"unsupported refraction_model {opts['refraction_model']!r}")
    o = _native.ViewshedOptions()
    o.width, o.height = int(opts["width"]), int(opts["height"])
    for name in ("observer_x", "observer_y", "observer_height_m", "target_height_m", "max_distance_m", "observer_latitude_rad",
                 "observer_longitude_rad", "left_unwrapped_deg", "top_deg", "longitude_step_deg", "latitude_step_deg",
                 "geodesic_sphere_radius_m"):
        setattr(o, name, float(opts[name]))
    o.earth_model = _native.EARTH_MODELS[opts["earth_model"]]
    o.refraction_model = _native.REFRACTION_MODELS[opts["refraction_model"]]
    for name in ("earth_latitude_deg", "sphere_radius_m", "refraction_k", "pressure_mbar", "temperature_c"):
        setattr(o, name, float(opts[name]))
    o.device = int(device)
    return o


def compute_viewshed(heights, positions_m, options: dict, *, device: int = 0) -> dict:
    """compute_viewshed(heights, positions_m, options) (viewshed.rs:341-347) on the CUDA backend."""
    dem = np.ascontiguousarray(heights, dtype=np.float32)
    pos = np.ascontiguousarray(positions_m, dtype=np.float32)
    if dem.ndim != 2 or pos.size != dem.size * 2 or dem.shape != (int(options["height"]), int(options["width"])):
        raise RuntimeError(f"DEM/position lengths do not match supported dimensions {options['width']}x{options['height']} "
                           "(both dimensions must be at least 2 and packed traversal supports at most 8192 cells per axis)")
    o = _native_options(options, device)
    vis = np.zeros(dem.shape, np.uint8)
    drop, gain, horizon = (np.zeros(dem.shape, np.float32) for _ in range(3))
    ms = C.c_double()
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    _native.check(_native.lib().f3d_viewshed(fp(dem), fp(pos), C.byref(o), vis.ctypes.data_as(C.POINTER(C.c_uint8)), fp(drop), fp(gain),
                                             fp(horizon), C.byref(ms)))
    return dict(visibility=vis.astype(np.bool_), curvature_drop_m=drop, refraction_gain_m=gain, horizon_distance_m=horizon,
                kernel_ms=float(ms.value))


def compute_shadow_mask(heights, geodetic_positions_and_sun_rad, options: dict, *, device: int = 0) -> np.ndarray:
    """compute_shadow_mask(heights, [lat, lon, sun azimuth, launch elevation] per cell in radians, options) (viewshed.rs:396-570)."""
    dem = np.ascontiguousarray(heights, dtype=np.float32)
    inp = np.ascontiguousarray(geodetic_positions_and_sun_rad, dtype=np.float32)
    if dem.ndim != 2 or inp.size != dem.size * 4 or not np.isfinite(inp).all():
        raise RuntimeError("shadow-mask geodetic/solar inputs do not match the DEM")
    o = _native_options(options, device)
    lit = np.zeros(dem.shape, np.uint8)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    _native.check(_native.lib().f3d_shadow_mask(fp(dem), fp(inp), C.byref(o), lit.ctypes.data_as(C.POINTER(C.c_uint8)), None))
    return lit.astype(np.bool_)


def _grid(dem, bounds, height_system):
    """terrain_grid_heights + the bounds checks shared by both entry points (geodesy.rs:214-262)."""
    left, bottom, right, top = (float(v) for v in bounds)
    right_unwrapped = right + 360.0 if right <= left else right
    span = right_unwrapped - left
    array = np.ascontiguousarray(dem, dtype=np.float32)
    if array.ndim != 2 or array.shape[0] < 2 or array.shape[1] < 2:
        raise ValueError("dem must be a two-dimensional array at least 2x2")
    if height_system == "orthometric_egm96":
        raise NotImplementedError("the EGM96 geoid model (src/geo/geoid.rs) is outside forge3d_b200: convert to ellipsoidal heights first")
    if height_system != "ellipsoidal":
        raise ValueError(f"unsupported height_system {height_system!r}; expected 'ellipsoidal' or 'orthometric_egm96'")
    height, width = array.shape
    return array, left, bottom, right_unwrapped, top, span, span / width, (top - bottom) / height


def viewshed_inputs(dem, observer, *, bounds, height_system, observer_height=1.7, target_height=0.0, max_distance=None,
                    earth_model="ellipsoid", sphere_radius_m=6_371_008.8, refraction_model="bennett", refraction_k=0.13,
                    pressure_mbar=1013.25, temperature_c=15.0):
    """Everything `terrain_viewshed` computes above the native seam: (heights, positions_m, options)."""
    observer_lat, observer_lon = float(observer[0]), float(observer[1])
    array, left, bottom, right_unwrapped, top, span, lon_step, lat_step = _grid(dem, bounds, height_system)
    obs_unwrapped = observer_lon + 360.0 if observer_lon < left else observer_lon
    ok = (-90.0 <= observer_lat <= 90.0 and -180.0 <= observer_lon <= 180.0 and -180.0 <= left <= 180.0 and
          -180.0 <= float(bounds[2]) <= 180.0 and span > 0.0 and top > bottom and span < 180.0 and top - bottom < 180.0 and
          -90.0 <= bottom <= 90.0 and -90.0 <= top <= 90.0 and left <= obs_unwrapped <= right_unwrapped and bottom <= observer_lat <= top)
    if not ok:
        raise ValueError("observer must be finite and inside local EPSG:4326 bounds spanning less than 180 degrees")
    if earth_model not in _native.EARTH_MODELS:
        raise ValueError(f"unsupported earth_model {earth_model!r}")
    if refraction_model not in _native.REFRACTION_MODELS:
        raise ValueError(f"unsupported refraction_model {refraction_model!r}")
    height, width = array.shape
    lat = top - (np.arange(height, dtype=np.float64) + 0.5) * lat_step
    lon = _wrap180(left + (np.arange(width, dtype=np.float64) + 0.5) * lon_step)
    lat2, lon2 = np.meshgrid(lat, lon, indexing="ij")
    obs_lon_n = _wrap180(observer_lon)
    if earth_model == "sphere":
        l1, l2, dl = math.radians(observer_lat), np.radians(lat2), np.radians(lon2 - obs_lon_n)
        central = np.arccos(np.clip(math.sin(l1) * np.sin(l2) + math.cos(l1) * np.cos(l2) * np.cos(dl), -1.0, 1.0))
        azimuth = np.arctan2(np.sin(dl) * np.cos(l2), math.cos(l1) * np.sin(l2) - math.sin(l1) * np.cos(l2) * np.cos(dl))
        distance = float(sphere_radius_m) * central
    else:
        distance, azimuth = vincenty_inverse(observer_lat, obs_lon_n, lat2, lon2)
    positions = np.stack([(distance * np.sin(azimuth)).astype(np.float32), (distance * np.cos(azimuth)).astype(np.float32)], axis=-1)
    max_distance_m = float(distance.max()) if max_distance is None else float(max_distance)
    opts = make_options(width, height,
                        observer_x=(obs_unwrapped - left) / lon_step - 0.5, observer_y=(top - observer_lat) / lat_step - 0.5,
                        observer_height_m=float(observer_height), target_height_m=float(target_height), max_distance_m=max_distance_m,
                        observer_latitude_rad=math.radians(observer_lat), observer_longitude_rad=math.radians(obs_lon_n),
                        left_unwrapped_deg=left, top_deg=top, longitude_step_deg=lon_step, latitude_step_deg=lat_step,
                        geodesic_sphere_radius_m=float(sphere_radius_m) if earth_model == "sphere" else 0.0,
                        earth_model=earth_model, earth_latitude_deg=observer_lat, sphere_radius_m=float(sphere_radius_m),
                        refraction_model=refraction_model, refraction_k=float(refraction_k), pressure_mbar=float(pressure_mbar),
                        temperature_c=float(temperature_c))
    return array, positions, opts


def viewshed(dem, observer, *, bounds, height_system, observer_height=1.7, target_height=0.0, max_distance=None,
             earth_model="ellipsoid", sphere_radius_m=6_371_008.8, refraction_model="bennett", refraction_k=0.13,
             pressure_mbar=1013.25, temperature_c=15.0, return_diagnostics=False):
    """`forge3d.viewshed` (python/forge3d/terrain.py:32-80): GPU visibility raster for an EPSG:4326 north-up DEM."""
    if len(observer) == 3:
        observer_height = float(observer[2])
    elif len(observer) != 2:
        raise ValueError("observer must be (lat, lon) or (lat, lon, h_agl)")
    heights, positions, opts = viewshed_inputs(dem, observer, bounds=bounds, height_system=height_system, observer_height=observer_height,
                                               target_height=target_height, max_distance=max_distance, earth_model=earth_model,
                                               sphere_radius_m=sphere_radius_m, refraction_model=refraction_model, refraction_k=refraction_k,
                                               pressure_mbar=pressure_mbar, temperature_c=temperature_c)
    try:
        out = compute_viewshed(heights, positions, opts)
    except RuntimeError as error:
        raise RuntimeError(f"viewshed failed: {error}") from None
    out.pop("kernel_ms")
    return out if return_diagnostics else out["visibility"]


def shadow_mask_inputs(dem, sun_azimuth_deg, sun_elevation_deg, *, bounds, height_system, observer=None, earth_model="ellipsoid",
                       sphere_radius_m=6_371_008.8, refraction_model="bennett", refraction_k=0.13, pressure_mbar=1013.25,
                       temperature_c=15.0):
    """What `terrain_shadow_mask` computes above the seam (geodesy.rs:405-520), with the sun's direction given instead of SPA."""
    array, left, bottom, right_unwrapped, top, span, lon_step, lat_step = _grid(dem, bounds, height_system)
    height, width = array.shape
    lat = top - (np.arange(height, dtype=np.float64) + 0.5) * lat_step
    lon = _wrap180(left + (np.arange(width, dtype=np.float64) + 0.5) * lon_step)
    lat2, lon2 = np.meshgrid(lat, lon, indexing="ij")
    az = np.broadcast_to(np.asarray(sun_azimuth_deg, np.float64), array.shape)
    el = np.broadcast_to(np.asarray(sun_elevation_deg, np.float64), array.shape)
    inputs = np.stack([np.radians(lat2), np.radians(lon2), np.radians(az), np.radians(el)], axis=-1).astype(np.float32)
    right = _wrap180(right_unwrapped)
    if earth_model == "sphere":
        def central(la1, lo1, la2, lo2):
            a, b, d = math.radians(la1), math.radians(la2), math.radians(lo2 - lo1)
            return math.acos(min(max(math.sin(a) * math.sin(b) + math.cos(a) * math.cos(b) * math.cos(d), -1.0), 1.0))
        diagonal = float(sphere_radius_m) * max(central(top, left, bottom, right), central(top, right, bottom, left))
    else:
        diagonal = max(float(vincenty_inverse(top, left, bottom, right)[0]), float(vincenty_inverse(top, right, bottom, left)[0]))
    obs_lat, obs_lon = (0.5 * (top + bottom), _wrap180(left + 0.5 * span)) if observer is None else (float(observer[0]), float(observer[1]))
    opts = make_options(width, height, max_distance_m=diagonal * 1.01, observer_latitude_rad=math.radians(obs_lat),
                        observer_longitude_rad=math.radians(obs_lon), left_unwrapped_deg=left, top_deg=top, longitude_step_deg=lon_step,
                        latitude_step_deg=lat_step, geodesic_sphere_radius_m=float(sphere_radius_m) if earth_model == "sphere" else 0.0,
                        earth_model=earth_model, earth_latitude_deg=obs_lat, sphere_radius_m=float(sphere_radius_m),
                        refraction_model=refraction_model, refraction_k=float(refraction_k), pressure_mbar=float(pressure_mbar),
                        temperature_c=float(temperature_c))
    return array, inputs, opts


def shadow_mask_from_angles(dem, sun_azimuth_deg, sun_elevation_deg, *, bounds, height_system, **kw) -> np.ndarray:
    """DEM-local terrain-to-sun visibility (True = lit) for a given sun direction; see the module docstring."""
    heights, inputs, opts = shadow_mask_inputs(dem, sun_azimuth_deg, sun_elevation_deg, bounds=bounds, height_system=height_system, **kw)
    try:
        return compute_shadow_mask(heights, inputs, opts)
    except RuntimeError as error:
        raise RuntimeError(f"shadow mask failed: {error}") from None
