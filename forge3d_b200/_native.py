"""ctypes binding of libforge3d_b200.so -- the stand-in for the reference's PyO3 module
`forge3d._forge3d` on the path-traced terrain path (src/py_functions/path_tracing/terrain_reference.rs).

`hybrid_render_terrain_reference` below has the native seam's positional order and defaults
(terrain_reference.rs:224-256).  There is no CPU fallback: if the library is missing or no CUDA
device is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
# F3D_B200_LIB: developer override used only for tuning experiments (alternative builds of the same sources)
LIB_PATH = Path(os.environ["F3D_B200_LIB"]) if os.environ.get("F3D_B200_LIB") else _PKG / "libforge3d_b200.so"

EARTH_MODELS = {"flat": 0, "sphere": 1, "ellipsoid": 2, "wgs84": 2}
REFRACTION_MODELS = {"none": 0, "bennett": 1, "saemundsson": 2, "effective_radius": 3}
IPC_HANDLE_BYTES = 64
IPC_HANDLES_PER_RANK = 3


class Atmosphere(C.Structure):
    """f3d_atmosphere (include/forge3d_b200.h)."""
    _fields_ = [
        ("transmittance", C.POINTER(C.c_uint16)), ("scattering", C.POINTER(C.c_uint16)), ("aerial", C.POINTER(C.c_uint16)),
        ("transmittance_mu", C.c_uint32), ("transmittance_height", C.c_uint32),
        ("scattering_mu_view", C.c_uint32), ("scattering_mu_sun", C.c_uint32),
        ("scattering_height", C.c_uint32), ("scattering_nu", C.c_uint32),
        ("aerial_distance", C.c_uint32), ("aerial_mu_view", C.c_uint32), ("aerial_height", C.c_uint32),
        ("bottom_radius_m", C.c_float), ("top_radius_m", C.c_float), ("max_aerial_distance_m", C.c_float),
        ("ozone_du", C.c_float), ("mie_g", C.c_float), ("turbidity", C.c_float),
        ("rayleigh_scale_height_m", C.c_float), ("mie_scale_height_m", C.c_float), ("ground_albedo", C.c_float),
    ]


def make_atmosphere(handle) -> tuple:
    """AtmosphereLutHandle (forge3d_b200.atmosphere) -> (f3d_atmosphere, keepalive)."""
    a = Atmosphere()
    u16 = lambda arr: arr.ctypes.data_as(C.POINTER(C.c_uint16))
    a.transmittance, a.scattering, a.aerial = u16(handle.transmittance), u16(handle.scattering), u16(handle.aerial)
    cfg, d = handle.config, handle.config.dimensions
    for name in ("transmittance_mu", "transmittance_height", "scattering_mu_view", "scattering_mu_sun", "scattering_height",
                 "scattering_nu", "aerial_distance", "aerial_mu_view", "aerial_height"):
        setattr(a, name, int(getattr(d, name)))
    for name in ("bottom_radius_m", "top_radius_m", "max_aerial_distance_m", "ozone_du", "mie_g", "turbidity",
                 "rayleigh_scale_height_m", "mie_scale_height_m", "ground_albedo"):
        setattr(a, name, float(getattr(cfg, name)))
    return a, [handle, a]


class SmokeVolume(C.Structure):
    """f3d_smoke_volume (include/forge3d_b200.h)."""
    _fields_ = [
        ("dims", C.c_uint32 * 3), ("voxel_size", C.c_float * 3), ("origin", C.c_float * 3),
        ("density", C.POINTER(C.c_float)), ("temperature", C.POINTER(C.c_float)), ("soot", C.POINTER(C.c_float)),
        ("humidity", C.POINTER(C.c_float)), ("emission_rate", C.POINTER(C.c_float)), ("particle_age", C.POINTER(C.c_float)),
        ("frame_index", C.c_uint64),
    ]


class SmokeSettings(C.Structure):
    """f3d_smoke_settings (include/forge3d_b200.h)."""
    _fields_ = [
        ("density_scale", C.c_float), ("extinction", C.c_float), ("scattering", C.c_float), ("absorption", C.c_float),
        ("phase_g", C.c_float), ("step_size", C.c_float), ("max_steps", C.c_uint32), ("self_shadow", C.c_int32),
        ("shadow_steps", C.c_uint32), ("shadow_step_size", C.c_float), ("jitter_strength", C.c_float), ("exposure", C.c_float),
        ("thin_color", C.c_float * 3), ("dense_color", C.c_float * 3), ("soot_absorption", C.c_float), ("fire_glow", C.c_float),
    ]


class ViewshedOptions(C.Structure):
    """f3d_viewshed_options (include/forge3d_b200.h)."""
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32)] + [(n, C.c_float) for n in (
        "observer_x", "observer_y", "observer_height_m", "target_height_m", "max_distance_m", "observer_latitude_rad",
        "observer_longitude_rad", "left_unwrapped_deg", "top_deg", "longitude_step_deg", "latitude_step_deg",
        "geodesic_sphere_radius_m")] + [
        ("earth_model", C.c_int32), ("earth_latitude_deg", C.c_double), ("sphere_radius_m", C.c_double),
        ("refraction_model", C.c_int32), ("refraction_k", C.c_double), ("pressure_mbar", C.c_double), ("temperature_c", C.c_double),
        ("device", C.c_int32)]


class WavefrontSceneC(C.Structure):
    """f3d_wavefront_scene (include/forge3d_b200.h)."""
    _fields_ = [
        ("cam_origin", C.c_float * 3), ("cam_forward", C.c_float * 3), ("cam_right", C.c_float * 3), ("cam_up", C.c_float * 3),
        ("fov_y_rad", C.c_float), ("exposure", C.c_float), ("seed_hi", C.c_uint32), ("seed_lo", C.c_uint32),
        ("spheres", C.POINTER(C.c_float)), ("nspheres", C.c_uint32),
        ("dir_lights", C.POINTER(C.c_float)), ("ndir", C.c_uint32),
        ("area_lights", C.POINTER(C.c_float)), ("narea", C.c_uint32),
        ("importance", C.POINTER(C.c_float)), ("nimportance", C.c_uint32),
        ("environment", C.c_float * 16),
        ("mesh_xyz", C.POINTER(C.c_float)), ("mesh_nverts", C.c_uint32),
        ("mesh_idx", C.POINTER(C.c_uint32)), ("mesh_ntris", C.c_uint32),
        ("instances", C.POINTER(C.c_float)), ("ninstances", C.c_uint32),
    ]


class WavefrontStats(C.Structure):
    """f3d_wavefront_stats (include/forge3d_b200.h)."""
    _fields_ = [("rays", C.c_uint64), ("max_rays_per_frame", C.c_uint64), ("min_iterations", C.c_uint32), ("launches", C.c_uint32),
                ("kernel_ms", C.c_double)]


class WavefrontPart(C.Structure):
    """f3d_wavefront_part (include/forge3d_b200.h)."""
    _fields_ = [("rank", C.c_uint32), ("world", C.c_uint32), ("block_rows", C.c_uint32), ("frame_iterations", C.POINTER(C.c_uint32)),
                ("frame_rays", C.POINTER(C.c_uint64))]


def fill_wavefront_scene(cs, s):
    """Fills a WavefrontSceneC-shaped ctypes struct from a normalized forge3d_b200.wavefront.WavefrontScene; returns the arrays
    that must outlive the call."""
    fpt, upt = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    for name in ("cam_origin", "cam_forward", "cam_right", "cam_up"):
        getattr(cs, name)[:] = [float(x) for x in getattr(s, name)]
    cs.fov_y_rad, cs.exposure, cs.seed_hi, cs.seed_lo = s.fov_y_rad, s.exposure, s.seed_hi, s.seed_lo
    cs.environment[:] = [float(x) for x in s.environment]
    keep = [s.spheres, s.dir_lights, s.area_lights, s.importance, s.mesh_xyz, s.mesh_idx, s.instances]
    ptr = lambda a, t: a.ctypes.data_as(t) if a.size else t()
    cs.spheres, cs.nspheres = ptr(s.spheres, fpt), s.spheres.shape[0]
    cs.dir_lights, cs.ndir = ptr(s.dir_lights, fpt), s.dir_lights.shape[0]
    cs.area_lights, cs.narea = ptr(s.area_lights, fpt), s.area_lights.shape[0]
    cs.importance, cs.nimportance = ptr(s.importance, fpt), s.importance.shape[0]
    cs.mesh_xyz, cs.mesh_nverts = ptr(s.mesh_xyz, fpt), s.mesh_xyz.shape[0]
    cs.mesh_idx, cs.mesh_ntris = ptr(s.mesh_idx, upt), s.mesh_idx.shape[0]
    cs.instances, cs.ninstances = ptr(s.instances, fpt), s.instances.shape[0]
    return keep


def make_wavefront_scene(s):
    cs = WavefrontSceneC()
    return cs, fill_wavefront_scene(cs, s)


def raise_last(rc: int) -> None:
    check(rc)


class TerrainDesc(C.Structure):
    """f3d_terrain_desc (include/forge3d_b200.h)."""
    _fields_ = [
        ("heights", C.POINTER(C.c_float)), ("dem_w", C.c_uint32), ("dem_h", C.c_uint32),
        ("spacing", C.c_float * 2), ("exaggeration", C.c_float), ("albedo", C.c_float * 3),
        ("cam_origin", C.c_float * 3), ("cam_look_at", C.c_float * 3), ("cam_up", C.c_float * 3),
        ("fov_y_deg", C.c_float), ("exposure", C.c_float),
        ("sun_az_deg", C.c_float), ("sun_el_deg", C.c_float), ("sun_intensity", C.c_float),
        ("sun_color", C.c_float * 3),
        ("observer_lat_deg", C.c_double), ("observer_lon_deg", C.c_double),
        ("earth_model", C.c_int32), ("sphere_radius_m", C.c_double),
        ("refraction_model", C.c_int32), ("refraction_k", C.c_double),
        ("pressure_mbar", C.c_double), ("temperature_c", C.c_double),
        ("env_rgb", C.POINTER(C.c_float)), ("env_w", C.c_uint32), ("env_h", C.c_uint32),
        ("env_intensity", C.c_float),
        ("mesh_xyz", C.POINTER(C.c_float)), ("mesh_nverts", C.c_uint32),
        ("mesh_idx", C.POINTER(C.c_uint32)), ("mesh_ntris", C.c_uint32),
        ("width", C.c_uint32), ("height", C.c_uint32), ("seed", C.c_uint32), ("spp", C.c_uint32),
        ("max_frames", C.c_uint32), ("min_frames", C.c_uint32), ("variance_threshold", C.c_float),
        ("device", C.c_int32), ("compat_512mib_gate", C.c_int32),
        ("part_rank", C.c_uint32), ("part_world", C.c_uint32), ("part_block_rows", C.c_uint32), ("part_mode", C.c_uint32),
        ("atmosphere", C.POINTER(Atmosphere)),
    ]


class TerrainOut(C.Structure):
    """f3d_terrain_out (include/forge3d_b200.h)."""
    _fields_ = [
        ("rgba", C.POINTER(C.c_uint8)), ("albedo", C.POINTER(C.c_float)),
        ("normal", C.POINTER(C.c_float)), ("depth", C.POINTER(C.c_float)), ("accum", C.POINTER(C.c_float)),
        ("frames", C.c_uint32), ("variance", C.c_float), ("converged", C.c_int32),
        ("peak_host_visible_bytes", C.c_uint64), ("minmax_pyramid_bytes", C.c_uint64),
        ("gpu_resource_bytes", C.c_uint64),
        ("rays_primary", C.c_uint64), ("rays_shadow", C.c_uint64), ("rays_ibl", C.c_uint64),
        ("nodes_popped", C.c_uint64),
        ("setup_ms", C.c_double), ("frames_ms", C.c_double), ("readback_ms", C.c_double),
        ("kernel_launches", C.c_uint64),
    ]


# every symbol include/forge3d_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "f3d_terrain_reference_render", "f3d_last_error", "f3d_abi_version", "f3d_device_count", "f3d_build_info", "f3d_host_alloc", "f3d_host_free", "f3d_cache_trim",
    "f3d_session_create", "f3d_session_render_frames", "f3d_session_variance",
    "f3d_session_resolve_device", "f3d_session_validity", "f3d_session_resolve_host", "f3d_session_frames",
    "f3d_session_stats", "f3d_session_sync", "f3d_session_last_frames_ms", "f3d_session_destroy",
    "f3d_session_ipc_export", "f3d_session_ipc_import", "f3d_trace_rays", "f3d_build_minmax",
    "f3d_smoke_create", "f3d_smoke_destroy", "f3d_smoke_raymarch_rgba", "f3d_smoke_raymarch_over_rgba", "f3d_smoke_raymarch_projection_rgba",
    "f3d_viewshed", "f3d_shadow_mask", "f3d_lbvh_build", "f3d_wavefront_render", "f3d_wavefront_render_part",
]

_libs = {}
NUMERICS = ("exact", "fast")


def default_numerics() -> str:
    """"exact" (bit-identical to the oracle: the numerics contract of DESIGN.md section 4) unless F3D_B200_NUMERICS=fast selects
    the throughput build (csrc/f3d_math.cuh F3D_FAST_NUMERICS; accepted by tolerance, tests/test_fast_numerics.py)."""
    return os.environ.get("F3D_B200_NUMERICS", "exact") or "exact"


def lib(numerics: str | None = None):
    """Loads the in-tree CUDA backend; fails loudly when it is missing (no fallback).  `numerics` picks the build: "exact"
    (libforge3d_b200.so) or "fast" (libforge3d_b200_fast.so); None = default_numerics().  Only the terrain session /
    render entry points are meant to be used from the fast build; the widened rows always bind the exact one."""
    numerics = numerics or default_numerics()
    if numerics not in NUMERICS:
        raise ValueError(f"numerics must be one of {NUMERICS}, got {numerics!r}")
    if numerics in _libs:
        return _libs[numerics]
    path = LIB_PATH if numerics == "exact" else _PKG / "libforge3d_b200_fast.so"
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build the CUDA backend with `python -m forge3d_b200.build`{' --fast' if numerics == 'fast' else ''} "
            "(or __graft_entry__.build()); forge3d_b200 has no CPU fallback")
    L = C.CDLL(str(path))
    vp, u8p, fp, u32p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    L.f3d_terrain_reference_render.argtypes = [C.POINTER(TerrainDesc), C.POINTER(TerrainOut)]
    L.f3d_last_error.restype = C.c_char_p
    L.f3d_build_info.restype = C.c_char_p
    L.f3d_host_alloc.restype = C.c_void_p
    L.f3d_host_alloc.argtypes = [C.c_uint64]
    L.f3d_host_free.argtypes = [C.c_void_p]
    L.f3d_cache_trim.restype = C.c_uint64
    L.f3d_cache_trim.argtypes = [C.c_int32]
    L.f3d_session_create.argtypes = [C.POINTER(TerrainDesc), vp, C.POINTER(vp)]
    L.f3d_session_render_frames.argtypes = [vp, C.c_uint32]
    L.f3d_session_variance.argtypes = [vp, fp, C.POINTER(C.c_int32)]
    L.f3d_session_resolve_device.argtypes = [vp, vp, vp, vp, vp, C.c_int32]
    L.f3d_session_validity.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.f3d_session_resolve_host.argtypes = [vp, C.POINTER(TerrainOut)]
    L.f3d_session_frames.argtypes = [vp, u32p]
    L.f3d_session_stats.argtypes = [vp, C.POINTER(TerrainOut)]
    L.f3d_session_sync.argtypes = [vp]
    L.f3d_session_last_frames_ms.argtypes = [vp, C.POINTER(C.c_double)]
    L.f3d_session_destroy.argtypes = [vp]
    L.f3d_session_destroy.restype = None
    L.f3d_session_ipc_export.argtypes = [vp, u8p]
    L.f3d_session_ipc_import.argtypes = [vp, u8p]
    L.f3d_trace_rays.argtypes = [fp, C.c_uint32, C.c_uint32, fp, fp, C.c_float, C.c_float, C.c_int32,
                                 fp, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, u8p, fp, fp,
                                 C.POINTER(C.c_uint64)]
    L.f3d_build_minmax.argtypes = [fp, C.c_uint32, C.c_uint32, C.c_int32, u32p, fp, C.c_uint64]
    f3 = C.c_float * 3
    L.f3d_smoke_create.argtypes = [C.POINTER(SmokeVolume), C.c_int32, C.POINTER(vp)]
    L.f3d_smoke_destroy.argtypes = [vp]
    L.f3d_smoke_destroy.restype = None
    L.f3d_smoke_raymarch_rgba.argtypes = [vp, C.POINTER(SmokeSettings), C.c_uint32, C.c_uint32, f3, f3, f3, C.c_float, f3, u8p,
                                          C.POINTER(C.c_double)]
    L.f3d_smoke_raymarch_over_rgba.argtypes = [vp, C.POINTER(SmokeSettings), C.c_uint32, C.c_uint32, f3, f3, f3, C.c_float, f3, u8p, fp, u8p,
                                               C.POINTER(C.c_double)]
    L.f3d_smoke_raymarch_projection_rgba.argtypes = [vp, C.POINTER(SmokeSettings), C.c_uint32, C.c_uint32, f3, f3, u8p,
                                                     C.POINTER(C.c_double)]
    L.f3d_viewshed.argtypes = [fp, fp, C.POINTER(ViewshedOptions), u8p, fp, fp, fp, C.POINTER(C.c_double)]
    L.f3d_shadow_mask.argtypes = [fp, fp, C.POINTER(ViewshedOptions), u8p, C.POINTER(C.c_double)]
    L.f3d_lbvh_build.argtypes = [fp, C.c_uint32, u32p, C.c_uint32, C.c_int32, u32p, u32p, u32p, u32p, u32p, fp]
    L.f3d_wavefront_render.argtypes = [C.POINTER(WavefrontSceneC), C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, fp, u8p,
                                       C.POINTER(WavefrontStats)]
    L.f3d_wavefront_render_part.argtypes = [C.POINTER(WavefrontSceneC), C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.POINTER(WavefrontPart),
                                            fp, u8p, C.POINTER(WavefrontStats)]
    for name in EXPORTS:   # fail at load time, not at first call, if the ABI drifted
        getattr(L, name)
    _libs[numerics] = L
    return L


def last_error(L=None) -> str:
    return (L or lib()).f3d_last_error().decode("utf-8", "replace")


def check(rc: int, L=None) -> None:
    """Maps f3d_status to the exception types PyO3 raises for RenderError (src/core/error.rs).  L = the library the call
    went to (the error text lives there); None = the default build."""
    if rc == 0:
        return
    msg = last_error(L)
    if rc == 5:
        raise ValueError(msg)
    if rc == 4:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class DeviceHeights:
    """A DEM that already lives in the memory of the CUDA device the session runs on (e.g. broadcast over NVLink by
    forge3d_b200.distributed): raw device pointer of a contiguous float32 (H, W) array + whatever keeps it alive."""

    def __init__(self, ptr: int, shape, keep=None):
        self.ptr, self.shape, self.keep = int(ptr), (int(shape[0]), int(shape[1])), keep


def make_desc(heightmap, width, height, cam, *, spacing, exaggeration, albedo, sun_azimuth_deg,
              sun_elevation_deg, sun_intensity, env_map, env_intensity, mesh_vertices, mesh_indices, spp,
              max_frames, min_frames, variance_threshold, seed, sun_color, observer_latitude_deg,
              observer_longitude_deg, earth_model, sphere_radius_m, refraction_model, refraction_k,
              pressure_mbar, temperature_c, device=0, compat_512mib_gate=False, part_rank=0, part_world=1,
              part_block_rows=0, atmosphere=None, part_mode=0):
    """Marshals the native seam's arguments into f3d_terrain_desc (terrain_reference.rs:295-414).
    Returns (desc, keepalive)."""
    if earth_model not in EARTH_MODELS:
        raise ValueError(f"unsupported earth_model {earth_model!r}")          # refraction.rs:52
    if refraction_model not in REFRACTION_MODELS:
        raise ValueError(f"unsupported refraction_model {refraction_model!r}")  # refraction.rs:97
    d = TerrainDesc()
    if isinstance(heightmap, DeviceHeights):
        keep = [heightmap]
        d.heights = C.cast(C.c_void_p(heightmap.ptr), C.POINTER(C.c_float))
        d.dem_h, d.dem_w = heightmap.shape
    else:
        dem = np.ascontiguousarray(heightmap, dtype=np.float32)
        if dem.ndim != 2:
            raise ValueError(f"heightmap must be 2D (H, W), got shape {dem.shape}")
        keep = [dem]
        d.heights = _fp(dem)
        d.dem_h, d.dem_w = dem.shape
    d.spacing = (C.c_float * 2)(float(spacing[0]), float(spacing[1]))
    d.exaggeration = float(exaggeration)
    d.albedo = (C.c_float * 3)(*[float(v) for v in albedo])
    cam = dict(cam or {})
    d.cam_origin = (C.c_float * 3)(*[float(v) for v in cam.get("origin", (0.0, 50.0, 120.0))])
    d.cam_look_at = (C.c_float * 3)(*[float(v) for v in cam.get("look_at", (0.0, 0.0, 0.0))])
    d.cam_up = (C.c_float * 3)(*[float(v) for v in cam.get("up", (0.0, 1.0, 0.0))])
    d.fov_y_deg = float(cam.get("fov_y", 45.0))
    d.exposure = float(cam.get("exposure", 1.0))
    d.sun_az_deg, d.sun_el_deg, d.sun_intensity = float(sun_azimuth_deg), float(sun_elevation_deg), float(sun_intensity)
    d.sun_color = (C.c_float * 3)(*[float(v) for v in sun_color])
    d.observer_lat_deg, d.observer_lon_deg = float(observer_latitude_deg), float(observer_longitude_deg)
    d.earth_model = EARTH_MODELS[earth_model]
    d.sphere_radius_m = float(sphere_radius_m)
    d.refraction_model = REFRACTION_MODELS[refraction_model]
    d.refraction_k, d.pressure_mbar, d.temperature_c = float(refraction_k), float(pressure_mbar), float(temperature_c)
    if env_map is not None:
        env = np.ascontiguousarray(env_map, dtype=np.float32)
        if env.ndim != 3 or env.shape[2] != 3:
            raise ValueError("env_map must have shape (H, W, 3)")
        keep.append(env)
        d.env_rgb = _fp(env)
        d.env_h, d.env_w = env.shape[0], env.shape[1]
    d.env_intensity = float(env_intensity)
    if (mesh_vertices is None) != (mesh_indices is None):
        raise ValueError("mesh_vertices and mesh_indices must be provided together")
    if mesh_vertices is not None:
        mv = np.ascontiguousarray(mesh_vertices, dtype=np.float32)
        mi = np.ascontiguousarray(mesh_indices, dtype=np.uint32)
        if mv.ndim != 2 or mv.shape[1] != 3:
            raise ValueError("mesh_vertices must have shape (N, 3)")
        if mi.ndim != 2 or mi.shape[1] != 3:
            raise ValueError("mesh_indices must have shape (M, 3)")
        keep += [mv, mi]
        d.mesh_xyz = _fp(mv)
        d.mesh_nverts = mv.shape[0]
        d.mesh_idx = mi.ctypes.data_as(C.POINTER(C.c_uint32))
        d.mesh_ntris = mi.shape[0]
    d.width, d.height, d.seed, d.spp = int(width), int(height), int(seed) & 0xFFFFFFFF, int(spp)
    d.max_frames, d.min_frames = int(max_frames), int(min_frames)
    d.variance_threshold = float(variance_threshold)
    d.device = int(device)
    d.compat_512mib_gate = int(bool(compat_512mib_gate))
    d.part_rank, d.part_world, d.part_block_rows = int(part_rank), int(part_world), int(part_block_rows)
    d.part_mode = int(part_mode)
    if atmosphere is not None:   # an AtmosphereLutHandle (see forge3d_b200.atmosphere.resolve_atmosphere)
        atm, atm_keep = make_atmosphere(atmosphere)
        keep += atm_keep
        d.atmosphere = C.pointer(atm)
    return d, keep


def extract_sun_color(obj):
    """extract_sun_color, terrain_reference.rs:13-43: exactly three finite non-negative numbers."""
    def reject():
        return ValueError("sun_color must be exactly three finite, non-negative numbers")
    if isinstance(obj, (str, bytes, bytearray, memoryview)):
        raise reject()
    try:
        items = list(iter(obj))
    except TypeError:
        raise reject()
    if len(items) != 3:
        raise reject()
    out = []
    for it in items:
        if isinstance(it, (str, bytes, bytearray, memoryview)):
            raise reject()
        try:
            out.append(float(it))
        except (TypeError, ValueError):
            raise reject()
    if any((not np.isfinite(c)) or c < 0.0 for c in out):
        raise reject()
    return tuple(out)


def host_array(shape, dtype):
    """numpy array over page-locked memory from the library's pool (f3d_host_alloc): the device writes it with one DMA, no page
    faults, no staging copy.  The block returns to the pool when the last view of the array is garbage-collected.  Falls back
    to ordinary memory when the pool cannot serve (no CUDA device: the call that follows fails anyway)."""
    import weakref

    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    L = lib()
    ptr = L.f3d_host_alloc(max(nbytes, 1)) if nbytes >= (1 << 16) else None
    if not ptr:
        return np.zeros(shape, dtype)
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    weakref.finalize(buf, L.f3d_host_free, C.c_void_p(ptr))
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr


def cache_trim(device: int = -1) -> int:
    """Hands the device buffers the library keeps parked for reuse back to the CUDA driver (all devices by default);
    returns the bytes released.  Call it before giving the GPU to another allocator (torch, NCCL) in the same process."""
    return int(lib().f3d_cache_trim(int(device)))


def alloc_outputs(width, height, want_accum=False):
    H, W = int(height), int(width)
    arrays = dict(rgba=host_array((H, W, 4), np.uint8), albedo=host_array((H, W, 3), np.float32),
                  normal=host_array((H, W, 3), np.float32), depth=host_array((H, W), np.float32))
    o = TerrainOut()
    o.rgba = arrays["rgba"].ctypes.data_as(C.POINTER(C.c_uint8))
    o.albedo, o.normal, o.depth = _fp(arrays["albedo"]), _fp(arrays["normal"]), _fp(arrays["depth"])
    if want_accum:
        arrays["accum"] = host_array((H, W, 4), np.float32)
        o.accum = _fp(arrays["accum"])
    return o, arrays


def result_dict(o: TerrainOut, arrays: dict, sun_azimuth_deg, sun_elevation_deg) -> dict:
    """The dict of terrain_reference.rs:419-451 plus this backend's measurement keys."""
    res = dict(arrays)
    res.update(frames=int(o.frames), variance=float(o.variance), converged=bool(o.converged),
               peak_host_visible_bytes=int(o.peak_host_visible_bytes),
               minmax_pyramid_bytes=int(o.minmax_pyramid_bytes), gpu_resource_bytes=int(o.gpu_resource_bytes),
               sun_source="manual_angles", solar_azimuth_deg=float(sun_azimuth_deg),
               solar_elevation_deg=float(sun_elevation_deg),
               rays_primary=int(o.rays_primary), rays_shadow=int(o.rays_shadow), rays_ibl=int(o.rays_ibl),
               nodes_popped=int(o.nodes_popped), setup_ms=float(o.setup_ms), frames_ms=float(o.frames_ms),
               readback_ms=float(o.readback_ms), kernel_launches=int(o.kernel_launches))
    return res


def hybrid_render_terrain_reference(heightmap, width, height, cam, spacing=(1.0, 1.0), exaggeration=1.0,
                                    albedo=(0.6, 0.6, 0.6), sun_azimuth_deg=315.0, sun_elevation_deg=45.0,
                                    sun_intensity=2.5, env_map=None, env_intensity=0.35, mesh_vertices=None,
                                    mesh_indices=None, spp=1, max_frames=512, min_frames=32,
                                    variance_threshold=1e-3, seed=7, certificate=None, sun_color=None,
                                    cache=None, observer_latitude_deg=0.0, observer_longitude_deg=0.0,
                                    earth_model="ellipsoid", sphere_radius_m=6371008.8,
                                    refraction_model="bennett", refraction_k=0.13, pressure_mbar=1013.25,
                                    temperature_c=15.0, atmosphere=None, *, device=0, compat_512mib_gate=False,
                                    want_accum=False):
    """Native seam `_forge3d.hybrid_render_terrain_reference` (terrain_reference.rs:224-457) on the
    CUDA backend.  `cache` is accepted and ignored (the reference ignores it, :291).  A truthy `certificate` raises:
    the reference writes / assembles a render certificate for it (emit_certificate_for_kwarg) and certificates are out
    of scope here - a drop-in must not silently skip a file the caller asked for.  `atmosphere`
    is an AtmosphereLutHandle, a mapping or a settings object (extract_atmosphere_lut_handle, :46-210) and enables the
    AETHER post.  Argument checks follow the PyO3 seam's order: sun_color, then seed / albedo extraction (u32 / [f32; 3]:
    out-of-range or mis-sized values raise as PyO3's extraction does), earth / refraction models, then the atmosphere."""
    _ = cache
    if certificate:
        raise NotImplementedError("certificate= is not supported by forge3d_b200 (render certificates are out of scope); "
                                  "pass certificate=None/False")
    from .atmosphere import resolve_atmosphere

    sun_rgb = (1.0, 0.97, 0.92) if sun_color is None else extract_sun_color(sun_color)
    if not (0 <= int(seed) < 2 ** 32):
        raise OverflowError("can't convert negative int to unsigned" if int(seed) < 0 else "Python int too large to convert to C unsigned long")
    if len(tuple(albedo)) != 3:
        raise ValueError(f"expected a sequence of length 3 (got {len(tuple(albedo))})")
    if earth_model not in EARTH_MODELS:
        raise ValueError(f"unsupported earth_model {earth_model!r}")
    if refraction_model not in REFRACTION_MODELS:
        raise ValueError(f"unsupported refraction_model {refraction_model!r}")
    atmosphere = resolve_atmosphere(atmosphere)
    d, keep = make_desc(heightmap, width, height, cam, spacing=spacing, exaggeration=exaggeration, albedo=albedo,
                        sun_azimuth_deg=sun_azimuth_deg, sun_elevation_deg=sun_elevation_deg,
                        sun_intensity=sun_intensity, env_map=env_map, env_intensity=env_intensity,
                        mesh_vertices=mesh_vertices, mesh_indices=mesh_indices, spp=spp, max_frames=max_frames,
                        min_frames=min_frames, variance_threshold=variance_threshold, seed=seed, sun_color=sun_rgb,
                        observer_latitude_deg=observer_latitude_deg, observer_longitude_deg=observer_longitude_deg,
                        earth_model=earth_model, sphere_radius_m=sphere_radius_m, refraction_model=refraction_model,
                        refraction_k=refraction_k, pressure_mbar=pressure_mbar, temperature_c=temperature_c,
                        device=device, compat_512mib_gate=compat_512mib_gate, atmosphere=atmosphere)
    o, arrays = alloc_outputs(width, height, want_accum)
    check(lib().f3d_terrain_reference_render(C.byref(d), C.byref(o)))
    del keep
    return result_dict(o, arrays, sun_azimuth_deg, sun_elevation_deg)


def trace_rays(heights, spacing, origin_xz, exaggeration, rays, *, any_hit, apply_curvature,
               inv_two_r_prime=0.0, curvature_enabled=False, device=0, variant=0, want_nodes=False):
    """GPU terrain_trace over a ray batch (KAT seam, terrain_heightfield.rs:1646-1671).
    variant 0 = production traversal, 1 = literal restatement of the WGSL loop."""
    dem = np.ascontiguousarray(heights, dtype=np.float32)
    r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
    n = r.shape[0]
    hit = np.zeros(n, np.uint8)
    t = np.zeros(n, np.float32)
    nrm = np.zeros((n, 3), np.float32)
    sp = (C.c_float * 2)(float(spacing[0]), float(spacing[1]))
    og = (C.c_float * 2)(float(origin_xz[0]), float(origin_xz[1]))
    nodes = C.c_uint64()
    check(lib().f3d_trace_rays(_fp(dem), dem.shape[1], dem.shape[0], sp, og, float(exaggeration),
                               float(inv_two_r_prime), int(bool(curvature_enabled)), _fp(r), n,
                               int(bool(any_hit)), int(bool(apply_curvature)), int(device), int(variant),
                               hit.ctypes.data_as(C.POINTER(C.c_uint8)), _fp(t), _fp(nrm), C.byref(nodes)))
    if want_nodes:
        return hit.astype(bool), t, nrm, int(nodes.value)
    return hit.astype(bool), t, nrm


def build_minmax(heights, device=0):
    """GPU min-max pyramid -> (levels finest first as (h, w, 2) float32, cell_w, cell_h)."""
    L = lib()
    dem = np.ascontiguousarray(heights, dtype=np.float32)
    h, w = dem.shape
    dims = (C.c_uint32 * 64)()
    n = L.f3d_build_minmax(_fp(dem), w, h, int(device), dims, None, 0)
    if n < 0:
        check(-n)
    total = sum(dims[2 * i] * dims[2 * i + 1] * 2 for i in range(n))
    buf = np.zeros(total, np.float32)
    n = L.f3d_build_minmax(_fp(dem), w, h, int(device), dims, _fp(buf), total)
    if n < 0:
        check(-n)
    levels, off = [], 0
    for i in range(n):
        lw, lh = dims[2 * i], dims[2 * i + 1]
        levels.append(buf[off:off + lw * lh * 2].reshape(lh, lw, 2))
        off += lw * lh * 2
    return levels, w - 1, h - 1


def lbvh_build(vertices, indices, device=0):
    """GPU LBVH build of a triangle mesh (test seam f3d_lbvh_build) -> dict(morton, order, left, right, parent, nodes)."""
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
    n = t.shape[0]
    out = dict(morton=np.zeros(n, np.uint32), order=np.zeros(n, np.uint32), left=np.zeros(max(n - 1, 0), np.uint32),
               right=np.zeros(max(n - 1, 0), np.uint32), parent=np.zeros(2 * n - 1, np.uint32), nodes=np.zeros((2 * n - 1, 8), np.float32))
    u = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
    check(lib().f3d_lbvh_build(_fp(v), v.shape[0], u(t), n, int(device), u(out["morton"]), u(out["order"]), u(out["left"]),
                               u(out["right"]), u(out["parent"]), _fp(out["nodes"])))
    return out
