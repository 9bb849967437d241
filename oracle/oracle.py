"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE -- never imported by forge3d_b200).

Only tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs and the
cpu_baseline legs of the side benches bench.py runs (tools/bench_*.py; tools/wavefront_golden_pin.py
is the pin generator) may import this module -- always as the checker or the timed CPU baseline,
never as a producer of product output.  It exposes the oracle through the same keyword surface as the
reference's native seam (src/py_functions/path_tracing/terrain_reference.rs:224-288).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libf3d_oracle.so"

EARTH_MODELS = {"flat": 0, "sphere": 1, "ellipsoid": 2, "wgs84": 2}
REFRACTION_MODELS = {"none": 0, "bennett": 1, "saemundsson": 2, "effective_radius": 3}


class OracleError(RuntimeError):
    pass


class _Atmosphere(C.Structure):
    """f3do_atmosphere (oracle/f3d_oracle.h)."""
    _fields_ = [
        ("transmittance", C.POINTER(C.c_uint16)), ("scattering", C.POINTER(C.c_uint16)), ("aerial", C.POINTER(C.c_uint16)),
        ("transmittance_dims", C.c_uint32 * 2), ("scattering_dims", C.c_uint32 * 3),
        ("scattering_height", C.c_uint32), ("scattering_nu", C.c_uint32), ("aerial_dims", C.c_uint32 * 3),
        ("bottom_radius_m", C.c_float), ("top_radius_m", C.c_float), ("max_aerial_distance_m", C.c_float),
        ("ozone_du", C.c_float), ("mie_g", C.c_float), ("turbidity", C.c_float),
        ("rayleigh_scale_height_m", C.c_float), ("mie_scale_height_m", C.c_float), ("ground_albedo", C.c_float),
    ]


class _AetherView(C.Structure):
    """f3do_aether_view (oracle/f3d_oracle.h)."""
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("cam_origin", C.c_float * 3), ("cam_right", C.c_float * 3), ("cam_up", C.c_float * 3), ("cam_forward", C.c_float * 3),
        ("tan_half_fov", C.c_float), ("aspect", C.c_float), ("exposure", C.c_float),
        ("light_dir", C.c_float * 3), ("sun_intensity", C.c_float),
    ]


def make_atmosphere(handle):
    """LUT handle (anything with .config, .transmittance, .scattering, .aerial as RGBA16F uint16 arrays shaped
    (height, mu, 4), (height*nu, mu_sun, mu_view, 4), (height, mu_view, distance, 4)) -> (f3do_atmosphere, keepalive)."""
    t = np.ascontiguousarray(handle.transmittance, dtype=np.uint16)
    s = np.ascontiguousarray(handle.scattering, dtype=np.uint16)
    e = np.ascontiguousarray(handle.aerial, dtype=np.uint16)
    a = _Atmosphere()
    u16 = lambda arr: arr.ctypes.data_as(C.POINTER(C.c_uint16))
    a.transmittance, a.scattering, a.aerial = u16(t), u16(s), u16(e)
    a.transmittance_dims = (C.c_uint32 * 2)(t.shape[1], t.shape[0])
    a.scattering_dims = (C.c_uint32 * 3)(s.shape[2], s.shape[1], s.shape[0])
    a.aerial_dims = (C.c_uint32 * 3)(e.shape[2], e.shape[1], e.shape[0])
    cfg = handle.config
    a.scattering_height, a.scattering_nu = int(cfg.dimensions.scattering_height), int(cfg.dimensions.scattering_nu)
    for name in ("bottom_radius_m", "top_radius_m", "max_aerial_distance_m", "ozone_du", "mie_g", "turbidity",
                 "rayleigh_scale_height_m", "mie_scale_height_m", "ground_albedo"):
        setattr(a, name, float(getattr(cfg, name)))
    return a, [t, s, e, a]


class _Desc(C.Structure):
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
        ("max_frames", C.c_uint32), ("min_frames", C.c_uint32),
        ("variance_threshold", C.c_float), ("compat_512mib_gate", C.c_int32),
        ("atmosphere", C.POINTER(_Atmosphere)),
    ]


class _Out(C.Structure):
    _fields_ = [
        ("rgba", C.POINTER(C.c_uint8)), ("albedo", C.POINTER(C.c_float)),
        ("normal", C.POINTER(C.c_float)), ("depth", C.POINTER(C.c_float)),
        ("accum", C.POINTER(C.c_float)),
        ("frames", C.c_uint32), ("variance", C.c_float), ("converged", C.c_int32),
        ("minmax_pyramid_bytes", C.c_uint64),
        ("rays_primary", C.c_uint64), ("rays_shadow", C.c_uint64), ("rays_ibl", C.c_uint64),
        ("nodes_popped", C.c_uint64),
        ("setup_seconds", C.c_double), ("frames_seconds", C.c_double),
        ("prev_m_max", C.c_uint32), ("prev_weight_max", C.c_float), ("prev_w_sum_max", C.c_float),
        ("prev_dir_min", C.c_float * 3), ("prev_dir_max", C.c_float * 3),
    ]


_lib = None


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc, no FMA contraction)."""
    src_mtime = max((_HERE / n).stat().st_mtime for n in ("f3d_oracle.c", "f3d_aether_oracle.c", "f3d_smoke_oracle.c", "f3d_viewshed_oracle.c", "f3d_lbvh_oracle.c", "f3d_wavefront_oracle.c", "f3d_oracle.h", "Makefile"))
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src_mtime:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", str(_HERE), "-B", "libf3d_oracle.so"], check=True, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        L.f3do_render.argtypes = [C.POINTER(_Desc), C.POINTER(_Out)]
        L.f3do_render.restype = C.c_int
        L.f3do_last_error.restype = C.c_char_p
        L.f3do_set_threads.argtypes = [C.c_int]
        L.f3do_get_threads.restype = C.c_int
        L.f3do_build_minmax.argtypes = [C.POINTER(C.c_float), C.c_uint32, C.c_uint32,
                                        C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.c_uint64]
        L.f3do_build_minmax.restype = C.c_int
        L.f3do_trace_rays.argtypes = [C.POINTER(C.c_float), C.c_uint32, C.c_uint32,
                                      C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float,
                                      C.c_float, C.c_int, C.POINTER(C.c_float), C.c_uint64,
                                      C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_float),
                                      C.POINTER(C.c_float)]
        L.f3do_trace_rays.restype = C.c_int
        L.f3do_earth_curvature.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_double,
                                           C.c_double, C.c_double, C.c_double,
                                           C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
        L.f3do_earth_curvature.restype = C.c_int
        L.f3do_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.f3do_atan2.argtypes = [C.c_float, C.c_float]
        L.f3do_atan2.restype = C.c_float
        L.f3do_acos.argtypes = [C.c_float]
        L.f3do_acos.restype = C.c_float
        L.f3do_f32_to_f16.argtypes = [C.c_float]
        L.f3do_f32_to_f16.restype = C.c_uint16
        L.f3do_f16_to_f32.argtypes = [C.c_uint16]
        L.f3do_f16_to_f32.restype = C.c_float
        L.f3do_exp2.argtypes = [C.c_float]
        L.f3do_exp2.restype = C.c_float
        L.f3do_aether_post.argtypes = [C.POINTER(_Atmosphere), C.POINTER(_AetherView), C.POINTER(C.c_float),
                                       C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_uint16)]
        L.f3do_aether_post.restype = C.c_int
        L.f3do_aether_validate.argtypes = [C.POINTER(_Atmosphere)]
        L.f3do_aether_validate.restype = C.c_char_p
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def set_threads(n: int) -> None:
    lib().f3do_set_threads(int(n))


def get_threads() -> int:
    return int(lib().f3do_get_threads())


def render(heightmap, width, height, cam=None, *, spacing=(1.0, 1.0), exaggeration=1.0,
           albedo=(0.6, 0.6, 0.6), sun_azimuth_deg=315.0, sun_elevation_deg=45.0, sun_intensity=2.5,
           env_map=None, env_intensity=0.35, mesh_vertices=None, mesh_indices=None, spp=1,
           max_frames=512, min_frames=32, variance_threshold=1e-3, seed=7, sun_color=None,
           observer_latitude_deg=0.0, observer_longitude_deg=0.0, earth_model="ellipsoid",
           sphere_radius_m=6371008.8, refraction_model="bennett", refraction_k=0.13,
           pressure_mbar=1013.25, temperature_c=15.0, compat_512mib_gate=False, want_accum=False, atmosphere=None):
    """Oracle render with the native seam's keyword surface; returns the reference's result dict
    (terrain_reference.rs:437-450) plus ray counters."""
    L = lib()
    cam = dict(cam or {})
    dem = np.ascontiguousarray(heightmap, dtype=np.float32)
    if dem.ndim != 2:
        raise ValueError(f"heightmap must be 2D (H, W), got shape {dem.shape}")
    if earth_model not in EARTH_MODELS:
        raise ValueError(f"unsupported earth_model {earth_model!r}")
    if refraction_model not in REFRACTION_MODELS:
        raise ValueError(f"unsupported refraction_model {refraction_model!r}")
    d = _Desc()
    d.heights = _fp(dem)
    d.dem_h, d.dem_w = dem.shape
    d.spacing = (C.c_float * 2)(*map(float, spacing))
    d.exaggeration = float(exaggeration)
    d.albedo = (C.c_float * 3)(*map(float, albedo))
    d.cam_origin = (C.c_float * 3)(*map(float, cam.get("origin", (0.0, 50.0, 120.0))))
    d.cam_look_at = (C.c_float * 3)(*map(float, cam.get("look_at", (0.0, 0.0, 0.0))))
    d.cam_up = (C.c_float * 3)(*map(float, cam.get("up", (0.0, 1.0, 0.0))))
    d.fov_y_deg = float(cam.get("fov_y", 45.0))
    d.exposure = float(cam.get("exposure", 1.0))
    d.sun_az_deg = float(sun_azimuth_deg)
    d.sun_el_deg = float(sun_elevation_deg)
    d.sun_intensity = float(sun_intensity)
    d.sun_color = (C.c_float * 3)(*map(float, sun_color if sun_color is not None else (1.0, 0.97, 0.92)))
    d.observer_lat_deg = float(observer_latitude_deg)
    d.observer_lon_deg = float(observer_longitude_deg)
    d.earth_model = EARTH_MODELS[earth_model]
    d.sphere_radius_m = float(sphere_radius_m)
    d.refraction_model = REFRACTION_MODELS[refraction_model]
    d.refraction_k = float(refraction_k)
    d.pressure_mbar = float(pressure_mbar)
    d.temperature_c = float(temperature_c)
    keep = [dem]
    if env_map is not None:
        env = np.ascontiguousarray(env_map, dtype=np.float32)
        keep.append(env)
        d.env_rgb = _fp(env)
        d.env_h, d.env_w = env.shape[0], env.shape[1]
    d.env_intensity = float(env_intensity)
    if mesh_vertices is not None:
        mv = np.ascontiguousarray(mesh_vertices, dtype=np.float32)
        mi = np.ascontiguousarray(mesh_indices, dtype=np.uint32)
        keep += [mv, mi]
        d.mesh_xyz = _fp(mv)
        d.mesh_nverts = mv.shape[0]
        d.mesh_idx = mi.ctypes.data_as(C.POINTER(C.c_uint32))
        d.mesh_ntris = mi.shape[0]
    d.width, d.height, d.seed, d.spp = int(width), int(height), int(seed), int(spp)
    d.max_frames, d.min_frames = int(max_frames), int(min_frames)
    d.variance_threshold = float(variance_threshold)
    d.compat_512mib_gate = int(bool(compat_512mib_gate))
    if atmosphere is not None:   # a resolved LUT handle (AETHER aerial-perspective post)
        atm, atm_keep = make_atmosphere(atmosphere)
        keep += atm_keep
        d.atmosphere = C.pointer(atm)

    H, W = int(height), int(width)
    rgba = np.zeros((H, W, 4), np.uint8)
    alb = np.zeros((H, W, 3), np.float32)
    nrm = np.zeros((H, W, 3), np.float32)
    dep = np.zeros((H, W), np.float32)
    acc = np.zeros((H, W, 4), np.float32) if want_accum else None
    o = _Out()
    o.rgba = rgba.ctypes.data_as(C.POINTER(C.c_uint8))
    o.albedo, o.normal, o.depth = _fp(alb), _fp(nrm), _fp(dep)
    if acc is not None:
        o.accum = _fp(acc)
    rc = L.f3do_render(C.byref(d), C.byref(o))
    if rc != 0:
        raise OracleError(L.f3do_last_error().decode("utf-8", "replace"))
    res = dict(rgba=rgba, albedo=alb, normal=nrm, depth=dep, frames=int(o.frames),
               variance=float(o.variance), converged=bool(o.converged),
               minmax_pyramid_bytes=int(o.minmax_pyramid_bytes),
               rays_primary=int(o.rays_primary), rays_shadow=int(o.rays_shadow),
               rays_ibl=int(o.rays_ibl), nodes_popped=int(o.nodes_popped),
               setup_seconds=float(o.setup_seconds), frames_seconds=float(o.frames_seconds),
               prev_m_max=int(o.prev_m_max), prev_weight_max=float(o.prev_weight_max),
               prev_w_sum_max=float(o.prev_w_sum_max), prev_dir_min=tuple(o.prev_dir_min), prev_dir_max=tuple(o.prev_dir_max),
               sun_source="manual_angles", solar_azimuth_deg=float(sun_azimuth_deg),
               solar_elevation_deg=float(sun_elevation_deg))
    if acc is not None:
        res["accum"] = acc
    return res


def aether_post(handle, accum, depth, visibility, *, cam_origin, cam_right, cam_up, cam_forward, tan_half_fov, aspect,
                exposure, light_dir, sun_intensity):
    """prometheus_aerial.wgsl `main` over given buffers -> (H, W, 4) RGBA16F bit patterns (uint16)."""
    L = lib()
    acc = np.ascontiguousarray(accum, np.float32)
    H, W = acc.shape[:2]
    dep = np.ascontiguousarray(depth, np.float32).reshape(H, W)
    vis = np.ascontiguousarray(visibility, np.uint8).reshape(H, W)
    atm, keep = make_atmosphere(handle)
    v = _AetherView()
    v.width, v.height = W, H
    v.cam_origin = (C.c_float * 3)(*map(float, cam_origin))
    v.cam_right = (C.c_float * 3)(*map(float, cam_right))
    v.cam_up = (C.c_float * 3)(*map(float, cam_up))
    v.cam_forward = (C.c_float * 3)(*map(float, cam_forward))
    v.tan_half_fov, v.aspect, v.exposure = float(tan_half_fov), float(aspect), float(exposure)
    v.light_dir = (C.c_float * 3)(*map(float, light_dir))
    v.sun_intensity = float(sun_intensity)
    out = np.zeros((H, W, 4), np.uint16)
    bad = L.f3do_aether_validate(C.byref(atm))
    if bad:
        raise OracleError(bad.decode())
    rc = L.f3do_aether_post(C.byref(atm), C.byref(v), _fp(acc), _fp(dep), vis.ctypes.data_as(C.POINTER(C.c_uint8)),
                            out.ctypes.data_as(C.POINTER(C.c_uint16)))
    if rc != 0:
        raise OracleError("f3do_aether_post failed")
    del keep
    return out


class _SmokeVolume(C.Structure):
    """f3do_smoke_volume (oracle/f3d_oracle.h)."""
    _fields_ = [
        ("dims", C.c_uint32 * 3), ("voxel_size", C.c_float * 3), ("origin", C.c_float * 3),
        ("density", C.POINTER(C.c_float)), ("temperature", C.POINTER(C.c_float)), ("soot", C.POINTER(C.c_float)),
        ("humidity", C.POINTER(C.c_float)), ("emission_rate", C.POINTER(C.c_float)), ("particle_age", C.POINTER(C.c_float)),
        ("frame_index", C.c_uint64),
    ]


class _SmokeSettings(C.Structure):
    """f3do_smoke_settings (oracle/f3d_oracle.h)."""
    _fields_ = [
        ("density_scale", C.c_float), ("extinction", C.c_float), ("scattering", C.c_float), ("absorption", C.c_float),
        ("phase_g", C.c_float), ("step_size", C.c_float), ("max_steps", C.c_uint32), ("self_shadow", C.c_int32),
        ("shadow_steps", C.c_uint32), ("shadow_step_size", C.c_float), ("jitter_strength", C.c_float), ("exposure", C.c_float),
        ("thin_color", C.c_float * 3), ("dense_color", C.c_float * 3), ("soot_absorption", C.c_float), ("fire_glow", C.c_float),
    ]


_SMOKE_FIELDS = ("density", "temperature", "soot", "humidity", "emission_rate", "particle_age")
_SMOKE_SCALARS = ("density_scale", "extinction", "scattering", "absorption", "phase_g", "step_size", "shadow_step_size",
                  "jitter_strength", "exposure", "soot_absorption", "fire_glow")


def _smoke_structs(domain, settings):
    """domain: .dims (x, y, z), .voxel_size, .origin, .frame_index and the six (z, y, x) float32 fields (None = zero);
    settings: an object with SmokeRenderSettings' attributes."""
    v = _SmokeVolume()
    v.dims = (C.c_uint32 * 3)(*map(int, domain.dims))
    v.voxel_size = (C.c_float * 3)(*map(float, domain.voxel_size))
    v.origin = (C.c_float * 3)(*map(float, domain.origin))
    keep = []
    for name in _SMOKE_FIELDS:
        arr = getattr(domain, name, None)
        if arr is None:
            continue
        a = np.ascontiguousarray(arr, dtype=np.float32)
        keep.append(a)
        setattr(v, name, _fp(a))
    v.frame_index = int(getattr(domain, "frame_index", 0))
    s = _SmokeSettings()
    for name in _SMOKE_SCALARS:
        setattr(s, name, float(getattr(settings, name)))
    s.max_steps, s.shadow_steps, s.self_shadow = int(settings.max_steps), int(settings.shadow_steps), int(bool(settings.self_shadow))
    s.thin_color = (C.c_float * 3)(*map(float, settings.thin_color))
    s.dense_color = (C.c_float * 3)(*map(float, settings.dense_color))
    return v, s, keep


def _smoke_lib():
    L = lib()
    if not getattr(L, "_smoke_ready", False):
        f3 = C.c_float * 3
        u8p = C.POINTER(C.c_uint8)
        L.f3do_smoke_raymarch_rgba.argtypes = [C.POINTER(_SmokeVolume), C.POINTER(_SmokeSettings), C.c_uint32, C.c_uint32, f3, f3, f3,
                                               C.c_float, f3, u8p]
        L.f3do_smoke_raymarch_over_rgba.argtypes = [C.POINTER(_SmokeVolume), C.POINTER(_SmokeSettings), C.c_uint32, C.c_uint32, f3, f3, f3,
                                                    C.c_float, f3, C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_uint8)]
        L.f3do_smoke_raymarch_projection_rgba.argtypes = [C.POINTER(_SmokeVolume), C.POINTER(_SmokeSettings), C.c_uint32,
                                                          C.c_uint32, f3, f3, u8p]
        L.f3do_smoke_sun_transmittance.argtypes = [C.POINTER(_SmokeVolume), C.POINTER(_SmokeSettings), f3, f3, C.c_float, C.c_uint32]
        L.f3do_smoke_sun_transmittance.restype = C.c_float
        L.f3do_smoke_last_error.restype = C.c_char_p
        L._smoke_ready = True
    return L


def smoke_raymarch_rgba(domain, settings, width, height, camera_pos, target, up=(0.0, 1.0, 0.0), fovy_deg=45.0,
                        sun_direction=(0.4, 0.8, -0.2)):
    """SmokeVolume::raymarch_rgba (src/smoke/render.rs:6-101) -> (height, width, 4) uint8."""
    L = _smoke_lib()
    v, s, keep = _smoke_structs(domain, settings)
    out = np.zeros((int(height), int(width), 4), np.uint8)
    f3 = C.c_float * 3
    rc = L.f3do_smoke_raymarch_rgba(C.byref(v), C.byref(s), int(width), int(height), f3(*map(float, camera_pos)),
                                    f3(*map(float, target)), f3(*map(float, up)), float(fovy_deg), f3(*map(float, sun_direction)),
                                    out.ctypes.data_as(C.POINTER(C.c_uint8)))
    if rc != 0:
        raise OracleError(L.f3do_smoke_last_error().decode())
    del keep
    return out


def composite_over_rgba(bottom, top):
    """_alpha_composite_rgba (python/forge3d/map_scene.py:1588-1604) on two (H, W, 4) uint8 images."""
    L = _smoke_lib()
    b, t = np.ascontiguousarray(bottom, np.uint8), np.ascontiguousarray(top, np.uint8)
    if b.shape != t.shape or b.shape[-1] != 4:
        raise OracleError("composite_over_rgba needs two RGBA images of the same shape")
    out = np.empty_like(b)
    L.f3do_composite_over_rgba.argtypes = [C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint8)]
    L.f3do_composite_over_rgba.restype = None
    L.f3do_composite_over_rgba(b.ctypes.data_as(C.POINTER(C.c_uint8)), t.ctypes.data_as(C.POINTER(C.c_uint8)), b.size // 4,
                               out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def smoke_raymarch_over_rgba(domain, settings, width, height, camera_pos, target, base_rgba, base_depth=None, up=(0.0, 1.0, 0.0),
                             fovy_deg=45.0, sun_direction=(0.4, 0.8, -0.2)):
    """The smoke layer (render.rs:6-101) composited over a terrain frame with _alpha_composite_rgba
    (python/forge3d/map_scene.py:1588-1604); base_depth (optional) ends each march at the terrain."""
    L = _smoke_lib()
    v, s, keep = _smoke_structs(domain, settings)
    f3 = C.c_float * 3
    base = np.ascontiguousarray(base_rgba, np.uint8)
    depth = None if base_depth is None else np.ascontiguousarray(base_depth, np.float32)
    out = np.zeros((int(height), int(width), 4), np.uint8)
    rc = L.f3do_smoke_raymarch_over_rgba(C.byref(v), C.byref(s), int(width), int(height), f3(*map(float, camera_pos)),
                                         f3(*map(float, target)), f3(*map(float, up)), float(fovy_deg), f3(*map(float, sun_direction)),
                                         base.ctypes.data_as(C.POINTER(C.c_uint8)),
                                         depth.ctypes.data_as(C.POINTER(C.c_float)) if depth is not None else None,
                                         out.ctypes.data_as(C.POINTER(C.c_uint8)))
    if rc != 0:
        raise OracleError(L.f3do_smoke_last_error().decode())
    del keep
    return out


def smoke_raymarch_projection_rgba(domain, settings, width, height, view_direction=(0.0, -1.0, 0.0), sun_direction=(0.4, 0.8, -0.2)):
    """SmokeVolume::raymarch_projection_rgba (src/smoke/render.rs:103-175) -> (height, width, 4) uint8."""
    L = _smoke_lib()
    v, s, keep = _smoke_structs(domain, settings)
    out = np.zeros((int(height), int(width), 4), np.uint8)
    f3 = C.c_float * 3
    rc = L.f3do_smoke_raymarch_projection_rgba(C.byref(v), C.byref(s), int(width), int(height), f3(*map(float, view_direction)),
                                               f3(*map(float, sun_direction)), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    if rc != 0:
        raise OracleError(L.f3do_smoke_last_error().decode())
    del keep
    return out


def smoke_sun_transmittance(domain, settings, start, sun_dir, step, steps):
    """SmokeVolume::sun_transmittance (src/smoke/render.rs:278-316)."""
    L = _smoke_lib()
    v, s, keep = _smoke_structs(domain, settings)
    f3 = C.c_float * 3
    out = float(L.f3do_smoke_sun_transmittance(C.byref(v), C.byref(s), f3(*map(float, start)), f3(*map(float, sun_dir)), float(step), int(steps)))
    del keep
    return out


class _ViewshedOptions(C.Structure):
    """f3do_viewshed_options (oracle/f3d_oracle.h)."""
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32)] + [(n, C.c_float) for n in (
        "observer_x", "observer_y", "observer_height_m", "target_height_m", "max_distance_m", "observer_latitude_rad",
        "observer_longitude_rad", "left_unwrapped_deg", "top_deg", "longitude_step_deg", "latitude_step_deg",
        "geodesic_sphere_radius_m")] + [("physics", C.c_float * 4)]


def _viewshed_options(opts: dict):
    """opts: the ViewshedOptions fields + earth_model / refraction_model names and their parameters (see
    forge3d_b200.viewshed.make_options)."""
    L = lib()
    o = _ViewshedOptions()
    o.width, o.height = int(opts["width"]), int(opts["height"])
    for name, _ in _ViewshedOptions._fields_[2:-1]:
        setattr(o, name, float(opts[name]))
    phys = (C.c_float * 4)()
    L.f3do_viewshed_physics.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_float * 4]
    rc = L.f3do_viewshed_physics(EARTH_MODELS[opts["earth_model"]], float(opts.get("earth_latitude_deg", 0.0)),
                                 float(opts.get("sphere_radius_m", 6371008.8)), REFRACTION_MODELS[opts["refraction_model"]],
                                 float(opts.get("refraction_k", 0.13)), float(opts.get("pressure_mbar", 1013.25)),
                                 float(opts.get("temperature_c", 15.0)), phys)
    if rc != 0:
        raise OracleError({1: "flat earth only supports refraction_model='none'", 3: "refraction k must be finite and less than 1",
                           4: "sphere radius must be finite and positive"}.get(rc, f"invalid physics ({rc})"))
    o.physics = phys
    return o


def viewshed(heights, positions_m, opts: dict):
    """compute_viewshed (viewshed.rs:161-347) -> dict(visibility bool, curvature_drop_m, refraction_gain_m, horizon_distance_m)."""
    L = lib()
    dem = np.ascontiguousarray(heights, np.float32)
    pos = np.ascontiguousarray(positions_m, np.float32).reshape(dem.shape + (2,))
    o = _viewshed_options(opts)
    vis = np.zeros(dem.shape, np.uint8)
    drop, gain, horizon = (np.zeros(dem.shape, np.float32) for _ in range(3))
    L.f3do_viewshed.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(_ViewshedOptions), C.POINTER(C.c_uint8),
                                C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    if L.f3do_viewshed(_fp(dem), _fp(pos), C.byref(o), vis.ctypes.data_as(C.POINTER(C.c_uint8)), _fp(drop), _fp(gain), _fp(horizon)) != 0:
        raise OracleError("f3do_viewshed failed")
    if (vis > 1).any():
        raise OracleError("viewshed geodesic leaves the DEM footprint")
    return dict(visibility=vis.astype(bool), curvature_drop_m=drop, refraction_gain_m=gain, horizon_distance_m=horizon)


def shadow_mask(heights, inputs, opts: dict):
    """compute_shadow_mask (viewshed.rs:396-570) -> bool (H, W), True = the sun is visible from the cell."""
    L = lib()
    dem = np.ascontiguousarray(heights, np.float32)
    inp = np.ascontiguousarray(inputs, np.float32).reshape(dem.shape + (4,))
    o = _viewshed_options(opts)
    lit = np.zeros(dem.shape, np.uint8)
    L.f3do_shadow_mask.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(_ViewshedOptions), C.POINTER(C.c_uint8)]
    if L.f3do_shadow_mask(_fp(dem), _fp(inp), C.byref(o), lit.ctypes.data_as(C.POINTER(C.c_uint8))) != 0:
        raise OracleError("f3do_shadow_mask failed")
    return lit.astype(bool)


def lbvh_build(vertices, indices, literal_split=False, pad_boxes=True):
    """LBVH build restatement -> dict(morton, order, left, right, parent, nodes); see oracle/f3d_lbvh_oracle.c."""
    L = lib()
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
    n = t.shape[0]
    out = dict(morton=np.zeros(n, np.uint32), order=np.zeros(n, np.uint32), left=np.zeros(max(n - 1, 0), np.uint32),
               right=np.zeros(max(n - 1, 0), np.uint32), parent=np.zeros(2 * n - 1, np.uint32), nodes=np.zeros((2 * n - 1, 8), np.float32))
    u = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
    u32p = C.POINTER(C.c_uint32)
    L.f3do_lbvh_build.argtypes = [C.POINTER(C.c_float), C.c_uint32, u32p, C.c_uint32, C.c_int, C.c_int, u32p, u32p, u32p, u32p, u32p,
                                  C.POINTER(C.c_float)]
    if L.f3do_lbvh_build(_fp(v), v.shape[0], u(t), n, int(bool(literal_split)), int(bool(pad_boxes)), u(out["morton"]), u(out["order"]),
                         u(out["left"]), u(out["right"]), u(out["parent"]), _fp(out["nodes"])) != 0:
        raise OracleError("f3do_lbvh_build failed")
    return out


class _WavefrontScene(C.Structure):
    """f3do_wavefront_scene (oracle/f3d_oracle.h)."""
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


def wavefront_render(scene, width, height, spp_frames, first_frame=0, num_frames=None, accum=None, resolve=True):
    """Wavefront path tracer restatement (oracle/f3d_wavefront_oracle.c).  `scene` carries the packed arrays of
    forge3d_b200.wavefront.WavefrontScene.normalized().  Returns dict(accum, hdr, rgba8, rays, max_rays_per_frame, min_iterations);
    pass `accum` back in to continue an accumulation in slices of frames."""
    L = lib()
    fpt, upt = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    cs = _WavefrontScene()
    for name in ("cam_origin", "cam_forward", "cam_right", "cam_up"):
        getattr(cs, name)[:] = [float(x) for x in getattr(scene, name)]
    cs.fov_y_rad, cs.exposure, cs.seed_hi, cs.seed_lo = scene.fov_y_rad, scene.exposure, scene.seed_hi, scene.seed_lo
    cs.environment[:] = [float(x) for x in scene.environment]
    arrs = {}
    for name, cnt, t, dt, shape in (("spheres", "nspheres", fpt, np.float32, (-1, 20)), ("dir_lights", "ndir", fpt, np.float32, (-1, 8)),
                                    ("area_lights", "narea", fpt, np.float32, (-1, 12)), ("importance", "nimportance", fpt, np.float32, (-1,)),
                                    ("mesh_xyz", "mesh_nverts", fpt, np.float32, (-1, 3)), ("mesh_idx", "mesh_ntris", upt, np.uint32, (-1, 3)),
                                    ("instances", "ninstances", fpt, np.float32, (-1, 36))):
        a = np.ascontiguousarray(np.asarray(getattr(scene, name), dtype=dt).reshape(shape))
        arrs[name] = a
        setattr(cs, name, a.ctypes.data_as(t) if a.size else t())
        setattr(cs, cnt, a.shape[0])
    if num_frames is None:
        num_frames = spp_frames - first_frame
    if accum is None:
        accum = np.zeros((height, width, 4), np.float32)
    hdr = np.zeros((height, width, 4), np.float32) if resolve else None
    rgba = np.zeros((height, width, 4), np.uint8) if resolve else None
    stats = (C.c_uint64 * 19)()
    L.f3do_wavefront_render.argtypes = [C.POINTER(_WavefrontScene), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, fpt, fpt,
                                        C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)]
    L.f3do_wavefront_last_error.restype = C.c_char_p
    rc = L.f3do_wavefront_render(C.byref(cs), width, height, spp_frames, first_frame, num_frames, accum.ctypes.data_as(fpt),
                                 hdr.ctypes.data_as(fpt) if resolve else fpt(), rgba.ctypes.data_as(C.POINTER(C.c_uint8)) if resolve else None,
                                 stats)
    if rc != 0:
        raise OracleError(L.f3do_wavefront_last_error().decode())
    return dict(accum=accum, hdr=hdr, rgba8=rgba, rays=int(stats[0]), max_rays_per_frame=int(stats[1]), min_iterations=int(stats[2]),
                rays_per_depth=[int(stats[3 + k]) for k in range(16)])


def log2(x: float) -> float:
    L = lib()
    L.f3do_log2.restype = C.c_float
    L.f3do_log2.argtypes = [C.c_float]
    return float(L.f3do_log2(float(x)))


def exp2(x: float) -> float:
    return float(lib().f3do_exp2(float(x)))


def build_minmax(heights):
    """build_minmax_mips -> (list of (h, w, 2) float32 levels finest first, cell_w, cell_h)."""
    L = lib()
    dem = np.ascontiguousarray(heights, dtype=np.float32)
    h, w = dem.shape
    dims = (C.c_uint32 * 64)()
    n = L.f3do_build_minmax(_fp(dem), w, h, dims, None, 0)
    if n < 0:
        raise OracleError(L.f3do_last_error().decode())
    total = sum(dims[2 * i] * dims[2 * i + 1] * 2 for i in range(n))
    buf = np.zeros(total, np.float32)
    n = L.f3do_build_minmax(_fp(dem), w, h, dims, _fp(buf), total)
    levels, off = [], 0
    for i in range(n):
        lw, lh = dims[2 * i], dims[2 * i + 1]
        levels.append(buf[off:off + lw * lh * 2].reshape(lh, lw, 2))
        off += lw * lh * 2
    return levels, w - 1, h - 1


def trace_rays(heights, spacing, origin_xz, exaggeration, rays, *, any_hit, apply_curvature,
               inv_two_r_prime=0.0, curvature_enabled=False):
    """terrain_trace over a ray batch; rays is (n, 8): origin.xyz, tmin, direction.xyz, tmax."""
    L = lib()
    dem = np.ascontiguousarray(heights, dtype=np.float32)
    r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
    n = r.shape[0]
    hit = np.zeros(n, np.uint8)
    t = np.zeros(n, np.float32)
    nrm = np.zeros((n, 3), np.float32)
    sp = (C.c_float * 2)(*map(float, spacing))
    og = (C.c_float * 2)(*map(float, origin_xz))
    rc = L.f3do_trace_rays(_fp(dem), dem.shape[1], dem.shape[0], sp, og, float(exaggeration),
                           float(inv_two_r_prime), int(bool(curvature_enabled)), _fp(r), n,
                           int(bool(any_hit)), int(bool(apply_curvature)),
                           hit.ctypes.data_as(C.POINTER(C.c_uint8)), _fp(t), _fp(nrm))
    if rc != 0:
        raise OracleError(L.f3do_last_error().decode())
    return hit.astype(bool), t, nrm


def earth_curvature(earth_model="ellipsoid", lat_deg=0.0, sphere_radius_m=6371008.8,
                    refraction_model="bennett", k=0.13, pressure_mbar=1013.25, temperature_c=15.0,
                    azimuth_deg=0.0):
    L = lib()
    inv = C.c_float()
    en = C.c_uint32()
    rc = L.f3do_earth_curvature(EARTH_MODELS[earth_model], lat_deg, sphere_radius_m,
                                REFRACTION_MODELS[refraction_model], k, pressure_mbar, temperature_c,
                                azimuth_deg, C.byref(inv), C.byref(en))
    if rc != 0:
        raise OracleError(L.f3do_last_error().decode())
    return float(inv.value), bool(en.value)
