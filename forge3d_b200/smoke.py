"""Smoke volumes rendered on the GPU: `SmokeDomain.render_rgba` / `render_projection_rgba`.

Host-side stand-in (Python, where the reference is Rust + PyO3) for the rendering half of `forge3d.smoke`
(/root/reference/python/forge3d/smoke.py, native classes in src/smoke/py.rs): `SmokeRenderSettings` (:266-341),
`SmokeEmitter` (:17-101), `SmokeDomain` (:343-628) with `from_density`, `set_density`, `add_emitter`, the numpy accessors
and the two render calls, which go through the C ABI (`f3d_smoke_*`, include/forge3d_b200.h) to the CUDA march kernel
instead of the reference's single-threaded CPU loops (src/smoke/render.rs).  Same signatures, defaults, validation text
and u8 RGBA result.  The fluid solver (`step`, src/smoke/sim.rs:47-899, a CPU pressure-projection code) is out of
scope: volumes come from `from_density`, the field setters, `add_emitter`, or an external simulation.
There is no CPU fallback: without the CUDA library / a device the render calls raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import _native

_F32_MAX = float(np.finfo(np.float32).max)


def _tuple3(value, name) -> tuple:
    out = tuple(float(v) for v in value)
    if len(out) != 3:
        raise ValueError(f"{name} must have exactly three components")
    return out


@dataclass
class SmokeRenderSettings:
    """SmokeRenderSettings (src/smoke/types.rs:225-266; constructor src/smoke/py.rs:273-332)."""
    density_scale: float = 1.0
    extinction: float = 2.6
    scattering: float = 0.85
    absorption: float = 0.45
    phase_g: float = 0.24
    step_size: float = 0.0
    max_steps: int = 256
    self_shadow: bool = True
    shadow_steps: int = 20
    shadow_step_size: float = 0.0
    jitter_strength: float = 0.5
    exposure: float = 1.0
    thin_color: Sequence[float] = (0.50, 0.54, 0.58)
    dense_color: Sequence[float] = (0.93, 0.91, 0.82)
    soot_absorption: float = 0.22
    fire_glow: float = 0.35

    def __post_init__(self) -> None:
        self.validate()

    def validate(self) -> None:
        """SmokeRenderSettings::validate, types.rs:268-317 (the native constructor raises ValueError with this text)."""
        for name in ("density_scale", "extinction", "scattering", "absorption", "phase_g", "step_size", "shadow_step_size",
                     "jitter_strength", "exposure", "soot_absorption", "fire_glow"):
            if not np.isfinite(np.float32(getattr(self, name))):
                raise ValueError(f"{name} must be finite")
        if self.density_scale < 0.0 or self.extinction < 0.0 or self.scattering < 0.0:
            raise ValueError("density_scale, extinction, and scattering must be >= 0")
        if self.absorption < 0.0 or self.soot_absorption < 0.0 or self.fire_glow < 0.0:
            raise ValueError("absorption, soot_absorption, and fire_glow must be >= 0")
        if not np.float32(-0.99) <= np.float32(self.phase_g) <= np.float32(0.99):
            raise ValueError("phase_g must be in [-0.99, 0.99]")
        if self.step_size < 0.0 or self.shadow_step_size < 0.0:
            raise ValueError("step sizes must be >= 0")
        if int(self.max_steps) == 0 or int(self.shadow_steps) == 0:
            raise ValueError("max_steps and shadow_steps must be >= 1")
        if not 0.0 <= self.jitter_strength <= 1.0:
            raise ValueError("jitter_strength must be in [0, 1]")
        for name in ("thin_color", "dense_color"):
            for axis, value in enumerate(_tuple3(getattr(self, name), name)):
                if not np.isfinite(value) or value < 0.0:
                    raise ValueError(f"{name}[{axis}] must be finite and >= 0")

    def _native(self) -> "_native.SmokeSettings":
        s = _native.SmokeSettings()
        for name in ("density_scale", "extinction", "scattering", "absorption", "phase_g", "step_size", "shadow_step_size",
                     "jitter_strength", "exposure", "soot_absorption", "fire_glow"):
            setattr(s, name, float(getattr(self, name)))
        s.max_steps, s.shadow_steps, s.self_shadow = int(self.max_steps), int(self.shadow_steps), int(bool(self.self_shadow))
        s.thin_color = (C.c_float * 3)(*_tuple3(self.thin_color, "thin_color"))
        s.dense_color = (C.c_float * 3)(*_tuple3(self.dense_color, "dense_color"))
        return s


@dataclass
class SmokeEmitter:
    """SmokeEmitter (src/smoke/types.rs:69-99)."""
    center: Sequence[float] = (0.0, 0.0, 0.0)
    radius: float = 1.0
    density_rate: float = 1.0
    temperature_rate: float = 1.0
    fuel_rate: float = 0.0
    soot_rate: float = 0.2
    humidity_rate: float = 0.0
    emission_rate: float = 1.0
    velocity: Sequence[float] = (0.0, 1.0, 0.0)
    start_time: float = 0.0
    end_time: float = _F32_MAX

    def __post_init__(self) -> None:
        self.center = _tuple3(self.center, "center")
        self.velocity = _tuple3(self.velocity, "velocity")
        if not all(np.isfinite(c) for c in self.center):
            raise ValueError("center must be finite")
        if not np.isfinite(self.radius) or self.radius <= 0.0:
            raise ValueError("radius must be finite and > 0")
        if self.end_time < self.start_time:
            raise ValueError("end_time must be >= start_time")


class SmokeDomain:
    """A dense smoke volume (six scalar fields, z-major (nz, ny, nx) float32 arrays) resident on one CUDA device."""

    _FIELDS = ("density", "temperature", "soot", "humidity", "emission_rate", "particle_age")

    def __init__(self, dims, voxel_size=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), brick_size=(16, 16, 16),
                 sparse_threshold=1.0e-5, *, device=0):
        self._dims = tuple(int(v) for v in dims)
        if len(self._dims) != 3:
            raise ValueError("dims must be (x, y, z)")
        for axis, value in enumerate(self._dims):   # SmokeDomainConfig::validate, types.rs:29-55
            if value < 2:
                raise ValueError(f"dims[{axis}] must be >= 2")
        self._voxel_size = _tuple3(voxel_size, "voxel_size")
        self._origin = _tuple3(origin, "origin")
        for axis, value in enumerate(self._voxel_size):
            if not np.isfinite(value) or value <= 0.0:
                raise ValueError(f"voxel_size[{axis}] must be finite and > 0")
        for axis, value in enumerate(self._origin):
            if not np.isfinite(value):
                raise ValueError(f"origin[{axis}] must be finite")
        self.brick_size = tuple(int(v) for v in brick_size)
        self.sparse_threshold = float(sparse_threshold)
        shape = (self._dims[2], self._dims[1], self._dims[0])
        self.density = np.zeros(shape, np.float32)
        self.temperature = np.zeros(shape, np.float32)
        self.soot = np.zeros(shape, np.float32)
        self.humidity = np.zeros(shape, np.float32)
        self.emission_rate = np.zeros(shape, np.float32)
        self.particle_age = np.full(shape, -1.0, np.float32)          # SmokeVolume::new, types.rs:348-365
        self.time_seconds = 0.0
        self.frame_index = 0
        self.device = int(device)
        self._handle = None

    # -- construction / state (src/smoke/py.rs:374-498) ---------------------------------------------------------
    @classmethod
    def from_density(cls, density, voxel_size=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), *, device=0) -> "SmokeDomain":
        d = np.asarray(density)
        if d.ndim != 3:
            raise ValueError("density must be a 3D (z, y, x) array")
        dom = cls((d.shape[2], d.shape[1], d.shape[0]), voxel_size, origin, device=device)
        dom.set_density(d)
        return dom

    def _checked(self, value, name) -> np.ndarray:
        a = np.ascontiguousarray(value, dtype=np.float32)
        if a.shape != self.density.shape:
            raise ValueError(f"{name} length {a.size} does not match voxel_count {self.density.size}")
        if not np.isfinite(a).all():
            raise ValueError(f"{name} contains non-finite values")
        return a

    def set_density(self, density) -> None:
        """SmokeVolume::set_density (types.rs:490-510): also resets the particle age (0 where there is smoke, else -1)."""
        self.density = self._checked(density, "density").copy()
        self.particle_age = np.where(self.density > np.float32(self.sparse_threshold), np.float32(0.0), np.float32(-1.0)).astype(np.float32)
        self._invalidate()

    def set_field(self, name: str, value) -> None:
        """Sets one of temperature / soot / humidity / emission_rate / particle_age (the reference fills these from its
        solver; external simulations hand them over here)."""
        if name not in self._FIELDS or name == "density":
            raise ValueError(f"unknown smoke field {name!r}")
        setattr(self, name, self._checked(value, name).copy())
        self._invalidate()

    def add_emitter(self, emitter: SmokeEmitter, dt: float) -> None:
        """SmokeVolume::add_emitter (src/smoke/sim.rs:7-45), f32 arithmetic, without the velocity / fuel fields."""
        dt = np.float32(dt)
        if not np.isfinite(dt) or dt <= 0.0:
            raise ValueError("dt must be finite and > 0")
        f = np.float32
        nx, ny, nz = self._dims
        wx = f(self._origin[0]) + (np.arange(nx, dtype=np.float32) + f(0.5)) * f(self._voxel_size[0])
        wy = f(self._origin[1]) + (np.arange(ny, dtype=np.float32) + f(0.5)) * f(self._voxel_size[1])
        wz = f(self._origin[2]) + (np.arange(nz, dtype=np.float32) + f(0.5)) * f(self._voxel_size[2])
        dx = (wx - f(emitter.center[0]))[None, None, :]
        dy = (wy - f(emitter.center[1]))[None, :, None]
        dz = (wz - f(emitter.center[2]))[:, None, None]
        dist = np.sqrt((dx * dx + dy * dy) + dz * dz).astype(np.float32)      # glam distance = sqrt(dot(d, d))
        radius = max(f(emitter.radius), f(1.0e-6))
        inside = ~(dist > radius)
        t = np.clip(dist / max(radius - f(0.0), f(1.0e-6)), f(0.0), f(1.0)).astype(np.float32)   # smoothstep(0, radius, d)
        falloff = (f(1.0) - t * t * (f(3.0) - f(2.0) * t)).astype(np.float32)
        amount = (dt * falloff).astype(np.float32)
        for name, rate in (("density", emitter.density_rate), ("temperature", emitter.temperature_rate),
                           ("soot", emitter.soot_rate), ("humidity", emitter.humidity_rate)):
            field = getattr(self, name)
            field[inside] = np.maximum(field[inside] + f(rate) * amount[inside], f(0.0))
        self.emission_rate[inside] += f(emitter.emission_rate) * falloff[inside]
        self.particle_age[inside] = 0.0
        self._invalidate()

    def step(self, settings=None, emitters=None):
        raise NotImplementedError("the smoke fluid solver (src/smoke/sim.rs) is outside forge3d_b200's scope: "
                                  "advance the volume with forge3d or an external simulation and hand the fields over")

    @property
    def dims(self):
        return self._dims

    @property
    def voxel_size(self):
        return self._voxel_size

    @property
    def origin(self):
        return self._origin

    def to_density_numpy(self) -> np.ndarray:
        return self.density.copy()

    def to_particle_age_numpy(self) -> np.ndarray:
        return self.particle_age.copy()

    # -- device residency -------------------------------------------------------------------------------------------
    def _invalidate(self) -> None:
        if self._handle is not None:
            _native.lib().f3d_smoke_destroy(self._handle)
            self._handle = None

    def _volume_struct(self):
        v = _native.SmokeVolume()
        v.dims = (C.c_uint32 * 3)(*self._dims)
        v.voxel_size = (C.c_float * 3)(*self._voxel_size)
        v.origin = (C.c_float * 3)(*self._origin)
        keep = []
        for name in self._FIELDS:
            a = np.ascontiguousarray(getattr(self, name), dtype=np.float32)
            keep.append(a)
            setattr(v, name, a.ctypes.data_as(C.POINTER(C.c_float)))
        v.frame_index = int(self.frame_index)
        return v, keep

    def _resident(self):
        if self._handle is None:
            v, keep = self._volume_struct()
            h = C.c_void_p()
            _native.check(_native.lib().f3d_smoke_create(C.byref(v), self.device, C.byref(h)))
            self._handle, self._uploaded_frame = h, int(self.frame_index)
            del keep
        elif self._uploaded_frame != int(self.frame_index):
            self._invalidate()
            return self._resident()
        return self._handle

    # -- rendering (src/smoke/py.rs:531-628) ------------------------------------------------------------------------
    def render_rgba(self, width, height, camera_pos, target, up=(0.0, 1.0, 0.0), fovy_deg=45.0,
                    sun_direction=(0.4, 0.8, -0.2), settings=None, certificate=None, cache=None) -> np.ndarray:
        """Perspective ray-march -> (height, width, 4) uint8 straight-alpha RGBA.  `certificate` / `cache` are accepted
        and ignored (certificates are out of scope; the reference ignores `cache`)."""
        del certificate, cache
        settings = settings or SmokeRenderSettings()
        out = np.zeros((int(height), int(width), 4), np.uint8)
        ms = C.c_double()
        f3 = C.c_float * 3
        _native.check(_native.lib().f3d_smoke_raymarch_rgba(
            self._resident(), C.byref(settings._native()), int(width), int(height), f3(*_tuple3(camera_pos, "camera_pos")),
            f3(*_tuple3(target, "target")), f3(*_tuple3(up, "up")), float(fovy_deg), f3(*_tuple3(sun_direction, "sun_direction")),
            out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(ms)))
        self.last_kernel_ms = float(ms.value)
        return out

    def render_over_rgba(self, base_rgba, camera_pos, target, up=(0.0, 1.0, 0.0), fovy_deg=45.0, sun_direction=(0.4, 0.8, -0.2),
                         settings=None, base_depth=None) -> np.ndarray:
        """BASELINE config 4: the smoke layer composited over a terrain frame, one kernel.  `base_rgba` = (H, W, 4) uint8 terrain
        snapshot (e.g. hybrid_render_terrain_reference(...)["rgba"] from the SAME camera); the composite is the reference's
        `_alpha_composite_rgba` (python/forge3d/map_scene.py:1588-1604).  `base_depth` = (H, W) float32 depth AOV of that
        snapshot (optional): ends each ray's march at the terrain, which the reference's caller-side composite cannot do."""
        settings = settings or SmokeRenderSettings()
        base = np.ascontiguousarray(base_rgba, np.uint8)
        if base.ndim != 3 or base.shape[2] != 4:
            raise ValueError(f"base_rgba must have shape (height, width, 4), got {base.shape}")
        height, width = base.shape[:2]
        depth = None
        if base_depth is not None:
            depth = np.ascontiguousarray(base_depth, np.float32)
            if depth.shape != (height, width):
                raise ValueError(f"base_depth must have shape {(height, width)}, got {depth.shape}")
        out = np.zeros((height, width, 4), np.uint8)
        ms = C.c_double()
        f3 = C.c_float * 3
        _native.check(_native.lib().f3d_smoke_raymarch_over_rgba(
            self._resident(), C.byref(settings._native()), int(width), int(height), f3(*_tuple3(camera_pos, "camera_pos")),
            f3(*_tuple3(target, "target")), f3(*_tuple3(up, "up")), float(fovy_deg), f3(*_tuple3(sun_direction, "sun_direction")),
            base.ctypes.data_as(C.POINTER(C.c_uint8)), depth.ctypes.data_as(C.POINTER(C.c_float)) if depth is not None else None,
            out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(ms)))
        self.last_kernel_ms = float(ms.value)
        return out

    def render_projection_rgba(self, width, height, view_direction=(0.0, -1.0, 0.0), sun_direction=(0.4, 0.8, -0.2),
                               settings=None, certificate=None, cache=None) -> np.ndarray:
        """Map-aligned parallel projection of the volume -> (height, width, 4) uint8."""
        del certificate, cache
        settings = settings or SmokeRenderSettings()
        out = np.zeros((int(height), int(width), 4), np.uint8)
        ms = C.c_double()
        f3 = C.c_float * 3
        _native.check(_native.lib().f3d_smoke_raymarch_projection_rgba(
            self._resident(), C.byref(settings._native()), int(width), int(height), f3(*_tuple3(view_direction, "view_direction")),
            f3(*_tuple3(sun_direction, "sun_direction")), out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(ms)))
        self.last_kernel_ms = float(ms.value)
        return out

    def close(self) -> None:
        self._invalidate()

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

    def __repr__(self) -> str:
        return f"SmokeDomain(dims={list(self._dims)}, time_seconds={self.time_seconds:.3f}, frame_index={self.frame_index})"


def domain_from_density(density, voxel_size=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0)) -> SmokeDomain:
    """forge3d.smoke.domain_from_density (python/forge3d/smoke.py:571-578)."""
    return SmokeDomain.from_density(np.ascontiguousarray(density, dtype=np.float32), voxel_size, origin)
