"""forge3d_b200.wavefront -- host side of the wavefront multi-bounce path tracer.

Mirrors the reference's scene description and entry points for this path:

* ``ReferenceSceneDesc`` / ``adjudication_scene()``   -- src/path_tracing/reference_scene.rs:19-100,119-240
* ``render_pt_reference(desc, w, h, spp_frames)``      -- src/path_tracing/adjudication.rs:76-331 (linear HDR mean, alpha 1)
* ``resolve`` to RGBA8 (Reinhard + sRGB)                -- src/core/tonemap.rs:11-32, done on the device by the same call
* ``render_adjudication_pt(w, h, spp)``                 -- the path-traced half of forge3d.render_adjudication_pair
                                                           (src/py_functions/adjudication.rs:19-170); the raster twin is the
                                                           reference's rasteriser and is out of this path's scope.

Everything is computed by libforge3d_b200.so on the GPU (csrc/f3d_wavefront.cu); there is no CPU fallback here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

from . import _native

f32 = np.float32


@dataclass
class SphereDesc:
    """reference_scene.rs:17-23"""
    center: Sequence[float]
    radius: float
    albedo: Sequence[float]
    roughness: float
    metallic: float = 0.0
    ior: float = 1.0
    emissive: Sequence[float] = (0.0, 0.0, 0.0)


def _v(a) -> np.ndarray:
    return np.asarray(a, dtype=f32).reshape(3)


def _normalize_glam(a: np.ndarray) -> np.ndarray:
    """glam 0.24 Vec3::normalize: self * (1 / sqrt(dot(self, self))), all in f32."""
    a = a.astype(f32)
    d = f32(f32(a[0] * a[0]) + f32(a[1] * a[1])) + f32(a[2] * a[2])
    inv = f32(1.0) / np.sqrt(f32(d), dtype=f32)
    return (a * f32(inv)).astype(f32)


def _cross_glam(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = a.astype(f32)
    b = b.astype(f32)
    return np.array([f32(a[1] * b[2]) - f32(b[1] * a[2]), f32(a[2] * b[0]) - f32(b[2] * a[0]), f32(a[0] * b[1]) - f32(b[0] * a[1])],
                    dtype=f32)


@dataclass
class ReferenceSceneDesc:
    """reference_scene.rs:26-49; every numeric field is rounded to f32 when packed."""
    cam_origin: Sequence[float]
    cam_look_at: Sequence[float]
    cam_up: Sequence[float]
    fov_y_deg: float
    exposure: float
    spheres: List[SphereDesc]
    sun_direction: Sequence[float]
    sun_intensity: float
    sun_color: Sequence[float]
    ambient_color: Sequence[float]
    sky_color: Sequence[float]
    plane_half_extent: float
    seed_hi: int
    seed_lo: int

    # -- reference_scene.rs:122-142
    def wavefront_spheres(self) -> np.ndarray:
        out = np.zeros((len(self.spheres), 20), f32)
        for i, s in enumerate(self.spheres):
            out[i, 0:3] = _v(s.center)
            out[i, 3] = s.radius
            out[i, 4:7] = _v(s.albedo)
            out[i, 7] = s.metallic
            out[i, 8] = s.roughness
            out[i, 9] = s.ior
            out[i, 12:15] = _v(s.emissive)
        return out

    # -- reference_scene.rs:144-151, lighting.rs:98-116 (direction normalised by division, intensity/importance clamped at 0)
    def directional_lights(self) -> np.ndarray:
        return pack_directional_light(self.sun_direction, self.sun_intensity, self.sun_color, 1.0).reshape(1, 8)

    # -- reference_scene.rs:157-166: one inert disc far below the plane, facing down
    def area_lights(self) -> np.ndarray:
        return pack_area_light([0.0, -1.0e4, 0.0], [0.0, -1.0, 0.0], 1.0e-6, 0.0, [0.0, 0.0, 0.0], 0.0).reshape(1, 12)

    def object_importance(self) -> np.ndarray:
        return np.ones(4, f32)

    # -- reference_scene.rs:178-187
    def environment_raw(self) -> np.ndarray:
        a, s = _v(self.ambient_color), _v(self.sky_color)
        e = np.zeros((4, 4), f32)
        e[0, :3] = a
        e[1, :3] = a
        e[2, :3] = s
        e[3, :3] = s
        return e.reshape(16)

    # -- reference_scene.rs:191-197
    def plane_mesh(self) -> Tuple[np.ndarray, np.ndarray]:
        e = f32(self.plane_half_extent)
        v = np.array([[-e, 0.0, -e], [-e, 0.0, e], [e, 0.0, e], [e, 0.0, -e]], f32)
        return v, np.array([[0, 1, 2], [0, 2, 3]], np.uint32)

    # -- reference_scene.rs:201-207
    def camera_basis(self):
        origin = _v(self.cam_origin)
        forward = _normalize_glam(_v(self.cam_look_at) - origin)
        right = _normalize_glam(_cross_glam(forward, _v(self.cam_up)))
        up = _normalize_glam(_cross_glam(right, forward))
        return origin, forward, right, up

    def fov_y_rad(self) -> np.float32:  # Rust f32::to_radians
        return f32(f32(self.fov_y_deg) * f32(f32(3.14159274101257324) / f32(180.0)))

    # -- reference_scene.rs:215-244
    def metadata_fields(self, width: int, height: int, spp: int) -> dict:
        sun = _normalize_glam(_v(self.sun_direction))
        d = lambda x: float(f32(x))
        o, l = _v(self.cam_origin), _v(self.cam_look_at)
        sc, am, sk = _v(self.sun_color), _v(self.ambient_color), _v(self.sky_color)
        return {
            "cam_origin_x": d(o[0]), "cam_origin_y": d(o[1]), "cam_origin_z": d(o[2]),
            "cam_look_at_x": d(l[0]), "cam_look_at_y": d(l[1]), "cam_look_at_z": d(l[2]),
            "fov_y_deg": d(self.fov_y_deg), "exposure": d(self.exposure),
            "sun_dir_x": d(sun[0]), "sun_dir_y": d(sun[1]), "sun_dir_z": d(sun[2]),
            "sun_intensity": d(self.sun_intensity),
            "sun_color_r": d(sc[0]), "sun_color_g": d(sc[1]), "sun_color_b": d(sc[2]),
            "ambient_r": d(am[0]), "ambient_g": d(am[1]), "ambient_b": d(am[2]),
            "sky_r": d(sk[0]), "sky_g": d(sk[1]), "sky_b": d(sk[2]),
            "width": float(width), "height": float(height), "spp": float(spp),
        }


def pack_directional_light(direction, intensity, color, importance) -> np.ndarray:
    """GpuDirectionalLight::new, lighting.rs:98-116."""
    d = _v(direction)
    ln = np.sqrt(f32(f32(f32(d[0] * d[0]) + f32(d[1] * d[1])) + f32(d[2] * d[2])), dtype=f32)
    dn = (d / ln).astype(f32) if ln > 0 else np.array([0.0, -1.0, 0.0], f32)
    out = np.zeros(8, f32)
    out[0:3] = dn
    out[3] = max(f32(intensity), f32(0.0))
    out[4:7] = _v(color)
    out[7] = max(f32(importance), f32(0.0))
    return out


def pack_area_light(position, normal, radius, intensity, color, importance) -> np.ndarray:
    """GpuAreaLight::disc, lighting.rs:32-48."""
    out = np.zeros(12, f32)
    out[0:3] = _v(position)
    out[3] = max(f32(radius), f32(0.0))
    out[4:7] = _v(normal)
    out[7] = max(f32(intensity), f32(0.0))
    out[8:11] = _v(color)
    out[11] = max(f32(importance), f32(0.0))
    return out


def pack_instance(object_to_world=None, world_to_object=None, blas_index: int = 0, material_id: int = 0) -> np.ndarray:
    """accel::instancing::InstanceData (column-major 4x4 pair + blas/material), adjudication.rs:141-147 -> 36 words."""
    ident = np.eye(4, dtype=f32).reshape(16)
    out = np.zeros(36, f32)
    out[0:16] = ident if object_to_world is None else np.asarray(object_to_world, f32).reshape(16)
    out[16:32] = ident if world_to_object is None else np.asarray(world_to_object, f32).reshape(16)
    out[32:34] = np.array([blas_index, material_id], np.uint32).view(f32)
    return out


def adjudication_scene() -> ReferenceSceneDesc:
    """The committed adjudication scene, reference_scene.rs:53-100."""
    return ReferenceSceneDesc(
        cam_origin=[0.0, 2.2, 6.5], cam_look_at=[0.0, 0.9, 0.0], cam_up=[0.0, 1.0, 0.0], fov_y_deg=40.0, exposure=1.0,
        spheres=[
            SphereDesc([-1.15, 1.0, 0.0], 1.0, [0.63, 0.28, 0.22], 0.70),
            SphereDesc([1.30, 0.8, 0.55], 0.8, [0.24, 0.40, 0.62], 0.55),
            SphereDesc([0.25, 0.5, -1.45], 0.5, [0.78, 0.68, 0.30], 0.85),
            SphereDesc([0.0, -1000.0, 0.0], 0.0, [0.42, 0.42, 0.42], 0.90),  # plane material slot (radius 0)
        ],
        sun_direction=[-0.45, -0.80, -0.30], sun_intensity=3.2, sun_color=[1.0, 0.97, 0.92],
        ambient_color=[0.40, 0.48, 0.62], sky_color=[0.35, 0.45, 0.70], plane_half_extent=40.0,
        seed_hi=0x9E3779B9, seed_lo=0x85EBCA6B)


@dataclass
class WavefrontScene:
    """The buffers render_pt_reference binds (adjudication.rs:97-176), packed as the C ABI takes them."""
    cam_origin: np.ndarray
    cam_forward: np.ndarray
    cam_right: np.ndarray
    cam_up: np.ndarray
    fov_y_rad: float
    exposure: float
    seed_hi: int
    seed_lo: int
    spheres: np.ndarray                                   # (n, 20) f32
    dir_lights: np.ndarray                                # (n, 8)
    area_lights: np.ndarray                               # (n, 12)
    importance: np.ndarray                                # (n,)
    environment: np.ndarray                               # (16,)
    mesh_xyz: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), f32))
    mesh_idx: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.uint32))
    instances: np.ndarray = field(default_factory=lambda: np.zeros((0, 36), f32))

    def normalized(self) -> "WavefrontScene":
        c = lambda a, shape, dt=f32: np.ascontiguousarray(np.asarray(a, dtype=dt).reshape(shape))
        return WavefrontScene(c(self.cam_origin, 3), c(self.cam_forward, 3), c(self.cam_right, 3), c(self.cam_up, 3),
                              float(f32(self.fov_y_rad)), float(f32(self.exposure)), int(self.seed_hi) & 0xFFFFFFFF,
                              int(self.seed_lo) & 0xFFFFFFFF, c(self.spheres, (-1, 20)), c(self.dir_lights, (-1, 8)),
                              c(self.area_lights, (-1, 12)), c(self.importance, -1), c(self.environment, 16),
                              c(self.mesh_xyz, (-1, 3)), c(self.mesh_idx, (-1, 3), np.uint32), c(self.instances, (-1, 36)))


def scene_from_desc(desc: ReferenceSceneDesc) -> WavefrontScene:
    """The wiring of render_pt_reference, adjudication.rs:91-223: plane mesh as instance 0 with material slot 3."""
    origin, forward, right, up = desc.camera_basis()
    v, t = desc.plane_mesh()
    return WavefrontScene(origin, forward, right, up, float(desc.fov_y_rad()), float(f32(desc.exposure)), desc.seed_hi, desc.seed_lo,
                          desc.wavefront_spheres(), desc.directional_lights(), desc.area_lights(), desc.object_importance(),
                          desc.environment_raw(), v, t, pack_instance(material_id=3).reshape(1, 36)).normalized()


@dataclass
class WavefrontStats:
    rays: int                 # rays traced over all frames (primary + continuation)
    max_rays_per_frame: int   # the reference's append-only ray queue holds 4*W*H of these (wavefront/mod.rs:34,77)
    min_iterations: int       # fewest wavefront iterations any frame ran (must be >= 2, adjudication.rs:259-265)
    launches: int             # kernels launched
    kernel_ms: float          # device time of the frames + resolve (CUDA events)
    frame_iterations: object = None   # per frame, partitioned renders only
    frame_rays: object = None


def render_pt_reference(scene, width: int, height: int, spp_frames: int, device: int = 0, return_rgba8: bool = False,
                        return_stats: bool = False, part=None):
    """Linear-HDR mean radiance over ``spp_frames`` frames, RGBA f32 (H, W, 4), alpha 1 (adjudication.rs:76-331).

    ``part=(rank, world, block_rows)`` renders only the rows that rank owns (interleaved blocks, see forge3d_b200.distributed); the
    returned images are full-size with the other rows untouched, the reference's two frame rules are left to the caller, and the
    stats carry ``frame_iterations`` / ``frame_rays`` (per frame, this rank) for the reduction over ranks."""
    if isinstance(scene, ReferenceSceneDesc):
        scene = scene_from_desc(scene)
    s = scene.normalized()
    if width <= 0 or height <= 0 or spp_frames <= 0:
        raise ValueError("adjudication PT reference requires non-zero width/height/spp")
    L = _native.lib()
    cs, keep = _native.make_wavefront_scene(s)
    hdr = np.zeros((height, width, 4), f32)
    rgba = np.zeros((height, width, 4), np.uint8)
    st = _native.WavefrontStats()
    fp, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    if part is None:
        rc = L.f3d_wavefront_render(C.byref(cs), width, height, spp_frames, device, hdr.ctypes.data_as(fp), rgba.ctypes.data_as(u8p),
                                    C.byref(st))
        frame_iters = frame_rays = None
    else:
        frame_iters, frame_rays = np.zeros(spp_frames, np.uint32), np.zeros(spp_frames, np.uint64)
        cp = _native.WavefrontPart(int(part[0]), int(part[1]), int(part[2]), frame_iters.ctypes.data_as(C.POINTER(C.c_uint32)),
                                   frame_rays.ctypes.data_as(C.POINTER(C.c_uint64)))
        rc = L.f3d_wavefront_render_part(C.byref(cs), width, height, spp_frames, device, C.byref(cp), hdr.ctypes.data_as(fp),
                                         rgba.ctypes.data_as(u8p), C.byref(st))
    del keep
    if rc != 0:
        _native.raise_last(rc)
    out = [hdr]
    if return_rgba8:
        out.append(rgba)
    if return_stats:
        ws = WavefrontStats(int(st.rays), int(st.max_rays_per_frame), int(st.min_iterations), int(st.launches), float(st.kernel_ms))
        ws.frame_iterations, ws.frame_rays = frame_iters, frame_rays
        out.append(ws)
    return out[0] if len(out) == 1 else tuple(out)


def check_frame_rules(frame_iterations, frame_rays, width: int, height: int) -> None:
    """The reference's two per-frame rules on whole-image numbers (render.rs:127-137, adjudication.rs:259-265), same texts."""
    capacity = 4 * width * height
    for fr, (it, rays) in enumerate(zip(frame_iterations, frame_rays)):
        if int(rays) > capacity:
            raise RuntimeError(f"wavefront frame {fr}: wavefront ray queue overflow: {int(rays)} rays pushed into capacity {capacity}")
        if int(it) < 2:
            raise RuntimeError(f"adjudication PT frame {fr} executed {int(it)} wavefront iteration(s); "
                               "a multi-bounce path-traced reference requires >= 2")


def render_adjudication_pt(width: int, height: int, spp: int, device: int = 0):
    """The path-traced half of forge3d.render_adjudication_pair: (pt_rgba uint8 (H, W, 4), metadata)."""
    if width <= 0 or height <= 0 or spp <= 0:
        raise ValueError("render_adjudication_pair requires width > 0, height > 0, spp > 0")
    desc = adjudication_scene()
    _, rgba = render_pt_reference(desc, width, height, spp, device=device, return_rgba8=True)
    return rgba, {"pt": desc.metadata_fields(width, height, spp)}
