"""Session wrapper over the C ABI's f3d_session_* entry points: the reference driver loop
(/root/reference/src/path_tracing/hybrid_compute/render_terrain.rs:1123-1244) split at its joints so
a host can keep the scene resident in HBM, time the frame loop alone, or drive one image partition
per GPU.  Used by bench.py and forge3d_b200.distributed; the one-call path does not need it."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native

DEFAULTS = dict(spacing=(1.0, 1.0), exaggeration=1.0, albedo=(0.6, 0.6, 0.6), sun_azimuth_deg=315.0,
                sun_elevation_deg=45.0, sun_intensity=2.5, env_map=None, env_intensity=0.35, mesh_vertices=None,
                mesh_indices=None, spp=1, max_frames=512, min_frames=32, variance_threshold=1e-3, seed=7,
                sun_color=(1.0, 0.97, 0.92), observer_latitude_deg=0.0, observer_longitude_deg=0.0,
                earth_model="ellipsoid", sphere_radius_m=6371008.8, refraction_model="bennett", refraction_k=0.13,
                pressure_mbar=1013.25, temperature_c=15.0)


class Session:
    """A resident scene + per-pixel state on one CUDA device."""

    def __init__(self, heightmap, width, height, cam=None, *, device=0, cuda_stream=0, part_rank=0, part_world=1,
                 part_block_rows=0, part_mode=0, compat_512mib_gate=False, atmosphere=None, numerics=None, **kw):
        from .atmosphere import resolve_atmosphere

        args = dict(DEFAULTS)
        unknown = set(kw) - set(args)
        if unknown:
            raise TypeError(f"unexpected arguments: {sorted(unknown)}")
        args.update(kw)
        args["sun_color"] = _native.extract_sun_color(args["sun_color"])
        self.width, self.height = int(width), int(height)
        self._args = args
        desc, keep = _native.make_desc(heightmap, width, height, cam, device=device,
                                       compat_512mib_gate=compat_512mib_gate, part_rank=part_rank,
                                       part_world=part_world, part_block_rows=part_block_rows, part_mode=part_mode,
                                       atmosphere=resolve_atmosphere(atmosphere), **args)
        self.numerics = numerics or _native.default_numerics()
        self._L = _native.lib(self.numerics)
        self._h = C.c_void_p()
        self._check(self._L.f3d_session_create(C.byref(desc), C.c_void_p(int(cuda_stream) or None), C.byref(self._h)))
        del keep

    # -- frame loop -------------------------------------------------------------------------------
    def _check(self, rc):
        _native.check(rc, self._L)

    def render_frames(self, n: int) -> None:
        self._check(self._L.f3d_session_render_frames(self._h, int(n)))

    def sync(self) -> None:
        self._check(self._L.f3d_session_sync(self._h))

    def last_frames_ms(self) -> float:
        ms = C.c_double()
        self._check(self._L.f3d_session_last_frames_ms(self._h, C.byref(ms)))
        return float(ms.value)

    @property
    def frames(self) -> int:
        n = C.c_uint32()
        self._check(self._L.f3d_session_frames(self._h, C.byref(n)))
        return int(n.value)

    def variance(self):
        v, bad = C.c_float(), C.c_int32()
        self._check(self._L.f3d_session_variance(self._h, C.byref(v), C.byref(bad)))
        return float(v.value), bool(bad.value)

    # -- outputs ----------------------------------------------------------------------------------
    def resolve_device(self, rgba=0, albedo=0, normal=0, depth=0, check_validity=True) -> None:
        """Resolve owned rows into device buffers given as raw pointers (e.g. torch .data_ptr())."""
        p = lambda v: C.c_void_p(int(v) or None)
        self._check(self._L.f3d_session_resolve_device(self._h, p(rgba), p(albedo), p(normal), p(depth),
                                                         int(bool(check_validity))))

    def validity(self):
        """(any_valid, required) of the last resolve_device(check_validity=True): see f3d_session_validity."""
        any_valid, required = C.c_int32(), C.c_int32()
        self._check(self._L.f3d_session_validity(self._h, C.byref(any_valid), C.byref(required)))
        return bool(any_valid.value), bool(required.value)

    def resolve_host(self, want_accum=False) -> dict:
        o, arrays = _native.alloc_outputs(self.width, self.height, want_accum)
        self._check(self._L.f3d_session_resolve_host(self._h, C.byref(o)))
        return _native.result_dict(o, arrays, self._args["sun_azimuth_deg"], self._args["sun_elevation_deg"])

    def stats(self) -> dict:
        o = _native.TerrainOut()
        self._check(self._L.f3d_session_stats(self._h, C.byref(o)))
        return dict(frames=int(o.frames), rays_primary=int(o.rays_primary), rays_shadow=int(o.rays_shadow),
                    rays_ibl=int(o.rays_ibl), nodes_popped=int(o.nodes_popped), setup_ms=float(o.setup_ms),
                    gpu_resource_bytes=int(o.gpu_resource_bytes), minmax_pyramid_bytes=int(o.minmax_pyramid_bytes),
                    kernel_launches=int(o.kernel_launches))

    # -- NVLink peer halo exchange ------------------------------------------------------------------
    def ipc_export(self) -> bytes:
        buf = (C.c_uint8 * (_native.IPC_HANDLE_BYTES * _native.IPC_HANDLES_PER_RANK))()
        self._check(self._L.f3d_session_ipc_export(self._h, buf))
        return bytes(buf)

    def ipc_import(self, all_handles: bytes) -> None:
        buf = (C.c_uint8 * len(all_handles)).from_buffer_copy(all_handles)
        self._check(self._L.f3d_session_ipc_import(self._h, buf))

    def close(self, trim: bool = False) -> None:
        """Destroys the session.  Its device buffers are parked in the library's cache for the next session (bounded, see
        f3d_cache_trim in include/forge3d_b200.h); trim=True hands them straight back to the CUDA driver instead."""
        if self._h:
            self._L.f3d_session_destroy(self._h)
            self._h = C.c_void_p()
            if trim:
                self._L.f3d_cache_trim(-1)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
