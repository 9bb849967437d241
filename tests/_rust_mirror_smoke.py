"""Second, independent restatement of the smoke ray-march (TEST INFRASTRUCTURE): written from the Rust text
(/root/reference/src/smoke/render.rs, sampling.rs, types.rs) in numpy-f32 scalars, one Rust expression per Python expression.
tests/test_smoke.py requires the C oracle to agree with it bit for bit on the u8 RGBA pixels, which pins the oracle's
operation order to the reference source under the stated pins (exp = exp2_pinned(x * log2 e), powf(d, 1.5) = d * sqrt(d))."""
import ctypes
import ctypes.util

import numpy as np

from _wgsl_mirror_aether import exp2

_libm = ctypes.CDLL(ctypes.util.find_library("m"))
_libm.tanf.argtypes = [ctypes.c_float]
_libm.tanf.restype = ctypes.c_float

f = np.float32
PI = f(3.14159274101257324)


def exp(x):
    return exp2(f(f(x) * f(1.4426950408889634)))


def clamp(x, lo, hi):          # f32::clamp
    x = f(x)
    return f(lo) if x < f(lo) else (f(hi) if x > f(hi) else x)


def fmax(a, b):                # f32::max (NaN-ignoring)
    a, b = f(a), f(b)
    if np.isnan(a):
        return b
    if np.isnan(b):
        return a
    return a if a > b else b


def fmin(a, b):
    a, b = f(a), f(b)
    if np.isnan(a):
        return b
    if np.isnan(b):
        return a
    return a if a < b else b


def lerp(a, b, t):
    return f(f(a) + f(f(f(b) - f(a)) * f(t)))


def dot(a, b):
    return f(f(f(a[0] * b[0]) + f(a[1] * b[1])) + f(a[2] * b[2]))


def length(a):
    return f(np.sqrt(dot(a, a)))


def normalize(a):
    r = f(f(1.0) / length(a))
    return tuple(f(c * r) for c in a)


def normalize_or_zero(a):
    with np.errstate(divide="ignore"):
        r = f(f(1.0) / length(a))
    if np.isfinite(r) and r > 0:
        return tuple(f(c * r) for c in a)
    return (f(0), f(0), f(0))


def cross(a, b):
    return (f(f(a[1] * b[2]) - f(b[1] * a[2])), f(f(a[2] * b[0]) - f(b[2] * a[0])), f(f(a[0] * b[1]) - f(b[0] * a[1])))


def smoothstep(e0, e1, x):
    t = clamp(f(f(f(x) - f(e0)) / fmax(f(f(e1) - f(e0)), 1.0e-6)), 0.0, 1.0)
    return f(f(t * t) * f(f(3.0) - f(f(2.0) * t)))


def hash01(v):
    v &= 0xFFFFFFFF
    v ^= v >> 16
    v = (v * 0x7FEB352D) & 0xFFFFFFFF
    v ^= v >> 15
    v = (v * 0x846CA68B) & 0xFFFFFFFF
    v ^= v >> 16
    return f(f(v) / f(4294967295))


def to_u8(v):
    c = f(f(clamp(v, 0.0, 1.0) * f(255.0)) + f(0.5))
    if np.isnan(c):
        return 0
    return int(min(max(int(c), 0), 255))


class Volume:
    def __init__(self, domain):
        self.dims = tuple(int(d) for d in domain.dims)
        self.voxel = tuple(f(v) for v in domain.voxel_size)
        self.origin = tuple(f(v) for v in domain.origin)
        self.fields = {n: np.ascontiguousarray(getattr(domain, n), np.float32).reshape(-1)
                       for n in ("density", "temperature", "soot", "humidity", "emission_rate", "particle_age")}
        self.frame_index = int(domain.frame_index)

    def bounds_max(self):
        return tuple(f(self.origin[a] + f(f(self.dims[a]) * self.voxel[a])) for a in range(3))

    def sample_scalar(self, name, p):
        fld, (nx, ny, nz) = self.fields[name], self.dims
        x, y, z = clamp(p[0], 0.0, nx - 1), clamp(p[1], 0.0, ny - 1), clamp(p[2], 0.0, nz - 1)
        x0, y0, z0 = int(np.floor(x)), int(np.floor(y)), int(np.floor(z))
        x1, y1, z1 = min(x0 + 1, nx - 1), min(y0 + 1, ny - 1), min(z0 + 1, nz - 1)
        fx, fy, fz = f(x - f(x0)), f(y - f(y0)), f(z - f(z0))
        idx = lambda X, Y, Z: (Z * ny + Y) * nx + X
        c00 = lerp(fld[idx(x0, y0, z0)], fld[idx(x1, y0, z0)], fx)
        c10 = lerp(fld[idx(x0, y1, z0)], fld[idx(x1, y1, z0)], fx)
        c01 = lerp(fld[idx(x0, y0, z1)], fld[idx(x1, y0, z1)], fx)
        c11 = lerp(fld[idx(x0, y1, z1)], fld[idx(x1, y1, z1)], fx)
        return lerp(lerp(c00, c10, fy), lerp(c01, c11, fy), fz)

    def sample(self, pos):
        p = tuple(f(f(f(pos[a] - self.origin[a]) / self.voxel[a]) - f(0.5)) for a in range(3))
        s = {n: self.sample_scalar(n, p) for n in self.fields}
        s["particle_age"] = fmax(s["particle_age"], 0.0)
        return s


def ray_box(o, d, bmin, bmax):
    with np.errstate(invalid="ignore", over="ignore"):
        inv = tuple(f(f(1.0) / d[a]) if abs(d[a]) > f(1.0e-12) else f(np.inf) for a in range(3))
        t0 = tuple(f(f(bmin[a] - o[a]) * inv[a]) for a in range(3))
        t1 = tuple(f(f(bmax[a] - o[a]) * inv[a]) for a in range(3))
    tmin = tuple(fmin(t0[a], t1[a]) for a in range(3))
    tmax = tuple(fmax(t0[a], t1[a]) for a in range(3))
    near = fmax(fmax(tmin[0], tmin[1]), tmin[2])
    far = fmin(fmin(tmax[0], tmax[1]), tmax[2])
    return (near, far) if far >= fmax(near, 0.0) else None


def add(a, b):
    return tuple(f(x + y) for x, y in zip(a, b))


def scale(a, s):
    return tuple(f(x * f(s)) for x in a)


def mul(a, b):
    return tuple(f(x * y) for x, y in zip(a, b))


def mix3(a, b, t):
    return tuple(lerp(x, y, t) for x, y in zip(a, b))


def sun_transmittance(V, start, sun_dir, step, steps, st):
    hit = ray_box(add(start, scale(sun_dir, step)), sun_dir, V.origin, V.bounds_max())
    if hit is None:
        return f(1.0)
    t0, t1 = fmax(hit[0], 0.0), hit[1]
    od = f(0.0)
    for i in range(int(steps)):
        t = f(t0 + f(f(f(i) + f(0.5)) * step))
        if t > t1:
            break
        s = V.sample(add(start, scale(sun_dir, f(step + t))))
        age_t = smoothstep(1.6, 17.0, s["particle_age"])
        gate = f(f(0.50) + f(f(0.50) * smoothstep(0.045, 0.34, s["density"])))
        term = f(s["density"] * f(st.density_scale))
        term = f(term * f(f(1.0) - f(f(0.58) * age_t)))
        term = f(term * gate)
        term = f(term * f(st.extinction))
        term = f(term * f(f(1.0) + f(s["soot"] * f(st.soot_absorption))))
        term = f(term * step)
        od = f(od + term)
        if od > f(8.0):
            break
    return clamp(exp(f(-od)), 0.0, 1.0)


def smoke_color(s, st):
    body = clamp(f(f(s["density"] * f(1.45)) + f(s["soot"] * f(1.35))), 0.0, 1.0)
    color = mix3(tuple(f(c) for c in st.thin_color), tuple(f(c) for c in st.dense_color), body)
    aged = clamp(f(s["particle_age"] / f(9.0)), 0.0, 1.0)
    color = mix3(color, (f(0.36), f(0.39), f(0.43)), f(aged * f(0.42)))
    milk = f(clamp(s["humidity"], 0.0, 1.0) * f(f(0.18) + f(f(0.42) * body)))
    color = mix3(color, (f(0.93), f(0.92), f(0.84)), clamp(milk, 0.0, 0.38))
    freshness = clamp(f(f(1.0) - f(s["particle_age"] / f(17.0))), 0.0, 1.0)
    heat = clamp(f(f(s["temperature"] * f(0.12)) * freshness), 0.0, 1.0)
    return mix3(color, (f(0.95), f(0.62), f(0.28)), f(heat * f(0.07)))


def march(V, origin, ray_dir, t0, t1, seed, step, shadow_step, sun_dir, st):
    jitter = f(f(f(hash01(seed) - f(0.5)) * f(st.jitter_strength)) * step)
    t = fmax(f(t0 + jitter), 0.0)
    tr = f(1.0)
    rgb = (f(0), f(0), f(0))
    steps = 0
    while t < t1 and steps < int(st.max_steps) and tr > f(0.01):
        p = add(origin, scale(ray_dir, t))
        s = V.sample(p)
        age_t = smoothstep(1.6, 17.0, s["particle_age"])
        gate = f(f(0.50) + f(f(0.50) * smoothstep(0.045, 0.34, s["density"])))
        density = fmax(f(f(f(s["density"] * f(st.density_scale)) * f(f(1.0) - f(f(0.58) * age_t))) * gate), 0.0)
        if density > f(1.0e-5):
            sigma_t = f(f(density * f(st.extinction)) * f(f(1.0) + f(f(s["soot"] * f(st.soot_absorption)) * f(0.85))))
            seg_t = clamp(exp(f(f(-sigma_t) * step)), 0.0, 1.0)
            seg_w = f(f(f(1.0) - seg_t) / sigma_t) if sigma_t > f(1.0e-6) else step
            light = sun_transmittance(V, p, sun_dir, shadow_step, st.shadow_steps, st) if st.self_shadow else f(1.0)
            cos_theta = clamp(dot(ray_dir, sun_dir), -1.0, 1.0)
            g = f(st.phase_g)
            g2 = f(g * g)
            denom = fmax(f(f(f(1.0) + g2) - f(f(f(2.0) * g) * cos_theta)), 1.0e-4)
            phase = f(f(f(1.0) - g2) / f(f(f(4.0) * PI) * f(denom * f(np.sqrt(denom)))))
            color = smoke_color(s, st)
            albedo = clamp(f(f(st.scattering) / f(f(f(f(st.scattering) + f(st.absorption)) + f(s["soot"] * f(0.55))) + f(1.0e-5))), 0.02, 0.98)
            sigma_s = f(sigma_t * albedo)
            sun_rad = scale((f(1.0), f(0.96), f(0.84)), 11.5)
            sky = scale(scale((f(0.52), f(0.60), f(0.72)), f(f(0.36) + f(f(0.26) * clamp(f(f(1.0) - light), 0.0, 1.0)))),
                        clamp(f(f(1.0) - f(s["soot"] * f(0.32))), 0.50, 1.0))
            ground = scale(scale((f(0.58), f(0.54), f(0.48)), 0.070),
                           clamp(f(f(1.0) - f(p[1] / fmax(V.bounds_max()[1], 1.0))), 0.0, 1.0))
            powder = clamp(f(f(1.0) - exp(f(f(f(-sigma_t) * step) * f(2.2)))), 0.0, 1.0)
            pw = f(f(powder * f(0.055)) * f(np.sqrt(light)))
            multiple = mul(scale(color, sigma_s), add(add(sky, ground), (pw, pw, pw)))
            direct = scale(scale(mul(scale(color, sigma_s), sun_rad), phase), light)
            freshness = clamp(f(f(1.0) - f(s["particle_age"] / f(17.0))), 0.0, 1.0)
            fresh_heat = f(f(s["temperature"] * freshness) * freshness)
            emission = scale((f(1.0), f(0.30), f(0.055)),
                             clamp(f(f(f(fresh_heat * f(0.10)) + f(s["emission_rate"] * f(1.18))) * f(st.fire_glow)), 0.0, 5.0))
            source = add(add(direct, multiple), emission)
            rgb = add(rgb, scale(scale(source, seg_w), tr))
            tr = f(tr * seg_t)
        t = f(t + step)
        steps += 1
    alpha = clamp(f(f(1.0) - tr), 0.0, 1.0)
    straight = tuple(f(c / alpha) for c in rgb) if alpha > f(1.0e-5) else rgb
    e = scale(straight, st.exposure)
    return [to_u8(f(c / f(f(1.0) + c))) for c in e] + [to_u8(alpha)]


def steps_for(V, st):
    min_step = fmax(fmin(fmin(fmin(np.inf, V.voxel[0]), V.voxel[1]), V.voxel[2]), 1.0e-4)
    step = f(st.step_size) if st.step_size > 0.0 else f(min_step * f(0.75))
    shadow = f(st.shadow_step_size) if st.shadow_step_size > 0.0 else f(step * f(2.0))
    return step, shadow


def perspective_pixel(V, st, width, height, x, y, camera_pos, target, up, fovy_deg, sun_direction):
    eye = tuple(f(c) for c in camera_pos)
    forward = normalize_or_zero(tuple(f(f(t) - e) for t, e in zip(target, eye)))
    upn = normalize_or_zero(tuple(f(c) for c in up))
    right = normalize_or_zero(cross(forward, upn))
    cam_up = normalize_or_zero(cross(right, forward))
    sun = normalize_or_zero(tuple(f(c) for c in sun_direction))
    step, shadow = steps_for(V, st)
    tan_half = f(_libm.tanf(float(f(f(f(fovy_deg) * f(PI / f(180.0))) * f(0.5)))))   # f32::tan -> libm tanf (host set-up)
    aspect = f(f(width) / f(height))
    px = f(f(f(f(f(f(f(x) + f(0.5)) / f(width)) * f(2.0)) - f(1.0)) * aspect) * tan_half)
    py = f(f(f(1.0) - f(f(f(f(y) + f(0.5)) / f(height)) * f(2.0))) * tan_half)
    ray_dir = normalize(add(add(forward, scale(right, px)), scale(cam_up, py)))
    hit = ray_box(eye, ray_dir, V.origin, V.bounds_max())
    if hit is None:
        return [0, 0, 0, 0]
    seed = (x * 73856093 + y * 19349663 + (V.frame_index & 0xFFFFFFFF)) & 0xFFFFFFFF
    return march(V, eye, ray_dir, fmax(hit[0], 0.0), hit[1], seed, step, shadow, sun, st)


def projection_pixel(V, st, width, height, px, py, view_direction, sun_direction):
    ray_dir = normalize_or_zero(tuple(f(c) for c in view_direction))
    sun = normalize_or_zero(tuple(f(c) for c in sun_direction))
    step, shadow = steps_for(V, st)
    bmin, bmax = V.origin, V.bounds_max()
    diagonal = fmax(length(tuple(f(b - a) for a, b in zip(bmin, bmax))), f(step * f(2.0)))
    fz = f(f(f(py) + f(0.5)) / f(height))
    z = lerp(bmin[2], bmax[2], fz)
    fx = f(f(f(px) + f(0.5)) / f(width))
    x = lerp(bmin[0], bmax[0], fx)
    plane = (x, f(f(bmin[1] + bmax[1]) * f(0.5)), z)
    origin = tuple(f(p - c) for p, c in zip(plane, scale(ray_dir, diagonal)))
    hit = ray_box(origin, ray_dir, bmin, bmax)
    if hit is None:
        return [0, 0, 0, 0]
    seed = (px * 73856093 + py * 19349663 + (V.frame_index & 0xFFFFFFFF) + 0x9E3779B9) & 0xFFFFFFFF
    return march(V, origin, ray_dir, fmax(hit[0], 0.0), hit[1], seed, step, shadow, sun, st)
