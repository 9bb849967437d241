"""A second, independent restatement of the reference's terrain path tracer in pure Python / numpy float32
scalars, transcribed directly from the WGSL (not from oracle/f3d_oracle.c), for TINY cases only.

Purpose: the C oracle is what every CUDA result is compared with; this mirror checks the oracle itself against
a fresh reading of /root/reference/src/shaders/{hybrid_terrain_traversal,hybrid_traversal,hybrid_kernel,
pt_restir_temporal,pt_restir_spatial}.wgsl and render_terrain.rs, so that a transcription slip in the C file
cannot silently become "the truth".  Same numerics contract: every operation rounds to
binary32, no fused operations, pinned sin/cos.  Terrain only (no mesh), constant environment.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math

import numpy as np

f = np.float32
ZERO, ONE = f(0.0), f(1.0)


_LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _name in ("cosf", "sinf", "tanf"):
    getattr(_LIBM, _name).restype = ctypes.c_float
    getattr(_LIBM, _name).argtypes = [ctypes.c_float]


def _libm(name, x):
    """Host-side trigonometry (Rust f32::cos/sin -> libm, render_terrain.rs:639-642); not the oracle's code."""
    return f(getattr(_LIBM, name)(float(x)))


def fmax(a, b):
    return a if a > b else b       # operands are never NaN here


def fmin(a, b):
    return a if a < b else b


def clamp(x, lo, hi):
    return fmin(fmax(x, lo), hi)


def dot3(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def normalize(v):
    inv = ONE / np.sqrt(dot3(v, v))
    return (v[0] * inv, v[1] * inv, v[2] * inv)


def cross(a, b):
    return (a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1])


def mix(a, b, t):
    return a * (ONE - t) + b * t


def lum(c):
    return dot3(c, (f(0.2126), f(0.7152), f(0.0722)))


def sincos_pinned(x):
    k = int(x * f(0.636619772) + f(0.5))
    fk = f(k)
    r = x - fk * f(1.5703125)
    r = r - fk * f(4.837512969970703125e-4)
    r = r - fk * f(7.549789954891882e-8)
    z = r * r
    sp = ((f(-1.9515295891e-4) * z + f(8.3321608736e-3)) * z - f(1.6666654611e-1)) * z * r + r
    cp = ((f(2.443315711809948e-5) * z - f(1.388731625493765e-3)) * z + f(4.166664568298827e-2)) * z * z - f(0.5) * z + ONE
    q = k & 3
    if q == 0:
        return sp, cp
    if q == 1:
        return cp, -sp
    if q == 2:
        return -sp, -cp
    return -cp, sp


class Rng:
    def __init__(self, state):
        self.s = state & 0xFFFFFFFF

    def next(self):                                   # hybrid_kernel.wgsl:78-85
        x = self.s
        x ^= (x << 13) & 0xFFFFFFFF
        x ^= x >> 17
        x ^= (x << 5) & 0xFFFFFFFF
        self.s = x
        return f(x) / f(4294967296.0)


def tent(u):                                          # hybrid_terrain_traversal.wgsl:409-414
    if u < f(0.5):
        return np.sqrt(f(2.0) * u) - ONE
    return ONE - np.sqrt(f(2.0) * (ONE - u))


class Scene:
    def __init__(self, dem, spacing, exaggeration, albedo, env_intensity, inv_two_r_prime, curvature_enabled):
        self.h = np.asarray(dem, np.float32)
        self.rows, self.cols = self.h.shape
        self.sx, self.sz = f(spacing[0]), f(spacing[1])
        self.ex = f(exaggeration)
        self.ox = f(-0.5) * (f(self.cols) - ONE) * self.sx        # terrain_heightfield.rs:359-360
        self.oz = f(-0.5) * (f(self.rows) - ONE) * self.sz
        self.albedo = tuple(f(a) for a in albedo)
        self.env = f(env_intensity)
        self.k = f(inv_two_r_prime)
        self.curv = bool(curvature_enabled)
        self.cw, self.ch = self.cols - 1, self.rows - 1
        # build_minmax_mips, terrain_heightfield.rs:132-202
        pw = 1
        while pw < self.cw:
            pw *= 2
        ph = 1
        while ph < self.ch:
            ph *= 2
        lv = np.empty((ph, pw, 2), np.float32)
        lv[..., 0] = np.inf
        lv[..., 1] = -np.inf
        for y in range(self.ch):
            for x in range(self.cw):
                c = self.h[y:y + 2, x:x + 2]
                lv[y, x, 0] = c.min()
                lv[y, x, 1] = c.max()
        self.levels = [lv]
        while self.levels[-1].shape[0] > 1 or self.levels[-1].shape[1] > 1:
            p = self.levels[-1]
            lh, lw = p.shape[:2]
            nh, nw = max(lh // 2, 1), max(lw // 2, 1)
            n = np.empty((nh, nw, 2), np.float32)
            for y in range(nh):
                for x in range(nw):
                    ys = [min(2 * y + d, lh - 1) for d in (0, 1)]
                    xs = [min(2 * x + d, lw - 1) for d in (0, 1)]
                    blk = p[np.ix_(ys, xs)]
                    n[y, x, 0] = blk[..., 0].min()
                    n[y, x, 1] = blk[..., 1].max()
            self.levels.append(n)
        self.mips = len(self.levels)

    # ---- hybrid_terrain_traversal.wgsl:88-141 ----
    @staticmethod
    def safe_inv(d):
        ad = fmax(abs(d), f(1e-12))
        return -(ONE / ad) if d < ZERO else ONE / ad

    def curved_height(self, o, d, t, apply):
        hd2 = t * t * (d[0] * d[0] + d[2] * d[2])
        corr = hd2 * self.k if (apply and self.curv) else ZERO
        return o[1] + t * d[1] + corr

    def height_range(self, o, d, t0, t1, apply):
        y0 = self.curved_height(o, d, t0, apply)
        y1 = self.curved_height(o, d, t1, apply)
        mn = fmin(y0, y1)
        if apply and self.curv:
            a = (d[0] * d[0] + d[2] * d[2]) * self.k
            if a > ZERO:
                v = -d[1] / (f(2.0) * a)
                if t0 <= v <= t1:
                    mn = fmin(mn, self.curved_height(o, d, v, True))
        return mn, fmax(y0, y1)

    def slab(self, o, d, x0, x1, z0, z1):
        ix, iz = self.safe_inv(d[0]), self.safe_inv(d[2])
        tx0, tx1 = (x0 - o[0]) * ix, (x1 - o[0]) * ix
        if tx0 > tx1:
            tx0, tx1 = tx1, tx0
        tz0, tz1 = (z0 - o[2]) * iz, (z1 - o[2]) * iz
        if tz0 > tz1:
            tz0, tz1 = tz1, tz0
        return fmax(tx0, tz0), fmin(tx1, tz1)

    def cell_heights(self, cx, cz):                    # :149-156
        e = self.ex
        return (self.h[cz, cx] * e, self.h[cz, cx + 1] * e, self.h[cz + 1, cx] * e, self.h[cz + 1, cx + 1] * e)

    def leaf(self, o, d, tmin, tmax, cx, cz, t0, t1, apply, any_hit):      # :167-235
        h = self.cell_heights(cx, cz)
        tm = f(0.5) * (t0 + t1)
        dev = []
        for t in (t0, tm, t1):
            px = o[0] + t * d[0]
            pz = o[2] + t * d[2]
            u = clamp((px - self.ox) / self.sx - f(cx), ZERO, ONE)
            v = clamp((pz - self.oz) / self.sz - f(cz), ZERO, ONE)
            hh = mix(mix(h[0], h[1], u), mix(h[2], h[3], u), v)
            dev.append(self.curved_height(o, d, t, apply) - hh)
        c = dev[0]
        a = f(2.0) * dev[2] + f(2.0) * dev[0] - f(4.0) * dev[1]
        b = dev[2] - dev[0] - a
        s_hit = f(1e30)
        if any_hit and c <= ZERO:
            s_hit = ZERO
        elif abs(a) < f(1e-12):
            if abs(b) > f(1e-12):
                s = -c / b
                if ZERO <= s <= ONE:
                    s_hit = s
        else:
            disc = b * b - f(4.0) * a * c
            if disc >= ZERO:
                sq = np.sqrt(disc)
                q = f(-0.5) * (b + (sq if b >= ZERO else -sq))
                r0 = q / a
                r1 = f(1e30) if abs(q) < f(1e-30) else c / q
                if r0 > r1:
                    r0, r1 = r1, r0
                if ZERO <= r0 <= ONE:
                    s_hit = r0
                elif ZERO <= r1 <= ONE:
                    s_hit = r1
        if s_hit <= ONE:
            t = t0 + s_hit * (t1 - t0)
            if tmin < t < tmax:
                return t
        return None

    def normal_at(self, p, cx, cz):                    # :239-248
        h = self.cell_heights(cx, cz)
        u = clamp((p[0] - self.ox) / self.sx - f(cx), ZERO, ONE)
        v = clamp((p[2] - self.oz) / self.sz - f(cz), ZERO, ONE)
        du = mix(h[1] - h[0], h[3] - h[2], v)
        dv = mix(h[2] - h[0], h[3] - h[1], u)
        return normalize((-du / self.sx, ONE, -dv / self.sz))

    def trace(self, o, d, tmin, tmax, any_hit, apply):  # :254-372 -> (hit, t, point, normal)
        best_t, hit, point, normal = tmax, False, None, None
        stack = [(self.mips - 1, 0, 0)]
        while stack:
            level, nx, ny = stack.pop()
            cx0, cz0 = nx << level, ny << level
            if cx0 >= self.cw or cz0 >= self.ch:
                continue
            cx1 = min((nx + 1) << level, self.cw)
            cz1 = min((ny + 1) << level, self.ch)
            s0, s1 = self.slab(o, d, self.ox + f(cx0) * self.sx, self.ox + f(cx1) * self.sx,
                               self.oz + f(cz0) * self.sz, self.oz + f(cz1) * self.sz)
            t_lo = fmax(s0, tmin)
            t_hi = fmin(s1, fmin(tmax, best_t))
            if t_lo > t_hi:
                continue
            mm = self.levels[level][ny, nx]
            mlo, mhi = mm[0] * self.ex, mm[1] * self.ex
            rlo, rhi = self.height_range(o, d, t_lo, t_hi, apply)
            if rlo > mhi or rhi < mlo:
                continue
            if level == 0:
                t = self.leaf(o, d, tmin, tmax, cx0, cz0, t_lo, t_hi, apply, any_hit)
                if t is not None and t < best_t:
                    hit, best_t = True, t
                    point = (o[0] + d[0] * t, o[1] + d[1] * t, o[2] + d[2] * t)
                    normal = self.normal_at(point, cx0, cz0)
                    if any_hit:
                        return hit, best_t, point, normal
                continue
            cl = level - 1
            kids = []
            for cy in (0, 1):
                for cxi in (0, 1):
                    ccx, ccy = nx * 2 + cxi, ny * 2 + cy
                    gx0, gz0 = ccx << cl, ccy << cl
                    if gx0 >= self.cw or gz0 >= self.ch:
                        continue
                    gx1 = min((ccx + 1) << cl, self.cw)
                    gz1 = min((ccy + 1) << cl, self.ch)
                    c0, c1 = self.slab(o, d, self.ox + f(gx0) * self.sx, self.ox + f(gx1) * self.sx,
                                       self.oz + f(gz0) * self.sz, self.oz + f(gz1) * self.sz)
                    lo, hi = fmax(c0, t_lo), fmin(c1, t_hi)
                    if lo > hi:
                        continue
                    kids.append((lo, (cl, ccx, ccy)))
            # insertion sort, descending t_enter (:351-363)
            for i in range(1, len(kids)):
                key = kids[i]
                j = i
                while not (j == 0 or kids[j - 1][0] >= key[0]):
                    kids[j] = kids[j - 1]
                    j -= 1
                kids[j] = key
            for _, node in kids:
                if len(stack) < 64:
                    stack.append(node)
        return hit, best_t, point, normal


def atan_pos(x):                                       # numerics contract: Cephes atanf
    if x > f(2.414213562373095):
        y, x = f(1.5707963267948966), -(ONE / x)
    elif x > f(0.4142135623730950):
        y, x = f(0.7853981633974483), (x - ONE) / (x + ONE)
    else:
        y = ZERO
    z = x * x
    return y + ((((f(8.05374449538e-2) * z - f(1.38776856032e-1)) * z + f(1.99777106478e-1)) * z
                 - f(3.33329491539e-1)) * z * x + x)


def atan2_pinned(y, x):
    pi, half_pi = f(3.14159265358979323846), f(1.5707963267948966)
    if x == ZERO:
        return half_pi if y > ZERO else (-half_pi if y < ZERO else ZERO)
    q = y / x
    a = atan_pos(abs(q))
    if q < ZERO:
        a = -a
    if x < ZERO:
        a = a + pi if y >= ZERO else a - pi
    return a


def asin_core(x):
    a = abs(x)
    if a > f(0.5):
        z = f(0.5) * (ONE - a)
        w, flag = np.sqrt(z), True
    else:
        w, flag = a, False
        z = w * w
    p = ((((f(4.2163199048e-2) * z + f(2.4181311049e-2)) * z + f(4.5470025998e-2)) * z
          + f(7.4953002686e-2)) * z + f(1.6666752422e-1)) * z * w + w
    if flag:
        p = p + p
        p = f(1.5707963267948966) - p
    return -p if x < ZERO else p


def acos_pinned(x):
    if x < f(-0.5):
        return f(3.14159265358979323846) - f(2.0) * asin_core(np.sqrt(f(0.5) * (ONE + x)))
    if x > f(0.5):
        return f(2.0) * asin_core(np.sqrt(f(0.5) * (ONE - x)))
    return f(1.5707963267948966) - asin_core(x)


class Hybrid:
    """intersect_hybrid / intersect_hybrid_optimized / get_surface_properties / terrain_env_radiance
    (hybrid_traversal.wgsl:84-259, hybrid_terrain_traversal.wgsl:384-402) over a Scene plus an optional triangle mesh
    and equirect environment map."""

    MESH_ALBEDO = (f(0.7), f(0.7), f(0.8))

    def __init__(self, scene, mesh_vertices=None, mesh_indices=None, env_map=None):
        self.S = scene
        self.verts = None if mesh_vertices is None else np.asarray(mesh_vertices, np.float32)
        self.tris = None if mesh_indices is None else np.asarray(mesh_indices, np.uint32).reshape(-1, 3)
        self.env = None if env_map is None else np.asarray(env_map, np.float32)

    def ray_triangle(self, o, d, tmin, tmax, v0, v1, v2):          # hybrid_traversal.wgsl:84-131
        e1 = tuple(v1[i] - v0[i] for i in range(3))
        e2 = tuple(v2[i] - v0[i] for i in range(3))
        h = cross(d, e2)
        a = dot3(e1, h)
        if abs(a) < f(1e-7):
            return None
        inv = ONE / a
        s = tuple(o[i] - v0[i] for i in range(3))
        u = inv * dot3(s, h)
        if u < ZERO or u > ONE:
            return None
        q = cross(s, e1)
        v = inv * dot3(d, q)
        if v < ZERO or u + v > ONE:
            return None
        t = inv * dot3(e2, q)
        if tmin < t < tmax:
            return t, normalize(cross(e1, e2))
        return None

    def intersect_mesh(self, o, d, tmin, tmax):                     # :136-172, index-order sweep
        best_t, best = tmax, None
        if self.tris is None:
            return None
        for tri in self.tris:
            v = [tuple(f(c) for c in self.verts[int(i)]) for i in tri]
            r = self.ray_triangle(o, d, tmin, tmax, *v)
            if r is not None and r[0] < best_t:
                best_t, best = r[0], r
        if best is None:
            return None
        t, n = best
        return t, tuple(o[i] + d[i] * t for i in range(3)), n

    def closest(self, o, d, tmin, tmax):                            # intersect_hybrid :175-201
        """-> (hit, t, point, normal, hit_type) with hit_type 0 = mesh, 3 = terrain."""
        best = (False, tmax, None, None, 3)
        m = self.intersect_mesh(o, d, tmin, tmax)
        if m is not None and m[0] < best[1]:
            best = (True, m[0], m[1], m[2], 0)
        hit, t, p, n = self.S.trace(o, d, tmin, best[1], False, False)
        if hit and t < best[1]:
            best = (True, t, p, n, 3)
        return best

    def occluded(self, o, d, tmin, tmax, apply_curvature):          # intersect_hybrid_optimized :204-233 + :248-259
        best_t = tmax
        m = self.intersect_mesh(o, d, tmin, tmax)
        if m is not None:
            if m[0] < f(0.01):
                return m[0] < f(1e30)
            if m[0] < best_t:
                best_t = m[0]
        hit, t, *_ = self.S.trace(o, d, tmin, best_t, True, apply_curvature)
        if hit and t < best_t:
            best_t = t
            return best_t < f(1e30)
        return m is not None and best_t < f(1e30)

    def surface_albedo(self, hit_type):                             # get_surface_properties :238-245
        return self.S.albedo if hit_type == 3 else self.MESH_ALBEDO

    def env_radiance(self, d):                                      # terrain_env_radiance :384-402
        if self.env is None:
            return (self.S.env, self.S.env, self.S.env)
        eh, ew = self.env.shape[:2]
        d = normalize(d)
        pi = f(3.14159265358979323846)
        uu = (atan2_pinned(d[2], d[0]) / (f(2.0) * pi)) + f(0.5)
        vv = acos_pinned(clamp(d[1], f(-1.0), ONE)) / pi
        px = min(int(uu * f(ew)), ew - 1)
        py = min(int(vv * f(eh)), eh - 1)
        return tuple(f(self.env[py, px, c]) * self.S.env for c in range(3))


def cosine_dir(n, u1, u2):                            # :421-431
    sign = f(-1.0) if n[2] < ZERO else ONE
    a = f(-1.0) / (sign + n[2])
    b = n[0] * n[1] * a
    t = (ONE + sign * n[0] * n[0] * a, sign * b, -sign * n[0])
    bt = (b, sign + n[1] * n[1] * a, -n[1])
    r = np.sqrt(u1)
    phi = f(2.0) * f(3.14159265358979323846) * u2
    s, c = sincos_pinned(phi)
    lx, ly, lz = r * c, r * s, np.sqrt(fmax(ZERO, ONE - u1))
    v = tuple((t[i] * lx + bt[i] * ly) + n[i] * lz for i in range(3))
    return normalize(v)


def render(dem, width, height, cam, *, spacing, exaggeration, albedo, sun_azimuth_deg, sun_elevation_deg,
           sun_intensity, sun_color, env_intensity, spp, frames, seed, inv_two_r_prime, curvature_enabled, aovs=None,
           mesh_vertices=None, mesh_indices=None, env_map=None):
    """Fixed-frame render; returns (accum[H,W,4] float32, depth[H,W] float32 with NaN on miss); `aovs`, when given a
    dict, is filled with rgba (uint8), normal and albedo (float32 after the rgba16float round trip)."""
    S = Scene(dem, spacing, exaggeration, albedo, env_intensity, inv_two_r_prime, curvature_enabled)
    X = Hybrid(S, mesh_vertices, mesh_indices, env_map)
    W, H = width, height
    # render_terrain.rs:635-661
    origin = tuple(f(v) for v in cam["origin"])
    look = tuple(f(v) for v in cam["look_at"])
    upv = tuple(f(v) for v in cam["up"])
    fwd = normalize(tuple(look[i] - origin[i] for i in range(3)))
    right = normalize(cross(fwd, upv))
    up = normalize(cross(right, fwd))
    to_rad = f(3.14159274101257324) / f(180.0)
    az, el = f(sun_azimuth_deg) * to_rad, f(sun_elevation_deg) * to_rad
    light_dir = (_libm("cosf", az) * _libm("cosf", el), _libm("sinf", el), _libm("sinf", az) * _libm("cosf", el))
    light_color = tuple(f(sun_intensity) * f(c) for c in sun_color)
    fov = f(cam["fov_y"]) * to_rad
    half_h = _libm("tanf", f(0.5) * fov)          # WGSL tan(); the numerics contract pins it to the host libm
    half_w = (f(W) / f(H)) * half_h
    seed_hi, seed_lo = seed & 0xFFFFFFFF, (seed ^ 0x85EBCA6B) & 0xFFFFFFFF
    wi = normalize(light_dir)

    def cam_ray(gx, gy, jx, jy):                      # hybrid_terrain_traversal.wgsl:481-485
        ndc_x = ((f(gx) + f(0.5) + jx) / f(W)) * f(2.0) - ONE
        ndc_y = (ONE - (f(gy) + f(0.5) + jy) / f(H)) * f(2.0) - ONE
        rd = normalize((ndc_x * half_w, ndc_y * half_h, f(-1.0)))
        v = tuple((rd[0] * right[i] + rd[1] * up[i]) + rd[2] * (-fwd[i]) for i in range(3))
        return normalize(v)

    npx = W * H
    zero_res = dict(type=0, dir=(ZERO, ZERO, ZERO), w_sum=ZERO, m=0, weight=ZERO, tpdf=ZERO)
    prev = [dict(zero_res) for _ in range(npx)]
    accum = np.zeros((H, W, 4), np.float32)
    depth = np.full((H, W), np.nan, np.float32)
    gb_n = [None] * npx
    aov_normal = np.zeros((H, W, 3), np.float32)
    aov_albedo = np.zeros((H, W, 3), np.float32)
    for gy in range(H):                               # main_terrain_gbuffer :619-644 == the frame-0 AOV block :576-606
        for gx in range(W):
            rd = cam_ray(gx, gy, ZERO, ZERO)
            hit, t, p, n, kind = X.closest(origin, rd, f(1e-3), f(1e30))
            gb_n[gy * W + gx] = n if hit else (ZERO, ZERO, ONE)
            if hit:
                depth[gy, gx] = t
                aov_normal[gy, gx] = n
                aov_albedo[gy, gx] = X.surface_albedo(kind)
    for frame in range(frames):
        curr = [None] * npx
        for gy in range(H):                           # main_terrain :445-610
            for gx in range(W):
                pix = gy * W + gx
                pr = prev[pix]
                if pr["m"] > 512:
                    scale = f(512.0) / f(pr["m"])
                    pr["w_sum"] = pr["w_sum"] * scale
                    pr["m"] = 512
                    if pr["tpdf"] > ZERO:
                        pr["weight"] = pr["w_sum"] / (f(pr["m"]) * pr["tpdf"])
                prev_valid = frame > 0 and pr["m"] > 0 and pr["weight"] > ZERO and pr["tpdf"] > ZERO and pr["type"] == 1
                rng = Rng(seed_hi ^ (gx * 1664525) ^ (gy * 1013904223) ^ (frame * 92837111) ^ seed_lo)
                fr = (ZERO, ZERO, ZERO)
                cand = dict(zero_res)
                for _ in range(spp):
                    jx = tent(rng.next()) * f(0.5)
                    jy = tent(rng.next()) * f(0.5)
                    rd = cam_ray(gx, gy, jx, jy)
                    hit, t, p, n, kind = X.closest(origin, rd, f(1e-3), f(1e30))
                    if not hit:
                        sky = X.env_radiance(rd)
                        fr = tuple(fr[i] + sky[i] for i in range(3))
                        continue
                    alb = X.surface_albedo(kind)
                    ndotl = fmax(dot3(n, wi), ZERO)
                    tp = lum(tuple(alb[i] * light_color[i] * ndotl for i in range(3)))
                    if tp > ZERO:
                        cand["type"] = 1
                        cand["dir"] = wi
                        cand["w_sum"] = cand["w_sum"] + tp
                        cand["m"] += 1
                        cand["tpdf"] = tp
                    sun_dir, reuse_w = wi, ONE
                    if prev_valid:
                        sun_dir = normalize(pr["dir"])
                        reuse_w = clamp(pr["weight"], ZERO, f(4.0))
                    sun = (ZERO, ZERO, ZERO)
                    nd = fmax(dot3(n, sun_dir), ZERO)
                    so = tuple(p[i] + n[i] * f(1e-3) for i in range(3))
                    if nd > ZERO:
                        occ = X.occluded(so, sun_dir, f(1e-3), f(1e30), True)
                        vis = ZERO if occ else ONE
                        sun = tuple(alb[i] * light_color[i] * nd * vis * reuse_w for i in range(3))
                    u1 = rng.next()
                    u2 = rng.next()
                    ei = cosine_dir(n, u1, u2)
                    occ = X.occluded(so, ei, f(1e-3), f(1e30), False)
                    ev = ZERO if occ else ONE
                    env_rgb = X.env_radiance(ei)
                    ibl = tuple(alb[i] * env_rgb[i] * ev for i in range(3))
                    fr = tuple(fr[i] + sun[i] + ibl[i] for i in range(3))
                fr = tuple(fr[i] / f(spp) for i in range(3))
                if cand["m"] > 0 and cand["w_sum"] > ZERO and cand["tpdf"] > ZERO:
                    cand["weight"] = cand["w_sum"] / (f(cand["m"]) * cand["tpdf"])
                curr[pix] = cand
                for i in range(3):
                    accum[gy, gx, i] = accum[gy, gx, i] + fr[i]
                accum[gy, gx, 3] = accum[gy, gx, 3] + ONE
        out = [None] * npx
        for i in range(npx):                          # pt_restir_temporal.wgsl:54-109
            rp, rc = prev[i], curr[i]
            pv = rp["m"] > 0 and rp["weight"] > ZERO and rp["tpdf"] > ZERO
            cv = rc["m"] > 0 and rc["weight"] > ZERO and rc["tpdf"] > ZERO
            if not pv:
                out[i] = dict(rc)
            elif not cv:
                out[i] = dict(rp)
            else:
                src = rp if rp["weight"] > rc["weight"] else rc
                ro = dict(type=src["type"], dir=src["dir"], tpdf=src["tpdf"], m=rp["m"] + rc["m"], w_sum=rp["w_sum"] + rc["w_sum"], weight=ZERO)
                if ro["w_sum"] > ZERO and ro["tpdf"] > ZERO:
                    ro["weight"] = ro["w_sum"] / (f(ro["m"]) * ro["tpdf"])
                out[i] = ro
        new_prev = [None] * npx
        for i in range(npx):                          # pt_restir_spatial.wgsl:158-224 (+ :45-117)
            x, y = i % W, i // W
            rng = Rng(((seed_hi ^ frame) + i * 1664525 + 1013904223) & 0xFFFFFFFF)
            rs = out[i]
            chosen_type, chosen_dir, chosen_pdf = rs["type"], rs["dir"], rs["tpdf"]
            st = dict(W=ZERO)
            m_total = 0
            N = normalize(gb_n[i])

            def consider(r):
                nonlocal chosen_type, chosen_dir, chosen_pdf
                if r["m"] == 0 or r["type"] != 1:
                    return
                if fmax(dot3(N, normalize(r["dir"])), ZERO) <= ZERO:
                    return
                p_curr = fmax(ONE, ZERO) / fmax(ONE, f(1e-8))
                if p_curr <= ZERO or r["tpdf"] <= ZERO:
                    return
                w = r["w_sum"] * (p_curr / fmax(r["tpdf"], f(1e-6)))
                if w <= ZERO:
                    return
                st["W"] = st["W"] + w
                if rng.next() < w / st["W"]:
                    chosen_type, chosen_dir, chosen_pdf = r["type"], r["dir"], p_curr

            consider(rs)
            m_total += rs["m"]
            for _ in range(8):
                rx = int(math.floor(rng.next() * f(7.0))) - 3
                ry = int(math.floor(rng.next() * f(7.0))) - 3
                if rx == 0 and ry == 0:
                    continue
                nxp = min(max(x + rx, 0), W - 1)
                nyp = min(max(y + ry, 0), H - 1)
                rn = out[nyp * W + nxp]
                consider(rn)
                m_total += rn["m"]
            o = dict(type=chosen_type, dir=chosen_dir, tpdf=chosen_pdf, w_sum=st["W"], m=m_total, weight=ZERO)
            if o["w_sum"] > ZERO and o["tpdf"] > ZERO:
                o["weight"] = o["w_sum"] / (f(o["m"]) * o["tpdf"])
            new_prev[i] = o
        prev = new_prev
    if aovs is not None:
        # resolve :570-573 + hybrid_kernel.wgsl:109-112 into rgba16float, read back as u8 (render_terrain.rs:1355-1366)
        exposure = fmin(fmax(f(cam.get("exposure", 1.0)), ZERO), f(65504.0))
        rgba = np.zeros((H, W, 4), np.uint8)
        for gy in range(H):
            for gx in range(W):
                for c in range(3):
                    mean = accum[gy, gx, c] / accum[gy, gx, 3]
                    exposed = mean * exposure
                    ldr = exposed / (ONE + exposed)
                    v = f(np.float16(ldr))
                    rgba[gy, gx, c] = int(clamp(v, ZERO, ONE) * f(255.0) + f(0.5))
                rgba[gy, gx, 3] = 255
        aovs["rgba"] = rgba
        aovs["normal"] = aov_normal.astype(np.float16).astype(np.float32)     # rgba16float AOV textures
        aovs["albedo"] = aov_albedo.astype(np.float16).astype(np.float32)
    return accum, depth
