"""Shared fixtures for the parity tests: the reference's locked golden scene, its drift-gate
metric, and the conservative-descent KAT of terrain_heightfield.rs (written independently in
numpy/f64).  Citations are into /root/reference."""
from __future__ import annotations

from pathlib import Path

import numpy as np

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"

# Locked scene: tests/test_hybrid_terrain_pt.py:31-79
SIZE = 256
SPAN = 100.0
RELIEF = 20.0
CAM = {"origin": (0.0, 35.0, 90.0), "look_at": (0.0, 5.0, 0.0), "up": (0.0, 1.0, 0.0),
       "fov_y": 45.0, "exposure": 1.0}
ALBEDO = (0.55, 0.52, 0.48)


def golden_dem() -> np.ndarray:
    return np.load(GOLDEN_DIR / "mini_dem_128.npy")


def golden_png() -> np.ndarray:
    from PIL import Image

    return np.array(Image.open(GOLDEN_DIR / "mini_dem_reference.png"))


def scene_kwargs(dem: np.ndarray) -> dict:
    spacing = SPAN / (dem.shape[1] - 1)
    return dict(spacing=(spacing, spacing), exaggeration=RELIEF, albedo=ALBEDO, sun_azimuth_deg=225.0,
                sun_elevation_deg=35.0, sun_intensity=2.5, env_intensity=0.35, max_frames=512,
                min_frames=32, variance_threshold=1e-3, seed=7)


def sine_dem(n: int = 128) -> np.ndarray:
    """BASELINE.json configs[0]: the 128x128 procedural sine-wave DEM."""
    x = np.arange(n, dtype=np.float64)
    h = 0.5 + 0.25 * np.sin(2 * np.pi * 3 * x[None, :] / (n - 1)) + 0.25 * np.sin(2 * np.pi * 2 * x[:, None] / (n - 1))
    return h.astype(np.float32)


def rainier_dem(n: int = 2048) -> np.ndarray:
    """SURVEY.md section 8d C2 'Rainier-shaped' closed-form DEM (metres), float32."""
    s = 256.0 / n
    x = np.arange(n, dtype=np.float64)[None, :]
    y = np.arange(n, dtype=np.float64)[:, None]
    cx, cy = 0.586 * n, 0.492 * n
    h = (900.0 + 180.0 * np.sin(0.071 * x * s) + 120.0 * np.cos(0.047 * y * s)
         + 3400.0 * np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (0.18 * n) ** 2))
    for k in range(5):
        h = h + 220.0 * 2.0 ** (-k) * np.sin(2.0 ** k * (0.113 * x * s) + 1.7 * k) * np.cos(2.0 ** k * (0.089 * y * s) - 0.9 * k)
    return h.astype(np.float32)


def rainier_camera(n: int, spacing: float, dem: np.ndarray) -> dict:
    """Orbit camera of SURVEY.md section 8d C2: phi=28 deg, theta=49 deg, radius 1.2*extent, fov 42."""
    phi, theta = np.radians(28.0), np.radians(49.0)
    radius = 1.2 * n * spacing
    target = np.array([0.0, float(dem.mean()), 0.0])
    origin = target + radius * np.array([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)])
    return {"origin": tuple(float(v) for v in origin), "look_at": tuple(float(v) for v in target),
            "up": (0.0, 1.0, 0.0), "fov_y": 42.0, "exposure": 1.0}


# --- reference drift-gate metric: tests/_ssim.py (Gaussian 11x11, sigma 1.5, zero padding) ---
def ssim(img1: np.ndarray, img2: np.ndarray, data_range: float = 255.0) -> float:
    from scipy.ndimage import convolve

    a = img1.astype(np.float64)
    b = img2.astype(np.float64)
    if a.ndim == 3:
        return float(np.mean([ssim(a[..., c], b[..., c], data_range) for c in range(a.shape[2])]))
    coords = np.arange(11, dtype=np.float64) - 5.0
    g = np.exp(-0.5 * (coords / 1.5) ** 2)
    g /= g.sum()
    win = np.outer(g, g)
    f = lambda im: convolve(im, win, mode="constant", cval=0.0)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu1, mu2 = f(a), f(b)
    s1 = f(a * a) - mu1 * mu1
    s2 = f(b * b) - mu2 * mu2
    s12 = f(a * b) - mu1 * mu2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s1 + s2 + c2))
    return float(m.mean())


# --- conservative-descent KAT: terrain_heightfield.rs:614-629,780-821,875-904,2081-2116 ---
PROOF_SIDE = 256
PROOF_CELLS = 255
PROOF_SPACING = 500.0
PROOF_INV_TWO_R = np.float32(1.0 / 14_650_000.0)
PROOF_TMAX = 200_000.0


def curvature_fixture() -> np.ndarray:
    i = np.arange(PROOF_SIDE * PROOF_SIDE)
    x = (i % PROOF_SIDE).astype(np.float32)
    y = (i // PROOF_SIDE).astype(np.float32)
    h = (np.float32(900.0) + np.float32(180.0) * np.sin(x * np.float32(0.071))
         + np.float32(120.0) * np.cos(y * np.float32(0.047))
         + np.float32(650.0) * np.exp(-((x - np.float32(150.0)) ** 2 + (y - np.float32(126.0)) ** 2) / np.float32(900.0)))
    return h.astype(np.float32).reshape(PROOF_SIDE, PROOF_SIDE)


def _xorshift_stream(state: int, count: int) -> np.ndarray:
    out = np.empty(count, dtype=np.uint64)
    s = state
    for k in range(count):
        s ^= (s << 13) & 0xFFFFFFFF
        s ^= s >> 17
        s ^= (s << 5) & 0xFFFFFFFF
        out[k] = s
    return out


def _proof_height(h: np.ndarray, cx, cz, x, z):
    u = x / PROOF_SPACING - cx
    v = z / PROOF_SPACING - cz
    h64 = h.astype(np.float64)
    h00, h10 = h64[cz, cx], h64[cz, cx + 1]
    h01, h11 = h64[cz + 1, cx], h64[cz + 1, cx + 1]
    return (h00 * (1 - u) + h10 * u) * (1 - v) + (h01 * (1 - u) + h11 * u) * v


def proof_rays(h: np.ndarray, x, z, az, el) -> np.ndarray:
    """proof_ray (:882-904) for arrays of float32 x, z (cell units), azimuth, elevation -> (n, 8) rays."""
    x = np.asarray(x, np.float32); z = np.asarray(z, np.float32)
    az = np.asarray(az, np.float32); el = np.asarray(el, np.float32)
    cx = np.floor(x).astype(np.int64); cz = np.floor(z).astype(np.int64)
    ox = (x * np.float32(PROOF_SPACING)).astype(np.float32)
    oz = (z * np.float32(PROOF_SPACING)).astype(np.float32)
    surf = _proof_height(h, cx, cz, ox.astype(np.float64), oz.astype(np.float64)).astype(np.float32)
    hor = np.cos(el).astype(np.float32)
    rays = np.zeros((x.size, 8), np.float32)
    rays[:, 0] = ox; rays[:, 1] = surf + np.float32(1.7); rays[:, 2] = oz; rays[:, 3] = 1e-3
    rays[:, 4] = hor * np.cos(az).astype(np.float32); rays[:, 5] = np.sin(el).astype(np.float32)
    rays[:, 6] = hor * np.sin(az).astype(np.float32); rays[:, 7] = PROOF_TMAX
    return rays


def kat_rays(h: np.ndarray):
    """The 10 000 xorshift rays (state 0x48454c49) + the 255x255 shadow-mask rays of the KAT."""
    s = _xorshift_stream(0x48454C49, 60_000)
    f32 = np.float32
    umax = f32(4294967295.0)
    s = s.reshape(10_000, 6)
    x = f32(1.0) + (s[:, 0] % 253).astype(f32) + (s[:, 1].astype(f32) / umax) * f32(0.999)
    z = f32(1.0) + (s[:, 2] % 253).astype(f32) + (s[:, 3].astype(f32) / umax) * f32(0.999)
    az = (s[:, 4].astype(f32) / umax) * f32(2.0 * np.pi)
    el = (f32(0.1) + (s[:, 5] % 790).astype(f32) / f32(100.0)) * f32(np.pi / 180.0)
    arb = proof_rays(h, x, z, az, el)
    gz, gx = np.meshgrid(np.arange(PROOF_CELLS), np.arange(PROOF_CELLS), indexing="ij")
    n = gx.size
    mask = proof_rays(h, gx.ravel().astype(f32) + f32(0.5), gz.ravel().astype(f32) + f32(0.5),
                      np.full(n, np.radians(f32(37.0)), f32), np.full(n, np.radians(f32(0.6)), f32))
    return arb, mask


def _root_in_span(a, b, c, t0, t1):
    """root_in_span (:670-687), vectorised f64."""
    lin = np.abs(a) < 1e-15
    with np.errstate(divide="ignore", invalid="ignore"):
        rl = -c / b
        lin_hit = (np.abs(b) >= 1e-15) & (rl >= t0) & (rl <= t1)
        disc = b * b - 4 * a * c
        sq = np.sqrt(np.maximum(disc, 0.0))
        q = -0.5 * (b + np.copysign(sq, b))
        first = q / a
        second = np.where(np.abs(q) < 1e-30, np.inf, c / q)
    quad_hit = (disc >= 0) & (((first >= t0) & (first <= t1)) | ((second >= t0) & (second <= t1)))
    return np.where(lin, lin_hit, quad_hit)


def brute_2d_hit(h: np.ndarray, rays: np.ndarray) -> np.ndarray:
    """Independent f64 cell-by-cell DDA oracle (brute_2d_hit, :780-821), vectorised over rays."""
    r = rays.astype(np.float64)
    o = r[:, 0:3]; d = r[:, 4:7]
    inv2r = float(PROOF_INV_TWO_R)
    extent = PROOF_CELLS * PROOF_SPACING
    n = r.shape[0]

    def axis(o_, d_):
        par = np.abs(d_) < 1e-12
        with np.errstate(divide="ignore", invalid="ignore"):
            a = (0.0 - o_) / d_
            b = (extent - o_) / d_
        lo = np.where(par, -np.inf, np.minimum(a, b))
        hi = np.where(par, np.inf, np.maximum(a, b))
        ok = np.where(par, (o_ >= 0.0) & (o_ <= extent), True)
        return lo, hi, ok

    lx, hx, okx = axis(o[:, 0], d[:, 0])
    lz, hz, okz = axis(o[:, 2], d[:, 2])
    enter = np.maximum(lx, lz); exit_ = np.minimum(hx, hz)
    alive = okx & okz & (enter <= exit_)
    t = np.maximum(enter, 1e-3)
    end = np.minimum(exit_, PROOF_TMAX)
    alive &= t <= end
    hit = np.zeros(n, bool)
    h64 = h.astype(np.float64)
    hs = d[:, 0] ** 2 + d[:, 2] ** 2
    for _ in range(4 * PROOF_SIDE):
        idx = np.nonzero(alive)[0]
        if idx.size == 0:
            break
        ti, ei = t[idx], end[idx]
        probe = np.minimum(ti + 1e-5, ei)
        x = o[idx, 0] + probe * d[idx, 0]
        z = o[idx, 2] + probe * d[idx, 2]
        cx = np.clip(np.floor(x / PROOF_SPACING), 0, 254).astype(np.int64)
        cz = np.clip(np.floor(z / PROOF_SPACING), 0, 254).astype(np.int64)
        dx, dz = d[idx, 0], d[idx, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            nx = np.where(dx > 0, ((cx + 1) * PROOF_SPACING - o[idx, 0]) / dx,
                          np.where(dx < 0, (cx * PROOF_SPACING - o[idx, 0]) / dx, np.inf))
            nz = np.where(dz > 0, ((cz + 1) * PROOF_SPACING - o[idx, 2]) / dz,
                          np.where(dz < 0, (cz * PROOF_SPACING - o[idx, 2]) / dz, np.inf))
        nxt = np.minimum(np.minimum(nx, nz), ei)
        # brute_cell_hit (:691-720)
        h00 = h64[cz, cx]
        hx_ = h64[cz, cx + 1] - h00
        hz_ = h64[cz + 1, cx] - h00
        hxz = h64[cz + 1, cx + 1] - h00 - hx_ - hz_
        u0 = o[idx, 0] / PROOF_SPACING - cx
        v0 = o[idx, 2] / PROOF_SPACING - cz
        du = dx / PROOF_SPACING; dv = dz / PROOF_SPACING
        ta = hxz * du * dv
        tb = hx_ * du + hz_ * dv + hxz * (u0 * dv + v0 * du)
        tc = h00 + hx_ * u0 + hz_ * v0 + hxz * u0 * v0
        got = _root_in_span(hs[idx] * inv2r - ta, d[idx, 1] - tb, o[idx, 1] - tc, ti, nxt)
        hit[idx[got]] = True
        done = got | (nxt >= ei)
        alive[idx[done]] = False
        t[idx] = nxt + 1e-7
        alive[idx[t[idx] > ei]] = False
    return hit


def _first_root_in_span(a, b, c, t0, t1):
    """Smallest root of a t^2 + b t + c in [t0, t1] (f64, vectorised); NaN where there is none."""
    with np.errstate(divide="ignore", invalid="ignore"):
        lin = np.abs(a) < 1e-18
        rl = np.where(np.abs(b) > 0, -c / b, np.nan)
        disc = b * b - 4 * a * c
        sq = np.sqrt(np.maximum(disc, 0.0))
        q = -0.5 * (b + np.copysign(sq, b))
        r0 = np.where(disc >= 0, q / a, np.nan)
        r1 = np.where((disc >= 0) & (np.abs(q) > 0), c / q, np.nan)
    lo = np.fmin(r0, r1); hi = np.fmax(r0, r1)
    in0 = (lo >= t0) & (lo <= t1); in1 = (hi >= t0) & (hi <= t1)
    quad = np.where(in0, lo, np.where(in1, hi, np.nan))
    linr = np.where((rl >= t0) & (rl <= t1), rl, np.nan)
    return np.where(lin, linr, quad)


def brute_first_hit_t(h: np.ndarray, rays: np.ndarray, spacing: float = PROOF_SPACING) -> np.ndarray:
    """Independent f64 closest-hit distance of straight rays (no curvature) against the bilinear heightfield with
    texel (0,0) at the world origin: cell-by-cell DDA + exact patch quadratic.  NaN = miss."""
    r = rays.astype(np.float64)
    o = r[:, 0:3]; d = r[:, 4:7]; tmin = r[:, 3]; tmax = r[:, 7]
    hh, ww = h.shape
    ex, ez = (ww - 1) * spacing, (hh - 1) * spacing
    n = r.shape[0]

    def axis(o_, d_, extent):
        par = np.abs(d_) < 1e-300
        with np.errstate(divide="ignore", invalid="ignore"):
            a = (0.0 - o_) / d_; b = (extent - o_) / d_
        return (np.where(par, -np.inf, np.minimum(a, b)), np.where(par, np.inf, np.maximum(a, b)),
                np.where(par, (o_ >= 0.0) & (o_ <= extent), True))

    lx, hx, okx = axis(o[:, 0], d[:, 0], ex)
    lz, hz, okz = axis(o[:, 2], d[:, 2], ez)
    t = np.maximum(np.maximum(lx, lz), tmin)
    end = np.minimum(np.minimum(hx, hz), tmax)
    alive = okx & okz & (t <= end)
    out = np.full(n, np.nan)
    h64 = h.astype(np.float64)
    for _ in range(4 * max(hh, ww)):
        idx = np.nonzero(alive)[0]
        if idx.size == 0:
            break
        ti, ei = t[idx], end[idx]
        probe = np.minimum(ti + 1e-7 * np.maximum(1.0, np.abs(ti)), ei)
        x = o[idx, 0] + probe * d[idx, 0]; z = o[idx, 2] + probe * d[idx, 2]
        cx = np.clip(np.floor(x / spacing), 0, ww - 2).astype(np.int64)
        cz = np.clip(np.floor(z / spacing), 0, hh - 2).astype(np.int64)
        dx, dz = d[idx, 0], d[idx, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            nx = np.where(dx > 0, ((cx + 1) * spacing - o[idx, 0]) / dx, np.where(dx < 0, (cx * spacing - o[idx, 0]) / dx, np.inf))
            nz = np.where(dz > 0, ((cz + 1) * spacing - o[idx, 2]) / dz, np.where(dz < 0, (cz * spacing - o[idx, 2]) / dz, np.inf))
        nxt = np.minimum(np.minimum(nx, nz), ei)
        h00 = h64[cz, cx]; hx_ = h64[cz, cx + 1] - h00; hz_ = h64[cz + 1, cx] - h00
        hxz = h64[cz + 1, cx + 1] - h00 - hx_ - hz_
        u0 = o[idx, 0] / spacing - cx; v0 = o[idx, 2] / spacing - cz
        du = dx / spacing; dv = dz / spacing
        ta = hxz * du * dv
        tb = hx_ * du + hz_ * dv + hxz * (u0 * dv + v0 * du)
        tc = h00 + hx_ * u0 + hz_ * v0 + hxz * u0 * v0
        root = _first_root_in_span(-ta, d[idx, 1] - tb, o[idx, 1] - tc, ti, nxt)
        got = np.isfinite(root)
        out[idx[got]] = root[got]
        done = got | (nxt >= ei)
        alive[idx[done]] = False
        t[idx] = nxt
    return out
