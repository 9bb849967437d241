"""Second, independent restatement of the AETHER post (TEST INFRASTRUCTURE): written from the WGSL text
(/root/reference/src/shaders/atmosphere/prometheus_aerial.wgsl, evaluation_core.wgsl) in numpy-f32 scalars, one WGSL
statement per Python statement, without looking at oracle/f3d_aether_oracle.c's structure.  tests/test_aether.py requires
the C oracle to agree with it bit for bit (RGBA16F texels), which pins the oracle's arithmetic to the shader text under
the numerics contract (exp2 = the pinned Cephes kernel, round = half-to-even, clamp = min(max()))."""
import numpy as np

f = np.float32
WAVELENGTHS = [380.0, 420.0, 460.0, 500.0, 540.0, 580.0, 620.0, 660.0, 700.0, 740.0, 780.0]
CIE = [(0.001368, 0.000039, 0.006450), (0.134380, 0.004000, 0.645600), (0.290800, 0.060000, 1.669200),
       (0.004900, 0.323000, 0.272000), (0.290400, 0.954000, 0.020300), (0.916300, 0.870000, 0.001650),
       (0.854450, 0.381000, 0.000190), (0.164900, 0.061000, 0.000000), (0.011359, 0.004102, 0.000000),
       (0.000690, 0.000249, 0.000000), (0.000042, 0.000015, 0.000000)]


def exp2(x):
    x = f(x)
    if np.isnan(x):
        return x
    if x >= f(128.0):
        return f(np.inf)
    if x < f(-126.0):
        return f(0.0)
    px = f(np.floor(x))
    i0 = int(px)
    fr = f(x - px)
    if fr > f(0.5):
        i0 += 1
        fr = f(fr - f(1.0))
    p = f(1.535336188319500e-4)
    for c in (1.339887440266574e-3, 9.618437357674640e-3, 5.550332471162809e-2, 2.402264791363012e-1, 6.931472028550421e-1):
        p = f(f(p * fr) + f(c))
    r = f(f(1.0) + f(fr * p))
    e1 = i0 >> 1
    e2 = i0 - e1
    return f(f(r * f(2.0 ** e1)) * f(2.0 ** e2))


def det_exp(x):
    return exp2(f(f(x) * f(1.4426950408889634)))


def clamp(x, lo, hi):
    return f(min(max(f(x), f(lo)), f(hi)))


def dot3(a, b):
    return f(f(f(a[0] * b[0]) + f(a[1] * b[1])) + f(a[2] * b[2]))


def normalize(v):
    inv = f(f(1.0) / f(np.sqrt(dot3(v, v))))
    return (f(v[0] * inv), f(v[1] * inv), f(v[2] * inv))


def h2f(bits):
    return f(np.array(bits, np.uint16).view(np.float16))


def f2h(v):
    with np.errstate(over="ignore"):
        return int(np.array(f(v)).astype(np.float16).view(np.uint16))


def clamp_hdr(c):
    return tuple(f(min(max(f(x), f(0.0)), f(65504.0))) for x in c)


def spectral_xyz(i, rc, mc, oc, turbidity):
    lam = f(WAVELENGTHS[i])
    ratio = f(f(550.0) / lam)
    r2 = f(ratio * ratio)
    rayleigh_beta = f(f(f(1.2989e-5) * r2) * r2)
    mie_beta = f(f(f(1.0e-5) * f(turbidity)) * ratio)
    d = f(f(lam - f(600.0)) / f(85.0))
    ozone_beta = f(f(1.2e-6) * det_exp(f(f(f(-0.5) * d) * d)))
    tau = f(f(f(rayleigh_beta * rc) + f(mie_beta * mc)) + f(ozone_beta * oc))
    t = det_exp(f(-max(tau, f(0.0))))
    w = f(0.5) if (i == 0 or i + 1 == 11) else f(1.0)
    return tuple(f(f(f(c) * t) * w) for c in CIE[i])


def mu_to_unit(mu):
    b = clamp(mu, -1.0, 1.0)
    m = f(np.sqrt(abs(b)))
    s = m if b >= f(0.0) else f(-m)
    return f(f(0.5) * f(s + f(1.0)))


def nu_to_unit(nu):
    return f(f(1.0) - f(np.sqrt(max(f(f(0.5) * f(f(1.0) - clamp(nu, -1.0, 1.0))), f(0.0)))))


class Luts:
    """handle: .transmittance (h, mu, 4), .scattering (h*nu, sun, view, 4), .aerial (h, mu, dist, 4) uint16 + .config"""

    def __init__(self, handle):
        self.t = np.asarray(handle.transmittance)
        self.s = np.asarray(handle.scattering)
        self.a = np.asarray(handle.aerial)
        self.cfg = handle.config


def sample_scattering(L, height_unit, mu_sun, mu_view, nu):
    dz, dy, dx = L.s.shape[:3]
    hc = max(int(L.cfg.dimensions.scattering_height), 2)
    nc = max(int(L.cfg.dimensions.scattering_nu), 2)
    coords = [f(mu_to_unit(mu_view) * f(dx - 1)), f(mu_to_unit(mu_sun) * f(dy - 1)),
              f(f(np.sqrt(clamp(height_unit, 0.0, 1.0))) * f(hc - 1)), f(nu_to_unit(nu) * f(nc - 1))]
    lims = [dx - 1, dy - 1, hc - 1, nc - 1]
    lower = [int(np.floor(c)) for c in coords]
    upper = [min(lo + 1, lim) for lo, lim in zip(lower, lims)]
    frac = [f(c - f(np.floor(c))) for c in coords]
    acc = [f(0.0)] * 4
    for hs in (0, 1):
        for ns in (0, 1):
            for ss in (0, 1):
                for vs in (0, 1):
                    vi = upper[0] if vs else lower[0]
                    si = upper[1] if ss else lower[1]
                    hi = upper[2] if hs else lower[2]
                    ni = upper[3] if ns else lower[3]
                    w = f(f(f((frac[0] if vs else f(f(1.0) - frac[0])) * (frac[1] if ss else f(f(1.0) - frac[1])))
                            * (frac[2] if hs else f(f(1.0) - frac[2]))) * (frac[3] if ns else f(f(1.0) - frac[3])))
                    x = min(max(vi, 0), dx - 1)
                    y = min(max(si, 0), dy - 1)
                    z = min(max(hi * nc + ni, 0), dz - 1)
                    tex = h2f(L.s[z, y, x])
                    acc = [f(a + f(w * t)) for a, t in zip(acc, tex)]
    return tuple(f(max(a, f(0.0))) for a in acc[:3])


def radius_m(camera_h, view_mu, dist, bottom):
    r = f(max(f(bottom), f(1.0)) + clamp(camera_h, 0.0, 100000.0))
    bd = clamp(dist, 0.0, 20000000.0)
    sq = f(f(f(r * r) + f(bd * bd)) + f(f(f(f(2.0) * r) * bd) * clamp(view_mu, -1.0, 1.0)))
    return f(np.sqrt(max(sq, f(0.0))))


def altitude(camera_h, view_mu, dist, bottom):
    return clamp(f(radius_m(camera_h, view_mu, dist, bottom) - max(f(bottom), f(1.0))), 0.0, 100000.0)


def endpoint_mus(camera_h, view_mu, sun_mu, nu, dist, bottom):
    r = f(max(f(bottom), f(1.0)) + clamp(camera_h, 0.0, 100000.0))
    bd = clamp(dist, 0.0, 20000000.0)
    er = f(max(radius_m(camera_h, view_mu, bd, bottom), f(1.0)))
    ev = f(f(f(r * clamp(view_mu, -1.0, 1.0)) + bd) / er)
    es = f(f(f(r * clamp(sun_mu, -1.0, 1.0)) + f(bd * clamp(nu, -1.0, 1.0))) / er)
    return clamp(ev, -1.0, 1.0), clamp(es, -1.0, 1.0)


def segment_transmittance(dist, camera_h, view_mu, bottom, density_scale, turbidity, ozone_du):
    bd = clamp(dist, 0.0, 20000000.0)
    bh = clamp(camera_h, 0.0, 100000.0)
    fractions = [0.03125, 0.09375, 0.15625, 0.21875, 0.28125, 0.34375, 0.40625, 0.46875,
                 0.53125, 0.59375, 0.65625, 0.71875, 0.78125, 0.84375, 0.90625, 0.96875]
    hs = [altitude(bh, view_mu, f(bd * f(fr)), bottom) for fr in fractions]
    ray = mie = oz = None
    for h in hs:
        a = det_exp(f(f(-h) / f(8000.0)))
        b = det_exp(f(f(-h) / f(1200.0)))
        c = f(max(f(f(1.0) - abs(f(f(h - f(25000.0)) / f(15000.0)))), f(0.0)))
        ray = a if ray is None else f(ray + a)
        mie = b if mie is None else f(mie + b)
        oz = c if oz is None else f(oz + c)
    pps = f(f(bd * f(density_scale)) * f(0.0625))
    rc = f(pps * ray)
    mc = f(pps * mie)
    oc = f(f(f(pps * oz) * f(ozone_du)) / f(300.0))
    xyz = spectral_xyz(0, rc, mc, oc, turbidity)
    for i in range(1, 11):
        s = spectral_xyz(i, rc, mc, oc, turbidity)
        xyz = tuple(f(a + b) for a, b in zip(xyz, s))
    rgb = (f(dot3((f(3.2404542), f(-1.5371385), f(-0.4985314)), xyz) / f(3.2613921)),
           f(dot3((f(-0.9692660), f(1.8760108), f(0.0415560)), xyz) / f(2.5069624)),
           f(dot3((f(0.0556434), f(-0.2040259), f(1.0572252)), xyz) / f(2.3679786)))
    return tuple(clamp(c, 0.0, 1.0) for c in rgb)


def rint(x):
    return int(np.rint(f(x)))   # numpy rint = round half to even, as WGSL round()


def boundary_transmittance(L, height_unit, mu):
    dy, dx = L.t.shape[:2]
    x = rint(f(f(f(0.5) * f(clamp(mu, -1.0, 1.0) + f(1.0))) * f(max(dx, 1) - 1)))
    y = rint(f(clamp(height_unit, 0.0, 1.0) * f(max(dy, 1) - 1)))
    return tuple(clamp(c, 0.0, 1.0) for c in h2f(L.t[y, x])[:3])


def aerial_transmittance(L, distance_unit, height_unit, mu_view):
    dz, dy, dx = L.a.shape[:3]
    x = rint(f(clamp(distance_unit, 0.0, 1.0) * f(max(dx, 1) - 1)))
    y = rint(f(f(f(0.5) * f(clamp(mu_view, -1.0, 1.0) + f(1.0))) * f(max(dy, 1) - 1)))
    z = rint(f(clamp(height_unit, 0.0, 1.0) * f(max(dz, 1) - 1)))
    return clamp(h2f(L.a[z, y, x])[3], 0.0, 1.0)


def reinhard(c):
    return tuple(f(x / f(f(1.0) + x)) for x in c)


def main_pixel(L, view, gx, gy, accumulated, depth, visibility):
    """prometheus_aerial.wgsl `main`; view: dict(width,height,cam_origin,cam_right,cam_up,cam_forward,tan_half_fov,aspect,
    exposure,light_dir,sun_intensity).  Returns 4 RGBA16F bit patterns."""
    cfg = L.cfg
    W, Hh = view["width"], view["height"]
    den = f(max(f(accumulated[3]), f(1.0)))
    surface = clamp_hdr(tuple(f(f(accumulated[c]) / den) for c in range(3)))
    ndc_x = f(f(f(f(f(gx) + f(0.5)) / f(W)) * f(2.0)) - f(1.0))
    ndc_y = f(f(f(f(1.0) - f(f(f(gy) + f(0.5)) / f(Hh))) * f(2.0)) - f(1.0))
    tan_half = f(view["tan_half_fov"])
    aspect = f(view["aspect"])
    sx = f(f(ndc_x * tan_half) * aspect)
    sy = f(ndc_y * tan_half)
    R, U, F = (tuple(f(c) for c in view[k]) for k in ("cam_right", "cam_up", "cam_forward"))
    ray = normalize(tuple(f(f(f(R[c] * sx) + f(U[c] * sy)) + F[c]) for c in range(3)))
    sun_dir = normalize(tuple(f(c) for c in view["light_dir"]))
    sun_intensity = clamp(view["sun_intensity"], 0.0, 65504.0)
    exposure = clamp(view["exposure"], 0.0, 65504.0)
    atmosphere_height = f(max(f(f(cfg.top_radius_m) - f(cfg.bottom_radius_m)), f(1.0)))
    camera_height = f(max(f(view["cam_origin"][1]), f(0.0)))
    camera_height_unit = clamp(f(camera_height / atmosphere_height), 0.0, 1.0)
    if f(visibility) < f(0.5):
        sc = sample_scattering(L, camera_height_unit, sun_dir[1], ray[1], dot3(ray, sun_dir))
        miss = clamp_hdr(tuple(f(c * sun_intensity) for c in sc))
        ldr = reinhard(tuple(f(c * exposure) for c in miss))
        return [f2h(c) for c in ldr] + [f2h(1.0)]
    bottom = f(cfg.bottom_radius_m)
    endpoint_height = altitude(camera_height, ray[1], depth, bottom)
    nu = dot3(ray, sun_dir)
    ev, es = endpoint_mus(camera_height, ray[1], sun_dir[1], nu, depth, bottom)
    analytic = segment_transmittance(depth, camera_height, ray[1], bottom, 1.0, cfg.turbidity, cfg.ozone_du)
    boundary = boundary_transmittance(L, camera_height_unit, ray[1])
    cam_sc = tuple(f(c * sun_intensity) for c in sample_scattering(L, camera_height_unit, sun_dir[1], ray[1], nu))
    endpoint_height_unit = clamp(f(endpoint_height / atmosphere_height), 0.0, 1.0)
    end_sc = tuple(f(c * sun_intensity) for c in sample_scattering(L, endpoint_height_unit, es, ev, nu))
    distance_unit = f(f(depth) / f(max(f(cfg.max_aerial_distance_m), f(1.0))))
    aerial_mean = aerial_transmittance(L, distance_unit, camera_height_unit, ray[1])
    analytic_mean = dot3(analytic, (f(0.2126), f(0.7152), f(0.0722)))
    k = f(aerial_mean / f(max(analytic_mean, f(1.0e-6))))
    tr = tuple(f(max(clamp(f(a * k), 0.0, 1.0), b)) for a, b in zip(analytic, boundary))
    fin = tuple(f(max(f(c - f(t * e)), f(0.0))) for c, t, e in zip(cam_sc, tr, end_sc))
    hdr = clamp_hdr(tuple(f(f(s * t) + i) for s, t, i in zip(surface, tr, fin)))
    ldr = reinhard(tuple(f(c * exposure) for c in hdr))
    return [f2h(c) for c in ldr] + [f2h(1.0)]
