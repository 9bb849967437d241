"""Second, independent restatement of the HELIOS viewshed / shadow-mask shader (TEST INFRASTRUCTURE): written from the WGSL text
(/root/reference/src/shaders/terrain_viewshed.wgsl and the det_* helpers of src/shaders/includes/determinism.wgsl) in numpy-f32
scalars, statement by statement.  tests/test_viewshed.py requires the C oracle to agree with it bit for bit, which pins the
oracle's operation order to the shader text (driver-defined operations pinned as in DESIGN.md section 4: IEEE `/` and sqrt,
atan2 of the azimuth = the Cephes kernel of the C oracle, degrees/radians as one multiply)."""
import numpy as np

f = np.float32
u32 = np.uint32
INVALID = 0xFFFFFFFF


def bits(x):
    return int(np.array(f(x)).view(np.uint32))


def from_bits(b):
    return f(np.array(b & 0xFFFFFFFF, np.uint32).view(np.float32))


def det_fma(a, b, c):
    p = f(f(a) * f(b))
    return f(p + f(c))


def det_mix(a, b, t):
    d = f(f(b) - f(a))
    s = f(d * f(t))
    return f(f(a) + s)


def det_inverse_sqrt(x):
    xc = f(max(f(x), f(1.17549435e-38)))
    y = from_bits(0x5F3759DF - (bits(xc) >> 1))
    half_x = f(f(0.5) * xc)
    for _ in range(3):
        y = f(y * f(f(1.5) - f(half_x * f(y * y))))
    return y


def det_rcp(x):
    ax = f(abs(f(x)))
    y = from_bits(0x7EF311C3 - bits(ax))
    for _ in range(3):
        y = f(y * f(f(2.0) - f(ax * y)))
    return f(-y) if f(x) < 0 else y


def det_div(a, b):
    return f(f(a) * det_rcp(b))


def det_sqrt(x):
    r = f(f(x) * det_inverse_sqrt(x))
    return f(0.0) if f(x) <= 0 else r


def det_sin(x):
    x = f(x)
    k = f(np.floor(f(f(x * f(0.6366197723675814)) + f(0.5))))
    r = f(x - f(k * f(1.5707963267948966)))
    q = int(k) & 3
    r2 = f(r * r)
    ps = det_fma(r2, -0.00019840874, 0.0083333310)
    ps = det_fma(r2, ps, -0.16666667)
    ps = det_fma(r2, ps, 1.0)
    s = f(r * ps)
    pc = det_fma(r2, -0.0013888378, 0.041666638)
    pc = det_fma(r2, pc, -0.5)
    pc = det_fma(r2, pc, 1.0)
    v = pc if (q & 1) == 1 else s
    return f(-v) if (q & 2) == 2 else v


def det_cos(x):
    return det_sin(f(f(x) + f(1.5707963267948966)))


def det_atan01(a):
    s = f(f(a) * f(a))
    p = det_fma(s, -0.0117212, 0.05265332)
    p = det_fma(s, p, -0.11643287)
    p = det_fma(s, p, 0.19354346)
    p = det_fma(s, p, -0.33262347)
    p = det_fma(s, p, 0.99997726)
    return f(f(a) * p)


def det_atan2(y, x):
    ax, ay = f(abs(f(x))), f(abs(f(y)))
    hi = max(ax, ay)
    if hi == 0:
        return f(0.0)
    lo = min(ax, ay)
    p = det_atan01(f(lo / hi))
    if ay > ax:
        p = f(f(1.5707963267948966) - p)
    if f(x) < 0:
        p = f(f(3.141592653589793) - p)
    return f(-p) if f(y) < 0 else p


def det_acos(x):
    xc = f(min(max(f(x), f(-1.0)), f(1.0)))
    x2 = f(xc * xc)
    s = f(np.sqrt(f(max(f(f(1.0) - x2), f(0.0)))))
    return det_atan2(s, xc)


def degrees(x):
    return f(f(x) * f(57.295779513082323))


def radians(x):
    return f(f(x) * f(0.017453292519943295))


class Scene:
    def __init__(self, heights, opts, physics, minmax_levels, atan2_native):
        self.h = np.ascontiguousarray(heights, np.float32)
        self.H, self.W = self.h.shape
        self.observer = [f(opts[k]) for k in ("observer_x", "observer_y", "observer_height_m", "target_height_m")]
        self.metric = [f(opts[k]) for k in ("max_distance_m", "longitude_step_deg", "latitude_step_deg", "geodesic_sphere_radius_m")]
        self.physics = [f(p) for p in physics]
        self.geodetic = [f(opts[k]) for k in ("observer_latitude_rad", "observer_longitude_rad", "left_unwrapped_deg", "top_deg")]
        self.levels = minmax_levels                      # list of (h, w, 2) arrays, finest first
        self.atan2_native = atan2_native


def height_at(S, px, py):
    x = f(min(max(f(px), f(0.0)), f(S.W - 1)))
    y = f(min(max(f(py), f(0.0)), f(S.H - 1)))
    x0, y0 = int(np.floor(x)), int(np.floor(y))
    x1, y1 = min(x0 + 1, S.W - 1), min(y0 + 1, S.H - 1)
    fx, fy = f(x - f(x0)), f(y - f(y0))
    return det_mix(det_mix(S.h[y0, x0], S.h[y0, x1], fx), det_mix(S.h[y1, x0], S.h[y1, x1], fx), fy)


def safe_inv(d):
    m = f(max(f(abs(f(d))), f(1e-12)))
    return f(f(-1.0) / m) if f(d) < 0 else f(f(1.0) / m)


def slab_xz(origin, direction, x0, x1, z0, z1):
    ix, iz = safe_inv(direction[0]), safe_inv(direction[1])
    tx0, tx1 = f(f(f(x0) - origin[0]) * ix), f(f(f(x1) - origin[0]) * ix)
    if tx0 > tx1:
        tx0, tx1 = tx1, tx0
    tz0, tz1 = f(f(f(z0) - origin[1]) * iz), f(f(f(z1) - origin[1]) * iz)
    if tz0 > tz1:
        tz0, tz1 = tz1, tz0
    return max(tx0, tz0), min(tx1, tz1)


def pack_node(level, x, y):
    return (level << 26) | (y << 13) | x


def height_limit(d, c):
    return det_fma(c[2], f(f(d) * f(d)), det_fma(c[1], d, c[0]))


def height_limit_range(d0, d1, c):
    h0, h1 = height_limit(d0, c), height_limit(d1, c)
    minimum = min(h0, h1)
    if c[2] > 0:
        vertex = f(f(-c[1]) / f(f(2.0) * c[2]))
        if min(f(d0), f(d1)) <= vertex <= max(f(d0), f(d1)):
            minimum = min(minimum, height_limit(vertex, c))
    return minimum, max(h0, h1)


def leaf_deviation(origin, direction, heights, cx, cy, t, d0, d1, c):
    px, py = f(origin[0] + f(f(t) * direction[0])), f(origin[1] + f(f(t) * direction[1]))
    u = f(min(max(f(px - f(cx)), f(0.0)), f(1.0)))
    v = f(min(max(f(py - f(cy)), f(0.0)), f(1.0)))
    th = det_mix(det_mix(heights[0], heights[1], u), det_mix(heights[2], heights[3], u), v)
    return f(th - height_limit(det_mix(d0, d1, t), c))


def leaf_occluded(S, origin, direction, cx, cy, t0, t1, d0, d1, c, tol):
    heights = (S.h[cy, cx], S.h[cy, cx + 1], S.h[cy + 1, cx], S.h[cy + 1, cx + 1])
    tm = f(f(0.5) * f(f(t0) + f(t1)))
    e = [leaf_deviation(origin, direction, heights, cx, cy, t, d0, d1, c) for t in (t0, tm, t1)]
    quadratic = f(f(f(f(2.0) * e[2]) + f(f(2.0) * e[0])) - f(f(4.0) * e[1]))
    linear = f(f(e[2] - e[0]) - quadratic)
    maximum = max(e[0], e[2])
    if abs(quadratic) > f(1e-12):
        vertex = f(f(-linear) / f(f(2.0) * quadratic))
        if f(0.0) < vertex < f(1.0):
            maximum = max(maximum, det_fma(quadratic, f(vertex * vertex), det_fma(linear, vertex, e[0])))
    return maximum > f(tol)


def select_child(S, origin, direction, plevel, px, py, after_t, after_id):
    cw, ch, cl = S.W - 1, S.H - 1, plevel - 1
    best_id, best_t = INVALID, f(2.0)
    for ci in range(4):
        nx, ny = px * 2 + (ci & 1), py * 2 + (ci >> 1)
        x0, y0 = nx << cl, ny << cl
        if x0 >= cw or y0 >= ch:
            continue
        x1, y1 = min((nx + 1) << cl, cw), min((ny + 1) << cl, ch)
        se, sx = slab_xz(origin, direction, x0, x1, y0, y1)
        entry, exit_t = max(se, f(0.0)), min(sx, f(1.0))
        if entry > exit_t:
            continue
        nid = pack_node(cl, nx, ny)
        follows = after_id == INVALID or entry > after_t or (entry == after_t and nid > after_id)
        if follows and (best_id == INVALID or entry < best_t or (entry == best_t and nid < best_id)):
            best_id, best_t = nid, entry
    return best_id


def trace_segment(S, origin, endpoint, d0, d1, c, tol):
    direction = (f(endpoint[0] - origin[0]), f(endpoint[1] - origin[1]))
    cw, ch = S.W - 1, S.H - 1
    root = len(S.levels) - 1
    node = pack_node(root, 0, 0)
    while True:
        level, ny, nx = node >> 26, (node >> 13) & 0x1FFF, node & 0x1FFF
        x0, y0 = nx << level, ny << level
        descend = False
        if x0 < cw and y0 < ch:
            x1, y1 = min((nx + 1) << level, cw), min((ny + 1) << level, ch)
            se, sx = slab_xz(origin, direction, x0, x1, y0, y1)
            t0, t1 = max(se, f(0.0)), min(sx, f(1.0))
            if t0 <= t1:
                hmin, _ = height_limit_range(det_mix(d0, d1, t0), det_mix(d0, d1, t1), c)
                mmax = S.levels[level][ny, nx, 1]
                if f(hmin + f(tol)) < mmax:
                    if level == 0:
                        if leaf_occluded(S, origin, direction, x0, y0, t0, t1, d0, d1, c, tol):
                            return True
                    else:
                        child = select_child(S, origin, direction, level, nx, ny, f(0.0), INVALID)
                        if child != INVALID:
                            node, descend = child, True
        if descend:
            continue
        cl, cx, cy = level, nx, ny
        advanced = False
        while cl < root:
            pl, ppx, ppy = cl + 1, cx >> 1, cy >> 1
            q0, r0 = cx << cl, cy << cl
            q1, r1 = min((cx + 1) << cl, cw), min((cy + 1) << cl, ch)
            se, _ = slab_xz(origin, direction, q0, q1, r0, r1)
            sib = select_child(S, origin, direction, pl, ppx, ppy, max(se, f(0.0)), pack_node(cl, cx, cy))
            if sib != INVALID:
                node, advanced = sib, True
                break
            cl, cx, cy = pl, ppx, ppy
        if not advanced:
            return False


def latlon_to_pixel(S, lat, lon):
    lon_deg = degrees(lon)
    if lon_deg < S.geodetic[2]:
        lon_deg = f(lon_deg + f(360.0))
    if lon_deg > f(S.geodetic[2] + f(180.0)):
        lon_deg = f(lon_deg - f(360.0))
    return (f(det_div(f(lon_deg - S.geodetic[2]), S.metric[1]) - f(0.5)),
            f(det_div(f(S.geodetic[3] - degrees(lat)), S.metric[2]) - f(0.5)))


def geodesic_sample_pixel(S, lat0, lon0, azimuth, distance_m):
    if S.metric[3] > 0:
        ad = det_div(distance_m, S.metric[3])
        sin_lat = det_fma(det_sin(lat0), det_cos(ad), f(f(det_sin(ad) * det_cos(lat0)) * det_cos(azimuth)))
        lat = f(f(1.5707963267948966) - det_acos(f(min(max(sin_lat, f(-1.0)), f(1.0)))))
        lon = f(f(lon0) + det_atan2(f(f(det_sin(azimuth) * det_sin(ad)) * det_cos(lat0)),
                                    f(det_cos(ad) - f(det_sin(lat0) * det_sin(lat)))))
        return latlon_to_pixel(S, lat, lon)
    fl = f(f(1.0) / f(298.257223563))
    a = f(6378137.0)
    b = f(a * f(f(1.0) - fl))
    reduced = det_atan2(f(f(f(1.0) - fl) * det_sin(lat0)), det_cos(lat0))
    su1, cu1 = det_sin(reduced), det_cos(reduced)
    saz, caz = det_sin(azimuth), det_cos(azimuth)
    sigma1 = det_atan2(su1, f(cu1 * caz))
    sin_alpha = f(cu1 * saz)
    c2a = f(f(1.0) - f(sin_alpha * sin_alpha))
    usq = f(f(c2a * f(f(a * a) - f(b * b))) / f(b * b))
    A = f(f(1.0) + f(f(usq / f(16384.0)) * f(f(4096.0) + f(usq * f(f(-768.0) + f(usq * f(f(320.0) - f(f(175.0) * usq))))))))
    B = f(f(usq / f(1024.0)) * f(f(256.0) + f(usq * f(f(-128.0) + f(usq * f(f(74.0) - f(f(47.0) * usq)))))))
    sigma = det_div(distance_m, f(b * A))
    for _ in range(4):
        tsm = f(f(f(2.0) * sigma1) + sigma)
        ss, cs, c2 = det_sin(sigma), det_cos(sigma), det_cos(tsm)
        inner = f(f(cs * f(f(-1.0) + f(f(f(2.0) * c2) * c2))) -
                  f(f(f(f(B / f(6.0)) * c2) * f(f(-3.0) + f(f(f(4.0) * ss) * ss))) * f(f(-3.0) + f(f(f(4.0) * c2) * c2))))
        delta = f(f(B * ss) * f(c2 + f(f(B / f(4.0)) * inner)))
        sigma = f(det_div(distance_m, f(b * A)) + delta)
    ss, cs = det_sin(sigma), det_cos(sigma)
    tsm = f(f(f(2.0) * sigma1) + sigma)
    tmp = f(f(su1 * ss) - f(f(cu1 * cs) * caz))
    lat = det_atan2(f(f(su1 * cs) + f(f(cu1 * ss) * caz)), f(f(f(1.0) - fl) * det_sqrt(f(f(sin_alpha * sin_alpha) + f(tmp * tmp)))))
    lam = det_atan2(f(ss * saz), f(f(cu1 * cs) - f(f(su1 * ss) * caz)))
    C = f(f(f(fl / f(16.0)) * c2a) * f(f(4.0) + f(fl * f(f(4.0) - f(f(3.0) * c2a)))))
    c2 = det_cos(tsm)
    dlon = f(lam - f(f(f(f(f(1.0) - C) * fl) * sin_alpha) *
                     f(sigma + f(f(C * ss) * f(c2 + f(f(C * cs) * f(f(-1.0) + f(f(f(2.0) * c2) * c2))))))))
    return latlon_to_pixel(S, lat, f(f(lon0) + dlon))


def local_inverse_radius(S, lat, azimuth):
    if S.physics[3] == 0:
        return f(0.0)
    if S.metric[3] > 0:
        return f(f(1.0) / S.metric[3])
    a, e2 = f(6378137.0), f(0.0066943799901413165)
    sl = det_sin(lat)
    w = det_sqrt(f(f(1.0) - f(f(e2 * sl) * sl)))
    meridional = det_div(f(a * f(f(1.0) - e2)), f(f(w * w) * w))
    prime_vertical = det_div(a, w)
    sa, ca = det_sin(azimuth), det_cos(azimuth)
    return f(det_div(f(ca * ca), meridional) + det_div(f(sa * sa), prime_vertical))


def shadow_step_m(S, lat, azimuth):
    sl = det_sin(lat)
    ft = f(f(1.0) - f(f(f(0.0066943799901413165) * sl) * sl))
    root = det_sqrt(ft)
    meridional = det_div(f(f(6378137.0) * f(f(1.0) - f(0.0066943799901413165))), f(ft * root))
    prime_vertical = det_div(f(6378137.0), root)
    hm = S.metric[3] if S.metric[3] > 0 else meridional
    hp = S.metric[3] if S.metric[3] > 0 else prime_vertical
    north_cell = f(hm * radians(S.metric[2]))
    east_cell = f(f(hp * det_cos(lat)) * radians(S.metric[1]))
    east_cross = det_div(east_cell, f(max(f(abs(det_sin(azimuth))), f(1e-6))))
    north_cross = det_div(north_cell, f(max(f(abs(det_cos(azimuth))), f(1e-6))))
    return f(max(f(0.1), f(f(0.5) * min(north_cross, east_cross))))


def viewshed_cell(S, positions, x, y):
    """main: returns (visible 0/1/2, drop, gain, horizon)."""
    mx, my = f(positions[y, x, 0]), f(positions[y, x, 1])
    d2 = f(f(mx * mx) + f(my * my))
    distance_m = f(np.sqrt(d2))
    azimuth = S.atan2_native(mx, my)
    inv_radius = f(0.0)
    if d2 != 0:
        e2, n2 = f(f(mx * mx) / d2), f(f(my * my) / d2)
        inv_radius = f(f(n2 * S.physics[0]) + f(e2 * S.physics[1]))
    vacuum_drop = f(f(f(f(0.5) * inv_radius) * distance_m) * distance_m)
    effective_drop = f(vacuum_drop * S.physics[2])
    gain = f(vacuum_drop - effective_drop)
    observer_elevation = f(height_at(S, S.observer[0], S.observer[1]) + S.observer[2])
    target_abs = f(S.h[y, x] + S.observer[3])
    horizon = S.metric[0]
    if inv_radius > 0:
        eff = f(inv_radius * S.physics[2])
        horizon = f(f(np.sqrt(f(f(f(2.0) * max(observer_elevation, f(0.0))) / eff))) + f(np.sqrt(f(f(f(2.0) * max(target_abs, f(0.0))) / eff))))
    if distance_m == 0:
        return 1, f(0.0), f(0.0), horizon
    if distance_m > S.metric[0]:
        return 0, vacuum_drop, gain, horizon
    target_elevation = f(target_abs - effective_drop)
    c = (observer_elevation, det_div(f(target_elevation - observer_elevation), distance_m), f(f(f(0.5) * inv_radius) * S.physics[2]))
    visible = 1
    start_d, sp = f(0.0), (S.observer[0], S.observer[1])
    while True:
        seg_lat = radians(det_fma(f(-f(sp[1] + f(0.5))), S.metric[2], S.geodetic[3]))
        end_d = min(f(start_d + shadow_step_m(S, seg_lat, azimuth)), distance_m)
        ep = geodesic_sample_pixel(S, S.geodetic[0], S.geodetic[1], azimuth, end_d)
        if ep[0] < f(-0.5) or ep[1] < f(-0.5) or ep[0] > f(f(S.W) - f(0.5)) or ep[1] > f(f(S.H) - f(0.5)):
            visible = 2
            break
        if trace_segment(S, sp, ep, start_d, end_d, c, 0.001):
            visible = 0
            break
        if end_d >= distance_m:
            break
        start_d, sp = end_d, ep
    return visible, vacuum_drop, gain, horizon


def shadow_cell(S, inputs, x, y):
    """shadow_mask_main: True = lit."""
    lat0, lon0, azimuth, elevation = (f(v) for v in inputs[y, x])
    if elevation <= 0:
        return False
    slope = det_div(det_sin(elevation), det_cos(elevation))
    eff = f(local_inverse_radius(S, lat0, azimuth) * S.physics[2])
    c = (S.h[y, x], slope, f(f(0.5) * eff))
    start_d, sp, seg_lat = f(0.0), (f(x), f(y)), lat0
    while True:
        end_d = min(f(start_d + shadow_step_m(S, seg_lat, azimuth)), S.metric[0])
        ep = geodesic_sample_pixel(S, lat0, lon0, azimuth, end_d)
        if trace_segment(S, sp, ep, start_d, end_d, c, 0.01):
            return False
        if ep[0] < f(-0.5) or ep[1] < f(-0.5) or ep[0] > f(f(S.W) - f(0.5)) or ep[1] > f(f(S.H) - f(0.5)):
            break
        if end_d >= S.metric[0]:
            break
        seg_lat = radians(det_fma(f(-f(ep[1] + f(0.5))), S.metric[2], S.geodetic[3]))
        start_d, sp = end_d, ep
    return True
