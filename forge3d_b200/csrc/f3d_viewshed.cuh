// forge3d_b200/csrc/f3d_viewshed.cuh
// HELIOS viewshed and solar shadow mask: any-hit descent of the same packed-node min-max
// chain as the path tracer, walked along geodesic chords in DEM-pixel space.  Replaces the WGSL entry points
//   /root/reference/src/shaders/terrain_viewshed.wgsl:470-560  main              (viewshed)
//   /root/reference/src/shaders/terrain_viewshed.wgsl:608-680  shadow_mask_main  (terrain-to-sun visibility)
// and everything they call (:24-468, :562-606) plus the det_* helpers they use
// (/root/reference/src/shaders/includes/determinism.wgsl:120-211,359-420).
//
// One thread per DEM cell.  Data layout is the path tracer's: `cells` holds the four corner heights of a DEM cell as
// one float4 (terrain_cell_heights :112-118 = one 128-bit load instead of four texel fetches), the min-max chain is the
// plain per-level [min,max] array built by k_build_level0 / k_reduce_level (exact min/max, so equal to the CPU
// build_minmax_mips the reference uploads, viewshed.rs:207-213).  The stackless ascend/descend order of
// terrain_trace_segment (:235-362, written that way for FXC) is kept as is: it needs no per-thread stack at all.
//
// Numerics: the det_* helpers are the reference's own software pins (bit-trick seeds + Newton steps, barriered mul/add,
// polynomial trig) restated literally; what WGSL still leaves to the driver is pinned as in DESIGN.md section 4
// (IEEE `/` and sqrt, atan2 of the target azimuth = atan2_pinned, dot/length left to right, degrees/radians as one
// multiply, no FMA contraction).
#pragma once
#include "f3d_math.cuh"

namespace f3d {

struct ViewshedParams {
    uint32_t w, h;                     // uniforms.dimensions.xy (texels)
    float observer[4];                 // x, y (pixels), observer height, target height
    float metric[4];                   // max distance, lon step deg, lat step deg, geodesic sphere radius (0 = WGS84)
    float physics[4];                  // 1/meridional, 1/prime-vertical, 1 - k, curved flag
    float geodetic[4];                 // observer lat, lon (rad), left (unwrapped deg), top (deg)
    float observer_elevation;          // height_at(observer.xy) + observer.z, evaluated once on the host
    const float* heights;              // raw DEM, row-major
    const float4* cells;               // (h00, h10, h01, h11) per cell
    const float2* mm[16];              // [min,max] per level, pitch mm_pitch[l]
    uint32_t mm_pitch[16];
    uint32_t root_level;
};

// ---- determinism.wgsl ------------------------------------------------------------------------------------------
__device__ __forceinline__ float vdet_fma(float a, float b, float c) { const float p = a * b; return p + c; }              // :133
__device__ __forceinline__ float vdet_mix(float a, float b, float t) { const float d = b - a; const float s = d * t; return a + s; }   // :143
__device__ __forceinline__ float vdet_inverse_sqrt(float x) {                                                             // :185
    const float xc = fmaxf(x, 1.17549435e-38f);
    float y = __uint_as_float(0x5f3759dfu - (__float_as_uint(xc) >> 1));
    const float half_x = 0.5f * xc;
#pragma unroll
    for (int i = 0; i < 3; i++) { const float yy = y * y; const float hh = half_x * yy; y = y * (1.5f - hh); }
    return y;
}
__device__ __forceinline__ float vdet_rcp(float x) {                                                                      // :195
    const float ax = fabsf(x);
    float y = __uint_as_float(0x7EF311C3u - __float_as_uint(ax));
#pragma unroll
    for (int i = 0; i < 3; i++) { const float p = ax * y; y = y * (2.0f - p); }
    return x < 0.0f ? -y : y;
}
__device__ __forceinline__ float vdet_div(float a, float b) { return a * vdet_rcp(b); }                                   // :204
__device__ __forceinline__ float vdet_sqrt(float x) { const float r = x * vdet_inverse_sqrt(x); return x <= 0.0f ? 0.0f : r; }   // :208
__device__ __forceinline__ float vdet_sin(float x) {                                                                      // :360
    const float k = floorf(x * 0.6366197723675814f + 0.5f);
    const float kp = k * 1.5707963267948966f;
    const float r = x - kp;
    const int q = (int)k & 3;
    const float r2 = r * r;
    float ps = vdet_fma(r2, -0.00019840874f, 0.0083333310f);
    ps = vdet_fma(r2, ps, -0.16666667f);
    ps = vdet_fma(r2, ps, 1.0f);
    const float s = r * ps;
    float pc = vdet_fma(r2, -0.0013888378f, 0.041666638f);
    pc = vdet_fma(r2, pc, -0.5f);
    pc = vdet_fma(r2, pc, 1.0f);
    const float v = (q & 1) == 1 ? pc : s;
    return (q & 2) == 2 ? -v : v;
}
__device__ __forceinline__ float vdet_cos(float x) { return vdet_sin(x + 1.5707963267948966f); }                          // :383
__device__ __forceinline__ float vdet_atan2(float y, float x) {                                                           // :388-412
    const float ax = fabsf(x), ay = fabsf(y);
    const float hi = fmaxf(ax, ay);
    if (hi == 0.0f) return 0.0f;
    const float a = fdiv(fminf(ax, ay), hi);
    const float s = a * a;
    float p = vdet_fma(s, -0.0117212f, 0.05265332f);
    p = vdet_fma(s, p, -0.11643287f);
    p = vdet_fma(s, p, 0.19354346f);
    p = vdet_fma(s, p, -0.33262347f);
    p = vdet_fma(s, p, 0.99997726f);
    p = a * p;
    p = ay > ax ? 1.5707963267948966f - p : p;
    p = x < 0.0f ? 3.141592653589793f - p : p;
    return y < 0.0f ? -p : p;
}
__device__ __forceinline__ float vdet_acos(float x) {                                                                     // :416
    const float xc = fminf(fmaxf(x, -1.0f), 1.0f);
    const float x2 = xc * xc;
    return vdet_atan2(fsqrt(fmaxf(1.0f - x2, 0.0f)), xc);
}
__device__ __forceinline__ float vdegrees(float r) { return r * 57.295779513082323f; }
__device__ __forceinline__ float vradians(float d) { return d * 0.017453292519943295f; }

// ---- terrain_viewshed.wgsl --------------------------------------------------------------------------------------
__device__ __forceinline__ float vs_safe_inv(float d) {                                        // :55
    const float m = fmaxf(fabsf(d), 1e-12f);
    return d < 0.0f ? fdiv(-1.0f, m) : fdiv(1.0f, m);
}
// terrain_slab_xz (:60-77) with the ray's two inverses hoisted out (they depend on the chord only)
__device__ __forceinline__ void vs_slab(float ox, float oy, float ix, float iz, float x0, float x1, float z0, float z1, float& te, float& tx) {
    float tx0 = (x0 - ox) * ix, tx1 = (x1 - ox) * ix;
    if (tx0 > tx1) { const float t = tx0; tx0 = tx1; tx1 = t; }
    float tz0 = (z0 - oy) * iz, tz1 = (z1 - oy) * iz;
    if (tz0 > tz1) { const float t = tz0; tz0 = tz1; tz1 = t; }
    te = fmaxf(tx0, tz0);
    tx = fminf(tx1, tz1);
}
__device__ __forceinline__ uint32_t vs_pack(uint32_t level, uint32_t x, uint32_t y) { return (level << 26) | (y << 13) | x; }
__device__ __forceinline__ float vs_height_limit(float d, const float c[3]) { return vdet_fma(c[2], d * d, vdet_fma(c[1], d, c[0])); }   // :83
__device__ __forceinline__ float vs_height_limit_min(float d0, float d1, const float c[3]) {   // :93-110 (only .x is consumed)
    const float h0 = vs_height_limit(d0, c), h1 = vs_height_limit(d1, c);
    float minimum = fminf(h0, h1);
    if (c[2] > 0.0f) {
        const float vertex = fdiv(-c[1], 2.0f * c[2]);
        if (vertex >= fminf(d0, d1) && vertex <= fmaxf(d0, d1)) minimum = fminf(minimum, vs_height_limit(vertex, c));
    }
    return minimum;
}
__device__ __forceinline__ float vs_leaf_deviation(float ox, float oy, float dx, float dy, float4 hts, uint32_t cx, uint32_t cy, float st,
                                                   float d0, float d1, const float c[3]) {     // :120-141
    const float px = ox + st * dx, py = oy + st * dy;
    const float u = fminf(fmaxf(px - (float)cx, 0.0f), 1.0f), v = fminf(fmaxf(py - (float)cy, 0.0f), 1.0f);
    const float th = vdet_mix(vdet_mix(hts.x, hts.y, u), vdet_mix(hts.z, hts.w, u), v);
    return th - vs_height_limit(vdet_mix(d0, d1, st), c);
}
__device__ __forceinline__ bool vs_leaf_occluded(const ViewshedParams& P, float ox, float oy, float dx, float dy, uint32_t cx, uint32_t cy,
                                                 float t0, float t1, float d0, float d1, const float c[3], float tol) {   // :143-181
    const float4 hts = __ldg(P.cells + (size_t)cy * (P.w - 1u) + cx);
    const float tm = 0.5f * (t0 + t1);
    const float e0 = vs_leaf_deviation(ox, oy, dx, dy, hts, cx, cy, t0, d0, d1, c), e1 = vs_leaf_deviation(ox, oy, dx, dy, hts, cx, cy, tm, d0, d1, c),
                e2 = vs_leaf_deviation(ox, oy, dx, dy, hts, cx, cy, t1, d0, d1, c);
    const float quadratic = 2.0f * e2 + 2.0f * e0 - 4.0f * e1;
    const float linear = e2 - e0 - quadratic;
    float maximum = fmaxf(e0, e2);
    if (fabsf(quadratic) > 1e-12f) {
        const float vertex = fdiv(-linear, 2.0f * quadratic);
        if (vertex > 0.0f && vertex < 1.0f) maximum = fmaxf(maximum, vdet_fma(quadratic, vertex * vertex, vdet_fma(linear, vertex, e0)));
    }
    return maximum > tol;
}
constexpr uint32_t kVsInvalid = 0xFFFFFFFFu;
// terrain_select_child, :186-229: the next intersecting child of (plevel, px, py) after (after_t, after_id) in (entry, id) order
__device__ __forceinline__ uint32_t vs_select_child(const ViewshedParams& P, float ox, float oy, float ix, float iz, uint32_t plevel,
                                                    uint32_t px, uint32_t py, float after_t, uint32_t after_id) {
    const uint32_t cw = P.w - 1u, ch = P.h - 1u, cl = plevel - 1u;
    uint32_t best_id = kVsInvalid;
    float best_t = 2.0f;
#pragma unroll
    for (uint32_t ci = 0; ci < 4u; ci++) {
        const uint32_t nx = px * 2u + (ci & 1u), ny = py * 2u + (ci >> 1);
        const uint32_t x0 = nx << cl, y0 = ny << cl;
        if (x0 >= cw || y0 >= ch) continue;
        const uint32_t x1 = min((nx + 1u) << cl, cw), y1 = min((ny + 1u) << cl, ch);
        float se, sx;
        vs_slab(ox, oy, ix, iz, (float)x0, (float)x1, (float)y0, (float)y1, se, sx);
        const float entry = fmaxf(se, 0.0f), exit_t = fminf(sx, 1.0f);
        if (entry > exit_t) continue;
        const uint32_t id = vs_pack(cl, nx, ny);
        const bool follows = after_id == kVsInvalid || entry > after_t || (entry == after_t && id > after_id);
        if (follows && (best_id == kVsInvalid || entry < best_t || (entry == best_t && id < best_id))) { best_id = id; best_t = entry; }
    }
    return best_id;
}
// terrain_trace_segment, :235-362
__device__ __forceinline__ bool vs_trace_segment(const ViewshedParams& P, float ox, float oy, float ex, float ey, float d0, float d1,
                                                 const float c[3], float tol) {
    const float dx = ex - ox, dy = ey - oy;
    const float ix = vs_safe_inv(dx), iz = vs_safe_inv(dy);
    const uint32_t cw = P.w - 1u, ch = P.h - 1u;
    // START BELOW THE ROOT.  The reference walks every segment down from the root (:247), ten levels for a half-cell chord.
    // A node that contains the whole segment with a margin yields t0 = 0 and t1 = 1 EXACTLY (its entry planes lie behind the
    // origin: sign-exact products <= 0; its exit planes lie >= 1/256 cell beyond the end point: products >= 1 after rounding),
    // so every such ancestor evaluates the SAME hmin and compares it with a maximum that only grows towards the root: the
    // lowest containing node passes its height test iff all its ancestors do, the root-to-there descent is forced (no other
    // node meets the segment), and nothing outside that node's subtree can be visited.  Starting there - and ending the
    // sibling search there - therefore returns the reference's flag.  A segment that touches the DEM border (clipped nodes)
    // or straddles a high-level split keeps the walk from the root.
    uint32_t root = P.root_level;
    uint32_t node = vs_pack(root, 0u, 0u);
    {
        const float m = 0.00390625f;
        const float lx = fminf(ox, ex) - m, hx = fmaxf(ox, ex) + m, ly = fminf(oy, ey) - m, hy = fmaxf(oy, ey) + m;
        if (lx >= 0.0f && ly >= 0.0f && hx <= (float)cw && hy <= (float)ch) {
            const uint32_t ax = (uint32_t)lx, bx = (uint32_t)hx, ay = (uint32_t)ly, by = (uint32_t)hy;   // truncation == floor (>= 0)
            const uint32_t diff = (ax ^ bx) | (ay ^ by);
            const uint32_t lv = 32u - (uint32_t)__clz((int)diff);             // 0 when both ends share a cell
            // the node's far planes must not be the clipped DEM edge (a clipped plane has no margin argument)
            if (lv < root && (((ax >> lv) + 1u) << lv) <= cw && (((ay >> lv) + 1u) << lv) <= ch) {
                root = lv;
                node = vs_pack(lv, ax >> lv, ay >> lv);
            }
        }
    }
    while (true) {
        const uint32_t level = node >> 26, ny = (node >> 13) & 0x1FFFu, nx = node & 0x1FFFu;
        const uint32_t x0 = nx << level, y0 = ny << level;
        bool descend = false;
        if (x0 < cw && y0 < ch) {
            const uint32_t x1 = min((nx + 1u) << level, cw), y1 = min((ny + 1u) << level, ch);
            float se, sx;
            vs_slab(ox, oy, ix, iz, (float)x0, (float)x1, (float)y0, (float)y1, se, sx);
            const float t0 = fmaxf(se, 0.0f), t1 = fminf(sx, 1.0f);
            if (t0 <= t1) {
                const float hmin = vs_height_limit_min(vdet_mix(d0, d1, t0), vdet_mix(d0, d1, t1), c);
                const float mmax = __ldg(P.mm[level] + (size_t)ny * P.mm_pitch[level] + nx).y;
                if (hmin + tol < mmax) {
                    if (level == 0u) {
                        if (vs_leaf_occluded(P, ox, oy, dx, dy, x0, y0, t0, t1, d0, d1, c, tol)) return true;
                    } else {
                        const uint32_t cid = vs_select_child(P, ox, oy, ix, iz, level, nx, ny, 0.0f, kVsInvalid);
                        if (cid != kVsInvalid) { node = cid; descend = true; }
                    }
                }
            }
        }
        if (descend) continue;
        uint32_t cl = level, cx = nx, cy = ny;
        bool advanced = false;
        while (cl < root) {
            const uint32_t pl = cl + 1u, ppx = cx >> 1, ppy = cy >> 1;
            const uint32_t q0 = cx << cl, r0 = cy << cl, q1 = min((cx + 1u) << cl, cw), r1 = min((cy + 1u) << cl, ch);
            float se, sx;
            vs_slab(ox, oy, ix, iz, (float)q0, (float)q1, (float)r0, (float)r1, se, sx);
            const uint32_t sid = vs_select_child(P, ox, oy, ix, iz, pl, ppx, ppy, fmaxf(se, 0.0f), vs_pack(cl, cx, cy));
            if (sid != kVsInvalid) { node = sid; advanced = true; break; }
            cl = pl; cx = ppx; cy = ppy;
        }
        if (!advanced) break;
    }
    return false;
}

__device__ __forceinline__ void vs_latlon_to_pixel(const ViewshedParams& P, float rcp_lon_step, float rcp_lat_step, float lat, float lon,
                                                   float& px, float& py) {   // :375-387; det_div(x, step) = x * det_rcp(step), reciprocals hoisted
    float lon_deg = vdegrees(lon);
    if (lon_deg < P.geodetic[2]) lon_deg += 360.0f;
    if (lon_deg > P.geodetic[2] + 180.0f) lon_deg -= 360.0f;
    px = (lon_deg - P.geodetic[2]) * rcp_lon_step - 0.5f;
    py = (P.geodetic[3] - vdegrees(lat)) * rcp_lat_step - 0.5f;
}
// geodesic_sample_pixel, :389-468: direct geodesic on the sphere (metric.w > 0) or Vincenty's direct formula on WGS84.
// The reference re-derives everything that depends on (lat0, azimuth) alone - reduced latitude, sigma1, alpha, u^2, A, B, C:
// ten polynomial trig evaluations - for EVERY half-cell segment of a chord; a chord has one start and one azimuth, so they
// are computed once per cell here (vs_geodesic_setup) with the same expressions in the same order.
struct VsGeodesic {
    float sin_lat0, cos_lat0, sin_az, cos_az;                      // sphere branch
    float flattening, semi_minor, sin_u1, cos_u1, sigma1, sin_alpha, cos_sq_alpha, ca, cb, cc;   // WGS84 branch
    // det_div(a, b) IS a * det_rcp(b) (determinism.wgsl:204): the reciprocals of per-chord constants are taken once
    float rcp_sigma_den;            // det_rcp(semi_minor * A)            (WGS84)  |  det_rcp(sphere radius)  (sphere)
    float rcp_lon_step, rcp_lat_step;   // det_rcp(metric.y), det_rcp(metric.z)   (latlon_to_pixel)
    float rcp_abs_sin_az, rcp_abs_cos_az;   // det_rcp(max(|sin az|, 1e-6)), det_rcp(max(|cos az|, 1e-6))   (shadow_step_m)
};
__device__ __forceinline__ VsGeodesic vs_geodesic_setup(const ViewshedParams& P, float lat0, float azimuth) {
    VsGeodesic G{};
    G.rcp_lon_step = vdet_rcp(P.metric[1]); G.rcp_lat_step = vdet_rcp(P.metric[2]);
    if (P.metric[3] > 0.0f) {
        G.sin_lat0 = vdet_sin(lat0); G.cos_lat0 = vdet_cos(lat0); G.sin_az = vdet_sin(azimuth); G.cos_az = vdet_cos(azimuth);
        G.rcp_sigma_den = vdet_rcp(P.metric[3]);
        G.rcp_abs_sin_az = vdet_rcp(fmaxf(fabsf(G.sin_az), 1e-6f)); G.rcp_abs_cos_az = vdet_rcp(fmaxf(fabsf(G.cos_az), 1e-6f));
        return G;
    }
    G.flattening = fdiv(1.0f, 298.257223563f);
    const float semi_major = 6378137.0f;
    G.semi_minor = semi_major * (1.0f - G.flattening);
    const float reduced = vdet_atan2((1.0f - G.flattening) * vdet_sin(lat0), vdet_cos(lat0));
    G.sin_u1 = vdet_sin(reduced); G.cos_u1 = vdet_cos(reduced);
    G.sin_az = vdet_sin(azimuth); G.cos_az = vdet_cos(azimuth);
    G.sigma1 = vdet_atan2(G.sin_u1, G.cos_u1 * G.cos_az);
    G.sin_alpha = G.cos_u1 * G.sin_az;
    G.cos_sq_alpha = 1.0f - G.sin_alpha * G.sin_alpha;
    const float u_sq = fdiv(G.cos_sq_alpha * (semi_major * semi_major - G.semi_minor * G.semi_minor), G.semi_minor * G.semi_minor);
    G.ca = 1.0f + fdiv(u_sq, 16384.0f) * (4096.0f + u_sq * (-768.0f + u_sq * (320.0f - 175.0f * u_sq)));
    G.cb = fdiv(u_sq, 1024.0f) * (256.0f + u_sq * (-128.0f + u_sq * (74.0f - 47.0f * u_sq)));
    G.cc = fdiv(G.flattening, 16.0f) * G.cos_sq_alpha * (4.0f + G.flattening * (4.0f - 3.0f * G.cos_sq_alpha));
    G.rcp_sigma_den = vdet_rcp(G.semi_minor * G.ca);
    G.rcp_abs_sin_az = vdet_rcp(fmaxf(fabsf(G.sin_az), 1e-6f)); G.rcp_abs_cos_az = vdet_rcp(fmaxf(fabsf(G.cos_az), 1e-6f));
    return G;
}
__device__ __forceinline__ void vs_geodesic_pixel(const ViewshedParams& P, const VsGeodesic& G, float lon0, float distance_m,
                                                  float& px, float& py) {
    if (P.metric[3] > 0.0f) {
        const float ad = distance_m * G.rcp_sigma_den;                      // det_div(distance, radius)
        const float sin_lat = vdet_fma(G.sin_lat0, vdet_cos(ad), vdet_sin(ad) * G.cos_lat0 * G.cos_az);
        const float lat = 1.5707963267948966f - vdet_acos(fminf(fmaxf(sin_lat, -1.0f), 1.0f));
        const float lon = lon0 + vdet_atan2(G.sin_az * vdet_sin(ad) * G.cos_lat0, vdet_cos(ad) - G.sin_lat0 * vdet_sin(lat));
        vs_latlon_to_pixel(P, G.rcp_lon_step, G.rcp_lat_step, lat, lon, px, py);
        return;
    }
    const float sigma0 = distance_m * G.rcp_sigma_den;                     // det_div(distance, semi_minor * A)
    float sigma = sigma0;
#pragma unroll 1
    for (int it = 0; it < 4; it++) {
        const float two_sigma_m = 2.0f * G.sigma1 + sigma;
        const float ss = vdet_sin(sigma), cs = vdet_cos(sigma), c2 = vdet_cos(two_sigma_m);
        const float delta = G.cb * ss * (c2 + fdiv(G.cb, 4.0f) * (cs * (-1.0f + 2.0f * c2 * c2) -
                                                                  fdiv(G.cb, 6.0f) * c2 * (-3.0f + 4.0f * ss * ss) * (-3.0f + 4.0f * c2 * c2)));
        sigma = sigma0 + delta;
    }
    const float ss = vdet_sin(sigma), cs = vdet_cos(sigma);
    const float two_sigma_m = 2.0f * G.sigma1 + sigma;
    const float tmp = G.sin_u1 * ss - G.cos_u1 * cs * G.cos_az;
    const float lat = vdet_atan2(G.sin_u1 * cs + G.cos_u1 * ss * G.cos_az, (1.0f - G.flattening) * vdet_sqrt(G.sin_alpha * G.sin_alpha + tmp * tmp));
    const float lambda = vdet_atan2(ss * G.sin_az, G.cos_u1 * cs - G.sin_u1 * ss * G.cos_az);
    const float c2 = vdet_cos(two_sigma_m);
    const float dlon = lambda - (1.0f - G.cc) * G.flattening * G.sin_alpha * (sigma + G.cc * ss * (c2 + G.cc * cs * (-1.0f + 2.0f * c2 * c2)));
    vs_latlon_to_pixel(P, G.rcp_lon_step, G.rcp_lat_step, lat, lon0 + dlon, px, py);
}
__device__ __forceinline__ float vs_shadow_step_m(const ViewshedParams& P, float lat, float rcp_abs_sin_az, float rcp_abs_cos_az) {   // :584-606 (sin / cos of the azimuth hoisted: VsGeodesic)
    const float sl = vdet_sin(lat);
    const float ft = 1.0f - 0.0066943799901413165f * sl * sl;
    const float root = vdet_sqrt(ft);
    const float meridional = vdet_div(6378137.0f * (1.0f - 0.0066943799901413165f), ft * root);
    const float prime_vertical = vdet_div(6378137.0f, root);
    const float hm = P.metric[3] > 0.0f ? P.metric[3] : meridional, hp = P.metric[3] > 0.0f ? P.metric[3] : prime_vertical;
    const float north_cell = hm * vradians(P.metric[2]);
    const float east_cell = hp * vdet_cos(lat) * vradians(P.metric[1]);
    const float east_cross = east_cell * rcp_abs_sin_az;         // det_div(east_cell, max(|sin az|, 1e-6))
    const float north_cross = north_cell * rcp_abs_cos_az;      // det_div(north_cell, max(|cos az|, 1e-6))
    return fmaxf(0.1f, 0.5f * fminf(north_cross, east_cross));
}

constexpr int kViewshedThreads = 64;   // 8 x 8 cells per CTA, the reference's workgroup shape

// ViewshedCell (:9-14) as one 16-byte record: visible (0 hidden, 1 visible, 2 geodesic left the DEM), drop, gain, horizon
__global__ void __launch_bounds__(kViewshedThreads) k_viewshed(const ViewshedParams P, const float2* __restrict__ positions_m,
                                                               float4* __restrict__ result) {
    const uint32_t x = blockIdx.x * 8u + (threadIdx.x & 7u), y = blockIdx.y * 8u + (threadIdx.x >> 3);
    if (x >= P.w || y >= P.h) return;
    const size_t index = (size_t)y * P.w + x;
    const float2 m = __ldg(positions_m + index);
    const float d2 = m.x * m.x + m.y * m.y;
    const float distance_m = fsqrt(d2);
    const float azimuth = atan2_pinned(m.x, m.y);
    float inv_radius = 0.0f;                                                                   // inverse_radius, :364-373
    if (d2 != 0.0f) inv_radius = fdiv(m.y * m.y, d2) * P.physics[0] + fdiv(m.x * m.x, d2) * P.physics[1];
    const float vacuum_drop = 0.5f * inv_radius * distance_m * distance_m;
    const float effective_drop = vacuum_drop * P.physics[2];
    const float refraction_gain = vacuum_drop - effective_drop;
    const float target_abs = __ldg(P.heights + index) + P.observer[3];
    float horizon = P.metric[0];
    if (inv_radius > 0.0f) {
        const float eff = inv_radius * P.physics[2];
        horizon = fsqrt(fdiv(2.0f * fmaxf(P.observer_elevation, 0.0f), eff)) + fsqrt(fdiv(2.0f * fmaxf(target_abs, 0.0f), eff));
    }
    if (distance_m == 0.0f) { result[index] = make_float4(__uint_as_float(1u), 0.0f, 0.0f, horizon); return; }
    uint32_t visible = 0u;
    if (!(distance_m > P.metric[0])) {
        const float target_elevation = target_abs - effective_drop;
        const float c[3] = {P.observer_elevation, vdet_div(target_elevation - P.observer_elevation, distance_m), 0.5f * inv_radius * P.physics[2]};
        visible = 1u;
        float start_d = 0.0f, spx = P.observer[0], spy = P.observer[1];
        const float maxx = (float)P.w - 0.5f, maxy = (float)P.h - 0.5f;
        const VsGeodesic G = vs_geodesic_setup(P, P.geodetic[0], azimuth);
        while (true) {
            const float seg_lat = vradians(vdet_fma(-(spy + 0.5f), P.metric[2], P.geodetic[3]));
            const float end_d = fminf(start_d + vs_shadow_step_m(P, seg_lat, G.rcp_abs_sin_az, G.rcp_abs_cos_az), distance_m);
            float epx, epy;
            vs_geodesic_pixel(P, G, P.geodetic[1], end_d, epx, epy);
            if (epx < -0.5f || epy < -0.5f || epx > maxx || epy > maxy) { visible = 2u; break; }
            if (vs_trace_segment(P, spx, spy, epx, epy, start_d, end_d, c, 0.001f)) { visible = 0u; break; }
            if (end_d >= distance_m) break;
            start_d = end_d; spx = epx; spy = epy;
        }
    }
    result[index] = make_float4(__uint_as_float(visible), vacuum_drop, refraction_gain, horizon);
}

// shadow_mask_main, :608-680.  lit[index] = 1 when no terrain inside the DEM blocks the sun (the reference packs bits
// with atomicOr and the host unpacks them again, viewshed.rs:556-566; bytes are written directly here).
__global__ void __launch_bounds__(kViewshedThreads) k_shadow_mask(const ViewshedParams P, const float4* __restrict__ inputs,
                                                                  uint8_t* __restrict__ lit) {
    const uint32_t x = blockIdx.x * 8u + (threadIdx.x & 7u), y = blockIdx.y * 8u + (threadIdx.x >> 3);
    if (x >= P.w || y >= P.h) return;
    const size_t index = (size_t)y * P.w + x;
    const float4 in = __ldg(inputs + index);
    uint8_t out = 0u;
    if (in.w > 0.0f) {
        const float lat0 = in.x, lon0 = in.y, azimuth = in.z;
        const float slope = vdet_div(vdet_sin(in.w), vdet_cos(in.w));
        float inv_radius;                                                                      // local_inverse_radius, :562-582
        if (P.physics[3] == 0.0f) inv_radius = 0.0f;
        else if (P.metric[3] > 0.0f) inv_radius = fdiv(1.0f, P.metric[3]);
        else {
            const float a = 6378137.0f, e2 = 0.0066943799901413165f;
            const float sl = vdet_sin(lat0);
            const float w = vdet_sqrt(1.0f - e2 * sl * sl);
            const float meridional = vdet_div(a * (1.0f - e2), w * w * w), prime_vertical = vdet_div(a, w);
            const float sa = vdet_sin(azimuth), ca = vdet_cos(azimuth);
            inv_radius = vdet_div(ca * ca, meridional) + vdet_div(sa * sa, prime_vertical);
        }
        const float c[3] = {__ldg(P.heights + index), slope, 0.5f * (inv_radius * P.physics[2])};
        float start_d = 0.0f, spx = (float)x, spy = (float)y, seg_lat = lat0;
        const float maxx = (float)P.w - 0.5f, maxy = (float)P.h - 0.5f;
        out = 1u;
        const VsGeodesic G = vs_geodesic_setup(P, lat0, azimuth);
        while (true) {
            const float end_d = fminf(start_d + vs_shadow_step_m(P, seg_lat, G.rcp_abs_sin_az, G.rcp_abs_cos_az), P.metric[0]);
            float epx, epy;
            vs_geodesic_pixel(P, G, lon0, end_d, epx, epy);
            if (vs_trace_segment(P, spx, spy, epx, epy, start_d, end_d, c, 0.01f)) { out = 0u; break; }
            if (epx < -0.5f || epy < -0.5f || epx > maxx || epy > maxy) break;
            if (end_d >= P.metric[0]) break;
            seg_lat = vradians(vdet_fma(-(epy + 0.5f), P.metric[2], P.geodetic[3]));
            start_d = end_d; spx = epx; spy = epy;
        }
    }
    lit[index] = out;
}

}  // namespace f3d
