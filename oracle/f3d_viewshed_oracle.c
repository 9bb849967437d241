/*
 * oracle/f3d_viewshed_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Plain-C f32 restatement of HELIOS' viewshed and solar shadow mask: the same packed-node
 * min-max descent as the path tracer, any-hit only, walked along geodesic chords in DEM-pixel space.
 * Reference sources followed (paths relative to /root/reference):
 *   src/shaders/terrain_viewshed.wgsl:1-680                      (everything)
 *   src/shaders/includes/determinism.wgsl:120-330,359-420        (det_barrier/fma/mix/rcp/div/sqrt/inverse_sqrt,
 *                                                                 det_sin/cos/atan01/atan2/acos)
 *   src/terrain/analysis/viewshed.rs:54-159,161-339,396-570      (physics_terms, validation, uniforms, result decode)
 *   src/geo/refraction.rs:6-13,100-144                           (principal radii, refraction k)
 *   src/path_tracing/hybrid_compute/terrain_heightfield.rs:132-202 (min-max chain, via f3do_build_minmax)
 *
 * Numerics: the reference itself pins most of this shader in software (det_* helpers: bit-trick seeds + Newton steps,
 * barriered mul/add, polynomial trig), which this file restates literally - those parts are bit-identical to ANY
 * conforming run of the reference.  What WGSL still leaves to the driver here is pinned by this project as in
 * `/` and sqrt = IEEE, atan2 (azimuth of the target, main:480) = the Cephes kernel f3do_atan2,
 * length(v) = sqrt(x*x + y*y), dot2 = x*x' + y*y', degrees(x) = x * 57.295779513082323, radians(x) = x * 0.017453292519943295,
 * no FMA contraction.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "f3d_oracle.h"

/* ------------------------------------------------------------------ determinism.wgsl ---------------------------------- */
static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float det_fma(float a, float b, float c) { float p = a * b; return p + c; }              /* :133-136 */
static inline float det_mix(float a, float b, float t) { float d = b - a; float s = d * t; return a + s; }   /* :143-147 */
static float det_inverse_sqrt(float x) {                                                               /* :185-193 */
    float xc = fmaxf(x, 1.17549435e-38f);
    float y = bitsf(0x5f3759dfu - (fbits(xc) >> 1));
    float half_x = 0.5f * xc;
    for (int i = 0; i < 3; i++) { float yy = y * y; float h = half_x * yy; y = y * (1.5f - h); }
    return y;
}
static float det_rcp(float x) {                                                                        /* :195-202 */
    float ax = fabsf(x);
    float y = bitsf(0x7EF311C3u - fbits(ax));
    for (int i = 0; i < 3; i++) { float p = ax * y; y = y * (2.0f - p); }
    return x < 0.0f ? -y : y;
}
static inline float det_div(float a, float b) { return a * det_rcp(b); }                               /* :204-206 */
static inline float det_sqrt(float x) { float r = x * det_inverse_sqrt(x); return x <= 0.0f ? 0.0f : r; }   /* :208-211 */
static float det_sin(float x) {                                                                        /* :360-381 */
    float k = floorf(x * 0.6366197723675814f + 0.5f);
    float kp = k * 1.5707963267948966f;
    float r = x - kp;
    int32_t q = (int32_t)k & 3;
    float r2 = r * r;
    float ps = det_fma(r2, -0.00019840874f, 0.0083333310f);
    ps = det_fma(r2, ps, -0.16666667f);
    ps = det_fma(r2, ps, 1.0f);
    float s = r * ps;
    float pc = det_fma(r2, -0.0013888378f, 0.041666638f);
    pc = det_fma(r2, pc, -0.5f);
    pc = det_fma(r2, pc, 1.0f);
    float v = (q & 1) == 1 ? pc : s;
    return (q & 2) == 2 ? -v : v;
}
static inline float det_cos(float x) { return det_sin(x + 1.5707963267948966f); }                      /* :383-385 */
static float det_atan01(float a) {                                                                     /* :388-397 */
    float s = a * a;
    float p = det_fma(s, -0.0117212f, 0.05265332f);
    p = det_fma(s, p, -0.11643287f);
    p = det_fma(s, p, 0.19354346f);
    p = det_fma(s, p, -0.33262347f);
    p = det_fma(s, p, 0.99997726f);
    return a * p;
}
static float det_atan2(float y, float x) {                                                             /* :399-412 */
    float ax = fabsf(x), ay = fabsf(y);
    float hi = fmaxf(ax, ay);
    if (hi == 0.0f) return 0.0f;
    float lo = fminf(ax, ay);
    float p = det_atan01(lo / hi);
    p = ay > ax ? 1.5707963267948966f - p : p;
    p = x < 0.0f ? 3.141592653589793f - p : p;
    return y < 0.0f ? -p : p;
}
static float det_acos(float x) {                                                                       /* :416-421 */
    float xc = fminf(fmaxf(x, -1.0f), 1.0f);
    float x2 = xc * xc;
    float s = sqrtf(fmaxf(1.0f - x2, 0.0f));
    return det_atan2(s, xc);
}
static inline float degrees_f(float r) { return r * 57.295779513082323f; }
static inline float radians_f(float d) { return d * 0.017453292519943295f; }

/* ------------------------------------------------------------------ scene -------------------------------------------- */
typedef struct {
    uint32_t w, h;                  /* uniforms.dimensions.xy (texels) */
    float observer[4], metric[4], physics[4], geodetic[4];
    const float* heights;
    int nlevels;
    uint32_t dims[64];
    float* levels;                  /* concatenated [min,max] levels, finest first */
    size_t level_off[32];           /* in floats */
} vs_scene;

static inline float height_texel(const vs_scene* S, uint32_t x, uint32_t y) { return S->heights[(size_t)y * S->w + x]; }

static float height_at(const vs_scene* S, float px, float py) {                                        /* terrain_viewshed.wgsl:24-43 */
    float lx = (float)(S->w - 1u), ly = (float)(S->h - 1u);
    float x = fminf(fmaxf(px, 0.0f), lx), y = fminf(fmaxf(py, 0.0f), ly);
    uint32_t x0 = (uint32_t)floorf(x), y0 = (uint32_t)floorf(y);
    uint32_t x1 = x0 + 1u < S->w - 1u ? x0 + 1u : S->w - 1u, y1 = y0 + 1u < S->h - 1u ? y0 + 1u : S->h - 1u;
    float fx = x - (float)x0, fy = y - (float)y0;
    return det_mix(det_mix(height_texel(S, x0, y0), height_texel(S, x1, y0), fx),
                   det_mix(height_texel(S, x0, y1), height_texel(S, x1, y1), fx), fy);
}

static inline float safe_inv(float d) {                                                                /* :55-58 */
    float m = fmaxf(fabsf(d), 1e-12f);
    return d < 0.0f ? -1.0f / m : 1.0f / m;
}
static void slab_xz(float ox, float oy, float dx, float dy, float x0, float x1, float z0, float z1, float* te, float* tx) {   /* :60-77 */
    float ix = safe_inv(dx), iz = safe_inv(dy);
    float tx0 = (x0 - ox) * ix, tx1 = (x1 - ox) * ix;
    if (tx0 > tx1) { float t = tx0; tx0 = tx1; tx1 = t; }
    float tz0 = (z0 - oy) * iz, tz1 = (z1 - oy) * iz;
    if (tz0 > tz1) { float t = tz0; tz0 = tz1; tz1 = t; }
    *te = fmaxf(tx0, tz0);
    *tx = fminf(tx1, tz1);
}
static inline uint32_t pack_node(uint32_t level, uint32_t x, uint32_t y) { return (level << 26) | (y << 13) | x; }
static inline float height_limit(float d, const float c[3]) { return det_fma(c[2], d * d, det_fma(c[1], d, c[0])); }   /* :83-89 */
static float height_limit_min(float d0, float d1, const float c[3]) {                                  /* :93-110 (.x only is consumed) */
    float h0 = height_limit(d0, c), h1 = height_limit(d1, c);
    float minimum = fminf(h0, h1);
    if (c[2] > 0.0f) {
        float vertex = -c[1] / (2.0f * c[2]);
        float dmin = fminf(d0, d1), dmax = fmaxf(d0, d1);
        if (vertex >= dmin && vertex <= dmax) minimum = fminf(minimum, height_limit(vertex, c));
    }
    return minimum;
}
static float leaf_deviation(float ox, float oy, float dx, float dy, const float hts[4], uint32_t cx, uint32_t cy, float st, float d0,
                            float d1, const float c[3]) {                                              /* :120-141 */
    float px = ox + st * dx, py = oy + st * dy;
    float u = fminf(fmaxf(px - (float)cx, 0.0f), 1.0f), v = fminf(fmaxf(py - (float)cy, 0.0f), 1.0f);
    float th = det_mix(det_mix(hts[0], hts[1], u), det_mix(hts[2], hts[3], u), v);
    return th - height_limit(det_mix(d0, d1, st), c);
}
static int leaf_occluded(const vs_scene* S, float ox, float oy, float dx, float dy, uint32_t cx, uint32_t cy, float t0, float t1, float d0,
                         float d1, const float c[3], float tol) {                                      /* :143-181 */
    float hts[4] = {height_texel(S, cx, cy), height_texel(S, cx + 1u, cy), height_texel(S, cx, cy + 1u), height_texel(S, cx + 1u, cy + 1u)};
    float tm = 0.5f * (t0 + t1);
    float e0 = leaf_deviation(ox, oy, dx, dy, hts, cx, cy, t0, d0, d1, c), e1 = leaf_deviation(ox, oy, dx, dy, hts, cx, cy, tm, d0, d1, c),
          e2 = leaf_deviation(ox, oy, dx, dy, hts, cx, cy, t1, d0, d1, c);
    float quadratic = 2.0f * e2 + 2.0f * e0 - 4.0f * e1;
    float linear = e2 - e0 - quadratic;
    float maximum = fmaxf(e0, e2);
    if (fabsf(quadratic) > 1e-12f) {
        float vertex = -linear / (2.0f * quadratic);
        if (vertex > 0.0f && vertex < 1.0f) maximum = fmaxf(maximum, det_fma(quadratic, vertex * vertex, det_fma(linear, vertex, e0)));
    }
    return maximum > tol;
}
#define VS_INVALID 0xFFFFFFFFu
static void select_child(const vs_scene* S, float ox, float oy, float dx, float dy, uint32_t plevel, uint32_t px, uint32_t py,
                         float after_t, uint32_t after_id, uint32_t* out_id, float* out_t) {            /* :186-229 */
    uint32_t cw = S->w - 1u, ch = S->h - 1u, cl = plevel - 1u;
    uint32_t best_id = VS_INVALID;
    float best_t = 2.0f;
    for (uint32_t ci = 0; ci < 4u; ci++) {
        uint32_t nx = px * 2u + (ci & 1u), ny = py * 2u + (ci >> 1);
        uint32_t x0 = nx << cl, y0 = ny << cl;
        if (x0 >= cw || y0 >= ch) continue;
        uint32_t x1 = ((nx + 1u) << cl) < cw ? ((nx + 1u) << cl) : cw, y1 = ((ny + 1u) << cl) < ch ? ((ny + 1u) << cl) : ch;
        float se, sx;
        slab_xz(ox, oy, dx, dy, (float)x0, (float)x1, (float)y0, (float)y1, &se, &sx);
        float entry = fmaxf(se, 0.0f), exit_t = fminf(sx, 1.0f);
        if (entry > exit_t) continue;
        uint32_t id = pack_node(cl, nx, ny);
        int follows = after_id == VS_INVALID || entry > after_t || (entry == after_t && id > after_id);
        if (follows && (best_id == VS_INVALID || entry < best_t || (entry == best_t && id < best_id))) { best_id = id; best_t = entry; }
    }
    *out_id = best_id;
    *out_t = best_t;
}
static int trace_segment(const vs_scene* S, float ox, float oy, float ex, float ey, float d0, float d1, const float c[3], float tol) {   /* :235-362 */
    float dx = ex - ox, dy = ey - oy;
    uint32_t cw = S->w - 1u, ch = S->h - 1u;
    uint32_t root = (uint32_t)S->nlevels - 1u;
    uint32_t node = pack_node(root, 0u, 0u);
    for (;;) {
        uint32_t level = node >> 26, ny = (node >> 13) & 0x1FFFu, nx = node & 0x1FFFu;
        uint32_t x0 = nx << level, y0 = ny << level;
        int descend = 0;
        if (x0 < cw && y0 < ch) {
            uint32_t x1 = ((nx + 1u) << level) < cw ? ((nx + 1u) << level) : cw, y1 = ((ny + 1u) << level) < ch ? ((ny + 1u) << level) : ch;
            float se, sx;
            slab_xz(ox, oy, dx, dy, (float)x0, (float)x1, (float)y0, (float)y1, &se, &sx);
            float t0 = fmaxf(se, 0.0f), t1 = fminf(sx, 1.0f);
            if (t0 <= t1) {
                float nd0 = det_mix(d0, d1, t0), nd1 = det_mix(d0, d1, t1);
                float hmin = height_limit_min(nd0, nd1, c);
                float mmax = S->levels[S->level_off[level] + 2u * ((size_t)ny * S->dims[2 * level] + nx) + 1u];
                if (hmin + tol < mmax) {
                    if (level == 0u) {
                        if (leaf_occluded(S, ox, oy, dx, dy, x0, y0, t0, t1, d0, d1, c, tol)) return 1;
                    } else {
                        uint32_t cid; float ct;
                        select_child(S, ox, oy, dx, dy, level, nx, ny, 0.0f, VS_INVALID, &cid, &ct);
                        if (cid != VS_INVALID) { node = cid; descend = 1; }
                    }
                }
            }
        }
        if (descend) continue;
        uint32_t cl = level, cx = nx, cy = ny;
        int advanced = 0;
        for (;;) {
            if (cl >= root) break;
            uint32_t pl = cl + 1u, ppx = cx >> 1, ppy = cy >> 1;
            uint32_t q0 = cx << cl, r0 = cy << cl;
            uint32_t q1 = ((cx + 1u) << cl) < cw ? ((cx + 1u) << cl) : cw, r1 = ((cy + 1u) << cl) < ch ? ((cy + 1u) << cl) : ch;
            float se, sx;
            slab_xz(ox, oy, dx, dy, (float)q0, (float)q1, (float)r0, (float)r1, &se, &sx);
            uint32_t sid; float stt;
            select_child(S, ox, oy, dx, dy, pl, ppx, ppy, fmaxf(se, 0.0f), pack_node(cl, cx, cy), &sid, &stt);
            if (sid != VS_INVALID) { node = sid; advanced = 1; break; }
            cl = pl; cx = ppx; cy = ppy;
        }
        if (!advanced) break;
    }
    return 0;
}

static float inverse_radius(const vs_scene* S, float mx, float my) {                                   /* :364-373 */
    float d2 = mx * mx + my * my;
    if (d2 == 0.0f) return 0.0f;
    float e2 = mx * mx / d2, n2 = my * my / d2;
    return n2 * S->physics[0] + e2 * S->physics[1];
}
static void latlon_to_pixel(const vs_scene* S, float lat, float lon, float* px, float* py) {             /* :375-387 */
    float lon_deg = degrees_f(lon);
    if (lon_deg < S->geodetic[2]) lon_deg += 360.0f;
    if (lon_deg > S->geodetic[2] + 180.0f) lon_deg -= 360.0f;
    *px = det_div(lon_deg - S->geodetic[2], S->metric[1]) - 0.5f;
    *py = det_div(S->geodetic[3] - degrees_f(lat), S->metric[2]) - 0.5f;
}
static void geodesic_sample_pixel(const vs_scene* S, float lat0, float lon0, float azimuth, float distance_m, float* px, float* py) {   /* :389-468 */
    if (S->metric[3] > 0.0f) {
        float ad = det_div(distance_m, S->metric[3]);
        float sin_lat = det_fma(det_sin(lat0), det_cos(ad), det_sin(ad) * det_cos(lat0) * det_cos(azimuth));
        float lat = 1.5707963267948966f - det_acos(fminf(fmaxf(sin_lat, -1.0f), 1.0f));
        float lon = lon0 + det_atan2(det_sin(azimuth) * det_sin(ad) * det_cos(lat0), det_cos(ad) - det_sin(lat0) * det_sin(lat));
        latlon_to_pixel(S, lat, lon, px, py);
        return;
    }
    const float flattening = 1.0f / 298.257223563f;
    const float semi_major = 6378137.0f;
    const float semi_minor = semi_major * (1.0f - flattening);
    float reduced = det_atan2((1.0f - flattening) * det_sin(lat0), det_cos(lat0));
    float sin_u1 = det_sin(reduced), cos_u1 = det_cos(reduced);
    float sin_az = det_sin(azimuth), cos_az = det_cos(azimuth);
    float sigma1 = det_atan2(sin_u1, cos_u1 * cos_az);
    float sin_alpha = cos_u1 * sin_az;
    float cos_sq_alpha = 1.0f - sin_alpha * sin_alpha;
    float u_sq = cos_sq_alpha * (semi_major * semi_major - semi_minor * semi_minor) / (semi_minor * semi_minor);
    float ca = 1.0f + u_sq / 16384.0f * (4096.0f + u_sq * (-768.0f + u_sq * (320.0f - 175.0f * u_sq)));
    float cb = u_sq / 1024.0f * (256.0f + u_sq * (-128.0f + u_sq * (74.0f - 47.0f * u_sq)));
    float sigma = det_div(distance_m, semi_minor * ca);
    for (int it = 0; it < 4; it++) {
        float two_sigma_m = 2.0f * sigma1 + sigma;
        float ss = det_sin(sigma), cs = det_cos(sigma), c2 = det_cos(two_sigma_m);
        float delta = cb * ss * (c2 + cb / 4.0f * (cs * (-1.0f + 2.0f * c2 * c2) - cb / 6.0f * c2 * (-3.0f + 4.0f * ss * ss) * (-3.0f + 4.0f * c2 * c2)));
        sigma = det_div(distance_m, semi_minor * ca) + delta;
    }
    float ss = det_sin(sigma), cs = det_cos(sigma);
    float two_sigma_m = 2.0f * sigma1 + sigma;
    float tmp = sin_u1 * ss - cos_u1 * cs * cos_az;
    float lat = det_atan2(sin_u1 * cs + cos_u1 * ss * cos_az, (1.0f - flattening) * det_sqrt(sin_alpha * sin_alpha + tmp * tmp));
    float lambda = det_atan2(ss * sin_az, cos_u1 * cs - sin_u1 * ss * cos_az);
    float cc = flattening / 16.0f * cos_sq_alpha * (4.0f + flattening * (4.0f - 3.0f * cos_sq_alpha));
    float c2 = det_cos(two_sigma_m);
    float dlon = lambda - (1.0f - cc) * flattening * sin_alpha * (sigma + cc * ss * (c2 + cc * cs * (-1.0f + 2.0f * c2 * c2)));
    latlon_to_pixel(S, lat, lon0 + dlon, px, py);
}
static float local_inverse_radius(const vs_scene* S, float lat, float azimuth) {                        /* :562-582 */
    if (S->physics[3] == 0.0f) return 0.0f;
    if (S->metric[3] > 0.0f) return 1.0f / S->metric[3];
    const float a = 6378137.0f, e2 = 0.0066943799901413165f;
    float sl = det_sin(lat);
    float w = det_sqrt(1.0f - e2 * sl * sl);
    float meridional = det_div(a * (1.0f - e2), w * w * w);
    float prime_vertical = det_div(a, w);
    float sa = det_sin(azimuth), ca = det_cos(azimuth);
    return det_div(ca * ca, meridional) + det_div(sa * sa, prime_vertical);
}
static float shadow_step_m(const vs_scene* S, float lat, float azimuth) {                               /* :584-606 */
    float sl = det_sin(lat);
    float ft = 1.0f - 0.0066943799901413165f * sl * sl;
    float root = det_sqrt(ft);
    float meridional = det_div(6378137.0f * (1.0f - 0.0066943799901413165f), ft * root);
    float prime_vertical = det_div(6378137.0f, root);
    float hm = S->metric[3] > 0.0f ? S->metric[3] : meridional;
    float hp = S->metric[3] > 0.0f ? S->metric[3] : prime_vertical;
    float north_cell = hm * radians_f(S->metric[2]);
    float east_cell = hp * det_cos(lat) * radians_f(S->metric[1]);
    float east_cross = det_div(east_cell, fmaxf(fabsf(det_sin(azimuth)), 1e-6f));
    float north_cross = det_div(north_cell, fmaxf(fabsf(det_cos(azimuth)), 1e-6f));
    return fmaxf(0.1f, 0.5f * fminf(north_cross, east_cross));
}

static void viewshed_cell(const vs_scene* S, const float* positions, uint32_t x, uint32_t y, float observer_elevation, uint32_t* visible,
                          float* drop, float* gain, float* horizon) {                                  /* main, :470-560 */
    size_t index = (size_t)y * S->w + x;
    float mx = positions[2 * index], my = positions[2 * index + 1];
    float distance_m = sqrtf(mx * mx + my * my);
    float azimuth = f3do_atan2(mx, my);
    float inv_radius = inverse_radius(S, mx, my);
    float vacuum_drop = 0.5f * inv_radius * distance_m * distance_m;
    float effective_drop = vacuum_drop * S->physics[2];
    float refraction_gain = vacuum_drop - effective_drop;
    float target_abs = height_texel(S, x, y) + S->observer[3];
    float horizon_distance = S->metric[0];
    if (inv_radius > 0.0f) {
        float eff = inv_radius * S->physics[2];
        horizon_distance = sqrtf(2.0f * fmaxf(observer_elevation, 0.0f) / eff) + sqrtf(2.0f * fmaxf(target_abs, 0.0f) / eff);
    }
    if (distance_m == 0.0f) { *visible = 1u; *drop = 0.0f; *gain = 0.0f; *horizon = horizon_distance; return; }
    *drop = vacuum_drop; *gain = refraction_gain; *horizon = horizon_distance;
    if (distance_m > S->metric[0]) { *visible = 0u; return; }
    float target_elevation = target_abs - effective_drop;
    float c[3] = {observer_elevation, det_div(target_elevation - observer_elevation, distance_m), 0.5f * inv_radius * S->physics[2]};
    uint32_t vis = 1u;
    float start_d = 0.0f, spx = S->observer[0], spy = S->observer[1];
    for (;;) {
        float seg_lat = radians_f(det_fma(-(spy + 0.5f), S->metric[2], S->geodetic[3]));
        float seg_len = shadow_step_m(S, seg_lat, azimuth);
        float end_d = fminf(start_d + seg_len, distance_m);
        float epx, epy;
        geodesic_sample_pixel(S, S->geodetic[0], S->geodetic[1], azimuth, end_d, &epx, &epy);
        float maxx = (float)S->w - 0.5f, maxy = (float)S->h - 0.5f;
        if (epx < -0.5f || epy < -0.5f || epx > maxx || epy > maxy) { vis = 2u; break; }
        if (trace_segment(S, spx, spy, epx, epy, start_d, end_d, c, 0.001f)) { vis = 0u; break; }
        if (end_d >= distance_m) break;
        start_d = end_d; spx = epx; spy = epy;
    }
    *visible = vis;
}

static int shadow_cell(const vs_scene* S, const float* inputs, uint32_t x, uint32_t y) {                 /* shadow_mask_main, :608-680; 1 = lit */
    size_t index = (size_t)y * S->w + x;
    const float* in = inputs + 4 * index;
    if (in[3] <= 0.0f) return 0;
    float lat0 = in[0], lon0 = in[1], azimuth = in[2];
    float origin_height = height_texel(S, x, y);
    float slope = det_div(det_sin(in[3]), det_cos(in[3]));
    float eff = local_inverse_radius(S, lat0, azimuth) * S->physics[2];
    float c[3] = {origin_height, slope, 0.5f * eff};
    float start_d = 0.0f, spx = (float)x, spy = (float)y, seg_lat = lat0;
    for (;;) {
        float end_d = fminf(start_d + shadow_step_m(S, seg_lat, azimuth), S->metric[0]);
        float epx, epy;
        geodesic_sample_pixel(S, lat0, lon0, azimuth, end_d, &epx, &epy);
        if (trace_segment(S, spx, spy, epx, epy, start_d, end_d, c, 0.01f)) return 0;
        float maxx = (float)S->w - 0.5f, maxy = (float)S->h - 0.5f;
        if (epx < -0.5f || epy < -0.5f || epx > maxx || epy > maxy) break;
        if (end_d >= S->metric[0]) break;
        seg_lat = radians_f(det_fma(-(epy + 0.5f), S->metric[2], S->geodetic[3]));
        start_d = end_d; spx = epx; spy = epy;
    }
    return 1;
}

/* physics_terms, viewshed.rs:54-78 (f64 host math, cast to f32) */
int f3do_viewshed_physics(int earth_model, double latitude_deg, double sphere_radius_m, int refraction_model, double k_in,
                          double pressure_mbar, double temperature_c, float physics[4]) {
    if (earth_model == 0 && refraction_model != 0) return 1;          /* "flat earth only supports refraction_model='none'" */
    double k;
    if (refraction_model == 0) k = 0.0;
    else if (refraction_model == 3) k = k_in;
    else {
        if (!isfinite(pressure_mbar) || pressure_mbar <= 0.0 || temperature_c <= -273.15) return 2;
        k = (refraction_model == 1 ? 0.13 : 1.0 / 7.0) * (pressure_mbar / 1013.25) * (288.15 / (273.15 + temperature_c));
    }
    if (!(isfinite(k) && k < 1.0)) return 3;
    double inv_m = 0.0, inv_p = 0.0;
    if (earth_model == 1) {
        if (!(isfinite(sphere_radius_m) && sphere_radius_m > 0.0)) return 4;
        inv_m = inv_p = 1.0 / sphere_radius_m;
    } else if (earth_model == 2) {
        if (!isfinite(latitude_deg) || latitude_deg < -90.0 || latitude_deg > 90.0) return 5;
        const double a = 6378137.0, e2 = 6.6943799901413165e-3;
        double phi = latitude_deg * (3.14159265358979323846 / 180.0), sp = sin(phi);
        double w = sqrt(1.0 - e2 * (sp * sp));
        inv_m = 1.0 / (a * (1.0 - e2) / (w * w * w));
        inv_p = 1.0 / (a / w);
    }
    physics[0] = (float)inv_m; physics[1] = (float)inv_p; physics[2] = (float)(1.0 - k); physics[3] = earth_model == 0 ? 0.0f : 1.0f;
    return 0;
}

static int scene_init(vs_scene* S, const float* heights, const f3do_viewshed_options* o) {
    memset(S, 0, sizeof *S);
    S->w = o->width; S->h = o->height; S->heights = heights;
    S->observer[0] = o->observer_x; S->observer[1] = o->observer_y; S->observer[2] = o->observer_height_m; S->observer[3] = o->target_height_m;
    S->metric[0] = o->max_distance_m; S->metric[1] = o->longitude_step_deg; S->metric[2] = o->latitude_step_deg; S->metric[3] = o->geodesic_sphere_radius_m;
    memcpy(S->physics, o->physics, sizeof S->physics);
    S->geodetic[0] = o->observer_latitude_rad; S->geodetic[1] = o->observer_longitude_rad; S->geodetic[2] = o->left_unwrapped_deg; S->geodetic[3] = o->top_deg;
    S->nlevels = f3do_build_minmax(heights, o->width, o->height, S->dims, NULL, 0);
    if (S->nlevels <= 0) return 1;
    size_t total = 0;
    for (int l = 0; l < S->nlevels; l++) { S->level_off[l] = total; total += (size_t)S->dims[2 * l] * S->dims[2 * l + 1] * 2u; }
    S->levels = (float*)malloc(total * sizeof(float));
    if (!S->levels) return 1;
    return f3do_build_minmax(heights, o->width, o->height, S->dims, S->levels, total) <= 0;
}

int f3do_viewshed(const float* heights, const float* positions_m, const f3do_viewshed_options* o, uint8_t* visible, float* drop,
                  float* gain, float* horizon) {
    vs_scene S;
    if (scene_init(&S, heights, o)) return 1;
    float observer_elevation = height_at(&S, S.observer[0], S.observer[1]) + S.observer[2];
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t y = 0; y < (int64_t)o->height; y++)
        for (uint32_t x = 0; x < o->width; x++) {
            size_t i = (size_t)y * o->width + x;
            uint32_t v;
            viewshed_cell(&S, positions_m, x, (uint32_t)y, observer_elevation, &v, &drop[i], &gain[i], &horizon[i]);
            visible[i] = (uint8_t)v;
        }
    free(S.levels);
    return 0;
}

int f3do_shadow_mask(const float* heights, const float* inputs, const f3do_viewshed_options* o, uint8_t* lit) {
    vs_scene S;
    if (scene_init(&S, heights, o)) return 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t y = 0; y < (int64_t)o->height; y++)
        for (uint32_t x = 0; x < o->width; x++) lit[(size_t)y * o->width + x] = (uint8_t)shadow_cell(&S, inputs, x, (uint32_t)y);
    free(S.levels);
    return 0;
}
