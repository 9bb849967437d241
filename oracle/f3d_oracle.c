/*
 * oracle/f3d_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Plain-C, IEEE-754 binary32, round-to-nearest, NO FMA contraction
 * (compile with -ffp-contract=off, never -ffast-math) restatement of the
 * reference's path-traced DEM snapshot path.  See f3d_oracle.h for the pin
 * status and the rule about who may load this library.
 *
 * Reference sources followed (paths relative to /root/reference):
 *   src/shaders/hybrid_terrain_traversal.wgsl      (all)
 *   src/shaders/hybrid_traversal.wgsl:19-34,86-259
 *   src/shaders/hybrid_kernel.wgsl:8-38,78-85,109-112
 *   src/shaders/pt_restir_temporal.wgsl:54-109
 *   src/shaders/pt_restir_spatial.wgsl:45-117,141-224
 *   src/path_tracing/hybrid_compute/terrain_heightfield.rs:52-84,132-202,348-369
 *   src/path_tracing/hybrid_compute/render_terrain.rs:453-557,563-1434
 *   src/path_tracing/restir/types.rs:3-60, restir/buffers.rs:238-267
 *   src/geo/refraction.rs:1-145
 *
 * Numerics contract: where WGSL leaves precision to the driver
 * (normalize, mix, dot, sin, cos, atan2, acos) this file pins ONE definition:
 *   dot3(a,b)    = (a.x*b.x + a.y*b.y) + a.z*b.z
 *   normalize(v) = v * (1 / sqrt(dot3(v,v)))         (glam 0.24.2 Vec3::normalize does the same)
 *   mix(a,b,t)   = a*(1-t) + b*t                      (WGSL spec formula)
 *   sin/cos/atan2/acos = the Cephes single-precision kernels written out below.
 * The CUDA kernels implement the same definitions independently, so the two
 * agree bit for bit.
 */
#include "f3d_oracle.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* error plumbing                                                             */
/* ------------------------------------------------------------------------- */
static _Thread_local char g_err[512];
static int g_threads = 0;

static int fail(int cls, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return cls;
}
const char* f3do_last_error(void) { return g_err; }
void f3do_set_threads(int n) { g_threads = n; }
int f3do_get_threads(void) {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}

static double now_seconds(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------------- */
/* small vector helpers (operation order is part of the contract)            */
/* ------------------------------------------------------------------------- */
typedef struct { float x, y, z; } v3;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 vneg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float dot2(float ax, float ay, float bx, float by) { return ax * bx + ay * by; }
static inline v3 cross3(v3 a, v3 b) {
    /* glam 0.24.2 Vec3::cross (src/f32/vec3.rs) and WGSL cross() */
    return V3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline v3 normalize3(v3 a) {
    float inv = 1.0f / sqrtf(dot3(a, a));
    return vscale(a, inv);
}
static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline float luminance(v3 c) { return dot3(c, V3(0.2126f, 0.7152f, 0.0722f)); } /* hybrid_terrain_traversal.wgsl:416-418 */

/* ------------------------------------------------------------------------- */
/* pinned elementary functions (Cephes sinf/cosf/atanf/asinf kernels)         */
/* ------------------------------------------------------------------------- */
void f3do_sincos(float x, float* s_out, float* c_out) {
    /* valid for 0 <= x <= ~8 (phi = 2*pi*u, u in [0,1]) */
    int k = (int)(x * 0.636619772f + 0.5f);
    float fk = (float)k;
    float r = x - fk * 1.5703125f;
    r = r - fk * 4.837512969970703125e-4f;
    r = r - fk * 7.549789954891882e-8f;
    float z = r * r;
    float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z
               - 0.5f * z + 1.0f;
    float s, c;
    switch (k & 3) {
        case 0: s = sp; c = cp; break;
        case 1: s = cp; c = -sp; break;
        case 2: s = -sp; c = -cp; break;
        default: s = -cp; c = sp; break;
    }
    *s_out = s;
    *c_out = c;
}

static float atan_pos(float x) {
    /* x >= 0 */
    float y;
    if (x > 2.414213562373095f) { y = 1.5707963267948966f; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = 0.7853981633974483f; x = (x - 1.0f) / (x + 1.0f); }
    else { y = 0.0f; }
    float z = x * x;
    y = y + ((((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z
              - 3.33329491539e-1f) * z * x + x);
    return y;
}

float f3do_atan2(float y, float x) {
    const float PI_F = 3.14159265358979323846f;
    const float HALF_PI_F = 1.5707963267948966f;
    if (x == 0.0f) {
        if (y > 0.0f) return HALF_PI_F;
        if (y < 0.0f) return -HALF_PI_F;
        return 0.0f;
    }
    float q = y / x;
    float a = atan_pos(fabsf(q));
    if (q < 0.0f) a = -a;
    if (x < 0.0f) a = (y >= 0.0f) ? a + PI_F : a - PI_F;
    return a;
}

static float asin_core(float x) {
    /* |x| <= 1 */
    float a = fabsf(x);
    float z, w;
    int flag = 0;
    if (a > 0.5f) { z = 0.5f * (1.0f - a); w = sqrtf(z); flag = 1; }
    else { w = a; z = w * w; }
    float p = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z
                + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * w + w;
    if (flag) { p = p + p; p = 1.5707963267948966f - p; }
    return x < 0.0f ? -p : p;
}

float f3do_acos(float x) {
    if (x < -0.5f) return 3.14159265358979323846f - 2.0f * asin_core(sqrtf(0.5f * (1.0f + x)));
    if (x > 0.5f) return 2.0f * asin_core(sqrtf(0.5f * (1.0f - x)));
    return 1.5707963267948966f - asin_core(x);
}

/* IEEE binary16 round-to-nearest-even, what a RGBA16F textureStore performs. */
uint16_t f3do_f32_to_f16(float v) {
    uint32_t x;
    memcpy(&x, &v, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return (uint16_t)(sign | (ax > 0x7f800000u ? 0x7e00u : 0x7c00u));
    if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);            /* rounds to >= 65520 -> inf */
    if (ax < 0x33000001u) return (uint16_t)sign;                            /* < 2^-25 (ties to even -> 0) */
    int32_t e = (int32_t)(ax >> 23) - 127;
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;
    uint32_t shift;
    uint32_t base;
    if (e < -14) { shift = (uint32_t)(13 + (-14 - e)); base = 0; }
    else { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7fffffu; }
    uint32_t q = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    return (uint16_t)(sign | (base + q));
}

float f3do_f16_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) x = sign;
        else {
            int sh = 0;
            while (!(m & 0x400u)) { m <<= 1; sh++; }
            m &= 0x3ffu;
            x = sign | ((uint32_t)(127 - 15 - sh + 1) << 23) | (m << 13);
        }
    } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e + 112u) << 23) | (m << 13);
    float f;
    memcpy(&f, &x, 4);
    return f;
}

/* ------------------------------------------------------------------------- */
/* geo/refraction.rs                                                          */
/* ------------------------------------------------------------------------- */
#define WGS84_A_M 6378137.0
#define WGS84_E2 6.6943799901413165e-3

static double deg2rad(double d) { return d * (3.14159265358979323846 / 180.0); } /* f64::to_radians */

int f3do_earth_curvature(int earth_model, double lat_deg, double sphere_radius_m,
                         int refraction_model, double k_in, double pressure_mbar, double temperature_c,
                         double azimuth_deg, float* inv_two_r_prime, uint32_t* enabled) {
    /* effective_radius_m, src/geo/refraction.rs:121-130 */
    if (earth_model == 0 && refraction_model != 0)
        return fail(1, "flat earth only supports refraction_model='none'");
    if (earth_model < 0 || earth_model > 2) return fail(1, "unsupported earth_model %d", earth_model);
    if (refraction_model < 0 || refraction_model > 3)
        return fail(1, "unsupported refraction_model %d", refraction_model);
    /* directional_radius_m, :57-76 */
    if (!isfinite(azimuth_deg)) return fail(1, "azimuth must be finite");
    double radius;
    if (earth_model == 0) radius = INFINITY;
    else if (earth_model == 1) {
        if (!(isfinite(sphere_radius_m) && sphere_radius_m > 0.0))
            return fail(1, "sphere radius must be finite and positive");
        radius = sphere_radius_m;
    } else {
        if (!(isfinite(lat_deg) && lat_deg >= -90.0 && lat_deg <= 90.0))
            return fail(1, "latitude must be finite and in [-90, 90]");
        /* principal_radii_m, :6-13 */
        double phi = deg2rad(lat_deg);
        double sp = sin(phi);
        double w = sqrt(1.0 - WGS84_E2 * (sp * sp));
        double meridional = WGS84_A_M * (1.0 - WGS84_E2) / (w * w * w);
        double prime_vertical = WGS84_A_M / w;
        double az = deg2rad(azimuth_deg);
        double ca = cos(az), sa = sin(az);
        radius = 1.0 / ((ca * ca) / meridional + (sa * sa) / prime_vertical);
    }
    /* RefractionModel::k, :101-118; standard_k :139-144 */
    double k;
    if (refraction_model == 0) k = 0.0;
    else if (refraction_model == 3) k = k_in;
    else {
        if (!isfinite(pressure_mbar) || pressure_mbar <= 0.0 || temperature_c <= -273.15)
            return fail(1, "pressure must be positive and temperature above absolute zero");
        double base = refraction_model == 1 ? 0.13 : 1.0 / 7.0;
        k = base * (pressure_mbar / 1013.25) * (288.15 / (273.15 + temperature_c));
    }
    if (!(isfinite(k) && k < 1.0)) return fail(1, "refraction k must be finite and less than 1");
    double eff = radius / (1.0 - k);
    /* EarthCurvatureUniforms::new, terrain_heightfield.rs:68-82 */
    int en = isfinite(eff);
    *inv_two_r_prime = en ? (float)(0.5 / eff) : 0.0f;
    *enabled = (uint32_t)en;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* min-max pyramid: build_minmax_mips, terrain_heightfield.rs:132-202         */
/* ------------------------------------------------------------------------- */
#define MAX_LEVELS 32

typedef struct {
    int nlevels;
    uint32_t dims[MAX_LEVELS][2];
    float* levels[MAX_LEVELS]; /* [min,max] pairs */
    uint32_t cell_w, cell_h;
    uint64_t total_floats;
} pyramid;

static uint32_t next_pow2(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

static void pyramid_free(pyramid* p) {
    for (int i = 0; i < p->nlevels; i++) free(p->levels[i]);
    p->nlevels = 0;
}

static int pyramid_build(const float* heights, uint32_t w, uint32_t h, pyramid* out) {
    memset(out, 0, sizeof *out);
    if (w < 2 || h < 2)
        return fail(2, "terrain heightfield must be at least 2x2 texels, got %ux%u", w, h);
    size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++)
        if (!isfinite(heights[i])) return fail(2, "terrain heightfield contains non-finite samples");
    uint32_t cw = w - 1, ch = h - 1;
    uint32_t pw = next_pow2(cw), ph = next_pow2(ch);
    float* l0 = (float*)malloc((size_t)pw * ph * 2 * sizeof(float));
    if (!l0) return fail(1, "oracle: out of memory");
    for (size_t i = 0; i < (size_t)pw * ph; i++) { l0[2 * i] = INFINITY; l0[2 * i + 1] = -INFINITY; }
    for (uint32_t y = 0; y < ch; y++)
        for (uint32_t x = 0; x < cw; x++) {
            size_t i00 = (size_t)y * w + x, i10 = i00 + 1, i01 = i00 + w, i11 = i01 + 1;
            float a = heights[i00], b = heights[i10], c = heights[i01], d = heights[i11];
            size_t o = ((size_t)y * pw + x) * 2;
            l0[o] = fminf(fminf(fminf(a, b), c), d);
            l0[o + 1] = fmaxf(fmaxf(fmaxf(a, b), c), d);
        }
    out->levels[0] = l0;
    out->dims[0][0] = pw;
    out->dims[0][1] = ph;
    out->nlevels = 1;
    out->total_floats = (uint64_t)pw * ph * 2;
    while (out->dims[out->nlevels - 1][0] > 1 || out->dims[out->nlevels - 1][1] > 1) {
        uint32_t lw = out->dims[out->nlevels - 1][0], lh = out->dims[out->nlevels - 1][1];
        uint32_t nw = lw / 2 > 1 ? lw / 2 : 1, nh = lh / 2 > 1 ? lh / 2 : 1;
        const float* prev = out->levels[out->nlevels - 1];
        float* next = (float*)malloc((size_t)nw * nh * 2 * sizeof(float));
        if (!next) { pyramid_free(out); return fail(1, "oracle: out of memory"); }
        for (uint32_t y = 0; y < nh; y++)
            for (uint32_t x = 0; x < nw; x++) {
                float mn = INFINITY, mx = -INFINITY;
                for (uint32_t dy = 0; dy < 2; dy++)
                    for (uint32_t dx = 0; dx < 2; dx++) {
                        uint32_t sx = 2 * x + dx; if (sx > lw - 1) sx = lw - 1;
                        uint32_t sy = 2 * y + dy; if (sy > lh - 1) sy = lh - 1;
                        const float* v = prev + ((size_t)sy * lw + sx) * 2;
                        mn = fminf(mn, v[0]);
                        mx = fmaxf(mx, v[1]);
                    }
                next[((size_t)y * nw + x) * 2] = mn;
                next[((size_t)y * nw + x) * 2 + 1] = mx;
            }
        out->levels[out->nlevels] = next;
        out->dims[out->nlevels][0] = nw;
        out->dims[out->nlevels][1] = nh;
        out->total_floats += (uint64_t)nw * nh * 2;
        out->nlevels++;
    }
    out->cell_w = cw;
    out->cell_h = ch;
    return 0;
}

int f3do_build_minmax(const float* heights, uint32_t w, uint32_t h,
                      uint32_t* dims, float* levels_out, uint64_t cap) {
    pyramid p;
    int rc = pyramid_build(heights, w, h, &p);
    if (rc) return -rc;
    if (levels_out && cap < p.total_floats) { pyramid_free(&p); fail(1, "levels_out too small"); return -1; }
    uint64_t off = 0;
    for (int l = 0; l < p.nlevels; l++) {
        dims[2 * l] = p.dims[l][0];
        dims[2 * l + 1] = p.dims[l][1];
        uint64_t nfl = (uint64_t)p.dims[l][0] * p.dims[l][1] * 2;
        if (levels_out) memcpy(levels_out + off, p.levels[l], nfl * sizeof(float));
        off += nfl;
    }
    int n = p.nlevels;
    pyramid_free(&p);
    return n;
}

/* ------------------------------------------------------------------------- */
/* scene + traversal                                                          */
/* ------------------------------------------------------------------------- */
typedef struct { v3 origin; float tmin; v3 direction; float tmax; } ray_t;   /* hybrid_traversal.wgsl:29-34 */

typedef struct {  /* HybridHitResult, hybrid_traversal.wgsl:19-27 */
    float t; v3 point; v3 normal; uint32_t material_id, hit_type, hit;
} hit_t;

typedef struct {
    /* TerrainPtUniforms, terrain_heightfield.rs:31-38 / :348-369 */
    float ox, oz, sx, sz;
    float exaggeration, env_intensity;
    v3 albedo;
    uint32_t dem_w, dem_h, cell_w, cell_h;
    uint32_t mip_count, env_w, env_h;
    const float* heights;
    const float* env_rgb;
    pyramid pyr;
    /* EarthCurvatureUniforms */
    float inv_two_r_prime;
    uint32_t curvature_enabled;
    /* mesh (HybridUniforms + buffers) */
    const float* mesh_xyz; uint32_t mesh_nverts;
    const uint32_t* mesh_idx; uint32_t mesh_index_count;
    uint32_t traversal_mode; /* 0 hybrid, 3 terrain only */
} scene_t;

typedef struct { uint64_t nodes; } trace_stats;

static inline float safe_inv(float d) {           /* terrain_safe_inv :88-91 */
    float ad = fmaxf(fabsf(d), 1e-12f);
    return d < 0.0f ? -1.0f / ad : 1.0f / ad;
}

static inline float curved_height(const scene_t* S, const ray_t* r, float t, int apply_curv) { /* :95-103 */
    float hd2 = t * t * dot2(r->direction.x, r->direction.z, r->direction.x, r->direction.z);
    float corr = (apply_curv && S->curvature_enabled) ? hd2 * S->inv_two_r_prime : 0.0f;
    return r->origin.y + t * r->direction.y + corr;
}

static inline void curved_height_range(const scene_t* S, const ray_t* r, float t0, float t1,
                                       int apply_curv, float* lo, float* hi) { /* :108-127 */
    float y0 = curved_height(S, r, t0, apply_curv);
    float y1 = curved_height(S, r, t1, apply_curv);
    float minimum = fminf(y0, y1);
    if (apply_curv && S->curvature_enabled) {
        float a = dot2(r->direction.x, r->direction.z, r->direction.x, r->direction.z) * S->inv_two_r_prime;
        if (a > 0.0f) {
            float vertex = -r->direction.y / (2.0f * a);
            if (vertex >= t0 && vertex <= t1) minimum = fminf(minimum, curved_height(S, r, vertex, 1));
        }
    }
    *lo = minimum;
    *hi = fmaxf(y0, y1);
}

static inline void slab_xz(const ray_t* r, float x0, float x1, float z0, float z1, float* te, float* tx) { /* :131-141 */
    float inv_x = safe_inv(r->direction.x);
    float inv_z = safe_inv(r->direction.z);
    float tx0 = (x0 - r->origin.x) * inv_x, tx1 = (x1 - r->origin.x) * inv_x;
    if (tx0 > tx1) { float tmp = tx0; tx0 = tx1; tx1 = tmp; }
    float tz0 = (z0 - r->origin.z) * inv_z, tz1 = (z1 - r->origin.z) * inv_z;
    if (tz0 > tz1) { float tmp = tz0; tz0 = tz1; tz1 = tmp; }
    *te = fmaxf(tx0, tz0);
    *tx = fminf(tx1, tz1);
}

static inline uint32_t pack_node(uint32_t level, uint32_t x, uint32_t y) { return (level << 26) | (y << 13) | x; } /* :144-146 */

static inline void cell_heights(const scene_t* S, uint32_t cx, uint32_t cz, float h[4]) { /* :149-156 */
    float ex = S->exaggeration;
    const float* H = S->heights;
    size_t w = S->dem_w;
    h[0] = H[(size_t)cz * w + cx] * ex;
    h[1] = H[(size_t)cz * w + cx + 1] * ex;
    h[2] = H[(size_t)(cz + 1) * w + cx] * ex;
    h[3] = H[(size_t)(cz + 1) * w + cx + 1] * ex;
}

static int leaf_intersect(const scene_t* S, const ray_t* r, uint32_t cx, uint32_t cz, float t0, float t1,
                          int apply_curv, int any_hit, float* t_out) { /* :167-235 */
    float h[4];
    cell_heights(S, cx, cz, h);
    float tm = 0.5f * (t0 + t1);
    float d3[3];
    for (int i = 0; i < 3; i++) {
        float t = i == 0 ? t0 : (i == 1 ? tm : t1);
        float px = r->origin.x + t * r->direction.x;
        float pz = r->origin.z + t * r->direction.z;
        float u = clampf((px - S->ox) / S->sx - (float)cx, 0.0f, 1.0f);
        float v = clampf((pz - S->oz) / S->sz - (float)cz, 0.0f, 1.0f);
        float hh = mixf(mixf(h[0], h[1], u), mixf(h[2], h[3], u), v);
        d3[i] = curved_height(S, r, t, apply_curv) - hh;
    }
    float c = d3[0];
    float a = 2.0f * d3[2] + 2.0f * d3[0] - 4.0f * d3[1];
    float b = d3[2] - d3[0] - a;
    float s_hit = 1e30f;
    if (any_hit && c <= 0.0f) s_hit = 0.0f;
    else if (fabsf(a) < 1e-12f) {
        if (fabsf(b) > 1e-12f) {
            float s = -c / b;
            if (s >= 0.0f && s <= 1.0f) s_hit = s;
        }
    } else {
        float disc = b * b - 4.0f * a * c;
        if (disc >= 0.0f) {
            float sq = sqrtf(disc);
            float q = -0.5f * (b + (b >= 0.0f ? sq : -sq));
            float r0 = q / a;
            float r1 = fabsf(q) < 1e-30f ? 1e30f : c / q;
            if (r0 > r1) { float tmp = r0; r0 = r1; r1 = tmp; }
            if (r0 >= 0.0f && r0 <= 1.0f) s_hit = r0;
            else if (r1 >= 0.0f && r1 <= 1.0f) s_hit = r1;
        }
    }
    if (s_hit <= 1.0f) {
        float t = t0 + s_hit * (t1 - t0);
        if (t > r->tmin && t < r->tmax) { *t_out = t; return 1; }
    }
    return 0;
}

static v3 normal_at(const scene_t* S, v3 p, uint32_t cx, uint32_t cz) { /* :239-248 */
    float h[4];
    cell_heights(S, cx, cz, h);
    float u = clampf((p.x - S->ox) / S->sx - (float)cx, 0.0f, 1.0f);
    float v = clampf((p.z - S->oz) / S->sz - (float)cz, 0.0f, 1.0f);
    float dh_du = mixf(h[1] - h[0], h[3] - h[2], v);
    float dh_dv = mixf(h[2] - h[0], h[3] - h[1], u);
    return normalize3(V3(-dh_du / S->sx, 1.0f, -dh_dv / S->sz));
}

#define STACK_SIZE 64u

static hit_t terrain_trace(const scene_t* S, const ray_t* r, int any_hit, int apply_curv, trace_stats* st) { /* :254-372 */
    hit_t res;
    memset(&res, 0, sizeof res);
    res.t = r->tmax;
    res.hit_type = 3u;
    const uint32_t cell_w = S->cell_w, cell_h = S->cell_h;
    const float ox = S->ox, oz = S->oz, sx = S->sx, sz = S->sz;
    uint32_t stack[STACK_SIZE];
    uint32_t sp = 0;
    stack[sp++] = pack_node(S->mip_count - 1u, 0u, 0u);
    while (sp != 0u) {
        sp--;
        if (st) st->nodes++;
        uint32_t node = stack[sp];
        uint32_t level = node >> 26, ny = (node >> 13) & 0x1FFFu, nx = node & 0x1FFFu;
        uint32_t cx0 = nx << level, cz0 = ny << level;
        if (cx0 >= cell_w || cz0 >= cell_h) continue;
        uint32_t cx1 = (nx + 1u) << level; if (cx1 > cell_w) cx1 = cell_w;
        uint32_t cz1 = (ny + 1u) << level; if (cz1 > cell_h) cz1 = cell_h;
        float s0, s1;
        slab_xz(r, ox + (float)cx0 * sx, ox + (float)cx1 * sx, oz + (float)cz0 * sz, oz + (float)cz1 * sz, &s0, &s1);
        float t_lo = fmaxf(s0, r->tmin);
        float t_hi = fminf(s1, fminf(r->tmax, res.t));
        if (t_lo > t_hi) continue;
        const float* mmp = S->pyr.levels[level] + ((size_t)ny * S->pyr.dims[level][0] + nx) * 2;
        float mm_lo = mmp[0] * S->exaggeration, mm_hi = mmp[1] * S->exaggeration;
        float ry_lo, ry_hi;
        curved_height_range(S, r, t_lo, t_hi, apply_curv, &ry_lo, &ry_hi);
        if (ry_lo > mm_hi || ry_hi < mm_lo) continue;
        if (level == 0u) {
            float lt;
            if (leaf_intersect(S, r, cx0, cz0, t_lo, t_hi, apply_curv, any_hit, &lt) && lt < res.t) {
                res.hit = 1u;
                res.t = lt;
                res.point = vadd(r->origin, vscale(r->direction, lt));
                res.normal = normal_at(S, res.point, cx0, cz0);
                res.material_id = 0u;
                res.hit_type = 3u;
                if (any_hit) return res;
            }
            continue;
        }
        uint32_t child_level = level - 1u;
        float child_t[4];
        uint32_t child_id[4];
        uint32_t child_count = 0;
        for (uint32_t cy = 0; cy < 2u; cy++)
            for (uint32_t cxi = 0; cxi < 2u; cxi++) {
                uint32_t ccx = nx * 2u + cxi, ccy = ny * 2u + cy;
                uint32_t gx0 = ccx << child_level, gz0 = ccy << child_level;
                if (gx0 >= cell_w || gz0 >= cell_h) continue;
                uint32_t gx1 = (ccx + 1u) << child_level; if (gx1 > cell_w) gx1 = cell_w;
                uint32_t gz1 = (ccy + 1u) << child_level; if (gz1 > cell_h) gz1 = cell_h;
                float c0, c1;
                slab_xz(r, ox + (float)gx0 * sx, ox + (float)gx1 * sx, oz + (float)gz0 * sz, oz + (float)gz1 * sz, &c0, &c1);
                float ct_lo = fmaxf(c0, t_lo), ct_hi = fminf(c1, t_hi);
                if (ct_lo > ct_hi) continue;
                child_t[child_count] = ct_lo;
                child_id[child_count] = pack_node(child_level, ccx, ccy);
                child_count++;
            }
        for (uint32_t i = 1; i < child_count; i++) {
            float kt = child_t[i];
            uint32_t kid = child_id[i];
            uint32_t j = i;
            while (!(j == 0u || child_t[j - 1u] >= kt)) {
                child_t[j] = child_t[j - 1u];
                child_id[j] = child_id[j - 1u];
                j--;
            }
            child_t[j] = kt;
            child_id[j] = kid;
        }
        for (uint32_t i = 0; i < child_count; i++)
            if (sp < STACK_SIZE) stack[sp++] = child_id[i];
    }
    return res;
}

/* ray_triangle_intersect, hybrid_traversal.wgsl:86-132 */
static hit_t ray_triangle(const ray_t* r, v3 v0, v3 v1, v3 v2) {
    hit_t res;
    memset(&res, 0, sizeof res);
    res.t = r->tmax;
    v3 e1 = vsub(v1, v0), e2 = vsub(v2, v0);
    v3 h = cross3(r->direction, e2);
    float a = dot3(e1, h);
    if (fabsf(a) < 1e-7f) return res;
    float f = 1.0f / a;
    v3 s = vsub(r->origin, v0);
    float u = f * dot3(s, h);
    if (u < 0.0f || u > 1.0f) return res;
    v3 q = cross3(s, e1);
    float v = f * dot3(r->direction, q);
    if (v < 0.0f || u + v > 1.0f) return res;
    float t = f * dot3(e2, q);
    if (t > r->tmin && t < r->tmax) {
        res.hit = 1u;
        res.t = t;
        res.point = vadd(r->origin, vscale(r->direction, t));
        res.normal = normalize3(cross3(e1, e2));
    }
    return res;
}

/* intersect_mesh (brute force), hybrid_traversal.wgsl:137-172 */
static hit_t intersect_mesh(const scene_t* S, const ray_t* r) {
    hit_t res;
    memset(&res, 0, sizeof res);
    res.t = r->tmax;
    uint32_t ic = S->mesh_index_count;
    if (ic < 3u) return res;
    for (uint32_t tri = 0; tri + 2u < ic; tri += 3u) {
        uint32_t i0 = S->mesh_idx[tri], i1 = S->mesh_idx[tri + 1], i2 = S->mesh_idx[tri + 2];
        if (i0 >= S->mesh_nverts || i1 >= S->mesh_nverts || i2 >= S->mesh_nverts) continue;
        const float* p0 = S->mesh_xyz + 3 * (size_t)i0;
        const float* p1 = S->mesh_xyz + 3 * (size_t)i1;
        const float* p2 = S->mesh_xyz + 3 * (size_t)i2;
        hit_t th = ray_triangle(r, V3(p0[0], p0[1], p0[2]), V3(p1[0], p1[1], p1[2]), V3(p2[0], p2[1], p2[2]));
        if (th.hit && th.t < res.t) res = th;
    }
    return res;
}

/* intersect_hybrid, hybrid_traversal.wgsl:175-201 */
static hit_t intersect_hybrid(const scene_t* S, const ray_t* r, trace_stats* st) {
    hit_t best;
    memset(&best, 0, sizeof best);
    best.t = r->tmax;
    if (S->traversal_mode == 0u || S->traversal_mode == 2u) {
        hit_t mh = intersect_mesh(S, r);
        if (mh.hit && mh.t < best.t) best = mh;
    }
    if (S->traversal_mode == 0u || S->traversal_mode == 3u) {
        ray_t tr = *r;
        tr.tmax = best.t;
        hit_t th = terrain_trace(S, &tr, 0, 0, st);
        if (th.hit && th.t < best.t) best = th;
    }
    return best;
}

/* intersect_hybrid_optimized + intersect_shadow_ray / intersect_ibl_occlusion_ray, :204-259 */
static int occluded(const scene_t* S, const ray_t* r, float early_exit, int apply_curv, float max_distance, trace_stats* st) {
    hit_t best;
    memset(&best, 0, sizeof best);
    best.t = r->tmax;
    if (S->traversal_mode == 0u || S->traversal_mode == 2u) {
        hit_t mh = intersect_mesh(S, r);
        if (mh.hit && mh.t < early_exit) return mh.t < max_distance;
        if (mh.hit && mh.t < best.t) best = mh;
    }
    if (S->traversal_mode == 0u || S->traversal_mode == 3u) {
        ray_t tr = *r;
        tr.tmax = best.t;
        hit_t th = terrain_trace(S, &tr, 1, apply_curv, st);
        if (th.hit && th.t < best.t) best = th;
    }
    return best.hit != 0u && best.t < max_distance;
}

static v3 surface_albedo(const scene_t* S, const hit_t* h) { /* get_surface_properties :238-245 */
    if (h->hit_type == 3u) return S->albedo;
    return V3(0.7f, 0.7f, 0.8f);
}

/* terrain_env_radiance, hybrid_terrain_traversal.wgsl:392-405 */
static v3 env_radiance(const scene_t* S, v3 dir) {
    float I = S->env_intensity;
    uint32_t ew = S->env_w, eh = S->env_h;
    if (ew == 0u || eh == 0u) return V3(I, I, I);
    const float PI_F = 3.14159265358979323846f;
    v3 d = normalize3(dir);
    float uu = (f3do_atan2(d.z, d.x) / (2.0f * PI_F)) + 0.5f;
    float vv = f3do_acos(clampf(d.y, -1.0f, 1.0f)) / PI_F;
    float fx = uu * (float)ew, fy = vv * (float)eh;
    /* WGSL u32(f32): truncate toward zero, saturating */
    uint32_t px = fx <= 0.0f ? 0u : (fx >= 4294967040.0f ? 0xFFFFFFFFu : (uint32_t)fx);
    uint32_t py = fy <= 0.0f ? 0u : (fy >= 4294967040.0f ? 0xFFFFFFFFu : (uint32_t)fy);
    if (px > ew - 1u) px = ew - 1u;
    if (py > eh - 1u) py = eh - 1u;
    const float* t = S->env_rgb + ((size_t)py * ew + px) * 3;
    return V3(t[0] * I, t[1] * I, t[2] * I);
}

static inline float xorshift32(uint32_t* st) { /* hybrid_kernel.wgsl:78-85 */
    uint32_t x = *st;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    *st = x;
    return (float)x / 4294967296.0f;
}

static inline float tent_offset(float u) { /* :409-414 */
    if (u < 0.5f) return sqrtf(2.0f * u) - 1.0f;
    return 1.0f - sqrtf(2.0f * (1.0f - u));
}

static v3 cosine_dir(v3 n, float u1, float u2) { /* :421-431 */
    const float PI_F = 3.14159265358979323846f;
    float sign = n.z < 0.0f ? -1.0f : 1.0f;
    float a = -1.0f / (sign + n.z);
    float b = n.x * n.y * a;
    v3 t = V3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    v3 bt = V3(b, sign + n.y * n.y * a, -n.y);
    float r = sqrtf(u1);
    float phi = 2.0f * PI_F * u2;
    float sphi, cphi;
    f3do_sincos(phi, &sphi, &cphi);
    v3 local = V3(r * cphi, r * sphi, sqrtf(fmaxf(0.0f, 1.0f - u1)));
    v3 d = vadd(vadd(vscale(t, local.x), vscale(bt, local.y)), vscale(n, local.z));
    return normalize3(d);
}

/* ------------------------------------------------------------------------- */
/* reservoirs: restir/types.rs:3-38, hybrid_terrain_traversal.wgsl:39-54       */
/* ------------------------------------------------------------------------- */
typedef struct {
    v3 position; uint32_t light_index; v3 direction; float intensity; uint32_t light_type;
    float w_sum; uint32_t m; float weight; float target_pdf;
} reservoir;

typedef struct {
    uint32_t width, height, frame_index, aov_flags;
    v3 cam_origin, cam_right, cam_up, cam_forward;
    float cam_exposure, half_w, half_h;
    uint32_t seed_hi, seed_lo;
    v3 light_dir, light_color;
    uint32_t spp, window;
} uniforms_t;

static ray_t camera_ray(const uniforms_t* U, uint32_t gx, uint32_t gy, float jx, float jy) { /* :480-485 */
    float ndc_x = (((float)gx + 0.5f + jx) / (float)U->width) * 2.0f - 1.0f;
    float ndc_y = (1.0f - ((float)gy + 0.5f + jy) / (float)U->height) * 2.0f - 1.0f;
    v3 rd = normalize3(V3(ndc_x * U->half_w, ndc_y * U->half_h, -1.0f));
    rd = normalize3(vadd(vadd(vscale(U->cam_right, rd.x), vscale(U->cam_up, rd.y)), vscale(vneg(U->cam_forward), rd.z)));
    ray_t r = {U->cam_origin, 1e-3f, rd, 1e30f};
    return r;
}

typedef struct {
    float* accum;       /* vec4 per pixel */
    float* welford;     /* vec2 per pixel */
    reservoir *curr, *out, *prev;
    float* gbuf_nr;     /* vec4 */
    float* gbuf_pos;    /* vec4 */
    uint16_t* aov_albedo;   /* rgba16f */
    uint16_t* aov_normal;   /* rgba16f */
    float* aov_depth;
    uint8_t* aov_vis;
} buffers_t;

typedef struct { uint64_t primary, shadow, ibl, nodes; } counters_t;

/* main_terrain, hybrid_terrain_traversal.wgsl:445-610 (one pixel) */
static void main_terrain_pixel(const scene_t* S, const uniforms_t* U, buffers_t* B, uint32_t gx, uint32_t gy, counters_t* C) {
    const uint32_t W = U->width;
    const uint32_t pix = gy * W + gx;
    trace_stats st = {0};

    reservoir prev_r = B->prev[pix];
    if (prev_r.m > 512u) {
        float scale = 512.0f / (float)prev_r.m;
        prev_r.w_sum = prev_r.w_sum * scale;
        prev_r.m = 512u;
        if (prev_r.target_pdf > 0.0f) prev_r.weight = prev_r.w_sum / ((float)prev_r.m * prev_r.target_pdf);
        B->prev[pix] = prev_r;
    }
    int prev_valid = U->frame_index > 0u && prev_r.m > 0u && prev_r.weight > 0.0f && prev_r.target_pdf > 0.0f
                     && prev_r.light_type == 1u;

    uint32_t rng = U->seed_hi ^ (gx * 1664525u) ^ (gy * 1013904223u) ^ (U->frame_index * 92837111u) ^ U->seed_lo;
    uint32_t spp = U->spp > 1u ? U->spp : 1u;

    v3 frame_radiance = V3(0, 0, 0);
    reservoir cand;
    memset(&cand, 0, sizeof cand);

    for (uint32_t s = 0; s < spp; s++) {
        float jx = tent_offset(xorshift32(&rng)) * 0.5f;
        float jy = tent_offset(xorshift32(&rng)) * 0.5f;
        ray_t ray = camera_ray(U, gx, gy, jx, jy);
        C->primary++;
        hit_t hit = intersect_hybrid(S, &ray, &st);
        if (hit.hit == 0u) {
            frame_radiance = vadd(frame_radiance, env_radiance(S, ray.direction));
            continue;
        }
        v3 n = hit.normal;
        v3 albedo = surface_albedo(S, &hit);

        v3 wi = normalize3(U->light_dir);
        float ndotl = fmaxf(dot3(n, wi), 0.0f);
        float target_pdf = luminance(vscale(vmul(albedo, U->light_color), ndotl));
        if (target_pdf > 0.0f) {
            cand.position = hit.point;
            cand.light_index = 0u;
            cand.direction = wi;
            cand.intensity = luminance(U->light_color);
            cand.light_type = 1u;
            cand.w_sum = cand.w_sum + target_pdf;
            cand.m = cand.m + 1u;
            cand.target_pdf = target_pdf;
        }

        v3 sun_dir = wi;
        float reuse_w = 1.0f;
        if (prev_valid) {
            sun_dir = normalize3(prev_r.direction);
            reuse_w = clampf(prev_r.weight, 0.0f, 4.0f);
        }
        v3 sun = V3(0, 0, 0);
        float nd = fmaxf(dot3(n, sun_dir), 0.0f);
        if (nd > 0.0f) {
            ray_t sray = {vadd(hit.point, vscale(n, 1e-3f)), 1e-3f, sun_dir, 1e30f};
            float vis = 1.0f;
            C->shadow++;
            if (occluded(S, &sray, 0.01f, 1, 1e30f, &st)) vis = 0.0f;
            sun = vscale(vscale(vscale(vmul(albedo, U->light_color), nd), vis), reuse_w);
        }

        float u1 = xorshift32(&rng);
        float u2 = xorshift32(&rng);
        v3 ei = cosine_dir(n, u1, u2);
        ray_t eray = {vadd(hit.point, vscale(n, 1e-3f)), 1e-3f, ei, 1e30f};
        float env_vis = 1.0f;
        C->ibl++;
        if (occluded(S, &eray, 0.01f, 0, 1e30f, &st)) env_vis = 0.0f;
        v3 ibl = vscale(vmul(albedo, env_radiance(S, ei)), env_vis);

        frame_radiance = vadd(vadd(frame_radiance, sun), ibl);
    }
    float fspp = (float)spp;
    frame_radiance = V3(frame_radiance.x / fspp, frame_radiance.y / fspp, frame_radiance.z / fspp);

    if (cand.m > 0u && cand.w_sum > 0.0f && cand.target_pdf > 0.0f)
        cand.weight = cand.w_sum / ((float)cand.m * cand.target_pdf);
    B->curr[pix] = cand;

    float* acc = B->accum + 4 * (size_t)pix;
    acc[0] = acc[0] + frame_radiance.x;
    acc[1] = acc[1] + frame_radiance.y;
    acc[2] = acc[2] + frame_radiance.z;
    acc[3] = acc[3] + 1.0f;

    uint32_t window = U->window > 2u ? U->window : 2u;
    float* wf = B->welford + 2 * (size_t)pix;
    float wmean = wf[0], wm2 = wf[1];
    if (U->frame_index % window == 0u) { wmean = 0.0f; wm2 = 0.0f; }
    float mean_lum = luminance(V3(acc[0] / acc[3], acc[1] / acc[3], acc[2] / acc[3]));
    float k = (float)(U->frame_index % window) + 1.0f;
    float delta = mean_lum - wmean;
    float mean = wmean + delta / k;
    float m2 = wm2 + delta * (mean_lum - mean);
    wf[0] = mean;
    wf[1] = m2;

    if (U->aov_flags != 0u) { /* :583-609 */
        ray_t cray = camera_ray(U, gx, gy, 0.0f, 0.0f);
        hit_t chit = intersect_hybrid(S, &cray, NULL);
        int is_hit = chit.hit != 0u;
        v3 calbedo = surface_albedo(S, &chit);
        if (chit.hit_type == 3u) calbedo = S->albedo;
        uint16_t* a = B->aov_albedo + 4 * (size_t)pix;
        uint16_t* nn = B->aov_normal + 4 * (size_t)pix;
        a[0] = f3do_f32_to_f16(is_hit ? calbedo.x : 0.0f);
        a[1] = f3do_f32_to_f16(is_hit ? calbedo.y : 0.0f);
        a[2] = f3do_f32_to_f16(is_hit ? calbedo.z : 0.0f);
        a[3] = f3do_f32_to_f16(1.0f);
        nn[0] = f3do_f32_to_f16(is_hit ? chit.normal.x : 0.0f);
        nn[1] = f3do_f32_to_f16(is_hit ? chit.normal.y : 0.0f);
        nn[2] = f3do_f32_to_f16(is_hit ? chit.normal.z : 0.0f);
        nn[3] = f3do_f32_to_f16(1.0f);
        if (is_hit) B->aov_depth[pix] = chit.t;
        else { uint32_t qn = 0x7fc00000u; memcpy(&B->aov_depth[pix], &qn, 4); }
        B->aov_vis[pix] = is_hit ? 255 : 0;
    }
    C->nodes += st.nodes;
}

/* main_terrain_gbuffer, hybrid_terrain_traversal.wgsl:619-644 */
static void gbuffer_pixel(const scene_t* S, const uniforms_t* U, buffers_t* B, uint32_t gx, uint32_t gy) {
    uint32_t pix = gy * U->width + gx;
    ray_t ray = camera_ray(U, gx, gy, 0.0f, 0.0f);
    hit_t hit = intersect_hybrid(S, &ray, NULL);
    float* nr = B->gbuf_nr + 4 * (size_t)pix;
    float* ps = B->gbuf_pos + 4 * (size_t)pix;
    if (hit.hit != 0u) {
        nr[0] = hit.normal.x; nr[1] = hit.normal.y; nr[2] = hit.normal.z; nr[3] = 1.0f;
        ps[0] = hit.point.x; ps[1] = hit.point.y; ps[2] = hit.point.z; ps[3] = 1.0f;
    } else {
        nr[0] = 0.0f; nr[1] = 0.0f; nr[2] = 1.0f; nr[3] = 1.0f;
        ps[0] = ps[1] = ps[2] = ps[3] = 0.0f;
    }
}

/* pt_restir_temporal::main, pt_restir_temporal.wgsl:54-109 */
static void temporal_pixel(const buffers_t* B, uint32_t idx) {
    reservoir rp = B->prev[idx], rc = B->curr[idx], ro;
    int prev_valid = rp.m > 0u && rp.weight > 0.0f && rp.target_pdf > 0.0f;
    int curr_valid = rc.m > 0u && rc.weight > 0.0f && rc.target_pdf > 0.0f;
    if (!prev_valid && !curr_valid) { B->out[idx] = rc; return; }
    if (!prev_valid) { B->out[idx] = rc; return; }
    if (!curr_valid) { B->out[idx] = rp; return; }
    int choose_prev = rp.weight > rc.weight;
    ro = choose_prev ? rp : rc;   /* sample + target_pdf */
    ro.m = rp.m + rc.m;
    ro.w_sum = rp.w_sum + rc.w_sum;
    if (ro.w_sum > 0.0f && ro.target_pdf > 0.0f) ro.weight = ro.w_sum / ((float)ro.m * ro.target_pdf);
    else ro.weight = 0.0f;
    B->out[idx] = ro;
}

/* consider_candidate (directional branch; this path binds exactly one directional light with
 * importance 1 and one zeroed area light), pt_restir_spatial.wgsl:45-117; render_terrain.rs:756-781 */
typedef struct { float Wsum; reservoir chosen; float chosen_pdf; uint32_t seed; } spatial_state;

static void consider_candidate(const buffers_t* B, const reservoir* r, uint32_t pix_idx, spatial_state* ss) {
    if (r->m == 0u) return;
    const float* nr = B->gbuf_nr + 4 * (size_t)pix_idx;
    v3 N = normalize3(V3(nr[0], nr[1], nr[2]));
    float p_curr = 0.0f;
    if (r->light_type == 1u) {
        const uint32_t dir_count = 1u;
        const float sum_imp_dir = 1.0f;
        float imp = fmaxf(1.0f, 0.0f);
        float p_sel = sum_imp_dir > 0.0f ? imp / fmaxf(sum_imp_dir, 1e-8f) : 1.0f / (float)dir_count;
        v3 wi = normalize3(r->direction);
        float cosTheta = fmaxf(dot3(N, wi), 0.0f);
        if (cosTheta <= 0.0f) return;
        p_curr = p_sel;
    } else if (r->light_type == 2u) {
        /* area lights never occur on this path (candidates are only ever light_type 1);
         * the zeroed area-light table would yield p_area = 1/(pi*1e-12): unreachable. */
        return;
    } else {
        return;
    }
    if (p_curr <= 0.0f || r->target_pdf <= 0.0f) return;
    float w = r->w_sum * (p_curr / fmaxf(r->target_pdf, 1e-6f));
    if (w <= 0.0f) return;
    ss->Wsum = ss->Wsum + w;
    float u = xorshift32(&ss->seed);
    if (u < w / ss->Wsum) {
        ss->chosen = *r;      /* only the LightSample part is consumed */
        ss->chosen_pdf = p_curr;
    }
}

/* pt_restir_spatial::main, pt_restir_spatial.wgsl:158-224 */
static void spatial_pixel(const uniforms_t* U, const buffers_t* B, uint32_t idx) {
    const uint32_t W = U->width, H = U->height;
    uint32_t x = idx % W, y = idx / W;
    const uint32_t K = 8u, R = 3u;
    spatial_state ss;
    ss.seed = (U->seed_hi ^ U->frame_index) + idx * 1664525u + 1013904223u;
    reservoir r_self = B->out[idx];
    ss.chosen = r_self;
    ss.chosen_pdf = r_self.target_pdf;
    ss.Wsum = 0.0f;
    uint32_t m_total = 0u;
    consider_candidate(B, &r_self, idx, &ss);
    m_total += r_self.m;
    for (uint32_t i = 0; i < K; i++) {
        int32_t rx = (int32_t)floorf(xorshift32(&ss.seed) * (float)(2u * R + 1u)) - (int32_t)R;
        int32_t ry = (int32_t)floorf(xorshift32(&ss.seed) * (float)(2u * R + 1u)) - (int32_t)R;
        if (rx == 0 && ry == 0) continue;
        int32_t nxi = (int32_t)x + rx; if (nxi < 0) nxi = 0; if (nxi > (int32_t)W - 1) nxi = (int32_t)W - 1;
        int32_t nyi = (int32_t)y + ry; if (nyi < 0) nyi = 0; if (nyi > (int32_t)H - 1) nyi = (int32_t)H - 1;
        uint32_t ni = (uint32_t)nyi * W + (uint32_t)nxi;
        reservoir rn = B->out[ni];
        consider_candidate(B, &rn, idx, &ss);
        m_total += rn.m;
    }
    reservoir o = ss.chosen;   /* sample fields */
    o.target_pdf = ss.chosen_pdf;
    o.w_sum = ss.Wsum;
    o.m = m_total;
    if (o.w_sum > 0.0f && o.target_pdf > 0.0f) o.weight = o.w_sum / ((float)o.m * o.target_pdf);
    else o.weight = 0.0f;
    B->prev[idx] = o;
}

/* ------------------------------------------------------------------------- */
/* validate_desc, render_terrain.rs:474-557                                   */
/* ------------------------------------------------------------------------- */
static int finite3(const float* v) { return isfinite(v[0]) && isfinite(v[1]) && isfinite(v[2]); }
static float vlen(v3 a) { return sqrtf(dot3(a, a)); }

static int validate_desc(const f3do_desc* d) {
    if (d->width == 0 || d->height == 0 || d->max_frames == 0)
        return fail(1, "terrain reference requires non-zero width/height/max_frames");
    if (d->min_frames > d->max_frames)
        return fail(1, "min_frames (%u) must be <= max_frames (%u)", d->min_frames, d->max_frames);
    if (d->spp == 0 || d->spp > 64) return fail(1, "spp must be in 1..=64, got %u", d->spp);
    if (!(isfinite(d->exaggeration) && d->exaggeration > 0.0f))
        return fail(1, "terrain exaggeration must be finite and > 0");
    if (!(finite3(d->cam_origin) && finite3(d->cam_look_at) && finite3(d->cam_up)))
        return fail(1, "camera origin/look_at/up must be finite");
    v3 origin = V3(d->cam_origin[0], d->cam_origin[1], d->cam_origin[2]);
    v3 fwd = vsub(V3(d->cam_look_at[0], d->cam_look_at[1], d->cam_look_at[2]), origin);
    if (vlen(fwd) < 1e-6f) return fail(1, "camera look_at must differ from origin");
    if (vlen(cross3(normalize3(fwd), V3(d->cam_up[0], d->cam_up[1], d->cam_up[2]))) < 1e-6f)
        return fail(1, "camera up vector must not be parallel to the view direction");
    if (!(isfinite(d->fov_y_deg) && d->fov_y_deg > 0.0f && d->fov_y_deg < 180.0f))
        return fail(1, "fov_y must be finite and in (0, 180) degrees, got %g", d->fov_y_deg);
    if (!(isfinite(d->exposure) && d->exposure > 0.0f)) return fail(1, "exposure must be finite and > 0");
    if (!(isfinite(d->sun_az_deg) && isfinite(d->sun_el_deg)))
        return fail(1, "sun azimuth/elevation must be finite");
    if (!(isfinite(d->sun_intensity) && d->sun_intensity >= 0.0f))
        return fail(1, "sun intensity must be finite and >= 0");
    if (!finite3(d->sun_color) || d->sun_color[0] < 0 || d->sun_color[1] < 0 || d->sun_color[2] < 0)
        return fail(1, "sun color must have three finite non-negative components");
    if (!(isfinite(d->env_intensity) && d->env_intensity >= 0.0f))
        return fail(1, "env intensity must be finite and >= 0");
    if (!(isfinite(d->variance_threshold) && d->variance_threshold > 0.0f))
        return fail(1, "variance threshold must be finite and > 0");
    if (!(isfinite(d->spacing[0]) && d->spacing[0] > 0.0f && isfinite(d->spacing[1]) && d->spacing[1] > 0.0f))
        return fail(1, "terrain spacing must be finite and > 0, got (%g, %g)", d->spacing[0], d->spacing[1]);
    if (d->mesh_xyz || d->mesh_idx) {
        if (!d->mesh_xyz || d->mesh_nverts == 0) return fail(1, "mesh vertices must be a non-empty flat [x,y,z] list");
        if (!d->mesh_idx || d->mesh_ntris == 0) return fail(1, "mesh indices must be a non-empty multiple of 3");
        for (size_t i = 0; i < (size_t)d->mesh_nverts * 3; i++)
            if (!isfinite(d->mesh_xyz[i])) return fail(1, "mesh vertices contain non-finite values");
        for (size_t i = 0; i < (size_t)d->mesh_ntris * 3; i++)
            if (d->mesh_idx[i] >= d->mesh_nverts) return fail(1, "mesh indices reference out-of-bounds vertices");
    }
    return 0;
}

static inline float clamp_radiometric(float v) { return clampf(v, 0.0f, 65504.0f); } /* render_terrain.rs:571-576 */
static inline float to_radians_f32(float d) { return d * (3.14159274101257324f / 180.0f); } /* Rust f32::to_radians */

/* ------------------------------------------------------------------------- */
/* the driver: render_terrain_reference, render_terrain.rs:563-1434           */
/* ------------------------------------------------------------------------- */
int f3do_render(const f3do_desc* d, f3do_out* out) {
    g_err[0] = 0;
    const double t_start = now_seconds();
    int rc = validate_desc(d);
    if (rc) return rc;
    const uint32_t W = d->width, H = d->height;
    const size_t npx = (size_t)W * H;
    float exposure = clamp_radiometric(d->exposure);
    float sun_intensity = clamp_radiometric(d->sun_intensity);
    float sun_color[3] = {clamp_radiometric(d->sun_color[0]), clamp_radiometric(d->sun_color[1]), clamp_radiometric(d->sun_color[2])};
    float env_intensity = clamp_radiometric(d->env_intensity);

    /* TerrainPtScene::new, terrain_heightfield.rs:390-494 */
    scene_t S;
    memset(&S, 0, sizeof S);
    if (!(isfinite(d->albedo[0]) && d->albedo[0] >= 0 && isfinite(d->albedo[1]) && d->albedo[1] >= 0 &&
          isfinite(d->albedo[2]) && d->albedo[2] >= 0))
        return fail(2, "terrain albedo must be finite and >= 0");
    rc = pyramid_build(d->heights, d->dem_w, d->dem_h, &S.pyr);
    if (rc) return rc;
    if (d->env_rgb) {
        if (d->env_w == 0 || d->env_h == 0) { pyramid_free(&S.pyr); return fail(2, "env map dims do not match data length"); }
        for (size_t i = 0; i < (size_t)d->env_w * d->env_h * 3; i++)
            if (!isfinite(d->env_rgb[i])) { pyramid_free(&S.pyr); return fail(2, "env map contains non-finite samples"); }
        S.env_rgb = d->env_rgb; S.env_w = d->env_w; S.env_h = d->env_h;
    }
    S.heights = d->heights;
    S.dem_w = d->dem_w; S.dem_h = d->dem_h;
    S.cell_w = S.pyr.cell_w; S.cell_h = S.pyr.cell_h;
    S.mip_count = (uint32_t)S.pyr.nlevels;
    S.sx = d->spacing[0]; S.sz = d->spacing[1];
    S.ox = -0.5f * ((float)d->dem_w - 1.0f) * S.sx;      /* terrain_heightfield.rs:359-360 */
    S.oz = -0.5f * ((float)d->dem_h - 1.0f) * S.sz;
    S.exaggeration = d->exaggeration;
    S.env_intensity = env_intensity;
    S.albedo = V3(d->albedo[0], d->albedo[1], d->albedo[2]);
    S.mesh_xyz = d->mesh_xyz; S.mesh_nverts = d->mesh_nverts;
    S.mesh_idx = d->mesh_idx; S.mesh_index_count = d->mesh_ntris * 3u;
    S.traversal_mode = d->mesh_xyz ? 0u : 3u;               /* render_terrain.rs:681-685 */
    out->minmax_pyramid_bytes = (uint64_t)d->dem_w * d->dem_h * 4 + S.pyr.total_floats * 4; /* :292-314 */

    rc = f3do_earth_curvature(d->earth_model, d->observer_lat_deg, d->sphere_radius_m, d->refraction_model,
                              d->refraction_k, d->pressure_mbar, d->temperature_c, (double)d->sun_az_deg,
                              &S.inv_two_r_prime, &S.curvature_enabled);
    if (rc == 0 && !(isfinite(d->observer_lat_deg) && d->observer_lat_deg >= -90.0 && d->observer_lat_deg <= 90.0 &&
                     isfinite(d->observer_lon_deg) && d->observer_lon_deg >= -180.0 && d->observer_lon_deg <= 180.0))
        rc = fail(1, "ray-origin latitude/longitude must be finite and in [-90,90]/[-180,180]");
    if (rc) { pyramid_free(&S.pyr); return rc; }

    /* camera + lighting uniforms, render_terrain.rs:635-661,697-708 */
    uniforms_t U;
    memset(&U, 0, sizeof U);
    v3 origin = V3(d->cam_origin[0], d->cam_origin[1], d->cam_origin[2]);
    v3 forward = normalize3(vsub(V3(d->cam_look_at[0], d->cam_look_at[1], d->cam_look_at[2]), origin));
    v3 right = normalize3(cross3(forward, V3(d->cam_up[0], d->cam_up[1], d->cam_up[2])));
    v3 up = normalize3(cross3(right, forward));
    float az = to_radians_f32(d->sun_az_deg), el = to_radians_f32(d->sun_el_deg);
    U.width = W; U.height = H;
    U.cam_origin = origin; U.cam_right = right; U.cam_up = up; U.cam_forward = forward;
    U.cam_exposure = exposure;
    float fov = to_radians_f32(d->fov_y_deg);
    float aspect = (float)W / (float)H;
    U.half_h = tanf(0.5f * fov);                              /* hybrid_terrain_traversal.wgsl:469-470 */
    U.half_w = aspect * U.half_h;
    U.seed_hi = d->seed;
    U.seed_lo = d->seed ^ 0x85EBCA6Bu;
    U.light_dir = V3(cosf(az) * cosf(el), sinf(el), sinf(az) * cosf(el));
    U.light_color = V3(sun_intensity * sun_color[0], sun_intensity * sun_color[1], sun_intensity * sun_color[2]);
    U.spp = d->spp > 1u ? d->spp : 1u;
    U.window = 32u;                                             /* WELFORD_WINDOW, render_terrain.rs:236 */

    if (d->compat_512mib_gate) { /* render_terrain.rs:785-888 working-set ledger */
        uint64_t total = (uint64_t)npx * (16 + 8 + 80 * 3 + 16 * 2 + 8 + 48) + out->minmax_pyramid_bytes + 16 +
                         (uint64_t)(d->env_rgb ? (uint64_t)d->env_w * d->env_h * 16 : 16) +
                         (uint64_t)d->mesh_nverts * 16 + (uint64_t)d->mesh_ntris * 12 + 96 + 32 + 80 + 96 + 24 + 32 + 32 + 48;
        uint64_t limit = 512ull * 1024 * 1024;
        if (total > limit) {
            pyramid_free(&S.pyr);
            return fail(1, "terrain PT exceeds the memory budget before rendering: tracked total %llu (host-visible %llu) > limit %llu",
                        (unsigned long long)total, 0ull, (unsigned long long)limit);
        }
    }

    buffers_t B;
    memset(&B, 0, sizeof B);
    B.accum = (float*)calloc(npx * 4, sizeof(float));
    B.welford = (float*)calloc(npx * 2, sizeof(float));
    B.curr = (reservoir*)calloc(npx, sizeof(reservoir));
    B.out = (reservoir*)calloc(npx, sizeof(reservoir));
    B.prev = (reservoir*)calloc(npx, sizeof(reservoir));
    B.gbuf_nr = (float*)calloc(npx * 4, sizeof(float));
    B.gbuf_pos = (float*)calloc(npx * 4, sizeof(float));
    B.aov_albedo = (uint16_t*)calloc(npx * 4, sizeof(uint16_t));
    B.aov_normal = (uint16_t*)calloc(npx * 4, sizeof(uint16_t));
    B.aov_depth = (float*)calloc(npx, sizeof(float));
    B.aov_vis = (uint8_t*)calloc(npx, 1);
    int result = 0;
    if (!B.accum || !B.welford || !B.curr || !B.out || !B.prev || !B.gbuf_nr || !B.gbuf_pos || !B.aov_albedo ||
        !B.aov_normal || !B.aov_depth || !B.aov_vis) { result = fail(1, "oracle: out of memory"); goto done; }

    int nthreads = f3do_get_threads();
    (void)nthreads;

    /* one-shot G-buffer pass, render_terrain.rs:1091-1121 */
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t y = 0; y < (int64_t)H; y++)
        for (uint32_t x = 0; x < W; x++) gbuffer_pixel(&S, &U, &B, x, (uint32_t)y);

    const double t_loop = now_seconds();
    uint32_t frames = 0;
    float variance = INFINITY;
    int converged = 0;
    uint64_t c_primary = 0, c_shadow = 0, c_ibl = 0, c_nodes = 0;
    while (frames < d->max_frames) { /* render_terrain.rs:1127-1234 */
        U.frame_index = frames;
        U.aov_flags = frames == 0 ? 0xFFu : 0u;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : c_primary, c_shadow, c_ibl, c_nodes)
        for (int64_t y = 0; y < (int64_t)H; y++) {
            counters_t C = {0, 0, 0, 0};
            for (uint32_t x = 0; x < W; x++) main_terrain_pixel(&S, &U, &B, x, (uint32_t)y, &C);
            c_primary += C.primary; c_shadow += C.shadow; c_ibl += C.ibl; c_nodes += C.nodes;
        }
#pragma omp parallel for schedule(static) num_threads(nthreads)
        for (int64_t i = 0; i < (int64_t)npx; i++) temporal_pixel(&B, (uint32_t)i);
#pragma omp parallel for schedule(static) num_threads(nthreads)
        for (int64_t i = 0; i < (int64_t)npx; i++) spatial_pixel(&U, &B, (uint32_t)i);
        frames++;

        int window_full = frames % 32u == 0u;
        if (window_full || frames == d->max_frames) {
            uint32_t n_window = ((frames - 1u) % 32u) + 1u;
            if (n_window >= 2u) {
                float n = (float)n_window;
                float vmax = 0.0f;
                int bad = 0;
                for (size_t i = 0; i < npx; i++) {
                    float m2 = B.welford[2 * i + 1];
                    if (!isfinite(m2)) { bad = 1; break; }
                    float v = m2 / (n - 1.0f);
                    /* f32::max: NaN-ignoring; v is finite here */
                    if (v > vmax) vmax = v;
                }
                if (bad) { result = fail(1, "terrain PT produced non-finite variance (NaN in accumulation)"); goto done; }
                variance = vmax;
                if (frames >= d->min_frames && variance < d->variance_threshold) { converged = 1; break; }
            }
        }
    }
    out->setup_seconds = t_loop - t_start;
    out->frames_seconds = now_seconds() - t_loop;
    if (!converged) {
        result = fail(1, "terrain PT did not converge: per-pixel luminance variance %.3e over the last 32-frame window after %u frames (threshold %.1e); raise max_frames or simplify the scene — refusing to return a fake reference",
                      (double)variance, frames, (double)d->variance_threshold);
        goto done;
    }

    /* reservoir validity, render_terrain.rs:1313-1337 */
    {
        int any_valid = 0;
        for (size_t i = 0; i < npx; i++) {
            const reservoir* r = &B.prev[i];
            if (!(isfinite(r->w_sum) && isfinite(r->weight) && isfinite(r->target_pdf))) {
                result = fail(1, "terrain PT reservoir bookkeeping produced non-finite values");
                goto done;
            }
            if (r->m > 0 && r->weight > 0.0f && r->target_pdf > 0.0f) any_valid = 1;
        }
        out->prev_m_max = 0; out->prev_weight_max = 0.0f; out->prev_w_sum_max = 0.0f;
        for (int a = 0; a < 3; a++) { out->prev_dir_min[a] = INFINITY; out->prev_dir_max[a] = -INFINITY; }
        for (size_t i = 0; i < npx; i++) {
            const reservoir* r = &B.prev[i];
            if (r->m > out->prev_m_max) out->prev_m_max = r->m;
            if (r->weight > out->prev_weight_max) out->prev_weight_max = r->weight;
            if (r->w_sum > out->prev_w_sum_max) out->prev_w_sum_max = r->w_sum;
            if (r->m > 0 && r->weight > 0.0f && r->target_pdf > 0.0f) {
                const float dv[3] = {r->direction.x, r->direction.y, r->direction.z};
                for (int a = 0; a < 3; a++) {
                    if (dv[a] < out->prev_dir_min[a]) out->prev_dir_min[a] = dv[a];
                    if (dv[a] > out->prev_dir_max[a]) out->prev_dir_max[a] = dv[a];
                }
            }
        }
        int require = d->sun_el_deg > 0.0f && sun_intensity > 0.0f &&
                      (sun_color[0] > 0.0f || sun_color[1] > 0.0f || sun_color[2] > 0.0f);
        if (require && !any_valid) {
            result = fail(1, "terrain PT ReSTIR reuse chain produced no valid reservoirs for a sun-lit scene — temporal/spatial reuse is broken");
            goto done;
        }
    }

    /* AETHER post (render_terrain.rs:1246-1311): overwrites the RGBA16F output of the last frame with
     * L_surface*T + L_inscatter resolved by the same Reinhard operator; AOVs are untouched. */
    uint16_t* aether_out = NULL;
    if (d->atmosphere) {
        const char* bad = f3do_aether_validate(d->atmosphere);   /* AetherPostPass::new, aether_post.rs:58-65 */
        if (bad) { result = fail(1, "%s", bad); goto done; }
        f3do_aether_view V;
        memset(&V, 0, sizeof V);
        V.width = W; V.height = H;
        V.cam_origin[0] = origin.x; V.cam_origin[1] = origin.y; V.cam_origin[2] = origin.z;
        V.cam_right[0] = right.x; V.cam_right[1] = right.y; V.cam_right[2] = right.z;
        V.cam_up[0] = up.x; V.cam_up[1] = up.y; V.cam_up[2] = up.z;
        V.cam_forward[0] = forward.x; V.cam_forward[1] = forward.y; V.cam_forward[2] = forward.z;
        V.tan_half_fov = tanf(0.5f * fov);                     /* (0.5 * fov_y_radians).tan(), aether_post.rs:139 */
        V.aspect = (float)W / (float)H;
        V.exposure = exposure;
        V.light_dir[0] = U.light_dir.x; V.light_dir[1] = U.light_dir.y; V.light_dir[2] = U.light_dir.z;
        V.sun_intensity = sun_intensity;
        aether_out = (uint16_t*)malloc(npx * 4 * sizeof(uint16_t));
        if (!aether_out) { result = fail(1, "oracle: out of memory"); goto done; }
        if (f3do_aether_post(d->atmosphere, &V, B.accum, B.aov_depth, B.aov_vis, aether_out) != 0) {
            free(aether_out);
            result = fail(1, "AETHER post failed");
            goto done;
        }
    }

    /* resolve: out_tex = reinhard(mean*exposure) as RGBA16F (hybrid_terrain_traversal.wgsl:576-579,
     * hybrid_kernel.wgsl:109-112), then f16 -> u8 (render_terrain.rs:1358-1366) */
    for (size_t i = 0; i < npx; i++) {
        const float* acc = B.accum + 4 * i;
        for (int c = 0; c < 3; c++) {
            float v;
            if (aether_out) v = f3do_f16_to_f32(aether_out[4 * i + c]);
            else {
                float mean = acc[c] / acc[3];
                float exposed = mean * exposure;
                float ldr = exposed / (1.0f + exposed);
                v = f3do_f16_to_f32(f3do_f32_to_f16(ldr));
            }
            out->rgba[4 * i + c] = (uint8_t)(clampf(v, 0.0f, 1.0f) * 255.0f + 0.5f);
        }
        out->rgba[4 * i + 3] = 255;
        for (int c = 0; c < 3; c++) {
            out->albedo[3 * i + c] = f3do_f16_to_f32(B.aov_albedo[4 * i + c]);
            out->normal[3 * i + c] = f3do_f16_to_f32(B.aov_normal[4 * i + c]);
        }
        out->depth[i] = B.aov_depth[i];
    }
    free(aether_out);
    if (out->accum) memcpy(out->accum, B.accum, npx * 4 * sizeof(float));
    out->frames = frames;
    out->variance = variance;
    out->converged = converged;
    out->rays_primary = c_primary;
    out->rays_shadow = c_shadow;
    out->rays_ibl = c_ibl;
    out->nodes_popped = c_nodes;

done:
    free(B.accum); free(B.welford); free(B.curr); free(B.out); free(B.prev);
    free(B.gbuf_nr); free(B.gbuf_pos); free(B.aov_albedo); free(B.aov_normal); free(B.aov_depth); free(B.aov_vis);
    pyramid_free(&S.pyr);
    return result;
}

/* ------------------------------------------------------------------------- */
/* KAT interface                                                              */
/* ------------------------------------------------------------------------- */
int f3do_trace_rays(const float* heights, uint32_t w, uint32_t h,
                    const float spacing[2], const float origin_xz[2], float exaggeration,
                    float inv_two_r_prime, int curvature_enabled,
                    const float* rays, uint64_t n, int any_hit, int apply_curvature,
                    uint8_t* hit, float* t, float* normal) {
    scene_t S;
    memset(&S, 0, sizeof S);
    int rc = pyramid_build(heights, w, h, &S.pyr);
    if (rc) return rc;
    S.heights = heights; S.dem_w = w; S.dem_h = h;
    S.cell_w = S.pyr.cell_w; S.cell_h = S.pyr.cell_h;
    S.mip_count = (uint32_t)S.pyr.nlevels;
    S.sx = spacing[0]; S.sz = spacing[1];
    S.ox = origin_xz[0]; S.oz = origin_xz[1];
    S.exaggeration = exaggeration;
    S.inv_two_r_prime = inv_two_r_prime;
    S.curvature_enabled = (uint32_t)(curvature_enabled != 0);
    S.traversal_mode = 3u;
    int nthreads = f3do_get_threads();
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const float* p = rays + 8 * i;
        ray_t r = {V3(p[0], p[1], p[2]), p[3], V3(p[4], p[5], p[6]), p[7]};
        hit_t hr = terrain_trace(&S, &r, any_hit, apply_curvature, NULL);
        hit[i] = (uint8_t)hr.hit;
        t[i] = hr.t;
        if (normal) { normal[3 * i] = hr.normal.x; normal[3 * i + 1] = hr.normal.y; normal[3 * i + 2] = hr.normal.z; }
    }
    pyramid_free(&S.pyr);
    return 0;
}
