/* oracle/f3d_wavefront_oracle.c
 *
 * TEST INFRASTRUCTURE -- CPU oracle for SURVEY section 8f row 2: the wavefront multi-bounce path tracer behind
 * forge3d.render_adjudication_pair's path-traced half.  Never linked into, imported by or called from the product; only
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use it.
 *
 * A restatement, per pixel, of what the reference's five wavefront stages do to one path:
 *   src/path_tracing/adjudication.rs:76-331        (render_pt_reference: scene wiring, splitmix32 frame seeds :226-236,
 *                                                   >= 2 iterations rule :259-265, mean over frames + alpha = 1)
 *   src/path_tracing/wavefront/render.rs:87-209    (render_frame_simple: raygen, then <= 16 x intersect/shade/shadow/scatter,
 *                                                   overflow error :127-137), mod.rs:32,34,77,104 (MAX_DEPTH 8, capacity 4*W*H,
 *                                                   settings {debug 0, qmc 1, adaptive bits(0.25)})
 *   src/shaders/pt_raygen.wgsl:74-224              (xorshift32, tent filter, Sobol/CP-rotated jitter, camera ray)
 *   src/shaders/pt_intersect.wgsl:84-216,374-558   (slab test, watertight triangle, spheres, instances, hit/miss records)
 *   src/shaders/pt_shade.wgsl:44-96,103-196,342-470,478-862 (BSDF eval, env mixture / disc / delta NEE, Lambert-GGX-dielectric
 *                                                   continuation, Russian roulette with the adaptive threshold)
 *   src/shaders/pt_shadow.wgsl:1-58,161-294        (any-hit spheres + Moeller-Trumbore mesh, visible => accumulate)
 *   src/shaders/pt_scatter.wgsl:77-132             (continuation re-queue, miss => throughput * miss gradient)
 *   src/core/tonemap.rs:11-32                      (Reinhard + sRGB resolve)
 *
 * With ReSTIR off (adjudication.rs:93-95) a pixel's path never reads another pixel's state, so running the stages pixel by pixel
 * gives exactly the arithmetic the wavefront performs; a frame's iteration k holds precisely the paths at depth k.
 *
 * WHERE THIS DEPARTS FROM THE SHADERS, all stated in DESIGN.md section 9e:
 *   - accumulation order.  The reference adds several shadow rays of one pixel from different threads of one dispatch with a
 *     non-atomic read-modify-write (pt_shadow.wgsl:288-291), so its own result is order- and timing-dependent.  Here (and in the
 *     CUDA build) a frame sums its contributions to a pixel from zero in push order (emissive, environment, directional, area, then
 *     the miss term) and the frame sums are added to the accumulator in frame order -- the same terms as the reference's in-place
 *     adds, associated per frame so that frames can be traced concurrently on the device.
 *   - the mesh is swept triangle by triangle (lowest index wins ties); the BVH the reference traverses only prunes that sweep.
 *   - not restated: hair segments, ReSTIR reservoirs, the fog medium, anisotropic GGX (ax != ay), debug AOV preview.
 *   - transcendental intrinsics are pinned as in the rest of the oracle: sin/cos = f3do_sincos, tan = sin/cos,
 *     pow(x, y) = f3do_exp2(y * f3do_log2(x)) with the integer powers 2, 5 and 16 written as products, mix(a,b,t) = a*(1-t)+b*t,
 *     dot/normalize as in f3d_oracle.c.  No FMA contraction.
 *
 * PARITY PIN: tests/test_wavefront.py renders the committed adjudication scene with this file and compares it with the reference's
 * own golden tests/golden/adjudication/pt_reference.png under the reference's own drift gate (tests/test_adjudication_gate.py:
 * SSIM >= 0.995, mean |diff| <= 2.0 at 512 x 512 x 4096 spp); the summary of that run is committed under tests/golden/.
 */
#include <math.h>
#include <omp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "f3d_oracle.h"

static _Thread_local char g_wf_err[320];
const char* f3do_wavefront_last_error(void) { return g_wf_err; }
static int wf_fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_wf_err, sizeof g_wf_err, fmt, ap);
    va_end(ap);
    return 1;
}

typedef struct { float x, y, z; } w3;
static inline w3 W3(float x, float y, float z) { w3 r = {x, y, z}; return r; }
static inline w3 P3(const float* p) { return W3(p[0], p[1], p[2]); }
static inline w3 wadd(w3 a, w3 b) { return W3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline w3 wsub(w3 a, w3 b) { return W3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline w3 wmul(w3 a, w3 b) { return W3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline w3 wscale(w3 a, float s) { return W3(a.x * s, a.y * s, a.z * s); }
static inline w3 wneg(w3 a) { return W3(-a.x, -a.y, -a.z); }
static inline float wdot(w3 a, w3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline w3 wcross(w3 a, w3 b) { return W3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
static inline w3 wnormalize(w3 a) { return wscale(a, 1.0f / sqrtf(wdot(a, a))); }
static inline float wlength(w3 a) { return sqrtf(wdot(a, a)); }
static inline float wmix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline w3 wmix3(w3 a, w3 b, float t) { return W3(wmix(a.x, b.x, t), wmix(a.y, b.y, t), wmix(a.z, b.z, t)); }
static inline float wsat(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
static inline float wget(w3 a, uint32_t i) { return i == 0u ? a.x : (i == 1u ? a.y : a.z); }

#define WF_PI 3.14159265358979323846f

/* Pinned log2 for normal positive x (Cephes log2f: frexp split at sqrt(1/2), degree-8 polynomial, two-part log2(e)).
 * x <= 0 or below the smallest normal -> -inf, +inf -> +inf, NaN -> NaN. */
float f3do_log2(float x) {
    if (x != x) return x;
    if (!(x >= 1.17549435e-38f)) return -INFINITY;
    if (x == INFINITY) return x;
    uint32_t b;
    memcpy(&b, &x, 4);
    int32_t e = (int32_t)(b >> 23) - 126;
    b = (b & 0x007FFFFFu) | 0x3F000000u; /* mantissa in [0.5, 1) */
    float m;
    memcpy(&m, &b, 4);
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.0f; } else { m = m - 1.0f; }
    float z = m * m;
    float p = 7.0376836292e-2f;
    p = p * m + -1.1514610310e-1f;
    p = p * m + 1.1676998740e-1f;
    p = p * m + -1.2420140846e-1f;
    p = p * m + 1.4249322787e-1f;
    p = p * m + -1.6668057665e-1f;
    p = p * m + 2.0000714765e-1f;
    p = p * m + -2.4999993993e-1f;
    p = p * m + 3.3333331174e-1f;
    float y = (m * z) * p;
    y = y - 0.5f * z;
    float r = y * 0.44269504088896340736f;
    r = r + m * 0.44269504088896340736f;
    r = r + y;
    r = r + m;
    r = r + (float)e;
    return r;
}
float f3do_pow(float x, float y) { return f3do_exp2(y * f3do_log2(x)); }
static inline float pow2f(float x) { return x * x; }
static inline float pow5f(float x) { float x2 = x * x; return (x2 * x2) * x; }
static inline float pow16f(float x) { float a = x * x; a = a * a; a = a * a; return a * a; }

/* xorshift32, pt_raygen.wgsl:74-81 / pt_shade.wgsl:342-349 */
static inline float wf_rand(uint32_t* st) {
    uint32_t x = *st;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    *st = x;
    return (float)x / 4294967296.0f;
}
static inline float tent_filter(float u) { /* pt_raygen.wgsl:87-92 */
    if (u < 0.5f) return sqrtf(2.0f * u) - 1.0f;
    return 1.0f - sqrtf(2.0f * (1.0f - u));
}
static inline float cp_rotate(float u, float r) { float x = u + r; return x - floorf(x); } /* :155-159 */

/* sobol2, pt_raygen.wgsl:122-153 (integer work; the final scale by 2^-32 is exact) */
void f3do_wavefront_sobol2(uint32_t i, float* ox, float* oy) {
    uint32_t xb = 0, yb = 0, idx = i;
    for (uint32_t j = 0; j < 32u; j++) {
        if (idx & 1u) {
            uint32_t base = 0x80000000u >> j;
            xb ^= base;
            yb ^= base ^ ((base >> 1) ^ (base >> 3));
        }
        idx >>= 1;
    }
    *ox = (float)xb * (1.0f / 4294967296.0f);
    *oy = (float)yb * (1.0f / 4294967296.0f);
}
uint32_t f3do_wavefront_splitmix32(uint32_t x) { /* adjudication.rs:226-232 */
    x += 0x9E3779B9u;
    uint32_t z = x;
    z = (z ^ (z >> 16)) * 0x21F0AAADu;
    z = (z ^ (z >> 15)) * 0x735A2D97u;
    return z ^ (z >> 15);
}

typedef struct { w3 t, b, n; } basis3; /* columns of make_tangent_basis, pt_shade.wgsl:352-360 */
static inline basis3 tangent_basis(w3 n) {
    float sign = n.z < 0.0f ? -1.0f : 1.0f;
    float a = -1.0f / (sign + n.z);
    float b = (n.x * n.y) * a;
    basis3 r;
    r.t = W3(1.0f + ((sign * n.x) * n.x) * a, sign * b, -sign * n.x);
    r.b = W3(b, sign + (n.y * n.y) * a, -n.y);
    r.n = n;
    return r;
}
static inline w3 to_world(const basis3* B, w3 v) { /* mat3x3 * v = (c0*v.x + c1*v.y) + c2*v.z */
    return wadd(wadd(wscale(B->t, v.x), wscale(B->b, v.y)), wscale(B->n, v.z));
}
static inline w3 cosine_hemisphere(float u1, float u2) { /* :363-370 */
    float r = sqrtf(u1), phi = (2.0f * WF_PI) * u2, s, c;
    f3do_sincos(phi, &s, &c);
    return W3(r * c, r * s, sqrtf(fmaxf(0.0f, 1.0f - u1)));
}
static inline w3 reflect3(w3 i, w3 n) { return wsub(i, wscale(n, 2.0f * wdot(n, i))); } /* WGSL reflect: e1 - 2 dot(e2,e1) e2 */
static inline w3 refract3(w3 i, w3 n, float eta) {                                       /* WGSL refract */
    float d = wdot(n, i);
    float k = 1.0f - (eta * eta) * (1.0f - d * d);
    if (k < 0.0f) return W3(0, 0, 0);
    return wsub(wscale(i, eta), wscale(n, eta * d + sqrtf(k)));
}

typedef struct {
    const float* sph; uint32_t nsph;
    const float* dirl; uint32_t ndir;
    const float* areal; uint32_t narea;
    const float* imp; uint32_t nimp;
    const float* env;
    const float* xyz; uint32_t nverts; const uint32_t* idx; uint32_t ntris;
    const float* inst; uint32_t ninst;
} wf_scene;

typedef struct { w3 o, d; float tmin, tmax; } wf_ray;

/* ray_sphere, pt_intersect.wgsl:374-386 */
static inline float sphere_t(w3 ro, w3 rd, w3 c, float r) {
    w3 oc = wsub(ro, c);
    float b = wdot(oc, rd);
    float cterm = wdot(oc, oc) - r * r;
    float disc = b * b - cterm;
    if (disc <= 0.0f) return 1e30f;
    float s = sqrtf(disc);
    float t0 = -b - s, t1 = -b + s;
    if (t0 > 1e-3f) return t0;
    if (t1 > 1e-3f) return t1;
    return 1e30f;
}
/* ray_sphere, pt_shadow.wgsl:161-174 */
static inline int sphere_any(w3 ro, w3 rd, w3 c, float r, float tmin, float tmax) {
    w3 oc = wsub(ro, c);
    float b = wdot(oc, rd);
    float cterm = wdot(oc, oc) - r * r;
    float disc = b * b - cterm;
    if (disc <= 0.0f) return 0;
    float s = sqrtf(disc);
    float t0 = -b - s, t1 = -b + s;
    return (t0 > tmin && t0 < tmax) || (t1 > tmin && t1 < tmax);
}

/* ray_triangle_intersect (watertight), pt_intersect.wgsl:101-166 */
static int tri_watertight(const wf_ray* r, w3 v0, w3 v1, w3 v2, float* t_out, w3* n_out) {
    w3 A = wsub(v0, r->o), B = wsub(v1, r->o), C = wsub(v2, r->o);
    float adx = fabsf(r->d.x), ady = fabsf(r->d.y), adz = fabsf(r->d.z);
    uint32_t kz = 2, kx = 0, ky = 1;
    if (adx > ady && adx > adz) { kz = 0; kx = 1; ky = 2; }
    else if (ady > adz) { kz = 1; kx = 2; ky = 0; }
    float Sz = 1.0f / wget(r->d, kz);
    float Sx = wget(r->d, kx) * Sz, Sy = wget(r->d, ky) * Sz;
    float ax = wget(A, kx) - Sx * wget(A, kz), ay = wget(A, ky) - Sy * wget(A, kz);
    float bx = wget(B, kx) - Sx * wget(B, kz), by = wget(B, ky) - Sy * wget(B, kz);
    float cx = wget(C, kx) - Sx * wget(C, kz), cy = wget(C, ky) - Sy * wget(C, kz);
    float az = wget(A, kz) * Sz, bz = wget(B, kz) * Sz, cz = wget(C, kz) * Sz;
    float U = (bx * cy) - (by * cx);
    float V = (cx * ay) - (cy * ax);
    float Wd = (ax * by) - (ay * bx);
    if ((U < 0.0f || V < 0.0f || Wd < 0.0f) && (U > 0.0f || V > 0.0f || Wd > 0.0f)) return 0;
    float det = (U + V) + Wd;
    if (det == 0.0f) return 0;
    float T = (U * az + V * bz) + Wd * cz;
    float t = T / det;
    if (t > r->tmin && t < r->tmax) {
        *t_out = t;
        *n_out = wnormalize(wcross(wsub(v1, v0), wsub(v2, v0)));
        return 1;
    }
    return 0;
}
/* bvh_intersect_mesh(_desc), pt_intersect.wgsl:168-216,301-371 -- the sweep the BVH prunes; ties go to the lowest triangle */
static int mesh_closest(const wf_scene* S, const wf_ray* r, float* t_out, w3* n_out) {
    int hit = 0;
    float best = r->tmax;
    for (uint32_t k = 0; k < S->ntris; k++) {
        uint32_t i0 = S->idx[3 * k], i1 = S->idx[3 * k + 1], i2 = S->idx[3 * k + 2];
        float t; w3 n;
        if (tri_watertight(r, P3(S->xyz + 3 * (size_t)i0), P3(S->xyz + 3 * (size_t)i1), P3(S->xyz + 3 * (size_t)i2), &t, &n) && t < best) {
            best = t; *n_out = n; hit = 1;
        }
    }
    *t_out = best;
    return hit;
}
/* mesh_any_hit(_desc), pt_shadow.wgsl:4-58,193-245 (Moeller-Trumbore) */
static int mesh_any(const wf_scene* S, w3 ro, w3 rd, float tmin, float tmax) {
    for (uint32_t k = 0; k < S->ntris; k++) {
        w3 v0 = P3(S->xyz + 3 * (size_t)S->idx[3 * k]), v1 = P3(S->xyz + 3 * (size_t)S->idx[3 * k + 1]), v2 = P3(S->xyz + 3 * (size_t)S->idx[3 * k + 2]);
        w3 e1 = wsub(v1, v0), e2 = wsub(v2, v0);
        w3 h = wcross(rd, e2);
        float a = wdot(e1, h);
        if (fabsf(a) < 1e-7f) continue;
        float f = 1.0f / a;
        w3 s = wsub(ro, v0);
        float u = f * wdot(s, h);
        if (u < 0.0f || u > 1.0f) continue;
        w3 q = wcross(s, e1);
        float v = f * wdot(rd, q);
        if (v < 0.0f || u + v > 1.0f) continue;
        float t = f * wdot(e2, q);
        if (t > tmin && t < tmax) return 1;
    }
    return 0;
}
/* mat4x4 (column-major) * vec4(p, w), pt_intersect.wgsl:273-287: ((c0*x + c1*y) + c2*z) + c3*w */
static inline w3 xform(const float* m, w3 p, float w) {
    return W3(((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12] * w, ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13] * w,
              ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14] * w);
}
/* transpose(world_to_object) * vec4(n, 0), pt_intersect.wgsl:289-294 */
static inline w3 xform_normal(const float* m, w3 n) {
    return W3((m[0] * n.x + m[1] * n.y) + m[2] * n.z, (m[4] * n.x + m[5] * n.y) + m[6] * n.z, (m[8] * n.x + m[9] * n.y) + m[10] * n.z);
}
static inline uint32_t inst_word(const float* inst, uint32_t ii, uint32_t w) {
    uint32_t v;
    memcpy(&v, inst + 36 * (size_t)ii + w, 4);
    return v;
}

static int shadow_occluded(const wf_scene* S, w3 ro, w3 rd, float tmin, float tmax) { /* pt_shadow.wgsl:248-293 */
    for (uint32_t i = 0; i < S->nsph; i++) {
        const float* s = S->sph + 20 * (size_t)i;
        if (sphere_any(ro, rd, P3(s), s[3], tmin, tmax)) return 1;
    }
    if (S->ninst == 0u) return S->ntris ? mesh_any(S, ro, rd, tmin, tmax) : 0;
    for (uint32_t ii = 0; ii < S->ninst; ii++) {
        const float* w2o = S->inst + 36 * (size_t)ii + 16;
        if (inst_word(S->inst, ii, 32) != 0u) continue; /* blas_index beyond the single BLAS: mesh_any_hit_desc returns false */
        w3 ro_o = xform(w2o, ro, 1.0f);
        w3 rd_o = wnormalize(xform(w2o, rd, 0.0f));
        if (mesh_any(S, ro_o, rd_o, tmin, tmax)) return 1;
    }
    return 0;
}

/* bsdf_eval_pdf, pt_shade.wgsl:44-96 (isotropic branch) */
static void bsdf_eval(w3 wo, w3 wi, w3 n, w3 albedo, float metallic, float roughness, w3* f, float* pdf) {
    float ndl = fmaxf(wdot(n, wi), 0.0f), ndv = fmaxf(wdot(n, wo), 0.0f);
    if (ndl <= 0.0f || ndv <= 0.0f) { *f = W3(0, 0, 0); *pdf = 0.0f; return; }
    float kd = wsat(1.0f - metallic);
    w3 fd = wscale(W3(albedo.x / WF_PI, albedo.y / WF_PI, albedo.z / WF_PI), kd);
    float pdf_d = ndl / WF_PI;
    float m = fmaxf(0.02f, roughness * roughness);
    w3 h = wnormalize(wadd(wi, wo));
    float ndh = fmaxf(wdot(n, h), 0.0f), vdh = fmaxf(wdot(wo, h), 0.0f);
    float a2 = m * m;
    float D = a2 / fmaxf(WF_PI * pow2f((ndh * ndh) * (a2 - 1.0f) + 1.0f), 1e-6f);                 /* ggx_D :385-390 */
    float k = pow2f(m + 1.0f) / 8.0f;                                                           /* smith_G1 :393-397 */
    float G = (ndl / (ndl * (1.0f - k) + k)) * (ndv / (ndv * (1.0f - k) + k));
    float sm = wsat(metallic);
    w3 F0 = W3(wmix(0.04f, albedo.x, sm), wmix(0.04f, albedo.y, sm), wmix(0.04f, albedo.z, sm));
    float fw = pow5f(1.0f - wsat(vdh));                                                         /* fresnel_schlick :380-382 */
    w3 F = W3(F0.x + (1.0f - F0.x) * fw, F0.y + (1.0f - F0.y) * fw, F0.z + (1.0f - F0.z) * fw);
    float spec = (D * G) / fmaxf((4.0f * ndl) * ndv, 1e-6f);
    w3 fs = wscale(F, spec);
    float pdf_s = (D * ndh) / fmaxf(4.0f * vdh, 1e-6f);
    float ks = 1.0f - kd;
    *f = wadd(fd, fs);
    *pdf = fmaxf(kd * pdf_d + ks * pdf_s, 1e-8f);
}
static inline w3 env_color(const wf_scene* S, w3 wi) { /* :101-104 */
    return wmix3(P3(S->env), P3(S->env + 4), 0.5f * (wi.y + 1.0f));
}
static inline float power_cosine_pdf_up(w3 w) { /* :157-161, m = 16 */
    float c = fmaxf(wdot(W3(0, 1, 0), wnormalize(w)), 0.0f);
    return (17.0f * pow16f(c)) / (2.0f * WF_PI);
}

typedef struct { float r, g, b; } acc3;
static inline void acc_add(float* px, w3 c) { px[0] += c.x; px[1] += c.y; px[2] += c.z; }

/* One frame of one pixel.  Returns the number of rays the path put on the ray queue (its depth count); iters[k] is bumped for
 * every iteration k the path was alive in. */
static uint32_t trace_pixel(const wf_scene* S, const f3do_wavefront_scene* D, uint32_t W, uint32_t H, uint32_t pix, uint32_t frame,
                            uint32_t seed_hi, uint32_t seed_lo, float u1, float u2, float half_h, float aspect, float* accum,
                            uint64_t* iters) {
    const uint32_t px = pix % W, py = pix / W;
    /* ---- pt_raygen.wgsl:161-224, one sample ---- */
    uint32_t rr = seed_lo ^ (px * 9781u) ^ (py * 6271u) ^ (seed_hi * 13007u);
    float r1 = wf_rand(&rr), r2 = wf_rand(&rr);
    float jx = tent_filter(cp_rotate(u1, r1)) * 0.5f, jy = tent_filter(cp_rotate(u2, r2)) * 0.5f;
    float ndc_x = ((((float)px + 0.5f) + jx) / (float)W) * 2.0f - 1.0f;
    float ndc_y = (1.0f - (((float)py + 0.5f) + jy) / (float)H) * 2.0f - 1.0f;
    float half_w = aspect * half_h;
    w3 rd = wnormalize(W3(ndc_x * half_w, ndc_y * half_h, -1.0f));
    rd = wnormalize(wadd(wadd(wscale(P3(D->cam_right), rd.x), wscale(P3(D->cam_up), rd.y)), wscale(wneg(P3(D->cam_forward)), rd.z)));
    wf_ray ray;
    ray.o = P3(D->cam_origin); ray.d = rd; ray.tmin = 1e-4f; ray.tmax = 1e30f;
    w3 thr = W3(1, 1, 1);
    uint32_t rng_hi = seed_hi ^ (pix * 9781u) ^ (frame * 6271u);
    float px_acc[3] = {0.0f, 0.0f, 0.0f}; /* this frame's radiance sum; added to the accumulator when the path ends */
    float* px_out = accum + 4 * (size_t)pix;
#define WF_RETURN(n) do { px_out[0] += px_acc[0]; px_out[1] += px_acc[1]; px_out[2] += px_acc[2]; return (n); } while (0)

    for (uint32_t depth = 0; depth < 16u; depth++) {
        iters[depth]++;
        /* ---- pt_intersect.wgsl:389-557 ---- */
        float t_best = 1e30f;
        w3 n_hit = W3(0, 1, 0);
        uint32_t mat = 0;
        for (uint32_t i = 0; i < S->nsph; i++) {
            const float* s = S->sph + 20 * (size_t)i;
            float t = sphere_t(ray.o, ray.d, P3(s), s[3]);
            if (t >= ray.tmin && t < fminf(t_best, ray.tmax)) {
                t_best = t;
                n_hit = wnormalize(wsub(wadd(ray.o, wscale(ray.d, t)), P3(s)));
                mat = i;
            }
        }
        if (S->ninst == 0u) {
            float t; w3 n;
            if (S->ntris && mesh_closest(S, &ray, &t, &n) && t < t_best) { t_best = t; n_hit = n; mat = 0; }
        } else {
            for (uint32_t ii = 0; ii < S->ninst; ii++) {
                const float* w2o = S->inst + 36 * (size_t)ii + 16;
                if (inst_word(S->inst, ii, 32) != 0u) continue;
                wf_ray ro = ray;
                ro.o = xform(w2o, ray.o, 1.0f);
                ro.d = wnormalize(xform(w2o, ray.d, 0.0f));
                float t; w3 n;
                if (mesh_closest(S, &ro, &t, &n) && t < t_best) {
                    t_best = t;
                    n_hit = wnormalize(xform_normal(w2o, n));
                    uint32_t mid = inst_word(S->inst, ii, 33);
                    mat = S->nsph ? (mid < S->nsph - 1u ? mid : S->nsph - 1u) : 0u;
                }
            }
        }
        if (!(t_best < 1e20f)) {
            /* ---- miss, pt_scatter.wgsl:108-131 ---- */
            w3 sky = wmix3(P3(S->env + 8), P3(S->env + 12), 0.5f * (ray.d.y + 1.0f));
            acc_add(px_acc, wmul(thr, sky));
            WF_RETURN(depth + 1u);
        }
        const w3 hp = wadd(ray.o, wscale(ray.d, t_best));
        const w3 hn = n_hit;
        const w3 wo_raw = wnormalize(wneg(ray.d));

        /* ---- pt_shade.wgsl:478-861 ---- */
        const uint32_t mi = mat < S->nsph ? mat : 0u;
        const float* M = S->sph + 20 * (size_t)mi;
        const w3 albedo = P3(M + 4);
        const float metallic = M[7], roughness = M[8], ior = M[9];
        const w3 emissive = P3(M + 12);
        if (emissive.x > 0.0f || emissive.y > 0.0f || emissive.z > 0.0f) acc_add(px_acc, wmul(thr, emissive));
        uint32_t rng = rng_hi ^ (pix * 26699u) ^ (frame * 30977u);
        const w3 n = wnormalize(hn), wo = wnormalize(wo_raw);
        const float ndv = fmaxf(wdot(n, wo), 0.0f);
        const basis3 basis = tangent_basis(n);
        const float a = fmaxf(0.02f, roughness * roughness);
        const float sm = wsat(metallic);
        const w3 F0 = W3(wmix(0.04f, albedo.x, sm), wmix(0.04f, albedo.y, sm), wmix(0.04f, albedo.z, sm));
        const float imp = mi < S->nimp ? S->imp[mi] : 1.0f;
        const w3 so = wadd(hp, wscale(n, 1e-3f));

        { /* environment NEE, :584-614, sample_env_mixture :166-183 */
            float e1 = wf_rand(&rng), e2 = wf_rand(&rng), e3 = wf_rand(&rng);
            w3 wi;
            if (e1 < 0.5f) { /* sample_power_cosine_about_up :146-155 */
                float phi = (2.0f * WF_PI) * e3, s, c;
                float ct = f3do_pow(1.0f - e2, 1.0f / (16.0f + 1.0f));
                float st = sqrtf(fmaxf(0.0f, 1.0f - ct * ct));
                f3do_sincos(phi, &s, &c);
                wi = W3(st * c, ct, st * s);
            } else {
                wi = to_world(&basis, cosine_hemisphere(e2, e3));
            }
            float pdf_up = power_cosine_pdf_up(wi);
            float pdf_cos = fmaxf(wdot(n, wi), 0.0f) / WF_PI;
            float pdf_light = 0.5f * pdf_up + (1.0f - 0.5f) * pdf_cos;
            float cos_surf = fmaxf(wdot(n, wi), 0.0f);
            if (cos_surf > 0.0f) {
                w3 f; float bpdf;
                bsdf_eval(wo, wi, n, albedo, metallic, roughness, &f, &bpdf);
                float w_mis = pdf_light / fmaxf(pdf_light + bpdf, 1e-8f);
                w3 c = wscale(wscale(wscale(wmul(wmul(thr, f), env_color(S, wi)), cos_surf / fmaxf(pdf_light, 1e-8f)), w_mis), imp);
                if (!shadow_occluded(S, so, wi, 1e-3f, 1e30f)) acc_add(px_acc, c);
            }
        }
        if (S->ndir) { /* delta lights, :617-657 */
            float sum_imp = 0.0f;
            for (uint32_t i = 0; i < S->ndir; i++) sum_imp = sum_imp + fmaxf(S->dirl[8 * (size_t)i + 7], 0.0f);
            uint32_t idx = 0;
            float u = wf_rand(&rng);
            if (sum_imp > 0.0f) {
                float rsel = u * sum_imp, acc = 0.0f;
                for (uint32_t i = 0; i < S->ndir; i++) { acc = acc + fmaxf(S->dirl[8 * (size_t)i + 7], 0.0f); if (rsel <= acc) { idx = i; break; } }
            } else {
                idx = (uint32_t)floorf(u * (float)S->ndir);
            }
            const float* L = S->dirl + 8 * (size_t)(idx < S->ndir - 1u ? idx : S->ndir - 1u);
            w3 wi = wnormalize(wneg(P3(L)));
            float cos_surf = fmaxf(wdot(n, wi), 0.0f);
            if (cos_surf > 0.0f) {
                w3 f; float bpdf;
                bsdf_eval(wo, wi, n, albedo, metallic, roughness, &f, &bpdf);
                w3 Li = wscale(P3(L + 4), L[3]);
                float p_sel = sum_imp > 0.0f ? fmaxf(L[7], 0.0f) / fmaxf(sum_imp, 1e-8f) : 1.0f / (float)S->ndir;
                w3 c = wscale(wscale(wmul(wmul(thr, f), Li), cos_surf / fmaxf(p_sel, 1e-8f)), imp);
                if (!shadow_occluded(S, so, wi, 1e-3f, 1e30f)) acc_add(px_acc, c);
            }
        }
        if (S->narea) { /* disc lights, :660-707, sample_area_light_disc :114-143 */
            float sum_imp = 0.0f;
            for (uint32_t i = 0; i < S->narea; i++) sum_imp = sum_imp + fmaxf(S->areal[12 * (size_t)i + 11], 0.0f);
            uint32_t idx = 0;
            float u = wf_rand(&rng);
            if (sum_imp > 0.0f) {
                float rsel = u * sum_imp, acc = 0.0f;
                for (uint32_t i = 0; i < S->narea; i++) { acc = acc + fmaxf(S->areal[12 * (size_t)i + 11], 0.0f); if (rsel <= acc) { idx = i; break; } }
            } else {
                idx = (uint32_t)floorf(u * (float)S->narea);
            }
            const float* L = S->areal + 12 * (size_t)(idx < S->narea - 1u ? idx : S->narea - 1u);
            float a1 = wf_rand(&rng), a2 = wf_rand(&rng);
            w3 nL = wnormalize(P3(L + 4));
            basis3 bl = tangent_basis(nL);
            float rad = fmaxf(L[3], 1e-6f);
            float r = sqrtf(a1) * rad, phi = (2.0f * WF_PI) * a2, s, c;
            f3do_sincos(phi, &s, &c);
            /* the shader reads basisL[0][0], basisL[1][0], basisL[2][0]: the first ROW of the (t, b, n) matrix, :118-119 */
            w3 tL = W3(bl.t.x, bl.b.x, bl.n.x), bL = W3(bl.t.y, bl.b.y, bl.n.y);
            w3 X = wadd(wadd(P3(L), wscale(tL, r * c)), wscale(bL, r * s));
            w3 dir = wsub(X, hp);
            float d = wlength(dir);
            if (d > 1e-6f) {
                w3 wi = W3(dir.x / d, dir.y / d, dir.z / d);
                float cos_s = fmaxf(wdot(n, wi), 0.0f), cos_l = fmaxf(wdot(nL, wneg(wi)), 0.0f);
                if (cos_s > 0.0f && cos_l > 0.0f) {
                    float area = (WF_PI * rad) * rad;
                    float pdf = ((1.0f / area) * (d * d)) / fmaxf(cos_l, 1e-6f);
                    if (pdf > 0.0f) {
                        w3 f; float bpdf;
                        bsdf_eval(wo, wi, n, albedo, metallic, roughness, &f, &bpdf);
                        float p_sel = sum_imp > 0.0f ? fmaxf(L[11], 0.0f) / fmaxf(sum_imp, 1e-8f) : 1.0f / (float)S->narea;
                        float pdf_light = p_sel * pdf;
                        float w_mis = pdf_light / fmaxf(pdf_light + bpdf, 1e-8f);
                        w3 Li = wscale(P3(L + 8), L[7]);
                        w3 cc = wscale(wscale(wscale(wmul(wmul(thr, f), Li), cos_s / fmaxf(pdf_light, 1e-8f)), w_mis), imp);
                        if (!shadow_occluded(S, so, wi, 1e-3f, d - 1e-3f)) acc_add(px_acc, cc);
                    }
                }
            }
        }

        /* continuation, :738-808 */
        w3 wi, nthr;
        if (metallic > 0.5f) {
            float m1 = wf_rand(&rng), m2 = wf_rand(&rng);
            float a2 = a * a; /* sample_ggx_isotropic :404-413 */
            float ch = sqrtf((1.0f - m1) / (1.0f + (a2 - 1.0f) * m1));
            float sh = sqrtf(fmaxf(0.0f, 1.0f - ch * ch));
            float phi = (2.0f * WF_PI) * m2, s, c;
            f3do_sincos(phi, &s, &c);
            w3 hw = wnormalize(to_world(&basis, W3(sh * c, sh * s, ch)));
            wi = wnormalize(reflect3(wneg(wo), hw));
            float ndl = fmaxf(wdot(n, wi), 0.0f), ndh = fmaxf(wdot(n, hw), 0.0f), vdh = fmaxf(wdot(wo, hw), 0.0f);
            if (!(ndl > 0.0f && ndv > 0.0f)) WF_RETURN(depth + 1u); /* invalid sample: the thread moves to its next hit, :771-774 */
            float D = a2 / fmaxf(WF_PI * pow2f((ndh * ndh) * (a2 - 1.0f) + 1.0f), 1e-6f);
            float k = pow2f(a + 1.0f) / 8.0f;
            float G = (ndl / (ndl * (1.0f - k) + k)) * (ndv / (ndv * (1.0f - k) + k));
            float fw = pow5f(1.0f - wsat(vdh));
            w3 F = W3(F0.x + (1.0f - F0.x) * fw, F0.y + (1.0f - F0.y) * fw, F0.z + (1.0f - F0.z) * fw);
            w3 spec = wscale(F, (D * G) / fmaxf((4.0f * ndl) * ndv, 1e-6f));
            float pdf = (D * ndh) / fmaxf(4.0f * vdh, 1e-6f);
            nthr = wscale(wmul(thr, spec), ndl / fmaxf(pdf, 1e-6f));
        } else if (ior > 1.01f) {
            float cosi = wsat(wdot(n, wo));
            float F0s = pow2f((ior - 1.0f) / (ior + 1.0f));
            float F = F0s + (1.0f - F0s) * pow5f(1.0f - cosi);
            float u = wf_rand(&rng);
            if (u < F) {
                wi = wnormalize(reflect3(wneg(wo), n));
            } else {
                int entering = wdot(n, wo) > 0.0f;
                float eta = entering ? 1.0f / ior : ior / 1.0f;
                w3 N = entering ? n : wneg(n);
                wi = wnormalize(refract3(wneg(wo), N, eta));
                if (!(wdot(wi, wi) >= 1e-12f)) wi = wnormalize(reflect3(wneg(wo), n)); /* also catches normalize(0) = NaN */
            }
            nthr = wmul(thr, W3(fmaxf(albedo.x, 0.0f), fmaxf(albedo.y, 0.0f), fmaxf(albedo.z, 0.0f)));
        } else {
            float l1 = wf_rand(&rng), l2 = wf_rand(&rng);
            wi = wnormalize(to_world(&basis, cosine_hemisphere(l1, l2)));
            float ct = fmaxf(0.0f, wdot(n, wi));
            float pdf = ct / WF_PI + 1e-8f;
            nthr = wscale(wmul(thr, W3(albedo.x / WF_PI, albedo.y / WF_PI, albedo.z / WF_PI)), ct / pdf);
        }
        /* Russian roulette, :811-829 (adaptive threshold 0.25, mod.rs:104) */
        float rr_scale = 1.0f;
        if (depth >= 4u) {
            float max_c = fmaxf(nthr.x, fmaxf(nthr.y, nthr.z));
            float q = fminf(fmaxf(1.0f - max_c, 0.0f), 0.95f);
            float q_extra = fminf(fmaxf(1.0f - max_c / 0.25f, 0.0f), 0.90f);
            q = fminf(fmaxf(q + q_extra, 0.0f), 0.95f);
            float u = wf_rand(&rng);
            if (u < q) WF_RETURN(depth + 1u);
            rr_scale = 1.0f / (1.0f - q);
        }
        if (!(depth + 1u < 16u)) WF_RETURN(depth + 1u);
        /* scatter, :831-848, pt_scatter.wgsl:77-106 */
        ray.o = wadd(hp, wscale(wnormalize(hn), 1e-3f));
        ray.d = wi;
        ray.tmin = 1e-3f;
        ray.tmax = 1e30f;
        thr = wscale(nthr, rr_scale);
        rng_hi = rng;
    }
    WF_RETURN(16u);
}
#undef WF_RETURN

static inline float srgb_encode(float c) { /* tonemap.rs:12-18; powf pinned */
    if (c <= 0.0031308f) return 12.92f * c;
    return 1.055f * f3do_pow(c, 1.0f / 2.4f) - 0.055f;
}

int f3do_wavefront_render(const f3do_wavefront_scene* D, uint32_t W, uint32_t H, uint32_t spp_frames, uint32_t first_frame,
                          uint32_t num_frames, float* accum_io, float* hdr_out, uint8_t* rgba8_out, uint64_t* stats_out) {
    g_wf_err[0] = 0;
    if (W == 0 || H == 0 || spp_frames == 0) return wf_fail("adjudication PT reference requires non-zero width/height/spp");
    if (D->nspheres == 0) return wf_fail("the scene needs at least one sphere / material slot");
    for (uint32_t i = 0; i < D->nspheres; i++) {
        const float* s = D->spheres + 20 * (size_t)i;
        if (fabsf(fmaxf(0.002f, s[15]) - fmaxf(0.002f, s[16])) >= 1e-4f) return wf_fail("anisotropic GGX (ax != ay) is not supported");
    }
    wf_scene S;
    S.sph = D->spheres; S.nsph = D->nspheres;
    S.dirl = D->dir_lights; S.ndir = D->ndir;
    S.areal = D->area_lights; S.narea = D->narea;
    S.imp = D->importance; S.nimp = D->nimportance;
    S.env = D->environment;
    S.xyz = D->mesh_xyz; S.nverts = D->mesh_nverts; S.idx = D->mesh_idx; S.ntris = D->mesh_ntris;
    S.inst = D->instances; S.ninst = D->ninstances;
    for (uint32_t k = 0; k < 3u * S.ntris; k++)
        if (S.idx[k] >= S.nverts) return wf_fail("mesh index %u out of range", S.idx[k]);
    const size_t npx = (size_t)W * H;
    const float aspect = (float)W / (float)H;
    float hs, hc;
    f3do_sincos(0.5f * D->fov_y_rad, &hs, &hc);
    const float half_h = hs / hc;
    const uint64_t capacity = 4ull * npx;
    uint64_t total_rays = 0, max_rays = 0;
    uint32_t min_iters = 0xFFFFFFFFu;
    uint64_t depth_hist[16] = {0};
    for (uint32_t fr = first_frame; fr < first_frame + num_frames; fr++) {
        const uint32_t seed_hi = f3do_wavefront_splitmix32(D->seed_hi ^ fr);
        const uint32_t seed_lo = f3do_wavefront_splitmix32(D->seed_lo ^ (fr * 0x00009E3Du));
        float u1, u2;
        f3do_wavefront_sobol2(fr, &u1, &u2); /* sidx = sample + frame * max(1, spp) with spp = 1 */
        uint64_t iters[16] = {0};
        uint64_t rays = 0;
#pragma omp parallel
        {
            uint64_t li[16] = {0};
            uint64_t lr = 0;
#pragma omp for schedule(dynamic, 1024) nowait
            for (size_t p = 0; p < npx; p++) lr += trace_pixel(&S, D, W, H, (uint32_t)p, fr, seed_hi, seed_lo, u1, u2, half_h, aspect, accum_io, li);
#pragma omp critical
            {
                rays += lr;
                for (int k = 0; k < 16; k++) iters[k] += li[k];
            }
        }
        uint32_t executed = 0;
        uint64_t cum = 0;
        for (int k = 0; k < 16; k++) {
            if (!iters[k]) break;
            cum += iters[k];
            if (cum > capacity) return wf_fail("wavefront frame %u: wavefront ray queue overflow: %llu rays pushed into capacity %llu", fr,
                                               (unsigned long long)cum, (unsigned long long)capacity);
            executed++;
        }
        if (executed < 2u)
            return wf_fail("adjudication PT frame %u executed %u wavefront iteration(s); a multi-bounce path-traced reference requires >= 2",
                           fr, executed);
        total_rays += rays;
        for (int k = 0; k < 16; k++) depth_hist[k] += iters[k];
        if (rays > max_rays) max_rays = rays;
        if (executed < min_iters) min_iters = executed;
    }
    if (stats_out) {
        stats_out[0] = total_rays; stats_out[1] = max_rays; stats_out[2] = min_iters;
        for (int k = 0; k < 16; k++) stats_out[3 + k] = depth_hist[k]; /* rays traced at each depth, summed over the frames */
    }
    if (hdr_out || rgba8_out) { /* adjudication.rs:318-331, tonemap.rs:11-32 */
        const float inv = 1.0f / (float)spp_frames;
        for (size_t p = 0; p < npx; p++) {
            for (int c = 0; c < 3; c++) {
                float v = accum_io[4 * p + c] * inv;
                if (hdr_out) hdr_out[4 * p + c] = v;
                if (rgba8_out) {
                    float x = fmaxf(v, 0.0f) * D->exposure;
                    float t = x / (1.0f + x);
                    float s = fminf(fmaxf(srgb_encode(t), 0.0f), 1.0f);
                    rgba8_out[4 * p + c] = (uint8_t)(s * 255.0f + 0.5f);
                }
            }
            if (hdr_out) hdr_out[4 * p + 3] = 1.0f;
            if (rgba8_out) rgba8_out[4 * p + 3] = 255u;
        }
    }
    return 0;
}
