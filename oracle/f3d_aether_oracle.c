/*
 * oracle/f3d_aether_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Plain-C f32 restatement of the AETHER aerial-perspective post of the path-traced DEM snapshot
 * (SURVEY section 8f row 1).  Reference sources followed (paths relative to /root/reference):
 *   src/shaders/atmosphere/prometheus_aerial.wgsl:51-231       (LUT loads, `main`)
 *   src/shaders/atmosphere/evaluation_core.wgsl:9-344          (spectral basis, LUT coordinates,
 *                                                               quadrilinear lookup, segment integral)
 *   src/shaders/includes/tonemap_common.wgsl:18-21             (tonemap_reinhard)
 *   src/shaders/includes/determinism.wgsl:324-327              (det_exp = exp2(x * log2 e))
 *   src/path_tracing/hybrid_compute/aether_post.rs:40-180,365-397 (uniforms, LUT validation, texel layout)
 *
 * Numerics contract extended for this pass: WGSL `exp2` is driver-defined; this
 * project pins it to the Cephes single-precision exp2f kernel written out below (f3do_exp2), `round`
 * is round-half-to-even (WGSL spec), `clamp(x,a,b) = min(max(x,a),b)`, `fract(x) = x - floor(x)`,
 * vector dot/normalize as in f3d_oracle.c.  No FMA contraction.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "f3d_oracle.h"

typedef struct { float x, y, z; } a3;
typedef struct { float x, y, z, w; } a4;

static inline a3 A3(float x, float y, float z) { a3 r = {x, y, z}; return r; }
static inline float adot(a3 a, a3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline a3 anormalize(a3 a) {
    float inv = 1.0f / sqrtf(adot(a, a));
    return A3(a.x * inv, a.y * inv, a.z * inv);
}
static inline float aclamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* Pinned exp2 (Cephes exp2f): round-to-floor split, |f| <= 0.5, degree-5 polynomial, exact scaling.
 * x >= 128 -> +inf, x < -126 -> 0 (no denormal results), NaN -> NaN. */
float f3do_exp2(float x) {
    if (x != x) return x;
    if (x >= 128.0f) return INFINITY;
    if (x < -126.0f) return 0.0f;
    float px = floorf(x);
    int32_t i0 = (int32_t)px;
    float f = x - px;
    if (f > 0.5f) { i0 += 1; f = f - 1.0f; }
    float p = 1.535336188319500e-4f;
    p = p * f + 1.339887440266574e-3f;
    p = p * f + 9.618437357674640e-3f;
    p = p * f + 5.550332471162809e-2f;
    p = p * f + 2.402264791363012e-1f;
    p = p * f + 6.931472028550421e-1f;
    float r = 1.0f + f * p;
    /* r * 2^i0 in two exact steps (both factors are normal numbers for i0 in [-126, 128]) */
    int32_t e1 = i0 >> 1, e2 = i0 - e1;
    uint32_t b1 = (uint32_t)(e1 + 127) << 23, b2 = (uint32_t)(e2 + 127) << 23;
    float s1, s2;
    memcpy(&s1, &b1, 4);
    memcpy(&s2, &b2, 4);
    return (r * s1) * s2;
}

static inline float det_exp(float x) { return f3do_exp2(x * 1.4426950408889634f); } /* determinism.wgsl:325-327 */

/* evaluation_core.wgsl:9-27 */
static const float WAVELENGTHS_NM[11] = {380.0f, 420.0f, 460.0f, 500.0f, 540.0f, 580.0f, 620.0f, 660.0f, 700.0f, 740.0f, 780.0f};
static const float CIE_XYZ[11][3] = {
    {0.001368f, 0.000039f, 0.006450f}, {0.134380f, 0.004000f, 0.645600f}, {0.290800f, 0.060000f, 1.669200f},
    {0.004900f, 0.323000f, 0.272000f}, {0.290400f, 0.954000f, 0.020300f}, {0.916300f, 0.870000f, 0.001650f},
    {0.854450f, 0.381000f, 0.000190f}, {0.164900f, 0.061000f, 0.000000f}, {0.011359f, 0.004102f, 0.000000f},
    {0.000690f, 0.000249f, 0.000000f}, {0.000042f, 0.000015f, 0.000000f},
};

static inline float clamp_radiometric_scale(float v) { return fminf(fmaxf(v, 0.0f), 65504.0f); }  /* :29-33 */
static inline a3 clamp_hdr(a3 c) {                                                                /* :35-37 */
    return A3(fminf(fmaxf(c.x, 0.0f), 65504.0f), fminf(fmaxf(c.y, 0.0f), 65504.0f), fminf(fmaxf(c.z, 0.0f), 65504.0f));
}

static a3 xyz_to_rgb(a3 xyz) { /* :39-45 */
    return A3(adot(A3(3.2404542f, -1.5371385f, -0.4985314f), xyz) / 3.2613921f,
              adot(A3(-0.9692660f, 1.8760108f, 0.0415560f), xyz) / 2.5069624f,
              adot(A3(0.0556434f, -0.2040259f, 1.0572252f), xyz) / 2.3679786f);
}

static a3 spectral_xyz(uint32_t i, float rayleigh_column, float mie_column, float ozone_column, float turbidity) { /* :47-76 */
    float lambda_nm = WAVELENGTHS_NM[i];
    float ratio = 550.0f / lambda_nm;
    float ratio2 = ratio * ratio;
    float rayleigh_beta = 1.2989e-5f * ratio2 * ratio2;
    float mie_beta = 1.0e-5f * turbidity * ratio;
    float delta = (lambda_nm - 600.0f) / 85.0f;
    float ozone_beta = 1.2e-6f * det_exp(-0.5f * delta * delta);
    float tau = rayleigh_beta * rayleigh_column + mie_beta * mie_column + ozone_beta * ozone_column;
    float spectral_t = det_exp(-fmaxf(tau, 0.0f));
    float w = (i == 0u || i + 1u == 11u) ? 0.5f : 1.0f;
    return A3(CIE_XYZ[i][0] * spectral_t * w, CIE_XYZ[i][1] * spectral_t * w, CIE_XYZ[i][2] * spectral_t * w);
}

static float mu_to_unit(float mu) { /* :78-86 */
    float bounded = aclamp(mu, -1.0f, 1.0f);
    float magnitude = sqrtf(fabsf(bounded));
    float signed_root = bounded >= 0.0f ? magnitude : -magnitude;
    return 0.5f * (signed_root + 1.0f);
}
static float nu_to_unit(float nu) { return 1.0f - sqrtf(fmaxf(0.5f * (1.0f - aclamp(nu, -1.0f, 1.0f)), 0.0f)); } /* :88-90 */
static float scattering_height_to_unit(float h) { return sqrtf(aclamp(h, 0.0f, 1.0f)); }                       /* :92-96 */

static inline int32_t iclamp(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* texel (x, y, z) of an RGBA16F LUT uploaded with bytes_per_row = w*8, rows_per_image = h (aether_post.rs:365-397) */
static a4 lut_texel(const uint16_t* lut, const uint32_t dims[3], int32_t x, int32_t y, int32_t z) {
    const uint16_t* p = lut + 4u * (((size_t)z * dims[1] + (size_t)y) * dims[0] + (size_t)x);
    a4 r = {f3do_f16_to_f32(p[0]), f3do_f16_to_f32(p[1]), f3do_f16_to_f32(p[2]), f3do_f16_to_f32(p[3])};
    return r;
}

static a4 load_scattering_texel(const f3do_atmosphere* A, int32_t view, int32_t sun, int32_t height, int32_t nu, int32_t nu_count) { /* :98-113 */
    const uint32_t* d = A->scattering_dims;
    return lut_texel(A->scattering, d, iclamp(view, 0, (int32_t)d[0] - 1), iclamp(sun, 0, (int32_t)d[1] - 1),
                     iclamp(height * nu_count + nu, 0, (int32_t)d[2] - 1));
}

static a3 sample_accumulated_scattering(const f3do_atmosphere* A, float height_unit, float mu_sun, float mu_view, float nu) { /* :115-170 */
    const uint32_t* d = A->scattering_dims;
    int32_t height_count = (int32_t)A->scattering_height > 2 ? (int32_t)A->scattering_height : 2;
    int32_t nu_count = (int32_t)A->scattering_nu > 2 ? (int32_t)A->scattering_nu : 2;
    float c[4] = {mu_to_unit(mu_view) * (float)(d[0] - 1u), mu_to_unit(mu_sun) * (float)(d[1] - 1u),
                  scattering_height_to_unit(height_unit) * (float)(height_count - 1), nu_to_unit(nu) * (float)(nu_count - 1)};
    int32_t hi_lim[4] = {(int32_t)d[0] - 1, (int32_t)d[1] - 1, height_count - 1, nu_count - 1};
    int32_t lower[4], upper[4];
    float fr[4];
    for (int k = 0; k < 4; k++) {
        float fl = floorf(c[k]);
        lower[k] = (int32_t)fl;
        upper[k] = lower[k] + 1 < hi_lim[k] ? lower[k] + 1 : hi_lim[k];
        fr[k] = c[k] - fl;
    }
    a4 acc = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int hs = 0; hs < 2; hs++)
        for (int ns = 0; ns < 2; ns++)
            for (int ss = 0; ss < 2; ss++)
                for (int vs = 0; vs < 2; vs++) {
                    float weight = (vs ? fr[0] : 1.0f - fr[0]) * (ss ? fr[1] : 1.0f - fr[1]) * (hs ? fr[2] : 1.0f - fr[2]) *
                                   (ns ? fr[3] : 1.0f - fr[3]);
                    a4 t = load_scattering_texel(A, vs ? upper[0] : lower[0], ss ? upper[1] : lower[1], hs ? upper[2] : lower[2],
                                                 ns ? upper[3] : lower[3], nu_count);
                    acc.x = acc.x + weight * t.x;
                    acc.y = acc.y + weight * t.y;
                    acc.z = acc.z + weight * t.z;
                    acc.w = acc.w + weight * t.w;
                }
    return A3(fmaxf(acc.x, 0.0f), fmaxf(acc.y, 0.0f), fmaxf(acc.z, 0.0f));
}

static float spherical_radius_m(float camera_height_m, float view_mu, float distance_m, float bottom_radius_m) { /* :172-186 */
    float radius_m = fmaxf(bottom_radius_m, 1.0f) + aclamp(camera_height_m, 0.0f, 100000.0f);
    float bd = aclamp(distance_m, 0.0f, 20000000.0f);
    float radial_squared = fmaxf(radius_m * radius_m + bd * bd + 2.0f * radius_m * bd * aclamp(view_mu, -1.0f, 1.0f), 0.0f);
    return sqrtf(radial_squared);
}
static float spherical_altitude(float camera_height_m, float view_mu, float distance_m, float bottom_radius_m) { /* :188-198 */
    float r = spherical_radius_m(camera_height_m, view_mu, distance_m, bottom_radius_m);
    return aclamp(r - fmaxf(bottom_radius_m, 1.0f), 0.0f, 100000.0f);
}
static void spherical_endpoint_mus(float camera_height_m, float view_mu, float sun_mu, float view_sun_nu, float distance_m,
                                   float bottom_radius_m, float* end_view_mu, float* end_sun_mu) { /* :200-223 */
    float radius_m = fmaxf(bottom_radius_m, 1.0f) + aclamp(camera_height_m, 0.0f, 100000.0f);
    float bd = aclamp(distance_m, 0.0f, 20000000.0f);
    float er = fmaxf(spherical_radius_m(camera_height_m, view_mu, bd, bottom_radius_m), 1.0f);
    float ev = (radius_m * aclamp(view_mu, -1.0f, 1.0f) + bd) / er;
    float es = (radius_m * aclamp(sun_mu, -1.0f, 1.0f) + bd * aclamp(view_sun_nu, -1.0f, 1.0f)) / er;
    *end_view_mu = aclamp(ev, -1.0f, 1.0f);
    *end_sun_mu = aclamp(es, -1.0f, 1.0f);
}

static a3 segment_transmittance(float distance_m, float camera_height_m, float view_mu, float bottom_radius_m, float density_scale,
                                float turbidity, float ozone_du) { /* :225-344 */
    float bd = aclamp(distance_m, 0.0f, 20000000.0f);
    float bh = aclamp(camera_height_m, 0.0f, 100000.0f);
    float h[16];
    for (int k = 0; k < 16; k++) h[k] = spherical_altitude(bh, view_mu, bd * ((float)(2 * k + 1) * 0.03125f), bottom_radius_m);
    float rayleigh = det_exp(-h[0] / 8000.0f), mie = det_exp(-h[0] / 1200.0f);
    float ozone = fmaxf(1.0f - fabsf((h[0] - 25000.0f) / 15000.0f), 0.0f);
    for (int k = 1; k < 16; k++) {
        rayleigh = rayleigh + det_exp(-h[k] / 8000.0f);
        mie = mie + det_exp(-h[k] / 1200.0f);
        ozone = ozone + fmaxf(1.0f - fabsf((h[k] - 25000.0f) / 15000.0f), 0.0f);
    }
    float path_per_sample = bd * density_scale * 0.0625f;
    float rayleigh_column = path_per_sample * rayleigh;
    float mie_column = path_per_sample * mie;
    float ozone_column = path_per_sample * ozone * ozone_du / 300.0f;
    a3 xyz = spectral_xyz(0u, rayleigh_column, mie_column, ozone_column, turbidity);
    for (uint32_t i = 1u; i < 11u; i++) {
        a3 s = spectral_xyz(i, rayleigh_column, mie_column, ozone_column, turbidity);
        xyz = A3(xyz.x + s.x, xyz.y + s.y, xyz.z + s.z);
    }
    a3 rgb = xyz_to_rgb(xyz);
    return A3(aclamp(rgb.x, 0.0f, 1.0f), aclamp(rgb.y, 0.0f, 1.0f), aclamp(rgb.z, 0.0f, 1.0f));
}

/* prometheus_aerial.wgsl:51-61.  WGSL round() is round-half-to-even == rintf under the default rounding mode. */
static a3 load_boundary_transmittance(const f3do_atmosphere* A, float height_unit, float mu) {
    const uint32_t d[3] = {A->transmittance_dims[0], A->transmittance_dims[1], 1u};
    uint32_t dx = d[0] > 1u ? d[0] : 1u, dy = d[1] > 1u ? d[1] : 1u;
    int32_t x = (int32_t)rintf((0.5f * (aclamp(mu, -1.0f, 1.0f) + 1.0f)) * (float)(dx - 1u));
    int32_t y = (int32_t)rintf(aclamp(height_unit, 0.0f, 1.0f) * (float)(dy - 1u));
    a4 t = lut_texel(A->transmittance, d, x, y, 0);
    return A3(aclamp(t.x, 0.0f, 1.0f), aclamp(t.y, 0.0f, 1.0f), aclamp(t.z, 0.0f, 1.0f));
}

/* prometheus_aerial.wgsl:84-96 */
static float load_aerial_transmittance(const f3do_atmosphere* A, float distance_unit, float height_unit, float mu_view) {
    const uint32_t* d = A->aerial_dims;
    uint32_t dx = d[0] > 1u ? d[0] : 1u, dy = d[1] > 1u ? d[1] : 1u, dz = d[2] > 1u ? d[2] : 1u;
    int32_t x = (int32_t)rintf(aclamp(distance_unit, 0.0f, 1.0f) * (float)(dx - 1u));
    int32_t y = (int32_t)rintf(0.5f * (aclamp(mu_view, -1.0f, 1.0f) + 1.0f) * (float)(dy - 1u));
    int32_t z = (int32_t)rintf(aclamp(height_unit, 0.0f, 1.0f) * (float)(dz - 1u));
    return aclamp(lut_texel(A->aerial, d, x, y, z).w, 0.0f, 1.0f);
}

static inline a3 reinhard(a3 c) { return A3(c.x / (1.0f + c.x), c.y / (1.0f + c.y), c.z / (1.0f + c.z)); } /* tonemap_common.wgsl:18-21 */

/* validate_luts + AtmosphereConfig::validate (aether_post.rs:345-363, src/core/atmosphere/bake.rs:165-215).
 * Returns NULL when valid, else the reference's message. */
const char* f3do_aether_validate(const f3do_atmosphere* A) {
    const float v[9] = {A->turbidity, A->ozone_du, A->mie_g, A->bottom_radius_m, A->top_radius_m, A->rayleigh_scale_height_m,
                        A->mie_scale_height_m, A->max_aerial_distance_m, A->ground_albedo};
    for (int i = 0; i < 9; i++)
        if (!isfinite(v[i])) return "invalid AETHER PT settings: invalid atmosphere configuration: all scalar parameters must be finite";
    if (!(A->turbidity >= 1.0f && A->turbidity <= 10.0f)) return "invalid AETHER PT settings: invalid atmosphere configuration: turbidity must be in [1, 10]";
    if (!(A->ozone_du >= 0.0f && A->ozone_du <= 600.0f)) return "invalid AETHER PT settings: invalid atmosphere configuration: ozone must be in [0, 600] DU";
    if (!(A->mie_g >= 0.0f && A->mie_g <= 0.99f)) return "invalid AETHER PT settings: invalid atmosphere configuration: mie_g must be in [0, 0.99]";
    if (A->bottom_radius_m <= 0.0f || A->top_radius_m <= A->bottom_radius_m)
        return "invalid AETHER PT settings: invalid atmosphere configuration: top radius must exceed a positive bottom radius";
    if (A->rayleigh_scale_height_m <= 0.0f || A->mie_scale_height_m <= 0.0f || A->max_aerial_distance_m <= 0.0f)
        return "invalid AETHER PT settings: invalid atmosphere configuration: scale heights and aerial distance must be positive";
    if (!(A->ground_albedo >= 0.0f && A->ground_albedo <= 1.0f)) return "invalid AETHER PT settings: invalid atmosphere configuration: ground albedo must be in [0, 1]";
    const uint32_t axes[9] = {A->transmittance_dims[0], A->transmittance_dims[1], A->scattering_dims[0], A->scattering_dims[1],
                              A->scattering_height, A->scattering_nu, A->aerial_dims[0], A->aerial_dims[1], A->aerial_dims[2]};
    for (int i = 0; i < 9; i++)   /* LutDimensions::validate, bake.rs:75-100 */
        if (axes[i] < 2u) return "invalid AETHER PT settings: invalid atmosphere configuration: every atmosphere LUT axis must contain at least two samples";
    for (int i = 0; i < 9; i++)
        if (axes[i] > 256u) return "invalid AETHER PT settings: invalid atmosphere configuration: atmosphere LUT axes are capped at 256 samples";
    if (!A->transmittance || !A->scattering || !A->aerial) return "PROMETHEUS AETHER LUT dimensions do not match metadata";
    const uint64_t packed = (uint64_t)A->scattering_height * A->scattering_nu;
    if (packed > 0xFFFFFFFFull) return "AETHER scattering depth overflow";
    if (A->scattering_dims[2] != (uint32_t)packed) return "PROMETHEUS AETHER LUT dimensions do not match metadata";
    return 0;
}

/* prometheus_aerial.wgsl:98-231 for one pixel; returns the RGBA16F texel written by textureStore. */
static void aether_pixel(const f3do_atmosphere* A, const f3do_aether_view* V, uint32_t gx, uint32_t gy, const float accum[4],
                         float depth, int visible, uint16_t out[4]) {
    a3 surface = clamp_hdr(A3(accum[0] / fmaxf(accum[3], 1.0f), accum[1] / fmaxf(accum[3], 1.0f), accum[2] / fmaxf(accum[3], 1.0f)));
    float ndc_x = (((float)gx + 0.5f) / (float)V->width) * 2.0f - 1.0f;
    float ndc_y = (1.0f - ((float)gy + 0.5f) / (float)V->height) * 2.0f - 1.0f;
    float sx = ndc_x * V->tan_half_fov * V->aspect, sy = ndc_y * V->tan_half_fov;
    a3 ray = anormalize(A3((V->cam_right[0] * sx + V->cam_up[0] * sy) + V->cam_forward[0],
                           (V->cam_right[1] * sx + V->cam_up[1] * sy) + V->cam_forward[1],
                           (V->cam_right[2] * sx + V->cam_up[2] * sy) + V->cam_forward[2]));
    a3 sun_dir = anormalize(A3(V->light_dir[0], V->light_dir[1], V->light_dir[2]));
    float sun_intensity = clamp_radiometric_scale(V->sun_intensity);
    float exposure = clamp_radiometric_scale(V->exposure);
    float atmosphere_height = fmaxf(A->top_radius_m - A->bottom_radius_m, 1.0f);
    float camera_height = fmaxf(V->cam_origin[1], 0.0f);
    float camera_height_unit = aclamp(camera_height / atmosphere_height, 0.0f, 1.0f);
    float view_sun_nu = adot(ray, sun_dir);
    a3 ldr;
    if (!visible) {
        a3 s = sample_accumulated_scattering(A, camera_height_unit, sun_dir.y, ray.y, view_sun_nu);
        a3 miss = clamp_hdr(A3(s.x * sun_intensity, s.y * sun_intensity, s.z * sun_intensity));
        ldr = reinhard(A3(miss.x * exposure, miss.y * exposure, miss.z * exposure));
    } else {
        float endpoint_height = spherical_altitude(camera_height, ray.y, depth, A->bottom_radius_m);
        float end_view_mu, end_sun_mu;
        spherical_endpoint_mus(camera_height, ray.y, sun_dir.y, view_sun_nu, depth, A->bottom_radius_m, &end_view_mu, &end_sun_mu);
        a3 seg = segment_transmittance(depth, camera_height, ray.y, A->bottom_radius_m, 1.0f, A->turbidity, A->ozone_du);
        a3 boundary_t = load_boundary_transmittance(A, camera_height_unit, ray.y);
        a3 cs = sample_accumulated_scattering(A, camera_height_unit, sun_dir.y, ray.y, view_sun_nu);
        cs = A3(cs.x * sun_intensity, cs.y * sun_intensity, cs.z * sun_intensity);
        float endpoint_height_unit = aclamp(endpoint_height / atmosphere_height, 0.0f, 1.0f);
        a3 es = sample_accumulated_scattering(A, endpoint_height_unit, end_sun_mu, end_view_mu, view_sun_nu);
        es = A3(es.x * sun_intensity, es.y * sun_intensity, es.z * sun_intensity);
        float distance_unit = depth / fmaxf(A->max_aerial_distance_m, 1.0f);
        float aerial_mean = load_aerial_transmittance(A, distance_unit, camera_height_unit, ray.y);
        float analytic_mean = adot(seg, A3(0.2126f, 0.7152f, 0.0722f));
        float k = aerial_mean / fmaxf(analytic_mean, 1.0e-6f);
        a3 tr = A3(fmaxf(aclamp(seg.x * k, 0.0f, 1.0f), boundary_t.x), fmaxf(aclamp(seg.y * k, 0.0f, 1.0f), boundary_t.y),
                   fmaxf(aclamp(seg.z * k, 0.0f, 1.0f), boundary_t.z));
        a3 fin = A3(fmaxf(cs.x - tr.x * es.x, 0.0f), fmaxf(cs.y - tr.y * es.y, 0.0f), fmaxf(cs.z - tr.z * es.z, 0.0f));
        a3 hdr = clamp_hdr(A3(surface.x * tr.x + fin.x, surface.y * tr.y + fin.y, surface.z * tr.z + fin.z));
        ldr = reinhard(A3(hdr.x * exposure, hdr.y * exposure, hdr.z * exposure));
    }
    out[0] = f3do_f32_to_f16(ldr.x);
    out[1] = f3do_f32_to_f16(ldr.y);
    out[2] = f3do_f32_to_f16(ldr.z);
    out[3] = f3do_f32_to_f16(1.0f);
}

int f3do_aether_post(const f3do_atmosphere* A, const f3do_aether_view* V, const float* accum, const float* depth,
                     const uint8_t* visibility, uint16_t* out_rgba16f) {
    if (!A || !V || !accum || !depth || !visibility || !out_rgba16f) return 1;
    if (f3do_aether_validate(A)) return 1;
    for (uint32_t y = 0; y < V->height; y++)
        for (uint32_t x = 0; x < V->width; x++) {
            size_t i = (size_t)y * V->width + x;
            aether_pixel(A, V, x, y, accum + 4 * i, depth[i], visibility[i] != 0, out_rgba16f + 4 * i);
        }
    return 0;
}
