/*
 * oracle/f3d_smoke_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Plain-C f32 restatement of the reference's smoke volume ray-march (BASELINE config 4):
 *   src/smoke/render.rs:6-101    SmokeVolume::raymarch_rgba          (perspective camera)
 *   src/smoke/render.rs:103-175  SmokeVolume::raymarch_projection_rgba (map-aligned parallel projection)
 *   src/smoke/render.rs:177-276  sample_render_fields, march_ray_rgba
 *   src/smoke/render.rs:278-316  sun_transmittance
 *   src/smoke/render.rs:328-401  smoke_color, ray_box_intersection, henyey_greenstein, tone_map, render_smoothstep, to_u8
 *   src/smoke/sampling.rs:1-32,83-103  sample_scalar (trilinear), index, lerp, hash01
 *   src/smoke/types.rs:268-317,387-405  settings validation, grid_coord_from_world, bounds
 * The reference runs this single-threaded on the CPU; pixels are independent, so the OpenMP loop here changes nothing.
 *
 * Numerics: Rust f32 semantics op for op (no FMA contraction).  `f32::clamp` propagates NaN (rs_clamp below),
 * `f32::min/max` ignore NaN (fminf/fmaxf), `as u8` saturates and maps NaN to 0, glam 0.24 Vec3 is scalar code
 * (dot = x*x + y*y + z*z left to right, normalize = v * (1/length)).  Two libm calls are pinned, as DESIGN.md section 4
 * does for the shader intrinsics: exp(x) = f3do_exp2(x * log2(e)) and powf(d, 1.5) = d * sqrt(d); `tan` (host set-up
 * only) stays libm tanf.  Pin status: the reference commits no image for this path; the oracle is checked against the
 * properties its own unit tests assert (render.rs:408-592) on the same scenes, see tests/test_smoke.py.
 */
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "f3d_oracle.h"

typedef struct { float x, y, z; } s3;
static inline s3 S3(float x, float y, float z) { s3 r = {x, y, z}; return r; }
static inline s3 sadd(s3 a, s3 b) { return S3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline s3 ssub(s3 a, s3 b) { return S3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline s3 smul(s3 a, s3 b) { return S3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline s3 sscale(s3 a, float s) { return S3(a.x * s, a.y * s, a.z * s); }
static inline float sdot(s3 a, s3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline s3 scross(s3 a, s3 b) { return S3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
static inline float slength(s3 a) { return sqrtf(sdot(a, a)); }
static inline s3 snormalize(s3 a) { return sscale(a, 1.0f / slength(a)); }                 /* glam Vec3::normalize */
static inline s3 snormalize_or_zero(s3 a) {                                                /* glam Vec3::normalize_or_zero */
    float rcp = 1.0f / slength(a);
    if (isfinite(rcp) && rcp > 0.0f) return sscale(a, rcp);
    return S3(0.0f, 0.0f, 0.0f);
}
static inline float rs_clamp(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }   /* f32::clamp */
static inline float lerpf(float a, float b, float t) { return a + (b - a) * t; }           /* sampling.rs:87-89 */
static inline s3 mix3(s3 a, s3 b, float t) { return S3(lerpf(a.x, b.x, t), lerpf(a.y, b.y, t), lerpf(a.z, b.z, t)); }
static inline float pinned_exp(float x) { return f3do_exp2(x * 1.4426950408889634f); }
static inline uint8_t to_u8(float v) {                                                     /* render.rs:399-401 */
    float c = rs_clamp(v, 0.0f, 1.0f) * 255.0f + 0.5f;
    if (!(c == c)) return 0;
    return c >= 255.0f ? 255 : (c <= 0.0f ? 0 : (uint8_t)c);
}
static inline float render_smoothstep(float e0, float e1, float x) {                       /* render.rs:386-389 */
    float t = rs_clamp((x - e0) / fmaxf(e1 - e0, 1.0e-6f), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
static inline float hash01(uint32_t v) {                                                   /* sampling.rs:96-103 */
    v ^= v >> 16; v *= 0x7FEB352Du; v ^= v >> 15; v *= 0x846CA68Bu; v ^= v >> 16;
    return (float)v / 4294967296.0f;                                                       /* u32::MAX as f32 == 2^32 */
}

static _Thread_local char g_smoke_err[256];
const char* f3do_smoke_last_error(void) { return g_smoke_err; }
static int sfail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_smoke_err, sizeof g_smoke_err, fmt, ap);
    va_end(ap);
    return 1;
}

/* SmokeRenderSettings::validate, types.rs:268-317 */
static int validate_settings(const f3do_smoke_settings* s) {
    const char* names[11] = {"density_scale", "extinction", "scattering", "absorption", "phase_g", "step_size", "shadow_step_size",
                             "jitter_strength", "exposure", "soot_absorption", "fire_glow"};
    const float vals[11] = {s->density_scale, s->extinction, s->scattering, s->absorption, s->phase_g, s->step_size,
                            s->shadow_step_size, s->jitter_strength, s->exposure, s->soot_absorption, s->fire_glow};
    for (int i = 0; i < 11; i++)
        if (!isfinite(vals[i])) return sfail("%s must be finite", names[i]);
    if (s->density_scale < 0.0f || s->extinction < 0.0f || s->scattering < 0.0f) return sfail("density_scale, extinction, and scattering must be >= 0");
    if (s->absorption < 0.0f || s->soot_absorption < 0.0f || s->fire_glow < 0.0f) return sfail("absorption, soot_absorption, and fire_glow must be >= 0");
    if (!(s->phase_g >= -0.99f && s->phase_g <= 0.99f)) return sfail("phase_g must be in [-0.99, 0.99]");
    if (s->step_size < 0.0f || s->shadow_step_size < 0.0f) return sfail("step sizes must be >= 0");
    if (s->max_steps == 0 || s->shadow_steps == 0) return sfail("max_steps and shadow_steps must be >= 1");
    if (!(s->jitter_strength >= 0.0f && s->jitter_strength <= 1.0f)) return sfail("jitter_strength must be in [0, 1]");
    for (int a = 0; a < 3; a++)
        if (!isfinite(s->thin_color[a]) || s->thin_color[a] < 0.0f) return sfail("thin_color[%d] must be finite and >= 0", a);
    for (int a = 0; a < 3; a++)
        if (!isfinite(s->dense_color[a]) || s->dense_color[a] < 0.0f) return sfail("dense_color[%d] must be finite and >= 0", a);
    return 0;
}

static float sample_scalar(const float* field, const uint32_t dims[3], const float p[3]) {   /* sampling.rs:1-32 */
    if (!field) return 0.0f;
    float x = rs_clamp(p[0], 0.0f, (float)(dims[0] - 1u));
    float y = rs_clamp(p[1], 0.0f, (float)(dims[1] - 1u));
    float z = rs_clamp(p[2], 0.0f, (float)(dims[2] - 1u));
    size_t x0 = (size_t)floorf(x), y0 = (size_t)floorf(y), z0 = (size_t)floorf(z);
    size_t x1 = x0 + 1 < dims[0] - 1u ? x0 + 1 : dims[0] - 1u;
    size_t y1 = y0 + 1 < dims[1] - 1u ? y0 + 1 : dims[1] - 1u;
    size_t z1 = z0 + 1 < dims[2] - 1u ? z0 + 1 : dims[2] - 1u;
    float fx = x - (float)x0, fy = y - (float)y0, fz = z - (float)z0;
#define IDX(X, Y, Z) (((Z) * dims[1] + (Y)) * dims[0] + (X))
    float c00 = lerpf(field[IDX(x0, y0, z0)], field[IDX(x1, y0, z0)], fx);
    float c10 = lerpf(field[IDX(x0, y1, z0)], field[IDX(x1, y1, z0)], fx);
    float c01 = lerpf(field[IDX(x0, y0, z1)], field[IDX(x1, y0, z1)], fx);
    float c11 = lerpf(field[IDX(x0, y1, z1)], field[IDX(x1, y1, z1)], fx);
#undef IDX
    return lerpf(lerpf(c00, c10, fy), lerpf(c01, c11, fy), fz);
}

typedef struct { float density, temperature, soot, humidity, emission, age; } render_sample;

static render_sample sample_render_fields(const f3do_smoke_volume* V, s3 pos) {             /* render.rs:177-188 */
    float p[3] = {(pos.x - V->origin[0]) / V->voxel_size[0] - 0.5f, (pos.y - V->origin[1]) / V->voxel_size[1] - 0.5f,
                  (pos.z - V->origin[2]) / V->voxel_size[2] - 0.5f};                          /* types.rs:387-393 */
    render_sample r;
    r.density = sample_scalar(V->density, V->dims, p);
    r.temperature = sample_scalar(V->temperature, V->dims, p);
    r.soot = sample_scalar(V->soot, V->dims, p);
    r.humidity = sample_scalar(V->humidity, V->dims, p);
    r.emission = sample_scalar(V->emission_rate, V->dims, p);
    r.age = fmaxf(sample_scalar(V->particle_age, V->dims, p), 0.0f);
    return r;
}

static s3 bounds_max(const f3do_smoke_volume* V) {                                           /* types.rs:399-405 */
    return S3(V->origin[0] + (float)V->dims[0] * V->voxel_size[0], V->origin[1] + (float)V->dims[1] * V->voxel_size[1],
              V->origin[2] + (float)V->dims[2] * V->voxel_size[2]);
}

static int ray_box(s3 o, s3 d, s3 bmin, s3 bmax, float* near_out, float* far_out) {          /* render.rs:348-378 */
    s3 inv = S3(fabsf(d.x) > 1.0e-12f ? 1.0f / d.x : INFINITY, fabsf(d.y) > 1.0e-12f ? 1.0f / d.y : INFINITY,
                fabsf(d.z) > 1.0e-12f ? 1.0f / d.z : INFINITY);
    s3 t0 = smul(ssub(bmin, o), inv), t1 = smul(ssub(bmax, o), inv);
    s3 tmin = S3(fminf(t0.x, t1.x), fminf(t0.y, t1.y), fminf(t0.z, t1.z));
    s3 tmax = S3(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y), fmaxf(t0.z, t1.z));
    float nr = fmaxf(fmaxf(tmin.x, tmin.y), tmin.z), fr = fminf(fminf(tmax.x, tmax.y), tmax.z);
    if (fr >= fmaxf(nr, 0.0f)) { *near_out = nr; *far_out = fr; return 1; }
    return 0;
}

static float henyey_greenstein(float cos_theta, float g) {                                   /* render.rs:380-384 */
    float g2 = g * g;
    float denom = fmaxf(1.0f + g2 - 2.0f * g * cos_theta, 1.0e-4f);
    return (1.0f - g2) / (4.0f * 3.14159274101257324f * (denom * sqrtf(denom)));              /* powf(denom, 1.5) pinned */
}

static s3 smoke_color(render_sample s, const f3do_smoke_settings* st) {                      /* render.rs:328-346 */
    float body = rs_clamp(s.density * 1.45f + s.soot * 1.35f, 0.0f, 1.0f);
    s3 color = mix3(S3(st->thin_color[0], st->thin_color[1], st->thin_color[2]),
                    S3(st->dense_color[0], st->dense_color[1], st->dense_color[2]), body);
    float aged = rs_clamp(s.age / 9.0f, 0.0f, 1.0f);
    color = mix3(color, S3(0.36f, 0.39f, 0.43f), aged * 0.42f);
    float humidity_milk = rs_clamp(s.humidity, 0.0f, 1.0f) * (0.18f + 0.42f * body);
    color = mix3(color, S3(0.93f, 0.92f, 0.84f), rs_clamp(humidity_milk, 0.0f, 0.38f));
    float freshness = rs_clamp(1.0f - s.age / 17.0f, 0.0f, 1.0f);
    float heat = rs_clamp(s.temperature * 0.12f * freshness, 0.0f, 1.0f);
    return mix3(color, S3(0.95f, 0.62f, 0.28f), heat * 0.07f);
}

static float sun_transmittance(const f3do_smoke_volume* V, s3 start, s3 sun_dir, float step, uint32_t steps,
                               const f3do_smoke_settings* st) {                              /* render.rs:278-316 */
    s3 bmin = S3(V->origin[0], V->origin[1], V->origin[2]), bmax = bounds_max(V);
    float t0, t1;
    if (!ray_box(sadd(start, sscale(sun_dir, step)), sun_dir, bmin, bmax, &t0, &t1)) return 1.0f;
    t0 = fmaxf(t0, 0.0f);
    float optical_depth = 0.0f;
    for (uint32_t i = 0; i < steps; i++) {
        float t = t0 + ((float)i + 0.5f) * step;
        if (t > t1) break;
        s3 p = sadd(start, sscale(sun_dir, step + t));
        render_sample s = sample_render_fields(V, p);
        float age_t = render_smoothstep(1.6f, 17.0f, s.age);
        float gate = 0.50f + 0.50f * render_smoothstep(0.045f, 0.34f, s.density);
        optical_depth += s.density * st->density_scale * (1.0f - 0.58f * age_t) * gate * st->extinction *
                         (1.0f + s.soot * st->soot_absorption) * step;
        if (optical_depth > 8.0f) break;
    }
    return rs_clamp(pinned_exp(-optical_depth), 0.0f, 1.0f);
}

float f3do_smoke_sun_transmittance(const f3do_smoke_volume* V, const f3do_smoke_settings* st, const float start[3],
                                   const float sun_dir[3], float step, uint32_t steps) {
    return sun_transmittance(V, S3(start[0], start[1], start[2]), S3(sun_dir[0], sun_dir[1], sun_dir[2]), step, steps, st);
}

static void march_ray_rgba(const f3do_smoke_volume* V, s3 origin, s3 dir, float t0, float t1, uint32_t jitter_seed, float step,
                           float shadow_step, s3 sun_dir, const f3do_smoke_settings* st, uint8_t out[4]) {   /* render.rs:190-276 */
    float jitter = (hash01(jitter_seed) - 0.5f) * st->jitter_strength * step;
    float t = fmaxf(t0 + jitter, 0.0f);
    float transmittance = 1.0f;
    s3 rgb = S3(0.0f, 0.0f, 0.0f);
    uint32_t steps = 0;
    const float by = fmaxf(bounds_max(V).y, 1.0f);
    while (t < t1 && steps < st->max_steps && transmittance > 0.01f) {
        s3 p = sadd(origin, sscale(dir, t));
        render_sample s = sample_render_fields(V, p);
        float age_t = render_smoothstep(1.6f, 17.0f, s.age);
        float gate = 0.50f + 0.50f * render_smoothstep(0.045f, 0.34f, s.density);
        float density = fmaxf(s.density * st->density_scale * (1.0f - 0.58f * age_t) * gate, 0.0f);
        if (density > 1.0e-5f) {
            float sigma_t = density * st->extinction * (1.0f + s.soot * st->soot_absorption * 0.85f);
            float seg_t = rs_clamp(pinned_exp(-sigma_t * step), 0.0f, 1.0f);
            float seg_w = sigma_t > 1.0e-6f ? (1.0f - seg_t) / sigma_t : step;
            float light = st->self_shadow ? sun_transmittance(V, p, sun_dir, shadow_step, st->shadow_steps, st) : 1.0f;
            float cos_theta = rs_clamp(sdot(dir, sun_dir), -1.0f, 1.0f);
            float phase = henyey_greenstein(cos_theta, st->phase_g);
            s3 color = smoke_color(s, st);
            float albedo = rs_clamp(st->scattering / (st->scattering + st->absorption + s.soot * 0.55f + 1.0e-5f), 0.02f, 0.98f);
            float sigma_s = sigma_t * albedo;
            s3 sun_radiance = sscale(S3(1.0f, 0.96f, 0.84f), 11.5f);
            s3 sky = sscale(sscale(S3(0.52f, 0.60f, 0.72f), 0.36f + 0.26f * rs_clamp(1.0f - light, 0.0f, 1.0f)),
                            rs_clamp(1.0f - s.soot * 0.32f, 0.50f, 1.0f));
            s3 ground = sscale(sscale(S3(0.58f, 0.54f, 0.48f), 0.070f), rs_clamp(1.0f - p.y / by, 0.0f, 1.0f));
            float powder = rs_clamp(1.0f - pinned_exp(-sigma_t * step * 2.2f), 0.0f, 1.0f);
            float pw = powder * 0.055f * sqrtf(light);
            s3 multiple = smul(sscale(color, sigma_s), sadd(sadd(sky, ground), S3(pw, pw, pw)));
            s3 direct = sscale(sscale(smul(sscale(color, sigma_s), sun_radiance), phase), light);
            float freshness = rs_clamp(1.0f - s.age / 17.0f, 0.0f, 1.0f);
            float fresh_heat = s.temperature * freshness * freshness;
            s3 emission = sscale(S3(1.0f, 0.30f, 0.055f), rs_clamp((fresh_heat * 0.10f + s.emission * 1.18f) * st->fire_glow, 0.0f, 5.0f));
            s3 source = sadd(sadd(direct, multiple), emission);
            rgb = sadd(rgb, sscale(sscale(source, seg_w), transmittance));
            transmittance *= seg_t;
        }
        t += step;
        steps += 1;
    }
    float alpha = rs_clamp(1.0f - transmittance, 0.0f, 1.0f);
    s3 straight = alpha > 1.0e-5f ? S3(rgb.x / alpha, rgb.y / alpha, rgb.z / alpha) : rgb;
    s3 e = sscale(straight, st->exposure);
    out[0] = to_u8(e.x / (1.0f + e.x));                                                      /* tone_map, :382-384 */
    out[1] = to_u8(e.y / (1.0f + e.y));
    out[2] = to_u8(e.z / (1.0f + e.z));
    out[3] = to_u8(alpha);
}

static void steps_for(const f3do_smoke_volume* V, const f3do_smoke_settings* st, float* step, float* shadow_step) {   /* :43-59 */
    float min_step = fmaxf(fminf(fminf(fminf(INFINITY, V->voxel_size[0]), V->voxel_size[1]), V->voxel_size[2]), 1.0e-4f);
    *step = st->step_size > 0.0f ? st->step_size : min_step * 0.75f;
    *shadow_step = st->shadow_step_size > 0.0f ? st->shadow_step_size : *step * 2.0f;
}

int f3do_smoke_raymarch_rgba(const f3do_smoke_volume* V, const f3do_smoke_settings* st, uint32_t width, uint32_t height,
                             const float camera_pos[3], const float target[3], const float up_in[3], float fovy_deg,
                             const float sun_direction[3], uint8_t* rgba) {
    g_smoke_err[0] = 0;
    if (validate_settings(st)) return 1;
    if (width == 0 || height == 0) return sfail("width and height must be >= 1");
    if (!isfinite(fovy_deg) || fovy_deg <= 0.0f || fovy_deg >= 179.0f) return sfail("fovy_deg must be finite and in (0, 179)");
    s3 eye = S3(camera_pos[0], camera_pos[1], camera_pos[2]);
    s3 forward = snormalize_or_zero(ssub(S3(target[0], target[1], target[2]), eye));
    if (sdot(forward, forward) < 1.0e-12f) return sfail("camera_pos and target must not be equal");
    s3 up = snormalize_or_zero(S3(up_in[0], up_in[1], up_in[2]));
    if (sdot(up, up) < 1.0e-12f) return sfail("up vector must not be zero");
    s3 right = snormalize_or_zero(scross(forward, up));
    s3 camera_up = snormalize_or_zero(scross(right, forward));
    s3 sun_dir = snormalize_or_zero(S3(sun_direction[0], sun_direction[1], sun_direction[2]));
    if (sdot(sun_dir, sun_dir) < 1.0e-12f) return sfail("sun_direction must not be zero");
    float step, shadow_step;
    steps_for(V, st, &step, &shadow_step);
    float tan_half_fov = tanf(fovy_deg * (3.14159274101257324f / 180.0f) * 0.5f);
    float aspect = (float)width / (float)height;
    s3 bmin = S3(V->origin[0], V->origin[1], V->origin[2]), bmax = bounds_max(V);
    memset(rgba, 0, (size_t)width * height * 4);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t y = 0; y < (int64_t)height; y++)
        for (uint32_t x = 0; x < width; x++) {
            float px = (((float)x + 0.5f) / (float)width * 2.0f - 1.0f) * aspect * tan_half_fov;
            float py = (1.0f - ((float)y + 0.5f) / (float)height * 2.0f) * tan_half_fov;
            s3 dir = snormalize(sadd(sadd(forward, sscale(right, px)), sscale(camera_up, py)));
            float t0, t1;
            if (!ray_box(eye, dir, bmin, bmax, &t0, &t1)) continue;
            t0 = fmaxf(t0, 0.0f);
            uint32_t seed = x * 73856093u + (uint32_t)y * 19349663u + (uint32_t)V->frame_index;
            march_ray_rgba(V, eye, dir, t0, t1, seed, step, shadow_step, sun_dir, st, rgba + ((size_t)y * width + x) * 4);
        }
    return 0;
}

/* Straight-alpha "over", python/forge3d/map_scene.py:1588-1604 (_alpha_composite_rgba): numpy float32 arithmetic,
 * truncating cast, alpha = max.  dst is updated in place. */
static void composite_over(uint8_t* dst, const uint8_t* src) {
    float alpha = (float)src[3] / 255.0f, keep = 1.0f - alpha;
    for (int c = 0; c < 3; c++) {
        float v = (float)dst[c] * keep + (float)src[c] * alpha;
        v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
        dst[c] = (uint8_t)v;
    }
    dst[3] = dst[3] > src[3] ? dst[3] : src[3];
}

/* The compositor alone, n pixels: out = over(bottom, top).  Pinned by tests/golden/alpha_composite_vectors.npz, which the
 * reference's own function produced (tools/make_composite_golden.py). */
void f3do_composite_over_rgba(const uint8_t* bottom, const uint8_t* top, uint64_t n, uint8_t* out) {
    memcpy(out, bottom, (size_t)n * 4);
    for (uint64_t i = 0; i < n; i++) composite_over(out + i * 4, top + i * 4);
}

/* The smoke layer over a terrain frame (BASELINE config 4): SmokeVolume::raymarch_rgba (render.rs:6-101) per pixel, the march
 * optionally ended at the terrain depth (NaN / <= 0 = sky; NULL = the reference's bare layer), then _alpha_composite_rgba. */
int f3do_smoke_raymarch_over_rgba(const f3do_smoke_volume* V, const f3do_smoke_settings* st, uint32_t width, uint32_t height,
                                  const float camera_pos[3], const float target[3], const float up_in[3], float fovy_deg,
                                  const float sun_direction[3], const uint8_t* base_rgba, const float* base_depth, uint8_t* rgba) {
    g_smoke_err[0] = 0;
    if (validate_settings(st)) return 1;
    if (width == 0 || height == 0) return sfail("width and height must be >= 1");
    if (!isfinite(fovy_deg) || fovy_deg <= 0.0f || fovy_deg >= 179.0f) return sfail("fovy_deg must be finite and in (0, 179)");
    s3 eye = S3(camera_pos[0], camera_pos[1], camera_pos[2]);
    s3 forward = snormalize_or_zero(ssub(S3(target[0], target[1], target[2]), eye));
    if (sdot(forward, forward) < 1.0e-12f) return sfail("camera_pos and target must not be equal");
    s3 up = snormalize_or_zero(S3(up_in[0], up_in[1], up_in[2]));
    if (sdot(up, up) < 1.0e-12f) return sfail("up vector must not be zero");
    s3 right = snormalize_or_zero(scross(forward, up));
    s3 camera_up = snormalize_or_zero(scross(right, forward));
    s3 sun_dir = snormalize_or_zero(S3(sun_direction[0], sun_direction[1], sun_direction[2]));
    if (sdot(sun_dir, sun_dir) < 1.0e-12f) return sfail("sun_direction must not be zero");
    float step, shadow_step;
    steps_for(V, st, &step, &shadow_step);
    float tan_half_fov = tanf(fovy_deg * (3.14159274101257324f / 180.0f) * 0.5f);
    float aspect = (float)width / (float)height;
    s3 bmin = S3(V->origin[0], V->origin[1], V->origin[2]), bmax = bounds_max(V);
    memcpy(rgba, base_rgba, (size_t)width * height * 4);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t y = 0; y < (int64_t)height; y++)
        for (uint32_t x = 0; x < width; x++) {
            float px = (((float)x + 0.5f) / (float)width * 2.0f - 1.0f) * aspect * tan_half_fov;
            float py = (1.0f - ((float)y + 0.5f) / (float)height * 2.0f) * tan_half_fov;
            s3 dir = snormalize(sadd(sadd(forward, sscale(right, px)), sscale(camera_up, py)));
            uint8_t layer[4] = {0, 0, 0, 0};
            float t0, t1;
            if (ray_box(eye, dir, bmin, bmax, &t0, &t1)) {
                if (base_depth) {
                    float d = base_depth[(size_t)y * width + x];
                    if (d > 0.0f) t1 = fminf(t1, d);
                }
                t0 = fmaxf(t0, 0.0f);
                uint32_t seed = x * 73856093u + (uint32_t)y * 19349663u + (uint32_t)V->frame_index;
                march_ray_rgba(V, eye, dir, t0, t1, seed, step, shadow_step, sun_dir, st, layer);
            }
            composite_over(rgba + ((size_t)y * width + x) * 4, layer);
        }
    return 0;
}

int f3do_smoke_raymarch_projection_rgba(const f3do_smoke_volume* V, const f3do_smoke_settings* st, uint32_t width, uint32_t height,
                                        const float view_direction[3], const float sun_direction[3], uint8_t* rgba) {
    g_smoke_err[0] = 0;
    if (validate_settings(st)) return 1;
    if (width == 0 || height == 0) return sfail("width and height must be >= 1");
    s3 dir = snormalize_or_zero(S3(view_direction[0], view_direction[1], view_direction[2]));
    if (sdot(dir, dir) < 1.0e-12f) return sfail("view_direction must not be zero");
    s3 sun_dir = snormalize_or_zero(S3(sun_direction[0], sun_direction[1], sun_direction[2]));
    if (sdot(sun_dir, sun_dir) < 1.0e-12f) return sfail("sun_direction must not be zero");
    float step, shadow_step;
    steps_for(V, st, &step, &shadow_step);
    s3 bmin = S3(V->origin[0], V->origin[1], V->origin[2]), bmax = bounds_max(V);
    float diagonal = fmaxf(slength(ssub(bmax, bmin)), step * 2.0f);
    memset(rgba, 0, (size_t)width * height * 4);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t py = 0; py < (int64_t)height; py++) {
        float fz = ((float)py + 0.5f) / (float)height;
        float z = lerpf(bmin.z, bmax.z, fz);
        for (uint32_t px = 0; px < width; px++) {
            float fx = ((float)px + 0.5f) / (float)width;
            float x = lerpf(bmin.x, bmax.x, fx);
            s3 plane = S3(x, (bmin.y + bmax.y) * 0.5f, z);
            s3 origin = ssub(plane, sscale(dir, diagonal));
            float t0, t1;
            if (!ray_box(origin, dir, bmin, bmax, &t0, &t1)) continue;
            t0 = fmaxf(t0, 0.0f);
            uint32_t seed = px * 73856093u + (uint32_t)py * 19349663u + (uint32_t)V->frame_index + 0x9e3779b9u;
            march_ray_rgba(V, origin, dir, t0, t1, seed, step, shadow_step, sun_dir, st, rgba + ((size_t)py * width + px) * 4);
        }
    }
    return 0;
}
