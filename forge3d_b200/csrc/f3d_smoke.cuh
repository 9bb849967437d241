// forge3d_b200/csrc/f3d_smoke.cuh
// Smoke volume ray-march on the GPU (BASELINE config 4 "heightfield + volume ray-march").
// Replaces the reference's single-threaded CPU loops
//   /root/reference/src/smoke/render.rs:6-101    SmokeVolume::raymarch_rgba            (perspective)
//   /root/reference/src/smoke/render.rs:103-175  SmokeVolume::raymarch_projection_rgba (map-aligned parallel projection)
//   /root/reference/src/smoke/render.rs:177-316  sample_render_fields, march_ray_rgba, sun_transmittance
//   /root/reference/src/smoke/render.rs:328-401  smoke_color, ray_box_intersection, henyey_greenstein, tone_map, to_u8
//   /root/reference/src/smoke/sampling.rs:1-32,83-103  trilinear sample_scalar, hash01
// with one thread per pixel (a warp owns an 8x4 pixel tile, so its rays walk neighbouring voxels).
//
// Data layout (B200-native, not the reference's six separate Vec<f32>): every march step samples density, soot and
// age, and - only where there is smoke - temperature, humidity and emission AT THE SAME POINT.  The fields are packed
// per voxel as one float4 A = (density, soot, age, temperature) and one float2 B = (humidity, emission): a trilinear
// tap is 8 x 128-bit loads (the x0/x1 pair of a row shares a 32-byte sector) instead of 8 x 6 scalar loads from six
// arrays, the self-shadow march (20 taps per lit sample, 95 % of all taps) touches A only, and B is read only inside
// the `density > 1e-5` branch.  Each component goes through the reference's own lerp chain, so values are identical.
//
// Empty-space skipping, exact: a march step whose eight trilinear corners all hold density 0 samples density exactly 0
// (lerp(0, 0, t) = 0 + (0 - 0) * t), so the main march's `density > 1e-5` test fails and the shadow march adds 0 to the optical
// depth - the step contributes nothing and only `t += step` remains.  k_smoke_pack marks, per 4x4x4 brick of sample cells, whether
// any corner of any cell in it is non-zero; a step in an unmarked brick skips its 8 (or 16) 128-bit loads and ~100 flops.  The
// skip is disabled when a field holds values beyond 1e15 (0 * inf would be NaN in the reference, not 0).
//
// Numerics: Rust f32 semantics op for op under the contract of DESIGN.md section 4 (no FMA contraction, IEEE
// division / sqrt).  f32::clamp propagates NaN (rs_clamp), f32::min/max ignore it (fminf/fmaxf), `as u8` saturates
// with NaN -> 0.  Pinned libm calls: exp(x) = exp2_pinned(x * log2 e), powf(d, 1.5) = d * sqrt(d).
#pragma once
#include "f3d_aether.cuh"   // exp2_pinned
#include "f3d_math.cuh"

namespace f3d {

struct SmokeSettings {      // SmokeRenderSettings, src/smoke/types.rs:225-266
    float density_scale, extinction, scattering, absorption, phase_g;
    uint32_t max_steps, self_shadow, shadow_steps;
    float jitter_strength, exposure, thin_color[3], dense_color[3], soot_absorption, fire_glow;
};

struct SmokeParams {
    const float4* volA;     // (density, soot, particle_age, temperature) per voxel, (z * ny + y) * nx + x
    const float2* volB;     // (humidity, emission_rate)
    uint32_t dims[3];
    float voxel[3], origin[3], bmax[3];
    SmokeSettings s;
    float step, shadow_step;
    uint32_t W, H, frame_index;
    uint32_t projection;    // 0: perspective camera, 1: parallel projection
    float eye[3], forward[3], right[3], up[3], tan_half_fov, aspect;   // perspective
    float dir[3], diagonal;                                            // projection
    float sun_dir[3];
    uint8_t* rgba;          // H x W x 4
    // smoke over a terrain frame (BASELINE config 4): both NULL = the bare layer, exactly the reference's raymarch_rgba
    const uchar4* base;     // H x W RGBA8 the layer is composited over (straight-alpha "over", map_scene.py:1588-1604), or NULL
    const float* depth;     // H x W distance along the camera ray to the opaque surface (NaN / <= 0 = none): clips the march, or NULL
    const uint8_t* occ;     // per 4x4x4 brick of sample cells: 1 = some corner density is non-zero; NULL = no skipping
    uint32_t occ_dims[2];   // bricks along x, y
};

__device__ __forceinline__ float rs_clamp(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }   // f32::clamp
__device__ __forceinline__ float lerp_rs(float a, float b, float t) { return a + (b - a) * t; }                      // sampling.rs:87
__device__ __forceinline__ v3 mix_rs(v3 a, v3 b, float t) { return V3(lerp_rs(a.x, b.x, t), lerp_rs(a.y, b.y, t), lerp_rs(a.z, b.z, t)); }
__device__ __forceinline__ float exp_pinned(float x) { return exp2_pinned(x * 1.4426950408889634f); }
__device__ __forceinline__ float smoothstep_rs(float e0, float e1, float x) {                                        // render.rs:386
    const float t = rs_clamp(fdiv(x - e0, fmaxf(e1 - e0, 1.0e-6f)), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
__device__ __forceinline__ uint8_t smoke_to_u8(float v) {                                                            // render.rs:399
    const float c = rs_clamp(v, 0.0f, 1.0f) * 255.0f + 0.5f;
    if (!(c == c)) return 0;
    return c >= 255.0f ? (uint8_t)255 : (c <= 0.0f ? (uint8_t)0 : (uint8_t)c);
}

// Trilinear tap geometry of sample_scalar (sampling.rs:1-17): corner indices + fractions for a grid-space point.
struct SmokeTap { uint32_t i000, i100, i010, i110, i001, i101, i011, i111; float fx, fy, fz; bool empty; };

__device__ __forceinline__ SmokeTap smoke_tap(const SmokeParams& P, v3 pos) {
    // grid_coord_from_world, types.rs:387-393
    const float gx = fdiv(pos.x - P.origin[0], P.voxel[0]) - 0.5f, gy = fdiv(pos.y - P.origin[1], P.voxel[1]) - 0.5f,
                gz = fdiv(pos.z - P.origin[2], P.voxel[2]) - 0.5f;
    const uint32_t nx = P.dims[0], ny = P.dims[1], nz = P.dims[2];
    const float x = rs_clamp(gx, 0.0f, (float)(nx - 1u)), y = rs_clamp(gy, 0.0f, (float)(ny - 1u)), z = rs_clamp(gz, 0.0f, (float)(nz - 1u));
    const float flx = floorf(x), fly = floorf(y), flz = floorf(z);
    const uint32_t x0 = (x == x) ? (uint32_t)flx : 0u, y0 = (y == y) ? (uint32_t)fly : 0u, z0 = (z == z) ? (uint32_t)flz : 0u;   // NaN as usize = 0
    const uint32_t x1 = min(x0 + 1u, nx - 1u), y1 = min(y0 + 1u, ny - 1u), z1 = min(z0 + 1u, nz - 1u);
    SmokeTap t;
    t.empty = P.occ != nullptr && __ldg(P.occ + ((size_t)(z0 >> 2) * P.occ_dims[1] + (y0 >> 2)) * P.occ_dims[0] + (x0 >> 2)) == 0u;
    t.fx = x - (float)x0; t.fy = y - (float)y0; t.fz = z - (float)z0;
    const uint32_t r00 = (z0 * ny + y0) * nx, r10 = (z0 * ny + y1) * nx, r01 = (z1 * ny + y0) * nx, r11 = (z1 * ny + y1) * nx;
    t.i000 = r00 + x0; t.i100 = r00 + x1; t.i010 = r10 + x0; t.i110 = r10 + x1;
    t.i001 = r01 + x0; t.i101 = r01 + x1; t.i011 = r11 + x0; t.i111 = r11 + x1;
    return t;
}

#define F3D_SMOKE_TRILERP(c000, c100, c010, c110, c001, c101, c011, c111, T)                                      \
    lerp_rs(lerp_rs(lerp_rs(c000, c100, (T).fx), lerp_rs(c010, c110, (T).fx), (T).fy),                               \
            lerp_rs(lerp_rs(c001, c101, (T).fx), lerp_rs(c011, c111, (T).fx), (T).fy), (T).fz)

__device__ __forceinline__ float4 smoke_sample_a(const SmokeParams& P, const SmokeTap& t) {
    const float4 a = __ldg(P.volA + t.i000), b = __ldg(P.volA + t.i100), c = __ldg(P.volA + t.i010), d = __ldg(P.volA + t.i110);
    const float4 e = __ldg(P.volA + t.i001), f = __ldg(P.volA + t.i101), g = __ldg(P.volA + t.i011), h = __ldg(P.volA + t.i111);
    return make_float4(F3D_SMOKE_TRILERP(a.x, b.x, c.x, d.x, e.x, f.x, g.x, h.x, t), F3D_SMOKE_TRILERP(a.y, b.y, c.y, d.y, e.y, f.y, g.y, h.y, t),
                       F3D_SMOKE_TRILERP(a.z, b.z, c.z, d.z, e.z, f.z, g.z, h.z, t), F3D_SMOKE_TRILERP(a.w, b.w, c.w, d.w, e.w, f.w, g.w, h.w, t));
}
__device__ __forceinline__ float2 smoke_sample_b(const SmokeParams& P, const SmokeTap& t) {
    const float2 a = __ldg(P.volB + t.i000), b = __ldg(P.volB + t.i100), c = __ldg(P.volB + t.i010), d = __ldg(P.volB + t.i110);
    const float2 e = __ldg(P.volB + t.i001), f = __ldg(P.volB + t.i101), g = __ldg(P.volB + t.i011), h = __ldg(P.volB + t.i111);
    return make_float2(F3D_SMOKE_TRILERP(a.x, b.x, c.x, d.x, e.x, f.x, g.x, h.x, t), F3D_SMOKE_TRILERP(a.y, b.y, c.y, d.y, e.y, f.y, g.y, h.y, t));
}

// ray_box_intersection, render.rs:348-378
__device__ __forceinline__ bool smoke_ray_box(v3 o, v3 d, v3 bmin, v3 bmax, float& near_t, float& far_t) {
    const float inf = __int_as_float(0x7f800000);
    const v3 inv = V3(fabsf(d.x) > 1.0e-12f ? fdiv(1.0f, d.x) : inf, fabsf(d.y) > 1.0e-12f ? fdiv(1.0f, d.y) : inf,
                      fabsf(d.z) > 1.0e-12f ? fdiv(1.0f, d.z) : inf);
    const v3 t0 = (bmin - o) * inv, t1 = (bmax - o) * inv;
    near_t = fmaxf(fmaxf(fminf(t0.x, t1.x), fminf(t0.y, t1.y)), fminf(t0.z, t1.z));
    far_t = fminf(fminf(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y)), fmaxf(t0.z, t1.z));
    return far_t >= fmaxf(near_t, 0.0f);
}

// sun_transmittance, render.rs:278-316
__device__ __forceinline__ float smoke_sun_transmittance(const SmokeParams& P, v3 start, v3 sun_dir) {
    const SmokeSettings& S = P.s;
    const float step = P.shadow_step;
    float t0, t1;
    if (!smoke_ray_box(start + sun_dir * step, sun_dir, V3(P.origin[0], P.origin[1], P.origin[2]), V3(P.bmax[0], P.bmax[1], P.bmax[2]), t0, t1))
        return 1.0f;
    t0 = fmaxf(t0, 0.0f);
    float optical_depth = 0.0f;
    for (uint32_t i = 0; i < S.shadow_steps; i++) {
        const float t = t0 + ((float)i + 0.5f) * step;
        if (t > t1) break;
        const v3 p = start + sun_dir * (step + t);
        const SmokeTap tap = smoke_tap(P, p);
        if (tap.empty) continue;                            // density samples exactly 0: optical_depth += 0
        const float4 a = smoke_sample_a(P, tap);
        const float age = fmaxf(a.z, 0.0f);
        const float age_t = smoothstep_rs(1.6f, 17.0f, age);
        const float gate = 0.50f + 0.50f * smoothstep_rs(0.045f, 0.34f, a.x);
        optical_depth += a.x * S.density_scale * (1.0f - 0.58f * age_t) * gate * S.extinction * (1.0f + a.y * S.soot_absorption) * step;
        if (optical_depth > 8.0f) break;
    }
    return rs_clamp(exp_pinned(-optical_depth), 0.0f, 1.0f);
}

// Straight-alpha "over" of `top` on `bottom`, the reference's _alpha_composite_rgba (python/forge3d/map_scene.py:1588-1604):
// alpha = a / 255 in f32, rgb = u8(clip(dst * (1 - alpha) + src * alpha, 0, 255)) (truncation), a = max(dst.a, src.a).
__device__ __forceinline__ uchar4 composite_over(uchar4 bottom, uchar4 top) {
    const float alpha = fdiv((float)top.w, 255.0f), keep = 1.0f - alpha;
    const float r = (float)bottom.x * keep + (float)top.x * alpha;
    const float g = (float)bottom.y * keep + (float)top.y * alpha;
    const float b = (float)bottom.z * keep + (float)top.z * alpha;
    return make_uchar4((unsigned char)rs_clamp(r, 0.0f, 255.0f), (unsigned char)rs_clamp(g, 0.0f, 255.0f),
                       (unsigned char)rs_clamp(b, 0.0f, 255.0f), bottom.w > top.w ? bottom.w : top.w);
}

// march_ray_rgba, render.rs:190-276
__device__ __forceinline__ uchar4 smoke_march(const SmokeParams& P, v3 origin, v3 dir, float t0, float t1, uint32_t seed, v3 sun_dir) {
    const SmokeSettings& S = P.s;
    const float step = P.step;
    uint32_t hv = seed;                                                               // hash01, sampling.rs:96-103
    hv ^= hv >> 16; hv *= 0x7FEB352Du; hv ^= hv >> 15; hv *= 0x846CA68Bu; hv ^= hv >> 16;
    const float jitter = (__uint2float_rn(hv) * 2.3283064365386963e-10f - 0.5f) * S.jitter_strength * step;
    float t = fmaxf(t0 + jitter, 0.0f);
    float transmittance = 1.0f;
    v3 rgb = V3(0.0f, 0.0f, 0.0f);
    uint32_t steps = 0u;
    const float by = fmaxf(P.bmax[1], 1.0f);
    const v3 thin = V3(S.thin_color[0], S.thin_color[1], S.thin_color[2]), dense = V3(S.dense_color[0], S.dense_color[1], S.dense_color[2]);
    while (t < t1 && steps < S.max_steps && transmittance > 0.01f) {
        const v3 p = origin + dir * t;
        const SmokeTap tap = smoke_tap(P, p);
        if (tap.empty) { t += step; steps += 1u; continue; }   // density samples exactly 0: nothing to add
        const float4 a = smoke_sample_a(P, tap);
        const float s_density = a.x, s_soot = a.y, s_age = fmaxf(a.z, 0.0f), s_temp = a.w;
        const float age_t = smoothstep_rs(1.6f, 17.0f, s_age);
        const float gate = 0.50f + 0.50f * smoothstep_rs(0.045f, 0.34f, s_density);
        const float density = fmaxf(s_density * S.density_scale * (1.0f - 0.58f * age_t) * gate, 0.0f);
        if (density > 1.0e-5f) {
            const float2 b = smoke_sample_b(P, tap);
            const float s_humidity = b.x, s_emission = b.y;
            const float sigma_t = density * S.extinction * (1.0f + s_soot * S.soot_absorption * 0.85f);
            const float seg_t = rs_clamp(exp_pinned(-sigma_t * step), 0.0f, 1.0f);
            const float seg_w = sigma_t > 1.0e-6f ? fdiv(1.0f - seg_t, sigma_t) : step;
            const float light = S.self_shadow ? smoke_sun_transmittance(P, p, sun_dir) : 1.0f;
            const float cos_theta = rs_clamp(dot3(dir, sun_dir), -1.0f, 1.0f);
            const float g2 = S.phase_g * S.phase_g;                                        // henyey_greenstein, :380-384
            const float hg_d = fmaxf(1.0f + g2 - 2.0f * S.phase_g * cos_theta, 1.0e-4f);
            const float phase = fdiv(1.0f - g2, 4.0f * 3.14159274101257324f * (hg_d * fsqrt(hg_d)));
            // smoke_color, :328-346
            const float body = rs_clamp(s_density * 1.45f + s_soot * 1.35f, 0.0f, 1.0f);
            v3 color = mix_rs(thin, dense, body);
            const float aged = rs_clamp(fdiv(s_age, 9.0f), 0.0f, 1.0f);
            color = mix_rs(color, V3(0.36f, 0.39f, 0.43f), aged * 0.42f);
            const float milk = rs_clamp(s_humidity, 0.0f, 1.0f) * (0.18f + 0.42f * body);
            color = mix_rs(color, V3(0.93f, 0.92f, 0.84f), rs_clamp(milk, 0.0f, 0.38f));
            const float freshness = rs_clamp(1.0f - fdiv(s_age, 17.0f), 0.0f, 1.0f);
            const float heat = rs_clamp(s_temp * 0.12f * freshness, 0.0f, 1.0f);
            color = mix_rs(color, V3(0.95f, 0.62f, 0.28f), heat * 0.07f);

            const float albedo = rs_clamp(fdiv(S.scattering, S.scattering + S.absorption + s_soot * 0.55f + 1.0e-5f), 0.02f, 0.98f);
            const float sigma_s = sigma_t * albedo;
            const v3 sun_radiance = V3(1.0f, 0.96f, 0.84f) * 11.5f;
            const v3 sky = (V3(0.52f, 0.60f, 0.72f) * (0.36f + 0.26f * rs_clamp(1.0f - light, 0.0f, 1.0f))) * rs_clamp(1.0f - s_soot * 0.32f, 0.50f, 1.0f);
            const v3 ground = (V3(0.58f, 0.54f, 0.48f) * 0.070f) * rs_clamp(1.0f - fdiv(p.y, by), 0.0f, 1.0f);
            const float powder = rs_clamp(1.0f - exp_pinned(-sigma_t * step * 2.2f), 0.0f, 1.0f);
            const float pw = powder * 0.055f * fsqrt(light);
            const v3 multiple = (color * sigma_s) * ((sky + ground) + V3(pw, pw, pw));
            const v3 direct = (((color * sigma_s) * sun_radiance) * phase) * light;
            const float fresh_heat = s_temp * freshness * freshness;
            const v3 emission = V3(1.0f, 0.30f, 0.055f) * rs_clamp((fresh_heat * 0.10f + s_emission * 1.18f) * S.fire_glow, 0.0f, 5.0f);
            const v3 source = (direct + multiple) + emission;
            rgb = rgb + (source * seg_w) * transmittance;
            transmittance *= seg_t;
        }
        t += step;
        steps += 1u;
    }
    const float alpha = rs_clamp(1.0f - transmittance, 0.0f, 1.0f);
    const v3 straight = alpha > 1.0e-5f ? V3(fdiv(rgb.x, alpha), fdiv(rgb.y, alpha), fdiv(rgb.z, alpha)) : rgb;
    const v3 e = straight * S.exposure;
    uchar4 px;
    px.x = smoke_to_u8(fdiv(e.x, 1.0f + e.x));
    px.y = smoke_to_u8(fdiv(e.y, 1.0f + e.y));
    px.z = smoke_to_u8(fdiv(e.z, 1.0f + e.z));
    px.w = smoke_to_u8(alpha);
    return px;
}

constexpr int kSmokeThreads = 128;   // 4 warps, each an 8 x 4 pixel tile: the CTA covers 16 x 8 pixels

__global__ void __launch_bounds__(kSmokeThreads) k_smoke_march(const SmokeParams P) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t x = blockIdx.x * 16u + (warp & 1u) * 8u + (lane & 7u);
    const uint32_t y = blockIdx.y * 8u + (warp >> 1) * 4u + (lane >> 3);
    if (x >= P.W || y >= P.H) return;
    const v3 bmin = V3(P.origin[0], P.origin[1], P.origin[2]), bmax = V3(P.bmax[0], P.bmax[1], P.bmax[2]);
    const v3 sun_dir = V3(P.sun_dir[0], P.sun_dir[1], P.sun_dir[2]);
    v3 origin, dir;
    uint32_t seed = x * 73856093u + y * 19349663u + P.frame_index;
    if (P.projection == 0u) {            // render.rs:66-79
        const float px = (fdiv((float)x + 0.5f, (float)P.W) * 2.0f - 1.0f) * P.aspect * P.tan_half_fov;
        const float py = (1.0f - fdiv((float)y + 0.5f, (float)P.H) * 2.0f) * P.tan_half_fov;
        origin = V3(P.eye[0], P.eye[1], P.eye[2]);
        const v3 d = (V3(P.forward[0], P.forward[1], P.forward[2]) + V3(P.right[0], P.right[1], P.right[2]) * px) + V3(P.up[0], P.up[1], P.up[2]) * py;
        dir = d * fdiv(1.0f, fsqrt(dot3(d, d)));                                             // glam Vec3::normalize
    } else {                             // render.rs:140-158
        const float fz = fdiv((float)y + 0.5f, (float)P.H), fx = fdiv((float)x + 0.5f, (float)P.W);
        const v3 plane = V3(lerp_rs(bmin.x, bmax.x, fx), (bmin.y + bmax.y) * 0.5f, lerp_rs(bmin.z, bmax.z, fz));
        dir = V3(P.dir[0], P.dir[1], P.dir[2]);
        origin = plane - dir * P.diagonal;
        seed += 0x9e3779b9u;
    }
    uchar4 out = make_uchar4(0, 0, 0, 0);
    float t0, t1;
    if (smoke_ray_box(origin, dir, bmin, bmax, t0, t1)) {
        if (P.depth != nullptr) {        // smoke behind the terrain is not seen: the march ends at the surface
            const float d = P.depth[(size_t)y * P.W + x];
            if (d > 0.0f) t1 = fminf(t1, d);
        }
        out = smoke_march(P, origin, dir, fmaxf(t0, 0.0f), t1, seed, sun_dir);
    }
    if (P.base != nullptr) out = composite_over(P.base[(size_t)y * P.W + x], out);
    reinterpret_cast<uchar4*>(P.rgba)[(size_t)y * P.W + x] = out;
}

// Packs the six host-layout fields into the A / B records (missing fields are zero).
// Also fills the brick occupancy (`occ` zeroed by the host) and raises *huge when a value rules the empty-space skip out.
__global__ void k_smoke_pack(const float* __restrict__ density, const float* __restrict__ temperature, const float* __restrict__ soot,
                             const float* __restrict__ humidity, const float* __restrict__ emission, const float* __restrict__ age,
                             size_t n, uint32_t nx, uint32_t ny, float4* __restrict__ volA, float2* __restrict__ volB,
                             uint8_t* __restrict__ occ, uint32_t bx, uint32_t by, uint32_t* __restrict__ huge) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = make_float4(density ? density[i] : 0.0f, soot ? soot[i] : 0.0f, age ? age[i] : 0.0f, temperature ? temperature[i] : 0.0f);
    const float2 b = make_float2(humidity ? humidity[i] : 0.0f, emission ? emission[i] : 0.0f);
    volA[i] = a;
    volB[i] = b;
    const float big = fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))), fmaxf(fabsf(b.x), fabsf(b.y)));
    if (!(big <= 1.0e15f)) *huge = 1u;                      // also catches NaN
    if (a.x != 0.0f) {
        // voxel (x, y, z) is a corner of the sample cells (x-1..x, y-1..y, z-1..z)
        const uint32_t x = (uint32_t)(i % nx), y = (uint32_t)((i / nx) % ny), z = (uint32_t)(i / ((size_t)nx * ny));
        for (uint32_t dz = 0; dz < 2u; dz++)
            for (uint32_t dy = 0; dy < 2u; dy++)
                for (uint32_t dx = 0; dx < 2u; dx++) {
                    const uint32_t cx = x >= dx ? x - dx : 0u, cy = y >= dy ? y - dy : 0u, cz = z >= dz ? z - dz : 0u;
                    occ[((size_t)(cz >> 2) * by + (cy >> 2)) * bx + (cx >> 2)] = 1u;
                }
    }
}

}  // namespace f3d
