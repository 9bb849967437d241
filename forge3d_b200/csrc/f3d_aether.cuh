// forge3d_b200/csrc/f3d_aether.cuh
// AETHER aerial-perspective post of the path-traced DEM snapshot: one kernel over
// the finished accumulation that replaces the reference's `hybrid-pt-aether-post` compute pass
//   /root/reference/src/shaders/atmosphere/prometheus_aerial.wgsl:51-231   (LUT loads, main)
//   /root/reference/src/shaders/atmosphere/evaluation_core.wgsl:9-344      (spectral basis, LUT coordinates,
//                                                                           quadrilinear lookup, segment integral)
//   /root/reference/src/path_tracing/hybrid_compute/aether_post.rs:40-340  (uniforms, LUT upload, dispatch)
// and the RGBA16F -> u8 read-back of render_terrain.rs:1358-1366, fused: L_out = L_surface*T + L_inscatter ->
// Reinhard -> f16 -> u8 in one pass, no intermediate RGBA16F image, no depth / visibility texture copies
// (aether_post.rs:296-331): the kernel reads the session's own depth AOV and hit-type bits.
//
// Numerics follow the contract of DESIGN.md section 4 (IEEE f32, no FMA contraction, pinned dot/normalize) plus
// the pins this pass adds: exp2 = the Cephes exp2f kernel below (WGSL leaves exp2 to the driver), round() =
// round-half-to-even, clamp = min(max()), fract(x) = x - floor(x).  The CPU oracle states the same definitions
// independently; results agree bit for bit.
//
// Data layout: the three LUTs stay RGBA16F exactly as the reference ships them (8 B per texel, one 64-bit load
// each, x fastest): transmittance 32x8, accumulated scattering 17x17x(8*16), aerial 8x8x8 = 302 KB, L2-resident.
#pragma once
#include <cuda_fp16.h>

#include "f3d_math.cuh"

namespace f3d {

struct AetherParams {
    const uint2* transmittance;       // [height][mu]
    const uint2* scattering;          // [height*nu_count + nu][mu_sun][mu_view]
    const uint2* aerial;              // [height][mu_view][distance]
    uint32_t t_dims[2];               // mu, height
    uint32_t s_dims[3];               // mu_view, mu_sun, height*nu
    uint32_t s_height, s_nu;
    uint32_t a_dims[3];               // distance, mu_view, height
    float bottom_radius_m, top_radius_m, max_aerial_distance_m, ozone_du, turbidity;
    float tan_half_fov, aspect;       // (0.5 * fov_y).tan(), W / H   (aether_post.rs:135-146)
    float sun_intensity;              // clamped scalar intensity (light_color carries intensity * colour)
};

// Pinned exp2: floor split, |f| <= 0.5, degree-5 polynomial, exact two-step scaling by 2^i.
// x >= 128 -> +inf, x < -126 -> 0 (no denormal results), NaN -> NaN.
__device__ __forceinline__ float exp2_pinned(float x) {
    if (x != x) return x;
    if (x >= 128.0f) return __int_as_float(0x7f800000);
    if (x < -126.0f) return 0.0f;
    const float px = floorf(x);
    int i0 = (int)px;
    float f = x - px;
    if (f > 0.5f) { i0 += 1; f = f - 1.0f; }
    float p = 1.535336188319500e-4f;
    p = p * f + 1.339887440266574e-3f;
    p = p * f + 9.618437357674640e-3f;
    p = p * f + 5.550332471162809e-2f;
    p = p * f + 2.402264791363012e-1f;
    p = p * f + 6.931472028550421e-1f;
    const float r = 1.0f + f * p;
    const int e1 = i0 >> 1, e2 = i0 - e1;
    return (r * __int_as_float((e1 + 127) << 23)) * __int_as_float((e2 + 127) << 23);
}
__device__ __forceinline__ float det_exp(float x) { return exp2_pinned(x * 1.4426950408889634f); }  // determinism.wgsl:325

__device__ __forceinline__ float4 aether_texel(const uint2* __restrict__ lut, uint32_t w, uint32_t h, int x, int y, int z) {
    const uint2 t = __ldg(lut + (((size_t)z * h + (size_t)y) * w + (size_t)x));
    return make_float4(__half2float(__ushort_as_half((unsigned short)(t.x & 0xFFFFu))),
                       __half2float(__ushort_as_half((unsigned short)(t.x >> 16))),
                       __half2float(__ushort_as_half((unsigned short)(t.y & 0xFFFFu))),
                       __half2float(__ushort_as_half((unsigned short)(t.y >> 16))));
}

__device__ __forceinline__ v3 clamp_hdr(v3 c) {                                          // evaluation_core.wgsl:35-37
    return V3(fminf(fmaxf(c.x, 0.0f), 65504.0f), fminf(fmaxf(c.y, 0.0f), 65504.0f), fminf(fmaxf(c.z, 0.0f), 65504.0f));
}
__device__ __forceinline__ float aether_mu_to_unit(float mu) {                           // :78-86
    const float b = clampf(mu, -1.0f, 1.0f);
    const float m = fsqrt(fabsf(b));
    return 0.5f * ((b >= 0.0f ? m : -m) + 1.0f);
}
__device__ __forceinline__ float aether_nu_to_unit(float nu) {                           // :88-90
    return 1.0f - fsqrt(fmaxf(0.5f * (1.0f - clampf(nu, -1.0f, 1.0f)), 0.0f));
}

// aether_eval_sample_accumulated_scattering, :115-170 (texel clamp of :98-113 folded in)
__device__ __forceinline__ v3 aether_scattering(const AetherParams& A, float height_unit, float mu_sun, float mu_view, float nu) {
    const int hc = max((int)A.s_height, 2), nc = max((int)A.s_nu, 2);
    const int dx = (int)A.s_dims[0], dy = (int)A.s_dims[1], dz = (int)A.s_dims[2];
    const float c[4] = {aether_mu_to_unit(mu_view) * (float)(A.s_dims[0] - 1u), aether_mu_to_unit(mu_sun) * (float)(A.s_dims[1] - 1u),
                        fsqrt(clampf(height_unit, 0.0f, 1.0f)) * (float)(hc - 1), aether_nu_to_unit(nu) * (float)(nc - 1)};
    const int lim[4] = {dx - 1, dy - 1, hc - 1, nc - 1};
    int lo[4], hi[4];
    float fr[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float fl = floorf(c[k]);
        lo[k] = (int)fl;
        hi[k] = min(lo[k] + 1, lim[k]);
        fr[k] = c[k] - fl;
    }
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
#pragma unroll
    for (int hs = 0; hs < 2; hs++)
#pragma unroll
        for (int ns = 0; ns < 2; ns++)
#pragma unroll
            for (int ss = 0; ss < 2; ss++)
#pragma unroll
                for (int vs = 0; vs < 2; vs++) {
                    const float wgt = (vs ? fr[0] : 1.0f - fr[0]) * (ss ? fr[1] : 1.0f - fr[1]) * (hs ? fr[2] : 1.0f - fr[2]) *
                                      (ns ? fr[3] : 1.0f - fr[3]);
                    const int x = min(max(vs ? hi[0] : lo[0], 0), dx - 1), y = min(max(ss ? hi[1] : lo[1], 0), dy - 1);
                    const int z = min(max((hs ? hi[2] : lo[2]) * nc + (ns ? hi[3] : lo[3]), 0), dz - 1);
                    const float4 t = aether_texel(A.scattering, A.s_dims[0], A.s_dims[1], x, y, z);
                    ax = ax + wgt * t.x;
                    ay = ay + wgt * t.y;
                    az = az + wgt * t.z;
                }
    return V3(fmaxf(ax, 0.0f), fmaxf(ay, 0.0f), fmaxf(az, 0.0f));
}

__device__ __forceinline__ float aether_radius(float camera_h, float view_mu, float dist, float bottom) {   // :172-186
    const float r = fmaxf(bottom, 1.0f) + clampf(camera_h, 0.0f, 100000.0f);
    const float bd = clampf(dist, 0.0f, 20000000.0f);
    return fsqrt(fmaxf(r * r + bd * bd + 2.0f * r * bd * clampf(view_mu, -1.0f, 1.0f), 0.0f));
}
__device__ __forceinline__ float aether_altitude(float camera_h, float view_mu, float dist, float bottom) { // :188-198
    return clampf(aether_radius(camera_h, view_mu, dist, bottom) - fmaxf(bottom, 1.0f), 0.0f, 100000.0f);
}

// aether_eval_spectral_xyz, :47-76.  The per-wavelength constants are evaluated at compile time by the same
// f32 operations the shader performs at run time (correctly rounded, so identical).
__device__ __forceinline__ v3 aether_spectral_xyz(int i, float rayleigh_col, float mie_col, float ozone_col, float turbidity) {
    const float lambda[11] = {380.0f, 420.0f, 460.0f, 500.0f, 540.0f, 580.0f, 620.0f, 660.0f, 700.0f, 740.0f, 780.0f};
    const float cie[11][3] = {
        {0.001368f, 0.000039f, 0.006450f}, {0.134380f, 0.004000f, 0.645600f}, {0.290800f, 0.060000f, 1.669200f},
        {0.004900f, 0.323000f, 0.272000f}, {0.290400f, 0.954000f, 0.020300f}, {0.916300f, 0.870000f, 0.001650f},
        {0.854450f, 0.381000f, 0.000190f}, {0.164900f, 0.061000f, 0.000000f}, {0.011359f, 0.004102f, 0.000000f},
        {0.000690f, 0.000249f, 0.000000f}, {0.000042f, 0.000015f, 0.000000f}};
    const float ratio = fdiv(550.0f, lambda[i]);
    const float ratio2 = ratio * ratio;
    const float rayleigh_beta = 1.2989e-5f * ratio2 * ratio2;
    const float mie_beta = 1.0e-5f * turbidity * ratio;
    const float delta = fdiv(lambda[i] - 600.0f, 85.0f);
    const float ozone_beta = 1.2e-6f * det_exp(-0.5f * delta * delta);
    const float tau = rayleigh_beta * rayleigh_col + mie_beta * mie_col + ozone_beta * ozone_col;
    const float t = det_exp(-fmaxf(tau, 0.0f));
    const float w = (i == 0 || i == 10) ? 0.5f : 1.0f;
    return V3(cie[i][0] * t * w, cie[i][1] * t * w, cie[i][2] * t * w);
}

// aether_eval_segment_transmittance, :225-344 (density_scale = 1)
__device__ __forceinline__ v3 aether_segment_transmittance(const AetherParams& A, float dist, float camera_h, float view_mu) {
    const float bd = clampf(dist, 0.0f, 20000000.0f);
    const float bh = clampf(camera_h, 0.0f, 100000.0f);
    float rayleigh = 0.0f, mie = 0.0f, ozone = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const float h = aether_altitude(bh, view_mu, bd * ((float)(2 * k + 1) * 0.03125f), A.bottom_radius_m);
        const float r = det_exp(fdiv(-h, 8000.0f)), m = det_exp(fdiv(-h, 1200.0f));
        const float o = fmaxf(1.0f - fabsf(fdiv(h - 25000.0f, 15000.0f)), 0.0f);
        rayleigh = k == 0 ? r : rayleigh + r;
        mie = k == 0 ? m : mie + m;
        ozone = k == 0 ? o : ozone + o;
    }
    const float pps = bd * 1.0f * 0.0625f;
    const float rayleigh_col = pps * rayleigh, mie_col = pps * mie;
    const float ozone_col = fdiv(pps * ozone * A.ozone_du, 300.0f);
    v3 xyz = aether_spectral_xyz(0, rayleigh_col, mie_col, ozone_col, A.turbidity);
#pragma unroll
    for (int i = 1; i < 11; i++) xyz = xyz + aether_spectral_xyz(i, rayleigh_col, mie_col, ozone_col, A.turbidity);
    const v3 rgb = V3(fdiv(dot3(V3(3.2404542f, -1.5371385f, -0.4985314f), xyz), 3.2613921f),
                      fdiv(dot3(V3(-0.9692660f, 1.8760108f, 0.0415560f), xyz), 2.5069624f),
                      fdiv(dot3(V3(0.0556434f, -0.2040259f, 1.0572252f), xyz), 2.3679786f));   // :39-45
    return V3(clampf(rgb.x, 0.0f, 1.0f), clampf(rgb.y, 0.0f, 1.0f), clampf(rgb.z, 0.0f, 1.0f));
}

struct AetherView {
    v3 cam_origin, cam_right, cam_up, cam_forward, sun_dir;   // sun_dir = normalize(light_dir)
    float exposure;                                            // clamped camera exposure
    uint32_t W, H;
};

// prometheus_aerial.wgsl `main` (:98-231) for one pixel; returns the LDR colour before the RGBA16F store.
__device__ __forceinline__ v3 aether_pixel(const AetherParams& A, const AetherView& V, uint32_t gx, uint32_t gy, float4 acc,
                                           float depth, bool visible) {
    const float denom = fmaxf(acc.w, 1.0f);
    const v3 surface = clamp_hdr(V3(fdiv(acc.x, denom), fdiv(acc.y, denom), fdiv(acc.z, denom)));
    const float ndc_x = fdiv((float)gx + 0.5f, (float)V.W) * 2.0f - 1.0f;
    const float ndc_y = (1.0f - fdiv((float)gy + 0.5f, (float)V.H)) * 2.0f - 1.0f;
    const float sx = ndc_x * A.tan_half_fov * A.aspect, sy = ndc_y * A.tan_half_fov;
    const v3 ray = normalize3((V.cam_right * sx + V.cam_up * sy) + V.cam_forward);
    const float sun_i = fminf(fmaxf(A.sun_intensity, 0.0f), 65504.0f);
    const float exposure = fminf(fmaxf(V.exposure, 0.0f), 65504.0f);
    const float atmosphere_height = fmaxf(A.top_radius_m - A.bottom_radius_m, 1.0f);
    const float camera_h = fmaxf(V.cam_origin.y, 0.0f);
    const float camera_hu = clampf(fdiv(camera_h, atmosphere_height), 0.0f, 1.0f);
    const float nu = dot3(ray, V.sun_dir);
    const v3 cs = aether_scattering(A, camera_hu, V.sun_dir.y, ray.y, nu) * sun_i;
    v3 hdr;
    if (!visible) hdr = clamp_hdr(cs);                                                              // :148-169
    else {
        const float end_h = aether_altitude(camera_h, ray.y, depth, A.bottom_radius_m);
        // aether_eval_spherical_endpoint_mus, :200-223
        const float r0 = fmaxf(A.bottom_radius_m, 1.0f) + clampf(camera_h, 0.0f, 100000.0f);
        const float bd = clampf(depth, 0.0f, 20000000.0f);
        const float er = fmaxf(aether_radius(camera_h, ray.y, bd, A.bottom_radius_m), 1.0f);
        const float end_view_mu = clampf(fdiv(r0 * clampf(ray.y, -1.0f, 1.0f) + bd, er), -1.0f, 1.0f);
        const float end_sun_mu = clampf(fdiv(r0 * clampf(V.sun_dir.y, -1.0f, 1.0f) + bd * clampf(nu, -1.0f, 1.0f), er), -1.0f, 1.0f);
        const v3 seg = aether_segment_transmittance(A, depth, camera_h, ray.y);
        // prometheus_load_boundary_transmittance, :51-61
        const int tx = (int)rintf((0.5f * (clampf(ray.y, -1.0f, 1.0f) + 1.0f)) * (float)(max(A.t_dims[0], 1u) - 1u));
        const int ty = (int)rintf(clampf(camera_hu, 0.0f, 1.0f) * (float)(max(A.t_dims[1], 1u) - 1u));
        const float4 bt4 = aether_texel(A.transmittance, A.t_dims[0], A.t_dims[1], tx, ty, 0);
        const v3 boundary_t = V3(clampf(bt4.x, 0.0f, 1.0f), clampf(bt4.y, 0.0f, 1.0f), clampf(bt4.z, 0.0f, 1.0f));
        const float end_hu = clampf(fdiv(end_h, atmosphere_height), 0.0f, 1.0f);
        const v3 es = aether_scattering(A, end_hu, end_sun_mu, end_view_mu, nu) * sun_i;
        // prometheus_load_aerial_transmittance, :84-96
        const float distance_unit = fdiv(depth, fmaxf(A.max_aerial_distance_m, 1.0f));
        const int ax = (int)rintf(clampf(distance_unit, 0.0f, 1.0f) * (float)(max(A.a_dims[0], 1u) - 1u));
        const int ay = (int)rintf(0.5f * (clampf(ray.y, -1.0f, 1.0f) + 1.0f) * (float)(max(A.a_dims[1], 1u) - 1u));
        const int az = (int)rintf(clampf(camera_hu, 0.0f, 1.0f) * (float)(max(A.a_dims[2], 1u) - 1u));
        const float aerial_mean = clampf(aether_texel(A.aerial, A.a_dims[0], A.a_dims[1], ax, ay, az).w, 0.0f, 1.0f);
        const float analytic_mean = dot3(seg, V3(0.2126f, 0.7152f, 0.0722f));
        const float k = fdiv(aerial_mean, fmaxf(analytic_mean, 1.0e-6f));
        const v3 tr = V3(fmaxf(clampf(seg.x * k, 0.0f, 1.0f), boundary_t.x), fmaxf(clampf(seg.y * k, 0.0f, 1.0f), boundary_t.y),
                         fmaxf(clampf(seg.z * k, 0.0f, 1.0f), boundary_t.z));
        const v3 fin = V3(fmaxf(cs.x - tr.x * es.x, 0.0f), fmaxf(cs.y - tr.y * es.y, 0.0f), fmaxf(cs.z - tr.z * es.z, 0.0f));
        hdr = clamp_hdr(surface * tr + fin);
    }
    const v3 e = hdr * exposure;
    return V3(fdiv(e.x, 1.0f + e.x), fdiv(e.y, 1.0f + e.y), fdiv(e.z, 1.0f + e.z));   // tonemap_reinhard
}

}  // namespace f3d
