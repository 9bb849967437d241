// forge3d_b200/csrc/f3d_smoke.cu
// Host side of the smoke volume ray-march entry points (f3d_smoke_*, include/forge3d_b200.h): volume upload + packing,
// SmokeRenderSettings validation and camera set-up in Rust f32 semantics, one kernel launch per render.
// Reference: /root/reference/src/smoke/render.rs:6-175, src/smoke/types.rs:29-55,268-317, src/smoke/py.rs:531-628.
#include "f3d_host.h"
#include "f3d_smoke.cuh"

using namespace f3d;
#define g_err g_f3d_err

// ------------------------------------------------------------------------------------------------
// smoke volume ray-march (src/smoke/render.rs:6-175): host set-up in Rust f32 semantics, one kernel per render
// ------------------------------------------------------------------------------------------------
struct f3d_smoke {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float4* volA = nullptr;
    float2* volB = nullptr;
    uint8_t* occ = nullptr;           // brick occupancy for the exact empty-space skip (NULL = disabled)
    uint32_t occ_dims[3] = {0, 0, 0};
    uint8_t* d_rgba = nullptr;
    size_t rgba_capacity = 0;
    uint8_t* d_base = nullptr;        // smoke over terrain: staged base RGBA and depth (f3d_smoke_raymarch_over_rgba)
    float* d_depth = nullptr;
    size_t base_capacity = 0;
    const uint8_t* over_base = nullptr;   // host pointers of the call in flight
    const float* over_depth = nullptr;
    uint32_t dims[3] = {0, 0, 0};
    float voxel[3] = {0, 0, 0}, origin[3] = {0, 0, 0};
    uint32_t frame_index = 0;
};

extern "C" void f3d_smoke_destroy(f3d_smoke* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    cached_free(s->volA, s->device); cached_free(s->volB, s->device); cached_free(s->d_rgba, s->device);
    cached_free(s->occ, s->device); cached_free(s->d_base, s->device); cached_free(s->d_depth, s->device);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

static int smoke_create_impl(const f3d_smoke_volume* v, int32_t device, f3d_smoke* s) {
    // SmokeDomainConfig::validate, src/smoke/types.rs:29-55 (the reference's CPU voxel cap is replaced by the 2^31 index limit)
    for (int a = 0; a < 3; a++)
        if (v->dims[a] < 2u) return fail(F3D_ERR_RENDER, "dims[%d] must be >= 2", a);
    const uint64_t n = (uint64_t)v->dims[0] * v->dims[1] * v->dims[2];
    if (n > (1ull << 31)) return fail(F3D_ERR_RENDER, "smoke domain has %llu voxels, exceeding the 2^31-voxel addressing limit", (unsigned long long)n);
    for (int a = 0; a < 3; a++)
        if (!isfinite(v->voxel_size[a]) || v->voxel_size[a] <= 0.0f) return fail(F3D_ERR_RENDER, "voxel_size[%d] must be finite and > 0", a);
    for (int a = 0; a < 3; a++)
        if (!isfinite(v->origin[a])) return fail(F3D_ERR_RENDER, "origin[%d] must be finite", a);
    if (!v->density) return fail(F3D_ERR_ARGUMENT, "density pointer is null");
    int rc = select_device(device);
    if (rc) return rc;
    s->device = device;
    CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&s->ev0));
    CUDA_TRY(cudaEventCreate(&s->ev1));
    memcpy(s->dims, v->dims, sizeof s->dims);
    memcpy(s->voxel, v->voxel_size, sizeof s->voxel);
    memcpy(s->origin, v->origin, sizeof s->origin);
    s->frame_index = (uint32_t)v->frame_index;                 // `self.frame_index as u32`, render.rs:78
    CUDA_TRY(cached_malloc((void**)&s->volA, n * sizeof(float4), device));
    CUDA_TRY(cached_malloc((void**)&s->volB, n * sizeof(float2), device));
    const float* host[6] = {v->density, v->temperature, v->soot, v->humidity, v->emission_rate, v->particle_age};
    float* dev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    struct Scratch {
        float** d; cudaStream_t st; int dv;
        ~Scratch() { cudaStreamSynchronize(st); for (int k = 0; k < 6; k++) cached_free(d[k], dv); }
    } scratch{dev, s->stream, device};
    for (int k = 0; k < 6; k++) {
        if (!host[k]) continue;
        CUDA_TRY(cached_malloc((void**)&dev[k], n * sizeof(float), device));
        CUDA_TRY(cudaMemcpyAsync(dev[k], host[k], n * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    }
    for (int a = 0; a < 3; a++) s->occ_dims[a] = (v->dims[a] + 3u) / 4u;
    const size_t nbricks = (size_t)s->occ_dims[0] * s->occ_dims[1] * s->occ_dims[2];
    uint32_t* d_huge = nullptr;
    CUDA_TRY(cached_malloc((void**)&s->occ, nbricks, device));
    CUDA_TRY(cached_malloc((void**)&d_huge, sizeof(uint32_t), device));
    CUDA_TRY(cudaMemsetAsync(s->occ, 0, nbricks, s->stream));
    CUDA_TRY(cudaMemsetAsync(d_huge, 0, sizeof(uint32_t), s->stream));
    k_smoke_pack<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(dev[0], dev[1], dev[2], dev[3], dev[4], dev[5], (size_t)n, v->dims[0],
                                                                     v->dims[1], s->volA, s->volB, s->occ, s->occ_dims[0], s->occ_dims[1], d_huge);
    uint32_t huge = 0;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&huge, d_huge, sizeof huge, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cached_free(d_huge, device);
    CUDA_TRY(e);
    if (huge || getenv("F3D_B200_SMOKE_NO_SKIP")) {         // 0 * inf would be NaN in the reference: march every step literally
        cached_free(s->occ, device);
        s->occ = nullptr;
    }
    return 0;
}

extern "C" int f3d_smoke_create(const f3d_smoke_volume* volume, int32_t device, f3d_smoke** out) {
    g_err[0] = 0;
    if (!volume || !out) return fail(F3D_ERR_ARGUMENT, "null argument");
    *out = nullptr;
    f3d_smoke* s = new f3d_smoke();
    int rc = smoke_create_impl(volume, device, s);
    if (rc) { f3d_smoke_destroy(s); return rc; }
    *out = s;
    return 0;
}

// SmokeRenderSettings::validate, src/smoke/types.rs:268-317
static int validate_smoke_settings(const f3d_smoke_settings* s) {
    const char* names[11] = {"density_scale", "extinction", "scattering", "absorption", "phase_g", "step_size", "shadow_step_size",
                             "jitter_strength", "exposure", "soot_absorption", "fire_glow"};
    const float vals[11] = {s->density_scale, s->extinction, s->scattering, s->absorption, s->phase_g, s->step_size,
                            s->shadow_step_size, s->jitter_strength, s->exposure, s->soot_absorption, s->fire_glow};
    for (int i = 0; i < 11; i++)
        if (!isfinite(vals[i])) return fail(F3D_ERR_RENDER, "%s must be finite", names[i]);
    if (s->density_scale < 0.0f || s->extinction < 0.0f || s->scattering < 0.0f)
        return fail(F3D_ERR_RENDER, "density_scale, extinction, and scattering must be >= 0");
    if (s->absorption < 0.0f || s->soot_absorption < 0.0f || s->fire_glow < 0.0f)
        return fail(F3D_ERR_RENDER, "absorption, soot_absorption, and fire_glow must be >= 0");
    if (!(s->phase_g >= -0.99f && s->phase_g <= 0.99f)) return fail(F3D_ERR_RENDER, "phase_g must be in [-0.99, 0.99]");
    if (s->step_size < 0.0f || s->shadow_step_size < 0.0f) return fail(F3D_ERR_RENDER, "step sizes must be >= 0");
    if (s->max_steps == 0 || s->shadow_steps == 0) return fail(F3D_ERR_RENDER, "max_steps and shadow_steps must be >= 1");
    if (!(s->jitter_strength >= 0.0f && s->jitter_strength <= 1.0f)) return fail(F3D_ERR_RENDER, "jitter_strength must be in [0, 1]");
    for (int a = 0; a < 3; a++)
        if (!isfinite(s->thin_color[a]) || s->thin_color[a] < 0.0f) return fail(F3D_ERR_RENDER, "thin_color[%d] must be finite and >= 0", a);
    for (int a = 0; a < 3; a++)
        if (!isfinite(s->dense_color[a]) || s->dense_color[a] < 0.0f) return fail(F3D_ERR_RENDER, "dense_color[%d] must be finite and >= 0", a);
    return 0;
}

static void smoke_common_params(const f3d_smoke* s, const f3d_smoke_settings* st, SmokeParams* P) {
    P->volA = s->volA; P->volB = s->volB;
    P->occ = s->occ; P->occ_dims[0] = s->occ_dims[0]; P->occ_dims[1] = s->occ_dims[1];
    for (int a = 0; a < 3; a++) {
        P->dims[a] = s->dims[a]; P->voxel[a] = s->voxel[a]; P->origin[a] = s->origin[a];
        P->bmax[a] = s->origin[a] + (float)s->dims[a] * s->voxel[a];                       // bounds_max, types.rs:399-405
    }
    SmokeSettings& S = P->s;
    S.density_scale = st->density_scale; S.extinction = st->extinction; S.scattering = st->scattering; S.absorption = st->absorption;
    S.phase_g = st->phase_g; S.max_steps = st->max_steps; S.self_shadow = st->self_shadow ? 1u : 0u; S.shadow_steps = st->shadow_steps;
    S.jitter_strength = st->jitter_strength; S.exposure = st->exposure; S.soot_absorption = st->soot_absorption; S.fire_glow = st->fire_glow;
    memcpy(S.thin_color, st->thin_color, sizeof S.thin_color);
    memcpy(S.dense_color, st->dense_color, sizeof S.dense_color);
    const float min_step = fmaxf(fminf(fminf(fminf(INFINITY, s->voxel[0]), s->voxel[1]), s->voxel[2]), 1.0e-4f);   // render.rs:43-59
    P->step = st->step_size > 0.0f ? st->step_size : min_step * 0.75f;
    P->shadow_step = st->shadow_step_size > 0.0f ? st->shadow_step_size : P->step * 2.0f;
    P->frame_index = s->frame_index;
}

static int smoke_launch(f3d_smoke* s, SmokeParams* P, uint32_t width, uint32_t height, uint8_t* rgba, double* kernel_ms,
                        const uint8_t* base_rgba = nullptr, const float* base_depth = nullptr) {
    if ((uint64_t)width * height > (1ull << 31)) return fail(F3D_ERR_ARGUMENT, "image of %ux%u pixels exceeds the 2^31-pixel addressing limit", width, height);
    const size_t bytes = (size_t)width * height * 4;
    CUDA_TRY(cudaSetDevice(s->device));
    if (s->rgba_capacity < bytes) {
        cudaStreamSynchronize(s->stream);
        cached_free(s->d_rgba, s->device);
        s->d_rgba = nullptr; s->rgba_capacity = 0;
        CUDA_TRY(cached_malloc((void**)&s->d_rgba, bytes, s->device));
        s->rgba_capacity = bytes;
    }
    P->W = width; P->H = height; P->rgba = s->d_rgba;
    P->base = nullptr; P->depth = nullptr;
    if (base_rgba || base_depth) {           // the terrain frame the layer goes over: staged beside the output
        if (s->base_capacity < bytes) {
            cudaStreamSynchronize(s->stream);
            cached_free(s->d_base, s->device); cached_free(s->d_depth, s->device);
            s->d_base = nullptr; s->d_depth = nullptr; s->base_capacity = 0;
            CUDA_TRY(cached_malloc((void**)&s->d_base, bytes, s->device));
            CUDA_TRY(cached_malloc((void**)&s->d_depth, bytes, s->device));
            s->base_capacity = bytes;
        }
        if (base_rgba) { CUDA_TRY(cudaMemcpyAsync(s->d_base, base_rgba, bytes, cudaMemcpyHostToDevice, s->stream)); P->base = (const uchar4*)s->d_base; }
        if (base_depth) { CUDA_TRY(cudaMemcpyAsync(s->d_depth, base_depth, bytes, cudaMemcpyHostToDevice, s->stream)); P->depth = s->d_depth; }
    }
    const dim3 grid((width + 15u) / 16u, (height + 7u) / 8u);
    CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
    k_smoke_march<<<grid, kSmokeThreads, 0, s->stream>>>(*P);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
    CUDA_TRY(cudaMemcpyAsync(rgba, s->d_rgba, bytes, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (kernel_ms) {
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        *kernel_ms = ms;
    }
    return 0;
}

extern "C" int f3d_smoke_raymarch_rgba(f3d_smoke* s, const f3d_smoke_settings* st, uint32_t width, uint32_t height,
                                       const float camera_pos[3], const float target[3], const float up_in[3], float fovy_deg,
                                       const float sun_direction[3], uint8_t* rgba, double* kernel_ms) {
    g_err[0] = 0;
    if (!s || !st || !camera_pos || !target || !up_in || !sun_direction || !rgba) return fail(F3D_ERR_ARGUMENT, "null argument");
    int rc = validate_smoke_settings(st);                                                    // render.rs:18-41, same order and text
    if (rc) return rc;
    if (width == 0 || height == 0) return fail(F3D_ERR_RENDER, "width and height must be >= 1");
    if (!isfinite(fovy_deg) || fovy_deg <= 0.0f || fovy_deg >= 179.0f) return fail(F3D_ERR_RENDER, "fovy_deg must be finite and in (0, 179)");
    const hv3 eye = HV(camera_pos);
    const hv3 forward = hnorm_or_zero(hsub(HV(target), eye));
    if (hdot(forward, forward) < 1.0e-12f) return fail(F3D_ERR_RENDER, "camera_pos and target must not be equal");
    const hv3 up = hnorm_or_zero(HV(up_in));
    if (hdot(up, up) < 1.0e-12f) return fail(F3D_ERR_RENDER, "up vector must not be zero");
    const hv3 right = hnorm_or_zero(hcross(forward, up));
    const hv3 camera_up = hnorm_or_zero(hcross(right, forward));
    const hv3 sun = hnorm_or_zero(HV(sun_direction));
    if (hdot(sun, sun) < 1.0e-12f) return fail(F3D_ERR_RENDER, "sun_direction must not be zero");
    SmokeParams P{};
    smoke_common_params(s, st, &P);
    P.projection = 0u;
    P.eye[0] = eye.x; P.eye[1] = eye.y; P.eye[2] = eye.z;
    P.forward[0] = forward.x; P.forward[1] = forward.y; P.forward[2] = forward.z;
    P.right[0] = right.x; P.right[1] = right.y; P.right[2] = right.z;
    P.up[0] = camera_up.x; P.up[1] = camera_up.y; P.up[2] = camera_up.z;
    P.tan_half_fov = tanf(to_radians_f32(fovy_deg) * 0.5f);
    P.aspect = (float)width / (float)height;
    P.sun_dir[0] = sun.x; P.sun_dir[1] = sun.y; P.sun_dir[2] = sun.z;
    return smoke_launch(s, &P, width, height, rgba, kernel_ms, s->over_base, s->over_depth);
}

extern "C" int f3d_smoke_raymarch_over_rgba(f3d_smoke* s, const f3d_smoke_settings* st, uint32_t width, uint32_t height,
                                            const float camera_pos[3], const float target[3], const float up_in[3], float fovy_deg,
                                            const float sun_direction[3], const uint8_t* base_rgba, const float* base_depth,
                                            uint8_t* rgba, double* kernel_ms) {
    g_err[0] = 0;
    if (!base_rgba) return fail(F3D_ERR_ARGUMENT, "null argument");
    s->over_base = base_rgba; s->over_depth = base_depth;
    const int rc = f3d_smoke_raymarch_rgba(s, st, width, height, camera_pos, target, up_in, fovy_deg, sun_direction, rgba, kernel_ms);
    s->over_base = nullptr; s->over_depth = nullptr;
    return rc;
}

extern "C" int f3d_smoke_raymarch_projection_rgba(f3d_smoke* s, const f3d_smoke_settings* st, uint32_t width, uint32_t height,
                                                  const float view_direction[3], const float sun_direction[3], uint8_t* rgba,
                                                  double* kernel_ms) {
    g_err[0] = 0;
    if (!s || !st || !view_direction || !sun_direction || !rgba) return fail(F3D_ERR_ARGUMENT, "null argument");
    int rc = validate_smoke_settings(st);                                                    // render.rs:111-123
    if (rc) return rc;
    if (width == 0 || height == 0) return fail(F3D_ERR_RENDER, "width and height must be >= 1");
    const hv3 dir = hnorm_or_zero(HV(view_direction));
    if (hdot(dir, dir) < 1.0e-12f) return fail(F3D_ERR_RENDER, "view_direction must not be zero");
    const hv3 sun = hnorm_or_zero(HV(sun_direction));
    if (hdot(sun, sun) < 1.0e-12f) return fail(F3D_ERR_RENDER, "sun_direction must not be zero");
    SmokeParams P{};
    smoke_common_params(s, st, &P);
    P.projection = 1u;
    P.dir[0] = dir.x; P.dir[1] = dir.y; P.dir[2] = dir.z;
    const hv3 ext = hv3{P.bmax[0] - P.origin[0], P.bmax[1] - P.origin[1], P.bmax[2] - P.origin[2]};
    P.diagonal = fmaxf(hlen(ext), P.step * 2.0f);                                            // render.rs:138
    P.sun_dir[0] = sun.x; P.sun_dir[1] = sun.y; P.sun_dir[2] = sun.z;
    return smoke_launch(s, &P, width, height, rgba, kernel_ms);
}

