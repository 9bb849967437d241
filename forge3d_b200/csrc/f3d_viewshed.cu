// forge3d_b200/csrc/f3d_viewshed.cu
// Host side of the HELIOS viewshed / solar shadow mask entry points (f3d_viewshed, f3d_shadow_mask): validation and physics
// terms (f64 host math), the device DEM + min-max chain shared with the path tracer, one kernel launch per call.
// Reference: /root/reference/src/terrain/analysis/viewshed.rs:54-159,161-347,396-570, src/geo/refraction.rs:6-13,100-144.
#include "f3d_host.h"
#include "f3d_viewshed.cuh"

using namespace f3d;
#define g_err g_f3d_err

// ------------------------------------------------------------------------------------------------
// HELIOS viewshed / shadow mask (src/terrain/analysis/viewshed.rs): validation, physics terms, one kernel
// ------------------------------------------------------------------------------------------------
// physics_terms, viewshed.rs:54-78 (+ RefractionModel::k, principal_radii_m: src/geo/refraction.rs:6-13,100-144)
static int viewshed_physics(const f3d_viewshed_options* o, float physics[4]) {
    if (o->earth_model < 0 || o->earth_model > 2) return fail(F3D_ERR_ARGUMENT, "unsupported earth_model %d", o->earth_model);
    if (o->refraction_model < 0 || o->refraction_model > 3) return fail(F3D_ERR_ARGUMENT, "unsupported refraction_model %d", o->refraction_model);
    if (o->earth_model == F3D_EARTH_FLAT && o->refraction_model != F3D_REFRACTION_NONE)
        return fail(F3D_ERR_RENDER, "flat earth only supports refraction_model='none'");
    double k;
    if (o->refraction_model == F3D_REFRACTION_NONE) k = 0.0;
    else if (o->refraction_model == F3D_REFRACTION_EFFECTIVE_RADIUS) k = o->refraction_k;
    else {
        if (!isfinite(o->pressure_mbar) || o->pressure_mbar <= 0.0 || o->temperature_c <= -273.15)
            return fail(F3D_ERR_RENDER, "pressure must be positive and temperature above absolute zero");
        k = (o->refraction_model == F3D_REFRACTION_BENNETT ? 0.13 : 1.0 / 7.0) * (o->pressure_mbar / 1013.25) * (288.15 / (273.15 + o->temperature_c));
    }
    if (!(isfinite(k) && k < 1.0)) return fail(F3D_ERR_RENDER, "refraction k must be finite and less than 1");
    double inv_m = 0.0, inv_p = 0.0;
    if (o->earth_model == F3D_EARTH_SPHERE) {
        if (!(isfinite(o->sphere_radius_m) && o->sphere_radius_m > 0.0)) return fail(F3D_ERR_RENDER, "sphere radius must be finite and positive");
        inv_m = inv_p = 1.0 / o->sphere_radius_m;
    } else if (o->earth_model == F3D_EARTH_ELLIPSOID) {
        if (!isfinite(o->earth_latitude_deg) || o->earth_latitude_deg < -90.0 || o->earth_latitude_deg > 90.0)
            return fail(F3D_ERR_RENDER, "latitude must be finite and in [-90, 90]");
        const double a_m = 6378137.0, e2 = 6.6943799901413165e-3;
        const double sp = sin(deg2rad(o->earth_latitude_deg));
        const double w = sqrt(1.0 - e2 * (sp * sp));
        inv_m = 1.0 / (a_m * (1.0 - e2) / (w * w * w));
        inv_p = 1.0 / (a_m / w);
    }
    physics[0] = (float)inv_m; physics[1] = (float)inv_p; physics[2] = (float)(1.0 - k);
    physics[3] = o->earth_model == F3D_EARTH_FLAT ? 0.0f : 1.0f;
    return 0;
}

// validate_common, viewshed.rs:80-143 (the element counts are implied by the pointer contract here)
static int viewshed_validate(const float* heights, const float* extra, size_t extra_per_cell, bool extra_is_position,
                             const f3d_viewshed_options* o, float physics[4]) {
    if (o->width < 2 || o->height < 2 || o->width - 1 > 8192u || o->height - 1 > 8192u)
        return fail(F3D_ERR_RENDER,
                    "DEM/position lengths do not match supported dimensions %ux%u (both dimensions must be at least 2 and packed traversal supports at most 8192 cells per axis)",
                    o->width, o->height);
    const size_t n = (size_t)o->width * o->height;
    bool finite = true;
    for (size_t i = 0; i < n && finite; i++) finite = isfinite(heights[i]);
    for (size_t i = 0; i < n * extra_per_cell && finite; i++) finite = isfinite(extra[i]);
    if (!finite) {
        if (!extra_is_position) {
            bool extra_ok = true;
            for (size_t i = 0; i < n * extra_per_cell && extra_ok; i++) extra_ok = isfinite(extra[i]);
            if (!extra_ok) return fail(F3D_ERR_RENDER, "shadow-mask geodetic/solar inputs do not match the DEM");
        }
        return fail(F3D_ERR_RENDER, "DEM heights and geodesic positions must be finite (%zu heights)", n);
    }
    const float f[12] = {o->observer_x, o->observer_y, o->observer_height_m, o->target_height_m, o->max_distance_m, o->observer_latitude_rad,
                         o->observer_longitude_rad, o->left_unwrapped_deg, o->top_deg, o->longitude_step_deg, o->latitude_step_deg,
                         o->geodesic_sphere_radius_m};
    bool ok = true;
    for (float v : f) ok = ok && isfinite(v);
    if (!ok || o->observer_x < -0.5f || o->observer_x > (float)o->width - 0.5f || o->observer_y < -0.5f || o->observer_y > (float)o->height - 0.5f ||
        o->observer_height_m < 0.0f || o->target_height_m < 0.0f || o->max_distance_m <= 0.0f || o->longitude_step_deg <= 0.0f ||
        o->latitude_step_deg <= 0.0f || o->geodesic_sphere_radius_m < 0.0f)
        return fail(F3D_ERR_RENDER, "viewshed dimensions, observer, heights, spacing, and distance are invalid");
    return viewshed_physics(o, physics);
}

// height_at(observer.xy), terrain_viewshed.wgsl:24-43, on the host (one value for the whole dispatch)
static float viewshed_height_at(const float* h, uint32_t w, uint32_t hh, float px, float py) {
    const float x = fminf(fmaxf(px, 0.0f), (float)(w - 1u)), y = fminf(fmaxf(py, 0.0f), (float)(hh - 1u));
    const uint32_t x0 = (uint32_t)floorf(x), y0 = (uint32_t)floorf(y);
    const uint32_t x1 = std::min(x0 + 1u, w - 1u), y1 = std::min(y0 + 1u, hh - 1u);
    const float fx = x - (float)x0, fy = y - (float)y0;
    auto mix = [](float a, float b, float t) { const float d = b - a; const float s = d * t; return a + s; };
    return mix(mix(h[(size_t)y0 * w + x0], h[(size_t)y0 * w + x1], fx), mix(h[(size_t)y1 * w + x0], h[(size_t)y1 * w + x1], fx), fy);
}

struct ViewshedRun {
    DeviceTerrain T;
    float* d_heights = nullptr;
    void* d_extra = nullptr;
    void* d_out = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int device = 0;
    ~ViewshedRun() {
        cudaDeviceSynchronize();
        T.release();
        cached_free(d_heights, device); cached_free(d_extra, device); cached_free(d_out, device);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
};

static int viewshed_setup(const float* heights, const void* extra, size_t extra_bytes, size_t out_bytes, const f3d_viewshed_options* o,
                          const float physics[4], ViewshedRun* R, ViewshedParams* P) {
    int rc = select_device(o->device);
    if (rc) return rc;
    R->device = o->device;
    uint64_t launches = 0;
    // the tracked DEM + min-max chain of TerrainMinMaxPyramid::from_heightfield (viewshed.rs:207-213), built on the device
    if ((rc = build_device_terrain(heights, o->width, o->height, 1.0f, nullptr, &R->T, &launches, true))) return rc;
    const size_t n = (size_t)o->width * o->height;
    CUDA_TRY(cached_malloc((void**)&R->d_heights, n * sizeof(float), o->device));
    CUDA_TRY(cached_malloc(&R->d_extra, extra_bytes, o->device));
    CUDA_TRY(cached_malloc(&R->d_out, out_bytes, o->device));
    CUDA_TRY(cudaMemcpy(R->d_heights, heights, n * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(R->d_extra, extra, extra_bytes, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaEventCreate(&R->ev0));
    CUDA_TRY(cudaEventCreate(&R->ev1));
    P->w = o->width; P->h = o->height;
    P->observer[0] = o->observer_x; P->observer[1] = o->observer_y; P->observer[2] = o->observer_height_m; P->observer[3] = o->target_height_m;
    P->metric[0] = o->max_distance_m; P->metric[1] = o->longitude_step_deg; P->metric[2] = o->latitude_step_deg; P->metric[3] = o->geodesic_sphere_radius_m;
    memcpy(P->physics, physics, sizeof P->physics);
    P->geodetic[0] = o->observer_latitude_rad; P->geodetic[1] = o->observer_longitude_rad; P->geodetic[2] = o->left_unwrapped_deg; P->geodetic[3] = o->top_deg;
    P->observer_elevation = viewshed_height_at(heights, o->width, o->height, o->observer_x, o->observer_y) + o->observer_height_m;
    P->heights = R->d_heights; P->cells = R->T.cells;
    for (int l = 0; l < 16; l++) {
        P->mm[l] = l < R->T.nlevels ? R->T.mm_base + R->T.level_off[l] : nullptr;
        P->mm_pitch[l] = l < R->T.nlevels ? R->T.dims[l][0] : 0u;
    }
    P->root_level = (uint32_t)R->T.nlevels - 1u;
    return 0;
}

extern "C" int f3d_viewshed(const float* heights, const float* positions_m, const f3d_viewshed_options* o, uint8_t* visibility,
                            float* drop, float* gain, float* horizon, double* kernel_ms) {
    g_err[0] = 0;
    if (!heights || !positions_m || !o || !visibility || !drop || !gain || !horizon) return fail(F3D_ERR_ARGUMENT, "null argument");
    float physics[4];
    int rc = viewshed_validate(heights, positions_m, 2, true, o, physics);
    if (rc) return rc;
    const size_t n = (size_t)o->width * o->height;
    ViewshedRun R;
    ViewshedParams P{};
    if ((rc = viewshed_setup(heights, positions_m, n * sizeof(float2), n * sizeof(float4), o, physics, &R, &P))) return rc;
    const dim3 grid((o->width + 7u) / 8u, (o->height + 7u) / 8u);
    CUDA_TRY(cudaEventRecord(R.ev0, 0));
    k_viewshed<<<grid, kViewshedThreads>>>(P, (const float2*)R.d_extra, (float4*)R.d_out);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(R.ev1, 0));
    std::vector<float4> cells(n);
    CUDA_TRY(cudaMemcpy(cells.data(), R.d_out, n * sizeof(float4), cudaMemcpyDeviceToHost));
    if (kernel_ms) { float ms = 0.0f; CUDA_TRY(cudaEventElapsedTime(&ms, R.ev0, R.ev1)); *kernel_ms = ms; }
    for (size_t i = 0; i < n; i++) {
        uint32_t v;
        memcpy(&v, &cells[i].x, 4);
        if (v > 1u) return fail(F3D_ERR_RENDER, "viewshed geodesic leaves the DEM footprint");   // viewshed.rs:320-327
    }
    for (size_t i = 0; i < n; i++) {
        uint32_t v;
        memcpy(&v, &cells[i].x, 4);
        visibility[i] = v != 0u; drop[i] = cells[i].y; gain[i] = cells[i].z; horizon[i] = cells[i].w;
    }
    return 0;
}

extern "C" int f3d_shadow_mask(const float* heights, const float* geodetic_and_sun, const f3d_viewshed_options* o, uint8_t* lit,
                               double* kernel_ms) {
    g_err[0] = 0;
    if (!heights || !geodetic_and_sun || !o || !lit) return fail(F3D_ERR_ARGUMENT, "null argument");
    float physics[4];
    int rc = viewshed_validate(heights, geodetic_and_sun, 4, false, o, physics);
    if (rc) return rc;
    const size_t n = (size_t)o->width * o->height;
    ViewshedRun R;
    ViewshedParams P{};
    if ((rc = viewshed_setup(heights, geodetic_and_sun, n * sizeof(float4), n, o, physics, &R, &P))) return rc;
    const dim3 grid((o->width + 7u) / 8u, (o->height + 7u) / 8u);
    CUDA_TRY(cudaEventRecord(R.ev0, 0));
    k_shadow_mask<<<grid, kViewshedThreads>>>(P, (const float4*)R.d_extra, (uint8_t*)R.d_out);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(R.ev1, 0));
    CUDA_TRY(cudaMemcpy(lit, R.d_out, n, cudaMemcpyDeviceToHost));
    if (kernel_ms) { float ms = 0.0f; CUDA_TRY(cudaEventElapsedTime(&ms, R.ev0, R.ev1)); *kernel_ms = ms; }
    return 0;
}

