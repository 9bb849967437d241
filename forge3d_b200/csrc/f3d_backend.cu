// forge3d_b200/csrc/f3d_backend.cu
// Host side of libforge3d_b200.so: trust-boundary validation, uniform setup, device memory, the
// accumulate-until-converged driver loop and the C ABI declared in include/forge3d_b200.h.
// This is the C++ stand-in for the reference's Rust driver
// /root/reference/src/path_tracing/hybrid_compute/render_terrain.rs:474-1434 (+ terrain_heightfield.rs
// :52-84,:224-369 and src/geo/refraction.rs): Rust is not available in this image, so the layer that
// is Rust in the reference is C++ here and exports the C ABI a Rust `extern "C"` shim would bind.
// There is no CPU fallback anywhere in this file: every path that cannot reach a CUDA device fails.
#ifndef EMU_SIMT
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost nothing unless a profiler is attached
#else
static inline void nvtxRangePushA(const char*) {}
static inline void nvtxRangePop() {}
#endif
#include <stdarg.h>

#include <map>
#include <mutex>
#include <string>
#include <unordered_map>

#include "f3d_host.h"
#include "f3d_kernels.cuh"
#include "f3d_lbvh.cuh"

using namespace f3d;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
thread_local char g_f3d_err[640];
#define g_err g_f3d_err

int f3d_fail(int cls, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_f3d_err, sizeof g_f3d_err, fmt, ap);
    va_end(ap);
    return cls;
}


// ------------------------------------------------------------------------------------------------
// Device-buffer cache.  A drop-in call allocates ~0.6 GB of state per 1080p render; cudaMalloc /
// cudaFree of that much memory costs 3-70 ms per call (page-table work), which is visible next to a
// 350 ms render.  Freed blocks are parked per device and handed back to the next session that asks
// for a similar size.  Buffers exported over CUDA IPC are never cached.
// ------------------------------------------------------------------------------------------------
namespace {
struct DevCache {
    std::mutex mu;
    std::multimap<size_t, void*> parked[16];
    std::unordered_map<void*, size_t> sizes;
    size_t parked_bytes[16] = {};
    // per-device cap: one session's working set (F3D_B200_CACHE_MB overrides, 0 = never park).  The per-step buffers of a session
    // are budgeted at 8 GB (session_create_impl), so a 4 GB cap made every partitioned 1080p session with batches of 8 steps
    // (4.6 GB) cudaMalloc / cudaFree its overflow on every call: 11-19 ms of session creation at N = 4 / 8 against 1.6 ms at
    // N = 2 (gpurun call X).  10 GB of a 180 GB device covers it.  The process may share the device with torch / NCCL: what is
    // parked here is invisible to them, so f3d_cache_trim() hands everything back on request.
    size_t cap() const {
        static const size_t c = [] { const char* e = getenv("F3D_B200_CACHE_MB"); return e ? (size_t)std::max(atoll(e), 0ll) << 20 : (size_t)10 << 30; }();
        return c;
    }
};
DevCache g_cache;
}  // namespace

cudaError_t cached_malloc(void** p, size_t bytes, int device) {
    bytes = std::max<size_t>((bytes + 511) & ~size_t(511), 512);
    {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        auto& m = g_cache.parked[device & 15];
        auto it = m.lower_bound(bytes);
        if (it != m.end() && it->first <= bytes + bytes / 4 + 4096) {
            *p = it->second;
            g_cache.parked_bytes[device & 15] -= it->first;
            m.erase(it);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {   // out of memory: drop everything parked on this device and retry once
        cudaGetLastError();
        std::vector<void*> drop;
        {
            std::lock_guard<std::mutex> lk(g_cache.mu);
            for (auto& kv : g_cache.parked[device & 15]) { drop.push_back(kv.second); g_cache.sizes.erase(kv.second); }
            g_cache.parked[device & 15].clear();
            g_cache.parked_bytes[device & 15] = 0;
        }
        for (void* q : drop) cudaFree(q);
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        g_cache.sizes[*p] = bytes;
    }
    return e;
}

// Caller guarantees no work that touches `p` is still in flight.
void cached_free(void* p, int device, bool allow_park) {
    if (!p) return;
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        auto it = g_cache.sizes.find(p);
        if (it != g_cache.sizes.end()) bytes = it->second;
        if (bytes && allow_park && g_cache.parked_bytes[device & 15] + bytes <= g_cache.cap()) {
            g_cache.parked[device & 15].emplace(bytes, p);
            g_cache.parked_bytes[device & 15] += bytes;
            return;
        }
        if (bytes) g_cache.sizes.erase(it);
    }
    cudaFree(p);
}

// ------------------------------------------------------------------------------------------------
// Pinned host memory.  Reading 66 MB of 1080p outputs back into freshly allocated pageable memory ran at ~4 GB/s in round 1
// (page faults on first touch + the driver's pageable staging): 17 ms of a 40 ms call.  Two remedies:
//   * f3d_host_alloc / f3d_host_free: a small pool of page-locked buffers.  The Python layer allocates its output arrays
//     from it, so the copy is one DMA at PCIe speed straight into the array the caller receives;
//   * copy_to_host: any other destination (a C caller's malloc'ed buffer) goes through two page-locked bounce buffers,
//     the DMA of chunk i overlapping the memcpy of chunk i-1.
// ------------------------------------------------------------------------------------------------
namespace {
struct HostPool {
    std::mutex mu;
    std::multimap<size_t, void*> parked;
    std::unordered_map<void*, size_t> sizes;
    size_t parked_bytes = 0;
    static constexpr size_t kMaxParked = 1ull << 30;
    void* bounce[2] = {nullptr, nullptr};
    cudaEvent_t bounce_ev[2] = {nullptr, nullptr};
    static constexpr size_t kBounce = 8ull << 20;
};
HostPool g_host;
}  // namespace

extern "C" void* f3d_host_alloc(uint64_t bytes) {
    bytes = std::max<uint64_t>((bytes + 4095) & ~uint64_t(4095), 4096);
    {
        std::lock_guard<std::mutex> lk(g_host.mu);
        auto it = g_host.parked.lower_bound((size_t)bytes);
        if (it != g_host.parked.end() && it->first <= bytes + bytes / 4) {
            void* p = it->second;
            g_host.parked_bytes -= it->first;
            g_host.parked.erase(it);
            return p;
        }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> lk(g_host.mu);
    g_host.sizes[p] = (size_t)bytes;
    return p;
}

extern "C" void f3d_host_free(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_host.mu);
        auto it = g_host.sizes.find(p);
        if (it == g_host.sizes.end()) return;
        if (g_host.parked_bytes + it->second <= HostPool::kMaxParked) {
            g_host.parked.emplace(it->second, p);
            g_host.parked_bytes += it->second;
            return;
        }
        g_host.sizes.erase(it);
    }
    cudaFreeHost(p);
}

// Device -> host copy on `stream`, synchronous on return for pageable destinations (page-locked ones are only enqueued).
static int copy_to_host(void* host, const void* dev, size_t bytes, cudaStream_t stream) {
    cudaPointerAttributes at{};
    const bool pinned = cudaPointerGetAttributes(&at, host) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned || bytes < (1u << 20)) {
        CUDA_TRY(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, stream));
        return 0;
    }
    std::lock_guard<std::mutex> lk(g_host.mu);          // one staged copy at a time per process
    for (int k = 0; k < 2; k++) {
        if (!g_host.bounce[k]) CUDA_TRY(cudaHostAlloc(&g_host.bounce[k], HostPool::kBounce, cudaHostAllocDefault));
        if (!g_host.bounce_ev[k]) CUDA_TRY(cudaEventCreateWithFlags(&g_host.bounce_ev[k], cudaEventDisableTiming));
    }
    const size_t nchunks = (bytes + HostPool::kBounce - 1) / HostPool::kBounce;
    for (size_t c = 0; c <= nchunks; c++) {
        if (c < nchunks) {
            const size_t off = c * HostPool::kBounce, len = std::min(HostPool::kBounce, bytes - off);
            CUDA_TRY(cudaMemcpyAsync(g_host.bounce[c & 1], (const char*)dev + off, len, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaEventRecord(g_host.bounce_ev[c & 1], stream));
        }
        if (c > 0) {
            const size_t off = (c - 1) * HostPool::kBounce, len = std::min(HostPool::kBounce, bytes - off);
            CUDA_TRY(cudaEventSynchronize(g_host.bounce_ev[(c - 1) & 1]));
            memcpy((char*)host + off, g_host.bounce[(c - 1) & 1], len);
        }
    }
    return 0;
}

// Returns every parked device buffer of `device` (all devices when negative) to the driver; returns the bytes released.
extern "C" uint64_t f3d_cache_trim(int32_t device) {
    std::vector<std::pair<int, void*>> drop;
    uint64_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        for (int d = 0; d < 16; d++) {
            if (device >= 0 && d != (device & 15)) continue;
            for (auto& kv : g_cache.parked[d]) { drop.emplace_back(d, kv.second); bytes += kv.first; g_cache.sizes.erase(kv.second); }
            g_cache.parked[d].clear();
            g_cache.parked_bytes[d] = 0;
        }
    }
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& dp : drop) { cudaSetDevice(dp.first); cudaFree(dp.second); }
    cudaSetDevice(cur);
    cudaGetLastError();
    return bytes;
}

extern "C" const char* f3d_last_error(void) { return g_err; }
extern "C" int f3d_abi_version(void) { return F3D_ABI_VERSION; }
// Identifies the build: content hash of the sources and the -D variant switches it was compiled with (forge3d_b200/build.py
// compares it to decide whether the in-tree library is the default build of the current sources; bench.py prints it).
#ifndef F3D_BUILD_INFO_STR
#define F3D_BUILD_INFO_STR "src=unknown;defines="
#endif
extern "C" const char* f3d_build_info(void) {
    static const char tagged[] = "F3D_BUILD_INFO:" F3D_BUILD_INFO_STR;
    return tagged + 15;
}
extern "C" int f3d_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ------------------------------------------------------------------------------------------------
// host math mirroring glam 0.24.2 / Rust f32 (render_terrain.rs:635-661)
// ------------------------------------------------------------------------------------------------

// validate_desc, render_terrain.rs:474-557 (same order, same message text)
static int validate_desc(const f3d_terrain_desc* d) {
    if (!d->heights) return fail(F3D_ERR_ARGUMENT, "heights pointer is null");
    if (d->width == 0 || d->height == 0 || d->max_frames == 0)
        return fail(F3D_ERR_RENDER, "terrain reference requires non-zero width/height/max_frames");
    if (d->min_frames > d->max_frames)
        return fail(F3D_ERR_RENDER, "min_frames (%u) must be <= max_frames (%u)", d->min_frames, d->max_frames);
    if (d->spp == 0 || d->spp > 64) return fail(F3D_ERR_RENDER, "spp must be in 1..=64, got %u", d->spp);
    if (!(isfinite(d->exaggeration) && d->exaggeration > 0.0f))
        return fail(F3D_ERR_RENDER, "terrain exaggeration must be finite and > 0");
    if (!(finite3(d->cam_origin) && finite3(d->cam_look_at) && finite3(d->cam_up)))
        return fail(F3D_ERR_RENDER, "camera origin/look_at/up must be finite");
    hv3 fwd = hsub(HV(d->cam_look_at), HV(d->cam_origin));
    if (hlen(fwd) < 1e-6f) return fail(F3D_ERR_RENDER, "camera look_at must differ from origin");
    if (hlen(hcross(hnorm(fwd), HV(d->cam_up))) < 1e-6f)
        return fail(F3D_ERR_RENDER, "camera up vector must not be parallel to the view direction");
    if (!(isfinite(d->fov_y_deg) && d->fov_y_deg > 0.0f && d->fov_y_deg < 180.0f))
        return fail(F3D_ERR_RENDER, "fov_y must be finite and in (0, 180) degrees, got %g", d->fov_y_deg);
    if (!(isfinite(d->exposure) && d->exposure > 0.0f)) return fail(F3D_ERR_RENDER, "exposure must be finite and > 0");
    if (!(isfinite(d->sun_az_deg) && isfinite(d->sun_el_deg)))
        return fail(F3D_ERR_RENDER, "sun azimuth/elevation must be finite");
    if (!(isfinite(d->sun_intensity) && d->sun_intensity >= 0.0f))
        return fail(F3D_ERR_RENDER, "sun intensity must be finite and >= 0");
    if (!finite3(d->sun_color) || d->sun_color[0] < 0 || d->sun_color[1] < 0 || d->sun_color[2] < 0)
        return fail(F3D_ERR_RENDER, "sun color must have three finite non-negative components");
    if (!(isfinite(d->env_intensity) && d->env_intensity >= 0.0f))
        return fail(F3D_ERR_RENDER, "env intensity must be finite and >= 0");
    if (!(isfinite(d->variance_threshold) && d->variance_threshold > 0.0f))
        return fail(F3D_ERR_RENDER, "variance threshold must be finite and > 0");
    if (!(isfinite(d->spacing[0]) && d->spacing[0] > 0.0f && isfinite(d->spacing[1]) && d->spacing[1] > 0.0f))
        return fail(F3D_ERR_RENDER, "terrain spacing must be finite and > 0, got (%g, %g)", d->spacing[0], d->spacing[1]);
    if (d->mesh_xyz || d->mesh_idx) {
        if (!d->mesh_xyz || d->mesh_nverts == 0)
            return fail(F3D_ERR_RENDER, "mesh vertices must be a non-empty flat [x,y,z] list");
        if (!d->mesh_idx || d->mesh_ntris == 0)
            return fail(F3D_ERR_RENDER, "mesh indices must be a non-empty multiple of 3");
        for (size_t i = 0; i < (size_t)d->mesh_nverts * 3; i++)
            if (!isfinite(d->mesh_xyz[i])) return fail(F3D_ERR_RENDER, "mesh vertices contain non-finite values");
        for (size_t i = 0; i < (size_t)d->mesh_ntris * 3; i++)
            if (d->mesh_idx[i] >= d->mesh_nverts)
                return fail(F3D_ERR_RENDER, "mesh indices reference out-of-bounds vertices");
    }
    // TerrainPtScene::new, terrain_heightfield.rs:402-421
    if (!(isfinite(d->albedo[0]) && d->albedo[0] >= 0 && isfinite(d->albedo[1]) && d->albedo[1] >= 0 &&
          isfinite(d->albedo[2]) && d->albedo[2] >= 0))
        return fail(F3D_ERR_UPLOAD, "terrain albedo must be finite and >= 0");
    // build_minmax_mips, terrain_heightfield.rs:133-137 (the non-finite scan runs on the device)
    if (d->dem_w < 2 || d->dem_h < 2)
        return fail(F3D_ERR_UPLOAD, "terrain heightfield must be at least 2x2 texels, got %ux%u", d->dem_w, d->dem_h);
    if (d->dem_w - 1 > 8192 || d->dem_h - 1 > 8192)
        return fail(F3D_ERR_UPLOAD, "terrain heightfield exceeds 8192 cells per axis (13-bit node packing), got %ux%u",
                    d->dem_w, d->dem_h);
    if (d->env_rgb) {
        if (d->env_w == 0 || d->env_h == 0) return fail(F3D_ERR_UPLOAD, "env map dims do not match data length");
        for (size_t i = 0; i < (size_t)d->env_w * d->env_h * 3; i++)
            if (!isfinite(d->env_rgb[i])) return fail(F3D_ERR_UPLOAD, "env map contains non-finite samples");
    }
    if ((uint64_t)d->width * d->height > (1ull << 31))   // kernels index pixels with 32-bit integers
        return fail(F3D_ERR_ARGUMENT, "image of %ux%u pixels exceeds the 2^31-pixel addressing limit", d->width, d->height);
    if (d->part_world > 1 && d->part_rank >= d->part_world)
        return fail(F3D_ERR_ARGUMENT, "part_rank (%u) must be < part_world (%u)", d->part_rank, d->part_world);
    if (d->part_mode > 1) return fail(F3D_ERR_ARGUMENT, "unsupported part_mode %u", d->part_mode);
    if (d->atmosphere && !(d->atmosphere->transmittance && d->atmosphere->scattering && d->atmosphere->aerial))
        return fail(F3D_ERR_ARGUMENT, "atmosphere LUT pointer is null");
    return 0;
}

// AtmosphereConfig::validate + LutDimensions::validate (src/core/atmosphere/bake.rs:75-100,165-220) as
// AetherPostPass::new reports them (aether_post.rs:58-65), then validate_luts / upload_lut (:343-385).
static int validate_atmosphere(const f3d_atmosphere& a) {
    const char* pre = "invalid AETHER PT settings: invalid atmosphere configuration: ";
    const float scalars[9] = {a.turbidity, a.ozone_du, a.mie_g, a.bottom_radius_m, a.top_radius_m, a.rayleigh_scale_height_m,
                              a.mie_scale_height_m, a.max_aerial_distance_m, a.ground_albedo};
    for (float v : scalars)
        if (!isfinite(v)) return fail(F3D_ERR_RENDER, "%sall scalar parameters must be finite", pre);
    if (!(a.turbidity >= 1.0f && a.turbidity <= 10.0f)) return fail(F3D_ERR_RENDER, "%sturbidity must be in [1, 10]", pre);
    if (!(a.ozone_du >= 0.0f && a.ozone_du <= 600.0f)) return fail(F3D_ERR_RENDER, "%sozone must be in [0, 600] DU", pre);
    if (!(a.mie_g >= 0.0f && a.mie_g <= 0.99f)) return fail(F3D_ERR_RENDER, "%smie_g must be in [0, 0.99]", pre);
    if (a.bottom_radius_m <= 0.0f || a.top_radius_m <= a.bottom_radius_m)
        return fail(F3D_ERR_RENDER, "%stop radius must exceed a positive bottom radius", pre);
    if (a.rayleigh_scale_height_m <= 0.0f || a.mie_scale_height_m <= 0.0f || a.max_aerial_distance_m <= 0.0f)
        return fail(F3D_ERR_RENDER, "%sscale heights and aerial distance must be positive", pre);
    if (!(a.ground_albedo >= 0.0f && a.ground_albedo <= 1.0f)) return fail(F3D_ERR_RENDER, "%sground albedo must be in [0, 1]", pre);
    const uint32_t axes[9] = {a.transmittance_mu, a.transmittance_height, a.scattering_mu_view, a.scattering_mu_sun,
                              a.scattering_height, a.scattering_nu, a.aerial_distance, a.aerial_mu_view, a.aerial_height};
    for (uint32_t ax : axes)
        if (ax < 2u) return fail(F3D_ERR_RENDER, "%severy atmosphere LUT axis must contain at least two samples", pre);
    for (uint32_t ax : axes)
        if (ax > 256u) return fail(F3D_ERR_RENDER, "%satmosphere LUT axes are capped at 256 samples", pre);
    return 0;
}

// effective_radius_m + EarthCurvatureUniforms::new (src/geo/refraction.rs:57-144, terrain_heightfield.rs:52-84)
static int earth_curvature(const f3d_terrain_desc* d, float* inv_two_r_prime, uint32_t* enabled) {
    if (!(isfinite(d->observer_lat_deg) && d->observer_lat_deg >= -90.0 && d->observer_lat_deg <= 90.0 &&
          isfinite(d->observer_lon_deg) && d->observer_lon_deg >= -180.0 && d->observer_lon_deg <= 180.0))
        return fail(F3D_ERR_RENDER, "ray-origin latitude/longitude must be finite and in [-90,90]/[-180,180]");
    if (d->earth_model < 0 || d->earth_model > 2) return fail(F3D_ERR_ARGUMENT, "unsupported earth_model %d", d->earth_model);
    if (d->refraction_model < 0 || d->refraction_model > 3)
        return fail(F3D_ERR_ARGUMENT, "unsupported refraction_model %d", d->refraction_model);
    if (d->earth_model == F3D_EARTH_FLAT && d->refraction_model != F3D_REFRACTION_NONE)
        return fail(F3D_ERR_RENDER, "flat earth only supports refraction_model='none'");
    const double az_deg = (double)d->sun_az_deg;
    if (!isfinite(az_deg)) return fail(F3D_ERR_RENDER, "azimuth must be finite");
    double radius;
    if (d->earth_model == F3D_EARTH_FLAT) radius = INFINITY;
    else if (d->earth_model == F3D_EARTH_SPHERE) {
        if (!(isfinite(d->sphere_radius_m) && d->sphere_radius_m > 0.0))
            return fail(F3D_ERR_RENDER, "sphere radius must be finite and positive");
        radius = d->sphere_radius_m;
    } else {
        const double a_m = 6378137.0, e2 = 6.6943799901413165e-3;
        double phi = deg2rad(d->observer_lat_deg), sp = sin(phi);
        double w = sqrt(1.0 - e2 * (sp * sp));
        double meridional = a_m * (1.0 - e2) / (w * w * w), prime_vertical = a_m / w;
        double az = deg2rad(az_deg), ca = cos(az), sa = sin(az);
        radius = 1.0 / ((ca * ca) / meridional + (sa * sa) / prime_vertical);
    }
    double k;
    if (d->refraction_model == F3D_REFRACTION_NONE) k = 0.0;
    else if (d->refraction_model == F3D_REFRACTION_EFFECTIVE_RADIUS) k = d->refraction_k;
    else {
        if (!isfinite(d->pressure_mbar) || d->pressure_mbar <= 0.0 || d->temperature_c <= -273.15)
            return fail(F3D_ERR_RENDER, "pressure must be positive and temperature above absolute zero");
        double base = d->refraction_model == F3D_REFRACTION_BENNETT ? 0.13 : 1.0 / 7.0;
        k = base * (d->pressure_mbar / 1013.25) * (288.15 / (273.15 + d->temperature_c));
    }
    if (!(isfinite(k) && k < 1.0)) return fail(F3D_ERR_RENDER, "refraction k must be finite and less than 1");
    double eff = radius / (1.0 - k);
    bool en = isfinite(eff);
    *inv_two_r_prime = en ? (float)(0.5 / eff) : 0.0f;
    *enabled = en ? 1u : 0u;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Mesh BVH (replaces the reference's unused upload of accel::build_bvh output, render_terrain.rs:597-632,
// and makes hybrid_traversal.wgsl:137-172's O(#tris) sweep O(log #tris) with identical hits).
// Median split on the widest centroid axis, leaves of <= 4 triangles, boxes padded against rounding.
// ------------------------------------------------------------------------------------------------
struct MeshBvh {
    std::vector<float4> nodes;      // 2 per node
    std::vector<uint32_t> tris;     // triangle ids in leaf order
};

static void bvh_build_rec(MeshBvh& B, std::vector<uint32_t>& ids, size_t lo, size_t hi, const std::vector<float>& tb /*6 per tri*/,
                          const std::vector<float>& cen /*3 per tri*/, size_t node) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    float cmn[3] = {INFINITY, INFINITY, INFINITY}, cmx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t k = lo; k < hi; k++) {
        const uint32_t t = ids[k];
        for (int a = 0; a < 3; a++) {
            mn[a] = fminf(mn[a], tb[6 * t + a]); mx[a] = fmaxf(mx[a], tb[6 * t + 3 + a]);
            cmn[a] = fminf(cmn[a], cen[3 * t + a]); cmx[a] = fmaxf(cmx[a], cen[3 * t + a]);
        }
    }
    const size_t count = hi - lo;
    if (count <= 4) {
        B.nodes[2 * node] = make_float4(mn[0], mn[1], mn[2], 0.0f);
        B.nodes[2 * node + 1] = make_float4(mx[0], mx[1], mx[2], 0.0f);
        const uint32_t first = 0x80000000u | (uint32_t)B.tris.size(), cnt = (uint32_t)count;   // leaf flag, see intersect_mesh
        for (size_t k = lo; k < hi; k++) B.tris.push_back(ids[k]);
        memcpy(&B.nodes[2 * node].w, &first, 4);
        memcpy(&B.nodes[2 * node + 1].w, &cnt, 4);
        return;
    }
    int axis = 0;
    if (cmx[1] - cmn[1] > cmx[axis] - cmn[axis]) axis = 1;
    if (cmx[2] - cmn[2] > cmx[axis] - cmn[axis]) axis = 2;
    const size_t mid = lo + count / 2;
    std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](uint32_t a, uint32_t b) {
        const float ca = cen[3 * a + axis], cb = cen[3 * b + axis];
        return ca < cb || (ca == cb && a < b);
    });
    const uint32_t left = (uint32_t)(B.nodes.size() / 2), zero = left + 1u;                // right child
    B.nodes.resize(B.nodes.size() + 4);
    B.nodes[2 * node] = make_float4(mn[0], mn[1], mn[2], 0.0f);
    B.nodes[2 * node + 1] = make_float4(mx[0], mx[1], mx[2], 0.0f);
    memcpy(&B.nodes[2 * node].w, &left, 4);
    memcpy(&B.nodes[2 * node + 1].w, &zero, 4);
    bvh_build_rec(B, ids, lo, mid, tb, cen, left);
    bvh_build_rec(B, ids, mid, hi, tb, cen, left + 1);
}

static void build_mesh_bvh(const float* xyz, const uint32_t* idx, uint32_t ntris, MeshBvh* B) {
    std::vector<float> tb((size_t)ntris * 6), cen((size_t)ntris * 3);
    std::vector<uint32_t> ids(ntris);
    for (uint32_t t = 0; t < ntris; t++) {
        ids[t] = t;
        float ext = 0.0f, scale = 0.0f;
        for (int a = 0; a < 3; a++) {
            const float p0 = xyz[3 * (size_t)idx[3 * t] + a], p1 = xyz[3 * (size_t)idx[3 * t + 1] + a], p2 = xyz[3 * (size_t)idx[3 * t + 2] + a];
            const float lo = fminf(p0, fminf(p1, p2)), hi = fmaxf(p0, fmaxf(p1, p2));
            tb[6 * t + a] = lo; tb[6 * t + 3 + a] = hi;
            cen[3 * t + a] = 0.5f * (lo + hi);
            ext = fmaxf(ext, hi - lo);
            scale = fmaxf(scale, fmaxf(fabsf(lo), fabsf(hi)));
        }
        // pad: Moeller-Trumbore's rounding error is ~1e-6 of the triangle's size / coordinates; 1e-3 is ample
        const float pad = 1e-3f * ext + 1e-5f * scale + 1e-6f;
        for (int a = 0; a < 3; a++) { tb[6 * t + a] -= pad; tb[6 * t + 3 + a] += pad; }
    }
    B->nodes.assign(2, make_float4(0, 0, 0, 0));
    B->tris.clear();
    B->tris.reserve(ntris);
    bvh_build_rec(*B, ids, 0, ntris, tb, cen, 0);
}

void host_build_mesh_bvh(const float* xyz, const uint32_t* idx, uint32_t ntris, std::vector<float4>* nodes, std::vector<uint32_t>* tris) {
    MeshBvh B;
    build_mesh_bvh(xyz, idx, ntris, &B);
    nodes->swap(B.nodes);
    tris->swap(B.tris);
}


// ------------------------------------------------------------------------------------------------
// GPU LBVH build (csrc/f3d_lbvh.cuh): d_verts (float4) and d_idx already on the device; writes 2 * (2n - 1) float4 nodes in
// the traversal format and the n triangle ids in leaf order.  The scene box is taken on the host (the vertices were just
// scanned there by validate_desc; compute_scene_aabb, src/accel/types.rs:298-305).
// ------------------------------------------------------------------------------------------------
struct LbvhScratch {
    unsigned long long* keys = nullptr;
    float4* tri_boxes = nullptr;
    uint32_t *left = nullptr, *right = nullptr, *parent = nullptr, *arrivals = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    ~LbvhScratch() {
        cudaStreamSynchronize(stream);
        cached_free(keys, device); cached_free(tri_boxes, device); cached_free(left, device); cached_free(right, device);
        cached_free(parent, device); cached_free(arrivals, device);
    }
};

static int build_mesh_lbvh(const float* xyz, uint32_t nverts, const uint32_t* idx, uint32_t ntris, const float4* d_verts,
                           const uint32_t* d_idx, int device, cudaStream_t stream, float4* d_nodes, uint32_t* d_order, uint64_t* launches,
                           LbvhScratch* keep = nullptr) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t t = 0; t < (size_t)ntris * 3; t++)               // only vertices referenced by a triangle count
        for (int a = 0; a < 3; a++) {
            const float v = xyz[3 * (size_t)idx[t] + a];
            mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v);
        }
    (void)nverts;
    v3 wmin, wext;                                               // MortonUniforms, lbvh_gpu/morton.rs:14-27
    wmin.x = mn[0]; wmin.y = mn[1]; wmin.z = mn[2];
    wext.x = fmaxf(mx[0] - mn[0], 1e-6f); wext.y = fmaxf(mx[1] - mn[1], 1e-6f); wext.z = fmaxf(mx[2] - mn[2], 1e-6f);
    const uint32_t padded = next_pow2(ntris);
    LbvhScratch local;
    LbvhScratch& W = keep ? *keep : local;
    W.device = device; W.stream = stream;
    CUDA_TRY(cached_malloc((void**)&W.keys, (size_t)padded * sizeof(unsigned long long), device));
    CUDA_TRY(cached_malloc((void**)&W.tri_boxes, (size_t)ntris * 2 * sizeof(float4), device));
    CUDA_TRY(cached_malloc((void**)&W.left, (size_t)std::max(ntris, 2u) * sizeof(uint32_t), device));
    CUDA_TRY(cached_malloc((void**)&W.right, (size_t)std::max(ntris, 2u) * sizeof(uint32_t), device));
    CUDA_TRY(cached_malloc((void**)&W.parent, (size_t)(2 * ntris) * sizeof(uint32_t), device));
    CUDA_TRY(cached_malloc((void**)&W.arrivals, (size_t)std::max(ntris, 2u) * sizeof(uint32_t), device));
    CUDA_TRY(cudaMemsetAsync(W.arrivals, 0, (size_t)std::max(ntris, 2u) * sizeof(uint32_t), stream));
    CUDA_TRY(cudaMemsetAsync(W.parent, 0xFF, (size_t)(2 * ntris) * sizeof(uint32_t), stream));
    const unsigned tb = 256;
    k_lbvh_prims<<<(padded + tb - 1) / tb, tb, 0, stream>>>(d_verts, d_idx, ntris, wmin, wext, W.keys, padded, W.tri_boxes);
    (*launches)++;
    // bitonic network over `padded` keys: every stage that stays inside a 2048-key tile runs in shared memory
    // (one launch sorts all tiles; each later merge step k needs log2(k / tile) global stages and one tile launch)
    const unsigned tiles = (padded + kBitonicTile - 1) / kBitonicTile;
    k_lbvh_bitonic_tile<<<tiles, 1024, kBitonicTile * sizeof(unsigned long long), stream>>>(W.keys, padded, 2u);
    (*launches)++;
    for (uint32_t k = 2 * kBitonicTile; k <= padded; k <<= 1) {
        for (uint32_t j = k >> 1; j >= kBitonicTile; j >>= 1) {
            k_lbvh_bitonic<<<(padded + tb - 1) / tb, tb, 0, stream>>>(W.keys, padded, k, j);
            (*launches)++;
        }
        k_lbvh_bitonic_tile<<<tiles, 1024, kBitonicTile * sizeof(unsigned long long), stream>>>(W.keys, padded, k);
        (*launches)++;
    }
    if (ntris > 1u) {
        k_lbvh_link<<<(ntris - 1u + tb - 1) / tb, tb, 0, stream>>>(W.keys, ntris, W.left, W.right, W.parent);
        (*launches)++;
    }
    k_lbvh_refit<<<(ntris + tb - 1) / tb, tb, 0, stream>>>(W.keys, ntris, W.left, W.right, W.parent, W.tri_boxes, W.arrivals, d_nodes, d_order);
    (*launches)++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// device terrain: packed cells + min-max levels
// ------------------------------------------------------------------------------------------------
static_assert(kHostMaxLevels == kMaxLevels, "f3d_host.h and f3d_trace.cuh disagree on the level count");

// Uploads the DEM, scans it for non-finite samples, and builds cells + pyramid on the device.
int build_device_terrain(const float* h_heights, uint32_t w, uint32_t h, float ex, cudaStream_t stream, DeviceTerrain* T,
                         uint64_t* launches, bool keep_plain) {
    const uint32_t cw = w - 1, ch = h - 1;
    uint32_t lw = next_pow2(cw), lh = next_pow2(ch);
    T->cell_w = cw; T->cell_h = ch;
    T->nlevels = 0; T->mm_total = 0;
    while (true) {
        if (T->nlevels >= kMaxLevels) return fail(F3D_ERR_UPLOAD, "min-max pyramid exceeds %d levels", kMaxLevels);
        T->dims[T->nlevels][0] = lw; T->dims[T->nlevels][1] = lh;
        T->level_off[T->nlevels] = T->mm_total;
        T->mm_total += (size_t)lw * lh;
        T->nlevels++;
        if (lw == 1 && lh == 1) break;
        lw = std::max(lw / 2, 1u); lh = std::max(lh / 2, 1u);
    }
    float* d_h = nullptr;
    uint32_t* d_flag = nullptr;
    const size_t n = (size_t)w * h;
    CUDA_TRY(cudaGetDevice(&T->device));
    struct Scratch {   // upload staging + finite flag: released on every return path, after the stream drained
        float*& h; uint32_t*& f; cudaStream_t st; int dev;
        ~Scratch() { if (h || f) cudaStreamSynchronize(st); cached_free(h, dev); cached_free(f, dev); }
    } scratch{d_h, d_flag, stream, T->device};
    // `h_heights` may already live on this device (a DEM broadcast over NVLink by the multi-GPU launcher): read it in place
    const float* src_dev = nullptr;
    {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, h_heights) == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged))
            src_dev = h_heights;
        cudaGetLastError();
    }
    if (!src_dev) CUDA_TRY(cached_malloc((void**)&d_h, n * sizeof(float), T->device));
    CUDA_TRY(cached_malloc((void**)&d_flag, sizeof(uint32_t), T->device));
    CUDA_TRY(cached_malloc((void**)&T->cells, (size_t)cw * ch * sizeof(float4), T->device));
    CUDA_TRY(cached_malloc((void**)&T->mm_base, T->mm_total * sizeof(float2), T->device));
    CUDA_TRY(cudaMemsetAsync(d_flag, 0, sizeof(uint32_t), stream));
    if (!src_dev) CUDA_TRY(cudaMemcpyAsync(d_h, h_heights, n * sizeof(float), cudaMemcpyHostToDevice, stream));
    const float* heights_dev = src_dev ? src_dev : d_h;
    k_check_finite<<<(unsigned)std::min<size_t>((n + 255) / 256, 1184), 256, 0, stream>>>(heights_dev, n, d_flag);
    dim3 blk(32, 8);
    dim3 g0((T->dims[0][0] + 31) / 32, (T->dims[0][1] + 7) / 8);
    k_build_level0<<<g0, blk, 0, stream>>>(heights_dev, w, h, T->dims[0][0], T->dims[0][1], ex, T->cells, T->mm_base);
    *launches += 2;
    for (int l = 1; l < T->nlevels; l++) {
        dim3 g((T->dims[l][0] + 31) / 32, (T->dims[l][1] + 7) / 8);
        k_reduce_level<<<g, blk, 0, stream>>>(T->mm_base + T->level_off[l - 1], T->dims[l - 1][0], T->dims[l - 1][1],
                                              T->mm_base + T->level_off[l], T->dims[l][0], T->dims[l][1]);
        (*launches)++;
    }
    // quad-packed levels: level l (< root) grouped by its level l+1 parent
    T->quad_total = 0;
    for (int l = 0; l + 1 < T->nlevels; l++) {
        T->quad_pitch[l] = T->dims[l + 1][0];
        T->quad_ph[l] = T->dims[l + 1][1];
        T->quad_off[l] = T->quad_total;
        T->quad_total += (size_t)T->quad_pitch[l] * T->quad_ph[l] * 4;
    }
    CUDA_TRY(cached_malloc((void**)&T->quad_base, std::max<size_t>(T->quad_total, 1) * sizeof(float2), T->device));
    for (int l = 0; l + 1 < T->nlevels; l++) {
        dim3 g((2 * T->quad_pitch[l] + 31) / 32, (2 * T->quad_ph[l] + 7) / 8);
        k_pack_quads<<<g, blk, 0, stream>>>(T->mm_base + T->level_off[l], T->dims[l][0], T->dims[l][1],
                                            T->quad_base + T->quad_off[l], T->quad_pitch[l], T->quad_ph[l]);
        (*launches)++;
    }
    uint32_t flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, d_flag, sizeof flag, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(&T->root_mm, T->mm_base + T->level_off[T->nlevels - 1], sizeof(float2), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    CUDA_TRY(cudaGetLastError());
    if (flag) return fail(F3D_ERR_UPLOAD, "terrain heightfield contains non-finite samples");
    if (!keep_plain) T->release_plain();
    T->bytes = (uint64_t)cw * ch * sizeof(float4) + T->quad_total * sizeof(float2) +
               (keep_plain ? T->mm_total * sizeof(float2) : 0);
    return 0;
}

static void fill_scene_terrain(SceneParams* S, const DeviceTerrain& T) {
    S->cells = T.cells;
    S->cell_w = T.cell_w; S->cell_h = T.cell_h;
    S->mip_count = (uint32_t)T.nlevels;
    for (int l = 0; l < kMaxLevels; l++) {
        S->mm[l] = (l < T.nlevels && T.mm_base) ? T.mm_base + T.level_off[l] : nullptr;
        S->mm_pitch[l] = l < T.nlevels ? T.dims[l][0] : 0;
    }
}

static void fill_fast_scene(FastScene* F, const SceneParams& S, const DeviceTerrain& T) {
    F->ox = S.ox; F->oz = S.oz; F->sx = S.sx; F->sz = S.sz;
    F->cell_w = T.cell_w; F->cell_h = T.cell_h; F->mip_count = (uint32_t)T.nlevels;
    F->cells = T.cells;
    for (int l = 0; l < kMaxLevels; l++) {
        F->q.lv[l] = l + 1 < T.nlevels ? T.quad_base + T.quad_off[l] : nullptr;
        F->q.parent_pitch[l] = l + 1 < T.nlevels ? T.quad_pitch[l] : 0;
    }
    F->root_mm = T.root_mm;
    F->inv_two_r_prime = S.inv_two_r_prime;
    fast_scene_finish(*F);
}

constexpr uint32_t kLbvhMinTris = 4096u;

static uint32_t stack_depth_for(int nlevels) { return 3u * (uint32_t)nlevels + 2u; }

template <typename K>
static int allow_smem(K kernel, size_t bytes) {
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// session
// ------------------------------------------------------------------------------------------------
struct f3d_session {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    FrameParams P{};
    DeviceTerrain terrain;
    float4* d_env = nullptr;
    float4* d_mesh_v = nullptr;
    uint32_t* d_mesh_i = nullptr;
    float4* d_bvh_nodes = nullptr;
    uint32_t* d_bvh_tris = nullptr;
    float4* d_accum = nullptr;
    float2* d_welford = nullptr;
    float4* d_resv[2] = {nullptr, nullptr};
    uint8_t* d_pixflags = nullptr;
    ushort4* d_aov_normal = nullptr;
    float* d_aov_depth = nullptr;
    unsigned long long* d_counters = nullptr;
    uint32_t* d_gate = nullptr;       // [0] vmax bits, [1] non-finite, [2] validity non-finite, [3] validity any
    uint32_t* h_gate = nullptr;       // pinned
    // output staging (host-facing resolve)
    uint8_t* d_rgba = nullptr; float* d_albedo = nullptr; float* d_normal = nullptr; float* d_depth = nullptr;
    // peers
    void* peer_ptrs[8 * F3D_IPC_HANDLES_PER_RANK] = {};
    int n_peer_ptrs = 0;
    uint32_t* d_sync = nullptr;     // [0] CTA counter, [1]/[2] neighbours' completed frames, [3] timeout flag (IPC-shared)
    // bookkeeping
    uint32_t frames = 0;
    uint32_t max_frames = 0, min_frames = 0;
    float variance_threshold = 0;
    float sun_el_deg = 0, sun_intensity = 0, sun_color[3] = {0, 0, 0};
    uint64_t gpu_bytes = 0, host_visible_bytes = 0, pyramid_bytes_ref = 0, launches = 0;
    double setup_ms = 0, frames_ms = 0, readback_ms = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    dim3 grid;
    size_t smem_bytes = 0;          // stack smem of the per-pixel kernels (kThreads)
    size_t trace_smem_bytes = 0;    // stack smem of k_trace (kTraceCtaThreads)
    float* d_hz = nullptr;          // sun horizon strips (SunHorizon::S)
    float* d_esc = nullptr;         // escape map (EscapeMap::E)
    size_t ptrace_smem_bytes = 0;   // k_ptrace: stacks + staged top levels (F3D_TMA_STAGE)
    size_t ascent_smem_bytes = 0;   // k_ascent: leaf rings of the near-field walk (+ staged top levels)
    int trace_grid = 0;             // persistent CTAs of k_trace
    int ascent_grid = 0;            // grid-stride CTAs of k_ascent
    bool ramp = true;               // half-size first batch of every render_frames call (F3D_B200_RAMP)
    cudaStream_t copy_stream = nullptr;   // one-call path: AOV decode + D2H beside the frame loop (early_aov_readback)
    bool aovs_early = false;              // the AOVs of this session are already on their way to the caller's arrays
    uint64_t aovs_early_bytes = 0;
    cudaStream_t prim_stream = nullptr;   // optional high-priority stream of the primary pass (F3D_B200_PRIM_PRIORITY=1)
    // Frame batching + pipelining.  k_primary(step+1) only depends on k_primary(step) (reservoir records); everything after it
    // (k_ascent, k_trace, k_accum) depends on k_primary of the same step and on k_accum of the step before.  Steps are
    // processed in BATCHES of up to `batch` steps: their primaries run back to back on the session stream, then ONE launch of
    // k_ascent / k_trace / k_accum serves the whole batch on the batch set's own stream, so the primaries of the next batch
    // overlap with it.  Two batch sets alternate.  On an image partition this is what keeps 148 SMs busy: a single step of a
    // 1/8 1080p frame is ~230 k rays.
    struct BatchSet {
        BatchSlot slots[kMaxBatch] = {};
        cudaEvent_t primary_done = nullptr, accum_done = nullptr;
        cudaStream_t stream = nullptr;     // k_ascent / k_trace / k_accum of the batches that use this set
        bool used = false;
    };
    static constexpr int kMaxSets = 2;
    BatchSet sets[kMaxSets];
    int n_sets = 1, batch = 1;
    bool split_primary = false;
    uint64_t steps = 0, batches = 0;
    cudaEvent_t join_ev = nullptr;
    float4* d_sstate = nullptr;
    // AETHER post (desc.atmosphere): payloads are copied at creation (the desc is borrowed for that call only) and
    // validated + uploaded by the first resolve, i.e. after the frame loop, as the reference does
    // (render_terrain.rs:1246-1271: "after PROMETHEUS has finished its frame-0 AOV writes and convergence loop")
    bool has_atmosphere = false, aether_ready = false;
    f3d_atmosphere atm{};
    std::vector<uint16_t> h_lut[3];
    uint2* d_lut[3] = {nullptr, nullptr, nullptr};
    AetherParams aether{};
};

static void session_free(f3d_session* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    for (int i = 0; i < s->n_peer_ptrs; i++)
        if (s->peer_ptrs[i]) cudaIpcCloseMemHandle(s->peer_ptrs[i]);
    // nothing may still touch buffers that get parked (render_frames joins the slot streams into the session
    // stream, but an error return in the middle of it does not)
    for (auto& bs : s->sets)
        if (bs.stream) cudaStreamSynchronize(bs.stream);
    if (s->stream) cudaStreamSynchronize(s->stream);
    cached_free(s->d_sync, s->device, false);
    const int dv = s->device;
    const bool ipc = s->P.part_world > 1u && s->P.part_mode == 0u;     // resv images may be mapped by peers: never park them
    s->terrain.release();
    cached_free(s->d_env, dv); cached_free(s->d_mesh_v, dv); cached_free(s->d_mesh_i, dv); cached_free(s->d_hz, dv); cached_free(s->d_esc, dv);
    cached_free(s->d_bvh_nodes, dv); cached_free(s->d_bvh_tris, dv);
    cached_free(s->d_accum, dv); cached_free(s->d_welford, dv);
    cached_free(s->d_resv[0], dv, !ipc); cached_free(s->d_resv[1], dv, !ipc);
    cached_free(s->d_pixflags, dv); cached_free(s->d_aov_normal, dv); cached_free(s->d_aov_depth, dv);
    cached_free(s->d_counters, dv); cached_free(s->d_gate, dv);
    for (auto& bs : s->sets) {
        if (bs.stream) { cudaStreamSynchronize(bs.stream); cudaStreamDestroy(bs.stream); }
        for (auto& sl : bs.slots) {
            cached_free(sl.prim, dv); cached_free(sl.rec, dv); cached_free(sl.occl_sun, dv); cached_free(sl.occl_ibl, dv);
            cached_free(sl.q_sun, dv); cached_free(sl.q_ibl, dv); cached_free(sl.q_counts, dv); cached_free(sl.qf_sun, dv); cached_free(sl.qf_ibl, dv);
            cached_free(sl.qn_sun, dv); cached_free(sl.qn_ibl, dv); cached_free(sl.q2_sun, dv); cached_free(sl.q2_ibl, dv);
        }
        if (bs.primary_done) cudaEventDestroy(bs.primary_done);
        if (bs.accum_done) cudaEventDestroy(bs.accum_done);
    }
    if (s->join_ev) cudaEventDestroy(s->join_ev);
    cached_free(s->d_sstate, dv);
    for (auto& p : s->d_lut) cached_free(p, dv);
    cached_free(s->d_rgba, dv); cached_free(s->d_albedo, dv); cached_free(s->d_normal, dv); cached_free(s->d_depth, dv);
    if (s->h_gate) cudaFreeHost(s->h_gate);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->prim_stream) { cudaStreamSynchronize(s->prim_stream); cudaStreamDestroy(s->prim_stream); }
    if (s->copy_stream) { cudaStreamSynchronize(s->copy_stream); cudaStreamDestroy(s->copy_stream); }
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

template <typename T>
static int dmalloc(f3d_session* s, T** p, size_t count, bool zero) {
    CUDA_TRY(cached_malloc((void**)p, count * sizeof(T), s->device));
    s->gpu_bytes += count * sizeof(T);
    if (zero) CUDA_TRY(cudaMemsetAsync(*p, 0, count * sizeof(T), s->stream));
    return 0;
}

// Sun horizon strips (SunHorizon, f3d_trace_fast.cuh): built once, for the direction every sun ray of the session shares.
// Not built (hz.S stays NULL, the tracer then behaves as before) for a sun below the horizon, a (near-)vertical sun, a
// negative curvature term, or when F3D_B200_SUN_HORIZON=0.
static int build_sun_horizon(f3d_session* s) {
    FrameParams& P = s->P;
    P.hz = SunHorizon{};
#if F3D_SUN_HORIZON
    if (const char* e = getenv("F3D_B200_SUN_HORIZON")) if (atoi(e) == 0) return 0;
    const double lx = P.light_dir[0], ly = P.light_dir[1], lz = P.light_dir[2];
    const double n = std::sqrt(lx * lx + ly * ly + lz * lz);
    if (!(n > 0.0) || ly < 0.0) return 0;
    if (P.scene.curvature_enabled && !(P.fast.inv_two_r_prime >= 0.0f)) return 0;
    const double du = lx / n / (double)P.fast.sx, dv = lz / n / (double)P.fast.sz, dy = ly / n;
    const bool xmajor = std::fabs(du) >= std::fabs(dv);
    const double da = xmajor ? du : dv, db = xmajor ? dv : du;
    if (!(std::fabs(da) > 0.0)) return 0;
    const double m = db / da, g = dy / std::fabs(da);
    if (!(g < 1.0e6)) return 0;                                   // the sun is (almost) overhead: every ray leaves its column upwards
    const uint32_t ncols = xmajor ? P.fast.cell_w : P.fast.cell_h, nrows = xmajor ? P.fast.cell_h : P.fast.cell_w;
    const double w_min = std::min(0.0, -m * (double)ncols), w_max = (double)nrows + std::max(0.0, -m * (double)ncols);
    const int32_t j0 = (int32_t)std::floor(w_min) - 1;
    const uint32_t nstrips = (uint32_t)((int32_t)std::floor(w_max) + 2 - j0);
    const size_t bytes = (size_t)ncols * nstrips * sizeof(float);
    if (bytes > ((size_t)1 << 30)) return 0;
    CUDA_TRY(cached_malloc((void**)&s->d_hz, bytes, s->device));
    SunHorizon Z{};
    Z.S = s->d_hz; Z.nstrips = nstrips; Z.ncols = ncols; Z.j0 = j0;
    Z.m = (float)m; Z.g = (float)g;
    Z.pad_rel = 1.9073486328125e-6f;                               // 2^-19, see the error budget in DESIGN.md section 6
    Z.mag = P.fast.mag_y + (float)ncols * (float)g;
    Z.xmajor = xmajor ? 1u : 0u; Z.forward = da > 0.0 ? 1u : 0u;
    k_hz_build<<<dim3((nstrips + 255u) / 256u, ncols), 256, 0, s->stream>>>(P.fast.cells, P.fast.cell_w, P.fast.cell_h, Z, m, g, s->d_hz);
    k_hz_suffix<<<(nstrips + 7u) / 8u, 256, 0, s->stream>>>(s->d_hz, nstrips, ncols);          // one warp per strip
    CUDA_TRY(cudaGetLastError());
    s->launches += 2;
    s->gpu_bytes += bytes;
    P.hz = Z;
    if (getenv("F3D_B200_DEBUG"))
        fprintf(stderr, "[forge3d_b200] sun horizon: %u strips x %u columns (%.1f MB), %s-major, m %.4f, g %.3f per column\n", nstrips, ncols,
                bytes / 1048576.0, xmajor ? "x" : "z", m, g);
#endif
    return 0;
}

// Escape map (EscapeMap, f3d_trace_fast.cuh): 32 bytes per DEM cell, built once per session from the min-max pyramid.
// Off unless F3D_B200_ESCAPE=1 (see below); without it the IBL rays all take the bottom-up start.
static int build_escape_map(f3d_session* s) {
    FrameParams& P = s->P;
    P.esc = EscapeMap{};
#if F3D_ESCAPE
    // OPT-IN (F3D_B200_ESCAPE=1).  Measured on the B200, C2 (profiles/r02_walks.md): the map takes 4.2 ms to build and saves
    // 0.008 ms per 1080p frame (0.859 -> 0.851 ms: k_trace -15 % warp-instructions, but the classify pass that reads the map
    // is latency-bound), i.e. it pays for itself only beyond ~500 frames of this size.  Bit-exact either way
    // (tests/test_gpu_parity.py::test_escape_map_is_exact_and_culls).
    const char* e = getenv("F3D_B200_ESCAPE");
    if (!e || atoi(e) == 0) return 0;
    const size_t ncells = (size_t)P.fast.cell_w * P.fast.cell_h;
    const size_t bytes = ncells * 8 * sizeof(float);
    if (bytes > ((size_t)4 << 30)) return 0;
    CUDA_TRY(cached_malloc((void**)&s->d_esc, bytes, s->device));
    // height pad: covers the rounding of the reference's ray height at the span start (2 ulp of |o.y| + the gain) and of
    // the ray origin against the cell's corner heights
    const float pad_abs = 1.52587890625e-5f * (P.fast.mag_y + 1.0f);          // 2^-16
    k_escape_build<<<dim3((P.fast.cell_w + 127u) / 128u, P.fast.cell_h), 128, 0, s->stream>>>(P.fast, escape_oct_table(), pad_abs, s->d_esc);
    CUDA_TRY(cudaGetLastError());
    s->launches += 1;
    s->gpu_bytes += bytes;
    P.esc.E = s->d_esc;
    if (getenv("F3D_B200_DEBUG"))
        fprintf(stderr, "[forge3d_b200] escape map: %zu cells x 8 octants (%.1f MB), near block radius %d\n", ncells, bytes / 1048576.0, kEscR);
#endif
    return 0;
}

// NVTX ranges carry the pass labels of the reference's render certificate (terrain_reference.rs:1410-1420:
// hybrid_pt.terrain_gbuffer | terrain | restir_temporal | restir_spatial | aether_aerial), so an nsys / ncu timeline of
// this backend lines up with the reference's own timing table.  Temporal + spatial reuse are fused into the shading kernel.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

static int session_create_impl(const f3d_terrain_desc* d, void* cuda_stream, f3d_session* s) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(F3D_ERR_DEVICE, "no CUDA device available: forge3d_b200 has no CPU fallback");
    }
    if (d->device < 0 || d->device >= ndev) return fail(F3D_ERR_DEVICE, "CUDA device %d out of range (%d devices)", d->device, ndev);
    s->device = d->device;
    CUDA_TRY(cudaSetDevice(s->device));
    if (cuda_stream) s->stream = (cudaStream_t)cuda_stream;
    else { CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)); s->own_stream = true; }
    CUDA_TRY(cudaEventCreate(&s->ev0));
    CUDA_TRY(cudaEventCreate(&s->ev1));
    CUDA_TRY(cudaEventRecord(s->ev0, s->stream));

    const uint32_t W = d->width, H = d->height;
    const size_t npx = (size_t)W * H;
    FrameParams& P = s->P;
    SceneParams& S = P.scene;

    // ---- clamp radiometric scalars, camera basis, light (render_terrain.rs:571-576,635-661,697-708) ----
    const float exposure = clamp_radiometric(d->exposure);
    const float sun_intensity = clamp_radiometric(d->sun_intensity);
    const float sun_color[3] = {clamp_radiometric(d->sun_color[0]), clamp_radiometric(d->sun_color[1]),
                                clamp_radiometric(d->sun_color[2])};
    const float env_intensity = clamp_radiometric(d->env_intensity);
    hv3 origin = HV(d->cam_origin);
    hv3 forward = hnorm(hsub(HV(d->cam_look_at), origin));
    hv3 right = hnorm(hcross(forward, HV(d->cam_up)));
    hv3 up = hnorm(hcross(right, forward));
    const float az = to_radians_f32(d->sun_az_deg), el = to_radians_f32(d->sun_el_deg);
    P.W = W; P.H = H; P.frame_index = 0; P.spp = std::max(d->spp, 1u); P.window = 32u;  // WELFORD_WINDOW :236
    memcpy(P.cam_origin, d->cam_origin, sizeof P.cam_origin);
    P.cam_right[0] = right.x; P.cam_right[1] = right.y; P.cam_right[2] = right.z;
    P.cam_up[0] = up.x; P.cam_up[1] = up.y; P.cam_up[2] = up.z;
    P.cam_forward[0] = forward.x; P.cam_forward[1] = forward.y; P.cam_forward[2] = forward.z;
    const float fov = to_radians_f32(d->fov_y_deg);
    P.half_h = tanf(0.5f * fov);                              // hybrid_terrain_traversal.wgsl:469
    P.half_w = ((float)W / (float)H) * P.half_h;              // cam_aspect * half_h
    P.exposure = exposure;
    P.seed_hi = d->seed;
    P.seed_lo = d->seed ^ 0x85EBCA6Bu;                        // render_terrain.rs:658-659
    P.light_dir[0] = cosf(az) * cosf(el); P.light_dir[1] = sinf(el); P.light_dir[2] = sinf(az) * cosf(el);
    for (int c = 0; c < 3; c++) P.light_color[c] = sun_intensity * sun_color[c];
    s->sun_el_deg = d->sun_el_deg; s->sun_intensity = sun_intensity;
    memcpy(s->sun_color, sun_color, sizeof sun_color);
    s->max_frames = d->max_frames; s->min_frames = d->min_frames; s->variance_threshold = d->variance_threshold;

    // ---- partition ----
    P.part_world = std::max(d->part_world, 1u);
    P.part_mode = P.part_world > 1u ? d->part_mode : 0u;
    P.part_rank = P.part_world > 1 ? d->part_rank : 0u;
    uint32_t block_rows = d->part_block_rows ? d->part_block_rows : 16u;
    block_rows = ((block_rows + kTileH - 1) / kTileH) * kTileH;
    if (P.part_world == 1) block_rows = ((H + kTileH - 1) / kTileH) * kTileH;   // one block = whole image
    P.block_rows = block_rows;
    P.tiles_per_block = block_rows / kTileH;
    P.nblocks = (H + block_rows - 1) / block_rows;
    const uint32_t owned_blocks = P.nblocks > P.part_rank ? (P.nblocks - P.part_rank + P.part_world - 1) / P.part_world : 0;
    s->grid = dim3((W + kTileW - 1) / kTileW, std::max(owned_blocks * P.tiles_per_block, 1u));

    // ---- terrain scene ----
    int rc = earth_curvature(d, &S.inv_two_r_prime, &S.curvature_enabled);
    if (rc) return rc;
    rc = build_device_terrain(d->heights, d->dem_w, d->dem_h, d->exaggeration, s->stream, &s->terrain, &s->launches, false);
    if (rc) return rc;
    s->gpu_bytes += s->terrain.bytes;
    fill_scene_terrain(&S, s->terrain);
    S.sx = d->spacing[0]; S.sz = d->spacing[1];
    S.ox = -0.5f * ((float)d->dem_w - 1.0f) * S.sx;           // terrain_heightfield.rs:359-360
    S.oz = -0.5f * ((float)d->dem_h - 1.0f) * S.sz;
    fill_fast_scene(&P.fast, S, s->terrain);
    P.stack_depth = stack_depth_for(s->terrain.nlevels);
    if ((rc = build_sun_horizon(s))) return rc;
    if ((rc = build_escape_map(s))) return rc;
    s->smem_bytes = stack_smem_bytes(P.stack_depth, kThreads);
    s->trace_smem_bytes = trace_smem_bytes_for(P.stack_depth, kTraceCtaThreads);
    s->ptrace_smem_bytes = s->smem_bytes;
    s->ascent_smem_bytes = kAscentRingBytes;
#if F3D_TMA_STAGE
    {   // TMA staging of the top pyramid levels (f3d_trace_fast.cuh): per kernel, as many whole levels from the top as fit the
        // kernel's shared-memory budget (k_ptrace 4 CTAs/SM beside 39 KB of stacks, k_ascent 3 CTAs/SM with nothing else,
        // k_trace 6 CTAs/SM beside 31 KB of stacks + leaf rings).  F3D_B200_TMA_STAGE=0 keeps the level table but stages nothing.
        const DeviceTerrain& T = s->terrain;
        const bool on = !(getenv("F3D_B200_TMA_STAGE") && atoi(getenv("F3D_B200_TMA_STAGE")) == 0);
        const uint32_t budget[3] = {12u << 10, 44u << 10, 3u << 10};
        for (int k = 0; k < 3; k++) {
            StageParams sp{};
            sp.first = (uint32_t)std::max(T.nlevels - 1, 0);
            sp.bytes = 0u;
            sp.src = T.quad_base;
            for (int l = T.nlevels - 2; l >= 0 && on; l--) {        // levels l .. nlevels-2 are the END of the quad arena
                const size_t bytes = (T.quad_total - T.quad_off[l]) * sizeof(float2);
                if (bytes > budget[k]) break;
                sp.first = (uint32_t)l; sp.bytes = (uint32_t)bytes; sp.src = T.quad_base + T.quad_off[l];
            }
            P.stage[k] = sp;
        }
        s->ptrace_smem_bytes = ((s->smem_bytes + 15) & ~(size_t)15) + stage_smem_bytes(P.stage[0].bytes);
        s->ascent_smem_bytes = kAscentRingBytes + stage_smem_bytes(P.stage[1].bytes);
        s->trace_smem_bytes = ((s->trace_smem_bytes + 15) & ~(size_t)15) + stage_smem_bytes(P.stage[2].bytes);
#define F3D_ALLOW_ASCENT(...) if ((rc = allow_smem(k_ascent<__VA_ARGS__>, s->ascent_smem_bytes))) return rc
        F3D_ALLOW_ASCENT(true, true, 1, 0); F3D_ALLOW_ASCENT(true, true, 1, 1); F3D_ALLOW_ASCENT(true, true, 1, 2);
        F3D_ALLOW_ASCENT(true, false, 1, 0); F3D_ALLOW_ASCENT(true, false, 1, 1); F3D_ALLOW_ASCENT(true, false, 1, 2);
        F3D_ALLOW_ASCENT(true, false, 0, 2);
        F3D_ALLOW_ASCENT(false, false, 0, 0); F3D_ALLOW_ASCENT(false, false, 0, 1); F3D_ALLOW_ASCENT(false, false, 0, 2);
#undef F3D_ALLOW_ASCENT
        if (getenv("F3D_B200_DEBUG"))
            fprintf(stderr, "[forge3d_b200] TMA staging: k_ptrace levels >= %u (%u B), k_ascent >= %u (%u B), k_trace >= %u (%u B)\n",
                    P.stage[0].first, P.stage[0].bytes, P.stage[1].first, P.stage[1].bytes, P.stage[2].first, P.stage[2].bytes);
    }
#endif
    if ((rc = allow_smem(k_primary, s->smem_bytes))) return rc;
    if ((rc = allow_smem(k_ptrace, s->ptrace_smem_bytes))) return rc;
    if ((rc = allow_smem(k_gbuffer, s->smem_bytes))) return rc;
    if ((rc = allow_smem(k_trace<true, 1>, s->trace_smem_bytes))) return rc;
    if ((rc = allow_smem(k_trace<true, 2>, s->trace_smem_bytes))) return rc;
    if ((rc = allow_smem(k_trace<false, 1>, s->trace_smem_bytes))) return rc;
    if ((rc = allow_smem(k_trace<false, 0>, s->trace_smem_bytes))) return rc;
    {   // persistent grid: every SM filled to the occupancy the traversal kernel reaches
        int per_sm = 0, sms = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace<true, 1>, kTraceCtaThreads, s->trace_smem_bytes));
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
        // Fewer persistent CTAs than the occupancy limit leave registers for the kernels of the other batch set (primaries of
        // the next batch) to run beside k_trace: measured 0.943 -> 0.933 ms/frame at 4 per SM on the full frame, 0.203 ->
        // 0.178 ms at 3 per SM on a 1/8 partition (profiles/README.md).
        per_sm = std::min(per_sm, P.part_world > 2u ? 3 : 4);
        if (const char* e = getenv("F3D_B200_TRACE_CTAS")) per_sm = std::max(atoi(e), 1);
        s->trace_grid = std::max(per_sm, 1) * std::max(sms, 1);
        s->ascent_grid = 8 * std::max(sms, 1);
        if (const char* e = getenv("F3D_B200_ASCENT_CTAS")) s->ascent_grid = std::max(atoi(e), 1) * std::max(sms, 1);   // grid-stride CTAs per SM (tuning)
        if (getenv("F3D_B200_DEBUG"))
            fprintf(stderr, "[forge3d_b200] k_trace: %d CTAs/SM x %d SMs, %zu B smem/CTA, stack depth %u\n", per_sm, sms,
                    s->trace_smem_bytes, P.stack_depth);
    }
    memcpy(S.albedo, d->albedo, sizeof S.albedo);
    S.env_intensity = env_intensity;
    {   // reference-compatible diagnostic: DEM R32F + RG32F chain (terrain_heightfield.rs:292-314)
        uint64_t b = (uint64_t)d->dem_w * d->dem_h * 4;
        for (int l = 0; l < s->terrain.nlevels; l++) b += (uint64_t)s->terrain.dims[l][0] * s->terrain.dims[l][1] * 8;
        s->pyramid_bytes_ref = b;
    }
    if (d->env_rgb) {
        const size_t ne = (size_t)d->env_w * d->env_h;
        std::vector<float4> rgba(ne);
        for (size_t i = 0; i < ne; i++) rgba[i] = make_float4(d->env_rgb[3 * i], d->env_rgb[3 * i + 1], d->env_rgb[3 * i + 2], 1.0f);
        rc = dmalloc(s, &s->d_env, ne, false);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(s->d_env, rgba.data(), ne * sizeof(float4), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        S.env = s->d_env; S.env_w = d->env_w; S.env_h = d->env_h;
    }
    S.traversal_mode = 3u;
    if (d->mesh_xyz) {
        std::vector<float4> v(d->mesh_nverts);
        for (uint32_t i = 0; i < d->mesh_nverts; i++) v[i] = make_float4(d->mesh_xyz[3 * i], d->mesh_xyz[3 * i + 1], d->mesh_xyz[3 * i + 2], 0.0f);
        rc = dmalloc(s, &s->d_mesh_v, (size_t)d->mesh_nverts, false);
        if (rc) return rc;
        rc = dmalloc(s, &s->d_mesh_i, (size_t)d->mesh_ntris * 3, false);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(s->d_mesh_v, v.data(), v.size() * sizeof(float4), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaMemcpyAsync(s->d_mesh_i, d->mesh_idx, (size_t)d->mesh_ntris * 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        S.mesh_v = s->d_mesh_v; S.mesh_i = s->d_mesh_i;
        S.mesh_index_count = d->mesh_ntris * 3u; S.mesh_nverts = d->mesh_nverts;
        // BVH choice: meshes above kLbvhMinTris are built on the device (LBVH, microseconds per 100 k triangles); small ones keep
        // the host median-split tree (shallower, 4 triangles per leaf).  F3D_B200_MESH_BVH=host|lbvh|none overrides.  Either
        // way the closest hit is the index-order sweep's (tests/test_gpu_parity.py::test_mesh_bvh_matches_index_order_sweep).
        const char* bvh_env = getenv("F3D_B200_MESH_BVH");
        const bool no_bvh = getenv("F3D_B200_NO_MESH_BVH") || (bvh_env && !strcmp(bvh_env, "none"));
        const bool use_lbvh = bvh_env ? !strcmp(bvh_env, "lbvh") : d->mesh_ntris >= kLbvhMinTris;
        if (d->mesh_ntris > 8u && !no_bvh && use_lbvh) {
            if ((rc = dmalloc(s, &s->d_bvh_nodes, (size_t)2 * (2 * (size_t)d->mesh_ntris - 1), false))) return rc;
            if ((rc = dmalloc(s, &s->d_bvh_tris, (size_t)d->mesh_ntris, false))) return rc;
            if ((rc = build_mesh_lbvh(d->mesh_xyz, d->mesh_nverts, d->mesh_idx, d->mesh_ntris, s->d_mesh_v, s->d_mesh_i, s->device, s->stream,
                                      s->d_bvh_nodes, s->d_bvh_tris, &s->launches)))
                return rc;
            S.bvh_nodes = s->d_bvh_nodes; S.bvh_tris = s->d_bvh_tris;
        } else if (d->mesh_ntris > 8u && !no_bvh) {
            MeshBvh bvh;
            build_mesh_bvh(d->mesh_xyz, d->mesh_idx, d->mesh_ntris, &bvh);
            if ((rc = dmalloc(s, &s->d_bvh_nodes, bvh.nodes.size(), false))) return rc;
            if ((rc = dmalloc(s, &s->d_bvh_tris, bvh.tris.size(), false))) return rc;
            CUDA_TRY(cudaMemcpyAsync(s->d_bvh_nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(float4), cudaMemcpyHostToDevice, s->stream));
            CUDA_TRY(cudaMemcpyAsync(s->d_bvh_tris, bvh.tris.data(), bvh.tris.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            S.bvh_nodes = s->d_bvh_nodes; S.bvh_tris = s->d_bvh_tris;
        }
        S.traversal_mode = 0u;                                   // TraversalMode::Hybrid, render_terrain.rs:681-685
    }

    // ---- the reference's 512 MiB working-set gate, opt-in (render_terrain.rs:785-888) ----
    if (d->compat_512mib_gate) {
        uint64_t total = (uint64_t)npx * (16 + 8 + 80 * 3 + 16 * 2 + 8 + 48) + s->pyramid_bytes_ref +
                         (d->env_rgb ? (uint64_t)d->env_w * d->env_h * 16 : 16) + (uint64_t)d->mesh_nverts * 16 +
                         (uint64_t)d->mesh_ntris * 12 + 96 + 32 + 80 + 96 + 24 + 32 + 32 + 48;
        const uint64_t limit = 512ull * 1024 * 1024;
        if (total > limit)
            return fail(F3D_ERR_BUDGET,
                        "terrain PT exceeds the memory budget before rendering: tracked total %llu (host-visible %llu) > limit %llu",
                        (unsigned long long)total, 0ull, (unsigned long long)limit);
    }

    // ---- per-pixel state ----
    if ((rc = dmalloc(s, &s->d_accum, npx, true))) return rc;
    if ((rc = dmalloc(s, &s->d_welford, npx, true))) return rc;
    if ((rc = dmalloc(s, &s->d_resv[0], npx, true))) return rc;
    if ((rc = dmalloc(s, &s->d_resv[1], npx, true))) return rc;
    if ((rc = dmalloc(s, &s->d_pixflags, npx, true))) return rc;
    if ((rc = dmalloc(s, &s->d_aov_normal, npx, true))) return rc;
    if ((rc = dmalloc(s, &s->d_aov_depth, npx, true))) return rc;
    if ((rc = dmalloc(s, &s->d_counters, (size_t)4, true))) return rc;
    if ((rc = dmalloc(s, &s->d_gate, (size_t)4, true))) return rc;
    CUDA_TRY(cudaMallocHost(&s->h_gate, 4 * sizeof(uint32_t)));
    s->host_visible_bytes += 4 * sizeof(uint32_t);
    // split primary pass (k_ptrace + k_shade, see f3d_kernels.cuh): terrain-only scenes with one sample per frame
    s->split_primary = P.spp == 1u && S.traversal_mode == 3u && !getenv("F3D_B200_NO_SPLIT");
    {
        // default: two batch sets of 4 steps; smaller batches when the per-step buffers (98 B/pixel) would exceed 8 GB in total.
        // F3D_B200_BATCH / F3D_B200_SETS override (F3D_B200_PIPELINE=1 is the old spelling of "no overlap": 1 set of 1).
        const int by_memory = (int)std::max<uint64_t>(1, (8ull << 30) / std::max<uint64_t>(1, (uint64_t)npx * (106 + (s->split_primary ? 32 : 0))));
        const char* eb = getenv("F3D_B200_BATCH");
        const char* es = getenv("F3D_B200_SETS");
        s->n_sets = es ? std::min(std::max(atoi(es), 1), (int)f3d_session::kMaxSets) : (by_memory >= 2 ? 2 : 1);
        s->batch = eb ? std::min(std::max(atoi(eb), 1), kMaxBatch) : std::min(P.part_world > 2u ? 8 : 4, std::max(by_memory / s->n_sets, 1));   // small partitions: bigger batches
        if (const char* e = getenv("F3D_B200_PIPELINE")) if (atoi(e) <= 1) { s->n_sets = 1; s->batch = 1; }
        if (const char* e = getenv("F3D_B200_RAMP")) s->ramp = atoi(e) != 0;
    }
    for (int k = 0; k < s->n_sets; k++) {
        f3d_session::BatchSet& bs = s->sets[k];
        for (int j = 0; j < s->batch; j++) {
            BatchSlot& sl = bs.slots[j];
            if (s->split_primary && (rc = dmalloc(s, &sl.prim, npx * 2, false))) return rc;
            if ((rc = dmalloc(s, &sl.rec, npx * 4, true))) return rc;
            if ((rc = dmalloc(s, &sl.occl_sun, npx, true))) return rc;
            if ((rc = dmalloc(s, &sl.occl_ibl, npx, true))) return rc;
            if ((rc = dmalloc(s, &sl.q_sun, npx, false))) return rc;
            if ((rc = dmalloc(s, &sl.q_ibl, npx, false))) return rc;
            if ((rc = dmalloc(s, &sl.q_counts, (size_t)kQCounts, true))) return rc;
            if ((rc = dmalloc(s, &sl.qf_sun, npx, false))) return rc;
            if ((rc = dmalloc(s, &sl.qf_ibl, npx, false))) return rc;
            if ((rc = dmalloc(s, &sl.q2_sun, npx, false))) return rc;
            if ((rc = dmalloc(s, &sl.q2_ibl, npx, false))) return rc;
            if ((rc = dmalloc(s, &sl.qn_sun, npx, false))) return rc;
            if ((rc = dmalloc(s, &sl.qn_ibl, npx, false))) return rc;
        }
        CUDA_TRY(cudaEventCreateWithFlags(&bs.primary_done, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&bs.accum_done, cudaEventDisableTiming));
        if (s->n_sets > 1) CUDA_TRY(cudaStreamCreateWithFlags(&bs.stream, cudaStreamNonBlocking));
    }
    if (s->n_sets > 1) CUDA_TRY(cudaEventCreateWithFlags(&s->join_ev, cudaEventDisableTiming));
    if (s->n_sets > 1) {
        // The k_ptrace / k_shade chain is the only frame-to-frame (and, partitioned, GPU-to-GPU) dependency; on its own
        // high-priority stream its CTAs are scheduled ahead of the secondary-ray kernels of the other batch set whenever an SM has
        // room.  Opt-in until measured at N = 8 (F3D_B200_PRIM_PRIORITY=1).
        const char* e = getenv("F3D_B200_PRIM_PRIORITY");
        if (e && atoi(e) != 0) {
            int least = 0, greatest = 0;
            CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            CUDA_TRY(cudaStreamCreateWithPriority(&s->prim_stream, cudaStreamNonBlocking, greatest));
        }
    }
    if (P.spp > 1u && (rc = dmalloc(s, &s->d_sstate, npx * 3, true))) return rc;
    P.sstate = s->d_sstate;
    P.sample_index = 0u;
    P.accum = s->d_accum; P.welford = s->d_welford; P.pixflags = s->d_pixflags; P.counters = s->d_counters;
    P.resv_in = s->d_resv[1]; P.resv_out = s->d_resv[0];
    P.peer_up = nullptr; P.peer_down = nullptr;
    P.sync_local = nullptr; P.peer_sync_up = nullptr; P.peer_sync_down = nullptr; P.sync_error = nullptr;
    if (P.part_world > 1u && P.part_mode == 0u) {
        CUDA_TRY(cached_malloc((void**)&s->d_sync, 8 * sizeof(uint32_t), s->device));      // IPC-shared: freed with allow_park = false
        CUDA_TRY(cudaMemsetAsync(s->d_sync, 0, 8 * sizeof(uint32_t), s->stream));
        P.sync_local = s->d_sync;
        P.sync_error = s->d_sync + 3;
        const char* te = getenv("F3D_B200_SYNC_TIMEOUT_MS");
        P.sync_timeout_ns = (unsigned long long)std::max(te ? atoll(te) : 2000ll, 1ll) * 1000000ull;
    }

    if (d->atmosphere) {
        const f3d_atmosphere& a = *d->atmosphere;
        s->has_atmosphere = true;
        s->atm = a;
        const size_t texels[3] = {(size_t)a.transmittance_mu * a.transmittance_height,
                                  (size_t)a.scattering_mu_view * a.scattering_mu_sun * a.scattering_height * a.scattering_nu,
                                  (size_t)a.aerial_distance * a.aerial_mu_view * a.aerial_height};
        const uint16_t* src[3] = {a.transmittance, a.scattering, a.aerial};
        for (int k = 0; k < 3; k++) {
            if (texels[k] > (64u << 20)) return fail(F3D_ERR_RENDER, "atmosphere LUT dimensions overflow");
            s->h_lut[k].assign(src[k], src[k] + texels[k] * 4);
        }
        s->atm.transmittance = s->atm.scattering = s->atm.aerial = nullptr;
        s->aether.tan_half_fov = tanf(0.5f * fov);            // (0.5 * fov_y_radians).tan(), aether_post.rs:139
        s->aether.aspect = (float)W / (float)H;
        s->aether.sun_intensity = sun_intensity;
    }

    // ---- one-shot G-buffer / centre-ray AOV pass (render_terrain.rs:1091-1121) ----
    GbufferOut G{s->d_pixflags, s->d_aov_normal, s->d_aov_depth};
    NvtxRange nvtx_gbuffer("hybrid_pt.terrain_gbuffer");
    k_gbuffer<<<s->grid, kThreads, s->smem_bytes, s->stream>>>(P, G);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->setup_ms = ms;
    return 0;
}

extern "C" int f3d_session_create(const f3d_terrain_desc* desc, void* cuda_stream, f3d_session** out_session) {
    g_err[0] = 0;
    if (!desc || !out_session) return fail(F3D_ERR_ARGUMENT, "null argument");
    *out_session = nullptr;
    int rc = validate_desc(desc);
    if (rc) return rc;
    f3d_session* s = new f3d_session();
    rc = session_create_impl(desc, cuda_stream, s);
    if (rc) { session_free(s); return rc; }
    *out_session = s;
    return 0;
}

extern "C" void f3d_session_destroy(f3d_session* s) { session_free(s); }

extern "C" int f3d_session_render_frames(f3d_session* s, uint32_t n) {
    if (!s) return fail(F3D_ERR_ARGUMENT, "null session");
    CUDA_TRY(cudaSetDevice(s->device));
    NvtxRange nvtx_frames("hybrid_pt.terrain");
    CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
    const bool pipelined = s->n_sets > 1;
    cudaStream_t ps = (pipelined && s->prim_stream) ? s->prim_stream : s->stream;      // stream of the primary pass
    if (pipelined) {   // the set streams start after whatever is already queued on the session stream
        CUDA_TRY(cudaEventRecord(s->join_ev, s->stream));
        for (int k = 0; k < s->n_sets; k++) CUDA_TRY(cudaStreamWaitEvent(s->sets[k].stream, s->join_ev, 0));
        if (ps != s->stream) CUDA_TRY(cudaStreamWaitEvent(ps, s->join_ev, 0));
    }
    FrameParams& P = s->P;
    const uint32_t spp = P.spp;
    uint64_t todo = (uint64_t)n * spp;            // steps; a call always starts and ends on a frame boundary
    uint32_t sample = 0u;
    bool first_batch = true;
    while (todo > 0) {
        // Ramp: the FIRST batch of a call is half a batch.  Its primaries run alone (nothing to overlap with yet), and the steps it
        // leaves over make the LAST batch, whose secondary rays run alone, half a batch as well: 20 frames = 2 + 4 x 4 + 2.
        // (F3D_B200_RAMP=0: full batches from the start.)
        uint32_t cap = (uint32_t)s->batch;
        if (pipelined && s->ramp && first_batch && s->batch >= 4 && todo > (uint64_t)s->batch) cap = (uint32_t)s->batch / 2u;
        first_batch = false;
        const uint32_t nb = (uint32_t)std::min<uint64_t>(todo, (uint64_t)cap);
        f3d_session::BatchSet& bs = s->sets[s->batches % (uint64_t)s->n_sets];
        f3d_session::BatchSet& prev = s->sets[(s->batches + (uint64_t)s->n_sets - 1u) % (uint64_t)s->n_sets];
        cudaStream_t ts = pipelined ? bs.stream : s->stream;
        if (pipelined && bs.used) CUDA_TRY(cudaStreamWaitEvent(ps, bs.accum_done, 0));   // buffer set free again
        const uint32_t frame0 = s->frames, sample0 = sample;
        // ---- the primaries of the batch, back to back (split path: ONE traversal launch for the batch, then the cheap
        // per-frame shading chain) ----
        if (s->split_primary) {
            P.frame_index = frame0;
            P.sample_index = 0u;
            P.n_batch = nb;
            for (uint32_t k = 0; k < nb; k++) P.slot[k] = bs.slots[k];
            k_ptrace<<<dim3(s->grid.x, s->grid.y, nb), kThreads, s->ptrace_smem_bytes, ps>>>(P);
            s->launches++;
        }
        for (uint32_t k = 0; k < nb; k++) {
            P.frame_index = s->frames;
            P.sample_index = sample;
            P.resv_in = s->d_resv[(s->frames + 1u) & 1u];
            P.resv_out = s->d_resv[s->frames & 1u];
            if (s->n_peer_ptrs) {
                // peer images of the buffer being written this frame (see f3d_session_ipc_import)
                const uint32_t up = (P.part_rank + P.part_world - 1u) % P.part_world, down = (P.part_rank + 1u) % P.part_world;
                P.peer_up = (float4*)s->peer_ptrs[up * F3D_IPC_HANDLES_PER_RANK + (s->frames & 1u)];
                P.peer_down = (float4*)s->peer_ptrs[down * F3D_IPC_HANDLES_PER_RANK + (s->frames & 1u)];
                P.peer_sync_up = (uint32_t*)s->peer_ptrs[up * F3D_IPC_HANDLES_PER_RANK + 2];
                P.peer_sync_down = (uint32_t*)s->peer_ptrs[down * F3D_IPC_HANDLES_PER_RANK + 2];
            }
            P.cur = bs.slots[k];
            P.n_batch = 0u;
            NvtxRange nvtx_shade("hybrid_pt.restir_temporal+restir_spatial");
            if (s->split_primary) k_shade<<<s->grid, kThreads, 0, ps>>>(P);
            else k_primary<<<s->grid, kThreads, s->smem_bytes, ps>>>(P);
            s->launches++;
            s->steps++;
            if (++sample == spp) { sample = 0u; s->frames++; }
        }
        if (pipelined) {
            CUDA_TRY(cudaEventRecord(bs.primary_done, ps));
            CUDA_TRY(cudaStreamWaitEvent(ts, bs.primary_done, 0));
        }
        // ---- one launch of each later kernel for the whole batch ----
        P.frame_index = frame0;
        P.sample_index = sample0;
        P.n_batch = nb;
        for (uint32_t k = 0; k < nb; k++) P.slot[k] = bs.slots[k];
        // sun above the horizon: every sun ray ascends (monotone height tests); curved sun rays that descend keep the
        // round-1 exact expansion (see F3D_CULL_FAST)
        const bool curv = P.scene.curvature_enabled != 0u, asc = P.light_dir[1] >= 0.0f;
#if F3D_TRACE_BOTTOM_UP
        // one launch per list (instruction-cache fit, see k_ascent); SUN_MODE 2 leaves the sun list to the top-down tracer
        // a list with a walk structure (sun horizon strips / escape map) takes two passes: classify + walk, then the far list
        {
            const size_t sm = s->ascent_smem_bytes;
            const int g = s->ascent_grid;
            const bool sun_walk = F3D_SUN_HORIZON && F3D_SUN_NEAR && asc && P.hz.S != nullptr, ibl_walk = F3D_ESCAPE && P.esc.E != nullptr;
            if (curv && asc) {
                if (sun_walk) { k_ascent<true, true, 1, 0><<<g, 256, sm, ts>>>(P); k_ascent<true, true, 1, 1><<<g, 256, sm, ts>>>(P); s->launches += 2; }
                else { k_ascent<true, true, 1, 2><<<g, 256, sm, ts>>>(P); s->launches++; }
            } else if (!curv && asc) {
                if (sun_walk) { k_ascent<true, false, 1, 0><<<g, 256, sm, ts>>>(P); k_ascent<true, false, 1, 1><<<g, 256, sm, ts>>>(P); s->launches += 2; }
                else { k_ascent<true, false, 1, 2><<<g, 256, sm, ts>>>(P); s->launches++; }
            } else if (!curv) { k_ascent<true, false, 0, 2><<<g, 256, sm, ts>>>(P); s->launches++; }
            if (ibl_walk) { k_ascent<false, false, 0, 0><<<g, 256, sm, ts>>>(P); k_ascent<false, false, 0, 1><<<g, 256, sm, ts>>>(P); s->launches += 2; }
            else { k_ascent<false, false, 0, 2><<<g, 256, sm, ts>>>(P); s->launches++; }
        }
        s->launches++;
#endif
        if (curv && asc) k_trace<true, 1><<<s->trace_grid, kTraceCtaThreads, s->trace_smem_bytes, ts>>>(P);
        else if (curv) k_trace<true, 2><<<s->trace_grid, kTraceCtaThreads, s->trace_smem_bytes, ts>>>(P);
        else if (asc) k_trace<false, 1><<<s->trace_grid, kTraceCtaThreads, s->trace_smem_bytes, ts>>>(P);
        else k_trace<false, 0><<<s->trace_grid, kTraceCtaThreads, s->trace_smem_bytes, ts>>>(P);
        if (pipelined && prev.used && &prev != &bs) CUDA_TRY(cudaStreamWaitEvent(ts, prev.accum_done, 0));     // accumulate in frame order
        k_accum<<<s->grid, kThreads, 0, ts>>>(P);
        if (pipelined) { CUDA_TRY(cudaEventRecord(bs.accum_done, ts)); bs.used = true; }
        s->launches += 2;
        s->batches++;
        todo -= nb;
    }
    if (pipelined)   // everything that follows on the session stream (variance, resolve, timing) sees all frames
        for (int k = 0; k < s->n_sets; k++)
            if (s->sets[k].used) CUDA_TRY(cudaStreamWaitEvent(s->stream, s->sets[k].accum_done, 0));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
    return 0;
}

extern "C" int f3d_session_sync(f3d_session* s) {
    if (!s) return fail(F3D_ERR_ARGUMENT, "null session");
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int f3d_session_last_frames_ms(f3d_session* s, double* ms) {
    if (!s || !ms) return fail(F3D_ERR_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaEventSynchronize(s->ev1));
    float f = 0;
    CUDA_TRY(cudaEventElapsedTime(&f, s->ev0, s->ev1));
    *ms = f;
    return 0;
}

extern "C" int f3d_session_frames(const f3d_session* s, uint32_t* frames) {
    if (!s || !frames) return fail(F3D_ERR_ARGUMENT, "null argument");
    *frames = s->frames;
    return 0;
}

// A neighbour wait of the multi-GPU frame barrier timed out at some point: nothing rendered since can be trusted.
static int check_sync_error(f3d_session* s) {
    if (!s->d_sync) return 0;
    uint32_t timed_out = 0;
    CUDA_TRY(cudaMemcpyAsync(&timed_out, s->d_sync + 3, sizeof timed_out, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (timed_out) return fail(F3D_ERR_DEVICE, "multi-GPU frame barrier timed out waiting for a neighbour rank (peer died?)");
    return 0;
}

extern "C" int f3d_session_variance(f3d_session* s, float* vmax, int32_t* nonfinite) {
    if (!s || !vmax || !nonfinite) return fail(F3D_ERR_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(s->device));
    if (s->frames == 0) return fail(F3D_ERR_ARGUMENT, "no frames rendered");
    const uint32_t n_window = ((s->frames - 1u) % 32u) + 1u;   // render_terrain.rs:1208
    CUDA_TRY(cudaMemsetAsync(s->d_gate, 0, 2 * sizeof(uint32_t), s->stream));
    k_variance<<<s->grid, kTileW * kTileH, 0, s->stream>>>(s->P, (float)n_window, s->d_gate);
    s->launches++;
    CUDA_TRY(cudaMemcpyAsync(s->h_gate, s->d_gate, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    memcpy(vmax, &s->h_gate[0], 4);
    *nonfinite = (int32_t)s->h_gate[1];
    return check_sync_error(s);
}

// AetherPostPass::new (aether_post.rs:40-291): validate the settings and LUTs, upload the three tables.
static int prepare_aether(f3d_session* s) {
    if (s->aether_ready) return 0;
    const f3d_atmosphere& a = s->atm;
    int rc = validate_atmosphere(a);
    if (rc) return rc;
    if ((uint64_t)a.scattering_height * a.scattering_nu > 0xFFFFFFFFull) return fail(F3D_ERR_RENDER, "AETHER scattering depth overflow");
    for (int k = 0; k < 3; k++) {
        const size_t texels = s->h_lut[k].size() / 4;
        if ((rc = dmalloc(s, &s->d_lut[k], texels, false))) return rc;
        CUDA_TRY(cudaMemcpyAsync(s->d_lut[k], s->h_lut[k].data(), texels * sizeof(uint2), cudaMemcpyHostToDevice, s->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    AetherParams& A = s->aether;
    A.transmittance = s->d_lut[0]; A.scattering = s->d_lut[1]; A.aerial = s->d_lut[2];
    A.t_dims[0] = a.transmittance_mu; A.t_dims[1] = a.transmittance_height;
    A.s_dims[0] = a.scattering_mu_view; A.s_dims[1] = a.scattering_mu_sun; A.s_dims[2] = a.scattering_height * a.scattering_nu;
    A.s_height = a.scattering_height; A.s_nu = a.scattering_nu;
    A.a_dims[0] = a.aerial_distance; A.a_dims[1] = a.aerial_mu_view; A.a_dims[2] = a.aerial_height;
    A.bottom_radius_m = a.bottom_radius_m; A.top_radius_m = a.top_radius_m; A.max_aerial_distance_m = a.max_aerial_distance_m;
    A.ozone_du = a.ozone_du; A.turbidity = a.turbidity;
    s->aether_ready = true;
    return 0;
}

static int resolve_device_impl(f3d_session* s, void* d_rgba, void* d_albedo, void* d_normal, void* d_depth,
                               int32_t check_validity) {
    if (s->frames == 0) return fail(F3D_ERR_ARGUMENT, "no frames rendered");
    const bool aether = s->has_atmosphere && d_rgba != nullptr;
    if (aether) {
        int rc = prepare_aether(s);
        if (rc) return rc;
    }
    FrameParams P = s->P;
    P.resv_in = s->d_resv[(s->frames + 1u) & 1u];   // out of the last frame
    ResolveOut R{};
    R.rgba = aether ? nullptr : (uint8_t*)d_rgba;   // with AETHER the beauty comes from k_aether below
    R.albedo = (float*)d_albedo; R.normal = (float*)d_normal; R.depth = (float*)d_depth;
    R.aov_normal = s->d_aov_normal; R.aov_depth = s->d_aov_depth;
    R.validity = s->d_gate + 2;
    R.last_frame = s->frames - 1u;
    CUDA_TRY(cudaMemsetAsync(s->d_gate + 2, 0, 2 * sizeof(uint32_t), s->stream));
    if (s->n_peer_ptrs) { k_wait_peers<<<1, 32, 0, s->stream>>>(P, s->frames); s->launches++; }
    k_resolve<<<s->grid, kTileW * kTileH, 0, s->stream>>>(P, R);
    s->launches++;
    if (aether) {   // the post pass of render_terrain.rs:1287-1311, after the traversal's own resolve work
        NvtxRange nvtx_aether("hybrid_pt.aether_aerial");
        k_aether<<<s->grid, kTileW * kTileH, 0, s->stream>>>(P, s->aether, s->d_aov_depth, (uint8_t*)d_rgba);
        s->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    if (!check_validity) { if (int rc2 = check_sync_error(s)) return rc2; }
    if (check_validity) {
        CUDA_TRY(cudaMemcpyAsync(s->h_gate + 2, s->d_gate + 2, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (int rc2 = check_sync_error(s)) return rc2;
        if (s->h_gate[2]) return fail(F3D_ERR_RENDER, "terrain PT reservoir bookkeeping produced non-finite values");
        const bool require = s->sun_el_deg > 0.0f && s->sun_intensity > 0.0f &&
                             (s->sun_color[0] > 0.0f || s->sun_color[1] > 0.0f || s->sun_color[2] > 0.0f);
        // With a row partition a rank may own only sky; the launcher ORs validity across ranks.
        if (require && !s->h_gate[3] && s->P.part_world == 1)
            return fail(F3D_ERR_RENDER,
                        "terrain PT ReSTIR reuse chain produced no valid reservoirs for a sun-lit scene — temporal/spatial reuse is broken");
    }
    return 0;
}

extern "C" int f3d_session_validity(const f3d_session* s, int32_t* any_valid, int32_t* required) {
    if (!s) return fail(F3D_ERR_ARGUMENT, "null session");
    if (any_valid) *any_valid = s->h_gate[3] != 0;
    if (required)
        *required = s->sun_el_deg > 0.0f && s->sun_intensity > 0.0f &&
                    (s->sun_color[0] > 0.0f || s->sun_color[1] > 0.0f || s->sun_color[2] > 0.0f);
    return 0;
}

extern "C" int f3d_session_resolve_device(f3d_session* s, void* d_rgba, void* d_albedo, void* d_normal, void* d_depth,
                                          int32_t check_validity) {
    if (!s) return fail(F3D_ERR_ARGUMENT, "null session");
    CUDA_TRY(cudaSetDevice(s->device));
    return resolve_device_impl(s, d_rgba, d_albedo, d_normal, d_depth, check_validity);
}

static int fill_stats(f3d_session* s, f3d_terrain_out* out) {
    unsigned long long c[4];
    CUDA_TRY(cudaMemcpyAsync(c, s->d_counters, sizeof c, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    out->rays_primary = c[0]; out->rays_shadow = c[1]; out->rays_ibl = c[2]; out->nodes_popped = c[3];
    out->frames = s->frames;
    out->minmax_pyramid_bytes = s->pyramid_bytes_ref;
    out->gpu_resource_bytes = s->gpu_bytes;
    out->peak_host_visible_bytes = s->host_visible_bytes;
    out->setup_ms = s->setup_ms; out->frames_ms = s->frames_ms; out->readback_ms = s->readback_ms;
    out->kernel_launches = s->launches;
    return 0;
}

extern "C" int f3d_session_stats(f3d_session* s, f3d_terrain_out* out) {
    if (!s || !out) return fail(F3D_ERR_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(s->device));
    return fill_stats(s, out);
}

// One-call path: the three AOVs are frame-0 data (G-buffer pass of session creation).  When the caller's AOV arrays are page-locked
// (the Python seam allocates them with f3d_host_alloc) they are decoded and copied on a side stream while the frames render, and
// f3d_session_resolve_host only has the 4 B/pixel beauty image left.  Pageable destinations keep the old path (bounce buffers).
// F3D_B200_EARLY_AOVS=0 switches it off.
static int early_aov_readback(f3d_session* s, f3d_terrain_out* out) {
    if (const char* e = getenv("F3D_B200_EARLY_AOVS")) if (atoi(e) == 0) return 0;
    if (!out->albedo && !out->normal && !out->depth) return 0;
    void* dst[3] = {out->albedo, out->normal, out->depth};
    for (void* p : dst) {
        if (!p) continue;
        cudaPointerAttributes at{};
        const bool pinned = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (!pinned) return 0;
    }
    const size_t npx = (size_t)s->P.W * s->P.H;
    int rc;
    if (out->albedo && !s->d_albedo && (rc = dmalloc(s, &s->d_albedo, npx * 3, false))) return rc;
    if (out->normal && !s->d_normal && (rc = dmalloc(s, &s->d_normal, npx * 3, false))) return rc;
    if (out->depth && !s->d_depth && (rc = dmalloc(s, &s->d_depth, npx, false))) return rc;
    if (!s->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamSynchronize(s->stream));          // the G-buffer pass and the allocations' memsets are done
    ResolveOut R{};
    R.albedo = out->albedo ? s->d_albedo : nullptr; R.normal = out->normal ? s->d_normal : nullptr; R.depth = out->depth ? s->d_depth : nullptr;
    R.aov_normal = s->d_aov_normal; R.aov_depth = s->d_aov_depth;
    k_resolve_aovs<<<s->grid, kTileW * kTileH, 0, s->copy_stream>>>(s->P, R);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    if (out->albedo) CUDA_TRY(cudaMemcpyAsync(out->albedo, s->d_albedo, npx * 12, cudaMemcpyDeviceToHost, s->copy_stream));
    if (out->normal) CUDA_TRY(cudaMemcpyAsync(out->normal, s->d_normal, npx * 12, cudaMemcpyDeviceToHost, s->copy_stream));
    if (out->depth) CUDA_TRY(cudaMemcpyAsync(out->depth, s->d_depth, npx * 4, cudaMemcpyDeviceToHost, s->copy_stream));
    s->aovs_early = true;
    s->aovs_early_bytes = (out->albedo ? npx * 12 : 0) + (out->normal ? npx * 12 : 0) + (out->depth ? npx * 4 : 0);
    return 0;
}

extern "C" int f3d_session_resolve_host(f3d_session* s, f3d_terrain_out* out) {
    if (!s || !out) return fail(F3D_ERR_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t npx = (size_t)s->P.W * s->P.H;
    CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
    int rc;
    if (out->rgba && !s->d_rgba && (rc = dmalloc(s, &s->d_rgba, npx * 4, false))) return rc;
    const bool early = s->aovs_early;                    // the AOVs already travel on the copy stream (early_aov_readback)
    if (!early && out->albedo && !s->d_albedo && (rc = dmalloc(s, &s->d_albedo, npx * 3, false))) return rc;
    if (!early && out->normal && !s->d_normal && (rc = dmalloc(s, &s->d_normal, npx * 3, false))) return rc;
    if (!early && out->depth && !s->d_depth && (rc = dmalloc(s, &s->d_depth, npx, false))) return rc;
    rc = resolve_device_impl(s, out->rgba ? s->d_rgba : nullptr, (!early && out->albedo) ? s->d_albedo : nullptr,
                             (!early && out->normal) ? s->d_normal : nullptr, (!early && out->depth) ? s->d_depth : nullptr, 1);
    if (rc) return rc;
    // Device -> caller memory: page-locked destinations (the Python layer's arrays come from f3d_host_alloc) are written by one
    // DMA each; pageable ones go through the bounce buffers of copy_to_host.
    uint64_t pulled = 0;
    auto pull = [&](void* host, const void* dev, size_t bytes) -> int {
        if (!host) return 0;
        if (int rc2 = copy_to_host(host, dev, bytes, s->stream)) return rc2;
        pulled += bytes;
        return 0;
    };
    if ((rc = pull(out->rgba, s->d_rgba, npx * 4))) return rc;
    if (!early) {
        if ((rc = pull(out->albedo, s->d_albedo, npx * 12))) return rc;
        if ((rc = pull(out->normal, s->d_normal, npx * 12))) return rc;
        if ((rc = pull(out->depth, s->d_depth, npx * 4))) return rc;
    } else {
        pulled += s->aovs_early_bytes;
        CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
        s->aovs_early = false;                           // a later resolve of the same session takes the ordinary path
    }
    if ((rc = pull(out->accum, s->d_accum, npx * 16))) return rc;
    s->host_visible_bytes = std::max<uint64_t>(s->host_visible_bytes, pulled + 16);
    CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->readback_ms = ms;
    return fill_stats(s, out);
}

// ------------------------------------------------------------------------------------------------
// NVLink peer halo exchange
// ------------------------------------------------------------------------------------------------
extern "C" int f3d_session_ipc_export(f3d_session* s, uint8_t* handles) {
    if (!s || !handles) return fail(F3D_ERR_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(s->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == F3D_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, s->d_resv[0]));
    memcpy(handles, &h, sizeof h);
    CUDA_TRY(cudaIpcGetMemHandle(&h, s->d_resv[1]));
    memcpy(handles + F3D_IPC_HANDLE_BYTES, &h, sizeof h);
    if (s->d_sync) {
        CUDA_TRY(cudaIpcGetMemHandle(&h, s->d_sync));
        memcpy(handles + 2 * F3D_IPC_HANDLE_BYTES, &h, sizeof h);
    } else memset(handles + 2 * F3D_IPC_HANDLE_BYTES, 0, F3D_IPC_HANDLE_BYTES);
    return 0;
}

extern "C" int f3d_session_ipc_import(f3d_session* s, const uint8_t* all) {
    if (!s || !all) return fail(F3D_ERR_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(s->device));
    const uint32_t world = s->P.part_world;
    if (world < 2 || s->P.part_mode == 1u) return 0;
    if (world > 8) return fail(F3D_ERR_ARGUMENT, "part_world > 8 not supported");
    for (uint32_t r = 0; r < world; r++) {
        if (r == s->P.part_rank) continue;
        const uint32_t up = (s->P.part_rank + world - 1u) % world, down = (s->P.part_rank + 1u) % world;
        if (r != up && r != down) continue;
        for (int k = 0; k < 3; k++) {
            cudaIpcMemHandle_t h;
            memcpy(&h, all + ((size_t)r * F3D_IPC_HANDLES_PER_RANK + k) * F3D_IPC_HANDLE_BYTES, sizeof h);
            void* p = nullptr;
            CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            s->peer_ptrs[r * F3D_IPC_HANDLES_PER_RANK + k] = p;
        }
    }
    s->n_peer_ptrs = (int)world * F3D_IPC_HANDLES_PER_RANK;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// one-call drop-in: the driver loop of render_terrain.rs:1123-1244
// ------------------------------------------------------------------------------------------------
extern "C" int f3d_terrain_reference_render(const f3d_terrain_desc* desc, f3d_terrain_out* out) {
    g_err[0] = 0;
    if (!desc || !out) return fail(F3D_ERR_ARGUMENT, "null argument");
    f3d_session* s = nullptr;
    int rc = f3d_session_create(desc, nullptr, &s);
    if (rc) return rc;
    if ((rc = early_aov_readback(s, out))) { f3d_session_destroy(s); return rc; }
    float variance = INFINITY;
    bool converged = false;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    cudaEventRecord(t0, s->stream);
    while (s->frames < desc->max_frames) {
        // run up to the next gate: frames % 32 == 0 or frames == max_frames (:1206-1207)
        uint32_t to_gate = 32u - (s->frames % 32u);
        uint32_t n = std::min(to_gate, desc->max_frames - s->frames);
        rc = f3d_session_render_frames(s, n);
        if (rc) break;
        const uint32_t n_window = ((s->frames - 1u) % 32u) + 1u;
        if (n_window >= 2u) {
            int32_t bad = 0;
            rc = f3d_session_variance(s, &variance, &bad);
            if (rc) break;
            if (bad) { rc = fail(F3D_ERR_RENDER, "terrain PT produced non-finite variance (NaN in accumulation)"); break; }
            if (s->frames >= desc->min_frames && variance < desc->variance_threshold) { converged = true; break; }
        }
    }
    if (!rc) {
        cudaEventRecord(t1, s->stream);
        cudaStreamSynchronize(s->stream);
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        s->frames_ms = ms;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    if (!rc && !converged)
        rc = fail(F3D_ERR_RENDER,
                  "terrain PT did not converge: per-pixel luminance variance %.3e over the last 32-frame window after %u frames (threshold %.1e); raise max_frames or simplify the scene — refusing to return a fake reference",
                  (double)variance, s->frames, (double)desc->variance_threshold);
    if (!rc) rc = f3d_session_resolve_host(s, out);
    if (!rc) { out->variance = variance; out->converged = 1; }
    f3d_session_destroy(s);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// KAT seams
// ------------------------------------------------------------------------------------------------
int select_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(F3D_ERR_DEVICE, "no CUDA device available: forge3d_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(F3D_ERR_DEVICE, "CUDA device %d out of range (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    return 0;
}

extern "C" int f3d_build_minmax(const float* heights, uint32_t w, uint32_t h, int32_t device, uint32_t* dims,
                                float* levels_out, uint64_t cap) {
    g_err[0] = 0;
    if (!heights || !dims) { fail(F3D_ERR_ARGUMENT, "null argument"); return -F3D_ERR_ARGUMENT; }
    if (w < 2 || h < 2) { fail(F3D_ERR_UPLOAD, "terrain heightfield must be at least 2x2 texels, got %ux%u", w, h); return -F3D_ERR_UPLOAD; }
    int rc = select_device(device);
    if (rc) return -rc;
    DeviceTerrain T;
    uint64_t launches = 0;
    rc = build_device_terrain(heights, w, h, 1.0f, nullptr, &T, &launches, true);
    if (rc) { T.release(); return -rc; }
    for (int l = 0; l < T.nlevels; l++) { dims[2 * l] = T.dims[l][0]; dims[2 * l + 1] = T.dims[l][1]; }
    if (levels_out) {
        if (cap < T.mm_total * 2) { T.release(); fail(F3D_ERR_ARGUMENT, "levels_out too small"); return -F3D_ERR_ARGUMENT; }
        if (cudaMemcpy(levels_out, T.mm_base, T.mm_total * sizeof(float2), cudaMemcpyDeviceToHost) != cudaSuccess) {
            T.release(); fail(F3D_ERR_DEVICE, "copy failed"); return -F3D_ERR_DEVICE;
        }
    }
    int n = T.nlevels;
    T.release();
    return n;
}

extern "C" int f3d_trace_rays(const float* heights, uint32_t w, uint32_t h, const float spacing[2], const float origin_xz[2],
                              float exaggeration, float inv_two_r_prime, int32_t curvature_enabled, const float* rays,
                              uint64_t n, int32_t any_hit, int32_t apply_curvature, int32_t device, int32_t variant,
                              uint8_t* hit, float* t, float* normal, uint64_t* nodes_popped) {
    g_err[0] = 0;
    if (!heights || !rays || !hit || !t) return fail(F3D_ERR_ARGUMENT, "null argument");
    if (w < 2 || h < 2) return fail(F3D_ERR_UPLOAD, "terrain heightfield must be at least 2x2 texels, got %ux%u", w, h);
    int rc = select_device(device);
    if (rc) return rc;
    DeviceTerrain T;
    uint64_t launches = 0;
    rc = build_device_terrain(heights, w, h, exaggeration, nullptr, &T, &launches, true);
    if (rc) { T.release(); return rc; }
    SceneParams S{};
    fill_scene_terrain(&S, T);
    S.ox = origin_xz[0]; S.oz = origin_xz[1]; S.sx = spacing[0]; S.sz = spacing[1];
    S.inv_two_r_prime = inv_two_r_prime; S.curvature_enabled = curvature_enabled ? 1u : 0u;
    S.traversal_mode = 3u;
    FastScene F{};
    fill_fast_scene(&F, S, T);
    const uint32_t depth = stack_depth_for(T.nlevels);
    const size_t smem = (size_t)depth * kTraceThreads * 4;
    float4* d_rays = nullptr; uint8_t* d_hit = nullptr; float* d_t = nullptr; float* d_n = nullptr;
    unsigned long long* d_nodes = nullptr;
    auto cleanup = [&]() {
        cudaDeviceSynchronize();
        cached_free(d_rays, device); cached_free(d_hit, device); cached_free(d_t, device); cached_free(d_n, device); cached_free(d_nodes, device);
        T.release();
    };
    if (nodes_popped) *nodes_popped = 0;
    if (n) {
        if (cached_malloc((void**)&d_rays, n * 32, device) != cudaSuccess || cached_malloc((void**)&d_hit, n, device) != cudaSuccess ||
            cached_malloc((void**)&d_t, n * 4, device) != cudaSuccess || cached_malloc((void**)&d_n, n * 12, device) != cudaSuccess ||
            cached_malloc((void**)&d_nodes, 8, device) != cudaSuccess) {
            cleanup();
            return fail(F3D_ERR_DEVICE, "device allocation failed");
        }
        cudaMemset(d_nodes, 0, 8);
        cudaMemcpy(d_rays, rays, n * 32, cudaMemcpyHostToDevice);
        if ((rc = allow_smem(k_trace_rays, smem))) { cleanup(); return rc; }
        k_trace_rays<<<(unsigned)((n + kTraceThreads - 1) / kTraceThreads), kTraceThreads, smem>>>(
            S, F, depth, variant, d_rays, n, any_hit, apply_curvature, d_hit, d_t, normal ? d_n : nullptr, d_nodes);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { cleanup(); return fail(F3D_ERR_DEVICE, "trace kernel failed: %s", cudaGetErrorString(e)); }
        cudaMemcpy(hit, d_hit, n, cudaMemcpyDeviceToHost);
        cudaMemcpy(t, d_t, n * 4, cudaMemcpyDeviceToHost);
        if (normal) cudaMemcpy(normal, d_n, n * 12, cudaMemcpyDeviceToHost);
        if (nodes_popped) { unsigned long long v = 0; cudaMemcpy(&v, d_nodes, 8, cudaMemcpyDeviceToHost); *nodes_popped = v; }
    }
    cleanup();
    return 0;
}

#ifdef F3D_SCHED_STATS
// Tuning builds only (see F3D_SCHED_STATS in f3d_kernels.cuh); not part of the public ABI.
extern "C" int f3d_debug_sched_stats(unsigned long long* out40, int reset) {
#ifdef EMU_SIMT
    memcpy(out40, f3d::g_sched_stats, 40 * sizeof(unsigned long long));
    if (reset) memset(f3d::g_sched_stats, 0, 40 * sizeof(unsigned long long));
#else
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out40, f3d::g_sched_stats, 40 * sizeof(unsigned long long));
    if (reset) { unsigned long long z[40] = {}; cudaMemcpyToSymbol(f3d::g_sched_stats, z, sizeof z); }
#endif
    return 0;
}
#endif

// ------------------------------------------------------------------------------------------------
// LBVH test seam: builds the tree for a host mesh and returns Morton order, topology and boxes
// ------------------------------------------------------------------------------------------------
extern "C" int f3d_lbvh_build(const float* xyz, uint32_t nverts, const uint32_t* idx, uint32_t ntris, int32_t device, uint32_t* morton,
                              uint32_t* order, uint32_t* left, uint32_t* right, uint32_t* parent, float* nodes /* (2n-1) x 8 */) {
    g_err[0] = 0;
    if (!xyz || !idx || nverts == 0 || ntris == 0) return fail(F3D_ERR_ARGUMENT, "empty mesh");
    if (ntris > (1u << 20)) return fail(F3D_ERR_RENDER, "Triangle count %u exceeds maximum of 1M triangles", ntris);   // lbvh_gpu/build.rs:11-16
    for (size_t i = 0; i < (size_t)ntris * 3; i++)
        if (idx[i] >= nverts) return fail(F3D_ERR_RENDER, "mesh indices reference out-of-bounds vertices");
    int rc = select_device(device);
    if (rc) return rc;
    std::vector<float4> v(nverts);
    for (uint32_t i = 0; i < nverts; i++) v[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.0f);
    float4 *d_v = nullptr, *d_nodes = nullptr;
    uint32_t *d_i = nullptr, *d_order = nullptr;
    struct Free { float4*& a; float4*& b; uint32_t*& c; uint32_t*& d; int dev; ~Free() { cudaDeviceSynchronize(); cached_free(a, dev); cached_free(b, dev); cached_free(c, dev); cached_free(d, dev); } }
        guard{d_v, d_nodes, d_i, d_order, device};
    const size_t nnodes = 2 * (size_t)ntris - 1;
    CUDA_TRY(cached_malloc((void**)&d_v, nverts * sizeof(float4), device));
    CUDA_TRY(cached_malloc((void**)&d_i, (size_t)ntris * 3 * sizeof(uint32_t), device));
    CUDA_TRY(cached_malloc((void**)&d_nodes, nnodes * 2 * sizeof(float4), device));
    CUDA_TRY(cached_malloc((void**)&d_order, (size_t)ntris * sizeof(uint32_t), device));
    CUDA_TRY(cudaMemcpy(d_v, v.data(), nverts * sizeof(float4), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_i, idx, (size_t)ntris * 3 * sizeof(uint32_t), cudaMemcpyHostToDevice));
    uint64_t launches = 0;
    LbvhScratch W;
    if ((rc = build_mesh_lbvh(xyz, nverts, idx, ntris, d_v, d_i, device, nullptr, d_nodes, d_order, &launches, &W))) return rc;
    std::vector<unsigned long long> keys(ntris);
    CUDA_TRY(cudaMemcpy(keys.data(), W.keys, (size_t)ntris * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (morton) for (uint32_t i = 0; i < ntris; i++) morton[i] = (uint32_t)(keys[i] >> 32);
    if (order) CUDA_TRY(cudaMemcpy(order, d_order, (size_t)ntris * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (ntris > 1u) {
        if (left) CUDA_TRY(cudaMemcpy(left, W.left, (size_t)(ntris - 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (right) CUDA_TRY(cudaMemcpy(right, W.right, (size_t)(ntris - 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    if (parent) CUDA_TRY(cudaMemcpy(parent, W.parent, nnodes * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (nodes) CUDA_TRY(cudaMemcpy(nodes, d_nodes, nnodes * 2 * sizeof(float4), cudaMemcpyDeviceToHost));
    return 0;
}
