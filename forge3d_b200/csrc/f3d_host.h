// forge3d_b200/csrc/f3d_host.h
// Host-side helpers shared by the translation units of libforge3d_b200.so (f3d_backend.cu defines them): the thread-local error
// text behind f3d_last_error(), the per-device buffer cache, device selection, glam-style f32 vector helpers, and the device
// terrain (packed cells + min-max chain) that both the path tracer and the viewshed build.  No CPU fallback lives here.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/forge3d_b200.h"

extern thread_local char g_f3d_err[640];
int f3d_fail(int cls, const char* fmt, ...);
#define fail f3d_fail

#define CUDA_TRY(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            return fail(F3D_ERR_DEVICE, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(_e), __FILE__,     \
                        __LINE__, cudaGetErrorString(_e));                                                 \
    } while (0)

// Device-buffer cache (see f3d_backend.cu): freed blocks are parked per device and reused by the next allocation of a similar size.
cudaError_t cached_malloc(void** p, size_t bytes, int device);
void cached_free(void* p, int device, bool allow_park = true);   // caller guarantees no work that touches `p` is in flight
int select_device(int device);                                    // cudaSetDevice with the "no CPU fallback" errors

// host math mirroring glam 0.24.2 / Rust f32 (render_terrain.rs:635-661)
struct hv3 { float x, y, z; };
static inline hv3 HV(const float* p) { return hv3{p[0], p[1], p[2]}; }
static inline hv3 hsub(hv3 a, hv3 b) { return hv3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline float hdot(hv3 a, hv3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline hv3 hcross(hv3 a, hv3 b) { return hv3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
static inline float hlen(hv3 a) { return sqrtf(hdot(a, a)); }
static inline hv3 hnorm(hv3 a) { float inv = 1.0f / hlen(a); return hv3{a.x * inv, a.y * inv, a.z * inv}; }
static inline hv3 hnorm_or_zero(hv3 a) {   // glam Vec3::normalize_or_zero
    const float rcp = 1.0f / hlen(a);
    if (isfinite(rcp) && rcp > 0.0f) return hv3{a.x * rcp, a.y * rcp, a.z * rcp};
    return hv3{0.0f, 0.0f, 0.0f};
}
static inline bool finite3(const float* v) { return isfinite(v[0]) && isfinite(v[1]) && isfinite(v[2]); }
static inline float clamp_radiometric(float v) { return fminf(fmaxf(v, 0.0f), 65504.0f); }  // render_terrain.rs:571-576
static inline float to_radians_f32(float d) { return d * (3.14159274101257324f / 180.0f); }
static inline double deg2rad(double d) { return d * (3.14159265358979323846 / 180.0); }
static inline uint32_t next_pow2(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

// device terrain: packed cells + min-max levels (+ the quad-packed copy the production traversal reads)
constexpr int kHostMaxLevels = 16;   // == f3d::kMaxLevels (checked in f3d_backend.cu)
struct DeviceTerrain {
    float4* cells = nullptr;
    float2* mm_base = nullptr;
    int nlevels = 0;
    uint32_t dims[kHostMaxLevels][2] = {};
    size_t level_off[kHostMaxLevels] = {};   // in float2 units
    size_t mm_total = 0;                     // float2 count
    uint32_t cell_w = 0, cell_h = 0;
    uint64_t bytes = 0;
    float2* quad_base = nullptr;             // levels 0..nlevels-2 grouped by their parent (f3d_trace_fast.cuh)
    size_t quad_off[kHostMaxLevels] = {};
    uint32_t quad_pitch[kHostMaxLevels] = {}, quad_ph[kHostMaxLevels] = {};
    size_t quad_total = 0;
    float2 root_mm = {0.0f, 0.0f};
    int device = 0;
    void release_plain() { cached_free(mm_base, device); mm_base = nullptr; }
    void release() {
        cached_free(cells, device);
        cached_free(quad_base, device);
        cells = nullptr; quad_base = nullptr;
        release_plain();
    }
};
// Uploads the DEM, scans it for non-finite samples, and builds cells + pyramid on the device (k_build_level0 / k_reduce_level /
// k_pack_quads).  keep_plain = keep the plain per-level chain (KAT seams, viewshed); otherwise only the quad-packed copy stays.
int build_device_terrain(const float* h_heights, uint32_t w, uint32_t h, float ex, cudaStream_t stream, DeviceTerrain* T,
                         uint64_t* launches, bool keep_plain);

// Host median-split BVH over a triangle mesh in the traversal format of intersect_mesh (f3d_trace.cuh) / wf_mesh (f3d_wavefront.cuh):
// 2 float4 per node, leaf boxes padded against rounding, `tris` = triangle ids in leaf order.
void host_build_mesh_bvh(const float* xyz, const uint32_t* idx, uint32_t ntris, std::vector<float4>* nodes, std::vector<uint32_t>* tris);
