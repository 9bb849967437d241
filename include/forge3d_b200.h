/*
 * include/forge3d_b200.h -- C ABI of libforge3d_b200.so, the B200-native backend for forge3d's
 * path-traced DEM snapshot path.
 *
 * The reference has no C ABI on this path; its seam is the two Rust lines
 *     let tracer = HybridPathTracer::new()?;
 *     let out = tracer.render_terrain_reference(&desc)?;
 * at /root/reference/src/py_functions/path_tracing/terrain_reference.rs:416-417.  Everything behind
 * those lines (src/path_tracing/hybrid_compute/render_terrain.rs:563-1434 and the WGSL kernels it
 * dispatches) is what this library replaces; everything above them (PyO3 marshalling, Python
 * validation) stays.  INTEGRATION.md shows the Rust `extern "C"` shim a maintainer would add.
 *
 * Conventions: plain pointers and sizes only, no C++/torch types.  Inputs are borrowed for the
 * duration of a call; outputs are caller-allocated.  Every function returns 0 on success or an
 * f3d_status error class; the message (reference-compatible text, see
 * tests/test_hybrid_terrain_pt.py:411-458 for the pinned substrings) is read with f3d_last_error(),
 * which is thread-local.  There is NO CPU fallback: without a CUDA device every entry point fails
 * with F3D_ERR_DEVICE.
 */
#ifndef FORGE3D_B200_H
#define FORGE3D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F3D_ABI_VERSION 2

/* Error classes; mirror RenderError::{Render,Upload,Device,Budget}, src/core/error.rs:10-32. */
typedef enum f3d_status {
    F3D_OK = 0,
    F3D_ERR_RENDER = 1,
    F3D_ERR_UPLOAD = 2,
    F3D_ERR_DEVICE = 3,
    F3D_ERR_BUDGET = 4,
    F3D_ERR_ARGUMENT = 5
} f3d_status;

/* EarthModel / RefractionModel::from_name, src/geo/refraction.rs:44-54,79-99 */
enum { F3D_EARTH_FLAT = 0, F3D_EARTH_SPHERE = 1, F3D_EARTH_ELLIPSOID = 2 };
enum { F3D_REFRACTION_NONE = 0, F3D_REFRACTION_BENNETT = 1, F3D_REFRACTION_SAEMUNDSSON = 2,
       F3D_REFRACTION_EFFECTIVE_RADIUS = 3 };

/* Replaces AtmosphereLutHandle as the AETHER post consumes it (src/core/atmosphere/runtime.rs:44-90; the uniforms
 * and LUT uploads of src/path_tracing/hybrid_compute/aether_post.rs:10-23,58-178,365-440).  The caller (the Rust
 * side, which owns the shipped LUT bank / a baked handle) passes the three RGBA16F payloads it would have uploaded
 * as textures: texel (x, y, z) at ((z * dim_y + y) * dim_x + x) * 4 halves, little-endian, borrowed for the call. */
typedef struct f3d_atmosphere {
    const uint16_t* transmittance;     /* dims transmittance_mu x transmittance_height */
    const uint16_t* scattering;        /* accumulated scattering: scattering_mu_view x scattering_mu_sun x
                                          (scattering_height * scattering_nu) */
    const uint16_t* aerial;            /* aerial_distance x aerial_mu_view x aerial_height (rgb = 0, a = transmittance) */
    uint32_t transmittance_mu, transmittance_height;
    uint32_t scattering_mu_view, scattering_mu_sun, scattering_height, scattering_nu;
    uint32_t aerial_distance, aerial_mu_view, aerial_height;
    /* AtmosphereConfig (src/core/atmosphere/bake.rs:132-162) */
    float bottom_radius_m, top_radius_m, max_aerial_distance_m, ozone_du;
    float mie_g, turbidity, rayleigh_scale_height_m, mie_scale_height_m, ground_albedo;
} f3d_atmosphere;

/* Replaces TerrainReferenceDesc, render_terrain.rs:239-282. */
typedef struct f3d_terrain_desc {
    const float* heights;          /* row-major dem_h x dem_w (desc.heights); host memory, or memory of `device` (read in place) */
    uint32_t dem_w, dem_h;
    float spacing[2];
    float exaggeration;
    float albedo[3];
    float cam_origin[3], cam_look_at[3], cam_up[3];
    float fov_y_deg, exposure;
    float sun_az_deg, sun_el_deg, sun_intensity, sun_color[3];
    double observer_lat_deg, observer_lon_deg;   /* observer_geodetic_deg */
    int32_t earth_model;           /* F3D_EARTH_* */
    double sphere_radius_m;
    int32_t refraction_model;      /* F3D_REFRACTION_* */
    double refraction_k, pressure_mbar, temperature_c;
    const float* env_rgb;          /* host, env_h x env_w x 3, or NULL (constant env) */
    uint32_t env_w, env_h;
    float env_intensity;
    const float* mesh_xyz;         /* host, mesh_nverts x 3, or NULL */
    uint32_t mesh_nverts;
    const uint32_t* mesh_idx;      /* host, mesh_ntris x 3 */
    uint32_t mesh_ntris;
    uint32_t width, height, seed, spp, max_frames, min_frames;
    float variance_threshold;
    /* ---- extensions (zero = reference behaviour on one GPU) ---- */
    int32_t device;                /* CUDA device ordinal */
    int32_t compat_512mib_gate;    /* 1 = enforce the reference's 512 MiB working-set gate
                                      (src/core/memory_tracker.rs:16, render_terrain.rs:875-888) */
    uint32_t part_rank, part_world;/* image-row partition: this process renders row blocks
                                      b with b % part_world == part_rank (world 0/1 = whole image) */
    uint32_t part_block_rows;      /* rows per block (0 = default 16; rounded up to a multiple of 16) */
    uint32_t part_mode;            /* 0 = exact: border rows of the reservoir image and a per-frame flag travel over peer
                                      memory, the result is bit-identical to one GPU;
                                      1 = gather-only: the spatial reuse pass clamps its neighbours to the row block it
                                      renders, ranks never talk during the frame loop; pixels within 4 rows
                                      of a block border differ slightly from the one-GPU image */
    /* ---- desc.atmosphere (render_terrain.rs:262-265): AETHER aerial-perspective post over the converged
     * accumulation and the frame-0 depth AOV (:1246-1311); NULL = none.  Only `rgba` changes; AOVs do not. ---- */
    const f3d_atmosphere* atmosphere;
} f3d_terrain_desc;

/* Replaces TerrainReferenceOutput, render_terrain.rs:285-299. */
typedef struct f3d_terrain_out {
    uint8_t* rgba;     /* host, W*H*4, caller allocated */
    float* albedo;     /* host, W*H*3 */
    float* normal;     /* host, W*H*3 */
    float* depth;      /* host, W*H (NaN 0x7fc00000 = miss) */
    float* accum;      /* optional host W*H*4 linear accumulation (rgb sum, frame count) or NULL */
    uint32_t frames;
    float variance;
    int32_t converged;
    uint64_t peak_host_visible_bytes, minmax_pyramid_bytes, gpu_resource_bytes;
    /* measurement extensions */
    uint64_t rays_primary, rays_shadow, rays_ibl, nodes_popped;
    double setup_ms, frames_ms, readback_ms;
    uint64_t kernel_launches;
} f3d_terrain_out;

/* One-call drop-in: validate, upload, build pyramid, accumulate until the windowed-variance
 * gate passes (or error), resolve, read back.  == HybridPathTracer::render_terrain_reference. */
int f3d_terrain_reference_render(const f3d_terrain_desc* desc, f3d_terrain_out* out);

const char* f3d_last_error(void);
int f3d_abi_version(void);
int f3d_device_count(void);
/* "src=<content hash of the sources>;defines=<-D variant switches, comma separated>": which build this library is.
 * No reference counterpart (the reference ships one wgpu pipeline); forge3d_b200/build.py and bench.py compare it so that
 * an A/B variant build can never pass for the validated default. */
const char* f3d_build_info(void);

/* Page-locked host memory from a small pool inside the library.  Output buffers allocated here are filled by one DMA at PCIe
 * speed; any other host pointer works too (the library bounces it through page-locked staging).  Replaces the reference's
 * mapped read-back buffers (render_terrain.rs:1339-1393 reads wgpu MAP_READ staging buffers).  NULL when no CUDA device. */
void* f3d_host_alloc(uint64_t bytes);
void f3d_host_free(void* p);
/* The library parks freed device buffers (<= 4 GiB per device, F3D_B200_CACHE_MB overrides, 0 disables) so that back-to-back
 * calls do not pay cudaMalloc's page-table work again; this returns them to the driver (device < 0: every device).  Returns the
 * bytes released.  No reference counterpart (wgpu owns its allocator). */
uint64_t f3d_cache_trim(int32_t device);

/* ---- session API: the same render split at the reference driver-loop's joints, so a host
 * (bench, multi-GPU launcher) can keep inputs resident and time the frame loop alone. ---- */
typedef struct f3d_session f3d_session;

/* TerrainPtScene::new + uniform/buffer setup + main_terrain_gbuffer (render_terrain.rs:581-1121).
 * `cuda_stream` is a cudaStream_t; NULL = the library creates its own non-blocking stream
 * (pass cudaStreamLegacy = (void*)1 to run on the legacy default stream). */
int f3d_session_create(const f3d_terrain_desc* desc, void* cuda_stream, f3d_session** out_session);
/* Enqueue `n` accumulation frames (one pass of render_terrain.rs:1127-1204 each); asynchronous. */
int f3d_session_render_frames(f3d_session* s, uint32_t n);
/* The convergence read-back (render_terrain.rs:1206-1233): max over owned pixels of m2/(n-1).
 * Synchronises the stream.  *nonfinite != 0 means NaN/inf variance. */
int f3d_session_variance(f3d_session* s, float* vmax, int32_t* nonfinite);
/* Final reuse pass + reservoir validity (render_terrain.rs:1313-1337) + resolve of owned rows
 * into DEVICE buffers (any may be NULL): rgba8 W*H*4, albedo/normal W*H*3 f32, depth W*H f32.
 * Rows not owned by this partition are left untouched.  Asynchronous on the session stream,
 * except that validity is checked (synchronises) when check_validity != 0. */
int f3d_session_resolve_device(f3d_session* s, void* d_rgba, void* d_albedo, void* d_normal, void* d_depth,
                               int32_t check_validity);
/* Reservoir validity seen by the last f3d_session_resolve_device(check_validity != 0) call, for launchers that
 * partition the image: *required != 0 when the scene is sun-lit (render_terrain.rs:465-471), *any_valid != 0
 * when an owned pixel holds a valid reservoir (:1313-1337).  A partitioned session never raises the
 * "no valid reservoirs" error itself (a rank may own only sky); the launcher ORs any_valid over ranks and
 * raises when required && !any.  No device work. */
int f3d_session_validity(const f3d_session* s, int32_t* any_valid, int32_t* required);
/* Same, then copies to host buffers of `out` and fills the scalar fields. */
int f3d_session_resolve_host(f3d_session* s, f3d_terrain_out* out);
int f3d_session_frames(const f3d_session* s, uint32_t* frames);
int f3d_session_stats(f3d_session* s, f3d_terrain_out* stats_only);   /* counters/bytes only; syncs */
int f3d_session_sync(f3d_session* s);
/* Device time (CUDA events on the session stream) of the most recent f3d_session_render_frames
 * call; synchronises. */
int f3d_session_last_frames_ms(f3d_session* s, double* ms);
void f3d_session_destroy(f3d_session* s);

/* ---- multi-GPU halo exchange over NVLink peer memory (exact mode) ----
 * Each rank exports one 64-byte CUDA IPC handle per shared allocation; the launcher all-gathers
 * them (any transport) and hands every rank the full table.  After import, the frame kernel
 * stores the border rows of its temporal-reuse records directly into the neighbours' images. */
#define F3D_IPC_HANDLE_BYTES 64
#define F3D_IPC_HANDLES_PER_RANK 3
int f3d_session_ipc_export(f3d_session* s, uint8_t* handles /* F3D_IPC_HANDLES_PER_RANK*64 bytes */);
int f3d_session_ipc_import(f3d_session* s, const uint8_t* all_handles /* part_world * per-rank bytes */);

/* ---- KAT seam (== the reference's test-only entry main_helios_production_terrain_trace_proof,
 * src/path_tracing/hybrid_compute/terrain_heightfield.rs:1646-1671): run terrain_trace over a
 * host ray batch (n x 8 floats: origin.xyz, tmin, direction.xyz, tmax) on the GPU. */
int f3d_trace_rays(const float* heights, uint32_t dem_w, uint32_t dem_h, const float spacing[2],
                   const float origin_xz[2], float exaggeration, float inv_two_r_prime,
                   int32_t curvature_enabled, const float* rays, uint64_t n, int32_t any_hit,
                   int32_t apply_curvature, int32_t device,
                   int32_t variant /* 0 = production traversal, 1 = literal restatement of the WGSL loop */,
                   uint8_t* hit, float* t, float* normal, uint64_t* nodes_popped /* may be NULL */);

/* GPU min-max pyramid build (== build_minmax_mips, terrain_heightfield.rs:132-202), copied back
 * to the host for parity checks: dims[2*l..] and levels finest first, [min,max] pairs. */
int f3d_build_minmax(const float* heights, uint32_t dem_w, uint32_t dem_h, int32_t device,
                     uint32_t* dims, float* levels_out, uint64_t levels_capacity_floats);

/* ---- smoke volume ray-march (BASELINE config 4) ----
 * Replaces SmokeVolume::raymarch_rgba / raymarch_projection_rgba, /root/reference/src/smoke/render.rs:6-175 (the
 * reference runs them single-threaded on the CPU behind PySmokeDomain.render_rgba / render_projection_rgba,
 * src/smoke/py.rs:531-628).  The volume is uploaded once and stays resident; a render is one kernel launch. */
typedef struct f3d_smoke_volume {     /* SmokeVolume + SmokeDomainConfig, src/smoke/types.rs:8-14,331-345 */
    uint32_t dims[3];                  /* x, y, z (each >= 2); voxel (x, y, z) at (z * dims[1] + y) * dims[0] + x */
    float voxel_size[3], origin[3];
    const float* density;              /* host, required */
    const float* temperature;          /* host; the remaining fields may be NULL = all zero */
    const float* soot;
    const float* humidity;
    const float* emission_rate;
    const float* particle_age;
    uint64_t frame_index;              /* seeds the per-pixel march jitter (render.rs:76-79) */
} f3d_smoke_volume;

typedef struct f3d_smoke_settings {   /* SmokeRenderSettings, src/smoke/types.rs:225-266 (defaults there) */
    float density_scale, extinction, scattering, absorption, phase_g, step_size;
    uint32_t max_steps;
    int32_t self_shadow;
    uint32_t shadow_steps;
    float shadow_step_size, jitter_strength, exposure;
    float thin_color[3], dense_color[3];
    float soot_absorption, fire_glow;
} f3d_smoke_settings;

typedef struct f3d_smoke f3d_smoke;
int f3d_smoke_create(const f3d_smoke_volume* volume, int32_t device, f3d_smoke** out);
void f3d_smoke_destroy(f3d_smoke* s);
/* rgba: host, height * width * 4, caller allocated.  *kernel_ms (may be NULL) = device time of the march kernel. */
int f3d_smoke_raymarch_rgba(f3d_smoke* s, const f3d_smoke_settings* settings, uint32_t width, uint32_t height,
                            const float camera_pos[3], const float target[3], const float up[3], float fovy_deg,
                            const float sun_direction[3], uint8_t* rgba, double* kernel_ms);
/* BASELINE config 4 ("volumetric smoke frame over terrain"): the perspective march and the composite over a terrain frame in
 * ONE kernel.  base_rgba = the terrain snapshot (same width x height, e.g. f3d_terrain_reference_render's rgba); the layer goes
 * over it with the reference's straight-alpha compositor (python/forge3d/map_scene.py:1588-1604 _alpha_composite_rgba:
 * rgb = u8(clip(dst (1 - a) + src a)), a = max).  base_depth (nullable) = distance along the camera ray to the terrain
 * (the depth AOV; NaN / <= 0 = sky): when given, each ray's march ends at the surface, so smoke behind a ridge is hidden -
 * an extension the reference's caller-side composite cannot do; NULL reproduces the reference pixel for pixel. */
int f3d_smoke_raymarch_over_rgba(f3d_smoke* s, const f3d_smoke_settings* settings, uint32_t width, uint32_t height,
                                 const float camera_pos[3], const float target[3], const float up[3], float fovy_deg,
                                 const float sun_direction[3], const uint8_t* base_rgba, const float* base_depth,
                                 uint8_t* rgba, double* kernel_ms);
int f3d_smoke_raymarch_projection_rgba(f3d_smoke* s, const f3d_smoke_settings* settings, uint32_t width, uint32_t height,
                                       const float view_direction[3], const float sun_direction[3], uint8_t* rgba,
                                       double* kernel_ms);

/* ---- HELIOS viewshed / solar shadow mask ----
 * Replace compute_viewshed / compute_shadow_mask, /root/reference/src/terrain/analysis/viewshed.rs:341-347,396-570 (called from
 * src/py_functions/geodesy.rs:351,502).  Inputs are what those functions take: the DEM, the per-cell geodesic offsets (or
 * geodetic position + sun angles) computed by the caller, and ViewshedOptions. */
typedef struct f3d_viewshed_options {    /* ViewshedOptions, viewshed.rs:8-25 */
    uint32_t width, height;
    float observer_x, observer_y, observer_height_m, target_height_m, max_distance_m;
    float observer_latitude_rad, observer_longitude_rad, left_unwrapped_deg, top_deg;
    float longitude_step_deg, latitude_step_deg, geodesic_sphere_radius_m;
    int32_t earth_model;                 /* F3D_EARTH_*; Ellipsoid { latitude_deg } = earth_latitude_deg */
    double earth_latitude_deg, sphere_radius_m;
    int32_t refraction_model;            /* F3D_REFRACTION_* */
    double refraction_k, pressure_mbar, temperature_c;
    int32_t device;
} f3d_viewshed_options;
/* heights: host height x width; positions_m: host n x 2 (east, north metres from the observer).  Outputs (host, n each):
 * visibility 0/1, curvature drop, refraction gain, horizon distance.  *kernel_ms may be NULL. */
int f3d_viewshed(const float* heights, const float* positions_m, const f3d_viewshed_options* options, uint8_t* visibility,
                 float* curvature_drop_m, float* refraction_gain_m, float* horizon_distance_m, double* kernel_ms);
/* geodetic_and_sun: host n x 4 (latitude rad, longitude rad, sun azimuth rad, launch elevation rad); lit: host n, 1 = lit. */
int f3d_shadow_mask(const float* heights, const float* geodetic_and_sun, const f3d_viewshed_options* options, uint8_t* lit,
                    double* kernel_ms);

/* ---- GPU LBVH build seam (replaces GpuBvhBuilder::build, /root/reference/src/accel/lbvh_gpu/build.rs:4-76:
 * Morton codes -> sort -> Karras link -> boxes).  Outputs (host, any may be NULL): morton[n] sorted codes, order[n] triangle ids in
 * leaf order, left/right[n-1] child NODE indices of the internal nodes (internal i, leaf n-1+i), parent[2n-1], nodes[(2n-1)*8] =
 * (min.xyz, w0, max.xyz, w1) per node in the traversal format (leaf boxes are padded, see csrc/f3d_lbvh.cuh). */
int f3d_lbvh_build(const float* xyz, uint32_t nverts, const uint32_t* idx, uint32_t ntris, int32_t device, uint32_t* morton,
                   uint32_t* order, uint32_t* left, uint32_t* right, uint32_t* parent, float* nodes);

/* ---- Wavefront multi-bounce path tracer.  Replaces render_pt_reference
 * (/root/reference/src/path_tracing/adjudication.rs:76-331) and the WavefrontScheduler frame loop behind it
 * (src/path_tracing/wavefront/render.rs:87-209, pt_raygen / pt_intersect / pt_shade / pt_shadow / pt_scatter.wgsl), plus the
 * Reinhard + sRGB resolve of src/core/tonemap.rs:11-32.  The buffers are the ones that function binds, in their GPU layouts. */
typedef struct f3d_wavefront_scene {
    float cam_origin[3], cam_forward[3], cam_right[3], cam_up[3];   /* ReferenceSceneDesc::camera_basis, reference_scene.rs:201-207 */
    float fov_y_rad, exposure;
    uint32_t seed_hi, seed_lo;                      /* per-frame seeds are splitmix32 hashes of these, adjudication.rs:226-236 */
    const float* spheres; uint32_t nspheres;        /* 20 floats each: WavefrontGpuSphere (80 B), reference_scene.rs:101-117 */
    const float* dir_lights; uint32_t ndir;         /* 8 floats each: GpuDirectionalLight, lighting.rs:88-96 */
    const float* area_lights; uint32_t narea;       /* 12 floats each: GpuAreaLight, lighting.rs:19-28 */
    const float* importance; uint32_t nimportance;  /* object_importance, one per material slot */
    float environment[16];                          /* ReferenceEnvironmentRaw: env_ground, env_sky, miss_ground, miss_sky */
    const float* mesh_xyz; uint32_t mesh_nverts;    /* one BLAS (may be empty) */
    const uint32_t* mesh_idx; uint32_t mesh_ntris;
    const float* instances; uint32_t ninstances;    /* 36 words each: InstanceData (transform, inv_transform column-major, blas_index,
                                                       material_id, pad); 0 instances = the mesh is traced un-instanced with material 0 */
} f3d_wavefront_scene;
typedef struct f3d_wavefront_stats {
    uint64_t rays;                 /* rays traced over all frames */
    uint64_t max_rays_per_frame;   /* the reference's append-only ray queue holds 4 * W * H per frame (wavefront/mod.rs:34,77) */
    uint32_t min_iterations;       /* fewest wavefront iterations in any frame */
    uint32_t launches;             /* kernels launched */
    double kernel_ms;              /* device time, CUDA events */
} f3d_wavefront_stats;
/* hdr_rgba: host W*H*4 floats (mean radiance, alpha 1) or NULL; rgba8: host W*H*4 bytes or NULL; stats may be NULL.
 * Errors as render_pt_reference raises them: zero size, a frame with < 2 wavefront iterations, ray-queue overflow. */
int f3d_wavefront_render(const f3d_wavefront_scene* scene, uint32_t width, uint32_t height, uint32_t spp_frames, int32_t device,
                         float* hdr_rgba, uint8_t* rgba8, f3d_wavefront_stats* stats);
/* Multi-GPU extension (the reference is single-GPU): the same render restricted to the rows one rank owns -- interleaved blocks of
 * block_rows rows, block b belongs to rank b % world.  Pixels are independent, so the owned rows are bit-identical to the one-GPU image.
 * hdr_rgba / rgba8 are full-size; only owned rows are meaningful.  The two frame rules are global properties, so they are NOT applied
 * here: frame_iterations[spp_frames] / frame_rays[spp_frames] (host, may be NULL) receive this rank's per-frame numbers for the caller to
 * reduce over ranks (MAX / SUM) and check (forge3d_b200/distributed.py::wavefront_partitioned). */
typedef struct f3d_wavefront_part {
    uint32_t rank, world, block_rows;
    uint32_t* frame_iterations;
    uint64_t* frame_rays;
} f3d_wavefront_part;
int f3d_wavefront_render_part(const f3d_wavefront_scene* scene, uint32_t width, uint32_t height, uint32_t spp_frames, int32_t device,
                              const f3d_wavefront_part* part, float* hdr_rgba, uint8_t* rgba8, f3d_wavefront_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
