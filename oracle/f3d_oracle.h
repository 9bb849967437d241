/*
 * oracle/f3d_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Plain-C f32 restatement of forge3d's path-traced DEM snapshot path
 * (`HybridPathTracer::render_terrain_reference` + WGSL `main_terrain`,
 * `main_terrain_gbuffer`, `pt_restir_temporal::main`, `pt_restir_spatial::main`).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product path
 * (forge3d_b200/csrc + libforge3d_b200.so) never links or calls it.
 *
 * Parity pin status: PINNED against the reference's own artefacts --
 *   tests/golden/hybrid_terrain/mini_dem_reference.png (drift gate SSIM>=0.995,
 *   mean-abs<=2.0 of tests/test_hybrid_terrain_pt.py:818-859), the AOV analytic
 *   gates (:290-380), the conservative-descent KAT of
 *   src/path_tracing/hybrid_compute/terrain_heightfield.rs:2001-2127 and the
 *   pyramid unit tests (:528-612).  See tests/test_oracle_*.py.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  The numerics contract (operation order, no FMA
 * contraction, pinned sin/cos/atan2/acos) is stated in DESIGN.md section 4.
 */
#ifndef F3D_ORACLE_H
#define F3D_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* AETHER LUT handle as the post consumes it (AtmosphereLutHandle, src/core/atmosphere/runtime.rs:44-90;
 * uniforms of src/path_tracing/hybrid_compute/aether_post.rs:10-23,131-178).  LUT payloads are RGBA16F texels,
 * x fastest, then y, then z (write_texture layout, aether_post.rs:365-397). */
typedef struct f3do_atmosphere {
    const uint16_t* transmittance;    /* [height][mu] x 4 halves */
    const uint16_t* scattering;       /* accumulated scattering, [height*nu + nu_i][mu_sun][mu_view] x 4 */
    const uint16_t* aerial;           /* [height][mu_view][distance] x 4 (rgb = 0, a = mean transmittance) */
    uint32_t transmittance_dims[2];   /* mu, height */
    uint32_t scattering_dims[3];      /* mu_view, mu_sun, scattering_height * scattering_nu */
    uint32_t scattering_height, scattering_nu;
    uint32_t aerial_dims[3];          /* distance, mu_view, height */
    float bottom_radius_m, top_radius_m, max_aerial_distance_m, ozone_du;
    float mie_g, turbidity, rayleigh_scale_height_m, mie_scale_height_m, ground_albedo;
} f3do_atmosphere;

/* Camera / light block of PrometheusAetherUniforms (aether_post.rs:131-156). */
typedef struct f3do_aether_view {
    uint32_t width, height;
    float cam_origin[3], cam_right[3], cam_up[3], cam_forward[3];
    float tan_half_fov, aspect, exposure;
    float light_dir[3], sun_intensity;
} f3do_aether_view;

/* Mirrors TerrainReferenceDesc, src/path_tracing/hybrid_compute/render_terrain.rs:239-282 */
typedef struct f3do_desc {
    const float* heights;        /* row-major dem_h x dem_w */
    uint32_t dem_w, dem_h;
    float spacing[2];
    float exaggeration;
    float albedo[3];
    float cam_origin[3], cam_look_at[3], cam_up[3];
    float fov_y_deg, exposure;
    float sun_az_deg, sun_el_deg, sun_intensity, sun_color[3];
    double observer_lat_deg, observer_lon_deg;
    int32_t earth_model;         /* 0 flat, 1 sphere, 2 ellipsoid */
    double sphere_radius_m;
    int32_t refraction_model;    /* 0 none, 1 bennett, 2 saemundsson, 3 effective_radius */
    double refraction_k, pressure_mbar, temperature_c;
    const float* env_rgb;        /* env_h x env_w x 3 or NULL */
    uint32_t env_w, env_h;
    float env_intensity;
    const float* mesh_xyz;       /* nverts x 3 or NULL */
    uint32_t mesh_nverts;
    const uint32_t* mesh_idx;    /* ntris x 3 */
    uint32_t mesh_ntris;
    uint32_t width, height, seed, spp, max_frames, min_frames;
    float variance_threshold;
    int32_t compat_512mib_gate;  /* 1 = enforce the reference's 512 MiB working-set gate */
    const f3do_atmosphere* atmosphere;   /* AETHER aerial-perspective post (render_terrain.rs:262-265,1246-1311) or NULL */
} f3do_desc;

/* Mirrors TerrainReferenceOutput, render_terrain.rs:285-299 (+ ray counters). */
typedef struct f3do_out {
    uint8_t* rgba;     /* W*H*4, caller allocated */
    float* albedo;     /* W*H*3 */
    float* normal;     /* W*H*3 */
    float* depth;      /* W*H   */
    float* accum;      /* optional W*H*4 linear accumulation (rgb sum, frame count) or NULL */
    uint32_t frames;
    float variance;
    int32_t converged;
    uint64_t minmax_pyramid_bytes;
    uint64_t rays_primary, rays_shadow, rays_ibl;
    uint64_t nodes_popped;   /* terrain_trace stack pops inside the frame loop */
    double setup_seconds;    /* validation + pyramid build + G-buffer pass (wall clock) */
    double frames_seconds;   /* the accumulation loop incl. convergence checks (wall clock) */
    /* diagnostics of the merged reservoirs after the last reuse pass (render_terrain.rs:1313-1337,
     * runtime contract :175-232) */
    uint32_t prev_m_max;
    float prev_weight_max, prev_w_sum_max;
    float prev_dir_min[3], prev_dir_max[3];   /* over valid reservoirs */
} f3do_out;

/* 0 = ok; otherwise an error class (1 render, 2 upload); message via f3do_last_error(). */
int f3do_render(const f3do_desc* desc, f3do_out* out);
const char* f3do_last_error(void);
void f3do_set_threads(int n);   /* OpenMP threads used by f3do_render (0 = default) */
int f3do_get_threads(void);

/* build_minmax_mips, terrain_heightfield.rs:132-202.  On success returns the number of
 * levels and fills dims[2*l], dims[2*l+1] (caller provides >= 2*32 entries) and, when
 * `levels_out` is non-NULL, the concatenated levels (finest first, [min,max] pairs). */
int f3do_build_minmax(const float* heights, uint32_t w, uint32_t h,
                      uint32_t* dims, float* levels_out, uint64_t levels_capacity_floats);

/* Batch ray KAT interface over terrain_trace (hybrid_terrain_traversal.wgsl:254-372).
 * rays: n x 8 floats (origin.xyz, tmin, direction.xyz, tmax).  origin_xz = world xz of
 * texel (0,0).  Outputs: hit[n] (0/1), t[n], normal[n*3] (may be NULL). */
int f3do_trace_rays(const float* heights, uint32_t w, uint32_t h,
                    const float spacing[2], const float origin_xz[2], float exaggeration,
                    float inv_two_r_prime, int curvature_enabled,
                    const float* rays, uint64_t n, int any_hit, int apply_curvature,
                    uint8_t* hit, float* t, float* normal);

/* effective_radius_m / EarthCurvatureUniforms::new, src/geo/refraction.rs:121-130,
 * terrain_heightfield.rs:52-84.  Returns 0 ok; fills inv_two_r_prime and enabled. */
int f3do_earth_curvature(int earth_model, double lat_deg, double sphere_radius_m,
                         int refraction_model, double k, double pressure_mbar, double temperature_c,
                         double azimuth_deg, float* inv_two_r_prime, uint32_t* enabled);

/* AETHER aerial-perspective post (src/shaders/atmosphere/prometheus_aerial.wgsl:98-231) over a finished
 * accumulation: accum W*H*4 (rgb sum, frame count), depth W*H (frame-0 AOV), visibility W*H (0/1), out W*H*4
 * RGBA16F bit patterns (what textureStore writes).  f3do_aether_validate returns NULL or the reference's message. */
int f3do_aether_post(const f3do_atmosphere* atm, const f3do_aether_view* view, const float* accum, const float* depth,
                     const uint8_t* visibility, uint16_t* out_rgba16f);
const char* f3do_aether_validate(const f3do_atmosphere* atm);
float f3do_exp2(float x);   /* pinned exp2 (Cephes exp2f kernel) */

/* ---- smoke volume ray-march (src/smoke/render.rs) ---- */
typedef struct f3do_smoke_volume {
    uint32_t dims[3];                  /* x, y, z; voxel (x, y, z) at (z * dims[1] + y) * dims[0] + x (sampling.rs:83-85) */
    float voxel_size[3], origin[3];
    const float* density;              /* any field may be NULL = all zero */
    const float* temperature;
    const float* soot;
    const float* humidity;
    const float* emission_rate;
    const float* particle_age;
    uint64_t frame_index;
} f3do_smoke_volume;

typedef struct f3do_smoke_settings {   /* SmokeRenderSettings, src/smoke/types.rs:225-266 */
    float density_scale, extinction, scattering, absorption, phase_g, step_size;
    uint32_t max_steps;
    int32_t self_shadow;
    uint32_t shadow_steps;
    float shadow_step_size, jitter_strength, exposure;
    float thin_color[3], dense_color[3];
    float soot_absorption, fire_glow;
} f3do_smoke_settings;

/* 0 = ok, 1 = error (reference message via f3do_smoke_last_error).  rgba: height x width x 4, caller allocated. */
int f3do_smoke_raymarch_rgba(const f3do_smoke_volume* vol, const f3do_smoke_settings* settings, uint32_t width, uint32_t height,
                             const float camera_pos[3], const float target[3], const float up[3], float fovy_deg,
                             const float sun_direction[3], uint8_t* rgba);
/* straight-alpha over, python/forge3d/map_scene.py:1588-1604 (_alpha_composite_rgba), n pixels */
void f3do_composite_over_rgba(const uint8_t* bottom, const uint8_t* top, uint64_t n, uint8_t* out);
/* the layer composited over a terrain frame (python/forge3d/map_scene.py:1588-1604); base_depth nullable */
int f3do_smoke_raymarch_over_rgba(const f3do_smoke_volume* vol, const f3do_smoke_settings* settings, uint32_t width, uint32_t height,
                                  const float camera_pos[3], const float target[3], const float up[3], float fovy_deg,
                                  const float sun_direction[3], const uint8_t* base_rgba, const float* base_depth, uint8_t* rgba);
int f3do_smoke_raymarch_projection_rgba(const f3do_smoke_volume* vol, const f3do_smoke_settings* settings, uint32_t width,
                                        uint32_t height, const float view_direction[3], const float sun_direction[3], uint8_t* rgba);
float f3do_smoke_sun_transmittance(const f3do_smoke_volume* vol, const f3do_smoke_settings* settings, const float start[3],
                                   const float sun_dir[3], float step, uint32_t steps);
const char* f3do_smoke_last_error(void);

/* ---- HELIOS viewshed / solar shadow mask (src/terrain/analysis/viewshed.rs, src/shaders/terrain_viewshed.wgsl;
 * SURVEY section 8f row 4) ---- */
typedef struct f3do_viewshed_options {   /* ViewshedOptions, viewshed.rs:8-25, with physics_terms (:54-78) resolved */
    uint32_t width, height;
    float observer_x, observer_y, observer_height_m, target_height_m, max_distance_m;
    float observer_latitude_rad, observer_longitude_rad, left_unwrapped_deg, top_deg;
    float longitude_step_deg, latitude_step_deg, geodesic_sphere_radius_m;
    float physics[4];                    /* 1/meridional, 1/prime-vertical, 1 - k, curved flag */
} f3do_viewshed_options;
/* 0 ok.  physics_terms: earth 0 flat / 1 sphere / 2 ellipsoid; refraction 0 none / 1 bennett / 2 saemundsson / 3 effective_radius. */
int f3do_viewshed_physics(int earth_model, double latitude_deg, double sphere_radius_m, int refraction_model, double k,
                          double pressure_mbar, double temperature_c, float physics[4]);
/* positions_m: n x 2 (east, north metres from the observer); outputs n each; visible: 0 hidden, 1 visible, 2 = the geodesic
 * left the DEM footprint (the reference turns that into an error, viewshed.rs:320-327). */
int f3do_viewshed(const float* heights, const float* positions_m, const f3do_viewshed_options* options, uint8_t* visible,
                  float* curvature_drop_m, float* refraction_gain_m, float* horizon_distance_m);
/* inputs: n x 4 (latitude rad, longitude rad, sun azimuth rad, sun elevation rad); lit: n (1 = sun visible). */
int f3do_shadow_mask(const float* heights, const float* inputs, const f3do_viewshed_options* options, uint8_t* lit);

/* ---- LBVH build (src/accel/lbvh_gpu, src/shaders/lbvh_morton.wgsl, lbvh_link.wgsl) ----
 * literal_split = 1: the shader's find_split on 32-bit codes (midpoint for equal codes; boxes are not computed in this mode);
 * 0: split on the composite key (see f3d_lbvh_oracle.c).  pad_boxes = 1: leaf boxes carry the traversal's safety pad.
 * Outputs as f3d_lbvh_build (include/forge3d_b200.h). */
int f3do_lbvh_build(const float* xyz, uint32_t nverts, const uint32_t* idx, uint32_t ntris, int literal_split, int pad_boxes,
                    uint32_t* morton, uint32_t* order, uint32_t* left, uint32_t* right, uint32_t* parent, float* nodes);

/* ---- wavefront multi-bounce path tracer (src/path_tracing/wavefront, pt_*.wgsl, adjudication.rs) ----
 * Field for field the scene f3d_wavefront_scene of include/forge3d_b200.h describes. */
typedef struct {
    float cam_origin[3], cam_forward[3], cam_right[3], cam_up[3];
    float fov_y_rad, exposure;
    uint32_t seed_hi, seed_lo;
    const float* spheres; uint32_t nspheres;       /* 20 floats each: WavefrontGpuSphere, reference_scene.rs:101-117 */
    const float* dir_lights; uint32_t ndir;        /* 8 floats each: direction, intensity, colour, importance */
    const float* area_lights; uint32_t narea;      /* 12 floats each: position, radius, normal, intensity, colour, importance */
    const float* importance; uint32_t nimportance; /* per material slot */
    float environment[16];                         /* env_ground, env_sky, miss_ground, miss_sky (vec4 each) */
    const float* mesh_xyz; uint32_t mesh_nverts;
    const uint32_t* mesh_idx; uint32_t mesh_ntris;
    const float* instances; uint32_t ninstances;   /* 36 words each: object_to_world, world_to_object (column-major), blas, material, pad */
} f3do_wavefront_scene;
/* Adds frames [first_frame, first_frame + num_frames) into accum_io (W*H*4 floats, caller-zeroed before the first call); when
 * hdr_out / rgba8_out are given they receive mean = accum / spp_frames (alpha 1) and its Reinhard + sRGB resolve.
 * stats_out[19] = {rays traced, most rays in one frame, fewest iterations in one frame, rays at depth 0..15}. */
int f3do_wavefront_render(const f3do_wavefront_scene* scene, uint32_t width, uint32_t height, uint32_t spp_frames, uint32_t first_frame,
                          uint32_t num_frames, float* accum_io, float* hdr_out, uint8_t* rgba8_out, uint64_t* stats_out);
const char* f3do_wavefront_last_error(void);
void f3do_wavefront_sobol2(uint32_t i, float* x, float* y);
uint32_t f3do_wavefront_splitmix32(uint32_t x);
float f3do_log2(float x);          /* pinned log2 (Cephes log2f kernel), normal positive arguments */
float f3do_pow(float x, float y);  /* f3do_exp2(y * f3do_log2(x)) */

/* Pinned elementary functions of the numerics contract (exposed for unit tests). */
void  f3do_sincos(float x, float* s, float* c);
float f3do_atan2(float y, float x);
float f3do_acos(float x);
uint16_t f3do_f32_to_f16(float v);
float f3do_f16_to_f32(uint16_t h);

#ifdef __cplusplus
}
#endif
#endif
