/* tests/c/abi_smoke.c -- a plain-C consumer of include/forge3d_b200.h (no Python, no C++, no torch):
 * what the reference's Rust `extern "C"` shim would do at terrain_reference.rs:416-417.
 * Exit code 0 = rendered (prints frames/variance), 3 = F3D_ERR_DEVICE (no GPU: the library must fail
 * loudly, not fall back), anything else = unexpected. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "forge3d_b200.h"

int main(void) {
    enum { DW = 16, DH = 16, W = 32, H = 24 };
    static float dem[DW * DH];
    for (int i = 0; i < DW * DH; i++) dem[i] = 0.1f * (float)((i % DW) + (i / DW) % 3);
    static unsigned char rgba[W * H * 4];
    static float albedo[W * H * 3], normal[W * H * 3], depth[W * H];
    f3d_terrain_desc d;
    memset(&d, 0, sizeof d);
    d.heights = dem; d.dem_w = DW; d.dem_h = DH;
    d.spacing[0] = d.spacing[1] = 5.0f; d.exaggeration = 4.0f;
    d.albedo[0] = d.albedo[1] = d.albedo[2] = 0.6f;
    d.cam_origin[0] = 0.0f; d.cam_origin[1] = 50.0f; d.cam_origin[2] = 120.0f;
    d.cam_up[1] = 1.0f; d.fov_y_deg = 45.0f; d.exposure = 1.0f;
    d.sun_az_deg = 315.0f; d.sun_el_deg = 45.0f; d.sun_intensity = 2.5f;
    d.sun_color[0] = 1.0f; d.sun_color[1] = 0.97f; d.sun_color[2] = 0.92f;
    d.earth_model = F3D_EARTH_ELLIPSOID; d.sphere_radius_m = 6371008.8;
    d.refraction_model = F3D_REFRACTION_BENNETT; d.refraction_k = 0.13; d.pressure_mbar = 1013.25; d.temperature_c = 15.0;
    d.env_intensity = 0.35f;
    d.width = W; d.height = H; d.seed = 7; d.spp = 1; d.max_frames = 4; d.min_frames = 2; d.variance_threshold = 1e30f;
    f3d_terrain_out o;
    memset(&o, 0, sizeof o);
    o.rgba = rgba; o.albedo = albedo; o.normal = normal; o.depth = depth;
    if (f3d_abi_version() != F3D_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 100; }
    int rc = f3d_terrain_reference_render(&d, &o);
    if (rc != 0) { printf("rc=%d error=%s\n", rc, f3d_last_error()); return rc; }
    printf("rc=0 frames=%u variance=%g converged=%d rays=%llu alpha=%u\n", o.frames, (double)o.variance, o.converged,
           (unsigned long long)(o.rays_primary + o.rays_shadow + o.rays_ibl), (unsigned)rgba[3]);
    return (o.frames == 4 && rgba[3] == 255) ? 0 : 101;
}
