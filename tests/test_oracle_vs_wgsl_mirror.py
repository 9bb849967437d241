"""The C oracle against a second, independent restatement of the WGSL (tests/_wgsl_mirror.py).

The golden image pins the oracle statistically (SSIM / mean-abs); this pins it ARITHMETICALLY: two separate
transcriptions of the same shaders, both under the numerics contract, must agree bit for bit on the accumulation
buffer and the depth AOV of a tiny scene - including the ReSTIR reuse chain (temporal + spatial, several frames),
curved-earth sun rays and multi-sample frames.  CPU only.
"""
import numpy as np
import pytest

import _wgsl_mirror as mirror
from oracle import oracle


def _bumpy_dem(rows, cols, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:rows, 0:cols].astype(np.float32)
    dem = 3.0 * np.sin(0.9 * xx) * np.cos(0.7 * yy) + 0.35 * xx + rng.random((rows, cols), dtype=np.float32) * 1.5
    return dem.astype(np.float32)


CASES = {
    # name: (dem rows, cols, W, H, spp, frames, earth_model, sphere_radius, sun az, sun el, seed)
    "flat_multi_frame": (6, 7, 11, 9, 1, 4, "flat", 6371008.8, 300.0, 35.0, 7),
    "curved_low_sun_spp2": (9, 5, 10, 8, 2, 3, "sphere", 900.0, 110.0, 12.0, 12345),
    "ragged_pow2_edge": (5, 10, 8, 8, 1, 3, "flat", 6371008.8, 20.0, 60.0, 0xFFFFFFFF),
}


def _mesh_and_env():
    """A slanted quad and a small pyramid floating over the DEM (visible, shadow-casting, partly behind terrain) and a
    smooth 8x4 equirect environment."""
    verts = np.array([[-3.0, 9.0, -2.0], [3.5, 9.5, -2.5], [3.0, 7.0, 3.0], [-3.5, 7.5, 2.5],
                      [5.0, 4.0, 6.0], [8.0, 4.0, 6.0], [6.5, 4.0, 9.0], [6.5, 8.0, 7.0]], np.float32)
    tris = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 7], [5, 6, 7], [6, 4, 7], [4, 6, 5]], np.uint32)
    yy, xx = np.mgrid[0:4, 0:8].astype(np.float32)
    env = np.stack([0.6 + 0.4 * np.sin(xx), 0.5 + 0.1 * yy, 0.9 - 0.08 * xx], axis=-1).astype(np.float32)
    return verts, tris, env


@pytest.mark.parametrize("name", sorted(CASES) + ["mesh_and_env_map"])
def test_oracle_matches_independent_wgsl_restatement(name):
    extra = {}
    if name == "mesh_and_env_map":
        verts, tris, env = _mesh_and_env()
        extra = dict(mesh_vertices=verts, mesh_indices=tris, env_map=env)
        name = "flat_multi_frame"
    rows, cols, W, H, spp, frames, earth, radius, az, el, seed = CASES[name]
    dem = _bumpy_dem(rows, cols, seed=rows * 100 + cols)
    spacing = (2.0, 3.0)
    cam = dict(origin=(1.5, 14.0, 22.0), look_at=(0.5, 1.0, 0.0), up=(0.0, 1.0, 0.0), fov_y=50.0, exposure=1.3)
    common = dict(spacing=spacing, exaggeration=1.25, albedo=(0.55, 0.6, 0.45), sun_azimuth_deg=az,
                  sun_elevation_deg=el, sun_intensity=2.0, sun_color=(1.0, 0.9, 0.8), env_intensity=0.4,
                  spp=spp, seed=seed)
    k, enabled = oracle.earth_curvature(earth, 0.0, radius, "none", 0.13, 1013.25, 15.0, azimuth_deg=az)
    want = oracle.render(dem, W, H, cam, max_frames=frames, min_frames=frames, variance_threshold=1e30,
                         earth_model=earth, sphere_radius_m=radius, refraction_model="none", want_accum=True, **common, **extra)
    assert want["frames"] == frames
    aovs = {}
    accum, depth = mirror.render(dem, W, H, cam, frames=frames, inv_two_r_prime=k, curvature_enabled=enabled, aovs=aovs, **common, **extra)

    hits = np.isfinite(depth)
    assert hits.any() and (~hits).any(), "the scene must contain terrain and sky pixels"
    if earth != "flat":
        assert enabled and k > 0.0
    np.testing.assert_array_equal(np.isnan(want["depth"]), ~hits)
    np.testing.assert_array_equal(want["depth"].view(np.uint32)[hits], depth.view(np.uint32)[hits])
    np.testing.assert_array_equal(want["accum"].view(np.uint32), accum.view(np.uint32))
    # the resolved image and the f16 AOVs (tonemap, rgba16float round trip, u8 quantisation)
    np.testing.assert_array_equal(want["rgba"], aovs["rgba"])
    np.testing.assert_array_equal(want["normal"].view(np.uint32), aovs["normal"].view(np.uint32))
    np.testing.assert_array_equal(want["albedo"].view(np.uint32), aovs["albedo"].view(np.uint32))
    if extra:   # mesh pixels exist (their albedo AOV is the legacy constant) and the sky is not constant
        assert (np.abs(aovs["albedo"][..., 2] - 0.8) < 1e-3).sum() >= 4
        assert np.unique(accum[~hits][:, 0]).size > 1
    # the chain was exercised: shading differs between pixels and frames actually accumulated
    assert float(accum[..., 3].min()) == frames and np.unique(accum[..., 0]).size > 8


def test_mirror_pyramid_equals_oracle_pyramid():
    dem = _bumpy_dem(7, 12, seed=3)
    S = mirror.Scene(dem, (1.0, 1.0), 1.0, (0.5, 0.5, 0.5), 0.3, 0.0, False)
    levels, cw, ch = oracle.build_minmax(dem)
    assert (cw, ch) == (S.cw, S.ch) and len(levels) == S.mips
    for a, b in zip(levels, S.levels):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("any_hit,apply_curvature", [(False, False), (True, True), (True, False), (False, True)])
def test_ray_level_agreement_including_curved_rays(any_hit, apply_curvature):
    """terrain_trace ray by ray: hit flag, t and normal bit-identical; strong curvature so that the quadratic
    ray-height term decides a good share of the outcomes (checked: curved and straight results differ)."""
    dem = _bumpy_dem(11, 14, seed=77)
    spacing, ex, k = (2.0, 3.0), 1.5, 4.0e-3
    S = mirror.Scene(dem, spacing, ex, (0.5, 0.5, 0.5), 0.3, k, True)
    rng = np.random.default_rng(5)
    n = 400
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0] = rng.uniform(-16, 16, n)
    rays[:, 1] = rng.uniform(-2, 14, n)
    rays[:, 2] = rng.uniform(-18, 18, n)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[:, 1] = rng.uniform(-0.6, 0.25, n)                       # mostly grazing, up and down
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 3], rays[:, 4:7], rays[:, 7] = 1e-3, d, 1e30
    rays[:40, 7] = rng.uniform(2, 20, 40)                      # bounded rays too
    hit, t, nrm = oracle.trace_rays(dem, spacing, (float(S.ox), float(S.oz)), ex, rays, any_hit=any_hit,
                                    apply_curvature=apply_curvature, inv_two_r_prime=k, curvature_enabled=True)
    straight_hits = 0
    for i in range(n):
        o = tuple(mirror.f(v) for v in rays[i, 0:3])
        dd = tuple(mirror.f(v) for v in rays[i, 4:7])
        h, tt, p, nn = S.trace(o, dd, mirror.f(rays[i, 3]), mirror.f(rays[i, 7]), any_hit, apply_curvature)
        assert bool(hit[i]) == h, i
        if h:
            assert np.float32(tt).view(np.uint32) == t[i].view(np.uint32), i
            np.testing.assert_array_equal(np.array(nn, np.float32).view(np.uint32), nrm[i].view(np.uint32))
        if apply_curvature:
            h0, t0, *_ = S.trace(o, dd, mirror.f(rays[i, 3]), mirror.f(rays[i, 7]), any_hit, False)
            straight_hits += int(h0 != h or (h and t0 != tt))
    assert 40 < int(hit.sum()) < n - 40
    if apply_curvature:
        assert straight_hits >= 8, "curvature must change outcomes in this fixture"
