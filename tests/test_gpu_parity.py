"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Everything here is exact: the numerics contract makes oracle and kernels agree
bit for bit, so the comparisons are np.array_equal, not tolerances.  The only statistical check is
the reference's own golden-image drift gate."""
import numpy as np
import pytest

import _helpers as H
from forge3d_b200 import _native, hybrid_render_terrain_reference
from oracle import oracle

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _assert_same_render(got, ref, label=""):
    assert got["frames"] == ref["frames"], label
    assert np.array_equal(_bits(got["accum"]), _bits(ref["accum"])), f"{label}: accumulation differs"
    assert np.array_equal(got["rgba"], ref["rgba"]), f"{label}: rgba differs"
    assert np.array_equal(_bits(got["albedo"]), _bits(ref["albedo"])), f"{label}: albedo AOV differs"
    assert np.array_equal(_bits(got["normal"]), _bits(ref["normal"])), f"{label}: normal AOV differs"
    assert np.array_equal(_bits(got["depth"]), _bits(ref["depth"])), f"{label}: depth AOV differs"
    # nodes_popped is not compared: the production traversal culls children at push time, so it
    # pops fewer nodes than the literal loop the oracle counts (same visit order, same hits)
    for key in ("rays_primary", "rays_shadow", "rays_ibl"):
        assert got[key] == ref[key], (label, key, got[key], ref[key])
    assert 0 < got["nodes_popped"] <= ref["nodes_popped"], (label, got["nodes_popped"], ref["nodes_popped"])
    assert np.float32(got["variance"]) == np.float32(ref["variance"]), label


def _both(dem, w, h, cam, **kw):
    got = _native.hybrid_render_terrain_reference(dem, w, h, cam, **kw, want_accum=True)
    ref = oracle.render(dem, w, h, cam, **kw, want_accum=True)
    return got, ref


@pytest.mark.parametrize("shape", [(256, 256), (37, 100), (2, 2), (3, 9), (129, 65)])
def test_pyramid_bit_exact(shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    dem = (rng.standard_normal(shape) * 100).astype(np.float32)
    g_levels, gcw, gch = _native.build_minmax(dem)
    o_levels, ocw, och = oracle.build_minmax(dem)
    assert (gcw, gch) == (ocw, och) and len(g_levels) == len(o_levels)
    for g, o in zip(g_levels, o_levels):
        assert g.shape == o.shape and np.array_equal(_bits(g), _bits(o))


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("any_hit,curv", [(True, True), (True, False), (False, False), (False, True)])
def test_kat_rays_bit_exact(any_hit, curv, variant):
    # variant 0 = production traversal (push-time culling, smem stack), 1 = literal WGSL loop
    h = H.curvature_fixture()
    arb, mask = H.kat_rays(h)
    rays = np.concatenate([arb, mask[::3]])
    kw = dict(any_hit=any_hit, apply_curvature=curv, inv_two_r_prime=float(H.PROOF_INV_TWO_R), curvature_enabled=True)
    gh, gt, gn = _native.trace_rays(h, (500.0, 500.0), (0.0, 0.0), 1.0, rays, variant=variant, **kw)
    oh, ot, on = oracle.trace_rays(h, (500.0, 500.0), (0.0, 0.0), 1.0, rays, **kw)
    assert np.array_equal(gh, oh)
    if any_hit and variant == 0:
        # the production any-hit order is the sign order (f3d_trace_fast.cuh, F3D_ANYHIT_SIGN_ORDER): the occlusion
        # flag is order-independent; WHICH hit is found first may differ from the sort order in exact key ties only
        same = _bits(gt) == _bits(ot)
        assert same[gh].mean() > 0.999
        assert np.array_equal(_bits(gn)[same & gh], _bits(on)[same & gh])
    else:
        assert np.array_equal(_bits(gt), _bits(ot))
        assert np.array_equal(_bits(gn), _bits(on))
    if any_hit and curv:  # the reference's KAT thresholds on the GPU path itself
        brute = H.brute_2d_hit(h, arb)
        assert int((brute & ~gh[:10_000]).sum()) == 0
        assert int((~brute & gh[:10_000]).sum()) / 10_000.0 < 0.001


def test_production_traversal_matches_literal_on_random_rays():
    # 400k random rays over a rough DEM with ragged (non power-of-two, non-square) cell counts:
    # closest-hit and any-hit, with and without curvature; the production traversal must return the
    # same hit flags, distances and normals as the literal loop while popping fewer nodes.
    rng = np.random.default_rng(11)
    w, h = 300, 173
    dem = (rng.standard_normal((h, w)).cumsum(0).cumsum(1) * 0.05 + rng.standard_normal((h, w)) * 2.0).astype(np.float32)
    n = 400_000
    rays = np.zeros((n, 8), np.float32)
    ext = np.array([(w - 1) * 7.5, (h - 1) * 7.5])
    rays[:, 0] = rng.uniform(-0.2 * ext[0], 1.2 * ext[0], n)
    rays[:, 2] = rng.uniform(-0.2 * ext[1], 1.2 * ext[1], n)
    rays[:, 1] = rng.uniform(dem.min() - 5, dem.max() + 60, n)
    d = rng.standard_normal((n, 3)); d[:, 1] *= 0.35
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 4:7] = d
    rays[::17, 4] = 0.0          # axis-parallel rays exercise terrain_safe_inv
    rays[::23, 6] = 0.0
    rays[:, 3] = 1e-3
    rays[:, 7] = np.where(rng.uniform(size=n) < 0.3, rng.uniform(50, 3000, n), 1e30)
    for any_hit, curv in [(False, False), (True, False), (True, True)]:
        kw = dict(any_hit=any_hit, apply_curvature=curv, inv_two_r_prime=3e-6, curvature_enabled=True)
        a = _native.trace_rays(dem, (7.5, 7.5), (0.0, 0.0), 1.3, rays, variant=0, want_nodes=True, **kw)
        b = _native.trace_rays(dem, (7.5, 7.5), (0.0, 0.0), 1.3, rays, variant=1, want_nodes=True, **kw)
        assert np.array_equal(a[0], b[0]) and 0.05 < a[0].mean() < 0.95
        if any_hit:   # sign-ordered any-hit rays: same flags; the first hit found may differ in exact key ties only
            same = _bits(a[1]) == _bits(b[1])
            assert same[a[0]].mean() > 0.999
            assert np.array_equal(_bits(a[2])[same & a[0]], _bits(b[2])[same & a[0]])
        else:
            assert np.array_equal(_bits(a[1]), _bits(b[1]))
            assert np.array_equal(_bits(a[2]), _bits(b[2]))
        assert a[3] < b[3]
    o = oracle.trace_rays(dem, (7.5, 7.5), (0.0, 0.0), 1.3, rays[:50_000], any_hit=False, apply_curvature=False)
    g = _native.trace_rays(dem, (7.5, 7.5), (0.0, 0.0), 1.3, rays[:50_000], any_hit=False, apply_curvature=False)
    assert np.array_equal(g[0], o[0]) and np.array_equal(_bits(g[1]), _bits(o[1])) and np.array_equal(_bits(g[2]), _bits(o[2]))


def test_render_bit_exact_small():
    dem = H.golden_dem()
    kw = {**H.scene_kwargs(dem), "max_frames": 8, "min_frames": 8, "variance_threshold": 1e30}
    got, ref = _both(dem, 64, 64, H.CAM, **kw)
    _assert_same_render(got, ref, "64x64x8")
    # non-multiple-of-tile image, non-square
    got, ref = _both(dem, 75, 41, H.CAM, **kw)
    _assert_same_render(got, ref, "75x41x8")


@pytest.mark.parametrize("shape,size", [((2, 2), (33, 20)), ((3, 9), (40, 24)), ((37, 100), (64, 48)), ((129, 65), (48, 48))])
def test_render_bit_exact_tiny_and_ragged_dems(shape, size):
    # smallest legal DEM (one cell = a 1-level pyramid whose root is a leaf), ragged non-power-of-two cell
    # counts and non-square DEMs: exercises the sentinel padding and the collapsed-axis levels
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    dem = rng.uniform(0.0, 1.0, shape).astype(np.float32)
    span = 80.0
    kw = dict(spacing=(span / max(shape[1] - 1, 1), span / max(shape[0] - 1, 1)), exaggeration=12.0, albedo=H.ALBEDO,
              sun_azimuth_deg=140.0, sun_elevation_deg=30.0, max_frames=5, min_frames=5, variance_threshold=1e30, spp=2)
    got, ref = _both(dem, size[0], size[1], H.CAM, **kw)
    _assert_same_render(got, ref, f"dem {shape}")
    assert np.isfinite(got["depth"]).any()


def test_native_seam_error_paths_on_gpu():
    # errors raised below the Python facade (device-side non-finite scan, C++ validate_desc)
    bad = np.full((16, 16), np.nan, np.float32)
    with pytest.raises(RuntimeError, match="non-finite"):
        _native.hybrid_render_terrain_reference(bad, 32, 32, H.CAM, max_frames=4, min_frames=2)
    dem = H.golden_dem()
    with pytest.raises(RuntimeError, match="at least 2x2"):
        _native.hybrid_render_terrain_reference(np.zeros((1, 5), np.float32), 32, 32, H.CAM, max_frames=4, min_frames=2)
    with pytest.raises(RuntimeError, match="min_frames"):
        _native.hybrid_render_terrain_reference(dem, 32, 32, H.CAM, max_frames=4, min_frames=8)
    with pytest.raises(RuntimeError, match="spp"):
        _native.hybrid_render_terrain_reference(dem, 32, 32, H.CAM, max_frames=4, min_frames=2, spp=65)
    with pytest.raises(RuntimeError, match="exaggeration"):
        _native.hybrid_render_terrain_reference(dem, 32, 32, H.CAM, max_frames=4, min_frames=2, exaggeration=0.0)
    with pytest.raises(RuntimeError, match="flat earth only supports"):
        _native.hybrid_render_terrain_reference(dem, 32, 32, H.CAM, max_frames=4, min_frames=2, earth_model="flat")
    with pytest.raises(RuntimeError, match="out-of-bounds"):
        _native.hybrid_render_terrain_reference(dem, 32, 32, H.CAM, max_frames=4, min_frames=2,
                                                mesh_vertices=np.zeros((3, 3), np.float32),
                                                mesh_indices=np.array([[0, 1, 7]], np.uint32))
    # a session survives an error in a previous call (no poisoned global state)
    ok = _native.hybrid_render_terrain_reference(dem, 32, 32, H.CAM, **{**H.scene_kwargs(dem), "max_frames": 4, "min_frames": 2,
                                                                        "variance_threshold": 1e30})
    assert ok["frames"] == 4


def test_curved_and_flat_earth_models_share_the_memory_footprint():
    # reference: tests/test_shadow_tip.py:248-430 (HELIOS memory gate): the curved-earth shadow policy must not
    # cost memory relative to the flat baseline; both models render the same workload.
    dem = H.golden_dem()[::2, ::2].copy()
    spacing = 100.0 / (dem.shape[1] - 1)
    kw = dict(spacing=(spacing, spacing), exaggeration=20.0, albedo=H.ALBEDO, sun_azimuth_deg=225.0, sun_elevation_deg=35.0,
              observer_latitude_deg=46.5, observer_longitude_deg=7.5, spp=1, min_frames=32, max_frames=32,
              variance_threshold=1e9, seed=7)
    res = {}
    for mode, (earth, refr) in {"flat_baseline": ("flat", "none"), "helios": ("ellipsoid", "effective_radius")}.items():
        got, ref = _both(dem, 64, 64, H.CAM, **kw, earth_model=earth, refraction_model=refr)
        _assert_same_render(got, ref, mode)
        res[mode] = got
    for key in ("peak_host_visible_bytes", "gpu_resource_bytes", "minmax_pyramid_bytes"):
        assert res["flat_baseline"][key] == res["helios"][key] > 0


def test_plain_c_consumer_renders_on_gpu():
    import subprocess

    from test_abi import _build_c_consumer

    res = subprocess.run([str(_build_c_consumer())], capture_output=True, text=True)
    assert res.returncode == 0 and "frames=4" in res.stdout and "alpha=255" in res.stdout, res.stdout + res.stderr


def test_session_api_matches_one_call_path():
    from forge3d_b200.session import Session

    dem = H.golden_dem()
    kw = {**H.scene_kwargs(dem), "max_frames": 64, "min_frames": 64, "variance_threshold": 1e30}
    one = _native.hybrid_render_terrain_reference(dem, 80, 60, H.CAM, **kw, want_accum=True)
    with Session(dem, 80, 60, H.CAM, **kw) as s:
        s.render_frames(20)
        s.render_frames(12)
        v32, bad = s.variance()          # the gate value at the first window boundary
        s.render_frames(32)
        v64, bad2 = s.variance()
        out = s.resolve_host(want_accum=True)
        assert s.frames == 64 and not bad and not bad2
    assert np.array_equal(_bits(out["accum"]), _bits(one["accum"])) and np.array_equal(out["rgba"], one["rgba"])
    assert np.float32(v64) == np.float32(one["variance"]) and v32 > v64


def test_render_bit_exact_spp_and_window():
    dem = H.golden_dem()
    kw = {**H.scene_kwargs(dem), "max_frames": 40, "min_frames": 40, "variance_threshold": 1e30, "spp": 3}
    got, ref = _both(dem, 96, 64, H.CAM, **kw)   # crosses a Welford window boundary (frame 32)
    _assert_same_render(got, ref, "96x64 spp3 x40")


def test_render_bit_exact_sine_dem_flat_earth():
    dem = H.sine_dem(128)   # BASELINE.json configs[0] DEM
    kw = dict(spacing=(100.0 / 127, 100.0 / 127), exaggeration=20.0, albedo=H.ALBEDO, sun_azimuth_deg=315.0,
              sun_elevation_deg=12.0, max_frames=6, min_frames=6, variance_threshold=1e30, earth_model="flat",
              refraction_model="none")
    got, ref = _both(dem, 128, 128, H.CAM, **kw)
    _assert_same_render(got, ref, "sine flat-earth")


def test_render_bit_exact_mesh_and_env():
    dem = H.golden_dem()
    quad_v = np.array([[-18.0, 22.0, -6.0], [18.0, 22.0, -6.0], [18.0, 40.0, -6.0], [-18.0, 40.0, -6.0]], np.float32)
    quad_i = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    rng = np.random.default_rng(5)
    env = rng.uniform(0.0, 2.0, (16, 32, 3)).astype(np.float32)
    kw = {**H.scene_kwargs(dem), "max_frames": 6, "min_frames": 6, "variance_threshold": 1e30}
    got, ref = _both(dem, 96, 96, H.CAM, **kw, mesh_vertices=quad_v, mesh_indices=quad_i)
    _assert_same_render(got, ref, "mesh")
    closer = np.isfinite(got["depth"]) & (got["albedo"][..., 2] > 0.75)
    assert closer.mean() > 0.01     # the quad is visible and carries the mesh albedo (0.7,0.7,0.8)
    got, ref = _both(dem, 96, 96, H.CAM, **kw, env_map=env)
    _assert_same_render(got, ref, "env map")
    got, ref = _both(dem, 64, 64, H.CAM, **kw, env_map=env, mesh_vertices=quad_v, mesh_indices=quad_i)
    _assert_same_render(got, ref, "mesh + env map")


def test_mesh_bvh_matches_index_order_sweep(monkeypatch):
    # 1500 random triangles (plus exact duplicates to force closest-hit ties) hovering over the DEM: the
    # GPU walks a BVH, the oracle sweeps every triangle in index order like hybrid_traversal.wgsl:137-172.
    dem = H.golden_dem()
    rng = np.random.default_rng(21)
    ntri = 1500
    centers = np.stack([rng.uniform(-45, 45, ntri), rng.uniform(12, 45, ntri), rng.uniform(-45, 45, ntri)], 1)
    verts = (centers[:, None, :] + rng.normal(0, 2.5, (ntri, 3, 3))).reshape(-1, 3).astype(np.float32)
    idx = np.arange(ntri * 3, dtype=np.uint32).reshape(ntri, 3)
    idx = np.concatenate([idx, idx[100:140]])            # duplicated triangles: ties must pick the lower index
    kw = {**H.scene_kwargs(dem), "max_frames": 4, "min_frames": 4, "variance_threshold": 1e30}
    got, ref = _both(dem, 128, 96, H.CAM, **kw, mesh_vertices=verts, mesh_indices=idx)
    _assert_same_render(got, ref, "mesh bvh vs sweep")
    mesh_px = np.isfinite(got["depth"]) & (got["albedo"][..., 2] > 0.75)
    assert mesh_px.mean() > 0.05
    monkeypatch.setenv("F3D_B200_NO_MESH_BVH", "1")
    sweep = _native.hybrid_render_terrain_reference(dem, 128, 96, H.CAM, **kw, mesh_vertices=verts, mesh_indices=idx,
                                                    want_accum=True)
    assert np.array_equal(_bits(sweep["accum"]), _bits(got["accum"]))


def test_render_bit_exact_large_relief_dem():
    # Rainier-shaped closed-form DEM at reduced size: metres-scale coordinates,
    # curvature active on the sun rays, deep pyramid (10 levels), orbit camera outside the DEM.
    n = 512
    dem = H.rainier_dem(n)
    spacing = 10.0 * 2048 / n
    cam = H.rainier_camera(n, spacing, dem)
    kw = dict(spacing=(spacing, spacing), exaggeration=1.0, albedo=H.ALBEDO, sun_azimuth_deg=302.0,
              sun_elevation_deg=24.0, max_frames=4, min_frames=4, variance_threshold=1e30)
    got, ref = _both(dem, 160, 90, cam, **kw)
    _assert_same_render(got, ref, "rainier 512")
    assert np.isfinite(got["depth"]).mean() > 0.2


@pytest.mark.parametrize("az,el,earth", [(20.0, 35.0, "flat"), (110.0, 12.0, "wgs84"), (200.0, 60.0, "flat"), (290.0, 3.0, "wgs84"),
                                         (45.0, 24.0, "flat"), (135.0, 0.5, "flat"), (180.0, 45.0, "wgs84"), (270.0, 89.97, "flat"),
                                         (0.0, 20.0, "wgs84"), (90.0, 0.0, "flat"), (315.0, 8.0, "flat")])
def test_sun_horizon_strips_are_exact_in_every_octant(az, el, earth, monkeypatch):
    """The sun horizon strips (csrc/f3d_trace_fast.cuh SunHorizon) skip parts of a sun ray's traversal; the render must stay
    bit-identical to the oracle for every major axis / travel direction / slope sign of the strips (8 octants + the axes),
    for grazing, ordinary and (almost) vertical suns, over a rough, ragged, non-square DEM - and the strips must really cull."""
    rng = np.random.default_rng(int(az * 7 + el * 13))
    shape = (83, 121)
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    dem = (0.6 * np.sin(xx / 9.0) * np.cos(yy / 7.0) + 0.4 * np.sin((xx + 2 * yy) / 17.0) + 0.35 * rng.uniform(0.0, 1.0, shape)).astype(np.float32)
    span = 120.0
    kw = dict(spacing=(span / (shape[1] - 1), 0.8 * span / (shape[0] - 1)), exaggeration=14.0, albedo=H.ALBEDO, sun_azimuth_deg=az,
              sun_elevation_deg=el, earth_model=earth, refraction_model="none" if earth == "flat" else "bennett", max_frames=4, min_frames=4,
              variance_threshold=1e30)
    got, ref = _both(dem, 96, 64, H.CAM, **kw)
    _assert_same_render(got, ref, f"sun az {az} el {el} {earth}")
    monkeypatch.setenv("F3D_B200_SUN_HORIZON", "0")
    off = _native.hybrid_render_terrain_reference(dem, 96, 64, H.CAM, **kw, want_accum=True)
    assert np.array_equal(_bits(off["accum"]), _bits(got["accum"]))
    # node counts of any-hit rays depend on the scheduling (how many nodes a ray expands before a queued leaf reports its hit):
    # equal paths can differ by a fraction of a percent from run to run
    assert got["nodes_popped"] <= 1.01 * off["nodes_popped"]
    if 1.0 <= el <= 60.0:
        assert got["nodes_popped"] < off["nodes_popped"], "the strips cleared nothing"


@pytest.mark.parametrize("shape,relief,earth", [((83, 121), 14.0, "flat"), ((64, 64), 40.0, "wgs84"), ((150, 97), 5.0, "flat"), ((33, 200), 25.0, "flat")])
def test_escape_map_is_exact_and_culls(shape, relief, earth, monkeypatch):
    """The escape map (csrc/f3d_trace_fast.cuh EscapeMap) lets an IBL ray whose slope clears every far cell of its octant skip the
    bottom-up start and k_trace; only the near block is walked.  The render must stay bit-identical to the oracle over rough,
    ragged, non-square DEMs (every octant is hit by the hemisphere samples), and the map must really cull."""
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    dem = (0.5 * np.sin(xx / 11.0) * np.cos(yy / 6.0) + 0.5 * np.cos((xx - yy) / 13.0) + 0.3 * rng.uniform(0.0, 1.0, shape)).astype(np.float32)
    span = 120.0
    kw = dict(spacing=(span / (shape[1] - 1), 1.3 * span / (shape[0] - 1)), exaggeration=relief, albedo=H.ALBEDO, sun_azimuth_deg=140.0,
              sun_elevation_deg=28.0, earth_model=earth, refraction_model="none" if earth == "flat" else "bennett", max_frames=4, min_frames=4,
              variance_threshold=1e30)
    monkeypatch.setenv("F3D_B200_ESCAPE", "1")          # opt-in: measured not to pay for its build below ~500 frames
    got, ref = _both(dem, 112, 80, H.CAM, **kw)
    _assert_same_render(got, ref, f"escape map {shape} relief {relief} {earth}")
    monkeypatch.setenv("F3D_B200_ESCAPE", "0")
    off = _native.hybrid_render_terrain_reference(dem, 112, 80, H.CAM, **kw, want_accum=True)
    assert np.array_equal(_bits(off["accum"]), _bits(got["accum"]))
    # node counts of any-hit rays vary a little with the scheduling; on the narrow DEM the near block is most of the terrain
    assert got["nodes_popped"] <= 1.01 * off["nodes_popped"]
    if min(shape) >= 64:
        assert got["nodes_popped"] < off["nodes_popped"], "the escape map cleared nothing"


class _variant_library:
    """Routes _native through variants/lib_<name>.so (a compile-time variant of the same sources) for the duration."""

    def __init__(self, name):
        from forge3d_b200 import build as b

        self.path = b.build_variant(name)        # prebuilt by __graft_entry__.build(); rebuilt here only if stale

    def __enter__(self):
        self._saved = (_native.LIB_PATH, dict(_native._libs))
        _native.LIB_PATH = self.path
        _native._libs.clear()
        return _native.lib()

    def __exit__(self, *exc):
        _native.LIB_PATH = self._saved[0]
        _native._libs.clear()
        _native._libs.update(self._saved[1])
        return False


def test_compile_time_variants_are_bit_identical():
    """variants/lib_tma.so stages the top pyramid levels into shared memory with cp.async.bulk + mbarrier (the north-star's TMA
    staging; measured SLOWER than the L1-cached loads, csrc/f3d_trace_fast.cuh F3D_TMA_STAGE, so not the default): it must render
    exactly what the default library renders - which the other tests pin to the oracle."""
    dem = H.golden_dem()
    kw = {**H.scene_kwargs(dem), "max_frames": 6, "min_frames": 6, "variance_threshold": 1e30}
    want = _native.hybrid_render_terrain_reference(dem, 160, 96, H.CAM, **kw, want_accum=True)
    with _variant_library("tma") as L:
        assert b"F3D_TMA_STAGE=1" in L.f3d_build_info()
        got = _native.hybrid_render_terrain_reference(dem, 160, 96, H.CAM, **kw, want_accum=True)
    assert np.array_equal(_bits(got["accum"]), _bits(want["accum"]))
    assert np.array_equal(_bits(got["depth"]), _bits(want["depth"])) and np.array_equal(got["rgba"], want["rgba"])


def test_throughput_numerics_build_is_bounded_but_not_within_tolerance():
    """libforge3d_b200_fast.so (SFU division / sqrt, FMA contraction: csrc/f3d_math.cuh F3D_FAST_NUMERICS) is an EXPERIMENT, not a
    product mode: measured on the B200 it differs from the oracle by RGBA RMSE 7.8e-3 on this scene (9.4e-3 on C2 at 256 spp),
    4 x the difference between two exact renders with different seeds, and its image mean is 0.5 % of full scale darker, so it meets neither BASELINE.json's same-seed tolerance
    (1e-3) nor Monte-Carlo equivalence.  This test only bounds the damage (same geometry, no gross bias) so the build keeps
    working as the A/B arm of the numerics cost; the default library is the bit-exact one."""
    dem = H.golden_dem()
    kw = {**H.scene_kwargs(dem), "max_frames": 64, "min_frames": 64, "variance_threshold": 1e30}
    ref = oracle.render(dem, 128, 128, H.CAM, **kw)
    from forge3d_b200.session import Session

    s = Session(dem, 128, 128, H.CAM, numerics="fast", **kw)
    s.render_frames(64)
    fast = s.resolve_host()
    s.close()
    d = (fast["rgba"][..., :3].astype(np.float64) - ref["rgba"][..., :3].astype(np.float64)) / 255.0
    assert float(np.sqrt(np.mean(d * d))) <= 2e-2
    assert np.abs(d.mean(axis=(0, 1))).max() <= 1e-2          # measured: 5.0e-3 darker (more occlusion flags flip to 'hit' than away from it)
    assert np.array_equal(fast["rgba"][..., 3], ref["rgba"][..., 3])
    hit_f, hit_r = np.isfinite(fast["depth"]) & (fast["depth"] < 1e29), np.isfinite(ref["depth"]) & (ref["depth"] < 1e29)
    assert (hit_f != hit_r).mean() <= 1e-3, "hit / miss classification differs on more than a few silhouette pixels"
    both = hit_f & hit_r
    assert np.abs(fast["depth"][both] - ref["depth"][both]).max() <= 1e-3 * np.abs(ref["depth"][both]).max()


def test_full_size_config3_bit_exact():
    # BASELINE.json configs[2] shape at full size: 4096x4096 DEM (13-level pyramid), 3840x2160 image, a few
    # samples: ~60 M rays through 67 M DEM cells, compared bit for bit with the oracle (host cores, ~10-20 s).
    n = 4096
    dem = H.rainier_dem(n)
    spacing = 10.0
    cam = H.rainier_camera(n, spacing, dem)
    kw = dict(spacing=(spacing, spacing), exaggeration=1.0, albedo=H.ALBEDO, sun_azimuth_deg=302.0,
              sun_elevation_deg=24.0, max_frames=2, min_frames=2, variance_threshold=1e30, spp=3)
    got, ref = _both(dem, 3840, 2160, cam, **kw)
    _assert_same_render(got, ref, "config 3 full size")
    assert got["rays_primary"] == 3840 * 2160 * 3 * 2
    assert np.isfinite(got["depth"]).mean() > 0.2


def test_golden_scene_converges_and_matches_reference_golden():
    dem = H.golden_dem()
    got, ref = _both(dem, H.SIZE, H.SIZE, H.CAM, **H.scene_kwargs(dem))
    _assert_same_render(got, ref, "golden 256x256")
    assert got["converged"] is True and got["variance"] < 1e-3
    gold = H.golden_png()
    rgba = got["rgba"]
    mean_abs = float(np.mean(np.abs(rgba[..., :3].astype(np.float32) - gold[..., :3].astype(np.float32))))
    score = H.ssim(rgba[..., :3], gold[..., :3], 255.0)
    print(f"CUDA vs reference golden: SSIM {score:.6f}, mean abs {mean_abs:.4f}, frames {got['frames']}")
    assert score >= 0.995 and mean_abs <= 2.0          # tests/test_hybrid_terrain_pt.py:853-859
    rmse = float(np.sqrt(np.mean((rgba.astype(np.float64) / 255.0 - ref["rgba"].astype(np.float64) / 255.0) ** 2)))
    assert rmse <= 1e-3                                  # BASELINE.json north_star tolerance (here exactly 0)


def test_public_api_contracts_on_gpu():
    dem = H.golden_dem()
    kw = H.scene_kwargs(dem)
    out = hybrid_render_terrain_reference(dem, 64, 64, H.CAM, **{**kw, "max_frames": 32, "min_frames": 2,
                                                                  "variance_threshold": 1e30})
    for key in ("rgba", "albedo", "normal", "depth", "frames", "variance", "converged", "peak_host_visible_bytes",
                "minmax_pyramid_bytes", "gpu_resource_bytes", "sun_source", "solar_azimuth_deg", "solar_elevation_deg"):
        assert key in out
    assert out["rgba"].shape == (64, 64, 4) and out["rgba"].dtype == np.uint8
    assert out["albedo"].shape == (64, 64, 3) and out["depth"].shape == (64, 64)
    assert out["gpu_resource_bytes"] > out["minmax_pyramid_bytes"] > 0
    assert out["sun_source"] == "manual_angles" and out["solar_azimuth_deg"] == 225.0
    with pytest.raises(Exception, match="did not converge"):
        hybrid_render_terrain_reference(dem, 128, 128, H.CAM, **{**kw, "max_frames": 8, "min_frames": 2,
                                                                  "variance_threshold": 1e-12})
    with pytest.raises(Exception, match="look_at"):
        hybrid_render_terrain_reference(dem, 64, 64, {**H.CAM, "look_at": H.CAM["origin"]}, **kw)
    with pytest.raises(Exception, match="fov"):
        hybrid_render_terrain_reference(dem, 64, 64, {**H.CAM, "fov_y": 0.0}, **kw)
    with pytest.raises(MemoryError, match="memory budget"):
        _native.hybrid_render_terrain_reference(dem, 1920, 1080, H.CAM, **kw, compat_512mib_gate=True)
    # zero sun: renders, no valid-reservoir requirement (render_terrain.rs:465-471)
    z = hybrid_render_terrain_reference(dem, 32, 32, H.CAM, **{**kw, "max_frames": 4, "min_frames": 2,
                                                                "variance_threshold": 1e30, "sun_color": (0, 0, 0)})
    assert z["frames"] == 4
