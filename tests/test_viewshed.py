"""HELIOS viewshed / solar shadow mask, CPU suite: the oracle against the reference's own known-answer
tests (tests/test_viewshed_curvature.py and the unit tests of src/terrain/analysis/viewshed.rs, restated on the same scenes), the
host-side contract of forge3d_b200.viewshed, and the product's CUDA source (k_viewshed, k_shadow_mask + host driver) under the SIMT
interpreter against the oracle, bit for bit.  The -m gpu tests make the same comparison on the device through the public API."""
import math

import numpy as np
import pytest

import _emu
from forge3d_b200 import viewshed as V
from oracle import oracle


def _oracle_viewshed(dem, observer, **kw):
    h, pos, opts = V.viewshed_inputs(dem, observer, **kw)
    return oracle.viewshed(h, pos, opts)


FLAT = dict(height_system="ellipsoidal", earth_model="flat", refraction_model="none")


# ---------------------------------------------------------------------------------------------------------------
# oracle pin: the reference's known-answer tests (tests/test_viewshed_curvature.py)
# ---------------------------------------------------------------------------------------------------------------
def test_ridge_blocks_targets_behind_it_and_physics_arrays_are_positive():         # :61-87
    dem = np.zeros((33, 33), np.float32)
    dem[:, 16] = 600.0
    kw = dict(bounds=(0.0, 0.0, 1.0, 1.0), height_system="ellipsoidal", observer_height=100.0, target_height=0.0, max_distance=120_000.0,
              earth_model="ellipsoid", refraction_model="bennett")
    first, second = _oracle_viewshed(dem, (0.5, 0.1), **kw), _oracle_viewshed(dem, (0.5, 0.1), **kw)
    assert first["visibility"].dtype == np.bool_ and first["visibility"].shape == dem.shape
    assert np.array_equal(first["visibility"], second["visibility"]) and np.array_equal(first["curvature_drop_m"], second["curvature_drop_m"])
    assert first["visibility"][16, 10] and not first["visibility"][16, 24]
    assert first["curvature_drop_m"][16, 32] > 0.0 and first["refraction_gain_m"][16, 32] > 0.0 and first["horizon_distance_m"][16, 32] > 0.0


def test_flat_empty_dem_is_fully_visible():                                          # :91-101
    assert _oracle_viewshed(np.zeros((8, 8), np.float32), (0.5, 0.5), bounds=(0.0, 0.0, 1.0, 1.0), **FLAT)["visibility"].all()


def test_curvature_drop_uses_the_geodesic_distance_at_high_latitude():               # :166-190
    d = _oracle_viewshed(np.zeros((3, 3), np.float32), (75.0, 0.0), bounds=(0.0, 70.0, 10.0, 80.0), height_system="ellipsoidal",
                         earth_model="ellipsoid", refraction_model="none")
    s12, azi1 = V.vincenty_inverse(75.0, 0.0, 78.33333333333333, 8.333333333333334)
    lat, e2 = math.radians(75.0), 6.694_379_990_141_316_5e-3
    w = math.sqrt(1.0 - e2 * math.sin(lat) ** 2)
    meridional, prime_vertical = 6_378_137.0 * (1.0 - e2) / w ** 3, 6_378_137.0 / w
    radius = 1.0 / (math.cos(float(azi1)) ** 2 / meridional + math.sin(float(azi1)) ** 2 / prime_vertical)
    assert d["curvature_drop_m"][0, 2] == pytest.approx(float(s12) ** 2 / (2.0 * radius), rel=2e-5)


def test_blockers_are_marched_on_the_geodesic():                                     # :193-205
    dem = np.zeros((11, 11), np.float32)
    dem[2, 4] = 1_500.0
    vis = _oracle_viewshed(dem, (75.0, 0.0), bounds=(0.0, 70.0, 10.0, 80.0), observer_height=1_000.0, **FLAT)["visibility"]
    assert not vis[0, 10]


def test_continuous_leaf_detects_blocker_between_half_cell_samples():               # :209-224 and viewshed.rs:659-667
    dem = np.array([[16.5, 28.5], [28.5, 24.5]], np.float32)
    vis = _oracle_viewshed(dem, (0.0015, 0.0005), bounds=(0.0, 0.0, 0.002, 0.002), observer_height=8.5, target_height=0.5, **FLAT)["visibility"]
    assert not vis[1, 1]


def test_horizon_includes_terrain_elevation_and_drop_scales_with_sphere_radius():    # :228-270
    dem = np.zeros((3, 3), np.float32)
    dem[1, 2] = 1_000.0
    kw = dict(bounds=(0.0, 0.0, 1.0, 1.0), height_system="ellipsoidal", earth_model="sphere", refraction_model="none")
    d = _oracle_viewshed(dem, (0.5, 0.5), observer_height=2.0, **kw)
    assert d["horizon_distance_m"][1, 2] > d["horizon_distance_m"][0, 2]
    small = _oracle_viewshed(np.zeros((3, 3), np.float32), (0.5, 0.5), sphere_radius_m=3_000_000.0, **kw)
    large = _oracle_viewshed(np.zeros((3, 3), np.float32), (0.5, 0.5), sphere_radius_m=6_000_000.0, **kw)
    assert large["curvature_drop_m"][0, 2] == pytest.approx(2.0 * small["curvature_drop_m"][0, 2], rel=2e-5)


def test_observer_on_the_raster_edge_and_geodesics_outside_the_footprint():         # :273-296
    assert _oracle_viewshed(np.zeros((3, 3), np.float32), (0.5, 0.0), bounds=(0.0, 0.0, 1.0, 1.0), **FLAT)["visibility"].shape == (3, 3)
    with pytest.raises(oracle.OracleError, match="geodesic leaves the DEM footprint"):
        _oracle_viewshed(np.zeros((3, 3), np.float32), (74.1666666667, -26.6666666667), bounds=(-40.0, 70.0, 40.0, 75.0), **FLAT)


def test_curvature_and_refraction_are_load_bearing():
    """test_real_dem_curvature_and_refraction_are_load_bearing (:372-401) on a synthetic 100 km ridge field: the curved and the
    flat viewshed differ, and refraction brings back part of what curvature hides."""
    n = 96
    y, x = np.mgrid[0:n, 0:n].astype(np.float64)
    dem = (120.0 * np.sin(x * 0.21) * np.cos(y * 0.17) + 180.0).astype(np.float32)
    kw = dict(bounds=(6.0, 46.0, 7.3, 46.9), height_system="ellipsoidal", observer_height=30.0)
    flat = _oracle_viewshed(dem, (46.45, 6.65), earth_model="flat", refraction_model="none", **kw)["visibility"]
    vac = _oracle_viewshed(dem, (46.45, 6.65), earth_model="ellipsoid", refraction_model="none", **kw)["visibility"]
    ref = _oracle_viewshed(dem, (46.45, 6.65), earth_model="ellipsoid", refraction_model="bennett", **kw)["visibility"]
    assert flat.sum() > ref.sum() > vac.sum() > 0
    iou = (flat & vac).sum() / (flat | vac).sum()
    assert iou <= 0.96


def test_shadow_mask_of_a_ridge_under_a_low_sun():
    dem = np.zeros((40, 80), np.float32)
    dem[:, 40] = 500.0                                            # north-south wall, cells are ~278 m wide at the equator
    kw = dict(bounds=(0.0, -0.05, 0.2, 0.05), height_system="ellipsoidal", earth_model="flat", refraction_model="none")
    h, inp, opts = V.shadow_mask_inputs(dem, 90.0, 10.0, **kw)  # sun due east, 10 degrees up: shadow length 500 / tan(10 deg) = 2.8 km
    lit = oracle.shadow_mask(h, inp, opts)
    assert lit[:, 41:].all() and lit[:, 40].all()                # east of the wall and its crest are lit
    assert not lit[5:35, 31:40].any() and lit[5:35, :28].all()   # ~10 cells of shadow to the west, sunlight beyond
    h, inp, opts = V.shadow_mask_inputs(dem, 90.0, -1.0, **kw)
    assert not oracle.shadow_mask(h, inp, opts).any()            # sun below the horizon: nothing is lit (shadow_mask_main:614-616)


def test_validation_contract():
    dem = np.zeros((4, 4), np.float32)
    with pytest.raises(ValueError, match="inside local EPSG:4326 bounds"):
        V.viewshed_inputs(dem, (2.0, 0.5), bounds=(0.0, 0.0, 1.0, 1.0), height_system="ellipsoidal")
    with pytest.raises(ValueError, match="spanning less than 180 degrees"):
        V.viewshed_inputs(dem, (0.0, 0.0), bounds=(-170.0, -80.0, 170.0, 80.0), height_system="ellipsoidal")      # :140-150
    h, pos, opts = V.viewshed_inputs(np.zeros((3, 3), np.float32), (0.0, 179.9), bounds=(179.0, -1.0, -179.0, 1.0),
                                     height_system="ellipsoidal")                                                # antimeridian :153-162
    assert oracle.viewshed(h, pos, opts)["visibility"].shape == (3, 3)
    with pytest.raises(ValueError, match="unsupported height_system"):
        V.viewshed_inputs(dem, (0.5, 0.5), bounds=(0.0, 0.0, 1.0, 1.0), height_system="msl")
    with pytest.raises(ValueError, match="unsupported earth_model"):
        V.viewshed_inputs(dem, (0.5, 0.5), bounds=(0.0, 0.0, 1.0, 1.0), height_system="ellipsoidal", earth_model="torus")
    with pytest.raises(oracle.OracleError, match="flat earth only supports refraction_model='none'"):
        _oracle_viewshed(dem, (0.5, 0.5), bounds=(0.0, 0.0, 1.0, 1.0), height_system="ellipsoidal", earth_model="flat", refraction_model="bennett")
    s, az = V.vincenty_inverse(0.0, 0.0, 0.0, 1.0)
    assert float(s) == pytest.approx(111_319.4908, abs=1e-3) and float(az) == pytest.approx(math.pi / 2)


# ---------------------------------------------------------------------------------------------------------------
# the product's CUDA source under the SIMT interpreter vs the oracle
# ---------------------------------------------------------------------------------------------------------------
def _rough_dem(h, w, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    return (300.0 * np.sin(x * 0.33 + seed) * np.cos(y * 0.27) + 80.0 * rng.standard_normal((h, w)).cumsum(0).cumsum(1) * 0.05 + 700.0).astype(np.float32)


MODELS = {"ellipsoid_bennett": dict(earth_model="ellipsoid", refraction_model="bennett"),
          "sphere_effective_radius": dict(earth_model="sphere", sphere_radius_m=5_000_000.0, refraction_model="effective_radius", refraction_k=0.2),
          "flat": dict(earth_model="flat", refraction_model="none")}


def _same(g, o):
    for k in ("visibility", "curvature_drop_m", "refraction_gain_m", "horizon_distance_m"):
        assert np.array_equal(np.ascontiguousarray(g[k]).view(np.uint8), np.ascontiguousarray(o[k]).view(np.uint8)), k


@pytest.mark.parametrize("model", sorted(MODELS))
def test_emulated_cuda_viewshed_and_shadow_mask_are_bit_identical_to_the_oracle(model):
    dem = _rough_dem(37, 53, 3)                                   # ragged, non power-of-two cell counts
    kw = dict(bounds=(7.0, 45.8, 7.6, 46.2), height_system="ellipsoidal", observer_height=12.0, target_height=1.5, **MODELS[model])
    h, pos, opts = V.viewshed_inputs(dem, (46.03, 7.31), **kw)
    o = oracle.viewshed(h, pos, opts)
    skw = {k: v for k, v in kw.items() if k not in ("observer_height", "target_height")}
    rng = np.random.default_rng(9)
    az = rng.uniform(60.0, 300.0, dem.shape)
    el = rng.uniform(-2.0, 35.0, dem.shape)                      # per-cell sun (what SPA would give), some below the horizon
    sh, sinp, sopts = V.shadow_mask_inputs(dem, az, el, **skw)
    so = oracle.shadow_mask(sh, sinp, sopts)
    with _emu.emulated_backend():
        g = V.compute_viewshed(h, pos, opts)
        sg = V.compute_shadow_mask(sh, sinp, sopts)
        limited = V.compute_viewshed(h, pos, {**opts, "max_distance_m": 9_000.0})
    _same(g, o)
    assert np.array_equal(sg, so)
    _same(limited, oracle.viewshed(h, pos, {**opts, "max_distance_m": 9_000.0}))
    assert 0.05 < o["visibility"].mean() < 0.95 and 0.05 < so.mean() < 0.95 and limited["visibility"].sum() < o["visibility"].sum()


def test_emulated_native_validation_and_footprint_error():
    dem = np.zeros((3, 3), np.float32)
    with _emu.emulated_backend():
        with pytest.raises(RuntimeError, match="viewshed failed: viewshed geodesic leaves the DEM footprint"):
            V.viewshed(dem, (74.1666666667, -26.6666666667), bounds=(-40.0, 70.0, 40.0, 75.0), **FLAT)
        with pytest.raises(RuntimeError, match="flat earth only supports refraction_model='none'"):
            V.viewshed(dem, (0.5, 0.5), bounds=(0.0, 0.0, 1.0, 1.0), height_system="ellipsoidal", earth_model="flat")
        bad = dem.copy()
        bad[1, 1] = np.nan
        with pytest.raises(RuntimeError, match="DEM heights and geodesic positions must be finite"):
            V.viewshed(bad, (0.5, 0.5), bounds=(0.0, 0.0, 1.0, 1.0), **FLAT)
        h, pos, opts = V.viewshed_inputs(dem, (0.5, 0.5), bounds=(0.0, 0.0, 1.0, 1.0), **FLAT)
        with pytest.raises(RuntimeError, match="dimensions, observer, heights, spacing, and distance are invalid"):
            V.compute_viewshed(h, pos, {**opts, "max_distance_m": 0.0})
        assert V.viewshed(np.zeros((8, 8), np.float32), (0.5, 0.5), bounds=(0.0, 0.0, 1.0, 1.0), **FLAT).all()


# ---------------------------------------------------------------------------------------------------------------
# GPU parity proper (run on the B200 box)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("model", sorted(MODELS))
def test_gpu_viewshed_and_shadow_mask_are_bit_identical_to_the_oracle(model):
    dem = _rough_dem(200, 260, 5)
    kw = dict(bounds=(7.0, 45.6, 7.9, 46.3), height_system="ellipsoidal", observer_height=25.0, target_height=2.0, **MODELS[model])
    h, pos, opts = V.viewshed_inputs(dem, (45.97, 7.43), **kw)
    g = V.compute_viewshed(h, pos, opts)
    _same(g, oracle.viewshed(h, pos, opts))
    skw = {k: v for k, v in kw.items() if k not in ("observer_height", "target_height")}
    sh, sinp, sopts = V.shadow_mask_inputs(dem, 245.0, 14.0, **skw)
    lit = V.compute_shadow_mask(sh, sinp, sopts)
    assert np.array_equal(lit, oracle.shadow_mask(sh, sinp, sopts))
    assert 0.02 < g["visibility"].mean() < 0.98 and 0.05 < lit.mean() < 0.98 and g["kernel_ms"] > 0.0


# ---------------------------------------------------------------------------------------------------------------
# oracle pin (arithmetic): an independent statement-by-statement restatement of the WGSL, bit for bit
# ---------------------------------------------------------------------------------------------------------------
def _physics(opts):
    """physics_terms (viewshed.rs:54-78) in float64, independently of the C oracle."""
    k = {"none": 0.0, "effective_radius": opts["refraction_k"]}.get(opts["refraction_model"])
    if k is None:
        base = 0.13 if opts["refraction_model"] == "bennett" else 1.0 / 7.0
        k = base * (opts["pressure_mbar"] / 1013.25) * (288.15 / (273.15 + opts["temperature_c"]))
    if opts["earth_model"] == "flat":
        inv_m = inv_p = 0.0
    elif opts["earth_model"] == "sphere":
        inv_m = inv_p = 1.0 / opts["sphere_radius_m"]
    else:
        phi, e2 = math.radians(opts["earth_latitude_deg"]), 6.694_379_990_141_316_5e-3
        w = math.sqrt(1.0 - e2 * math.sin(phi) ** 2)
        inv_m, inv_p = 1.0 / (6_378_137.0 * (1.0 - e2) / w ** 3), 1.0 / (6_378_137.0 / w)
    return [inv_m, inv_p, 1.0 - k, 0.0 if opts["earth_model"] == "flat" else 1.0]


@pytest.mark.parametrize("model", sorted(MODELS))
def test_oracle_matches_the_wgsl_mirror_bit_for_bit(model):
    import _wgsl_mirror_viewshed as M

    dem = _rough_dem(11, 14, 8)
    kw = dict(bounds=(7.0, 45.9, 7.2, 46.05), height_system="ellipsoidal", observer_height=9.0, target_height=1.0, **MODELS[model])
    h, pos, opts = V.viewshed_inputs(dem, (45.97, 7.08), **kw)
    o = oracle.viewshed(h, pos, opts)
    levels, _, _ = oracle.build_minmax(h)
    atan2 = lambda y, x: np.float32(oracle.lib().f3do_atan2(float(y), float(x)))
    S = M.Scene(h, opts, _physics(opts), levels, atan2)
    with np.errstate(all="ignore"):
        for y in range(h.shape[0]):
            for x in range(h.shape[1]):
                vis, drop, gain, horizon = M.viewshed_cell(S, pos, x, y)
                assert vis == int(o["visibility"][y, x]), (x, y)
                for got, want in ((drop, o["curvature_drop_m"]), (gain, o["refraction_gain_m"]), (horizon, o["horizon_distance_m"])):
                    assert np.float32(got).tobytes() == want[y, x].tobytes(), (x, y)
        skw = {k: v for k, v in kw.items() if k not in ("observer_height", "target_height")}
        rng = np.random.default_rng(4)
        sh, sinp, sopts = V.shadow_mask_inputs(dem, rng.uniform(70.0, 290.0, dem.shape), rng.uniform(-1.0, 30.0, dem.shape), **skw)
        so = oracle.shadow_mask(sh, sinp, sopts)
        S2 = M.Scene(sh, sopts, _physics(sopts), levels, atan2)
        for y in range(h.shape[0]):
            for x in range(h.shape[1]):
                assert M.shadow_cell(S2, sinp, x, y) == bool(so[y, x]), (x, y)
    assert 0 < o["visibility"].sum() < o["visibility"].size and 0 < so.sum() < so.size
