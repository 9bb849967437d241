"""The PRODUCTION traversal source (forge3d_b200/csrc/f3d_trace_fast.cuh), compiled for the host and run as a one-lane
warp, against the oracle - on the CPU.  This is the pre-flight check for any restructuring of the CUDA traversal: the
same header the GPU runs must stay bit-identical to the WGSL restatement before GPU minutes are spent on it
(the -m gpu tests then confirm it on the device, where the warp is 32 lanes wide)."""
import numpy as np
import pytest

import _emu
import _helpers as H
from oracle import oracle

# compile-time variants of the traversal that must all be exact (the default build is the first one)
VARIANTS = {"default": (), "anyhit_sorted": ("F3D_ANYHIT_SIGN_ORDER=0",), "literal_push_clip": ("F3D_PUSH_CLIP_FOLDED=0",),
            "parked_leaf": ("F3D_PRIMARY_PARK=1",)}
# variants whose any-hit rays may report a different (equally valid) first hit: flags are compared, not t
FLAG_ONLY_ANYHIT = {"default", "literal_push_clip"}


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _random_rays(n, w, h, spacing, dem, seed):
    rng = np.random.default_rng(seed)
    rays = np.zeros((n, 8), np.float32)
    ext = np.array([(w - 1) * spacing, (h - 1) * spacing])
    rays[:, 0] = rng.uniform(-0.2 * ext[0], 1.2 * ext[0], n)
    rays[:, 2] = rng.uniform(-0.2 * ext[1], 1.2 * ext[1], n)
    rays[:, 1] = rng.uniform(dem.min() - 5, dem.max() + 60, n)
    d = rng.standard_normal((n, 3))
    d[:, 1] *= 0.35
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 4:7] = d
    rays[::17, 4] = 0.0          # axis-parallel rays exercise terrain_safe_inv
    rays[::23, 6] = 0.0
    rays[:, 3] = 1e-3
    rays[:, 7] = np.where(rng.uniform(size=n) < 0.3, rng.uniform(50, 3000, n), 1e30)
    return rays


@pytest.mark.parametrize("variant", sorted(VARIANTS))
@pytest.mark.parametrize("any_hit,curv", [(True, True), (True, False), (False, False), (False, True)])
def test_emulated_production_traversal_on_kat_rays(any_hit, curv, variant):
    h = H.curvature_fixture()
    arb, mask = H.kat_rays(h)
    rays = np.concatenate([arb, mask[::5]])
    kw = dict(any_hit=any_hit, apply_curvature=curv, inv_two_r_prime=float(H.PROOF_INV_TWO_R), curvature_enabled=True)
    eh, et, en, nodes = _emu.trace_rays(h, (500.0, 500.0), (0.0, 0.0), 1.0, rays, defines=VARIANTS[variant], **kw)
    oh, ot, on = oracle.trace_rays(h, (500.0, 500.0), (0.0, 0.0), 1.0, rays, **kw)
    assert np.array_equal(eh, oh)
    if any_hit and variant in FLAG_ONLY_ANYHIT:
        assert (_bits(et)[eh] == _bits(ot)[oh]).mean() > 0.99    # same front-to-back order except in exact key ties
    else:
        assert np.array_equal(_bits(et), _bits(ot))
    if not any_hit:   # normals are a closest-hit output (finish_hit); any-hit callers only read the flag
        assert np.array_equal(_bits(en)[eh], _bits(on)[oh])
    assert nodes > 0


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_emulated_production_traversal_on_random_rays_over_a_ragged_dem(variant):
    rng = np.random.default_rng(11)
    w, h = 300, 173
    dem = (rng.standard_normal((h, w)).cumsum(0).cumsum(1) * 0.05 + rng.standard_normal((h, w)) * 2.0).astype(np.float32)
    rays = _random_rays(120_000, w, h, 7.5, dem, seed=12)
    for any_hit, curv in [(False, False), (True, False), (True, True)]:
        kw = dict(any_hit=any_hit, apply_curvature=curv, inv_two_r_prime=3e-6, curvature_enabled=True)
        eh, et, en, _ = _emu.trace_rays(dem, (7.5, 7.5), (0.0, 0.0), 1.3, rays, defines=VARIANTS[variant], **kw)
        oh, ot, on = oracle.trace_rays(dem, (7.5, 7.5), (0.0, 0.0), 1.3, rays, **kw)
        assert np.array_equal(eh, oh)
        if any_hit and variant in FLAG_ONLY_ANYHIT:
            assert (_bits(et)[eh] == _bits(ot)[oh]).mean() > 0.99
        else:
            assert np.array_equal(_bits(et), _bits(ot))
        if not any_hit:
            assert np.array_equal(_bits(en)[eh], _bits(on)[oh])
        assert 0.05 < eh.mean() < 0.95


def _surface_rays(dem, spacing, n, seed):
    """Secondary rays as the renderer makes them: origins 1e-3 above random points of the surface (many of them in the border
    cells, where a ray can start outside every cell's slab), directions = one low sun or a hemisphere sample."""
    rng = np.random.default_rng(seed)
    h, w = dem.shape
    gx = rng.uniform(0, w - 1, n)
    gz = rng.uniform(0, h - 1, n)
    edge = rng.uniform(size=n) < 0.25
    gx[edge] = np.where(rng.uniform(size=edge.sum()) < 0.5, rng.uniform(0, 1.5, edge.sum()), rng.uniform(w - 2.5, w - 1, edge.sum()))
    gz[edge & (rng.uniform(size=n) < 0.5)] = rng.uniform(h - 2.2, h - 1, 1)[0]
    on_plane = rng.uniform(size=n) < 0.15          # exactly on a cell border: the ray starts inside no cell's open slab
    gx[on_plane] = np.round(gx[on_plane])
    gz[on_plane & (rng.uniform(size=n) < 0.5)] = h - 1.0
    ix, iz = np.minimum(gx.astype(int), w - 2), np.minimum(gz.astype(int), h - 2)
    u, v = gx - ix, gz - iz
    y = (dem[iz, ix] * (1 - u) + dem[iz, ix + 1] * u) * (1 - v) + (dem[iz + 1, ix] * (1 - u) + dem[iz + 1, ix + 1] * u) * v
    rays = np.zeros((n, 8), np.float32)
    ox = -0.5 * (w - 1) * spacing
    oz = -0.5 * (h - 1) * spacing
    rays[:, 0] = ox + gx * spacing
    rays[:, 2] = oz + gz * spacing
    rays[:, 1] = y + rng.choice([1e-3, 2e-3, -1e-3, 0.05], n)
    rays[:, 3] = 1e-3
    rays[:, 7] = np.where(rng.uniform(size=n) < 0.2, rng.uniform(20, 3000, n), 1e30)
    return rays, (ox, oz)


@pytest.mark.parametrize("curv", [True, False])
def test_emulated_bottom_up_start_matches_the_oracle(curv):
    """k_ascent + k_trace's seed traversal (csrc/f3d_trace_fast.cuh: ascent_seeds) as a one-lane warp: the occlusion flag of
    every ray must equal the oracle's top-down terrain_trace, for ascending sun rays (monotone tests), descending ones and
    hemisphere rays, on a ragged DEM whose border cells are over-sampled."""
    rng = np.random.default_rng(21)
    w, h = 333, 190
    dem = (rng.standard_normal((h, w)).cumsum(0).cumsum(1) * 0.4 + rng.standard_normal((h, w)) * 3.0).astype(np.float32)
    rays, origin = _surface_rays(dem, 10.0, 90_000, 22)
    n = rays.shape[0]
    d = rng.standard_normal((n, 3))
    d[:, 1] = np.abs(d[:, 1]) * rng.choice([1.0, 1.0, -0.3], n)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    az, el = np.radians(302.0), np.radians(9.0)
    d[: n // 3] = [np.cos(az) * np.cos(el), np.sin(el), np.sin(az) * np.cos(el)]
    d[::29, 0] = 0.0
    rays[:, 4:7] = d
    kw = dict(apply_curvature=curv, inv_two_r_prime=2e-6, curvature_enabled=True)
    bh, nodes = _emu.trace_rays_bottom_up(dem, (10.0, 10.0), origin, 1.0, rays, **kw)
    oh, _, _ = oracle.trace_rays(dem, (10.0, 10.0), origin, 1.0, rays, any_hit=True, **kw)
    assert np.array_equal(bh, oh), int((bh != oh).sum())
    assert 0.1 < oh.mean() < 0.9 and nodes > n
