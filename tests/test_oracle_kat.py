"""Oracle terrain_trace vs the reference's conservative-descent KAT
(terrain_heightfield.rs:2001-2127 / :2129-2285): 10 000 xorshift rays + 255x255 shadow-mask
rays over `curvature_fixture`, compared with an independent f64 brute-force cell DDA.
Reference thresholds: false misses == 0, false-hit rate < 0.1 %, mask agreement >= 99.9 %."""
import numpy as np

from oracle import oracle
import _helpers as H


def _trace(h, rays):
    hit, t, _ = oracle.trace_rays(h, (H.PROOF_SPACING, H.PROOF_SPACING), (0.0, 0.0), 1.0, rays, any_hit=True,
                                  apply_curvature=True, inv_two_r_prime=float(H.PROOF_INV_TWO_R),
                                  curvature_enabled=True)
    return hit


def test_conservative_descent_kat():
    h = H.curvature_fixture()
    arb, mask = H.kat_rays(h)
    assert arb.shape == (10_000, 8) and mask.shape == (255 * 255, 8)
    b_arb, b_mask = H.brute_2d_hit(h, arb), H.brute_2d_hit(h, mask)
    d_arb, d_mask = _trace(h, arb), _trace(h, mask)
    false_miss = int((b_arb & ~d_arb).sum())
    false_hit = int((~b_arb & d_arb).sum())
    agree = float((b_mask == d_mask).mean())
    print(f"KAT: rays=10000 hits={int(b_arb.sum())} false_misses={false_miss} false_hits={false_hit} "
          f"mask_hits={int(b_mask.sum())} agreement={agree:.6f}")
    assert 0.02 < b_arb.mean() < 0.95 and 0.05 < b_mask.mean() < 0.99  # the fixture exercises both outcomes
    assert false_miss == 0
    assert false_hit / 10_000.0 < 0.001
    assert agree >= 0.999
    assert int((b_mask & ~d_mask).sum()) == 0


def test_captured_regression_ray():
    # terrain_heightfield.rs:1970-1989: the ray NVIDIA Vulkan rounded onto a shared leaf boundary
    h = H.curvature_fixture()
    ray = np.array([[125_750.0, 870.54614, 67_750.0, 1e-3, 0.79859173, 0.010471784, 0.60178196, 200_000.0]], np.float32)
    assert H.brute_2d_hit(h, ray)[0]
    hit, _, _ = oracle.trace_rays(h, (500.0, 500.0), (0.0, 0.0), 1.0, ray, any_hit=True, apply_curvature=True,
                                  inv_two_r_prime=6.8259382e-8, curvature_enabled=True)
    assert hit[0]


def test_curvature_policy_changes_long_shadow_rays():
    # mutation-test analogue (:2017-2071): dropping the curvature term must change some outcomes
    h = H.curvature_fixture()
    _, mask = H.kat_rays(h)
    with_c = _trace(h, mask)
    no_c, _, _ = oracle.trace_rays(h, (500.0, 500.0), (0.0, 0.0), 1.0, mask, any_hit=True, apply_curvature=False,
                                   inv_two_r_prime=float(H.PROOF_INV_TWO_R), curvature_enabled=True)
    assert (with_c != no_c).sum() > 0
    assert (no_c & ~with_c).sum() >= 0 and (~no_c & with_c).sum() == 0  # curvature only ever lifts the ray


def test_closest_hit_matches_any_hit_occlusion():
    h = H.curvature_fixture()
    arb, _ = H.kat_rays(h)
    hit_any, _, _ = oracle.trace_rays(h, (500.0, 500.0), (0.0, 0.0), 1.0, arb, any_hit=True, apply_curvature=False)
    hit_cl, t_cl, n = oracle.trace_rays(h, (500.0, 500.0), (0.0, 0.0), 1.0, arb, any_hit=False, apply_curvature=False)
    # any-hit additionally accepts "entered below the surface" leaves (hybrid_terrain_traversal.wgsl:202-207)
    assert (hit_cl & ~hit_any).sum() == 0
    assert (hit_any != hit_cl).mean() < 0.01
    assert np.allclose(np.linalg.norm(n[hit_cl], axis=1), 1.0, atol=1e-5)
    assert (t_cl[hit_cl] > 1e-3).all() and (t_cl[~hit_cl] == 200_000.0).all()


def test_closest_hit_distance_matches_f64_reference():
    """Depth accuracy (the depth AOV is f32 `t`): the oracle's closest-hit distance against an independent f64
    cell DDA with the exact patch quadratic, on rays started ABOVE the surface so that both report the first
    crossing.  f32 evaluation at kilometre scale is good to ~1e-5 relative."""
    h = H.curvature_fixture()
    rng = np.random.default_rng(3)
    n = 20000
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0] = rng.uniform(2000, 125000, n)
    rays[:, 2] = rng.uniform(2000, 125000, n)
    rays[:, 1] = rng.uniform(2500, 6000, n)                 # fixture heights stay below ~1800 m
    d = rng.standard_normal((n, 3)); d[:, 1] = -np.abs(d[:, 1]) * 0.4 - 0.02
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 4:7] = d
    rays[:, 3] = 1e-3
    rays[:, 7] = 1e30
    hit, t, nrm = oracle.trace_rays(h, (500.0, 500.0), (0.0, 0.0), 1.0, rays, any_hit=False, apply_curvature=False)
    ref = H.brute_first_hit_t(h, rays)
    ref_hit = np.isfinite(ref)
    assert (hit == ref_hit).mean() > 0.9995                  # silhouette-grazing rays may flip
    both = hit & ref_hit
    assert both.mean() > 0.5
    rel = np.abs(t[both].astype(np.float64) - ref[both]) / ref[both]
    print(f"closest-hit t vs f64: n={int(both.sum())} max rel err {rel.max():.2e}, median {np.median(rel):.2e}")
    assert np.percentile(rel, 99.9) < 2e-4 and np.median(rel) < 5e-6
    # the hit point lies on the bilinear surface and the normal is the unit patch normal
    assert np.allclose(np.linalg.norm(nrm[both], axis=1), 1.0, atol=1e-5)
