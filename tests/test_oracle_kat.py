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
