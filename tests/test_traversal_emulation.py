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
VARIANTS = {"default": (), "anyhit_sorted": ("F3D_ANYHIT_SIGN_ORDER=0",), "literal_push_clip": ("F3D_PUSH_CLIP_FOLDED=0",)}
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
