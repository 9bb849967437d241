"""The product's CUDA source on the CPU: host driver (csrc/f3d_backend.cu) + every kernel (csrc/*.cuh) compiled by g++
and executed by the cooperative SIMT interpreter in tests/c/emu (CTAs of fibers, real 32-lane warp collectives), compared
bit for bit with the oracle.  This is NOT a product path (the product has no CPU fallback and never loads this library);
it is the pre-flight check that lets kernel restructurings be proven exact - and free of warp-synchronisation deadlocks -
without GPU time.  The -m gpu tests make the same comparisons on the device."""
import numpy as np
import pytest

import _emu
import _helpers as H
from oracle import oracle

VARIANTS = {"default": (), "anyhit_sorted": ("F3D_ANYHIT_SIGN_ORDER=0",), "literal_push_clip": ("F3D_PUSH_CLIP_FOLDED=0",),
            "deferred_leaves": ("F3D_TRACE_DEFER_LEAVES=4", "F3D_DEFER_RULE=2", "F3D_DEFER_STALL=4", "F3D_DEFER_LEAF_BATCH=20"),
            "parked_primary_leaf": ("F3D_PRIMARY_PARK=1",), "no_sun_walk": ("F3D_SUN_NEAR=0",), "tma_staging": ("F3D_TMA_STAGE=1",)}


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def _same_render(g, o):
    for k in ("rgba", "depth", "normal", "albedo", "accum"):
        assert np.array_equal(_bits(g[k]), _bits(o[k])), k
    assert g["frames"] == o["frames"] and g["converged"] == o["converged"]
    assert np.float32(g["variance"]) == np.float32(o["variance"])
    for k in ("rays_primary", "rays_shadow", "rays_ibl", "minmax_pyramid_bytes"):
        assert g[k] == o[k], k


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_emulated_render_is_bit_identical_to_the_oracle(variant):
    dem = H.golden_dem()
    kw = {**H.scene_kwargs(dem), "max_frames": 5, "min_frames": 5, "variance_threshold": 1e30}
    o = oracle.render(dem, 56, 40, H.CAM, want_accum=True, **kw)
    with _emu.emulated_backend(VARIANTS[variant]) as native:
        g = native.hybrid_render_terrain_reference(dem, 56, 40, H.CAM, want_accum=True, **kw)
    _same_render(g, o)
    assert o["rays_shadow"] > 1000 and np.isfinite(o["depth"]).any() and np.isnan(o["depth"]).any()


@pytest.mark.parametrize("az,el,earth,escape", [(20.0, 35.0, "flat", "1"), (110.0, 12.0, "wgs84", "0"), (200.0, 60.0, "flat", "1"), (290.0, 3.0, "wgs84", "1"),
                                                (135.0, 0.5, "flat", "0"), (270.0, 89.97, "flat", "1")])
def test_emulated_near_field_walks_are_bit_identical_to_the_oracle(az, el, earth, escape, monkeypatch):
    """The CPU twin of tests/test_gpu_parity.py::test_sun_horizon_strips_are_exact_in_every_octant / test_escape_map_is_exact_and_culls:
    sun horizon strips + the near-field walk of k_ascent (two passes per list, far lists) and, with F3D_B200_ESCAPE=1, the escape map
    for the IBL rays, over a rough, ragged, non-square DEM for several octants of the sun direction - under the SIMT interpreter, so
    that the warp-convergent walk loops and the leaf rings are exercised with real 32-lane collectives before any GPU time."""
    rng = np.random.default_rng(int(az * 7 + el * 13))
    shape = (83, 121)
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    dem = (0.6 * np.sin(xx / 9.0) * np.cos(yy / 7.0) + 0.4 * np.sin((xx + 2 * yy) / 17.0) + 0.35 * rng.uniform(0.0, 1.0, shape)).astype(np.float32)
    span = 120.0
    kw = dict(spacing=(span / (shape[1] - 1), 0.8 * span / (shape[0] - 1)), exaggeration=14.0, albedo=H.ALBEDO, sun_azimuth_deg=az,
              sun_elevation_deg=el, earth_model=earth, refraction_model="none" if earth == "flat" else "bennett", max_frames=3, min_frames=3,
              variance_threshold=1e30)
    monkeypatch.setenv("F3D_B200_ESCAPE", escape)
    o = oracle.render(dem, 72, 48, H.CAM, want_accum=True, **kw)
    with _emu.emulated_backend() as native:
        g = native.hybrid_render_terrain_reference(dem, 72, 48, H.CAM, want_accum=True, **kw)
    _same_render(g, o)


def test_emulated_early_aov_readback_matches_the_ordinary_path(monkeypatch):
    """With page-locked output arrays the one-call path decodes and copies the three AOVs on a side stream right after set-up
    (csrc/f3d_backend.cu early_aov_readback, k_resolve_aovs) while the frames render: same bytes as k_resolve's AOV half."""
    dem = H.golden_dem()
    kw = {**H.scene_kwargs(dem), "max_frames": 4, "min_frames": 4, "variance_threshold": 1e30}
    o = oracle.render(dem, 56, 40, H.CAM, want_accum=True, **kw)
    monkeypatch.setenv("F3D_EMU_HOST_PINNED", "1")
    with _emu.emulated_backend() as native:
        g = native.hybrid_render_terrain_reference(dem, 56, 40, H.CAM, want_accum=True, **kw)
        monkeypatch.setenv("F3D_B200_EARLY_AOVS", "0")
        g_off = native.hybrid_render_terrain_reference(dem, 56, 40, H.CAM, want_accum=True, **kw)
    _same_render(g, o)
    _same_render(g_off, o)
    assert g["kernel_launches"] == g_off["kernel_launches"] + 1          # k_resolve_aovs ran


def test_emulated_render_with_mesh_env_map_spp_and_curvature():
    dem = H.sine_dem(64)
    rng = np.random.default_rng(2)
    env = rng.random((8, 16, 3), dtype=np.float32) + 0.2
    verts = np.array([[-8, 14, -6], [9, 15, -5], [8, 11, 7], [-9, 12, 6], [0, 22, 0]], np.float32)
    tris = np.array([[0, 1, 2], [0, 2, 3], [0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4], [0, 2, 1], [1, 3, 2], [2, 0, 3]], np.uint32)
    kw = dict(spacing=(1.5, 1.5), exaggeration=1.0, albedo=(0.5, 0.55, 0.45), sun_azimuth_deg=290.0, sun_elevation_deg=18.0,
              env_map=env, mesh_vertices=verts, mesh_indices=tris, spp=3, seed=99, max_frames=3, min_frames=3,
              variance_threshold=1e30, earth_model="sphere", sphere_radius_m=5000.0, refraction_model="none")
    cam = {"origin": (0.0, 30.0, 70.0), "look_at": (0.0, 6.0, 0.0), "up": (0.0, 1.0, 0.0), "fov_y": 45.0}
    o = oracle.render(dem, 40, 32, cam, want_accum=True, **kw)
    with _emu.emulated_backend() as native:
        g = native.hybrid_render_terrain_reference(dem, 40, 32, cam, want_accum=True, **kw)
    _same_render(g, o)


def test_emulated_seams_pyramid_and_cooperative_ray_batches():
    rng = np.random.default_rng(4)
    dem = (rng.standard_normal((37, 70)) * 30).astype(np.float32)
    h = H.curvature_fixture()
    arb, _ = H.kat_rays(h)
    rays = arb[:4096]
    with _emu.emulated_backend() as native:
        g_levels, gcw, gch = native.build_minmax(dem)
        got = {}
        for any_hit, curv in [(True, True), (False, False)]:
            kw = dict(any_hit=any_hit, apply_curvature=curv, inv_two_r_prime=float(H.PROOF_INV_TWO_R), curvature_enabled=True)
            got[any_hit] = (native.trace_rays(h, (500.0, 500.0), (0.0, 0.0), 1.0, rays, variant=0, **kw),
                            native.trace_rays(h, (500.0, 500.0), (0.0, 0.0), 1.0, rays, variant=1, **kw),
                            oracle.trace_rays(h, (500.0, 500.0), (0.0, 0.0), 1.0, rays, **kw))
    o_levels, ocw, och = oracle.build_minmax(dem)
    assert (gcw, gch) == (ocw, och) and len(g_levels) == len(o_levels)
    for a, b in zip(g_levels, o_levels):
        assert np.array_equal(_bits(a), _bits(b))
    for any_hit, (fast, literal, want) in got.items():
        for have in (fast, literal):          # 32-lane cooperative production traversal and the literal loop
            assert np.array_equal(have[0], want[0])
            if any_hit and have is fast:      # sign-ordered any-hit: the flag is exact, the first hit found may differ in ties
                assert (have[1].view(np.uint32) == want[1].view(np.uint32))[want[0]].mean() > 0.999
            else:
                assert np.array_equal(_bits(have[1]), _bits(want[1]))
            if not any_hit:
                assert np.array_equal(_bits(have[2]), _bits(want[2]))


@pytest.mark.parametrize("world,width,height,frames", [(2, 48, 72, 6), (3, 40, 100, 4)])
def test_emulated_row_partition_is_bit_identical_to_one_session(world, width, height, frames):
    """SURVEY section 8e on the CPU: `world` sessions (one per rank) in ONE process, stepped frame by frame in rank order.
    Their peer pointers are real pointers here, so k_primary's halo stores (4 rows up / 3 rows down) and the in-kernel
    frame barrier (wait_neighbours / signal_neighbours) run exactly as over NVLink; every rank resolves its owned rows into
    the same buffers.  The result must equal the unpartitioned session bit for bit."""
    from forge3d_b200.session import Session

    dem = H.golden_dem()
    kw = {**H.scene_kwargs(dem), "max_frames": frames, "min_frames": frames, "variance_threshold": 1e30}
    outs = {}
    with _emu.emulated_backend():
        for tag, nranks in (("single", 1), ("split", world)):
            sessions = [Session(dem, width, height, H.CAM, part_rank=r, part_world=nranks, part_block_rows=16, **kw)
                        for r in range(nranks)]
            if nranks > 1:
                table = b"".join(s.ipc_export() for s in sessions)
                for s in sessions:
                    s.ipc_import(table)
            for _ in range(frames):
                for s in sessions:
                    s.render_frames(1)
            rgba = np.zeros((height, width, 4), np.uint8)
            depth = np.zeros((height, width), np.float32)
            normal = np.zeros((height, width, 3), np.float32)
            albedo = np.zeros((height, width, 3), np.float32)
            variances = []
            for s in sessions:
                variances.append(s.variance()[0])
                s.resolve_device(rgba.ctypes.data, albedo.ctypes.data, normal.ctypes.data, depth.ctypes.data, check_validity=True)
            outs[tag] = (rgba, depth, normal, albedo, max(variances))
            for s in sessions:
                s.close()
    for a, b in zip(outs["single"][:4], outs["split"][:4]):
        assert np.array_equal(_bits(a), _bits(b))
    assert np.float32(outs["single"][4]) == np.float32(outs["split"][4])
    assert outs["single"][0][..., :3].max() > 0
