"""Wavefront multi-bounce path tracer: oracle pins, the emulated CUDA build against the oracle bit for bit,
host-side scene packing, the reference's error rules, and the GPU parity cases."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

import _emu
from _png import read_png
from _ssim_gate import ssim
from forge3d_b200 import wavefront as wf
from oracle import oracle

GOLDEN = Path(__file__).parent / "golden"
REF_GOLDEN = Path("/root/reference/tests/golden/adjudication/pt_reference.png")
f32 = np.float32


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def _adjudication():
    return wf.scene_from_desc(wf.adjudication_scene())


def _sheet(res, seed, y0=0.0):
    rng = np.random.default_rng(seed)
    xs = np.linspace(-6.0, 6.0, res + 1, dtype=f32)
    x, z = np.meshgrid(xs, xs)
    y = (y0 + 0.25 * np.sin(x * 0.9) * np.cos(z * 0.7) + rng.uniform(-0.03, 0.03, x.shape)).astype(f32)
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(f32)
    i = (np.arange(res)[:, None] * (res + 1) + np.arange(res)[None, :]).reshape(-1)
    tris = np.concatenate([np.stack([i, i + res + 1, i + res + 2], 1), np.stack([i, i + res + 2, i + 1], 1)]).astype(np.uint32)
    return verts, tris[rng.permutation(len(tris))]


def _rich_scene(instanced: bool, res: int = 6) -> wf.WavefrontScene:
    """Every branch the kernels have: Lambert, GGX metal, glass, an emitter, two suns, a live disc light, a bumpy mesh (BVH when
    res > 2), optionally instanced through a rigid transform."""
    d = wf.adjudication_scene()
    d.spheres = [
        wf.SphereDesc([-1.15, 1.0, 0.0], 1.0, [0.63, 0.28, 0.22], 0.70),
        wf.SphereDesc([1.30, 0.8, 0.55], 0.8, [0.95, 0.93, 0.88], 0.25, metallic=1.0),
        wf.SphereDesc([0.25, 0.5, 1.6], 0.5, [0.98, 0.98, 0.98], 0.10, ior=1.5),
        wf.SphereDesc([-0.4, 0.35, 2.2], 0.35, [0.8, 0.8, 0.8], 0.5, emissive=[2.0, 1.2, 0.4]),
        wf.SphereDesc([0.0, -1000.0, 0.0], 0.0, [0.42, 0.45, 0.40], 0.90),
    ]
    s = wf.scene_from_desc(d)
    s.dir_lights = np.stack([wf.pack_directional_light([-0.45, -0.80, -0.30], 2.0, [1.0, 0.97, 0.92], 1.0),
                             wf.pack_directional_light([0.6, -0.5, 0.2], 0.7, [0.6, 0.7, 1.0], 0.5)])
    s.area_lights = np.stack([wf.pack_area_light([0.5, 3.5, 1.0], [0.0, -1.0, 0.1], 0.8, 6.0, [1.0, 0.8, 0.6], 1.0),
                              wf.pack_area_light([-2.0, 2.5, 2.0], [0.5, -1.0, -0.4], 0.4, 9.0, [0.5, 0.7, 1.0], 2.0)])
    s.importance = np.array([1.0, 0.9, 1.0, 1.0], f32)        # shorter than the sphere list: slot 4 falls back to 1
    e = s.environment.reshape(4, 4).copy()
    e[0, :3], e[1, :3] = [0.25, 0.22, 0.20], [0.40, 0.48, 0.62]  # a real gradient this time
    e[2, :3], e[3, :3] = [0.30, 0.28, 0.26], [0.35, 0.45, 0.70]
    s.environment = e.reshape(16)
    s.mesh_xyz, s.mesh_idx = _sheet(res, 11)
    if instanced:
        a = 0.3
        rot = np.array([[np.cos(a), 0, np.sin(a), 0], [0, 1, 0, 0], [-np.sin(a), 0, np.cos(a), 0], [0.4, 0.05, -0.3, 1]], f32)  # columns
        o2w = rot.reshape(16)                                    # column-major: rows of this array are the matrix columns
        m = rot.T.astype(np.float64)                             # the matrix itself
        w2o = np.linalg.inv(m).T.astype(f32).reshape(16)
        s.instances = np.stack([wf.pack_instance(o2w, w2o, 0, 4), wf.pack_instance(blas_index=1, material_id=0)])  # 2nd: unknown BLAS
    else:
        s.instances = np.zeros((0, 36), f32)
    return s.normalized()


# ------------------------------------------------------------------ oracle pins
def test_pinned_log2_and_pow():
    xs = np.concatenate([np.linspace(1e-6, 4.0, 4001), 2.0 ** np.arange(-120, 120, 7.3)]).astype(f32)
    got = np.array([oracle.log2(float(x)) for x in xs])
    assert np.max(np.abs(got - np.log2(xs.astype(np.float64))) / np.maximum(1.0, np.abs(np.log2(xs.astype(np.float64))))) < 2e-7
    assert oracle.log2(1.0) == 0.0 and oracle.log2(0.25) == -2.0 and oracle.log2(0.0) == -np.inf and oracle.log2(-1.0) == -np.inf
    L = oracle.lib()
    import ctypes as C
    L.f3do_pow.restype = C.c_float
    L.f3do_pow.argtypes = [C.c_float, C.c_float]
    for x, y in [(0.5, 1 / 17), (0.9, 1 / 2.4), (0.0031308, 1 / 2.4), (1.0, 3.3)]:
        assert abs(L.f3do_pow(x, y) - x ** y) < 3e-7 * max(1.0, x ** y)
    assert L.f3do_pow(0.0, 1 / 17) == 0.0                        # pow(1 - u, 1/17) at u = 1


def test_frame_seeds_and_sobol_known_answers():
    import ctypes as C
    L = oracle.lib()
    L.f3do_wavefront_splitmix32.restype = C.c_uint32
    L.f3do_wavefront_splitmix32.argtypes = [C.c_uint32]

    def splitmix(x):                                             # adjudication.rs:226-232, in Python integers
        x = (x + 0x9E3779B9) & 0xFFFFFFFF
        z = x
        z = ((z ^ (z >> 16)) * 0x21F0AAAD) & 0xFFFFFFFF
        z = ((z ^ (z >> 15)) * 0x735A2D97) & 0xFFFFFFFF
        return z ^ (z >> 15)

    for x in (0, 1, 0x9E3779B9, 0x85EBCA6B, 0xFFFFFFFF, 123456789):
        assert L.f3do_wavefront_splitmix32(x) == splitmix(x)
    x, y = C.c_float(), C.c_float()
    L.f3do_wavefront_sobol2.argtypes = [C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    want_x = [0.0, 0.5, 0.25, 0.75, 0.125]                       # dimension 0 is the van der Corput sequence
    for i, wx in enumerate(want_x):
        L.f3do_wavefront_sobol2(i, C.byref(x), C.byref(y))
        assert x.value == wx
        yb = 0
        for j in range(32):                                      # pt_raygen.wgsl:107-120: base ^ (base >> 1) ^ (base >> 3)
            if (i >> j) & 1:
                b = 0x80000000 >> j
                yb ^= b ^ (b >> 1) ^ (b >> 3)
        assert y.value == float(f32(yb) * f32(2.0 ** -32))


def test_oracle_render_is_pinned_by_checksum():
    """The oracle's arithmetic must not drift: a small render of the adjudication scene is pinned by the SHA-256 of its radiance."""
    r = oracle.wavefront_render(_adjudication(), 64, 48, 6)
    pin = json.loads((GOLDEN / "wavefront_pin.json").read_text())
    assert hashlib.sha256(_bits(r["hdr"]).tobytes()).hexdigest() == pin["oracle_64x48x6_hdr_sha256"]
    assert r["min_iterations"] >= 2 and r["max_rays_per_frame"] <= 4 * 64 * 48


def test_golden_pin_record_passes_the_reference_drift_gate():
    """tools/wavefront_golden_pin.py rendered 512 x 512 x 4096 spp with the oracle and scored it against the reference's own
    tests/golden/adjudication/pt_reference.png with the reference's drift gate (tests/test_adjudication_gate.py:48-49,136-153)."""
    pin = json.loads((GOLDEN / "wavefront_pin.json").read_text())
    assert (pin["width"], pin["height"], pin["spp"]) == (512, 512, 4096)
    assert pin["ssim_vs_reference_golden"] >= 0.995 and pin["mean_abs_vs_reference_golden"] <= 2.0
    # ... and through the reference's adjudication gate itself (dE2000 < 2 on >= 95 % of lit pixels, shadow-band SSIM > 0.96,
    # tests/test_adjudication_gate.py:156-200) against the reference's committed RASTER golden, scored with the reference's own helpers
    adj = pin["adjudication_vs_reference_raster_golden"]
    assert adj["oracle_pt"]["delta_e2000_below_2_fraction_of_lit"] >= 0.95 and adj["oracle_pt"]["shadow_band_ssim"] > 0.96
    assert adj["oracle_pt"]["mean_delta_e2000_lit"] <= adj["reference_golden_pt"]["mean_delta_e2000_lit"]   # at least as close as its own PT
    img = read_png(GOLDEN / "wavefront_oracle_512.png")
    assert img.shape == (512, 512, 4) and (img[..., 3] == 255).all()
    if REF_GOLDEN.exists():                                      # in the build container the score is recomputed from the two images
        ref = read_png(REF_GOLDEN)
        assert abs(ssim(img[..., :3], ref[..., :3]) - pin["ssim_vs_reference_golden"]) < 1e-9
        assert float(np.abs(img[..., :3].astype(f32) - ref[..., :3].astype(f32)).mean()) <= 2.0


def test_oracle_low_spp_converges_to_the_committed_render():
    img = read_png(GOLDEN / "wavefront_oracle_512.png")[..., :3].astype(f32)
    errs = []
    for spp in (4, 64):
        r = oracle.wavefront_render(_adjudication(), 512, 512, spp)
        errs.append(float(np.abs(r["rgba8"][..., :3].astype(f32) - img).mean()))
    assert errs[1] < 1.0 and errs[1] < 0.45 * errs[0]           # Monte-Carlo noise falls ~ 1/sqrt(spp): x16 samples, < half the error


def test_oracle_slices_of_frames_accumulate_identically():
    s = _adjudication()
    whole = oracle.wavefront_render(s, 40, 30, 5)
    part = oracle.wavefront_render(s, 40, 30, 5, first_frame=0, num_frames=2, resolve=False)
    part = oracle.wavefront_render(s, 40, 30, 5, first_frame=2, num_frames=3, accum=part["accum"])
    assert np.array_equal(_bits(whole["hdr"]), _bits(part["hdr"]))


def test_oracle_branches_are_all_exercised():
    """The rich scene must really light up what it claims to: emitter, metal, glass, disc lights, mesh."""
    s = _rich_scene(False)
    base = oracle.wavefront_render(s, 48, 36, 4)["hdr"]
    for mutate in ("emissive", "metal", "glass", "area", "mesh", "sun2"):
        t = _rich_scene(False)
        if mutate == "emissive":
            t.spheres[3, 12:15] = 0
        elif mutate == "metal":
            t.spheres[1, 7] = 0
        elif mutate == "glass":
            t.spheres[2, 9] = 1.0
        elif mutate == "area":
            t.area_lights[:, 7] = 0
        elif mutate == "mesh":
            t.mesh_xyz[:, 1] -= 0.2
        else:
            t.dir_lights[1, 3] = 0
        other = oracle.wavefront_render(t, 48, 36, 4)["hdr"]
        assert not np.array_equal(base, other), mutate
    assert np.isfinite(base).all()


# ------------------------------------------------------------------ host side
def test_scene_packing_matches_the_reference_layouts():
    d = wf.adjudication_scene()
    s = wf.scene_from_desc(d)
    assert s.spheres.shape == (4, 20) and s.spheres.itemsize * 20 == 80          # WavefrontGpuSphere stride, reference_scene.rs:249
    assert s.spheres[1, 8] == f32(0.55) and (s.spheres[:, 9] == 1.0).all() and (s.spheres[:, 12:17] == 0).all()
    assert s.spheres[3, 3] == 0.0                                                # slot 3: plane material, radius 0
    assert s.dir_lights.shape == (1, 8) and abs(float(np.linalg.norm(s.dir_lights[0, :3])) - 1.0) < 1e-6
    assert s.area_lights.shape == (1, 12) and s.area_lights[0, 11] == 0.0
    assert np.array_equal(s.environment.reshape(4, 4)[0], s.environment.reshape(4, 4)[1])       # flat ambient
    assert np.array_equal(s.environment.reshape(4, 4)[2, :3], np.array([0.35, 0.45, 0.70], f32))
    assert s.instances.shape == (1, 36) and s.instances[0, 32:34].view(np.uint32).tolist() == [0, 3]
    o, f, r, u = d.camera_basis()                                                # reference_scene.rs:271-275
    assert abs(float(f @ r)) < 1e-6 and abs(float(f @ u)) < 1e-6 and abs(float(r @ u)) < 1e-6 and abs(float(np.linalg.norm(f)) - 1) < 1e-6
    v, t = d.plane_mesh()
    for tri in t:                                                                # plane_winding_points_up, :338-347
        assert np.cross(v[tri[1]] - v[tri[0]], v[tri[2]] - v[tri[0]])[1] > 0
    m = d.metadata_fields(8, 4, 2)
    assert {"ambient_r", "sky_b", "sun_dir_x", "width", "spp"} <= set(m) and "env_ground_r" not in m
    assert abs(d.fov_y_rad() - np.deg2rad(40.0)) < 1e-6


def test_argument_errors_need_no_device():
    with pytest.raises(ValueError, match="non-zero width/height/spp"):
        wf.render_pt_reference(wf.adjudication_scene(), 0, 8, 1)
    with pytest.raises(ValueError, match="width > 0, height > 0, spp > 0"):
        wf.render_adjudication_pt(8, 8, 0)


# ------------------------------------------------------------------ the CUDA build on the SIMT interpreter
def _compare(native, scene, w, h, spp):
    o = oracle.wavefront_render(scene, w, h, spp)
    hdr, rgba, st = wf.render_pt_reference(scene, w, h, spp, return_rgba8=True, return_stats=True)
    if not np.array_equal(_bits(hdr), _bits(o["hdr"])):         # say how far apart, for whoever reads the GPU log
        bad = (hdr.view(np.uint32) != o["hdr"].view(np.uint32)).any(axis=2)
        rel = np.abs(hdr - o["hdr"]) / np.maximum(np.abs(o["hdr"]), 1e-12)
        first = tuple(int(v) for v in np.argwhere(bad)[0])
        raise AssertionError(f"radiance differs from the oracle on {int(bad.sum())} of {bad.size} pixels, max rel {float(rel.max()):.3e}, "
                             f"first at (y, x) = {first}: got {hdr[first]}, want {o['hdr'][first]}; rays {st.rays} vs {o['rays']}")
    assert np.array_equal(rgba, o["rgba8"])
    assert (st.rays, st.max_rays_per_frame, st.min_iterations) == (o["rays"], o["max_rays_per_frame"], o["min_iterations"])
    return o, st


def test_emulated_cuda_build_is_bit_identical_on_the_adjudication_scene(monkeypatch):
    with _emu.emulated_backend() as native:
        o, st = _compare(native, _adjudication(), 48, 40, 3)     # small image: all three frames share one batch
        assert st.launches == 6 + 1 and o["min_iterations"] >= 2
        _compare(native, _adjudication(), 33, 7, 2)              # ragged: not a multiple of the warp or the block
        for batch, launches in (("1", 5 * 6 + 1), ("2", 3 * 6 + 1), ("16", 6 + 1)):   # the image does not depend on the batch size
            monkeypatch.setenv("F3D_B200_WF_BATCH", batch)
            _, st = _compare(native, _adjudication(), 40, 24, 5)
            assert st.launches == launches
        monkeypatch.delenv("F3D_B200_WF_BATCH")
        for wide in ("1", "2", "9", "15"):                       # ... nor on where the compacted waves hand over to the tail kernel
            monkeypatch.setenv("F3D_B200_WF_WIDE_DEPTH", wide)
            _, st = _compare(native, _rich_scene(False, 2), 32, 24, 3)
            assert st.launches == int(wide) + 2 + 1


@pytest.mark.parametrize("instanced,res", [(False, 6), (True, 6), (False, 2), (True, 1)])
def test_emulated_cuda_build_is_bit_identical_on_every_branch(instanced, res, monkeypatch):
    with _emu.emulated_backend() as native:
        _compare(native, _rich_scene(instanced, res), 40, 30, 3)
        if res == 6:                                             # the BVH only prunes: same image without it
            monkeypatch.setenv("F3D_B200_NO_MESH_BVH", "1")
            _compare(native, _rich_scene(instanced, res), 40, 30, 1)


def _empty_scene():
    s = _adjudication()
    s.spheres[:, 3] = 0.0
    s.mesh_idx = np.zeros((0, 3), np.uint32)
    s.mesh_xyz = np.zeros((0, 3), f32)
    s.instances = np.zeros((0, 36), f32)
    return s.normalized()


def _furnace_scene():
    s = _adjudication()                                          # paths rattle between the ground plane and the underside of a huge sphere
    s.spheres[:, 3] = 0.0
    s.spheres[0, :4] = [0.0, 1005.0, 0.0, 1000.0]
    s.spheres[:, 4:7] = 0.99
    return s.normalized()


def test_emulated_frame_rules_raise_the_reference_errors():
    with _emu.emulated_backend():
        with pytest.raises(RuntimeError, match=r"adjudication PT frame 0 executed 1 wavefront iteration\(s\); a multi-bounce"):
            wf.render_pt_reference(_empty_scene(), 16, 8, 2)
        with pytest.raises(RuntimeError, match=r"wavefront frame 0: wavefront ray queue overflow: \d+ rays pushed into capacity 512"):
            wf.render_pt_reference(_furnace_scene(), 16, 8, 1)
        s = _adjudication()
        s.spheres[0, 15], s.spheres[0, 16] = 0.3, 0.1
        with pytest.raises(ValueError, match="anisotropic GGX"):
            wf.render_pt_reference(s, 8, 8, 1)
        s = _adjudication()
        s.spheres = np.zeros((0, 20), f32)
        with pytest.raises(ValueError, match="at least one sphere / material slot"):
            wf.render_pt_reference(s, 8, 8, 1)
        s = _adjudication()
        s.mesh_idx = np.array([[0, 1, 9]], np.uint32)
        with pytest.raises(ValueError, match="mesh index 9 out of range"):
            wf.render_pt_reference(s, 8, 8, 1)
    for scene, text in ((_empty_scene(), "executed 1 wavefront iteration"), (_furnace_scene(), "ray queue overflow")):
        with pytest.raises(oracle.OracleError, match=text) as want:
            oracle.wavefront_render(scene, 16, 8, 2)
        with _emu.emulated_backend():
            with pytest.raises(RuntimeError) as got:
                wf.render_pt_reference(scene, 16, 8, 2)
        assert str(got.value) == str(want.value)                 # same frame, same ray count in the text


# ------------------------------------------------------------------ image partition (multi-GPU extension)
def _assemble_parts(scene, w, h, spp, world, block_rows):
    from forge3d_b200 import distributed as D

    hdr = np.zeros((h, w, 4), f32)
    rgba = np.zeros((h, w, 4), np.uint8)
    iters, rays = np.zeros(spp, np.int64), np.zeros(spp, np.int64)
    for rank in range(world):
        ph, pr, st = wf.render_pt_reference(scene, w, h, spp, return_rgba8=True, return_stats=True, part=(rank, world, block_rows))
        rows = D.wavefront_owned_rows(h, world, rank, block_rows)
        hdr[rows], rgba[rows] = ph[rows], pr[rows]
        other = np.setdiff1d(np.arange(h), rows)
        assert not ph[other, :, :3].any()                        # a rank never touches rows it does not own
        iters = np.maximum(iters, st.frame_iterations.astype(np.int64))
        rays += st.frame_rays.astype(np.int64)
    return hdr, rgba, iters, rays


@pytest.mark.parametrize("world,block_rows,h", [(2, 16, 40), (3, 4, 30), (4, 16, 24)])   # last: rank 2 and 3 own nothing or a ragged block
def test_emulated_row_partition_is_bit_identical_to_one_device(world, block_rows, h):
    scene = _rich_scene(True, 6)
    o = oracle.wavefront_render(scene, 36, h, 2)
    with _emu.emulated_backend():
        hdr, rgba, iters, rays = _assemble_parts(scene, 36, h, 2, world, block_rows)
    assert np.array_equal(_bits(hdr), _bits(o["hdr"])) and np.array_equal(rgba, o["rgba8"])
    assert int(rays.sum()) == o["rays"] and int(rays.max()) == o["max_rays_per_frame"] and int(iters.min()) == o["min_iterations"]
    wf.check_frame_rules(iters, rays, 36, h)


def _free_port():
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import os
    import sys

    sys.path[:0] = [str(Path(__file__).parent), str(Path(__file__).parent.parent)]
    import torch.distributed as dist

    from forge3d_b200 import distributed as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with _emu.emulated_backend():
            hdr, rgba = D.wavefront_partitioned(_rich_scene(False, 6), 36, 30, 2, block_rows=4)
            err = ""
            try:
                D.wavefront_partitioned(_empty_scene(), 16, 8, 1, block_rows=4)   # every rank must raise together
            except RuntimeError as e:
                err = str(e)
        q.put((rank, hdr, rgba, err))
    finally:
        dist.destroy_process_group()


def test_gloo_two_ranks_assemble_the_one_device_image():
    import torch.multiprocessing as mp

    o = oracle.wavefront_render(_rich_scene(False, 6), 36, 30, 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, hdr, rgba, err in results:
        assert np.array_equal(_bits(hdr), _bits(o["hdr"])) and np.array_equal(rgba, o["rgba8"]), rank
        assert "executed 1 wavefront iteration" in err, (rank, err)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_is_bit_identical_to_the_oracle():
    _compare(None, _adjudication(), 160, 120, 8)
    _compare(None, _rich_scene(False, 24), 128, 96, 6)           # 1152 triangles through the BVH
    _compare(None, _rich_scene(True, 6), 96, 64, 4)
    _compare(None, _adjudication(), 33, 7, 2)


@pytest.mark.gpu
def test_gpu_gate_render_equals_the_committed_oracle_render():
    """512 x 512 x 4096 spp, the reference gate's configuration: the GPU image must be the oracle's byte for byte (the committed
    tests/golden/wavefront_oracle_512.png, itself inside the reference's drift gate against the reference's golden)."""
    want = read_png(GOLDEN / "wavefront_oracle_512.png")
    rgba, meta = wf.render_adjudication_pt(512, 512, 4096)
    assert rgba.shape == (512, 512, 4) and rgba.dtype == np.uint8
    diff = np.abs(rgba.astype(np.int16) - want.astype(np.int16))
    assert np.array_equal(rgba, want), f"{int((diff > 0).any(axis=2).sum())} pixels differ, max |diff| {int(diff.max())}"
    assert meta["pt"]["spp"] == 4096.0 and meta["pt"]["sky_b"] == float(f32(0.70))


@pytest.mark.gpu
def test_gpu_row_partition_is_bit_identical_to_one_device():
    scene = _rich_scene(True, 6)
    whole = wf.render_pt_reference(scene, 96, 70, 3)
    hdr, rgba, iters, rays = _assemble_parts(scene, 96, 70, 3, 3, 16)
    assert np.array_equal(_bits(hdr), _bits(whole))
    wf.check_frame_rules(iters, rays, 96, 70)


@pytest.mark.gpu
def test_gpu_frame_rules_raise_the_reference_errors():
    with pytest.raises(RuntimeError, match="executed 1 wavefront iteration"):
        wf.render_pt_reference(_empty_scene(), 64, 32, 2)
    with pytest.raises(RuntimeError, match="ray queue overflow"):
        wf.render_pt_reference(_furnace_scene(), 64, 32, 1)
