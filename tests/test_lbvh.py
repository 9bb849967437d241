"""GPU LBVH build (second half; the north-star's `src/accel`), CPU suite: the oracle against known Morton
answers and tree invariants, the product's CUDA build kernels under the SIMT interpreter against the oracle bit for bit, and renders
through the LBVH against the oracle's index-order sweep.  The -m gpu tests repeat the comparisons on the device."""
import os

import numpy as np
import pytest

import _emu
import _helpers as H
from oracle import oracle


def _mesh(n, seed, clusters=0):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-40.0, 40.0, (n, 3)).astype(np.float32)
    if clusters:                                             # many triangles share a centroid cell: equal Morton codes
        c = c[rng.integers(0, clusters, n)]
    v = (c[:, None, :] + rng.uniform(-1.0, 1.0, (n, 3, 3)).astype(np.float32))
    if clusters:
        v = np.repeat(c[:, None, :], 3, axis=1) + np.tile(np.array([[0.5, 0, 0], [-0.25, 0.4, 0], [-0.25, -0.4, 0.1]], np.float32), (n, 1, 1))
    return v.reshape(-1, 3).astype(np.float32), np.arange(3 * n, dtype=np.uint32).reshape(n, 3)


def _check_tree(b, n, verts, tris):
    """Every node has one parent, every leaf is reachable exactly once, every box contains its subtree's triangles."""
    if n == 1:
        assert int(b["nodes"].view(np.uint32)[0, 3]) == 0x80000000
        return
    seen = np.zeros(2 * n - 1, np.int32)
    stack = [0]
    while stack:
        x = stack.pop()
        seen[x] += 1
        if x < n - 1:
            for child in (int(b["left"][x]), int(b["right"][x])):
                assert int(b["parent"][child]) == x
                lo, hi = b["nodes"][child, :3], b["nodes"][child, 4:7]
                assert (b["nodes"][x, :3] <= lo).all() and (b["nodes"][x, 4:7] >= hi).all()
                stack.append(child)
    assert (seen == 1).all()
    assert sorted(b["order"].tolist()) == list(range(n))
    leaf_tri = b["order"]
    tri_lo = verts[tris].min(axis=1)[leaf_tri]
    tri_hi = verts[tris].max(axis=1)[leaf_tri]
    assert (b["nodes"][n - 1:, :3] <= tri_lo).all() and (b["nodes"][n - 1:, 4:7] >= tri_hi).all()


def test_morton_known_answers_and_order():
    # lbvh_morton.wgsl:24-38: bit k of x lands on bit 3k, y on 3k+1, z on 3k+2
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 0], [0, 1, 0], [0, 0, 1], [3, 3, 3], [3, 3, 2.99], [3, 2.99, 3]], np.float32)
    t = np.array([[0, 1, 2], [3, 4, 5], [6, 7, 8]], np.uint32)
    b = oracle.lbvh_build(v, t)
    # centroids (1/3, 1/3, 0), (0, 1/3, 1/3), (3, 2.9967, 2.9967) in a [0,3]^3 box -> grid (113, 113, 0), (0, 113, 113), (1023, 1021, 1021)
    def morton(x, y, z):
        out = 0
        for k in range(10):
            out |= ((x >> k) & 1) << (3 * k) | ((y >> k) & 1) << (3 * k + 1) | ((z >> k) & 1) << (3 * k + 2)
        return out
    want = sorted([(morton(113, 113, 0), 0), (morton(0, 113, 113), 1), (morton(1023, 1021, 1021), 2)])
    assert b["morton"].tolist() == [w[0] for w in want] and b["order"].tolist() == [w[1] for w in want]
    assert morton(1023, 1023, 1023) == 0x3FFFFFFF


@pytest.mark.parametrize("n,clusters", [(1, 0), (2, 0), (3, 0), (64, 0), (1000, 0), (777, 5), (300, 1)])
def test_oracle_tree_invariants_and_split_rules(n, clusters):
    v, t = _mesh(n, n + clusters, clusters)
    b = oracle.lbvh_build(v, t)
    _check_tree(b, n, v, t)
    assert (np.diff(b["morton"].astype(np.int64)) >= 0).all()
    lit = oracle.lbvh_build(v, t, literal_split=True)
    if len(np.unique(b["morton"])) == n:       # distinct codes: the composite-key split IS the shader's find_split
        assert np.array_equal(lit["left"], b["left"]) and np.array_equal(lit["right"], b["right"])
    raw = oracle.lbvh_build(v, t, pad_boxes=False)
    assert np.array_equal(raw["order"], b["order"]) and (raw["nodes"][:, :3] >= b["nodes"][:, :3]).all()


@pytest.mark.parametrize("n,clusters", [(1, 0), (2, 0), (5, 0), (257, 0), (3000, 0), (900, 7), (200, 1)])
def test_emulated_cuda_build_is_bit_identical_to_the_oracle(n, clusters):
    v, t = _mesh(n, 100 + n, clusters)
    o = oracle.lbvh_build(v, t)
    with _emu.emulated_backend() as native:
        g = native.lbvh_build(v, t)
    for k in o:
        assert np.array_equal(g[k].view(np.uint32), o[k].view(np.uint32)), k
    _check_tree(g, n, v, t)


def _bumpy_mesh(res, seed):
    """A wavy sheet of 2 * res^2 triangles hovering over the golden DEM scene."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-30.0, 30.0, res + 1, dtype=np.float32)
    x, z = np.meshgrid(xs, xs)
    y = (16.0 + 2.5 * np.sin(x * 0.4) * np.cos(z * 0.3) + rng.uniform(-0.3, 0.3, x.shape)).astype(np.float32)
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    i = (np.arange(res)[:, None] * (res + 1) + np.arange(res)[None, :]).reshape(-1)
    tris = np.concatenate([np.stack([i, i + 1, i + res + 2], 1), np.stack([i, i + res + 2, i + res + 1], 1)]).astype(np.uint32)
    return verts, tris[rng.permutation(len(tris))]


def test_emulated_render_through_the_lbvh_matches_the_index_order_sweep(monkeypatch):
    dem = H.golden_dem()
    verts, tris = _bumpy_mesh(12, 3)                          # 288 triangles
    kw = {**H.scene_kwargs(dem), "mesh_vertices": verts, "mesh_indices": tris, "max_frames": 3, "min_frames": 3, "variance_threshold": 1e30}
    o = oracle.render(dem, 48, 36, H.CAM, want_accum=True, **kw)
    out = {}
    with _emu.emulated_backend() as native:
        for mode in ("lbvh", "host"):
            monkeypatch.setenv("F3D_B200_MESH_BVH", mode)
            out[mode] = native.hybrid_render_terrain_reference(dem, 48, 36, H.CAM, want_accum=True, **kw)
    for mode, g in out.items():
        for k in ("rgba", "depth", "normal", "albedo", "accum"):
            assert np.array_equal(np.ascontiguousarray(g[k]).view(np.uint8), np.ascontiguousarray(o[k]).view(np.uint8)), (mode, k)
    assert (o["albedo"][..., 2] > 0.75).mean() > 0.05           # the mesh (albedo 0.7, 0.7, 0.8) is really in view


@pytest.mark.gpu
def test_gpu_build_is_bit_identical_to_the_oracle():
    from forge3d_b200 import _native

    for n, clusters in [(1, 0), (2, 0), (4097, 0), (60_000, 0), (20_000, 40)]:
        v, t = _mesh(n, 7 + n, clusters)
        o, g = oracle.lbvh_build(v, t), _native.lbvh_build(v, t)
        for k in o:
            assert np.array_equal(g[k].view(np.uint32), o[k].view(np.uint32)), (n, k)


@pytest.mark.gpu
def test_gpu_render_through_the_lbvh_matches_the_index_order_sweep():
    from forge3d_b200 import _native

    dem = H.golden_dem()
    verts, tris = _bumpy_mesh(48, 4)                          # 4608 triangles >= the LBVH threshold: the default path
    kw = {**H.scene_kwargs(dem), "mesh_vertices": verts, "mesh_indices": tris, "max_frames": 3, "min_frames": 3, "variance_threshold": 1e30}
    o = oracle.render(dem, 96, 72, H.CAM, want_accum=True, **kw)
    g = _native.hybrid_render_terrain_reference(dem, 96, 72, H.CAM, want_accum=True, **kw)
    for k in ("rgba", "depth", "normal", "albedo", "accum"):
        assert np.array_equal(np.ascontiguousarray(g[k]).view(np.uint8), np.ascontiguousarray(o[k]).view(np.uint8)), k
    assert (o["albedo"][..., 2] > 0.75).mean() > 0.05
