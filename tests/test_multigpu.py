"""Multi-GPU image partition on real devices (NCCL): a render dealt to N ranks in interleaved row blocks,
with halo rows and the frame barrier pushed over NVLink peer memory inside k_primary, must be
BIT-IDENTICAL to the single-GPU render (SURVEY section 8e "exact mode").  Skipped when too few GPUs are visible."""
import os
import socket

import numpy as np
import pytest

import _helpers as H

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, frames, width, height, block_rows, q, seed=7, mode="exact"):
    import torch
    import torch.distributed as dist

    from forge3d_b200.distributed import PartitionedRender

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dem = H.golden_dem()
        kw = {**H.scene_kwargs(dem), "max_frames": frames, "min_frames": frames, "variance_threshold": 1e30, "seed": seed}
        pr = PartitionedRender(dem, width, height, H.CAM, block_rows=block_rows, mode=mode, **kw)
        pr.render_frames(frames)
        var, bad = pr.variance()
        out = pr.resolve(aovs=True)
        pr.close()
        if rank == 0:
            q.put({"rgba": out["rgba"], "depth": out["depth"], "normal": out["normal"], "albedo": out["albedo"],
                   "variance": var, "bad": bad})
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,width,height,block_rows,frames", [(2, 96, 80, 16, 40), (2, 64, 50, 32, 6),
                                                                      (4, 96, 120, 16, 24), (3, 80, 100, 16, 8)])
def test_partition_is_bit_identical(world, width, height, block_rows, frames):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    from forge3d_b200 import _native

    dem = H.golden_dem()
    kw = {**H.scene_kwargs(dem), "max_frames": frames, "min_frames": frames, "variance_threshold": 1e30}
    ref = _native.hybrid_render_terrain_reference(dem, width, height, H.CAM, **kw)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, frames, width, height, block_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=900)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert np.array_equal(got["rgba"], ref["rgba"])
    assert np.array_equal(got["depth"].view(np.uint32), ref["depth"].view(np.uint32))
    assert np.array_equal(got["normal"], ref["normal"]) and np.array_equal(got["albedo"], ref["albedo"])
    assert np.float32(got["variance"]) == np.float32(ref["variance"]) and not got["bad"]


def test_plus4_offset_needs_a_four_row_halo():
    """The reference's neighbour offset floor(u*7)-3 reaches +4 when xorshift32 returns exactly 1.0 (about 3e-8 per
    draw).  At 1920x1080 x 48 frames that happens ~45 times, ~25 % of them across a 16-row block boundary: with a
    3-row halo the partitioned image differs from the 1-GPU image in a few LSBs (observed in round 1)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from forge3d_b200 import _native

    width, height, block_rows, frames = 1920, 1080, 16, 48
    dem = H.golden_dem()
    kw = {**H.scene_kwargs(dem), "max_frames": frames, "min_frames": frames, "variance_threshold": 1e30}
    ref = _native.hybrid_render_terrain_reference(dem, width, height, H.CAM, **kw)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, frames, width, height, block_rows, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=900)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert np.array_equal(got["rgba"], ref["rgba"])
    assert np.float32(got["variance"]) == np.float32(ref["variance"])


def _validity_worker(rank, world, port, cam, height, q):
    import torch
    import torch.distributed as dist

    from forge3d_b200.distributed import PartitionedRender

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dem = H.golden_dem()
        kw = {**H.scene_kwargs(dem), "max_frames": 4, "min_frames": 4, "variance_threshold": 1e30}
        pr = PartitionedRender(dem, 64, height, cam, block_rows=16, **kw)
        pr.render_frames(4)
        try:
            out = pr.resolve(aovs=True)
            q.put((rank, "ok", int(np.isfinite(out["depth"][:16]).sum()), int(np.isfinite(out["depth"][16:]).sum())))
        except RuntimeError as exc:
            q.put((rank, str(exc), 0, 0))
        pr.close()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["rank0_owns_only_sky", "all_sky"])
def test_reservoir_validity_is_checked_across_ranks(case):
    """render_terrain.rs:1313-1337 under the row partition: a rank that owns only sky rows must not fail the
    render, and a sun-lit render in which NO rank holds a valid reservoir raises the reference's error on every rank."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    ox, oy, oz = H.CAM["origin"]
    if case == "all_sky":
        cam = {**H.CAM, "look_at": (ox, oy + 100.0, oz - 1.0)}       # straight up
    else:
        cam = {**H.CAM, "look_at": (0.0, oy, 0.0), "fov_y": 70.0}     # level view: horizon at mid-height
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_validity_worker, args=(r, 2, port, cam, 32, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=900) for _ in range(2)]
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    for rank, status, top_hits, bottom_hits in results:
        if case == "all_sky":
            assert "produced no valid reservoirs for a sun-lit scene" in status
        else:
            assert status == "ok" and top_hits == 0 and bottom_hits > 0, (status, top_hits, bottom_hits)


def _wavefront_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    import test_wavefront as T
    from forge3d_b200 import distributed as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        hdr, rgba = D.wavefront_partitioned(T._rich_scene(True, 6), 96, 70, 3, block_rows=16, device=rank, tensor_device="cuda")
        err = ""
        try:
            D.wavefront_partitioned(T._empty_scene(), 32, 32, 1, block_rows=16, device=rank, tensor_device="cuda")
        except RuntimeError as e:      # every rank raises the reference's error together
            err = str(e)
        q.put((rank, hdr, rgba, err))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_wavefront_tracer_partition_over_nccl_is_bit_identical():
    """SURVEY section 8f row 2 over 2 GPUs: rows dealt in 16-row blocks, no data-path collective, one all-gather per image."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    import test_wavefront as T
    from forge3d_b200 import wavefront as wf

    whole_hdr, whole_rgba = wf.render_pt_reference(T._rich_scene(True, 6), 96, 70, 3, return_rgba8=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_wavefront_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=900) for _ in range(2)]
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    for rank, hdr, rgba, err in results:
        assert np.array_equal(hdr.view(np.uint32), whole_hdr.view(np.uint32)) and np.array_equal(rgba, whole_rgba), rank
        assert "executed 1 wavefront iteration" in err, (rank, err)


def _ipc_worker(rank, world, width, height, frames, block_rows, inbox, outbox, result_q):
    """One rank of a partitioned render WITHOUT NCCL: plain Session + CUDA-IPC handles exchanged through pipes, every rank
    on device 0.  The two processes time-share the GPU, so a neighbour wait costs a context time slice - slow, but it drives
    exactly the product path of a real multi-GPU run: peer mappings, halo stores and the frame flags inside k_primary."""
    from forge3d_b200 import distributed as D
    from forge3d_b200.session import Session

    dem = H.golden_dem()
    kw = {**H.scene_kwargs(dem), "max_frames": frames, "min_frames": frames, "variance_threshold": 1e30}
    s = Session(dem, width, height, H.CAM, device=0, part_rank=rank, part_world=world, part_block_rows=block_rows, **kw)
    outbox.put((rank, s.ipc_export()))
    handles = inbox.get(timeout=120)
    s.ipc_import(handles)
    outbox.put((rank, "imported"))
    assert inbox.get(timeout=120) == "go"
    for _ in range(frames):          # frame by frame: a wait then always refers to work its neighbour has already been given
        s.render_frames(1)
    s.sync()
    var, bad = s.variance()
    out = s.resolve_host()
    rows = D.owned_rows(height, world, rank, block_rows)
    result_q.put((rank, {k: np.array(out[k][rows]) for k in ("rgba", "depth", "normal", "albedo")}, var, bad))
    assert inbox.get(timeout=120) == "done"      # nobody frees memory a peer may still store into
    s.close()


def test_two_ranks_on_one_device_over_cuda_ipc_are_bit_identical():
    """The multi-GPU exact mode on a ONE-GPU box: two processes, both on device 0, linked by CUDA IPC exactly as two GPUs
    are (forge3d_b200.session ipc_export / ipc_import), render interleaved 16-row blocks; the assembled image must equal the
    one-session render bit for bit.  This is the test the driver's single-GPU run can execute; the NCCL tests above need
    2-4 devices."""
    import torch.multiprocessing as mp

    from forge3d_b200 import _native
    from forge3d_b200 import distributed as D

    world, width, height, frames, block_rows = 2, 96, 80, 12, 16
    dem = H.golden_dem()
    kw = {**H.scene_kwargs(dem), "max_frames": frames, "min_frames": frames, "variance_threshold": 1e30}
    ref = _native.hybrid_render_terrain_reference(dem, width, height, H.CAM, **kw)
    ctx = mp.get_context("spawn")
    inboxes = [ctx.Queue() for _ in range(world)]
    outbox, result_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, width, height, frames, block_rows, inboxes[r], outbox, result_q))
             for r in range(world)]
    for p in procs:
        p.start()
    try:
        exported = dict(outbox.get(timeout=300) for _ in range(world))
        blob = b"".join(exported[r] for r in range(world))
        for q in inboxes:
            q.put(blob)
        for _ in range(world):
            assert outbox.get(timeout=300)[1] == "imported"
        for q in inboxes:
            q.put("go")
        parts = dict((r, (imgs, var, bad)) for r, imgs, var, bad in (result_q.get(timeout=600) for _ in range(world)))
        for q in inboxes:
            q.put("done")
    finally:
        for p in procs:
            p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs)
    for key in ("rgba", "depth", "normal", "albedo"):
        full = np.zeros_like(ref[key])
        for r in range(world):
            full[D.owned_rows(height, world, r, block_rows)] = parts[r][0][key]
        assert np.array_equal(full.view(np.uint8), np.ascontiguousarray(ref[key]).view(np.uint8)), key
    assert np.float32(max(parts[r][1] for r in range(world))) == np.float32(ref["variance"])
    assert not any(parts[r][2] for r in range(world))


# ---- gather-only partition: no exchange during the frame loop ----
# MEASURED AND REJECTED as a default: clamping the spatial reuse pass to the row block changes the
# reservoir chain (the reference's M-weighted combination is not invariant to the neighbour set), the change feeds back
# through the temporal pass and spreads ~4 rows per frame: against the one-GPU image of the same seed the RGB RMSE is
# 1.6e-2 (2 bands) to 2.6e-2 (16-row blocks) where two different SEEDS of the one-GPU render differ by 2e-3, and a
# ~6 % darker band follows every seam.  That is 20x the north-star tolerance (1e-3), so the mode stays opt-in and the
# exact mode (peer-memory halo, no NCCL call in the frame loop either) is what bench.py and PartitionedRender default to.
# The tests below pin what the mode does guarantee.
GATHER_ONLY_SANITY_RMSE = 5e-2


def _rgb_rmse(a, b):
    d = a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64)
    return float(np.sqrt(np.mean((d / 255.0) ** 2)))


def _gather_only_rank_by_rank(dem, width, height, world, block_rows, kw):
    """Every rank of a gather-only partition rendered in turn on ONE device (the mode has no communication, so this is the
    real computation), rows assembled as the final gather would."""
    from forge3d_b200 import distributed as D
    from forge3d_b200.session import Session

    got = None
    for r in range(world):
        s = Session(dem, width, height, H.CAM, part_rank=r, part_world=world, part_block_rows=block_rows, part_mode=1, **kw)
        s.render_frames(kw["max_frames"])
        out = s.resolve_host()
        if got is None:
            got = {k: np.zeros_like(out[k]) for k in ("rgba", "depth", "normal", "albedo")}
        rows = D.owned_rows(height, world, r, block_rows)
        for k in got:
            got[k][rows] = out[k][rows]
        s.close()
    return got


@pytest.mark.parametrize("world,block_rows,frames", [(2, 64, 8), (4, 16, 24)])
def test_gather_only_partition_keeps_aovs_exact_and_rgb_close(world, block_rows, frames):
    from forge3d_b200 import _native

    width, height = 128, 128
    dem = H.golden_dem()
    kw = {**H.scene_kwargs(dem), "max_frames": frames, "min_frames": frames, "variance_threshold": 1e30}
    ref = _native.hybrid_render_terrain_reference(dem, width, height, H.CAM, **kw)
    got = _gather_only_rank_by_rank(dem, width, height, world, block_rows, kw)
    for k in ("depth", "normal", "albedo"):          # the AOVs never read a neighbour: identical
        assert np.array_equal(np.ascontiguousarray(got[k]).view(np.uint8), np.ascontiguousarray(ref[k]).view(np.uint8)), k
    rmse = _rgb_rmse(got["rgba"], ref["rgba"])
    assert 0.0 < rmse <= GATHER_ONLY_SANITY_RMSE, rmse      # close, but NOT the same image (see the note above)
    # the change starts at the seams and can travel at most 4 rows per frame: rows further away are the one-GPU rows
    from forge3d_b200 import distributed as D

    br = D.effective_block_rows(block_rows, height, world)
    seams = np.arange(br, height, br)
    y = np.arange(height)
    untouched = np.abs(y[:, None] - seams[None, :] + 0.5).min(axis=1) > 4 * frames + 4
    if untouched.any():
        assert np.array_equal(got["rgba"][untouched], ref["rgba"][untouched])


def test_gather_only_partition_over_nccl_matches_the_rank_by_rank_render():
    """Two GPUs, mode="gather_only": PartitionedRender exchanges nothing but the final gather; the image equals the one the
    single-device rank-by-rank render of the same partition produces, bit for bit."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    world, width, height, block_rows, frames = 2, 96, 80, 16, 40
    dem = H.golden_dem()
    kw = {**H.scene_kwargs(dem), "max_frames": frames, "min_frames": frames, "variance_threshold": 1e30}
    local = _gather_only_rank_by_rank(dem, width, height, world, block_rows, kw)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, frames, width, height, block_rows, q, 7, "gather_only")) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=900)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert np.array_equal(got["rgba"], local["rgba"])
    assert np.array_equal(got["depth"].view(np.uint32), local["depth"].view(np.uint32))
