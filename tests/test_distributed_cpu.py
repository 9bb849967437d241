"""Host-side logic of the multi-GPU image partition on CPU: ownership maps and the single
all-gather framebuffer assembly over torch.distributed (gloo, world_size 2 and 3)."""
import os
import socket

import numpy as np
import pytest

from forge3d_b200 import distributed as D


def test_partition_is_a_disjoint_cover_in_interleaved_blocks():
    for height, world, br in [(1080, 8, 0), (1080, 2, 0), (270, 4, 16), (7200, 8, 64), (17, 3, 0), (256, 1, 0)]:
        seen = np.zeros(height, int)
        for r in range(world):
            rows = D.owned_rows(height, world, r, br)
            seen[rows] += 1
            eff = D.effective_block_rows(br, height, world)
            assert ((rows // eff) % world == r).all()
        assert (seen == 1).all()
        assert sum(D.row_counts(height, world, br)) == height
    # load balance of the headline config: no rank owns more than one extra block
    counts = D.row_counts(1080, 8)
    assert max(counts) - min(counts) <= 32


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, height, width, block_rows, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)
        full = rng.integers(0, 255, (height, width, 4), dtype=np.uint8)
        depth = rng.standard_normal((height, width)).astype(np.float32)
        depth[::7] = np.nan
        rows = D.owned_rows(height, world, rank, block_rows)
        local = np.zeros_like(full)
        local[rows] = full[rows]          # a rank only ever fills the rows it owns
        ldepth = np.zeros_like(depth)
        ldepth[rows] = depth[rows]
        got = D.gather_rows(torch.from_numpy(local), height, None, block_rows).numpy()
        gotd = D.gather_rows(torch.from_numpy(ldepth), height, None, block_rows).numpy()
        ok = np.array_equal(got, full) and np.array_equal(gotd.view(np.uint32), depth.view(np.uint32))
        # the variance gate: max over ranks
        t = torch.tensor([float(rank + 1) * 0.25])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t) == world * 0.25
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,height,block_rows", [(2, 270, 0), (3, 100, 16), (2, 33, 0)])
def test_gather_rows_gloo(world, height, block_rows):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, height, 48, block_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(r, True) for r in range(world)]


def test_spatial_offsets_reach_plus_four_and_kernel_halo_covers_it():
    """pt_restir_spatial.wgsl:199-200 draws floor(u*7)-3 with u = f32(x)/2^32.  u32 -> f32 rounds to nearest, so
    x >= 0xFFFFFF80 gives u == 1.0 and the offset +4 (never -4): the multi-GPU halo must be 4 rows towards the
    block above and 3 towards the block below.  Checked against the constants compiled into k_primary."""
    import re
    from pathlib import Path

    x = np.array([0xFFFFFF7F, 0xFFFFFF80, 0xFFFFFFFF, 0, 1], dtype=np.uint32)
    u = x.astype(np.float32) / np.float32(4294967296.0)
    off = np.floor(u * np.float32(7.0)).astype(np.int64) - 3
    assert list(off) == [3, 4, 4, -3, -3] and u[1] == np.float32(1.0)
    src = (Path(__file__).resolve().parent.parent / "forge3d_b200" / "csrc" / "f3d_kernels.cuh").read_text()
    assert re.search(r"in_y < 4u && b > 0u && P\.peer_up", src), "upward halo must be 4 rows"
    assert re.search(r"in_y \+ 3u >= P\.block_rows && b \+ 1u < P\.nblocks && P\.peer_down", src), "downward halo must be 3 rows"


def _validity_worker(rank, world, port, flags, required, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        try:
            D.check_validity_across_ranks(flags[rank], required)
            q.put((rank, "ok"))
        except RuntimeError as exc:
            q.put((rank, str(exc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("flags,required,raises", [
    ((False, True), True, False),     # rank 0 owns only sky: fine, rank 1 has valid reservoirs
    ((False, False), True, True),     # sun-lit scene and nobody has one: every rank raises the reference's error
    ((False, False), False, False),   # sun off / below horizon: the check does not apply
])
def test_reservoir_validity_is_or_reduced_over_ranks_gloo(flags, required, raises):
    """render_terrain.rs:1313-1337 under a row partition: OR over ranks, same error text."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_validity_worker, args=(r, 2, port, flags, required, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(2):
        if raises:
            assert "produced no valid reservoirs for a sun-lit scene" in results[rank]
        else:
            assert results[rank] == "ok"


def test_validity_check_without_a_process_group_is_the_single_rank_rule():
    D.check_validity_across_ranks(True, True)
    D.check_validity_across_ranks(False, False)
    with pytest.raises(RuntimeError, match="no valid reservoirs"):
        D.check_validity_across_ranks(False, True)
