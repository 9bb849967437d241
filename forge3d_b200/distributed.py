"""Image-row partition of one frame across the GPUs of a node (one process per GPU, torch.distributed).

The reference is single-device (no NCCL/MPI anywhere); this is new work specified by
BASELINE.json's north_star and SURVEY section 8e.  Design:

  * rows are dealt to ranks in interleaved blocks (block b -> rank b % world) so sky and terrain
    rows balance; every index/seed computation uses GLOBAL pixel coordinates, so a partitioned
    render is bit-identical to the single-GPU render;
  * the only per-frame cross-pixel dependency is the spatial reuse pass reading temporal records
    within -3..+4 px (pt_restir_spatial.wgsl:170-171; +4 because xorshift32 can return exactly 1.0).
    k_primary stores the 4 top / 3 bottom border rows of each owned
    block straight into the neighbours' images over NVLink peer memory (CUDA IPC mappings) while it
    computes -- no staging copy, no separate exchange kernel;
  * a frame may only start when the neighbours' previous frame (and its halo stores) completed: the
    last CTA of k_primary publishes "frame f done" into the neighbours' memory after a system-scope
    fence and the next k_primary spins on its local copy (csrc/f3d_kernels.cuh: wait_neighbours /
    signal_neighbours) -- no host round trip and no NCCL call inside the frame loop;
  * the convergence gate needs max over ranks of the windowed variance: one NCCL MAX all-reduce per
    32-frame window; validity flags are OR-reduced the same way;
  * the framebuffer is assembled by ONE NCCL all-gather of the packed owned rows (RGBA8 + AOVs).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

TILE_H = 16  # kTileH in csrc/f3d_kernels.cuh: block_rows is rounded up to a multiple of it


def effective_block_rows(block_rows: int, height: int, world: int) -> int:
    """Mirrors session_create_impl (csrc/f3d_backend.cu): default 16, multiple of 16; one block when world == 1."""
    if world <= 1:
        return ((height + TILE_H - 1) // TILE_H) * TILE_H
    b = block_rows if block_rows else 16
    return ((b + TILE_H - 1) // TILE_H) * TILE_H


def owned_rows(height: int, world: int, rank: int, block_rows: int = 0) -> np.ndarray:
    """Global row indices rendered by `rank` (ascending)."""
    br = effective_block_rows(block_rows, height, world)
    rows = np.arange(height)
    return rows[(rows // br) % max(world, 1) == rank]


def row_counts(height: int, world: int, block_rows: int = 0) -> List[int]:
    return [int(owned_rows(height, world, r, block_rows).size) for r in range(world)]


def pack_rows(image, rows):
    """image: (H, ...) tensor/array; returns the owned rows as one contiguous block."""
    return image[rows]


def assemble(gathered, height: int, world: int, block_rows: int = 0):
    """gathered[r]: (max_rows, ...) block of rank r (padded); returns the (H, ...) image."""
    first = gathered[0]
    if hasattr(first, "new_zeros"):
        out = first.new_zeros((height,) + tuple(first.shape[1:]))
    else:
        out = np.zeros((height,) + tuple(first.shape[1:]), dtype=first.dtype)
    for r in range(world):
        rows = owned_rows(height, world, r, block_rows)
        if hasattr(first, "new_zeros"):
            import torch

            out[torch.as_tensor(rows, device=first.device)] = gathered[r][: rows.size]
        else:
            out[rows] = gathered[r][: rows.size]
    return out


def gather_rows(local_image, height: int, group=None, block_rows: int = 0):
    """ONE all-gather of every rank's packed owned rows; every rank returns the assembled image.
    Works on CUDA tensors (NCCL) and CPU tensors (gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = row_counts(height, world, block_rows)
    mx = max(counts)
    rows = torch.as_tensor(owned_rows(height, world, rank, block_rows), device=local_image.device)
    packed = local_image.new_zeros((mx,) + tuple(local_image.shape[1:]))
    packed[: rows.numel()] = local_image[rows]
    out = local_image.new_empty((world * mx,) + tuple(local_image.shape[1:]))
    dist.all_gather_into_tensor(out, packed.contiguous(), group=group)
    return assemble([out[r * mx:(r + 1) * mx] for r in range(world)], height, world, block_rows)


NO_VALID_RESERVOIRS = ("terrain PT ReSTIR reuse chain produced no valid reservoirs for a sun-lit scene "
                       "\u2014 temporal/spatial reuse is broken")


def check_validity_across_ranks(any_valid: bool, required: bool, group=None, device="cpu") -> None:
    """The reference's end-of-render reservoir check (render_terrain.rs:1313-1337) for a row partition: a rank
    may legitimately own only sky, so the per-rank flags are OR-reduced (one MAX all-reduce of one int32) and
    every rank raises the reference's error together when the scene is sun-lit and no rank holds a valid
    reservoir.  Works over NCCL (device="cuda") and gloo (device="cpu")."""
    import torch
    import torch.distributed as dist

    flag = torch.tensor([1 if any_valid else 0], dtype=torch.int32, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    if required and int(flag.item()) == 0:
        raise RuntimeError(NO_VALID_RESERVOIRS)


def wavefront_block_rows(block_rows: int, world: int) -> int:
    """Row-block size of the wavefront tracer's partition: any positive value works (no tile constraint); default 16."""
    return int(block_rows) if block_rows else 16


def wavefront_owned_rows(height: int, world: int, rank: int, block_rows: int = 0) -> np.ndarray:
    br = wavefront_block_rows(block_rows, world)
    rows = np.arange(height)
    return rows if world <= 1 else rows[(rows // br) % world == rank]


def wavefront_partitioned(scene, width: int, height: int, spp_frames: int, *, group=None, block_rows: int = 0, device: int = 0,
                          tensor_device: str = "cpu"):
    """The wavefront multi-bounce tracer (forge3d_b200.wavefront.render_pt_reference) over the ranks of a process group: each rank
    traces the rows it owns (pixels are independent, so they are bit-identical to the one-GPU image; no data-path collective), then
    ONE all-gather per image assembles (hdr f32, rgba8) on every rank, and the reference's two per-frame rules are applied to the
    per-frame iteration counts (MAX over ranks) and ray counts (SUM over ranks).  NCCL with tensor_device="cuda", gloo on the CPU."""
    import torch
    import torch.distributed as dist

    from . import wavefront as wf

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    br = wavefront_block_rows(block_rows, world)
    hdr, rgba, st = wf.render_pt_reference(scene, width, height, spp_frames, device=device, return_rgba8=True, return_stats=True,
                                           part=(rank, world, br))
    iters = torch.from_numpy(st.frame_iterations.astype(np.int64)).to(tensor_device)
    rays = torch.from_numpy(st.frame_rays.astype(np.int64)).to(tensor_device)
    if world > 1:
        dist.all_reduce(iters, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM, group=group)
    wf.check_frame_rules(iters.cpu().numpy(), rays.cpu().numpy(), width, height)
    if world == 1:
        return hdr, rgba
    rows = torch.as_tensor(wavefront_owned_rows(height, world, rank, br), device=tensor_device)
    counts = [int(wavefront_owned_rows(height, world, r, br).size) for r in range(world)]
    mx = max(counts)
    out = []
    for img in (hdr, rgba):
        t = torch.from_numpy(img).to(tensor_device)
        packed = t.new_zeros((mx,) + tuple(t.shape[1:]))
        packed[: rows.numel()] = t[rows]
        allr = t.new_empty((world * mx,) + tuple(t.shape[1:]))
        dist.all_gather_into_tensor(allr, packed.contiguous(), group=group)
        full = t.new_zeros(t.shape)
        for r in range(world):
            rr = torch.as_tensor(wavefront_owned_rows(height, world, r, br), device=tensor_device)
            full[rr] = allr[r * mx:r * mx + rr.numel()]
        out.append(full.cpu().numpy())
    return out[0], out[1]


def gather_rows_to(local_image, height: int, dst: int = 0, group=None, block_rows: int = 0):
    """ONE gather of every rank's packed owned rows to rank `dst`; that rank returns the assembled (H, ...) image, the others
    None.  NCCL on CUDA tensors, gloo on CPU tensors."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = row_counts(height, world, block_rows)
    mx = max(counts)
    rows = torch.as_tensor(owned_rows(height, world, rank, block_rows), device=local_image.device)
    packed = local_image.new_zeros((mx,) + tuple(local_image.shape[1:]))
    packed[: rows.numel()] = local_image[rows]
    parts = [torch.empty_like(packed) for _ in range(world)] if rank == dst else None
    dist.gather(packed.contiguous(), parts, dst=dst, group=group)
    if rank != dst:
        return None
    return assemble(parts, height, world, block_rows)


def _to_host(t):
    """Device tensor -> numpy array in page-locked memory from the library's pool (_native.host_array): one DMA, no pageable
    bounce, no first-touch page faults; the block goes back to the pool when the array is garbage-collected."""
    import torch

    from . import _native

    if t.device.type != "cuda":
        return t.numpy()
    arr = _native.host_array(tuple(t.shape), np.dtype(str(t.dtype).replace("torch.", "")))
    torch.from_numpy(arr).copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return arr


class PartitionedRender:
    """One rank's share of a partitioned render (CUDA + NCCL)."""

    def __init__(self, heightmap, width, height, cam=None, *, block_rows: int = 0, group=None, mode: str = "exact", **scene_kw):
        """mode="exact" (default): bit-identical to one GPU - halo rows of the reservoir image and the frame flags travel over
        NVLink peer memory inside k_shade / k_primary.  mode="gather_only" (SURVEY section 8e-ii): the spatial reuse pass stays
        inside each row block, ranks exchange NOTHING during the frame loop, the only collective is the final gather; pixels
        within 4 rows of a block border differ slightly from the one-GPU image (tests/test_multigpu.py bounds the RMSE)."""
        import torch
        import torch.distributed as dist

        from . import _native
        from .session import Session

        import time

        self.torch, self.dist, self.group = torch, dist, group
        self.timings = {}            # wall-clock ms of the host-visible phases (set-up and resolve), for bench.py's e2e record
        t_mark = time.perf_counter()

        def lap(name):
            nonlocal t_mark
            torch.cuda.synchronize()
            now = time.perf_counter()
            self.timings[name] = self.timings.get(name, 0.0) + (now - t_mark) * 1e3
            t_mark = now

        initialised = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if initialised else 1
        self.rank = dist.get_rank(group) if initialised else 0
        if mode not in ("exact", "gather_only"):
            raise ValueError(f"mode must be 'exact' or 'gather_only', got {mode!r}")
        self.mode = mode
        self.width, self.height, self.block_rows = int(width), int(height), int(block_rows)
        self.device = torch.cuda.current_device()
        self.stream = torch.cuda.current_stream()
        if self.world > 1:
            # The DEM crosses PCIe ONCE (rank 0) and reaches the other GPUs over NVLink (one NCCL broadcast); every session
            # then reads it in place instead of staging its own pageable host copy.
            shape = tuple(np.shape(heightmap))
            if self.rank == 0:
                host = torch.from_numpy(np.ascontiguousarray(heightmap, dtype=np.float32))
                dem_dev = host.to("cuda", non_blocking=False)
            else:
                dem_dev = torch.empty(shape, dtype=torch.float32, device="cuda")
            dist.broadcast(dem_dev, src=0, group=group)
            self._dem_dev = dem_dev
            heightmap = _native.DeviceHeights(dem_dev.data_ptr(), shape, keep=dem_dev)
            lap("dem_h2d_broadcast_ms")
        # torch's default stream has handle 0, which the C ABI reads as "create your own stream":
        # pass cudaStreamLegacy (0x1) so kernels, torch events and NCCL share one stream.
        self.session = Session(heightmap, width, height, cam, device=self.device,
                               cuda_stream=self.stream.cuda_stream or 1, part_rank=self.rank, part_world=self.world,
                               part_block_rows=block_rows, part_mode=1 if mode == "gather_only" else 0, **scene_kw)
        lap("session_create_ms")
        if self.world > 1 and mode == "exact":
            # CUDA-IPC handles of the reservoir images and the frame-barrier words: one all-gather of 192 bytes per rank
            mine = torch.frombuffer(bytearray(self.session.ipc_export()), dtype=torch.uint8).cuda()
            every = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(every, mine, group=group)
            self.session.ipc_import(every.cpu().numpy().tobytes())
            lap("ipc_exchange_ms")

    def render_frames(self, n: int) -> None:
        """n accumulation frames, enqueued back to back; with world > 1 the cross-GPU ordering is done
        on the devices by k_primary's peer-memory frame barrier."""
        self.session.render_frames(n)

    def variance(self) -> Tuple[float, bool]:
        v, bad = self.session.variance()
        if self.world > 1:
            t = self.torch.tensor([v, 1.0 if bad else 0.0], dtype=self.torch.float32, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
            v, bad = float(t[0]), bool(t[1] > 0)
        return v, bad

    def resolve(self, aovs: bool = True, dst=None):
        """Resolve owned rows on the device, then ONE collective per output.  dst=None: all-gather, every rank returns the
        numpy images; dst=r: gather to rank r only (the consumer), the other ranks return None and copy nothing to the host."""
        import time

        torch = self.torch
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        H, W = self.height, self.width
        rgba = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
        bufs = {"rgba": rgba}
        if aovs:
            bufs["albedo"] = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
            bufs["normal"] = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
            bufs["depth"] = torch.zeros((H, W), dtype=torch.float32, device="cuda")
        err = None
        try:
            self.session.resolve_device(rgba.data_ptr(), bufs["albedo"].data_ptr() if aovs else 0,
                                        bufs["normal"].data_ptr() if aovs else 0,
                                        bufs["depth"].data_ptr() if aovs else 0, check_validity=True)
        except RuntimeError as exc:      # every rank must reach the collectives below, or the others hang in them
            err = exc
        if self.world > 1:
            any_valid, required = self.session.validity() if err is None else (False, False)
            flags = torch.tensor([1 if err is not None else 0, 1 if any_valid else 0], dtype=torch.int32, device="cuda")
            self.dist.all_reduce(flags, op=self.dist.ReduceOp.MAX, group=self.group)
            if err is not None:
                raise err
            if int(flags[0]) != 0:
                raise RuntimeError("another rank of the partitioned render failed in resolve (see its error)")
            if required and int(flags[1]) == 0:
                raise RuntimeError(NO_VALID_RESERVOIRS)
        elif err is not None:
            raise err
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        self.timings["resolve_kernels_ms"] = (t1 - t0) * 1e3
        out = {}
        for k, t in bufs.items():
            if self.world == 1:
                full = t
            elif dst is None:
                full = gather_rows(t, H, self.group, self.block_rows)
            else:
                full = gather_rows_to(t, H, int(dst), self.group, self.block_rows)
            if full is not None:
                out[k] = _to_host(full) if dst is not None else full.cpu().numpy()
        torch.cuda.synchronize()
        self.timings["gather_and_d2h_ms"] = (time.perf_counter() - t1) * 1e3
        return out if (self.world == 1 or dst is None or self.rank == int(dst)) else None

    def close(self):
        # peers store halo rows and barrier flags into this rank's memory: nobody may free before everybody is done
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.group)
        self.session.close()
