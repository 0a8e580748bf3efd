"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed (NCCL over NVLink on
the GPU box, gloo in CPU tests).  Two sharding modes, neither needs a collective on the data path:

* frame sharding  -- every rank accumulates a contiguous block of the frame pairs into a PRIVATE full grid; the run
  ends with ONE sum-reduce of the N^3 cells to rank 0 (fp32: tolerance mode, int64 fixed point: bit-exact and
  order independent).
* slab sharding   -- every rank owns the x-planes [lo, hi) of a grid too large for one GPU; all ranks walk all rays with
  full-grid arithmetic and suppress writes outside their slab (a clipped restart would change the reference's
  accumulated t_max roundings).  No reduce: x is the slowest index (ray_voxel.cpp:527), so slabs are contiguous byte
  ranges of the .bin and each rank writes its own range.
"""
from __future__ import annotations

import os
import struct

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `total` items for `rank` of `world` (same split as pvp_ray_voxel_accumulate)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    return total * rank // world, total * (rank + 1) // world


def slab_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """x-plane slab [lo, hi) of an n^3 grid owned by `rank`."""
    return shard_range(n, rank, world)


def split_pairs(pairs: list[tuple[int, int]], rank: int, world: int) -> list[tuple[int, int]]:
    """Frame-pair shard of this rank.  Pairs are independent (each needs only its two frames), so a contiguous split
    keeps every rank's frame reads contiguous with one frame of overlap at the shard edge."""
    lo, hi = shard_range(len(pairs), rank, world)
    return pairs[lo:hi]


def reduce_grid(cells: torch.Tensor, dst: int = 0, group=None, all_ranks: bool = False) -> torch.Tensor:
    """The run's single grid merge: in-place sum over ranks (to `dst`, or to everyone with all_ranks)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return cells
    if all_ranks:
        dist.all_reduce(cells, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(cells, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return cells


def write_bin_header(path: str, n: int, voxel_size: float) -> None:
    """Create/size the .bin (ray_voxel.cpp:549-550 header: int32 N, float32 voxel_size) for slab-wise writes."""
    with open(path, "wb") as f:
        f.write(struct.pack("<if", int(n), float(voxel_size)))
        f.truncate(8 + 4 * n ** 3)


def write_bin_slab(path: str, n: int, slab: tuple[int, int], cells_f32: np.ndarray) -> None:
    """Write the float32 x-planes [lo, hi) at their byte range of an existing .bin (payload layout :552)."""
    lo, hi = slab
    a = np.ascontiguousarray(cells_f32, np.float32)
    if a.shape != (hi - lo, n, n):
        raise ValueError(f"slab cells must have shape {(hi - lo, n, n)}, got {a.shape}")
    fd = os.open(path, os.O_WRONLY)
    try:
        os.pwrite(fd, a.tobytes(), 8 + 4 * lo * n * n)
    finally:
        os.close(fd)


def ray_voxel_distributed(metadata_json: str, image_folder: str, output_bin: str | None, mode: str = "frames", n: int = 500,
                          voxel_size: float = 6.0, center=(0.0, 0.0, 500.0), threshold: float = 2.0, accum: str = "f32",
                          frac_bits: int = 16, group=None) -> dict:
    """The ray_voxel pipeline (ray_voxel.cpp:400-558) over all ranks of the process group.  Returns this rank's counters
    summed over ranks (n_steps, n_rays ...).  Rank 0 (frames) / every rank (slabs) writes `output_bin`."""
    from . import motion

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    if mode == "frames":
        grid = motion.VoxelGrid(n, voxel_size, center, accum=accum, frac_bits=frac_bits, device=dev)
        rep = motion.ray_voxel_accumulate(metadata_json, image_folder, grid, threshold, shard=(rank, world))
        reduce_grid(grid.data, dst=0, group=group)
        if rank == 0 and output_bin:
            grid.save(output_bin)
    elif mode == "slabs":
        slab = slab_range(n, rank, world)
        grid = motion.VoxelGrid(n, voxel_size, center, accum=accum, frac_bits=frac_bits, slab=slab, device=dev)
        rep = motion.ray_voxel_accumulate(metadata_json, image_folder, grid, threshold, shard=(0, 1))
        if output_bin:
            if rank == 0:
                write_bin_header(output_bin, n, voxel_size)
            if world > 1:
                dist.barrier(group)
            write_bin_slab(output_bin, n, slab, grid.to_float32().cpu().numpy())
            if world > 1:
                dist.barrier(group)
    else:
        raise ValueError("mode must be 'frames' or 'slabs'")
    keys = ["n_rays", "n_steps", "n_written", "n_atomics", "n_pairs", "n_frames_loaded"]
    t = torch.tensor([rep[k] for k in keys], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, group=group)
    out = dict(zip(keys, (int(x) for x in t.tolist())))
    if mode == "slabs":  # every rank walked every ray: ray/step counts are replicated, writes are not
        for k in ("n_rays", "n_steps", "n_pairs", "n_frames_loaded"):
            out[k] //= world
    return out
