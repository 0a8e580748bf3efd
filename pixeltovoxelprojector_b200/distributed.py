"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed (NCCL over NVLink on
the GPU box, gloo in CPU tests).  Two sharding modes, neither needs a collective on the data path:

* frame sharding  -- every rank accumulates a contiguous block of the frame pairs into a PRIVATE full grid; the run
  ends with ONE sum-reduce of the N^3 cells to rank 0 (fp32: tolerance mode, int64 fixed point: bit-exact and
  order independent).
* slab sharding   -- every rank owns the x-planes [lo, hi) of a grid too large for one GPU; all ranks see all rays and walk
  them from their start with full-grid arithmetic (a walk restarted at the slab face would change the reference's
  accumulated t_max roundings), writing only inside their slab.  The clipping is exact: a ray whose x range misses the
  slab is not walked, and a walk stops once x has passed the slab.  No reduce: x is the slowest index
  (ray_voxel.cpp:527), so slabs are contiguous byte ranges of the .bin and each rank writes its own range.
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


def balanced_slabs(plane_work, world: int) -> list[tuple[int, int]]:
    """x-plane slabs [lo, hi) of about equal WORK instead of equal width.  `plane_work[i]` estimates the voxel steps written into
    x-plane i (see :func:`estimate_plane_work`).  With the camera inside the grid the planes around it take most of the steps --
    every ray starts there -- so equal-width slabs leave the camera's rank as the bottleneck (measured at 1024^3 / 8 slabs: per-rank
    step times 1.8 ... 5.8 ms, mean 4.3).  Every slab keeps at least one plane; the slabs partition [0, n)."""
    w = np.asarray(plane_work, np.float64)
    n = len(w)
    if world < 1 or n < world:
        raise ValueError(f"cannot cut {n} planes into {world} slabs")
    total = float(w.sum())
    if not (total > 0) or not np.all(np.isfinite(w)):
        return [slab_range(n, r, world) for r in range(world)]
    c = np.concatenate([[0.0], np.cumsum(np.maximum(w, 0.0))])
    bounds = [0]
    for r in range(1, world):
        b = int(np.searchsorted(c, total * r / world, side="left"))
        b = min(max(b, bounds[-1] + 1), n - (world - r))      # at least one plane for this slab and for every later one
        bounds.append(b)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def refine_slabs_by_time(slabs, plane_work, seconds) -> list[tuple[int, int]]:
    """One correction step of the slab split from MEASURED per-rank times of a pilot batch: the estimate of :func:`estimate_plane_work`
    counts the voxel steps a slab receives, not what a rank spends on ray set-up, on fast-forwarding rays to its slab, or on lanes that
    idle while the longest ray of their chunk finishes.  Each slab's measured time is spread over its planes in proportion to the
    estimated work, and the planes are re-cut into slabs of equal measured cost."""
    w = np.maximum(np.asarray(plane_work, np.float64), 0.0)
    cost = np.zeros_like(w)
    for (lo, hi), t in zip(slabs, seconds):
        seg = w[lo:hi]
        tot = float(seg.sum())
        cost[lo:hi] = float(t) * (seg / tot if tot > 0 else np.full(hi - lo, 1.0 / max(1, hi - lo)))
    return balanced_slabs(cost, len(slabs))


def estimate_plane_work(frames, poses, n: int, voxel_size: float, center, pairs=None, coarse: int = 128, threshold: float = 2.0, ctx=None) -> np.ndarray:
    """Per-x-plane work estimate for :func:`balanced_slabs`: the sample pairs are accumulated into a COARSE grid of the same extent
    (`coarse` planes; a 1080p pair costs ~0.2 ms there) and the brightness that lands in each coarse x-plane is spread over the fine planes
    it covers.  Voxel steps per plane scale with it (exactly for uniformly bright motion).  `frames` / `poses` / `pairs` as in
    accumulate_motion; every rank of a slab-sharded run sees the same frames, so every rank computes the same split."""
    from . import motion

    nc = int(min(n, max(1, coarse)))
    dev = frames.device if isinstance(frames, torch.Tensor) and frames.is_cuda else torch.device("cuda", torch.cuda.current_device())
    g = motion.VoxelGrid(nc, float(voxel_size) * n / nc, center, accum="fixed", frac_bits=8, device=dev)
    width = int(frames.shape[2])
    tab, _ = motion.make_pairs(poses, width, pairs=pairs)
    motion.accumulate_motion(frames, tab, g, threshold, ctx=ctx, want_stats=False)
    wc = g.data.sum(dim=(1, 2)).double().cpu().numpy()
    idx = (np.arange(n) * nc) // n                                 # fine plane -> the coarse plane that covers it
    return wc[idx] * (nc / n)


def split_pairs(pairs: list[tuple[int, int]], rank: int, world: int) -> list[tuple[int, int]]:
    """Frame-pair shard of this rank.  Pairs are independent (each needs only its two frames), so a contiguous split
    keeps every rank's frame reads contiguous with one frame of overlap at the shard edge."""
    lo, hi = shard_range(len(pairs), rank, world)
    return pairs[lo:hi]


def reduce_grid(cells: torch.Tensor, dst: int = 0, group=None, all_ranks: bool = False) -> torch.Tensor:
    """The run's single grid merge through the process group's collective (NCCL on GPUs, gloo in CPU tests): in-place sum
    over ranks (to `dst`, or to everyone with all_ranks).  Grids made by :func:`symmetric_grid` use
    :func:`reduce_grid_peers` instead (our own kernel over NVLink peer memory / NVSwitch multicast)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return cells
    if all_ranks:
        dist.all_reduce(cells, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(cells, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return cells


def symmetric_grid(n: int, voxel_size: float, center, accum: str = "f32", frac_bits: int = 16, device=None, group=None):
    """A private full grid allocated in NVLink symmetric memory (torch.distributed._symmetric_memory: every rank can
    address every rank's copy, plus the NVSwitch multicast address), so that the run's grid merge can be done by
    pvp_grid_reduce_peers instead of a library collective.  Collective: every rank of the group must call it."""
    import torch.distributed._symmetric_memory as symm_mem
    from . import motion

    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    dtype = torch.float32 if accum == "f32" else torch.int64
    cells = symm_mem.empty((n, n, n), dtype=dtype, device=dev)
    cells.zero_()
    handle = symm_mem.rendezvous(cells, group if group is not None else dist.group.WORLD)
    grid = motion.VoxelGrid(n, voxel_size, center, accum=accum, frac_bits=frac_bits, device=dev, data=cells)
    grid.symm = handle
    return grid


def reduce_grid_peers(grid, dst: int | None = 0, ctx=None, multicast: bool | None = None) -> None:
    """The run's single grid merge with our own kernel over NVLink peer memory (SURVEY 8e mode 1): every rank sums 1/world
    of the cells over all ranks' grids -- one multimem.ld_reduce per 16 bytes when the NVSwitch multicast address exists --
    and stores them into rank `dst`'s grid (None: into every rank's).  Stream-ordered; the two symmetric-memory barriers
    keep every rank's accumulate kernels before, and every later use of the grids after, all ranks' reduce kernels.
    multicast=None picks by world size: plain peer loads for 2 GPUs (a multicast load would also pull the local copy through
    the switch: 1.19 ms vs 0.85 ms for 512 MiB), the switch reduction from 4 GPUs up (8 GPUs: 1.44 ms vs 1.52 ms to one rank,
    1.22 ms to all ranks; NCCL reduce / all_reduce on the same box: 1.05 / 1.34 ms -- tools/reduce_bench.py)."""
    import ctypes as C

    from . import motion
    from ._lib import check

    h = grid.symm
    world, rank = h.world_size, h.rank
    if world == 1:
        return
    ctx = ctx or motion.default_context(grid.data.device)
    ctx.use_current_stream()
    ptrs = (C.c_void_p * world)(*[int(p) for p in h.buffer_ptrs])
    if multicast is None:
        multicast = world > 2
    mc = int(h.multicast_ptr) if (multicast and h.multicast_ptr) else None
    h.barrier(channel=0)
    check(ctx._lib.pvp_grid_reduce_peers(ctx.handle, ptrs, world, rank, -1 if dst is None else int(dst), mc, grid.data.numel(),
                                         0 if grid.accum == "f32" else 1))
    h.barrier(channel=1)


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


def pick_mode(n: int, accum: str = "f32", world: int = 1, hbm_bytes: int | None = None) -> str:
    """mode="auto": frame sharding whenever a private full grid fits a GPU, slab sharding only when it cannot.
    Frames scale with the number of GPUs (every rank walks only its own frames, one merge of N^3 cells at the end: 512 MiB in
    ~1.3 ms, 4 GiB in ~12 ms over NVSwitch); slabs replicate the walk up to each slab (measured at 1024^3 on 8 GPUs: 2.7x one
    GPU, against ~7.9x for frames), so they are the fallback for grids beyond one GPU's memory.  A private grid needs
    n^3 cells (4 B float32, 8 B int64 fixed point) plus the ray queue and frame rings (< 2 GiB); half of the HBM is left to the caller."""
    if world <= 1:
        return "frames"
    if hbm_bytes is None:
        hbm_bytes = torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory if torch.cuda.is_available() else 180 << 30
    need = n ** 3 * (4 if accum == "f32" else 8) + (2 << 30)
    return "frames" if need <= hbm_bytes // 2 else "slabs"


def ray_voxel_distributed(metadata_json: str, image_folder: str, output_bin: str | None, mode: str = "auto", n: int = 500,
                          voxel_size: float = 6.0, center=(0.0, 0.0, 500.0), threshold: float = 2.0, accum: str = "f32",
                          frac_bits: int = 16, group=None, peer_reduce: bool = True, balance_slabs: bool = True) -> dict:
    """The ray_voxel pipeline (ray_voxel.cpp:400-558) over all ranks of the process group.  Returns this rank's counters
    summed over ranks (n_steps, n_rays ...).  Rank 0 (frames) / every rank (slabs) writes `output_bin`.
    mode: "frames", "slabs" or "auto" (see :func:`pick_mode`; the returned dict says which ran).  balance_slabs: cut the slabs by
    estimated work (a pilot over up to 8 frame pairs of the sequence into a coarse grid, the same on every rank) instead of equal width."""
    from . import motion

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    if mode == "auto":
        mode = pick_mode(n, accum, world)
    if mode == "frames":
        grid = None
        if world > 1 and peer_reduce:
            try:
                grid = symmetric_grid(n, voxel_size, center, accum=accum, frac_bits=frac_bits, device=dev, group=group)
            except Exception:  # no symmetric memory on this system: library collective
                grid = None
        if grid is None:
            grid = motion.VoxelGrid(n, voxel_size, center, accum=accum, frac_bits=frac_bits, device=dev)
        rep = motion.ray_voxel_accumulate(metadata_json, image_folder, grid, threshold, shard=(rank, world))
        if getattr(grid, "symm", None) is not None:
            # measured (tools/reduce_bench.py, 512 MiB): to every rank through multicast ld_reduce + multicast store is faster
            # than to one rank from 4 GPUs up (8 GPUs: 1.22 ms vs 1.39 ms); with 2 GPUs plain peer loads into rank 0 win
            reduce_grid_peers(grid, dst=0 if world == 2 else None)
        else:
            reduce_grid(grid.data, dst=0, group=group)
        if rank == 0 and output_bin:
            grid.save(output_bin)
    elif mode == "slabs":
        slab = slab_range(n, rank, world)
        if balance_slabs and world > 1:
            try:
                slab = balanced_slabs(_pilot_plane_work(metadata_json, image_folder, n, voxel_size, center, threshold), world)[rank]
            except Exception:      # no usable sample (e.g. every sampled file missing): equal widths
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
        raise ValueError("mode must be 'frames', 'slabs' or 'auto'")
    keys = ["n_rays", "n_steps", "n_written", "n_atomics", "n_pairs", "n_frames_loaded"]
    t = torch.tensor([rep[k] for k in keys], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, group=group)
    out = dict(zip(keys, (int(x) for x in t.tolist())))
    if mode == "slabs":  # every rank saw every ray (replicated counts); it walks only the rays that can reach its slab, and the
        for k in ("n_rays", "n_pairs", "n_frames_loaded"):   # writes partition the run's voxel steps
            out[k] //= world
        out["n_steps"] = out["n_written"]
    out["mode"] = mode
    return out


# ---- path B (process_image_cpp) over several GPUs: SURVEY.md 8e, "B path" -------------------------------------------
# Every pixel's ray is independent and the only shared state is additive (voxel grid, celestial-sphere texture:
# process_image.cpp:309-313, :356-357), so a rank processes a block of image ROWS (one big image) or a block of the
# IMAGES of the per-file loop (spacevoxelviewer.py:300-310) into PRIVATE grid + texture tensors, and the run ends with
# one sum of both.  int64 fixed-point tensors merge bit-exactly in any order; float64 differs by summation order only.
# Background subtraction reads the texture while it is being accumulated (:296-307): across ranks that read would see a
# partial sum, so it is refused for world > 1 (the reference's only call site passes False, spacevoxelviewer.py:250).

def _pilot_plane_work(metadata_json: str, image_folder: str, n: int, voxel_size: float, center, threshold: float, max_pairs: int = 8) -> np.ndarray:
    """estimate_plane_work on up to `max_pairs` consecutive-frame pairs spread over the first camera's frames (deterministic: every
    rank reads the same files and gets the same estimate)."""
    from . import motion

    poses = motion.load_metadata(metadata_json)
    cam0 = min(p.camera_index for p in poses)
    seq = sorted((p for p in poses if p.camera_index == cam0), key=lambda p: p.frame_index)
    starts = sorted({int(k) for k in np.linspace(0, max(0, len(seq) - 2), num=min(max_pairs, max(1, len(seq) - 1)))})
    frames, used, pairs = [], [], []
    for k in starts:
        try:
            a = motion.load_image_gray(os.path.join(image_folder, seq[k].image_file))
            b = motion.load_image_gray(os.path.join(image_folder, seq[k + 1].image_file))
        except Exception:
            continue
        if a.shape != b.shape or (frames and a.shape != frames[0].shape):
            continue
        pairs.append((len(frames), len(frames) + 1))
        frames += [a, b]
        used += [seq[k], seq[k + 1]]
    if not pairs:
        raise ValueError("no loadable sample pair")
    dev = torch.device("cuda", torch.cuda.current_device())
    return estimate_plane_work(torch.as_tensor(np.stack(frames), device=dev), used, n, voxel_size, center, pairs=pairs, threshold=threshold)


def _world_rank(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _check_shardable(perform_background_subtraction, world: int) -> None:
    if perform_background_subtraction and world > 1:
        raise ValueError("perform_background_subtraction=True reads the texture being accumulated (process_image.cpp:296-307): "
                         "run it on one GPU (replicas only), not sharded")


def process_image_sharded(image, earth_position, pointing_direction, fov, image_width, image_height, voxel_grid, voxel_grid_extent,
                          max_distance, num_steps, celestial_sphere_texture, center_ra_rad, center_dec_rad, angular_width_rad,
                          angular_height_rad, update_celestial_sphere, perform_background_subtraction, *, group=None,
                          rank: int | None = None, world: int | None = None, **kw) -> tuple[int, int]:
    """process_image_cpp (same 17 arguments) for ONE image shared by all ranks: this rank casts the rays of its block of
    image rows into its private `voxel_grid` / `celestial_sphere_texture`.  Returns the row block.  Follow the last image
    with :func:`reduce_sky`.  rank / world override the process group (tests run the ranks one after another)."""
    from .process_image import process_image_cpp

    w, r = _world_rank(group)
    world, rank = (w if world is None else world), (r if rank is None else rank)
    _check_shardable(perform_background_subtraction, world)
    rows = shard_range(int(image_height), rank, world)
    process_image_cpp(image, earth_position, pointing_direction, fov, image_width, image_height, voxel_grid, voxel_grid_extent,
                      max_distance, num_steps, celestial_sphere_texture, center_ra_rad, center_dec_rad, angular_width_rad,
                      angular_height_rad, update_celestial_sphere, perform_background_subtraction, row_range=rows, **kw)
    return rows


def process_images_sharded(observations, image_width, image_height, voxel_grid, voxel_grid_extent, max_distance, num_steps,
                           celestial_sphere_texture, center_ra_rad, center_dec_rad, angular_width_rad, angular_height_rad,
                           update_celestial_sphere=True, perform_background_subtraction=False, *, group=None,
                           rank: int | None = None, world: int | None = None, **kw) -> tuple[int, int]:
    """The per-file loop of spacevoxelviewer.py:300-310 with the FILES sharded: `observations` is the ordered list of
    (image, earth_position, pointing_direction, fov) records (or a callable index -> record, so a rank only loads its own
    files, with `n_observations=` in kw); this rank processes its contiguous block into its private tensors.  Returns the
    block.  Follow with :func:`reduce_sky`."""
    from .process_image import process_image_cpp

    w, r = _world_rank(group)
    world, rank = (w if world is None else world), (r if rank is None else rank)
    _check_shardable(perform_background_subtraction, world)
    n_obs = kw.pop("n_observations", None)
    n_obs = len(observations) if n_obs is None else int(n_obs)
    lo, hi = shard_range(n_obs, rank, world)
    for k in range(lo, hi):
        image, earth_position, pointing_direction, fov = observations(k) if callable(observations) else observations[k]
        process_image_cpp(image, earth_position, pointing_direction, fov, image_width, image_height, voxel_grid, voxel_grid_extent,
                          max_distance, num_steps, celestial_sphere_texture, center_ra_rad, center_dec_rad, angular_width_rad,
                          angular_height_rad, update_celestial_sphere, perform_background_subtraction, **kw)
    return lo, hi


def reduce_sky(voxel_grid: torch.Tensor | None, celestial_sphere_texture: torch.Tensor | None, dst: int | None = 0, group=None) -> None:
    """The single merge that ends a sharded path-B run: in-place sum over ranks of the private voxel grids and textures
    (to rank `dst`, or to every rank with dst=None).  Tensors must be contiguous (strided views would be copied by the
    collective and the sum lost)."""
    for t in (voxel_grid, celestial_sphere_texture):
        if t is None or t.numel() == 0:
            continue
        if not t.is_contiguous():
            raise ValueError("reduce_sky needs contiguous tensors")
        reduce_grid(t, dst=0 if dst is None else dst, group=group, all_ranks=dst is None)
