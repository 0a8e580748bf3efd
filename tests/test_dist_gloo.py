"""The N>1 host logic on CPU: world_size-2 gloo process group (127.0.0.1).  Covers the shard arithmetic, the
single grid sum-reduce (int64 fixed point: bit-exact; fp32), and slab-wise .bin assembly.  The kernels
themselves need a GPU and are covered by `-m gpu` (test_python_pipeline_sharded_equals_whole)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT  # noqa: F401  (puts the repo on sys.path)

from pixeltovoxelprojector_b200 import distributed as pd
from pixeltovoxelprojector_b200.motion import read_bin, write_bin


def test_shard_ranges_partition():
    for total in (0, 1, 7, 500, 1001):
        for world in (1, 2, 3, 8):
            blocks = [pd.shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert max(h - l for l, h in blocks) - min(h - l for l, h in blocks) <= 1
    pairs = [(i, i + 1) for i in range(11)]
    got = sum((pd.split_pairs(pairs, r, 4) for r in range(4)), [])
    assert got == pairs
    with pytest.raises(ValueError):
        pd.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 12
        rng = np.random.default_rng(100 + rank)
        # frame sharding: private grids, one sum-reduce to rank 0
        fixed = torch.from_numpy(rng.integers(0, 1 << 40, (n, n, n), dtype=np.int64))
        f32 = torch.from_numpy(rng.integers(0, 255, (n, n, n)).astype(np.float32))
        mine_fixed, mine_f32 = fixed.clone(), f32.clone()
        pd.reduce_grid(fixed, dst=0)
        pd.reduce_grid(f32, dst=0)
        everyone = mine_fixed.clone()
        pd.reduce_grid(everyone, all_ranks=True)
        parts = [torch.zeros_like(mine_fixed) for _ in range(world)]
        dist.all_gather(parts, mine_fixed)
        want = sum(parts)
        assert torch.equal(everyone, want)
        if rank == 0:
            assert torch.equal(fixed, want)
            parts32 = [torch.from_numpy(np.random.default_rng(100 + r).integers(0, 1 << 40, (n, n, n), dtype=np.int64)) for r in range(world)]
            assert torch.equal(want, sum(parts32))
            assert torch.equal(f32, sum(torch.from_numpy(_f32_of(r, n)) for r in range(world)))
        # slab sharding: every rank writes its byte range of the shared .bin
        whole = np.random.default_rng(7).random((n, n, n)).astype(np.float32)
        lo, hi = pd.slab_range(n, rank, world)
        path = os.path.join(tmp, "slabs.bin")
        if rank == 0:
            pd.write_bin_header(path, n, 6.0)
        dist.barrier()
        pd.write_bin_slab(path, n, (lo, hi), whole[lo:hi])
        dist.barrier()
        if rank == 0:
            ref = os.path.join(tmp, "whole.bin")
            write_bin(ref, whole, 6.0)
            assert open(path, "rb").read() == open(ref, "rb").read()
            back, vs = read_bin(path)
            assert np.array_equal(back, whole) and vs == 6.0
        _path_b_host_logic(rank, world)
    finally:
        dist.destroy_process_group()


def _path_b_host_logic(rank, world):
    """SURVEY 8e "B path": file sharding + the merge of grid and texture, with the kernel call replaced by a CPU stand-in
    (the real kernel is covered on the GPU by test_gpu_process_image.py::test_row_and_file_shards_sum_to_whole)."""
    import pixeltovoxelprojector_b200.process_image as pi
    calls = []

    def fake(image, earth_position, pointing_direction, fov, w, h, grid, ext, md, ns, tex, *rest, row_range=None, **kw):
        lo, hi = row_range if row_range is not None else (0, h)
        calls.append((float(fov), lo, hi))
        grid += int(image[lo:hi].sum())
        tex += float(fov)
    real, pi.process_image_cpp = pi.process_image_cpp, fake
    try:
        H, W, n_obs = 11, 6, 5
        obs = [(torch.full((H, W), k + 1, dtype=torch.int64), [0, 0, 0], [0, 0, 1], float(k)) for k in range(n_obs)]
        grid, tex = torch.zeros((3, 3, 3), dtype=torch.int64), torch.zeros((2, 2), dtype=torch.float64)
        common = (W, H, grid, [(0, 1)] * 3, 10.0, 10, tex, 0.0, 0.0, 1.0, 1.0)
        lo, hi = pd.process_images_sharded(obs, *common)
        assert (lo, hi) == pd.shard_range(n_obs, rank, world) and [c[0] for c in calls] == [float(k) for k in range(lo, hi)]
        loaded = []
        pd.process_images_sharded(lambda k: (loaded.append(k), obs[k])[1], *common, n_observations=n_obs)   # lazy: only this rank's files
        assert loaded == list(range(lo, hi))
        rows = pd.process_image_sharded(obs[0][0], [0, 0, 0], [0, 0, 1], 0.5, *common, True, False)
        assert rows == pd.shard_range(H, rank, world) and calls[-1] == (0.5, *rows)
        pd.reduce_sky(grid, tex, dst=None)
        per_image = [(k + 1) * H * W for k in range(n_obs)]
        assert int(grid[0, 0, 0]) == 2 * sum(per_image) + H * W                 # both passes over the files + the row-sharded image
        assert float(tex[0, 0]) == 2 * sum(range(n_obs)) + 0.5 * world
        with pytest.raises(ValueError, match="background_subtraction"):
            pd.process_image_sharded(obs[0][0], [0, 0, 0], [0, 0, 1], 0.5, *common, True, True)
        with pytest.raises(ValueError, match="contiguous"):
            pd.reduce_sky(grid[:, ::2], None)
    finally:
        pi.process_image_cpp = real


def _f32_of(rank, n):
    rng = np.random.default_rng(100 + rank)
    rng.integers(0, 1 << 40, (n, n, n), dtype=np.int64)
    return rng.integers(0, 255, (n, n, n)).astype(np.float32)


def test_world2_gloo_reduce_and_slab_assembly(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)


def test_auto_mode_prefers_frames_while_a_private_grid_fits():
    """mode="auto" of ray_voxel_distributed: frames + one merge whenever a private full grid fits the GPU, slabs only beyond that."""
    hbm = 180 << 30
    assert pd.pick_mode(512, "f32", world=8, hbm_bytes=hbm) == "frames"
    assert pd.pick_mode(1024, "f32", world=8, hbm_bytes=hbm) == "frames"         # BASELINE config 4's grid: 4 GiB
    assert pd.pick_mode(2048, "fixed", world=8, hbm_bytes=hbm) == "frames"       # 64 GiB
    assert pd.pick_mode(3072, "f32", world=8, hbm_bytes=hbm) == "slabs"          # 108 GiB: more than half of one GPU
    assert pd.pick_mode(4096, "f32", world=1, hbm_bytes=hbm) == "frames"         # one rank: nothing to shard


def test_balanced_slabs_partition_and_balance():
    rng = np.random.default_rng(4)
    for n, world in ((1024, 8), (100, 7), (8, 8), (33, 2)):
        w = rng.gamma(0.5, 1.0, n) * np.exp(-((np.arange(n) - n * 0.45) / (n * 0.12)) ** 2)      # work concentrated around the camera
        slabs = pd.balanced_slabs(w, world)
        assert slabs[0][0] == 0 and slabs[-1][1] == n and all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
        assert all(hi > lo for lo, hi in slabs)
        if n == 1024:
            per = [w[lo:hi].sum() for lo, hi in slabs]
            eq = [w[lo:hi].sum() for lo, hi in (pd.slab_range(n, r, world) for r in range(world))]
            assert max(per) < 0.6 * max(eq) and max(per) < 1.3 * w.sum() / world
    assert pd.balanced_slabs(np.zeros(64), 4) == [pd.slab_range(64, r, 4) for r in range(4)]     # no information: equal widths
    # measured-time refinement: a rank that took twice as long as the others gives planes away
    w = np.ones(800)
    first = pd.balanced_slabs(w, 8)
    again = pd.refine_slabs_by_time(first, w, [1.0, 1.0, 2.0, 1.0, 1.0, 1.0, 1.0, 1.0])
    assert again[2][1] - again[2][0] < first[2][1] - first[2][0] and again[0][0] == 0 and again[-1][1] == 800
    assert abs((again[2][1] - again[2][0]) - 100 * 9 / 8 / 2) <= 2
    with pytest.raises(ValueError):
        pd.balanced_slabs(np.ones(3), 4)
