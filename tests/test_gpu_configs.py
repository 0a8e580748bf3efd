"""BASELINE.json's multi-GPU configurations at their full shapes, on one GPU with the ranks run one after another
(the collective itself is covered by tests/test_dist_gloo.py on CPU and tools/dist_check.py under torchrun):

* config 3: 4 cameras x 3840x2160 frames into a 512^3 grid, frame pairs sharded over 2 / 4 / 8 ranks, private grids,
  one sum-reduce.  Property: the sum of the shard grids equals the whole run (fixed point: bit for bit; fp32 with
  uint8 frames: bit for bit too, every addend is an integer and the sums stay below 2^24).
* config 4: a 1024^3 grid (4 GiB fp32) sharded into 8 x-slabs.  Property: a slab walks the rays that can reach it with full-grid
  arithmetic (from their start, until x has passed the slab), writes partition the voxel steps, and the concatenated slabs equal the whole grid.

No CPU oracle at these sizes; the small-size bit-exact parity against the oracle is in test_gpu_motion.py."""
import numpy as np
import pytest
import torch

import pixeltovoxelprojector_b200 as pvp
from pixeltovoxelprojector_b200 import synth
from pixeltovoxelprojector_b200.distributed import shard_range, slab_range

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _four_camera_scene(frames_per_cam, H, W, N, vs, center):
    frames, poses, pairs = [], [], []
    for cam in range(4):
        f = synth.sparse_frames(frames_per_cam, H, W, seed=40 + cam, blob=160, blobs=5)
        p = synth.orbit_poses(frames_per_cam, N, vs, center, seed=50 + cam, camera_index=cam)
        base = len(poses)
        frames.append(f)
        poses += p
        pairs += [(base + i - 1, base + i) for i in range(1, frames_per_cam)]   # consecutive frames of ONE camera
    return np.concatenate(frames), poses, pairs


def test_config3_four_4k_cameras_frame_sharded_512(ctx):
    H, W, N, vs, center = 2160, 3840, 512, 2.0, (0.0, 0.0, 500.0)
    frames_np, poses, pairs = _four_camera_scene(3, H, W, N, vs, center)
    frames = torch.as_tensor(frames_np, device=DEV)
    whole_fx = pvp.VoxelGrid(N, vs, center, accum="fixed", frac_bits=8)
    tab, n = pvp.make_pairs(poses, W, pairs=pairs)
    st = pvp.accumulate_motion(frames, tab, whole_fx, 2.0)
    assert n == 8 and st["n_rays"] > 20000 and st["n_steps"] > 100 * st["n_rays"]   # moving blobs: only their edges change
    whole_f32 = pvp.VoxelGrid(N, vs, center)
    pvp.accumulate_motion(frames, tab, whole_f32, 2.0)
    assert float(whole_f32.data.max()) < 2 ** 24
    assert torch.equal(whole_f32.data, whole_fx.to_float32())           # exact integer sums in both modes
    for world in (2, 4, 8):
        total_fx = torch.zeros_like(whole_fx.data)
        total_f32 = torch.zeros_like(whole_f32.data)
        steps = 0
        for rank in range(world):
            lo, hi = shard_range(len(pairs), rank, world)
            if lo == hi:
                continue
            t, _ = pvp.make_pairs(poses, W, pairs=pairs[lo:hi])
            g = pvp.VoxelGrid(N, vs, center, accum="fixed", frac_bits=8)
            steps += pvp.accumulate_motion(frames, t, g, 2.0)["n_steps"]
            total_fx += g.data                                            # the rank-0 side of the sum-reduce
            g32 = pvp.VoxelGrid(N, vs, center)
            pvp.accumulate_motion(frames, t, g32, 2.0)
            total_f32 += g32.data
            del g, g32
        assert steps == st["n_steps"]
        assert torch.equal(total_fx, whole_fx.data), world
        assert torch.equal(total_f32, whole_f32.data), world


def test_config4_1024_grid_eight_x_slabs(ctx):
    H, W, N, vs, center = 1080, 1920, 1024, 1.0, (0.0, 0.0, 500.0)
    frames = torch.as_tensor(synth.sparse_frames(3, H, W, seed=60, blob=120, blobs=4), device=DEV)
    poses = synth.orbit_poses(3, N, vs, center, seed=61)
    tab, _ = pvp.make_pairs(poses, W)
    whole = pvp.VoxelGrid(N, vs, center)                                   # 4 GiB
    st = pvp.accumulate_motion(frames, tab, whole, 2.0)
    assert st["n_steps"] > 200 * st["n_rays"] > 0 and st["n_written"] == st["n_steps"]
    written, walked, world = 0, 0, 8
    for rank in range(world):
        lo, hi = slab_range(N, rank, world)
        assert hi - lo == 128
        slab = pvp.VoxelGrid(N, vs, center, slab=(lo, hi))                 # 512 MiB: x-planes [lo, hi) only
        s = pvp.accumulate_motion(frames, tab, slab, 2.0)
        assert s["n_rays"] == st["n_rays"] and s["n_written"] <= s["n_steps"] < st["n_steps"]   # clipped: rays that cannot reach the slab are not walked
        walked += s["n_steps"]
        written += s["n_written"]
        assert torch.equal(slab.data, whole.data[lo:hi]), rank             # a contiguous byte range of the .bin
        del slab
    assert written == st["n_steps"]
    assert walked < 4 * st["n_steps"]                 # 8 ranks walking every ray in full would be 8x
    # the fixed-point slab path (64-bit cells) on one slab
    lo, hi = slab_range(N, 3, world)
    fx = pvp.VoxelGrid(N, vs, center, accum="fixed", frac_bits=6, slab=(lo, hi))
    pvp.accumulate_motion(frames, tab, fx, 2.0)
    assert torch.equal(fx.to_float32(), whole.data[lo:hi])


def test_headline_shape_all_variants_agree_1080p_dense_512(ctx):
    """The bench workload's shape (1080p dense motion -> 512^3, 2 M rays and 1.2 G voxel steps per pair): every K2 variant --
    the default auto pick (lean2 for this batch of two pairs), lean, lean2, lean4, the first scan kernel and the plain one-RED-per-step kernel, whose code shares nothing
    with the lock-step kernels beyond the ray set-up -- must produce the same int64 grid and the same counters, and the float32
    grid must equal the fixed-point one exactly wherever the integer sum is below 2^24."""
    H, W, N, vs, center = 1080, 1920, 512, 2.0, (0.0, 0.0, 500.0)
    frames = torch.as_tensor(synth.dense_frames(3, H, W, seed=77), device=DEV)
    poses = synth.orbit_poses(3, N, vs, center, seed=78)
    tab, _ = pvp.make_pairs(poses, W)
    ref_grid, ref_stats = None, None
    try:
        ctx.set_option("count_atomics", 1)     # lean4 leaves its RED counter out unless asked
        for variant in (3, 0, 4, 5, 6, 1):
            ctx.set_option("dda_variant", variant)
            g = pvp.VoxelGrid(N, vs, center, accum="fixed", frac_bits=4)
            st = pvp.accumulate_motion(frames, tab, g, 2.0)
            key = (st["n_rays"], st["n_steps"], st["n_written"])
            if ref_grid is None:
                ref_grid, ref_stats = g.data, key
                assert st["n_rays"] > 3_900_000 and st["n_steps"] > 2_000_000_000 and st["n_atomics"] == st["n_written"]
            else:
                assert key == ref_stats, variant
                assert torch.equal(g.data, ref_grid), variant
                assert st["n_atomics"] < 0.3 * st["n_written"]          # aggregated
                del g
        exact = pvp.VoxelGrid(N, vs, center, accum="fixed", frac_bits=4, data=ref_grid).to_float32()
        small = exact < 2 ** 24                     # integer sums below 2^24 are exact in fp32 whatever the order
        for variant, want_pick in ((0, 2), (6, 4), (4, 1)):     # float32 cells: the auto pick (4 M rays: lean2), lean4 and lean
            ctx.set_option("dda_variant", variant)
            gf = pvp.VoxelGrid(N, vs, center)
            st = pvp.accumulate_motion(frames, tab, gf, 2.0)
            assert ctx.last_k2_pick() == want_pick, variant
            assert (st["n_rays"], st["n_steps"], st["n_written"]) == ref_stats and 0 < st["n_atomics"] < 0.3 * st["n_written"], variant
            assert torch.equal(gf.data[small], exact[small]) and int((~small).sum()) > 0, variant
            # the few cells every ray crosses (around the camera: up to 1.8e8 here) round at each of ~10^5 atomic adds
            torch.testing.assert_close(gf.data[~small], exact[~small], rtol=1e-4, atol=0)
            del gf
    finally:
        ctx.set_option("dda_variant", 0)
        ctx.set_option("count_atomics", 0)


def test_grid_near_the_32bit_offset_limit_variants_agree(ctx):
    """N = 1500: 3.4e9 cells, voxel offsets just below the 2^32 key space of the lock-step kernels (N >= 1626 falls back to the
    plain kernel), x-planes of 9 MB (auto mode: one-ray kernel on 8-row tiles).  The auto pick, lean2 and the plain one-RED-per-step
    kernel must agree bit for bit in fixed point, counters included."""
    H, W, N, vs, center = 480, 640, 1500, 0.7, (0.0, 0.0, 500.0)
    frames = torch.as_tensor(synth.dense_frames(2, H, W, seed=91), device=DEV)
    poses = synth.orbit_poses(2, N, vs, center, seed=92)
    tab, _ = pvp.make_pairs(poses, W)
    ref, ref_key = None, None
    try:
        for variant in (3, 0, 5):
            ctx.set_option("dda_variant", variant)
            g = pvp.VoxelGrid(N, vs, center, accum="fixed", frac_bits=4)          # 27 GB
            st = pvp.accumulate_motion(frames, tab, g, 2.0)
            key = (st["n_rays"], st["n_steps"], st["n_written"])
            if ref is None:
                ref, ref_key = g, key
                assert st["n_rays"] > 290_000 and st["n_steps"] > 500 * st["n_rays"]
                assert int(g.data.sum()) > 0
            else:
                assert key == ref_key, variant
                assert torch.equal(g.data, ref.data), variant
                del g
    finally:
        ctx.set_option("dda_variant", 0)
        del ref
        torch.cuda.empty_cache()


def test_slab_of_a_grid_beyond_one_gpu_vs_oracle(ctx, oracle):
    """A grid too large for one GPU is what slab sharding exists for: N = 3000 float32 is 108 GB of cells.  One rank's slab (x-planes
    [1496, 1512): 576 MB) of that grid against the C port walking the same rays with full-grid arithmetic and writing into the same
    slab -- counters and every cell.  (N > 2048 runs the reference-structured kernel: the lock-step kernels' per-ray limits are proven
    for N <= 2048.)"""
    H, W, N, vs, center = 60, 80, 3000, 0.34, (0.0, 0.0, 500.0)
    frames = synth.dense_frames(2, H, W, seed=301)
    poses = synth.orbit_poses(2, N, vs, center, seed=302)
    lo, hi = 1496, 1512
    want = np.zeros((hi - lo, N, N), np.float32)
    R = oracle.rotation_matrix(poses[1].yaw, poses[1].pitch, poses[1].roll)
    nr, ns, nw = oracle.accumulate_pair(frames[0], frames[1], np.array(poses[1].camera_position, np.float32), R, oracle.focal_length(poses[1].fov_degrees, W),
                                        want, N, vs, np.array(center, np.float32), 2.0, slab=(lo, hi))
    tab, _ = pvp.make_pairs(poses, W)
    slab = pvp.VoxelGrid(N, vs, center, slab=(lo, hi))
    st = pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), tab, slab, 2.0)
    assert (st["n_rays"], st["n_written"]) == (nr, nw) and nw > 1000 and ns > 1000 * nr > 0
    got = slab.data.cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    with pytest.raises(pvp.PvpError):
        pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), tab, pvp.VoxelGrid(4097, 1.0, center, slab=(0, 0)), 2.0)   # n <= 4096


def test_slab_of_a_2048_grid_runs_on_the_lockstep_kernels_vs_oracle(ctx, oracle):
    """2048^3 float32 is 32 GiB: the size slab sharding is for.  Since the exact fast-forward a walked ray never leaves its slab, so the
    32-bit voxel keys only have to hold the slab's own cells and the lock-step kernels (lean by default at this plane size, lean2 and
    lean4 when forced) take grids up to n = 2048 under slabs; before, every grid above 1290^3 fell back to the one-RED-per-step kernel.
    One rank's slab (x-planes [1016, 1032): 256 MiB) against the C port walking the same rays with full-grid arithmetic -- counters and
    every cell, for every kernel, float32 and fixed point; and a slab at the grid's low end whose rays start inside it."""
    H, W, N, vs, center = 60, 80, 2048, 0.5, (0.0, 0.0, 500.0)
    frames = synth.dense_frames(2, H, W, seed=311)
    poses = synth.orbit_poses(2, N, vs, center, seed=312)
    R = oracle.rotation_matrix(poses[1].yaw, poses[1].pitch, poses[1].roll)
    tab, _ = pvp.make_pairs(poses, W)
    fr = torch.as_tensor(frames, device=DEV)
    cam_plane = int((poses[1].camera_position[0] - (center[0] - N * vs / 2)) / vs)
    for lo, hi in ((1016, 1032), (max(0, cam_plane - 8), max(0, cam_plane - 8) + 16)):
        want = np.zeros((hi - lo, N, N), np.float32)
        nr, ns, nw = oracle.accumulate_pair(frames[0], frames[1], np.array(poses[1].camera_position, np.float32), R, oracle.focal_length(poses[1].fov_degrees, W),
                                            want, N, vs, np.array(center, np.float32), 2.0, slab=(lo, hi))
        assert nw > 1000 and ns > 500 * nr > 0
        try:
            for variant, pick in ((0, 1), (4, 1), (5, 2), (6, 4), (3, 0)):
                ctx.set_option("dda_variant", variant)
                slab = pvp.VoxelGrid(N, vs, center, slab=(lo, hi))
                st = pvp.accumulate_motion(fr, tab, slab, 2.0)
                assert ctx.last_k2_pick() == pick, (variant, lo)
                assert (st["n_rays"], st["n_written"]) == (nr, nw), (variant, lo, st)
                assert np.array_equal(slab.data.cpu().numpy().view(np.uint32), want.view(np.uint32)), (variant, lo)
                del slab
        finally:
            ctx.set_option("dda_variant", 0)
