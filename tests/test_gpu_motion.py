"""Parity of path A on the GPU, through the C ABI: motion masks, ray set-up, visited voxel lists (bit-exact),
accumulated grids (bit-exact for integer-valued frames and for fixed point; rel. 1e-5 for float frames with
fp32 atomics), slabs, the host-staged path, and size-independent properties at full sizes."""
import numpy as np
import pytest
import torch

from conftest import bits, golden

import pixeltovoxelprojector_b200 as pvp
from pixeltovoxelprojector_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL_F32_ATOMICS = 1e-5   # BASELINE.json north_star: fp32 atomics within a relative tolerance of 1e-5


def grid_like(g, accum="f32", slab=None, frac_bits=16):
    return pvp.VoxelGrid(int(g["N"]), float(g["voxel_size"]), tuple(float(c) for c in g["center"]), accum=accum, slab=slab, frac_bits=frac_bits)


def pose_of(g, W):
    return pvp.FramePose(tuple(float(c) for c in g["cam"]), *(float(a) for a in g["ypr"]), fov_degrees=float(g["fov"]))


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
def test_detect_motion_bit_exact(ctx, oracle, dtype):
    rng = np.random.default_rng(0)
    for shape in ((1, 1), (7, 13), (480, 640), (1080, 1920)):
        prev = rng.integers(0, 256, shape, dtype=np.uint8)
        curr = np.where(rng.random(shape) < 0.5, prev, rng.integers(0, 256, shape, dtype=np.uint8)).astype(np.uint8)
        if dtype == torch.float32:
            prev = prev + rng.uniform(-1, 1, shape).astype(np.float32)
            curr = curr + rng.uniform(-1, 1, shape).astype(np.float32)
        want_c, want_d = oracle.detect_motion(prev.astype(np.float32), curr.astype(np.float32), 2.0)
        c, d = pvp.detect_motion(torch.as_tensor(prev, device=DEV).to(dtype), torch.as_tensor(curr, device=DEV).to(dtype), 2.0)
        assert np.array_equal(c.cpu().numpy(), want_c)
        assert np.array_equal(bits(d.cpu().numpy()), bits(want_d))
    with pytest.raises(ValueError, match="differ in size"):
        pvp.detect_motion(torch.zeros((4, 4), device=DEV, dtype=dtype), torch.zeros((4, 5), device=DEV, dtype=dtype), 2.0)


def test_pixel_rays_bit_exact(ctx, oracle):
    rng = np.random.default_rng(1)
    for W, H in ((640, 480), (1920, 1080), (3840, 2160), (4096, 4096), (17, 5)):
        pose = pvp.FramePose((0, 0, 0), float(np.float32(rng.uniform(-180, 180))), float(np.float32(rng.uniform(-90, 90))),
                             float(np.float32(rng.uniform(-180, 180))), fov_degrees=float(np.float32(rng.uniform(20, 120))))
        pix = rng.integers(0, W * H, 4000)
        pix[:4] = [0, W - 1, W * (H - 1), W * H - 1]
        got = pvp.pixel_rays(torch.as_tensor(pix, device=DEV), W, H, pose).cpu().numpy()
        R, f = oracle.rotation_matrix(pose.yaw, pose.pitch, pose.roll), oracle.focal_length(pose.fov_degrees, W)
        want = np.stack([oracle.pixel_ray(int(p % W), int(p // W), W, H, f, R) for p in pix])
        assert np.array_equal(bits(got), bits(want))


def test_cast_rays_golden_bit_exact(ctx):
    g = golden("rays_golden.npz")
    off = np.concatenate([[0], np.cumsum(g["counts"])])
    for i in range(len(g["N"])):
        grid = pvp.VoxelGrid(int(g["N"][i]), float(g["voxel_size"][i]), tuple(float(c) for c in g["center"][i]))
        counts, offsets, vox, t = pvp.cast_rays(torch.as_tensor(g["cam"][i:i + 1], device=DEV), torch.as_tensor(g["dir"][i:i + 1], device=DEV), grid)
        assert int(counts[0]) == g["counts"][i], i
        assert np.array_equal(vox.cpu().numpy(), g["vox"][off[i]:off[i + 1]]), i
        assert np.array_equal(bits(t.cpu().numpy()), bits(g["t"][off[i]:off[i + 1]])), i


@pytest.mark.parametrize("N,vs", [(64, 2.0), (256, 1.5), (500, 6.0)])
def test_cast_rays_random_bit_exact(ctx, oracle, N, vs):
    """Visited voxel indices must match the reference bit-exactly (order, count, t)."""
    rng = np.random.default_rng(N)
    n = 600
    center = np.array([0, 0, 500], np.float32)
    half = N * vs / 2
    cam = (center + rng.uniform(-1.5, 1.5, (n, 3)) * half).astype(np.float32)
    tgt = center + rng.uniform(-1, 1, (n, 3)) * half
    d = (tgt - cam).astype(np.float32)
    d[::9, rng.integers(0, 3)] = 0.0
    d[::11] = rng.choice([-1.0, 1.0, 0.0, -0.0], (len(d[::11]), 3))
    cam[::11] = (center - half + vs * rng.integers(0, N + 1, (len(cam[::11]), 3))).astype(np.float32)  # on lattice planes: ties
    nrm = np.linalg.norm(d, axis=1, keepdims=True)
    d = np.where(nrm > 0, d / np.maximum(nrm, 1e-30), d).astype(np.float32)
    grid = pvp.VoxelGrid(N, vs, tuple(center))
    counts, offsets, vox, t = pvp.cast_rays(torch.as_tensor(cam, device=DEV), torch.as_tensor(d, device=DEV), grid)
    counts, offsets, vox, t = counts.cpu().numpy(), offsets.cpu().numpy(), vox.cpu().numpy(), t.cpu().numpy()
    total = 0
    for i in range(n):
        w_vox, w_t = oracle.cast_ray(cam[i], d[i], N, vs, center, want_t=True)
        assert counts[i] == len(w_vox), i
        assert np.array_equal(vox[offsets[i]:offsets[i + 1]], w_vox), i
        assert np.array_equal(bits(t[offsets[i]:offsets[i + 1]]), bits(w_t)), i
        total += len(w_vox)
    assert total > 20 * n // 4


@pytest.mark.parametrize("accum", ["f32", "fixed"])
@pytest.mark.parametrize("frames_on", ["device", "host"])
@pytest.mark.parametrize("dtype", ["u8", "f32"])
def test_pair_golden(ctx, accum, frames_on, dtype):
    """ray_voxel.cpp:484-531 on the committed reference output: integer-valued frames => exact sums."""
    g = golden("pair_golden.npz")
    H, W = g["prev"].shape
    frames = np.stack([g["prev"], g["curr"]])
    frames = frames if dtype == "u8" else frames.astype(np.float32)
    pairs, n = pvp.make_pairs([pose_of(g, W)] * 2, W)
    assert np.array_equal(bits(np.array(pairs[0].R[:], np.float32)), bits(g["R"])) and bits(np.float32(pairs[0].focal)) == bits(g["focal"])
    grid = grid_like(g, accum)
    fr = torch.as_tensor(frames, device=DEV) if frames_on == "device" else frames
    st = pvp.accumulate_motion(fr, pairs, grid, float(g["threshold"]))
    assert (st["n_rays"], st["n_steps"], st["n_written"]) == (int(g["n_rays"]), int(g["n_steps"]), int(g["n_steps"]))
    out = grid.to_float32().cpu().numpy().reshape(-1)
    nz = np.flatnonzero(out)
    assert np.array_equal(nz, g["nz_index"])
    assert np.array_equal(bits(out[nz]), bits(g["nz_value"]))


def test_slabs_partition_the_grid(ctx):
    g = golden("pair_golden.npz")
    H, W = g["prev"].shape
    frames = torch.as_tensor(np.stack([g["prev"], g["curr"]]), device=DEV)
    pairs, _ = pvp.make_pairs([pose_of(g, W)] * 2, W)
    N = int(g["N"])
    parts, written = [], 0
    for lo, hi in ((0, 9), (9, 10), (10, 31), (31, N), (N, N)):
        grid = grid_like(g, slab=(lo, hi))
        st = pvp.accumulate_motion(frames, pairs, grid, float(g["threshold"]))
        assert st["n_written"] <= st["n_steps"] <= int(g["n_steps"])   # a slab walks the rays that can reach it, until they have passed it
        written += st["n_written"]
        parts.append(grid.data)
    assert written == int(g["n_steps"])
    out = torch.cat(parts).cpu().numpy().reshape(-1)
    assert np.array_equal(np.flatnonzero(out), g["nz_index"]) and np.array_equal(out[g["nz_index"]], g["nz_value"])


def _scene(rng, F, H, W, N, vs, center, dense):
    frames = synth.dense_frames(F, H, W, seed=int(rng.integers(1 << 30))) if dense else synth.sparse_frames(F, H, W, seed=int(rng.integers(1 << 30)), blobs=3)
    poses = synth.orbit_poses(F, N, vs, center, seed=int(rng.integers(1 << 30)))
    return frames, poses


def _oracle_run(oracle, frames, poses, N, vs, center, thr, accum_mode=0, frac_bits=16, pair_idx=None):
    grid = np.zeros((N, N, N), np.float32 if accum_mode == 0 else np.int64)
    tot = np.zeros(3, np.int64)
    W = frames.shape[2]
    for (p, c) in (pair_idx or [(i - 1, i) for i in range(1, len(frames))]):
        R = oracle.rotation_matrix(poses[c].yaw, poses[c].pitch, poses[c].roll)
        f = oracle.focal_length(poses[c].fov_degrees, W)
        tot += oracle.accumulate_pair(frames[p], frames[c], np.array(poses[c].camera_position, np.float32), R, f, grid, N, vs,
                                      np.array(center, np.float32), thr, accum_mode=accum_mode, frac_bits=frac_bits)
    return grid, tot


@pytest.mark.parametrize("dense", [False, True])
def test_sequence_u8_bit_exact(ctx, oracle, dense):
    """Multi-pair batch, moving camera, uint8 frames: every addend is an integer <= 255, so fp32 sums are exact and
    order independent (< 2^24) -- the grid must equal the reference's serial sum bit for bit."""
    rng = np.random.default_rng(10 + dense)
    F, H, W, N, vs, center = 5, 60, 96, 96, 4.0, (0.0, 0.0, 500.0)
    frames, poses = _scene(rng, F, H, W, N, vs, center, dense)
    want, tot = _oracle_run(oracle, frames, poses, N, vs, center, 2.0)
    assert want.max() < 2 ** 24 and tot[1] > 1000
    pairs, n = pvp.make_pairs(poses, W)
    for fr in (torch.as_tensor(frames, device=DEV), frames, torch.as_tensor(frames).pin_memory()):
        grid = pvp.VoxelGrid(N, vs, center)
        st = pvp.accumulate_motion(fr, pairs, grid, 2.0)
        assert (st["n_rays"], st["n_steps"], st["n_written"]) == tuple(tot)
        assert np.array_equal(bits(grid.data.cpu().numpy()), bits(want))


def test_sequence_float_frames_tolerance_and_fixed_exact(ctx, oracle):
    """Float frames (loader-style noise, seeded): fp32 atomics within rel 1e-5 of the serial reference sum;
    fixed-point mode bit-exact against the oracle's fixed-point accumulation and run-to-run deterministic."""
    rng = np.random.default_rng(12)
    F, H, W, N, vs, center = 4, 48, 80, 80, 5.0, (0.0, 0.0, 500.0)
    frames, poses = _scene(rng, F, H, W, N, vs, center, True)
    frames = np.clip(frames.astype(np.float32) + rng.uniform(-1, 1, frames.shape).astype(np.float32), 0, 255)  # ray_voxel.cpp:213-218
    want, tot = _oracle_run(oracle, frames, poses, N, vs, center, 2.0)
    pairs, _ = pvp.make_pairs(poses, W)
    grid = pvp.VoxelGrid(N, vs, center)
    st = pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, grid, 2.0)
    assert (st["n_rays"], st["n_steps"]) == (tot[0], tot[1])
    got = grid.data.cpu().numpy()
    assert np.array_equal(got != 0, want != 0)                      # identical support (visited set)
    np.testing.assert_allclose(got, want, rtol=RTOL_F32_ATOMICS, atol=0)
    want_fx, _ = _oracle_run(oracle, frames, poses, N, vs, center, 2.0, accum_mode=1, frac_bits=20)
    runs = []
    for _ in range(2):
        gfx = pvp.VoxelGrid(N, vs, center, accum="fixed", frac_bits=20)
        pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, gfx, 2.0)
        runs.append(gfx.data.cpu().numpy())
    assert np.array_equal(runs[0], want_fx) and np.array_equal(runs[0], runs[1])
    assert np.array_equal(bits(pvp.VoxelGrid(N, vs, center, accum="fixed", frac_bits=20, data=torch.as_tensor(want_fx, device=DEV)).to_float32().cpu().numpy()),
                          bits(oracle.fixed_to_f32(want_fx, 20)))


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("accum", ["f32", "fixed"])
def test_k2_variants_bit_exact(ctx, oracle, variant, accum):
    """Every K2 variant (segmented-scan aggregation, match.any aggregation, plain per-lane RED, lean lock-step) must give the
    reference's grid bit for bit on integer-valued frames, for whole grids and slabs, and the same counters."""
    rng = np.random.default_rng(20 + variant)
    F, H, W, N, vs, center = 4, 50, 131, 72, 5.0, (0.0, 0.0, 500.0)   # odd width: partial warps, unaligned rows
    frames, poses = _scene(rng, F, H, W, N, vs, center, True)
    pairs, _ = pvp.make_pairs(poses, W)
    mode = 0 if accum == "f32" else 1
    want, tot = _oracle_run(oracle, frames, poses, N, vs, center, 2.0, accum_mode=mode, frac_bits=10)
    ctx.set_option("dda_variant", variant)
    try:
        grid = pvp.VoxelGrid(N, vs, center, accum=accum, frac_bits=10)
        st = pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, grid, 2.0)
        assert (st["n_rays"], st["n_steps"], st["n_written"]) == tuple(tot)
        assert st["n_atomics"] <= st["n_written"] and (variant != 3 or st["n_atomics"] == st["n_written"])
        assert np.array_equal(grid.data.cpu().numpy(), want)
        parts = []
        for lo, hi in ((0, 30), (30, 31), (31, N)):
            s = pvp.VoxelGrid(N, vs, center, accum=accum, frac_bits=10, slab=(lo, hi))
            pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, s, 2.0)
            parts.append(s.data)
        assert np.array_equal(torch.cat(parts).cpu().numpy(), want)
        # float frames: same support, tolerance for fp32, exact for fixed
        ff = np.clip(frames.astype(np.float32) + rng.uniform(-1, 1, frames.shape).astype(np.float32), 0, 255)
        wantf, _ = _oracle_run(oracle, ff, poses, N, vs, center, 2.0, accum_mode=mode, frac_bits=10)
        gridf = pvp.VoxelGrid(N, vs, center, accum=accum, frac_bits=10)
        pvp.accumulate_motion(torch.as_tensor(ff, device=DEV), pairs, gridf, 2.0)
        if accum == "fixed":
            assert np.array_equal(gridf.data.cpu().numpy(), wantf)
        else:
            got = gridf.data.cpu().numpy()
            assert np.array_equal(got != 0, wantf != 0)
            np.testing.assert_allclose(got, wantf, rtol=RTOL_F32_ATOMICS, atol=0)
    finally:
        ctx.set_option("dda_variant", 0)


@pytest.mark.parametrize("rows", [1, 2, 4, 8])
def test_queue_tile_heights_bit_exact(ctx, oracle, rows):
    """K1 writes the ray queue in tiles of 1 / 2 / 4 / 8 image rows (the auto mode picks the height from the size of one
    x-plane of the grid: 8 rows + the one-ray kernel for 1024^3); the order must not change anything: same grid as the
    reference bit for bit with both lock-step kernels, uint8 and float frames, heights that leave a partial last tile."""
    rng = np.random.default_rng(300 + rows)
    F, H, W, N, vs, center = 3, 53, 131, 64, 5.0, (0.0, 0.0, 500.0)
    frames, poses = _scene(rng, F, H, W, N, vs, center, True)
    ff = np.clip(frames.astype(np.float32) + rng.uniform(-1, 1, frames.shape).astype(np.float32), 0, 255)
    pairs, _ = pvp.make_pairs(poses, W)
    want, tot = _oracle_run(oracle, frames, poses, N, vs, center, 2.0, accum_mode=1, frac_bits=10)
    wantf, totf = _oracle_run(oracle, ff, poses, N, vs, center, 2.0, accum_mode=1, frac_bits=10)
    ctx.set_option("k1_row_group", rows)
    try:
        for variant in (4, 5, 6):
            ctx.set_option("dda_variant", variant)
            for fr, w, t in ((frames, want, tot), (ff, wantf, totf)):
                grid = pvp.VoxelGrid(N, vs, center, accum="fixed", frac_bits=10)
                st = pvp.accumulate_motion(torch.as_tensor(fr, device=DEV), pairs, grid, 2.0)
                assert (st["n_rays"], st["n_steps"], st["n_written"]) == tuple(t), (variant, rows)
                assert np.array_equal(grid.data.cpu().numpy(), w), (variant, rows)
    finally:
        ctx.set_option("k1_row_group", 0)
        ctx.set_option("dda_variant", 0)


def test_edge_cases(ctx, oracle):
    N, vs, center = 32, 1.0, (0.0, 0.0, 0.0)
    W, H = 16, 8
    frames = np.zeros((3, H, W), np.uint8)
    frames[1, 2, 3] = 2      # |d| == threshold: not changed
    frames[2, 2, 3] = 5      # |5-2| = 3 > 2: one ray
    inside = pvp.FramePose((0.3, 0.2, 0.1))
    pairs, n = pvp.make_pairs([inside] * 3, W)
    grid = pvp.VoxelGrid(N, vs, center)
    assert pvp.accumulate_motion(torch.as_tensor(frames[:2], device=DEV), pairs, grid, 2.0, n_pairs=1)["n_rays"] == 0
    st = pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, grid, 2.0)
    want, tot = _oracle_run(oracle, frames, [inside] * 3, N, vs, center, 2.0)
    assert st["n_rays"] == 1 and st["n_steps"] == tot[1] > 0 and np.array_equal(grid.data.cpu().numpy(), want)
    # zero pairs: a no-op that still reports counters
    assert pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, grid, 2.0, n_pairs=0)["n_steps"] == 0
    # camera outside the +z face looking in: every ray enters through a max face and is dropped (ray_voxel.cpp:329-334)
    outside = pvp.FramePose((0.0, 0.0, 100.0))
    p2, _ = pvp.make_pairs([outside] * 2, W)
    g2 = pvp.VoxelGrid(N, vs, center)
    st = pvp.accumulate_motion(torch.as_tensor(synth.dense_frames(2, H, W), device=DEV), p2, g2, 2.0)
    assert st["n_rays"] > 100 and st["n_steps"] == 0 and not g2.data.any()
    # N = 1 grid, odd sizes, unaligned frame stride (scalar-load path of K1)
    for (h, w) in ((1, 1), (3, 5), (7, 33)):
        fr = synth.dense_frames(3, h, w, seed=h * w)
        pz = [pvp.FramePose((0.1, 0.1, 0.1), yaw=5.0 * i) for i in range(3)]
        pp, _ = pvp.make_pairs(pz, w)
        for n_ in (1, 5):
            gg = pvp.VoxelGrid(n_, 2.0, (0, 0, 0))
            pvp.accumulate_motion(torch.as_tensor(fr, device=DEV), pp, gg, 2.0)
            w_, _ = _oracle_run(oracle, fr, pz, n_, 2.0, (0, 0, 0), 2.0)
            assert np.array_equal(gg.data.cpu().numpy(), w_)
    with pytest.raises(pvp.PvpError, match="outside"):
        bad, _ = pvp.make_pairs([inside] * 8, W, pairs=[(0, 7)])
        pvp.accumulate_motion(frames, bad, grid, 2.0)


def test_full_size_properties_1080p_256(ctx):
    """BASELINE config 2 shape (1080p -> 256^3): size-independent properties instead of a CPU oracle run.
    (a) grid total == sum over rays of diff * steps (checked through the fixed-point grid, exactly);
    (b) linearity: accumulating pairs A then B equals accumulating A+B in one batch;
    (c) fixed-point determinism across runs; (d) fp32 result within 1e-5 of the fixed-point result."""
    F, H, W, N, vs, center = 3, 1080, 1920, 256, 2.0, (0.0, 0.0, 500.0)
    frames = torch.as_tensor(synth.sparse_frames(F, H, W, seed=5, blob=96, blobs=4), device=DEV)
    poses = synth.orbit_poses(F, N, vs, center, seed=6)
    pairs, n = pvp.make_pairs(poses, W)
    gfx = pvp.VoxelGrid(N, vs, center, accum="fixed", frac_bits=8)
    st = pvp.accumulate_motion(frames, pairs, gfx, 2.0)
    assert st["n_rays"] > 5000 and st["n_steps"] > st["n_rays"] * 50
    g2 = pvp.VoxelGrid(N, vs, center, accum="fixed", frac_bits=8)
    a, _ = pvp.make_pairs(poses, W, pairs=[(0, 1)])
    b, _ = pvp.make_pairs(poses, W, pairs=[(1, 2)])
    sa = pvp.accumulate_motion(frames, a, g2, 2.0)
    sb = pvp.accumulate_motion(frames, b, g2, 2.0)
    assert sa["n_steps"] + sb["n_steps"] == st["n_steps"]
    assert torch.equal(gfx.data, g2.data)
    # (a): per-ray step counts from the stage API, diff from detect_motion
    total = 0
    for (p, c) in ((0, 1), (1, 2)):
        ch, d = pvp.detect_motion(frames[p], frames[c], 2.0)
        pix = torch.nonzero(ch.reshape(-1)).reshape(-1)
        dirs = pvp.pixel_rays(pix, W, H, poses[c])
        org = torch.tensor(poses[c].camera_position, device=DEV, dtype=torch.float32).expand(len(pix), 3)
        counts, _, _, _ = pvp.cast_rays(org, dirs, pvp.VoxelGrid(N, vs, center, data=torch.zeros((N, N, N), device=DEV)), want_steps=False)
        total += int((counts.to(torch.int64) * d.reshape(-1)[pix].to(torch.int64)).sum())
    assert int(gfx.data.sum()) == total * 256
    gf = pvp.VoxelGrid(N, vs, center)
    pvp.accumulate_motion(frames, pairs, gf, 2.0)
    torch.testing.assert_close(gf.data, gfx.to_float32(), rtol=RTOL_F32_ATOMICS, atol=0)


def test_rays_leaving_through_low_corners_and_tiny_grids(ctx, oracle):
    """Regression: the lock-step kernel keys voxels by their 32-bit linear offset and reserves 0xffffffff.  A ray that steps
    out of the grid downwards from offset 0 / N-1 / N^2-1 wraps to exactly that value, and in a slab kernel a LIVE voxel
    (lo-1, N-1, N-1) has that relative offset.  Counters and grids must still equal the reference: camera near the low corner
    looking at it, tiny grids, every slab split."""
    rng = np.random.default_rng(77)
    H, W = 24, 40
    for N, vs in ((1, 16.0), (2, 8.0), (3, 4.0), (5, 3.0), (8, 2.0)):
        center = (0.0, 0.0, 0.0)
        half = N * vs / 2
        frames = synth.dense_frames(4, H, W, seed=N)
        # inside the grid, looking (mostly) towards the (-,-,-) corner / the -z face, wide field of view
        poses = [pvp.FramePose((float(np.float32(-half + vs * 0.6 + 0.3 * i)), float(np.float32(-half + vs * 0.7)), float(np.float32(-half + vs * (1.1 if N > 1 else 0.6) + 0.2 * i))),
                               yaw=float(rng.uniform(-40, 40)), pitch=float(rng.uniform(-40, 40)), roll=float(rng.uniform(-20, 20)), fov_degrees=100.0)
                 for i in range(4)]
        pairs, _ = pvp.make_pairs(poses, W)
        want, tot = _oracle_run(oracle, frames, poses, N, vs, center, 2.0)
        assert tot[1] > 500
        for variant in (0, 1, 3, 5, 6):
            ctx.set_option("dda_variant", variant)
            try:
                grid = pvp.VoxelGrid(N, vs, center)
                st = pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, grid, 2.0)
                assert (st["n_rays"], st["n_steps"], st["n_written"]) == tuple(tot), (N, variant)
                assert np.array_equal(grid.data.cpu().numpy(), want)
                planes_written = 0
                for lo in range(N + 1):
                    for hi in sorted({lo, min(N, lo + 1), N}):
                        s = pvp.VoxelGrid(N, vs, center, slab=(lo, hi))
                        ss = pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, s, 2.0)
                        # variants 1-3 walk every ray in full on a slab; the default kernels clip (rays that cannot reach the slab are
                        # not walked, a walk stops once x has passed it)
                        assert ss["n_rays"] == tot[0] and ss["n_written"] <= ss["n_steps"] <= tot[1], (N, variant, lo, hi)
                        assert variant in (0, 5, 6) or ss["n_steps"] == tot[1], (N, variant, lo, hi)
                        assert np.array_equal(s.data.cpu().numpy(), want[lo:hi]), (N, variant, lo, hi)
                        if hi == lo + 1:
                            planes_written += ss["n_written"]
                assert planes_written == tot[1]     # the single-plane slabs partition the voxel steps
            finally:
                ctx.set_option("dda_variant", 0)


@pytest.mark.parametrize("variant", [0, 5, 6])
def test_random_geometry_sweep_vs_oracle(ctx, oracle, variant):
    """Broad randomised parity sweep of the default (lean lock-step) kernel: cameras inside and outside the grid (outside ones
    look in through min or max faces -- the latter cast no steps, ray_voxel.cpp:329-334), full-range yaw / pitch / roll, fov
    20..140 degrees, odd and even N, frame sizes that leave partial warps, both accumulation modes and a random slab."""
    rng = np.random.default_rng(4242)
    total_steps = 0
    ctx.set_option("dda_variant", variant)
    try:
        total_steps = _geometry_sweep(ctx, oracle, rng)
    finally:
        ctx.set_option("dda_variant", 0)
    assert total_steps > 150000


def _geometry_sweep(ctx, oracle, rng):
    total_steps = 0
    for trial in range(60):
        N = int(rng.choice([1, 2, 3, 7, 16, 31, 40]))
        vs = float(rng.choice([0.5, 1.0, 2.5, 6.0]))
        center = tuple(float(np.float32(c)) for c in rng.uniform(-50, 50, 3))
        half = N * vs / 2
        H, W = int(rng.integers(1, 40)), int(rng.integers(1, 70))
        F = 3
        frames = synth.dense_frames(F, H, W, seed=1000 + trial)
        poses = []
        for f in range(F):
            where = rng.choice(["inside", "outside", "far"])
            scale = {"inside": 0.95, "outside": 2.0, "far": 6.0}[where]
            pos = np.array(center) + rng.uniform(-1, 1, 3) * half * scale
            poses.append(pvp.FramePose(tuple(float(np.float32(p)) for p in pos), yaw=float(np.float32(rng.uniform(-180, 180))),
                                       pitch=float(np.float32(rng.uniform(-90, 90))), roll=float(np.float32(rng.uniform(-180, 180))),
                                       fov_degrees=float(np.float32(rng.uniform(20, 140)))))
        pairs, _ = pvp.make_pairs(poses, W)
        for accum, mode in (("f32", 0), ("fixed", 1)):
            want, tot = _oracle_run(oracle, frames, poses, N, vs, center, 2.0, accum_mode=mode, frac_bits=9)
            grid = pvp.VoxelGrid(N, vs, center, accum=accum, frac_bits=9)
            st = pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, grid, 2.0)
            assert (st["n_rays"], st["n_steps"], st["n_written"]) == tuple(tot), (trial, accum)
            assert np.array_equal(grid.data.cpu().numpy(), want), (trial, accum)
            lo = int(rng.integers(0, N + 1))
            hi = int(rng.integers(lo, N + 1))
            s = pvp.VoxelGrid(N, vs, center, accum=accum, frac_bits=9, slab=(lo, hi))
            ss = pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, s, 2.0)
            assert np.array_equal(s.data.cpu().numpy(), want[lo:hi]), (trial, accum, lo, hi)
            assert ss["n_rays"] == tot[0] and ss["n_written"] <= ss["n_steps"] <= tot[1], (trial, accum, lo, hi)   # slab walks are clipped
        total_steps += tot[1]
    return total_steps


def test_opt_in_distance_attenuation(ctx, oracle):
    """attenuation_alpha > 0 applies the factor the reference computes and discards (ray_voxel.cpp:524-526),
    val = pix_val / (1 + alpha * distance) per voxel step.  Opt-in extension: checked against the oracle's restatement of those
    lines with the factor applied (fixed point: bit for bit -- every addend is quantised from the same IEEE fp32 product;
    float32: same support, rtol 1e-5); alpha = 0 stays the reference's behaviour bit for bit."""
    rng = np.random.default_rng(88)
    F, H, W, N, vs, center = 4, 33, 52, 40, 5.0, (0.0, 0.0, 500.0)
    frames, poses = _scene(rng, F, H, W, N, vs, center, True)
    pairs, _ = pvp.make_pairs(poses, W)
    alpha = 0.1

    def oracle_att(mode, slab=None):
        lo, hi = (0, N) if slab is None else slab
        grid = np.zeros((hi - lo, N, N), np.float32 if mode == 0 else np.int64)
        tot = np.zeros(3, np.int64)
        for c in range(1, F):
            R = oracle.rotation_matrix(poses[c].yaw, poses[c].pitch, poses[c].roll)
            f = oracle.focal_length(poses[c].fov_degrees, W)
            tot += oracle.accumulate_pair(frames[c - 1], frames[c], np.array(poses[c].camera_position, np.float32), R, f, grid, N, vs,
                                          np.array(center, np.float32), 2.0, accum_mode=mode, frac_bits=14, slab=slab, attenuation_alpha=alpha)
        return grid, tot

    plain, _ = _oracle_run(oracle, frames, poses, N, vs, center, 2.0)
    for variant in (0, 3, 5):
        ctx.set_option("dda_variant", variant)
        try:
            want_fx, tot = oracle_att(1)
            g = pvp.VoxelGrid(N, vs, center, accum="fixed", frac_bits=14, attenuation_alpha=alpha)
            st = pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, g, 2.0)
            assert (st["n_rays"], st["n_steps"], st["n_written"]) == tuple(tot)
            assert np.array_equal(g.data.cpu().numpy(), want_fx), variant
            want_f, _ = oracle_att(0)
            gf = pvp.VoxelGrid(N, vs, center, attenuation_alpha=alpha)
            pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, gf, 2.0)
            got = gf.data.cpu().numpy()
            assert np.array_equal(got != 0, want_f != 0)
            np.testing.assert_allclose(got, want_f, rtol=RTOL_F32_ATOMICS, atol=0)
            assert (got < plain).any() and (got <= plain + 1e-3).all()      # the factor is < 1 for distance > 0
            want_s, _ = oracle_att(1, slab=(7, 29))
            gs = pvp.VoxelGrid(N, vs, center, accum="fixed", frac_bits=14, slab=(7, 29), attenuation_alpha=alpha)
            pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, gs, 2.0)
            assert np.array_equal(gs.data.cpu().numpy(), want_s), variant
            g0 = pvp.VoxelGrid(N, vs, center, attenuation_alpha=0.0)
            pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, g0, 2.0)
            assert np.array_equal(bits(g0.data.cpu().numpy()), bits(plain))
        finally:
            ctx.set_option("dda_variant", 0)
    with pytest.raises(pvp.PvpError, match="attenuation_alpha"):
        pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, pvp.VoxelGrid(N, vs, center, attenuation_alpha=-1.0), 2.0)


@pytest.mark.parametrize("kind", ["dense", "scatter", "object", "float"])
def test_zorder_queue_all_motion_kinds_vs_oracle(ctx, oracle, kind):
    """The Z-order ray queue (motion_compact_morton_kernel) on frames that span several 64 x 64 footprints, with widths that are
    and are not multiples of 4 (vector / scalar loads), for every kind of motion the order treats differently: dense frames (strip
    order inside footprints, lean2 picked), iid-scattered motion (Morton order, lean picked by the paired share), a moving textured
    object (dense and empty footprints side by side), float frames; blocks of 4 and of 8 rows.  Counters and grids against the C port,
    bit-exact (uint8 frames, and fixed point always)."""
    rng = np.random.default_rng({"dense": 1, "scatter": 2, "object": 3, "float": 4}[kind])
    N, vs, center = 48, 6.0, (0.0, 0.0, 500.0)
    for (H, W) in ((150, 200), (97, 131), (64, 64), (65, 257)):
        F = 3
        if kind == "dense":
            frames = synth.dense_frames(F, H, W, seed=int(rng.integers(1 << 30)))
        elif kind == "scatter":
            frames = synth.scatter_frames(F, H, W, seed=int(rng.integers(1 << 30)), fraction=0.1)
        elif kind == "object":
            frames = synth.object_frames(F, H, W, seed=int(rng.integers(1 << 30)))
        else:
            frames = (synth.dense_frames(F, H, W, seed=int(rng.integers(1 << 30))).astype(np.float32) + rng.random((F, H, W)).astype(np.float32))
        poses = synth.orbit_poses(F, N, vs, center, seed=int(rng.integers(1 << 30)))
        pairs, _ = pvp.make_pairs(poses, W)
        for mode, accum in ((1, "fixed"), (0, "f32")):
            want, tot = _oracle_run(oracle, frames, poses, N, vs, center, 2.0, accum_mode=mode, frac_bits=10)
            for rg in (-4, -8):
                ctx.set_option("k1_row_group", rg)
                try:
                    grid = pvp.VoxelGrid(N, vs, center, accum=accum, frac_bits=10)
                    st = pvp.accumulate_motion(torch.as_tensor(frames, device=DEV), pairs, grid, 2.0)
                finally:
                    ctx.set_option("k1_row_group", 0)
                assert (st["n_rays"], st["n_steps"], st["n_written"]) == tuple(tot), (kind, H, W, accum, rg)
                assert st["n_rays"] > 0
                got = grid.data.cpu().numpy()
                if accum == "fixed" or kind != "float":
                    assert np.array_equal(got, want), (kind, H, W, accum, rg)
                else:
                    np.testing.assert_allclose(got, want, rtol=1e-5, atol=0)


def test_lockstep_pick_follows_ray_count_pairing_and_cell_type(ctx):
    """The device-side choice between the three lock-step kernels (select_lockstep_kernel): dense batches of >= 6 M rays into float32
    cells run on lean4 (four rays per lane), smaller or int64 batches on lean2, scattered motion and tiny batches on lean (one ray per
    lane); option lean4 = -1 keeps lean4 out.  lean4 reports n_atomics only when asked (count_atomics).  Whatever is picked, the grid is
    the same (uint8 frames: exact)."""
    H, W, N, vs, center = 1080, 1920, 256, 4.0, (0.0, 0.0, 500.0)
    F = 5
    dense = torch.as_tensor(synth.dense_frames(F, H, W, seed=5), device=DEV)
    scat = torch.as_tensor(synth.scatter_frames(F, H, W, seed=6, fraction=0.1), device=DEV)
    poses = synth.orbit_poses(F, N, vs, center, seed=7)
    tab, _ = pvp.make_pairs(poses, W)                         # 4 pairs: 8 M rays when everything moves
    one, _ = pvp.make_pairs(poses[:2], W)

    def run(frames, table, accum="f32", **opts):
        for k, v in opts.items():
            ctx.set_option(k, v)
        try:
            g = pvp.VoxelGrid(N, vs, center, accum=accum, frac_bits=4)
            st = pvp.accumulate_motion(frames, table, g, 2.0)
            return ctx.last_k2_pick(), st, g.data
        finally:
            for k in opts:
                ctx.set_option(k, 0)

    def same(a, b):     # float32 sums are exact and order-independent below 2^24; the few hotter cells around the camera round per RED
        small = (a < 2 ** 24) & (b < 2 ** 24)
        return torch.equal(a[small], b[small]) and torch.allclose(a[~small], b[~small], rtol=1e-4, atol=0)

    p4, st4, g4 = run(dense, tab)
    assert p4 == 4 and st4["n_rays"] > 6_100_000 and st4["n_atomics"] == 0
    p4c, st4c, g4c = run(dense, tab, count_atomics=1)
    assert p4c == 4 and 0 < st4c["n_atomics"] < 0.3 * st4c["n_written"] and same(g4c, g4)
    p2, st2, g2 = run(dense, tab, lean4=-1)
    assert p2 == 2 and 0 < st2["n_atomics"] < 0.3 * st2["n_written"]
    assert (st2["n_rays"], st2["n_steps"], st2["n_written"]) == (st4["n_rays"], st4["n_steps"], st4["n_written"]) and same(g2, g4)
    assert run(dense, tab, accum="fixed")[0] == 2             # int64 cells: lean4 is no faster there
    assert run(dense, one)[0] == 2                            # 2 M rays: below lean4's threshold
    assert run(scat, tab)[0] == 1                             # 10 % iid motion: few vertical pairs
    assert run(dense[:, :64, :64].contiguous(), one)[0] == 1  # 4 k rays: too few for 64-ray chunks
    p1, st1, g1 = run(dense, tab, dda_variant=4)
    assert p1 == 1 and same(g1, g4)
    assert run(dense, tab, dda_variant=3)[0] == 0             # not a lock-step kernel
