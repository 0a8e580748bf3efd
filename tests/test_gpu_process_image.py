"""Parity of path B (process_image_cpp) on the GPU: the reference module's own outputs (golden, generated with
OMP_NUM_THREADS=1) and the oracle on fresh inputs.  Voxel indices are bit-reproducible; fp64 sums differ from the
serial reference only by summation order (rtol 1e-12); fixed-point mode is bit-exact against the oracle."""
import numpy as np
import pytest
import torch

from conftest import bits, golden

import pixeltovoxelprojector_b200 as pvp
import process_image_cpp as dropin

pytestmark = pytest.mark.gpu
RTOL_F64_ATOMICS = 1e-12


def call(fn, g, name, img, grid, tex, **kw):
    H, W = g["image"].shape
    return fn(img, list(g["earth_position"]), list(g[f"{name}_pointing"]), float(g["fov"]), W, H, grid, [tuple(r) for r in g["extent"]],
              float(g["max_distance"]), int(g["num_steps"]), tex, *[float(s) for s in g["sky"]], bool(g[f"{name}_update"]), bool(g[f"{name}_bgsub"]), **kw)


@pytest.mark.parametrize("name", ["update", "bgsub", "zenith"])
def test_golden_numpy_dropin(ctx, name):
    g = golden("process_image_golden.npz")
    grid = np.zeros_like(g[f"{name}_grid_out"])
    tex = g[f"{name}_tex_in"].copy()
    assert call(dropin.process_image_cpp, g, name, g["image"], grid, tex) is None   # mutates in place, returns None
    want_g, want_t = g[f"{name}_grid_out"], g[f"{name}_tex_out"]
    assert np.array_equal(grid != 0, want_g != 0)                                   # identical visited voxels
    np.testing.assert_allclose(grid, want_g, rtol=RTOL_F64_ATOMICS, atol=0)
    assert np.array_equal(tex != 0, want_t != 0)
    np.testing.assert_allclose(tex, want_t, rtol=RTOL_F64_ATOMICS, atol=0)


def test_golden_device_tensors_and_fixed_point(ctx, oracle):
    g = golden("process_image_golden.npz")
    name = "update"
    img = torch.as_tensor(g["image"], device="cuda")
    grid = torch.zeros(g[f"{name}_grid_out"].shape, dtype=torch.float64, device="cuda")
    tex = torch.as_tensor(g[f"{name}_tex_in"], device="cuda").clone()
    call(pvp.process_image_cpp, g, name, img, grid, tex, want_stats=True)
    from pixeltovoxelprojector_b200 import process_image as pi
    np.testing.assert_allclose(grid.cpu().numpy(), g[f"{name}_grid_out"], rtol=RTOL_F64_ATOMICS, atol=0)
    H, W = g["image"].shape
    og, ot = np.zeros(grid.shape, np.int64), np.zeros(tex.shape, np.int64)
    nr, ns = call(oracle.process_image, g, name, g["image"], og, ot, accum_mode=1, frac_bits=30)
    assert pi.last_stats == {"n_rays": nr, "n_samples": ns}
    runs = []
    for _ in range(2):
        gi = torch.zeros(grid.shape, dtype=torch.int64, device="cuda")
        ti = torch.zeros(tex.shape, dtype=torch.int64, device="cuda")
        call(pvp.process_image_cpp, g, name, img, gi, ti, accum_mode="fixed", frac_bits=30)
        runs.append((gi.cpu().numpy(), ti.cpu().numpy()))
    assert np.array_equal(runs[0][0], og) and np.array_equal(runs[0][1], ot)
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])


def test_strides_empty_grid_and_errors(ctx):
    g = golden("process_image_golden.npz")
    name = "update"
    want_g, want_t = g[f"{name}_grid_out"], g[f"{name}_tex_out"]
    big = np.zeros((want_g.shape[0], want_g.shape[1] * 2, want_g.shape[2]))
    view = big[:, ::2, :]                                     # strided view: updated in place like the reference
    tex = np.asfortranarray(g[f"{name}_tex_in"].copy())       # F-order texture
    call(dropin.process_image_cpp, g, name, g["image"], view, tex)
    np.testing.assert_allclose(view, want_g, rtol=RTOL_F64_ATOMICS, atol=0)
    np.testing.assert_allclose(tex, want_t, rtol=RTOL_F64_ATOMICS, atol=0)
    assert not big[:, 1::2, :].any()
    rev = np.zeros_like(want_g)[::-1]                         # negative stride: copy-in / copy-out
    call(dropin.process_image_cpp, g, name, g["image"], rev, g[f"{name}_tex_in"].copy())
    np.testing.assert_allclose(rev, want_g, rtol=RTOL_F64_ATOMICS, atol=0)
    tex2 = g[f"{name}_tex_in"].copy()                         # empty grid of any ndim => voxel stage skipped (:152)
    call(dropin.process_image_cpp, g, name, g["image"], np.zeros((0,)), tex2)
    np.testing.assert_allclose(tex2, want_t, rtol=RTOL_F64_ATOMICS, atol=0)
    with pytest.raises(ValueError, match="incorrect number of dimensions: 2; expected 3"):
        call(dropin.process_image_cpp, g, name, g["image"], np.zeros((4, 4)), tex2)
    with pytest.raises(TypeError):
        dropin.process_image_cpp(g["image"])                  # arity
    g32 = np.zeros(want_g.shape, np.float32)                  # forcecast: silently copied, caller sees no update
    with pytest.warns(UserWarning, match="not float64"):
        call(dropin.process_image_cpp, g, name, g["image"], g32, g[f"{name}_tex_in"].copy())
    assert not g32.any()


def test_division_by_extent_equals_ieee_division(ctx):
    """The lock-step kernel divides by the grid extent with a host-prepared reciprocal (Markstein sequence); it must
    equal IEEE division bit for bit wherever the library chooses that kernel.  4e8 numerators per extent, plus +-64 ulps
    around every voxel boundary; extents the sequence is not proven for must be declared ineligible."""
    import ctypes as C
    from pixeltovoxelprojector_b200._lib import check
    rng = np.random.default_rng(77)
    extents = [20.0, 17.0, 512.0, 0.3, 1e-3, 123456.789, 3.0, 2.0 / 3.0, float(np.nextafter(2.0, 0)) / 4, 1e200, 7e-190]
    extents += [float(x) for x in np.exp(rng.uniform(-40, 40, 12))]
    for b in extents:
        bad, ok = C.c_uint64(0), C.c_int32(0)
        check(ctx._lib.pvp_selftest_div_by_extent(ctx.handle, b, 40_000_000, 512, int(rng.integers(1 << 62)), C.byref(bad), C.byref(ok)))
        all_ones = np.frexp(b)[0] == 1.0 - 2.0 ** -53
        too_big = not (2.0 ** -500 < b < 2.0 ** 499)
        assert bool(ok.value) == (not all_ones and not too_big), b
        assert bad.value == 0, (b, bad.value)


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_random_scenes_vs_oracle(ctx, oracle, variant):
    """variant 1 = generic per-pixel kernel, 2 = lock-step kernel with warp-aggregated REDs, one pixel per lane, 3 = two pixels
    per lane (the default where eligible)."""
    ctx.set_option("image_variant", variant)
    try:
        _random_scenes(oracle)
    finally:
        ctx.set_option("image_variant", 0)


def test_lockstep_fixed_point_wide_and_ineligible_extent(ctx, oracle):
    """Fixed point with values >= 2^27 takes the 64-bit run sums; an extent whose significand is all ones falls back to
    the generic kernel; both must match the oracle exactly."""
    rng = np.random.default_rng(31)
    H, W = 40, 52
    img = rng.random((H, W)) * 1000.0          # * 2^24 >> 2^27: wide sums
    pos, pointing = [0.1, -0.2, 0.3], [0.05, 0.02, 1.0]
    odd = float(np.nextafter(32.0, 0))          # 31.999...: significand all ones
    for ext in ([(-8.0, 9.0), (-7.0, 6.5), (10.0, 28.0)], [(0.0, odd), (-7.0, 6.5), (10.0, 28.0)]):
        og, ot = np.zeros((20, 24, 28), np.int64), np.zeros((8, 8), np.int64)
        nr, ns = oracle.process_image(img, pos, pointing, 0.9, W, H, og, ext, 60.0, 240, ot, 0.0, 0.0, 1.0, 1.0, True, False, accum_mode=1, frac_bits=24)
        gi = torch.zeros(og.shape, dtype=torch.int64, device="cuda")
        ti = torch.zeros(ot.shape, dtype=torch.int64, device="cuda")
        pvp.process_image_cpp(torch.as_tensor(img, device="cuda"), pos, pointing, 0.9, W, H, gi, ext, 60.0, 240, ti, 0.0, 0.0, 1.0, 1.0, True, False,
                              accum_mode="fixed", frac_bits=24, want_stats=True)
        from pixeltovoxelprojector_b200 import process_image as pi
        assert pi.last_stats == {"n_rays": nr, "n_samples": ns} and ns > 1000
        assert np.array_equal(gi.cpu().numpy(), og) and np.array_equal(ti.cpu().numpy(), ot)


def _random_scenes(oracle):
    rng = np.random.default_rng(9)
    for trial in range(6):
        H, W = int(rng.integers(8, 70)), int(rng.integers(8, 70))
        img = rng.random((H, W))
        img[img < 0.3] = 0
        pos = rng.uniform(-5, 5, 3)
        pointing = rng.normal(size=3) if trial else np.array([0.0, 0.0, -3.0])
        ext = [(-8.0, 9.0), (-7.0, 6.5), (-6.0, 8.0)] if trial % 2 else [(10.0, 30.0), (-10.0, 10.0), (-10.0, 10.0)]
        shape = tuple(int(x) for x in rng.integers(4, 40, 3))
        sky = [float(rng.uniform(0, 6.28)), float(rng.uniform(-1.5, 1.5)), float(rng.uniform(0.5, 6.0)), float(rng.uniform(0.5, 3.0))]
        for upd, bg in ((True, False), (False, True)):
            g1, g2 = np.zeros(shape), np.zeros(shape)
            t1 = rng.random((16, 24)) * 0.4
            t2 = t1.copy()
            oracle.process_image(img, pos, pointing, 0.9, W, H, g1, ext, 60.0, 300, t1, *sky, upd, bg)
            dropin.process_image_cpp(img, list(pos), list(pointing), 0.9, W, H, g2, ext, 60.0, 300, t2, *sky, upd, bg)
            assert np.array_equal(g1 != 0, g2 != 0), (trial, upd, bg)
            np.testing.assert_allclose(g2, g1, rtol=RTOL_F64_ATOMICS, atol=0)
            np.testing.assert_allclose(t2, t1, rtol=RTOL_F64_ATOMICS, atol=0)


def test_full_size_properties_4096_512_fixed(ctx):
    """BASELINE config 5 shape (4096^2 image -> 512^3 grid, fixed point): determinism, linearity in brightness
    (2x image => exactly 2x grid) and conservation (grid total == n_samples * brightness for a constant image)."""
    from pixeltovoxelprojector_b200 import process_image as pi
    H = W = 4096
    n = 512
    ext = [(-256.0, 256.0), (-256.0, 256.0), (100.0, 612.0)]
    args = ([0.0, 0.0, 0.0], [0.02, 0.01, 1.0], 0.8, W, H)
    tail = (ext, 700.0, 700, None, 0.0, 0.0, 1.0, 1.0, False, False)
    img = torch.full((H, W), 0.5, dtype=torch.float64, device="cuda")
    grids = []
    for scale in (1.0, 1.0, 2.0):
        gi = torch.zeros((n, n, n), dtype=torch.int64, device="cuda")
        ti = torch.zeros((8, 8), dtype=torch.int64, device="cuda")
        pvp.process_image_cpp(img * scale, *args, gi, tail[0], tail[1], tail[2], ti, *tail[4:], accum_mode="fixed", frac_bits=20, want_stats=True)
        grids.append(gi)
        assert int(gi.sum()) == pi.last_stats["n_samples"] * int(0.5 * scale * (1 << 20))
    assert pi.last_stats["n_rays"] == H * W and pi.last_stats["n_samples"] > 50 * H * W
    assert torch.equal(grids[0], grids[1]) and torch.equal(grids[2], grids[0] * 2)


@pytest.mark.parametrize("variant", [2, 3, 4])
def test_lockstep_random_sweep_fixed_point_exact(ctx, oracle, variant):
    """Broad randomised sweep of the lock-step kernel in fixed-point mode (bit-exact against the oracle): observer inside
    and outside the grid, non-cubic and strided grids, steps longer and shorter than a voxel, max_distance clipping inside
    the grid, pointing along +-z (the up-vector switch, process_image.cpp:203-211), widths that leave partial warps, zero
    pixels, texture updates."""
    from pixeltovoxelprojector_b200 import process_image as pi
    rng = np.random.default_rng(515)
    ctx.set_option("image_variant", variant)
    try:
        total = _lockstep_sweep(oracle, pi, rng)
    finally:
        ctx.set_option("image_variant", 0)
    assert total > 300000


def _lockstep_sweep(oracle, pi, rng):
    total = 0
    for trial in range(36):
        H, W = int(rng.integers(1, 50)), int(rng.integers(1, 90))
        img = rng.random((H, W))
        img[rng.random((H, W)) < 0.25] = 0.0
        shape = tuple(int(x) for x in rng.integers(1, 36, 3))
        lo = rng.uniform(-20, 5, 3)
        ext = [(float(lo[k]), float(lo[k] + rng.uniform(0.5, 40))) for k in range(3)]
        mid = np.array([(a + b) / 2 for a, b in ext])
        size = np.array([b - a for a, b in ext])
        kind = trial % 4
        pos = mid + rng.uniform(-0.45, 0.45, 3) * size if kind == 0 else mid + rng.normal(size=3) * size * 1.5
        if kind == 2:
            pointing = np.array([0.0, 0.0, float(rng.choice([-2.0, 3.0]))])       # parallel to z: up switches to y
        else:
            pointing = (mid - pos) + rng.normal(size=3) * size * 0.3
            if not np.any(pointing):
                pointing = np.array([0.3, 0.2, 1.0])
        max_distance = float(rng.uniform(0.3, 3.0) * np.linalg.norm(size))
        num_steps = int(rng.choice([7, 40, 300, 1500]))
        fov = float(rng.uniform(0.2, 2.2))
        sky = [float(rng.uniform(0, 6.28)), float(rng.uniform(-1.5, 1.5)), float(rng.uniform(0.5, 6.0)), float(rng.uniform(0.5, 3.0))]
        upd = bool(trial % 3 == 0)
        og, ot = np.zeros(shape, np.int64), np.zeros((6, 9), np.int64)
        nr, ns = oracle.process_image(img, pos, pointing, fov, W, H, og, ext, max_distance, num_steps, ot, *sky, upd, False, accum_mode=1, frac_bits=22)
        # a strided device grid: every other plane of a larger tensor along a random axis
        ax = int(rng.integers(0, 3))
        big_shape = list(shape)
        big_shape[ax] *= 2
        big = torch.zeros(big_shape, dtype=torch.int64, device="cuda")
        view = big[tuple(slice(None, None, 2) if k == ax else slice(None) for k in range(3))]
        ti = torch.zeros(ot.shape, dtype=torch.int64, device="cuda")
        pvp.process_image_cpp(torch.as_tensor(img, device="cuda"), list(pos), list(pointing), fov, W, H, view, ext, max_distance, num_steps, ti, *sky, upd, False,
                              accum_mode="fixed", frac_bits=22, want_stats=True)
        assert pi.last_stats == {"n_rays": nr, "n_samples": ns}, trial
        assert np.array_equal(view.cpu().numpy(), og), trial
        assert int(big.sum()) == int(og.sum())                                   # nothing written between the planes of the view
        if upd:
            assert np.array_equal(ti.cpu().numpy(), ot), trial
        total += ns
    return total


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_row_and_file_shards_sum_to_whole(ctx, oracle, variant):
    """SURVEY 8e, path B over several GPUs (the ranks run one after another here): row shards of one image -- ragged,
    odd boundaries that split the kernels' 2-row tiles, an empty shard -- and file shards of an observation list, each
    into private int64 tensors, sum to exactly the single-GPU result, which equals the oracle; float64 mode agrees to
    summation order.  A row range is part of the C ABI (pvp_image_args.row_begin / row_end)."""
    from pixeltovoxelprojector_b200 import distributed as pd, process_image as pi
    rng = np.random.default_rng(2024)
    H, W = 37, 45
    ext, shape = [(-8.0, 9.0), (-7.0, 6.5), (10.0, 28.0)], (20, 24, 28)
    sky = (0.3, -0.2, 1.0, 1.5)
    obs = []
    for k in range(5):
        img = rng.random((H, W))
        img[rng.random((H, W)) < 0.2] = 0.0
        obs.append((img, [0.1 * k, -0.2, 0.3], [0.05 - 0.02 * k, 0.02, 1.0], 0.9))
    og, ot = np.zeros(shape, np.int64), np.zeros((16, 24), np.int64)
    want_stats = [0, 0]
    for img, pos, pointing, fov in obs:
        nr, ns = oracle.process_image(img, pos, pointing, fov, W, H, og, ext, 60.0, 240, ot, *sky, True, False, accum_mode=1, frac_bits=24)
        want_stats[0] += nr
        want_stats[1] += ns
    dev_obs = [(torch.as_tensor(img, device="cuda"), pos, pointing, fov) for img, pos, pointing, fov in obs]

    def zeros(dtype=torch.int64):
        return torch.zeros(shape, dtype=dtype, device="cuda"), torch.zeros((16, 24), dtype=dtype, device="cuda")
    ctx.set_option("image_variant", variant)
    try:
        for world in (1, 2, 3, 5, 64):                              # 64 ranks for 37 rows / 5 files: most shards are empty
            tg, tt = zeros()
            tg_files, tt_files = zeros()
            rays = samples = 0
            for rank in range(world):
                g, t = zeros()
                for image, pos, pointing, fov in dev_obs:
                    rows = pd.process_image_sharded(image, pos, pointing, fov, W, H, g, ext, 60.0, 240, t, *sky, True, False,
                                                    rank=rank, world=world, accum_mode="fixed", frac_bits=24, want_stats=True)
                    assert rows == pd.shard_range(H, rank, world)
                    rays += pi.last_stats["n_rays"]
                    samples += pi.last_stats["n_samples"]
                tg += g
                tt += t
                g, t = zeros()
                pd.process_images_sharded(dev_obs, W, H, g, ext, 60.0, 240, t, *sky, rank=rank, world=world, accum_mode="fixed", frac_bits=24)
                tg_files += g
                tt_files += t
            assert [rays, samples] == want_stats, world
            assert np.array_equal(tg.cpu().numpy(), og) and np.array_equal(tt.cpu().numpy(), ot), world
            assert np.array_equal(tg_files.cpu().numpy(), og) and np.array_equal(tt_files.cpu().numpy(), ot), world
        # float64 accumulation: 3 row shards against the whole image
        wg, wt = zeros(torch.float64)
        sg, st = zeros(torch.float64)
        image, pos, pointing, fov = dev_obs[1]
        pvp.process_image_cpp(image, pos, pointing, fov, W, H, wg, ext, 60.0, 240, wt, *sky, True, False)
        for rank in range(3):
            pd.process_image_sharded(image, pos, pointing, fov, W, H, sg, ext, 60.0, 240, st, *sky, True, False, rank=rank, world=3)
        assert torch.equal(sg != 0, wg != 0)
        torch.testing.assert_close(sg, wg, rtol=RTOL_F64_ATOMICS, atol=0)
        torch.testing.assert_close(st, wt, rtol=RTOL_F64_ATOMICS, atol=0)
        # numpy host path takes the row range too
        hg, ht = np.zeros(shape), np.zeros((16, 24))
        for lo, hi in ((0, 9), (9, 10), (10, H)):
            dropin.process_image_cpp(obs[1][0], pos, pointing, fov, W, H, hg, ext, 60.0, 240, ht, *sky, True, False, row_range=(lo, hi))
        np.testing.assert_allclose(hg, wg.cpu().numpy(), rtol=RTOL_F64_ATOMICS, atol=0)
        with pytest.raises(ValueError, match="row_range"):
            pvp.process_image_cpp(image, pos, pointing, fov, W, H, wg, ext, 60.0, 240, wt, *sky, True, False, row_range=(3, H + 1))
        with pytest.raises(ValueError, match="background_subtraction"):
            pd.process_image_sharded(image, pos, pointing, fov, W, H, wg, ext, 60.0, 240, wt, *sky, True, True, rank=0, world=2)
    finally:
        ctx.set_option("image_variant", 0)


@pytest.mark.parametrize("variant", [2, 3, 4])
def test_bench_geometry_and_samples_on_voxel_faces_exact(ctx, oracle, variant):
    """The bench scene's geometry at a size the oracle finishes in seconds (512^2 image -> 128^3 grid, one sample per voxel
    pitch, ~50 M samples), and an axis-aligned scene whose samples land EXACTLY on voxel faces (observer on the grid axis,
    unit step, integer extent: the scaled coordinate of the lock-step kernels' screening is an exact integer there and the
    exact division must decide) -- both bit for bit against the oracle in fixed-point mode, counters included."""
    from pixeltovoxelprojector_b200 import process_image as pi
    rng = np.random.default_rng(99)
    ctx.set_option("image_variant", variant)
    try:
        n, H, W = 128, 512, 512
        half = n / 2.0
        scenes = [
            (rng.random((H, W)), [0.0, 0.0, 0.0], [0.02, 0.01, 1.0], 0.8, [(-half, half), (-half, half), (100.0, 100.0 + n)], float(100 + n + 88), 100 + n + 88, (n, n, n)),
            (rng.random((96, 96)), [0.0, 0.0, -10.0], [0.0, 0.0, 1.0], 0.5, [(-32.0, 32.0), (-32.0, 32.0), (0.0, 64.0)], 80.0, 80, (64, 64, 64)),
            (rng.random((96, 96)), [0.5, 0.5, -10.0], [0.0, 0.0, 2.0], 0.5, [(-32.0, 32.0), (-32.0, 32.0), (0.0, 64.0)], 80.0, 160, (64, 32, 128)),
            (rng.random((64, 80)), [-40.0, 0.0, 0.25], [1.0, 0.0, 0.0], 0.7, [(-32.0, 32.0), (-32.0, 32.0), (-32.0, 32.0)], 90.0, 360, (64, 64, 64)),
        ]
        for img, pos, pointing, fov, ext, md, ns, shape in scenes:
            h, w = img.shape
            og, ot = np.zeros(shape, np.int64), np.zeros((8, 8), np.int64)
            nr, nsamp = oracle.process_image(img, pos, pointing, fov, w, h, og, ext, md, ns, ot, 0.0, 0.0, 1.0, 1.0, False, False, accum_mode=1, frac_bits=20)
            gi = torch.zeros(shape, dtype=torch.int64, device="cuda")
            ti = torch.zeros((8, 8), dtype=torch.int64, device="cuda")
            pvp.process_image_cpp(torch.as_tensor(img, device="cuda"), pos, pointing, fov, w, h, gi, ext, md, ns, ti, 0.0, 0.0, 1.0, 1.0, False, False,
                                  accum_mode="fixed", frac_bits=20, want_stats=True)
            assert pi.last_stats == {"n_rays": nr, "n_samples": nsamp} and nsamp > 50 * nr > 0
            assert np.array_equal(gi.cpu().numpy(), og)
    finally:
        ctx.set_option("image_variant", 0)
