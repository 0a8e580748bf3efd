"""Pin the C restatement (oracle/pvp_oracle.c) to the reference: against the committed golden vectors
(generated from oracle/_ref by tests/golden/make_golden.py) and, where oracle/_ref exists, against the
compiled reference itself on fresh random inputs.  Everything is bit-exact."""
import json
import os

import numpy as np
import pytest

from conftest import bits, golden


def test_pose_golden(oracle):
    g = golden("pose_golden.npz")
    for ypr, R, fov, w, f in zip(g["ypr"], g["R"], g["fov"], g["width"], g["focal"]):
        assert np.array_equal(bits(oracle.rotation_matrix(*ypr)), bits(R))
        assert bits(np.float32(oracle.focal_length(fov, int(w)))) == bits(np.float32(f))


def test_rays_golden(oracle):
    g = golden("rays_golden.npz")
    off = np.concatenate([[0], np.cumsum(g["counts"])])
    for i in range(len(g["N"])):
        vox, t = oracle.cast_ray(g["cam"][i], g["dir"][i], int(g["N"][i]), g["voxel_size"][i], g["center"][i], want_t=True)
        assert len(vox) == g["counts"][i], i
        assert np.array_equal(vox, g["vox"][off[i]:off[i + 1]]), i
        assert np.array_equal(bits(t), bits(g["t"][off[i]:off[i + 1]])), i


def test_pair_golden(oracle):
    g = golden("pair_golden.npz")
    N = int(g["N"])
    prev, curr = g["prev"], g["curr"]
    changed, diff = oracle.detect_motion(prev.astype(np.float32), curr.astype(np.float32), float(g["threshold"]))
    assert np.array_equal(changed, g["changed"]) and np.array_equal(bits(diff), bits(g["diff"]))
    R = oracle.rotation_matrix(*g["ypr"])
    f = oracle.focal_length(g["fov"], prev.shape[1])
    assert np.array_equal(bits(R), bits(g["R"])) and bits(np.float32(f)) == bits(g["focal"])
    for frames in ((prev, curr), (prev.astype(np.float32), curr.astype(np.float32))):
        grid = np.zeros((N, N, N), np.float32)
        nr, ns, nw = oracle.accumulate_pair(frames[0], frames[1], g["cam"], R, f, grid, N, g["voxel_size"], g["center"], float(g["threshold"]))
        assert (nr, ns, nw) == (int(g["n_rays"]), int(g["n_steps"]), int(g["n_steps"]))
        nz = np.flatnonzero(grid)
        assert np.array_equal(nz, g["nz_index"]) and np.array_equal(bits(grid.reshape(-1)[nz]), bits(g["nz_value"]))
    # fixed point: integer inputs => identical values after conversion
    gi = np.zeros((N, N, N), np.int64)
    oracle.accumulate_pair(prev, curr, g["cam"], R, f, gi, N, g["voxel_size"], g["center"], float(g["threshold"]), accum_mode=1, frac_bits=16)
    gf = oracle.fixed_to_f32(gi, 16)
    assert np.array_equal(np.flatnonzero(gf), g["nz_index"]) and np.array_equal(gf.reshape(-1)[g["nz_index"]], g["nz_value"])
    # slabs partition the grid
    parts = []
    for lo, hi in ((0, 13), (13, 14), (14, N)):
        s = np.zeros((hi - lo, N, N), np.float32)
        oracle.accumulate_pair(prev, curr, g["cam"], R, f, s, N, g["voxel_size"], g["center"], float(g["threshold"]), slab=(lo, hi))
        parts.append(s)
    assert np.array_equal(np.concatenate(parts).reshape(-1)[g["nz_index"]], g["nz_value"])


def test_process_image_golden(oracle):
    g = golden("process_image_golden.npz")
    img = g["image"]
    H, W = img.shape
    for name in g["names"]:
        grid = np.zeros_like(g[f"{name}_grid_out"])
        tex = g[f"{name}_tex_in"].copy()
        oracle.process_image(img, g["earth_position"], g[f"{name}_pointing"], float(g["fov"]), W, H, grid, g["extent"], float(g["max_distance"]),
                             int(g["num_steps"]), tex, *g["sky"], bool(g[f"{name}_update"]), bool(g[f"{name}_bgsub"]))
        assert np.array_equal(bits(grid), bits(g[f"{name}_grid_out"])), name
        assert np.array_equal(bits(tex), bits(g[f"{name}_tex_out"])), name


def test_process_image_strided_and_empty_grid(oracle):
    g = golden("process_image_golden.npz")
    img = g["image"]
    H, W = img.shape
    name = "update"
    want = g[f"{name}_grid_out"]
    big = np.zeros((want.shape[0], want.shape[1] * 2, want.shape[2]))
    view = big[:, ::2, :]
    tex = np.asfortranarray(g[f"{name}_tex_in"].copy())
    oracle.process_image(img, g["earth_position"], g[f"{name}_pointing"], float(g["fov"]), W, H, view, g["extent"], float(g["max_distance"]),
                         int(g["num_steps"]), tex, *g["sky"], True, False)
    assert np.array_equal(view, want) and np.array_equal(tex, g[f"{name}_tex_out"]) and not big[:, 1::2, :].any()
    tex2 = g[f"{name}_tex_in"].copy()
    nr, ns = oracle.process_image(img, g["earth_position"], g[f"{name}_pointing"], float(g["fov"]), W, H, np.zeros((0,)), g["extent"],
                                  float(g["max_distance"]), int(g["num_steps"]), tex2, *g["sky"], True, False)
    assert ns == 0 and np.array_equal(tex2, g[f"{name}_tex_out"])


# ---- against the compiled reference itself (only where oracle/_ref exists) -------------------
def test_pose_vs_reference(oracle, reflib):
    rng = np.random.default_rng(1)
    for _ in range(1500):
        ypr = rng.uniform(-360, 360, 3).astype(np.float32)
        assert np.array_equal(bits(oracle.rotation_matrix(*ypr)), bits(reflib.rotation_matrix(*ypr)))
        fov, w = np.float32(rng.uniform(1, 179)), int(rng.integers(1, 8192))
        assert bits(np.float32(oracle.focal_length(fov, w))) == bits(np.float32(reflib.focal_length(fov, w)))


def test_pixel_rays_vs_reference(oracle, reflib):
    rng = np.random.default_rng(2)
    for _ in range(200):
        W, H = int(rng.integers(2, 4000)), int(rng.integers(2, 3000))
        R = reflib.rotation_matrix(*rng.uniform(-180, 180, 3).astype(np.float32))
        f = reflib.focal_length(np.float32(rng.uniform(20, 120)), W)
        for _ in range(10):
            u, v = int(rng.integers(0, W)), int(rng.integers(0, H))
            assert np.array_equal(bits(oracle.pixel_ray(u, v, W, H, f, R)), bits(reflib.pixel_ray(u, v, W, H, f, R)))


def test_cast_ray_vs_reference(oracle, reflib):
    rng = np.random.default_rng(3)
    steps = 0
    for k in range(2500):
        N = int(rng.integers(1, 160))
        vs = np.float32(rng.uniform(0.05, 12))
        center = rng.uniform(-200, 200, 3).astype(np.float32)
        half = N * vs / 2
        cam = (center + rng.uniform(-1.8, 1.8, 3) * half).astype(np.float32)
        d = ((center + rng.uniform(-1, 1, 3) * half) - cam).astype(np.float32) if k % 2 else rng.normal(size=3).astype(np.float32)
        if k % 7 == 0:
            d[rng.integers(0, 3)] = rng.choice([0.0, -0.0, 1e-13])
        if k % 13 == 0:
            cam = (center - half + vs * rng.integers(0, N + 1, 3)).astype(np.float32)
            d = rng.choice([-1.0, 1.0, 0.0], 3).astype(np.float32)
        nrm = np.linalg.norm(d)
        if nrm > 0:
            d = (d / nrm).astype(np.float32)
        a, ta = oracle.cast_ray(cam, d, N, vs, center, want_t=True)
        b, tb = reflib.cast_ray(cam, d, N, vs, center, want_t=True)
        assert a.shape == b.shape and np.array_equal(a, b) and np.array_equal(bits(ta), bits(tb)), k
        steps += len(a)
    assert steps > 20000


@pytest.mark.parametrize("dense", [False, True])
def test_accumulate_pair_vs_reference(oracle, reflib, dense):
    rng = np.random.default_rng(4 + dense)
    H, W, N = 54, 72, 72
    vs, center = np.float32(3.0), np.array([5, -3, 400], np.float32)
    prev = rng.integers(0, 256, (H, W)).astype(np.float32)
    curr = rng.integers(0, 256, (H, W)).astype(np.float32) if dense else prev.copy()
    if not dense:
        curr[20:30, 30:44] += 50
    prev += rng.uniform(-1, 1, prev.shape).astype(np.float32)  # float frames with (seeded) loader-style noise
    curr += rng.uniform(-1, 1, curr.shape).astype(np.float32)
    R = reflib.rotation_matrix(20, -10, 170)   # roll 170 -> looks towards +z
    f = reflib.focal_length(70, W)
    cam = np.array([10, 20, 250], np.float32)  # outside, below the min-z face
    g1, g2 = np.zeros((N, N, N), np.float32), np.zeros((N, N, N), np.float32)
    s1 = oracle.accumulate_pair(prev, curr, cam, R, f, g1, N, vs, center, 2.0)
    s2 = reflib.accumulate_pair(prev, curr, cam, R, f, g2, N, vs, center, 2.0)
    assert s1[:2] == s2 and s1[1] > 0
    assert np.array_equal(bits(g1), bits(g2))


def test_process_image_vs_reference(oracle):
    from oracle import REF_DIR, ref_process_image_module
    if not os.path.isdir(REF_DIR):
        pytest.skip("oracle/_ref not built")
    m = ref_process_image_module()
    rng = np.random.default_rng(5)
    for trial in range(4):
        H, W = int(rng.integers(8, 40)), int(rng.integers(8, 40))
        img = rng.random((H, W))
        img[img < 0.3] = 0
        pos = rng.uniform(-5, 5, 3)
        pointing = rng.normal(size=3) if trial else np.array([0.0, 0.0, -3.0])
        ext = [(-8.0, 9.0), (-7.0, 6.5), (-6.0, 8.0)] if trial % 2 else [(10.0, 30.0), (-10.0, 10.0), (-10.0, 10.0)]
        shape = (int(rng.integers(4, 20)), int(rng.integers(4, 20)), int(rng.integers(4, 20)))
        sky = (float(rng.uniform(0, 6.28)), float(rng.uniform(-1.5, 1.5)), float(rng.uniform(0.5, 6.0)), float(rng.uniform(0.5, 3.0)))
        for upd, bg in ((True, False), (False, True)):
            g1, g2 = np.zeros(shape), np.zeros(shape)
            t1 = rng.random((16, 24)) * 0.4
            t2 = t1.copy()
            m.process_image_cpp(img, list(pos), list(pointing), 0.9, W, H, g1, ext, 60.0, 300, t1, *sky, upd, bg)
            oracle.process_image(img, pos, pointing, 0.9, W, H, g2, ext, 60.0, 300, t2, *sky, upd, bg)
            assert np.array_equal(bits(g1), bits(g2)) and np.array_equal(bits(t1), bits(t2)), (trial, upd, bg)


def test_cast_ray_vs_reference_exotic_floats(oracle, reflib):
    """hypothesis over raw float32 bit patterns (zeros of both signs, subnormals, huge values, infinities, NaN) for camera,
    direction, voxel size and centre: the restatement and the compiled reference must visit the same voxels with the same t
    bits, or both nothing -- the regimes where fminf/fmaxf, the 1e-12 tests and int() of NaN / out-of-range values decide."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    f32 = st.one_of(st.floats(width=32, allow_nan=True, allow_infinity=True, allow_subnormal=True),
                    st.sampled_from([0.0, -0.0, 1e-12, -1e-12, 9.9e-13, 1e-13, 1.0, -1.0, 0.5, 3.4e38, -3.4e38, 1e25, 1e-38, 2.0 ** -126, 2.0 ** -149]),
                    st.floats(min_value=-600.0, max_value=600.0, width=32))
    vec = st.tuples(f32, f32, f32)
    n_checked = [0, 0]

    @hyp.settings(max_examples=1500, deadline=None, derandomize=True, suppress_health_check=list(hyp.HealthCheck))
    @hyp.given(cam=vec, d=vec, n=st.integers(1, 48), vs=f32, center=vec)
    def run(cam, d, n, vs, center):
        cam, d, center = (np.array(v, np.float32) for v in (cam, d, center))
        vs = np.float32(vs)
        a, ta = oracle.cast_ray(cam, d, n, vs, center, cap=4 * n + 8, want_t=True)
        b, tb = reflib.cast_ray(cam, d, n, vs, center, cap=4 * n + 8, want_t=True)
        assert a.shape == b.shape and np.array_equal(a, b) and np.array_equal(bits(ta), bits(tb)), (cam, d, n, vs, center)
        n_checked[0] += 1
        n_checked[1] += len(a) > 0

    run()
    assert n_checked[0] >= 1000 and n_checked[1] > 50      # a good share of the exotic inputs still produce steps


def test_process_image_vs_reference_property(oracle):
    """hypothesis over path B's geometry: observer inside / outside / on a face of the grid, pointing along the axes (the `up`
    switch of process_image.cpp:203-211), tiny and huge fields of view, one march step, textures of one cell, images with zero,
    negative and huge brightness, thin grids.  The restatement and the compiled reference module must agree bit for bit."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    from oracle import REF_DIR, ref_process_image_module
    if not os.path.isdir(REF_DIR):
        pytest.skip("oracle/_ref not built")
    m = ref_process_image_module()
    coord = st.one_of(st.sampled_from([0.0, -0.0, 1.0, -1.0, 5.0, -5.0, 10.0, 1e-9, -1e-9, 1e6, -1e6]), st.floats(-30.0, 30.0, allow_nan=False, width=64))
    direction = st.one_of(st.sampled_from([(0.0, 0.0, 1.0), (0.0, 0.0, -2.0), (1e-9, 0.0, 1.0), (1.0, 0.0, 0.0), (0.0, -3.0, 0.0), (1.0, 1.0, 1.0), (1e-8, 1e-8, -1.0)]),
                          st.tuples(coord, coord, coord).filter(lambda v: any(abs(c) > 1e-6 for c in v)))
    lohi = st.tuples(st.floats(-20.0, 19.0, width=64), st.floats(0.015625, 25.0, width=64)).map(lambda p: (p[0], p[0] + p[1]))
    pixel = st.one_of(st.sampled_from([0.0, -0.0, -1.0, 1.0, 1e-300, 1e300, 0.5]), st.floats(0.0, 1.0, width=64))
    n_calls = [0, 0]

    @hyp.settings(max_examples=800, deadline=None, derandomize=True, suppress_health_check=list(hyp.HealthCheck))
    @hyp.given(data=st.data(), h=st.integers(1, 5), w=st.integers(1, 5), pos=st.tuples(coord, coord, coord), pointing=direction,
               fov=st.sampled_from([1e-6, 0.01, 0.5, 0.9, 1.5, 3.0, 3.1415]), shape=st.tuples(st.integers(1, 5), st.integers(1, 5), st.integers(1, 5)),
               ext=st.tuples(lohi, lohi, lohi), max_distance=st.sampled_from([0.5, 7.0, 60.0, 1e3]), num_steps=st.sampled_from([1, 2, 17, 64]),
               tex=st.tuples(st.integers(1, 4), st.integers(1, 4)), sky=st.tuples(st.floats(0.0, 6.28), st.floats(-1.5, 1.5), st.floats(0.01, 6.0), st.floats(0.01, 3.0)),
               flags=st.sampled_from([(True, False), (False, True), (False, False), (True, True)]), with_grid=st.booleans(), aim=st.integers(0, 3))
    def run(data, h, w, pos, pointing, fov, shape, ext, max_distance, num_steps, tex, sky, flags, with_grid, aim):
        if aim:      # three times out of four look at the grid (from wherever the observer is), so that rays do cross it
            to_centre = [0.5 * (lo + hi) - p for (lo, hi), p in zip(ext, pos)]
            if any(abs(c) > 1e-6 for c in to_centre):
                pointing = tuple(to_centre)
        img = np.array(data.draw(st.lists(pixel, min_size=h * w, max_size=h * w)), np.float64).reshape(h, w)
        g1 = np.zeros(shape if with_grid else (0,))
        g2 = g1.copy()
        t1 = np.full(tex, 0.25)
        t2 = t1.copy()
        upd, bg = flags
        # bg-subtraction while updating reads the texture another pixel of the same call may have written: serial in both (1 thread)
        m.process_image_cpp(img, list(pos), list(pointing), fov, w, h, g1, [tuple(e) for e in ext], max_distance, num_steps, t1, *sky, upd, bg)
        oracle.process_image(img, pos, pointing, fov, w, h, g2, [tuple(e) for e in ext], max_distance, num_steps, t2, *sky, upd, bg)
        assert np.array_equal(bits(g1), bits(g2)) and np.array_equal(bits(t1), bits(t2)), (pos, pointing, fov, shape, ext, max_distance, num_steps, tex, sky, flags)
        n_calls[0] += 1
        n_calls[1] += bool(g1.any())

    os.environ.setdefault("OMP_NUM_THREADS", "1")
    run()
    assert n_calls[0] >= 500 and n_calls[1] >= 40      # calls in which rays did cross the grid


def test_config1_oracle_reproduces_the_reference_cli_output(oracle, tmp_path):
    """BASELINE configs[0] (640x480 short sequence, default grid): the restated frame loop, fed the fixture's frames and poses,
    rebuilds the reference CLI's 500 000 008-byte .bin byte for byte (tests/golden/config1_golden.npz holds its SHA-256)."""
    import hashlib
    import json
    import struct
    from golden.make_golden_config1 import H, W, SQUARE, VALUE
    g = golden("config1_golden.npz")
    entries = json.loads(str(g["metadata"]))
    N, vs, center = 500, np.float32(6.0), np.array([0.0, 0.0, 500.0], np.float32)      # ray_voxel.cpp:434-437
    grid = np.zeros((N, N, N), np.float32)
    prev = None
    for e, (x, y) in zip(entries, g["track"]):
        img = g["bg"].copy()
        img[y:y + SQUARE, x:x + SQUARE] = VALUE
        if prev is not None:
            R = oracle.rotation_matrix(e["yaw"], e["pitch"], e["roll"])
            oracle.accumulate_pair(prev, img, np.array(e["camera_position"], np.float32), R, oracle.focal_length(e["fov_degrees"], W), grid, N, vs, center, 2.0)
        prev = img
    assert img.shape == (H, W)
    sha = hashlib.sha256(struct.pack("<if", N, 6.0))
    sha.update(grid.tobytes())
    assert int(np.count_nonzero(grid)) == int(g["nnz"]) and float(grid.astype(np.float64).sum()) == float(g["total"])
    assert sha.hexdigest() == str(g["sha256"])


def test_accumulate_pair_vs_reference_exotic_poses(oracle, reflib):
    """The whole per-pair loop (ray_voxel.cpp:484-531: detect_motion, ray set-up, cast_ray_into_grid, accumulate) on small
    frames with camera positions, grid centres, voxel sizes, focal lengths and rotation entries drawn from raw float32
    patterns (NaN, infinities, subnormals, the neighbourhood of 1e-12): the restatement and the compiled reference loop
    must produce the same ray / step counts and the same grid bits.  tests/exotic_pose_check.py asks the CUDA kernels the
    same question against the oracle."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    f32 = st.one_of(st.floats(width=32, allow_nan=True, allow_infinity=True, allow_subnormal=True),
                    st.sampled_from([0.0, -0.0, 1e-12, -1e-12, 9.9e-13, 1.0, -1.0, 3.4e38, 1e25, 2.0 ** -126, 16777216.0, 2147483648.0]),
                    st.floats(min_value=-80.0, max_value=80.0, width=32))
    tame = st.floats(min_value=-60.0, max_value=60.0, width=32)
    coord = st.one_of(f32, tame, tame)
    seen = [0, 0]

    @hyp.settings(max_examples=600, deadline=None, derandomize=True, suppress_health_check=list(hyp.HealthCheck))
    @hyp.given(n=st.integers(1, 24), h=st.integers(1, 6), w=st.integers(1, 8), seed=st.integers(0, 10 ** 6), cam=st.tuples(coord, coord, coord),
               center=st.tuples(coord, coord, coord), vs=st.one_of(f32, st.floats(min_value=0.25, max_value=8.0, width=32)),
               ypr=st.tuples(st.sampled_from([0.0, 90.0, 180.0, -90.0, 33.0]), st.sampled_from([0.0, 90.0, -45.0, 10.0]), st.sampled_from([0.0, 180.0, 170.0, -20.0])),
               fov=st.sampled_from([1.0, 60.0, 120.0, 179.0, 180.0, 0.0]), poke=st.one_of(st.none(), st.tuples(st.integers(0, 8), f32)))
    def run(n, h, w, seed, cam, center, vs, ypr, fov, poke):
        rng = np.random.default_rng(seed)
        prev = rng.integers(0, 256, (h, w)).astype(np.float32)
        curr = rng.integers(0, 256, (h, w)).astype(np.float32)
        R = reflib.rotation_matrix(*ypr).copy()
        if poke is not None:
            R[poke[0]] = np.float32(poke[1])          # an arbitrary float in the rotation (the loop takes R as data)
        f = reflib.focal_length(fov, w)                # fov 0 / 180 give an infinite / ~zero focal length
        cam, center, vs = np.array(cam, np.float32), np.array(center, np.float32), np.float32(vs)
        g1, g2 = np.zeros((n, n, n), np.float32), np.zeros((n, n, n), np.float32)
        s1 = oracle.accumulate_pair(prev, curr, cam, R, f, g1, n, vs, center, 2.0)
        s2 = reflib.accumulate_pair(prev, curr, cam, R, f, g2, n, vs, center, 2.0)
        assert s1[:2] == s2 and np.array_equal(bits(g1), bits(g2)), (n, h, w, seed, cam, center, vs, ypr, fov, poke)
        seen[0] += 1
        seen[1] += s2[1] > 0

    run()
    assert seen[0] >= 400 and seen[1] >= 40
