"""End to end through the CLI boundary (B-2): metadata.json + image folder -> .bin, compared byte-for-value with the
output of the reference CLI itself (tests/golden/cli_golden.npz: noise-off build of the unmodified ray_voxel.cpp)."""
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, bits, golden

import pixeltovoxelprojector_b200 as pvp
from pixeltovoxelprojector_b200 import synth

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "pixeltovoxelprojector_b200", "ray_voxel")


def materialise(tmp_path):
    g = golden("cli_golden.npz")
    for name, present in zip(g["names"], g["present"]):
        if present:
            synth.write_pgm(str(tmp_path / str(name)), g["img_" + str(name)])
    meta = tmp_path / "metadata.json"
    meta.write_text(str(g["metadata"]))
    return g, str(meta)


def check_bin(path, g):
    grid, vs = pvp.read_bin(path)
    assert grid.shape == (500, 500, 500) and vs == g["voxel_size"]
    assert os.path.getsize(path) == 500_000_008                    # 8 + 4 N^3
    flat = grid.reshape(-1)
    nz = np.flatnonzero(flat)
    assert np.array_equal(nz, g["nz_index"])
    assert np.array_equal(bits(flat[nz]), bits(g["nz_value"]))


def test_cli_binary_matches_reference_cli(tmp_path):
    g, meta = materialise(tmp_path)
    out = str(tmp_path / "voxel_grid.bin")
    res = subprocess.run([EXE, meta, str(tmp_path), out], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert res.stdout == f"Saved voxel grid to {out}\n"
    # same diagnostics as the reference: one failed load (+ skip notice) and one size mismatch
    assert res.stderr.count("Skipping frame due to load error.") == str(g["stderr"]).count("Skipping frame due to load error.") == 1
    assert res.stderr.count("Images differ in size") == str(g["stderr"]).count("Images differ in size") == 2
    check_bin(out, g)
    # fixed-point mode writes the same float32 payload for integer frames
    out2 = str(tmp_path / "fx.bin")
    res = subprocess.run([EXE, meta, str(tmp_path), out2, "--fixed-point", "12", "--quiet", "--stats"], capture_output=True, text=True)
    assert res.returncode == 0 and "voxel_steps=" in res.stderr
    check_bin(out2, g)


def test_cli_error_paths(tmp_path):
    res = subprocess.run([EXE], capture_output=True, text=True)
    assert res.returncode == 1 and res.stderr.startswith("Usage:")
    res = subprocess.run([EXE, str(tmp_path / "nope.json"), str(tmp_path), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert res.returncode == 1 and "Cannot open" in res.stderr
    (tmp_path / "empty.json").write_text("[]")
    res = subprocess.run([EXE, str(tmp_path / "empty.json"), str(tmp_path), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert res.returncode == 1 and "No frames loaded." in res.stderr
    g, meta = materialise(tmp_path)
    res = subprocess.run([EXE, meta, str(tmp_path), str(tmp_path / "no_dir" / "o.bin"), "--quiet"], capture_output=True, text=True)
    assert res.returncode == 1 and "Cannot open output file" in res.stderr


def test_python_pipeline_sharded_equals_whole(tmp_path):
    """Frame sharding (SURVEY 8e mode 1) on one GPU: the per-shard grids of 1, 2, 3 and 5 shards sum to the reference grid
    (fixed point => bit-exact), and every shard resolves the skipped-frame pairing like the reference."""
    g, meta = materialise(tmp_path)
    whole = None
    for count in (1, 2, 3, 5):
        total = torch.zeros((500, 500, 500), dtype=torch.int64, device="cuda")
        pairs = 0
        for rank in range(count):
            grid = pvp.VoxelGrid(accum="fixed", frac_bits=8)
            rep = pvp.ray_voxel_accumulate(meta, str(tmp_path), grid, shard=(rank, count))
            total += grid.data
            pairs += rep["n_pairs"]
            del grid
        assert pairs == 4     # cam0: (0,1),(1,3),(3,4); cam3: (0,1); the two size-mismatched pairs cast no rays
        if whole is None:
            whole = total
            flat = pvp.VoxelGrid(accum="fixed", frac_bits=8, data=total).to_float32().cpu().numpy().reshape(-1)
            nz = np.flatnonzero(flat)
            assert np.array_equal(nz, g["nz_index"]) and np.array_equal(bits(flat[nz]), bits(g["nz_value"]))
        else:
            assert torch.equal(whole, total)


def test_decode_pool_does_not_change_results(tmp_path):
    """frames are decoded ahead by a pool of host threads; skip / size-mismatch / pairing semantics and the
    grid must not depend on the number of threads (1 = the old serial loader)."""
    g, meta = materialise(tmp_path)
    ref = None
    for threads in (1, 2, 7):
        grid = pvp.VoxelGrid(accum="fixed", frac_bits=8)
        rep = pvp.ray_voxel_accumulate(meta, str(tmp_path), grid, decode_threads=threads)
        key = (rep["n_pairs"], rep["n_frames_loaded"], rep["n_frames_skipped"], rep["n_size_mismatch"], rep["n_rays"], rep["n_steps"])
        if ref is None:
            ref = (key, grid.data.clone())
            assert rep["n_pairs"] == 4 and rep["n_frames_skipped"] == 1
        else:
            assert key == ref[0] and torch.equal(grid.data, ref[1])
        del grid


def test_cli_loader_noise_matches_seeded_reference(tmp_path):
    """--loader-noise SEED reproduces the reference's own load_image_gray noise (ray_voxel.cpp:203-218: one mt19937 for the
    whole run, one uniform [-1,1) draw per pixel in load order, clamp to [0,255]) when the reference's random_device is
    pinned to the same seed (tests/golden/make_golden_noise.py).  Same visited voxels; float frames with fp32 atomics differ
    from the reference's serial sum only by summation order (rtol 1e-5); fixed point agrees to the quantisation step."""
    g, meta = materialise(tmp_path)
    gn = golden("cli_noise_golden.npz")
    out = str(tmp_path / "noisy.bin")
    res = subprocess.run([EXE, meta, str(tmp_path), out, "--loader-noise", str(int(gn["seed"])), "--quiet"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    grid, _ = pvp.read_bin(out)
    flat = grid.reshape(-1)
    nz = np.flatnonzero(flat)
    assert not np.allclose(gn["nz_value"], g["nz_value"], rtol=1e-4)     # the noise really changes the accumulated values
    assert np.array_equal(nz, gn["nz_index"])                             # identical support
    np.testing.assert_allclose(flat[nz], gn["nz_value"], rtol=1e-5, atol=0)
    # python entry, other seed => different grid; same seed twice => identical (fixed point: bit for bit)
    a = pvp.VoxelGrid(accum="fixed", frac_bits=16)
    b = pvp.VoxelGrid(accum="fixed", frac_bits=16)
    c = pvp.VoxelGrid(accum="fixed", frac_bits=16)
    pvp.ray_voxel_accumulate(meta, str(tmp_path), a, loader_noise_seed=int(gn["seed"]))
    pvp.ray_voxel_accumulate(meta, str(tmp_path), b, loader_noise_seed=int(gn["seed"]), decode_threads=3, chunk_frames=2)
    pvp.ray_voxel_accumulate(meta, str(tmp_path), c, loader_noise_seed=7)
    assert torch.equal(a.data, b.data) and not torch.equal(a.data, c.data)
    fa = a.to_float32().cpu().numpy().reshape(-1)
    assert np.array_equal(np.flatnonzero(fa), gn["nz_index"])
    np.testing.assert_allclose(fa[gn["nz_index"]], gn["nz_value"], rtol=1e-5, atol=2.0 ** -16 * 64)
    with pytest.raises(pvp.PvpError, match="cannot be frame-sharded"):
        pvp.ray_voxel_accumulate(meta, str(tmp_path), c, loader_noise_seed=1, shard=(0, 2))


def test_write_bin_from_device_matches_host_writer(tmp_path):
    """pvp_write_bin_dev (the CLI's writer: chunked D2H overlapping the file writes) produces the bytes of pvp_write_bin
    (ray_voxel.cpp:542-555), for a grid smaller than one chunk and for one that ends in a ragged chunk."""
    import ctypes as C
    ctx = pvp.default_context()
    for n in (1, 37, 170):     # 170^3 * 4 B = 19.7 MB: one full 16 MiB chunk + a tail
        g = torch.randn((n, n, n), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        a, b = str(tmp_path / f"dev_{n}.bin"), str(tmp_path / f"host_{n}.bin")
        ctx.use_current_stream()
        pvp._lib.check(ctx._lib.pvp_write_bin_dev(ctx.handle, os.fsencode(a), n, C.c_float(0.75), g.data_ptr()))
        pvp.write_bin(b, g.cpu().numpy(), 0.75)
        assert open(a, "rb").read() == open(b, "rb").read()
    with pytest.raises(pvp.PvpError, match="Cannot open output file"):
        pvp._lib.check(ctx._lib.pvp_write_bin_dev(ctx.handle, os.fsencode(str(tmp_path / "no_such_dir" / "x.bin")), 1, C.c_float(1.0), g.data_ptr()))

