"""BASELINE configs[0] through the CLI boundary (B-2): the reference's own CPU-runnable case, byte for byte.  Kept in a file of its
own that sorts last, so that it runs after every other parity test."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden

from pixeltovoxelprojector_b200 import synth

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "pixeltovoxelprojector_b200", "ray_voxel")


def test_config1_640x480_default_grid_matches_reference_cli_byte_for_byte(tmp_path):
    """BASELINE configs[0]: the short 640x480 sequence of tests/golden/config1_golden.npz through OUR CLI with the reference's
    three positional arguments only (default grid N = 500, voxel 6, centre (0, 0, 500)): the .bin equals the reference CLI's
    byte for byte (SHA-256 of all 500 000 008 bytes; uint8 frames => every voxel is an exact integer sum)."""
    import hashlib
    g = golden("config1_golden.npz")
    H, W, SQUARE, VALUE = 480, 640, 20, 220          # tests/golden/make_golden_config1.py
    assert g["bg"].shape == (H, W)
    for f, (x, y) in enumerate(g["track"]):
        img = g["bg"].copy()
        img[y:y + SQUARE, x:x + SQUARE] = VALUE
        synth.write_pgm(str(tmp_path / f"frame_{f:03d}.pgm"), img)
    meta = tmp_path / "metadata.json"
    meta.write_text(str(g["metadata"]))
    out = tmp_path / "voxel_grid.bin"
    res = subprocess.run([os.environ.get("PVP_TEST_CLI", EXE), str(meta), str(tmp_path), str(out)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert res.stdout.strip().endswith(f"Saved voxel grid to {out}")
    raw = out.read_bytes()
    assert len(raw) == 500_000_008
    grid = np.frombuffer(raw, np.float32, offset=8)
    assert int(np.count_nonzero(grid)) == int(g["nnz"]) and float(grid.astype(np.float64).sum()) == float(g["total"])
    assert hashlib.sha256(raw).hexdigest() == str(g["sha256"])
