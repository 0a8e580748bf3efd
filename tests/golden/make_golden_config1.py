"""Generates tests/golden/config1_golden.npz: BASELINE.json configs[0] -- the reference CLI on a short synthetic 640x480 sequence with
the default grid (N = 500, voxel 6, centre (0, 0, 500), threshold 2; ray_voxel.cpp:434-447).  The sequence is SURVEY 8d's C1: static
noise background (uint8 20..60) with a moving bright square (value 220), one camera inside the grid whose pose drifts from frame to
frame.  The REFERENCE binary (oracle/_ref/ray_voxel_ref: the unmodified ray_voxel.cpp, loader noise off) writes the .bin; the fixture
keeps what rebuilds the inputs (background, square track, metadata.json text) and the SHA-256 of the 500 000 008-byte output -- uint8
frames make every voxel an exact integer sum, so a bit-exact implementation reproduces the file byte for byte.
Run where /root/reference is mounted:   make -C oracle all ref && python tests/golden/make_golden_config1.py"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import REF_CLI  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
H, W, F, SQUARE, VALUE = 480, 640, 12, 20, 220


def build(folder, bg, track, metadata_text):
    """Materialise the sequence (shared with the tests: tests/test_gpu_cli.py, tests/test_oracle_pin.py)."""
    for f, (x, y) in enumerate(track):
        img = bg.copy()
        img[y:y + SQUARE, x:x + SQUARE] = VALUE
        with open(os.path.join(folder, f"frame_{f:03d}.pgm"), "wb") as fh:
            fh.write(b"P5\n%d %d\n255\n" % (W, H))
            fh.write(img.tobytes())
    meta = os.path.join(folder, "metadata.json")
    with open(meta, "w") as fh:
        fh.write(metadata_text)
    return meta


def main():
    rng = np.random.default_rng(640480)
    bg = rng.integers(20, 61, (H, W), dtype=np.uint8)
    track = np.stack([100 + 37 * np.arange(F), 60 + 23 * np.arange(F)], 1).astype(np.int32)   # the square moves right and down
    entries = []
    for f in range(F):
        entries.append({"camera_index": 0, "frame_index": f, "camera_position": [-200.0 + 3.0 * f, 10.0 - 1.5 * f, 1200.0 + 2.0 * f],
                        "yaw": 2.0 + 0.5 * f, "pitch": -3.0 + 0.25 * f, "roll": 1.0 - 0.125 * f, "fov_degrees": 60.0, "image_file": f"frame_{f:03d}.pgm"})
    metadata_text = json.dumps(entries)
    with tempfile.TemporaryDirectory() as d:
        meta = build(d, bg, track, metadata_text)
        out = os.path.join(d, "voxel_grid.bin")
        r = subprocess.run([REF_CLI, meta, d, out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        raw = open(out, "rb").read()
    assert len(raw) == 8 + 4 * 500 ** 3
    grid = np.frombuffer(raw, np.float32, offset=8)
    nz = np.flatnonzero(grid)
    np.savez_compressed(os.path.join(OUT, "config1_golden.npz"), bg=bg, track=track, metadata=metadata_text, sha256=hashlib.sha256(raw).hexdigest(),
                        nnz=np.int64(nz.size), total=np.float64(grid.astype(np.float64).sum()), vmax=np.float32(grid.max()),
                        first_nz=nz[:64].astype(np.int64), first_values=grid[nz[:64]])
    print("wrote config1_golden.npz: nnz", nz.size, "sum", grid.astype(np.float64).sum(), "max", grid.max(), "sha", hashlib.sha256(raw).hexdigest()[:16])


if __name__ == "__main__":
    main()
