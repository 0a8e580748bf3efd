"""Generates tests/golden/cli_noise_golden.npz: the reference CLI WITH its loader noise (ray_voxel.cpp:203-218), made
reproducible by oracle/_ref/ray_voxel_ref_seeded (the unmodified reference TU; only the token std::random_device is
re-pointed at a device that returns $PVP_REF_NOISE_SEED -- see oracle/ref_harness_ray_voxel.cpp), on the inputs of
cli_golden.npz.  Run where /root/reference is mounted, after `make -C oracle ref`."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SEEDED = os.path.join(ROOT, "oracle", "_ref", "ray_voxel_ref_seeded")
SEED = 12345


def write_pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def main():
    g = np.load(os.path.join(HERE, "cli_golden.npz"), allow_pickle=False)
    with tempfile.TemporaryDirectory() as td:
        for name, present in zip(g["names"], g["present"]):
            if present:
                write_pgm(os.path.join(td, str(name)), g["img_" + str(name)])
        meta = os.path.join(td, "metadata.json")
        open(meta, "w").write(str(g["metadata"]))
        out_bin = os.path.join(td, "out.bin")
        runs = []
        for _ in range(2):
            res = subprocess.run([REF_SEEDED, meta, td, out_bin], capture_output=True, text=True, env={**os.environ, "PVP_REF_NOISE_SEED": str(SEED)})
            assert res.returncode == 0, res.stderr
            runs.append(np.fromfile(out_bin, np.uint8))
        assert np.array_equal(runs[0], runs[1]), "seeded reference is not reproducible"
        grid = runs[0][8:].view(np.float32)
        nz = np.flatnonzero(grid)
        quiet = np.load(os.path.join(HERE, "cli_golden.npz"))["nz_index"]
        np.savez_compressed(os.path.join(HERE, "cli_noise_golden.npz"), seed=np.int64(SEED), nz_index=nz.astype(np.int64), nz_value=grid[nz])
        print("noisy reference:", len(nz), "non-zero voxels (noise-free:", len(quiet), ")")


if __name__ == "__main__":
    sys.exit(main())
