"""Generate the golden fixtures under tests/golden/ from the REAL reference (oracle/_ref, i.e. the
unmodified /root/reference sources compiled by oracle/Makefile).  Run in the container that mounts
/root/reference:

    make -C oracle all ref && OMP_NUM_THREADS=1 python tests/golden/make_golden.py

The fixtures are small .npz files; tests compare both the C restatement (oracle/pvp_oracle.c) and the
CUDA path against them, so parity stays pinned where /root/reference does not exist (the GPU box).
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")

from oracle import REF_CLI, RefLib, ref_process_image_module  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def write_pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def golden_pose(ref, rng):
    ypr = rng.uniform(-180, 180, (64, 3)).astype(np.float32)
    ypr[:4] = [[0, 0, 0], [90, 0, 0], [0, 90, 0], [0, 0, 180]]
    R = np.stack([ref.rotation_matrix(*row) for row in ypr])
    fov = rng.uniform(10, 150, 64).astype(np.float32)
    width = rng.integers(16, 4097, 64).astype(np.int32)
    focal = np.array([ref.focal_length(f, w) for f, w in zip(fov, width)], np.float32)
    np.savez_compressed(os.path.join(OUT, "pose_golden.npz"), ypr=ypr, R=R, fov=fov, width=width, focal=focal)


def golden_rays(ref, rng):
    """cast_ray_into_grid on rays that exercise: inside starts, min-face entry, max-face drop, axis-parallel
    directions (zero / negative-zero components), misses, grazing corners."""
    cases = []
    for k in range(160):
        N = int(rng.integers(2, 48))
        vs = np.float32(rng.choice([0.5, 1.0, 2.0, 6.0, 0.37]))
        center = rng.uniform(-50, 50, 3).astype(np.float32)
        half = N * vs / 2
        cam = (center + rng.uniform(-1.6, 1.6, 3) * half).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        if k % 3:  # aim at a point inside the grid so most outside cameras hit it
            d = ((center + rng.uniform(-0.9, 0.9, 3) * half) - cam).astype(np.float32)
        if k % 5 == 0:
            d[rng.integers(0, 3)] = 0.0
        if k % 7 == 0:
            d[rng.integers(0, 3)] = -0.0
        if k % 9 == 0:  # exactly on a face / corner
            cam = (center + np.array([half, -half, half]) * rng.choice([-1, 1], 3)).astype(np.float32)
        if k % 11 == 0:  # integer-aligned diagonal: ties in t_max
            cam = (center - half + vs * rng.integers(0, N, 3)).astype(np.float32)
            d = rng.choice([-1.0, 1.0], 3).astype(np.float32)
        d = (d / np.linalg.norm(d)).astype(np.float32)
        cases.append((N, vs, center, cam, d))
    vox_all, t_all, counts = [], [], []
    for N, vs, center, cam, d in cases:
        vox, t = ref.cast_ray(cam, d, N, vs, center, want_t=True)
        vox_all.append(vox)
        t_all.append(t)
        counts.append(len(vox))
    np.savez_compressed(os.path.join(OUT, "rays_golden.npz"),
                        N=np.array([c[0] for c in cases], np.int32), voxel_size=np.array([c[1] for c in cases], np.float32),
                        center=np.stack([c[2] for c in cases]), cam=np.stack([c[3] for c in cases]), dir=np.stack([c[4] for c in cases]),
                        counts=np.array(counts, np.int64), vox=np.concatenate(vox_all), t=np.concatenate(t_all))
    print("rays:", len(cases), "rays,", sum(counts), "steps,", sum(1 for c in counts if c == 0), "empty")


def golden_pair(ref, rng):
    """ray_voxel.cpp:484-531 on one small frame pair (uint8 values -> exact fp32 sums)."""
    H, W, N = 36, 48, 40
    vs, center = np.float32(2.5), np.array([3.0, -2.0, 80.0], np.float32)
    prev = rng.integers(0, 256, (H, W), dtype=np.uint8)
    curr = prev.copy()
    curr[8:30, 10:40] = rng.integers(0, 256, (22, 30), dtype=np.uint8)
    curr[0, 0] = prev[0, 0] + 2 if prev[0, 0] < 250 else prev[0, 0] - 2   # |d| == threshold -> not changed
    curr[0, 1] = prev[0, 1] + 3 if prev[0, 1] < 250 else prev[0, 1] - 3   # just above
    ypr = np.array([12.0, -7.0, 5.0], np.float32)
    fov = np.float32(55.0)
    cam = np.array([-10.0, 6.0, 110.0], np.float32)  # inside the grid
    R = ref.rotation_matrix(*ypr)
    focal = ref.focal_length(fov, W)
    grid = np.zeros((N, N, N), np.float32)
    n_rays, n_steps = ref.accumulate_pair(prev.astype(np.float32), curr.astype(np.float32), cam, R, focal, grid, N, vs, center, 2.0)
    changed, diff = ref.detect_motion(prev.astype(np.float32), curr.astype(np.float32), 2.0)
    nz = np.flatnonzero(grid)
    np.savez_compressed(os.path.join(OUT, "pair_golden.npz"), prev=prev, curr=curr, ypr=ypr, fov=fov, cam=cam, R=R, focal=np.float32(focal),
                        N=np.int32(N), voxel_size=vs, center=center, threshold=np.float32(2.0), changed=changed, diff=diff,
                        nz_index=nz.astype(np.int64), nz_value=grid.reshape(-1)[nz], n_rays=np.int64(n_rays), n_steps=np.int64(n_steps))
    print("pair:", n_rays, "rays,", n_steps, "steps,", len(nz), "non-zero voxels")


def golden_cli(rng):
    """The reference CLI itself (noise-off build, ray_voxel.cpp:400-558 incl. its constants N=500, voxel_size 6,
    centre (0,0,500)) on a 2-camera sequence with shuffled metadata, one missing file, one size mismatch and
    a camera with a single frame."""
    H, W = 30, 40
    with tempfile.TemporaryDirectory() as td:
        entries, frames = [], {}
        for cam, nf in ((0, 5), (3, 4), (7, 1)):
            bg = rng.integers(20, 60, (H, W), dtype=np.uint8)
            for f in range(nf):
                img = bg.copy()
                x = 4 + 6 * f
                img[10:16, x:x + 5] = 220
                name = f"c{cam}_f{f}.pgm"
                if cam == 3 and f == 2:
                    img = rng.integers(0, 255, (H + 2, W), dtype=np.uint8)  # size mismatch: no rays, prev advances
                if not (cam == 0 and f == 2):  # cam0 frame2 is missing on disk -> skipped, pair (1,3) forms
                    write_pgm(os.path.join(td, name), img)
                frames[name] = img
                entries.append({"camera_index": cam, "frame_index": 10 * f, "camera_position": [float(20 * cam - 40), 15.0 * f, 300.0 + 10 * f],
                                "yaw": 4.0 * f - 5, "pitch": 2.0 * cam, "roll": 1.5 * f, "fov_degrees": 50.0 + cam, "image_file": name})
        order = rng.permutation(len(entries))
        entries = [entries[i] for i in order]
        del entries[0]["pitch"]  # exercise the defaults of load_metadata
        meta = os.path.join(td, "metadata.json")
        with open(meta, "w") as f:
            json.dump(entries, f)
        out_bin = os.path.join(td, "out.bin")
        res = subprocess.run([REF_CLI, meta, td, out_bin], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        raw = np.fromfile(out_bin, np.uint8)
        n = int(raw[:4].view(np.int32)[0])
        vs = raw[4:8].view(np.float32)[0]
        grid = raw[8:].view(np.float32)
        assert n == 500 and grid.size == n ** 3
        nz = np.flatnonzero(grid)
        names = sorted(frames)
        np.savez_compressed(os.path.join(OUT, "cli_golden.npz"), metadata=json.dumps(entries), names=np.array(names),
                            present=np.array([os.path.exists(os.path.join(td, k)) for k in names]),
                            **{"img_" + k: frames[k] for k in names}, n=np.int32(n), voxel_size=np.float32(vs),
                            nz_index=nz.astype(np.int64), nz_value=grid[nz], stdout=res.stdout, stderr=res.stderr)
        print("cli:", len(nz), "non-zero voxels; stdout:", res.stdout.strip(), "| stderr lines:", len(res.stderr.splitlines()))


def golden_process_image(rng):
    """The reference's own pybind11 module (process_image.cpp, -O2 -fopenmp, OMP_NUM_THREADS=1)."""
    m = ref_process_image_module()
    H, W = 40, 56
    img = rng.random((H, W))
    img[img < 0.25] = 0.0
    cases = {}
    common = dict(earth_position=[0.5, 0.2, -30.0], pointing_direction=[0.05, 0.02, 1.0], fov=0.6, ext=[(-10.0, 10.0), (-12.0, 12.0), (20.0, 50.0)],
                  max_distance=120.0, num_steps=700, sky=(1.52, 1.42, 0.9, 0.7))
    for name, (upd, bg, pointing) in {"update": (True, False, common["pointing_direction"]), "bgsub": (False, True, common["pointing_direction"]),
                                      "zenith": (True, False, [0.0, 0.0, 2.0])}.items():
        grid = np.zeros((24, 30, 20))
        tex = np.zeros((48, 96)) if name != "bgsub" else rng.random((48, 96)) * 0.5
        tex0 = tex.copy()
        m.process_image_cpp(img, common["earth_position"], pointing, common["fov"], W, H, grid, common["ext"], common["max_distance"],
                            common["num_steps"], tex, *common["sky"], upd, bg)
        cases[name] = (pointing, upd, bg, tex0, tex, grid)
        print("process_image", name, "grid sum", grid.sum(), "tex delta", (tex - tex0).sum())
    np.savez_compressed(os.path.join(OUT, "process_image_golden.npz"), image=img, earth_position=np.array(common["earth_position"]), fov=common["fov"],
                        extent=np.array(common["ext"]), max_distance=common["max_distance"], num_steps=common["num_steps"], sky=np.array(common["sky"]),
                        names=np.array(list(cases)),
                        **{f"{k}_{field}": np.asarray(v[i]) for k, v in cases.items()
                           for i, field in enumerate(["pointing", "update", "bgsub", "tex_in", "tex_out", "grid_out"])})


def main():
    ref = RefLib()
    rng = np.random.default_rng(20261019)
    golden_pose(ref, rng)
    golden_rays(ref, rng)
    golden_pair(ref, rng)
    golden_cli(rng)
    golden_process_image(rng)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
