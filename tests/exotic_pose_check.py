"""K2 against the oracle on exotic geometry (a GPU script for the start of a round; lives under tests/ because it uses the oracle, but is not collected by pytest):
    python tests/exotic_pose_check.py [--cases 400] [--seed 0]
Random small scenes whose camera positions, grid centres and voxel sizes are drawn from raw float32 patterns -- zeros of both
signs, subnormals, values next to the 1e-12 / 1e25 thresholds of the code, huge magnitudes, infinities, NaN -- and whose poses
include axis-aligned views (rays with zero direction components).  Every K2 variant (auto, 1-5) in float32 and fixed point must
reproduce the oracle's restatement of ray_voxel.cpp:484-531: same ray and step counts, same grid (uint8 frames => bit-exact).
The CPU side of the same question -- oracle against the compiled reference on such inputs -- is tests/test_oracle_pin.py::
test_cast_ray_vs_reference_exotic_floats.  Prints one line per mismatch and a summary; exit code 1 on any mismatch."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pixeltovoxelprojector_b200 as pvp  # noqa: E402
from oracle import Oracle  # noqa: E402  (a checker, like tests/)
from pixeltovoxelprojector_b200 import synth  # noqa: E402

SPECIAL = np.array([0.0, -0.0, 1e-12, -1e-12, 9.9e-13, 1e-13, 1.0, -1.0, 0.5, 3.4e38, -3.4e38, 1e25, 9.9e24, 1.1e25, 1e-38, 2.0 ** -126, 2.0 ** -149,
                    np.inf, -np.inf, np.nan, 1e10, -1e10, 16777216.0, 2147483648.0, -2147483904.0], np.float32)


def draw(rng, scale):
    k = rng.integers(0, 4)
    if k == 0:
        return np.float32(rng.choice(SPECIAL))
    if k == 1:
        return np.frombuffer(rng.integers(0, 2 ** 32, dtype=np.uint32).tobytes(), np.float32)[0]     # any bit pattern
    return np.float32(rng.uniform(-scale, scale))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=400)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--oracle-only", action="store_true", help="CPU dry run: build the scenes and run the oracle half only")
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    o = Oracle()
    ctx = None if a.oracle_only else pvp.default_context()
    bad = checked = with_steps = 0
    for case in range(a.cases):
        N = int(rng.integers(1, 40))
        H, W, F = int(rng.integers(1, 9)), int(rng.integers(1, 13)), 3
        exotic_grid = case % 3 == 0
        vs = float(abs(draw(rng, 8.0))) if exotic_grid else float(np.float32(rng.uniform(0.1, 8.0)))
        center = tuple(float(draw(rng, 50.0)) for _ in range(3)) if exotic_grid else tuple(float(np.float32(c)) for c in rng.uniform(-50, 50, 3))
        if not np.isfinite(vs) or vs <= 0:
            vs = 1.0            # the grid descriptor itself must be valid; the exotic part is where the camera sits
        half = 0.5 * N * vs
        poses = []
        for f in range(F):
            cam = [float(draw(rng, 2.0 * half + 1.0)) if rng.random() < 0.5 else float(np.float32(center[k] + rng.uniform(-1.2, 1.2) * half)) if np.isfinite(center[k]) else 0.0
                   for k in range(3)]
            ypr = [float(rng.choice([0.0, 90.0, 180.0, -90.0, 45.0])) if rng.random() < 0.4 else float(np.float32(rng.uniform(-180, 180))) for _ in range(3)]
            poses.append(pvp.FramePose(tuple(cam), yaw=ypr[0], pitch=ypr[1], roll=ypr[2], fov_degrees=float(rng.choice([1.0, 60.0, 120.0, 179.0]))))
        frames = synth.dense_frames(F, H, W, seed=case)
        pairs, n_pairs = pvp.make_pairs(poses, W)
        want = {}
        for accum, mode in (("f32", 0), ("fixed", 1)):
            grid = np.zeros((N, N, N), np.float32 if mode == 0 else np.int64)
            tot = np.zeros(3, np.int64)
            for i in range(1, F):
                R = o.rotation_matrix(poses[i].yaw, poses[i].pitch, poses[i].roll)
                tot += o.accumulate_pair(frames[i - 1], frames[i], np.array(poses[i].camera_position, np.float32), R, o.focal_length(poses[i].fov_degrees, W),
                                         grid, N, np.float32(vs), np.array(center, np.float32), 2.0, accum_mode=mode, frac_bits=8)
            want[accum] = (grid, tot)
        with_steps += want["f32"][1][1] > 0
        for variant in (() if a.oracle_only else (0, 1, 2, 3, 4, 5, 6)):
            ctx.set_option("dda_variant", variant)
            for accum in ("f32", "fixed"):
                g = pvp.VoxelGrid(N, vs, center, accum=accum, frac_bits=8)
                st = pvp.accumulate_motion(torch.as_tensor(frames, device="cuda"), pairs, g, 2.0)
                grid, tot = want[accum]
                got = g.data.cpu().numpy()
                same = np.array_equal(got.view(np.uint32 if accum == "f32" else np.uint64), grid.view(np.uint32 if accum == "f32" else np.uint64))
                checked += 1
                if (st["n_rays"], st["n_steps"]) != (tot[0], tot[1]) or not same:
                    bad += 1
                    print(f"MISMATCH case {case} variant {variant} {accum}: N={N} vs={vs!r} center={center!r} poses={[(p.camera_position, p.yaw, p.pitch, p.roll, p.fov_degrees) for p in poses]} "
                          f"rays/steps {st['n_rays']}/{st['n_steps']} want {tot[0]}/{tot[1]} grid_equal={same}", flush=True)
    if ctx is not None:
        ctx.set_option("dda_variant", 0)
    print(f"exotic_pose_check: {a.cases} scenes ({with_steps} with voxel steps), {checked} kernel runs, {bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    with np.errstate(over="ignore", invalid="ignore"):
        main()
