"""Smallest end-to-end case touching every kernel (for compute-sanitizer): K1 (row-major + row-group), K2 (every variant, whole
grid + slab, f32 + fixed), K3 (generic, one-pixel, two-pixel), the analysis kernels and the fixed-point converters."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pixeltovoxelprojector_b200 as pvp  # noqa: E402
from pixeltovoxelprojector_b200 import analysis, synth  # noqa: E402

ctx = pvp.default_context()
F, H, W, N, vs, center = 3, 20, 37, 24, 4.0, (0.0, 0.0, 500.0)
frames = torch.as_tensor(synth.dense_frames(F, H, W, seed=1), device="cuda")
poses = synth.orbit_poses(F, N, vs, center, seed=2)
pairs, _ = pvp.make_pairs(poses, W)
for variant in (0, 1, 2, 3, 4, 5, 6):
    ctx.set_option("dda_variant", variant)
    for accum in ("f32", "fixed"):
        for slab in (None, (5, 17)):
            g = pvp.VoxelGrid(N, vs, center, accum=accum, frac_bits=10, slab=slab)
            st = pvp.accumulate_motion(frames, pairs, g, 2.0)
            assert st["n_steps"] > 0
ctx.set_option("dda_variant", 0)
pvp.accumulate_motion(frames.float(), pairs, pvp.VoxelGrid(N, vs, center), 2.0)
pvp.accumulate_motion(frames.cpu().numpy(), pairs, pvp.VoxelGrid(N, vs, center), 2.0)
rng = np.random.default_rng(3)
img = torch.as_tensor(rng.random((19, 45)), device="cuda")
ext = [(-10.0, 10.0), (-10.0, 10.0), (20.0, 40.0)]
for variant in (1, 2, 3, 4):
    ctx.set_option("image_variant", variant)
    for mode, dt in (("f64", torch.float64), ("fixed", torch.int64)):
        g = torch.zeros((12, 14, 16), dtype=dt, device="cuda")
        t = torch.zeros((8, 8), dtype=dt, device="cuda")
        pvp.process_image_cpp(img, [0.1, 0.2, -5.0], [0.02, 0.01, 1.0], 0.7, 45, 19, g, ext, 80.0, 200, t, 1.5, 1.5, 0.8, 0.8, True, False, accum_mode=mode, frac_bits=20)
ctx.set_option("image_variant", 0)
grid = pvp.VoxelGrid(N, vs, center)
pvp.accumulate_motion(frames, pairs, grid, 2.0)
pts, inten = analysis.extract_top_percentile_z_up(grid.data, np.float32(vs), np.array(center), percentile=90.0)
fx = pvp.VoxelGrid(N, vs, center, accum="fixed", frac_bits=10)
pvp.accumulate_motion(frames, pairs, fx, 2.0)
fx.to_float32()
torch.cuda.synchronize()
print("sanitize_case ok", len(inten))
