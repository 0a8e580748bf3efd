"""Generates tests/golden/sky_golden.npz by executing the REFERENCE's own pre- and post-processing lines of spacevoxelviewer.py on
synthetic inputs:
  * :179-184  image_data = np.nan_to_num(...) ... image = (image_data - image_min) / (image_max - image_min)
  * :320-347  voxel_grid_avg = voxel_grid / len(fits_files_sorted) ... brightest_z = z_coords[brightest_idx]
The lines sit inside functions of a module that needs astropy to import, so the statements are cut out of the unmodified source text
(located by their first and last statement) and executed with numpy only.  Run where /root/reference is mounted:
    python tests/golden/make_golden_sky.py"""
import os
import textwrap

import numpy as np

REF = "/root/reference/spacevoxelviewer.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def block(first, last):
    lines = open(REF).read().splitlines()
    a = next(i for i, l in enumerate(lines) if l.strip() == first)
    b = next(i for i, l in enumerate(lines) if i >= a and l.strip() == last)
    return compile(textwrap.dedent("\n".join(lines[a:b + 1])), REF, "exec")


def main():
    normalize = block("image_data = np.nan_to_num(image_data, nan=0.0, posinf=0.0, neginf=0.0)", "image = (image_data - image_min) / (image_max - image_min)")
    analyze = block("voxel_grid_avg = voxel_grid / len(fits_files_sorted)", "brightest_z = z_coords[brightest_idx]")
    rng = np.random.default_rng(31)
    out = {}
    # ---- normalisation: float32 and float64 images with NaN / +-inf, and a constant image (ValueError)
    n_norm = 0
    for dt in (np.float32, np.float64, np.int16):
        img = (rng.normal(size=(37, 53)) * 100).astype(dt)
        if dt != np.int16:
            img[3, 4], img[5, 6], img[7, 8] = np.nan, np.inf, -np.inf
        ns = {"np": np, "image_data": img.copy()}
        exec(normalize, ns)
        out[f"norm_{n_norm}_in"], out[f"norm_{n_norm}_out"] = img, ns["image"]
        n_norm += 1
    try:
        exec(normalize, {"np": np, "image_data": np.full((4, 4), 2.5)})
        raised = ""
    except ValueError as e:
        raised = str(e)
    out["norm_constant_error"] = np.array(raised)
    out["n_norm"] = np.int64(n_norm)
    # ---- analysis: float64 grids (what spacevoxelviewer allocates), several percentiles, an all-zero grid, ties
    cases = []
    g = np.where(rng.random((14, 11, 9)) < 0.3, rng.gamma(2.0, 3.0, (14, 11, 9)), 0.0)
    cases.append((g, 7, [(-3.0, 5.0), (10.0, 12.5), (-1e3, 1e3)], .1))
    cases.append((g, 7, [(-3.0, 5.0), (10.0, 12.5), (-1e3, 1e3)], 90.0))
    cases.append((g, 3, [(0.1, 0.7), (-0.3, 0.3), (2.0, 2.5)], 99.9))
    cases.append((np.zeros((5, 6, 7)), 4, [(-1.0, 1.0)] * 3, .1))                       # nothing positive: threshold 0, no object voxels
    cases.append((rng.integers(0, 4, (8, 8, 8)).astype(np.float64), 2, [(-8.0, 8.0)] * 3, 50.0))   # heavy ties
    cases.append((rng.random((6, 5, 4)).astype(np.float32), 5, [(0.0, 6.0), (0.0, 5.0), (0.0, 4.0)], 25.0))  # a float32 grid
    for k, (grid, n_files, extent, pct) in enumerate(cases):
        ns = {"np": np, "voxel_grid": grid.copy(), "fits_files_sorted": [None] * n_files, "voxel_grid_extent": extent, "brightness_threshold_percentile": pct}
        exec(analyze, ns)
        out[f"an_{k}_grid"], out[f"an_{k}_n_files"], out[f"an_{k}_extent"], out[f"an_{k}_pct"] = grid, np.int64(n_files), np.array(extent), np.float64(pct)
        out[f"an_{k}_avg"], out[f"an_{k}_threshold"] = ns["voxel_grid_avg"], np.float64(ns["threshold"])
        for name in ("x_coords", "y_coords", "z_coords", "intensities"):
            out[f"an_{k}_{name}"] = ns[name]
        has = ns["intensities"].size > 0
        out[f"an_{k}_brightest"] = np.array([ns["brightest_x"], ns["brightest_y"], ns["brightest_z"]]) if has else np.zeros(0)
    out["n_an"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "sky_golden.npz"), **out)
    print("wrote sky_golden.npz:", n_norm, "normalisation cases,", len(cases), "analysis cases; constant image ->", repr(raised))


if __name__ == "__main__":
    main()
