"""Generates tests/golden/viewer_golden.npz by running the REFERENCE's own functions (voxelmotionviewer.py:28-96,
`load_voxel_grid` and `extract_top_percentile_z_up`) on synthetic .bin files.  The module cannot be imported (pyvista is
not installed), so the two function definitions are cut out of the unmodified source with `ast` and executed with numpy
only.  Run where /root/reference is mounted:  python tests/golden/make_golden_viewer.py"""
import ast
import os
import struct
import sys
import tempfile

import numpy as np

REF = "/root/reference/voxelmotionviewer.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_functions():
    src = open(REF).read()
    tree = ast.parse(src)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("load_voxel_grid", "extract_top_percentile_z_up"):
            exec(compile(ast.Module([node], []), REF, "exec"), ns)
    return ns["load_voxel_grid"], ns["extract_top_percentile_z_up"]


def main():
    load_voxel_grid, extract = reference_functions()
    rng = np.random.default_rng(2024)
    out = {}
    cases = {
        # integer-valued sums, mostly zeros (what the motion pipeline writes for uint8 frames)
        "sparse": (40, 6.0, np.where(rng.random((40, 40, 40)) < 0.03, rng.integers(3, 2000, (40, 40, 40)), 0).astype(np.float32)),
        # dense float data, no ties
        "dense": (33, 1.5, rng.gamma(2.0, 50.0, (33, 33, 33)).astype(np.float32)),
        # all zeros: nothing above the threshold
        "empty": (16, 2.0, np.zeros((16, 16, 16), np.float32)),
        # voxel size that is not a float32 number (N * voxel_size must round in float32, :68) and cells EQUAL to float32(hard_thresh) for a
        # hard threshold that is not one either (`grid > 0.3` compares in float32: float32(0.3) > 0.3 is False there, True in double)
        "inexact": (25, 0.1, rng.choice(np.array([0.0, 0.1, 0.2, 0.3, 0.4, 0.7], np.float32), (25, 25, 25))),
    }
    params = {"sparse": [dict(percentile=99.5), dict(percentile=97.0), dict(use_hard_thresh=True, hard_thresh=700)],
              "dense": [dict(percentile=99.5), dict(percentile=12.5), dict(percentile=100.0), dict(percentile=0.0)],
              "empty": [dict(percentile=99.5)],
              "inexact": [dict(use_hard_thresh=True, hard_thresh=0.3), dict(percentile=60.0), dict(use_hard_thresh=True, hard_thresh=np.float64(0.3))]}
    centers = {"sparse": np.array([0, 0, 500]), "dense": np.array([10.5, -3.25, 7.0]), "empty": np.array([0, 0, 0]), "inexact": np.array([0.3, -1.7, 12.9])}
    for name, (n, vs, grid) in cases.items():
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "g.bin")
            with open(path, "wb") as f:
                f.write(struct.pack("<if", n, vs))
                f.write(grid.tobytes())
            g, voxel_size = load_voxel_grid(path)
            assert np.array_equal(g, grid)
            out[f"{name}_grid"] = grid
            out[f"{name}_voxel_size"] = np.float32(voxel_size)
            out[f"{name}_center"] = centers[name]
            for i, kw in enumerate(params[name]):
                pts, inten = extract(g, voxel_size, centers[name], **kw)
                out[f"{name}_{i}_percentile"] = np.float64(kw.get("percentile", 99.5))
                out[f"{name}_{i}_hard"] = np.float64(kw.get("hard_thresh", 700) if kw.get("use_hard_thresh") else -1)
                out[f"{name}_{i}_hard_is_np64"] = np.bool_(isinstance(kw.get("hard_thresh"), np.float64))
                out[f"{name}_{i}_points"] = np.zeros((0, 3)) if pts is None else pts
                out[f"{name}_{i}_intensities"] = np.zeros((0,), np.float32) if inten is None else inten
                out[f"{name}_{i}_none"] = np.bool_(pts is None)
        out[f"{name}_ncases"] = np.int64(len(params[name]))
    np.savez_compressed(os.path.join(HERE, "viewer_golden.npz"), **out)
    print("wrote viewer_golden.npz:", {k: v.shape for k, v in out.items() if k.endswith("points")})


if __name__ == "__main__":
    sys.exit(main())
