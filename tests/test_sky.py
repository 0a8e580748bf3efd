"""the numpy steps around process_image_cpp in spacevoxelviewer.py (:179-184 and :317-347) as tensor ops.
The reference evaluates them inline (inside functions that need astropy), so the recipe is re-evaluated here with numpy,
line for line, and compared value for value.  Runs on the CPU (torch CPU tensors) and, with -m gpu, on the device."""
import numpy as np
import pytest
import torch

from conftest import bits

from pixeltovoxelprojector_b200 import sky


def numpy_normalize(image_data):                      # spacevoxelviewer.py:179-184
    image_data = np.nan_to_num(image_data, nan=0.0, posinf=0.0, neginf=0.0)
    image_min = np.min(image_data)
    image_max = np.max(image_data)
    if image_max - image_min == 0:
        raise ValueError("Image data has zero dynamic range.")
    return (image_data - image_min) / (image_max - image_min)


def numpy_analyze(voxel_grid, n_files, voxel_grid_extent, brightness_threshold_percentile):   # spacevoxelviewer.py:317-347
    voxel_grid_avg = voxel_grid / n_files
    if np.any(voxel_grid_avg > 0):
        threshold = np.percentile(voxel_grid_avg[voxel_grid_avg > 0], brightness_threshold_percentile)
    else:
        threshold = 0
    object_voxels = voxel_grid_avg > threshold
    x_indices, y_indices, z_indices = np.nonzero(object_voxels)
    nx, ny, nz = voxel_grid_avg.shape
    x_min, x_max = voxel_grid_extent[0]
    y_min, y_max = voxel_grid_extent[1]
    z_min, z_max = voxel_grid_extent[2]
    x_coords = x_indices / nx * (x_max - x_min) + x_min
    y_coords = y_indices / ny * (y_max - y_min) + y_min
    z_coords = z_indices / nz * (z_max - z_min) + z_min
    intensities = voxel_grid_avg[object_voxels]
    brightest = None
    if intensities.size > 0:
        b = np.argmax(intensities)
        brightest = (x_coords[b], y_coords[b], z_coords[b])
    return threshold, x_coords, y_coords, z_coords, intensities, brightest


def _check(device):
    rng = np.random.default_rng(17)
    for dt in (np.float32, np.float64):
        img = (rng.normal(size=(37, 53)) * 100).astype(dt)
        img[3, 4], img[5, 6], img[7, 8] = np.nan, np.inf, -np.inf
        got = sky.normalize_image(torch.as_tensor(img, device=device))
        want = numpy_normalize(img)
        assert got.dtype == {np.float32: torch.float32, np.float64: torch.float64}[dt]
        assert np.array_equal(bits(got.cpu().numpy()), bits(want))
    with pytest.raises(ValueError, match="zero dynamic range"):
        sky.normalize_image(torch.full((4, 4), 2.5, device=device))
    with pytest.raises(ValueError, match="not 2D"):
        sky.normalize_image(torch.zeros((2, 3, 4), device=device))
    ext = ((-3e12 + 1.1e12, 3e12 + 1.1e12), (-3e12 - 5.7e12, 3e12 - 5.7e12), (-3e12 - 2.2e12, 3e12 - 2.2e12))
    for trial, pct in enumerate((0.1, 50.0, 99.0, 0.0, 100.0)):
        grid = np.where(rng.random((21, 17, 29)) < 0.2, rng.gamma(2.0, 3.0, (21, 17, 29)), 0.0)
        if trial == 3:
            grid[:] = 0.0                                  # nothing positive: threshold 0, no object voxels
        thr, xc, yc, zc, inten, brightest = numpy_analyze(grid, 7, ext, pct)
        out = sky.analyze_voxel_grid(torch.as_tensor(grid, device=device), 7, ext, pct)
        assert float(out["threshold"]) == float(thr), (trial, out["threshold"], thr)
        assert np.array_equal(bits(out["x_coords"].cpu().numpy()), bits(xc)) and np.array_equal(bits(out["y_coords"].cpu().numpy()), bits(yc))
        assert np.array_equal(bits(out["z_coords"].cpu().numpy()), bits(zc))
        assert np.array_equal(bits(out["intensities"].cpu().numpy()), bits(inten))
        assert (out["brightest"] is None) == (brightest is None)
        if brightest is not None:
            assert out["brightest"] == tuple(float(v) for v in brightest)


def test_sky_pre_and_post_processing_cpu():
    _check("cpu")


@pytest.mark.gpu
def test_sky_pre_and_post_processing_gpu():
    _check("cuda")


@pytest.mark.gpu
def test_fits_loop_stays_on_device(ctx):
    """The loop of spacevoxelviewer.py:300-320 with everything resident: normalise -> process_image_cpp -> analyse."""
    import pixeltovoxelprojector_b200 as pvp
    rng = np.random.default_rng(3)
    n = 48
    grid = torch.zeros((n, n, n), dtype=torch.float64, device="cuda")
    tex = torch.zeros((32, 32), dtype=torch.float64, device="cuda")
    ext = [(-10.0, 10.0), (-10.0, 10.0), (20.0, 40.0)]
    files = 3
    for f in range(files):
        raw = torch.as_tensor((rng.gamma(2.0, 50.0, (40, 56))).astype(np.float32), device="cuda")
        image = sky.normalize_image(raw).to(torch.float64)                     # :184, :234
        pvp.process_image_cpp(image, [0.1 * f, 0.2, -5.0], [0.02, 0.01, 1.0], 0.7, 56, 40, grid, ext, 80.0, 400, tex, 1.5, 1.5, 0.8, 0.8, True, False)
    out = sky.analyze_voxel_grid(grid, files, ext, 0.1)
    thr, xc, yc, zc, inten, brightest = numpy_analyze(grid.cpu().numpy(), files, ext, 0.1)
    assert float(out["threshold"]) == float(thr) and len(inten) > 100
    assert np.array_equal(bits(out["intensities"].cpu().numpy()), bits(inten)) and out["brightest"] == tuple(float(v) for v in brightest)


def test_field_of_view_matches_the_reference_lines():
    """tests/golden/fov_golden.npz: the reference's own statements (spacevoxelviewer.py:197-215) executed on synthetic headers."""
    import json
    import os
    from conftest import ROOT
    import pixeltovoxelprojector_b200 as pvp
    g = np.load(os.path.join(ROOT, "tests", "golden", "fov_golden.npz"))
    cases = json.loads(str(g["cases"]))
    assert len(cases) >= 20
    for c, want in zip(cases, g["fov_rad"]):
        got = pvp.field_of_view_rad(c["header"], c["width"], c["height"])
        assert np.float64(got).tobytes() == np.float64(want).tobytes(), c


def _golden_check(device):
    """tests/golden/sky_golden.npz: outputs of the reference's OWN statements (spacevoxelviewer.py:179-184 and :320-347, cut out of the
    unmodified source and executed by tests/golden/make_golden_sky.py) -- this is what pins SURVEY 8f N2, bit for bit."""
    import os
    from conftest import ROOT
    g = np.load(os.path.join(ROOT, "tests", "golden", "sky_golden.npz"))
    for k in range(int(g["n_norm"])):
        src, want = g[f"norm_{k}_in"], g[f"norm_{k}_out"]
        if src.dtype.kind != "f":
            continue                                        # integer FITS data: the reference promotes through numpy; torch callers pass floats
        got = sky.normalize_image(torch.as_tensor(src, device=device))
        assert got.cpu().numpy().dtype == want.dtype and np.array_equal(bits(got.cpu().numpy()), bits(want)), k
    with pytest.raises(ValueError, match=str(g["norm_constant_error"])):
        sky.normalize_image(torch.full((4, 4), 2.5, device=device))
    assert int(g["n_an"]) >= 6
    for k in range(int(g["n_an"])):
        grid = g[f"an_{k}_grid"]
        ext = [tuple(r) for r in g[f"an_{k}_extent"]]
        out = sky.analyze_voxel_grid(torch.as_tensor(grid, device=device), int(g[f"an_{k}_n_files"]), ext, float(g[f"an_{k}_pct"]))
        assert np.array_equal(bits(out["voxel_grid_avg"].cpu().numpy()), bits(g[f"an_{k}_avg"])), k
        assert float(out["threshold"]) == float(g[f"an_{k}_threshold"]), (k, out["threshold"], g[f"an_{k}_threshold"])
        for name in ("x_coords", "y_coords", "z_coords", "intensities"):
            assert np.array_equal(bits(out[name].cpu().numpy()), bits(g[f"an_{k}_{name}"])), (k, name)
        want_b = g[f"an_{k}_brightest"]
        assert (out["brightest"] is None) == (want_b.size == 0)
        if want_b.size:
            assert out["brightest"] == tuple(float(v) for v in want_b), k


def test_sky_matches_the_reference_lines_cpu():
    _golden_check("cpu")


@pytest.mark.gpu
def test_sky_matches_the_reference_lines_gpu():
    _golden_check("cuda")
