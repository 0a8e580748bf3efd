"""The COMPILED drop-in for the reference's Python extension (setup.py:27-36, process_image.cpp:369-390):
pixeltovoxelprojector_b200/compiled/process_image_cpp.<abi>.so, pybind11 over the C ABI (csrc/process_image_cpp_module.cpp).
CPU: it imports without a GPU, has the reference's 17 arguments by name and order, and pybind11 raises the reference's errors.
GPU: the reference module's golden outputs, in-place semantics through strides, forcecast behaviour; and the per-call cost of the two
drop-ins (compiled, pure Python) for the reference's own call pattern -- many small images into one large grid."""
import glob
import importlib.util
import json
import os
import time

import numpy as np
import pytest

from conftest import ROOT, golden

NAMES = ["image", "earth_position", "pointing_direction", "fov", "image_width", "image_height", "voxel_grid", "voxel_grid_extent", "max_distance",
         "num_steps", "celestial_sphere_texture", "center_ra_rad", "center_dec_rad", "angular_width_rad", "angular_height_rad",
         "update_celestial_sphere", "perform_background_subtraction"]


@pytest.fixture(scope="module")
def compiled():
    hits = glob.glob(os.path.join(ROOT, "pixeltovoxelprojector_b200", "compiled", "process_image_cpp*.so"))
    if not hits:
        pytest.skip("compiled extension not built (make -C pixeltovoxelprojector_b200/csrc ext)")
    spec = importlib.util.spec_from_file_location("process_image_cpp", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_module_surface_is_the_reference_s(compiled):
    doc = compiled.process_image_cpp.__doc__
    pos = [doc.index(n + ":") for n in NAMES]                 # every argument, by name
    assert pos == sorted(pos)                                  # in the reference's order (process_image.cpp:372-388)
    assert [n for n in dir(compiled) if not n.startswith("_")] == ["process_image_cpp"]
    with pytest.raises(TypeError):
        compiled.process_image_cpp(np.zeros((4, 4)))           # arity
    base = dict(image=np.zeros((4, 4)), earth_position=[0, 0, 0], pointing_direction=[0, 0, 1], fov=0.5, image_width=4, image_height=4,
                voxel_grid=np.zeros((3, 3, 3)), voxel_grid_extent=[(0, 1)] * 3, max_distance=10.0, num_steps=10,
                celestial_sphere_texture=np.zeros((4, 4)), center_ra_rad=0, center_dec_rad=0, angular_width_rad=1, angular_height_rad=1,
                update_celestial_sphere=True, perform_background_subtraction=False)
    with pytest.raises(ValueError, match="incorrect number of dimensions: 2; expected 3"):
        compiled.process_image_cpp(**{**base, "voxel_grid": np.zeros((3, 3))})
    with pytest.raises(ValueError, match="incorrect number of dimensions: 1; expected 2"):
        compiled.process_image_cpp(**{**base, "image": np.zeros(4)})
    with pytest.raises(TypeError):
        compiled.process_image_cpp(**{**base, "earth_position": [0, 0]})       # std::array<double, 3>
    with pytest.raises(ValueError, match="incorrect number of dimensions: 0; expected 3"):
        compiled.process_image_cpp(**{**base, "voxel_grid": None})              # the reference's own message for None (probed)


def _call(fn, g, name, img, grid, tex):
    H, W = g["image"].shape
    return fn(img, list(g["earth_position"]), list(g[f"{name}_pointing"]), float(g["fov"]), W, H, grid, [tuple(r) for r in g["extent"]],
              float(g["max_distance"]), int(g["num_steps"]), tex, *[float(s) for s in g["sky"]], bool(g[f"{name}_update"]), bool(g[f"{name}_bgsub"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["update", "bgsub", "zenith"])
def test_golden_of_the_reference_module(compiled, ctx, name):
    g = golden("process_image_golden.npz")
    grid = np.zeros_like(g[f"{name}_grid_out"])
    tex = g[f"{name}_tex_in"].copy()
    assert _call(compiled.process_image_cpp, g, name, g["image"], grid, tex) is None
    want_g, want_t = g[f"{name}_grid_out"], g[f"{name}_tex_out"]
    assert np.array_equal(grid != 0, want_g != 0)
    np.testing.assert_allclose(grid, want_g, rtol=1e-12, atol=0)
    np.testing.assert_allclose(tex, want_t, rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_in_place_through_strides_and_forcecast(compiled, ctx):
    g = golden("process_image_golden.npz")
    name = "update"
    want_g, want_t = g[f"{name}_grid_out"], g[f"{name}_tex_out"]
    big = np.zeros((want_g.shape[0], want_g.shape[1] * 2, want_g.shape[2]))
    view = big[:, ::2, :]
    tex = np.asfortranarray(g[f"{name}_tex_in"].copy())
    _call(compiled.process_image_cpp, g, name, g["image"], view, tex)
    np.testing.assert_allclose(view, want_g, rtol=1e-12, atol=0)
    np.testing.assert_allclose(tex, want_t, rtol=1e-12, atol=0)
    assert not big[:, 1::2, :].any()
    rev = np.zeros_like(want_g)[::-1]                          # negative stride
    _call(compiled.process_image_cpp, g, name, g["image"], rev, g[f"{name}_tex_in"].copy())
    np.testing.assert_allclose(rev, want_g, rtol=1e-12, atol=0)
    tex2 = g[f"{name}_tex_in"].copy()                          # empty grid of any ndim: voxel stage skipped (:152)
    _call(compiled.process_image_cpp, g, name, g["image"], np.zeros((0,)), tex2)
    np.testing.assert_allclose(tex2, want_t, rtol=1e-12, atol=0)
    g32 = np.zeros(want_g.shape, np.float32)                   # forcecast: silently copied, the caller sees no update
    _call(compiled.process_image_cpp, g, name, g["image"], g32, g[f"{name}_tex_in"].copy())
    assert not g32.any()


@pytest.mark.gpu
def test_per_call_cost_for_the_reference_call_pattern(compiled, ctx):
    """spacevoxelviewer.py:300-310 calls the extension once per FITS file with a small image and one large grid (400^3 float64 = 512 MB
    there).  With numpy arrays both drop-ins must move the whole grid over PCIe twice per call; with CUDA tensors (Python drop-in)
    nothing but the image moves.  The figures go to gpurun_out/dropin_call_latency.json (copied into profiles/ by the round)."""
    import torch
    import process_image_cpp as dropin
    import pixeltovoxelprojector_b200 as pvp
    rng = np.random.default_rng(5)
    n, hw, calls = 256, 128, 6
    ext = [(-128.0, 128.0), (-128.0, 128.0), (100.0, 356.0)]
    imgs = rng.random((calls, hw, hw))
    args = ([0.0, 0.0, 0.0], [0.02, 0.01, 1.0], 0.05, hw, hw)
    tail = (ext, 400.0, 400)
    sky = (0.0, 0.0, 1.0, 1.0, False, False)
    out = {"grid": f"{n}^3 float64 ({n ** 3 * 8 >> 20} MiB)", "image": f"{hw}x{hw}", "calls": calls}
    grids = {}
    for label, fn in (("compiled_numpy", compiled.process_image_cpp), ("python_numpy", dropin.process_image_cpp)):
        grid, tex = np.zeros((n, n, n)), np.zeros((8, 8))
        fn(imgs[0], *args, grid, *tail, tex, *sky)                            # warm-up (context, kernels)
        grid[:] = 0
        t0 = time.perf_counter()
        for k in range(calls):
            fn(imgs[k], *args, grid, *tail, tex, *sky)
        out[label + "_ms_per_call"] = 1e3 * (time.perf_counter() - t0) / calls
        grids[label] = grid
    dgrid, dtex = torch.zeros((n, n, n), dtype=torch.float64, device="cuda"), torch.zeros((8, 8), dtype=torch.float64, device="cuda")
    dimgs = torch.as_tensor(imgs, device="cuda")
    pvp.process_image_cpp(dimgs[0], *args, dgrid, *tail, dtex, *sky)
    dgrid.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(calls):
        pvp.process_image_cpp(dimgs[k], *args, dgrid, *tail, dtex, *sky)
    torch.cuda.synchronize()
    out["python_cuda_tensors_ms_per_call"] = 1e3 * (time.perf_counter() - t0) / calls
    np.testing.assert_allclose(grids["compiled_numpy"], grids["python_numpy"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(dgrid.cpu().numpy(), grids["python_numpy"], rtol=1e-12, atol=0)
    assert np.count_nonzero(grids["python_numpy"]) > 1000
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dropin_call_latency.json"), "w"), indent=1)
    assert out["python_cuda_tensors_ms_per_call"] < out["python_numpy_ms_per_call"]
