"""The COMPILED drop-in for the reference's Python extension (setup.py:27-36, process_image.cpp:369-390):
pixeltovoxelprojector_b200/compiled/process_image_cpp.<abi>.so, pybind11 over the C ABI (csrc/process_image_cpp_module.cpp).
CPU: it imports without a GPU, has the reference's 17 arguments by name and order, and pybind11 raises the reference's errors.
GPU: the reference module's golden outputs, in-place semantics through strides, forcecast behaviour; and the per-call cost of the two
drop-ins (compiled, pure Python) for the reference's own call pattern -- many small images into one large grid.

Every check runs in a subprocess with compiled/ first on sys.path -- the way a user would switch -- because the module has, by design, the
NAME of the reference's module: pybind11 keeps one module object per name and process, so it cannot share a process with oracle/_ref's
build of the reference extension (which other tests of this suite load)."""
import glob
import os
import subprocess
import sys

import pytest

from conftest import ROOT

COMPILED = os.path.join(ROOT, "pixeltovoxelprojector_b200", "compiled")

PRELUDE = r'''
import json, os, sys, time
import numpy as np
ROOT = %r
sys.path.insert(0, os.path.join(ROOT, "pixeltovoxelprojector_b200", "compiled"))   # what a user of the compiled drop-in does
import process_image_cpp as compiled
assert compiled.__file__.endswith(".so") and os.sep + "compiled" + os.sep in compiled.__file__, compiled.__file__
sys.path.insert(1, ROOT)
golden = lambda name: np.load(os.path.join(ROOT, "tests", "golden", name), allow_pickle=False)
def call(fn, g, name, img, grid, tex):
    H, W = g["image"].shape
    return fn(img, list(g["earth_position"]), list(g[name + "_pointing"]), float(g["fov"]), W, H, grid, [tuple(r) for r in g["extent"]],
              float(g["max_distance"]), int(g["num_steps"]), tex, *[float(s) for s in g["sky"]], bool(g[name + "_update"]), bool(g[name + "_bgsub"]))
''' % ROOT


def run(code):
    if not glob.glob(os.path.join(COMPILED, "process_image_cpp*.so")):
        pytest.skip("compiled extension not built (make -C pixeltovoxelprojector_b200/csrc ext)")
    r = subprocess.run([sys.executable, "-c", PRELUDE + code], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "CHECKS PASSED" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
    return r.stdout


def test_module_surface_is_the_reference_s():
    run(r'''
NAMES = ["image", "earth_position", "pointing_direction", "fov", "image_width", "image_height", "voxel_grid", "voxel_grid_extent", "max_distance",
         "num_steps", "celestial_sphere_texture", "center_ra_rad", "center_dec_rad", "angular_width_rad", "angular_height_rad",
         "update_celestial_sphere", "perform_background_subtraction"]
doc = compiled.process_image_cpp.__doc__
pos = [doc.index(n + ":") for n in NAMES]                  # every argument, by name
assert pos == sorted(pos)                                  # in the reference's order (process_image.cpp:372-388)
assert [n for n in dir(compiled) if not n.startswith("_")] == ["process_image_cpp"]
def raises(exc, match, **kw):
    base = dict(image=np.zeros((4, 4)), earth_position=[0, 0, 0], pointing_direction=[0, 0, 1], fov=0.5, image_width=4, image_height=4,
                voxel_grid=np.zeros((3, 3, 3)), voxel_grid_extent=[(0, 1)] * 3, max_distance=10.0, num_steps=10,
                celestial_sphere_texture=np.zeros((4, 4)), center_ra_rad=0, center_dec_rad=0, angular_width_rad=1, angular_height_rad=1,
                update_celestial_sphere=True, perform_background_subtraction=False)
    try:
        compiled.process_image_cpp(**{**base, **kw})
    except exc as e:
        assert match in str(e), str(e)
        return
    raise AssertionError("no %s for %s" % (exc.__name__, list(kw)))
try:
    compiled.process_image_cpp(np.zeros((4, 4)))           # arity
    raise AssertionError("arity")
except TypeError:
    pass
raises(ValueError, "incorrect number of dimensions: 2; expected 3", voxel_grid=np.zeros((3, 3)))
raises(ValueError, "incorrect number of dimensions: 1; expected 2", image=np.zeros(4))
raises(TypeError, "", earth_position=[0, 0])               # std::array<double, 3>
raises(ValueError, "incorrect number of dimensions: 0; expected 3", voxel_grid=None)   # the reference's own message for None (SURVEY 8b, probed)
print("CHECKS PASSED")
''')


@pytest.mark.gpu
def test_golden_of_the_reference_module_strides_and_forcecast():
    run(r'''
g = golden("process_image_golden.npz")
for name in ("update", "bgsub", "zenith"):
    grid = np.zeros_like(g[name + "_grid_out"])
    tex = g[name + "_tex_in"].copy()
    assert call(compiled.process_image_cpp, g, name, g["image"], grid, tex) is None     # mutates in place, returns None
    want_g, want_t = g[name + "_grid_out"], g[name + "_tex_out"]
    assert np.array_equal(grid != 0, want_g != 0)                                       # identical visited voxels
    np.testing.assert_allclose(grid, want_g, rtol=1e-12, atol=0)
    np.testing.assert_allclose(tex, want_t, rtol=1e-12, atol=0)
name = "update"
want_g, want_t = g[name + "_grid_out"], g[name + "_tex_out"]
big = np.zeros((want_g.shape[0], want_g.shape[1] * 2, want_g.shape[2]))
view = big[:, ::2, :]                                      # strided view: updated in place like the reference
tex = np.asfortranarray(g[name + "_tex_in"].copy())        # F-order texture
call(compiled.process_image_cpp, g, name, g["image"], view, tex)
np.testing.assert_allclose(view, want_g, rtol=1e-12, atol=0)
np.testing.assert_allclose(tex, want_t, rtol=1e-12, atol=0)
assert not big[:, 1::2, :].any()
rev = np.zeros_like(want_g)[::-1]                          # negative stride
call(compiled.process_image_cpp, g, name, g["image"], rev, g[name + "_tex_in"].copy())
np.testing.assert_allclose(rev, want_g, rtol=1e-12, atol=0)
tex2 = g[name + "_tex_in"].copy()                          # empty grid of any ndim: voxel stage skipped (:152)
call(compiled.process_image_cpp, g, name, g["image"], np.zeros((0,)), tex2)
np.testing.assert_allclose(tex2, want_t, rtol=1e-12, atol=0)
g32 = np.zeros(want_g.shape, np.float32)                   # forcecast: silently copied, the caller sees no update
call(compiled.process_image_cpp, g, name, g["image"], g32, g[name + "_tex_in"].copy())
assert not g32.any()
print("CHECKS PASSED")
''')


@pytest.mark.gpu
def test_per_call_cost_for_the_reference_call_pattern():
    """spacevoxelviewer.py:300-310 calls the extension once per FITS file with a small image and one large grid (400^3 float64 = 512 MB
    there).  With numpy arrays both drop-ins must move the whole grid over PCIe twice per call; with CUDA tensors (Python drop-in)
    nothing but the image moves.  The figures go to gpurun_out/dropin_call_latency.json (copied into profiles/ by the round)."""
    run(r'''
import torch
import importlib.util
spec = importlib.util.spec_from_file_location("pvp_python_dropin", os.path.join(ROOT, "process_image_cpp.py"))   # the pure-Python drop-in (same name: load by path)
dropin = importlib.util.module_from_spec(spec); spec.loader.exec_module(dropin)
import pixeltovoxelprojector_b200 as pvp
rng = np.random.default_rng(5)
n, hw, calls = 256, 128, 6
ext = [(-128.0, 128.0), (-128.0, 128.0), (100.0, 356.0)]
imgs = rng.random((calls, hw, hw))
args = ([0.0, 0.0, 0.0], [0.02, 0.01, 1.0], 0.05, hw, hw)
tail = (ext, 400.0, 400)
sky = (0.0, 0.0, 1.0, 1.0, False, False)
out = {"grid": "%d^3 float64 (%d MiB)" % (n, n ** 3 * 8 >> 20), "image": "%dx%d" % (hw, hw), "calls": calls}
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
print("CHECKS PASSED")
''')
