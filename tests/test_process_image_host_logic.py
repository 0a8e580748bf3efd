"""Host side of boundary B-1 without a GPU: what `process_image_cpp(...)` (process_image.cpp:369-390, call site
spacevoxelviewer.py:233-251) marshals into the C ABI for numpy callers.  The library call is replaced by a recorder that
reads the image and writes grid / texture THROUGH the pointers and strides it was given -- the way K3 addresses them -- so the
test sees whether the caller's arrays are the ones updated: C and Fortran order, strided views, negative strides (copy in /
copy back), the reference's silent temporary for non-float64 arrays, the empty grid that skips the voxel stage, and the error
behaviour SURVEY 8b lists.  No compute happens here."""
import ctypes as C
import warnings

import numpy as np
import pytest

from pixeltovoxelprojector_b200 import process_image as pi
from pixeltovoxelprojector_b200._lib import ACCUM_F64, ImageArgs

FIELDS = [f for f, _ in ImageArgs._fields_]


class RecordingLib:
    """Stands in for libpvp_b200.so: pvp_process_image_host(ctx, image, grid, tex, args, stats)."""

    def __init__(self):
        self.calls = []

    def pvp_process_image_host(self, handle, image, grid, tex, args_ref, stats_ref):
        a = args_ref._obj
        rec = {f: (list(getattr(a, f)) if hasattr(getattr(a, f), "__len__") else getattr(a, f)) for f in FIELDS}
        rec["grid_is_null"] = grid is None
        img = C.cast(image, C.POINTER(C.c_double))
        rec["image_1_2"] = img[1 * a.image_stride[0] + 2 * a.image_stride[1]]
        t = C.cast(tex, C.POINTER(C.c_double))
        t[1 * a.tex_stride[0] + 2 * a.tex_stride[1]] += 7.0
        if grid is not None:
            g = C.cast(grid, C.POINTER(C.c_double))
            g[1 * a.grid_stride[0] + 0 * a.grid_stride[1] + 2 * a.grid_stride[2]] += 5.0
        if stats_ref is not None:
            stats_ref._obj.n_rays, stats_ref._obj.n_samples = 11, 22
        self.calls.append(rec)
        return 0


class RecordingCtx:
    handle = None

    def __init__(self):
        self._lib = RecordingLib()

    def use_current_stream(self):
        pass


def call(ctx, image, grid, tex, extent=((-1.0, 1.0), (-2.0, 2.0), (-3.0, 3.0)), **kw):
    return pi.process_image_cpp(image, [1.0, 2.0, 3.0], (0.0, 0.0, 2.0), 0.5, image.shape[-1], image.shape[0], grid, list(extent), 10.0, 20, tex,
                                0.1, 0.2, 0.3, 0.4, True, False, ctx=ctx, **kw)


def fresh():
    rng = np.random.default_rng(0)
    return rng.random((4, 6)), np.zeros((3, 4, 5)), np.zeros((7, 9))


def test_arguments_reach_the_c_abi_as_the_reference_means_them():
    ctx = RecordingCtx()
    image, grid, tex = fresh()
    assert call(ctx, image, grid, tex) is None                      # returns None like the pybind11 function
    a = ctx._lib.calls[0]
    assert a["earth_position"] == [1.0, 2.0, 3.0] and a["pointing_direction"] == [0.0, 0.0, 2.0] and a["fov"] == 0.5
    assert (a["image_width"], a["image_height"]) == (6, 4) and a["image_stride"] == [6, 1]
    assert a["grid_shape"] == [3, 4, 5] and a["grid_stride"] == [20, 5, 1] and a["extent"] == [-1.0, 1.0, -2.0, 2.0, -3.0, 3.0]
    assert a["max_distance"] == 10.0 and a["num_steps"] == 20 and a["tex_shape"] == [7, 9] and a["tex_stride"] == [9, 1]
    assert (a["center_ra_rad"], a["center_dec_rad"], a["angular_width_rad"], a["angular_height_rad"]) == (0.1, 0.2, 0.3, 0.4)
    assert a["update_celestial_sphere"] == 1 and a["perform_background_subtraction"] == 0 and a["accum_mode"] == ACCUM_F64
    assert (a["row_begin"], a["row_end"]) == (0, 0) and not a["grid_is_null"]
    assert a["image_1_2"] == image[1, 2]
    assert grid[1, 0, 2] == 5.0 and grid.sum() == 5.0 and tex[1, 2] == 7.0 and tex.sum() == 7.0      # mutated in place


def test_fortran_order_and_strided_views_are_updated_in_place():
    ctx = RecordingCtx()
    image, grid, tex = fresh()
    gf, tf, imf = np.asfortranarray(grid), np.asfortranarray(tex), np.asfortranarray(image)
    call(ctx, imf, gf, tf)
    a = ctx._lib.calls[-1]
    assert a["grid_stride"] == [1, 3, 12] and a["tex_stride"] == [1, 7] and a["image_stride"] == [1, 4] and a["image_1_2"] == image[1, 2]
    assert gf[1, 0, 2] == 5.0 and gf.sum() == 5.0 and tf[1, 2] == 7.0
    base_g, base_t, base_i = np.zeros((6, 4, 15)), np.zeros((7, 27)), np.random.default_rng(1).random((8, 12))
    call(ctx, base_i[::2, ::2], base_g[::2, :, ::3], base_t[:, ::3])
    a = ctx._lib.calls[-1]
    assert a["grid_shape"] == [3, 4, 5] and a["grid_stride"] == [120, 15, 3] and a["tex_stride"] == [27, 3] and a["image_stride"] == [24, 2]
    assert a["image_1_2"] == base_i[2, 4] and base_g[2, 0, 6] == 5.0 and base_g.sum() == 5.0 and base_t[1, 6] == 7.0


def test_negative_strides_go_through_a_temporary_and_come_back():
    ctx = RecordingCtx()
    image, grid, tex = fresh()
    call(ctx, image[::-1], grid[::-1], tex[:, ::-1])
    a = ctx._lib.calls[-1]
    assert min(a["grid_stride"] + a["tex_stride"] + a["image_stride"]) > 0
    assert a["image_1_2"] == image[::-1][1, 2]
    assert grid[::-1][1, 0, 2] == 5.0 and grid[1, 0, 2] == 5.0 and grid.sum() == 5.0       # view[1] is base[3-1-1] = base[1]
    assert tex[:, ::-1][1, 2] == 7.0 and tex[1, 9 - 1 - 2] == 7.0 and tex.sum() == 7.0


def test_non_float64_arrays_become_silent_temporaries_like_pybind11_forcecast():
    ctx = RecordingCtx()
    image, grid, tex = fresh()
    g32, t32 = grid.astype(np.float32), tex.astype(np.float32)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        call(ctx, image.astype(np.float32), g32, t32)
    assert len(w) == 2 and all("not float64" in str(x.message) for x in w)     # grid and texture; the image is only read
    assert g32.sum() == 0 and t32.sum() == 0                             # SURVEY 8b [probed]: float32 grid -> grid.sum() == 0 after the call
    assert ctx._lib.calls[-1]["image_1_2"] == float(np.float32(image[1, 2]))


def test_empty_grid_skips_the_voxel_stage_whatever_its_ndim():
    ctx = RecordingCtx()
    image, _, tex = fresh()
    for empty in (np.zeros((0,)), np.zeros((0, 3, 3)), np.zeros((2, 0)), np.array([])):      # process_image.cpp:152: size() == 0
        call(ctx, image, empty, tex, extent=())
        assert ctx._lib.calls[-1]["grid_is_null"] and ctx._lib.calls[-1]["grid_shape"] == [0, 0, 0]
    assert tex[1, 2] == 7.0 * 4


def test_errors_follow_the_reference():
    ctx = RecordingCtx()
    image, grid, tex = fresh()
    with pytest.raises(ValueError, match="incorrect number of dimensions: 2; expected 3"):    # [probed] message of unchecked<3>
        call(ctx, image, np.zeros((3, 4)), tex)
    with pytest.raises(ValueError, match="incorrect number of dimensions: 3; expected 2"):
        call(ctx, np.zeros((2, 3, 4)), grid, tex)
    with pytest.raises(ValueError, match="incorrect number of dimensions: 1; expected 2"):
        call(ctx, image, grid, np.zeros(5))
    with pytest.raises(ValueError):                                                              # None is not an array (pybind11: 0-dim)
        call(ctx, image, None, tex)
    with pytest.raises(TypeError):
        pi.process_image_cpp(image, [1.0, 2.0], (0.0, 0.0, 1.0), 0.5, 6, 4, grid, [(-1, 1)] * 3, 10.0, 20, tex, 0, 0, 1, 1, True, False, ctx=ctx)
    with pytest.raises(TypeError):                                                               # arity, like pybind11
        pi.process_image_cpp(image, [1.0, 2.0, 3.0])
    with pytest.raises(IndexError):
        call(ctx, image, grid, tex, extent=((-1.0, 1.0), (-2.0, 2.0)))
    with pytest.raises(ValueError, match="fixed-point"):
        call(ctx, image, grid, tex, accum_mode="fixed")
    with pytest.raises(ValueError, match="accum_mode"):
        call(ctx, image, grid, tex, accum_mode="f32")
    assert ctx._lib.calls == [] and grid.sum() == 0 and tex.sum() == 0       # nothing reached the library


def test_row_range_and_stats():
    ctx = RecordingCtx()
    image, grid, tex = fresh()
    call(ctx, image, grid, tex, row_range=(1, 3), want_stats=True)
    assert (ctx._lib.calls[-1]["row_begin"], ctx._lib.calls[-1]["row_end"]) == (1, 3) and pi.last_stats == {"n_rays": 11, "n_samples": 22}
    call(ctx, image, grid, tex, row_range=(2, 2), want_stats=True)            # an empty shard never reaches the library
    assert len(ctx._lib.calls) == 1 and pi.last_stats == {"n_rays": 0, "n_samples": 0}
    for bad in ((-1, 2), (3, 2), (0, 5)):
        with pytest.raises(ValueError, match="row_range"):
            call(ctx, image, grid, tex, row_range=bad)
