"""Path B host side: the Python-callable ``process_image_cpp`` entry point (boundary B-1).

Mirrors the pybind11 module of the reference (process_image.cpp:369-390): one function,
17 keyword-capable arguments in the same order, arrays mutated in place, returns None.
The call site it serves is spacevoxelviewer.py:233-251.  numpy arrays take the host path
(copy in, kernel, copy back in place, honouring arbitrary positive strides like the
reference's ``unchecked<>`` accessors); CUDA torch tensors are used in place and stay on the
device across calls.  All compute happens in libpvp_b200.so (K3); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np
import torch

from . import _lib
from ._lib import ACCUM_F64, ACCUM_FIXED, ImageArgs, ImageStats, check
from .motion import Context, default_context

last_stats: dict | None = None  # counters of the most recent call (n_rays, n_samples), for benchmarks


def _as_f64_array(x, name: str, ndim: int | None, mutated: bool = True):
    """pybind11's ``py::array_t<double>`` (forcecast) conversion: a non-float64 input is silently
    converted into a temporary, so in-place updates are lost to the caller -- as in the reference
    (hence the warning for the two arrays the call mutates; the image is only read)."""
    a = x if isinstance(x, np.ndarray) else np.asarray(x)
    if a.dtype != np.float64:
        if mutated:
            warnings.warn(f"process_image_cpp: {name} is {a.dtype}, not float64; it is converted to a temporary copy and "
                          "updates will not be visible to the caller (same as the reference's pybind11 forcecast)", stacklevel=3)
        a = a.astype(np.float64)
    if ndim is not None and a.ndim != ndim:
        raise ValueError(f"array has incorrect number of dimensions: {a.ndim}; expected {ndim}")
    return a


def _vec3(x, name):
    v = [float(c) for c in x]
    if len(v) != 3:
        raise TypeError(f"process_image_cpp(): incompatible function arguments: {name} must have 3 elements")
    return v


def _fill_common(a: ImageArgs, earth_position, pointing_direction, fov, image_width, image_height, voxel_grid_extent, max_distance,
                 num_steps, center_ra_rad, center_dec_rad, angular_width_rad, angular_height_rad, update_celestial_sphere,
                 perform_background_subtraction):
    a.earth_position[:] = _vec3(earth_position, "earth_position")
    a.pointing_direction[:] = _vec3(pointing_direction, "pointing_direction")
    a.fov = float(fov)
    a.image_width, a.image_height = int(image_width), int(image_height)
    ext = [(float(lo), float(hi)) for lo, hi in voxel_grid_extent]
    a.extent[:] = [c for pair in (ext + [(0.0, 0.0)] * 3)[:3] for c in pair]
    a.max_distance, a.num_steps = float(max_distance), int(num_steps)
    a.center_ra_rad, a.center_dec_rad = float(center_ra_rad), float(center_dec_rad)
    a.angular_width_rad, a.angular_height_rad = float(angular_width_rad), float(angular_height_rad)
    a.update_celestial_sphere = int(bool(update_celestial_sphere))
    a.perform_background_subtraction = int(bool(perform_background_subtraction))
    return len(ext)


def process_image_cpp(image, earth_position, pointing_direction, fov, image_width, image_height, voxel_grid, voxel_grid_extent,
                      max_distance, num_steps, celestial_sphere_texture, center_ra_rad, center_dec_rad, angular_width_rad,
                      angular_height_rad, update_celestial_sphere, perform_background_subtraction, *, accum_mode: str = "f64",
                      frac_bits: int = 32, ctx: Context | None = None, want_stats: bool = False,
                      row_range: tuple[int, int] | None = None) -> None:
    """Process image and update voxel grid (process_image.cpp:125-366) on the B200.

    The 17 positional arguments are the reference's.  Keyword-only extras: ``accum_mode="fixed"``
    (int64 CUDA tensors holding value * 2**frac_bits; deterministic), ``ctx``, ``want_stats``
    (fills ``process_image.last_stats``; forces a sync on the device path), ``row_range=(lo, hi)``
    (only image rows lo <= i < hi cast rays -- a row shard of a multi-GPU run, see ``distributed.process_image_sharded``).
    """
    global last_stats
    on_device = isinstance(image, torch.Tensor) and image.is_cuda
    args = ImageArgs()
    n_ext = _fill_common(args, earth_position, pointing_direction, fov, image_width, image_height, voxel_grid_extent, max_distance,
                         num_steps, center_ra_rad, center_dec_rad, angular_width_rad, angular_height_rad, update_celestial_sphere,
                         perform_background_subtraction)
    if accum_mode not in ("f64", "fixed"):
        raise ValueError("accum_mode must be 'f64' or 'fixed'")
    args.accum_mode = ACCUM_F64 if accum_mode == "f64" else ACCUM_FIXED
    args.frac_bits = int(frac_bits)
    if row_range is not None:
        lo, hi = int(row_range[0]), int(row_range[1])
        if not 0 <= lo <= hi <= int(image_height):
            raise ValueError(f"row_range {row_range} outside the image's {image_height} rows")
        if lo == hi:      # an empty shard: nothing to do (the C ABI reads 0, 0 as "all rows")
            last_stats = {"n_rays": 0, "n_samples": 0} if want_stats else last_stats
            return None
        args.row_begin, args.row_end = lo, hi
    st = ImageStats()
    st_ref = C.byref(st) if want_stats else None

    if on_device:
        cell = torch.float64 if accum_mode == "f64" else torch.int64
        tex, grid = celestial_sphere_texture, voxel_grid
        if image.dtype != torch.float64 or image.ndim != 2:
            raise ValueError(f"array has incorrect number of dimensions: {image.ndim}; expected 2" if image.ndim != 2
                             else "image must be a float64 CUDA tensor")
        if image.shape[0] < int(image_height) or image.shape[1] < int(image_width):
            raise ValueError(f"image of shape {tuple(image.shape)} is smaller than image_height x image_width = {int(image_height)} x {int(image_width)}")
        if not (isinstance(tex, torch.Tensor) and tex.is_cuda and tex.dtype == cell):
            raise TypeError(f"celestial_sphere_texture must be a CUDA {cell} tensor on the device path")
        if tex.ndim != 2:
            raise ValueError(f"array has incorrect number of dimensions: {tex.ndim}; expected 2")
        grid_ptr = None
        if grid is not None and isinstance(grid, torch.Tensor) and grid.numel() > 0:
            if not (grid.is_cuda and grid.dtype == cell):
                raise TypeError(f"voxel_grid must be a CUDA {cell} tensor on the device path")
            if grid.ndim != 3:
                raise ValueError(f"array has incorrect number of dimensions: {grid.ndim}; expected 3")
            if n_ext < 3:
                raise IndexError("voxel_grid_extent needs three (min, max) pairs")
            args.grid_shape[:] = list(grid.shape)
            args.grid_stride[:] = list(grid.stride())
            grid_ptr = grid.data_ptr()
        args.image_stride[:] = list(image.stride())
        args.tex_shape[:] = list(tex.shape)
        args.tex_stride[:] = list(tex.stride())
        if min(list(image.stride()) + list(tex.stride()) + list(args.grid_stride)) < 0:
            raise ValueError("negative strides are not supported on the device path")
        ctx = ctx or default_context(image.device)
        ctx.use_current_stream()
        check(ctx._lib.pvp_process_image(ctx.handle, image.data_ptr(), grid_ptr, tex.data_ptr(), C.byref(args), st_ref))
    else:
        if accum_mode != "f64":
            raise ValueError("fixed-point accumulation needs CUDA int64 tensors (device path)")
        img = _as_f64_array(image, "image", 2, mutated=False)
        if img.shape[0] < int(image_height) or img.shape[1] < int(image_width):   # the reference reads out of bounds here (unchecked<2>)
            raise ValueError(f"image of shape {img.shape} is smaller than image_height x image_width = {int(image_height)} x {int(image_width)}")
        tex = _as_f64_array(celestial_sphere_texture, "celestial_sphere_texture", 2)
        grid = voxel_grid if isinstance(voxel_grid, np.ndarray) else np.asarray(voxel_grid)
        have_grid = grid.size > 0  # process_image.cpp:152 -- an empty array of any ndim skips the voxel stage
        if have_grid:
            grid = _as_f64_array(grid, "voxel_grid", 3)
            if n_ext < 3:
                raise IndexError("voxel_grid_extent needs three (min, max) pairs")

        # negative strides: operate on a contiguous copy, copy back afterwards
        def _positive(a):
            return a if all(s >= 0 for s in a.strides) else np.ascontiguousarray(a)
        img_w, tex_w = _positive(img), _positive(tex)
        grid_w = _positive(grid) if have_grid else None
        args.image_stride[:] = [s // 8 for s in img_w.strides]
        args.tex_shape[:] = list(tex_w.shape)
        args.tex_stride[:] = [s // 8 for s in tex_w.strides]
        if have_grid:
            args.grid_shape[:] = list(grid_w.shape)
            args.grid_stride[:] = [s // 8 for s in grid_w.strides]
        ctx = ctx or default_context()
        ctx.use_current_stream()
        check(ctx._lib.pvp_process_image_host(ctx.handle, img_w.ctypes.data, grid_w.ctypes.data if have_grid else None,
                                              tex_w.ctypes.data, C.byref(args), st_ref))
        if tex_w is not tex:
            tex[...] = tex_w
        if have_grid and grid_w is not grid:
            grid[...] = grid_w
    if want_stats:
        last_stats = {"n_rays": int(st.n_rays), "n_samples": int(st.n_samples)}
    return None
