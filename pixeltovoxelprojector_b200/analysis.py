"""The consumer side of the .bin with the grid kept in HBM: drop-in mirrors of
voxelmotionviewer.py's ``load_voxel_grid`` (:28-53) and ``extract_top_percentile_z_up`` (:56-96).

The data-parallel parts -- the order statistics behind ``np.percentile`` over N^3 cells and the ordered
``np.argwhere(grid > thresh)`` extraction -- run in libpvp_b200.so (csrc/pvp_analysis.cu).  The scalar arithmetic
around them (numpy's "linear" percentile interpolation, grid_min) is evaluated here with the same numpy expressions
the reference evaluates, so dtypes and roundings follow the installed numpy exactly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import check
from .motion import Context, default_context, read_bin


def load_voxel_grid(filename, device: torch.device | str | None = "cuda"):
    """voxelmotionviewer.py:28-53: returns (voxel_grid [N,N,N] float32, voxel_size np.float32); the grid is a CUDA
    tensor (``device=None`` keeps the numpy array the reference returns)."""
    grid, voxel_size = read_bin(os.fspath(filename))
    if device is None:
        return grid, voxel_size
    return torch.as_tensor(grid, device=device), voxel_size


def percentile_indices(n: int, percentile, dtype=np.float32):
    """Index arithmetic of ``np.percentile(flat, percentile)`` (method "linear") for an array of n cells of `dtype`:
    returns (previous_index, next_index, gamma) exactly as numpy derives them -- the same expressions on the same
    scalar types (numpy/lib/_function_base_impl.py: percentile, _quantile, _get_indexes, _get_gamma)."""
    dtype = np.dtype(dtype)
    q = np.true_divide(percentile, dtype.type(100) if dtype.kind == "f" else 100)
    q = np.asanyarray(q)
    if not (0 <= q <= 1):
        raise ValueError("Percentiles must be in the range [0, 100]")
    virtual = np.asanyarray((n - 1) * q)
    previous = np.asanyarray(np.floor(virtual))
    nxt = np.asanyarray(previous + 1)
    if virtual >= n - 1:
        previous, nxt = np.asanyarray(-1.0), np.asanyarray(-1.0)
    if virtual < 0:
        previous, nxt = np.asanyarray(0.0), np.asanyarray(0.0)
    previous_i, next_i = int(previous.astype(np.intp)), int(nxt.astype(np.intp))
    gamma = np.asanyarray(np.asanyarray(virtual - previous_i), dtype=virtual.dtype)   # previous_i may be -1, as in numpy
    return previous_i % n, next_i % n, gamma


def lerp_like_numpy(a, b, t):
    """numpy's _lerp for two 0-d values of the array dtype and the weight gamma."""
    a, b, t = np.asanyarray(a), np.asanyarray(b), np.asanyarray(t)
    diff_b_a = np.subtract(b, a)
    out = np.asanyarray(np.add(a, diff_b_a * t))
    np.subtract(b, diff_b_a * (1 - t), out=out, where=t >= 0.5, casting="unsafe", dtype=type(out.dtype))
    return out[()]


def grid_percentile(voxel_grid: torch.Tensor, percentile, ctx: Context | None = None):
    """``np.percentile(voxel_grid.ravel(), percentile)`` for a float32 CUDA tensor: np.float32 scalar, same value."""
    if not (isinstance(voxel_grid, torch.Tensor) and voxel_grid.is_cuda and voxel_grid.dtype == torch.float32):
        raise TypeError("voxel_grid must be a float32 CUDA tensor")
    g = voxel_grid.contiguous()
    n = g.numel()
    if n == 0:
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")
    ctx = ctx or default_context(g.device)
    ctx.use_current_stream()
    prev_i, next_i, gamma = percentile_indices(n, percentile, np.float32)
    lo, hi = C.c_float(), C.c_float()
    check(ctx._lib.pvp_grid_order_stats(ctx.handle, g.data_ptr(), n, prev_i, C.byref(lo), C.byref(hi)))
    a = np.float32(lo.value)
    b = a if next_i == prev_i else np.float32(hi.value)
    return lerp_like_numpy(a, b, gamma)


def _threshold_as_numpy_compares(thresh) -> float:
    """The value `float32_array > thresh` (:77) really compares with: a Python scalar is weak and is rounded to float32 first; a numpy
    scalar promotes by its dtype (float32 / small ints stay float32, float64 makes the comparison double)."""
    if isinstance(thresh, np.generic) or isinstance(thresh, np.ndarray):
        return float(thresh) if np.result_type(np.float32, np.asarray(thresh).dtype) == np.float64 else float(np.float32(thresh))
    return float(np.float32(thresh))


def extract_top_percentile_z_up(voxel_grid, voxel_size, grid_center, percentile=99.5, use_hard_thresh=False, hard_thresh=700,
                                ctx: Context | None = None):
    """voxelmotionviewer.py:56-96 on the device.  Returns (points [M,3] float64, intensities [M] float32) as CUDA
    tensors in the reference's order (np.argwhere's C order), or (None, None) after the reference's message when no
    voxel exceeds the threshold."""
    if not (isinstance(voxel_grid, torch.Tensor) and voxel_grid.is_cuda and voxel_grid.dtype == torch.float32 and voxel_grid.ndim == 3):
        raise TypeError("voxel_grid must be a 3-D float32 CUDA tensor (see load_voxel_grid)")
    g = voxel_grid.contiguous()
    N = g.shape[0]  # assume shape is (N,N,N), like the reference
    if g.shape != (N, N, N):
        raise ValueError("voxel_grid must be N x N x N")
    ctx = ctx or default_context(g.device)
    ctx.use_current_stream()
    # :67-69 -- N is a Python int there (shape[0]), so with the np.float32 voxel_size load_voxel_grid returns the product and the
    # half stay float32 (a weak scalar does not promote); np.int32(N) would promote to float64 and move grid_min by an ulp of float32
    half_side = (int(N) * voxel_size) * 0.5
    grid_min = np.asarray(grid_center) - half_side
    if use_hard_thresh:
        thresh = hard_thresh
    else:
        thresh = grid_percentile(g, percentile, ctx)
    thresh_cmp = _threshold_as_numpy_compares(thresh)
    gm = (C.c_double * 3)(*[float(x) for x in np.asarray(grid_min, np.float64)])
    count = C.c_int64()
    lib = ctx._lib
    check(lib.pvp_grid_extract_above(ctx.handle, g.data_ptr(), N, thresh_cmp, float(voxel_size), gm, 0, None, None, None, C.byref(count)))
    m = count.value
    if m == 0:
        print(f"No voxels above threshold {thresh}. Nothing to display.")
        return None, None
    points = torch.empty((m, 3), dtype=torch.float64, device=g.device)
    intens = torch.empty((m,), dtype=torch.float32, device=g.device)
    check(lib.pvp_grid_extract_above(ctx.handle, g.data_ptr(), N, thresh_cmp, float(voxel_size), gm, m, points.data_ptr(), intens.data_ptr(), None,
                                     C.byref(count)))
    return points, intens


def argwhere_above(voxel_grid: torch.Tensor, thresh, ctx: Context | None = None) -> torch.Tensor:
    """``np.argwhere(voxel_grid > thresh)`` ([M,3] int32, C order) on the device."""
    g = voxel_grid.contiguous()
    N = g.shape[0]
    ctx = ctx or default_context(g.device)
    ctx.use_current_stream()
    count = C.c_int64()
    thresh = _threshold_as_numpy_compares(thresh)
    check(ctx._lib.pvp_grid_extract_above(ctx.handle, g.data_ptr(), N, float(thresh), 1.0, None, 0, None, None, None, C.byref(count)))
    coords = torch.empty((count.value, 3), dtype=torch.int32, device=g.device)
    if count.value:
        check(ctx._lib.pvp_grid_extract_above(ctx.handle, g.data_ptr(), N, float(thresh), 1.0, None, count.value, None, None, coords.data_ptr(),
                                              C.byref(count)))
    return coords
