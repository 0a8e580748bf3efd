"""Tensor-level pre- and post-processing around process_image_cpp: the numpy steps that feed and
drain the extension inside spacevoxelviewer.py, on the device, so that a FITS loop keeps image, grid and texture in HBM.

* :func:`normalize_image`      spacevoxelviewer.py:179-184  (nan_to_num, min-max normalise, zero-range error)
* :func:`field_of_view_rad`    spacevoxelviewer.py:197-215  (FITS header 'FOV' | CD matrix | default 2.7 arcmin -> radians; host scalars)
* :func:`analyze_voxel_grid`   spacevoxelviewer.py:317-347  (average, percentile threshold over the positive cells,
                               object voxels -> world coordinates, intensities, brightest voxel)

The reference evaluates these inline in ``process_image`` / ``main`` (astropy I/O around them, not importable here); the
functions below restate the same numpy expressions as torch expressions, operation for operation, so dtypes and roundings
are those of the reference.  torch supplies the elementwise arithmetic and the selection primitives -- plumbing around the
hot path, like PyTorch everywhere else in this package; astropy's ephemeris / FITS handling stays on the host.
"""
from __future__ import annotations

import numpy as np
import torch

from .analysis import lerp_like_numpy, percentile_indices


def normalize_image(image_data: torch.Tensor) -> torch.Tensor:
    """spacevoxelviewer.py:179-184: ``nan_to_num(nan=0, posinf=0, neginf=0)`` then ``(x - min) / (max - min)``; raises the
    reference's ValueError for a constant image.  The dtype is kept (the caller applies ``astype(float64)``, :234)."""
    if image_data.ndim != 2:
        raise ValueError("Image data is not 2D.")                       # :176-177
    x = torch.nan_to_num(image_data, nan=0.0, posinf=0.0, neginf=0.0)
    image_min, image_max = x.min(), x.max()
    if bool(image_max - image_min == 0):
        raise ValueError("Image data has zero dynamic range.")
    return (x - image_min) / (image_max - image_min)


def _percentile_of(values: torch.Tensor, percentile) -> np.generic:
    """np.percentile(values, percentile) (method "linear") for a 1-D float tensor, value-identical to numpy."""
    n = values.numel()
    dt = {torch.float64: np.float64, torch.float32: np.float32}[values.dtype]
    prev_i, next_i, gamma = percentile_indices(n, percentile, dt)
    lo = torch.kthvalue(values, prev_i + 1).values
    hi = lo if next_i == prev_i else torch.kthvalue(values, next_i + 1).values
    return lerp_like_numpy(dt(lo.item()), dt(hi.item()), gamma)


def field_of_view_rad(header, width: int, height: int, default_fov_arcminutes: float = 2.7) -> np.float64:
    """spacevoxelviewer.py:197-215: the ``fov`` argument of ``process_image_cpp`` (radians) from a FITS header mapping: the
    'FOV' card in degrees if present, else the larger side of the image from the CD matrix' pixel scales, else the reference's
    default of 2.7 arcminutes (``:44``).  Per-file host scalars (numpy float64, the reference's own expressions)."""
    fov = header.get('FOV')
    if fov is None:
        cd = [header.get(k) for k in ('CD1_1', 'CD1_2', 'CD2_1', 'CD2_2')]
        if all(c is not None for c in cd):
            cd1_1, cd1_2, cd2_1, cd2_2 = cd
            pixel_scale_x = np.sqrt(cd1_1**2 + cd2_1**2)
            pixel_scale_y = np.sqrt(cd1_2**2 + cd2_2**2)
            fov = max(pixel_scale_x * width, pixel_scale_y * height)
        else:
            fov = default_fov_arcminutes / 60  # degrees
    else:
        fov = float(fov)
    return np.deg2rad(fov)


def analyze_voxel_grid(voxel_grid: torch.Tensor, n_images: int, voxel_grid_extent, brightness_threshold_percentile=.1) -> dict:
    """spacevoxelviewer.py:317-347 on the device.  Returns the quantities the reference computes there:
    ``voxel_grid_avg``, ``threshold``, ``x_coords / y_coords / z_coords`` (float64, np.nonzero order), ``intensities`` and
    ``brightest`` = (x, y, z) of the first maximum (None when no voxel is significant)."""
    # torch divides by a PYTHON scalar as a multiplication by its reciprocal on CUDA; numpy divides.  Divisors are therefore
    # passed as 0-dim device tensors, which takes the true (IEEE) division kernel.
    def dev(v):
        return torch.tensor(float(v), dtype=torch.float64, device=voxel_grid.device)

    voxel_grid_avg = voxel_grid / dev(n_images).to(voxel_grid.dtype)                     # :320
    positive = voxel_grid_avg > 0
    if bool(positive.any()):                                                             # :323-326
        threshold = _percentile_of(voxel_grid_avg[positive], brightness_threshold_percentile)
    else:
        threshold = 0
    object_voxels = voxel_grid_avg > float(threshold)                                    # :328
    idx = torch.nonzero(object_voxels)                                                   # :329 (C order, like np.nonzero)
    nx, ny, nz = voxel_grid_avg.shape                                                    # :331
    (x_min, x_max), (y_min, y_max), (z_min, z_max) = voxel_grid_extent                   # :332-334
    x_coords = idx[:, 0].double() / dev(nx) * (x_max - x_min) + x_min                    # :336-338 (numpy: int64 / int -> float64)
    y_coords = idx[:, 1].double() / dev(ny) * (y_max - y_min) + y_min
    z_coords = idx[:, 2].double() / dev(nz) * (z_max - z_min) + z_min
    intensities = voxel_grid_avg[object_voxels]                                          # :340
    brightest = None
    if intensities.numel() > 0:                                                          # :343-347
        b = int(torch.argmax(intensities))
        brightest = (float(x_coords[b]), float(y_coords[b]), float(z_coords[b]))
    return {"voxel_grid_avg": voxel_grid_avg, "threshold": threshold, "x_coords": x_coords, "y_coords": y_coords, "z_coords": z_coords,
            "intensities": intensities, "brightest": brightest}
