"""pixeltovoxelprojector_b200 -- B200-native (sm_100a) implementation of the Pixeltovoxelprojector hot path.

Path A: per-frame motion pipeline of ray_voxel.cpp (diff/threshold -> camera rays -> Amanatides-Woo DDA ->
grid accumulation -> .bin).  Path B: the process_image_cpp extension (fixed-step ray march).
All compute runs in libpvp_b200.so (hand-written CUDA behind the C ABI of include/pvp_b200.h).
"""
from ._lib import LIB_PATH, PvpError  # noqa: F401
from .motion import (Context, FramePose, VoxelGrid, accumulate_motion, cast_rays, default_context, detect_motion,  # noqa: F401
                     focal_length, load_image_gray, load_metadata, make_pairs, pixel_rays, ray_voxel_accumulate, ray_voxel_main,
                     read_bin, rotation_matrix, write_bin)
from .process_image import process_image_cpp  # noqa: F401
from .analysis import extract_top_percentile_z_up, grid_percentile, load_voxel_grid  # noqa: F401
from .sky import analyze_voxel_grid, field_of_view_rad, normalize_image  # noqa: F401

__version__ = "0.1.0"
