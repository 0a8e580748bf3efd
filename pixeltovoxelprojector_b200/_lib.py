"""ctypes binding of libpvp_b200.so (include/pvp_b200.h).

The library is built in-tree by ``pixeltovoxelprojector_b200/csrc/Makefile`` (nvcc, sm_100a).
There is no Python, PyTorch or CPU fallback: if the shared object is missing, importing the
compute entry points raises; if no B200 is visible, ``pvp_ctx_create`` fails.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpvp_b200.so")

FRAME_U8, FRAME_F32 = 0, 1
ACCUM_F32, ACCUM_FIXED, ACCUM_F64 = 0, 1, 0


class PvpError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libpvp_b200 error {status}: {message}")
        self.status = status


class GridDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("voxel_size", C.c_float), ("center", C.c_float * 3), ("accum_mode", C.c_int32),
                ("frac_bits", C.c_int32), ("slab_lo", C.c_int32), ("slab_hi", C.c_int32),
                ("attenuation_alpha", C.c_float)]


class Pair(C.Structure):
    _fields_ = [("cam", C.c_float * 3), ("R", C.c_float * 9), ("focal", C.c_float), ("prev", C.c_int32), ("curr", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n_rays", C.c_uint64), ("n_steps", C.c_uint64), ("n_written", C.c_uint64), ("n_atomics", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class FrameInfo(C.Structure):
    _fields_ = [("camera_index", C.c_int32), ("frame_index", C.c_int32), ("camera_position", C.c_float * 3),
                ("yaw", C.c_float), ("pitch", C.c_float), ("roll", C.c_float), ("fov_degrees", C.c_float),
                ("has_position", C.c_int32), ("image_file", C.c_char * 512)]


class RunOptions(C.Structure):
    _fields_ = [("grid", GridDesc), ("threshold", C.c_float), ("shard_rank", C.c_int32), ("shard_count", C.c_int32),
                ("chunk_frames", C.c_int32), ("verbose", C.c_int32), ("decode_threads", C.c_int32),
                ("loader_noise", C.c_int32), ("noise_seed", C.c_uint32)]


class RunReport(C.Structure):
    _fields_ = [("n_frames_listed", C.c_uint64), ("n_frames_loaded", C.c_uint64), ("n_frames_skipped", C.c_uint64),
                ("n_pairs", C.c_uint64), ("n_size_mismatch", C.c_uint64),
                ("seconds_total", C.c_double), ("seconds_decode_wait", C.c_double), ("seconds_upload", C.c_double), ("seconds_launch", C.c_double),
                ("seconds_setup", C.c_double), ("seconds_drain", C.c_double), ("decode_threads", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self):
        return {k: (float if k.startswith("seconds") else int)(getattr(self, k)) for k, _ in self._fields_ if k != "reserved"}


class ImageArgs(C.Structure):
    _fields_ = [("earth_position", C.c_double * 3), ("pointing_direction", C.c_double * 3), ("fov", C.c_double),
                ("image_width", C.c_int64), ("image_height", C.c_int64), ("image_stride", C.c_int64 * 2),
                ("grid_shape", C.c_int64 * 3), ("grid_stride", C.c_int64 * 3), ("extent", C.c_double * 6),
                ("max_distance", C.c_double), ("num_steps", C.c_int32), ("tex_shape", C.c_int64 * 2),
                ("tex_stride", C.c_int64 * 2), ("center_ra_rad", C.c_double), ("center_dec_rad", C.c_double),
                ("angular_width_rad", C.c_double), ("angular_height_rad", C.c_double),
                ("update_celestial_sphere", C.c_int32), ("perform_background_subtraction", C.c_int32),
                ("accum_mode", C.c_int32), ("frac_bits", C.c_int32),
                ("row_begin", C.c_int64), ("row_end", C.c_int64)]


class Profile(C.Structure):
    _fields_ = [("launches_k1", C.c_uint64), ("launches_k2", C.c_uint64), ("launches_k3", C.c_uint64), ("launches_other", C.c_uint64),
                ("ms_k1", C.c_double), ("ms_k2", C.c_double), ("ms_k3", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ImageStats(C.Structure):
    _fields_ = [("n_rays", C.c_uint64), ("n_samples", C.c_uint64)]


# name -> (restype, argtypes); every symbol include/pvp_b200.h declares
_P = C.c_void_p
SIGNATURES = {
    "pvp_version": (C.c_int, []),
    "pvp_build_flags": (C.c_int, []),
    "pvp_last_error": (C.c_char_p, []),
    "pvp_ctx_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "pvp_ctx_destroy": (C.c_int, [_P]),
    "pvp_ctx_set_stream": (C.c_int, [_P, _P]),
    "pvp_ctx_sync": (C.c_int, [_P]),
    "pvp_ctx_profile": (C.c_int, [_P, C.c_int]),
    "pvp_ctx_get_profile": (C.c_int, [_P, C.POINTER(Profile), C.c_int]),
    "pvp_ctx_last_k2_pick": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "pvp_ctx_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "pvp_dev_malloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "pvp_dev_free": (C.c_int, [_P, _P]),
    "pvp_dev_memset": (C.c_int, [_P, _P, C.c_int, C.c_size_t]),
    "pvp_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "pvp_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "pvp_host_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "pvp_host_free_pinned": (C.c_int, [_P]),
    "pvp_rotation_matrix": (None, [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float)]),
    "pvp_focal_length": (C.c_float, [C.c_float, C.c_int]),
    "pvp_detect_motion": (C.c_int, [_P, _P, _P, C.c_int, C.c_int64, C.c_float, _P, _P]),
    "pvp_pixel_rays": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.POINTER(Pair), _P]),
    "pvp_cast_rays": (C.c_int, [_P, _P, _P, C.c_int64, C.POINTER(GridDesc), _P, _P, _P, _P]),
    "pvp_motion_accumulate": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(Pair), C.c_int, C.c_float,
                                        _P, C.POINTER(GridDesc), C.POINTER(Stats)]),
    "pvp_motion_accumulate_host": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(Pair), C.c_int,
                                             C.c_float, _P, C.POINTER(GridDesc), C.POINTER(Stats)]),
    "pvp_load_metadata": (C.c_int, [C.c_char_p, C.POINTER(C.POINTER(FrameInfo)), C.POINTER(C.c_int)]),
    "pvp_free_metadata": (None, [C.POINTER(FrameInfo)]),
    "pvp_load_image_gray": (C.c_int, [C.c_char_p, _P, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pvp_run_options_default": (None, [C.POINTER(RunOptions)]),
    "pvp_ray_voxel_accumulate": (C.c_int, [_P, C.c_char_p, C.c_char_p, _P, C.POINTER(RunOptions), C.POINTER(Stats),
                                           C.POINTER(RunReport)]),
    "pvp_ray_voxel_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "pvp_grid_bytes": (C.c_size_t, [C.POINTER(GridDesc)]),
    "pvp_fixed_to_f32": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int]),
    "pvp_fixed_to_f64": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int]),
    "pvp_write_bin": (C.c_int, [C.c_char_p, C.c_int32, C.c_float, _P]),
    "pvp_write_bin_dev": (C.c_int, [_P, C.c_char_p, C.c_int32, C.c_float, _P]),
    "pvp_read_bin_header": (C.c_int, [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)]),
    "pvp_read_bin": (C.c_int, [C.c_char_p, C.c_int32, _P]),
    "pvp_grid_reduce_peers": (C.c_int, [_P, C.POINTER(_P), C.c_int, C.c_int, C.c_int, _P, C.c_int64, C.c_int]),
    "pvp_grid_order_stats": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "pvp_grid_extract_above": (C.c_int, [_P, _P, C.c_int32, C.c_double, C.c_double, C.POINTER(C.c_double), C.c_int64, _P, _P, _P, C.POINTER(C.c_int64)]),
    "pvp_process_image": (C.c_int, [_P, _P, _P, _P, C.POINTER(ImageArgs), C.POINTER(ImageStats)]),
    "pvp_process_image_host": (C.c_int, [_P, _P, _P, _P, C.POINTER(ImageArgs), C.POINTER(ImageStats)]),
    "pvp_selftest_div_by_extent": (C.c_int, [_P, C.c_double, C.c_int64, C.c_int32, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]),
    "pvp_atomic_microbench": (C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_float)]),
    "pvp_atomic_microbench_ops": (C.c_int64, [_P, C.c_int64]),
}

_lib = None


def load() -> C.CDLL:
    """Load libpvp_b200.so; raise loudly when it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH
    variant = os.environ.get("PVP_LIB_VARIANT", "")
    if variant:   # "checked": the same sources with -DPVP_CHECKED, every index a kernel computes is tested (make checked); other names: kernel
        # experiments built from the same sources with extra -D flags (tools/build_exp.sh), for A/B runs on one GPU box
        if not variant.replace("_", "").isalnum():
            raise ImportError(f"bad PVP_LIB_VARIANT {variant!r}")
        path = LIB_PATH.replace("libpvp_b200.so", f"libpvp_b200_{variant}.so")
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing. Build it with `make -C pixeltovoxelprojector_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError => header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        raise PvpError(status, load().pvp_last_error().decode("utf-8", "replace"))
