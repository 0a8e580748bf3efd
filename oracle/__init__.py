"""ctypes bindings for the parity oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing in pixeltovoxelprojector_b200/ imports it.

Two libraries are wrapped with the same Python surface:

* ``Oracle``  -- oracle/libpvp_oracle.so, the C restatement (pvp_oracle.c).  Travels to the GPU box.
* ``RefLib``  -- oracle/_ref/libref_ray_voxel.so, the UNMODIFIED reference ray_voxel.cpp compiled
  where it lies (oracle/Makefile target ``ref``).  Built in the container that mounts
  /root/reference; the built .so travels with the snapshot, the sources do not.
* ``ref_process_image_module()`` -- oracle/_ref/process_image_cpp*.so, the reference's own
  pybind11 extension (process_image.cpp, -O2 -fopenmp).
"""
from __future__ import annotations

import ctypes as C
import glob
import importlib.util
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libpvp_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_SO = os.path.join(REF_DIR, "libref_ray_voxel.so")
REF_CLI = os.path.join(REF_DIR, "ray_voxel_ref")
REF_CLI_NOISY = os.path.join(REF_DIR, "ray_voxel_ref_noisy")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u64ref = C.POINTER(C.c_uint64)


def build(ref: bool | None = None) -> None:
    """Compile the restatement, and the real reference when /root/reference is mounted."""
    if ref is None:
        ref = os.path.isdir("/root/reference")
    targets = ["all"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", HERE, "-s"] + targets, check=True)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _f3(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(3))


class _PathA:
    """Shared Python surface over the pose / motion / DDA entry points."""

    _prefix = ""
    _lib = None

    def rotation_matrix(self, yaw, pitch, roll) -> np.ndarray:
        R = np.zeros(9, np.float32)
        getattr(self._lib, self._prefix + "rotation_matrix")(C.c_float(yaw), C.c_float(pitch), C.c_float(roll), R)
        return R

    def focal_length(self, fov_degrees, width) -> np.float32:
        return np.float32(getattr(self._lib, self._prefix + "focal_length")(C.c_float(fov_degrees), int(width)))

    def pixel_ray(self, u, v, width, height, focal, R) -> np.ndarray:
        d = np.zeros(3, np.float32)
        getattr(self._lib, self._prefix + "pixel_ray")(int(u), int(v), int(width), int(height), C.c_float(focal),
                                                        np.ascontiguousarray(R, np.float32), d)
        return d

    def cast_ray(self, cam, direction, N, voxel_size, center, cap=None, want_t=False):
        """Visited voxel list [(ix,iy,iz)...] of one ray, in visit order (ray_voxel.cpp:274-395)."""
        fn = getattr(self._lib, self._prefix + "cast_ray")
        if cap is None:
            cap = 3 * int(N) + 8
        vox = np.zeros((cap, 3), np.int32)
        t = np.zeros(cap, np.float32)
        n = fn(_f3(cam), _f3(direction), int(N), C.c_float(voxel_size), _f3(center), vox, t, cap)
        assert n <= cap, (n, cap)
        return (vox[:n], t[:n]) if want_t else vox[:n]


class Oracle(_PathA):
    """oracle/libpvp_oracle.so (the C restatement)."""

    _prefix = "pvpo_"

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = self._lib = C.CDLL(ORACLE_SO)
        L.pvpo_rotation_matrix.argtypes = [C.c_float, C.c_float, C.c_float, _f32p]
        L.pvpo_rotation_matrix.restype = None
        L.pvpo_focal_length.argtypes = [C.c_float, C.c_int]
        L.pvpo_focal_length.restype = C.c_float
        L.pvpo_detect_motion.argtypes = [_f32p, _f32p, C.c_int64, C.c_float, _u8p, _f32p]
        L.pvpo_detect_motion.restype = None
        L.pvpo_pixel_ray.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _f32p, _f32p]
        L.pvpo_pixel_ray.restype = None
        L.pvpo_cast_ray.argtypes = [_f32p, _f32p, C.c_int, C.c_float, _f32p, _i32p, _f32p, C.c_int64]
        L.pvpo_cast_ray.restype = C.c_int64
        common = [C.c_int, C.c_int, _f32p, _f32p, C.c_float, C.c_int, C.c_float, _f32p, C.c_float,
                  C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _u64ref, _u64ref, _u64ref]
        L.pvpo_accumulate_pair.argtypes = [_f32p, _f32p] + common
        L.pvpo_accumulate_pair.restype = C.c_int
        L.pvpo_accumulate_pair_u8.argtypes = [_u8p, _u8p] + common
        L.pvpo_accumulate_pair_u8.restype = C.c_int
        L.pvpo_accumulate_pair_att.argtypes = [_f32p, _f32p, C.c_int, C.c_int, _f32p, _f32p, C.c_float, C.c_int, C.c_float, _f32p, C.c_float, C.c_float,
                                               C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _u64ref, _u64ref, _u64ref]
        L.pvpo_accumulate_pair_att.restype = C.c_int
        L.pvpo_fixed_to_f32.argtypes = [_i64p, _f32p, C.c_int64, C.c_int]
        L.pvpo_fixed_to_f32.restype = None
        L.pvpo_quantize_f64.argtypes = [C.c_double, C.c_int]
        L.pvpo_quantize_f64.restype = C.c_int64
        L.pvpo_max_threads.restype = C.c_int
        L.pvpo_process_image.argtypes = [
            C.c_void_p, C.c_int64, C.c_int64, _f64p, _f64p, C.c_double, C.c_int64, C.c_int64,
            C.c_void_p, _i64p, _i64p, _f64p, C.c_double, C.c_int,
            C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
            C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
            C.c_int, C.c_int, C.c_int, _u64ref, _u64ref]
        L.pvpo_process_image.restype = C.c_int

    def max_threads(self) -> int:
        return int(self._lib.pvpo_max_threads())

    def detect_motion(self, prev, nxt, threshold):
        prev = np.ascontiguousarray(prev, np.float32)
        nxt = np.ascontiguousarray(nxt, np.float32)
        changed = np.zeros(prev.shape, np.uint8)
        diff = np.zeros(prev.shape, np.float32)
        self._lib.pvpo_detect_motion(prev, nxt, prev.size, C.c_float(threshold), changed, diff)
        return changed, diff

    def accumulate_pair(self, prev, curr, cam, R, focal, grid, N, voxel_size, center, threshold,
                        accum_mode=0, frac_bits=16, slab=None, attenuation_alpha=0.0):
        """ray_voxel.cpp:484-531 on one frame pair; ``grid`` (f32 or i64, C order, shape
        [slab_hi-slab_lo, N, N]) is updated in place.  Returns (n_rays, n_steps, n_written).
        attenuation_alpha > 0: the opt-in extension that applies the factor of :524-525 (unpinned, see pvp_oracle.c)."""
        H, W = prev.shape
        lo, hi = (0, N) if slab is None else slab
        assert grid.flags.c_contiguous and grid.size == (hi - lo) * N * N
        assert grid.dtype == (np.float32 if accum_mode == 0 else np.int64)
        nr, ns, nw = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        if attenuation_alpha:
            a, b = np.ascontiguousarray(prev, np.float32), np.ascontiguousarray(curr, np.float32)
            rc = self._lib.pvpo_accumulate_pair_att(a, b, W, H, _f3(cam), np.ascontiguousarray(R, np.float32), C.c_float(focal), int(N),
                                                    C.c_float(voxel_size), _f3(center), C.c_float(threshold), C.c_float(attenuation_alpha),
                                                    grid.ctypes.data, int(accum_mode), int(frac_bits), int(lo), int(hi), C.byref(nr), C.byref(ns), C.byref(nw))
            assert rc == 0
            return nr.value, ns.value, nw.value
        if prev.dtype == np.uint8:
            fn, a, b = self._lib.pvpo_accumulate_pair_u8, np.ascontiguousarray(prev), np.ascontiguousarray(curr, np.uint8)
        else:
            fn = self._lib.pvpo_accumulate_pair
            a, b = np.ascontiguousarray(prev, np.float32), np.ascontiguousarray(curr, np.float32)
        rc = fn(a, b, W, H, _f3(cam), np.ascontiguousarray(R, np.float32), C.c_float(focal), int(N),
                C.c_float(voxel_size), _f3(center), C.c_float(threshold), grid.ctypes.data, int(accum_mode),
                int(frac_bits), int(lo), int(hi), C.byref(nr), C.byref(ns), C.byref(nw))
        assert rc == 0
        return nr.value, ns.value, nw.value

    def fixed_to_f32(self, g: np.ndarray, frac_bits: int) -> np.ndarray:
        out = np.zeros(g.shape, np.float32)
        self._lib.pvpo_fixed_to_f32(np.ascontiguousarray(g), out, g.size, int(frac_bits))
        return out

    def process_image(self, image, earth_position, pointing_direction, fov, image_width, image_height,
                      voxel_grid, voxel_grid_extent, max_distance, num_steps, celestial_sphere_texture,
                      center_ra_rad, center_dec_rad, angular_width_rad, angular_height_rad,
                      update_celestial_sphere, perform_background_subtraction,
                      accum_mode=0, frac_bits=32, num_threads=1):
        """process_image.cpp:125-366 with the reference's argument order; arrays are updated in place
        through their numpy strides.  Returns (n_rays, n_samples)."""
        want = np.float64 if accum_mode == 0 else np.int64
        assert image.dtype == np.float64 and celestial_sphere_texture.dtype == want
        es = 8
        if voxel_grid is None or voxel_grid.size == 0:
            gptr, gshape, gstride = None, np.zeros(3, np.int64), np.zeros(3, np.int64)
        else:
            assert voxel_grid.dtype == want and voxel_grid.ndim == 3
            gptr = voxel_grid.ctypes.data
            gshape = np.array(voxel_grid.shape, np.int64)
            gstride = np.array(voxel_grid.strides, np.int64) // es
        ext = np.ascontiguousarray(np.asarray(voxel_grid_extent, np.float64).reshape(6))
        tex = celestial_sphere_texture
        nr, ns = C.c_uint64(0), C.c_uint64(0)
        rc = self._lib.pvpo_process_image(
            image.ctypes.data, image.strides[0] // es, image.strides[1] // es,
            np.ascontiguousarray(earth_position, np.float64), np.ascontiguousarray(pointing_direction, np.float64),
            float(fov), int(image_width), int(image_height), gptr, gshape, gstride, ext, float(max_distance),
            int(num_steps), tex.ctypes.data, tex.shape[0], tex.shape[1], tex.strides[0] // es, tex.strides[1] // es,
            float(center_ra_rad), float(center_dec_rad), float(angular_width_rad), float(angular_height_rad),
            int(bool(update_celestial_sphere)), int(bool(perform_background_subtraction)),
            int(accum_mode), int(frac_bits), int(num_threads), C.byref(nr), C.byref(ns))
        assert rc == 0
        return nr.value, ns.value


class RefLib(_PathA):
    """oracle/_ref/libref_ray_voxel.so (the unmodified reference TU behind a C ABI)."""

    _prefix = "ref_"

    def __init__(self):
        if not have_ref():
            raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference is mounted")
        L = self._lib = C.CDLL(REF_SO)
        L.ref_rotation_matrix.argtypes = [C.c_float, C.c_float, C.c_float, _f32p]
        L.ref_rotation_matrix.restype = None
        L.ref_focal_length.argtypes = [C.c_float, C.c_int]
        L.ref_focal_length.restype = C.c_float
        L.ref_detect_motion.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_float, _u8p, _f32p]
        L.ref_detect_motion.restype = C.c_int
        L.ref_pixel_ray.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _f32p, _f32p]
        L.ref_pixel_ray.restype = None
        L.ref_cast_ray.argtypes = [_f32p, _f32p, C.c_int, C.c_float, _f32p, _i32p, _f32p, C.c_int64]
        L.ref_cast_ray.restype = C.c_int64
        L.ref_accumulate_pair.argtypes = [_f32p, _f32p, C.c_int, C.c_int, _f32p, _f32p, C.c_float, C.c_int,
                                          C.c_float, _f32p, C.c_float, _f32p, _u64ref, _u64ref]
        L.ref_accumulate_pair.restype = C.c_int
        L.ref_cli.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.ref_cli.restype = C.c_int

    def detect_motion(self, prev, nxt, threshold):
        prev = np.ascontiguousarray(prev, np.float32)
        nxt = np.ascontiguousarray(nxt, np.float32)
        H, W = prev.shape
        changed = np.zeros(prev.shape, np.uint8)
        diff = np.zeros(prev.shape, np.float32)
        self._lib.ref_detect_motion(prev, nxt, W, H, C.c_float(threshold), changed, diff)
        return changed, diff

    def accumulate_pair(self, prev, curr, cam, R, focal, grid, N, voxel_size, center, threshold):
        H, W = prev.shape
        assert grid.dtype == np.float32 and grid.flags.c_contiguous and grid.size == N ** 3
        a = np.ascontiguousarray(prev, np.float32)
        b = np.ascontiguousarray(curr, np.float32)
        nr, ns = C.c_uint64(0), C.c_uint64(0)
        rc = self._lib.ref_accumulate_pair(a, b, W, H, _f3(cam), np.ascontiguousarray(R, np.float32),
                                           C.c_float(focal), int(N), C.c_float(voxel_size), _f3(center),
                                           C.c_float(threshold), grid.reshape(-1), C.byref(nr), C.byref(ns))
        assert rc == 0
        return nr.value, ns.value


def ref_process_image_module():
    """Import oracle/_ref/process_image_cpp*.so (the reference's own pybind11 extension) under a private
    name so it cannot shadow the drop-in module of the same name."""
    hits = glob.glob(os.path.join(REF_DIR, "process_image_cpp*.so"))
    if not hits:
        raise FileNotFoundError("oracle/_ref/process_image_cpp*.so missing: run `make -C oracle ref`")
    spec = importlib.util.spec_from_file_location("process_image_cpp", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def parity_report(got: np.ndarray, ref: np.ndarray, exact: np.ndarray | None = None) -> dict:
    """Distance between an accumulated float32 grid (``got``: the CUDA path) and the reference's own grid (``ref``: the serial fp32
    sum of ray_voxel.cpp:523-529 from RefLib / the C port) for integer-valued addends (uint8 frames).  Below 2^24 every partial sum of
    the reference is an exact integer, so the two must agree bit for bit; from 2^24 on both sides round at every addition (the
    reference at each of its serial adds, the GPU at each RED of a warp-aggregated run sum) and only a relative distance is defined.
    ``exact``: the same grid as exact integers (fixed-point accumulation) -- then both float grids are also measured against it."""
    got, ref = np.ascontiguousarray(got, np.float32).reshape(-1), np.ascontiguousarray(ref, np.float32).reshape(-1)
    assert got.shape == ref.shape
    big = ref >= np.float32(2 ** 24)
    if exact is not None:
        ex = np.asarray(exact).reshape(-1).astype(np.float64)
        big = ex >= 2.0 ** 24
    small = ~big
    rep = {
        "cells": int(got.size), "cells_nonzero": int(np.count_nonzero(ref)),
        "support_equal": bool(np.array_equal(got != 0, ref != 0)),
        "cells_below_2p24": int(small.sum()), "bit_mismatches_below_2p24": int(np.count_nonzero(got[small].view(np.uint32) != ref[small].view(np.uint32))),
        "cells_at_or_above_2p24": int(big.sum()),
    }

    def rel(a, b):
        b = b.astype(np.float64)
        return float(np.max(np.abs(a.astype(np.float64) - b) / b)) if b.size else 0.0

    rep["max_rel_err_vs_reference_at_or_above_2p24"] = rel(got[big], ref[big])
    rep["max_value"] = float(ref.max()) if ref.size else 0.0
    if exact is not None:
        rep["gpu_max_rel_err_vs_exact"] = rel(got[big], ex[big])
        rep["reference_max_rel_err_vs_exact"] = rel(ref[big], ex[big])
        rep["reference_equals_exact_below_2p24"] = bool(np.array_equal(ref[small].astype(np.float64), ex[small]))
    return rep
