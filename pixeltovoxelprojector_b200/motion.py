"""Path A host side: the motion pipeline of ray_voxel.cpp driven from Python with torch tensors.

Everything here is plumbing above the C ABI (include/pvp_b200.h): torch owns device memory and
streams, libpvp_b200.so does the work.  Names follow the reference (ray_voxel.cpp): frames,
pairs, poses, voxel grid, .bin.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
from typing import Iterable, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import ACCUM_F32, ACCUM_FIXED, FRAME_F32, FRAME_U8, GridDesc, Pair, PvpError, Stats, check


# ---------------------------------------------------------------------------------------------
# context
# ---------------------------------------------------------------------------------------------
class Context:
    """A libpvp_b200 context: one CUDA device, one stream, the ray-queue workspace."""

    def __init__(self, device: int | str | torch.device | None = None, stream: torch.cuda.Stream | None = None):
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("pixeltovoxelprojector_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise ValueError(f"device must be a CUDA device, got {dev}")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self._lib = lib
        self._stream_handle = None
        h = C.c_void_p()
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        check(lib.pvp_ctx_create(self.device.index, C.c_void_p(self._handle_of(s)), C.byref(h)))
        self.handle = h
        self._stream_handle = self._handle_of(s)
        self._pinned_stream = stream is not None

    @staticmethod
    def _handle_of(stream: torch.cuda.Stream) -> int:
        # torch's default stream is the legacy default stream, whose handle is 0 == NULL; the C ABI reads
        # NULL as "make a private stream", so name the legacy stream explicitly (cudaStreamLegacy == 0x1).
        return stream.cuda_stream or 1

    def use_current_stream(self) -> None:
        """Enqueue on torch's current stream (unless a stream was pinned at construction)."""
        if self._pinned_stream:
            return
        s = self._handle_of(torch.cuda.current_stream(self.device))
        if s != self._stream_handle:
            check(self._lib.pvp_ctx_set_stream(self.handle, C.c_void_p(s)))
            self._stream_handle = s

    def set_option(self, key: str, value: int) -> None:
        check(self._lib.pvp_ctx_set_option(self.handle, key.encode(), int(value)))

    def profile(self, enable: bool = True) -> None:
        """Bracket every K1/K2/K3 launch with CUDA events on the context stream (live kernel timing)."""
        check(self._lib.pvp_ctx_profile(self.handle, int(enable)))

    def get_profile(self, reset: bool = True) -> dict:
        p = _lib.Profile()
        check(self._lib.pvp_ctx_get_profile(self.handle, C.byref(p), int(reset)))
        return p.as_dict()

    def last_k2_pick(self) -> int:
        """Which lock-step kernel took the last K2 batch: 1 lean, 2 lean2, 4 lean4 (0: another variant, or nothing ran)."""
        v = C.c_int(0)
        check(self._lib.pvp_ctx_last_k2_pick(self.handle, C.byref(v)))
        return int(v.value)

    def sync(self) -> None:
        check(self._lib.pvp_ctx_sync(self.handle))

    def close(self) -> None:
        if getattr(self, "handle", None) is not None and self.handle:
            self._lib.pvp_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: dict[int, Context] = {}


def default_context(device: torch.device | int | None = None) -> Context:
    idx = torch.cuda.current_device() if device is None else torch.device("cuda", device).index if isinstance(device, int) else (
        torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device())
    if idx not in _default_ctx:
        _default_ctx[idx] = Context(idx)
    return _default_ctx[idx]


# ---------------------------------------------------------------------------------------------
# pose and pairs
# ---------------------------------------------------------------------------------------------
def rotation_matrix(yaw_deg: float, pitch_deg: float, roll_deg: float) -> np.ndarray:
    """rotation_matrix_yaw_pitch_roll (ray_voxel.cpp:84-135): Rz(yaw)*Ry(roll)*Rx(pitch), fp32, row major."""
    R = (C.c_float * 9)()
    _lib.load().pvp_rotation_matrix(C.c_float(yaw_deg), C.c_float(pitch_deg), C.c_float(roll_deg), R)
    return np.frombuffer(R, np.float32).copy()


def focal_length(fov_degrees: float, width: int) -> np.float32:
    """(width*0.5f)/tanf(deg2rad(fov)*0.5f) (ray_voxel.cpp:490-491)."""
    return np.float32(_lib.load().pvp_focal_length(C.c_float(fov_degrees), int(width)))


@dataclasses.dataclass
class FramePose:
    """Per-frame record of metadata.json (FrameInfo, ray_voxel.cpp:47-55)."""
    camera_position: Sequence[float]
    yaw: float = 0.0
    pitch: float = 0.0
    roll: float = 0.0
    fov_degrees: float = 60.0
    camera_index: int = 0
    frame_index: int = 0
    image_file: str = ""


def make_pairs(poses: Sequence[FramePose], width: int, pairs: Iterable[tuple[int, int]] | None = None):
    """Build the C pair table.  Pair (p, c) uses the pose of frame c (ray_voxel.cpp:486-491).
    Default pairing: consecutive frames (0,1), (1,2), ..."""
    idx = list(pairs) if pairs is not None else [(i - 1, i) for i in range(1, len(poses))]
    arr = (Pair * max(1, len(idx)))()
    for k, (p, c) in enumerate(idx):
        pose = poses[c]
        arr[k].cam[:] = [np.float32(x) for x in pose.camera_position]
        arr[k].R[:] = rotation_matrix(pose.yaw, pose.pitch, pose.roll).tolist()
        arr[k].focal = focal_length(pose.fov_degrees, width)
        arr[k].prev, arr[k].curr = int(p), int(c)
    return arr, len(idx)


# ---------------------------------------------------------------------------------------------
# voxel grid
# ---------------------------------------------------------------------------------------------
class VoxelGrid:
    """The N^3 grid of ray_voxel.cpp:434-440 on the device.  Storage [x][y][z], z fastest (:527).

    accum="f32": float32 cells, reference arithmetic.  accum="fixed": int64 cells holding
    value * 2**frac_bits -- deterministic, order independent.  ``slab=(lo, hi)`` stores only the
    x-planes [lo, hi) (slab sharding, SURVEY 8e mode 2).
    """

    def __init__(self, n: int = 500, voxel_size: float = 6.0, center: Sequence[float] = (0.0, 0.0, 500.0), accum: str = "f32",
                 frac_bits: int = 16, slab: tuple[int, int] | None = None, device: torch.device | str | int | None = None,
                 data: torch.Tensor | None = None, attenuation_alpha: float = 0.0):
        if accum not in ("f32", "fixed"):
            raise ValueError("accum must be 'f32' or 'fixed'")
        self.n, self.voxel_size, self.center = int(n), float(voxel_size), tuple(float(c) for c in center)
        self.accum, self.frac_bits = accum, int(frac_bits)
        self.attenuation_alpha = float(attenuation_alpha)   # 0 = reference behaviour (ray_voxel.cpp:526 discards the factor)
        self.slab = (0, self.n) if slab is None else (int(slab[0]), int(slab[1]))
        if not (0 <= self.slab[0] <= self.slab[1] <= self.n):
            raise ValueError(f"bad slab {self.slab} for n={self.n}")
        dtype = torch.float32 if accum == "f32" else torch.int64
        shape = (self.slab[1] - self.slab[0], self.n, self.n)
        if data is None:
            dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
            data = torch.zeros(shape, dtype=dtype, device=dev)
        if tuple(data.shape) != shape or data.dtype != dtype or not data.is_cuda or not data.is_contiguous():
            raise ValueError(f"grid data must be a contiguous CUDA {dtype} tensor of shape {shape}")
        self.data = data

    @property
    def desc(self) -> GridDesc:
        d = GridDesc()
        d.n, d.voxel_size = self.n, self.voxel_size
        d.center[:] = self.center
        d.accum_mode = ACCUM_F32 if self.accum == "f32" else ACCUM_FIXED
        d.frac_bits, d.slab_lo, d.slab_hi = self.frac_bits, self.slab[0], self.slab[1]
        d.attenuation_alpha = self.attenuation_alpha
        return d

    def zero_(self) -> "VoxelGrid":
        self.data.zero_()
        return self

    def to_float32(self, ctx: Context | None = None) -> torch.Tensor:
        """The .bin payload: float32 cells (fixed point is converted with one rounding)."""
        if self.accum == "f32":
            return self.data
        ctx = ctx or default_context(self.data.device)
        ctx.use_current_stream()
        out = torch.empty(self.data.shape, dtype=torch.float32, device=self.data.device)
        check(ctx._lib.pvp_fixed_to_f32(ctx.handle, self.data.data_ptr(), out.data_ptr(), self.data.numel(), self.frac_bits))
        return out

    def save(self, path: str, ctx: Context | None = None) -> None:
        """Write the reference's .bin (ray_voxel.cpp:542-555): int32 N, float32 voxel_size, N^3 float32."""
        if self.slab != (0, self.n):
            raise ValueError("save() needs the whole grid; gather the slabs first")
        host = self.to_float32(ctx).cpu().contiguous().numpy()
        write_bin(path, host, self.voxel_size)


def write_bin(path: str, grid: np.ndarray, voxel_size: float) -> None:
    g = np.ascontiguousarray(grid, np.float32)
    assert g.ndim == 3 and g.shape[0] == g.shape[1] == g.shape[2]
    check(_lib.load().pvp_write_bin(os.fsencode(path), g.shape[0], C.c_float(voxel_size), g.ctypes.data))


def read_bin(path: str) -> tuple[np.ndarray, np.float32]:
    """Reader of the .bin layout (the contract of voxelmotionviewer.py:28-53)."""
    lib = _lib.load()
    n, vs = C.c_int32(), C.c_float()
    check(lib.pvp_read_bin_header(os.fsencode(path), C.byref(n), C.byref(vs)))
    out = np.empty((n.value,) * 3, np.float32)
    check(lib.pvp_read_bin(os.fsencode(path), n.value, out.ctypes.data))
    return out, np.float32(vs.value)


# ---------------------------------------------------------------------------------------------
# the tensor-level entry (boundary B-3)
# ---------------------------------------------------------------------------------------------
def _frame_dtype(t) -> int:
    dt = t.dtype
    if dt in (torch.uint8, np.uint8):
        return FRAME_U8
    if dt in (torch.float32, np.float32):
        return FRAME_F32
    raise TypeError(f"frames must be uint8 or float32, got {dt}")


def accumulate_motion(frames, pairs, grid: VoxelGrid, threshold: float = 2.0, ctx: Context | None = None,
                      n_pairs: int | None = None, want_stats: bool = True) -> dict | None:
    """ray_voxel.cpp:484-531 for a batch of frame pairs.

    frames: [F, H, W] uint8 or float32.  A CUDA tensor is consumed in place; a CPU tensor or numpy
            array goes through the staged host path (H2D copies overlapped with the kernels).
    pairs:  a ctypes Pair array from :func:`make_pairs` (``n_pairs`` entries).
    grid:   accumulated in place.
    Returns the counters {n_rays, n_steps, n_written, n_atomics} (forces a stream sync) or None.
    """
    ctx = ctx or default_context(grid.data.device)
    ctx.use_current_stream()
    lib = ctx._lib
    if n_pairs is None:
        n_pairs = len(pairs)
    if frames.ndim != 3:
        raise ValueError("frames must be [F, H, W]")
    F, H, W = (int(x) for x in frames.shape)
    st = Stats()
    st_ref = C.byref(st) if want_stats else None
    desc = grid.desc
    if isinstance(frames, torch.Tensor) and frames.is_cuda:
        if frames.stride(2) != 1 or frames.stride(1) != W:
            raise ValueError("each frame must be contiguous")
        check(lib.pvp_motion_accumulate(ctx.handle, frames.data_ptr(), _frame_dtype(frames), frames.stride(0), W, H, F, pairs, n_pairs,
                                        C.c_float(threshold), grid.data.data_ptr(), C.byref(desc), st_ref))
    else:
        if isinstance(frames, torch.Tensor):
            if not frames.is_contiguous():
                frames = frames.contiguous()
            ptr, dt = frames.data_ptr(), _frame_dtype(frames)
        else:
            frames = np.ascontiguousarray(frames)
            ptr, dt = frames.ctypes.data, _frame_dtype(frames)
        check(lib.pvp_motion_accumulate_host(ctx.handle, ptr, dt, H * W, W, H, F, pairs, n_pairs, C.c_float(threshold),
                                             grid.data.data_ptr(), C.byref(desc), st_ref))
    return st.as_dict() if want_stats else None


# ---------------------------------------------------------------------------------------------
# stage-level entries (parity hooks)
# ---------------------------------------------------------------------------------------------
def detect_motion(prev: torch.Tensor, nxt: torch.Tensor, threshold: float, ctx: Context | None = None):
    """detect_motion (ray_voxel.cpp:236-255) on the device: returns (changed uint8, diff float32)."""
    ctx = ctx or default_context(prev.device)
    ctx.use_current_stream()
    if prev.shape != nxt.shape or prev.dtype != nxt.dtype:
        raise ValueError("Images differ in size. Can't do motion detection!")
    prev, nxt = prev.contiguous(), nxt.contiguous()
    changed = torch.empty(prev.shape, dtype=torch.uint8, device=prev.device)
    diff = torch.empty(prev.shape, dtype=torch.float32, device=prev.device)
    check(ctx._lib.pvp_detect_motion(ctx.handle, prev.data_ptr(), nxt.data_ptr(), _frame_dtype(prev), prev.numel(), C.c_float(threshold),
                                     changed.data_ptr(), diff.data_ptr()))
    return changed, diff


def pixel_rays(pix: torch.Tensor, width: int, height: int, pose: FramePose, ctx: Context | None = None) -> torch.Tensor:
    """Ray set-up of ray_voxel.cpp:506-515 for pixel ids pix = v*width+u (int64/int32 CUDA tensor) -> [n,3] fp32."""
    ctx = ctx or default_context(pix.device)
    ctx.use_current_stream()
    arr, _ = make_pairs([pose, pose], width)
    p32 = pix.to(torch.int64).to(torch.int32).contiguous()  # bit pattern == uint32 for ids < 2^31
    out = torch.empty((p32.numel(), 3), dtype=torch.float32, device=pix.device)
    check(ctx._lib.pvp_pixel_rays(ctx.handle, p32.data_ptr(), p32.numel(), int(width), int(height), C.byref(arr[0]), out.data_ptr()))
    return out


def cast_rays(origins: torch.Tensor, dirs: torch.Tensor, grid: VoxelGrid, want_steps: bool = True, ctx: Context | None = None):
    """cast_ray_into_grid (ray_voxel.cpp:274-395) for n rays.  Returns (counts[n] int32, offsets[n+1] int64,
    vox[total,3] int32, t[total] float32); the step lists are in visit order."""
    ctx = ctx or default_context(origins.device)
    ctx.use_current_stream()
    o = origins.to(torch.float32).contiguous()
    d = dirs.to(torch.float32).contiguous()
    n = o.shape[0]
    desc = grid.desc
    counts = torch.zeros(n, dtype=torch.int32, device=o.device)
    check(ctx._lib.pvp_cast_rays(ctx.handle, o.data_ptr(), d.data_ptr(), n, C.byref(desc), counts.data_ptr(), None, None, None))
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=o.device)
    offsets[1:] = torch.cumsum(counts.to(torch.int64), 0)
    if not want_steps:
        return counts, offsets, None, None
    total = int(offsets[-1].item())
    vox = torch.empty((max(total, 1), 3), dtype=torch.int32, device=o.device)
    t = torch.empty(max(total, 1), dtype=torch.float32, device=o.device)
    check(ctx._lib.pvp_cast_rays(ctx.handle, o.data_ptr(), d.data_ptr(), n, C.byref(desc), counts.data_ptr(), offsets.data_ptr(),
                                 vox.data_ptr(), t.data_ptr()))
    return counts, offsets, vox[:total], t[:total]


# ---------------------------------------------------------------------------------------------
# whole pipeline (boundary B-2 from Python)
# ---------------------------------------------------------------------------------------------
def load_metadata(json_path: str) -> list[FramePose]:
    """load_metadata (ray_voxel.cpp:140-178) through the library's own parser."""
    lib = _lib.load()
    arr = C.POINTER(_lib.FrameInfo)()
    n = C.c_int()
    check(lib.pvp_load_metadata(os.fsencode(json_path), C.byref(arr), C.byref(n)))
    try:
        return [FramePose(tuple(arr[i].camera_position), arr[i].yaw, arr[i].pitch, arr[i].roll, arr[i].fov_degrees,
                          arr[i].camera_index, arr[i].frame_index, arr[i].image_file.decode()) for i in range(n.value)]
    finally:
        lib.pvp_free_metadata(arr)


def load_image_gray(path: str) -> np.ndarray:
    """stbi_load(path, ..., 1) of ray_voxel.cpp:195 (no noise): uint8 [H, W]."""
    lib = _lib.load()
    w, h = C.c_int(), C.c_int()
    check(lib.pvp_load_image_gray(os.fsencode(path), None, 0, C.byref(w), C.byref(h)))
    out = np.empty((h.value, w.value), np.uint8)
    check(lib.pvp_load_image_gray(os.fsencode(path), out.ctypes.data, out.size, C.byref(w), C.byref(h)))
    return out


def ray_voxel_accumulate(metadata_json: str, image_folder: str, grid: VoxelGrid, threshold: float = 2.0, shard: tuple[int, int] = (0, 1),
                         chunk_frames: int = 0, verbose: bool = False, ctx: Context | None = None, decode_threads: int = 0,
                         loader_noise_seed: int | None = None) -> dict:
    """main's frame loop (ray_voxel.cpp:450-537) into ``grid``.  ``shard=(rank, count)`` makes this
    process handle one contiguous block of the frames (frame sharding, SURVEY 8e mode 1).  ``decode_threads`` host
    threads decode the image files ahead of the GPU (0 = $PVP_DECODE_THREADS or min(16, cores)); results do not depend on it.
    ``loader_noise_seed``: add the reference loader's per-pixel noise (ray_voxel.cpp:203-218) from std::mt19937(seed)."""
    ctx = ctx or default_context(grid.data.device)
    ctx.use_current_stream()
    opt = _lib.RunOptions()
    ctx._lib.pvp_run_options_default(C.byref(opt))
    opt.grid = grid.desc
    opt.threshold = threshold
    opt.shard_rank, opt.shard_count = int(shard[0]), int(shard[1])
    opt.chunk_frames, opt.verbose = int(chunk_frames), int(bool(verbose))
    opt.decode_threads = int(decode_threads)
    if loader_noise_seed is not None:
        opt.loader_noise, opt.noise_seed = 1, int(loader_noise_seed) & 0xffffffff
    st, rep = Stats(), _lib.RunReport()
    check(ctx._lib.pvp_ray_voxel_accumulate(ctx.handle, os.fsencode(metadata_json), os.fsencode(image_folder), grid.data.data_ptr(),
                                            C.byref(opt), C.byref(st), C.byref(rep)))
    return {**st.as_dict(), **rep.as_dict()}


def ray_voxel_main(argv: Sequence[str]) -> int:
    """The ray_voxel CLI (ray_voxel.cpp:400-558) in-process: argv[0] is the program name."""
    arr = (C.c_char_p * len(argv))(*[os.fsencode(a) for a in argv])
    return int(_lib.load().pvp_ray_voxel_main(len(argv), arr))


__all__ = ["Context", "default_context", "rotation_matrix", "focal_length", "FramePose", "make_pairs", "VoxelGrid", "write_bin",
           "read_bin", "accumulate_motion", "detect_motion", "pixel_rays", "cast_rays", "load_metadata", "load_image_gray",
           "ray_voxel_accumulate", "ray_voxel_main", "PvpError"]
