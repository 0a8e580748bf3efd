"""Synthetic inputs of the shapes BASELINE.json names: seeded frame sequences, camera
pose tables, metadata.json + PGM folders.  Used by tests/ and bench.py; nothing here is on the hot path."""
from __future__ import annotations

import json
import os

import numpy as np

from .motion import FramePose


def write_pgm(path: str, img: np.ndarray) -> None:
    img = np.ascontiguousarray(img, np.uint8)
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())


def write_ppm(path: str, img: np.ndarray) -> None:
    img = np.ascontiguousarray(img, np.uint8)
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())


def dense_frames(n_frames: int, height: int, width: int, seed: int = 0) -> np.ndarray:
    """iid uint8 frames: ~98 % of the pixels change between consecutive frames (threshold 2)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n_frames, height, width), dtype=np.uint8)


def sparse_frames(n_frames: int, height: int, width: int, seed: int = 0, blob: int | None = None, blobs: int = 1) -> np.ndarray:
    """Static noise background (uint8 20..60) with `blobs` moving bright squares (value 220)."""
    rng = np.random.default_rng(seed)
    bg = rng.integers(20, 61, (height, width), dtype=np.uint8)
    blob = blob or max(2, height // 24)
    frames = np.repeat(bg[None], n_frames, 0)
    for b in range(blobs):
        x0, y0 = rng.integers(0, max(1, width - blob)), rng.integers(0, max(1, height - blob))
        vx, vy = rng.integers(-7, 8), rng.integers(-5, 6)
        for f in range(n_frames):
            x = int((x0 + f * vx) % max(1, width - blob))
            y = int((y0 + f * vy) % max(1, height - blob))
            frames[f, y:y + blob, x:x + blob] = 220
    return frames


def scatter_frames(n_frames: int, height: int, width: int, seed: int = 0, fraction: float = 0.10) -> np.ndarray:
    """Scattered motion: a static background in which, between consecutive frames, a fresh iid-random `fraction` of the pixels jumps by
    64 grey levels (well above the threshold) -- the motion pixels of a pair are isolated, not neighbours of each other."""
    rng = np.random.default_rng(seed)
    frames = np.empty((n_frames, height, width), np.uint8)
    frames[0] = rng.integers(40, 200, (height, width), dtype=np.uint8)
    for f in range(1, n_frames):
        hit = rng.random((height, width)) < fraction
        frames[f] = np.where(hit, frames[f - 1] + np.uint8(64), frames[f - 1])   # uint8 wrap-around: |diff| is 64 or 192
    return frames


def object_frames(n_frames: int, height: int, width: int, seed: int = 0) -> np.ndarray:
    """A moving textured object at a realistic density: a static noise background and one textured disc (radius height/7, ~3.6 % of a
    16:9 frame) drifting a few pixels per frame; its whole area changes between frames (the texture moves with it), the rest is still."""
    rng = np.random.default_rng(seed)
    bg = rng.integers(20, 61, (height, width), dtype=np.uint8)
    r = max(2, height // 7)
    tex = rng.integers(90, 256, (2 * r + 1, 2 * r + 1), dtype=np.uint8)
    yy, xx = np.mgrid[-r:r + 1, -r:r + 1]
    disc = (yy * yy + xx * xx) <= r * r
    frames = np.repeat(bg[None], n_frames, 0)
    x0, y0 = rng.integers(r, max(r + 1, width - r)), rng.integers(r, max(r + 1, height - r))
    vx, vy = int(rng.integers(3, 8)), int(rng.integers(2, 6))
    for f in range(n_frames):
        cx = r + (x0 - r + f * vx) % max(1, width - 2 * r)
        cy = r + (y0 - r + f * vy) % max(1, height - 2 * r)
        ys, xs = slice(cy - r, cy + r + 1), slice(cx - r, cx + r + 1)
        h, w = frames[f, ys, xs].shape
        frames[f, ys, xs] = np.where(disc[:h, :w], tex[:h, :w], frames[f, ys, xs])
    return frames


def orbit_poses(n_frames: int, n: int, voxel_size: float, center, fov_degrees: float = 60.0, seed: int = 0,
                camera_index: int = 0, radius_frac: float = 0.35) -> list[FramePose]:
    """A camera moving INSIDE the grid (rays entering through a max face are dropped by the reference,
    ray_voxel.cpp:329-334, so a useful pose sits inside or looks in through a min face).  The pose changes
    every frame (slow orbit + jitter) so no two pairs share a ray set."""
    rng = np.random.default_rng(seed)
    half = 0.5 * n * voxel_size
    out = []
    for f in range(n_frames):
        ang = 2 * np.pi * f / max(8, n_frames) + 0.3 * camera_index
        pos = (center[0] + radius_frac * half * np.cos(ang), center[1] + radius_frac * half * np.sin(ang),
               center[2] + half * 0.6 + 0.05 * half * np.sin(3 * ang))
        out.append(FramePose(tuple(float(np.float32(p)) for p in pos), yaw=float(np.float32(rng.uniform(-25, 25))),
                             pitch=float(np.float32(rng.uniform(-12, 12))), roll=float(np.float32(rng.uniform(-12, 12))),
                             fov_degrees=fov_degrees, camera_index=camera_index, frame_index=f, image_file=f"cam{camera_index}_{f:05d}.pgm"))
    return out


def write_sequence(folder: str, frames_by_cam: dict[int, np.ndarray], poses_by_cam: dict[int, list[FramePose]],
                   shuffle_seed: int | None = None) -> str:
    """metadata.json + one PGM per frame, the inputs of the ray_voxel CLI (ray_voxel.cpp:401-407, :155-175)."""
    os.makedirs(folder, exist_ok=True)
    entries = []
    for cam, frames in frames_by_cam.items():
        for f, pose in enumerate(poses_by_cam[cam]):
            write_pgm(os.path.join(folder, pose.image_file), frames[f])
            entries.append({"camera_index": cam, "frame_index": pose.frame_index, "camera_position": list(pose.camera_position),
                            "yaw": pose.yaw, "pitch": pose.pitch, "roll": pose.roll, "fov_degrees": pose.fov_degrees,
                            "image_file": pose.image_file})
    if shuffle_seed is not None:
        np.random.default_rng(shuffle_seed).shuffle(entries)
    meta = os.path.join(folder, "metadata.json")
    with open(meta, "w") as f:
        json.dump(entries, f)
    return meta
