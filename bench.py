#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: voxel-steps/s and frames/s, 512^3 grid, 1080p).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the motion pipeline (K1 diff/threshold/compact + K2 DDA/accumulate) over one batch of
synthetic frame pairs.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL prints its version banner to STDOUT at NCCL_DEBUG=VERSION; the contract is ONE JSON line on stdout
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

WORKLOADS = {
    # the configuration BASELINE.json's metric is quoted on: 1080p frames into a 512^3 fp32 grid, one camera whose
    # pose changes every frame, dense motion (iid uint8 frames: ~98 % of the pixels cast a ray)
    "1080p_dense_512": dict(H=1080, W=1920, N=512, voxel_size=2.0, center=(0.0, 0.0, 500.0), pairs_per_step=64, pool=136, motion="dense", accum="f32"),
    # BASELINE config[1]: 256^3 grid (L2 resident)
    "1080p_dense_256": dict(H=1080, W=1920, N=256, voxel_size=4.0, center=(0.0, 0.0, 500.0), pairs_per_step=64, pool=136, motion="dense", accum="f32"),
    # sparse motion (moving blobs, ~1 % of the pixels): frames/s matters, many pairs per step
    "1080p_sparse_512": dict(H=1080, W=1920, N=512, voxel_size=2.0, center=(0.0, 0.0, 500.0), pairs_per_step=64, pool=72, motion="sparse", accum="f32"),
    # scattered motion: 10 % of the pixels, iid positions, change between consecutive frames -- consecutive queue entries are NOT image
    # neighbours, so run merging gets little; and a moving textured object at a realistic density (~3.6 % of the pixels, one compact region)
    "1080p_scatter10_512": dict(H=1080, W=1920, N=512, voxel_size=2.0, center=(0.0, 0.0, 500.0), pairs_per_step=64, pool=72, motion="scatter10", accum="f32"),
    "1080p_scatter30_512": dict(H=1080, W=1920, N=512, voxel_size=2.0, center=(0.0, 0.0, 500.0), pairs_per_step=64, pool=72, motion="scatter30", accum="f32"),
    "1080p_scatter60_512": dict(H=1080, W=1920, N=512, voxel_size=2.0, center=(0.0, 0.0, 500.0), pairs_per_step=64, pool=72, motion="scatter60", accum="f32"),
    "1080p_scatter2_512": dict(H=1080, W=1920, N=512, voxel_size=2.0, center=(0.0, 0.0, 500.0), pairs_per_step=64, pool=72, motion="scatter2", accum="f32"),
    "1080p_object_512": dict(H=1080, W=1920, N=512, voxel_size=2.0, center=(0.0, 0.0, 500.0), pairs_per_step=64, pool=72, motion="object", accum="f32"),
    # BASELINE config[2] shape: 3840x2160 frames into a 512^3 grid (frame-sharded over N GPUs with --gpus N)
    "4k_dense_512": dict(H=2160, W=3840, N=512, voxel_size=2.0, center=(0.0, 0.0, 500.0), pairs_per_step=2, pool=20, motion="dense", accum="f32"),
    # BASELINE config[3] shape on ONE GPU: the whole 1024^3 grid (4 GiB; slab sharding itself is covered by tests/test_gpu_configs.py)
    # BASELINE config[3]: the 1024^3 grid sharded into 8 x-slabs (x is the slowest index: contiguous byte ranges of the .bin); every rank walks
    # every ray with full-grid arithmetic and writes only inside its slab.  On one GPU this runs rank 4 of 8; with --gpus N the grid is cut into N slabs, one per rank (N = 8: config 4).
    # Same frames on every rank, no merge; `value` counts each voxel step once (the job's useful work), whatever the number of ranks.
    "1080p_dense_1024_slab8": dict(H=1080, W=1920, N=1024, voxel_size=1.0, center=(0.0, 0.0, 500.0), pairs_per_step=4, pool=24, motion="dense", accum="f32", slabs=8),
    "1080p_dense_1024": dict(H=1080, W=1920, N=1024, voxel_size=1.0, center=(0.0, 0.0, 500.0), pairs_per_step=4, pool=24, motion="dense", accum="f32"),
    "1080p_dense_768": dict(H=1080, W=1920, N=768, voxel_size=4.0 / 3.0, center=(0.0, 0.0, 500.0), pairs_per_step=4, pool=24, motion="dense", accum="f32"),
    "1080p_dense_512_fixed": dict(H=1080, W=1920, N=512, voxel_size=2.0, center=(0.0, 0.0, 500.0), pairs_per_step=4, pool=72, motion="dense", accum="fixed"),
    # BASELINE config[4]: process_image_cpp path (fp64 fixed-step march) on a 4096x4096 image into a 512^3 grid,
    # fixed-point deterministic accumulation; ~1 sample per voxel pitch
    "image_4096_512_fixed": dict(path="B", H=4096, W=4096, N=512, accum="fixed", frac_bits=20, pool=3),
    "image_4096_512_f64": dict(path="B", H=4096, W=4096, N=512, accum="f64", frac_bits=0, pool=3),
    "image_1024_256_f64": dict(path="B", H=1024, W=1024, N=256, accum="f64", frac_bits=0, pool=3),
    "image_tiny": dict(path="B", H=96, W=96, N=32, accum="f64", frac_bits=0, pool=2),
    # files -> .bin (BASELINE config[1] end to end through the reference's own formats): 1000 1080p PGM frames + metadata.json on disk
    # (page cache) -> decode pool -> pinned ring -> device ring -> K1/K2 -> 256^3 grid -> .bin file; sparse (moving blobs) and dense motion
    "files_1080p_1000": dict(path="files", H=1080, W=1920, N=256, voxel_size=4.0, center=(0.0, 0.0, 500.0), frames=1000, motion="sparse", accum="f32"),
    "files_1080p_1000_dense": dict(path="files", H=1080, W=1920, N=256, voxel_size=4.0, center=(0.0, 0.0, 500.0), frames=1000, motion="dense", accum="f32"),
    "files_tiny": dict(path="files", H=120, W=160, N=64, voxel_size=16.0, center=(0.0, 0.0, 500.0), frames=12, motion="dense", accum="f32"),
    # small smoke-sized case for CPU-side checks of this script
    "tiny": dict(H=120, W=160, N=64, voxel_size=16.0, center=(0.0, 0.0, 500.0), pairs_per_step=2, pool=8, motion="dense", accum="f32"),
}
ALGO_BYTES_PER_STEP = {"f32": 8, "fixed": 16}   # one atomic RMW: 4 B read + 4 B write (fp32), 8 + 8 (int64) -- SURVEY 8d


def make_inputs(wl, seed, pose_seed=1001):
    """Frames differ per rank (seed); the pose table is the same on every rank, so that the per-GPU work of the weak-scaling
    run is the same everywhere (ray lengths depend on the pose, not on the pixel values)."""
    from pixeltovoxelprojector_b200 import synth
    F = wl["pool"]
    if wl["motion"] == "dense":
        frames = synth.dense_frames(F, wl["H"], wl["W"], seed=seed)
    elif wl["motion"].startswith("scatter"):
        frames = synth.scatter_frames(F, wl["H"], wl["W"], seed=seed, fraction=int(wl["motion"][7:]) / 100.0)
    elif wl["motion"] == "object":
        frames = synth.object_frames(F, wl["H"], wl["W"], seed=seed)
    else:
        frames = synth.sparse_frames(F, wl["H"], wl["W"], seed=seed, blobs=6)
    poses = synth.orbit_poses(F, wl["N"], wl["voxel_size"], wl["center"], seed=pose_seed)
    return frames, poses


def step_pairs(step, wl):
    """Frame indices of step `step`: B consecutive pairs out of the pool, wrapping."""
    B, F = wl["pairs_per_step"], wl["pool"]
    start = (step * B) % (F - B)
    return start, [(start + i, start + i + 1) for i in range(B)]


class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (pynvml, 100 ms period)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._t = [], set(), None, threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_mhz_min": None, "sm_max_mhz": self.max_mhz, "samples": 0, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_mhz_min": float(np.min(self.samples)), "sm_max_mhz": self.max_mhz, "samples": len(self.samples),
                "reasons": sorted(self.reasons)}


def cpu_reference(wl, frames, poses, pairs, rows=None, want_grid=False):
    """Time the reference CPU implementation (oracle/_ref = the unmodified ray_voxel.cpp compiled; else the C port) on
    the given pairs, optionally restricted to a band of rows (pixels outside the band are made motion-free).
    Returns (voxel_steps, seconds, kind) -- with want_grid also the grid it accumulated and its ray count (for `parity`).
    The reference is serial: 1 thread."""
    import oracle
    use_ref = oracle.have_ref()
    lib = oracle.RefLib() if use_ref else oracle.Oracle()
    N = wl["N"]
    grid = np.zeros((N, N, N), np.float32)
    steps, secs, rays = 0, 0.0, 0
    for (p, c) in pairs:
        prev, curr = frames[p].astype(np.float32), frames[c].astype(np.float32)
        if rows is not None:
            keep = np.zeros(prev.shape[0], bool)
            keep[rows[0]:rows[1]] = True
            curr = np.where(keep[:, None], curr, prev)
        R = lib.rotation_matrix(poses[c].yaw, poses[c].pitch, poses[c].roll)
        f = lib.focal_length(poses[c].fov_degrees, wl["W"])
        t0 = time.perf_counter()
        if use_ref:
            nr, ns = lib.accumulate_pair(prev, curr, np.array(poses[c].camera_position, np.float32), R, f, grid, N, wl["voxel_size"], wl["center"], 2.0)
        else:
            nr, ns, _ = lib.accumulate_pair(prev, curr, np.array(poses[c].camera_position, np.float32), R, f, grid, N, wl["voxel_size"], np.array(wl["center"], np.float32), 2.0)
        secs += time.perf_counter() - t0
        steps += ns
        rays += nr
    kind = "reference" if use_ref else "port"
    return (steps, secs, kind, grid, rays) if want_grid else (steps, secs, kind)


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames, poses = make_inputs(wl, seed=1000)
    H = wl["H"]
    bands = 4   # each step = one quarter of the rows of one frame pair (bounded sample of the same workload)

    def one(step):
        _, pairs = step_pairs(step // bands, wl)
        b = step % bands
        return cpu_reference(wl, frames, poses, pairs[:1], rows=(H * b // bands, H * (b + 1) // bands))

    for s in range(args.warmup):
        one(s)
    tot_steps, tot_s, kind = 0, 0.0, "port"
    for s in range(args.steps):
        st, sec, kind = one(args.warmup + s)
        tot_steps += st
        tot_s += sec
    v = tot_steps / tot_s if tot_s > 0 else 0.0
    sample = f"per step: rows [b*{H // bands},(b+1)*{H // bands}) of one {wl['W']}x{H} frame pair, {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": "voxel_steps_per_sec", "value": v, "unit": "voxel-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, wl),
        "cpu_baseline": {"value": v, "unit": "voxel-steps/s", "cores": 1, "kind": kind, "sample": sample,
                         "note": "ray_voxel.cpp is single-threaded (no OpenMP/threads, SURVEY 2): 1 of %d host cores" % (os.cpu_count() or 1)},
        "e2e": {"value": v, "unit": "voxel-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_s": (args.steps / bands) / tot_s if tot_s > 0 else 0.0,
    }))


# ---- path B (process_image_cpp) ---------------------------------------------------------------------------------
def image_scene(wl, seed):
    """Synthetic C5 scene: rng.random() images, observer at the origin looking at a cube in front of it,
    one march sample per voxel pitch."""
    rng = np.random.default_rng(seed)
    n = wl["N"]
    images = rng.random((wl["pool"], wl["H"], wl["W"]))
    half = n / 2.0
    ext = [(-half, half), (-half, half), (100.0, 100.0 + n)]
    geom = dict(earth_position=[0.0, 0.0, 0.0], pointing_direction=[0.02, 0.01, 1.0], fov=0.8, extent=ext,
                max_distance=float(100 + n + 88), num_steps=int(100 + n + 88), sky=(0.0, 0.0, 1.0, 1.0))
    return images, geom


def image_config(args, wl):
    return {"workload": args.workload, "path": "process_image_cpp (fp64 fixed-step march)", "image": f"{wl['W']}x{wl['H']} float64",
            "grid": f"{wl['N']}^3 {'int64 fixed point, frac_bits %d' % wl['frac_bits'] if wl['accum'] == 'fixed' else 'float64'}",
            "images_per_step": 1, "parallelism": f"images sharded over {args.gpus} GPU(s), private grids, one NCCL reduce per run" if args.gpus > 1 else "1 GPU",
            "l2": "grid (1 GiB) and image (134 MB) exceed L2; 256 MiB buffer written between timed steps"}


def run_reference_image(args, wl):
    """--impl reference for path B: the unmodified process_image.cpp (oracle/_ref, -O2 -fopenmp) on all host cores,
    each step a band of image rows (bounded sample)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import oracle
    images, geom = image_scene(wl, seed=2000)
    n = wl["N"]
    use_ref = oracle.have_ref()
    mod = oracle.ref_process_image_module() if use_ref else None
    port = None if use_ref else oracle.Oracle()
    cores = os.cpu_count() or 1
    rows = max(8, wl["H"] // 128)                 # 32 rows of 4096 px: ~60 M samples per step
    stride = wl["H"] // rows                      # spread over the image so that every OpenMP thread (static row blocks) gets work
    grid = np.zeros((n, n, n), np.float64)
    tex = np.zeros((8, 8), np.float64)

    def band(step):
        img = np.zeros((wl["H"], wl["W"]))
        sel = (np.arange(rows) * stride + (step * 7) % stride) % wl["H"]
        img[sel] = images[0][sel]
        return img

    def one(step):
        img = band(step)
        t0 = time.perf_counter()
        if use_ref:
            mod.process_image_cpp(img, geom["earth_position"], geom["pointing_direction"], geom["fov"], wl["W"], wl["H"], grid, geom["extent"],
                                  geom["max_distance"], geom["num_steps"], tex, *geom["sky"], False, False)
        else:
            port.process_image(img, geom["earth_position"], geom["pointing_direction"], geom["fov"], wl["W"], wl["H"], grid, geom["extent"],
                               geom["max_distance"], geom["num_steps"], tex, *geom["sky"], False, False)
        dt = time.perf_counter() - t0
        return dt

    import pixeltovoxelprojector_b200.synth  # noqa: F401  (keeps import order identical to the path-A arm)
    for s in range(args.warmup):
        one(s)
    tot_s = sum(one(args.warmup + s) for s in range(args.steps))
    # samples are counted by the C port on the same bands (the reference module has no counter)
    o = oracle.Oracle()
    samples = 0
    for s in range(args.steps):
        img = band(args.warmup + s)
        samples += o.process_image(img, geom["earth_position"], geom["pointing_direction"], geom["fov"], wl["W"], wl["H"], np.zeros((n, n, n)), geom["extent"],
                                   geom["max_distance"], geom["num_steps"], np.zeros((8, 8)), *geom["sky"], False, False)[1]
    v = samples / tot_s if tot_s > 0 else 0.0
    kind = "reference" if use_ref else "port"
    sample = f"per step: {rows} rows (every {stride}th) of one {wl['W']}x{wl['H']} image ({samples // max(1, args.steps)} samples), {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": "voxel_steps_per_sec", "value": v, "unit": "voxel-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": image_config(args, wl),
        "cpu_baseline": {"value": v, "unit": "voxel-steps/s", "cores": cores if use_ref else 1, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "voxel-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, wl):
    return {"workload": args.workload, "k2_variant": getattr(args, "variant", 0), "frame": f"{wl['W']}x{wl['H']} uint8", "grid": f"{wl['N']}^3 {wl['accum']}", "voxel_size": wl["voxel_size"],
            "pairs_per_step": wl["pairs_per_step"], "motion": wl["motion"], "camera": "pose changes every frame (orbit inside the grid); N>1: same pose table, different frames on every rank",
            "parallelism": (f"grid sharded into {args.gpus if args.gpus > 1 else wl['slabs']} x-slabs; " + (f"{args.gpus} ranks, rank r owns slab r" if args.gpus > 1 else f"1 GPU running rank {wl['slabs'] // 2}'s slab") + f"; {getattr(args, 'reduce_how', '')}") if wl.get("slabs") else
                           f"frames sharded over {args.gpus} GPU(s), private grids, one grid merge per run inside the timed region: {getattr(args, 'reduce_how', 'NCCL reduce')}" if args.gpus > 1 else "1 GPU",
            "l2": "256 MiB buffer written between timed steps (L2 flush); frame pool 149 MB and grid > L2" if wl["N"] >= 512 else
                  "256 MiB buffer written between timed steps (L2 flush)"}



# ---- our arm ---------------------------------------------------------------------------------------------------------
class Env:
    """One process per GPU (torchrun exports RANK / LOCAL_RANK / WORLD_SIZE); the process group, the library context and the
    L2-flush buffer are made once and shared by the primary measurement and the `also` runs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import pixeltovoxelprojector_b200 as pvp
        self.torch, self.dist, self.pvp = torch, dist, pvp
        self.rank, self.world, self.local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = pvp.default_context(self.dev)
        self.args = args
        if args.variant:
            self.ctx.set_option("dda_variant", args.variant)
        for kv in args.opt or []:
            k, v = kv.split("=")
            self.ctx.set_option(k, int(v))
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.peak_gbs, self.peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        self.sms = torch.cuda.get_device_properties(self.dev).multi_processor_count

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def reduce(self, values, op="max"):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def sass_fresh(workload):
    """Does the kernel the committed ncu figures (profiles/traffic.json) were captured from still have the same SASS in the library
    that is loaded?  None when it cannot be told (no cuobjdump, no entry)."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sass_md5
        return sass_md5.check().get(workload)
    except Exception:
        return None


def rooflines(env, workload, kernel, units_per_launch, launch_ms, algo_bytes, clocks, share, pick=None):
    """`roofline` of the dominant kernel.  PRIMARY bound = the one the lock-step kernels sit on: warp-instruction issue (ncu: ~80 % of the
    issue slots busy, DRAM at 1-2 % of its peak).  achieved = warp instructions per voxel step (committed ncu capture of this very SASS,
    see sass_md5_ok) x the live rate; peak = SMs x 4 schedulers x the SM clock sampled in this run.  `algorithmic_model` keeps SURVEY 8(d)'s
    figure -- the read-modify-write bytes the reference's loop implies (8 B per voxel step, 16 B for 64-bit cells) against the measured HBM
    copy peak -- and says how much of that traffic exists at all (`traffic_ratio` = measured DRAM bytes / algorithmic bytes): warp
    aggregation issues ~0.18 REDs per voxel step and the L2 absorbs nearly all of those, so frac > 1 there is un-issued, not hidden, traffic."""
    rate = units_per_launch / (launch_ms * 1e-3) if launch_ms > 0 else 0.0
    e = {}
    try:
        e = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload) or {}
    except Exception:
        pass
    if pick and e.get("k2_pick") not in (None, pick):   # the capture belongs to another lock-step kernel than the one that ran: its figures do not apply
        e = {}
    approx_from = None
    if not e and pick:   # no capture of this workload: the capture of the SAME kernel on another workload gives its instructions per voxel step approximately
        try:
            for k, v in json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).items():
                if isinstance(v, dict) and v.get("k2_pick") == pick and "limiter" not in v:
                    e, approx_from = {kk: vv for kk, vv in v.items() if kk != "bytes_per_launch"}, k
                    break
        except Exception:
            pass
    algo_bytes_launch = algo_bytes * units_per_launch
    traffic = None
    if "bytes_per_launch" in e and e.get("voxel_steps_per_launch"):
        traffic = e["bytes_per_launch"] / e["voxel_steps_per_launch"] * units_per_launch    # the capture's DRAM bytes per unit x this run's units per launch
    model = {"bound": "hbm", "achieved": algo_bytes * rate / 1e9, "peak": env.peak_gbs, "unit": "GB/s", "frac": algo_bytes * rate / 1e9 / env.peak_gbs,
             "peak_source": env.peak_src, "algorithmic_bytes_per_voxel_step": algo_bytes, "algorithmic_bytes_per_launch": algo_bytes_launch,
             "traffic": traffic, "traffic_ratio": (traffic / algo_bytes_launch) if (traffic and algo_bytes_launch) else None,
             "note": "nominal: the RMW bytes of the reference's loop against the HBM copy peak; traffic_ratio says how little of it reaches DRAM"}
    c = clocks.summary()
    mhz = c["sm_mhz"] or c["sm_max_mhz"] or 0
    peak_issue = env.sms * 4 * float(mhz) * 1e6
    out = {"kernel": kernel, "voxel_steps_per_launch": units_per_launch, "avg_launch_ms": launch_ms, "share_of_step": share, "traffic": traffic,
           "algorithmic_model": model}
    if e.get("warp_inst_per_launch") and e.get("voxel_steps_per_launch") and peak_issue > 0:
        per_unit = e["warp_inst_per_launch"] / e["voxel_steps_per_launch"]
        out["kernel"] = e.get("kernel", kernel)
        out = {"bound": "issue", "achieved": per_unit * rate, "peak": peak_issue, "unit": "warp-instructions/s", "frac": per_unit * rate / peak_issue,
               "warp_inst_per_voxel_step": per_unit, "sass_md5_ok": sass_fresh(workload),
               "peak_source": f"{env.sms} SMs x 4 schedulers x {mhz:.0f} MHz (SM clock sampled during the timed region)",
               "source": f"{e.get('source', 'profiles/traffic.json')}: smsp__inst_executed.sum per voxel step x live rate", **out}
        if approx_from:
            out["approx_from"] = f"instructions per voxel step of the same kernel captured on {approx_from} (no capture of this workload)"
            out["sass_md5_ok"] = sass_fresh(approx_from)
        lim = e.get("limiter")
        if lim:   # ncu found another unit busier than the issue slots (e.g. the L2 / RED path on scattered motion): that is the bound
            cap_rate = e["voxel_steps_per_launch"] / (lim["launch_ms"] * 1e-3)
            out["issue"] = {k: out[k] for k in ("achieved", "peak", "unit", "frac")}
            out.update({"bound": lim["bound"], "unit": "% of peak (" + lim["ncu_metric"] + ")", "peak": 100.0, "achieved": lim["ncu_pct"] * rate / cap_rate,
                        "frac": lim["ncu_pct"] / 100.0 * rate / cap_rate, "limiter_note": lim.get("note")})
    else:   # no ncu capture committed for this workload: only the nominal model can be stated
        out = {"bound": "hbm", "achieved": model["achieved"], "peak": model["peak"], "unit": "GB/s", "frac": model["frac"], **out}
    return out


def measure_image(env, args, name, steps, warmup, full):
    """Path B (process_image_cpp): one image per step."""
    torch, dist, pvp, ctx, dev, world, rank = env.torch, env.dist, env.pvp, env.ctx, env.dev, env.world, env.rank
    from pixeltovoxelprojector_b200 import process_image as pi
    wl = WORKLOADS[name]
    images_np, geom = image_scene(wl, seed=2000 + rank)
    n, H, W = wl["N"], wl["H"], wl["W"]
    cell = torch.int64 if wl["accum"] == "fixed" else torch.float64
    images_dev = torch.as_tensor(images_np, device=dev)
    images_pin = torch.as_tensor(images_np).pin_memory()
    stage = torch.empty((H, W), dtype=torch.float64, device=dev)
    symm = None
    if world > 1 and args.reduce == "peers":
        try:
            import torch.distributed._symmetric_memory as symm_mem
            grid = symm_mem.empty((n, n, n), dtype=cell, device=dev)
            grid.zero_()
            symm = symm_mem.rendezvous(grid, dist.group.WORLD)
        except Exception as e:
            print(f"bench.py: symmetric memory unavailable ({e!r}); using the NCCL reduce", file=sys.stderr)
            symm = None
    if symm is None:
        grid = torch.zeros((n, n, n), dtype=cell, device=dev)
    tex = torch.zeros((8, 8), dtype=cell, device=dev)
    flush = env.flush
    kw = dict(accum_mode=wl["accum"], frac_bits=wl["frac_bits"], ctx=ctx)

    def call(img, want_stats):
        pvp.process_image_cpp(img, geom["earth_position"], geom["pointing_direction"], geom["fov"], W, H, grid, geom["extent"], geom["max_distance"],
                              geom["num_steps"], tex, *geom["sky"], False, False, want_stats=want_stats, **kw)
        return pi.last_stats if want_stats else None

    for s in range(warmup):
        call(images_dev[s % wl["pool"]], False)
    samples_per_image = [call(images_dev[i], True)["n_samples"] for i in range(wl["pool"])]
    grid.zero_()
    ctx.get_profile(reset=True)
    ctx.profile(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps + 1)]
    sampler = ClockSampler(env.local)   # NVML start-up (100s of ms, different on every rank) before the barrier, not between it and the timed steps
    env.barrier()
    my_samples = 0
    with sampler as clocks:
        for s in range(steps):
            flush.fill_(s & 0xFF)
            ev[s][0].record()
            call(images_dev[s % wl["pool"]], False)
            ev[s][1].record()
            my_samples += samples_per_image[s % wl["pool"]]
        ev[-1][0].record()
        if world > 1:
            if symm is not None and wl["accum"] == "fixed":
                import ctypes as C
                from pixeltovoxelprojector_b200._lib import check
                ptrs = (C.c_void_p * world)(*[int(p) for p in symm.buffer_ptrs])
                symm.barrier(channel=0)
                mc = int(symm.multicast_ptr) if (symm.multicast_ptr and world > 2) else None
                check(ctx._lib.pvp_grid_reduce_peers(ctx.handle, ptrs, world, rank, 0 if world == 2 else -1, mc, grid.numel(), 1))
                symm.barrier(channel=1)
            else:
                dist.reduce(grid, dst=0)
        ev[-1][1].record()
        env.barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    prof = ctx.get_profile(reset=True)
    ctx.profile(False)
    # e2e: the image of every step comes from pinned HOST memory (H2D inside), the step's counters go back (D2H)
    grid.zero_()
    e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e_samples = 0
    env.barrier()
    # double-buffered staging, as a host loop over image files would do it: the H2D copy of image s+1 is enqueued on a copy
    # stream at the start of step s and overlaps step s's kernel; every copy starts inside a timed interval
    stages = [stage, torch.empty_like(stage)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    main = torch.cuda.current_stream(dev)

    def prefetch(step):
        b = step % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[b])
            stages[b].copy_(images_pin[step % wl["pool"]], non_blocking=True)
            ready[b].record(copy_stream)
    for b in range(2):
        freed[b].record(main)
    for s in range(steps):
        flush.fill_(s & 0xFF)
        e_ev[s][0].record()
        if s == 0:
            prefetch(0)
        if s + 1 < steps:
            prefetch(s + 1)
        main.wait_event(ready[s % 2])
        e_samples += call(stages[s % 2], True)["n_samples"]
        freed[s % 2].record(main)
        e_ev[s][1].record()
    env.barrier()
    t = env.reduce([ms, sum(a.elapsed_time(b) for a, b in e_ev)], "max")
    tot = env.reduce([my_samples, e_samples], "sum")
    del grid, images_dev, stages, stage
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    launches = max(1, int(prof["launches_k3"]))
    roofline = rooflines(env, name, "process_image_lockstepN_kernel (K3)", my_samples / launches, prof["ms_k3"] / launches, 16, clocks,
                         prof["ms_k3"] / ms if ms > 0 else None)
    if not full or args.no_cpu:
        cpu = {"value": 0.0, "unit": "voxel-steps/s", "cores": 0, "kind": "skipped", "sample": ""}
    else:
        import contextlib
        import io
        buf = io.StringIO()
        a2 = argparse.Namespace(**{**vars(args), "steps": 3, "warmup": 1, "workload": name})
        with contextlib.redirect_stdout(buf):
            run_reference_image(a2, wl)
        cpu = json.loads(buf.getvalue().strip().splitlines()[-1])["cpu_baseline"]
    a2 = argparse.Namespace(**{**vars(args), "workload": name})
    return {
        "metric": "voxel_steps_per_sec", "value": tot[0] / (t[0] * 1e-3), "unit": "voxel-steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": t[0] / steps, "timed_region_s": t[0] * 1e-3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": image_config(a2, wl), "frames_per_s": steps * world / (t[0] * 1e-3),
        "e2e": {"value": tot[1] / (t[1] * 1e-3), "unit": "voxel-steps/s", "h2d_bytes_per_step": H * W * 8, "d2h_bytes_per_step": 16,
                "frames_per_s": steps * world / (t[1] * 1e-3),
                "api": "pixeltovoxelprojector_b200.process_image_cpp(CUDA tensors; image staged from pinned host memory, double buffered) -> pvp_process_image"},
        "gpu_launches": int(prof["launches_k3"]), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks.summary(),
    }


def parity_of_pair(env, wl, frames_dev, poses, pair, ref_grid, ref_rays, ref_steps, pick=0):
    """The pair the cpu_baseline leg has just run through the reference: the same pair on the GPU (float32 atomics and fixed point),
    compared cell by cell (oracle.parity_report; this is the checker at work inside the cpu_baseline leg, not a product path).
    pick: the lock-step kernel the timed region's batches ran on (Context.last_k2_pick) -- the float32 grid of this single pair is made
    by that same kernel (a one-pair batch alone would fall below lean4's ray threshold and go to lean2)."""
    import oracle
    pvp, ctx = env.pvp, env.ctx
    tab, _ = pvp.make_pairs(poses, wl["W"], pairs=[pair])
    g32 = pvp.VoxelGrid(wl["N"], wl["voxel_size"], wl["center"], device=env.dev)
    forced = {1: 4, 2: 5, 4: 6}.get(pick, 0) if not getattr(env.args, "variant", 0) else 0
    if forced:
        ctx.set_option("dda_variant", forced)
    try:
        st = pvp.accumulate_motion(frames_dev, tab, g32, 2.0, ctx=ctx)
    finally:
        if forced:
            ctx.set_option("dda_variant", 0)
    gfx = pvp.VoxelGrid(wl["N"], wl["voxel_size"], wl["center"], accum="fixed", frac_bits=8, device=env.dev)
    stx = pvp.accumulate_motion(frames_dev, tab, gfx, 2.0, ctx=ctx)
    rep = oracle.parity_report(g32.data.cpu().numpy(), ref_grid, gfx.data.cpu().numpy() >> 8)
    rep["counters_equal"] = bool((st["n_rays"], st["n_steps"]) == (ref_rays, ref_steps) == (stx["n_rays"], stx["n_steps"]))
    rep["k2_kernel"] = {1: "lean", 2: "lean2", 4: "lean4"}.get(pick, "as configured")
    rep["against"] = "unmodified ray_voxel.cpp (oracle/_ref), one frame pair of the workload, serial fp32 sum; exact = the GPU's int64 fixed-point grid"
    rep["fixed_point_equals_reference_below_2p24"] = rep.pop("reference_equals_exact_below_2p24")
    return rep


def measure_motion(env, args, name, steps, warmup, full):
    """Path A (K1 diff/threshold/compact + K2 DDA/accumulate): `pairs_per_step` frame pairs per step."""
    torch, dist, pvp, ctx, dev, world, rank = env.torch, env.dist, env.pvp, env.ctx, env.dev, env.world, env.rank
    wl = WORKLOADS[name]
    a2 = argparse.Namespace(**{**vars(args), "workload": name})
    slabs = wl.get("slabs")
    frames_np, poses = make_inputs(wl, seed=1000 + (0 if slabs else rank))   # weak scaling: every rank has its own frames, same shape (slab mode: the same frames)
    frames_dev = torch.as_tensor(frames_np, device=dev)
    frames_pin = torch.as_tensor(frames_np).pin_memory()
    grid, reduce_how = None, "none (1 GPU)"
    if slabs:
        from pixeltovoxelprojector_b200.distributed import slab_range
        n_slabs = world if world > 1 else slabs     # --gpus N: N slabs, one per rank (N = 8 is BASELINE config 4); 1 GPU: rank 4 of 8
        one_gpu_rank = int(os.environ.get("PVP_BENCH_SLAB_RANK", slabs // 2))       # which of the slabs a 1-GPU run plays
        lo, hi = slab_range(wl["N"], rank if world > 1 else one_gpu_rank, n_slabs)
        if os.environ.get("PVP_BENCH_SLAB_BALANCE", "1") != "0":   # slabs of equal estimated work (pilot over a sample of the pose table, same on every rank)
            from pixeltovoxelprojector_b200.distributed import balanced_slabs, estimate_plane_work
            sample = [(i, i + 1) for i in range(0, wl["pool"] - 1, max(1, (wl["pool"] - 1) // 8))]
            work = estimate_plane_work(frames_dev, poses, wl["N"], wl["voxel_size"], wl["center"], pairs=sample, ctx=ctx)
            cut = balanced_slabs(work, n_slabs)
            if world > 1:    # one correction from measured times: every rank runs the sample on its slab, the times are gathered, the planes re-cut
                from pixeltovoxelprojector_b200.distributed import refine_slabs_by_time
                tab_s, n_s = pvp.make_pairs(poses, wl["W"], pairs=sample)
                trial = pvp.VoxelGrid(wl["N"], wl["voxel_size"], wl["center"], accum=wl["accum"], device=dev, slab=cut[rank])
                for rep in range(2):
                    a_ev, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a_ev.record()
                    pvp.accumulate_motion(frames_dev, tab_s, trial, 2.0, ctx=ctx, want_stats=False)
                    b_ev.record()
                    torch.cuda.synchronize(dev)
                times = torch.zeros(world, dtype=torch.float64, device=dev)
                times[rank] = a_ev.elapsed_time(b_ev)
                dist.all_reduce(times)
                del trial
                torch.cuda.empty_cache()
                cut = refine_slabs_by_time(cut, work, times.tolist())
            lo, hi = cut[rank if world > 1 else one_gpu_rank]
        grid = pvp.VoxelGrid(wl["N"], wl["voxel_size"], wl["center"], accum=wl["accum"], device=dev, slab=(lo, hi))
        reduce_how = f"none (slab sharding: rank writes x-planes [{lo}, {hi}) of the .bin)"
    if world > 1 and not slabs:
        from pixeltovoxelprojector_b200 import distributed as pd
        if args.reduce == "peers":
            try:   # private grid in NVLink symmetric memory: the merge is our own kernel (multimem.ld_reduce over NVSwitch)
                grid = pd.symmetric_grid(wl["N"], wl["voxel_size"], wl["center"], accum=wl["accum"], device=dev)
                reduce_how = "pvp_grid_reduce_peers (own kernel, %s)" % ("NVSwitch multicast ld_reduce + multicast store, sum on every rank" if (grid.symm.multicast_ptr and world > 2) else "NVLink peer loads into rank 0")
            except Exception as e:
                print(f"bench.py: symmetric memory unavailable ({e!r}); using the NCCL reduce", file=sys.stderr)
        if grid is None:
            reduce_how = "NCCL reduce"
    if grid is None:
        grid = pvp.VoxelGrid(wl["N"], wl["voxel_size"], wl["center"], accum=wl["accum"], device=dev)
    a2.reduce_how = reduce_how
    flush = env.flush
    B, H, W = wl["pairs_per_step"], wl["H"], wl["W"]
    pair_tabs = {}

    def pairs_for(step, rebase=False):
        start, idx = step_pairs(step, wl)
        key = (start, rebase)
        if key not in pair_tabs:
            if rebase:
                pair_tabs[key] = pvp.make_pairs(poses[start:start + B + 1], W)[0]
            else:
                pair_tabs[key] = pvp.make_pairs(poses, W, pairs=idx)[0]
        return start, pair_tabs[key]

    def step_device(step):
        _, tab = pairs_for(step)
        pvp.accumulate_motion(frames_dev, tab, grid, 2.0, ctx=ctx, n_pairs=B, want_stats=False)

    def step_e2e(step):
        start, tab = pairs_for(step, rebase=True)
        return pvp.accumulate_motion(frames_pin[start:start + B + 1], tab, grid, 2.0, ctx=ctx, n_pairs=B, want_stats=True)

    # ---- warm-up -------------------------------------------------------------------------------------------------
    for s in range(warmup):
        step_device(s)
        step_e2e(s)
    pvp.accumulate_motion(frames_dev, pairs_for(0)[1], grid, 2.0, ctx=ctx, n_pairs=0, want_stats=True)  # drain counters
    grid.zero_()
    ctx.get_profile(reset=True)

    # ---- timed region 1: inputs resident in HBM (`value`) ------------------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps + 1)]
    ctx.profile(True)
    sampler = ClockSampler(env.local)   # NVML start-up (100s of ms, different on every rank) before the barrier: otherwise the skew between
    env.barrier()                       # the ranks lands in the wait of the grid merge (measured once at 8 GPUs: 434 ms instead of 1.3 ms)
    with sampler as clocks:
        for s in range(steps):
            flush.fill_(s & 0xFF)            # L2 flush, outside the timed events
            ev[s][0].record()
            step_device(warmup + s)
            ev[s][1].record()
        ev[-1][0].record()
        if world > 1 and not slabs:          # the run's single grid merge over NVLink (SURVEY 8e mode 1)
            if getattr(grid, "symm", None) is not None:
                pd.reduce_grid_peers(grid, dst=0 if world == 2 else None, ctx=ctx)   # see distributed.ray_voxel_distributed
            else:
                dist.reduce(grid.data, dst=0)
        ev[-1][1].record()
        env.barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    reduce_ms = ev[-1][0].elapsed_time(ev[-1][1])
    prof = ctx.get_profile(reset=True)
    ctx.profile(False)
    k2_pick = ctx.last_k2_pick()    # which lock-step kernel the device-side pick gave the batches to: 1 lean, 2 lean2, 4 lean4 (0: a forced other variant)
    stats = pvp.accumulate_motion(frames_dev, pairs_for(0)[1], grid, 2.0, ctx=ctx, n_pairs=0, want_stats=True)
    if k2_pick == 4:
        # lean4 leaves the RED counter out of its loop (4 of 106 instructions, ~5 %); one extra, untimed step through the instantiation
        # that counts gives the REDs per voxel step of this workload
        ctx.set_option("count_atomics", 1)
        try:
            c = pvp.accumulate_motion(frames_dev, pairs_for(warmup)[1], grid, 2.0, ctx=ctx, n_pairs=B, want_stats=True)
        finally:
            ctx.set_option("count_atomics", 0)
        stats["n_atomics"] = int(round(c["n_atomics"] / max(1, c["n_steps"]) * stats["n_steps"]))
    # slab mode over several ranks: the job's voxel steps are the ranks' WRITES (each step lands in exactly one slab); a rank's n_steps is what it walked
    ms_max, = env.reduce([ms], "max")
    all_steps, all_rays, all_atomics = env.reduce([stats["n_written"] if (slabs and world > 1) else stats["n_steps"], stats["n_rays"], stats["n_atomics"]], "sum")
    if slabs and world > 1:   # every rank saw the same rays
        all_rays = all_rays / world

    # ---- timed region 2: end to end through the public API with HOST (pinned) frames ------------------------------
    grid.zero_()
    e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e_steps = 0
    env.barrier()
    for s in range(steps):
        flush.fill_(s & 0xFF)
        e_ev[s][0].record()
        e_steps += step_e2e(warmup + s)["n_written" if (slabs and world > 1) else "n_steps"]      # H2D of B+1 frames inside; D2H of the step's counters
        e_ev[s][1].record()
    env.barrier()
    e_ms, = env.reduce([sum(a.elapsed_time(b) for a, b in e_ev)], "max")
    e_tot, = env.reduce([e_steps], "sum")
    del grid
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    # ---- roofline of the dominant kernel (K2), live CUDA-event timing ---------------------------------------------
    k2_s = prof["ms_k2"] * 1e-3
    per_rank_steps = stats["n_steps"]
    # a batch (at most max_queue / pixels = 32 1080p pairs) is one K1 + one K2 launch; in the default auto mode a batch launches both
    # lock-step kernels and the one the ray count does not pick exits in ~2.5 us, so "launch" here means the batch's working K2 launch
    n_batches = max(1, int(prof["launches_k1"]))
    k2_name = {1: "dda_accumulate_lean_kernel (K2)", 2: "dda_accumulate_lean2_kernel (K2)", 4: "dda_accumulate_lean4_kernel (K2)"}.get(k2_pick, "K2 as configured")
    roofline = rooflines(env, name, k2_name, per_rank_steps / n_batches, prof["ms_k2"] / n_batches, ALGO_BYTES_PER_STEP[wl["accum"]],
                         clocks, prof["ms_k2"] / ms if ms > 0 else None, pick=k2_pick)
    roofline["k2_pick"] = k2_pick
    roofline["k2_kernels_launched"] = int(prof["launches_k2"])
    roofline["k1_ms_total"] = prof["ms_k1"]
    line = {
        "metric": "voxel_steps_per_sec", "value": all_steps / (ms_max * 1e-3), "unit": "voxel-steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_max / steps, "timed_region_s": ms_max * 1e-3, "higher_is_better": True, "scaling": "strong" if slabs else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(a2, wl),
    }
    frames_total = steps * B * (1 if slabs else world)
    line.update({
        "frames_per_s": frames_total / (ms_max * 1e-3), "rays_per_s": all_rays / (ms_max * 1e-3), "voxel_steps_per_ray": all_steps / max(1.0, all_rays),
        "e2e": {"value": e_tot / (e_ms * 1e-3), "unit": "voxel-steps/s", "h2d_bytes_per_step": (B + 1) * H * W, "d2h_bytes_per_step": 32,
                "frames_per_s": frames_total / (e_ms * 1e-3), "api": "pixeltovoxelprojector_b200.accumulate_motion(pinned host frames) -> pvp_motion_accumulate_host"},
        "gpu_launches": int(prof["launches_k1"] + prof["launches_k2"]), "roofline": roofline,
        "atomics_per_voxel_step": all_atomics / max(1.0, all_steps), "grid_merge_ms": reduce_ms if (world > 1 and not slabs) else None,
        "clocks": clocks.summary(),
    })
    if not full:
        return line
    # L2-atomic roofline: measured peak of independent red.global.add on this GPU, same run
    l2 = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import atomic_roofline
        elem = 4 if wl["accum"] == "f32" else 8
        nbytes = wl["N"] ** 3 * elem                                                    # same footprint as the grid
        buf = torch.zeros(nbytes // 8, dtype=torch.int64, device=dev)
        peak_ops = {nm: atomic_roofline.measure(ctx, buf, nbytes, elem, pat, 1 << 30) for pat, nm in ((0, "random_per_lane"), (3, "random_line_per_warp"))}
        k2_rate = per_rank_steps / k2_s if k2_s > 0 else 0.0
        l2 = {"unit": "atomic adds/s", "buffer_bytes": nbytes, "peak_random_per_lane": peak_ops["random_per_lane"],
              "peak_coalesced_line_per_warp": peak_ops["random_line_per_warp"], "achieved_voxel_steps_per_s_in_k2": k2_rate,
              "frac_of_random_peak": k2_rate / peak_ops["random_per_lane"]}
        del buf
    except Exception as e:  # the micro-benchmark must never break the bench line
        l2 = {"error": repr(e)}
    line["l2_atomic_roofline"] = l2

    # ---- CPU baseline on this box's host cores (bounded sample: one frame pair of the same workload) + parity of that pair ------
    _, p0 = step_pairs(warmup, wl)
    if args.no_cpu:
        line["cpu_baseline"] = {"value": 0.0, "unit": "voxel-steps/s", "cores": 0, "kind": "skipped", "sample": ""}
    else:
        c_steps, c_s, kind, c_grid, c_rays = cpu_reference(wl, frames_np, poses, p0[:1], want_grid=True)
        line["cpu_baseline"] = {"value": c_steps / c_s, "unit": "voxel-steps/s", "cores": 1, "kind": kind,
                                "sample": f"one {W}x{H} frame pair of the workload ({c_steps} voxel-steps, {c_s:.1f} s); reference is single-threaded: 1 of {os.cpu_count()} host cores"}
        try:
            line["parity"] = parity_of_pair(env, wl, frames_dev, poses, p0[0], c_grid, c_rays, c_steps, pick=k2_pick)
        except Exception as e:
            line["parity"] = {"error": repr(e)}
    return line


def measure_files(env, args, name, steps, warmup, full):
    """The whole CLI path in-process (pvp_ray_voxel_accumulate + pvp_write_bin_dev = what `ray_voxel` runs after context creation): a step
    reads the sequence's image files, decodes them on the pool's host threads, stages them over PCIe once each, runs K1 / K2 per batch
    and writes the .bin.  Timed with CUDA events around the whole step (the .bin write ends with a stream sync) and with the host clock."""
    import ctypes as C
    import shutil
    import tempfile
    torch, pvp, ctx, dev, world, rank = env.torch, env.pvp, env.ctx, env.dev, env.world, env.rank
    from pixeltovoxelprojector_b200 import synth
    wl = WORKLOADS[name]
    F, H, W, N = wl["frames"], wl["H"], wl["W"], wl["N"]
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 2 * F * H * W + (1 << 30) else None
    folder = tempfile.mkdtemp(prefix=f"pvp_bench_r{rank}_", dir=base)
    try:
        frames = synth.dense_frames(F, H, W, seed=3000 + rank) if wl["motion"] == "dense" else synth.sparse_frames(F, H, W, seed=3000 + rank, blobs=6)
        poses = synth.orbit_poses(F, N, wl["voxel_size"], wl["center"], seed=3001)
        meta = synth.write_sequence(folder, {0: frames}, {0: poses})
        del frames
        out = os.path.join(folder, "voxel_grid.bin")
        grid = pvp.VoxelGrid(N, wl["voxel_size"], wl["center"], accum=wl["accum"], device=dev)
        reps, ms, wall, bin_s = [], [], [], []

        def one():
            grid.zero_()
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            rep = pvp.ray_voxel_accumulate(meta, folder, grid, 2.0, ctx=ctx)
            t1 = time.perf_counter()
            pvp._lib.check(ctx._lib.pvp_write_bin_dev(ctx.handle, os.fsencode(out), N, C.c_float(wl["voxel_size"]), grid.data.data_ptr()))
            bin_s.append(time.perf_counter() - t1)
            b.record()
            torch.cuda.synchronize(dev)
            return rep, a.elapsed_time(b), time.perf_counter() - t0

        for _ in range(warmup):
            one()
        env.barrier()
        sampler = ClockSampler(env.local)
        with sampler as clocks:
            for _ in range(steps):
                r, m, w = one()
                reps.append(r), ms.append(m), wall.append(w)
        env.barrier()
    finally:
        shutil.rmtree(folder, ignore_errors=True)
    t_ms, = env.reduce([sum(ms)], "max")
    w_s, = env.reduce([sum(wall)], "max")
    tot_steps, tot_frames = env.reduce([sum(r["n_steps"] for r in reps), sum(r["n_frames_loaded"] for r in reps)], "sum")
    del grid
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    v = tot_steps / (t_ms * 1e-3)
    r0 = reps[-1]
    e2e = {"value": tot_steps / w_s, "unit": "voxel-steps/s", "h2d_bytes_per_step": int(r0["n_frames_loaded"]) * H * W, "d2h_bytes_per_step": 8 + 4 * N ** 3,
           "frames_per_s": tot_frames / w_s, "api": "pixeltovoxelprojector_b200.ray_voxel_accumulate(metadata.json, folder) + pvp_write_bin_dev; host clock, files in the page cache"}
    return {
        "metric": "voxel_steps_per_sec", "value": v, "unit": "voxel-steps/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": t_ms / steps,
        "timed_region_s": t_ms * 1e-3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "path": "files -> .bin: metadata.json + %d %dx%d PGM files per rank -> %d^3 %s grid -> .bin" % (F, W, H, N, wl["accum"]), "grid": f"{N}^3 {wl['accum']}",
                   "motion": wl["motion"], "parallelism": f"{world} GPU(s), own sequence each" if world > 1 else "1 GPU", "l2": "inputs come from files every step (2 GB per step >> L2)"},
        "frames_per_s": tot_frames / (t_ms * 1e-3), "e2e": e2e, "gpu_launches": None,
        "stages_s_per_step": {k: r0[k] for k in ("seconds_total", "seconds_setup", "seconds_decode_wait", "seconds_upload", "seconds_launch", "seconds_drain")} | {"seconds_bin_write": bin_s[-1]}, "decode_threads": r0["decode_threads"],
        "host_cores": os.cpu_count(), "clocks": clocks.summary(),
        "roofline": {"bound": "host", "kernel": "none dominant: decode pool + PCIe + K1/K2 overlap", "frac": None},
        "cpu_baseline": {"value": 0.0, "unit": "voxel-steps/s", "cores": 0, "kind": "skipped", "sample": ""},
    }


def compact(line):
    """An `also` entry: the figures of a secondary workload without the long descriptions."""
    if line is None:
        return None
    r = line.get("roofline", {})
    out = {k: line[k] for k in ("value", "unit", "steps", "warmup", "ms_per_step", "timed_region_s", "frames_per_s", "scaling", "gpu_launches", "atomics_per_voxel_step", "grid_merge_ms",
                                "stages_s_per_step", "decode_threads", "host_cores") if k in line}
    out["e2e"] = {k: line["e2e"][k] for k in ("value", "frames_per_s", "h2d_bytes_per_step", "d2h_bytes_per_step")}
    out["config"] = {k: line["config"][k] for k in ("workload", "grid", "parallelism") if k in line["config"]}
    out["roofline"] = {k: r.get(k) for k in ("bound", "frac", "kernel", "share_of_step", "avg_launch_ms", "warp_inst_per_voxel_step", "sass_md5_ok")}
    am = r.get("algorithmic_model") or {}
    out["roofline"]["algorithmic_model"] = {k: am.get(k) for k in ("frac", "traffic_ratio", "algorithmic_bytes_per_voxel_step")}
    out["clocks"] = line.get("clocks")
    return out


# secondary workloads measured by the default run (the contract is ONE line: they ride in its `also` dict), short runs
ALSO_DEFAULT = [("image_4096_512_fixed", 5, 3), ("1080p_dense_256", 3, 3), ("1080p_scatter10_512", 3, 3), ("1080p_object_512", 3, 3), ("files_1080p_1000", 3, 1)]
ALSO_8GPU = [("1080p_dense_1024_slab8", 10, 3), ("1080p_dense_1024", 10, 3)]


def run_ours(args):
    # the contract is ONE JSON line on stdout: libraries that print there from C (NCCL's version banner, whatever NCCL_DEBUG says in
    # this environment) are sent to stderr for the duration of the run; the line itself goes to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    env = Env(args)
    wl = WORKLOADS[args.workload]
    pick = lambda w: measure_image if w.get("path") == "B" else measure_files if w.get("path") == "files" else measure_motion   # noqa: E731
    line = pick(wl)(env, args, args.workload, args.steps, args.warmup, full=True)
    if args.workload == "1080p_dense_512" and not args.no_also:
        also = {}
        for name, st, wu in ALSO_DEFAULT + (ALSO_8GPU if (env.world == 8 or (args.also_multi and env.world > 1)) else []):
            if name not in WORKLOADS:
                continue
            try:
                also[name] = compact(pick(WORKLOADS[name])(env, args, name, st, wu, full=False))
            except Exception as e:   # a secondary workload must never cost the primary line (collective calls inside are symmetric over ranks)
                also[name] = {"error": repr(e)}
        if line is not None:
            line["also"] = also
    env.close()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1080p_dense_512", choices=sorted(WORKLOADS))
    ap.add_argument("--opt", action="append", help="context option key=value (pvp_ctx_set_option), repeatable; kernel experiments")
    ap.add_argument("--variant", type=int, default=0, help="K2 variant (0 = library default; see pvp_motion.cu launch_k2)")
    ap.add_argument("--reduce", default="peers", choices=["peers", "nccl"], help="N>1: grid merge by our peer-memory kernel (default) or by NCCL")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (kernel experiments only)")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads of the default run")
    ap.add_argument("--also-multi", action="store_true", help="run the 8-GPU secondary workloads (1024^3 as slabs and as frames + merge) at any N > 1")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        if wl.get("path") == "files":
            raise SystemExit("bench.py: the files workloads have no reference arm of their own (the reference arm of the kernels' workloads times the same loop body)")
        (run_reference_image if wl.get("path") == "B" else run_reference)(args, wl)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
