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
    "1080p_dense_512": dict(H=1080, W=1920, N=512, voxel_size=2.0, center=(0.0, 0.0, 500.0), pairs_per_step=4, pool=72, motion="dense", accum="f32"),
    # BASELINE config[1]: 256^3 grid (L2 resident)
    "1080p_dense_256": dict(H=1080, W=1920, N=256, voxel_size=4.0, center=(0.0, 0.0, 500.0), pairs_per_step=4, pool=72, motion="dense", accum="f32"),
    # sparse motion (moving blobs, ~1 % of the pixels): frames/s matters, many pairs per step
    "1080p_sparse_512": dict(H=1080, W=1920, N=512, voxel_size=2.0, center=(0.0, 0.0, 500.0), pairs_per_step=64, pool=72, motion="sparse", accum="f32"),
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
    # small smoke-sized case for CPU-side checks of this script
    "tiny": dict(H=120, W=160, N=64, voxel_size=16.0, center=(0.0, 0.0, 500.0), pairs_per_step=2, pool=8, motion="dense", accum="f32"),
}
def measured_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[workload]["bytes_per_launch"]
    except Exception:
        return None


ALGO_BYTES_PER_STEP = {"f32": 8, "fixed": 16}   # one atomic RMW: 4 B read + 4 B write (fp32), 8 + 8 (int64) -- SURVEY 8d


def make_inputs(wl, seed, pose_seed=1001):
    """Frames differ per rank (seed); the pose table is the same on every rank, so that the per-GPU work of the weak-scaling
    run is the same everywhere (ray lengths depend on the pose, not on the pixel values)."""
    from pixeltovoxelprojector_b200 import synth
    F = wl["pool"]
    if wl["motion"] == "dense":
        frames = synth.dense_frames(F, wl["H"], wl["W"], seed=seed)
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
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def issue_roofline(workload, units_per_s, clocks, dev):
    """The bound the lock-step kernels actually sit on: warp-instruction issue.  Instructions per voxel step come from the
    committed ncu capture (profiles/traffic.json), the rate from this run's CUDA-event timing, the peak from the SM count x 4
    schedulers x the SM clock sampled during this run.  Extra information next to `roofline`; never breaks the bench line."""
    try:
        import torch
        e = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
        if not e or "warp_inst_per_launch" not in e:
            return None   # no ncu capture committed for this workload
        per_unit = e["warp_inst_per_launch"] / e["voxel_steps_per_launch"]
        c = clocks.summary()
        mhz = c["sm_mhz"] or c["sm_max_mhz"]
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        peak = sms * 4 * float(mhz) * 1e6
        achieved = per_unit * units_per_s
        return {"bound": "issue", "unit": "warp instructions/s", "achieved": achieved, "peak": peak, "frac": achieved / peak,
                "warp_inst_per_voxel_step": per_unit, "source": "profiles/traffic.json (ncu smsp__inst_executed.sum) x live rate; peak = SMs x 4 x sampled SM clock"}
    except Exception as ex:
        return {"error": repr(ex)}


def cpu_reference(wl, frames, poses, pairs, rows=None):
    """Time the reference CPU implementation (oracle/_ref = the unmodified ray_voxel.cpp compiled; else the C port) on
    the given pairs, optionally restricted to a band of rows (pixels outside the band are made motion-free).
    Returns (voxel_steps, seconds, kind).  The reference is serial: 1 thread."""
    import oracle
    use_ref = oracle.have_ref()
    lib = oracle.RefLib() if use_ref else oracle.Oracle()
    N = wl["N"]
    grid = np.zeros((N, N, N), np.float32)
    steps, secs = 0, 0.0
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
            _, ns = lib.accumulate_pair(prev, curr, np.array(poses[c].camera_position, np.float32), R, f, grid, N, wl["voxel_size"], wl["center"], 2.0)
        else:
            _, ns, _ = lib.accumulate_pair(prev, curr, np.array(poses[c].camera_position, np.float32), R, f, grid, N, wl["voxel_size"], np.array(wl["center"], np.float32), 2.0)
        secs += time.perf_counter() - t0
        steps += ns
    return steps, secs, ("reference" if use_ref else "port")


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


def run_ours_image(args, wl):
    import torch
    import torch.distributed as dist

    import pixeltovoxelprojector_b200 as pvp
    from pixeltovoxelprojector_b200 import process_image as pi

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = pvp.default_context(dev)
    for kv in args.opt or []:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    kw = dict(accum_mode=wl["accum"], frac_bits=wl["frac_bits"], ctx=ctx)

    def call(img, want_stats):
        pvp.process_image_cpp(img, geom["earth_position"], geom["pointing_direction"], geom["fov"], W, H, grid, geom["extent"], geom["max_distance"],
                              geom["num_steps"], tex, *geom["sky"], False, False, want_stats=want_stats, **kw)
        return pi.last_stats if want_stats else None

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for s in range(args.warmup):
        call(images_dev[s % wl["pool"]], False)
    samples_per_image = [call(images_dev[i], True)["n_samples"] for i in range(wl["pool"])]
    grid.zero_()
    ctx.get_profile(reset=True)
    ctx.profile(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps + 1)]
    sampler = ClockSampler(local)   # NVML start-up (100s of ms, different on every rank) before the barrier, not between it and the timed steps
    barrier()
    my_samples = 0
    with sampler as clocks:
        for s in range(args.steps):
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
        barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    prof = ctx.get_profile(reset=True)
    ctx.profile(False)
    # e2e: the image of every step comes from pinned HOST memory (H2D inside), the step's counters go back (D2H)
    grid.zero_()
    e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_samples = 0
    barrier()
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
    for s in range(args.steps):
        flush.fill_(s & 0xFF)
        e_ev[s][0].record()
        if s == 0:
            prefetch(0)
        if s + 1 < args.steps:
            prefetch(s + 1)
        main.wait_event(ready[s % 2])
        e_samples += call(stages[s % 2], True)["n_samples"]
        freed[s % 2].record(main)
        e_ev[s][1].record()
    barrier()
    t = torch.tensor([ms, sum(a.elapsed_time(b) for a, b in e_ev)], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(my_samples), float(e_samples)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    k3_s = prof["ms_k3"] * 1e-3
    achieved = 16 * my_samples / k3_s / 1e9 if k3_s > 0 else 0.0     # one 8-byte atomic RMW per sample
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": measured_traffic(args.workload), "kernel": "process_image_lockstep2_kernel (K3)", "peak_source": peak_src, "algorithmic_bytes_per_voxel_step": 16,
                "voxel_steps_per_launch": my_samples / max(1, prof["launches_k3"]), "avg_launch_ms": prof["ms_k3"] / max(1, prof["launches_k3"]),
                "k3_share_of_step": prof["ms_k3"] / ms if ms > 0 else None}
    if args.no_cpu:
        cpu = {"value": 0.0, "unit": "voxel-steps/s", "cores": 0, "kind": "skipped", "sample": ""}
    else:
        import io, contextlib
        buf = io.StringIO()
        a2 = argparse.Namespace(**{**vars(args), "steps": 3, "warmup": 1})
        with contextlib.redirect_stdout(buf):
            run_reference_image(a2, wl)
        cpu = json.loads(buf.getvalue().strip().splitlines()[-1])["cpu_baseline"]
    value = float(tot[0].item()) / (float(t[0].item()) * 1e-3)
    e_value = float(tot[1].item()) / (float(t[1].item()) * 1e-3)
    print(json.dumps({
        "metric": "voxel_steps_per_sec", "value": value, "unit": "voxel-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(t[0].item()) / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": image_config(args, wl), "frames_per_s": args.steps * world / (float(t[0].item()) * 1e-3),
        "e2e": {"value": e_value, "unit": "voxel-steps/s", "h2d_bytes_per_step": H * W * 8, "d2h_bytes_per_step": 16,
                "frames_per_s": args.steps * world / (float(t[1].item()) * 1e-3),
                "api": "pixeltovoxelprojector_b200.process_image_cpp(CUDA tensors; image staged from pinned host memory, double buffered) -> pvp_process_image"},
        "gpu_launches": int(prof["launches_k3"]), "roofline": roofline,
        "issue_roofline": issue_roofline(args.workload, my_samples / k3_s if k3_s > 0 else 0.0, clocks, dev),
        "cpu_baseline": cpu, "clocks": clocks.summary(),
    }))
    if world > 1:
        dist.destroy_process_group()


def workload_config(args, wl):
    return {"workload": args.workload, "k2_variant": getattr(args, "variant", 0), "frame": f"{wl['W']}x{wl['H']} uint8", "grid": f"{wl['N']}^3 {wl['accum']}", "voxel_size": wl["voxel_size"],
            "pairs_per_step": wl["pairs_per_step"], "motion": wl["motion"], "camera": "pose changes every frame (orbit inside the grid); N>1: same pose table, different frames on every rank",
            "parallelism": (f"grid sharded into {args.gpus if args.gpus > 1 else wl['slabs']} x-slabs; " + (f"{args.gpus} ranks, rank r owns slab r" if args.gpus > 1 else f"1 GPU running rank {wl['slabs'] // 2}'s slab") + f"; {getattr(args, 'reduce_how', '')}") if wl.get("slabs") else
                           f"frames sharded over {args.gpus} GPU(s), private grids, one grid merge per run inside the timed region: {getattr(args, 'reduce_how', 'NCCL reduce')}" if args.gpus > 1 else "1 GPU",
            "l2": "256 MiB buffer written between timed steps (L2 flush); frame pool 149 MB and grid > L2" if wl["N"] >= 512 else
                  "256 MiB buffer written between timed steps (L2 flush)"}


def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    import pixeltovoxelprojector_b200 as pvp

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = pvp.default_context(dev)
    if args.variant:
        ctx.set_option("dda_variant", args.variant)
    for kv in args.opt or []:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))

    slabs = wl.get("slabs")
    frames_np, poses = make_inputs(wl, seed=1000 + (0 if slabs else rank))   # weak scaling: every rank has its own frames, same shape (slab mode: the same frames)
    frames_dev = torch.as_tensor(frames_np, device=dev)
    frames_pin = torch.as_tensor(frames_np).pin_memory()
    grid, reduce_how = None, "none (1 GPU)"
    if slabs:
        from pixeltovoxelprojector_b200.distributed import slab_range
        n_slabs = world if world > 1 else slabs     # --gpus N: N slabs, one per rank (N = 8 is BASELINE config 4); 1 GPU: rank 4 of 8
        lo, hi = slab_range(wl["N"], rank if world > 1 else slabs // 2, n_slabs)
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
    args.reduce_how = reduce_how
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
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

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up -------------------------------------------------------------------------------------------------
    for s in range(args.warmup):
        step_device(s)
        step_e2e(s)
    pvp.accumulate_motion(frames_dev, pairs_for(0)[1], grid, 2.0, ctx=ctx, n_pairs=0, want_stats=True)  # drain counters
    grid.zero_()
    ctx.get_profile(reset=True)

    # ---- timed region 1: inputs resident in HBM (`value`) ------------------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps + 1)]
    ctx.profile(True)
    sampler = ClockSampler(local)   # NVML start-up (100s of ms, different on every rank) before the barrier: otherwise the skew between
    barrier()                       # the ranks lands in the wait of the grid merge (measured once at 8 GPUs: 434 ms instead of 1.3 ms)
    with sampler as clocks:
        for s in range(args.steps):
            flush.fill_(s & 0xFF)            # L2 flush, outside the timed events
            ev[s][0].record()
            step_device(args.warmup + s)
            ev[s][1].record()
        ev[-1][0].record()
        if world > 1 and not slabs:          # the run's single grid merge over NVLink (SURVEY 8e mode 1)
            if getattr(grid, "symm", None) is not None:
                pd.reduce_grid_peers(grid, dst=0 if world == 2 else None, ctx=ctx)   # see distributed.ray_voxel_distributed
            else:
                dist.reduce(grid.data, dst=0)
        ev[-1][1].record()
        barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    reduce_ms = ev[-1][0].elapsed_time(ev[-1][1])
    prof = ctx.get_profile(reset=True)
    ctx.profile(False)
    stats = pvp.accumulate_motion(frames_dev, pairs_for(0)[1], grid, 2.0, ctx=ctx, n_pairs=0, want_stats=True)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    # slab mode over several ranks: the job's voxel steps are the ranks' WRITES (each step lands in exactly one slab); a rank's n_steps is what it walked
    tot = torch.tensor([stats["n_written"] if (slabs and world > 1) else stats["n_steps"], stats["n_rays"], stats["n_atomics"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max, all_steps, all_rays = float(t.item()), float(tot[0].item()), float(tot[1].item())
    if slabs and world > 1:   # every rank saw the same rays
        all_rays = all_rays / world

    # ---- timed region 2: end to end through the public API with HOST (pinned) frames ------------------------------
    grid.zero_()
    e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_steps = 0
    barrier()
    for s in range(args.steps):
        flush.fill_(s & 0xFF)
        e_ev[s][0].record()
        e_steps += step_e2e(args.warmup + s)["n_written" if (slabs and world > 1) else "n_steps"]      # H2D of B+1 frames inside; D2H of the step's counters
        e_ev[s][1].record()
    barrier()
    e_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in e_ev)], dtype=torch.float64, device=dev)
    e_tot = torch.tensor([float(e_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e_tot, op=dist.ReduceOp.SUM)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K2), live CUDA-event timing ---------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    k2_s = prof["ms_k2"] * 1e-3
    per_rank_steps = stats["n_steps"]
    achieved = ALGO_BYTES_PER_STEP[wl["accum"]] * per_rank_steps / k2_s / 1e9 if k2_s > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": measured_traffic(args.workload), "kernel": "dda_accumulate_lean2_kernel (K2)", "peak_source": peak_src,
                "algorithmic_bytes_per_voxel_step": ALGO_BYTES_PER_STEP[wl["accum"]],
                # one batch per step; in the default auto mode a batch launches both lock-step kernels and the one the ray count
                # does not pick exits in ~2.5 us, so "launch" here means the batch's working K2 launch
                "voxel_steps_per_launch": per_rank_steps / max(1, args.steps), "avg_launch_ms": prof["ms_k2"] / max(1, args.steps),
                "k2_kernels_launched": int(prof["launches_k2"]),
                "k2_share_of_step": prof["ms_k2"] / ms if ms > 0 else None, "k1_ms_total": prof["ms_k1"]}
    # L2-atomic roofline: measured peak of independent red.global.add on this GPU, same run
    l2 = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import atomic_roofline
        elem = 4 if wl["accum"] == "f32" else 8
        buf = torch.zeros(grid.data.numel() * grid.data.element_size() // 8, dtype=torch.int64, device=dev)  # same footprint as the grid
        nbytes = buf.numel() * 8
        peak_ops = {name: atomic_roofline.measure(ctx, buf, nbytes, elem, pat, 1 << 30) for pat, name in ((0, "random_per_lane"), (3, "random_line_per_warp"))}
        k2_rate = per_rank_steps / k2_s if k2_s > 0 else 0.0
        l2 = {"unit": "atomic adds/s", "buffer_bytes": nbytes, "peak_random_per_lane": peak_ops["random_per_lane"],
              "peak_coalesced_line_per_warp": peak_ops["random_line_per_warp"], "achieved_voxel_steps_per_s_in_k2": k2_rate,
              "frac_of_random_peak": k2_rate / peak_ops["random_per_lane"]}
    except Exception as e:  # the micro-benchmark must never break the bench line
        l2 = {"error": repr(e)}

    # ---- CPU baseline on this box's host cores (bounded sample: one frame pair of the same workload) -------------
    _, p0 = step_pairs(args.warmup, wl)
    if args.no_cpu:
        c_steps, c_s, kind = 0, 1.0, "skipped"
    else:
        c_steps, c_s, kind = cpu_reference(wl, frames_np, poses, p0[:1])
    cpu = {"value": c_steps / c_s, "unit": "voxel-steps/s", "cores": 1, "kind": kind,
           "sample": f"one {W}x{H} frame pair of the workload ({c_steps} voxel-steps, {c_s:.1f} s); reference is single-threaded: 1 of {os.cpu_count()} host cores"}

    value = all_steps / (ms_max * 1e-3)
    e_value = float(e_tot.item()) / (float(e_ms.item()) * 1e-3)
    frames_total = args.steps * B * (1 if slabs else world)
    line = {
        "metric": "voxel_steps_per_sec", "value": value, "unit": "voxel-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if slabs else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, wl),
        "frames_per_s": frames_total / (ms_max * 1e-3), "rays_per_s": all_rays / (ms_max * 1e-3), "voxel_steps_per_ray": all_steps / max(1.0, all_rays),
        "e2e": {"value": e_value, "unit": "voxel-steps/s", "h2d_bytes_per_step": (B + 1) * H * W, "d2h_bytes_per_step": 32,
                "frames_per_s": frames_total / (float(e_ms.item()) * 1e-3), "api": "pixeltovoxelprojector_b200.accumulate_motion(pinned host frames) -> pvp_motion_accumulate_host"},
        "gpu_launches": int(prof["launches_k1"] + prof["launches_k2"]),
        "roofline": roofline, "l2_atomic_roofline": l2,
        "issue_roofline": issue_roofline(args.workload, per_rank_steps / k2_s if k2_s > 0 else 0.0, clocks, dev),
        "cpu_baseline": cpu, "clocks": clocks.summary(),
        "atomics_per_voxel_step": float(tot[2].item()) / max(1.0, all_steps),
        "grid_merge_ms": reduce_ms if world > 1 else None,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1080p_dense_512", choices=sorted(WORKLOADS))
    ap.add_argument("--opt", action="append", help="context option key=value (pvp_ctx_set_option), repeatable; kernel experiments")
    ap.add_argument("--variant", type=int, default=0, help="K2 variant (0 = library default; see pvp_motion.cu launch_k2)")
    ap.add_argument("--reduce", default="peers", choices=["peers", "nccl"], help="N>1: grid merge by our peer-memory kernel (default) or by NCCL")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (kernel experiments only)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if wl.get("path") == "B":
        (run_reference_image if args.impl == "reference" else run_ours_image)(args, wl)
    elif args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
