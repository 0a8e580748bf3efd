"""files -> .bin over all ranks (frame sharding, one merge): how well does the host feed N GPUs at once?
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/dist_files_bench.py [--frames 2000] [--motion sparse|dense]
Rank 0 writes a synthetic 1080p PGM sequence + metadata.json into /dev/shm, every rank runs pixeltovoxelprojector_b200.distributed.
ray_voxel_distributed on it (each rank decodes and uploads only its own block of the frames; private 256^3 grids; one merge; rank 0
writes the .bin).  Prints one JSON line: frames/s of the whole job (host clock around the collective call, max over ranks) and the
per-rank stage split.  With N = 1 the same script gives the single-GPU figure to compare with."""
import argparse
import json
import os
import shutil
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pixeltovoxelprojector_b200 as pvp  # noqa: E402
from pixeltovoxelprojector_b200 import synth  # noqa: E402
from pixeltovoxelprojector_b200.distributed import ray_voxel_distributed  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--motion", default="sparse", choices=["sparse", "dense"])
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if world > 1 else 0
    H, W, N, vs, center = 1080, 1920, a.n, 1024.0 / a.n, (0.0, 0.0, 500.0)
    folder = "/dev/shm/pvp_dist_files"
    if rank == 0:
        shutil.rmtree(folder, ignore_errors=True)
        frames = synth.dense_frames(a.frames, H, W, seed=1) if a.motion == "dense" else synth.sparse_frames(a.frames, H, W, seed=1, blobs=6)
        synth.write_sequence(folder, {0: frames}, {0: synth.orbit_poses(a.frames, N, vs, center, seed=2)})
        del frames
    if world > 1:
        dist.barrier()
    meta, out = os.path.join(folder, "metadata.json"), os.path.join(folder, "voxel_grid.bin")
    times, rep = [], None
    for r in range(a.reps + 1):                    # first repetition warms up (context, kernels, rings, page cache)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        rep = ray_voxel_distributed(meta, folder, out, mode="frames", n=N, voxel_size=vs, center=center)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if r:
            times.append(float(dt.item()))
    if rank == 0:
        best = min(times)
        print(json.dumps({"n_gpus": world, "frames": a.frames, "motion": a.motion, "grid": f"{N}^3 f32", "seconds": times, "frames_per_s": a.frames / best,
                          "voxel_steps_per_s": rep["n_steps"] / best, "pairs": rep["n_pairs"], "host_cores": os.cpu_count(), "mode": rep["mode"]}))
        shutil.rmtree(folder, ignore_errors=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
