"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
Frame-sharded and slab-sharded runs of the ray_voxel pipeline must write the same .bin as one GPU alone
(fixed point: byte-identical; fp32 with uint8 frames: byte-identical too, the sums are exact)."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pixeltovoxelprojector_b200 as pvp  # noqa: E402
from pixeltovoxelprojector_b200 import synth  # noqa: E402
from pixeltovoxelprojector_b200.distributed import ray_voxel_distributed  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n, vs, center = 192, 5.0, (0.0, 0.0, 500.0)
    folder = os.path.join(tempfile.gettempdir(), "pvp_dist_check")
    if rank == 0:
        frames = {c: synth.dense_frames(7, 120, 160, seed=c) for c in (0, 1, 2)}
        poses = {c: synth.orbit_poses(7, n, vs, center, seed=10 + c, camera_index=c) for c in (0, 1, 2)}
        synth.write_sequence(folder, frames, poses, shuffle_seed=3)
        os.remove(os.path.join(folder, poses[1][3].image_file))  # a skipped frame inside a shard
    dist.barrier()
    meta = os.path.join(folder, "metadata.json")
    ok = True
    for accum in ("fixed", "f32"):
        outs = {}
        for mode in ("frames", "slabs"):
            outs[mode] = os.path.join(folder, f"{mode}_{accum}.bin")
            rep = ray_voxel_distributed(meta, folder, outs[mode], mode=mode, n=n, voxel_size=vs, center=center, accum=accum, frac_bits=12)
            dist.barrier()
            if rank == 0:
                print(f"[{accum}/{mode}] world={world} pairs={rep['n_pairs']} rays={rep['n_rays']} steps={rep['n_steps']} written={rep['n_written']}")
        if rank == 0:
            grid = pvp.VoxelGrid(n, vs, center, accum=accum, frac_bits=12)
            one = pvp.ray_voxel_accumulate(meta, folder, grid)
            single = os.path.join(folder, f"single_{accum}.bin")
            grid.save(single)
            ref = open(single, "rb").read()
            for mode, path in outs.items():
                same = open(path, "rb").read() == ref
                ok &= same
                print(f"[{accum}/{mode}] identical to 1-GPU .bin: {same} (1-GPU steps={one['n_steps']})")
    # the peer-memory reduce kernel against the NCCL collective on random grids: multicast and plain peer loads, to one rank
    # and to all ranks, float32 (integer-valued => exact in any order) and int64, sizes that do / do not fill 16-byte vectors
    from pixeltovoxelprojector_b200.distributed import reduce_grid, reduce_grid_peers, symmetric_grid
    for accum, nn in (("f32", 96), ("fixed", 96), ("f32", 3), ("fixed", 5)):
        for dst in (0, None):
            for mc in (True, False):
                g = symmetric_grid(nn, vs, center, accum=accum, frac_bits=12)
                gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
                vals = torch.randint(0, 1 << 16, g.data.shape, device="cuda", generator=gen)
                g.data.copy_(vals.to(g.data.dtype))
                want = vals.to(torch.int64).clone()
                dist.all_reduce(want)
                reduce_grid_peers(g, dst=dst, multicast=mc)
                torch.cuda.synchronize()
                if dst is None or rank == dst:
                    same = bool(torch.equal(g.data.to(torch.int64), want))
                    ok &= same
                    if rank == 0:
                        print(f"[peer reduce {accum} n={nn} dst={dst} multicast={mc and bool(g.symm.multicast_ptr)}] equals NCCL all_reduce: {same}")
                dist.barrier()
    # path B: row shards of one image and file shards of an observation list into private int64 tensors, one merge of grid +
    # texture (NCCL), equal to one GPU alone bit for bit
    from pixeltovoxelprojector_b200.distributed import process_image_sharded, process_images_sharded, reduce_sky
    rng = np.random.default_rng(5)
    H = W = 384
    ext, shape, sky = [(-64.0, 64.0), (-64.0, 64.0), (40.0, 168.0)], (128, 128, 128), (0.3, -0.2, 1.0, 1.5)
    obs = [(torch.as_tensor(rng.random((H, W)), device="cuda"), [0.5 * k, -0.2, 0.3], [0.05 - 0.02 * k, 0.02, 1.0], 0.9) for k in range(6)]
    tail = (W, H)
    for what in ("rows", "files"):
        g = torch.zeros(shape, dtype=torch.int64, device="cuda")
        t = torch.zeros((64, 96), dtype=torch.int64, device="cuda")
        if what == "rows":
            for image, pos, pointing, fov in obs:
                process_image_sharded(image, pos, pointing, fov, W, H, g, ext, 200.0, 400, t, *sky, True, False, accum_mode="fixed", frac_bits=24)
        else:
            process_images_sharded(obs, W, H, g, ext, 200.0, 400, t, *sky, accum_mode="fixed", frac_bits=24)
        reduce_sky(g, t, dst=0)
        if rank == 0:
            g1 = torch.zeros_like(g)
            t1 = torch.zeros_like(t)
            for image, pos, pointing, fov in obs:
                pvp.process_image_cpp(image, pos, pointing, fov, W, H, g1, ext, 200.0, 400, t1, *sky, True, False, accum_mode="fixed", frac_bits=24)
            same = bool(torch.equal(g, g1) and torch.equal(t, t1)) and int(g1.sum()) > 0
            ok &= same
            print(f"[path B {what}] world={world} merged grid + texture identical to 1 GPU: {same} (grid sum {int(g1.sum())})")
        dist.barrier()
    flag = torch.tensor([int(ok)], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
