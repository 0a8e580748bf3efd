"""One workload's dominant kernel on IDENTICAL launches, for an exact ncu capture: every launch processes the same batch, the script prints
the voxel steps (samples) of one launch, so warp instructions and DRAM bytes per voxel step follow from the capture without guessing which
batch ncu caught.
    python tools/ncu_case.py 1080p_dense_512 [launches [pairs]] -> "voxel_steps_per_launch ..." (K2: one batch of 32 frame pairs; 4K frames: pass 8 pairs,
                                                                   the largest batch the ray queue holds)
    python tools/ncu_case.py image_4096_512_fixed [launches]   -> one image (K3)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pixeltovoxelprojector_b200 as pvp  # noqa: E402

name = sys.argv[1]
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wl = bench.WORKLOADS[name]
ctx = pvp.default_context()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
if wl.get("path") == "B":
    from pixeltovoxelprojector_b200 import process_image as pi
    images, geom = bench.image_scene(wl, seed=2000)
    n, H, W = wl["N"], wl["H"], wl["W"]
    cell = torch.int64 if wl["accum"] == "fixed" else torch.float64
    img = torch.as_tensor(images[0], device="cuda")
    grid, tex = torch.zeros((n, n, n), dtype=cell, device="cuda"), torch.zeros((8, 8), dtype=cell, device="cuda")
    per = None
    for _ in range(launches):
        flush.fill_(1)
        pvp.process_image_cpp(img, geom["earth_position"], geom["pointing_direction"], geom["fov"], W, H, grid, geom["extent"], geom["max_distance"],
                              geom["num_steps"], tex, *geom["sky"], False, False, accum_mode=wl["accum"], frac_bits=wl["frac_bits"], want_stats=True)
        per = pi.last_stats["n_samples"]
else:
    wl = dict(wl, pool=(int(sys.argv[3]) if len(sys.argv) > 3 else 32) + 1)
    frames, poses = bench.make_inputs(wl, seed=1000)
    fr = torch.as_tensor(frames, device="cuda")
    tab, n_pairs = pvp.make_pairs(poses, wl["W"])
    grid = pvp.VoxelGrid(wl["N"], wl["voxel_size"], wl["center"], accum=wl["accum"])
    per = None
    for _ in range(launches):
        flush.fill_(1)
        per = pvp.accumulate_motion(fr, tab, grid, 2.0)
torch.cuda.synchronize()
print(json.dumps({"workload": name, "launches": launches, "per_launch": per}))
