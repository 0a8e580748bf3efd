"""One process, several (grid size, K2 variant, K1 tile rows) combinations of the dense 1080p workload -- run under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:dda_`
to see how the queue order changes DRAM traffic per K2 launch.  usage: locality_probe.py N:variant:rows[:extra_roll_degrees] [...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pixeltovoxelprojector_b200 as pvp
from pixeltovoxelprojector_b200 import synth

H, W, center = 1080, 1920, (0.0, 0.0, 500.0)
from pixeltovoxelprojector_b200.motion import default_context
ctx = default_context()
frames = torch.as_tensor(synth.dense_frames(3, H, W, seed=7), device="cuda")
for spec in sys.argv[1:]:
    f = [int(x) for x in spec.split(":")]
    N, variant, rows, roll = f[0], f[1], f[2], (f[3] if len(f) > 3 else 0)
    vs = 1024.0 / N
    poses = synth.orbit_poses(3, N, vs, center, seed=8)
    if roll:
        import dataclasses
        poses = [dataclasses.replace(p, roll=p.roll + roll) for p in poses]
    tab, _ = pvp.make_pairs(poses, W)
    ctx.set_option("dda_variant", variant)
    ctx.set_option("k1_row_group", rows)
    g = pvp.VoxelGrid(N, vs, center)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    pvp.accumulate_motion(frames, tab, g, 2.0)
    torch.cuda.synchronize()
    ev[0].record()
    st = pvp.accumulate_motion(frames, tab, g, 2.0)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    print(spec, "steps", st["n_steps"], "ms", round(ms, 2), "G steps/s", round(st["n_steps"] / ms / 1e6, 1), "atomics/step", round(st["n_atomics"] / st["n_steps"], 3), flush=True)
    del g
