"""Path B orientation probe: the 4096^2 -> 512^3 fixed-point workload with the observer looking along +z (bench.py's scene),
+y and +x, i.e. with the image columns running along world y, x and y -- the grid is [x][y][z] with z fastest, so a warp's
32-column tile spreads over x-planes (2 MiB apart) only in the second case.  usage: image_probe.py [variant ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pixeltovoxelprojector_b200 as pvp
from pixeltovoxelprojector_b200 import process_image as pi
from pixeltovoxelprojector_b200.motion import default_context

H = W = 4096
n = 512
ctx = default_context()
img = torch.as_tensor(np.random.default_rng(1).random((H, W)), device="cuda")
half = n / 2.0
scenes = {
    "look +z (columns ~ y)": ([0.02, 0.01, 1.0], [(-half, half), (-half, half), (100.0, 100.0 + n)]),
    "look +y (columns ~ x)": ([0.02, 1.0, 0.01], [(-half, half), (100.0, 100.0 + n), (-half, half)]),
    "look +x (columns ~ y)": ([1.0, 0.02, 0.01], [(100.0, 100.0 + n), (-half, half), (-half, half)]),
}
for variant in [int(v) for v in sys.argv[1:]] or [0]:
    ctx.set_option("image_variant", variant)
    for name, (pointing, ext) in scenes.items():
        grid = torch.zeros((n, n, n), dtype=torch.int64, device="cuda")
        tex = torch.zeros((8, 8), dtype=torch.int64, device="cuda")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for it in range(3):
            if it == 2:
                ev[0].record()
            pvp.process_image_cpp(img, [0.0, 0.0, 0.0], pointing, 0.8, W, H, grid, ext, float(100 + n + 88), 100 + n + 88, tex,
                                  0.0, 0.0, 1.0, 1.0, False, False, accum_mode="fixed", frac_bits=20, want_stats=(it == 1))
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1])
        print(f"variant {variant} {name}: {ms:.2f} ms, {pi.last_stats['n_samples'] / ms / 1e6:.1f} G samples/s", flush=True)
