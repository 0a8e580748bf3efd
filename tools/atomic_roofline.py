"""Measure the L2-atomic / HBM-atomic roofline denominators on this GPU: device throughput of
independent no-return atomic adds for several element types, buffer sizes and address patterns.
Writes one JSON object (stdout, and --out FILE)."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pixeltovoxelprojector_b200 as pvp  # noqa: E402

PATTERNS = {0: "random_per_lane", 1: "random_sector_8lanes", 2: "same_address_per_warp", 3: "random_line_per_warp"}
ELEMS = {4: "f32", 8: "u64", 9: "f64"}


def measure(ctx, buf, nbytes, elem, pattern, n_ops, reps=3):
    lib = ctx._lib
    best = None
    for _ in range(reps + 1):  # first run warms up
        ms = C.c_float()
        pvp._lib.check(lib.pvp_atomic_microbench(ctx.handle, buf.data_ptr(), nbytes, elem, pattern, n_ops, C.byref(ms)))
        best = ms.value if best is None else min(best, ms.value)
    ops = lib.pvp_atomic_microbench_ops(ctx.handle, n_ops)
    return ops / (best * 1e-3)


def run(sizes_mib=(64, 512, 4096), n_ops=1 << 31, quick=False):
    ctx = pvp.default_context()
    out = {"gpu": torch.cuda.get_device_name(0), "unit": "atomic ops/s", "results": {}}
    if quick:
        sizes_mib, n_ops = (64, 512), 1 << 29
    for mib in sizes_mib:
        nbytes = mib << 20
        buf = torch.zeros(nbytes // 8, dtype=torch.int64, device="cuda")
        for elem, ename in ELEMS.items():
            for pat, pname in PATTERNS.items():
                key = f"{ename}/{mib}MiB/{pname}"
                out["results"][key] = measure(ctx, buf, nbytes, elem, pat, n_ops)
                print(f"{key:48s} {out['results'][key] / 1e9:10.2f} G/s", file=sys.stderr)
        del buf
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    res = run(quick=a.quick)
    txt = json.dumps(res, indent=1)
    print(txt)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        open(a.out, "w").write(txt)
