"""Generates tests/golden/fov_golden.npz by executing the REFERENCE's own field-of-view lines (spacevoxelviewer.py:197-215: header
'FOV' | CD matrix | default 2.7 arcmin, then np.deg2rad) on synthetic FITS headers.  The lines sit inside `process_image`, whose
module needs astropy to import, so the statements between `fov = header.get('FOV')` and `fov_rad = np.deg2rad(fov)` are cut out of
the unmodified source text and executed with numpy only.  Run where /root/reference is mounted:
    python tests/golden/make_golden_fov.py"""
import json
import os
import textwrap

import numpy as np

REF = "/root/reference/spacevoxelviewer.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_lines():
    lines = open(REF).read().splitlines()
    a = next(i for i, l in enumerate(lines) if l.strip() == "fov = header.get('FOV')")
    b = next(i for i, l in enumerate(lines) if l.strip() == "fov_rad = np.deg2rad(fov)")
    default = next(l for l in lines if l.startswith("default_fov_arcminutes"))
    return compile(textwrap.dedent("\n".join(lines[a:b + 1])), REF, "exec"), default


def headers():
    rng = np.random.default_rng(77)
    out = [{"FOV": 0.5}, {"FOV": "1.25"}, {"FOV": 3}, {}, {"CD1_1": -2.5e-4, "CD1_2": 1e-6, "CD2_1": 2e-6, "CD2_2": 2.5e-4},
           {"CD1_1": -2.5e-4, "CD1_2": 0.0, "CD2_1": 0.0},                       # incomplete CD matrix -> default
           {"CD1_1": 0.0, "CD1_2": 0.0, "CD2_1": 0.0, "CD2_2": 0.0},             # present but zero -> fov 0
           {"FOV": 0.75, "CD1_1": 1e-3, "CD1_2": 0.0, "CD2_1": 0.0, "CD2_2": 1e-3}]   # FOV wins
    for _ in range(20):
        out.append({k: float(rng.normal(0, 3e-4)) for k in ("CD1_1", "CD1_2", "CD2_1", "CD2_2")})
    return out


def main():
    code, default = reference_lines()
    cases, fov_rad = [], []
    rng = np.random.default_rng(5)
    for h in headers():
        width, height = int(rng.integers(64, 5000)), int(rng.integers(64, 5000))
        ns = {"np": np, "header": h, "width": width, "height": height}
        exec(default, ns)
        exec(code, ns)
        cases.append({"header": h, "width": width, "height": height})
        fov_rad.append(np.float64(ns["fov_rad"]))
    np.savez(os.path.join(HERE, "fov_golden.npz"), cases=json.dumps(cases), fov_rad=np.array(fov_rad, np.float64))
    print("wrote fov_golden.npz:", len(cases), "cases")


if __name__ == "__main__":
    main()
