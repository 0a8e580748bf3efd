"""Frames/s through the ray_voxel CLI boundary (files on disk -> .bin) as a function of the decode threads.
    python tools/cli_throughput.py [--frames 120] [--format pgm|png|jpg]
Writes a synthetic single-camera 1080p sequence (sparse motion), runs the CLI with 1 decode thread and with the default pool,
checks that both .bin files are identical, prints frames/s."""
import argparse
import os
import struct
import subprocess
import sys
import tempfile
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixeltovoxelprojector_b200 import synth  # noqa: E402


def write_png_gray(path, img):
    raw = b"".join(b"\x00" + img[r].tobytes() for r in range(img.shape[0]))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", img.shape[1], img.shape[0], 8, 0, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--format", default="png", choices=["pgm", "png", "jpg"], help="jpg needs PIL (quality 90, 4:2:0)")
    ap.add_argument("--n", type=int, default=512, help="grid size (BASELINE config 2: 256)")
    ap.add_argument("--ab-lib", default=None, help="directory with another build of libpvp_b200.so: extra runs with it, interleaved (A/B)")
    a = ap.parse_args()
    H, W, N, center = 1080, 1920, a.n, (0.0, 0.0, 500.0)
    vs = 1024.0 / N
    exe = os.path.join(ROOT, "pixeltovoxelprojector_b200", "ray_voxel")
    with tempfile.TemporaryDirectory() as d:
        frames = synth.sparse_frames(a.frames, H, W, seed=1, blobs=6)
        poses = synth.orbit_poses(a.frames, N, vs, center, seed=2)
        meta = synth.write_sequence(d, {0: frames}, {0: poses})
        if a.format != "pgm":
            import json
            entries = json.load(open(meta))
            for e in entries:
                src = os.path.join(d, e["image_file"])
                e["image_file"] = e["image_file"].replace(".pgm", "." + a.format)
                if a.format == "png":
                    write_png_gray(os.path.join(d, e["image_file"]), frames[e["frame_index"]])
                else:
                    from PIL import Image
                    Image.fromarray(frames[e["frame_index"]]).save(os.path.join(d, e["image_file"]), quality=90)
                os.remove(src)
            json.dump(entries, open(meta, "w"))
        outs = {}
        plan = [("warm-up", 0, None), ("1 decode thread", 1, None), ("default pool", 0, None)]
        if a.ab_lib:
            plan += [("B: 1 decode thread", 1, a.ab_lib), ("B: default pool", 0, a.ab_lib), ("default pool, again", 0, None)]
        for label, threads, libdir in plan:
            out = os.path.join(d, f"out_{threads}.bin")
            env = dict(os.environ, LD_LIBRARY_PATH=os.path.abspath(libdir)) if libdir else None
            t0 = time.perf_counter()
            r = subprocess.run([exe, meta, d, out, "--n", str(N), "--voxel-size", str(vs), "--quiet", "--stats"] +
                               (["--decode-threads", str(threads)] if threads else []), capture_output=True, text=True, env=env)
            dt = time.perf_counter() - t0
            assert r.returncode == 0, r.stderr
            outs[label] = open(out, "rb").read()
            if label != "warm-up":
                print(f"{a.format} 1080p x {a.frames} frames, {label}: {dt:.2f} s wall = {a.frames / dt:.0f} frames/s (incl. process start, {4 * N ** 3 >> 20} MiB .bin write)  {r.stderr.strip()}", flush=True)
        assert outs["1 decode thread"] == outs["default pool"], ".bin differs between decode-thread settings"
        print("identical .bin for both settings; host cores:", os.cpu_count())


if __name__ == "__main__":
    main()
