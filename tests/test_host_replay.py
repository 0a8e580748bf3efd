"""Host logic of path A without a GPU: the product's own orchestration sources (csrc/pvp_pipeline.cu + csrc/pvp_imageio.cpp)
are compiled with the device calls replaced by a recorder (tests/host_replay/replay_harness.cpp), and the list of frame pairs they
hand to K1/K2 -- which two images, in which order, with whose pose -- is compared with a Python model of the reference's frame loop
(ray_voxel.cpp:419-429 grouping/sort, :450-537 loop: :454 cameras with < 2 frames, :470-473 skip on load error, :475-481 first
valid frame, :238-243 + :534 size mismatch casts no rays but becomes prev, :486-491 pose of the CURRENT frame), for every decode
thread count, chunk size and frame shard."""
import json
import os
import shutil
import struct
import subprocess
import zlib

import numpy as np
import pytest

from conftest import ROOT

from pixeltovoxelprojector_b200 import synth

CSRC = os.path.join(ROOT, "pixeltovoxelprojector_b200", "csrc")
HARNESS = os.path.join(ROOT, "tests", "host_replay", "replay_harness.cpp")


@pytest.fixture(scope="module")
def replay(tmp_path_factory):
    cxx = shutil.which("g++")
    cuda_inc = "/usr/local/cuda/include"
    if cxx is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path_factory.mktemp("replay") / "replay")
    cmd = [cxx, "-std=c++17", "-O1", "-x", "c++", "-include", "cuda_runtime.h", "-I", CSRC, "-I", cuda_inc, HARNESS, "-o", exe, "-lz", "-lpthread", "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]

    def run(meta, folder, threads=1, chunk=0, shard=(0, 1)):
        r = subprocess.run([exe, meta, folder, str(threads), str(chunk), str(shard[0]), str(shard[1])], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        lines = r.stdout.splitlines()
        assert not [l for l in lines if l.startswith("BAD")], r.stdout
        again = lines.index("AGAIN")                      # the harness runs the loop twice on one context (ring reuse): same record
        assert [l for l in lines[:again] if l.startswith("PAIR")] == [l for l in lines[again + 1:] if l.startswith("PAIR")]
        lines = lines[again + 1:]
        pairs = [l for l in lines if l.startswith("PAIR")]
        rep = [l for l in lines if l.startswith("REPORT")][0].split()
        return pairs, {rep[i]: int(rep[i + 1]) for i in range(1, len(rep), 2)}
    return run


def c_hex(x):
    """printf("%a") of a float promoted to double, as glibc prints it."""
    x = float(np.float32(x))
    if x == 0:
        return "-0x0p+0" if np.signbit(x) else "0x0p+0"
    m, e = np.frexp(abs(x))          # x = m * 2^e, 0.5 <= m < 1  ->  1.xxx * 2^(e-1)
    mant = int(m * (1 << 53)) - (1 << 52)
    digits = f"{mant:013x}".rstrip("0")
    return ("-" if x < 0 else "") + "0x1" + ("." + digits if digits else "") + f"p{e - 1:+d}"


def make_sequence(folder, rng, n_cams=3, frames_per_cam=(9, 1, 41)):
    """Frames of several cameras in shuffled metadata order, with missing files, a corrupt file and size changes."""
    entries, images = [], {}
    for cam, n in zip((5, 2, 0)[:n_cams], frames_per_cam):
        size = (12, 16)
        for k in range(n):
            if rng.random() < 0.12:
                size = (int(rng.integers(6, 14)), int(rng.integers(6, 18)))     # the camera changes resolution
            name = f"cam{cam}_{k:03d}.pgm"
            fate = rng.random()
            img = rng.integers(0, 256, size, dtype=np.uint8)
            if fate < 0.12:
                img = None                                                       # file missing
            elif fate < 0.2:
                open(os.path.join(folder, name), "wb").write(b"P5\n16 12\n255\n" + b"\0" * 7)   # truncated file
                img = None
            else:
                synth.write_pgm(os.path.join(folder, name), img)
            images[name] = img
            entries.append({"camera_index": cam, "frame_index": 3 * k + 1, "camera_position": [float(cam), float(k), -1.5], "yaw": float(k),
                            "pitch": cam + 0.25, "roll": -float(k) / 2, "fov_degrees": 40.0 + cam, "image_file": name})
    order = rng.permutation(len(entries))
    meta = os.path.join(folder, "metadata.json")
    json.dump([entries[i] for i in order], open(meta, "w"))
    return meta, entries, images


def reference_pairs(entries, images, threshold=2.0):
    """The frame loop of ray_voxel.cpp:419-537 as data: one record per detect_motion call that can cast rays."""
    cams = {}
    for e in entries:                                      # :419-422 (std::map: ascending camera id)
        cams.setdefault(e["camera_index"], []).append(e)
    out, n_loaded, n_skipped, n_mismatch = [], 0, 0, 0
    slots = []                                            # (camera, position) of every frame of a camera with >= 2 frames
    for cam in sorted(cams):
        fr = sorted(cams[cam], key=lambda e: e["frame_index"])   # :424-429
        if len(fr) < 2:                                   # :454
            continue
        prev = None
        for pos, e in enumerate(fr):
            slots.append((cam, pos))
            img = images[e["image_file"]]
            if img is None:                               # :470-473
                n_skipped += 1
                continue
            n_loaded += 1
            if prev is None:                              # :475-481
                prev = img
                continue
            if img.shape != prev.shape:                   # :238-243 -> no rays; :534 prev <- curr
                n_mismatch += 1
                prev = img
                continue
            h, w = img.shape
            rec = "PAIR %08x %08x %d %d cam %s %s %s ypr %s %s %s focal %s thr %s dtype 0" % (
                zlib.crc32(prev.tobytes()), zlib.crc32(img.tobytes()), w, h, *(c_hex(c) for c in e["camera_position"]),
                c_hex(e.get("yaw", 0.0)), c_hex(e.get("pitch", 0.0)), c_hex(e.get("roll", 0.0)),      # defaults of :157-163
                c_hex(np.float32(e.get("fov_degrees", 60.0)) * np.float32(65536) + np.float32(w)), c_hex(threshold))
            out.append(((cam, pos), rec))
            prev = img
    return out, slots, {"listed": len(entries), "loaded": n_loaded, "skipped": n_skipped, "pairs": len(out), "mismatch": n_mismatch}


def test_c_hex_matches_python_hex():
    for x in (0.0, 1.0, -1.5, 2.0, 40.0, 3.25, 1e-3, 65536.0 * 45 + 16, -0.0, 7.0, 0.1):
        assert float.fromhex(c_hex(x)) == float(np.float32(x))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_frame_loop_hands_the_reference_pairs_to_the_kernels(replay, tmp_path, seed):
    rng = np.random.default_rng(seed)
    meta, entries, images = make_sequence(str(tmp_path), rng)
    want, slots, want_rep = reference_pairs(entries, images)
    assert len(want) > 10 and want_rep["skipped"] > 0 and want_rep["mismatch"] > 0
    want_lines = [rec for _, rec in want]
    for threads, chunk in ((1, 0), (4, 0), (1, 2), (3, 3), (16, 5), (2, 64)):
        got, rep = replay(meta, str(tmp_path), threads=threads, chunk=chunk)
        assert got == want_lines, (threads, chunk)
        assert rep == want_rep, (threads, chunk)
    # frame sharding: rank r takes slots [S r / n, S (r+1) / n); a pair belongs to the shard that holds its CURRENT frame, its
    # previous frame may lie in an earlier shard (look-back to the last loadable frame)
    S = len(slots)
    for count in (2, 3, 7, S, S + 3):
        union, n_pairs = [], 0
        for rank in range(count):
            lo, hi = S * rank // count, S * (rank + 1) // count
            mine = set(slots[lo:hi])
            got, rep = replay(meta, str(tmp_path), threads=2, chunk=4, shard=(rank, count))
            assert got == [rec for key, rec in want if key in mine], (rank, count)
            union += got
            n_pairs += rep["pairs"]
        assert union == want_lines and n_pairs == len(want_lines), count


def test_frame_loop_edge_cases(replay, tmp_path):
    # a camera whose every file is missing, a camera with one frame, a camera whose first frames are missing
    rng = np.random.default_rng(9)
    entries, images = [], {}
    def add(cam, k, present=True, size=(8, 8)):
        name = f"c{cam}_{k}.pgm"
        img = rng.integers(0, 256, size, dtype=np.uint8) if present else None
        if present:
            synth.write_pgm(str(tmp_path / name), img)
        images[name] = img
        entries.append({"camera_index": cam, "frame_index": k, "camera_position": [0.0, 0.0, float(cam)], "image_file": name})
    for k in range(4): add(0, k, present=False)
    add(1, 0)
    for k in range(5): add(2, k, present=k >= 2)
    for k in range(3): add(3, k, size=(8, 8 + k))      # every pair mismatched
    meta = str(tmp_path / "metadata.json")
    json.dump(entries, open(meta, "w"))
    want, slots, want_rep = reference_pairs(entries, images)
    assert [key for key, _ in want] == [(2, 3), (2, 4)] and want_rep["mismatch"] == 2 and want_rep["skipped"] == 6
    for threads in (1, 5):
        got, rep = replay(meta, str(tmp_path), threads=threads)
        assert got == [rec for _, rec in want] and rep == want_rep
    for rank in range(4):      # the look-back of a shard whose camera has no loadable frame yet finds nothing: first valid frame primes
        got, rep = replay(meta, str(tmp_path), threads=2, shard=(rank, 4))
        lo, hi = len(slots) * rank // 4, len(slots) * (rank + 1) // 4
        assert got == [rec for key, rec in want if key in set(slots[lo:hi])]


def test_frame_loop_is_blind_to_the_container_format(replay, tmp_path):
    """The same frames stored as PGM, PNG and BMP (what stbi_load would read) give the same pairs."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    frames = [rng.integers(0, 256, (10, 14), dtype=np.uint8) for _ in range(7)]
    lists = {}
    for ext in ("pgm", "png", "bmp"):
        d = tmp_path / ext
        d.mkdir()
        entries = []
        for k, img in enumerate(frames):
            name = f"f{k}.{ext}"
            if ext == "pgm":
                synth.write_pgm(str(d / name), img)
            else:
                Image.fromarray(img).save(d / name)
            entries.append({"camera_index": 0, "frame_index": k, "camera_position": [0.0, 1.0, 2.0], "yaw": 3.0 * k, "image_file": name})
        json.dump(entries, open(d / "metadata.json", "w"))
        lists[ext], rep = replay(str(d / "metadata.json"), str(d), threads=3, chunk=3)
        assert rep["pairs"] == 6 and rep["loaded"] == 7
    assert lists["pgm"] == lists["png"] == lists["bmp"] and len(lists["pgm"]) == 6


def test_pgm_header_forms_take_the_same_path(replay, tmp_path):
    """The decode pool reads a plain binary PGM straight into its pinned slot (header parsed from the first 64 bytes, payload by pread);
    headers with comments, long runs of white space, 16-bit samples or a header longer than 64 bytes fall back to the general decoder.
    All forms of the same frames must give the same pairs; a payload one byte short is a failed load in either path."""
    rng = np.random.default_rng(12)
    frames = [rng.integers(0, 256, (9, 13), dtype=np.uint8) for _ in range(8)]

    def write(path, img, form):
        h, w = img.shape
        if form == "plain":
            head, body = b"P5\n%d %d\n255\n" % (w, h), img.tobytes()
        elif form == "comment":
            head, body = b"P5\n# made by a camera\n%d %d\n# another\n255\n" % (w, h), img.tobytes()
        elif form == "spaces":
            head, body = b"P5 \t\r\n%d\n\n%d\t255 " % (w, h), img.tobytes()
        elif form == "long":
            head, body = b"P5\n" + b" " * 80 + b"%d %d\n255\n" % (w, h), img.tobytes()
        elif form == "maxval16":     # 16-bit big-endian samples: stb_image's 16 -> 8 bit conversion keeps the high byte
            head, body = b"P5\n%d %d\n65535\n" % (w, h), (img.astype(np.uint16) * 256 + 7).astype(">u2").tobytes()
        open(path, "wb").write(head + body)

    lists = {}
    for form in ("plain", "comment", "spaces", "long", "maxval16"):
        d = tmp_path / form
        d.mkdir()
        entries = []
        for k, img in enumerate(frames):
            name = f"f{k}.pgm"
            write(str(d / name), img, form)
            entries.append({"camera_index": 0, "frame_index": k, "camera_position": [0.0, 1.0, 2.0], "image_file": name})
        json.dump(entries, open(d / "metadata.json", "w"))
        lists[form], rep = replay(str(d / "metadata.json"), str(d), threads=3, chunk=4)
        assert rep["pairs"] == 7 and rep["loaded"] == 8, form
    for f in lists:
        assert lists[f] == lists["plain"], f
    # truncated payload: one byte short -> the frame is skipped (ray_voxel.cpp:470-473), its neighbours pair up across it
    d = tmp_path / "plain"
    data = open(d / "f3.pgm", "rb").read()
    open(d / "f3.pgm", "wb").write(data[:-1])
    got, rep = replay(str(d / "metadata.json"), str(d), threads=2)
    assert rep["skipped"] == 1 and rep["pairs"] == 6 and rep["loaded"] == 7
