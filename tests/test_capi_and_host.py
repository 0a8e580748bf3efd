"""CPU-side checks of the product: the C-ABI library loads and exports every symbol include/pvp_b200.h declares,
the host-side logic (pose, metadata, image decode, .bin) matches the oracle, and the product fails loudly
without a GPU.  No compute entry point is called here."""
import ctypes as C
import json
import os
import re
import struct
import subprocess
import zlib

import numpy as np
import pytest

from conftest import ROOT, bits, golden

import pixeltovoxelprojector_b200 as pvp
from pixeltovoxelprojector_b200 import _lib, synth

HEADER = os.path.join(ROOT, "include", "pvp_b200.h")


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(pvp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pvp_b200.h but not exported by libpvp_b200.so"
    assert set(_lib.SIGNATURES) == set(names), set(_lib.SIGNATURES) ^ set(names)
    assert lib.pvp_version() == 100


def test_struct_layouts_match_header_sizes():
    assert C.sizeof(_lib.Pair) == 60 and C.sizeof(_lib.GridDesc) == 40 and C.sizeof(_lib.Stats) == 32
    assert C.sizeof(_lib.FrameInfo) == 4 * 10 + 512
    opt = _lib.RunOptions()
    _lib.load().pvp_run_options_default(C.byref(opt))
    # the hard-coded constants of ray_voxel.cpp:434-447
    assert (opt.grid.n, opt.grid.voxel_size, tuple(opt.grid.center), opt.threshold) == (500, 6.0, (0.0, 0.0, 500.0), 2.0)
    assert (opt.grid.slab_lo, opt.grid.slab_hi, opt.shard_rank, opt.shard_count) == (0, 500, 0, 1)
    assert opt.grid.attenuation_alpha == 0.0 and opt.loader_noise == 0      # reference behaviour by default (:526, no seeded noise)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pvp.Context()
    h = C.c_void_p()
    rc = _lib.load().pvp_ctx_create(0, None, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in _lib.load().pvp_last_error()


def test_host_pose_matches_oracle(oracle):
    g = golden("pose_golden.npz")
    for ypr, R, fov, w, f in zip(g["ypr"], g["R"], g["fov"], g["width"], g["focal"]):
        assert np.array_equal(bits(pvp.rotation_matrix(*ypr)), bits(R))
        assert bits(pvp.focal_length(fov, int(w))) == bits(np.float32(f))
    rng = np.random.default_rng(7)
    for _ in range(500):
        ypr = rng.uniform(-720, 720, 3).astype(np.float32)
        assert np.array_equal(bits(pvp.rotation_matrix(*ypr)), bits(oracle.rotation_matrix(*ypr)))


def test_metadata_defaults_and_order(tmp_path):
    entries = [{"camera_index": 2, "frame_index": 5, "camera_position": [1.5, -2, 3e2], "yaw": 10, "pitch": -1.25, "roll": 0.5,
                "fov_degrees": 75.5, "image_file": "a b/é.pgm", "object_name": "ignored"},
               {},
               {"camera_position": [1, 2], "image_file": "x.pgm", "frame_index": 1.0}]
    p = tmp_path / "m.json"
    p.write_text(json.dumps(entries))
    fr = pvp.load_metadata(str(p))
    assert len(fr) == 3
    assert (fr[0].camera_index, fr[0].frame_index, fr[0].camera_position, fr[0].yaw, fr[0].pitch, fr[0].roll, fr[0].fov_degrees) == \
        (2, 5, (1.5, -2.0, 300.0), 10.0, -1.25, 0.5, 75.5)
    assert fr[0].image_file == "a b/é.pgm"
    # defaults of ray_voxel.cpp:157-163
    assert (fr[1].camera_index, fr[1].frame_index, fr[1].yaw, fr[1].pitch, fr[1].roll, fr[1].fov_degrees, fr[1].image_file) == (0, 0, 0, 0, 0, 60.0, "")
    assert fr[2].camera_position == (0.0, 0.0, 0.0) and fr[2].frame_index == 1  # short array is ignored (:168)
    (tmp_path / "bad.json").write_text('{"not": "an array"}')
    with pytest.raises(pvp.PvpError, match="top level is not an array"):
        pvp.load_metadata(str(tmp_path / "bad.json"))
    with pytest.raises(pvp.PvpError, match="Cannot open"):
        pvp.load_metadata(str(tmp_path / "missing.json"))


def _png_scanlines(arr, ctype, depth, rng):
    """Filtered scanlines (filter byte + data per row) of one image or one Adam7 pass; random filter types."""
    h, w = arr.shape[:2]
    raw = b""
    rows = arr.reshape(h, -1)
    if depth == 16:
        rows = rows.astype(">u2").view(np.uint8).reshape(h, -1)
    elif depth < 8:
        per = 8 // depth
        pad = (-rows.shape[1]) % per
        r = np.pad(rows, ((0, 0), (0, pad))).reshape(h, -1, per).astype(np.uint32)
        rows = sum(r[:, :, k] << (8 - depth * (k + 1)) for k in range(per)).astype(np.uint8)
    prev = np.zeros(rows.shape[1], np.int32)
    bpp = max(1, {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype] * depth // 8)
    for y in range(h):  # exercise all five filter types
        cur = rows[y].astype(np.int32)
        ft = int(rng.integers(0, 5))
        left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])[:len(cur)]
        ul = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])[:len(cur)]
        if ft == 0: f = cur
        elif ft == 1: f = cur - left
        elif ft == 2: f = cur - prev
        elif ft == 3: f = cur - ((left + prev) >> 1)
        else:
            p = left + prev - ul
            pa, pb, pc = abs(p - left), abs(p - prev), abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
            f = cur - pred
        raw += bytes([ft]) + (f & 255).astype(np.uint8).tobytes()
        prev = cur
    return raw


ADAM7 = ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2))


def _png(path, arr, ctype, depth=8, palette=None, interlace=False):
    h, w = arr.shape[:2]
    rng = np.random.default_rng(0)
    if interlace:   # seven reduced images, each with its own scanlines; empty passes contribute nothing
        raw = b"".join(_png_scanlines(arr[y0::dy, x0::dx], ctype, depth, rng) for x0, y0, dx, dy in ADAM7 if x0 < w and y0 < h)
    else:
        raw = _png_scanlines(arr, ctype, depth, rng)

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    comp = zlib.compress(raw)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, int(interlace)))
    if palette is not None:
        out += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    out += chunk(b"IDAT", comp[: len(comp) // 2]) + chunk(b"IDAT", comp[len(comp) // 2:]) + chunk(b"IEND", b"")
    open(path, "wb").write(out)


def luma(r, g, b):
    return ((r.astype(np.uint32) * 77 + g.astype(np.uint32) * 150 + b.astype(np.uint32) * 29) >> 8)


def test_image_decode_pnm_png(tmp_path):
    rng = np.random.default_rng(3)
    gray = rng.integers(0, 256, (13, 17), dtype=np.uint8)
    rgb = rng.integers(0, 256, (13, 17, 3), dtype=np.uint8)
    synth.write_pgm(str(tmp_path / "g.pgm"), gray)
    synth.write_ppm(str(tmp_path / "c.ppm"), rgb)
    assert np.array_equal(pvp.load_image_gray(str(tmp_path / "g.pgm")), gray)
    want = luma(rgb[..., 0], rgb[..., 1], rgb[..., 2]).astype(np.uint8)   # stb_image's integer luma
    assert np.array_equal(pvp.load_image_gray(str(tmp_path / "c.ppm")), want)
    (tmp_path / "h.pgm").write_bytes(b"P5\n# comment\n17 13\n# another\n255\n" + gray.tobytes())
    assert np.array_equal(pvp.load_image_gray(str(tmp_path / "h.pgm")), gray)
    g16 = rng.integers(0, 65536, (5, 7)).astype(np.uint16)
    (tmp_path / "g16.pgm").write_bytes(b"P5\n7 5\n65535\n" + g16.astype(">u2").tobytes())
    assert np.array_equal(pvp.load_image_gray(str(tmp_path / "g16.pgm")), (g16 >> 8).astype(np.uint8))
    # PNG: gray8, rgb8, rgba8, gray+alpha, gray16, rgb16, palette8, gray4, gray1
    _png(tmp_path / "g.png", gray, 0)
    assert np.array_equal(pvp.load_image_gray(str(tmp_path / "g.png")), gray)
    _png(tmp_path / "c.png", rgb, 2)
    assert np.array_equal(pvp.load_image_gray(str(tmp_path / "c.png")), want)
    rgba = np.concatenate([rgb, rng.integers(0, 256, (13, 17, 1), dtype=np.uint8)], -1)
    _png(tmp_path / "ca.png", rgba, 6)
    assert np.array_equal(pvp.load_image_gray(str(tmp_path / "ca.png")), want)
    ga = np.stack([gray, 255 - gray], -1)
    _png(tmp_path / "ga.png", ga, 4)
    assert np.array_equal(pvp.load_image_gray(str(tmp_path / "ga.png")), gray)
    g16b = rng.integers(0, 65536, (13, 17)).astype(np.uint16)
    _png(tmp_path / "g16.png", g16b, 0, depth=16)
    assert np.array_equal(pvp.load_image_gray(str(tmp_path / "g16.png")), (g16b >> 8).astype(np.uint8))
    c16 = rng.integers(0, 65536, (13, 17, 3)).astype(np.uint16)
    _png(tmp_path / "c16.png", c16, 2, depth=16)
    assert np.array_equal(pvp.load_image_gray(str(tmp_path / "c16.png")), (luma(c16[..., 0], c16[..., 1], c16[..., 2]) >> 8).astype(np.uint8))
    pal = rng.integers(0, 256, (40, 3), dtype=np.uint8)
    idx = rng.integers(0, 40, (13, 17), dtype=np.uint8)
    _png(tmp_path / "p.png", idx, 3, palette=pal)
    assert np.array_equal(pvp.load_image_gray(str(tmp_path / "p.png")), luma(pal[idx, 0], pal[idx, 1], pal[idx, 2]).astype(np.uint8))
    g4 = rng.integers(0, 16, (13, 17), dtype=np.uint8)
    _png(tmp_path / "g4.png", g4, 0, depth=4)
    assert np.array_equal(pvp.load_image_gray(str(tmp_path / "g4.png")), g4 * 17)
    g1 = rng.integers(0, 2, (13, 17), dtype=np.uint8)
    _png(tmp_path / "g1.png", g1, 0, depth=1)
    assert np.array_equal(pvp.load_image_gray(str(tmp_path / "g1.png")), g1 * 255)
    try:  # cross-check the PNG decoder against an independent one when PIL is present
        from PIL import Image
        Image.fromarray(gray).save(tmp_path / "pil.png")
        assert np.array_equal(pvp.load_image_gray(str(tmp_path / "pil.png")), gray)
        Image.fromarray(rgb).save(tmp_path / "pilc.png", optimize=True)
        assert np.array_equal(pvp.load_image_gray(str(tmp_path / "pilc.png")), want)
    except ImportError:
        pass
    for bad in ("nope.pgm",):
        with pytest.raises(pvp.PvpError, match="Failed to load image"):
            pvp.load_image_gray(str(tmp_path / bad))
    (tmp_path / "junk.jpg").write_bytes(b"\xff\xd8\xff\xe0 not really a jpeg")   # SOI + a segment that runs past the end
    with pytest.raises(pvp.PvpError, match="Failed to load image"):
        pvp.load_image_gray(str(tmp_path / "junk.jpg"))


def test_image_decode_interlaced_png_and_bmp(tmp_path):
    """stb_image also reads Adam7-interlaced PNG and uncompressed BMP; same gray bytes as the non-interlaced / PNM file."""
    rng = np.random.default_rng(4)
    for h, w in ((13, 17), (1, 1), (2, 3), (5, 1), (1, 9), (8, 8), (9, 16), (33, 20)):   # incl. sizes with empty Adam7 passes
        gray = rng.integers(0, 256, (h, w), dtype=np.uint8)
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = luma(rgb[..., 0], rgb[..., 1], rgb[..., 2]).astype(np.uint8)
        _png(tmp_path / "ig.png", gray, 0, interlace=True)
        assert np.array_equal(pvp.load_image_gray(str(tmp_path / "ig.png")), gray), (h, w)
        _png(tmp_path / "ic.png", rgb, 2, interlace=True)
        assert np.array_equal(pvp.load_image_gray(str(tmp_path / "ic.png")), want), (h, w)
        rgba = np.concatenate([rgb, rng.integers(0, 256, (h, w, 1), dtype=np.uint8)], -1)
        _png(tmp_path / "ica.png", rgba, 6, interlace=True)
        assert np.array_equal(pvp.load_image_gray(str(tmp_path / "ica.png")), want), (h, w)
        g16 = rng.integers(0, 65536, (h, w)).astype(np.uint16)
        _png(tmp_path / "ig16.png", g16, 0, depth=16, interlace=True)
        assert np.array_equal(pvp.load_image_gray(str(tmp_path / "ig16.png")), (g16 >> 8).astype(np.uint8)), (h, w)
        for depth in (1, 2, 4):
            gd = rng.integers(0, 1 << depth, (h, w), dtype=np.uint8)
            _png(tmp_path / "igd.png", gd, 0, depth=depth, interlace=True)
            assert np.array_equal(pvp.load_image_gray(str(tmp_path / "igd.png")), gd * (255 // ((1 << depth) - 1))), (h, w, depth)
        pal = rng.integers(0, 256, (16, 3), dtype=np.uint8)
        idx = rng.integers(0, 16, (h, w), dtype=np.uint8)
        _png(tmp_path / "ip4.png", idx, 3, depth=4, palette=pal, interlace=True)
        assert np.array_equal(pvp.load_image_gray(str(tmp_path / "ip4.png")), luma(pal[idx, 0], pal[idx, 1], pal[idx, 2]).astype(np.uint8)), (h, w)
    Image = pytest.importorskip("PIL.Image")
    # the test's own Adam7 writer agrees with an independent reader
    assert np.array_equal(np.asarray(Image.open(tmp_path / "ig.png")), gray)
    # BMP as PIL writes it: 8-bit palettised gray, 24-bit, 32-bit, 1-bit, 8-bit colour palette
    for h, w in ((13, 17), (1, 1), (4, 6), (7, 33)):
        gray = rng.integers(0, 256, (h, w), dtype=np.uint8)
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = luma(rgb[..., 0], rgb[..., 1], rgb[..., 2]).astype(np.uint8)
        Image.fromarray(gray).save(tmp_path / "g.bmp")
        gg = gray.astype(np.uint32)                      # a gray palette entry (v, v, v) -> (77 v + 150 v + 29 v) >> 8 = v
        assert np.array_equal(pvp.load_image_gray(str(tmp_path / "g.bmp")), luma(gg, gg, gg).astype(np.uint8)) and np.array_equal(luma(gg, gg, gg), gg)
        Image.fromarray(rgb).save(tmp_path / "c.bmp")
        assert np.array_equal(pvp.load_image_gray(str(tmp_path / "c.bmp")), want), (h, w)
        rgba = np.concatenate([rgb, rng.integers(0, 256, (h, w, 1), dtype=np.uint8)], -1)
        Image.fromarray(rgba).save(tmp_path / "ca.bmp")
        assert np.array_equal(pvp.load_image_gray(str(tmp_path / "ca.bmp")), want), (h, w)
        bits1 = rng.integers(0, 2, (h, w), dtype=np.uint8)
        Image.fromarray(bits1 * 255).convert("1").save(tmp_path / "b1.bmp")
        assert np.array_equal(pvp.load_image_gray(str(tmp_path / "b1.bmp")), bits1 * 255), (h, w)
        pimg = Image.fromarray(rgb).quantize(16)
        pimg.save(tmp_path / "p.bmp")
        prgb = np.asarray(pimg.convert("RGB")).astype(np.uint32)
        assert np.array_equal(pvp.load_image_gray(str(tmp_path / "p.bmp")), luma(prgb[..., 0], prgb[..., 1], prgb[..., 2]).astype(np.uint8)), (h, w)
    # top-down BMP (negative height) built by hand from the bottom-up 24-bit file
    data = bytearray((tmp_path / "c.bmp").read_bytes())
    off = struct.unpack_from("<I", data, 10)[0]
    hh = struct.unpack_from("<i", data, 22)[0]
    stride = (w * 3 + 3) // 4 * 4
    rows = [bytes(data[off + r * stride: off + (r + 1) * stride]) for r in range(hh)]
    struct.pack_into("<i", data, 22, -hh)
    data[off:off + hh * stride] = b"".join(reversed(rows))
    (tmp_path / "td.bmp").write_bytes(bytes(data))
    assert np.array_equal(pvp.load_image_gray(str(tmp_path / "td.bmp")), want)
    # RLE-compressed BMP is refused (frame skipped), as documented
    struct.pack_into("<I", data, 30, 1)
    (tmp_path / "rle.bmp").write_bytes(bytes(data))
    with pytest.raises(pvp.PvpError, match="Failed to load image"):
        pvp.load_image_gray(str(tmp_path / "rle.bmp"))


def test_image_decode_jpeg_luma_plane(tmp_path):
    """JPEG (sequential and progressive Huffman): stbi_load(..., 1) returns the luma plane of its integer IDCT.  No stb_image exists in the image and the
    reference pins no version, so this format is UNPINNED; the decoder is held to +-1 grey level of libjpeg's luma plane (two
    integer IDCTs of different fixed-point precision), to exact values where both are exact (flat blocks), and to refusing
    what it does not implement (CMYK)."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(6)

    def picture(h, w, c):
        y, x = np.mgrid[0:h, 0:w]
        img = np.stack([128 + 100 * np.sin(x / 17.0 + k) * np.cos(y / 23.0 - k) + 20 * rng.standard_normal((h, w)) for k in range(c)], -1)
        img[h // 3:h // 2, w // 4:w // 2] = 240
        return np.clip(img, 0, 255).astype(np.uint8)

    def libjpeg_luma(path):
        im = Image.open(path)
        if im.mode != "L":
            im.draft("L", im.size)          # libjpeg hands out the Y plane, no colour conversion
        assert im.mode == "L"
        return np.asarray(im)

    variants = (("L", {}), ("RGB", dict(subsampling=0)), ("RGB", dict(subsampling=1)), ("RGB", dict(subsampling=2)),
                ("RGB", dict(subsampling=2, optimize=True)), ("L", dict(optimize=True, quality=95)), ("RGB", dict(quality=30)),
                ("RGB", dict(quality=100, subsampling=0)), ("RGB", dict(subsampling=2, restart_marker_blocks=3)), ("L", dict(restart_marker_rows=1)))
    f = str(tmp_path / "t.jpg")
    for h, w in ((64, 64), (37, 53), (8, 8), (1, 1), (17, 16), (120, 161)):
        for mode, kw in variants:
            img = picture(h, w, 1 if mode == "L" else 3)
            Image.fromarray(img[..., 0] if mode == "L" else img, mode).save(f, **kw)
            got, want = pvp.load_image_gray(f), libjpeg_luma(f)
            assert got.shape == want.shape == (h, w)
            diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
            assert diff.max() <= 1 and (diff == 0).mean() >= 0.95, (h, w, mode, kw, int(diff.max()), float((diff == 0).mean()))
    # flat 8x8 blocks at quality 100: only DC terms, where every integer IDCT gives 128 + ((dc + 4) >> 3) -- exact equality
    flat = np.kron(rng.integers(0, 256, (5, 7)), np.ones((8, 8))).astype(np.uint8)
    Image.fromarray(flat).save(f, quality=100)
    assert np.array_equal(pvp.load_image_gray(f), libjpeg_luma(f))
    Image.fromarray(np.stack([flat, flat, flat], -1)).save(f, quality=100, subsampling=2)
    assert np.array_equal(pvp.load_image_gray(f), libjpeg_luma(f))
    # progressive files (spectral selection + successive approximation): the same picture stored progressively and sequentially
    # holds the same quantised coefficients, so the two decodes must agree bit for bit; and +-1 of libjpeg as above
    f2 = str(tmp_path / "s.jpg")
    for h, w in ((64, 64), (37, 53), (8, 8), (1, 1), (120, 161)):
        for mode, kw in (("L", {}), ("RGB", dict(subsampling=0)), ("RGB", dict(subsampling=2)), ("RGB", dict(subsampling=1, optimize=True, quality=30)),
                         ("L", dict(quality=95)), ("RGB", dict(subsampling=2, restart_marker_blocks=3))):
            img = picture(h, w, 1 if mode == "L" else 3)
            im = Image.fromarray(img[..., 0] if mode == "L" else img, mode)
            im.save(f, progressive=True, **kw)
            im.save(f2, **{k: v for k, v in kw.items() if not k.startswith("restart")})
            got, want = pvp.load_image_gray(f), libjpeg_luma(f)
            diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
            assert diff.max() <= 1 and (diff == 0).mean() >= 0.95, (h, w, mode, kw)
            assert np.array_equal(got, pvp.load_image_gray(f2)), (h, w, mode, kw)
    # not implemented -> the reference's failed-load path (frame skipped), never a wrong image
    Image.fromarray(picture(32, 32, 4), "CMYK").save(f)
    with pytest.raises(pvp.PvpError, match="Failed to load image"):
        pvp.load_image_gray(f)


def test_image_decode_survives_damaged_files(tmp_path):
    """A damaged or hostile file is the reference's failed stbi_load (frame skipped, ray_voxel.cpp:470-473): the decoder must
    answer "Failed to load image" -- never crash, never allocate what a forged header asks for.  It also runs on the decode
    pool's threads, where an escaping exception would end the process."""
    rng = np.random.default_rng(11)
    gray = rng.integers(0, 256, (24, 31), dtype=np.uint8)
    rgb = rng.integers(0, 256, (24, 31, 3), dtype=np.uint8)
    seeds = {}
    _png(tmp_path / "g.png", gray, 0); seeds["g.png"] = (tmp_path / "g.png").read_bytes()
    _png(tmp_path / "c.png", rgb, 2); seeds["c.png"] = (tmp_path / "c.png").read_bytes()
    _png(tmp_path / "g2.png", gray & 3, 0, depth=2); seeds["g2.png"] = (tmp_path / "g2.png").read_bytes()
    pal = rng.integers(0, 256, (7, 3), dtype=np.uint8)
    _png(tmp_path / "p.png", gray % 7, 3, palette=pal); seeds["p.png"] = (tmp_path / "p.png").read_bytes()
    _png(tmp_path / "ic.png", rgb, 2, interlace=True); seeds["ic.png"] = (tmp_path / "ic.png").read_bytes()
    synth.write_pgm(str(tmp_path / "g.pgm"), gray); seeds["g.pgm"] = (tmp_path / "g.pgm").read_bytes()
    synth.write_ppm(str(tmp_path / "c.ppm"), rgb); seeds["c.ppm"] = (tmp_path / "c.ppm").read_bytes()
    try:
        from PIL import Image
        Image.fromarray(rgb).save(tmp_path / "c.bmp"); seeds["c.bmp"] = (tmp_path / "c.bmp").read_bytes()
        Image.fromarray(rgb).quantize(16).save(tmp_path / "p.bmp"); seeds["p.bmp"] = (tmp_path / "p.bmp").read_bytes()
        Image.fromarray(rgb).save(tmp_path / "c.jpg", subsampling=2, restart_marker_blocks=2); seeds["c.jpg"] = (tmp_path / "c.jpg").read_bytes()
        Image.fromarray(gray).save(tmp_path / "g.jpg", optimize=True); seeds["g.jpg"] = (tmp_path / "g.jpg").read_bytes()
        Image.fromarray(rgb).save(tmp_path / "p.jpg", progressive=True); seeds["p.jpg"] = (tmp_path / "p.jpg").read_bytes()
    except ImportError:
        pass

    def probe(data):
        f = tmp_path / "x.bin"
        f.write_bytes(data)
        try:
            img = pvp.load_image_gray(str(f))
        except pvp.PvpError as e:
            assert "Failed to load image" in str(e)
            return None
        assert img.ndim == 2 and img.size > 0
        return img

    n_ok = 0
    for name, data in seeds.items():
        assert probe(data) is not None, name
        for cut in sorted(set(int(c) for c in rng.integers(0, len(data), 25)) | {0, 1, 8, len(data) - 1}):
            img = probe(data[:cut])       # truncated
            # a cut inside the trailing IEND chunk leaves a complete image; anything shorter must fail
            # (a JPEG cut inside its entropy-coded data still decodes -- zero bits are fed, as stb_image does -- with a wrong tail)
            assert img is None or (name.endswith(".png") and cut >= len(data) - 12) or name.endswith(".jpg"), (name, cut)
        for _ in range(60):               # random byte damage (headers included)
            b = bytearray(data)
            for pos in rng.integers(0, len(b), int(rng.integers(1, 4))):
                b[int(pos)] = int(rng.integers(0, 256))
            n_ok += probe(bytes(b)) is not None
    assert n_ok > 0   # PNG CRCs are not checked (stb_image does not check them either), so pixel damage still decodes

    # forged sizes: headers that announce far more pixels than the file can hold
    g = seeds["g.png"]
    ihdr = g.index(b"IHDR") + 4
    for w, h in ((1 << 24, 1 << 24), (1 << 23, 3), (31, 1 << 24), (0, 24), (31, 0), (0xFFFFFFFF, 0xFFFFFFFF)):
        assert probe(g[:ihdr] + struct.pack(">II", w, h) + g[ihdr + 8:]) is None
    for hdr in (b"P5\n1000000 1000000\n255\n", b"P6\n99999999 99999999\n255\n", b"P5\n31 24\n0\n", b"P5\n31 24\n70000\n",
                b"P5\n-3 24\n255\n", b"P5\n99999999999999999999 24\n255\n", b"P5\n31\n"):
        assert probe(hdr + gray.tobytes()) is None
    assert probe(b"") is None and probe(b"P") is None and probe(b"\x89PNG\r\n\x1a\n") is None


def test_bin_roundtrip_matches_reference_layout(tmp_path):
    rng = np.random.default_rng(0)
    grid = rng.random((7, 7, 7)).astype(np.float32)
    path = str(tmp_path / "g.bin")
    pvp.write_bin(path, grid, 6.0)
    raw = open(path, "rb").read()
    assert len(raw) == 8 + 4 * 7 ** 3                      # ray_voxel.cpp:549-552
    assert struct.unpack("<if", raw[:8]) == (7, 6.0)
    assert np.array_equal(np.frombuffer(raw[8:], np.float32).reshape(7, 7, 7), grid)   # voxelmotionviewer.py:41-51
    back, vs = pvp.read_bin(path)
    assert np.array_equal(back, grid) and vs == np.float32(6.0)
    with pytest.raises(pvp.PvpError, match="Cannot open output file"):
        pvp.write_bin(str(tmp_path / "no_such_dir" / "g.bin"), grid, 1.0)


def test_cli_usage_error_without_gpu():
    exe = os.path.join(ROOT, "pixeltovoxelprojector_b200", "ray_voxel")
    if not os.path.exists(exe):
        pytest.skip("CLI not built")
    res = subprocess.run([exe, "only", "two"], capture_output=True, text=True)
    assert res.returncode == 1 and "Usage:" in res.stderr and "<metadata.json> <image_folder> <output_voxel_bin>" in res.stderr


def test_product_never_imports_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "pixeltovoxelprojector_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libref" not in out


def test_multiply_high_division_used_by_the_k2_slow_path():
    """csrc/pvp_motion.cu::udiv_by decodes a voxel offset with q = umulhi(x, floor(2^32 / d)) and ONE correction step
    (q += x - q d >= d) for d = n and n^2.  The claim "one step suffices for every 32-bit x" restated in numpy over all grid
    sizes the library accepts: random x, the neighbourhood of every kind of multiple of d, and the top of the range."""
    rng = np.random.default_rng(0)
    for n in range(2, 2049):
        for d in (n, n * n):
            m = (1 << 32) // d
            k = rng.integers(0, (1 << 32) // d, 300, dtype=np.uint64)
            xs = np.concatenate([rng.integers(0, 1 << 32, 500, dtype=np.uint64), k * np.uint64(d), k * np.uint64(d) + np.uint64(d - 1),
                                 np.uint64((1 << 32) - 1) - np.arange(0, 300, dtype=np.uint64), np.arange(0, 300, dtype=np.uint64)])
            xs = xs[xs < (1 << 32)]
            q = (xs * np.uint64(m)) >> np.uint64(32)
            q = q + ((xs - q * np.uint64(d)) >= d)
            assert np.array_equal(q, xs // np.uint64(d)), (n, d)


@pytest.mark.gpu
def test_checked_build_catches_an_injected_out_of_bounds_index():
    """libpvp_b200_checked.so (make checked; tools/checked_run.sh runs the whole GPU suite against it) must actually see a bad index:
    told a quarter of the queue's capacity (K1) or half of the grid's span (K3), the kernels skip the offending writes, count them,
    and the call returns PVP_ERR_CHECK.  Runs in a subprocess: the library variant is chosen at import time."""
    import subprocess
    import sys
    checked = os.path.join(ROOT, "pixeltovoxelprojector_b200", "libpvp_b200_checked.so")
    if not os.path.exists(checked):
        pytest.skip("libpvp_b200_checked.so not built (make -C pixeltovoxelprojector_b200/csrc checked)")
    code = r'''
import numpy as np, torch
import pixeltovoxelprojector_b200 as pvp
from pixeltovoxelprojector_b200 import synth, _lib
assert _lib.load().pvp_build_flags() & 1
ctx = pvp.default_context()
frames = torch.as_tensor(synth.dense_frames(3, 64, 96, seed=1), device="cuda")
poses = synth.orbit_poses(3, 32, 8.0, (0.0, 0.0, 500.0), seed=2)
pairs, _ = pvp.make_pairs(poses, 96)
grid = pvp.VoxelGrid(32, 8.0, (0.0, 0.0, 500.0))
assert pvp.accumulate_motion(frames, pairs, grid, 2.0)["n_steps"] > 0           # clean run: no violation
ctx.set_option("inject_fault", 1)
try:
    pvp.accumulate_motion(frames, pairs, grid, 2.0)
    print("K1 MISSED")
except pvp.PvpError as e:
    print("K1 CAUGHT", e.status, "queue write" in str(e))
ctx.set_option("inject_fault", 0)
assert pvp.accumulate_motion(frames, pairs, grid, 2.0)["n_steps"] > 0           # and clean again
img = torch.as_tensor(np.random.default_rng(3).random((32, 32)), device="cuda")
g = torch.zeros((24, 24, 24), dtype=torch.float64, device="cuda"); t = torch.zeros((16, 16), dtype=torch.float64, device="cuda")
call = lambda: pvp.process_image_cpp(img, [0.1, 0.2, -5.0], [0.02, 0.01, 1.0], 0.7, 32, 32, g, [(-10.0, 10.0), (-10.0, 10.0), (20.0, 40.0)], 80.0, 400, t, 1.5, 1.5, 0.8, 0.8, True, False)
call()
ctx.set_option("inject_fault", 2)
try:
    call()
    print("K3 MISSED")
except pvp.PvpError as e:
    print("K3 CAUGHT", e.status, "voxel grid" in str(e))
'''
    env = dict(os.environ, PVP_LIB_VARIANT="checked", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "K1 CAUGHT -6 True" in r.stdout and "K3 CAUGHT -6 True" in r.stdout, r.stdout
