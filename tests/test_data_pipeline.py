"""Input pipeline (utils.py:12-121, BigGAN.py:771-787).  Host half: the PNG decoder of
libbiggan_io.so against hand-written PNG streams (every filter type, palette bit depths) and against PIL, label files,
the shuffle-and-repeat / batch iterator.  Device half: `bg_image_prepare` executed on the CPU through the host emulation
build of csrc/next_ops.cu against a numpy statement of TF-1's bilinear resize + flip + normalise."""
import ctypes
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

try:
    from biggan_tensorflow_b200 import data as D
except RuntimeError:                       # libbiggan_io.so not built yet (fresh checkout): build, then import
    import __graft_entry__
    __graft_entry__.build()
    from biggan_tensorflow_b200 import data as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _chunk(tag, body):
    return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if pa <= pb and pa <= pc else (b if pb <= pc else c)


def write_png(img, color_type, filters, depth=8, palette=None, packed_rows=None):
    """img uint8 [h, w, c]; filters: one filter type per row (PNG spec 9.2).  packed_rows: pre-packed scanlines (palette < 8 bit)."""
    h, w = img.shape[:2]
    rows = packed_rows if packed_rows is not None else [img[y].reshape(-1).astype(np.int32) for y in range(h)]
    bpp = 1 if packed_rows is not None else img.shape[2]
    raw = bytearray()
    prev = np.zeros(len(rows[0]), np.int32)
    for y, row in enumerate(rows):
        ft = filters[y % len(filters)]
        out = np.zeros_like(row)
        for i in range(len(row)):
            a = row[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = [0, a, b, (a + b) // 2, _paeth(a, b, c)][ft]
            out[i] = (row[i] - pred) & 0xFF
        raw.append(ft)
        raw += bytes(out.astype(np.uint8))
        prev = row
    comp = zlib.compress(bytes(raw), 6)
    body = _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color_type, 0, 0, 0))
    if palette is not None:
        body += _chunk(b"PLTE", bytes(np.asarray(palette, np.uint8).reshape(-1)))
    half = len(comp) // 2                                             # two IDAT chunks: the stream must be concatenated
    body += _chunk(b"tEXt", b"Comment\x00test") + _chunk(b"IDAT", comp[:half]) + _chunk(b"IDAT", comp[half:]) + _chunk(b"IEND", b"")
    return b"\x89PNG\r\n\x1a\n" + body


@pytest.mark.parametrize("color_type,c", [(0, 1), (2, 3), (4, 2), (6, 4)])
def test_png_decoder_every_filter_type(color_type, c):
    rng = np.random.default_rng(c)
    img = rng.integers(0, 256, (11, 7, c), dtype=np.uint8)
    img[3:6] = np.arange(7 * c, dtype=np.uint8).reshape(7, c)          # smooth rows: the predictors actually predict
    png = write_png(img, color_type, filters=[0, 1, 2, 3, 4])
    assert D.png_info(png) == (7, 11, c, 8)
    assert np.array_equal(D.decode_png(png), img)
    want3 = np.repeat(img[..., :1], 3, axis=2) if c <= 2 else img[..., :3]
    assert np.array_equal(D.decode_png(png, channels=3), want3)
    alpha = img[..., -1:] if c in (2, 4) else np.full((11, 7, 1), 255, np.uint8)
    assert np.array_equal(D.decode_png(png, channels=4), np.concatenate([want3, alpha], axis=2))
    if c >= 3:
        with pytest.raises(ValueError, match="gray"):
            D.decode_png(png, channels=1)


@pytest.mark.parametrize("depth", [1, 2, 4, 8])
def test_png_decoder_palette_images(depth):
    rng = np.random.default_rng(depth)
    ncol = 1 << depth
    palette = rng.integers(0, 256, (ncol, 3), dtype=np.uint8)
    h, w = 5, 13                                                       # 13 is not a multiple of the pixels per byte
    idx = rng.integers(0, ncol, (h, w))
    per = 8 // depth
    rows = []
    for y in range(h):
        row = np.zeros((w + per - 1) // per, np.int32)
        for x in range(w):
            row[x // per] |= int(idx[y, x]) << ((per - 1 - x % per) * depth)
        rows.append(row)
    png = write_png(np.zeros((h, w, 1), np.uint8), 3, filters=[0, 2, 4], depth=depth, palette=palette, packed_rows=rows)
    assert D.png_info(png) == (w, h, 3, depth)
    assert np.array_equal(D.decode_png(png, channels=3), palette[idx])


def test_png_decoder_against_pil(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:37, 0:53]
    smooth = np.stack([(yy * 5 + xx) % 256, (xx * 3) % 256, (yy * xx) % 256], axis=2).astype(np.uint8)
    noise = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    for name, arr, mode in (("smooth", smooth, "RGB"), ("noise", noise, "RGB"), ("gray", smooth[..., 0], "L"),
                            ("rgba", np.concatenate([smooth, noise[..., :1]], axis=2), "RGBA")):
        path = str(tmp_path / (name + ".png"))
        Image.fromarray(arr, mode).save(path, optimize=(name != "noise"))
        data = open(path, "rb").read()
        for ch in (0, 3, 4):
            want = np.asarray(Image.open(path).convert({0: mode, 3: "RGB", 4: "RGBA"}[ch]))
            got = D.decode_png(data, channels=ch)
            assert np.array_equal(got.reshape(want.shape), want), (name, ch)
    pal = Image.fromarray(smooth, "RGB").quantize(16)
    pal.save(str(tmp_path / "pal.png"))
    assert np.array_equal(D.decode_png(open(str(tmp_path / "pal.png"), "rb").read(), channels=3), np.asarray(pal.convert("RGB")))


def test_png_decoder_rejects_bad_input():
    img = np.zeros((4, 4, 3), np.uint8)
    png = bytearray(write_png(img, 2, [0]))
    with pytest.raises(ValueError, match="not a PNG"):
        D.decode_png(b"GIF89a" + bytes(60))
    bad = bytearray(png)
    bad[20] ^= 1                                                       # inside IHDR
    with pytest.raises(ValueError, match="CRC"):
        D.decode_png(bytes(bad))
    with pytest.raises(ValueError, match="truncated"):
        D.decode_png(bytes(png[:-20]))
    sixteen = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 16, 2, 0, 0, 0)) + _chunk(b"IDAT", zlib.compress(bytes(26))) + _chunk(b"IEND", b"")
    with pytest.raises(ValueError, match="bit depth"):
        D.decode_png(sixteen)
    interlaced = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 2, 0, 0, 1)) + _chunk(b"IDAT", zlib.compress(bytes(14))) + _chunk(b"IEND", b"")
    with pytest.raises(ValueError, match="interlaced"):
        D.decode_png(interlaced)
    with pytest.raises(ValueError, match="`out`"):
        D.decode_png(bytes(png), out=np.zeros((4, 4, 4), np.uint8))
    # crafted header: dimensions near 2^31 would wrap the size arithmetic -> rejected before any allocation
    for w, h in ((0x7FFFFFFF, 0x7FFFFFFF), (0x7FFFFFFF, 1), (1, 0x7FFFFFFF), (0, 4), (70000, 2)):
        huge = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + _chunk(b"IDAT", zlib.compress(bytes(14))) + _chunk(b"IEND", b"")
        with pytest.raises(ValueError, match="out of range|unsupported PNG header"):
            D.decode_png(huge)
    # a stream longer than the image needs (trailing scanline data) is rejected, not silently truncated
    over = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 2, 0, 0, 0)) + _chunk(b"IDAT", zlib.compress(bytes(14 + 7))) + _chunk(b"IEND", b"")
    with pytest.raises(ValueError, match="inflate"):
        D.decode_png(over)


def _dataset(tmp_path, n, size=6):
    root = tmp_path / "dataset" / "toy"
    root.mkdir(parents=True)
    rows = []
    for i in range(n):
        img = np.full((size, size, 3), i, np.uint8)
        (root / ("img%03d.png" % i)).write_bytes(write_png(img, 2, [1, 2]))
        rows.append("img%03d.png\t%d\t%.1f" % (i, i % 3, i / 10))
    (tmp_path / "labels.tsv").write_text("\n".join(rows) + "\n")
    return str(tmp_path / "dataset"), str(tmp_path / "labels.tsv")


def test_load_data_labels_and_the_shuffle_repeat_batch_iterator(tmp_path):
    root, label_file = _dataset(tmp_path, 10)
    paths, labels = D.load_data("toy", label_file, root=root)
    assert len(paths) == 10 and labels[3] == [0.0, 0.3] and os.path.basename(paths[3]) == "img003.png"
    os.remove(paths[9])
    os.rename(os.path.join(root, "toy", "img008.png"), os.path.join(root, "toy", "zzz.png"))
    with pytest.raises(RuntimeError, match="No label found for file zzz.png"):
        D.load_data("toy", label_file, root=root)
    p2, l2 = D.load_data("toy", label_file, ignore_missing=True, n_labels=2, root=root)
    assert l2[-1] == [0.0, 0.0] and len(p2) == 9
    paths, labels = paths[:8], labels[:8]
    it = D.HostBatches(paths, labels, batch_size=3, seed=5, num_threads=4)
    seen = []
    for _ in range(8):                                                 # 24 samples = 3 epochs of 8, batches straddle epochs
        batch, lab, flips = next(it)
        assert batch.shape == (3, 6, 6, 3) and batch.dtype == np.uint8 and lab.shape == (3, 2) and flips.shape == (3,)
        ids = batch[:, 0, 0, 0].tolist()
        assert all((batch[k] == ids[k]).all() for k in range(3))
        assert np.allclose(lab[:, 0], np.array(ids) % 3) and np.allclose(lab[:, 1], np.array(ids) / 10)
        seen += ids
    assert all(sorted(seen[e * 8:(e + 1) * 8]) == list(range(8)) for e in range(3))       # every epoch is a permutation
    assert seen[:8] != seen[8:16]
    again = D.HostBatches(paths, labels, batch_size=3, seed=5, num_threads=2)
    assert [next(again)[0][:, 0, 0, 0].tolist() for _ in range(8)] == [seen[i:i + 3] for i in range(0, 24, 3)]
    b, l, f = next(D.HostBatches(paths, None, 2, flip=False))
    assert l is None and not f.any()


def test_device_prefetcher_refuses_to_run_without_cuda(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    root, label_file = _dataset(tmp_path, 4)
    paths, labels = D.load_data("toy", label_file, root=root)
    with pytest.raises(RuntimeError, match="no CPU path"):
        D.DevicePrefetcher(D.HostBatches(paths, labels, 2), img_size=8)


def _tf1_resize_flip_normalise(src, flips, S, Cd):
    """numpy statement of resize_images (TF-1 bilinear, align_corners=False) -> flip_left_right -> x / 127.5 - 1."""
    B, h, w, c = src.shape
    out = np.zeros((B, S, S, Cd), np.float64)
    ys, xs = np.arange(S) * (h / S), np.arange(S) * (w / S)
    y0, x0 = np.floor(ys).astype(int), np.floor(xs).astype(int)
    y1, x1 = np.minimum(np.ceil(ys).astype(int), h - 1), np.minimum(np.ceil(xs).astype(int), w - 1)
    ly, lx = (ys - y0)[None, :, None, None], (xs - x0)[None, None, :, None]
    s = src.astype(np.float64)
    top = s[:, y0][:, :, x0] + (s[:, y0][:, :, x1] - s[:, y0][:, :, x0]) * lx
    bot = s[:, y1][:, :, x0] + (s[:, y1][:, :, x1] - s[:, y1][:, :, x0]) * lx
    res = top + (bot - top) * ly
    for b in range(B):
        if flips[b]:
            res[b] = res[b][:, ::-1]
    out[..., :c] = res / 127.5 - 1
    return out


@pytest.mark.parametrize("h,w,S,Cd", [(6, 6, 6, 3), (10, 7, 4, 3), (5, 9, 16, 16), (12, 12, 8, 4)])
def test_image_prepare_kernel_on_the_host_emulation(tmp_path, h, w, S, Cd):
    from biggan_tensorflow_b200._lib import parse_header
    so = str(tmp_path / "next_ops_host.so")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-DBG_HOST_EMULATION", "-x", "c++", "-shared", "-fPIC",
                    os.path.join(ROOT, "biggan_tensorflow_b200", "csrc", "next_ops.cu"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    res, args = parse_header()["bg_image_prepare"]
    lib.bg_image_prepare.restype, lib.bg_image_prepare.argtypes = res, args
    rng = np.random.default_rng(h * 100 + S)
    B, c = 3, 3
    src = rng.integers(0, 256, (B, h, w, c), dtype=np.uint8)
    flips = np.array([0, 1, 1], np.int32)
    dst = np.full((B, S, S, Cd), 9.0, np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.bg_image_prepare(p(src), p(flips), p(dst), B, h, w, c, S, Cd, 0, None) == 0
    want = _tf1_resize_flip_normalise(src, flips, S, Cd)
    assert np.abs(dst - want).max() < 2e-6
    if h == S and w == S:                                              # stored at the training size: an exact rescale of the bytes
        assert np.array_equal(dst[0, ..., :3], (src[0].astype(np.float32) / np.float32(127.5) - np.float32(1.0)))
    dst2 = np.zeros_like(dst)
    assert lib.bg_image_prepare(p(src), None, p(dst2), B, h, w, c, S, Cd, 0, None) == 0   # flip = NULL: no flips
    assert np.abs(dst2 - _tf1_resize_flip_normalise(src, [0, 0, 0], S, Cd)).max() < 2e-6
    assert lib.bg_image_prepare(p(src), None, p(dst2), B, h, w, c, S, 2, 0, None) != 0
