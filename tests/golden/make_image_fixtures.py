"""Generates tests/golden/image_fixtures.npz: small PNG / JPEG files and the pixels an
independent decoder (Pillow = libpng / libjpeg-turbo) reads from them, mapped to RGBA8 the way
the reference's `rgba8_image` does (crates/lib/src/loaders/gltf.rs:12-44: a source with c < 4
channels fills the first c bytes of a zeroed texel).

Run here (needs Pillow); the tests only read the committed .npz.
    python tests/golden/make_image_fixtures.py
"""
from __future__ import annotations

import io
import struct
import zlib
from pathlib import Path

import numpy as np
from PIL import Image

OUT = Path(__file__).resolve().parent / "image_fixtures.npz"


def pattern(w, h, channels, seed):
    """smooth gradients + a few hard edges + noise: exercises every PNG filter and the DCT"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    out = np.zeros((h, w, channels))
    for c in range(channels):
        out[..., c] = 127 + 90 * np.sin(x * (0.21 + 0.07 * c) + c) * np.cos(y * (0.17 + 0.05 * c))
    out[h // 3: h // 2, w // 4: w // 2] = 240
    out += rng.normal(0, 6, out.shape)
    return np.clip(out, 0, 255).astype(np.uint8)


def expand(arr):
    """rgba8_image: copy the c source channels into a zeroed RGBA texel"""
    if arr.ndim == 2:
        arr = arr[..., None]
    h, w, c = arr.shape
    out = np.zeros((h, w, 4), dtype=np.uint8)
    out[..., :c] = arr
    return out


def png_bytes(img: Image.Image, **kw) -> bytes:
    b = io.BytesIO()
    img.save(b, format="PNG", **kw)
    return b.getvalue()


def jpeg_bytes(img: Image.Image, **kw) -> bytes:
    b = io.BytesIO()
    img.save(b, format="JPEG", **kw)
    return b.getvalue()


def adam7_png(rgba: np.ndarray) -> bytes:
    """hand-written interlaced RGBA8 PNG (Pillow cannot write Adam7); filter type 0 and 2"""
    h, w, _ = rgba.shape
    x0, y0 = [0, 4, 0, 2, 0, 1, 0], [0, 0, 4, 0, 2, 0, 1]
    dx, dy = [8, 8, 4, 4, 2, 2, 1], [8, 8, 8, 4, 4, 2, 2]
    raw = bytearray()
    for p in range(7):
        sub = rgba[y0[p]::dy[p], x0[p]::dx[p]]
        if sub.size == 0:
            continue
        prev = np.zeros_like(sub[0])
        for r, row in enumerate(sub):
            if r % 2:  # Up filter on odd rows
                raw.append(2)
                raw += ((row.astype(np.int16) - prev.astype(np.int16)) & 0xFF).astype(np.uint8).tobytes()
            else:
                raw.append(0)
                raw += row.tobytes()
            prev = row

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body))
    ihdr = struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 1)
    comp = zlib.compress(bytes(raw), 6)
    # two IDAT chunks: the decoder must concatenate them
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", comp[: len(comp) // 2]) +
            chunk(b"IDAT", comp[len(comp) // 2:]) + chunk(b"IEND", b""))


def main():
    files, expect = {}, {}
    w, h = 37, 23
    rgb = pattern(w, h, 3, 1)
    rgba = pattern(w, h, 4, 2)
    gray = pattern(w, h, 1, 3)[..., 0]
    la = pattern(w, h, 2, 4)

    files["png_rgb8"] = png_bytes(Image.fromarray(rgb, "RGB"))
    expect["png_rgb8"] = expand(rgb)
    files["png_rgba8"] = png_bytes(Image.fromarray(rgba, "RGBA"), compress_level=9)
    expect["png_rgba8"] = rgba
    files["png_rgba8_stored"] = png_bytes(Image.fromarray(rgba, "RGBA"), compress_level=0)
    expect["png_rgba8_stored"] = rgba
    files["png_gray8"] = png_bytes(Image.fromarray(gray, "L"))
    expect["png_gray8"] = expand(gray)
    files["png_la8"] = png_bytes(Image.fromarray(la, "LA"))
    expect["png_la8"] = expand(la)
    pal = Image.fromarray(rgb, "RGB").quantize(colors=13)
    files["png_palette"] = png_bytes(pal)
    expect["png_palette"] = expand(np.asarray(pal.convert("RGB")))
    pal2 = Image.fromarray(rgb, "RGB").quantize(colors=200)
    files["png_palette8"] = png_bytes(pal2)
    expect["png_palette8"] = expand(np.asarray(pal2.convert("RGB")))
    palt = pal.copy()
    files["png_palette_trns"] = png_bytes(palt, transparency=bytes([0, 50, 100, 150, 200]))
    expect["png_palette_trns"] = np.asarray(Image.open(io.BytesIO(files["png_palette_trns"])).convert("RGBA"))
    bw = (gray > 127)
    files["png_gray1"] = png_bytes(Image.fromarray(bw).convert("1"))
    expect["png_gray1"] = expand((bw * 255).astype(np.uint8))
    g16 = (pattern(w, h, 1, 5)[..., 0].astype(np.uint16) << 8) | 0x5A
    files["png_gray16"] = png_bytes(Image.fromarray(g16, "I;16"))
    expect["png_gray16"] = expand((g16 >> 8).astype(np.uint8))
    files["png_adam7"] = adam7_png(rgba)
    expect["png_adam7"] = np.asarray(Image.open(io.BytesIO(files["png_adam7"])).convert("RGBA"))
    assert np.array_equal(expect["png_adam7"], rgba)
    big = pattern(300, 70, 3, 6)  # > 32 KiB raw: back-references across the whole window
    files["png_rgb8_wide"] = png_bytes(Image.fromarray(big, "RGB"), compress_level=9)
    expect["png_rgb8_wide"] = expand(big)

    smooth = pattern(w, h, 3, 7)
    for name, kw in {
        "jpeg_444": dict(quality=92, subsampling=0),
        "jpeg_422": dict(quality=90, subsampling=1),
        "jpeg_420": dict(quality=90, subsampling=2),
        "jpeg_420_q50_opt": dict(quality=50, subsampling=2, optimize=True),
    }.items():
        files[name] = jpeg_bytes(Image.fromarray(smooth, "RGB"), **kw)
        expect[name] = expand(np.asarray(Image.open(io.BytesIO(files[name])).convert("RGB")))
    files["jpeg_gray"] = jpeg_bytes(Image.fromarray(gray, "L"), quality=90)
    expect["jpeg_gray"] = expand(np.asarray(Image.open(io.BytesIO(files["jpeg_gray"]))))
    wide = pattern(130, 50, 3, 8)
    files["jpeg_420_restart"] = jpeg_bytes(Image.fromarray(wide, "RGB"), quality=85, subsampling=2,
                                           restart_marker_blocks=3)
    expect["jpeg_420_restart"] = expand(np.asarray(Image.open(io.BytesIO(files["jpeg_420_restart"])).convert("RGB")))
    files["jpeg_progressive"] = jpeg_bytes(Image.fromarray(smooth, "RGB"), quality=90, progressive=True)

    payload = {}
    for k, v in files.items():
        payload["file_" + k] = np.frombuffer(v, dtype=np.uint8)
    for k, v in expect.items():
        payload["expect_" + k] = np.ascontiguousarray(v)
    np.savez_compressed(OUT, **payload)
    print(OUT, OUT.stat().st_size, "bytes;", len(files), "files")


if __name__ == "__main__":
    main()
