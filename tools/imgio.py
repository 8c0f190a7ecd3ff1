"""Tiny PNG writer (no external deps) for eyeballing renders."""
import struct
import zlib

import numpy as np


def tonemap(rgb, exposure=1.0):
    x = np.clip(np.nan_to_num(rgb) * exposure, 0, None)
    x = x / (1.0 + x)
    return (np.clip(x, 0, 1) ** (1 / 2.2) * 255 + 0.5).astype(np.uint8)


def write_png(path, img8):
    h, w, _ = img8.shape
    raw = b"".join(b"\x00" + img8[y].tobytes() for y in range(h))

    def chunk(t, d):
        c = struct.pack(">I", len(d)) + t + d
        return c + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
