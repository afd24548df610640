"""Turns an image the REFERENCE ITSELF rendered — /root/reference/output/output.png, written by /root/reference/src/main.rs:17-20
(Cornell box 300x300, 300 spp, depth 20, background 0.001) — into a small fixture the CPU suite can use where /root/reference
does not exist (the GPU box): tests/golden/reference_output_300.npz = the decoded 8-bit pixels, nothing else.
It is the only artifact of the reference's own renderer that pins BVH + materials + camera + tone map end to end
(tests/test_oracle_vs_reference_png.py). Run here (the container with /root/reference):  python tests/golden/make_reference_png_fixture.py"""
import os
import struct
import sys
import zlib

import numpy as np

SRC = "/root/reference/output/output.png"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_output_300.npz")


def decode_png_rgb8(path):
    """Minimal PNG reader: 8-bit RGB, non-interlaced, all five filter types (PNG spec §9)."""
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, hdr = 8, b"", None
    while pos < len(b):
        n, = struct.unpack(">I", b[pos:pos + 4]); tag = b[pos + 4:pos + 8]; data = b[pos + 8:pos + 8 + n]
        assert zlib.crc32(tag + data) == struct.unpack(">I", b[pos + 8 + n:pos + 12 + n])[0]
        if tag == b"IHDR": hdr = struct.unpack(">IIBBBBB", data)
        if tag == b"IDAT": idat += data
        pos += 12 + n
    w, h, depth, ctype, _, _, interlace = hdr
    assert (depth, ctype, interlace) == (8, 2, 0), hdr
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)
    out = np.zeros((h, 3 * w), np.int32)
    prev = np.zeros(3 * w, np.int32)
    for y in range(h):
        f, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        cur = np.zeros(3 * w, np.int32)
        if f == 0: cur = line
        elif f == 2: cur = (line + prev) & 255
        else:
            for i in range(3 * w):
                a = cur[i - 3] if i >= 3 else 0
                bb = prev[i]
                c = prev[i - 3] if i >= 3 else 0
                if f == 1: p = a
                elif f == 3: p = (a + bb) >> 1
                else:
                    pa, pb, pc = abs(bb - c), abs(a - c), abs(a + bb - 2 * c)
                    p = a if (pa <= pb and pa <= pc) else (bb if pb <= pc else c)
                cur[i] = (line[i] + p) & 255
        out[y] = cur; prev = cur
    return out.astype(np.uint8).reshape(h, w, 3)


if __name__ == "__main__":
    img = decode_png_rgb8(SRC)
    np.savez_compressed(DST, rgb8=img, source=np.array(SRC))
    print("wrote", DST, img.shape, "mean", img.mean(), "bytes", os.path.getsize(DST))
