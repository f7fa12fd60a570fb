"""Minimal lossless PNG reader / writer (8-bit RGB and RGBA, non-interlaced) on zlib + struct: the golden pipeline must not
depend on an imaging library that this image does not ship. Reads every filter type (the reference's `png` crate picks
filters adaptively)."""
import struct
import zlib

import numpy as np

_SIG = b"\x89PNG\r\n\x1a\n"


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def write_png(path, img):
    """img: (H, W, 3 | 4) uint8."""
    h, w, c = img.shape
    assert img.dtype == np.uint8 and c in (3, 4)
    raw = np.zeros((h, 1 + w * c), np.uint8)  # filter type 0 on every scanline
    raw[:, 1:] = img.reshape(h, w * c)
    with open(path, "wb") as f:
        f.write(_SIG)
        f.write(_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 6, 0, 0, 0)))
        f.write(_chunk(b"IDAT", zlib.compress(raw.tobytes(), 6)))
        f.write(_chunk(b"IEND", b""))


def read_png(path):
    """-> (H, W, 3 | 4) uint8."""
    data = open(path, "rb").read()
    assert data[:8] == _SIG, "not a PNG file"
    pos, idat, w = 8, b"", None
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        pos += 12 + n
        if tag == b"IHDR":
            w, h, depth, ctype, _, _, interlace = struct.unpack(">IIBBBBB", body)
            assert depth == 8 and ctype in (2, 6) and interlace == 0, "only 8-bit RGB / RGBA, non-interlaced"
            c = 3 if ctype == 2 else 4
        elif tag == b"IDAT":
            idat += body
        elif tag == b"IEND":
            break
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * c)
    out = np.zeros((h, w * c), np.int32)
    prev = np.zeros(w * c, np.int32)
    for y in range(h):
        ft, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        cur = np.zeros(w * c, np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:  # Sub, Average, Paeth depend on the pixel to the left: byte by byte
            for i in range(w * c):
                a = cur[i - c] if i >= c else 0
                b = prev[i]
                cc = prev[i - c] if i >= c else 0
                if ft == 1:
                    pred = a
                elif ft == 3:
                    pred = (a + b) >> 1
                else:
                    p = a + b - cc
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - cc)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else cc)
                cur[i] = (line[i] + pred) & 255
        out[y] = cur
        prev = cur
    return out.astype(np.uint8).reshape(h, w, c)
