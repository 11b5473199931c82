"""Headless dumps of the rendered frame (there is no swapchain on the GPU box): binary PPM for the rgba8 target / present
image, PFM and OpenEXR for the f32 accumulation buffer.  Pure host-side file writers, no third-party codec.

  write_ppm(path, rgba8)          P6, alpha dropped                                   (H, W, 3|4) uint8
  write_pfm(path, rgb)            Portable Float Map, little endian, bottom row first (H, W, 3|4) float32
  write_exr(path, rgb)            OpenEXR 2.0 single-part scanline file, uncompressed, FLOAT channels B, G, R
  read_exr(path)                  reader for exactly that subset (used by the tests)
"""
import struct

import numpy as np


def write_ppm(path, img):
    img = np.ascontiguousarray(np.asarray(img, np.uint8)[..., :3])
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())


def write_pfm(path, img):
    img = np.asarray(img, "<f4")[..., :3]
    with open(path, "wb") as f:
        f.write(b"PF\n%d %d\n-1.0\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img[::-1]).tobytes())


def _attr(name, typ, payload):
    return name + b"\0" + typ + b"\0" + struct.pack("<i", len(payload)) + payload


def write_exr(path, img):
    """Scanline OpenEXR, NO_COMPRESSION, one chunk per scanline, three FLOAT channels stored in the mandatory alphabetical
    order (B, G, R); data window = display window = the whole image."""
    img = np.asarray(img, "<f4")[..., :3]
    h, w = img.shape[:2]
    chlist = b"".join(c + b"\0" + struct.pack("<iBBBBii", 2, 0, 0, 0, 0, 1, 1) for c in (b"B", b"G", b"R")) + b"\0"
    box = struct.pack("<4i", 0, 0, w - 1, h - 1)
    header = (struct.pack("<I", 20000630) + struct.pack("<I", 2) +
              _attr(b"channels", b"chlist", chlist) +
              _attr(b"compression", b"compression", b"\0") +
              _attr(b"dataWindow", b"box2i", box) +
              _attr(b"displayWindow", b"box2i", box) +
              _attr(b"lineOrder", b"lineOrder", b"\0") +
              _attr(b"pixelAspectRatio", b"float", struct.pack("<f", 1.0)) +
              _attr(b"screenWindowCenter", b"v2f", struct.pack("<2f", 0.0, 0.0)) +
              _attr(b"screenWindowWidth", b"float", struct.pack("<f", 1.0)) + b"\0")
    line_bytes = 3 * w * 4
    first = len(header) + 8 * h
    offsets = first + (8 + line_bytes) * np.arange(h, dtype="<u8")
    # per scanline: y, byte count, then all B, all G, all R of the line
    planar = np.ascontiguousarray(img[..., ::-1].transpose(0, 2, 1))          # (h, 3 = B,G,R, w)
    chunks = np.zeros((h, 8 + line_bytes), np.uint8)
    chunks[:, 0:4] = np.arange(h, dtype="<i4").view(np.uint8).reshape(h, 4)
    chunks[:, 4:8] = np.full(h, line_bytes, "<i4").view(np.uint8).reshape(h, 4)
    chunks[:, 8:] = planar.view(np.uint8).reshape(h, line_bytes)
    with open(path, "wb") as f:
        f.write(header)
        f.write(offsets.tobytes())
        f.write(chunks.tobytes())


def read_exr(path):
    """Reads the subset write_exr produces (uncompressed scanline file, FLOAT channels); returns (H, W, 3) float32 RGB."""
    with open(path, "rb") as f:
        blob = f.read()
    if struct.unpack_from("<I", blob, 0)[0] != 20000630:
        raise ValueError("failed to open %s: not an OpenEXR file" % path)
    if struct.unpack_from("<I", blob, 4)[0] != 2:
        raise ValueError("failed to open %s: only version-2 single-part scanline files are supported" % path)
    pos, attrs = 8, {}
    while blob[pos] != 0:
        end = blob.index(b"\0", pos); name = blob[pos:end]; pos = end + 1
        end = blob.index(b"\0", pos); typ = blob[pos:end]; pos = end + 1
        size = struct.unpack_from("<i", blob, pos)[0]; pos += 4
        attrs[name] = (typ, blob[pos:pos + size]); pos += size
    pos += 1
    if attrs[b"compression"][1] != b"\0":
        raise ValueError("failed to open %s: compressed EXR" % path)
    names, p, ch = [], 0, attrs[b"channels"][1]
    while ch[p] != 0:
        end = ch.index(b"\0", p); names.append(ch[p:end]); p = end + 1
        if struct.unpack_from("<i", ch, p)[0] != 2:
            raise ValueError("failed to open %s: only FLOAT channels are supported" % path)
        p += 16
    x0, y0, x1, y1 = struct.unpack("<4i", attrs[b"dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    offsets = np.frombuffer(blob, "<u8", h, pos)
    out = np.zeros((h, w, 3), np.float32)
    for off in offsets:
        y, nbytes = struct.unpack_from("<ii", blob, int(off))
        line = np.frombuffer(blob, "<f4", nbytes // 4, int(off) + 8).reshape(len(names), w)
        for c, nm in enumerate(names):
            if nm in (b"R", b"G", b"B"):
                out[y - y0, :, (b"R", b"G", b"B").index(nm)] = line[c]
    return out
