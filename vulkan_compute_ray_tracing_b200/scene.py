"""Scene containers for the hot path: the five storage buffers of main.cpp:84-106 as raw reference-layout records.

`.vcrt` file = "VCRTSCN1" | 5 x u32 counts (triangles, materials, bvh, lights, spheres) | 3 x u32 reserved | the arrays.
"""
import struct

import numpy as np

from .api import RECORD_BYTES

ORDER = ("triangles", "materials", "bvh", "lights", "spheres")
CAMERA_START = (1.8, 8.6, 1.1)   # main.cpp:37


def load_scene(path):
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:8] != b"VCRTSCN1":
        raise ValueError("failed to open scene %s: bad magic" % path)
    counts = struct.unpack_from("<5I", blob, 8)
    off, out = 40, {}
    for name, n in zip(ORDER, counts):
        size = n * RECORD_BYTES[name]
        out[name] = np.frombuffer(blob, dtype=np.uint8, count=size, offset=off).copy()
        off += size
    if off != len(blob):
        raise ValueError("failed to open scene %s: truncated or oversized" % path)
    return out


def save_scene(path, scene):
    with open(path, "wb") as f:
        f.write(b"VCRTSCN1")
        f.write(struct.pack("<8I", *[len(scene[n]) // RECORD_BYTES[n] for n in ORDER], 0, 0, 0))
        for n in ORDER:
            f.write(np.ascontiguousarray(scene[n]).view(np.uint8).tobytes())


def pack_ubo(cam_pos, current_sample, scene, time=0.0, num_triangles=None):
    """UniformBufferObject of main.cpp:39-47 / :174.  numTriangles = rtScene.triangles.size() there; pass num_triangles to differ."""
    return struct.pack("<3ffIIII", cam_pos[0], cam_pos[1], cam_pos[2], time, current_sample,
                       len(scene["triangles"]) // 48 if num_triangles is None else num_triangles, len(scene["lights"]) // 8, len(scene["spheres"]) // 32)
