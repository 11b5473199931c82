"""ctypes harness for oracle/_ref/libvcrt_ref.so (the reference's own shader text compiled as C++).

TEST INFRASTRUCTURE ONLY.  Used by tests/ (to pin the restated oracle) and by
tests/golden/make_golden.py (to generate committed fixtures).  The product never imports this.
"""
import ctypes as C
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libvcrt_ref.so")

STRIDES = {"triangles": 48, "materials": 32, "bvh": 48, "lights": 8, "spheres": 32}
BINDING = {"triangles": 3, "materials": 4, "bvh": 5, "lights": 6, "spheres": 7}


class Image2D(C.Structure):
    _fields_ = [("data", C.c_void_p), ("w", C.c_int), ("h", C.c_int)]


class Bindings(C.Structure):
    _fields_ = [("ubo", C.c_void_p), ("images", Image2D * 3), ("ssbo", C.c_void_p * 8),
                ("ssboCount", C.c_int * 8), ("ssboLen", C.c_int * 8)]


def load_scene(path):
    """Reads a .vcrt scene container -> dict of uint8 numpy arrays (raw reference-layout records)."""
    with open(path, "rb") as f:
        blob = f.read()
    assert blob[:8] == b"VCRTSCN1", "bad magic"
    nt, nm, nb, nl, ns = struct.unpack_from("<5I", blob, 8)
    off = 8 + 32
    out = {}
    for name, n in (("triangles", nt), ("materials", nm), ("bvh", nb), ("lights", nl), ("spheres", ns)):
        size = n * STRIDES[name]
        out[name] = np.frombuffer(blob, dtype=np.uint8, count=size, offset=off).copy()
        off += size
    assert off == len(blob)
    return out


def pack_ubo(cam_pos, current_sample, scene, time=0.0, num_triangles=None):
    """std140 UBO of ray-trace-compute.comp:8-15 / main.cpp:39-47 (32 bytes)."""
    return struct.pack("<3ffIIII", cam_pos[0], cam_pos[1], cam_pos[2], time, current_sample,
                       len(scene["triangles"]) // 48 if num_triangles is None else num_triangles, len(scene["lights"]) // 8, len(scene["spheres"]) // 32)


def have_ref():
    return os.path.exists(REF_SO)


class Ref:
    def __init__(self):
        self.lib = C.CDLL(REF_SO)

    def _bind(self, scene, ubo, target, accum, w, h, lights_length=None):
        b = Bindings()
        self._keep = [scene, ubo, target, accum]
        self._ubo = C.create_string_buffer(ubo, 32)
        b.ubo = C.cast(self._ubo, C.c_void_p)
        b.images[1] = Image2D(target.ctypes.data if target is not None else None, w, h)
        b.images[2] = Image2D(accum.ctypes.data if accum is not None else None, w, h)
        for name, idx in BINDING.items():
            arr = scene[name]
            n = len(arr) // STRIDES[name]
            b.ssbo[idx] = arr.ctypes.data
            b.ssboCount[idx] = n
            b.ssboLen[idx] = n
        if lights_length is not None:
            b.ssboLen[6] = lights_length
        return b

    def dispatch(self, variant, scene, ubo, target, accum, gx, gy, lights_length=None):
        """One vkCmdDispatch(gx, gy, 1) of the shader variant; writes `target` (h, w, 4) uint8."""
        h, w = target.shape[:2]
        b = self._bind(scene, ubo, target, accum, w, h, lights_length)
        fn = getattr(self.lib, "ref_dispatch_" + variant)
        fn.argtypes = [C.POINTER(Bindings), C.c_int, C.c_int]
        fn.restype = None
        fn(C.byref(b), gx, gy)

    def render_frames(self, variant, scene, cam_pos, w, h, frames, lights_length=None, full_cover=True, num_triangles=None):
        """The reference frame loop (main.cpp:166-183, :228, :253-261): dispatch, then target->accum copy."""
        target = np.zeros((h, w, 4), np.uint8)
        accum = np.zeros((h, w, 4), np.uint8)
        gx = (w + 31) // 32 if full_cover else w // 32
        gy = (h + 31) // 32 if full_cover else h // 32
        for s in range(frames):
            self.dispatch(variant, scene, pack_ubo(cam_pos, s, scene, num_triangles=num_triangles), target, accum, gx, gy, lights_length)
            accum[...] = target
        return target

    def post_process(self, tex, denoise=False):
        """The reference's post-process fragment shader over the rgba8 image `tex`: as shipped (gamma 2.2 only) or with its own
        commented-out smartDeNoise line enabled (post-process-shader.frag:64: mix 0.5, sigma 2, kSigma 2, threshold 0.05)."""
        tex = np.ascontiguousarray(tex, np.uint8)
        h, w = tex.shape[:2]
        out = np.zeros_like(tex)
        fn = getattr(self.lib, "ref_post_denoise" if denoise else "ref_post")
        fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        fn.restype = None
        fn(tex.ctypes.data, w, h, out.ctypes.data)
        return out

    def hit_bvh(self, variant, scene, org_dir):
        org_dir = np.ascontiguousarray(org_dir, np.float32)
        n = org_dir.shape[0]
        out = np.zeros((n, 10), np.uint32)
        dummy = np.zeros((1, 1, 4), np.uint8)
        b = self._bind(scene, pack_ubo((0, 0, 0), 0, scene), dummy, dummy, 1, 1)
        fn = getattr(self.lib, "ref_hit_bvh_" + variant)
        fn.argtypes = [C.POINTER(Bindings), C.c_void_p, C.c_int, C.c_void_p]
        fn.restype = None
        fn(C.byref(b), org_dir.ctypes.data, n, out.ctypes.data)
        return out

    def random(self, variant, seed, n):
        out = np.zeros(n, np.float32)
        dummy = np.zeros((1, 1, 4), np.uint8)
        scene = {k: np.zeros(STRIDES[k], np.uint8) for k in STRIDES}
        b = self._bind(scene, pack_ubo((0, 0, 0), 0, scene), dummy, dummy, 1, 1)
        fn = getattr(self.lib, "ref_random_" + variant)
        fn.argtypes = [C.POINTER(Bindings), C.c_uint32, C.c_int, C.c_void_p]
        fn.restype = None
        fn(C.byref(b), seed, n, out.ctypes.data)
        return out
