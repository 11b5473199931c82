"""ctypes harness for tests/hostemu/libvcrt_hostemu.so: the product's host/device path code compiled for the CPU.
TEST INFRASTRUCTURE ONLY (GPU-less debugging of traversal/shading logic against the oracle)."""
import ctypes as C
import os

import numpy as np

from oracleharness import AOV_DTYPE, RenderParams, Ubo, make_ubo

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("VCRT_HOSTEMU_SO") or os.path.join(HERE, "hostemu", "libvcrt_hostemu.so")   # VCRT_HOSTEMU_SO: development A/B builds


def have_hostemu():
    return os.path.exists(SO)


class HostEmu:
    def __init__(self):
        self.lib = C.CDLL(SO)
        self.lib.hostemu_render.restype = C.c_int

    def render(self, scene, cam_pos, w, h, params, want_aov=False):
        target = np.zeros((h, w, 4), np.uint8)
        accum8 = np.zeros((h, w, 4), np.uint8)
        accumf = np.zeros((h, w, 4), np.float32)
        aov = np.zeros((h, w), AOV_DTYPE)
        if want_aov:
            params.flags |= 2
        u = make_ubo(cam_pos, scene)
        cnt = (C.c_ulonglong * 3)()
        err = C.create_string_buffer(512)
        p = lambda a: C.c_void_p(a.ctypes.data if len(a) else None)
        rc = self.lib.hostemu_render(p(scene["triangles"]), C.c_uint32(len(scene["triangles"]) // 48), p(scene["materials"]),
                                     C.c_uint32(len(scene["materials"]) // 32), p(scene["bvh"]), C.c_uint32(len(scene["bvh"]) // 48),
                                     p(scene["lights"]), C.c_uint32(len(scene["lights"]) // 8), p(scene["spheres"]),
                                     C.c_uint32(len(scene["spheres"]) // 32), C.byref(u), C.byref(params), C.c_uint32(w), C.c_uint32(h),
                                     C.c_void_p(target.ctypes.data), C.c_void_p(accum8.ctypes.data), C.c_void_p(accumf.ctypes.data),
                                     C.c_void_p(aov.ctypes.data), cnt, err, 512)
        if rc != 0:
            raise RuntimeError(err.value.decode())
        return dict(target=target, accum8=accum8, accumf=accumf, aov=aov, rays=cnt[0], nodes=cnt[1], tris=cnt[2])
