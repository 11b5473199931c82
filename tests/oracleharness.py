"""ctypes harness for oracle/libvcrt_oracle.so (the restated CPU oracle).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libvcrt_oracle.so")

SHADER = {"full": 0, "simple": 1}
TRAVERSAL = {"reference": 0, "fast": 1, "brute_force": 2}
RNG = {"pcg_ref": 0, "philox": 1}
ACCUM = {"rgba8_ref": 0, "f32": 1}
TRIG = {"libm": 0, "portable": 1}
FLAG_REF_COVERAGE, FLAG_AOV, FLAG_COUNT = 1, 2, 4

AOV_DTYPE = np.dtype([("triangle", "<i4"), ("material", "<i4"), ("t", "<f4"), ("backFace", "<u4")])


class RenderParams(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "struct_size", "shader", "traversal", "rng_mode", "accum_mode", "trig_mode", "max_bounces", "stack_depth",
        "lights_length", "sample_begin", "sample_count", "tile_rank", "tile_count", "philox_seed", "flags", "_reserved")]


class Ubo(C.Structure):
    _fields_ = [("camPos", C.c_float * 3), ("time", C.c_float), ("currentSample", C.c_uint32),
                ("numTriangles", C.c_uint32), ("numLights", C.c_uint32), ("numSpheres", C.c_uint32)]


class OracleScene(C.Structure):
    _fields_ = [("triangles", C.c_void_p), ("num_triangles", C.c_uint32),
                ("materials", C.c_void_p), ("num_materials", C.c_uint32),
                ("bvh", C.c_void_p), ("num_bvh_nodes", C.c_uint32),
                ("lights", C.c_void_p), ("num_lights", C.c_uint32),
                ("spheres", C.c_void_p), ("num_spheres", C.c_uint32)]


class OracleCounters(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("ref_nodes", C.c_uint64), ("ref_triangles", C.c_uint64),
                ("canon_nodes", C.c_uint64), ("canon_triangles", C.c_uint64), ("max_stack", C.c_uint32), ("_pad", C.c_uint32)]


def make_params(shader="full", traversal="reference", rng="pcg_ref", accum="rgba8_ref", trig="libm", max_bounces=0,
                stack_depth=0, lights_length=0, sample_begin=0, sample_count=1, tile_rank=0, tile_count=0,
                philox_seed=0, flags=0):
    p = RenderParams()
    p.struct_size = C.sizeof(RenderParams)
    p.shader = SHADER[shader] if isinstance(shader, str) else shader
    p.traversal = TRAVERSAL[traversal] if isinstance(traversal, str) else traversal
    p.rng_mode = RNG[rng] if isinstance(rng, str) else rng
    p.accum_mode = ACCUM[accum] if isinstance(accum, str) else accum
    p.trig_mode = TRIG[trig] if isinstance(trig, str) else trig
    p.max_bounces, p.stack_depth, p.lights_length = max_bounces, stack_depth, lights_length
    p.sample_begin, p.sample_count, p.tile_rank, p.tile_count = sample_begin, sample_count, tile_rank, tile_count
    p.philox_seed, p.flags = philox_seed, flags
    return p


def make_ubo(cam_pos, scene, current_sample=0, num_triangles=None):
    u = Ubo()
    u.camPos[0], u.camPos[1], u.camPos[2] = cam_pos
    u.time = 0.0
    u.currentSample = current_sample
    u.numTriangles = len(scene["triangles"]) // 48 if num_triangles is None else num_triangles
    u.numLights = len(scene["lights"]) // 8
    u.numSpheres = len(scene["spheres"]) // 32
    return u


def have_oracle():
    return os.path.exists(ORACLE_SO)


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.vcrt_oracle_render.restype = C.c_int
        self.lib.vcrt_oracle_render.argtypes = [C.POINTER(OracleScene), C.POINTER(Ubo), C.POINTER(RenderParams), C.c_uint32,
                                               C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(OracleCounters)]

    @staticmethod
    def scene_struct(scene):
        s = OracleScene()
        for name, stride in (("triangles", 48), ("materials", 32), ("bvh", 48), ("lights", 8), ("spheres", 32)):
            arr = scene[name]
            setattr(s, name, arr.ctypes.data if len(arr) else None)
        s.num_triangles = len(scene["triangles"]) // 48
        s.num_materials = len(scene["materials"]) // 32
        s.num_bvh_nodes = len(scene["bvh"]) // 48
        s.num_lights = len(scene["lights"]) // 8
        s.num_spheres = len(scene["spheres"]) // 32
        return s

    def render(self, scene, cam_pos, w, h, params, target=None, accum8=None, accumf=None, want_aov=False, num_triangles=None):
        """Returns dict(target, accum8, accumf, aov, counters).  num_triangles: ubo.numTriangles when it differs from the buffer length."""
        s = self.scene_struct(scene)
        u = make_ubo(cam_pos, scene, num_triangles=num_triangles)
        if target is None:
            target = np.zeros((h, w, 4), np.uint8)
        if accum8 is None:
            accum8 = np.zeros((h, w, 4), np.uint8)
        if accumf is None:
            accumf = np.zeros((h, w, 4), np.float32)
        aov = np.zeros((h, w), AOV_DTYPE) if want_aov else None
        if want_aov:
            params.flags |= FLAG_AOV
        cnt = OracleCounters()
        rc = self.lib.vcrt_oracle_render(C.byref(s), C.byref(u), C.byref(params), w, h, target.ctypes.data, accum8.ctypes.data,
                                         accumf.ctypes.data, aov.ctypes.data if want_aov else None, C.byref(cnt))
        if rc != 0:
            raise RuntimeError("vcrt_oracle_render failed: %d" % rc)
        return dict(target=target, accum8=accum8, accumf=accumf, aov=aov, counters=cnt)

    def post_process(self, tex, mix=0.0, sigma=2.0, k_sigma=2.0, threshold=0.05, gamma=2.2):
        tex = np.ascontiguousarray(tex, np.uint8)
        h, w = tex.shape[:2]
        out = np.zeros_like(tex)
        fn = self.lib.vcrt_oracle_post_process
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        assert fn(tex.ctypes.data, w, h, mix, sigma, k_sigma, threshold, gamma, out.ctypes.data) == 0
        return out

    def hit_bvh(self, scene, org_dir, stack_depth=16):
        org_dir = np.ascontiguousarray(org_dir, np.float32)
        n = org_dir.shape[0]
        out = np.zeros((n, 10), np.uint32)
        tri = np.zeros(n, np.int32)
        s = self.scene_struct(scene)
        self.lib.vcrt_oracle_hit_bvh.argtypes = [C.POINTER(OracleScene), C.c_uint32, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        rc = self.lib.vcrt_oracle_hit_bvh(C.byref(s), stack_depth, org_dir.ctypes.data, n, out.ctypes.data, tri.ctypes.data)
        assert rc == 0
        return out, tri

    def random(self, seed, n):
        out = np.zeros(n, np.float32)
        self.lib.vcrt_oracle_random.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
        self.lib.vcrt_oracle_random.restype = None
        self.lib.vcrt_oracle_random(seed, n, out.ctypes.data)
        return out

    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self.lib.vcrt_oracle_philox4x32_10.restype = None
        self.lib.vcrt_oracle_philox4x32_10(c, k, o)
        return list(o)

    def sincos_portable(self, x):
        s, c = C.c_float(), C.c_float()
        self.lib.vcrt_oracle_sincos_portable.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        self.lib.vcrt_oracle_sincos_portable.restype = None
        self.lib.vcrt_oracle_sincos_portable(x, C.byref(s), C.byref(c))
        return s.value, c.value
