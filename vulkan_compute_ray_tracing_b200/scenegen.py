"""Python binding of lib/libvcrt_scene.so (include/vcrt_scene.h): the reference's BVH builder (Bvh.h:141-209), its light
list (RtScene.h:87-96) and the seeded synthetic scenes of BASELINE configs 3-5.  Host C++; no GPU involved."""
import ctypes as C
import os

import numpy as np

from ._native import VcrtError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvcrt_scene.so")
_lib = None

SIGNATURES = {
    "vcrt_scene_build_bvh": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]),
    "vcrt_scene_collect_lights": (C.c_uint32, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]),
    "vcrt_scene_generate_box": (C.c_uint32, [C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "vcrt_scene_load_obj": (C.c_uint32, [C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32]),
    "vcrt_scene_default_materials": (C.c_uint32, [C.c_void_p, C.c_uint32]),
    "vcrt_scene_glibc_rand": (None, [C.c_uint32, C.c_uint32, C.c_void_p]),
    "vcrt_scene_last_error": (C.c_char_p, []),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VcrtError("failed to load %s: not built" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def build_bvh(triangles, axis_seed=0):
    """Bvh::createBvh on raw 48-byte triangle records -> raw 48-byte BvhNode records (uint8 arrays)."""
    tris = np.ascontiguousarray(triangles).view(np.uint8).reshape(-1)
    n = tris.nbytes // 48
    if n == 0:
        raise VcrtError("failed to build BVH: no triangles")
    nodes = np.zeros((2 * n - 1) * 48, np.uint8)
    cnt = C.c_uint32()
    if lib().vcrt_scene_build_bvh(tris.ctypes.data, n, axis_seed, nodes.ctypes.data, C.byref(cnt)) != 0:
        raise VcrtError(lib().vcrt_scene_last_error().decode())
    return nodes[: cnt.value * 48]


def collect_lights(triangles, materials):
    tris = np.ascontiguousarray(triangles).view(np.uint8).reshape(-1)
    mats = np.ascontiguousarray(materials).view(np.uint8).reshape(-1)
    n, nm = tris.nbytes // 48, mats.nbytes // 32
    k = lib().vcrt_scene_collect_lights(tris.ctypes.data, n, mats.ctypes.data, nm, None)
    out = np.zeros(k * 8, np.uint8)
    if k:
        lib().vcrt_scene_collect_lights(tris.ctypes.data, n, mats.ctypes.data, nm, out.ctypes.data)
    return out


def generate_box_scene(target_triangles, seed=1234, axis_seed=0):
    """The synthetic lit-box scene (terrain + displaced spheres), with its reference-layout BVH and light list."""
    tris = np.zeros(target_triangles * 48, np.uint8)
    mats = np.zeros(64 * 32, np.uint8)
    nm = C.c_uint32()
    n = lib().vcrt_scene_generate_box(target_triangles, seed, tris.ctypes.data, target_triangles, mats.ctypes.data, 64, C.byref(nm))
    if n == 0:
        raise VcrtError(lib().vcrt_scene_last_error().decode())
    tris, mats = tris[: n * 48].copy(), mats[: nm.value * 32].copy()
    spheres = np.zeros(32, np.uint8)
    spheres.view(np.float32)[:4] = (0.6, 1.0, -1.0, 0.6)   # RtScene.h:98 (never intersected on the active path)
    spheres.view(np.uint32)[4] = 3
    return {"triangles": tris, "materials": mats, "bvh": build_bvh(tris, axis_seed), "lights": collect_lights(tris, mats), "spheres": spheres}


def load_obj(path, material_index):
    """Triangles of an OBJ file as the reference extracts them (mesh.cpp:96-139 + RtScene.h:13-30): raw 48-byte records."""
    n = lib().vcrt_scene_load_obj(path.encode(), material_index, None, 0)
    if n == 0:
        raise VcrtError(lib().vcrt_scene_last_error().decode())
    out = np.zeros(n * 48, np.uint8)
    lib().vcrt_scene_load_obj(path.encode(), material_index, out.ctypes.data, n)
    return out


def default_materials():
    out = np.zeros(6 * 32, np.uint8)
    lib().vcrt_scene_default_materials(out.ctypes.data, 6)
    return out


# GpuModel::Scene::Scene (RtScene.h:62-85): (file, material) in concatenation order
DEFAULT_SCENE_MESHES = (("buff-doge.obj", 0), ("cheems.obj", 0), ("right.obj", 1), ("left.obj", 2), ("back.obj", 0), ("ceil.obj", 0),
                        ("floor.obj", 0), ("light.obj", 3))


def assemble_scene(meshes, materials=None, axis_seed=0):
    """RtScene-equivalent assembly (RtScene.h:44-101): concatenate meshes, light list, the never-intersected sphere, BVH.
    meshes: iterable of (obj path, material index)."""
    mats = default_materials() if materials is None else np.ascontiguousarray(materials).view(np.uint8).reshape(-1)
    tris = np.concatenate([load_obj(p, m) for p, m in meshes])
    spheres = np.zeros(32, np.uint8)
    spheres.view(np.float32)[:4] = (0.6, 1.0, -1.0, 0.6)   # RtScene.h:98
    spheres.view(np.uint32)[4] = 5
    return {"triangles": tris, "materials": mats, "bvh": build_bvh(tris, axis_seed), "lights": collect_lights(tris, mats), "spheres": spheres}


def load_default_scene(models_dir, extra=()):
    """The bundled scene from its OBJ files (resources/models/doge_scene); `extra` appends (file, material) pairs, e.g.
    (("box1.obj", 5),) for the glass-box variant of BASELINE config 2 (the reference has that line commented out, RtScene.h:67)."""
    meshes = [(os.path.join(models_dir, f), m) for f, m in tuple(DEFAULT_SCENE_MESHES) + tuple(extra)]
    return assemble_scene(meshes)


def glibc_rand(seed, n):
    out = np.zeros(n, np.int32)
    lib().vcrt_scene_glibc_rand(seed, n, out.ctypes.data)
    return out
