"""ctypes binding of lib/libvcrt.so (the C ABI declared in include/vcrt.h).

There is no fallback: if the CUDA library is missing or fails to load, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VCRT_LIB") or os.path.join(_HERE, "lib", "libvcrt.so")   # VCRT_LIB: development A/B builds only


class VcrtError(RuntimeError):
    """Raised where the reference throws std::runtime_error("failed to ...")."""


class Ubo(C.Structure):
    """std140 uniform block, ray-trace-compute.comp:8-15 / main.cpp:39-47 (32 bytes)."""
    _fields_ = [("camPos", C.c_float * 3), ("time", C.c_float), ("currentSample", C.c_uint32),
                ("numTriangles", C.c_uint32), ("numLights", C.c_uint32), ("numSpheres", C.c_uint32)]


class RenderParams(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "struct_size", "shader", "traversal", "rng_mode", "accum_mode", "trig_mode", "max_bounces", "stack_depth",
        "lights_length", "sample_begin", "sample_count", "tile_rank", "tile_count", "philox_seed", "flags", "_reserved")]


class Counters(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("nodes", C.c_uint64), ("triangles", C.c_uint64), ("kernel_ms", C.c_double),
                ("launches", C.c_uint64), ("trace_ms", C.c_double), ("trace_launches", C.c_uint64), ("traversals", C.c_uint64),
                ("primary_rays", C.c_uint64), ("primary_trace_ms", C.c_double)]


assert C.sizeof(Ubo) == 32 and C.sizeof(RenderParams) == 64

# every entry point include/vcrt.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "vcrt_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "vcrt_destroy": (C.c_int, [_P]),
    "vcrt_last_error": (C.c_char_p, [_P]),
    "vcrt_version": (C.c_char_p, []),
    "vcrt_set_shader": (C.c_int, [_P, C.c_char_p]),
    "vcrt_set_buffer": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "vcrt_set_buffer_device": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "vcrt_set_image_size": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "vcrt_set_ubo": (C.c_int, [_P, C.POINTER(Ubo)]),
    "vcrt_dispatch": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32]),
    "vcrt_render": (C.c_int, [_P, C.POINTER(RenderParams)]),
    "vcrt_clear_accum": (C.c_int, [_P]),
    "vcrt_resolve": (C.c_int, [_P, C.c_uint32, C.c_float]),
    "vcrt_post_process": (C.c_int, [_P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]),
    "vcrt_read_present_rgba8": (C.c_int, [_P, _P, C.c_size_t]),
    "vcrt_read_target_rgba8": (C.c_int, [_P, _P, C.c_size_t]),
    "vcrt_read_accum_rgba8": (C.c_int, [_P, _P, C.c_size_t]),
    "vcrt_read_accum_f32": (C.c_int, [_P, _P, C.c_size_t]),
    "vcrt_read_aov": (C.c_int, [_P, _P, C.c_size_t]),
    "vcrt_write_accum_f32": (C.c_int, [_P, _P, C.c_size_t]),
    "vcrt_device_ptr": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "vcrt_set_option": (C.c_int, [_P, C.c_char_p, C.c_char_p]),
    "vcrt_pack_tiles": (C.c_int, [_P, C.c_int, C.c_uint32, C.c_uint32, _P, C.c_size_t]),
    "vcrt_unpack_tiles": (C.c_int, [_P, C.c_int, C.c_uint32, C.c_uint32, _P, C.c_size_t]),
    "vcrt_get_info": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_size_t]),
    "vcrt_set_stream": (C.c_int, [_P, _P]),
    "vcrt_synchronize": (C.c_int, [_P]),
    "vcrt_get_counters": (C.c_int, [_P, C.POINTER(Counters)]),
    "vcrt_reset_counters": (C.c_int, [_P]),
    "vcrt_frames_begin": (C.c_int, [_P, C.c_uint32]),
    "vcrt_frame_submit": (C.c_int, [_P, C.POINTER(RenderParams), C.c_uint32, C.c_float, _P, C.c_size_t, C.POINTER(C.c_uint32)]),
    "vcrt_frame_dispatch": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, _P, C.c_size_t, C.POINTER(C.c_uint32)]),
    "vcrt_frame_wait": (C.c_int, [_P, C.c_uint32]),
    "vcrt_frames_end": (C.c_int, [_P]),
    "vcrt_alloc_host": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "vcrt_free_host": (C.c_int, [_P]),
    "vcrt_group_unique_id": (C.c_int, [_P]),
    "vcrt_group_create_local": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(_P)]),
    "vcrt_group_create_rank": (C.c_int, [_P, _P, C.c_int, C.c_int, C.POINTER(_P)]),
    "vcrt_group_destroy": (C.c_int, [_P]),
    "vcrt_group_last_error": (C.c_char_p, [_P]),
    "vcrt_group_size": (C.c_int, [_P]),
    "vcrt_group_local_count": (C.c_int, [_P]),
    "vcrt_group_rank": (C.c_int, [_P, C.c_int]),
    "vcrt_group_ctx": (_P, [_P, C.c_int]),
    "vcrt_group_set_shader": (C.c_int, [_P, C.c_char_p]),
    "vcrt_group_set_buffer": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "vcrt_group_set_image_size": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "vcrt_group_set_ubo": (C.c_int, [_P, C.POINTER(Ubo)]),
    "vcrt_group_set_option": (C.c_int, [_P, C.c_char_p, C.c_char_p]),
    "vcrt_group_render": (C.c_int, [_P, C.POINTER(RenderParams), C.c_int, C.c_float]),
    "vcrt_group_read_target_rgba8": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "vcrt_group_synchronize": (C.c_int, [_P]),
}

_lib = None


def lib():
    """Loads libvcrt.so once.  Fails loudly: the product has no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VcrtError("failed to load %s: not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "or `make -C vulkan_compute_ray_tracing_b200/csrc`)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError if a declared symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib
