"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/vcrt.h declares, has the
reference's data layouts, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

import vulkan_compute_ray_tracing_b200 as vcrt
from vulkan_compute_ray_tracing_b200 import _native


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vcrt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vcrt_[a-z0-9_]+)\s*\(", src)))


def test_header_compiles_as_c_and_layouts_match(tmp_path):
    """include/vcrt.h is plain C; sizes/offsets equal GpuModel::* (GpuModels.h:26-63, probed in SURVEY 8a A0/A1)."""
    prog = tmp_path / "layout.c"
    prog.write_text('''#include <stdio.h>
#include <stddef.h>
#include "vcrt.h"
int main(void){
 printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(vcrt_material), sizeof(vcrt_triangle), sizeof(vcrt_sphere), sizeof(vcrt_bvh_node), sizeof(vcrt_light), sizeof(vcrt_ubo), sizeof(vcrt_render_params), sizeof(vcrt_aov));
 printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", offsetof(vcrt_material, albedo), offsetof(vcrt_triangle, v1), offsetof(vcrt_triangle, v2), offsetof(vcrt_triangle, materialIndex),
   offsetof(vcrt_sphere, materialIndex), offsetof(vcrt_bvh_node, max), offsetof(vcrt_bvh_node, leftNodeIndex), offsetof(vcrt_bvh_node, rightNodeIndex), offsetof(vcrt_bvh_node, objectIndex), offsetof(vcrt_ubo, currentSample));
 return 0; }''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split("\n")
    assert out[0].split() == ["32", "48", "32", "48", "8", "32", "64", "16"]
    assert out[1].split() == ["16", "16", "32", "44", "16", "16", "28", "32", "36", "16"]


def test_library_exports_every_declared_symbol():
    syms = declared_symbols()
    assert len(syms) >= 20 and set(syms) == set(_native.SIGNATURES)
    lib = C.CDLL(_native.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), s
    assert b"sm_100a" in _native.lib().vcrt_version()


def test_library_is_built_for_sm_100a():
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", _native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_cpu_fallback_without_device(gpu_available):
    """Without a CUDA device the product refuses to run (reference convention: throw 'failed to ...')."""
    if gpu_available:
        pytest.skip("a CUDA device is present")
    ctx = C.c_void_p()
    assert _native.lib().vcrt_create(0, C.byref(ctx)) < 0
    assert b"failed to" in _native.lib().vcrt_last_error(None)
    m = vcrt.ComputeMaterial("shaders/generated/ray-trace-compute.spv")
    ubo = vcrt.BufferUtils.createBundle(vcrt.BufferBundle(1), bytes(32))
    m.addUniformBufferBundle(ubo)
    m.addStorageImage(vcrt.Image(64, 64)); m.addStorageImage(vcrt.Image(64, 64))
    for n in (48, 32, 48, 8, 32):
        m.addStorageBufferBundle(vcrt.BufferUtils.createBundle(vcrt.BufferBundle(1), bytes(n)))
    with pytest.raises(vcrt.VcrtError, match="failed to"):
        vcrt.ComputeModel(m)


def test_product_does_not_reference_the_oracle():
    """The product tree must not include, import, link or load anything under oracle/ or tests/ (comments aside)."""
    pkg = os.path.join(ROOT, "vulkan_compute_ray_tracing_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dp, f)
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".inl")):
                txt = re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S))
            elif f.endswith(".py"):
                txt = re.sub(r"#[^\n]*", "", re.sub(r'""".*?"""', "", open(path).read(), flags=re.S))
            elif f == "Makefile":
                txt = re.sub(r"#[^\n]*", "", open(path).read())
            else:
                continue
            for word in ("oracle", "hostemu", "refharness", "tests/"):
                assert word not in txt, (path, word)
    ldd = subprocess.check_output(["ldd", _native.LIB_PATH]).decode()
    assert "oracle" not in ldd and "hostemu" not in ldd and "vcrt_ref" not in ldd


def test_binding_order_and_shader_names():
    m = vcrt.ComputeMaterial("x/ray-trace-compute-simple.spv")
    b = vcrt.BufferBundle(3)
    vcrt.BufferUtils.createBundle(b, np.arange(32, dtype=np.uint8))
    assert len(b.buffers) == 3 and all(x.size == 32 for x in b.buffers)
    m.addUniformBufferBundle(b, vcrt.VK_SHADER_STAGE_COMPUTE_BIT)
    assert m.getUniformBufferBundles()[0].data is b and m.getUniformBufferBundles()[0].shaderStageFlags == 0x20
    assert m.getStorageImages() == [] and m.getStorageBufferBundles() == []
    with pytest.raises(vcrt.VcrtError, match="failed to"):
        m.bind(None, 0)


def test_camera_matches_reference_defaults():
    """Camera.h:28-134 restated: yaw 180 / pitch 0 looks down -x; W/S/A/D/UP/DOWN move by SPEED * dt along Front/Right/Up."""
    import vulkan_compute_ray_tracing_b200 as vcrt
    c = vcrt.Camera()
    assert np.allclose(c.Front, (-1, 0, 0), atol=1e-6) and np.allclose(c.Right, (0, 0, -1), atol=1e-6) and np.allclose(c.Up, (0, 1, 0), atol=1e-6)
    p0 = c.Position.copy()
    c.ProcessKeyboard("w", 0.1)
    assert np.allclose(c.Position - p0, (-0.25, 0, 0), atol=1e-6)
    c.ProcessKeyboard("d", 0.1); c.ProcessKeyboard("u", 0.2)
    assert np.allclose(c.Position - p0, (-0.25, 0.5, -0.25), atol=1e-6)
    c.ProcessKeyboard(".", 1.0)
    assert np.allclose(c.Position - p0, (-0.25, 0.5, -0.25), atol=1e-6)


def build_cpp_example(dst):
    """examples/headless_main.cpp (the reference's main.cpp without the window) against the header-only facade + libvcrt.so."""
    lib_dir = os.path.join(ROOT, "vulkan_compute_ray_tracing_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", str(dst), os.path.join(ROOT, "examples", "headless_main.cpp"),
                           "-L", lib_dir, "-lvcrt", "-Wl,-rpath," + lib_dir])
    return str(dst)


def test_cpp_facade_example_builds(tmp_path):
    exe = build_cpp_example(tmp_path / "headless_main")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode != 0 and "usage:" in out.stderr          # no arguments: usage text, EXIT_FAILURE (no device touched)

