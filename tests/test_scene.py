"""Host-side scene producers (include/vcrt_scene.h): reference-identical BVH builder, light list,
OBJ ingestion and default-scene assembly, synthetic scenes.  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

MODELS = "/root/reference/resources/models/doge_scene"


@pytest.fixture(scope="module")
def scenegen():
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "vulkan_compute_ray_tracing_b200", "csrc", "scene")], stdout=subprocess.DEVNULL)
    from vulkan_compute_ray_tracing_b200 import scenegen as sg
    return sg


def test_bvh_builder_reproduces_reference_tree(scenegen, doge):
    """Bvh::createBvh restated (random axis from glibc's unseeded rand(), median split, stack-order numbering): the tree
    the reference built for the bundled scene (fixture dumped by its own code), bit for bit; likewise the light list."""
    assert np.array_equal(scenegen.build_bvh(doge["triangles"]), doge["bvh"])
    assert np.array_equal(scenegen.collect_lights(doge["triangles"], doge["materials"]), doge["lights"])
    assert np.array_equal(scenegen.default_materials(), doge["materials"])


def test_glibc_rand_sequence(scenegen):
    # first outputs of glibc rand() after srand(1) (what an unseeded program sees)
    assert list(scenegen.glibc_rand(1, 5)) == [1804289383, 846930886, 1681692777, 1714636915, 1957747793]


def test_bvh_builder_vs_live_reference(scenegen, ref):
    """Where oracle/_ref is built: the reference's own createBvh on a seeded random soup, bit for bit."""
    import ctypes as C
    from conftest import small_scene
    sc = small_scene(n_tris=3000, seed=21)
    tris = sc["triangles"]
    n = len(tris) // 48
    out = np.zeros((2 * n - 1) * 48, np.uint8)
    fn = ref.lib.ref_create_bvh
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
    for seed in (0, 7):      # 0 = the reference's unseeded sequence
        cnt = fn(tris.ctypes.data, n, seed, out.ctypes.data, 2 * n - 1)
        assert cnt == 2 * n - 1
        assert np.array_equal(scenegen.build_bvh(tris, axis_seed=seed), out[: cnt * 48])


@pytest.mark.skipif(not os.path.isdir(MODELS), reason="the reference's OBJ files are only present in the build container")
def test_obj_ingestion_reproduces_reference_scene(scenegen, doge, doge_glass):
    """vcrt_scene_load_obj + assembly == the reference's tinyobjloader/Mesh/RtScene pipeline on its own OBJ files."""
    sc = scenegen.load_default_scene(MODELS)
    for k in doge:
        assert np.array_equal(sc[k], doge[k]), k
    g = scenegen.load_default_scene(MODELS, extra=(("box1.obj", 5), ("box2.obj", 4)))
    for k in doge_glass:
        assert np.array_equal(g[k], doge_glass[k]), k


def test_obj_reader_details(scenegen, tmp_path):
    p = tmp_path / "quad.obj"
    p.write_text("# comment\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf 1/1/1 2/2/1 3/3/1 4/4/1\nf -4//1 -3//1 -2//1\n")
    t = scenegen.load_obj(str(p), 7).view(np.float32).reshape(-1, 12)
    assert t.shape[0] == 3                                    # quad fanned into 2 + 1 triangle
    assert np.array_equal(t[0, [0, 1, 2, 4, 5, 6, 8, 9, 10]], [0, 0, 0, 1, 0, 0, 1, 1, 0])
    assert np.array_equal(t[1, [0, 1, 2, 4, 5, 6, 8, 9, 10]], [0, 0, 0, 1, 1, 0, 0, 1, 0])
    assert np.array_equal(t[2, [0, 1, 2, 4, 5, 6, 8, 9, 10]], [0, 0, 0, 1, 0, 0, 1, 1, 0])   # negative (relative) indices
    assert (scenegen.load_obj(str(p), 7).view(np.uint32).reshape(-1, 12)[:, 11] == 7).all()
    from vulkan_compute_ray_tracing_b200 import VcrtError
    with pytest.raises(VcrtError, match="failed to open"):
        scenegen.load_obj(str(tmp_path / "missing.obj"), 0)
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nf 1 2 3\n")
    with pytest.raises(VcrtError, match="failed to parse"):
        scenegen.load_obj(str(bad), 0)


def test_synthetic_scene_is_deterministic_and_lit(scenegen, oracle):
    from oracleharness import make_params
    a = scenegen.generate_box_scene(20000, seed=5)
    b = scenegen.generate_box_scene(20000, seed=5)
    c = scenegen.generate_box_scene(20000, seed=6)
    assert all(np.array_equal(a[k], b[k]) for k in a) and not np.array_equal(a["triangles"], c["triangles"])
    n = len(a["triangles"]) // 48
    assert 0.95 * 20000 <= n <= 20000 and len(a["bvh"]) // 48 == 2 * n - 1 and len(a["lights"]) // 8 == 2
    img = oracle.render(a, (1.8, 8.6, 1.1), 96, 64, make_params(shader="full", max_bounces=8, stack_depth=64, accum="f32", sample_count=4))["accumf"]
    assert img[..., :3].mean() > 0.02          # the emitter lights the box


def test_image_writers_roundtrip(tmp_path):
    """Headless dumps: PPM / PFM byte layout and the uncompressed scanline OpenEXR subset (header attributes, offset table,
    B-G-R planar scanlines) read back exactly."""
    import struct
    from vulkan_compute_ray_tracing_b200 import imageio
    rng = np.random.default_rng(3)
    img = rng.random((5, 7, 4)).astype(np.float32) * 3.0
    imageio.write_exr(tmp_path / "a.exr", img)
    back = imageio.read_exr(tmp_path / "a.exr")
    assert back.shape == (5, 7, 3) and np.array_equal(back, img[..., :3])
    blob = (tmp_path / "a.exr").read_bytes()
    assert blob[:4] == bytes([0x76, 0x2F, 0x31, 0x01]) and blob[4:8] == b"\x02\0\0\0"
    assert b"channels\0chlist\0" in blob and b"dataWindow\0box2i\0" in blob and b"compression\0compression\0" in blob
    first = struct.unpack_from("<Q", blob, blob.index(b"screenWindowWidth") + len(b"screenWindowWidth\0float\0") + 8 + 1)[0]
    assert struct.unpack_from("<ii", blob, first) == (0, 3 * 7 * 4) and len(blob) == first + 5 * (8 + 3 * 7 * 4)
    imageio.write_pfm(tmp_path / "a.pfm", img)
    raw = (tmp_path / "a.pfm").read_bytes()
    assert raw.startswith(b"PF\n7 5\n-1.0\n") and np.array_equal(np.frombuffer(raw[len(b"PF\n7 5\n-1.0\n"):], "<f4").reshape(5, 7, 3)[::-1], img[..., :3])
    u8 = (rng.random((4, 6, 4)) * 255).astype(np.uint8)
    imageio.write_ppm(tmp_path / "a.ppm", u8)
    raw = (tmp_path / "a.ppm").read_bytes()
    assert raw.startswith(b"P6\n6 4\n255\n") and raw[len(b"P6\n6 4\n255\n"):] == u8[..., :3].tobytes()
    try:
        import os
        os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
        import cv2
        cvimg = cv2.imread(str(tmp_path / "a.exr"), cv2.IMREAD_UNCHANGED)
    except Exception:
        cvimg = None
    if cvimg is not None:                     # an independent decoder, when this OpenCV build has one
        assert np.array_equal(cvimg[..., ::-1], img[..., :3])
