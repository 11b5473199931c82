import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

GOLDEN = os.path.join(HERE, "golden")
CAM = (1.8, 8.6, 1.1)  # main.cpp:37


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _make(path, target=None):
    cmd = ["make", "-C", path] + ([target] if target else [])
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle():
    """The restated CPU oracle (oracle/libvcrt_oracle.so); built on demand (plain C, ~1 s)."""
    _make(os.path.join(ROOT, "oracle"))
    from oracleharness import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """oracle/_ref (the reference's own shader compiled as C++); only where it was built (needs /root/reference)."""
    from refharness import Ref, have_ref
    if not have_ref():
        if os.path.isdir("/root/reference"):
            _make(os.path.join(ROOT, "oracle"), "ref")
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return Ref()


@pytest.fixture(scope="session")
def hostemu():
    _make(os.path.join(HERE, "hostemu"))
    from hostemuharness import HostEmu
    return HostEmu()


@pytest.fixture(scope="session")
def doge():
    from refharness import load_scene
    return load_scene(os.path.join(GOLDEN, "doge_scene.vcrt"))


@pytest.fixture(scope="session")
def doge_glass():
    """Bundled scene + box1.obj as glass + box2.obj as metal (tests/golden/make_golden.py)."""
    from refharness import load_scene
    return load_scene(os.path.join(GOLDEN, "doge_glass_scene.vcrt"))


def load_png(name):
    from PIL import Image
    return np.array(Image.open(os.path.join(GOLDEN, name)).convert("RGBA"))


def small_scene(n_tris=64, seed=7, glass=True, metal=True):
    """A seeded random-triangle scene with every material type and a reference-style BVH (via tests/tinybvh.py)."""
    from tinybvh import build_scene
    return build_scene(n_tris, seed, glass=glass, metal=metal)


@pytest.fixture(scope="session")
def gpu_available():
    try:
        import ctypes as C
        from vulkan_compute_ray_tracing_b200 import _native
        ctx = C.c_void_p()
        rc = _native.lib().vcrt_create(0, C.byref(ctx))
        if rc == 0:
            _native.lib().vcrt_destroy(ctx)
        return rc == 0
    except Exception:
        return False
