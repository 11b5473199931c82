#!/usr/bin/env python3
"""Host-side check of the fast tree's quality on the bench scene (C3): node visits / triangle tests per ray of the 4-wide
tree, counted by the host emulation (tests/hostemu) on a small frame.  TEST/DEV TOOL -- uses test infrastructure."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from hostemuharness import HostEmu
from oracleharness import make_params
from vulkan_compute_ray_tracing_b200 import scenegen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
scene = scenegen.generate_box_scene(n, seed=1234)
h = HostEmu()
for res, name in ((8, "q15x4"), (0, "q15")):
    p = make_params(shader="full", traversal="fast", rng="philox", accum="f32", max_bounces=8, sample_count=1, stack_depth=64)
    p._reserved = res
    t0 = time.time()
    r = h.render(scene, (1.8, 8.6, 1.1), 480, 270, p)
    print("%-6s rays %d nodes/ray %.2f tris/ray %.3f  (%.1f s)" % (name, r["rays"], r["nodes"] / r["rays"], r["tris"] / r["rays"], time.time() - t0))
