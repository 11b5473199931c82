#!/usr/bin/env python3
"""Generates the committed golden vectors from oracle/_ref (the reference's own shader text and host code
compiled here from /root/reference by oracle/ref/Makefile).  Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden.py

Outputs (all small, committed):
  doge_scene.vcrt                  the five buffers main.cpp:84-106 uploads, produced by the reference's unmodified
                                   RtScene.h/Bvh.h/mesh.cpp (oracle/_ref/ref_scene_dump)
  ref_<variant>_800x600_f<N>.png   rgba8 target image after N frames of the reference frame loop (dispatch + copy),
                                   camera main.cpp:37, full-cover dispatch (25 x 19 groups)
  doge_glass_scene.vcrt            BASELINE config 2's scene variant: the bundled scene plus box1.obj as glass (material 5) and
                                   box2.obj as metal (material 4) -- the meshes RtScene.h:67,69 keeps commented out -- assembled by
                                   the product's scene library (vcrt_scene_load_obj / vcrt_scene_build_bvh), whose output for the
                                   default scene is checked to equal ref_scene_dump's bit for bit
  ref_glass_full_b8_s16_800x600_f2.png   the reference shader's frame for that scene (depth 8, 2 frames)
  ref_post_800x600_f4.png, ref_post_denoise_800x600_f4.png
                                   the reference's post-process fragment shader (post-process-shader.frag compiled by oracle/_ref)
                                   over ref_full_b2_s16_800x600_f4.png: as shipped (gamma 2.2) and with its own commented-out
                                   smartDeNoise line enabled (:64)
  ref_brute_96x64.npz              frames of the shader with its commented-out brute-force hit_scene line enabled (:322), full and
                                   simple, on a 30-triangle scene with one sphere; ubo.numTriangles = all and = 20
  ref_hits.npz                     reference hit_bvh records for 4096 seeded rays (primary + random)
  ref_facts.json                   PCG stream KATs and primary-hit material histograms
"""
import json
import os
import subprocess
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from refharness import Ref, load_scene  # noqa: E402

CAM = (1.8, 8.6, 1.1)


def main():
    root = os.path.dirname(os.path.dirname(HERE))
    subprocess.check_call([os.path.join(root, "oracle", "_ref", "ref_scene_dump"), os.path.join(HERE, "doge_scene.vcrt")])
    scene = load_scene(os.path.join(HERE, "doge_scene.vcrt"))
    ref = Ref()
    for variant, frames in (("full_b2_s16", 1), ("full_b2_s16", 4), ("simple_b4_s16", 1), ("full_b8_s16", 2), ("full_b4_s16", 1)):
        img = ref.render_frames(variant, scene, CAM, 800, 600, frames)
        Image.fromarray(img, "RGBA").save(os.path.join(HERE, "ref_%s_800x600_f%d.png" % (variant, frames)), optimize=True)
    # the reference's literal dispatch extent (main.cpp:228): floor(W/32) x floor(H/32) groups
    img = ref.render_frames("full_b2_s16", scene, CAM, 800, 600, 1, full_cover=False)
    Image.fromarray(img, "RGBA").save(os.path.join(HERE, "ref_full_b2_s16_800x600_f1_refdispatch.png"), optimize=True)
    # lights.length() == 1 (descriptor-range quirk)
    img = ref.render_frames("full_b2_s16", scene, CAM, 800, 600, 2, lights_length=1)
    Image.fromarray(img, "RGBA").save(os.path.join(HERE, "ref_full_b2_s16_800x600_f2_lights1.png"), optimize=True)

    # glass + metal variant (config 2), built by the product's scene library from the reference's OBJ files
    sys.path.insert(0, root)
    from vulkan_compute_ray_tracing_b200 import save_scene, scenegen
    models = "/root/reference/resources/models/doge_scene"
    again = scenegen.load_default_scene(models)
    assert all(np.array_equal(again[k], scene[k]) for k in scene), "scene library differs from the reference's own scene dump"
    glass = scenegen.load_default_scene(models, extra=(("box1.obj", 5), ("box2.obj", 4)))
    save_scene(os.path.join(HERE, "doge_glass_scene.vcrt"), glass)
    img = ref.render_frames("full_b8_s16", glass, CAM, 800, 600, 2)
    Image.fromarray(img, "RGBA").save(os.path.join(HERE, "ref_glass_full_b8_s16_800x600_f2.png"), optimize=True)

    # post-process pass over the 4-frame golden
    f4 = np.array(Image.open(os.path.join(HERE, "ref_full_b2_s16_800x600_f4.png")).convert("RGBA"))
    Image.fromarray(ref.post_process(f4, denoise=False), "RGBA").save(os.path.join(HERE, "ref_post_800x600_f4.png"), optimize=True)
    Image.fromarray(ref.post_process(f4, denoise=True), "RGBA").save(os.path.join(HERE, "ref_post_denoise_800x600_f4.png"), optimize=True)

    # brute-force hit_scene (the shader's commented-out alternative), seeded test scene (tests/tinybvh.py)
    from tinybvh import build_scene
    sc = build_scene(30, 4)
    brute = {}
    for variant in ("full_b2_s16_brute", "simple_b4_s16_brute"):
        for nt in (None, 20):
            brute["%s_%s" % (variant, "all" if nt is None else nt)] = ref.render_frames(variant, sc, (0.0, 6.0, 1.5), 96, 64, 2, num_triangles=nt)
    np.savez_compressed(os.path.join(HERE, "ref_brute_96x64.npz"), **brute)

    rs = np.random.RandomState(1234)
    n = 4096
    rays = np.zeros((n, 6), np.float32)
    # half: rays from the camera through random image points; half: random origins inside the box, random directions
    rays[: n // 2, 0:3] = (-1.1, 1.8, 8.6)
    tgt = rs.uniform([-1.6, 0.0, -3.0], [1.6, 3.1, 0.1], (n // 2, 3))
    d = tgt - rays[: n // 2, 0:3]
    rays[: n // 2, 3:6] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[n // 2:, 0:3] = rs.uniform([-1.4, 0.1, -2.8], [1.4, 2.9, 0.0], (n // 2, 3))
    d = rs.normal(size=(n // 2, 3))
    rays[n // 2:, 3:6] = d / np.linalg.norm(d, axis=1, keepdims=True)
    hits = ref.hit_bvh("full_b2_s16", scene, rays)
    np.savez_compressed(os.path.join(HERE, "ref_hits.npz"), rays=rays, hits=hits)

    facts = {"pcg": {}, "material_histogram": {}}
    for seed in (0, 1, 600, 479999, 0xFFFFFFFF):
        facts["pcg"][str(seed)] = [float(x).hex() for x in ref.random("full_b2_s16", seed, 8)]
    for (w, h) in ((800, 600), (1920, 1080)):
        # primary-hit materials through the reference's hit_bvh on the primary rays of main():352-373
        ys, xs = np.mgrid[0:h, 0:w]
        rays = primary_rays(w, h, xs.ravel(), ys.ravel())
        hr = ref.hit_bvh("full_b2_s16", scene, rays)
        mats = np.where(hr[:, 0] == 1, hr[:, 1], 0xFFFFFFFF)
        facts["material_histogram"]["%dx%d" % (w, h)] = {str(int(m)): int((mats == m).sum()) for m in np.unique(mats)}
    ys, xs = np.mgrid[0:600, 0:800]
    hr = ref.hit_bvh("full_b8_s16", glass, primary_rays(800, 600, xs.ravel(), ys.ravel()))
    mats = np.where(hr[:, 0] == 1, hr[:, 1], 0xFFFFFFFF)
    facts["material_histogram"]["glass_800x600"] = {str(int(m)): int((mats == m).sum()) for m in np.unique(mats)}
    with open(os.path.join(HERE, "ref_facts.json"), "w") as f:
        json.dump(facts, f, indent=1, sort_keys=True)
    print(json.dumps(facts["material_histogram"]))


def primary_rays(w, h, xs, ys):
    """fp32 restatement of ray generation (ray-trace-compute.comp:355-372) -- only used to feed ref_hit_bvh."""
    f = np.float32
    pi = f(3.1415926535897932385)
    theta = f(30.0) * pi / f(180.0)
    hh = np.tan(theta / f(2.0), dtype=np.float32)
    vh = f(2.0) * hh
    vw = f(w) / f(h) * vh
    origin = np.array([-CAM[2], CAM[0], CAM[1]], np.float32)
    llc = np.array([origin[0] - vw / f(2.0), origin[1] - (-vh) / f(2.0), origin[2] - f(1.0)], np.float32)
    u = xs.astype(np.float32) / f(w)
    v = ys.astype(np.float32) / f(h)
    d = np.stack([(llc[0] + u * vw) - origin[0], (llc[1] + v * (-vh)) - origin[1], np.full_like(u, llc[2] - origin[2])], 1).astype(np.float32)
    inv = f(1.0) / np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2], dtype=np.float32)
    d = (d * inv[:, None]).astype(np.float32)
    out = np.zeros((len(xs), 6), np.float32)
    out[:, 0:3] = origin
    out[:, 3:6] = d
    return out


if __name__ == "__main__":
    main()
