"""Headless render-to-file: the reference application's job (src/main.cpp: load the scene, accumulate samples, show the
frame) without the window.

  python -m vulkan_compute_ray_tracing_b200.render scene.vcrt out.{ppm,pfm,exr} [--width 1920 --height 1080 --spp 64 ...]
  python -m vulkan_compute_ray_tracing_b200.render --obj-dir resources/models/doge_scene out.ppm      (RtScene.h's bundled scene)

.ppm gets the resolved rgba8 frame after the reference's post-process pass (gamma 2.2, post-process-shader.frag:67-68;
--denoise enables its smartDeNoise call); .pfm / .exr get the linear f32 mean of the samples.
"""
import argparse
import sys
import time

from . import CAMERA_START, BufferBundle, BufferUtils, ComputeMaterial, ComputeModel, Image, load_scene, pack_ubo, render_params
from . import imageio


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("scene", nargs="?", help=".vcrt scene container (see scene.py); or use --obj-dir")
    ap.add_argument("out", help="output file: .ppm (8-bit, post-processed), .pfm or .exr (linear float)")
    ap.add_argument("--obj-dir", help="directory with the reference's doge_scene OBJ files: assemble the scene as RtScene.h does")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--bounces", type=int, default=8, help="NUM_BOUNCES of the shader (ray-trace-compute.comp:313 ships 2)")
    ap.add_argument("--shader", default="full", choices=["full", "simple"])
    ap.add_argument("--camera", type=float, nargs=3, default=list(CAMERA_START), metavar=("X", "Y", "Z"), help="camera.Position (main.cpp:37)")
    ap.add_argument("--denoise", action="store_true", help="run the post-process shader's smartDeNoise pass (mix 0.5)")
    ap.add_argument("--gamma", type=float, default=2.2)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    if a.obj_dir:
        from . import scenegen
        scene = scenegen.load_default_scene(a.obj_dir)
    elif a.scene:
        scene = load_scene(a.scene)
    else:
        ap.error("give a .vcrt scene or --obj-dir")

    ubo = BufferUtils.createBundle(BufferBundle(1), pack_ubo(tuple(a.camera), 0, scene))
    target, accum = Image(a.width, a.height), Image(a.width, a.height)
    mat = ComputeMaterial("resources/shaders/generated/ray-trace-compute%s.spv" % ("-simple" if a.shader == "simple" else ""), device=a.device)
    mat.addUniformBufferBundle(ubo)
    mat.addStorageImage(target)
    mat.addStorageImage(accum)
    for name in ("triangles", "materials", "bvh", "lights", "spheres"):
        mat.addStorageBufferBundle(BufferUtils.createBundle(BufferBundle(1), scene[name]))
    model = ComputeModel(mat)
    p = render_params(shader=a.shader, traversal="fast", rng="philox", accum="f32", max_bounces=a.bounces, sample_count=a.spp)
    t0 = time.perf_counter()
    mat.clearAccum()
    model.renderCommand(None, 0, p)
    ext = a.out.rsplit(".", 1)[-1].lower()
    if ext == "ppm":
        mat.resolve(a.spp, 0.0)
        img = mat.postProcess(mix=0.5 if a.denoise else 0.0, gamma=a.gamma)
        imageio.write_ppm(a.out, img)
    elif ext in ("pfm", "exr"):
        acc = mat.readAccumF32()
        mean = acc[..., :3] / acc[..., 3:4].clip(min=1.0)
        (imageio.write_pfm if ext == "pfm" else imageio.write_exr)(a.out, mean)
    else:
        ap.error("output must end in .ppm, .pfm or .exr")
    c = mat.counters()
    dt = time.perf_counter() - t0
    print("%s: %dx%d, %d spp, depth %d: %d rays, %.1f ms on the GPU (%.0f Mrays/s), %.2f s wall" %
          (a.out, a.width, a.height, a.spp, a.bounces, c.rays, c.kernel_ms, c.rays / max(c.kernel_ms, 1e-9) / 1e3, dt), file=sys.stderr)
    mat.destroy()
    return 0


if __name__ == "__main__":
    sys.exit(main())
