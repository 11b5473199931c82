#!/usr/bin/env python3
"""1-spp frame latency (the reference's own unit, ms/frame: main.cpp:397-413) of the progressive frame loop -- UBO in, one sample
on top of the accumulation, resolve, rgba8 frame back to pinned host memory -- per kernel family.  DEV TOOL.
usage: python tools/latency.py [--triangles N | --bundled] [--frames 64]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vulkan_compute_ray_tracing_b200 as vcrt
from vulkan_compute_ray_tracing_b200 import scenegen, _native

ap = argparse.ArgumentParser()
ap.add_argument("--triangles", type=int, default=1000000)
ap.add_argument("--bundled", action="store_true")
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--bounces", default="8", help="comma-separated depths")
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--in-flight", default="2,3,4", help="comma-separated numbers of frames in flight for the pipelined loop")
a = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scene = vcrt.load_scene(os.path.join(root, "tests", "golden", "doge_scene.vcrt")) if a.bundled else scenegen.generate_box_scene(a.triangles, seed=1234)
w, h = a.width, a.height
ubo = vcrt.BufferUtils.createBundle(vcrt.BufferBundle(1), vcrt.pack_ubo(vcrt.CAMERA_START, 0, scene))
m = vcrt.ComputeMaterial("ray-trace-compute.spv")
m.addUniformBufferBundle(ubo)
m.addStorageImage(vcrt.Image(w, h)); m.addStorageImage(vcrt.Image(w, h))
for n in ("triangles", "materials", "bvh", "lights", "spheres"):
    m.addStorageBufferBundle(vcrt.BufferUtils.createBundle(vcrt.BufferBundle(1), scene[n]))
model = vcrt.ComputeModel(m)
L = _native.lib()
out = torch.empty((h, w, 4), dtype=torch.uint8, pin_memory=True).numpy()
for nb, name, flags, opts in [(int(nb), n_, f_, o_) for nb in a.bounces.split(",") for n_, f_, o_ in (
        ("auto", 0, {}), ("wavefront", 32, {}), ("wavefront with trace timing", 32, {"trace_timing": "on"}), ("megakernel", 16, {}), ("one thread per pixel", 8, {}))]:
    for k, v in {"trace_timing": "auto", "wf_streams": "auto", **opts}.items():
        m.setOption(k, v)
    p = vcrt.render_params(shader="full", traversal="fast", rng="philox", accum="f32", max_bounces=nb, sample_count=1, flags=flags)
    name = "depth %d %s" % (nb, name)

    def frame(k):
        ubo.buffers[0].write(vcrt.pack_ubo(vcrt.CAMERA_START, k, scene))
        p.sample_begin = k
        model.renderCommand(None, 0, p)
        m.resolve(k + 1, 0.0)
        m._check(L.vcrt_read_target_rgba8(m._ctx, out.ctypes.data, out.nbytes))
    m.clearAccum()
    for k in range(4):
        frame(k)
    m.resetCounters()
    t0 = time.perf_counter()
    for k in range(4, 4 + a.frames):
        frame(k)
    dt = time.perf_counter() - t0
    c = m.counters()
    print("%-36s %.3f ms/frame  %.0f Mrays/s  (%d launches/frame, kernels %.3f ms/frame)" % (name, 1e3 * dt / a.frames, c.rays / dt / 1e6, c.launches / a.frames, c.kernel_ms / a.frames), flush=True)


m.setOption("trace_timing", "auto")
# ---- the pipelined loop (vcrt_frames_begin / vcrt_frame_submit / vcrt_frame_wait: the reference's MAX_FRAMES_IN_FLIGHT, main.cpp:68)
for nb in [int(x) for x in a.bounces.split(",")]:
    for name, flags in (("wavefront", 32), ("megakernel", 16), ("one thread per pixel", 8)):
        for nfl in [int(x) for x in a.in_flight.split(",")]:
            p = vcrt.render_params(shader="full", traversal="fast", rng="philox", accum="f32", max_bounces=nb, sample_count=1, flags=flags)
            bufs = [vcrt.PinnedFrame(w, h) for _ in range(nfl)]
            m.clearAccum()
            m.framesBegin(nfl)

            def submit(k):
                ubo.buffers[0].write(vcrt.pack_ubo(vcrt.CAMERA_START, k, scene))
                p.sample_begin = k
                m.frameSubmit(p, total_samples=k + 1, gamma=0.0, out=bufs[k % nfl])   # waits for the slot's fence first
            for k in range(8):
                submit(k)
            m.synchronize()
            t0 = time.perf_counter()
            for k in range(8, 8 + a.frames):
                submit(k)
            for s_ in range(nfl):
                m.frameWait(s_)
            dt = time.perf_counter() - t0
            m.framesEnd()
            for b in bufs:
                b.free()
            print("depth %d %-22s %d frames in flight  %.3f ms/frame" % (nb, name, nfl, 1e3 * dt / a.frames), flush=True)
