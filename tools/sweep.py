#!/usr/bin/env python3
"""Kernel-only A/B sweeps on the bench workload (C3 scene): prints Mrays/s per option set.
usage: python tools/sweep.py [--spp 16] key=v1,v2 key2=v1,v2 ...   (options of vcrt_set_option)"""
import argparse
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vulkan_compute_ray_tracing_b200 as vcrt
from vulkan_compute_ray_tracing_b200 import scenegen

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--triangles", type=int, default=1000000)
ap.add_argument("--bounces", type=int, default=8)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--count", action="store_true")
ap.add_argument("--trace", action="store_true", help="also print the trace kernel's own time")
ap.add_argument("--flags", type=int, default=0, help="8 static kernel, 16 megakernel")
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("opts", nargs="*")
a = ap.parse_args()
scene = scenegen.generate_box_scene(a.triangles, seed=1234)
w, h = a.width, a.height
ubo = vcrt.BufferUtils.createBundle(vcrt.BufferBundle(1), vcrt.pack_ubo(vcrt.CAMERA_START, 0, scene))
m = vcrt.ComputeMaterial("ray-trace-compute.spv")
m.addUniformBufferBundle(ubo)
m.addStorageImage(vcrt.Image(w, h)); m.addStorageImage(vcrt.Image(w, h))
for n in ("triangles", "materials", "bvh", "lights", "spheres"):
    m.addStorageBufferBundle(vcrt.BufferUtils.createBundle(vcrt.BufferBundle(1), scene[n]))
model = vcrt.ComputeModel(m)
if a.trace:
    m.setOption("trace_timing", "on")   # "auto" leaves 1-spp renders untimed
keys = [o.split("=")[0] for o in a.opts]
vals = [o.split("=")[1].split(",") for o in a.opts]
p = vcrt.render_params(shader="full", traversal="fast", rng="philox", accum="f32", max_bounces=a.bounces, sample_count=a.spp,
                       flags=(vcrt.FLAG_COUNT_TRAVERSAL if a.count else 0) | a.flags)
for combo in itertools.product(*vals) if vals else [()]:
    for k, v in zip(keys, combo):
        m.setOption(k, v)
    best = 0.0
    for rep in range(a.reps + 1):
        m.clearAccum(); m.resetCounters()
        model.renderCommand(None, 0, p)
        c = m.counters()
        if rep:
            best = max(best, c.rays / c.kernel_ms / 1e3)
    extra = " nodes/ray %.1f tris/ray %.2f" % (c.nodes / c.rays, c.triangles / c.rays) if a.count else ""
    if a.trace:
        bounce_ms = c.trace_ms - c.primary_trace_ms
        extra += " trace %.2f ms in %d launches; bounce launches %.2f ms = %.0f Mrays/s; primary launches %.3f ms" % (
            c.trace_ms, c.trace_launches, bounce_ms, (c.rays - c.primary_rays) / max(bounce_ms, 1e-9) / 1e3, c.primary_trace_ms)
    print(dict(zip(keys, combo)), "%.0f Mrays/s (best of %d, %.1f ms)%s" % (best, a.reps, c.kernel_ms, extra), flush=True)
