#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path: Mrays/s (primary + bounce) of the path-tracing kernel.

Workload at N=1 (BASELINE.json configs[2], "C3"): seeded synthetic 1M-triangle lit-box scene (terrain + displaced
spheres, reference-layout BVH built by the reference's median-split algorithm), 1920x1080, 64 spp, depth 8, full
shader (Lambertian + light sampling), Philox RNG, f32 accumulation, fast traversal.  One "step" = one 64-spp render
of the whole frame.  At N>1 the scene is replicated, rank r renders sample slice [r*64, (r+1)*64) of an N*64-spp
image (weak scaling) and the f32 accumulation buffers are summed onto rank 0 with an NCCL reduce inside the timed
region.

  value     whole-job Mrays/s with the scene resident in HBM; timed per step with CUDA events on the launch stream
            (L2 flushed between steps), max over ranks.
  e2e       the same metric through the public API with HOST buffers, following the reference's frame loop: per step the
            32-byte UBO goes host -> device, the frame is rendered and resolved, and the rgba8 target is read back into
            pinned host memory (scene resident, as after the reference's initScene); `with_scene_upload` also re-uploads the
            five scene buffers and rebuilds the traversal records every step.  Wall clock between synchronised barriers,
            max over ranks.
  roofline  HBM-bound traversal roofline for the dominant kernel (wf_trace_kernel, timed per launch with CUDA events on
            its stream inside the timed region): algorithmic bytes per ray B_ray = 48*(nodes + triangles) of the canonical
            (reference-order, t-culled) traversal, counted by the CPU oracle on a tile sample of the same ray set
            (SURVEY.md 8d).  The kernel walks a SAH tree rebuilt over the same leaves and fetches far fewer bytes, so
            `frac` exceeds 1; `own_*` gives the bytes the kernel actually requests (its own node/triangle counters) and
            `l1_gather` the kernel's 32-byte gather rate against the L1TEX gather ceiling measured by tools/ubench.
  cpu_baseline / --impl reference
            the reference's own shader text compiled for the CPU (oracle/_ref, kind "reference"; falls back to the
            restated oracle, kind "port") on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CAM = (1.8, 8.6, 1.1)   # main.cpp:37
HBM_FALLBACK_GBS = 6650.0
L1_GATHER_PEAK_G = 270.0   # measured: profiles/r01_v10_gather_tex_ubench.log (modes L / LL, 32 MB and 2 MB record sets)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.rows, self.proc = device, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, reasons, smax = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        loaded = [x for x in sm if smax and x > 0.5 * smax] or sm
        return {"sm_mhz": float(np.median(loaded)) if loaded else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def pinned_copy(arr):
    import torch
    t = torch.empty(max(arr.nbytes, 1), dtype=torch.uint8, pin_memory=True)
    out = t.numpy()[: arr.nbytes]
    out[:] = arr
    return out, t


def build_workload(args):
    t0 = time.time()
    if args.config == "c2":   # BASELINE.json configs[1]: the bundled scene (glass/metal variant, as assembled by csrc/scene from the reference's OBJ files)
        import vulkan_compute_ray_tracing_b200 as vcrt
        scene = vcrt.load_scene(os.path.join(ROOT, "tests", "golden", "doge_glass_scene.vcrt"))
    else:
        from vulkan_compute_ray_tracing_b200 import scenegen
        scene = scenegen.generate_box_scene(args.triangles, seed=args.scene_seed)
    return scene, time.time() - t0


def oracle_bray(scene, w, h, bounces, tile_count):
    """Algorithmic bytes per ray: canonical (reference order + t-culling) traversal counted by the oracle on 1/tile_count
    of the 32x32 tiles at 1 spp (pcg_ref), SURVEY.md 8d."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracleharness import Oracle, make_params
    o = Oracle()
    p = make_params(shader="full", max_bounces=bounces, stack_depth=64, sample_count=1, accum="f32", tile_rank=0, tile_count=tile_count, flags=4)
    c = o.render(scene, CAM, w, h, p)["counters"]
    return 48.0 * (c.canon_nodes + c.canon_triangles) / max(c.rays, 1), dict(rays=c.rays, nodes_per_ray=c.canon_nodes / max(c.rays, 1),
                                                                             tris_per_ray=c.canon_triangles / max(c.rays, 1),
                                                                             ref_nodes_per_ray=c.ref_nodes / max(c.rays, 1))


def cpu_reference_step(scene, w, h, bounces, frames, first_sample=0):
    """One bounded sample of the workload on the host cores: `frames` 1-spp frames of the full image.
    Returns (seconds, rays, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracleharness import Oracle, make_params
    from refharness import Ref, have_ref
    o = Oracle()
    variant = {2: "full_b2_s16", 4: "full_b4_s16", 8: "full_b8_s32"}.get(bounces)
    if have_ref() and variant:
        import refharness as rh
        r = Ref()
        target = np.zeros((h, w, 4), np.uint8)
        accum = np.zeros((h, w, 4), np.uint8)
        t0 = time.perf_counter()
        for s in range(first_sample, first_sample + frames):
            r.dispatch(variant, scene, rh.pack_ubo(CAM, s, scene), target, accum, (w + 31) // 32, (h + 31) // 32)
            accum[...] = target
        dt = time.perf_counter() - t0
        kind = "reference"
        # the reference shader keeps no counters: count the rays of the identical frames with the restated oracle (untimed)
        stack = 32 if variant.endswith("s32") else 16
        rays = o.render(scene, CAM, w, h, make_params(shader="full", max_bounces=bounces, stack_depth=stack, sample_begin=first_sample,
                                                     sample_count=frames))["counters"].rays
    else:
        p = make_params(shader="full", max_bounces=bounces, stack_depth=64, sample_begin=first_sample, sample_count=frames)
        t0 = time.perf_counter()
        res = o.render(scene, CAM, w, h, p)
        dt = time.perf_counter() - t0
        rays = res["counters"].rays
        kind = "port"
    return dt, rays, kind


def run_reference_arm(args, rank):
    if rank != 0:
        return
    subprocess.call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    scene, _ = build_workload(args)
    frames = args.ref_frames
    for _ in range(args.warmup):
        cpu_reference_step(scene, args.width, args.height, args.bounces, 1)
    tot_t, tot_r, kind = 0.0, 0, "port"
    for k in range(args.steps):
        dt, rays, kind = cpu_reference_step(scene, args.width, args.height, args.bounces, frames, first_sample=k * frames)
        tot_t += dt; tot_r += rays
    v = tot_r / tot_t / 1e6
    cores = os.cpu_count()
    sample = "%d x 1-spp frames of the %dx%d image per step (of %d spp), reference traversal, pcg_ref RNG" % (frames, args.width, args.height, args.spp)
    line = {"impl": "reference", "metric": "Mrays/s (primary+bounce)", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, scene),
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# BASELINE.json configs[2..4].  c3 is the headline (weak scaling: 64 spp per GPU); c4 and c5 are fixed-size jobs (strong scaling).
CONFIGS = {
    "c2": dict(triangles=0, width=1920, height=1080, spp=16, sharding="samples", scaling="weak"),
    "c3": dict(triangles=1000000, width=1920, height=1080, spp=64, sharding="samples", scaling="weak"),
    "c4": dict(triangles=10000000, width=3840, height=2160, spp=16, sharding="tiles", scaling="strong"),
    "c5": dict(triangles=1000000, width=3840, height=2160, spp=1024, sharding="samples", scaling="strong"),
}


_HASH = {}


def scene_hash(scene):
    """sha256 over the five reference-layout buffers in binding order (the bytes both arms render)."""
    key = id(scene)
    if key not in _HASH:
        import hashlib
        h = hashlib.sha256()
        for name in ("triangles", "materials", "bvh", "lights", "spheres"):
            h.update(np.ascontiguousarray(scene[name]).tobytes())
        _HASH[key] = h.hexdigest()[:16]
    return _HASH[key]


def workload_config(args, scene):
    spp_txt = "%d spp/GPU" % args.spp if args.scaling == "weak" else "%d spp in total" % args.spp
    shard = {"samples": "sample slices + NCCL reduce of the f32 accumulation buffers", "tiles": "32x32 tile interleave + NCCL all-gather of packed rgba8 tiles"}[args.sharding]
    what = "bundled scene, glass/metal variant (%d triangles)" % (len(scene["triangles"]) // 48) if args.config == "c2" else \
        "synthetic %d-triangle lit box (seed %d)" % (len(scene["triangles"]) // 48, args.scene_seed)
    return {"workload": "%s: %s, %dx%d, %s, depth %d, full shader, philox, f32 accum, fast traversal"
                        % (args.config.upper(), what, args.width, args.height, spp_txt, args.bounces),
            "triangles": len(scene["triangles"]) // 48, "bvh_nodes": len(scene["bvh"]) // 48, "width": args.width, "height": args.height,
            "spp": args.spp, "spp_is": "per GPU" if args.scaling == "weak" else "total", "max_bounces": args.bounces,
            "scene_sha256": scene_hash(scene),
            "sharding": (shard + " (%s)" % args.scaling) if args.gpus > 1 else "none",
            "l2": "flushed between timed steps (256 MiB memset)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS), help="BASELINE.json workload: c3 (headline), c2, c4, c5")
    ap.add_argument("--triangles", type=int, default=None)
    ap.add_argument("--scene-seed", type=int, default=1234)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--spp", type=int, default=None)
    ap.add_argument("--sharding", default=None, choices=["samples", "tiles"])
    ap.add_argument("--bounces", type=int, default=8)
    ap.add_argument("--traversal", default="fast")
    ap.add_argument("--ref-frames", type=int, default=1, help="1-spp frames per step of the CPU reference arm")
    ap.add_argument("--static-kernel", action="store_true", help="A/B: one-thread-per-pixel launch of the fast traversal")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    for k, v in CONFIGS[args.config].items():
        if getattr(args, k, None) is None:
            setattr(args, k, v)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # host-side scene generation / record builds use OpenMP; torchrun exports OMP_NUM_THREADS=1, which would serialise them
    host_threads = max(1, (os.cpu_count() or 1) // max(world, 1))
    if args.impl == "ours":
        os.environ["OMP_NUM_THREADS"] = str(host_threads)
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    import vulkan_compute_ray_tracing_b200 as vcrt
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene, gen_s = build_workload(args)
    w, h = args.width, args.height

    # ---- host object, as main.cpp:76-153 builds it
    pinned = {k: pinned_copy(v) for k, v in scene.items()}
    ubo = vcrt.BufferUtils.createBundle(vcrt.BufferBundle(1), vcrt.pack_ubo(CAM, 0, scene))
    target, accum = vcrt.Image(w, h), vcrt.Image(w, h)
    mat = vcrt.ComputeMaterial("resources/shaders/generated/ray-trace-compute.spv", device=local_rank)
    mat.addUniformBufferBundle(ubo)
    mat.addStorageImage(target)
    mat.addStorageImage(accum)
    order = ("triangles", "materials", "bvh", "lights", "spheres")
    for name in order:
        mat.addStorageBufferBundle(vcrt.BufferUtils.createBundle(vcrt.BufferBundle(1), scene[name]))
    model = vcrt.ComputeModel(mat)
    mat.setOption("host_threads", str(host_threads))
    stream = torch.cuda.Stream()          # a real (non-default) stream shared by the kernels, NCCL and the timing events
    torch.cuda.set_stream(stream)
    mat.setStream(stream.cuda_stream)

    from vulkan_compute_ray_tracing_b200 import sharding
    # weak scaling: the job is an (spp x world)-sample frame; strong scaling: an spp-sample frame whatever the world size
    total_spp = args.spp * world if args.scaling == "weak" else args.spp
    base = vcrt.render_params(shader="full", traversal=args.traversal, rng="philox", accum="f32", trig="libm", max_bounces=args.bounces,
                              stack_depth=64, sample_begin=0, sample_count=total_spp, philox_seed=args.scene_seed,
                              flags=vcrt.FLAG_STATIC_KERNEL if args.static_kernel else 0)
    params, active = sharding.shard_params(base, args.sharding, rank, world)
    ptr, nbytes = mat.devicePtr(2)

    class _Wrap:
        __cuda_array_interface__ = {"shape": (h, w, 4), "typestr": "<f4", "data": (ptr, False), "version": 2}
    accum_t = torch.as_tensor(_Wrap(), device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        mat.clearAccum()
        if active:
            model.renderCommand(None, 0, params)
        if args.sharding == "samples":
            if world > 1:
                sharding.reduce_accumulation(accum_t, dst=0)      # NCCL SUM reduce on the render stream
            if rank == 0:
                mat.resolve(total_spp, 0.0)
        else:
            mat.resolve(total_spp, 0.0)                           # every rank resolves its own tiles ...
            sharding.gather_tiles_device(mat, 0, rank, world)     # ... and the packed rgba8 tiles are all-gathered

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    mat.resetCounters()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = []
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    c = mat.counters()
    rays, kernel_ms, launches = int(c.rays), float(c.kernel_ms), int(c.launches)
    trace_ms, trace_launches = float(c.trace_ms), int(c.trace_launches)
    tot = torch.tensor([float(rays), float(launches)], dtype=torch.float64, device="cuda")
    mx = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    total_rays, total_launches = float(tot[0]), int(tot[1])
    max_ms = float(mx[0])
    value = total_rays / (max_ms * 1e-3) / 1e6

    # ---- e2e: through the public API with host buffers, host<->device copies inside the timed region.
    # Headline protocol = the reference's own frame loop (main.cpp:166-183, :228, :323-395), which is also what the reference
    # arm times: scene prepared once outside the timed region; per step the 32-byte UBO goes host -> device, the frame is
    # rendered and resolved, and the rgba8 target comes back into pinned host memory.  `with_scene_upload` additionally
    # re-uploads the five scene buffers from pinned host memory and rebuilds the traversal records on the host every step.
    e2e = None
    if not args.no_e2e:
        out_host = torch.empty((h, w, 4), dtype=torch.uint8, pin_memory=True).numpy()
        L = __import__("vulkan_compute_ray_tracing_b200._native", fromlist=["lib"]).lib()
        h2d = sum(v[0].nbytes for v in pinned.values()) + 32

        def frame_step(upload):
            if upload:
                for i, name in enumerate(order):
                    a = pinned[name][0]
                    mat._check(L.vcrt_set_buffer(mat._ctx, 3 + i, a.ctypes.data if a.nbytes else None, a.nbytes))
            ubo.buffers[0].write(vcrt.pack_ubo(CAM, 0, scene))
            step()
            if rank == 0:
                mat._check(L.vcrt_read_target_rgba8(mat._ctx, out_host.ctypes.data, out_host.nbytes))
            else:
                mat.synchronize()

        def timed(upload, steps):
            frame_step(upload)
            barrier()
            mat.resetCounters()
            t0 = time.perf_counter()
            for _ in range(steps):
                frame_step(upload)
            barrier()
            dt = time.perf_counter() - t0
            cc = mat.counters()
            tr = torch.tensor([float(cc.rays)], dtype=torch.float64, device="cuda")
            tm = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tr, op=dist.ReduceOp.SUM)
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            return float(tr[0]) / float(tm[0]) / 1e6, 1e3 * float(tm[0]) / steps

        v, ms = timed(False, args.steps)
        e2e = {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 32, "d2h_bytes_per_step": int(out_host.nbytes), "ms_per_step": ms,
               "protocol": "reference frame loop: UBO write + computeCommand-equivalent render + resolve + rgba8 read-back to pinned host memory; scene resident"}
        # the reference's own unit of work, "ms/frame" (main.cpp:397-413): one 1-spp frame per iteration of the frame loop --
        # UBO with the frame's sample index in, one sample rendered on top of the accumulation, resolve, rgba8 frame back
        one = vcrt.render_params(shader="full", traversal=args.traversal, rng="philox", accum="f32", trig="libm", max_bounces=args.bounces,
                                 stack_depth=64, sample_begin=0, sample_count=1, philox_seed=args.scene_seed)
        if world == 1:
            def one_frame(k):
                ubo.buffers[0].write(vcrt.pack_ubo(CAM, k, scene))
                one.sample_begin = k
                model.renderCommand(None, 0, one)
                mat.resolve(k + 1, 0.0)
                mat._check(L.vcrt_read_target_rgba8(mat._ctx, out_host.ctypes.data, out_host.nbytes))
            mat.clearAccum()
            for k in range(4):
                one_frame(k)
            mat.resetCounters()
            nfr = 32
            t0 = time.perf_counter()
            for k in range(4, 4 + nfr):
                one_frame(k)
            dt = time.perf_counter() - t0
            e2e["frame_1spp"] = {"ms_per_frame": 1e3 * dt / nfr, "value": mat.counters().rays / dt / 1e6, "unit": "Mrays/s", "frames": nfr,
                                 "protocol": "progressive frame loop, one sample per frame, rgba8 frame read back every frame"}
        v, ms = timed(True, min(args.steps, 3))
        e2e["with_scene_upload"] = {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(out_host.nbytes), "ms_per_step": ms,
                                    "includes": "upload of the 5 scene buffers from pinned host memory + host rebuild of the traversal records, every step"}

    if rank == 0:
        peak, peak_src = measured_peak()
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        bray, bray_info = oracle_bray(scene, w, h, args.bounces, tile_count=32)
        # dominant kernel = wf_trace_kernel: every ray passes through exactly one of its launches
        if trace_launches:
            rays_per_launch = rays / trace_launches
            kernel_s = trace_ms * 1e-3 / trace_launches   # average launch duration, CUDA events on the launching stream, timed region
            kname = "wf_trace_kernel"
        else:                                              # A/B variants without a separate trace kernel: the whole render
            rays_per_launch = rays / max(args.steps, 1)
            kernel_s = kernel_ms * 1e-3 / max(args.steps, 1)
            kname = "render kernel"
        achieved = bray * rays_per_launch / kernel_s / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        # what the kernel itself requests: its node/triangle counters on one 1-spp pass (untimed, counting build of the kernel)
        cp = vcrt.render_params(shader="full", traversal=args.traversal, rng="philox", accum="f32", trig="libm", max_bounces=args.bounces,
                                stack_depth=64, sample_begin=0, sample_count=1, philox_seed=args.scene_seed, flags=vcrt.FLAG_COUNT_TRAVERSAL)
        mat.clearAccum(); mat.resetCounters()
        model.renderCommand(None, 0, cp)
        cc = mat.counters()
        node_bytes = 32 if mat.getInfo("fast_nodes") == "q15" else 64
        tri_bytes = 64                                  # the traversal's triangle record (two 256-bit loads per test)
        own_bray = (node_bytes * cc.nodes + tri_bytes * cc.triangles) / max(cc.rays, 1)
        # the memory-side ceiling that actually binds the kernel: 32-byte gathers through L1TEX;
        # peak = dependent random 32-byte gathers missing L1, tools/ubench/gather_tex.cu on this GPU model (profiles/r01_v10_gather_tex_ubench.log)
        gathers_per_ray = (node_bytes // 32) * cc.nodes / max(cc.rays, 1) + 2.0 * cc.triangles / max(cc.rays, 1) + 1.0
        g_rate = gathers_per_ray * rays_per_launch / kernel_s / 1e9
        own = {"own_bytes_per_ray": own_bray, "own_nodes_per_ray": cc.nodes / max(cc.rays, 1), "own_tris_per_ray": cc.triangles / max(cc.rays, 1),
               "own_node_bytes": node_bytes, "own_tri_bytes": tri_bytes, "own_achieved": own_bray * rays_per_launch / kernel_s / 1e9,
               "own_frac": own_bray * rays_per_launch / kernel_s / 1e9 / peak,
               "l1_gather": {"gathers_per_ray": gathers_per_ray, "achieved": g_rate, "peak": L1_GATHER_PEAK_G, "unit": "G 32-byte gathers/s",
                             "frac": g_rate / L1_GATHER_PEAK_G, "peak_source": "tools/ubench/gather_tex.cu, L1-missing chains, B200"}}
        line = {"metric": "Mrays/s (primary+bounce)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, scene), "clocks": clocks, "gpu_launches": total_launches,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                             "peak_source": peak_src, "kernel": kname, "bytes_per_ray": bray, "roofline_mrays": peak * 1e3 / bray,
                             "kernel_ms_per_launch": kernel_s * 1e3, "rays_per_launch": rays_per_launch, "launches_per_step": trace_launches / max(args.steps, 1),
                             "kernel_share_of_step": (trace_ms / max(args.steps, 1)) / (max_ms / args.steps) if trace_launches else 1.0,
                             "kernel_mrays": rays_per_launch / kernel_s / 1e6, "canonical_traversal": bray_info, **own},
                "rays_per_step": total_rays / args.steps, "scene_build_s": gen_s}
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline and args.config in ("c2", "c3"):
            dt1, rays1, kind = cpu_reference_step(scene, w, h, args.bounces, 1)
            frames = int(min(max(round(15.0 / max(dt1, 1e-3)), 1), 16))
            dt, r, kind = cpu_reference_step(scene, w, h, args.bounces, frames, first_sample=1)
            line["cpu_baseline"] = {"value": r / dt / 1e6, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": kind,
                                    "sample": "%d x 1-spp frames of the %dx%d image (of %d spp), reference traversal, pcg_ref RNG, %.1f s" % (frames, w, h, args.spp, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries exactly one line, the JSON result: anything libraries print on the way (e.g. NCCL's version banner)
    # goes to stderr instead
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _print = print

    def print(*a, **k):   # noqa: A001 -- the result line goes to the real stdout
        sys.stdout.flush()
        os.write(_real_stdout, (" ".join(str(x) for x in a) + "\n").encode())

    main()
