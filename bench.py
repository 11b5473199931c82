#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path: Mrays/s (primary + bounce) of the path-tracing kernel.

Workload at N=1 (BASELINE.json configs[2], "C3"): seeded synthetic 1M-triangle lit-box scene (terrain + displaced
spheres, reference-layout BVH built by the reference's median-split algorithm), 1920x1080, 64 spp, depth 8, full
shader (Lambertian + light sampling), Philox RNG, f32 accumulation, fast traversal.  One "step" = one 64-spp render
of the whole frame.  At N>1 the scene is replicated, rank r renders sample slice [r*64, (r+1)*64) of an N*64-spp
image (weak scaling) and the f32 accumulation buffers are summed onto rank 0 with ONE NCCL reduce issued by the library
(vcrt_group_render) on the render stream, inside the timed region.

  value     whole-job Mrays/s with the scene resident in HBM; timed per step with CUDA events on the launch stream
            (L2 flushed between steps), max over ranks.  A "ray" is a closest-hit query of the shader's ray_color
            (ray-trace-compute.comp:321-323), one per sample and bounce -- the same count the reference arm reports.  The
            shader's primary ray does not depend on the sample (:352-373, no jitter), so the wavefront pipeline walks the tree
            once per pixel for bounce 0 and shares the answer among the pixel's samples: `primary_mrays` / `bounce_mrays`
            give the two populations separately, `traversals_per_step` what was actually walked.
  e2e       the same metric through the public API with HOST buffers, following the reference's frame loop: per step the
            32-byte UBO goes host -> device, the frame is rendered and resolved, and the rgba8 target is read back into
            pinned host memory (scene resident, as after the reference's initScene); `with_scene_upload` also re-uploads the
            five scene buffers and rebuilds the traversal records every step.  Wall clock between synchronised barriers,
            max over ranks.
  roofline  for the dominant kernel (wf_trace_kernel, timed per launch with CUDA events on its stream).  When the timed region
            runs as one pipeline those are the launches of the timed region; when it runs as two (wf_streams "auto", large
            renders: one pipeline's shade launches overlap the other's trace launches) a launch's events also span the time it
            shares the SMs, so the kernel is timed alone in `steps` further steps of the same workload with one pipeline, same
            process, right after the timed region (`roofline.timed_as`, `roofline.timed_region`).  C3's working set (21 MB of nodes + 64 MB of triangle records) is L2-resident on a B200, so HBM does not
            bind it; what does is the rate at which L1TEX pulls divergent 32-byte sectors out of L2.  `peak` is that rate
            MEASURED IN THIS PROCESS before the timed region (tools/ubench/probe.cu: dependent random 32-byte LDG.256 gathers
            over a 32 MB L2-resident table, the kernel's own launch shape), `achieved` the kernel's own scene-record sector
            requests (its node/triangle counters x 2 sectors) per second of kernel time, both x 32 B -> GB/s.
            `roofline.canonical` keeps SURVEY 8d's figure for the record (48 B x the canonical reference-order traversal's
            fetches over the measured HBM peak; the kernel walks a 4-wide SAH tree and fetches ~14x fewer bytes, so that
            fraction is not a ceiling).  `roofline.hbm` = real DRAM bytes per launch (ncu capture committed under profiles/)
            over kernel time and the measured HBM peak.  `c4` (N=1 only) runs BASELINE configs[3]'s 10M-triangle scene, whose
            0.8 GB of records do not fit L2, with the same accounting against a DRAM-resident gather probe.
  cpu_baseline / --impl reference
            the reference's own shader text compiled for the CPU (oracle/_ref, kind "reference"; falls back to the
            restated oracle, kind "port") on all host cores, on a bounded sample of the same workload.
"""
import argparse
import ctypes
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


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


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


def build_workload(args, config=None, triangles=None):
    t0 = time.time()
    config = config or args.config
    if config == "c2":   # BASELINE.json configs[1]: the bundled scene (glass/metal variant, as assembled by csrc/scene from the reference's OBJ files)
        import vulkan_compute_ray_tracing_b200 as vcrt
        scene = vcrt.load_scene(os.path.join(ROOT, "tests", "golden", "doge_glass_scene.vcrt"))
    else:
        from vulkan_compute_ray_tracing_b200 import scenegen
        scene = scenegen.generate_box_scene(triangles or args.triangles, seed=args.scene_seed)
    return scene, time.time() - t0


def set_host_threads(n):
    """OpenMP threads of everything host-side in this process (launchers such as torchrun export OMP_NUM_THREADS=1, which would
    serialise the scene generator, the record build and -- 16x slower -- the CPU reference arm)."""
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


def oracle_bray(scene, w, h, bounces, tile_count):
    """Algorithmic bytes per ray of SURVEY.md 8d: canonical (reference order + t-culling) traversal counted by the oracle on
    1/tile_count of the 32x32 tiles at 1 spp (pcg_ref)."""
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
    """The reference's own CPU implementation of the path (oracle/_ref), all host threads, rank 0 only.  Bounded: every step is
    `--ref-frames` 1-spp frames of the workload's image (a 1/spp sample of the step), and the whole run is cut short once
    `--ref-budget` seconds of CPU rendering have been spent, so that it ends within a few minutes on any core count."""
    if rank != 0:
        return
    set_host_threads(os.cpu_count() or 1)
    subprocess.call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    scene, _ = build_workload(args)
    frames = args.ref_frames
    spent = 0.0
    for _ in range(min(args.warmup, 1)):          # one warm-up frame pages the scene in; more would only burn the budget
        dt, _, _ = cpu_reference_step(scene, args.width, args.height, args.bounces, 1)
        spent += dt
    tot_t, tot_r, kind, done = 0.0, 0, "port", 0
    for k in range(args.steps):
        dt, rays, kind = cpu_reference_step(scene, args.width, args.height, args.bounces, frames, first_sample=k * frames)
        tot_t += dt; tot_r += rays; done += 1
        if spent + tot_t > args.ref_budget:
            break
    v = tot_r / tot_t / 1e6
    cores = os.cpu_count()
    sample = "%d step(s) x %d 1-spp frame(s) of the %dx%d image (the workload's step is %d spp), %.1f s of CPU time" % (done, frames, args.width, args.height, args.spp, tot_t)
    cfg = workload_config(args, scene)
    cfg["implementation"] = ("the reference's shader text compiled for the host (oracle/_ref, OpenMP over rows, %d threads): PCG RNG of random.glsl, literal hit_bvh "
                             "on the bound median-split tree with a 32-entry stack, rgba8 running mean; %s" % (cores, sample)) if kind == "reference" else \
        ("restated oracle (oracle/vcrt_oracle.c), %d threads; %s" % (cores, sample))
    cfg["sharding"] = "none (rank 0 only)"
    line = {"impl": "reference", "metric": "Mrays/s (primary+bounce)", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": done,
            "warmup": min(args.warmup, 1), "steps_requested": args.steps, "warmup_requested": args.warmup, "ms_per_step": 1e3 * tot_t / max(done, 1), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
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
    """What is rendered (identical for both arms); HOW it is rendered goes into config["implementation"] per arm."""
    spp_txt = "%d spp/GPU" % args.spp if args.scaling == "weak" else "%d spp in total" % args.spp
    shard = {"samples": "sample slices + NCCL reduce of the f32 accumulation buffers", "tiles": "32x32 tile interleave + NCCL all-gather of packed rgba8 tiles"}[args.sharding]
    what = "bundled scene, glass/metal variant (%d triangles)" % (len(scene["triangles"]) // 48) if args.config == "c2" else \
        "synthetic %d-triangle lit box (seed %d)" % (len(scene["triangles"]) // 48, args.scene_seed)
    return {"workload": "%s: %s, %dx%d, %s, depth %d, full shader (Lambertian + light sampling)" % (args.config.upper(), what, args.width, args.height, spp_txt, args.bounces),
            "triangles": len(scene["triangles"]) // 48, "bvh_nodes": len(scene["bvh"]) // 48, "width": args.width, "height": args.height,
            "spp": args.spp, "spp_is": "per GPU" if args.scaling == "weak" else "total", "max_bounces": args.bounces,
            "scene_sha256": scene_hash(scene),
            "sharding": (shard + " (%s), vcrt_group_render" % args.scaling) if args.gpus > 1 else "none",
            "l2": "flushed between timed steps (256 MiB memset)"}


class Probe:
    """tools/ubench/libvcrt_probe.so: peaks measured in this process, on this GPU, before the timed region."""

    def __init__(self, device):
        path = os.path.join(ROOT, "tools", "ubench", "libvcrt_probe.so")
        if not os.path.exists(path):
            subprocess.call(["make", "-C", os.path.dirname(path)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        self.lib = ctypes.CDLL(path)
        self.device = device

    def gather(self, records_log2, steps=64, chains=1, reps=3):
        out = ctypes.c_double()
        rc = self.lib.vcrt_probe_gather(self.device, records_log2, steps, chains, reps, ctypes.byref(out))
        return out.value if rc == 0 else None

    def gather64(self, records_log2, steps=64, chains=1, reps=3):
        out = ctypes.c_double()
        rc = self.lib.vcrt_probe_gather64(self.device, records_log2, steps, chains, reps, ctypes.byref(out))
        return out.value if rc == 0 else None

    def gather128(self, records_log2, steps=32, chains=1, reps=3):
        out = ctypes.c_double()
        rc = self.lib.vcrt_probe_gather128(self.device, records_log2, steps, chains, reps, ctypes.byref(out))
        return out.value if rc == 0 else None

    def stream(self, nbytes, passes=8, reps=3):
        out = ctypes.c_double()
        self.lib.vcrt_probe_stream.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        rc = self.lib.vcrt_probe_stream(self.device, nbytes, passes, reps, ctypes.byref(out))
        return out.value if rc == 0 else None


def ncu_traffic(key):
    """DRAM bytes per trace launch from the committed ncu capture of this round (profiles/ncu_traffic.json; `ncu --metrics
    dram__bytes_read.sum,dram__bytes_write.sum` over the bench command).  None when no capture is on record for `key`."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            d = json.load(f)
        return d.get(key)
    except Exception:
        return None


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
    ap.add_argument("--ref-budget", type=float, default=60.0, help="seconds of CPU rendering after which the reference arm stops early")
    ap.add_argument("--static-kernel", action="store_true", help="A/B: one-thread-per-pixel launch of the fast traversal")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the HBM-sized C4 leg of the N=1 run")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sub-record of an N>1 run")
    ap.add_argument("--option", action="append", default=[], help="key=value for vcrt_set_option (A/B runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    for k, v in CONFIGS[args.config].items():
        if getattr(args, k, None) is None:
            setattr(args, k, v)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    # host-side scene generation / record builds use OpenMP; torchrun exports OMP_NUM_THREADS=1, which would serialise them
    host_threads = max(1, (os.cpu_count() or 1) // max(world, 1))
    set_host_threads(host_threads)

    import torch
    import torch.distributed as dist
    import vulkan_compute_ray_tracing_b200 as vcrt
    from vulkan_compute_ray_tracing_b200 import sharding
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- peaks of the levels that can bind the trace kernel, measured here and now (before anything else touches the GPU)
    probes = None
    if rank == 0:
        pr = Probe(local_rank)
        probes = {"l2_gather_gps": pr.gather(20), "l2_gather_2chains_gps": pr.gather(20, chains=2), "l1_gather_gps": pr.gather(11), "l2_gather64_gps": pr.gather64(19),
                  "l2_stream_gbs": pr.stream(48 << 20, passes=16), "hbm_stream_gbs": pr.stream(4 << 30, passes=1),
                  "how": "tools/ubench/probe.cu in this process: dependent random 32-byte LDG.256 gathers, 148 x 10 x 6 blocks of 128 (table 32 MB: L2-resident, "
                         "L1-missing; 64 KB: L1-resident); coalesced 128-bit reads of 48 MB (L2) / 4 GB (HBM); best of 3 after a warm-up"}

    scene, gen_s = build_workload(args)
    w, h = args.width, args.height

    def make_host_object(sc, ww, hh):
        """main.cpp:76-153"""
        ubo_ = vcrt.BufferUtils.createBundle(vcrt.BufferBundle(1), vcrt.pack_ubo(CAM, 0, sc))
        target_, accum_ = vcrt.Image(ww, hh), vcrt.Image(ww, hh)
        mat_ = vcrt.ComputeMaterial("resources/shaders/generated/ray-trace-compute.spv", device=local_rank)
        mat_.addUniformBufferBundle(ubo_)
        mat_.addStorageImage(target_)
        mat_.addStorageImage(accum_)
        for name in order:
            mat_.addStorageBufferBundle(vcrt.BufferUtils.createBundle(vcrt.BufferBundle(1), sc[name]))
        model_ = vcrt.ComputeModel(mat_)
        mat_.setOption("host_threads", str(host_threads))
        for kv in args.option:
            k_, v_ = kv.split("=", 1)
            mat_.setOption(k_, v_)
        return ubo_, target_, mat_, model_

    order = ("triangles", "materials", "bvh", "lights", "spheres")
    pinned = {k: pinned_copy(v) for k, v in scene.items()}
    ubo, target, mat, model = make_host_object(scene, w, h)
    stream = torch.cuda.Stream()          # a real (non-default) stream shared by the kernels, the library's NCCL calls and the timing events
    torch.cuda.set_stream(stream)
    mat.setStream(stream.cuda_stream)
    group = sharding.Group.from_torch(mat) if world > 1 else None

    # weak scaling: the job is an (spp x world)-sample frame; strong scaling: an spp-sample frame whatever the world size
    total_spp = args.spp * world if args.scaling == "weak" else args.spp

    def params_for(spp):
        return vcrt.render_params(shader="full", traversal=args.traversal, rng="philox", accum="f32", trig="libm", max_bounces=args.bounces,
                                  stack_depth=64, sample_begin=0, sample_count=spp, philox_seed=args.scene_seed,
                                  flags=vcrt.FLAG_STATIC_KERNEL if args.static_kernel else 0)
    base = params_for(total_spp)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step(p=base, mode=None):
        if group is None:
            mat.clearAccum()
            model.renderCommand(None, 0, p)
            mat.resolve(p.sample_count, 0.0)
        else:
            group.render(model, p, mode or args.sharding, 0.0)     # clear + this rank's share + ONE NCCL collective + resolve, all on `stream`

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(p, steps, warmup, mode=None):
        for _ in range(warmup):
            step(p, mode)
        barrier()
        mat.resetCounters()
        evs = []
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step(p, mode)
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        wall_ = time.perf_counter() - t0
        dev_ms_ = sum(a.elapsed_time(b) for a, b in evs)
        c_ = mat.counters()
        tot_ = torch.tensor([float(c_.rays), float(c_.launches), float(c_.traversals), float(c_.primary_rays)], dtype=torch.float64, device="cuda")
        mx_ = torch.tensor([dev_ms_, wall_ * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot_, op=dist.ReduceOp.SUM)
            dist.all_reduce(mx_, op=dist.ReduceOp.MAX)
        return c_, [float(x) for x in tot_], float(mx_[0])

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step()
    barrier()
    if rank == 0:
        sampler.start()
    c, tot, max_ms = timed_steps(base, args.steps, 0)
    clocks = sampler.stop() if rank == 0 else None
    rays, kernel_ms, launches = int(c.rays), float(c.kernel_ms), int(c.launches)
    # ---- the dominant kernel timed ALONE.  A large render runs as two pipelines (option wf_streams, "auto"): one pipeline's shade
    # launches overlap the other's trace launches, so inside the timed region the CUDA events around a trace launch also span the
    # time it shares the SMs with other kernels.  The roofline figures come from `steps` more steps of the same workload, same
    # process, L2 flushed likewise, with ONE pipeline -- every trace launch then has the GPU to itself -- right after the timed region.
    pipelines = int(mat.getInfo("wf_pipelines") or 1)
    if world > 1:       # every rank takes the same branch (tile shards can differ by one tile)
        pl_ = torch.tensor([float(pipelines)], dtype=torch.float64, device="cuda")
        dist.all_reduce(pl_, op=dist.ReduceOp.MAX)
        pipelines = int(pl_.item())
    overlapped = None
    c_roof, roof_ms = c, max_ms
    if pipelines > 1:
        user_streams = dict(kv.split("=", 1) for kv in args.option).get("wf_streams", "auto")
        overlapped = {"pipelines": pipelines, "trace_ms_per_step_summed_over_overlapping_launches": float(c.trace_ms) / args.steps,
                      "trace_launches_per_step": int(c.trace_launches) / args.steps}
        mat.setOption("wf_streams", "1")
        c_roof, _, roof_ms = timed_steps(base, args.steps, 1)
        mat.setOption("wf_streams", user_streams)
        step()      # queues back to the two-pipeline layout before anything else is timed
        barrier()
    trace_ms, trace_launches = float(c_roof.trace_ms), int(c_roof.trace_launches)
    total_rays, total_launches, total_trav, total_prim = tot[0], int(tot[1]), tot[2], tot[3]
    value = total_rays / (max_ms * 1e-3) / 1e6

    # ---- correctness of the combined multi-GPU frame (outside the timed region): the N-GPU frame against ONE GPU rendering the
    # same samples.  Tile shards: bit-identical on every rank; sample slices: same sample set, fp32 summation order differs ->
    # the resolved rgba8 frames agree within 1 LSB.
    frame_check = None
    if world > 1:
        step()
        mat.synchronize()
        got = target.read() if (rank == 0 or args.sharding == "tiles") else None
        ok = True
        detail = None
        if got is not None:
            mat.clearAccum()
            model.renderCommand(None, 0, base)
            mat.resolve(total_spp, 0.0)
            want = target.read()
            d = np.abs(got.astype(np.int16) - want.astype(np.int16))
            ok = bool(d.max() == 0) if args.sharding == "tiles" else bool(d.max() <= 1)
            detail = {"max_abs_diff_lsb": int(d.max()), "differing_channels": int((d > 0).sum()), "sha256_combined": __import__("hashlib").sha256(got.tobytes()).hexdigest()[:16]}
        flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        frame_check = {"frame_matches_1gpu": bool(flag.item() == 1.0), "rule": "bit-identical on every rank" if args.sharding == "tiles" else "rgba8 within 1 LSB on rank 0",
                       "samples_compared": total_spp, **(detail or {})}

    # ---- strong scaling (N > 1): the SAME spp-sample frame split over the N GPUs, collective included, and -- the anchor --
    # that frame on rank 0's GPU alone, in the same run
    strong = None
    if world > 1 and args.scaling == "weak" and not args.no_strong:
        ps = params_for(args.spp)
        # both partitions of the frame: sample slices (every GPU all pixels, spp / N samples; f32 sums reduced onto rank 0) and 32x32
        # tiles (every GPU all samples of its tiles -- bounce 0 traced for 1/N of the pixels only; packed rgba8 tiles all-gathered,
        # bit-identical to one GPU); the headline of the record is the faster one
        runs = {}
        for mode_ in ("samples", "tiles"):
            _, tot_m, ms_m = timed_steps(ps, args.steps, 2, mode_)
            runs[mode_] = (tot_m, ms_m)
        best_mode = min(runs, key=lambda k: runs[k][1])
        tot_s, ms_s = runs[best_mode]
        one_ms = 0.0
        if rank == 0:
            def one():
                mat.clearAccum(); model.renderCommand(None, 0, ps); mat.resolve(args.spp, 0.0)
            for _ in range(2):
                one()
            mat.synchronize()
            evs = []
            for _ in range(args.steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); one(); e1.record(stream)
                evs.append((e0, e1))
            torch.cuda.synchronize()
            one_ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
        barrier()
        strong = {"what": "the %d-spp frame of the N=1 workload split over %d GPUs (%s), collective and resolve inside the timed region" % (args.spp, world, best_mode),
                  "sharding": best_mode,
                  "ms_per_step": ms_s / args.steps, "value": tot_s[0] / (ms_s * 1e-3) / 1e6, "unit": "Mrays/s", "spp_total": args.spp, "steps": args.steps,
                  "n1_ms_per_step_same_run": one_ms, "speedup_vs_n1_same_run": one_ms / (ms_s / args.steps) if ms_s > 0 else None,
                  "efficiency_vs_n1_same_run": one_ms / (ms_s / args.steps) / world if ms_s > 0 else None,
                  "by_sharding": {k: {"ms_per_step": v[1] / args.steps, "value": v[0][0] / (v[1] * 1e-3) / 1e6,
                                      "efficiency_vs_n1_same_run": one_ms / (v[1] / args.steps) / world if v[1] > 0 else None} for k, v in runs.items()}}

    # ---- e2e: through the public API with host buffers, host<->device copies inside the timed region.
    # Headline protocol = the reference's own frame loop (main.cpp:166-183, :228, :323-395), which is also what the reference
    # arm times: scene prepared once outside the timed region; per step the 32-byte UBO goes host -> device, the frame is
    # rendered and resolved, and the rgba8 target comes back into pinned host memory.  `with_scene_upload` additionally
    # re-uploads the five scene buffers from pinned host memory and rebuilds the traversal records every step.
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
        one = params_for(1)
        if world == 1:
            def one_frame(k):
                ubo.buffers[0].write(vcrt.pack_ubo(CAM, k, scene))
                one.sample_begin = k
                model.renderCommand(None, 0, one)
                mat.resolve(k + 1, 0.0)
                mat._check(L.vcrt_read_target_rgba8(mat._ctx, out_host.ctypes.data, out_host.nbytes))

            def frame_loop(nfr):
                mat.clearAccum()
                for k in range(4):
                    one_frame(k)
                mat.resetCounters()
                t0 = time.perf_counter()
                for k in range(4, 4 + nfr):
                    one_frame(k)
                dt = time.perf_counter() - t0
                return 1e3 * dt / nfr, mat.counters().rays / dt / 1e6
            nfr = 64
            ms1, v1 = frame_loop(nfr)

            # The same loop as the reference runs it: MAX_FRAMES_IN_FLIGHT = 2 frames going at once (main.cpp:68; fences :298-316,
            # :325, :394) -- vcrt_frames_begin / vcrt_frame_submit / vcrt_frame_wait.  drawFrame waits for the fence of the slot it
            # reuses, writes the UBO, submits and moves on; every frame's rgba8 image still lands in (page-locked) host memory.
            def frame_loop_in_flight(nfl, nfr):
                bufs = [vcrt.PinnedFrame(w, h) for _ in range(nfl)]
                mat.clearAccum()
                mat.framesBegin(nfl)

                def submit(k):
                    ubo.buffers[0].write(vcrt.pack_ubo(CAM, k, scene))
                    one.sample_begin = k
                    mat.frameSubmit(one, total_samples=k + 1, gamma=0.0, out=bufs[k % nfl])
                for k in range(8):
                    submit(k)
                mat.synchronize()
                c0 = mat.counters().rays
                t0 = time.perf_counter()
                for k in range(8, 8 + nfr):
                    submit(k)
                for s_ in range(nfl):
                    mat.frameWait(s_)
                dt = time.perf_counter() - t0
                r_ = mat.counters().rays - c0
                mat.framesEnd()
                for b_ in bufs:
                    b_.free()
                return 1e3 * dt / nfr, r_ / dt / 1e6
            mat.resetCounters()
            ms2, v2 = frame_loop_in_flight(2, nfr)
            ms4, v4 = frame_loop_in_flight(4, nfr)
            e2e["frame_1spp"] = {"ms_per_frame": ms2, "value": v2, "unit": "Mrays/s", "frames": nfr, "frames_in_flight": 2,
                                 "protocol": "progressive frame loop with the reference's MAX_FRAMES_IN_FLIGHT = 2 (main.cpp:68): per frame UBO write, one sample rendered, folded into "
                                             "the accumulation in frame order, resolved, rgba8 frame read back into page-locked host memory; frames bit-identical to the synchronous loop",
                                 "synchronous": {"ms_per_frame": ms1, "value": v1, "what": "one frame at a time: render, resolve, blocking read-back (the latency of a single frame)"},
                                 "frames_in_flight_4": {"ms_per_frame": ms4, "value": v4}}
        # a scene change every step: the records are rebuilt on the device (fast_build=device: CUDA kernels over the uploaded buffers);
        # `host_build` = the same with the host builder (binned SAH on the CPU cores + upload of its records)
        mat.setOption("fast_build", "device")
        v, ms = timed(True, min(args.steps, 3))
        e2e["with_scene_upload"] = {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(out_host.nbytes), "ms_per_step": ms,
                                    "record_build": mat.getInfo("fast_build"), "record_build_ms": float(mat.getInfo("fast_build_ms")), "record_build_stages": mat.getInfo("fast_build_stages"),
                                    "includes": "upload of the 5 scene buffers from pinned host memory + rebuild of the traversal records on the device, every step"}
        mat.setOption("fast_build", "host")
        v, ms = timed(True, min(args.steps, 2))
        e2e["with_scene_upload"]["host_build"] = {"value": v, "ms_per_step": ms, "record_build_ms": float(mat.getInfo("fast_build_ms"))}
        mat.setOption("fast_build", "auto")

    def kernel_accounting(mat_, model_, spp_, c_, steps_, step_ms, probes_, cfg_key, hbm_peak):
        """Roofline records of the dominant kernel from live counters: c_ = counters of the timed region (steps of spp_ samples)."""
        rays_ = int(c_.rays)
        trav_ = int(c_.traversals)
        t_ms, t_n, p_ms = float(c_.trace_ms), int(c_.trace_launches), float(c_.primary_trace_ms)
        if not t_n:
            return None
        # what the kernel itself requests: its node/triangle counters over one more step of the same shape (untimed, counting
        # build of the kernel; same mix of primary and bounce traversals as the timed steps)
        cp = params_for(spp_)
        cp.flags |= vcrt.FLAG_COUNT_TRAVERSAL
        mat_.clearAccum(); mat_.resetCounters()
        model_.renderCommand(None, 0, cp)
        cc = mat_.counters()
        node_bytes = 32 if mat_.getInfo("fast_nodes") == "q15" else 64
        tri_bytes = 64
        sectors_per_trav = ((node_bytes // 32) * cc.nodes + 2.0 * cc.triangles) / max(cc.traversals, 1)
        kernel_s = t_ms * 1e-3 / t_n
        trav_per_launch = trav_ / t_n
        g_rate = sectors_per_trav * trav_ / (t_ms * 1e-3) / 1e9          # G sectors/s over all trace launches of the timed region
        peak_g = probes_["l2_gather_gps"]
        rec = {"bound": "l2", "achieved": g_rate * 32.0, "peak": peak_g * 32.0, "unit": "GB/s", "frac": g_rate / peak_g,
               "what": "scene-record sector requests of the kernel (2 x 32 B per node visit and per triangle test, own counters) x 32 B per second of kernel time, against the "
                       "rate of dependent random 32-byte gathers from an L2-resident table measured by the in-process probe",
               "peak_source": "in-process probe (tools/ubench/probe.cu), this GPU, this run", "kernel": "wf_trace_kernel",
               "gather_rate_gps": g_rate, "gather_peak_gps": peak_g, "gather_peak_l1_hit_gps": probes_["l1_gather_gps"], "frac_vs_l1_hit_peak": g_rate / probes_["l1_gather_gps"],
               "sectors_per_traversal": sectors_per_trav, "nodes_per_traversal": cc.nodes / max(cc.traversals, 1), "tris_per_traversal": cc.triangles / max(cc.traversals, 1),
               "node_bytes": node_bytes, "tri_bytes": tri_bytes, "kernel_ms_per_launch": kernel_s * 1e3, "traversals_per_launch": trav_per_launch,
               "launches_per_step": t_n / max(steps_, 1), "kernel_share_of_step": (t_ms / max(steps_, 1)) / step_ms,
               "kernel_mrays_traversed": trav_ / (t_ms * 1e-3) / 1e6,
               "bounce_launches": {"ms_per_step": (t_ms - p_ms) / max(steps_, 1), "mrays": (rays_ - int(c_.primary_rays)) / max((t_ms - p_ms) * 1e-3, 1e-12) / 1e6},
               "primary_launches": {"ms_per_step": p_ms / max(steps_, 1), "traversals_per_step": (trav_ - (rays_ - int(c_.primary_rays))) / max(steps_, 1),
                                    "queries_answered_per_step": int(c_.primary_rays) / max(steps_, 1)}}
        tr = ncu_traffic(cfg_key)
        rec["traffic"] = tr.get("dram_bytes_per_launch") if tr else None
        if tr and tr.get("dram_bytes_per_launch"):
            per_launch = float(tr["dram_bytes_per_launch"])
            ach = per_launch / kernel_s / 1e9
            rec["hbm"] = {"traffic_bytes_per_launch": per_launch, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "source": tr.get("source"),
                          "dram_bytes_per_traversal": per_launch / trav_per_launch, "record_bytes_requested_per_traversal": 32.0 * sectors_per_trav}
        return rec

    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        steps = args.steps
        step_ms = max_ms / steps
        # the committed ncu DRAM capture applies to the headline shape only (C3 at its default size and sample count, per GPU)
        traffic_key = "c3" if (args.config == "c3" and (w, h, args.spp, args.triangles) == (1920, 1080, 64, 1000000) and args.scaling == "weak") else "none"
        if pipelines > 1:
            mat.setOption("wf_streams", "1")      # the counting step of the accounting: same launch shape as c_roof
        roof = kernel_accounting(mat, model, base.sample_count if world == 1 else max(base.sample_count // world, 1), c_roof, steps, roof_ms / steps, probes, traffic_key, peak)
        if pipelines > 1:
            mat.setOption("wf_streams", user_streams)
            if roof is not None:
                roof["timed_as"] = ("the kernel alone: %d steps of the same workload with one pipeline (wf_streams=1, %.2f ms per step), in this process right after the "
                                    "timed region; the timed region itself runs %d pipelines whose trace and shade launches overlap" % (steps, roof_ms / steps, pipelines))
                roof["timed_region"] = overlapped
        if roof is None:      # A/B variants without a separate trace kernel
            roof = {"bound": "l2", "achieved": None, "peak": probes["l2_gather_gps"] * 32.0, "unit": "GB/s", "frac": None, "traffic": None, "kernel": "render kernel (one launch)"}
        bray, bray_info = oracle_bray(scene, w, h, args.bounces, tile_count=32)
        if trace_launches:
            ach_c = bray * (int(c_roof.rays) / trace_launches) / (trace_ms * 1e-3 / trace_launches) / 1e9
            roof["canonical"] = {"bytes_per_ray": bray, "achieved": ach_c, "peak": peak, "unit": "GB/s", "frac": ach_c / peak, "peak_source": peak_src,
                                 "note": "SURVEY 8d's figure, kept for the record: 48 B x fetches of the reference-order t-culled traversal of the bound median-split tree, per "
                                         "closest-hit query; the kernel walks its own 4-wide SAH tree (and bounce 0 once per pixel), so this is not a ceiling", **bray_info}
        roof["probes"] = probes
        prim_q, bounce_q = total_prim, total_rays - total_prim
        cfg = workload_config(args, scene)
        cfg["implementation"] = "Philox RNG, f32 accumulation, fast traversal (4-wide quantised SAH tree rebuilt over the bound leaves), wavefront pipeline; bounce 0 traced once per pixel"
        line = {"metric": "Mrays/s (primary+bounce)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg, "clocks": clocks, "gpu_launches": total_launches,
                "roofline": roof, "rays_per_step": total_rays / steps, "traversals_per_step": total_trav / steps,
                "primary_queries_per_step": prim_q / steps, "bounce_rays_per_step": bounce_q / steps,
                "bounce_mrays": roof.get("bounce_launches", {}).get("mrays"), "primary_mrays": (prim_q / steps / world) / max(roof.get("primary_launches", {}).get("ms_per_step", 0.0) * 1e-3, 1e-12) / 1e6
                if roof.get("primary_launches") else None,
                "scene_build_s": gen_s}
        if frame_check:
            line.update(frame_check)
        if strong:
            line["strong"] = strong
        if e2e:
            line["e2e"] = e2e

    # ---- C4 on one GPU (N=1, default config only): the configuration whose traversal records (0.8 GB) do NOT fit L2, i.e. the
    # one where HBM can bind.  Same kernel, same accounting, at 1080p / 8 spp (the trace kernel's behaviour per ray does not depend
    # on the image size; BASELINE's 4K 16-spp frame of this scene is the 8-GPU tile-sharded job).
    if world == 1 and rank == 0 and args.config == "c3" and not args.no_c4 and not args.static_kernel:
        try:
            mat.destroy()
            del flush
            torch.cuda.empty_cache()
            sc4, gen4 = build_workload(args, "c3", triangles=10000000)
            ubo4, target4, mat4, model4 = make_host_object(sc4, 1920, 1080)
            mat4.setStream(stream.cuda_stream)
            p4 = params_for(8)
            pr4 = dict(probes)
            pr = Probe(local_rank)
            dram_gather = pr.gather(26, steps=32)        # 2 GB table of 32-byte records: DRAM-resident
            # 2 GB tables of 64-byte and of 128-byte records (2 / 4 adjacent sectors per gather; 1, 2 and 4 independent chains per lane):
            # HBM serves uniformly random reads at ~49 G sectors/s = 1.5-1.6 TB/s whatever the record size and however many reads a lane
            # has in flight -- the worst case for DRAM (every access opens a new page), reported beside the streaming rate
            dram_gather64_by_chains = {ch: pr.gather64(25, steps=32 // ch, chains=ch) for ch in (1, 2, 4)}
            dram_gather64 = max([v for v in dram_gather64_by_chains.values() if v] or [0.0]) or None
            dram_gather128 = max([v for v in (pr.gather128(24, steps=32, chains=1), pr.gather128(24, steps=16, chains=2)) if v] or [0.0]) or None
            hbm_stream4 = pr.stream(4 << 30, passes=1)
            flush4 = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

            def step4():
                mat4.clearAccum(); model4.renderCommand(None, 0, p4); mat4.resolve(8, 0.0)
            for _ in range(3):
                step4()
            torch.cuda.synchronize()
            mat4.resetCounters()
            evs = []
            for _ in range(3):
                flush4.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); step4(); e1.record(stream)
                evs.append((e0, e1))
            torch.cuda.synchronize()
            ms4 = sum(a.elapsed_time(b) for a, b in evs) / 3
            c4 = mat4.counters()
            roof4 = kernel_accounting(mat4, model4, 8, c4, 3, ms4, pr4, "c4_1080p_8spp", peak)
            # This working set is DRAM-resident: what binds is the rate of RANDOM 64-byte reads HBM delivers (far below its streaming
            # bandwidth), measured by the probe; the kernel's real DRAM bytes per launch come from the committed ncu capture.
            roof4["dram_gather32_peak_gbs"] = dram_gather * 32.0 if dram_gather else None
            roof4["dram_gather64_peak_gbs"] = dram_gather64 * 64.0 if dram_gather64 else None
            roof4["dram_gather64_gbs_by_chains_per_lane"] = {str(k): (v * 64.0 if v else None) for k, v in dram_gather64_by_chains.items()}
            roof4["dram_gather128_peak_gbs"] = dram_gather128 * 128.0 if dram_gather128 else None
            if roof4.get("hbm") and dram_gather64 and hbm_stream4:
                rnd = max(dram_gather64 * 64.0, (dram_gather128 or 0.0) * 128.0)
                roof4["hbm"]["peak"] = hbm_stream4
                roof4["hbm"]["frac"] = roof4["hbm"]["achieved"] / hbm_stream4
                roof4["hbm"]["peak_source"] = "in-process probe: coalesced read of 4 GB"
                roof4["hbm"]["uniform_random_rate"] = rnd
                roof4["hbm"]["achieved_over_uniform_random_rate"] = roof4["hbm"]["achieved"] / rnd
                roof4["bound"] = "hbm"
                roof4["l2_gather"] = {k: roof4[k] for k in ("achieved", "peak", "frac")}
                roof4["achieved"], roof4["peak"], roof4["frac"] = roof4["hbm"]["achieved"], hbm_stream4, roof4["hbm"]["frac"]
                roof4["what"] = ("DRAM bytes per trace launch (ncu dram__bytes_read+write, committed capture of this round) per second of kernel time (live CUDA events), against the "
                                 "HBM streaming rate measured by the in-process probe -- the only figure that bounds every access pattern.  hbm.uniform_random_rate is what the same probe gets "
                                 "for uniformly random 64- / 128-byte reads of a 2 GB table (the worst case: a new DRAM page per access); the kernel's record fetches are random too, but "
                                 "siblings and neighbouring leaves sit next to each other, and it runs above that rate (hbm.achieved_over_uniform_random_rate)")
            line["c4"] = {"workload": "synthetic %d-triangle lit box (seed %d), 1920x1080, 8 spp, depth 8, one GPU" % (len(sc4["triangles"]) // 48, args.scene_seed),
                          "fast_nodes": mat4.getInfo("fast_nodes"), "record_bytes": int(mat4.getInfo("fast_node_count")) * 64 + (len(sc4["triangles"]) // 48) * 64,
                          "value": c4.rays / 3 / (ms4 * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms4, "roofline": roof4, "scene_build_s": gen4}
            mat4.destroy()
        except Exception as ex:      # the extra leg must never take the headline down with it
            line["c4"] = {"error": repr(ex)}

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline and args.config in ("c2", "c3"):
            set_host_threads(os.cpu_count() or 1)
            dt1, rays1, kind = cpu_reference_step(scene, w, h, args.bounces, 1)
            frames = int(min(max(round(15.0 / max(dt1, 1e-3)), 1), 16))
            dt, r, kind = cpu_reference_step(scene, w, h, args.bounces, frames, first_sample=1)
            line["cpu_baseline"] = {"value": r / dt / 1e6, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": kind,
                                    "sample": "%d x 1-spp frames of the %dx%d image (of %d spp), reference traversal, pcg_ref RNG, %.1f s" % (frames, w, h, args.spp, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        if group is not None:
            group.close()
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
