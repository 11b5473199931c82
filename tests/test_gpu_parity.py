"""GPU parity tests (pytest -m gpu, run on the B200 box): the CUDA path, called through the C ABI via the
ComputeMaterial/ComputeModel mirror, against the CPU oracle on the same inputs and against the committed goldens
rendered by the reference's own shader (oracle/_ref).

Tolerances, stated once:
  * integer / index outputs (primary-hit triangle, material, back-face flag, ray counts in portable-trig mode): bit-exact;
  * primary-hit distance t: bit-exact (same fp32 operation order, -fmad=false);
  * trig_mode=portable: every output bit-exact (rgba8 frames, f32 accumulation);
  * trig_mode=libm (CUDA sinf/cosf vs glibc, <= 2 ulp apart): rgba8 frames within 1 LSB on >= 99.9 % of pixel-channels,
    f32 accumulation PSNR >= 45 dB against the oracle at equal sample count.
"""
import os

import numpy as np
import pytest

from conftest import CAM, GOLDEN, load_png, small_scene
from oracleharness import make_params

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def psnr(a, b, peak=1.0):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10 * np.log10(peak * peak / mse)


def frac_within_1lsb(a, b):
    return float((np.abs(a.astype(int) - b.astype(int)) <= 1).mean())


@pytest.fixture(scope="module")
def gpu_doge(doge):
    from gpuharness import GpuScene
    g = GpuScene(doge, 800, 600)
    yield g
    g.close()


@pytest.mark.parametrize("name,frames,kw", [
    ("ref_full_b2_s16_800x600_f1.png", 1, {}),
    ("ref_full_b2_s16_800x600_f4.png", 4, {}),
    ("ref_full_b2_s16_800x600_f1_refdispatch.png", 1, dict(full_cover=False)),
])
def test_dispatch_matches_reference_frames(gpu_doge, name, frames, kw):
    """computeCommand (vcrt_dispatch) vs golden rgba8 frames of the reference shader (libm trig -> 1-LSB tolerance)."""
    gpu_doge.material.clearAccum()
    got = gpu_doge.frames(CAM, frames, **kw)
    want = load_png(name)
    assert frac_within_1lsb(got, want) >= 0.999
    # pixels whose path never reaches a Lambertian bounce carry no sin/cos: alpha and all misses are exact
    assert np.array_equal(got[..., 3], want[..., 3])


def test_dispatch_simple_shader_matches_reference_frame(doge):
    from gpuharness import GpuScene
    g = GpuScene(doge, 800, 600, shader="ray-trace-compute-simple")
    got = g.frames(CAM, 1)
    g.close()
    # the simple shader has no transcendental: bit-exact
    assert np.array_equal(got, load_png("ref_simple_b4_s16_800x600_f1.png"))


def trav_kw(trav):
    """'fast' = wavefront pipeline; 'fast_mega' = persistent-warps megakernel; 'fast_static' = one thread per pixel; 'fast_auto' = what
    the library picks by the shape of the call."""
    if trav == "fast_static":
        return dict(traversal="fast", flags=8)
    if trav == "fast_mega":
        return dict(traversal="fast", flags=16)
    if trav == "fast":
        return dict(traversal="fast", flags=32)
    if trav == "fast_auto":
        return dict(traversal="fast")
    return dict(traversal=trav)


@pytest.mark.parametrize("trav", ["reference", "fast", "fast_mega", "fast_static"])
def test_primary_hits_bit_exact(gpu_doge, oracle, doge, trav):
    a = oracle.render(doge, CAM, 800, 600, make_params(sample_count=1), want_aov=True)
    b = gpu_doge.render(CAM, accum="rgba8_ref", sample_count=1, want_aov=True, **trav_kw(trav))
    for f in ("triangle", "material", "backFace"):
        assert np.array_equal(a["aov"][f], b["aov"][f]), f
    assert same_bits(a["aov"]["t"], b["aov"]["t"])
    assert int((a["aov"]["triangle"] >= 0).sum()) == 131420 + 34614 + 6093 + 1762


@pytest.mark.parametrize("trav", ["reference", "fast", "fast_mega", "fast_static"])
@pytest.mark.parametrize("shader,nb", [("full", 2), ("full", 8), ("simple", 4)])
def test_portable_trig_everything_bit_exact(gpu_doge, oracle, doge, trav, shader, nb):
    w, h = 800, 600
    for accum, rng, spp in (("rgba8_ref", "pcg_ref", 3), ("f32", "philox", 2)):
        kw = dict(shader=shader, max_bounces=nb, sample_count=spp, accum=accum, rng=rng, trig="portable")
        a = oracle.render(doge, CAM, w, h, make_params(traversal="reference", **kw))
        b = gpu_doge.render(CAM, **trav_kw(trav), **kw)
        key = "target" if accum == "rgba8_ref" else "accumf"
        assert same_bits(a[key], b[key]), (accum, rng)
        assert a["counters"].rays == b["counters"].rays


def test_libm_trig_within_tolerance(gpu_doge, oracle, doge):
    kw = dict(shader="full", max_bounces=8, sample_count=16, accum="f32", rng="philox", trig="libm")
    a = oracle.render(doge, CAM, 800, 600, make_params(**kw))["accumf"]
    b = gpu_doge.render(CAM, traversal="fast", **kw)["accumf"]
    assert np.array_equal(a[..., 3], b[..., 3])
    assert psnr(a[..., :3] / 16.0, b[..., :3] / 16.0) >= 45.0
    assert float((np.abs(a - b) > 1e-3).mean()) < 5e-3


def test_c1_config_reference_frame(gpu_doge, oracle, doge):
    """BASELINE configs[0] exactly: bundled scene, 800x600, 1 spp, max depth 4, as the reference's own shader renders it
    (golden from oracle/_ref built with NUM_BOUNCES 4).  The CUDA frame is within 1 LSB of the golden on >= 99.9 % of the
    pixel-channels (libm sinf/cosf), and bit-identical to the oracle in portable-trig mode for the reference-shaped and
    the fast traversal alike."""
    want = load_png("ref_full_b4_s16_800x600_f1.png")
    kw = dict(shader="full", max_bounces=4, sample_count=1, accum="rgba8_ref", rng="pcg_ref")
    for trav in ("reference", "fast"):
        got = gpu_doge.render(CAM, traversal=trav, trig="libm", **kw)["target"]
        assert frac_within_1lsb(got, want) >= 0.999, trav
        a = oracle.render(doge, CAM, 800, 600, make_params(traversal="reference", trig="portable", **kw))["target"]
        b = gpu_doge.render(CAM, traversal=trav, trig="portable", **kw)["target"]
        assert np.array_equal(a, b), trav


def test_c2_config_1080p(oracle, doge):
    """BASELINE config 2 shape: bundled scene, 1920x1080, depth 8, light sampling (4 of the 16 spp checked against the oracle)."""
    from gpuharness import GpuScene
    g = GpuScene(doge, 1920, 1080)
    kw = dict(shader="full", max_bounces=8, sample_count=4, accum="f32", rng="pcg_ref", trig="portable")
    a = oracle.render(doge, CAM, 1920, 1080, make_params(**kw), want_aov=True)
    b = g.render(CAM, traversal="fast", want_aov=True, **kw)
    g.close()
    assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"])
    mats = b["aov"]["material"]
    assert [int((mats == m).sum()) for m in (0, 1, 2, 3)] == [425240, 112140, 20164, 5706]


def test_c2_glass_scene_1080p(oracle, doge_glass):
    """BASELINE config 2 in full: bundled scene + glass box + metal box, 1920x1080, depth 8, light sampling.  (a) the
    reference shader's golden frame for this scene (800x600, 2 frames, rgba8 running mean) within 1 LSB (libm trig);
    (b) 1080p, portable trig: f32 accumulation and primary hits bit-exact against the oracle on 4 of the 16 spp;
    (c) the remaining 12 spp are covered by the slice property: 16 spp == 4 + 12 on the GPU."""
    from gpuharness import GpuScene
    g = GpuScene(doge_glass, 800, 600)
    got = g.render(CAM, traversal="fast", shader="full", max_bounces=8, sample_count=2, accum="rgba8_ref", rng="pcg_ref", trig="libm")["target"]
    g.close()
    want = load_png("ref_glass_full_b8_s16_800x600_f2.png")
    assert frac_within_1lsb(got, want) >= 0.999 and np.array_equal(got[..., 3], want[..., 3])
    g = GpuScene(doge_glass, 1920, 1080)
    kw = dict(shader="full", max_bounces=8, accum="f32", rng="philox", trig="portable", philox_seed=2)
    a = oracle.render(doge_glass, CAM, 1920, 1080, make_params(sample_count=4, **kw), want_aov=True)
    b = g.render(CAM, traversal="fast", want_aov=True, sample_count=4, **kw)
    assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"])
    mats = b["aov"]["material"]
    assert int((mats == 5).sum()) > 60000 and int((mats == 4).sum()) > 20000      # glass and metal are in view
    rest = g.render(CAM, traversal="fast", sample_begin=4, sample_count=12, clear=False, **kw)["accumf"]
    full = g.render(CAM, traversal="fast", sample_count=16, **kw)["accumf"]
    g.close()
    assert same_bits(rest, full)


@pytest.mark.parametrize("fmt", ["q15x4", "q15", "f32"])
def test_node_formats_bit_exact(oracle, fmt):
    """Every node format of the fast traversal (4-wide quantised = the default, binary quantised, 64-byte float) through
    the wavefront trace kernel and the one-thread-per-pixel kernel: boxes only cull, so every output bit equals the
    reference traversal's; the format actually in use is read back through vcrt_get_info."""
    from gpuharness import GpuScene
    sc = small_scene(n_tris=6000, seed=21)
    cam = (0.0, 6.0, 1.5)
    g = GpuScene(sc, 256, 192)
    g.material.setOption("fast_nodes", fmt)
    kw = dict(shader="full", max_bounces=6, sample_count=2, accum="f32", trig="portable", rng="philox", stack_depth=64)
    a = oracle.render(sc, cam, 256, 192, make_params(traversal="reference", **kw), want_aov=True)
    import vulkan_compute_ray_tracing_b200 as vcrt
    for trav in ("fast", "fast_static"):
        args = dict(kw, **trav_kw(trav))
        args["flags"] = args.get("flags", 0) | vcrt.FLAG_COUNT_TRAVERSAL
        b = g.render(cam, want_aov=True, **args)
        assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"]), (fmt, trav)
        assert a["counters"].rays == b["counters"].rays and b["counters"].nodes > 0
    assert g.material.getInfo("fast_nodes") == fmt
    g.close()


def test_glass_metal_deep_tree(oracle):
    from gpuharness import GpuScene
    sc = small_scene(n_tris=5000, seed=5)
    cam = (0.0, 6.0, 1.5)
    g = GpuScene(sc, 320, 240)
    kw = dict(shader="full", max_bounces=8, sample_count=2, accum="f32", trig="portable", stack_depth=64)
    a = oracle.render(sc, cam, 320, 240, make_params(traversal="reference", **kw), want_aov=True)
    for trav in ("reference", "fast", "fast_mega", "fast_static"):
        b = g.render(cam, want_aov=True, **trav_kw(trav), **kw)
        assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"]), trav
    g.close()
    # MAX_STACK_DEPTH 16 quirk (tree depth > 13 <=> more than 8192 triangles): the reference traversal truncates,
    # identically on both sides
    sc = small_scene(n_tris=20000, seed=3)
    g = GpuScene(sc, 96, 64)
    kw.update(sample_count=1, max_bounces=2)
    a64 = oracle.render(sc, cam, 96, 64, make_params(traversal="reference", **kw))
    kw["stack_depth"] = 16
    a16 = oracle.render(sc, cam, 96, 64, make_params(traversal="reference", **kw))
    b16 = g.render(cam, traversal="reference", **kw)
    g.close()
    assert a16["counters"].max_stack == 16
    assert same_bits(a16["accumf"], b16["accumf"]) and not same_bits(a16["accumf"], a64["accumf"])


def test_brute_force_spheres(oracle):
    from gpuharness import GpuScene
    sc = small_scene(n_tris=30, seed=4)
    cam = (0.0, 6.0, 1.5)
    g = GpuScene(sc, 64, 48)
    kw = dict(shader="full", traversal="brute_force", max_bounces=4, sample_count=2, accum="rgba8_ref", trig="portable")
    a = oracle.render(sc, cam, 64, 48, make_params(**kw), want_aov=True)
    b = g.render(cam, want_aov=True, **kw)
    g.close()
    assert same_bits(a["target"], b["target"]) and same_bits(a["aov"], b["aov"])


def test_brute_force_loop_bounds(oracle):
    """hit_scene's triangle loop runs to ubo.numTriangles (ray-trace-compute.comp:229) -- less than, equal to or beyond what the
    buffer holds -- and the simple shader's hit_scene has no sphere loop: against the goldens of the reference's own shader
    text (commented-out hit_scene line enabled) and against the oracle."""
    from gpuharness import GpuScene
    gold = np.load(os.path.join(GOLDEN, "ref_brute_96x64.npz"))
    sc = small_scene(n_tris=30, seed=4)
    cam = (0.0, 6.0, 1.5)
    for shader, variant in (("full", "full_b2_s16_brute"), ("simple", "simple_b4_s16_brute")):
        g = GpuScene(sc, 96, 64, shader="ray-trace-compute" if shader == "full" else "ray-trace-compute-simple")
        for nt in (None, 20, 50):
            kw = dict(shader=shader, traversal="brute_force", sample_count=2, accum="rgba8_ref", num_triangles=nt)
            b = g.render(cam, trig="libm", **kw)["target"]
            if nt != 50:
                assert frac_within_1lsb(b, gold["%s_%s" % (variant, "all" if nt is None else nt)]) >= 0.999, (variant, nt)
            kw.pop("num_triangles")
            a = oracle.render(sc, cam, 96, 64, make_params(trig="portable", **kw), num_triangles=nt, want_aov=True)
            b = g.render(cam, trig="portable", num_triangles=nt, want_aov=True, **kw)
            assert same_bits(a["target"], b["target"]) and same_bits(a["aov"], b["aov"]), (variant, nt)
        g.close()


def test_lights_length_quirk(gpu_doge, oracle, doge):
    """lights.length() is what the host's descriptor range makes it (ray-trace-compute.comp:76-77, Buffer.h:85: 1
    for the unmodified reference on a conformant driver), so light sampling only ever picks lights[0].  Golden frame of the
    reference shader rendered with that length; bit-exact against the oracle in portable-trig mode on every kernel."""
    want = load_png("ref_full_b2_s16_800x600_f2_lights1.png")
    kw = dict(shader="full", sample_count=2, accum="rgba8_ref", rng="pcg_ref", lights_length=1)
    got = gpu_doge.render(CAM, traversal="fast", trig="libm", **kw)["target"]
    assert frac_within_1lsb(got, want) >= 0.999 and np.array_equal(got[..., 3], want[..., 3])
    both = load_png("ref_full_b2_s16_800x600_f4.png")
    assert not np.array_equal(want, both)
    a = oracle.render(doge, CAM, 800, 600, make_params(traversal="reference", trig="portable", **kw))["target"]
    for trav in ("reference", "fast", "fast_mega", "fast_static"):
        b = gpu_doge.render(CAM, trig="portable", **trav_kw(trav), **kw)["target"]
        assert np.array_equal(a, b), trav
    deep = dict(kw, max_bounces=8, sample_count=3, accum="f32", rng="philox")      # the wavefront pipeline
    a = oracle.render(doge, CAM, 800, 600, make_params(traversal="reference", trig="portable", **deep))["accumf"]
    b = gpu_doge.render(CAM, traversal="fast", trig="portable", **deep)["accumf"]
    c = gpu_doge.render(CAM, traversal="fast", trig="portable", **dict(deep, lights_length=0))["accumf"]
    assert same_bits(a, b) and not same_bits(b, c)


def test_duplicate_triangles_tie_rule(oracle):
    """Every triangle twice: both copies are hit at the same t, and the reference keeps the first leaf of its right-first
    visiting order (strict `t < closest_so_far`, ray-trace-compute.comp:209).  The fast traversal visits in distance order and
    resolves the tie by slot rank (vcrt_fast.cuh: trav_leaf_test) -- on the GPU, for every node format and kernel."""
    from gpuharness import GpuScene
    import tinybvh
    sc = small_scene(n_tris=400, seed=9)
    tri = sc["triangles"].reshape(-1, 48)
    ntri = len(tri)
    perm = np.random.RandomState(3).permutation(2 * ntri)          # the two copies of a triangle land anywhere in the buffer
    t2 = np.concatenate([tri, tri])[perm].reshape(-1).copy()
    twin = np.empty(2 * ntri, np.int64)
    where = np.argsort(perm)                                          # position of original record k (k and k + ntri are the two copies)
    twin[where[:ntri]], twin[where[ntri:]] = where[ntri:], where[:ntri]
    sc2 = dict(sc)
    sc2["triangles"] = t2
    lights = sc2["lights"].view(tinybvh.LIGHT).copy()
    lights["triangleIndex"] = where[lights["triangleIndex"]]
    sc2["lights"] = lights.view(np.uint8).reshape(-1).copy()
    sc2["bvh"] = tinybvh.build_bvh(t2.view(tinybvh.TRI), seed=5, tie_seed=8).view(np.uint8).reshape(-1).copy()
    cam = (0.0, 6.0, 1.5)
    kw = dict(shader="full", max_bounces=4, sample_count=2, accum="f32", trig="portable", stack_depth=64)
    a = oracle.render(sc2, cam, 256, 192, make_params(traversal="reference", **kw), want_aov=True)
    hit = a["aov"]["triangle"][a["aov"]["triangle"] >= 0]
    assert len(hit) > 5000 and (hit < twin[hit]).any() and (hit > twin[hit]).any()      # neither "lowest index wins" nor "highest index wins"
    g = GpuScene(sc2, 256, 192)
    for fmt in ("q15x4", "q15", "f32"):
        g.material.setOption("fast_nodes", fmt)
        for trav in ("fast", "fast_mega", "fast_static"):
            b = g.render(cam, want_aov=True, **trav_kw(trav), **kw)
            assert same_bits(a["aov"], b["aov"]) and same_bits(a["accumf"], b["accumf"]), (fmt, trav)
    g.material.setOption("fast_nodes", "auto")
    g.material.setOption("fast_bvh", "topology")
    b = g.render(cam, want_aov=True, traversal="fast", **kw)
    g.close()
    assert same_bits(a["aov"], b["aov"]) and same_bits(a["accumf"], b["accumf"])


def test_absent_child_slots(oracle):
    """Inner nodes with one child, or with a child that has neither triangle nor children, under fast_bvh=topology: the 64-byte
    float nodes encode the absent slot as the box (+inf, -inf), which must not count as hit (it used to end the traversal early).
    Every node format, wavefront and one-launch kernels, bit-exact against the oracle."""
    from gpuharness import GpuScene
    import tinybvh
    sc = dict(small_scene(n_tris=900, seed=33))
    sc["bvh"] = tinybvh.add_degenerate_inner_nodes(sc["bvh"].view(tinybvh.NODE)).view(np.uint8).reshape(-1).copy()
    cam = (0.0, 6.0, 1.5)
    kw = dict(shader="full", max_bounces=5, sample_count=2, accum="f32", rng="philox", trig="portable", stack_depth=64)
    a = oracle.render(sc, cam, 192, 128, make_params(traversal="reference", **kw), want_aov=True)
    g = GpuScene(sc, 192, 128)
    for bvh in ("topology", "sah"):
        g.material.setOption("fast_bvh", bvh)
        for fmt in ("f32", "q15", "q15x4"):
            g.material.setOption("fast_nodes", fmt)
            for trav in ("fast", "fast_mega", "fast_static"):
                b = g.render(cam, want_aov=True, **trav_kw(trav), **kw)
                assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"]), (bvh, fmt, trav)
    g.close()


def test_wavefront_pipelines_and_counters(gpu_doge, oracle, doge):
    """(1) A wavefront render cut into parallel pipelines on 1, 2 or 4 streams (option wf_streams) gives the same bits.
    (2) Bounce 0 is traced once per pixel and shared by the pixel's samples: `rays` still counts every closest-hit query the
    shader asks for (equal to the oracle's count), `traversals` what was actually walked."""
    kw = dict(shader="full", max_bounces=8, sample_count=1, accum="f32", rng="philox", trig="portable")
    a = oracle.render(doge, CAM, 800, 600, make_params(traversal="reference", **kw))
    for streams in ("1", "2", "4", "auto"):
        gpu_doge.material.setOption("wf_streams", streams)
        b = gpu_doge.render(CAM, traversal="fast", flags=32, **kw)
        assert same_bits(a["accumf"], b["accumf"]) and a["counters"].rays == b["counters"].rays, streams
        assert b["counters"].launches > 1
    kw["sample_count"] = 4
    a = oracle.render(doge, CAM, 800, 600, make_params(traversal="reference", **kw))
    for streams in ("3", "auto"):
        gpu_doge.material.setOption("wf_streams", streams)
        b = gpu_doge.render(CAM, traversal="fast", **kw)
        c = b["counters"]
        assert same_bits(a["accumf"], b["accumf"]) and a["counters"].rays == c.rays
        assert c.primary_rays == 800 * 600 * 4 and c.traversals == c.rays - 800 * 600 * 3
    s = gpu_doge.render(CAM, traversal="fast", flags=8, **kw)["counters"]      # the one-launch kernel walks every query
    assert s.rays == c.rays and s.traversals == s.rays and s.primary_rays == c.primary_rays
    gpu_doge.material.setOption("wf_streams", "auto")
    # (3) "auto" runs calls of 32 Mi paths or more as two pipelines that share one batch's queue memory: same bits as one pipeline,
    # also when the batches are small and alternate between the pipelines many times
    from gpuharness import GpuScene
    g = GpuScene(doge, 1920, 1080)
    kw["sample_count"] = 17      # 1920 x 1056 covered pixels x 17 > 2^25
    one = None
    for streams, batch, pipelines in (("1", 256 << 20, 1), ("auto", 256 << 20, 2), ("auto", 3 << 20, 2), ("3", 5 << 20, 3)):
        g.material.setOption("wf_streams", streams)
        g.material.setOption("wf_batch_paths", str(batch))
        b = g.render(CAM, traversal="fast", **kw)
        assert int(g.material.getInfo("wf_pipelines")) == pipelines, (streams, batch)
        one = one or b
        assert same_bits(one["accumf"], b["accumf"]) and one["counters"].rays == b["counters"].rays, (streams, batch)
    kw["sample_count"] = 4
    g.material.setOption("wf_batch_paths", str(256 << 20))
    g.material.setOption("wf_streams", "auto")
    g.render(CAM, traversal="fast", **kw)
    assert int(g.material.getInfo("wf_pipelines")) == 1      # a small call stays one pipeline
    g.close()


def test_c4_size_scene(oracle):
    """BASELINE.json configs[3] at full size: 10 M triangles (0.8 GB of traversal records: 32-bit slot indices, the q15 extent
    rule and the stack4 bound of vcrt_repack.cpp in a regime the smaller scenes never reach), 3840x2160, depth 8.  (1) the
    oracle on 1/256 of the tiles, bit-exact (portable trig); (2) the 4-wide quantised tree and the 64-byte float nodes agree on
    every bit of the full 4K frame and on the ray count; (3) tile shards partition the frame."""
    from gpuharness import GpuScene
    from vulkan_compute_ray_tracing_b200 import scenegen
    sc = scenegen.generate_box_scene(10000000, seed=1234)
    w, h = 3840, 2160
    g = GpuScene(sc, w, h)
    kw = dict(shader="full", max_bounces=8, sample_count=1, accum="f32", rng="philox", trig="portable", stack_depth=64)
    full = g.render(CAM, traversal="fast", want_aov=True, **kw)
    assert g.material.getInfo("fast_nodes") == "q15x4" and int(g.material.getInfo("fast_node_count")) > 2000000
    a = oracle.render(sc, CAM, w, h, make_params(traversal="reference", tile_rank=77, tile_count=256, **kw), want_aov=True)
    own = a["accumf"][..., 3] > 0
    assert own.sum() > 30000 and (a["aov"]["triangle"][own] >= 0).mean() > 0.1
    assert same_bits(full["accumf"][own], a["accumf"][own]) and same_bits(full["aov"][own], a["aov"][own])
    b = g.render(CAM, traversal="fast", tile_rank=77, tile_count=256, **kw)
    assert a["counters"].rays == b["counters"].rays
    g.material.setOption("fast_nodes", "f32")
    o = g.render(CAM, traversal="fast", want_aov=True, **kw)
    assert same_bits(o["accumf"], full["accumf"]) and same_bits(o["aov"], full["aov"]) and o["counters"].rays == full["counters"].rays
    g.material.setOption("fast_nodes", "auto")
    parts = [g.render(CAM, traversal="fast", tile_rank=r, tile_count=8, **kw)["accumf"] for r in range(8)]
    assert same_bits(sum(parts), full["accumf"])
    g.close()


def test_device_record_build(oracle, doge):
    """fast_build=device: tie ranks, PLOC topology, 4-wide collapse and quantisation as CUDA kernels over the bound buffers where
    they lie (vcrt_devbuild.cu).  Any valid tree gives the same bits: against the oracle on small and awkward scenes, against the
    host-built records on the full-size C3 frame; scenes the device builder declines fall back to the host builder under "auto"
    and fail loudly under "device"."""
    from gpuharness import GpuScene
    from vulkan_compute_ray_tracing_b200 import scenegen
    import vulkan_compute_ray_tracing_b200 as vcrt
    import tinybvh
    cam = (0.0, 6.0, 1.5)
    cases = [(doge, CAM, 400, 300)]
    for seed, n in ((52, 2), (53, 3), (55, 40), (57, 4000)):
        cases.append((small_scene(n_tris=n, seed=seed), cam, 160, 120))
    dup = dict(small_scene(n_tris=200, seed=58))
    tri = dup["triangles"].reshape(-1, 48)
    dup["triangles"] = np.concatenate([tri, tri]).reshape(-1).copy()
    dup["bvh"] = tinybvh.build_bvh(dup["triangles"].view(tinybvh.TRI), seed=6, tie_seed=4).view(np.uint8).reshape(-1).copy()
    cases.append((dup, cam, 160, 120))
    odd = dict(small_scene(n_tris=300, seed=59))
    odd["bvh"] = tinybvh.add_degenerate_inner_nodes(odd["bvh"].view(tinybvh.NODE)).view(np.uint8).reshape(-1).copy()
    cases.append((odd, cam, 160, 120))
    for sc, c, w, h in cases:
        kw = dict(shader="full", max_bounces=5, sample_count=2, accum="f32", rng="philox", trig="portable", stack_depth=64)
        a = oracle.render(sc, c, w, h, make_params(traversal="reference", **kw), want_aov=True)
        g = GpuScene(sc, w, h)
        g.material.setOption("fast_build", "device")
        for trav in ("fast", "fast_static", "fast_mega"):
            b = g.render(c, want_aov=True, **trav_kw(trav), **kw)
            assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"]), (len(sc["triangles"]) // 48, trav)
            assert g.material.getInfo("fast_build") == "device"
        g.close()
    # a scene too large for 15-bit bounds: declined by the device builder
    big = dict(small_scene(n_tris=800, seed=12))
    t = big["triangles"].copy().view(np.float32).reshape(-1, 12)
    t[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]] *= 1000.0
    big["triangles"] = t.view(np.uint8).reshape(-1)
    big["bvh"] = tinybvh.build_bvh(big["triangles"].view(tinybvh.TRI), seed=2).view(np.uint8).reshape(-1).copy()
    g = GpuScene(big, 96, 64)
    g.material.setOption("fast_build", "device")
    with pytest.raises(vcrt.VcrtError, match="15-bit"):
        g.render((0.0, 6000.0, 1500.0), traversal="fast", accum="f32")
    g.material.setOption("fast_build", "auto")
    g.render((0.0, 6000.0, 1500.0), traversal="fast", accum="f32")
    assert g.material.getInfo("fast_build") == "host" and g.material.getInfo("fast_nodes") == "f32"
    g.close()
    # full size: 1 M triangles, 1080p -- device-built and host-built records give the same frame, bit for bit
    sc = scenegen.generate_box_scene(1000000, seed=1234)
    g = GpuScene(sc, 1920, 1080)
    kw = dict(shader="full", traversal="fast", max_bounces=8, sample_count=2, accum="f32", rng="philox", trig="portable", want_aov=True)
    host = g.render(CAM, **kw)
    assert g.material.getInfo("fast_build") == "host"
    g.material.setOption("fast_build", "device")
    dev = g.render(CAM, **kw)
    assert g.material.getInfo("fast_build") == "device" and g.material.getInfo("fast_nodes") == "q15x4"
    assert same_bits(host["accumf"], dev["accumf"]) and same_bits(host["aov"], dev["aov"]) and host["counters"].rays == dev["counters"].rays
    # the first device build of a process pays for loading the builder's kernels (hundreds of milliseconds on a cold box): what is
    # asserted is a repeated build -- the scene-change case the builder exists for
    g.material.updateStorageBuffer(0, sc["triangles"])
    assert g.material.getInfo("fast_build") == "device" and float(g.material.getInfo("fast_build_ms")) < 200.0
    g.close()


def test_edge_cases(oracle):
    """Empty scene, single triangle, ragged image sizes (not multiples of 32 / 8 / 4)."""
    from gpuharness import GpuScene
    import tinybvh
    cam = (0.0, 6.0, 1.5)
    for n, (w, h) in ((1, (33, 17)), (2, (70, 45)), (9, (1, 1)), (50, (37, 61))):
        sc = small_scene(n_tris=n, seed=n)
        g = GpuScene(sc, w, h)
        kw = dict(shader="full", max_bounces=4, sample_count=2, accum="f32", trig="portable")
        a = oracle.render(sc, cam, w, h, make_params(**kw), want_aov=True)
        for trav in ("reference", "fast", "fast_mega", "fast_static"):
            b = g.render(cam, want_aov=True, **trav_kw(trav), **kw)
            assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"]), (n, w, h, trav)
        g.close()
    sc = small_scene(n_tris=4, seed=1)
    empty = dict(sc)
    empty["triangles"] = np.zeros(0, np.uint8); empty["bvh"] = np.zeros(0, np.uint8); empty["lights"] = np.zeros(0, np.uint8)
    g = GpuScene(empty, 40, 40)
    for trav in ("reference", "fast", "fast_mega", "fast_static"):
        b = g.render(cam, accum="f32", sample_count=1, **trav_kw(trav))
        assert np.all(b["accumf"][..., :3] == 0) and np.all(b["accumf"][..., 3] == 1)
    g.close()


def test_sharding_properties(gpu_doge):
    """Size-independent properties at full size: tile shards partition the frame bit-exactly; sample slices add up."""
    kw = dict(shader="full", traversal="fast", max_bounces=4, accum="f32", rng="philox", trig="libm")
    full = gpu_doge.render(CAM, sample_count=8, **kw)["accumf"]
    parts = [gpu_doge.render(CAM, sample_count=8, tile_rank=r, tile_count=4, **kw)["accumf"] for r in range(4)]
    assert same_bits(sum(parts), full)
    cover = sum((p[..., 3] > 0).astype(int) for p in parts)
    assert cover.min() == 1 and cover.max() == 1
    a = gpu_doge.render(CAM, sample_begin=0, sample_count=4, **kw)["accumf"]
    b = gpu_doge.render(CAM, sample_begin=4, sample_count=4, **kw)["accumf"]
    assert np.allclose(a + b, full, rtol=1e-6, atol=1e-6)
    # the same for a shallow 1-spp frame, which the library renders with the one-launch kernel
    one = dict(kw, max_bounces=2)
    full1 = gpu_doge.render(CAM, sample_count=1, **one)
    assert full1["counters"].launches == 1
    parts1 = [gpu_doge.render(CAM, sample_count=1, tile_rank=r, tile_count=3, **one)["accumf"] for r in range(3)]
    assert same_bits(sum(parts1), full1["accumf"])
    deep1 = gpu_doge.render(CAM, sample_count=1, **dict(kw, max_bounces=8))      # deep 1-spp paths: one launch of the megakernel
    assert deep1["counters"].launches == 1
    many = gpu_doge.render(CAM, sample_count=2, **dict(kw, max_bounces=8))       # more samples: the wavefront pipeline
    assert many["counters"].launches > 1
    # resume: continuing on top of a reloaded accumulation equals the uninterrupted render bit-for-bit
    gpu_doge.material.clearAccum()
    gpu_doge.material.writeAccumF32(a)
    c = gpu_doge.render(CAM, sample_begin=4, sample_count=4, clear=False, **kw)["accumf"]
    assert same_bits(c, full)


def test_resolve_and_gamma(gpu_doge):
    kw = dict(shader="full", traversal="fast", max_bounces=2, accum="f32", sample_count=4)
    acc = gpu_doge.render(CAM, **kw)["accumf"]
    gpu_doge.material.resolve(4, 0.0)
    img = gpu_doge.target.read()
    want = np.rint(np.clip(acc / 4.0, 0, 1) * 255.0).astype(np.uint8)
    assert np.abs(img.astype(int) - want.astype(int)).max() <= 1
    gpu_doge.material.resolve(4, 2.2)      # post-process-shader.frag:67-68
    img = gpu_doge.target.read()
    want = np.rint(np.clip(acc[..., :3] / 4.0, 0, 1) ** (1 / 2.2) * 255.0).astype(np.uint8)
    assert np.abs(img[..., :3].astype(int) - want.astype(int)).max() <= 1


def test_c3_full_size_properties(oracle):
    """BASELINE.json configs[2] at full size (999 158 triangles, 1920x1080, depth 8), where the oracle cannot render the
    whole frame in seconds: (1) the oracle on 1/64 of the tiles, bit-exact (portable trig); (2) four independent
    traversals -- the wavefront kernel over the 4-wide tree, over the binary quantised tree, over float nodes, and the
    one-thread-per-pixel kernel -- agree on every bit of the full frame and on the ray count; (3) tile shards partition the
    frame; (4) a render continued on top of its first half equals the uninterrupted one."""
    from gpuharness import GpuScene
    from vulkan_compute_ray_tracing_b200 import scenegen
    import vulkan_compute_ray_tracing_b200 as vcrt
    sc = scenegen.generate_box_scene(1000000, seed=1234)
    w, h = 1920, 1080
    g = GpuScene(sc, w, h)
    kw = dict(shader="full", max_bounces=8, sample_count=2, accum="f32", rng="philox", trig="portable", stack_depth=64)
    full = g.render(CAM, traversal="fast", want_aov=True, **kw)
    assert g.material.getInfo("fast_nodes") == "q15x4"
    # (1) the oracle on every 64th tile
    a = oracle.render(sc, CAM, w, h, make_params(traversal="reference", tile_rank=5, tile_count=64, **kw), want_aov=True)
    b = g.render(CAM, traversal="fast", want_aov=True, tile_rank=5, tile_count=64, **kw)
    assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"])
    assert a["counters"].rays == b["counters"].rays
    own = a["accumf"][..., 3] > 0
    assert own.any() and same_bits(full["accumf"][own], a["accumf"][own]) and same_bits(full["aov"][own], a["aov"][own])
    # (2) independent traversals
    for fmt, flags in (("q15", 0), ("f32", 0), ("q15x4", vcrt.FLAG_STATIC_KERNEL)):
        g.material.setOption("fast_nodes", fmt)
        o = g.render(CAM, traversal="fast", want_aov=True, flags=flags, **kw)
        assert same_bits(o["accumf"], full["accumf"]) and same_bits(o["aov"], full["aov"]), (fmt, flags)
        assert o["counters"].rays == full["counters"].rays
    g.material.setOption("fast_nodes", "auto")
    # (3) tile shards
    parts = [g.render(CAM, traversal="fast", tile_rank=r, tile_count=3, **kw)["accumf"] for r in range(3)]
    assert same_bits(sum(parts), full["accumf"])
    # (4) resume
    half = dict(kw, sample_count=1)
    g.render(CAM, traversal="fast", **half)
    rest = g.render(CAM, traversal="fast", sample_begin=1, clear=False, **half)["accumf"]
    assert same_bits(rest, full["accumf"])
    g.close()


def test_batching_does_not_change_results(gpu_doge):
    """The wavefront pipeline cuts a call into batches of whole pixels (option wf_batch_paths); any batch size gives the same
    bits, ragged last batch and batches smaller than a tile included."""
    kw = dict(shader="full", traversal="fast", max_bounces=6, sample_count=3, accum="f32", rng="philox", trig="libm")
    want = gpu_doge.render(CAM, **kw)
    default = gpu_doge.material.getInfo("wf_batch_paths")
    for batch in (1024, 7 * 1024 + 3, 100000):
        gpu_doge.material.setOption("wf_batch_paths", str(batch))
        got = gpu_doge.render(CAM, **kw)
        assert same_bits(got["accumf"], want["accumf"]) and got["counters"].rays == want["counters"].rays, batch
        assert got["counters"].launches > want["counters"].launches
    gpu_doge.material.setOption("wf_batch_paths", default)


def test_headless_render_cli(gpu_doge, tmp_path):
    """python -m vulkan_compute_ray_tracing_b200.render: the EXR holds the linear mean of the samples the API returns, the
    PPM the post-processed 8-bit frame."""
    from vulkan_compute_ray_tracing_b200 import imageio, render
    scene_path = os.path.join(GOLDEN, "doge_scene.vcrt")
    kw = dict(shader="full", traversal="fast", max_bounces=8, sample_count=4, accum="f32", rng="philox")
    want = gpu_doge.render(CAM, **kw)["accumf"]
    assert render.main([scene_path, str(tmp_path / "o.exr"), "--width", "800", "--height", "600", "--spp", "4"]) == 0
    got = imageio.read_exr(tmp_path / "o.exr")
    assert same_bits(got, np.ascontiguousarray(want[..., :3] / 4.0))
    assert render.main([scene_path, str(tmp_path / "o.ppm"), "--width", "800", "--height", "600", "--spp", "4"]) == 0
    raw = (tmp_path / "o.ppm").read_bytes()
    assert raw.startswith(b"P6\n800 600\n255\n")
    img = np.frombuffer(raw[len(b"P6\n800 600\n255\n"):], np.uint8).reshape(600, 800, 3)
    ref = np.rint(np.clip(want[..., :3] / 4.0, 0, 1) * 255.0) / 255.0
    assert np.abs(img.astype(int) - np.rint(ref ** (1 / 2.2) * 255.0).astype(int)).max() <= 1


def test_cpp_example_matches_python_frames(tmp_path):
    """examples/headless_main.cpp -- the reference's frame loop written against include/vcrt/ComputeMaterial.hpp -- renders
    the same rgba8 frames as the Python mirror of the same classes (both are thin layers over the C ABI)."""
    import subprocess
    from gpuharness import GpuScene
    from test_boundary import build_cpp_example
    import vulkan_compute_ray_tracing_b200 as vcrt
    exe = build_cpp_example(tmp_path / "headless_main")
    scene_path = os.path.join(GOLDEN, "doge_scene.vcrt")
    out = subprocess.run([exe, scene_path, str(tmp_path / "o.ppm"), "3", "320", "200"], capture_output=True, text=True)
    assert out.returncode == 0 and "ms/frame" in out.stdout, out.stderr
    raw = (tmp_path / "o.ppm").read_bytes()
    hdr = b"P6\n320 200\n255\n"
    assert raw.startswith(hdr)
    img = np.frombuffer(raw[len(hdr):], np.uint8).reshape(200, 320, 3)
    g = GpuScene(vcrt.load_scene(scene_path), 320, 200)
    want = g.frames(CAM, 3)
    g.close()
    assert np.array_equal(img, want[..., :3])


@pytest.mark.parametrize("shader", ["ray-trace-compute", "ray-trace-compute-simple"])
def test_dispatch_auto_takes_the_fast_traversal(doge, shader):
    """computeCommand (vcrt_dispatch) walks the fast tree in the one-launch kernel whenever the reference's 16-entry stack
    cannot overflow on the bound tree -- same hit records by construction, so the frames are identical bit for bit to the
    literal traversal's; deeper trees keep the literal traversal (its truncation is part of the reference's behaviour)."""
    from gpuharness import GpuScene
    g = GpuScene(doge, 800, 600, shader=shader)
    g.material.setOption("dispatch_traversal", "reference")
    want = g.frames(CAM, 4)
    assert g.material.getInfo("dispatch_kernel") == "reference"
    g.material.setOption("dispatch_traversal", "auto")
    got = g.frames(CAM, 4)
    assert g.material.getInfo("dispatch_kernel") == "fast"
    assert np.array_equal(got, want)
    got = g.frames(CAM, 2, full_cover=False)          # the reference's floor(W/32) x floor(H/32) dispatch
    g.material.setOption("dispatch_traversal", "reference")
    assert np.array_equal(got, g.frames(CAM, 2, full_cover=False))
    g.close()
    deep = small_scene(n_tris=20000, seed=3)          # depth > 13: the shader's stack overflows, only the literal traversal reproduces that
    g = GpuScene(deep, 96, 64, shader=shader)
    g.frames((0.0, 6.0, 1.5), 1)
    assert g.material.getInfo("dispatch_kernel") == "reference"
    g.close()


def test_long_frame_loop_keeps_counting(doge):
    """A caller that renders frame after frame and reads the counters only at the end (the reference's mainLoop never reads
    any): timing events of finished frames are folded in on the way, nothing is lost and the frames still accumulate."""
    from gpuharness import GpuScene
    g = GpuScene(doge, 64, 64)
    g.material.resetCounters()
    got = g.frames(CAM, 600)
    c = g.material.counters()
    assert c.launches == 600 and c.kernel_ms > 0.0 and c.rays > 0
    assert got.shape == (64, 64, 4) and int(got[..., 3].min()) == 255
    g.close()


def test_error_behaviour(doge):
    import ctypes as C
    import vulkan_compute_ray_tracing_b200 as vcrt
    from vulkan_compute_ray_tracing_b200 import _native
    L = _native.lib()
    ctx = C.c_void_p()
    assert L.vcrt_create(0, C.byref(ctx)) == 0
    assert L.vcrt_dispatch(ctx, 1, 1, 1) < 0 and b"no storage images" in L.vcrt_last_error(ctx)
    assert L.vcrt_set_buffer(ctx, 3, doge["triangles"].ctypes.data, 47) < 0
    assert L.vcrt_set_buffer(ctx, 9, None, 0) < 0
    assert L.vcrt_set_shader(ctx, b"nope.spv") < 0 and b"failed to" in L.vcrt_last_error(ctx)
    assert L.vcrt_create(99, C.byref(C.c_void_p())) < 0
    L.vcrt_destroy(ctx)
    with pytest.raises(vcrt.VcrtError):
        vcrt.ComputeMaterial("bogus.spv").bind(None, 0)


def test_converged_image_psnr(oracle, doge):
    """North-star image gate: the converged f32 image (1024 spp, depth 8, light sampling) against the oracle's render of the
    same sample set: PSNR >= 45 dB and every pixel-channel within 2e-3 (libm trig differs by ulps between glibc and CUDA).
    A render with a different Philox seed must agree within Monte-Carlo noise (PSNR >= 30 dB at this sample count)."""
    from gpuharness import GpuScene
    w, h, spp = 160, 120, 1024
    kw = dict(shader="full", max_bounces=8, sample_count=spp, accum="f32", rng="philox", trig="libm", philox_seed=11)
    want = oracle.render(doge, CAM, w, h, make_params(**kw))["accumf"][..., :3] / spp
    g = GpuScene(doge, w, h)
    got = g.render(CAM, traversal="fast", **kw)["accumf"][..., :3] / spp
    kw["philox_seed"] = 12
    other = g.render(CAM, traversal="fast", **kw)["accumf"][..., :3] / spp
    g.close()
    assert psnr(np.clip(got, 0, 1), np.clip(want, 0, 1)) >= 45.0
    assert float(np.abs(got - want).max()) <= 2e-3
    assert psnr(np.clip(other, 0, 1), np.clip(want, 0, 1)) >= 30.0


def test_tile_pack_unpack(gpu_doge):
    """vcrt_pack_tiles / vcrt_unpack_tiles against the host-side layout definition (sharding.tile_pixel_index)."""
    import torch
    from vulkan_compute_ray_tracing_b200 import sharding
    m = gpu_doge.material
    w, h = gpu_doge.w, gpu_doge.h                      # 800 x 600: ragged bottom tile row
    full = gpu_doge.render(CAM, traversal="fast", shader="full", max_bounces=4, accum="f32", rng="philox", sample_count=2)["accumf"]
    m.resolve(2, 0.0)
    rgba = gpu_doge.target.read()
    for what, image, elem, dt in ((2, full, 16, np.float32), (0, rgba, 4, np.uint8)):
        flat = image.reshape(w * h, 4)
        for rank, world in ((0, 1), (1, 3), (2, 3), (7, 8)):
            n_tiles = sharding.max_owned_tiles(w, h, world)
            buf = torch.zeros(n_tiles * 1024 * elem, dtype=torch.uint8, device="cuda")
            m.packTiles(what, rank, world, buf.data_ptr(), buf.numel())
            m.synchronize()
            packed = buf.cpu().numpy().view(dt).reshape(-1, 4)
            idx = sharding.tile_pixel_index(w, h, rank, world, pad_tiles=n_tiles)
            want = np.zeros_like(packed)
            want[idx >= 0] = flat[idx[idx >= 0]]
            assert same_bits(packed, want), (what, rank, world)
    # unpack restores exactly the owned pixels
    m.clearAccum()
    idx = sharding.tile_pixel_index(w, h, 1, 3)
    n = len(idx) * 16
    src = np.zeros((len(idx), 4), np.float32)
    src[idx >= 0] = full.reshape(-1, 4)[idx[idx >= 0]]
    buf = torch.from_numpy(src.view(np.uint8).reshape(-1)).cuda()
    m.unpackTiles(2, 1, 3, buf.data_ptr(), n)
    got = m.readAccumF32().reshape(-1, 4)
    want = np.zeros_like(got)
    want[idx[idx >= 0]] = full.reshape(-1, 4)[idx[idx >= 0]]
    assert same_bits(got, want)
    with pytest.raises(Exception):
        m.packTiles(1, 0, 1, buf.data_ptr(), n)          # the rgba8 accumulation image is not tile-packed
    with pytest.raises(Exception):
        m.packTiles(2, 0, 1, buf.data_ptr(), 16)         # too small


def _two_gpu_worker(rank, world, port, mode, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import vulkan_compute_ray_tracing_b200 as vcrt
        from vulkan_compute_ray_tracing_b200 import sharding
        from gpuharness import GpuScene
        from refharness import load_scene
        scene = load_scene(os.path.join(GOLDEN, "doge_scene.vcrt"))
        w, h, spp = 800, 600, 6
        g = GpuScene(scene, w, h, device=rank)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        g.material.setStream(stream.cuda_stream)
        base = vcrt.render_params(shader="full", traversal="fast", rng="philox", accum="f32", max_bounces=4, sample_begin=3, sample_count=spp, philox_seed=5)
        g.set_camera(CAM, 0)
        g.material.clearAccum()
        p, active = sharding.shard_params(base, "tiles" if mode == "gather" else mode, rank, world)
        if active:
            g.model.renderCommand(None, 0, p)
        ptr, nbytes = g.material.devicePtr(2)

        class _Wrap:
            __cuda_array_interface__ = {"shape": (h, w, 4), "typestr": "<f4", "data": (ptr, False), "version": 2}
        acc = torch.as_tensor(_Wrap(), device="cuda")
        if mode == "gather":
            sharding.gather_tiles_device(g.material, 2, rank, world)
        else:
            sharding.reduce_accumulation(acc, dst=0)
        torch.cuda.synchronize()
        got = g.material.readAccumF32()
        if rank == 0 or mode == "gather":
            g.material.clearAccum()
            g.model.renderCommand(None, 0, base)
            want = g.material.readAccumF32()
            if mode == "samples":
                out[rank] = bool(np.allclose(got, want, rtol=1e-6, atol=1e-6)) and bool(np.array_equal(got[..., 3], want[..., 3]))
            else:
                out[rank] = bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
        else:
            out[rank] = True
        g.close()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["tiles", "samples", "gather"])
def test_two_gpus_match_one(mode):
    """N=2 over NCCL: tile shards (SUM reduce, and packed all-gather) reproduce the 1-GPU frame bit-for-bit; sample slices
    within fp32 reassociation.  Skipped on single-GPU boxes (tests/test_sharding.py covers the same logic with gloo)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_two_gpu_worker, args=(r, 2, port, mode, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=300)
        for p in procs:
            if p.is_alive():
                p.terminate()
            assert p.exitcode == 0
        assert dict(out) == {0: True, 1: True}


def test_group_api_single_process(oracle, doge, tmp_path):
    """vcrt_group_* (multi-GPU through the C ABI, NCCL inside the library).  (1) A local group of every GPU of the box -- one on
    the driver's box, where the path is exercised with world = 1 -- renders the same frame as a plain context, in both sharding
    modes.  (2) With two or more GPUs: tile shards reproduce the 1-GPU frame bit for bit ON EVERY GPU, sample slices within
    1 LSB of the resolved rgba8 frame.  (3) The C++ example's multi-GPU leg (examples/headless_main.cpp) writes the same
    picture for gpus = 1 and gpus = all."""
    import subprocess
    import torch
    import vulkan_compute_ray_tracing_b200 as vcrt
    from vulkan_compute_ray_tracing_b200 import sharding
    from gpuharness import GpuScene
    from test_boundary import build_cpp_example
    w, h, spp = 800, 600, 6
    p = vcrt.render_params(shader="full", traversal="fast", rng="philox", accum="f32", max_bounces=4, sample_begin=3, sample_count=spp, philox_seed=5)
    one = GpuScene(doge, w, h)
    one.set_camera(CAM, 0)
    one.material.clearAccum()
    one.model.renderCommand(None, 0, p)
    one.material.resolve(spp, 2.2)
    want = one.target.read()
    one.close()
    ubo = vcrt.pack_ubo(CAM, 0, doge)
    ngpu = torch.cuda.device_count()
    for n in sorted({1, min(ngpu, 2), ngpu}):
        g = sharding.LocalGroup(n, doge, w, h)
        g.render(ubo, p, "tiles", 2.2)
        for i in range(n):
            assert np.array_equal(g.read_target(i), want), ("tiles", n, i)
        g.render(ubo, p, "samples", 2.2)
        got = g.read_target(0)
        assert (np.array_equal(got, want) if n == 1 else frac_within_1lsb(got, want) == 1.0), ("samples", n)
        with pytest.raises(vcrt.VcrtError, match="f32 accumulation"):
            g.render(ubo, vcrt.render_params(accum="rgba8_ref"), "tiles")
        g.close()
    exe = build_cpp_example(tmp_path / "headless_main")
    scene_path = os.path.join(GOLDEN, "doge_scene.vcrt")
    outs = []
    for n in sorted({1, ngpu}):
        r = subprocess.run([exe, scene_path, str(tmp_path / ("g%d.ppm" % n)), "4", "320", "200", "full", str(n)], capture_output=True, text=True)
        assert r.returncode == 0 and "tile-sharded" in r.stdout, r.stderr
        outs.append((tmp_path / ("g%d.ppm" % n)).read_bytes())
    assert all(o == outs[0] for o in outs) and len(outs[0]) == len(b"P6\n320 200\n255\n") + 320 * 200 * 3


def _group_rank_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import vulkan_compute_ray_tracing_b200 as vcrt
        from vulkan_compute_ray_tracing_b200 import sharding
        from gpuharness import GpuScene
        from refharness import load_scene
        scene = load_scene(os.path.join(GOLDEN, "doge_scene.vcrt"))
        w, h, spp = 800, 600, 5
        g = GpuScene(scene, w, h, device=rank)
        grp = sharding.Group.from_torch(g.material)
        base = vcrt.render_params(shader="full", traversal="fast", rng="philox", accum="f32", max_bounces=4, sample_count=spp, philox_seed=9)
        g.set_camera(CAM, 0)
        ok = True
        for mode in ("tiles", "samples"):
            grp.render(g.model, base, mode, 0.0)
            got = g.target.read()
            g.material.clearAccum()
            g.model.renderCommand(None, 0, base)
            g.material.resolve(spp, 0.0)
            want = g.target.read()
            if mode == "tiles":
                ok = ok and bool(np.array_equal(got, want))
            elif rank == 0:
                ok = ok and bool((np.abs(got.astype(int) - want.astype(int)) <= 1).all())
        out[rank] = ok
        grp.close()
        g.close()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_group_api_one_process_per_gpu():
    """vcrt_group_create_rank under a torchrun-style launch (one process per GPU, the NCCL id carried by torch.distributed):
    the tile-sharded frame is bit-identical to the 1-GPU frame on every rank, sample slices within 1 LSB on rank 0."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_group_rank_worker, args=(r, 2, port, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=300)
        for p in procs:
            if p.is_alive():
                p.terminate()
            assert p.exitcode == 0
        assert dict(out) == {0: True, 1: True}


def test_post_process_matches_oracle(gpu_doge, oracle):
    """vcrt_post_process (CUDA) vs the restated fragment shader: gamma-only exactly as shipped, and with smartDeNoise enabled
    as its commented-out call would (mix 0.5, sigma 2, kSigma 2, threshold 0.05).  expf/powf differ by ulps between CUDA and
    glibc: <= 1 LSB on >= 99.9 % of channels, never more than 2."""
    # the reference's own fragment shader text (oracle/_ref) over the reference's own 4-frame golden, as shipped and with the
    # denoiser line enabled: the golden is put into the target image through the f32 accumulation (k/255 resolves back to k)
    gold = load_png("ref_full_b2_s16_800x600_f4.png")
    gpu_doge.material.clearAccum()
    gpu_doge.material.writeAccumF32((gold.astype(np.float32) / np.float32(255.0)))
    gpu_doge.material.resolve(1, 0.0)
    assert np.array_equal(gpu_doge.target.read(), gold)
    for name, kw in (("ref_post_800x600_f4.png", dict(mix=0.0, gamma=2.2)), ("ref_post_denoise_800x600_f4.png", dict(mix=0.5, sigma=2.0, k_sigma=2.0, threshold=0.05, gamma=2.2))):
        d = np.abs(gpu_doge.material.postProcess(**kw).astype(int) - load_png(name).astype(int))
        assert (d <= 1).mean() >= 0.999 and d.max() <= 2, name
    gpu_doge.material.clearAccum()
    img = gpu_doge.frames(CAM, 3)
    for kw in (dict(mix=0.0, gamma=2.2), dict(mix=0.5, sigma=2.0, k_sigma=2.0, threshold=0.05, gamma=2.2), dict(mix=1.0, sigma=3.0, k_sigma=2.0, threshold=0.2, gamma=0.0)):
        got = gpu_doge.material.postProcess(**kw)
        want = oracle.post_process(img, **kw)
        d = np.abs(got.astype(int) - want.astype(int))
        assert (d <= 1).mean() >= 0.999 and d.max() <= 2, kw
        assert (got[..., 3] == 255).all()
    with pytest.raises(Exception):
        gpu_doge.material.postProcess(mix=0.5, sigma=0.0)


def test_scripted_frame_loop(oracle, doge):
    """The reference's frame loop with a scripted keyboard (frameloop.py): accumulation restarts on every camera move
    (main.cpp:169-173), so after "..w.." the target holds frames 0..2 of the moved camera -- checked against the oracle."""
    import vulkan_compute_ray_tracing_b200 as vcrt
    from gpuharness import GpuScene
    g = GpuScene(doge, 320, 240)
    loop = vcrt.FrameLoop(g.model, doge, 320, 240, frame_time=0.5)
    img = loop.run("...")
    want = oracle.render(doge, CAM, 320, 240, make_params(sample_count=3))["target"]
    assert frac_within_1lsb(img, want) >= 0.999
    img = loop.run("w..")
    cam = tuple(float(x) for x in loop.camera.Position)
    assert cam != tuple(CAM) and loop.currentSample == 3 and loop.frames == 6
    want = oracle.render(doge, cam, 320, 240, make_params(sample_count=3))["target"]
    assert frac_within_1lsb(img, want) >= 0.999
    with pytest.raises(ValueError):
        loop.run("x")
    g.close()


@pytest.mark.parametrize("nslots", [1, 2, 4])
def test_frames_in_flight_dispatch_bit_identical(doge, nslots):
    """The reference's pipelined loop (MAX_FRAMES_IN_FLIGHT, main.cpp:68, :325, :394): every presented frame of the pipelined
    FrameLoop -- camera moves (accumulation restarts) included -- equals the frame the synchronous computeCommand sequence leaves in
    the target, bit for bit."""
    import vulkan_compute_ray_tracing_b200 as vcrt
    from gpuharness import GpuScene
    script = "....w..a.d..."
    g = GpuScene(doge, 320, 240)
    sync = vcrt.FrameLoop(g.model, doge, 320, 240, frame_time=0.5)
    want = []
    for key in script:
        sync.drawFrame(key)
        want.append(g.target.read())
    g.close()
    g = GpuScene(doge, 320, 240)
    got = {}
    loop = vcrt.FrameLoop(g.model, doge, 320, 240, frame_time=0.5, frames_in_flight=nslots, on_present=lambda i, a: got.__setitem__(i, a.copy()))
    final = loop.run(script)
    assert sorted(got) == list(range(len(script)))
    for i in range(len(script)):
        assert same_bits(got[i], want[i]), i
    assert same_bits(final, want[-1]) and same_bits(g.accum.read(), want[-1])
    # the synchronous interface works again after the loop has finished, and refuses to run while frames are in flight
    g.material.framesBegin(2)
    with pytest.raises(vcrt.VcrtError, match="frames are in flight"):
        g.target.read()
    with pytest.raises(vcrt.VcrtError, match="frames are in flight"):
        g.model.computeCommand(None, 0, 10, 8, 1)
    g.material.framesEnd()
    assert same_bits(g.target.read(), want[-1])
    with pytest.raises(vcrt.VcrtError, match="vcrt_frames_begin"):
        g.model.frameCommand(None, 0, 10, 8, 1)
    g.close()


@pytest.mark.parametrize("family", ["static", "mega", "wavefront", "reference"])
@pytest.mark.parametrize("accum", ["f32", "rgba8_ref"])
def test_frames_in_flight_render_params(doge, family, accum):
    """vcrt_frame_submit with run-time parameters, every kernel family: n progressive 1-spp frames in flight against the same
    frames through vcrt_render + vcrt_resolve + read-back, bit for bit (f32: resolved frame and the accumulation buffer itself)."""
    import vulkan_compute_ray_tracing_b200 as vcrt
    from gpuharness import GpuScene
    w, h, n = 200, 136, 7      # ragged: not a multiple of the 32x32 tiles
    flags = {"static": vcrt.FLAG_STATIC_KERNEL, "mega": vcrt.FLAG_MEGAKERNEL, "wavefront": vcrt.FLAG_WAVEFRONT, "reference": 0}[family]
    kw = dict(shader="full", traversal="reference" if family == "reference" else "fast", rng="philox", accum=accum, trig="libm", max_bounces=5,
              stack_depth=64, sample_count=1, philox_seed=7, flags=flags)
    g = GpuScene(doge, w, h)
    want = []
    g.set_camera(CAM, 0)
    g.material.clearAccum()
    for k in range(n):
        p = vcrt.render_params(**kw, sample_begin=k)
        g.model.renderCommand(None, 0, p)
        if accum == "f32":
            g.material.resolve(k + 1, 2.2)
        want.append(g.target.read())
    want_acc = g.material.readAccumF32() if accum == "f32" else g.accum.read()
    g.close()
    g = GpuScene(doge, w, h)
    g.set_camera(CAM, 0)
    # stale accumulation: sample 0 restarts it (f32 explicitly, rgba8 through the running mean's zero weight)
    g.model.renderCommand(None, 0, vcrt.render_params(**kw, sample_begin=3))
    bufs = [vcrt.PinnedFrame(w, h) for _ in range(3)]
    g.material.framesBegin(3)
    got, inflight = [], {}
    for k in range(n):
        p = vcrt.render_params(**kw, sample_begin=k)
        buf = bufs[k % 3]
        if (k % 3) in inflight:
            g.material.frameWait(k % 3)
            got.append(buf.array.copy())
        slot = g.material.frameSubmit(p, total_samples=k + 1, gamma=2.2, out=buf)
        assert slot == k % 3
        inflight[slot] = k
    for k in range(n, n + 3):
        if len(got) < n:
            g.material.frameWait(k % 3)
            got.append(bufs[k % 3].array.copy())
    g.material.framesEnd()
    assert len(got) == n
    for k in range(n):
        assert same_bits(got[k], want[k]), (family, accum, k)
    assert same_bits(g.target.read(), want[-1])
    assert same_bits(g.material.readAccumF32() if accum == "f32" else g.accum.read(), want_acc)
    c = g.material.counters()
    assert c.rays > 0 and c.kernel_ms > 0.0
    for b in bufs:
        b.free()
    g.close()


@pytest.mark.parametrize("fmt", ["q15x4", "q15", "f32"])
def test_tail_loop_work_sharing(oracle, fmt):
    """The trace kernel's tail loop (idle lanes of a warp walk subtrees handed over by its busy lanes, results merged by the
    traversal's own rule): launches with fewer rays than lanes on a random-triangle soup -- long rays, deep stacks, plenty of
    equal-t candidates among the 200 duplicated triangles -- run almost entirely in it.  Every node format, against the oracle and
    against the kernels that have no such loop: bit-identical f32 accumulation and primary hits."""
    from gpuharness import GpuScene
    import tinybvh
    sc = dict(small_scene(n_tris=20000, seed=11))
    tris = sc["triangles"].reshape(-1, 48)
    t2 = np.concatenate([tris, tris[:200]]).reshape(-1).copy()      # 200 triangles twice: the tie rule must pick the same copy whichever lane finds which
    sc["triangles"] = t2
    sc["bvh"] = tinybvh.build_bvh(t2.view(tinybvh.TRI), seed=4, tie_seed=6).view(np.uint8).reshape(-1).copy()
    cam = (0.0, 6.0, 1.5)
    w, h = 96, 64
    kw = dict(shader="full", max_bounces=8, sample_count=3, accum="f32", rng="philox", trig="portable", philox_seed=5, stack_depth=64)
    a = oracle.render(sc, cam, w, h, make_params(**kw), want_aov=True)
    g = GpuScene(sc, w, h)
    g.material.setOption("fast_nodes", fmt)
    b = g.render(cam, traversal="fast", flags=32, want_aov=True, **kw)      # wavefront pipeline (tail loop)
    assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"]), fmt
    for flags in (8, 16):                                                     # one thread per pixel, megakernel
        c = g.render(cam, traversal="fast", flags=flags, want_aov=True, **kw)
        assert same_bits(b["accumf"], c["accumf"]) and same_bits(b["aov"], c["aov"]), (fmt, flags)
    g.close()


@pytest.mark.parametrize("family", ["static", "wavefront"])
def test_frames_in_flight_tile_shard_and_coverage(doge, family):
    """Frames in flight that cover only part of the image -- a tile shard (tile_rank / tile_count) and the reference's dispatch
    coverage (floor(W/32) x floor(H/32) groups, main.cpp:228): the fold touches the covered pixels only, the rest of the target and of
    the accumulation stay as the synchronous calls leave them."""
    import vulkan_compute_ray_tracing_b200 as vcrt
    from gpuharness import GpuScene
    w, h, n = 200, 136, 4
    flags = {"static": vcrt.FLAG_STATIC_KERNEL, "wavefront": vcrt.FLAG_WAVEFRONT}[family] | vcrt.FLAG_REF_DISPATCH_COVERAGE
    kw = dict(shader="full", traversal="fast", rng="philox", accum="f32", trig="libm", max_bounces=3, sample_count=1, philox_seed=3, flags=flags,
              tile_rank=1, tile_count=3)
    frames = {}
    for mode in ("sync", "in_flight"):
        g = GpuScene(doge, w, h)
        g.set_camera(CAM, 0)
        g.material.clearAccum()
        out = []
        if mode == "sync":
            for k in range(n):
                g.model.renderCommand(None, 0, vcrt.render_params(**kw, sample_begin=k))
                g.material.resolve(k + 1, 0.0)
                out.append(g.target.read())
        else:
            bufs = [vcrt.PinnedFrame(w, h) for _ in range(2)]
            g.material.framesBegin(2)
            for k in range(n):
                g.material.frameWait(k % 2)
                if k >= 2:
                    out.append(bufs[k % 2].array.copy())
                g.material.frameSubmit(vcrt.render_params(**kw, sample_begin=k), total_samples=k + 1, out=bufs[k % 2])
            for k in range(n, n + 2):
                g.material.frameWait(k % 2)
                out.append(bufs[k % 2].array.copy())
            g.material.framesEnd()
            for b in bufs:
                b.free()
        frames[mode] = (out, g.material.readAccumF32(), g.target.read())
        g.close()
    assert len(frames["sync"][0]) == len(frames["in_flight"][0]) == n
    for k in range(n):
        assert same_bits(frames["sync"][0][k], frames["in_flight"][0][k]), (family, k)
    assert same_bits(frames["sync"][1], frames["in_flight"][1]) and same_bits(frames["sync"][2], frames["in_flight"][2])
    covered = frames["sync"][1][..., 3] > 0
    assert 0 < covered.sum() < w * h / 2      # a third of the tiles, minus the uncovered border


@pytest.mark.parametrize("family,nslots", [("wavefront", 4), ("mega", 2), ("auto", 3)])
def test_frames_in_flight_long_run_equals_one_render(doge, family, nslots):
    """300 progressive 1-spp frames in flight leave exactly the f32 accumulation that ONE 300-spp render leaves (samples folded in
    sample order in both): the frame-order chain of the folds holds under sustained overlap, and the trace kernel's tail loop -- which
    most of every small launch runs in -- never changes a hit."""
    import vulkan_compute_ray_tracing_b200 as vcrt
    from gpuharness import GpuScene
    w, h, n = 320, 240, 300
    flags = {"wavefront": vcrt.FLAG_WAVEFRONT, "mega": vcrt.FLAG_MEGAKERNEL, "auto": 0}[family]
    kw = dict(shader="full", traversal="fast", rng="philox", accum="f32", trig="libm", max_bounces=8, philox_seed=11)
    g = GpuScene(doge, w, h)
    g.set_camera(CAM, 0)
    g.material.clearAccum()
    g.model.renderCommand(None, 0, vcrt.render_params(**kw, sample_begin=0, sample_count=n, flags=vcrt.FLAG_WAVEFRONT))
    want = g.material.readAccumF32()
    g.material.clearAccum()
    g.material.framesBegin(nslots)
    for k in range(n):
        g.material.frameSubmit(vcrt.render_params(**kw, sample_begin=k, sample_count=1, flags=flags), total_samples=k + 1)
    g.material.framesEnd()
    got = g.material.readAccumF32()
    assert same_bits(got, want), family
    assert float(got[..., 3].min()) == n
    g.close()
