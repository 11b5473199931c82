"""CPU tests of the oracle: known-answer vectors, the committed goldens generated from oracle/_ref, and -- where
oracle/_ref exists -- a live bit-for-bit comparison with the reference's own shader text."""
import json
import os

import numpy as np
import pytest

from conftest import CAM, GOLDEN, load_png, small_scene
from oracleharness import make_params

FACTS = json.load(open(os.path.join(GOLDEN, "ref_facts.json")))


def test_pcg_known_answers(oracle):
    # SURVEY.md 8c "PCG KAT" (random.glsl:4-17 evaluated exactly) and the stream dumped from oracle/_ref
    assert [float(x).hex() for x in oracle.random(0, 4)] == ["0x1.08ef2a0000000p-4", "0x1.9b14000000000p-15", "0x1.b998420000000p-1", "0x1.a7766a0000000p-1"]
    assert float(oracle.random(600, 1)[0]).hex() == "0x1.3e42b20000000p-1"
    for seed, vals in FACTS["pcg"].items():
        assert [float(x).hex() for x in oracle.random(int(seed), len(vals))] == vals


def test_pcg_words(oracle):
    import ctypes as C
    for seed, words in ((0, [0x108ef29b, 0x00033628, 0xdccc2102, 0xd3bb3506]), (1, [0x00033628, 0xdccc2102, 0xd3bb3506, 0xd977a704]),
                        (479999, [0x8faa0ffb, 0x9b2b74a8, 0xcb176003, 0x6120954f]), (0xFFFFFFFF, [0x106ee0ab, 0xef78d393, 0x52c6e4ab, 0xb320119d])):
        st = C.c_uint32(seed)
        oracle.lib.vcrt_oracle_pcg_next.restype = C.c_uint32
        assert [oracle.lib.vcrt_oracle_pcg_next(C.byref(st)) for _ in range(4)] == words


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_portable_sincos_accuracy(oracle):
    xs = np.linspace(0, 2 * np.pi, 4001).astype(np.float32)
    err = max(max(abs(oracle.sincos_portable(float(x))[0] - np.sin(np.float64(x))), abs(oracle.sincos_portable(float(x))[1] - np.cos(np.float64(x)))) for x in xs)
    assert err < 2e-7


@pytest.mark.parametrize("name,kw", [
    ("ref_full_b2_s16_800x600_f1.png", dict(shader="full", sample_count=1)),
    ("ref_full_b2_s16_800x600_f4.png", dict(shader="full", sample_count=4)),
    ("ref_full_b8_s16_800x600_f2.png", dict(shader="full", max_bounces=8, sample_count=2)),
    ("ref_full_b4_s16_800x600_f1.png", dict(shader="full", max_bounces=4, sample_count=1)),       # BASELINE configs[0]: 800x600, 1 spp, depth 4
    ("ref_simple_b4_s16_800x600_f1.png", dict(shader="simple", sample_count=1)),
    ("ref_full_b2_s16_800x600_f1_refdispatch.png", dict(shader="full", sample_count=1, flags=1)),
    ("ref_full_b2_s16_800x600_f2_lights1.png", dict(shader="full", sample_count=2, lights_length=1)),
])
def test_oracle_reproduces_reference_frames(oracle, doge, name, kw):
    """Golden rgba8 frames rendered by the reference's own shader (oracle/_ref) -- bit-exact."""
    want = load_png(name)
    got = oracle.render(doge, CAM, 800, 600, make_params(**kw))["target"]
    assert np.array_equal(got, want)


def test_oracle_reproduces_reference_glass_scene_frame(oracle, doge_glass):
    """BASELINE config 2's variant (glass box + metal box, depth 8): golden frame of the reference shader, bit-exact; the
    primary-hit material histogram includes glass (5) and metal (4)."""
    got = oracle.render(doge_glass, CAM, 800, 600, make_params(shader="full", max_bounces=8, sample_count=2), want_aov=True)
    assert np.array_equal(got["target"], load_png("ref_glass_full_b8_s16_800x600_f2.png"))
    want = FACTS["material_histogram"]["glass_800x600"]
    mats = got["aov"]["material"]
    for m in (0, 1, 2, 3, 4, 5):
        assert int((mats == m).sum()) == want[str(m)]
    assert want["5"] > 20000 and want["4"] > 9000


def test_oracle_reproduces_reference_hit_records(oracle, doge):
    g = np.load(os.path.join(GOLDEN, "ref_hits.npz"))
    out, tri = oracle.hit_bvh(doge, g["rays"])
    assert np.array_equal(out, g["hits"])
    assert (out[:, 0] == 1).sum() > 2000 and ((tri >= 0) == (out[:, 0] == 1)).all()


@pytest.mark.parametrize("w,h", [(800, 600), (1920, 1080)])
def test_primary_hit_material_histogram(oracle, doge, w, h):
    """SURVEY.md 8c soft goldens == histogram measured through oracle/_ref's hit_bvh."""
    res = oracle.render(doge, CAM, w, h, make_params(sample_count=1, flags=4), want_aov=True)
    mats = res["aov"]["material"]
    want = FACTS["material_histogram"]["%dx%d" % (w, h)]
    for m in (0, 1, 2, 3):
        assert int((mats == m).sum()) == want[str(m)]
    assert int((mats == -1).sum()) == want["4294967295"]
    c = res["counters"]
    assert c.max_stack <= 13          # bundled scene never reaches the 16-entry limit
    if (w, h) == (800, 600):          # BASELINE.md section 2 work-per-ray facts (primary rays only -> use a 1-bounce render)
        r1 = oracle.render(doge, CAM, w, h, make_params(sample_count=1, max_bounces=1, flags=4))["counters"]
        assert r1.rays == w * h
        assert abs(r1.ref_nodes / r1.rays - 24.7) < 0.1 and abs(r1.ref_triangles / r1.rays - 1.04) < 0.01
        assert abs(r1.canon_nodes / r1.rays - 22.4) < 0.1 and abs(r1.canon_triangles / r1.rays - 0.78) < 0.01


def test_live_reference_comparison(oracle, ref, doge):
    """Where oracle/_ref is built: other resolutions, depths and the stack-16 quirk, bit-for-bit."""
    for variant, kw in (("full_b2_s16", dict(shader="full")), ("full_b4_s16", dict(shader="full", max_bounces=4)),
                        ("simple_b4_s32", dict(shader="simple", stack_depth=32)), ("full_b8_s32", dict(shader="full", max_bounces=8, stack_depth=32))):
        a = ref.render_frames(variant, doge, CAM, 320, 200, 3)
        b = oracle.render(doge, CAM, 320, 200, make_params(sample_count=3, **kw))["target"]
        assert np.array_equal(a, b), variant
    # a deeper synthetic tree: with MAX_STACK_DEPTH 16 the reference drops part of the traversal; the oracle must too
    sc = small_scene(n_tris=20000, seed=3)
    a = ref.render_frames("full_b2_s16", sc, (0.0, 6.0, 1.5), 96, 64, 1)
    b = oracle.render(sc, (0.0, 6.0, 1.5), 96, 64, make_params(sample_count=1))
    assert np.array_equal(a, b["target"]) and b["counters"].max_stack == 16
    a32 = ref.render_frames("full_b8_s32", sc, (0.0, 6.0, 1.5), 96, 64, 1)
    b32 = oracle.render(sc, (0.0, 6.0, 1.5), 96, 64, make_params(sample_count=1, max_bounces=8, stack_depth=32))["target"]
    assert np.array_equal(a32, b32) and not np.array_equal(a, a32)


def test_glass_metal_scene_vs_reference(oracle, ref):
    sc = small_scene(n_tris=300, seed=11)
    a = ref.render_frames("full_b8_s32", sc, (0.0, 6.0, 1.5), 160, 120, 2)
    b = oracle.render(sc, (0.0, 6.0, 1.5), 160, 120, make_params(sample_count=2, max_bounces=8, stack_depth=32))["target"]
    assert np.array_equal(a, b)
    assert a[..., :3].max() > 0


def test_oracle_modes_consistency(oracle, doge):
    """f32 accumulation of N samples == sum of N single-sample renders; tile shards partition the image."""
    w, h = 160, 120
    full = oracle.render(doge, CAM, w, h, make_params(accum="f32", sample_count=4))["accumf"]
    acc = np.zeros_like(full)
    for s in range(4):
        oracle.render(doge, CAM, w, h, make_params(accum="f32", sample_begin=s, sample_count=1), accumf=acc)
    assert np.array_equal(full, acc)
    parts = [oracle.render(doge, CAM, w, h, make_params(accum="f32", sample_count=4, tile_rank=r, tile_count=3))["accumf"] for r in range(3)]
    assert np.array_equal(sum(parts), full)
    assert all((p[..., 3] > 0).sum() > 0 for p in parts)
    cover = sum((p[..., 3] > 0).astype(int) for p in parts)
    assert cover.max() == 1 and cover.min() == 1


def test_post_process_reproduces_reference_shader(oracle):
    """post-process-shader.frag:26-70 -- golden frames written by the reference's own fragment shader text (oracle/_ref), as
    shipped and with its commented-out smartDeNoise line enabled (:64: mix 0.5, sigma 2, kSigma 2, threshold 0.05): bit-exact."""
    img = load_png("ref_full_b2_s16_800x600_f4.png")
    assert np.array_equal(oracle.post_process(img, mix=0.0, gamma=2.2), load_png("ref_post_800x600_f4.png"))
    den = oracle.post_process(img, mix=0.5, sigma=2.0, k_sigma=2.0, threshold=0.05, gamma=2.2)
    want = load_png("ref_post_denoise_800x600_f4.png")
    assert np.array_equal(den, want)
    assert (want != load_png("ref_post_800x600_f4.png")).mean() > 0.01      # the denoiser does something


def test_post_process_live_reference_comparison(oracle, ref, doge):
    """Where oracle/_ref is built: other image sizes (odd sizes: REPEAT addressing at the borders), other frames."""
    for w, h, n in ((200, 150, 2), (97, 61, 1), (33, 17, 3)):
        img = oracle.render(doge, CAM, w, h, make_params(shader="full", sample_count=n))["target"]
        assert np.array_equal(ref.post_process(img, denoise=False), oracle.post_process(img, mix=0.0, gamma=2.2)), (w, h)
        assert np.array_equal(ref.post_process(img, denoise=True), oracle.post_process(img, mix=0.5, sigma=2.0, k_sigma=2.0, threshold=0.05, gamma=2.2)), (w, h)


def test_brute_force_reproduces_reference_shader(oracle):
    """hit_scene (ray-trace-compute.comp:222-247, the shader's commented-out alternative to hit_bvh at :322): goldens rendered by
    the reference's own text with that line enabled.  The triangle loop runs to ubo.numTriangles (:229) -- also when that is
    less than the buffer holds -- the sphere loop restarts from t_max (:238), and the simple shader has no sphere loop."""
    g = np.load(os.path.join(GOLDEN, "ref_brute_96x64.npz"))
    sc = small_scene(n_tris=30, seed=4)
    for shader, variant in (("full", "full_b2_s16_brute"), ("simple", "simple_b4_s16_brute")):
        for nt in (None, 20):
            got = oracle.render(sc, (0.0, 6.0, 1.5), 96, 64, make_params(shader=shader, traversal="brute_force", sample_count=2), num_triangles=nt)["target"]
            assert np.array_equal(got, g["%s_%s" % (variant, "all" if nt is None else nt)]), (variant, nt)
    assert not np.array_equal(g["full_b2_s16_brute_all"], g["full_b2_s16_brute_20"])
    assert not np.array_equal(g["full_b2_s16_brute_all"], g["simple_b4_s16_brute_all"])


def test_brute_force_live_reference_comparison(oracle, ref):
    for seed, n, nt in ((5, 12, None), (6, 40, 25), (7, 40, 60)):      # 60 > buffer length: reads past the end return the zero triangle
        sc = small_scene(n_tris=n, seed=seed)
        for shader, variant in (("full", "full_b2_s16_brute"), ("simple", "simple_b4_s16_brute")):
            a = ref.render_frames(variant, sc, (0.0, 6.0, 1.5), 64, 32, 2, num_triangles=nt)
            b = oracle.render(sc, (0.0, 6.0, 1.5), 64, 32, make_params(shader=shader, traversal="brute_force", sample_count=2), num_triangles=nt)["target"]
            assert np.array_equal(a, b), (seed, variant, nt)


def test_post_process_restatement(oracle, doge):
    """post-process-shader.frag restated: gamma-only path (the shipped shader) against numpy; smartDeNoise keeps flat
    regions, smooths Monte-Carlo noise and leaves hard edges in place."""
    img = oracle.render(doge, CAM, 200, 150, make_params(shader="full", sample_count=2))["target"]
    out = oracle.post_process(img, mix=0.0, gamma=2.2)
    want = np.rint(np.clip((img[..., :3] / np.float32(255.0)) ** np.float32(1 / 2.2), 0, 1) * 255.0)
    assert np.abs(out[..., :3].astype(int) - want.astype(int)).max() <= 1 and (out[..., 3] == 255).all()
    assert np.array_equal(oracle.post_process(img, mix=0.0, gamma=0.0)[..., :3], img[..., :3])
    flat = np.full((40, 50, 4), 77, np.uint8)
    assert np.array_equal(oracle.post_process(flat, mix=0.5, gamma=0.0)[..., :3], flat[..., :3])
    den = oracle.post_process(img, mix=1.0, sigma=2.0, k_sigma=2.0, threshold=0.3, gamma=0.0)
    lit = img[..., :3].astype(float).sum(-1) > 0
    rough = lambda a: np.abs(np.diff(a[..., :3].astype(float), axis=1))[lit[:, 1:] & lit[:, :-1]].mean()
    assert rough(den) < 0.8 * rough(img)
