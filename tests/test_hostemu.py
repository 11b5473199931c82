"""CPU tests: the product's own path code (vcrt_path.cuh / vcrt_fast.cuh / vcrt_repack.cpp), compiled for the host by
tests/hostemu, against the oracle -- bit-exact.  This catches logic errors in the kernels' source before any GPU time
is spent; the GPU parity tests (test_gpu_parity.py) are the real gate."""
import numpy as np
import pytest

from conftest import CAM, small_scene
from oracleharness import make_params


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


@pytest.mark.parametrize("trav", ["reference", "fast"])
@pytest.mark.parametrize("shader,nb", [("full", 2), ("full", 8), ("simple", 4)])
def test_bundled_scene(oracle, hostemu, doge, trav, shader, nb):
    w, h = 200, 150
    for accum, rng in (("rgba8_ref", "pcg_ref"), ("f32", "philox")):
        kw = dict(shader=shader, max_bounces=nb, sample_count=3, accum=accum, rng=rng)
        a = oracle.render(doge, CAM, w, h, make_params(traversal="reference", **kw), want_aov=True)
        b = hostemu.render(doge, CAM, w, h, make_params(traversal=trav, **kw), want_aov=True)
        key = "target" if accum == "rgba8_ref" else "accumf"
        assert same_bits(a[key], b[key])
        assert same_bits(a["aov"], b["aov"])
        assert a["counters"].rays == b["rays"]


def test_fast_equals_reference_on_deep_random_trees(oracle, hostemu):
    """Ties, degenerate triangles, glass/metal, trees deeper than 16: fast traversal == oracle (stack 64)."""
    for seed, n in ((1, 1), (2, 2), (3, 7), (4, 500), (5, 5000)):
        sc = small_scene(n_tris=n, seed=seed)
        kw = dict(shader="full", max_bounces=6, sample_count=2, accum="f32", stack_depth=64)
        a = oracle.render(sc, (0.0, 6.0, 1.5), 96, 64, make_params(traversal="reference", **kw), want_aov=True)
        b = hostemu.render(sc, (0.0, 6.0, 1.5), 96, 64, make_params(traversal="fast", **kw), want_aov=True)
        assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"]), (seed, n)


@pytest.mark.parametrize("reserved", [8, 12])
def test_wide_tree_on_small_and_random_scenes(oracle, hostemu, reserved):
    """The 4-wide form of the quantised tree (build_wide_bvh) on the awkward sizes -- one to five triangles (root with fewer
    than four children, a single leaf, no inner node at all), duplicated triangles (ties across sibling leaves), a few
    hundred random ones with glass and metal: every bit equals the reference traversal's.  `reserved` 8 = automatic
    quantisation limit, 12 = quantised whatever the scene extent."""
    cam = (0.0, 6.0, 1.5)
    for seed, n in ((11, 1), (12, 2), (13, 3), (14, 4), (15, 5), (16, 9), (17, 33), (18, 300), (19, 2000)):
        sc = small_scene(n_tris=n, seed=seed)
        kw = dict(shader="full", max_bounces=6, sample_count=2, accum="f32", rng="philox", stack_depth=64)
        a = oracle.render(sc, cam, 80, 56, make_params(traversal="reference", **kw), want_aov=True)
        p = make_params(traversal="fast", **kw)
        p._reserved = reserved
        b = hostemu.render(sc, cam, 80, 56, p, want_aov=True)
        assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"]), (seed, n)
        assert a["counters"].rays == b["rays"]


def test_duplicate_triangles_tie_rule(oracle, hostemu):
    """Every triangle duplicated: equal t for both copies; the reference keeps the first leaf in its visiting order."""
    sc = small_scene(n_tris=40, seed=9)
    tri = sc["triangles"].reshape(-1, 48)
    import tinybvh
    t2 = np.concatenate([tri, tri]).reshape(-1).copy()
    tt = t2.view(tinybvh.TRI)
    sc2 = dict(sc)
    sc2["triangles"] = t2
    sc2["bvh"] = tinybvh.build_bvh(tt, seed=5, tie_seed=8).view(np.uint8).reshape(-1).copy()
    kw = dict(shader="full", max_bounces=4, sample_count=1, accum="f32", stack_depth=64)
    a = oracle.render(sc2, (0.0, 6.0, 1.5), 128, 96, make_params(traversal="reference", **kw), want_aov=True)
    b = hostemu.render(sc2, (0.0, 6.0, 1.5), 128, 96, make_params(traversal="fast", **kw), want_aov=True)
    assert same_bits(a["aov"], b["aov"]) and same_bits(a["accumf"], b["accumf"])
    hit = a["aov"]["triangle"][a["aov"]["triangle"] >= 0]
    n = len(tri)
    assert len(hit) > 1000 and (hit < n).any() and (hit >= n).any()      # winners come from both copies: neither lowest nor highest index wins


def test_absent_child_slots_in_float_nodes(oracle, hostemu):
    """The bound topology kept as it is (fast_bvh=topology) with 64-byte float nodes, on a tree whose inner nodes have one
    child or a child with neither triangle nor children: an absent slot is the box (+inf, -inf), which a min/max slab test
    reports as hit unless the verdict is gated on the child code (it used to end the ray's traversal early: wrong hits)."""
    import tinybvh
    for seed, n in ((31, 7), (32, 64), (33, 900)):
        sc = dict(small_scene(n_tris=n, seed=seed))
        nodes = tinybvh.add_degenerate_inner_nodes(sc["bvh"].view(tinybvh.NODE))
        sc["bvh"] = nodes.view(np.uint8).reshape(-1).copy()
        kw = dict(shader="full", max_bounces=5, sample_count=2, accum="f32", rng="philox", stack_depth=64)
        a = oracle.render(sc, (0.0, 6.0, 1.5), 96, 64, make_params(traversal="reference", **kw), want_aov=True)
        for reserved in (3, 1, 9, 0):   # topology + f32 | topology + q15 | topology + q15x4 | SAH rebuild + q15
            p = make_params(traversal="fast", **kw)
            p._reserved = reserved
            b = hostemu.render(sc, (0.0, 6.0, 1.5), 96, 64, p, want_aov=True)
            assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"]), (seed, n, reserved)
        assert (a["aov"]["triangle"] >= 0).sum() > 100


def test_deep_bound_tree_is_fine_once_rebuilt(oracle, hostemu):
    """A degenerate (list-shaped) bound tree, deeper than the fast traversal's stack: keeping its topology is refused, the
    SAH rebuild (the default) walks its own shallow tree and reproduces the reference traversal's hits."""
    import tinybvh
    sc = dict(small_scene(n_tris=46, seed=41))   # 50 triangles with the floor and the emitter: depth 49 > 45, reference stack 51 < 64
    tri = sc["triangles"].view(tinybvh.TRI)
    n = len(tri)
    lo = np.minimum(np.minimum(tri["v0"], tri["v1"]), tri["v2"]) - np.float32(1e-4)
    hi = np.maximum(np.maximum(tri["v0"], tri["v1"]), tri["v2"]) + np.float32(1e-4)
    nodes = np.zeros(2 * n - 1, tinybvh.NODE)
    nodes["left"] = nodes["right"] = nodes["object"] = -1
    for i in range(n - 1):          # inner node 2i: {leaf 2i+1, rest 2i+2}
        nodes[2 * i]["left"], nodes[2 * i]["right"] = 2 * i + 1, 2 * i + 2
        nodes[2 * i]["min"], nodes[2 * i]["max"] = lo[i:].min(axis=0), hi[i:].max(axis=0)
        nodes[2 * i + 1]["object"] = i
        nodes[2 * i + 1]["min"], nodes[2 * i + 1]["max"] = lo[i], hi[i]
    nodes[2 * n - 2]["object"] = n - 1
    nodes[2 * n - 2]["min"], nodes[2 * n - 2]["max"] = lo[n - 1], hi[n - 1]
    sc["bvh"] = nodes.view(np.uint8).reshape(-1).copy()
    kw = dict(shader="full", max_bounces=3, sample_count=1, accum="f32", stack_depth=64)
    a = oracle.render(sc, (0.0, 6.0, 1.5), 64, 48, make_params(traversal="reference", **kw), want_aov=True)
    b = hostemu.render(sc, (0.0, 6.0, 1.5), 64, 48, make_params(traversal="fast", **kw), want_aov=True)
    assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"])
    p = make_params(traversal="fast", **kw)
    p._reserved = 1
    with pytest.raises(RuntimeError, match="exceeds the fast traversal stack"):
        hostemu.render(sc, (0.0, 6.0, 1.5), 64, 48, p)


def test_device_build_algorithm(oracle, hostemu, doge):
    """The on-device record build (vcrt_devbuild.cuh: tie ranks by counting, PLOC topology, level-by-level 4-wide collapse) run
    sequentially on the CPU through the same per-element code: every bit equals the reference traversal's, on the bundled
    scene, on tiny trees, on duplicated triangles (ties), on trees with one-child / childless inner nodes, and with forced
    quantisation of a scene 1000x larger than the automatic limit."""
    import tinybvh
    cases = [(doge, CAM, 160, 120, 16)]
    for seed, n in ((51, 1), (52, 2), (53, 3), (54, 5), (55, 40), (56, 700), (57, 4000)):
        cases.append((small_scene(n_tris=n, seed=seed), (0.0, 6.0, 1.5), 96, 64, 16))
    dup = dict(small_scene(n_tris=60, seed=58))
    tri = dup["triangles"].reshape(-1, 48)
    dup["triangles"] = np.concatenate([tri, tri]).reshape(-1).copy()
    dup["bvh"] = tinybvh.build_bvh(dup["triangles"].view(tinybvh.TRI), seed=6, tie_seed=4).view(np.uint8).reshape(-1).copy()
    cases.append((dup, (0.0, 6.0, 1.5), 96, 64, 16))
    odd = dict(small_scene(n_tris=300, seed=59))
    odd["bvh"] = tinybvh.add_degenerate_inner_nodes(odd["bvh"].view(tinybvh.NODE)).view(np.uint8).reshape(-1).copy()
    cases.append((odd, (0.0, 6.0, 1.5), 96, 64, 16))
    big = dict(small_scene(n_tris=800, seed=12))
    t = big["triangles"].copy().view(np.float32).reshape(-1, 12)
    t[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]] *= 1000.0
    big["triangles"] = t.view(np.uint8).reshape(-1)
    big["bvh"] = tinybvh.build_bvh(big["triangles"].view(tinybvh.TRI), seed=2).view(np.uint8).reshape(-1).copy()
    cases.append((big, (0.0, 6000.0, 1500.0), 96, 64, 16 | 4))
    for sc, cam, w, h, reserved in cases:
        kw = dict(shader="full", max_bounces=5, sample_count=2, accum="f32", rng="philox", stack_depth=64)
        a = oracle.render(sc, cam, w, h, make_params(traversal="reference", **kw), want_aov=True)
        p = make_params(traversal="fast", **kw)
        p._reserved = reserved
        b = hostemu.render(sc, cam, w, h, p, want_aov=True)
        assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"]), len(sc["triangles"]) // 48
        assert a["counters"].rays == b["rays"]
    # scenes the device builder declines (the product then falls back to the host builder)
    p = make_params(traversal="fast")
    p._reserved = 16
    with pytest.raises(RuntimeError, match="too large for 15-bit"):
        hostemu.render(big, (0.0, 6000.0, 1500.0), 32, 32, p)
    sc = small_scene(n_tris=8, seed=1)
    nodes = sc["bvh"].view(tinybvh.NODE).copy()
    nodes[1]["left"] = 0     # cycle through the root
    sc["bvh"] = nodes.view(np.uint8).reshape(-1)
    with pytest.raises(RuntimeError, match="not a plain tree"):
        hostemu.render(sc, (0.0, 6.0, 1.5), 32, 32, p)


def test_brute_force_with_spheres(oracle, hostemu):
    sc = small_scene(n_tris=30, seed=4)
    kw = dict(shader="full", traversal="brute_force", max_bounces=4, sample_count=2)
    a = oracle.render(sc, (0.0, 6.0, 1.5), 64, 48, make_params(**kw), want_aov=True)
    b = hostemu.render(sc, (0.0, 6.0, 1.5), 64, 48, make_params(**kw), want_aov=True)
    assert same_bits(a["target"], b["target"]) and same_bits(a["aov"], b["aov"])
    assert (a["aov"]["triangle"] <= -2).sum() > 0   # sphere hits are encoded as -2 - index


def test_repack_rejects_unsupported_trees(hostemu):
    import tinybvh
    sc = small_scene(n_tris=8, seed=1)
    nodes = sc["bvh"].view(tinybvh.NODE).copy()
    nodes[1]["left"] = 0     # cycle
    sc["bvh"] = nodes.view(np.uint8).reshape(-1)
    with pytest.raises(RuntimeError, match="reachable twice"):
        hostemu.render(sc, (0.0, 6.0, 1.5), 32, 32, make_params(traversal="fast"))


@pytest.mark.parametrize("fmt", ["auto", "f32", "q15_forced", "q15x4", "q15x4_forced"])
def test_node_formats_give_identical_results(oracle, hostemu, doge, fmt):
    """32-byte quantised nodes (one 256-bit load per visit) vs 64-byte float nodes: boxes only cull, so every output
    bit is the same; the quantised tree may only visit MORE nodes.  `q15_forced` also runs a scene 1000x larger than the
    quantiser's automatic limit (coarse quanta: correctness must not depend on them).  `q15x4` walks the 4-wide form of the
    quantised tree (four children per visit, vcrt_repack.h: build_wide_bvh): same bits again, and far fewer visits."""
    reserved = {"auto": 0, "f32": 2, "q15_forced": 4, "q15x4": 8, "q15x4_forced": 12}[fmt]
    cases = [(doge, CAM, 160, 120), (small_scene(n_tris=3000, seed=11), (0.0, 6.0, 1.5), 96, 64)]
    if fmt.endswith("_forced"):
        big = small_scene(n_tris=800, seed=12)
        big = dict(big)
        t = big["triangles"].copy().view(np.float32).reshape(-1, 12)
        t[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]] *= 1000.0
        big["triangles"] = t.view(np.uint8).reshape(-1)
        import tinybvh
        big["bvh"] = tinybvh.build_bvh(big["triangles"].view(tinybvh.TRI), seed=2).view(np.uint8).reshape(-1).copy()
        cases.append((big, (0.0, 6000.0, 1500.0), 96, 64))
    for sc, cam, w, h in cases:
        kw = dict(shader="full", max_bounces=5, sample_count=2, accum="f32", rng="philox", stack_depth=64)
        a = oracle.render(sc, cam, w, h, make_params(traversal="reference", **kw), want_aov=True)
        p = make_params(traversal="fast", **kw)
        p._reserved = reserved
        b = hostemu.render(sc, cam, w, h, p, want_aov=True)
        assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"])
        pf = make_params(traversal="fast", **kw)
        pf._reserved = 2
        f = hostemu.render(sc, cam, w, h, pf, want_aov=True)
        assert b["rays"] == f["rays"]
        if fmt.startswith("q15x4"):
            assert b["nodes"] < 0.75 * f["nodes"]          # a 4-wide visit replaces about two binary ones
            continue
        assert b["nodes"] >= f["nodes"]
        if fmt != "f32" and sc is doge:
            assert b["nodes"] <= 1.1 * f["nodes"]        # quantisation costs only a few per cent more visits here


@pytest.mark.parametrize("passes,percent", [(1, 10), (3, 10), (4, 50)])
def test_reinsertion_optimised_tree(oracle, hostemu, passes, percent):
    """optimize_fast_bvh_reinsert (Bittner et al. 2013) moves subtrees about; leaves keep their slots, so every bit of the result must
    stay what the reference traversal gives -- on random scenes with duplicated triangles (ties), on the awkward sizes, and with fewer
    node visits than the tree as built on a scene large enough to have something to optimise."""
    cam = (0.0, 6.0, 1.5)
    for seed, n in ((21, 8), (22, 9), (23, 40), (24, 700), (25, 6000)):
        sc = small_scene(n_tris=n, seed=seed)
        kw = dict(shader="full", max_bounces=6, sample_count=2, accum="f32", rng="philox", stack_depth=64)
        a = oracle.render(sc, cam, 80, 56, make_params(traversal="reference", **kw), want_aov=True)
        for wide in (0, 8):
            p = make_params(traversal="fast", **kw)
            p._reserved = wide | (passes << 8) | (percent << 16)
            b = hostemu.render(sc, cam, 80, 56, p, want_aov=True)
            assert same_bits(a["accumf"], b["accumf"]) and same_bits(a["aov"], b["aov"]), (seed, n, wide)
            assert a["counters"].rays == b["rays"]
    if (passes, percent) == (3, 10):
        from vulkan_compute_ray_tracing_b200 import scenegen
        sc = scenegen.generate_box_scene(60000, seed=5)
        kw = dict(shader="full", max_bounces=4, sample_count=1, accum="f32", rng="philox", stack_depth=64)
        nodes = []
        for res in (8, 8 | (3 << 8) | (10 << 16)):
            p = make_params(traversal="fast", **kw)
            p._reserved = res
            r = hostemu.render(sc, (1.8, 8.6, 1.1), 160, 90, p)
            nodes.append(r["nodes"] / r["rays"])
        assert nodes[1] < nodes[0]
