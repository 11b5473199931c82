"""Test-side scene synthesis: random triangles + a median-split BVH in the reference's layout (numpy, small sizes).
Independent of the product's builder so that parity tests do not depend on it."""
import numpy as np

TRI = np.dtype([("v0", "<f4", 3), ("p0", "<u4"), ("v1", "<f4", 3), ("p1", "<u4"), ("v2", "<f4", 3), ("materialIndex", "<u4")])
MAT = np.dtype([("type", "<u4"), ("p0", "<u4", 3), ("albedo", "<f4", 3), ("p1", "<u4")])
NODE = np.dtype([("min", "<f4", 3), ("p0", "<u4"), ("max", "<f4", 3), ("left", "<i4"), ("right", "<i4"), ("object", "<i4"), ("p1", "<u4", 2)])
LIGHT = np.dtype([("triangleIndex", "<u4"), ("area", "<f4")])
SPHERE = np.dtype([("s", "<f4", 4), ("materialIndex", "<u4"), ("p", "<u4", 3)])
assert TRI.itemsize == 48 and MAT.itemsize == 32 and NODE.itemsize == 48 and LIGHT.itemsize == 8 and SPHERE.itemsize == 32


def build_bvh(tris, seed=0, numbering="reference", tie_seed=None):
    """Median split on a seeded random axis, one triangle per leaf, nodes numbered like Bvh.h:141-209
    (children get consecutive indices when their parent is popped; right subtree processed first).
    tie_seed: triangles with equal sort keys (duplicates) keep a seeded random relative order instead of their index order, so
    that either copy may end up first in the traversal order."""
    rs = np.random.RandomState(seed)
    n = len(tris)
    lo = np.minimum(np.minimum(tris["v0"], tris["v1"]), tris["v2"]) - np.float32(1e-4)
    hi = np.maximum(np.maximum(tris["v0"], tris["v1"]), tris["v2"]) + np.float32(1e-4)
    nodes = np.zeros(max(2 * n - 1, 1), NODE)
    nodes["left"] = nodes["right"] = nodes["object"] = -1
    if n == 0:
        return nodes[:0]
    counter = 1
    stack = [(0, np.arange(n) if tie_seed is None else np.random.RandomState(tie_seed).permutation(n))]
    while stack:
        idx, ids = stack.pop()
        nodes[idx]["min"] = lo[ids].min(axis=0)
        nodes[idx]["max"] = hi[ids].max(axis=0)
        axis = rs.randint(3)
        if len(ids) == 1:
            nodes[idx]["object"] = ids[0]
            continue
        order = ids[np.argsort(lo[ids, axis], kind="stable")]
        mid = len(order) // 2
        nodes[idx]["left"], nodes[idx]["right"] = counter, counter + 1
        stack.append((counter, order[:mid]))
        stack.append((counter + 1, order[mid:]))
        counter += 2
    return nodes[:counter]


def build_scene(n_tris=64, seed=7, glass=True, metal=True, extent=1.0):
    rs = np.random.RandomState(seed)
    mats = np.zeros(6, MAT)
    mats["type"] = [1, 1, 1, 0, 2, 3]
    mats["albedo"] = [(.3, .3, .3), (.9, .1, .1), (.1, .9, .1), (2, 2, 2), (1, 1, 1), (1, 1, 1)]
    tris = np.zeros(n_tris + 4, TRI)
    c = rs.uniform(-extent, extent, (n_tris, 3)) + np.array([0, 1.5, -1.5])
    for k in ("v0", "v1", "v2"):
        tris[k][:n_tris] = (c + rs.uniform(-0.35, 0.35, (n_tris, 3))).astype(np.float32)
    choices = [0, 1, 2] + ([4] if metal else []) + ([5] if glass else [])
    tris["materialIndex"][:n_tris] = rs.choice(choices, n_tris)
    # floor quad + horizontal emitter quad
    f = [(-3, -0.2, -4), (3, -0.2, -4), (3, -0.2, 1), (-3, -0.2, 1)]
    e = [(-0.8, 3.2, -2.2), (0.8, 3.2, -2.2), (0.8, 3.2, -0.8), (-0.8, 3.2, -0.8)]
    for base, q, m in ((n_tris, f, 0), (n_tris + 2, e, 3)):
        tris[base]["v0"], tris[base]["v1"], tris[base]["v2"] = q[0], q[1], q[2]
        tris[base + 1]["v0"], tris[base + 1]["v1"], tris[base + 1]["v2"] = q[0], q[2], q[3]
        tris["materialIndex"][base:base + 2] = m
    lights = np.zeros(2, LIGHT)
    lights["triangleIndex"] = [n_tris + 2, n_tris + 3]
    lights["area"] = 1.0
    spheres = np.zeros(1, SPHERE)
    spheres["s"] = (0.6, 1.0, -1.0, 0.6)
    spheres["materialIndex"] = 5
    nodes = build_bvh(tris, seed)
    return {"triangles": tris.view(np.uint8).reshape(-1).copy(), "materials": mats.view(np.uint8).reshape(-1).copy(),
            "bvh": nodes.view(np.uint8).reshape(-1).copy(), "lights": lights.view(np.uint8).reshape(-1).copy(),
            "spheres": spheres.view(np.uint8).reshape(-1).copy()}


def add_degenerate_inner_nodes(nodes, every=3):
    """Returns a tree with the same leaves in the same visiting order, where every `every`-th leaf is wrapped in an inner node
    that has ONE child (the leaf, on alternating sides) and every other `every`-th one gets a sibling with neither a triangle
    nor children.  The reference shader walks such trees without noticing (-1 children are skipped,
    ray-trace-compute.comp:279-281; a node without an object only forwards its children, :290-306); a repacked traversal
    has to represent the absent slots explicitly."""
    nodes = nodes.copy()
    extra = []
    base = len(nodes)
    k = 0
    for i in range(base):
        if nodes[i]["object"] < 0:
            continue
        k += 1
        if k % every == 0:          # leaf i becomes an inner node with one child: the leaf, moved to a new slot
            leaf = nodes[i].copy()
            idx = base + len(extra)
            extra.append(leaf)
            nodes[i]["object"] = -1
            if (k // every) % 2:
                nodes[i]["left"], nodes[i]["right"] = idx, -1
            else:
                nodes[i]["left"], nodes[i]["right"] = -1, idx
        elif k % every == 1:        # leaf i becomes an inner node {leaf, childless node}
            leaf = nodes[i].copy()
            hollow = nodes[i].copy()
            hollow["object"] = -1
            idx = base + len(extra)
            extra.extend([leaf, hollow])
            nodes[i]["object"] = -1
            nodes[i]["left"], nodes[i]["right"] = idx, idx + 1
    if extra:
        nodes = np.concatenate([nodes, np.array(extra, NODE)])
    return nodes
