// vcrt_devbuild.cuh -- traversal records built ON THE DEVICE from the bound bvh[] / triangles[] (the producer
// side of Bvh::createBvh, Bvh.h:141-209): what vcrt_repack.cpp does on the host in 0.4 s (1 M triangles) .. 4 s (10 M), as a few
// dozen kernel launches that take milliseconds, so that a scene change costs about as much as a frame.
//
// Same contract as the host builder (vcrt_repack.h): triangle slot = the reference's tie rank (position of the leaf in
// hit_bvh's right-child-first visiting order, ray-trace-compute.comp:301-306), boxes only cull and are rounded outwards, so
// RESULTS ARE IDENTICAL whatever tree is built; only the number of visits per ray depends on the topology.
//
// Steps (each a data-parallel pass; the per-element bodies below are `__host__ __device__` so that tests/hostemu can run the very
// same code sequentially on the CPU against the oracle):
//   A  tie ranks.  parent links (atomicCAS detects shared subtrees / cycles), leaves per subtree bottom-up (arrival counters),
//      then every leaf walks to the root adding count[right sibling] whenever it sits in a left subtree: its rank in the
//      right-first order.  No DFS, no recursion.
//   B  64-byte triangle records in slot order (a = v0 - v1, b = v2 - v0, n = cross(b, a) with the shader's own fp32 operations,
//      cf. precompute_triangles), padded leaf boxes, 63-bit Morton codes of the centroids, radix sort.
//   C  topology by PLOC (parallel locally-ordered clustering, Meister & Bittner 2018): clusters in Morton order look for the
//      neighbour within +-R positions that minimises the surface area of the union; mutual nearest neighbours merge; repeat
//      until one cluster is left.  Surface-area cost within 1 % of the host's binned-SAH tree on the C3 scene (55.6 vs 55.2),
//      where a plain LBVH is 28 % worse.  Node and cluster indices come from prefix sums, so the layout is deterministic.
//   D  4-wide collapse (open the inner child with the largest area until four children, as build_wide_bvh) level by level,
//      15-bit outward quantisation in the scene-wide frame, exact traversal-stack bound.
#pragma once

#include "vcrt_core.cuh"

namespace vcrt {
namespace devbuild {

#define VCRT_DB_EMPTY ((int32_t)0x80000000)
#ifndef VCRT_PLOC_RADIUS
#define VCRT_PLOC_RADIUS 4   /* clusters look this far to either side for their nearest neighbour.  4-wide visits / triangle tests per ray on the host emulation (480x270, depth 8), radius 2 | 4 | 8 | 16, host SAH tree last: 1 M-triangle box 12.00/3.05 | 12.12/3.07 | 12.53/3.13 | 12.60/3.13 | 11.73/3.11; 100 K box 10.24/2.90 | 10.12/2.92 | 10.46/2.99 | - | 9.78/3.09; random 20 K-triangle soup 41.85/20.55 | 39.30/20.10 | 38.95/20.04 | - | 39.72/19.85 */
#endif

// ---- atomics that degrade to plain operations in the sequential host emulation
VCRT_HD uint32_t atom_add(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    const uint32_t o = *p; *p += v; return o;
#endif
}
VCRT_HD int32_t atom_cas(int32_t* p, int32_t cmp, int32_t val) {
#ifdef __CUDA_ARCH__
    return atomicCAS(p, cmp, val);
#else
    const int32_t o = *p; if (o == cmp) *p = val; return o;
#endif
}
VCRT_HD void atom_max(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
VCRT_HD void atom_min(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
VCRT_HD void fence() {
#ifdef __CUDA_ARCH__
    __threadfence();
#endif
}
VCRT_HD uint32_t load_relaxed(const uint32_t* p) {
#ifdef __CUDA_ARCH__
    return *(const volatile uint32_t*)p;
#else
    return *p;
#endif
}
// floats as order-preserving unsigned keys (for atomicMin/Max on bounds)
VCRT_HD uint32_t f2ord(float f) { const uint32_t u = f2u(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
VCRT_HD float ord2f(uint32_t k) { return u2f((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// status words (device memory, zero-initialised)
enum { ST_ERROR = 0, ST_MAXDEPTH = 1, ST_STACK = 2, ST_CENTROID = 4 /* 6 words: lo.xyz hi.xyz as ordered keys */, ST_WORDS = 16 };
enum { ERR_SHARED = 1u, ERR_LEAF_CHILDREN = 2u, ERR_DEEP = 4u, ERR_NONFINITE = 8u };

struct View {
    const vcrt_bvh_node* bvh; uint32_t nbvh;
    const vcrt_triangle* tris; uint32_t ntris;
    int32_t* parent;        // [nbvh]
    uint32_t* count;        // [nbvh] reachable-or-not leaves below
    uint32_t* arrive;       // [nbvh]
    uint32_t* leaf_node;    // [nleaves] slot -> bvh node
    uint32_t* status;       // [ST_WORDS]
};

VCRT_HD bool in_range(int32_t c, uint32_t n) { return c >= 0 && (uint32_t)c < n; }
VCRT_HD bool is_leaf(const vcrt_bvh_node& nd) { return nd.objectIndex != -1; }

// A2: parent links
VCRT_HD void link_children(const View& v, uint32_t i) {
    const vcrt_bvh_node nd = v.bvh[i];
    const bool has_children = nd.leftNodeIndex != -1 || nd.rightNodeIndex != -1;
    if (is_leaf(nd)) { if (has_children) atom_max(v.status + ST_ERROR, ERR_LEAF_CHILDREN); return; }
    const int32_t ch[2] = {nd.leftNodeIndex, nd.rightNodeIndex};
    for (int s = 0; s < 2; ++s) {
        if (!in_range(ch[s], v.nbvh)) continue;
        if (ch[s] == 0 || atom_cas(v.parent + ch[s], -1, (int32_t)i) != -1) atom_max(v.status + ST_ERROR, ERR_SHARED);
    }
}

// A3: leaves per subtree.  Sources are the nodes nothing reports to: leaves, and nodes without any valid child.
VCRT_HD void count_up(const View& v, uint32_t i) {
    const vcrt_bvh_node nd = v.bvh[i];
    const int valid = is_leaf(nd) ? 0 : (int)in_range(nd.leftNodeIndex, v.nbvh) + (int)in_range(nd.rightNodeIndex, v.nbvh);
    if (valid != 0) return;
    uint32_t mine = is_leaf(nd) ? 1u : 0u;
    v.count[i] = mine;
    uint32_t cur = i;
    for (int guard = 0; guard < 4096; ++guard) {
        const int32_t p = v.parent[cur];
        if (p < 0) return;
        const vcrt_bvh_node pn = v.bvh[p];
        const uint32_t need = (uint32_t)in_range(pn.leftNodeIndex, v.nbvh) + (uint32_t)in_range(pn.rightNodeIndex, v.nbvh);
        atom_add(v.count + p, mine);
        fence();
        if (atom_add(v.arrive + p, 1u) + 1u < need) return;    // a sibling subtree is still counting: it will carry on
        fence();
        mine = load_relaxed(v.count + p);
        cur = (uint32_t)p;
    }
    atom_max(v.status + ST_ERROR, ERR_DEEP);
}

// A4: rank of leaf node i in the reference's right-child-first visiting order; false when the leaf is not reachable from node 0
VCRT_HD bool leaf_rank(const View& v, uint32_t i, uint32_t& rank, uint32_t& depth) {
    rank = 0u; depth = 0u;
    uint32_t cur = i;
    for (int guard = 0; guard < 4096; ++guard) {
        const int32_t p = v.parent[cur];
        if (p < 0) return cur == 0u;
        const vcrt_bvh_node pn = v.bvh[p];
        if (pn.leftNodeIndex == (int32_t)cur && in_range(pn.rightNodeIndex, v.nbvh)) rank += load_relaxed(v.count + pn.rightNodeIndex);   // the right subtree is visited first
        cur = (uint32_t)p;
        ++depth;
    }
    atom_max(v.status + ST_ERROR, ERR_DEEP);
    return false;
}

struct Box6 { float lo[3], hi[3]; };
VCRT_HD float half_area(const Box6& b) { const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2]; return dx * dy + dy * dz + dz * dx; }
VCRT_HD Box6 box_union(const Box6& a, const Box6& b) {
    Box6 r;
    for (int k = 0; k < 3; ++k) { r.lo[k] = fminf(a.lo[k], b.lo[k]); r.hi[k] = fmaxf(a.hi[k], b.hi[k]); }
    return r;
}

// B: the 64-byte triangle record of slot s (layout: vcrt_fast.cuh) and its padded box (the reference's leaf padding, Bvh.h:16)
VCRT_HD void make_slot(const View& v, uint32_t slot, float4* tris64, float4* cl_lo, float4* cl_hi) {
    const uint32_t leaf = v.leaf_node[slot];
    vcrt_bvh_node nd;
    nd.objectIndex = -1;
    if (leaf < v.nbvh) nd = v.bvh[leaf]; else atom_max(v.status + ST_ERROR, ERR_SHARED);     // a rank nobody claimed: the counts were inconsistent
    vcrt_triangle t;
    for (int k = 0; k < 3; ++k) t.v0[k] = t.v1[k] = t.v2[k] = 0.0f;
    t.materialIndex = 0u;
    if (nd.objectIndex >= 0 && (uint32_t)nd.objectIndex < v.ntris) t = v.tris[nd.objectIndex];      // out of range -> the zero triangle (never hit), as robust reads give
    // every operation rounded to fp32 on its own (this translation unit is compiled with -fmad=false)
    const float3 v0 = f3(t.v0[0], t.v0[1], t.v0[2]), v1 = f3(t.v1[0], t.v1[1], t.v1[2]), v2 = f3(t.v2[0], t.v2[1], t.v2[2]);
    const float3 a = sub(v0, v1), b = sub(v2, v0), n = cross(b, a);
    tris64[4 * (size_t)slot + 0] = make_float4(v0.x, v0.y, v0.z, u2f((uint32_t)nd.objectIndex));
    tris64[4 * (size_t)slot + 1] = make_float4(a.x, a.y, a.z, u2f(t.materialIndex));
    tris64[4 * (size_t)slot + 2] = make_float4(b.x, b.y, b.z, 0.0f);
    tris64[4 * (size_t)slot + 3] = make_float4(n.x, n.y, n.z, 0.0f);
    const float eps = 0.0001f;
    float lo[3], hi[3];
    for (int k = 0; k < 3; ++k) {
        const float l = fminf(fminf(t.v0[k], t.v1[k]), t.v2[k]), h = fmaxf(fmaxf(t.v0[k], t.v1[k]), t.v2[k]);
        lo[k] = l - eps; hi[k] = h + eps;
        if (!(l - l == 0.0f) || !(h - h == 0.0f)) atom_max(v.status + ST_ERROR, ERR_NONFINITE);
        const float c = 0.5f * (l + h);
        atom_min(v.status + ST_CENTROID + k, f2ord(c));
        atom_max(v.status + ST_CENTROID + 3 + k, f2ord(c));
    }
    cl_lo[slot] = make_float4(lo[0], lo[1], lo[2], u2f((uint32_t)~(int32_t)slot));   // w = code: leaf ~slot
    cl_hi[slot] = make_float4(hi[0], hi[1], hi[2], u2f(0u));                         // w = depth of the cluster's subtree
}

VCRT_HD uint64_t spread21(uint64_t x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
VCRT_HD uint64_t morton_of(const float4 lo, const float4 hi, const float clo[3], const float chi[3]) {
    const float c[3] = {0.5f * ((lo.x + 0.0001f) + (hi.x - 0.0001f)), 0.5f * ((lo.y + 0.0001f) + (hi.y - 0.0001f)), 0.5f * ((lo.z + 0.0001f) + (hi.z - 0.0001f))};
    uint64_t q[3];
    for (int k = 0; k < 3; ++k) {
        const float e = chi[k] - clo[k];
        float u = e > 0.0f ? (c[k] - clo[k]) / e : 0.0f;
        u = fminf(fmaxf(u * 2097152.0f, 0.0f), 2097151.0f);
        q[k] = (uint64_t)u;
    }
    return spread21(q[0]) << 2 | spread21(q[1]) << 1 | spread21(q[2]);
}

// C: nearest neighbour of cluster i among positions i-R .. i+R (smallest union area; ties: the lower position)
VCRT_HD uint32_t nearest(const float4* lo, const float4* hi, uint32_t m, uint32_t i) {
    Box6 me = {{lo[i].x, lo[i].y, lo[i].z}, {hi[i].x, hi[i].y, hi[i].z}};
    const uint32_t b = i > VCRT_PLOC_RADIUS ? i - VCRT_PLOC_RADIUS : 0u, e = i + VCRT_PLOC_RADIUS < m - 1u ? i + VCRT_PLOC_RADIUS : m - 1u;
    float best = u2f(0x7f800000u);
    uint32_t bj = i;
    for (uint32_t j = b; j <= e; ++j) {
        if (j == i) continue;
        const Box6 o = {{lo[j].x, lo[j].y, lo[j].z}, {hi[j].x, hi[j].y, hi[j].z}};
        const float a = half_area(box_union(me, o));
        if (a < best || bj == i) { best = a; bj = j; }     // `bj == i`: the first candidate, whatever its area (NaN-proof)
    }
    return bj;
}
// flags of position i: bit 0 = survives into the next round (as itself or as the merged cluster), bit 32 = creates a node
VCRT_HD uint64_t merge_flags(const uint32_t* nn, uint32_t i) {
    const uint32_t j = nn[i];
    const bool mutual = j != i && nn[j] == i;
    if (!mutual) return 1ull;
    return i < j ? (1ull | 1ull << 32) : 0ull;
}
// writes position i's contribution: the merged node + cluster (lower partner), nothing (upper partner), or a copy
VCRT_HD void merge_write(const float4* lo, const float4* hi, const uint32_t* nn, const uint64_t* scan, uint32_t i, uint32_t node_base,
                         float4* out_lo, float4* out_hi, float* nodes) {
    const uint64_t f = merge_flags(nn, i);
    if (!(f & 1ull)) return;
    const uint32_t pos = (uint32_t)(scan[i] & 0xffffffffull);
    if (!(f >> 32)) { out_lo[pos] = lo[i]; out_hi[pos] = hi[i]; return; }
    const uint32_t j = nn[i];
    const uint32_t node = node_base + (uint32_t)(scan[i] >> 32);
    const float4 al = lo[i], ah = hi[i], bl = lo[j], bh = hi[j];
    float* p = nodes + (size_t)node * 16;      // binary node, float layout of vcrt_fast.cuh
    p[0] = al.x; p[1] = ah.x; p[2] = al.y; p[3] = ah.y; p[8] = al.z; p[9] = ah.z;
    p[4] = bl.x; p[5] = bh.x; p[6] = bl.y; p[7] = bh.y; p[10] = bl.z; p[11] = bh.z;
    p[12] = al.w; p[13] = bl.w; p[14] = 0.0f; p[15] = 0.0f;
    const uint32_t d = (f2u(ah.w) > f2u(bh.w) ? f2u(ah.w) : f2u(bh.w)) + 1u;
    out_lo[pos] = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), u2f(node));
    out_hi[pos] = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), u2f(d));
}

// D: 4-wide collapse.  One work item = one 4-wide node to be written: {binary node, stack entries above it}.
struct WideItem { int32_t node2; uint32_t stack_above; };
struct WideKids { int32_t code[4]; Box6 box[4]; int n; };

VCRT_HD void binary_children(const float* nodes, int32_t node, Box6 out_box[2], int32_t out_code[2]) {
    const float* p = nodes + (size_t)node * 16;
    out_box[0] = {{p[0], p[2], p[8]}, {p[1], p[3], p[9]}};
    out_box[1] = {{p[4], p[6], p[10]}, {p[5], p[7], p[11]}};
    out_code[0] = (int32_t)f2u(p[12]); out_code[1] = (int32_t)f2u(p[13]);
}
// the (up to four) children of the 4-wide node rooted at binary node `node`: open the inner child with the largest area
VCRT_HD void wide_children(const float* nodes, int32_t node, WideKids& k) {
    Box6 b[2]; int32_t c[2];
    binary_children(nodes, node, b, c);
    k.n = 2; k.box[0] = b[0]; k.box[1] = b[1]; k.code[0] = c[0]; k.code[1] = c[1];
    while (k.n < 4) {
        int best = -1; float ba = -1.0f;
        for (int i = 0; i < k.n; ++i) if (k.code[i] >= 0 && half_area(k.box[i]) > ba) { ba = half_area(k.box[i]); best = i; }
        if (best < 0) break;
        binary_children(nodes, k.code[best], b, c);
        k.box[best] = b[0]; k.code[best] = c[0];
        k.box[k.n] = b[1]; k.code[k.n] = c[1];
        ++k.n;
    }
}
VCRT_HD uint32_t wide_inner_count(const float* nodes, const WideItem& it) {
    WideKids k;
    wide_children(nodes, it.node2, k);
    uint32_t n = 0;
    for (int i = 0; i < k.n; ++i) n += k.code[i] >= 0 ? 1u : 0u;
    return n;
}

struct QFrame { float org[3], ext[3]; };
// one axis of one box in the frame: {lo | hi << 16}, rounded outwards against the floats the kernel decodes with (cf. quant_bounds)
VCRT_HD uint32_t quant_axis(const QFrame& q, int a, float mnf, float mxf) {
    const double mn = mnf, mx = mxf, org = q.org[a], ext = q.ext[a];
    double l = floor((mn - org - ext) / ext * 32768.0), h = ceil((mx - org - ext) / ext * 32768.0);
    while (l > 0.0 && org + (1.0 + l / 32768.0) * ext > mn) l -= 1.0;
    while (h < 32767.0 && org + (1.0 + h / 32768.0) * ext < mx) h += 1.0;
    l = l < 0.0 ? 0.0 : (l > 32767.0 ? 32767.0 : l);
    h = h < 0.0 ? 0.0 : (h > 32767.0 ? 32767.0 : h);
    return (uint32_t)l | ((uint32_t)h << 16);
}
// writes 4-wide node `node4` for work item `it`; its inner children get the consecutive indices child_base, child_base + 1, ...
// and are appended to the next level's work list at next_pos, next_pos + 1, ...
VCRT_HD void wide_write(const float* nodes, const QFrame& q, const WideItem& it, uint32_t node4, uint32_t child_base, uint32_t next_pos, WideItem* next,
                        uint32_t* q4nodes, uint32_t* status) {
    WideKids k;
    wide_children(nodes, it.node2, k);
    const uint32_t below = it.stack_above + (uint32_t)(k.n - 1);     // a visit leaves at most n - 1 entries behind and descends into one child
    atom_max(status + ST_STACK, below);
    uint32_t w[16];
    uint32_t inner = 0;
    for (int i = 0; i < 4; ++i) {
        uint32_t* half = w + (i / 2) * 8;
        const int s = i & 1;
        int32_t code = VCRT_DB_EMPTY;
        for (int a = 0; a < 3; ++a) half[s * 3 + a] = 32767u;          // the empty box: min > max on every axis
        if (i < k.n) {
            for (int a = 0; a < 3; ++a) half[s * 3 + a] = quant_axis(q, a, k.box[i].lo[a], k.box[i].hi[a]);
            code = k.code[i];
            if (code >= 0) {
                next[next_pos + inner].node2 = code;
                next[next_pos + inner].stack_above = below;
                code = (int32_t)(child_base + inner);
                ++inner;
            }
        }
        half[6 + s] = (uint32_t)code;
    }
    uint32_t* o = q4nodes + (size_t)node4 * 16;
    for (int i = 0; i < 16; ++i) o[i] = w[i];
}

}  // namespace devbuild
}  // namespace vcrt
