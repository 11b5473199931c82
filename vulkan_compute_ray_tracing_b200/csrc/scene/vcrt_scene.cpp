// vcrt_scene.cpp -- host-side producers of the hot path's inputs (C ABI: include/vcrt_scene.h).
#include "../../../include/vcrt_scene.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

// ---------------------------------------------------------------------------------------------- glibc rand()
// The reference never seeds rand() (Bvh.h:42-45), so its axis sequence is glibc's TYPE_3 generator after srand(1):
// r[i] = r[i-3] + r[i-31] over a 34-word table initialised by the minimal-standard LCG, first 310 outputs dropped.
struct GlibcRand {
    uint32_t r[34];
    int pos = 0;
    std::vector<uint32_t> hist;
    explicit GlibcRand(uint32_t seed) {
        if (seed == 0) seed = 1;
        hist.resize(344);
        int32_t word = (int32_t)seed;
        hist[0] = (uint32_t)word;
        for (int i = 1; i < 31; ++i) {
            long hi = word / 127773, lo = word % 127773;
            word = (int32_t)(16807 * lo - 2836 * hi);
            if (word < 0) word += 2147483647;
            hist[i] = (uint32_t)word;
        }
        for (int i = 31; i < 34; ++i) hist[i] = hist[i - 31];
        for (int i = 34; i < 344; ++i) hist[i] = hist[i - 31] + hist[i - 3];
    }
    int32_t next() {
        size_t i = hist.size();
        uint32_t v = hist[i - 31] + hist[i - 3];
        hist.push_back(v);
        if (hist.size() > 4096) hist.erase(hist.begin(), hist.begin() + 2048);
        return (int32_t)(v >> 1);
    }
};

struct Key { float k[3]; uint32_t idx; };

struct Plan { uint32_t begin, count; int32_t left, right; uint8_t axis; };

inline void tri_box(const vcrt_triangle& t, float lo[3], float hi[3]) {
    const float eps = 0.0001f;   // Bvh.h:16
    for (int a = 0; a < 3; ++a) {
        lo[a] = std::min(std::min(t.v0[a], t.v1[a]), t.v2[a]) - eps;
        hi[a] = std::max(std::max(t.v0[a], t.v1[a]), t.v2[a]) + eps;
    }
}

void build_subtree(const std::vector<Plan>& plan, std::vector<Key>& keys, const vcrt_triangle* tris, vcrt_bvh_node* nodes, int32_t idx) {
    const Plan& p = plan[idx];
    vcrt_bvh_node& nd = nodes[idx];
    std::memset(&nd, 0, sizeof nd);
    nd.leftNodeIndex = nd.rightNodeIndex = nd.objectIndex = -1;
    const int axis = p.axis;
    // Bvh.h:166 sorts every popped node's list (leaves included; a 1-element sort is a no-op)
    std::sort(keys.begin() + p.begin, keys.begin() + p.begin + p.count, [axis](const Key& a, const Key& b) { return a.k[axis] < b.k[axis]; });
    if (p.count <= 1) {
        const uint32_t ti = keys[p.begin].idx;
        tri_box(tris[ti], nd.min, nd.max);
        nd.objectIndex = (int32_t)ti;
        return;
    }
    if (p.count > 4096) {
#pragma omp task shared(plan, keys) firstprivate(tris, nodes)
        build_subtree(plan, keys, tris, nodes, p.left);
#pragma omp task shared(plan, keys) firstprivate(tris, nodes)
        build_subtree(plan, keys, tris, nodes, p.right);
#pragma omp taskwait
    } else {
        build_subtree(plan, keys, tris, nodes, p.left);
        build_subtree(plan, keys, tris, nodes, p.right);
    }
    // objectListBoundingBox (Bvh.h:100-114) is a min/max fold: exact, hence equal to the union of the children
    for (int a = 0; a < 3; ++a) {
        nd.min[a] = std::min(nodes[p.left].min[a], nodes[p.right].min[a]);
        nd.max[a] = std::max(nodes[p.left].max[a], nodes[p.right].max[a]);
    }
    nd.leftNodeIndex = p.left;
    nd.rightNodeIndex = p.right;
}

// ---------------------------------------------------------------------------------------------- synthetic scenes
struct SplitMix {
    uint64_t s;
    explicit SplitMix(uint64_t seed) : s(seed) {}
    uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    float uni() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }
    float range(float a, float b) { return a + (b - a) * uni(); }
};

struct V { float x, y, z; };

inline void put_tri(vcrt_triangle* out, uint32_t& n, uint32_t cap, V a, V b, V c, uint32_t mat) {
    if (n >= cap) return;
    vcrt_triangle t;
    std::memset(&t, 0, sizeof t);
    t.v0[0] = a.x; t.v0[1] = a.y; t.v0[2] = a.z;
    t.v1[0] = b.x; t.v1[1] = b.y; t.v1[2] = b.z;
    t.v2[0] = c.x; t.v2[1] = c.y; t.v2[2] = c.z;
    t.materialIndex = mat;
    out[n++] = t;
}

inline void put_quad(vcrt_triangle* out, uint32_t& n, uint32_t cap, V a, V b, V c, V d, uint32_t mat) {
    put_tri(out, n, cap, a, b, c, mat);
    put_tri(out, n, cap, a, c, d, mat);
}

}  // namespace

extern "C" {

const char* vcrt_scene_last_error(void) { return g_err.c_str(); }

void vcrt_scene_glibc_rand(uint32_t seed, uint32_t n, int32_t* out) {
    GlibcRand g(seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = g.next();
}

int vcrt_scene_build_bvh(const vcrt_triangle* tris, uint32_t n, uint32_t axis_seed, vcrt_bvh_node* nodes, uint32_t* num_nodes) {
    if (num_nodes) *num_nodes = 0;
    if (!tris || !nodes || !num_nodes) { g_err = "vcrt_scene_build_bvh: NULL argument"; return VCRT_ERR_INVALID; }
    if (n == 0) { g_err = "vcrt_scene_build_bvh: no triangles (the reference dereferences objects[0] of an empty root)"; return VCRT_ERR_INVALID; }
    if (n > 0x3fffffffu) { g_err = "vcrt_scene_build_bvh: too many triangles"; return VCRT_ERR_INVALID; }
    // 1. the tree shape, node numbering and axis choices depend on n only: replay the reference's stack over sizes
    std::vector<Plan> plan((size_t)2 * n - 1);
    {
        GlibcRand rng(axis_seed ? axis_seed : 1u);
        std::vector<int32_t> stack;
        stack.push_back(0);
        plan[0] = {0u, n, -1, -1, 0};
        int32_t counter = 1;
        while (!stack.empty()) {
            const int32_t idx = stack.back();
            stack.pop_back();
            Plan& p = plan[idx];
            p.axis = (uint8_t)(rng.next() % 3);      // Bvh.h:160: drawn for every popped node, leaves included
            if (p.count <= 1) continue;
            const uint32_t mid = p.count / 2;        // Bvh.h:175
            p.left = counter++;
            p.right = counter++;
            plan[p.left] = {p.begin, mid, -1, -1, 0};
            plan[p.right] = {p.begin + mid, p.count - mid, -1, -1, 0};
            stack.push_back(p.left);                 // left pushed first, right popped first (Bvh.h:177-193)
            stack.push_back(p.right);
        }
    }
    // 2. geometry: per-node sort by padded-box minimum on the node's axis, in parallel over subtrees
    std::vector<Key> keys(n);
    for (uint32_t i = 0; i < n; ++i) {
        float lo[3], hi[3];
        tri_box(tris[i], lo, hi);
        keys[i] = {{lo[0], lo[1], lo[2]}, i};
    }
#pragma omp parallel
#pragma omp single
    build_subtree(plan, keys, tris, nodes, 0);
    *num_nodes = 2 * n - 1;
    return VCRT_OK;
}

uint32_t vcrt_scene_collect_lights(const vcrt_triangle* tris, uint32_t n, const vcrt_material* mats, uint32_t nmats, vcrt_light* lights) {
    uint32_t k = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const vcrt_triangle& t = tris[i];
        if (t.materialIndex < nmats && mats[t.materialIndex].type == VCRT_MAT_LIGHT) {
            if (lights) {
                // glm::length(glm::cross(t.v0, t.v1)) * 0.5f  (RtScene.h:93)
                const float cx = t.v0[1] * t.v1[2] - t.v1[1] * t.v0[2], cy = t.v0[2] * t.v1[0] - t.v1[2] * t.v0[0], cz = t.v0[0] * t.v1[1] - t.v1[0] * t.v0[1];
                lights[k].triangleIndex = i;
                lights[k].area = std::sqrt(cx * cx + cy * cy + cz * cz) * 0.5f;
            }
            ++k;
        }
    }
    return k;
}

uint32_t vcrt_scene_generate_box(uint32_t target, uint32_t seed, vcrt_triangle* out, uint32_t cap, vcrt_material* mats, uint32_t max_mats, uint32_t* num_mats) {
    if (!out || !mats || !num_mats || max_mats < 8) { g_err = "vcrt_scene_generate_box: bad arguments"; return 0; }
    SplitMix rng(0x5eedull * 0x100000001b3ull + seed);
    // materials: the reference's table head (RtScene.h:48-60: gray, red, green, light) then random Lambertians
    const uint32_t nlamb = std::min<uint32_t>(32u, max_mats - 4u);
    std::memset(mats, 0, sizeof(vcrt_material) * (4 + nlamb));
    const float head[4][3] = {{.3f, .3f, .3f}, {.9f, .1f, .1f}, {.1f, .9f, .1f}, {2.f, 2.f, 2.f}};
    for (int i = 0; i < 4; ++i) { mats[i].type = i == 3 ? VCRT_MAT_LIGHT : VCRT_MAT_LAMBERTIAN; std::memcpy(mats[i].albedo, head[i], 12); }
    for (uint32_t i = 0; i < nlamb; ++i) { mats[4 + i].type = VCRT_MAT_LAMBERTIAN; for (int c = 0; c < 3; ++c) mats[4 + i].albedo[c] = rng.range(0.2f, 0.9f); }
    *num_mats = 4 + nlamb;

    // the bundled scene's box (scene AABB (-1.574,-0.062,-2.999)..(1.574,3.086,0.149)), open towards the camera (+z)
    const float x0 = -1.574f, x1 = 1.574f, y0 = -0.062f, y1 = 3.086f, z0 = -2.999f, z1 = 0.149f;
    uint32_t n = 0;
    const uint32_t capn = std::min(cap, target);
    put_quad(out, n, capn, {x1, y0, z0}, {x1, y1, z0}, {x1, y1, z1}, {x1, y0, z1}, 1);   // right, red
    put_quad(out, n, capn, {x0, y0, z0}, {x0, y0, z1}, {x0, y1, z1}, {x0, y1, z0}, 2);   // left, green
    put_quad(out, n, capn, {x0, y0, z0}, {x0, y1, z0}, {x1, y1, z0}, {x1, y0, z0}, 0);   // back
    put_quad(out, n, capn, {x0, y1, z0}, {x0, y1, z1}, {x1, y1, z1}, {x1, y1, z0}, 0);   // ceiling
    put_quad(out, n, capn, {x0, y0, z0}, {x1, y0, z0}, {x1, y0, z1}, {x0, y0, z1}, 0);   // floor
    put_quad(out, n, capn, {-0.6f, y1 - 0.01f, -2.0f}, {-0.6f, y1 - 0.01f, -0.9f}, {0.6f, y1 - 0.01f, -0.9f}, {0.6f, y1 - 0.01f, -2.0f}, 3);  // emitter
    if (target <= n) return n;
    const uint32_t budget = capn - n;

    // terrain: displaced grid over the floor, ~40 % of the budget
    const uint32_t G = (uint32_t)std::floor(std::sqrt(0.4 * budget / 2.0));
    const float ph[6] = {rng.range(0, 6.28f), rng.range(0, 6.28f), rng.range(0, 6.28f), rng.range(0, 6.28f), rng.range(0, 6.28f), rng.range(0, 6.28f)};
    auto height = [&](float u, float v) {
        return y0 + 0.02f + 0.10f * (1.0f + std::sin(9.0f * u + ph[0]) * std::cos(7.0f * v + ph[1])) + 0.04f * std::sin(31.0f * u + ph[2]) * std::sin(29.0f * v + ph[3]) +
               0.015f * std::sin(97.0f * u + ph[4]) * std::cos(101.0f * v + ph[5]);
    };
    for (uint32_t j = 0; j < G; ++j)
        for (uint32_t i = 0; i < G; ++i) {
            const float u0 = (float)i / G, u1 = (float)(i + 1) / G, v0 = (float)j / G, v1 = (float)(j + 1) / G;
            auto P = [&](float u, float v) { return V{x0 + (x1 - x0) * u, height(u, v), z0 + (z1 - z0) * v}; };
            const uint32_t m = 4 + (uint32_t)(((i * 8) / std::max(G, 1u)) + 8 * ((j * 4) / std::max(G, 1u))) % nlamb;
            put_quad(out, n, capn, P(u0, v0), P(u0, v1), P(u1, v1), P(u1, v0), m);
        }
    // blobs: displaced tessellated spheres share what is left
    const uint32_t B = 8;
    for (uint32_t b = 0; b < B && n < capn; ++b) {
        const uint32_t share = (capn - n) / (B - b);
        const uint32_t R = std::max<uint32_t>(2u, (uint32_t)std::floor(std::sqrt(share / 4.0)));
        const uint32_t S = 2 * R;
        const float rad = rng.range(0.28f, 0.5f);
        const V c = {rng.range(x0 + 0.6f, x1 - 0.6f), rng.range(y0 + 0.7f, y1 - 0.8f), rng.range(z0 + 0.6f, z1 - 0.6f)};
        const float f1 = rng.range(3.f, 9.f), f2 = rng.range(3.f, 9.f), p1 = rng.range(0, 6.28f), p2 = rng.range(0, 6.28f);
        const uint32_t mbase = 4 + (uint32_t)(rng.next() % nlamb);
        auto P = [&](uint32_t ring, uint32_t seg) {
            const float th = 3.14159265f * (float)ring / R, ph2 = 6.28318531f * (float)(seg % S) / S;
            const float d = rad * (1.0f + 0.10f * std::sin(f1 * th + p1) * std::cos(f2 * ph2 + p2) + 0.03f * std::sin(23.f * th) * std::sin(19.f * ph2));
            return V{c.x + d * std::sin(th) * std::cos(ph2), c.y + d * std::cos(th), c.z + d * std::sin(th) * std::sin(ph2)};
        };
        for (uint32_t r = 0; r < R; ++r)
            for (uint32_t s = 0; s < S; ++s) {
                const uint32_t m = 4 + (mbase + (r * 4 / R)) % nlamb;
                if (r != 0) put_tri(out, n, capn, P(r, s), P(r + 1, s), P(r, s + 1), m);
                if (r != R - 1) put_tri(out, n, capn, P(r, s + 1), P(r + 1, s), P(r + 1, s + 1), m);
            }
    }
    return n;
}


// ---------------------------------------------------------------------------------------------- OBJ ingestion
// What mesh.cpp:96-139 + RtScene.h:13-30 extract from an OBJ file: the positions of every face corner, faces in file
// order, polygons fanned around their first corner (tinyobjloader's default triangulation; the bundled files hold triangles
// only).  Texture coordinates, normals, groups and materials in the file are ignored, as they are by the reference's path
// tracer.  Indices may be negative (relative).  Returns the triangle count; `out` may be NULL to query it.
uint32_t vcrt_scene_load_obj(const char* path, uint32_t material_index, vcrt_triangle* out, uint32_t max_triangles) {
    g_err.clear();
    FILE* f = path ? std::fopen(path, "rb") : nullptr;
    if (!f) { g_err = std::string("failed to open file: ") + (path ? path : "(null)"); return 0; }
    std::vector<float> pos;
    std::vector<char> line(1 << 16);
    uint32_t n = 0;
    bool bad = false;
    while (std::fgets(line.data(), (int)line.size(), f)) {
        const char* p = line.data();
        while (*p == ' ' || *p == '\t') ++p;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            char* e = nullptr;
            p += 2;
            for (int k = 0; k < 3; ++k) { pos.push_back((float)std::strtod(p, &e)); if (e == p) bad = true; p = e; }
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            p += 2;
            long idx[64];
            int cnt = 0;
            for (;;) {
                while (*p == ' ' || *p == '\t') ++p;
                if (*p == 0 || *p == '\n' || *p == '\r') break;
                char* e = nullptr;
                long v = std::strtol(p, &e, 10);
                if (e == p) { bad = true; break; }
                const long nv = (long)(pos.size() / 3);
                v = v < 0 ? nv + v : v - 1;
                if (v < 0 || v >= nv) { bad = true; break; }
                if (cnt < 64) idx[cnt++] = v;
                p = e;
                while (*p && *p != ' ' && *p != '\t' && *p != '\n' && *p != '\r') ++p;   // skip /vt/vn
            }
            for (int k = 1; k + 1 < cnt; ++k) {
                if (out && n < max_triangles) {
                    vcrt_triangle t;
                    std::memset(&t, 0, sizeof t);
                    std::memcpy(t.v0, &pos[3 * idx[0]], 12); std::memcpy(t.v1, &pos[3 * idx[k]], 12); std::memcpy(t.v2, &pos[3 * idx[k + 1]], 12);
                    t.materialIndex = material_index;
                    out[n] = t;
                }
                ++n;
            }
        }
        if (bad) break;
    }
    std::fclose(f);
    if (bad) { g_err = std::string("failed to parse OBJ: ") + path; return 0; }
    if (n == 0) g_err = std::string("failed to load OBJ: no faces in ") + path;
    return n;
}

// The material table of RtScene.h:48-60: gray, red, green, white light (2,2,2), metal, glass.
uint32_t vcrt_scene_default_materials(vcrt_material* out, uint32_t max_materials) {
    static const struct { uint32_t type; float a[3]; } tab[6] = {
        {VCRT_MAT_LAMBERTIAN, {0.3f, 0.3f, 0.3f}}, {VCRT_MAT_LAMBERTIAN, {0.9f, 0.1f, 0.1f}}, {VCRT_MAT_LAMBERTIAN, {0.1f, 0.9f, 0.1f}},
        {VCRT_MAT_LIGHT, {2.0f, 2.0f, 2.0f}}, {VCRT_MAT_METAL, {1.0f, 1.0f, 1.0f}}, {VCRT_MAT_GLASS, {1.0f, 1.0f, 1.0f}}};
    for (uint32_t i = 0; i < 6 && out && i < max_materials; ++i) {
        std::memset(&out[i], 0, sizeof out[i]);
        out[i].type = tab[i].type;
        std::memcpy(out[i].albedo, tab[i].a, 12);
    }
    return 6;
}

}  // extern "C"
