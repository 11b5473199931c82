// TEST INFRASTRUCTURE ONLY -- runs the product's `__host__ __device__` path code (vcrt_path.cuh, vcrt_fast.cuh,
// vcrt_repack.cpp) on the CPU, pixel by pixel, so that traversal/shading logic can be checked against the
// oracle in the GPU-less container.  It is NOT a product path: nothing in vulkan_compute_ray_tracing_b200/
// links or loads it, and it is built only by tests (tests/hostemu/Makefile).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include <algorithm>
#include <cmath>
#include <numeric>

#include "../../vulkan_compute_ray_tracing_b200/csrc/vcrt_devbuild.cuh"
#include "../../vulkan_compute_ray_tracing_b200/csrc/vcrt_host_setup.h"
#include "../../vulkan_compute_ray_tracing_b200/csrc/vcrt_repack.h"

using namespace vcrt;

// The on-device record build (vcrt_devbuild.cu) run sequentially on the CPU: the same per-element bodies (vcrt_devbuild.cuh) in
// plain loops, std::stable_sort for the radix sort, running sums for the scans.  Returns an empty string on success, else why not.
struct DevBuildOut { std::vector<float4> tris64; std::vector<uint32_t> q4; float qorg[3], qext[3]; uint32_t nwide = 0, stack4 = 0, depth = 0, rounds = 0; };
static std::string devbuild_host(const vcrt_bvh_node* bvh, uint32_t nbvh, const vcrt_triangle* tris, uint32_t ntris, float max_quantum, DevBuildOut& o) {
    using namespace devbuild;
    if (nbvh < 3) return "fewer than two leaves";
    std::vector<int32_t> parent(nbvh, -1);
    std::vector<uint32_t> count(nbvh, 0), arrive(nbvh, 0), status(ST_WORDS, 0);
    for (int k = 0; k < 3; ++k) status[ST_CENTROID + k] = 0xffffffffu;
    View v;
    v.bvh = bvh; v.nbvh = nbvh; v.tris = tris; v.ntris = ntris; v.parent = parent.data(); v.count = count.data(); v.arrive = arrive.data(); v.status = status.data(); v.leaf_node = nullptr;
    for (uint32_t i = 0; i < nbvh; ++i) link_children(v, i);
    for (uint32_t i = 0; i < nbvh; ++i) count_up(v, i);
    if (status[ST_ERROR]) return "not a plain tree: flags " + std::to_string(status[ST_ERROR]);
    const uint32_t n = count[0];
    if (n < 2) return "fewer than two leaves";
    std::vector<uint32_t> leaf_node(n, 0xffffffffu);
    v.leaf_node = leaf_node.data();
    for (uint32_t i = 0; i < nbvh; ++i) {
        if (!is_leaf(bvh[i])) continue;
        uint32_t rank, depth;
        if (!leaf_rank(v, i, rank, depth)) continue;
        if (rank < n) leaf_node[rank] = i; else status[ST_ERROR] |= ERR_SHARED;
    }
    o.tris64.assign((size_t)n * 4, make_float4(0, 0, 0, 0));
    std::vector<float4> lo_a(n), hi_a(n), lo_b(n), hi_b(n);
    for (uint32_t s = 0; s < n; ++s) make_slot(v, s, o.tris64.data(), lo_a.data(), hi_a.data());
    if (status[ST_ERROR]) return "flags " + std::to_string(status[ST_ERROR]);
    float clo[3], chi[3];
    for (int k = 0; k < 3; ++k) { clo[k] = ord2f(status[ST_CENTROID + k]); chi[k] = ord2f(status[ST_CENTROID + 3 + k]); }
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    for (uint32_t i = 0; i < n; ++i) keys[i] = morton_of(lo_a[i], hi_a[i], clo, chi);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    for (uint32_t i = 0; i < n; ++i) { lo_b[i] = lo_a[order[i]]; hi_b[i] = hi_a[order[i]]; }
    std::vector<float> nodes((size_t)(n - 1) * 16, 0.0f);
    std::vector<uint32_t> nn(n);
    std::vector<uint64_t> scan(n);
    float4 *cl = lo_b.data(), *ch = hi_b.data(), *nl = lo_a.data(), *nh = hi_a.data();
    uint32_t m = n, node_base = 0;
    while (m > 1) {
        for (uint32_t i = 0; i < m; ++i) nn[i] = nearest(cl, ch, m, i);
        uint64_t run = 0;
        for (uint32_t i = 0; i < m; ++i) { scan[i] = run; run += merge_flags(nn.data(), i); }
        for (uint32_t i = 0; i < m; ++i) merge_write(cl, ch, nn.data(), scan.data(), i, node_base, nl, nh, nodes.data());
        const uint32_t next = (uint32_t)(run & 0xffffffffull), made = (uint32_t)(run >> 32);
        if (next >= m || next + made != m) return "PLOC made no progress";
        m = next; node_base += made;
        std::swap(cl, nl); std::swap(ch, nh);
        ++o.rounds;
    }
    if (node_base != n - 1) return "PLOC node count";
    QFrame q;
    const double blo[3] = {cl[0].x, cl[0].y, cl[0].z}, bhi[3] = {ch[0].x, ch[0].y, ch[0].z};
    for (int a = 0; a < 3; ++a) {
        double ext = bhi[a] - blo[a];
        if (!(ext > 0.0)) ext = 1e-3;
        const double quantum = ext / 32764.0, base = blo[a] - quantum, E = 32768.0 * quantum;
        if (quantum > (double)max_quantum) return "scene extent too large for 15-bit bounds";
        q.org[a] = (float)(base - E); q.ext[a] = (float)E;
        if (std::fabs((double)q.org[a] - (base - E)) > 0.01 * quantum || std::fabs((double)q.ext[a] - E) > 1e-6 * E) return "frame rounding";
    }
    o.q4.assign((size_t)(n - 1) * 16, 0u);
    std::vector<WideItem> items_a(n), items_b(n);
    std::vector<uint32_t> inner(n), offs(n);
    items_a[0].node2 = (int32_t)f2u(cl[0].w); items_a[0].stack_above = 0u;
    uint32_t cnt = 1, level_base = 0;
    while (cnt > 0) {
        uint32_t run = 0;
        for (uint32_t i = 0; i < cnt; ++i) { inner[i] = wide_inner_count(nodes.data(), items_a[i]); offs[i] = run; run += inner[i]; }
        for (uint32_t i = 0; i < cnt; ++i) wide_write(nodes.data(), q, items_a[i], level_base + i, level_base + cnt + offs[i], offs[i], items_b.data(), o.q4.data(), status.data());
        level_base += cnt; cnt = run;
        items_a.swap(items_b);
        if (level_base + cnt > n - 1) return "4-wide node count";
    }
    o.nwide = level_base; o.stack4 = status[ST_STACK] + 2; o.depth = f2u(ch[0].w);
    std::memcpy(o.qorg, q.org, sizeof o.qorg); std::memcpy(o.qext, q.ext, sizeof o.qext);
    if (o.stack4 > VCRT_FAST_STACK) return "stack4 " + std::to_string(o.stack4);
    return "";
}

template <int SHADER, int TRAV, int RNG_MODE, int TRIG>
static void run(const KernelArgs& a, unsigned long long* counters) {
    const uint32_t items = a.owned_tiles * 1024u;
    unsigned long long rays = 0, nodes = 0, tris = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : rays, nodes, tris)
    for (uint32_t item = 0; item < items; ++item) {
        uint32_t x, y;
        TraceStats st = {0u, 0u, 0u, 0u};
        if (item_to_pixel(a, item, x, y)) render_pixel<SHADER, TRAV, RNG_MODE, TRIG, true>(a, x, y, st);
        rays += st.rays; nodes += st.nodes; tris += st.tris;
    }
    counters[0] += rays; counters[1] += nodes; counters[2] += tris;
}

template <int SHADER, int TRAV>
static void run2(const KernelArgs& a, int rng, int trig, unsigned long long* c) {
    if (rng == VCRT_RNG_PHILOX) { if (trig) run<SHADER, TRAV, VCRT_RNG_PHILOX, VCRT_TRIG_PORTABLE>(a, c); else run<SHADER, TRAV, VCRT_RNG_PHILOX, VCRT_TRIG_LIBM>(a, c); }
    else { if (trig) run<SHADER, TRAV, VCRT_RNG_PCG_REF, VCRT_TRIG_PORTABLE>(a, c); else run<SHADER, TRAV, VCRT_RNG_PCG_REF, VCRT_TRIG_LIBM>(a, c); }
}

extern "C" int hostemu_render(const void* tris, uint32_t ntris, const void* mats, uint32_t nmats, const void* bvh, uint32_t nbvh,
                              const void* lights, uint32_t nlights, const void* spheres, uint32_t nspheres, const vcrt_ubo* ubo,
                              const vcrt_render_params* p, uint32_t W, uint32_t H, void* target, void* accum8, void* accumf, void* aov,
                              unsigned long long* counters, char* err, int errlen) {
    KernelArgs a;
    std::memset(&a, 0, sizeof a);
    SceneView& s = a.scene;
    s.tris = (const float4*)tris; s.ntris = ntris; s.mats = (const float4*)mats; s.nmats = nmats; s.bvh = (const float4*)bvh; s.nbvh = nbvh;
    s.lights = (const vcrt_light*)lights; s.nlights = nlights; s.spheres = (const float4*)spheres; s.nspheres = nspheres;
    s.froot = (int32_t)0x80000000;
    FastBvh fb;
    DevBuildOut db;
    if (p->traversal == VCRT_TRAVERSAL_FAST && (p->_reserved & 16u)) {   // test hook: _reserved bit 4 = records by the device builder's algorithm
        const std::string e = devbuild_host((const vcrt_bvh_node*)bvh, nbvh, (const vcrt_triangle*)tris, ntris, (p->_reserved & 4u) ? 3.0e38f : 2.5e-4f, db);
        if (!e.empty()) {
            if (err && errlen > 0) { std::strncpy(err, ("devbuild: " + e).c_str(), errlen - 1); err[errlen - 1] = 0; }
            return -2;
        }
        s.ftris = db.tris64.data(); s.q4nodes = (const Words8*)db.q4.data(); s.froot4 = 0;
        s.qorg = make_float3(db.qorg[0], db.qorg[1], db.qorg[2]);
        s.qext = make_float3(db.qext[0], db.qext[1], db.qext[2]);
    } else if (p->traversal == VCRT_TRAVERSAL_FAST) {
        std::string e;
        if (!build_fast_bvh((const vcrt_bvh_node*)bvh, nbvh, (const vcrt_triangle*)tris, ntris, fb, e)) {
            if (err && errlen > 0) { std::strncpy(err, e.c_str(), errlen - 1); err[errlen - 1] = 0; }
            return -1;
        }
        // test hooks: _reserved bit 0 keeps the bound topology; bits 8..15 = reinsertion passes, bits 16..23 = percent of the nodes per pass
        if ((!(p->_reserved & 1u) && (!rebuild_fast_bvh_sah(fb, e) || (((p->_reserved >> 8) & 255u) && !optimize_fast_bvh_reinsert(fb, (int)((p->_reserved >> 8) & 255u), 0.01f * (float)((p->_reserved >> 16) & 255u), e)))) || !check_fast_depth(fb, e)) {
            if (err && errlen > 0) { std::strncpy(err, e.c_str(), errlen - 1); err[errlen - 1] = 0; }
            return -1;
        }
        precompute_triangles(fb);
        s.fnodes = (const float4*)fb.nodes.data(); s.ftris = (const float4*)fb.tris64.data(); s.nfnodes = fb.num_nodes(); s.froot = fb.root;
        // test hooks: _reserved bit 1 keeps the 64-byte float nodes, bit 2 quantises whatever the scene extent
        if (!(p->_reserved & 2u) && quantize_fast_bvh(fb, (p->_reserved & 4u) ? 3.0e38f : 2.5e-4f)) {
            s.qnodes = (const Words8*)fb.qnodes.data();   // std::vector storage is 16-byte aligned; the host load is a plain copy
            s.qorg = make_float3(fb.qorg[0], fb.qorg[1], fb.qorg[2]);
            s.qext = make_float3(fb.qext[0], fb.qext[1], fb.qext[2]);
            // test hook: _reserved bit 3 walks the 4-wide form of the quantised tree
            if ((p->_reserved & 8u) && build_wide_bvh(fb, VCRT_FAST_STACK)) { s.q4nodes = (const Words8*)fb.q4nodes.data(); s.froot4 = fb.root4; }
        }
    }
    const bool ref_cov = (p->flags & VCRT_FLAG_REF_DISPATCH_COVERAGE) != 0;
    setup_args(a, *ubo, *p, W, H, ref_cov ? (W / 32) * 32 : W, ref_cov ? (H / 32) * 32 : H, nlights);
    a.target = (uchar4*)target; a.accum8 = (uchar4*)accum8; a.accumf = (float4*)accumf; a.aov = (vcrt_aov*)aov;
    const int sh = (int)p->shader, rng = (int)p->rng_mode, trig = (int)p->trig_mode;
    switch (p->traversal) {
        case VCRT_TRAVERSAL_FAST: if (sh) run2<1, 1>(a, rng, trig, counters); else run2<0, 1>(a, rng, trig, counters); break;
        case VCRT_TRAVERSAL_BRUTE_FORCE: if (sh) run2<1, 2>(a, rng, trig, counters); else run2<0, 2>(a, rng, trig, counters); break;
        default: if (sh) run2<1, 0>(a, rng, trig, counters); else run2<0, 0>(a, rng, trig, counters); break;
    }
    return 0;
}
