// TEST INFRASTRUCTURE ONLY -- runs the product's `__host__ __device__` path code (vcrt_path.cuh, vcrt_fast.cuh,
// vcrt_repack.cpp) on the CPU, pixel by pixel, so that traversal/shading logic can be checked against the
// oracle in the GPU-less container.  It is NOT a product path: nothing in vulkan_compute_ray_tracing_b200/
// links or loads it, and it is built only by tests (tests/hostemu/Makefile).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../vulkan_compute_ray_tracing_b200/csrc/vcrt_host_setup.h"
#include "../../vulkan_compute_ray_tracing_b200/csrc/vcrt_repack.h"

using namespace vcrt;

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
    if (p->traversal == VCRT_TRAVERSAL_FAST) {
        std::string e;
        if (!build_fast_bvh((const vcrt_bvh_node*)bvh, nbvh, (const vcrt_triangle*)tris, ntris, fb, e)) {
            if (err && errlen > 0) { std::strncpy(err, e.c_str(), errlen - 1); err[errlen - 1] = 0; }
            return -1;
        }
        if ((!(p->_reserved & 1u) && !rebuild_fast_bvh_sah(fb, e)) || !check_fast_depth(fb, e)) {   // test hook: _reserved bit 0 keeps the bound topology
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
