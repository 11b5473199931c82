// vcrt_persistent.cuh -- persistent-warps megakernel for the fast traversal (GPU only).
//
// Why: in the one-thread-per-pixel kernel a lane whose path ends after one ray idles while its neighbours bounce up
// to max_bounces times, and lanes at leaves / inner nodes serialise; ncu on the 1M-triangle scene showed 5.3 of 32
// lanes active per issued instruction (profiles/r01_v1_static_fast_ncu_summary.json).  Here every lane is a small
// state machine that owns one pixel at a time (all of its samples, so the f32 sum keeps the oracle's order) and
// the warp alternates between three phases, each entered only when enough lanes want it:
//
//   T  inner-node steps for all lanes that have a node to visit           (the hot loop, ~95 % of instructions)
//   L  triangle tests for every lane with a postponed leaf                (entered when >= LEAF_T lanes are blocked)
//   S  shading / next ray / next sample / next pixel for finished lanes   (entered when >= SHADE_T lanes are waiting)
//
// A lane postpones one leaf while it keeps traversing (speculative traversal); work is fetched per lane with an
// atomic counter over the item enumeration of item_to_pixel.  The arithmetic of every step is the shared
// __host__ __device__ code of vcrt_fast.cuh / vcrt_core.cuh, so results are bit-identical to the static kernel.
#pragma once

#include "vcrt_path.cuh"

namespace vcrt {


template <int SHADER, int RNG_MODE, int TRIG, bool COUNT>
__global__ void __launch_bounds__(VCRT_PBLOCK, VCRT_MEGA_MINB) render_persistent_kernel(const __grid_constant__ KernelArgs a) {
    const unsigned FULL = 0xffffffffu;
    const uint32_t total_items = a.owned_tiles * 1024u;
    const bool f32 = a.accum_mode == VCRT_ACCUM_F32;
    const SceneView& s = a.scene;
    const bool wide = s.q4nodes != nullptr;   // uniform: 4-wide quantised nodes (the default records) ...
    const bool qn = s.qnodes != nullptr;      // ... else binary quantised 32-byte nodes, else 64-byte float nodes

    // ---- lane state
    bool has_pixel = false, done = false;
    uint32_t x = 0, y = 0, pix = 0, k = 0, bounce = 0, rng_saved = 0;
    float4 acc = make_float4(0, 0, 0, 0);
    uchar4 px8 = make_uchar4(0, 0, 0, 0);
    Ray cur; cur.o = cur.d = f3(0, 0, 0);
    float3 thr = f3(1, 1, 1);
    TravState t;
    t.idir = t.ood = f3(0, 0, 0); t.closest = VCRT_T_MAX; t.best = -1; t.node = VCRT_FAST_EMPTY; t.sp = 0;
    t.selx = t.sely = t.selz = VCRT_Q15_SEL_LO;
    int32_t pending = VCRT_FAST_EMPTY;
    int32_t stack[VCRT_FAST_STACK];
    bool have_ray = false;          // a ray is in flight (or just finished traversal and awaits shading)
    TraceStats st = {0u, 0u, 0u, 0u};

    for (;;) {
        // =========================================================== S: lanes whose ray is finished (or that have none)
        const bool waiting = !done && t.node == VCRT_FAST_EMPTY && pending == VCRT_FAST_EMPTY;
        if (waiting) {
            bool path_over = !have_ray;
            if (have_ray) {
                // ---- ray_color body for the ray that just finished (ray-trace-compute.comp:321-340)
                Hit rec;
                const bool hit = trav_finish(t, s, cur, rec);
                if (bounce == 0 && k == 0 && (a.flags & VCRT_FLAG_WRITE_AOV)) {
                    vcrt_aov o;
                    if (hit) { o.triangle = rec.triangle; o.material = (int32_t)rec.materialIndex; o.t = rec.t; o.backFace = (uint32_t)rec.backFaceInt; }
                    else { o.triangle = -1; o.material = -1; o.t = 0.0f; o.backFace = 0u; }
                    a.aov[pix] = o;
                }
                if (hit) {
                    Rng g;   // PCG: one word carried across the path; Philox: rebuilt from (pixel, sample, bounce)
                    g.pcg = rng_saved;
                    g.key0 = pix; g.key1 = a.philox_seed; g.ctr0 = a.sample_begin + k;
                    rng_begin_bounce<RNG_MODE>(g, bounce);
                    float3 albedo;
                    Ray next;
                    const bool emits = scatter<SHADER, RNG_MODE, TRIG>(s, a.env, cur, rec, albedo, next, g);
                    rng_saved = g.pcg;
                    cur = next;
                    thr = mul(thr, albedo);
                    bounce++;
                    path_over = emits || bounce >= a.env.max_bounces;
                } else {
                    thr = scale(thr, 0.0f);
                    path_over = true;
                }
                if (path_over) {   // ray-trace-compute.comp:375-379
                    if (f32) { acc.x += thr.x; acc.y += thr.y; acc.z += thr.z; acc.w += 1.0f; }
                    else running_mean_rgba8(px8, thr, a.sample_begin + k);
                    k++;
                    if (k == a.sample_count) {
                        if (a.sample_out) a.sample_out[pix] = make_float4(thr.x, thr.y, thr.z, 1.0f);   // a frame in flight (one sample): folded in frame order later
                        else if (f32) a.accumf[pix] = acc;
                        else { a.target[pix] = px8; a.accum8[pix] = px8; }
                        has_pixel = false;
                    }
                }
            }
            if (!has_pixel) {
                for (;;) {   // next work item; items outside the covered extent are skipped
                    const uint32_t item = atomicAdd(a.work_counter, 1u);
                    if (item >= total_items) { done = true; break; }
                    if (item_to_pixel(a, item, x, y)) break;
                }
                if (!done) {
                    has_pixel = true;
                    pix = y * a.W + x;
                    k = 0;
                    if (!a.sample_out) { if (f32) acc = a.accumf[pix]; else px8 = a.accum8[pix]; }
                }
            }
            have_ray = !done;
            if (!done) {
                if (path_over) {   // new sample: primary ray, ray-trace-compute.comp:352-373, :317-319
                    const Ray pr = primary_ray(a.cam, x, y);
                    cur.o = pr.o;
                    cur.d = normalize(pr.d);
                    thr = f3(1.0f, 1.0f, 1.0f);
                    bounce = 0;
                    rng_saved = (600u * x + y) * (a.sample_begin + k + 1u);   // random.glsl:19 (unused by Philox)
                }
                if (wide) trav_begin<2>(t, s, cur); else if (qn) trav_begin<1>(t, s, cur); else trav_begin<0>(t, s, cur);
                st.rays++;
                if (bounce == 0) st.prim++;
            }
        }
        if (__all_sync(FULL, done)) break;

        // =========================================================== T / L until enough lanes wait for S
        for (;;) {
            if (t.node >= 0) {
                if (COUNT) st.nodes++;
                if (wide) trav_inner_step4(t, s, stack); else if (qn) trav_inner_step<1>(t, s, stack); else trav_inner_step<0>(t, s, stack);
            }
            if (t.node < 0 && t.node != VCRT_FAST_EMPTY && pending == VCRT_FAST_EMPTY) {   // postpone one leaf, keep going
                pending = t.node;
                t.node = t.sp ? stack[--t.sp] : VCRT_FAST_EMPTY;
            }
            const bool inner = t.node >= 0;
            const bool blocked = !inner && pending != VCRT_FAST_EMPTY;   // needs the leaf phase before it can go on
            const unsigned mi = __ballot_sync(FULL, inner);
            const unsigned mb = __ballot_sync(FULL, blocked);
            if (mb != 0u && (__popc(mb) >= (int)a.leaf_threshold || mi == 0u)) {
                if (pending != VCRT_FAST_EMPTY) {
                    if (COUNT) st.tris++;
                    trav_leaf_test(t, s, cur, pending);
                    pending = VCRT_FAST_EMPTY;
                    if (t.node < 0 && t.node != VCRT_FAST_EMPTY) {   // the leaf it was blocked on becomes the new postponed one
                        pending = t.node;
                        t.node = t.sp ? stack[--t.sp] : VCRT_FAST_EMPTY;
                    }
                }
                continue;
            }
            const bool finished = !done && t.node == VCRT_FAST_EMPTY && pending == VCRT_FAST_EMPTY;
            const unsigned mf = __ballot_sync(FULL, finished);
            if (mf != 0u && (__popc(mf) >= (int)a.shade_threshold || mi == 0u)) break;
            if (mi == 0u && mb == 0u) break;   // nothing left to traverse (only done lanes and finished ones)
        }
    }
    flush_stats(a, st);
}

}  // namespace vcrt
