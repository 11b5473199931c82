// vcrt_path.cuh -- per-pixel path loop shared by every kernel variant (and by tests/hostemu).
// ray_color: ray-trace-compute.comp:314-350; pixel epilogue: :375-379.
#pragma once

#include "vcrt_core.cuh"
#include "vcrt_fast.cuh"

namespace vcrt {

struct KernelArgs {
    SceneView scene;
    Camera cam;
    ShadeEnv env;
    uint32_t W, H, covW, covH;           // image size; covered extent (reference dispatch coverage or full)
    uint32_t tilesX, tilesY;             // 32x32 tiles (the reference's workgroup footprint, ray-trace-compute.comp:3)
    uint32_t tile_rank, tile_count;      // this call renders tile k iff k % tile_count == tile_rank
    uint32_t owned_tiles;                // number of such tiles
    uint32_t sample_begin, sample_count;
    uint32_t accum_mode, philox_seed, flags;
    uint32_t num_triangles;              // ubo.numTriangles: the loop bound of the brute-force hit_scene (ray-trace-compute.comp:229)
    uchar4* target;                      // binding 1
    uchar4* accum8;                      // binding 2
    float4* accumf;                      // f32 accumulation (sum of samples, w = sample count)
    vcrt_aov* aov;
    unsigned long long* counters;        // [0] rays [1] nodes [2] triangles
    unsigned int* work_counter;          // persistent kernels: next work item
    uint32_t leaf_threshold, shade_threshold;   // persistent kernel phase thresholds (lanes)
    uint32_t continue_threshold;                // trace kernel: lanes on inner nodes at or above which the phase votes are skipped
    float4* sample_out;                  // frames in flight (vcrt_frame_submit): the one-launch kernels store the frame's sample colour here, per pixel,
                                         // instead of folding it into the accumulation -- the fold happens in frame order (vcrt_post.cu: frame_fold_kernel)
};

// internal flag (not part of the C ABI's VCRT_FLAG_*): the accumulation of the covered pixels restarts with this call's samples
#define VCRT_FLAG_INTERNAL_RESTART 0x80000000u

// Work item -> pixel.  Items enumerate the owned 32x32 tiles; inside a tile, 32 consecutive items form an
// 8x4 pixel block so that a warp's primary rays are coherent.  Returns false for pixels outside the coverage.
VCRT_HD bool item_to_pixel(const KernelArgs& a, uint32_t item, uint32_t& x, uint32_t& y) {
    uint32_t j = item >> 10, inner = item & 1023u;
    uint32_t tile = j * (a.tile_count ? a.tile_count : 1u) + a.tile_rank;
    uint32_t ty = tile / a.tilesX, tx = tile - ty * a.tilesX;
    uint32_t sub = inner >> 5, lane = inner & 31u;
    x = tx * 32u + (sub & 3u) * 8u + (lane & 7u);
    y = ty * 32u + (sub >> 2) * 4u + (lane >> 3);
    return x < a.covW && y < a.covH;
}

template <int SHADER, int TRAV, bool COUNT>
VCRT_HD bool closest_hit(const KernelArgs& a, const Ray& r, Hit& rec, TraceStats& st) {
    st.rays++;
    if (TRAV == VCRT_TRAVERSAL_BRUTE_FORCE) return hit_scene<SHADER>(a.scene, a.num_triangles, r, rec, st);
    if (TRAV == VCRT_TRAVERSAL_FAST) return hit_bvh_fast<COUNT>(a.scene, r, rec, st);
    return hit_bvh_reference(a.scene, r, rec, a.env.stack_depth, st);
}

// One sample of one pixel: ray_color, ray-trace-compute.comp:314-350.
template <int SHADER, int TRAV, int RNG_MODE, int TRIG, bool COUNT>
VCRT_HD float3 ray_color(const KernelArgs& a, const Ray& primary, Rng& g, TraceStats& st, vcrt_aov* aov) {
    Hit rec;
    rec.p = rec.normal = f3(0, 0, 0); rec.materialIndex = 0; rec.t = 0; rec.backFaceInt = 0; rec.triangle = -1;
    float3 final_color = f3(1.0f, 1.0f, 1.0f);
    Ray cur;
    cur.o = primary.o;
    cur.d = normalize(primary.d);
    for (uint32_t i = 0; i < a.env.max_bounces; ++i) {
        rng_begin_bounce<RNG_MODE>(g, i);
        bool hit = closest_hit<SHADER, TRAV, COUNT>(a, cur, rec, st);
        if (i == 0) st.prim++;
        if (i == 0 && aov) {
            if (hit) { aov->triangle = rec.triangle; aov->material = (int32_t)rec.materialIndex; aov->t = rec.t; aov->backFace = (uint32_t)rec.backFaceInt; }
            else { aov->triangle = -1; aov->material = -1; aov->t = 0.0f; aov->backFace = 0u; }
        }
        if (hit) {
            float3 albedo;
            Ray next;
            bool emits = scatter<SHADER, RNG_MODE, TRIG>(a.scene, a.env, cur, rec, albedo, next, g);
            cur = next;
            final_color = mul(final_color, albedo);
            if (emits) break;
        } else {
            final_color = scale(final_color, 0.0f);
            break;
        }
    }
    return final_color;
}

// All samples of one pixel, in sample order (the order fixes the f32 sum and the rgba8 history).
template <int SHADER, int TRAV, int RNG_MODE, int TRIG, bool COUNT>
VCRT_HD void render_pixel(const KernelArgs& a, uint32_t x, uint32_t y, TraceStats& st) {
    const size_t pix = (size_t)y * a.W + x;
    const Ray primary = primary_ray(a.cam, x, y);
    const bool f32 = a.accum_mode == VCRT_ACCUM_F32;
    float4 acc = make_float4(0, 0, 0, 0);
    uchar4 px = make_uchar4(0, 0, 0, 0);
    if (a.sample_out) {   // a frame in flight (one sample): hand the colour over, the fold runs in frame order
        Rng g;
        rng_init<RNG_MODE>(g, x, y, (uint32_t)pix, a.sample_begin, a.philox_seed);
        const float3 c = ray_color<SHADER, TRAV, RNG_MODE, TRIG, COUNT>(a, primary, g, st, (a.flags & VCRT_FLAG_WRITE_AOV) ? &a.aov[pix] : nullptr);
        a.sample_out[pix] = make_float4(c.x, c.y, c.z, 1.0f);
        return;
    }
    if (f32) acc = a.accumf[pix];
    else px = a.accum8[pix];
    for (uint32_t k = 0; k < a.sample_count; ++k) {
        const uint32_t s = a.sample_begin + k;
        Rng g;
        rng_init<RNG_MODE>(g, x, y, (uint32_t)pix, s, a.philox_seed);
        vcrt_aov* aov = (k == 0 && (a.flags & VCRT_FLAG_WRITE_AOV)) ? &a.aov[pix] : nullptr;
        float3 c = ray_color<SHADER, TRAV, RNG_MODE, TRIG, COUNT>(a, primary, g, st, aov);
        if (f32) { acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += 1.0f; }
        else running_mean_rgba8(px, c, s);
    }
    if (f32) a.accumf[pix] = acc;
    else { a.target[pix] = px; a.accum8[pix] = px; }
}

}  // namespace vcrt
