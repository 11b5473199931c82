// vcrt_host_setup.h -- host-side evaluation of the kernel arguments that do not depend on device state:
// the scalar prologue of main() (ray-trace-compute.comp:355-369) and the work decomposition.
#pragma once
#include <cmath>
#include <cstring>
#include "vcrt_path.cuh"

namespace vcrt {

inline void setup_camera(Camera& cam, const vcrt_ubo& ubo, uint32_t W, uint32_t H) {
    // same operation order as the shader; tanf is libm's, as in the oracle
    const float pi = 3.1415926535897932385f;
    cam.imW = (float)W; cam.imH = (float)H;
    const float vfov = 30.0f;
    const float theta = vfov * pi / 180.0f;
    const float hh = tanf(theta / 2.0f);
    cam.viewport_height = 2.0f * hh;
    cam.viewport_width = cam.imW / cam.imH * cam.viewport_height;
    const float focal_length = 1.0f;
    cam.origin = make_float3(ubo.camPos[2] * -1.0f, ubo.camPos[0] * 1.0f, ubo.camPos[1] * 1.0f);
    const float3 half_h = make_float3(cam.viewport_width / 2.0f, 0.0f / 2.0f, 0.0f / 2.0f);
    const float3 half_v = make_float3(0.0f / 2.0f, -cam.viewport_height / 2.0f, 0.0f / 2.0f);
    cam.llc = make_float3(((cam.origin.x - half_h.x) - half_v.x) - 0.0f, ((cam.origin.y - half_h.y) - half_v.y) - 0.0f,
                          ((cam.origin.z - half_h.z) - half_v.z) - focal_length);
}

// Fills everything except pointers (scene, images, counters).
inline void setup_args(KernelArgs& a, const vcrt_ubo& ubo, const vcrt_render_params& p, uint32_t W, uint32_t H, uint32_t covW, uint32_t covH, uint32_t nlights) {
    setup_camera(a.cam, ubo, W, H);
    a.env.lights_length = p.lights_length ? p.lights_length : nlights;
    a.env.max_bounces = p.max_bounces ? p.max_bounces : (p.shader == VCRT_SHADER_SIMPLE ? 4u : 2u);
    a.env.stack_depth = (int)(p.stack_depth ? p.stack_depth : 16u);
    a.W = W; a.H = H; a.covW = covW; a.covH = covH;
    a.tilesX = (W + 31) / 32; a.tilesY = (H + 31) / 32;
    a.tile_count = p.tile_count > 1 ? p.tile_count : 1u;
    a.tile_rank = p.tile_count > 1 ? p.tile_rank : 0u;
    const uint32_t tiles = a.tilesX * a.tilesY;
    a.owned_tiles = tiles > a.tile_rank ? (tiles - a.tile_rank + a.tile_count - 1) / a.tile_count : 0u;
    a.sample_begin = p.sample_begin;
    a.sample_count = p.sample_count ? p.sample_count : 1u;
    a.accum_mode = p.accum_mode; a.philox_seed = p.philox_seed; a.flags = p.flags;
    a.num_triangles = ubo.numTriangles;
}

}  // namespace vcrt
