// TEST INFRASTRUCTURE ONLY -- wraps one rewritten reference shader (SHADER_INC) as a C entry point.
// One `Invocation` object == one GLSL invocation (globals of the shader are its members, so the
// per-invocation global initialiser `rngState = ...` of random.glsl:19 runs in the constructor).
// Built by oracle/ref/Makefile into oracle/_ref/libvcrt_ref.so; compiled once per variant with
//   -DSHADER_INC="..." -DVARIANT=full_b2_s16
#include "glsl_prelude.hpp"
#include <cstdio>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

namespace CAT(ns_, VARIANT) {

struct Invocation : InvocationBase {
    Invocation(const InvocationBase& b) : InvocationBase(b) {}
#include SHADER_INC
};

static_assert(sizeof(Invocation::material) == 32 && offsetof(Invocation::material, albedo) == 16, "std430 material");
static_assert(sizeof(Invocation::triangle) == 48 && offsetof(Invocation::triangle, v1) == 16 &&
              offsetof(Invocation::triangle, v2) == 32 && offsetof(Invocation::triangle, materialIndex) == 44, "std430 triangle");
static_assert(sizeof(Invocation::sphere) == 32 && offsetof(Invocation::sphere, materialIndex) == 16, "std430 sphere");
static_assert(sizeof(Invocation::bvhNode) == 48 && offsetof(Invocation::bvhNode, max) == 16 &&
              offsetof(Invocation::bvhNode, leftNodeIndex) == 28 && offsetof(Invocation::bvhNode, rightNodeIndex) == 32 &&
              offsetof(Invocation::bvhNode, objectIndex) == 36, "std430 bvhNode");
static_assert(sizeof(Invocation::light) == 8, "std430 light");

}  // namespace

using CAT(ns_, VARIANT)::Invocation;

// Runs gx*gy workgroups of 32x32 invocations (vkCmdDispatch(gx, gy, 1), main.cpp:228).
extern "C" void CAT(ref_dispatch_, VARIANT)(const Bindings* b, int gx, int gy) {
    const int W = gx * 32, H = gy * 32;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            InvocationBase base{{(uint)x, (uint)y, 0u}, b};
            Invocation inv(base);
            inv.main();
        }
    }
}

// Closest-hit query of the reference's hit_bvh for caller-supplied rays (dir is used as given).
// out[i] = {hit, materialIndex, backFaceInt, t bits, p.xyz bits, normal.xyz bits} as 10 x u32.
extern "C" void CAT(ref_hit_bvh_, VARIANT)(const Bindings* b, const float* org_dir6, int n, uint32_t* out10) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        InvocationBase base{{0u, 0u, 0u}, b};
        Invocation inv(base);
        Invocation::ray r{vec3(org_dir6[6 * i], org_dir6[6 * i + 1], org_dir6[6 * i + 2]),
                          vec3(org_dir6[6 * i + 3], org_dir6[6 * i + 4], org_dir6[6 * i + 5])};
        Invocation::hit_record rec;
        std::memset((void*)&rec, 0, sizeof(rec));
        bool hit = inv.hit_bvh(r, rec);
        uint32_t* o = out10 + 10 * i;
        std::memset(o, 0, 40);
        o[0] = hit ? 1u : 0u;
        if (hit) {
            o[1] = rec.materialIndex;
            o[2] = (uint32_t)rec.backFaceInt;
            std::memcpy(o + 3, &rec.t, 4);
            std::memcpy(o + 4, &rec.p, 12);
            std::memcpy(o + 7, &rec.normal, 12);
        }
    }
}

// The reference's PCG stream for a given seed (random.glsl:4-22): n floats.
extern "C" void CAT(ref_random_, VARIANT)(const Bindings* b, uint32_t seed, int n, float* out) {
    InvocationBase base{{0u, 0u, 0u}, b};
    Invocation inv(base);
    inv.rngState = seed;
    for (int i = 0; i < n; ++i) out[i] = inv.random();
}
