// vcrt_core.cuh -- device-side building blocks of the path-tracing hot path (sm_100a).
//
// Everything here is `__host__ __device__` so that tests/hostemu can run the very same code on the CPU
// as a debugging aid; the product only ever launches it on the GPU (vcrt_kernels.cu).
//
// Floating-point contract: this translation unit is compiled with -fmad=false, IEEE division and sqrt
// (nvcc defaults -prec-div=true -prec-sqrt=true).  Every value that decides a hit or a colour is
// computed with the same operation order as the reference shader evaluated through glm (dot = (x+y)+z,
// normalize = v * (1/sqrt(dot)), min = (y<x)?y:x ...).  Fused multiply-adds appear only where written
// explicitly (fmaf) -- in the conservative box tests of the fast traversal, never in a hit decision.
//
// Reference citations are file:line in /root/reference/resources/shaders/source.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vcrt.h"

#define VCRT_HD __host__ __device__ __forceinline__

namespace vcrt {

// ------------------------------------------------------------------------------------------ vectors
VCRT_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
VCRT_HD float3 add(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
VCRT_HD float3 sub(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
VCRT_HD float3 mul(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
VCRT_HD float3 scale(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
VCRT_HD float3 neg(float3 a) { return f3(-a.x, -a.y, -a.z); }
VCRT_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
VCRT_HD float3 cross(float3 x, float3 y) { return f3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
VCRT_HD float3 normalize(float3 v) { return scale(v, 1.0f / sqrtf(dot(v, v))); }
VCRT_HD float glsl_min(float x, float y) { return (y < x) ? y : x; }
VCRT_HD float glsl_max(float x, float y) { return (x < y) ? y : x; }
VCRT_HD float3 reflect(float3 I, float3 N) { return sub(I, scale(scale(N, dot(N, I)), 2.0f)); }
VCRT_HD float3 refract(float3 I, float3 N, float eta) {
    float d = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k >= 0.0f) return sub(scale(I, eta), scale(N, eta * d + sqrtf(k)));
    return f3(0.0f, 0.0f, 0.0f);
}
VCRT_HD float3 xyz(float4 v) { return f3(v.x, v.y, v.z); }

VCRT_HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
VCRT_HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

VCRT_HD float4 ldg4(const float4* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// One 256-bit read-only load (LDG.E.256 on sm_100a): a divergent lane costs one L1 data-pipe wavefront per 32-byte
// sector it touches, so 32 bytes per instruction halves the wavefronts of two 128-bit loads.  p must be 32-byte aligned.
#ifndef VCRT_L2HINT
#define VCRT_L2HINT 1
#endif
struct __align__(32) Words8 { uint32_t w[8]; };
VCRT_HD Words8 ldg8(const Words8* p) {
#ifdef __CUDA_ARCH__
    Words8 r;
    // VCRT_L2HINT: scene records are marked evict-last in L2 (LDG.E.ELL2.256) and the ray queues evict-first (stream_*
    // below): a trace launch streams ~1 GB of queue records past a ~100 MB scene, and without the hints 27 % of the
    // kernel's L2 sector requests went to DRAM (ncu r01_v8), i.e. most warp iterations waited on at least one DRAM miss.
#if VCRT_L2HINT
    asm("ld.global.nc.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
        : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
    return r;
#else
    Words8 r;   // host emulation: the source may be only 4-byte aligned
    for (int i = 0; i < 8; ++i) r.w[i] = ((const uint32_t*)p)[i];
    return r;
#endif
}

// Streaming accesses of the wavefront queues (read once / written once per bounce): evict-first when VCRT_L2HINT.
#if defined(__CUDA_ARCH__) && VCRT_L2HINT
__device__ __forceinline__ float4 stream_ld(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ uint2 stream_ld(const uint2* p) { return __ldcs(p); }
__device__ __forceinline__ void stream_st(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void stream_st(uint2* p, uint2 v) { __stcs(p, v); }
#else
VCRT_HD float4 stream_ld(const float4* p) { return *p; }
VCRT_HD uint2 stream_ld(const uint2* p) { return *p; }
VCRT_HD void stream_st(float4* p, float4 v) { *p = v; }
VCRT_HD void stream_st(uint2* p, uint2 v) { *p = v; }
#endif

// Hint: bring the line holding p into L1 (no register, no scoreboard wait).  No-op on the host.
VCRT_HD void prefetch_l1(const void* p) {
#ifdef __CUDA_ARCH__
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// 15-bit fixed point -> float m = 0.5 * (1 + q / 32768) in [0.5, 1), exactly, in one PRMT: the 16-bit field becomes
// bits 8..23 of a float whose exponent byte is 0x3F (bit 23 = the field's top bit = 0).
// Same with the half chosen by a PRMT selector held in a register: 0x7104 = low half, 0x7324 = high half.  The
// traversal keeps one selector per axis (by the sign of the ray direction) so that the near and far planes of a box come
// out of the decode directly, without the min/max pair of the generic slab test.
#define VCRT_Q15_SEL_LO 0x7104u
#define VCRT_Q15_SEL_HI 0x7324u
VCRT_HD float q15_sel(uint32_t w, uint32_t sel) {
#ifdef __CUDA_ARCH__
    // prmt.b32 directly: __byte_perm would first mask the selector (one LOP3 per axis and visit)
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0x3F000000u), "r"(sel));
    return __uint_as_float(r);
#else
    return u2f(0x3F000000u | ((sel == VCRT_Q15_SEL_LO ? (w & 0xffffu) : (w >> 16)) << 8));
#endif
}
VCRT_HD float q15_lo(uint32_t w) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(__byte_perm(w, 0x3F000000u, 0x7104));
#else
    return u2f(0x3F000000u | ((w & 0xffffu) << 8));
#endif
}
VCRT_HD float q15_hi(uint32_t w) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(__byte_perm(w, 0x3F000000u, 0x7324));
#else
    return u2f(0x3F000000u | ((w >> 16) << 8));
#endif
}

// ------------------------------------------------------------------------------------------ trig
// VCRT_TRIG_PORTABLE: fixed fp32 sequence, bit-identical to the oracle's (no FMA, IEEE ops only).
VCRT_HD void sincos_portable(float x, float* s, float* c) {
    int q = (int)(x * 0.636619772367581343f + 0.5f);
    float fq = (float)q;
    float r = ((x - fq * 1.5703125f) - fq * 4.837512969970703125e-4f) - fq * 7.549789948768648e-8f;
    float z = r * r;
    float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    float pc = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    switch (q & 3) {
        case 0: *s = ps;  *c = pc;  break;
        case 1: *s = pc;  *c = -ps; break;
        case 2: *s = -ps; *c = -pc; break;
        default: *s = -pc; *c = ps; break;
    }
}

// ------------------------------------------------------------------------------------------ RNG
// PCG-RXS-M-XS-32 exactly as include/random.glsl:4-22; Philox4x32-10 for the production mode.
// Philox stream: key = (pixel, seed), counter = (sample, bounce, block-in-bounce, 0): every bounce starts a fresh
// block, so no generator state has to survive a traversal and a bounce costs one block (two with light sampling).
struct Rng {
    uint32_t pcg;
    uint32_t key0, key1, ctr0, ctr1, ctr2;
    uint32_t buf[4];
    uint32_t have;
};

VCRT_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

template <int RNG_MODE>
VCRT_HD void rng_init(Rng& g, uint32_t x, uint32_t y, uint32_t pix, uint32_t sample, uint32_t seed) {
    if (RNG_MODE == VCRT_RNG_PCG_REF) {
        g.pcg = (600u * x + y) * (sample + 1u);  // random.glsl:19
    } else {
        g.key0 = pix; g.key1 = seed; g.ctr0 = sample; g.ctr1 = 0u; g.ctr2 = 0u; g.have = 0u;
    }
}

template <int RNG_MODE>
VCRT_HD void rng_begin_bounce(Rng& g, uint32_t bounce) {
    if (RNG_MODE == VCRT_RNG_PHILOX) { g.ctr1 = bounce; g.ctr2 = 0u; g.have = 0u; }
}

template <int RNG_MODE>
VCRT_HD float rng_next(Rng& g) {
    if (RNG_MODE == VCRT_RNG_PCG_REF) {
        g.pcg = g.pcg * 747796405u + 1u;
        uint32_t s = g.pcg;
        uint32_t word = ((s >> ((s >> 28) + 4u)) ^ s) * 277803737u;
        word = (word >> 22) ^ word;
        // float(word) / 4294967295.0f with float(2^32-1) == 2^32: exact scaling, range [0,1] inclusive
#ifdef __CUDA_ARCH__
        return __uint2float_rn(word) * 2.3283064365386963e-10f;
#else
        return (float)word * 2.3283064365386963e-10f;
#endif
    } else {
        if (g.have == 0u) {
            philox4x32_10(g.ctr0, g.ctr1, g.ctr2, 0u, g.key0, g.key1, g.buf);
            g.ctr2++;
            g.have = 4u;
        }
        uint32_t w = g.have == 4u ? g.buf[0] : g.have == 3u ? g.buf[1] : g.have == 2u ? g.buf[2] : g.buf[3];
        g.have--;
        return (float)(w >> 8) * 5.9604644775390625e-8f;
    }
}

// ------------------------------------------------------------------------------------------ scene view
// Device pointers to the bound buffers in the reference's own layouts, read as 128-bit words:
//   triangle (48 B) = {v0.xyz,-} {v1.xyz,-} {v2.xyz, materialIndex}          GpuModels.h:32-38
//   bvhNode  (48 B) = {min.xyz,-} {max.xyz, left} {right, object, -, -}      GpuModels.h:47-54
//   material (32 B) = {type,-,-,-} {albedo.xyz,-}                            GpuModels.h:26-30
//   sphere   (32 B) = {c.xyz, r} {materialIndex,-,-,-}                       GpuModels.h:40-44
struct SceneView {
    const float4* tris;
    const float4* mats;
    const float4* bvh;
    const vcrt_light* lights;
    const float4* spheres;
    uint32_t ntris, nmats, nbvh, nlights, nspheres;
    // repacked records for the fast traversal (vcrt_fast.cuh); null until built
    const float4* fnodes;   // 64 B per inner node
    const float4* ftris;    // 64 B per triangle, leaf order: {v0, original index} {a, materialIndex} {b, -} {n, -}
    uint32_t nfnodes;
    int32_t froot;          // >= 0 inner node, < 0 leaf (~triangle slot), INT_MIN empty
    // 32-byte quantised inner nodes (vcrt_repack.h: quantize_fast_bvh); null = walk the 64-byte float nodes
    const Words8* qnodes;
    const Words8* q4nodes;  // 4-wide quantised nodes (two Words8 per node) or null; root = froot4
    int32_t froot4;
    float3 qorg, qext;      // coordinate = qorg + 2m * qext
};

struct Ray { float3 o, d; };
struct Hit { float3 p, normal; uint32_t materialIndex; float t; int backFaceInt; int triangle; };

struct TraceStats { uint32_t rays, nodes, tris, prim; };   // prim: bounce-0 rays among `rays`

// Out-of-range reads return zero (robustBufferAccess), like oracle/_ref's Ssbo::operator[].
VCRT_HD void load_tri(const SceneView& s, uint32_t i, float3& v0, float3& v1, float3& v2, uint32_t& mat) {
    if (i >= s.ntris) { v0 = v1 = v2 = f3(0, 0, 0); mat = 0u; return; }
    float4 a = ldg4(s.tris + 3 * (size_t)i), b = ldg4(s.tris + 3 * (size_t)i + 1), c = ldg4(s.tris + 3 * (size_t)i + 2);
    v0 = xyz(a); v1 = xyz(b); v2 = xyz(c); mat = f2u(c.w);
}
VCRT_HD void load_mat(const SceneView& s, uint32_t i, uint32_t& type, float3& albedo) {
    if (i >= s.nmats) { type = 0u; albedo = f3(0, 0, 0); return; }
    float4 a = ldg4(s.mats + 2 * (size_t)i), b = ldg4(s.mats + 2 * (size_t)i + 1);
    type = f2u(a.x); albedo = xyz(b);
}

// ------------------------------------------------------------------------------------------ intersection
// hit_triangle + triIntersect, ray-trace-compute.comp:205-220, :157-173.  Returns the u/v verdict and t;
// the hit record is completed by finish_triangle_hit only for the accepted triangle (same values: they depend
// on the triangle and the ray only).
VCRT_HD bool tri_test(float3 v0, float3 v1, float3 v2, const Ray& r, float& t) {
    float3 a = sub(v0, v1), b = sub(v2, v0), p = sub(v0, r.o);
    float3 n = cross(b, a);
    float3 q = cross(p, r.d);
    float idet = 1.0f / dot(r.d, n);
    float u = dot(q, b) * idet, v = dot(q, a) * idet;
    t = dot(n, p) * idet;
    return !(u < 0.0f || u > 1.0f || v < 0.0f || (u + v) > 1.0f);
}

// The same test on a repacked record that carries a = v0 - v1, b = v2 - v0, n = cross(b, a) (vcrt_repack.h:
// precompute_triangles): identical operations, the ray-independent ones done once at upload.
VCRT_HD bool tri_test_pre(float3 v0, float3 a, float3 b, float3 n, const Ray& r, float& t) {
    float3 p = sub(v0, r.o);
    float3 q = cross(p, r.d);
    float idet = 1.0f / dot(r.d, n);
    float u = dot(q, b) * idet, v = dot(q, a) * idet;
    t = dot(n, p) * idet;
    return !(u < 0.0f || u > 1.0f || v < 0.0f || (u + v) > 1.0f);
}

VCRT_HD void finish_triangle_hit_pre(float3 n, uint32_t mat, int tri, const Ray& r, float t, Hit& rec) {
    rec.p = add(r.o, scale(r.d, t));
    rec.normal = normalize(n);
    rec.backFaceInt = dot(r.d, rec.normal) > 0.0f ? 1 : 0;
    rec.normal = scale(rec.normal, (float)(1 - 2 * rec.backFaceInt));
    rec.p = add(rec.p, scale(rec.normal, 0.0001f));
    rec.t = t;
    rec.materialIndex = mat;
    rec.triangle = tri;
}

VCRT_HD void finish_triangle_hit(float3 v0, float3 v1, float3 v2, uint32_t mat, int tri, const Ray& r, float t, Hit& rec) {
    float3 a = sub(v0, v1), b = sub(v2, v0);
    float3 n = cross(b, a);
    rec.p = add(r.o, scale(r.d, t));
    rec.normal = normalize(n);
    rec.backFaceInt = dot(r.d, rec.normal) > 0.0f ? 1 : 0;
    rec.normal = scale(rec.normal, (float)(1 - 2 * rec.backFaceInt));
    rec.p = add(rec.p, scale(rec.normal, 0.0001f));
    rec.t = t;
    rec.materialIndex = mat;
    rec.triangle = tri;
}

// hit_sphere, ray-trace-compute.comp:175-203
VCRT_HD bool hit_sphere(const SceneView& s, uint32_t si, const Ray& r, float tMin, float tMax, Hit& rec) {
    float4 c4 = ldg4(s.spheres + 2 * (size_t)si), m4 = ldg4(s.spheres + 2 * (size_t)si + 1);
    float3 center = xyz(c4);
    float radius = c4.w;
    float3 oc = sub(r.o, center);
    float a = dot(r.d, r.d), half_b = dot(oc, r.d), c = dot(oc, oc) - radius * radius;
    float disc = half_b * half_b - a * c;
    if (disc < 0.0f) return false;
    float sqrtd = sqrtf(disc);
    rec.backFaceInt = 0;
    float root = (-half_b - sqrtd) / a;
    if (root < tMin || tMax < root) {
        root = (-half_b + sqrtd) / a;
        rec.backFaceInt = 1;
        if (root < tMin || tMax < root) return false;
    }
    rec.t = root;
    rec.p = add(r.o, scale(r.d, root));
    float3 d = scale(sub(rec.p, center), (float)(1 - 2 * rec.backFaceInt));
    rec.normal = f3(d.x / radius, d.y / radius, d.z / radius);
    rec.materialIndex = f2u(m4.x);
    rec.triangle = -2 - (int)si;
    return true;
}

#define VCRT_T_MIN 0.001f
#define VCRT_T_MAX 10000.0f

// hit_scene, ray-trace-compute.comp:222-247 (brute force; the sphere loop restarts from t_max).  The triangle loop runs to
// ubo.numTriangles (:229), not to the buffer length: reads past the bound buffer return the zero triangle (never hit).
// The simple shader's hit_scene (ray-trace-compute-simple.comp:106-123) has no sphere loop.
template <int SHADER>
VCRT_HD bool hit_scene(const SceneView& s, uint32_t num_triangles, const Ray& r, Hit& rec, TraceStats& st) {
    bool any = false;
    float closest = VCRT_T_MAX;
    int best = -1;
    for (uint32_t i = 0; i < num_triangles; ++i) {
        float3 v0, v1, v2; uint32_t mat; float t;
        load_tri(s, i, v0, v1, v2, mat);
        st.tris++;
        if (tri_test(v0, v1, v2, r, t) && t > VCRT_T_MIN && t < closest) { any = true; closest = t; best = (int)i; }
    }
    if (best >= 0) {
        float3 v0, v1, v2; uint32_t mat;
        load_tri(s, (uint32_t)best, v0, v1, v2, mat);
        finish_triangle_hit(v0, v1, v2, mat, best, r, closest, rec);
    }
    if (SHADER == VCRT_SHADER_SIMPLE) return any;
    closest = VCRT_T_MAX;
    for (uint32_t j = 0; j < s.nspheres; ++j) {
        Hit tmp;
        if (hit_sphere(s, j, r, VCRT_T_MIN, closest, tmp)) { any = true; closest = tmp.t; rec = tmp; }
    }
    return any;
}

// intersectAABB, ray-trace-compute.comp:250-258 -- true divisions, glm min/max NaN behaviour
VCRT_HD void intersect_aabb_ref(const Ray& r, float3 bmin, float3 bmax, float& tNear, float& tFar) {
    float3 tMin = f3((bmin.x - r.o.x) / r.d.x, (bmin.y - r.o.y) / r.d.y, (bmin.z - r.o.z) / r.d.z);
    float3 tMax = f3((bmax.x - r.o.x) / r.d.x, (bmax.y - r.o.y) / r.d.y, (bmax.z - r.o.z) / r.d.z);
    float3 t1 = f3(glsl_min(tMin.x, tMax.x), glsl_min(tMin.y, tMax.y), glsl_min(tMin.z, tMax.z));
    float3 t2 = f3(glsl_max(tMin.x, tMax.x), glsl_max(tMin.y, tMax.y), glsl_max(tMin.z, tMax.z));
    tNear = glsl_max(glsl_max(t1.x, t1.y), t1.z);
    tFar = glsl_min(glsl_min(t2.x, t2.y), t2.z);
}

#define VCRT_MAX_STACK 64

// hit_bvh, ray-trace-compute.comp:263-311, literally: right child popped first, no t-culling, -1 children pushed,
// loop ends (dropping the stack) when stackIndex reaches `depth` (MAX_STACK_DEPTH quirk).
VCRT_HD bool hit_bvh_reference(const SceneView& s, const Ray& r, Hit& rec, int depth, TraceStats& st) {
    float closest = VCRT_T_MAX;
    int best = -1;
    int stack[VCRT_MAX_STACK];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0 && sp < depth) {
        sp--;
        int cur = stack[sp];
        if (cur == -1) continue;
        float3 bmin = f3(0, 0, 0), bmax = f3(0, 0, 0);
        int left = 0, right = 0, obj = 0;
        if ((uint32_t)cur < s.nbvh) {
            float4 a = ldg4(s.bvh + 3 * (size_t)cur), b = ldg4(s.bvh + 3 * (size_t)cur + 1), c = ldg4(s.bvh + 3 * (size_t)cur + 2);
            bmin = xyz(a); bmax = xyz(b);
            left = (int)f2u(b.w); right = (int)f2u(c.x); obj = (int)f2u(c.y);
        }
        st.nodes++;
        float tN, tF;
        intersect_aabb_ref(r, bmin, bmax, tN, tF);
        if (tN > tF) continue;
        if (obj != -1) {
            float3 v0, v1, v2; uint32_t mat; float t;
            load_tri(s, (uint32_t)obj, v0, v1, v2, mat);
            st.tris++;
            if (tri_test(v0, v1, v2, r, t) && t > VCRT_T_MIN && t < closest) { closest = t; best = obj; }
        }
        stack[sp++] = left;
        stack[sp++] = right;
    }
    if (best < 0) return false;
    float3 v0, v1, v2; uint32_t mat;
    load_tri(s, (uint32_t)best, v0, v1, v2, mat);
    finish_triangle_hit(v0, v1, v2, mat, best, r, closest, rec);
    return true;
}

// ------------------------------------------------------------------------------------------ shading
struct ShadeEnv {
    uint32_t lights_length;
    uint32_t max_bounces;
    int stack_depth;
};

// Onb + random_cosine_direction + sampleLambertian: definitions.glsl:42-53, random.glsl:42-52, ray-trace-compute.comp:91-99
template <int RNG_MODE, int TRIG>
VCRT_HD float3 sample_lambertian(float3 normal, Rng& g) {
    float3 w = normalize(normal);
    float3 a = (fabsf(w.x) > 0.9f) ? f3(0, 1, 0) : f3(1, 0, 0);
    float3 v = normalize(cross(w, a));
    float3 u = cross(w, v);
    float r1 = rng_next<RNG_MODE>(g), r2 = rng_next<RNG_MODE>(g);
    float z = sqrtf(1.0f - r2);
    float phi = 2.0f * 3.1415926535897932385f * r1;
    float sn, cs;
    if (TRIG == VCRT_TRIG_PORTABLE) sincos_portable(phi, &sn, &cs);
    else { cs = cosf(phi); sn = sinf(phi); }
    float sr2 = sqrtf(r2);
    float x = cs * sr2, y = sn * sr2;
    return normalize(add(add(scale(u, x), scale(v, y)), scale(w, z)));
}

// sampleGlass, ray-trace-compute.comp:106-116
VCRT_HD float3 sample_glass(float3 I, const Hit& rec) {
    float ir = 1.5f;
    float ratio = (float)(1 - rec.backFaceInt) * 1.0f / ir + (float)rec.backFaceInt * ir;
    float3 i = normalize(I);
    float cos_theta = glsl_min(dot(neg(i), rec.normal), 1.0f);
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    float t = floorf(glsl_min(glsl_max(ratio * sin_theta, 0.0f), 1.0f));
    return add(scale(reflect(i, rec.normal), t), scale(refract(i, rec.normal, ratio), 1.0f - t));
}

// sampleLight + randomOnATriangle, ray-trace-compute.comp:67-89
template <int RNG_MODE>
VCRT_HD float3 sample_light(const SceneView& s, const ShadeEnv& e, float3 p, Rng& g, float& lightCosine) {
    int lightIndex = (int)floorf((float)(int)e.lights_length * rng_next<RNG_MODE>(g));
    uint32_t ti = 0u;
    if (lightIndex >= 0 && (uint32_t)lightIndex < s.nlights) ti = s.lights[lightIndex].triangleIndex;
    float sa = rng_next<RNG_MODE>(g), tb = rng_next<RNG_MODE>(g);
    float3 v0, v1, v2; uint32_t mat;
    load_tri(s, ti, v0, v1, v2, mat);
    float3 v01 = add(neg(v0), v1), v02 = add(neg(v0), v2);
    float3 onLight = add(add(v0, scale(v01, sa)), scale(v02, tb));
    float3 toLight = normalize(sub(onLight, p));
    lightCosine = fabsf(toLight.y);
    return toLight;
}

// scatter, ray-trace-compute.comp:118-155; SHADER_SIMPLE: ray-trace-compute-simple.comp:62-68 + random.glsl:24-40
template <int SHADER, int RNG_MODE, int TRIG>
VCRT_HD bool scatter(const SceneView& s, const ShadeEnv& e, const Ray& r_in, const Hit& rec, float3& albedo, Ray& scattered, Rng& g) {
    uint32_t type;
    load_mat(s, rec.materialIndex, type, albedo);
    if (SHADER == VCRT_SHADER_SIMPLE) {
        float lo = -0.3f, hi = 0.3f;
        float px = lo + (hi - lo) * rng_next<RNG_MODE>(g);
        float py = lo + (hi - lo) * rng_next<RNG_MODE>(g);
        float pz = lo + (hi - lo) * rng_next<RNG_MODE>(g);
        float3 p = normalize(f3(px, py, pz));
        if (!(dot(p, rec.normal) > 0.0f)) p = neg(p);
        scattered.o = rec.p;
        scattered.d = normalize(p);
        return type == VCRT_MAT_LIGHT;
    }
    float3 materialSample = f3(0, 0, 0);
    if (type == VCRT_MAT_LAMBERTIAN) materialSample = sample_lambertian<RNG_MODE, TRIG>(rec.normal, g);
    else if (type == VCRT_MAT_METAL) materialSample = reflect(r_in.d, rec.normal);
    else if (type == VCRT_MAT_GLASS) { materialSample = sample_glass(r_in.d, rec); albedo = f3(1.0f, 1.0f, 1.0f); }
    float3 finalSample = materialSample;
    float coin = rng_next<RNG_MODE>(g);  // always drawn (:138)
    if (coin < 0.5f && type == VCRT_MAT_LAMBERTIAN) {
        float lightCosine;
        finalSample = sample_light<RNG_MODE>(s, e, rec.p, g, lightCosine);
        if (fabsf(lightCosine) < 0.001f) finalSample = materialSample;
    }
    scattered.o = rec.p;
    scattered.d = finalSample;
    return type == VCRT_MAT_LIGHT;
}

// ------------------------------------------------------------------------------------------ camera + accumulation
// main(), ray-trace-compute.comp:352-373.  The host evaluates the scalar prologue (tanf via libm, identical to
// the oracle's) and passes the resulting vectors; the per-pixel arithmetic below is the shader's.
struct Camera {
    float3 origin, llc;          // origin = camPos.zxy * (-1,1,1); lower_left_corner
    float viewport_width, viewport_height;
    float imW, imH;
};

VCRT_HD Ray primary_ray(const Camera& c, uint32_t x, uint32_t y) {
    float u = (float)x / c.imW, v = (float)y / c.imH;
    float3 horizontal = f3(c.viewport_width, 0.0f, 0.0f);
    float3 vertical = f3(0.0f, -c.viewport_height, 0.0f);
    Ray r;
    r.o = c.origin;
    r.d = sub(add(add(c.llc, scale(horizontal, u)), scale(vertical, v)), c.origin);
    return r;
}

VCRT_HD uint8_t unorm8(float f) {
    if (!(f == f)) return 0;
    f = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);
    return (uint8_t)rintf(f * 255.0f);
}

// One step of the rgba8 running mean: ray-trace-compute.comp:375-379 followed by the host's
// target -> accumulation copy (main.cpp:253-261).  `px` holds the accumulation texel and receives the new one.
VCRT_HD void running_mean_rgba8(uchar4& px, float3 c, uint32_t sample) {
    float fs = (float)sample;
    float m = glsl_min(fs, 1.0f);
    float col[4] = {c.x, c.y, c.z, 1.0f};
    uint8_t in[4] = {px.x, px.y, px.z, px.w}, out[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        float cur = ((float)in[ch] / 255.0f) * m;
        out[ch] = unorm8((col[ch] + cur * fs) / (fs + 1.0f));
    }
    px = make_uchar4(out[0], out[1], out[2], out[3]);
}

}  // namespace vcrt
