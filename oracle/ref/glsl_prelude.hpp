// TEST INFRASTRUCTURE ONLY -- GLSL-on-C++ runtime for oracle/_ref (the reference's own compute
// shader text, rewritten syntactically by glsl2cpp.py, compiled with g++ against the reference's
// vendored glm 0.9.9.9).  Never linked into, imported by or shipped with the product.
//
// Semantics fixed here (GLSL leaves them implementation-defined; these are the canonical choices
// the restated oracle and the CUDA kernels follow):
//   * vector math = glm's formulas (normalize = v * (1/sqrt(dot)), reflect, refract, (y<x)?y:x min ...)
//   * sqrt / division IEEE-754 fp32, sin/cos/tan = libm sinf/cosf/tanf, no FMA contraction
//   * rgba8 imageLoad = c / 255.0f; imageStore = clamp -> *255 -> round-half-even; NaN -> 0
//   * out-of-range SSBO reads return zero (robustBufferAccess); buffer.length() is host-provided
//   * texture(sampler2D, uv) = the reference's sampler (Image.cpp:353-364: LINEAR mag/min filter, REPEAT addressing, normalised
//     coordinates, one mip level) evaluated as the Vulkan specification states it (texel coordinate u*W - 0.5, floor + fraction,
//     weights at 8 bits of sub-texel precision = VkPhysicalDeviceLimits::subTexelPrecisionBits of every desktop driver and of
//     lavapipe, the four taps combined in the order of the spec's formula), texels = c / 255.0f
#pragma once
#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>
#include <cmath>
#include <cstdint>
#include <cstring>

using namespace glm;

struct GlobalInvocationId {
    uint x, y, z;
    // uvec2 -> vec2 implicit conversion of GLSL (exact for pixel coordinates)
    vec2 xy() const { return vec2(float(x), float(y)); }
};

struct image2D {
    uint8_t* data;
    int w, h;
};

inline ivec2 imageSize_(const image2D& im) { return ivec2(im.w, im.h); }
inline vec4 imageLoad(const image2D& im, ivec2 p) {
    if (p.x < 0 || p.y < 0 || p.x >= im.w || p.y >= im.h) return vec4(0.0f);
    const uint8_t* c = im.data + 4 * (size_t(p.y) * im.w + p.x);
    return vec4(c[0] / 255.0f, c[1] / 255.0f, c[2] / 255.0f, c[3] / 255.0f);
}
inline uint8_t unorm8(float f) {
    if (!(f == f)) return 0;
    f = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);
    return (uint8_t)rintf(f * 255.0f);
}
inline void imageStore(const image2D& im, ivec2 p, vec4 v) {
    if (p.x < 0 || p.y < 0 || p.x >= im.w || p.y >= im.h) return;
    uint8_t* c = im.data + 4 * (size_t(p.y) * im.w + p.x);
    c[0] = unorm8(v.x); c[1] = unorm8(v.y); c[2] = unorm8(v.z); c[3] = unorm8(v.w);
}

struct sampler2D {
    const uint8_t* data;   // rgba8
    int w, h;
};
inline ivec2 textureSize(const sampler2D& s, int) { return ivec2(s.w, s.h); }
inline vec4 samplerTexel(const sampler2D& s, int x, int y) {   // VK_SAMPLER_ADDRESS_MODE_REPEAT
    x %= s.w; if (x < 0) x += s.w;
    y %= s.h; if (y < 0) y += s.h;
    const uint8_t* c = s.data + 4 * (size_t(y) * s.w + x);
    return vec4(c[0] / 255.0f, c[1] / 255.0f, c[2] / 255.0f, c[3] / 255.0f);
}
inline vec4 texture(const sampler2D& s, vec2 uv) {   // VK_FILTER_LINEAR
    const float u = uv.x * float(s.w) - 0.5f, v = uv.y * float(s.h) - 0.5f;
    const float fu = floorf(u), fv = floorf(v);
    const float a = rintf((u - fu) * 256.0f) / 256.0f, b = rintf((v - fv) * 256.0f) / 256.0f;
    const int i0 = (int)fu, j0 = (int)fv;
    return ((1.0f - a) * (1.0f - b)) * samplerTexel(s, i0, j0) + (a * (1.0f - b)) * samplerTexel(s, i0 + 1, j0) +
           ((1.0f - a) * b) * samplerTexel(s, i0, j0 + 1) + (a * b) * samplerTexel(s, i0 + 1, j0 + 1);
}

template <typename T>
struct Ssbo {
    const T* p;
    int count;    // elements really present (bounds for robust reads)
    int len;      // what buffer.length() reports (descriptor range / stride)
    int length() const { return len; }
    T operator[](long long i) const {
        if (i < 0 || i >= count) { T z; std::memset((void*)&z, 0, sizeof(T)); return z; }
        return p[i];
    }
};

// GLSL implicit int/uint -> float conversions that glm's templates do not deduce
inline float min(uint a, float b) { float x = float(a); return (b < x) ? b : x; }
inline float clamp(float x, int lo, int hi) { return glm::min(glm::max(x, float(lo)), float(hi)); }
inline vec3 operator/(const vec3& v, int s) { return v / float(s); }
inline vec3 operator*(int s, const vec3& v) { return float(s) * v; }
inline vec4 operator*(const vec4& v, uint s) { return v * float(s); }

struct Bindings {
    const void* ubo;
    image2D images[3];              // [1] target, [2] accumulation; the post-process pass samples images[0] (its binding 0)
    const void* ssbo[8];            // [3..7]
    int ssboCount[8];
    int ssboLen[8];
};

struct InvocationBase {
    GlobalInvocationId gl_GlobalInvocationID;
    const Bindings* B;
    template <typename U> U bindUbo(int) const { U u; static_assert(sizeof(U) == 32, "std140 UBO"); std::memcpy(&u, B->ubo, sizeof(U)); return u; }
    image2D bindImage(int b) const { return B->images[b]; }
    sampler2D bindSampler(int b) const { return sampler2D{B->images[b].data, B->images[b].w, B->images[b].h}; }
    template <typename T> Ssbo<T> bindSsbo(int b) const { return Ssbo<T>{(const T*)B->ssbo[b], B->ssboCount[b], B->ssboLen[b]}; }
};
