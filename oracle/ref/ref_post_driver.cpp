// TEST INFRASTRUCTURE ONLY -- wraps the rewritten reference post-process fragment shader (SHADER_INC, from
// post-process-shader.frag) as a C entry point: one `Invocation` == one fragment of the full-screen quad the reference
// draws (mesh.cpp:58-93: positions (-1,-1)..(1,1), texture coordinates (0,0)..(1,1); post-process-shader.vert passes them
// through), i.e. fragTexCoord = (pixel centre) / (image size).  Built by oracle/ref/Makefile into oracle/_ref/libvcrt_ref.so,
// once as shipped (VARIANT = post) and once with the shader's own commented-out smartDeNoise line enabled (post_denoise).
#include "glsl_prelude.hpp"

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

namespace CAT(ns_, VARIANT) {
struct Invocation : InvocationBase {
    Invocation(const InvocationBase& b) : InvocationBase(b) {}
#include SHADER_INC
};
}  // namespace

using CAT(ns_, VARIANT)::Invocation;

// tex, out: rgba8 w x h.  out = the colour attachment as an rgba8 UNORM image (clamp, *255, round half to even).
extern "C" void CAT(ref_, VARIANT)(const uint8_t* tex, int w, int h, uint8_t* out) {
    Bindings b;
    std::memset(&b, 0, sizeof b);
    b.images[0] = image2D{const_cast<uint8_t*>(tex), w, h};
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            InvocationBase base{{(uint)x, (uint)y, 0u}, &b};
            Invocation inv(base);
            inv.fragTexCoord = vec2((float(x) + 0.5f) / float(w), (float(y) + 0.5f) / float(h));
            inv.main();
            uint8_t* c = out + 4 * (size_t(y) * w + x);
            c[0] = unorm8(inv.outColor.x); c[1] = unorm8(inv.outColor.y); c[2] = unorm8(inv.outColor.z); c[3] = unorm8(inv.outColor.w);
        }
    }
}
