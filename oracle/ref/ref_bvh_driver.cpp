// TEST INFRASTRUCTURE ONLY -- the reference's own Bvh::createBvh (src/ray-tracing/Bvh.h:141-209, unmodified,
// included from where it lies) behind a C entry point, so that the product's builder can be compared with it on
// arbitrary triangle sets.  srand(seed) first: the reference never seeds, i.e. it sees glibc's srand(1) sequence.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <ray-tracing/Bvh.h>

extern "C" int ref_create_bvh(const void* triangles48, uint32_t n, uint32_t seed, void* nodes48, uint32_t capacity) {
    static_assert(sizeof(GpuModel::Triangle) == 48 && sizeof(GpuModel::BvhNode) == 48, "GpuModels.h layouts");
    std::vector<Bvh::Object0> objects(n);
    const GpuModel::Triangle* t = (const GpuModel::Triangle*)triangles48;
    for (uint32_t i = 0; i < n; ++i) { objects[i].index = i; objects[i].t = t[i]; }
    srand(seed ? seed : 1u);
    std::vector<GpuModel::BvhNode> out = Bvh::createBvh(objects);
    if (out.size() > capacity) return -1;
    std::memset(nodes48, 0, out.size() * 48);
    GpuModel::BvhNode* o = (GpuModel::BvhNode*)nodes48;
    for (size_t i = 0; i < out.size(); ++i) {
        o[i].min = out[i].min; o[i].max = out[i].max;
        o[i].leftNodeIndex = out[i].leftNodeIndex; o[i].rightNodeIndex = out[i].rightNodeIndex; o[i].objectIndex = out[i].objectIndex;
    }
    return (int)out.size();
}
