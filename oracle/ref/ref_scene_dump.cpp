// TEST INFRASTRUCTURE ONLY -- runs the reference's own, unmodified host scene code
// (src/ray-tracing/RtScene.h:44-101 -> Bvh.h:141-209, src/scene/mesh.cpp:96-139, tinyobjloader)
// and dumps the five buffers main.cpp:84-106 uploads, in the .vcrt scene container:
//   "VCRTSCN1" | u32 nTriangles nMaterials nBvhNodes nLights nSpheres | 3 x u32 reserved |
//   Triangle[48 B] | Material[32 B] | BvhNode[48 B] | Light[8 B] | Sphere[32 B]
// Padding bytes of the structs are zeroed so the dump is byte-reproducible.
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <ray-tracing/RtScene.h>

template <typename T> static void put(FILE* f, const std::vector<T>& v) { fwrite(v.data(), sizeof(T), v.size(), f); }

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s out.vcrt\n", argv[0]); return 2; }
    static_assert(sizeof(GpuModel::Triangle) == 48 && sizeof(GpuModel::Material) == 32 && sizeof(GpuModel::BvhNode) == 48 &&
                  sizeof(GpuModel::Light) == 8 && sizeof(GpuModel::Sphere) == 32, "GpuModels.h layouts");
    GpuModel::Scene s;
    // zero the padding: rebuild each record into zeroed storage
    std::vector<GpuModel::Triangle> tris(s.triangles.size());
    std::memset((void*)tris.data(), 0, tris.size() * sizeof(GpuModel::Triangle));
    for (size_t i = 0; i < tris.size(); ++i) { tris[i].v0 = s.triangles[i].v0; tris[i].v1 = s.triangles[i].v1; tris[i].v2 = s.triangles[i].v2; tris[i].materialIndex = s.triangles[i].materialIndex; }
    std::vector<GpuModel::Material> mats(s.materials.size());
    std::memset((void*)mats.data(), 0, mats.size() * sizeof(GpuModel::Material));
    for (size_t i = 0; i < mats.size(); ++i) { mats[i].type = s.materials[i].type; mats[i].albedo = s.materials[i].albedo; }
    std::vector<GpuModel::BvhNode> nodes(s.bvhNodes.size());
    std::memset((void*)nodes.data(), 0, nodes.size() * sizeof(GpuModel::BvhNode));
    for (size_t i = 0; i < nodes.size(); ++i) { nodes[i].min = s.bvhNodes[i].min; nodes[i].max = s.bvhNodes[i].max; nodes[i].leftNodeIndex = s.bvhNodes[i].leftNodeIndex; nodes[i].rightNodeIndex = s.bvhNodes[i].rightNodeIndex; nodes[i].objectIndex = s.bvhNodes[i].objectIndex; }
    std::vector<GpuModel::Sphere> sph(s.spheres.size());
    std::memset((void*)sph.data(), 0, sph.size() * sizeof(GpuModel::Sphere));
    for (size_t i = 0; i < sph.size(); ++i) { sph[i].s = s.spheres[i].s; sph[i].materialIndex = s.spheres[i].materialIndex; }

    FILE* f = fopen(argv[1], "wb");
    if (!f) { perror("fopen"); return 1; }
    uint32_t hdr[8] = {(uint32_t)tris.size(), (uint32_t)mats.size(), (uint32_t)nodes.size(), (uint32_t)s.lights.size(), (uint32_t)sph.size(), 0, 0, 0};
    fwrite("VCRTSCN1", 1, 8, f);
    fwrite(hdr, 4, 8, f);
    put(f, tris); put(f, mats); put(f, nodes); put(f, s.lights); put(f, sph);
    fclose(f);
    printf("triangles %u materials %u bvh %u lights %u spheres %u\n", hdr[0], hdr[1], hdr[2], hdr[3], hdr[4]);
    for (auto& l : s.lights) printf("light tri %u area %f\n", l.triangleIndex, l.area);
    printf("root min %f %f %f max %f %f %f children %d %d\n", nodes[0].min.x, nodes[0].min.y, nodes[0].min.z, nodes[0].max.x, nodes[0].max.y, nodes[0].max.z, nodes[0].leftNodeIndex, nodes[0].rightNodeIndex);
    return 0;
}
