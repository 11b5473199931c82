// headless_main.cpp -- the reference's application (src/main.cpp) without the window: same scene set-up calls
// (initScene, main.cpp:76-153), same per-frame protocol (updateScene :166-183, computeCommand :228), rendering to a PPM
// instead of a swapchain.  Usage: headless_main scene.vcrt out.ppm [frames=16] [width=800] [height=600] [simple]
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "../include/vcrt/ComputeMaterial.hpp"

struct UniformBufferObject {   // main.cpp:39-47
    alignas(16) float camPosition[3];
    alignas(4) float time;
    alignas(4) uint32_t currentSample, numTriangles, numLights, numSpheres;
};
static_assert(sizeof(UniformBufferObject) == 32, "std140 UBO");

struct RtScene {   // the five arrays of GpuModel::Scene (RtScene.h:35-42), read from a .vcrt container
    std::vector<vcrt_triangle> triangles; std::vector<vcrt_material> materials; std::vector<vcrt_bvh_node> bvhNodes;
    std::vector<vcrt_light> lights; std::vector<vcrt_sphere> spheres;
    explicit RtScene(const std::string& path) {
        std::ifstream f(path, std::ios::binary);
        char magic[8]; uint32_t hdr[8];
        if (!f.read(magic, 8) || std::memcmp(magic, "VCRTSCN1", 8) || !f.read((char*)hdr, 32)) throw std::runtime_error("failed to open file: " + path);
        triangles.resize(hdr[0]); materials.resize(hdr[1]); bvhNodes.resize(hdr[2]); lights.resize(hdr[3]); spheres.resize(hdr[4]);
        f.read((char*)triangles.data(), 48ull * hdr[0]); f.read((char*)materials.data(), 32ull * hdr[1]); f.read((char*)bvhNodes.data(), 48ull * hdr[2]);
        f.read((char*)lights.data(), 8ull * hdr[3]); f.read((char*)spheres.data(), 32ull * hdr[4]);
        if (!f) throw std::runtime_error("failed to read scene: " + path);
    }
};

int main(int argc, char** argv) {
    try {
        using namespace mcvkp;
        if (argc < 3) { std::cerr << "usage: " << argv[0] << " scene.vcrt out.ppm [frames] [width] [height] [simple]\n"; return EXIT_FAILURE; }
        const int frames = argc > 3 ? atoi(argv[3]) : 16;
        const uint32_t W = argc > 4 ? atoi(argv[4]) : 800, H = argc > 5 ? atoi(argv[5]) : 600;   // VulkanApplicationContext.h:10-11
        const bool simple = argc > 6 && !strcmp(argv[6], "simple");
        const uint32_t descriptorSetsSize = 1;   // swapchain image count in the reference
        auto rtScene = std::make_shared<RtScene>(argv[1]);

        auto uniformBufferBundle = std::make_shared<BufferBundle>(descriptorSetsSize);
        BufferUtils::createBundle<UniformBufferObject>(uniformBufferBundle.get(), UniformBufferObject(), VK_BUFFER_USAGE_UNIFORM_BUFFER_BIT, VMA_MEMORY_USAGE_CPU_TO_GPU);
        auto triangleBufferBundle = std::make_shared<BufferBundle>(descriptorSetsSize);
        BufferUtils::createBundle<vcrt_triangle>(triangleBufferBundle.get(), rtScene->triangles.data(), rtScene->triangles.size(), VK_BUFFER_USAGE_STORAGE_BUFFER_BIT, VMA_MEMORY_USAGE_CPU_TO_GPU);
        auto materialBufferBundle = std::make_shared<BufferBundle>(descriptorSetsSize);
        BufferUtils::createBundle<vcrt_material>(materialBufferBundle.get(), rtScene->materials.data(), rtScene->materials.size(), VK_BUFFER_USAGE_STORAGE_BUFFER_BIT, VMA_MEMORY_USAGE_CPU_TO_GPU);
        auto aabbBufferBundle = std::make_shared<BufferBundle>(descriptorSetsSize);
        BufferUtils::createBundle<vcrt_bvh_node>(aabbBufferBundle.get(), rtScene->bvhNodes.data(), rtScene->bvhNodes.size(), VK_BUFFER_USAGE_STORAGE_BUFFER_BIT, VMA_MEMORY_USAGE_CPU_TO_GPU);
        auto lightsBufferBundle = std::make_shared<BufferBundle>(descriptorSetsSize);
        BufferUtils::createBundle<vcrt_light>(lightsBufferBundle.get(), rtScene->lights.data(), rtScene->lights.size(), VK_BUFFER_USAGE_STORAGE_BUFFER_BIT, VMA_MEMORY_USAGE_CPU_TO_GPU);
        auto spheresBufferBundle = std::make_shared<BufferBundle>(descriptorSetsSize);
        BufferUtils::createBundle<vcrt_sphere>(spheresBufferBundle.get(), rtScene->spheres.data(), rtScene->spheres.size(), VK_BUFFER_USAGE_STORAGE_BUFFER_BIT, VMA_MEMORY_USAGE_CPU_TO_GPU);

        auto accumulationTexture = std::make_shared<Image>(W, H);
        auto targetTexture = std::make_shared<Image>(W, H);

        auto computeMaterial = std::make_shared<ComputeMaterial>(simple ? "shaders/generated/ray-trace-compute-simple.spv" : "shaders/generated/ray-trace-compute.spv");
        computeMaterial->addUniformBufferBundle(uniformBufferBundle, VK_SHADER_STAGE_COMPUTE_BIT);
        computeMaterial->addStorageImage(targetTexture, VK_SHADER_STAGE_COMPUTE_BIT);
        computeMaterial->addStorageImage(accumulationTexture, VK_SHADER_STAGE_COMPUTE_BIT);
        computeMaterial->addStorageBufferBundle(triangleBufferBundle, VK_SHADER_STAGE_COMPUTE_BIT);
        computeMaterial->addStorageBufferBundle(materialBufferBundle, VK_SHADER_STAGE_COMPUTE_BIT);
        computeMaterial->addStorageBufferBundle(aabbBufferBundle, VK_SHADER_STAGE_COMPUTE_BIT);
        computeMaterial->addStorageBufferBundle(lightsBufferBundle, VK_SHADER_STAGE_COMPUTE_BIT);
        computeMaterial->addStorageBufferBundle(spheresBufferBundle, VK_SHADER_STAGE_COMPUTE_BIT);
        auto computeModel = std::make_shared<ComputeModel>(computeMaterial);

        VkCommandBuffer cmd = nullptr;
        uint32_t currentSample = 0;
        const float camera[3] = {1.8f, 8.6f, 1.1f};   // main.cpp:37
        auto t0 = std::chrono::steady_clock::now();
        for (int frame = 0; frame < frames; ++frame) {
            // updateScene, main.cpp:166-183
            UniformBufferObject ubo = {{camera[0], camera[1], camera[2]}, 0.0f, currentSample, (uint32_t)rtScene->triangles.size(), (uint32_t)rtScene->lights.size(), (uint32_t)rtScene->spheres.size()};
            auto& buffer = computeModel->getMaterial()->getUniformBufferBundles()[0].data->buffers[0];
            std::memcpy(buffer->map(), &ubo, sizeof(ubo));
            buffer->unmap();
            currentSample++;
            // main.cpp:228 (ceil-div instead of the reference's floor so the bottom rows are rendered too)
            computeModel->computeCommand(cmd, 0, (W + 31) / 32, (H + 31) / 32, 1);
        }
        std::vector<uint8_t> px = targetTexture->read();
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        printf("%f ms/frame\n", ms / frames);   // main.cpp:409
        std::ofstream out(argv[2], std::ios::binary);
        out << "P6\n" << W << " " << H << "\n255\n";
        for (size_t i = 0; i < (size_t)W * H; ++i) out.write((const char*)&px[4 * i], 3);
    } catch (const std::exception& e) {   // main.cpp:449-457
        std::cerr << e.what() << std::endl;
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}
