// headless_main.cpp -- the reference's application (src/main.cpp) without the window: same scene set-up calls
// (initScene, main.cpp:76-153), same per-frame protocol (updateScene :166-183, computeCommand :228), rendering to a PPM
// instead of a swapchain.  Usage: headless_main scene.vcrt out.ppm [frames=16] [width=800] [height=600] [simple|full] [gpus=1]
// With gpus given, the same frames (samples 0..frames-1, f32 accumulation, Philox RNG, fast traversal) are rendered as ONE call
// sharded over that many GPUs by 32x32 tiles (vcrt_group_*: NCCL all-gather of the packed tiles inside the library); gpus=1
// takes the same code path on one GPU, and the two outputs are bit-identical.
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
        if (argc > 7) {   // ---- multi-GPU leg: the scene set-up calls of initScene once per GPU, through the group
            const int gpus = atoi(argv[7]);
            vcrt_group* group = nullptr;
            auto gcheck = [&](int rc) { if (rc != VCRT_OK) throw std::runtime_error(vcrt_group_last_error(group)); };
            gcheck(vcrt_group_create_local(gpus, nullptr, &group));
            gcheck(vcrt_group_set_shader(group, simple ? "ray-trace-compute-simple" : "ray-trace-compute"));
            gcheck(vcrt_group_set_image_size(group, W, H));
            gcheck(vcrt_group_set_buffer(group, VCRT_BINDING_TRIANGLES, rtScene->triangles.data(), rtScene->triangles.size() * sizeof(vcrt_triangle)));
            gcheck(vcrt_group_set_buffer(group, VCRT_BINDING_MATERIALS, rtScene->materials.data(), rtScene->materials.size() * sizeof(vcrt_material)));
            gcheck(vcrt_group_set_buffer(group, VCRT_BINDING_BVH, rtScene->bvhNodes.data(), rtScene->bvhNodes.size() * sizeof(vcrt_bvh_node)));
            gcheck(vcrt_group_set_buffer(group, VCRT_BINDING_LIGHTS, rtScene->lights.data(), rtScene->lights.size() * sizeof(vcrt_light)));
            gcheck(vcrt_group_set_buffer(group, VCRT_BINDING_SPHERES, rtScene->spheres.data(), rtScene->spheres.size() * sizeof(vcrt_sphere)));
            vcrt_ubo ubo = {{1.8f, 8.6f, 1.1f}, 0.0f, 0u, (uint32_t)rtScene->triangles.size(), (uint32_t)rtScene->lights.size(), (uint32_t)rtScene->spheres.size()};
            gcheck(vcrt_group_set_ubo(group, &ubo));
            vcrt_render_params p;
            std::memset(&p, 0, sizeof p);
            p.struct_size = sizeof p;
            p.shader = simple ? VCRT_SHADER_SIMPLE : VCRT_SHADER_FULL;
            p.traversal = VCRT_TRAVERSAL_FAST; p.rng_mode = VCRT_RNG_PHILOX; p.accum_mode = VCRT_ACCUM_F32; p.max_bounces = 8;
            p.sample_count = (uint32_t)frames;
            gcheck(vcrt_group_render(group, &p, VCRT_SHARD_TILES, 2.2f));   // warm-up: record builds, NCCL channels
            gcheck(vcrt_group_synchronize(group));
            auto g0 = std::chrono::steady_clock::now();
            gcheck(vcrt_group_render(group, &p, VCRT_SHARD_TILES, 2.2f));
            std::vector<uint8_t> px((size_t)W * H * 4), other((size_t)W * H * 4);
            gcheck(vcrt_group_read_target_rgba8(group, 0, px.data(), px.size()));
            double gms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g0).count();
            for (int i = 1; i < gpus; ++i) {   // tile sharding leaves the complete frame on every GPU
                gcheck(vcrt_group_read_target_rgba8(group, i, other.data(), other.size()));
                if (other != px) throw std::runtime_error("failed to gather tiles: GPU " + std::to_string(i) + " holds a different frame");
            }
            printf("%f ms/frame (%d samples on %d GPUs, tile-sharded)\n", gms, frames, gpus);
            std::ofstream gout(argv[2], std::ios::binary);
            gout << "P6\n" << W << " " << H << "\n255\n";
            for (size_t i = 0; i < (size_t)W * H; ++i) gout.write((const char*)&px[4 * i], 3);
            vcrt_group_destroy(group);
            return EXIT_SUCCESS;
        }

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
        // main.cpp:68, :298-316: two frames in flight, each slot behind a fence; what the reference presents to the swapchain lands
        // in the slot's page-locked buffer here
        const int MAX_FRAMES_IN_FLIGHT = 2;
        void* presented[MAX_FRAMES_IN_FLIGHT] = {nullptr, nullptr};
        const size_t frameBytes = (size_t)W * H * 4;
        for (auto& p : presented)
            if (vcrt_alloc_host(frameBytes, &p) != VCRT_OK) throw std::runtime_error(vcrt_last_error(nullptr));
        computeMaterial->framesBegin(MAX_FRAMES_IN_FLIGHT);
        size_t currentFrame = 0;
        auto t0 = std::chrono::steady_clock::now();
        for (int frame = 0; frame < frames; ++frame) {
            computeMaterial->frameWait((uint32_t)currentFrame);   // vkWaitForFences(inFlightFences[currentFrame]), main.cpp:325
            // updateScene, main.cpp:166-183
            UniformBufferObject ubo = {{camera[0], camera[1], camera[2]}, 0.0f, currentSample, (uint32_t)rtScene->triangles.size(), (uint32_t)rtScene->lights.size(), (uint32_t)rtScene->spheres.size()};
            auto& buffer = computeModel->getMaterial()->getUniformBufferBundles()[0].data->buffers[0];
            std::memcpy(buffer->map(), &ubo, sizeof(ubo));
            buffer->unmap();
            currentSample++;
            // main.cpp:228 (ceil-div instead of the reference's floor so the bottom rows are rendered too)
            computeModel->frameCommand(cmd, 0, (W + 31) / 32, (H + 31) / 32, 1, presented[currentFrame], frameBytes);
            currentFrame = (currentFrame + 1) % MAX_FRAMES_IN_FLIGHT;   // main.cpp:394
        }
        computeMaterial->framesEnd();   // vkDeviceWaitIdle, main.cpp:419
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        std::vector<uint8_t> px = targetTexture->read();
        if (frames > 0 && std::memcmp(px.data(), presented[(frames - 1) % MAX_FRAMES_IN_FLIGHT], frameBytes) != 0)
            throw std::runtime_error("failed to present: the last presented frame differs from the target image");
        for (auto& p : presented) vcrt_free_host(p);
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
