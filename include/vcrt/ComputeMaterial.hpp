// ComputeMaterial.hpp -- header-only C++ facade over the C ABI (vcrt.h) with the reference's class and method names.
//
// A maintainer of the reference keeps main.cpp's call sites (main.cpp:84-153, :166-183, :228) and swaps the Vulkan-backed
// mcvkp::{BufferBundle, Image, ComputeMaterial, ComputeModel} for these; see INTEGRATION.md.  Everything Vulkan-specific in
// the signatures (VkCommandBuffer, VkShaderStageFlags, VkBufferUsageFlags, VmaMemoryUsage) is kept as an opaque placeholder
// so the calls read the same.  Errors: std::runtime_error("failed to ..."), as in the reference.
#pragma once

#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../vcrt.h"

namespace mcvkp {

using VkShaderStageFlags = uint32_t;
using VkBufferUsageFlags = uint32_t;
using VmaMemoryUsage = uint32_t;
struct CommandBufferStandIn {};
using VkCommandBuffer = CommandBufferStandIn*;   // commands execute in submission order on the context's CUDA stream
constexpr VkShaderStageFlags VK_SHADER_STAGE_COMPUTE_BIT = 0x20;
constexpr VkBufferUsageFlags VK_BUFFER_USAGE_UNIFORM_BUFFER_BIT = 0x10, VK_BUFFER_USAGE_STORAGE_BUFFER_BIT = 0x20;
constexpr VmaMemoryUsage VMA_MEMORY_USAGE_CPU_TO_GPU = 3;

// Buffer.h:15-40 -- host-visible memory the application maps and memcpy's into.
struct Buffer {
    std::vector<uint8_t> data;
    size_t size = 0;
    void* map() { return data.data(); }       // vmaMapMemory
    void unmap() {}                           // vmaUnmapMemory
};

// Buffer.h:42-53 -- one buffer per swapchain image; headless default is a bundle of 1.
struct BufferBundle {
    std::vector<std::shared_ptr<Buffer>> buffers;
    explicit BufferBundle(size_t bundleSize = 1) { for (size_t i = 0; i < bundleSize; ++i) buffers.push_back(std::make_shared<Buffer>()); }
};

namespace BufferUtils {
// BufferUtils::createBundle<T>(bundle, data, count, usage, memoryUsage), Buffer.h:95-108 (whole array, not sizeof(T): see SURVEY 8a A12)
template <typename T>
inline void createBundle(BufferBundle* bundle, const T* data, size_t count, VkBufferUsageFlags = 0, VmaMemoryUsage = 0) {
    for (auto& b : bundle->buffers) {
        b->size = sizeof(T) * count;
        b->data.resize(b->size);
        if (b->size) std::memcpy(b->data.data(), data, b->size);
    }
}
template <typename T>
inline void createBundle(BufferBundle* bundle, const T& value, VkBufferUsageFlags u = 0, VmaMemoryUsage m = 0) { createBundle<T>(bundle, &value, 1, u, m); }
}  // namespace BufferUtils

class ComputeMaterial;

// Image.h -- rgba8 storage image (ImageUtils::createImage, main.cpp:108-140); texels live in the material's context.
struct Image {
    uint32_t width = 0, height = 0;
    ComputeMaterial* owner = nullptr;
    int slot = -1;   // 0 = target (binding 1), 1 = accumulation (binding 2)
    Image() = default;
    Image(uint32_t w, uint32_t h) : width(w), height(h) {}
    inline std::vector<uint8_t> read() const;
};

template <typename T>
struct Descriptor { std::shared_ptr<T> data; VkShaderStageFlags shaderStageFlags; };   // Material.h:12-17

class ComputeMaterial {
public:
    // ComputeMaterial.cpp:9-13; `device` is the CUDA ordinal (the reference has one global VkDevice)
    explicit ComputeMaterial(const std::string& computeShaderPath, int device = 0) : m_computeShaderPath(computeShaderPath), m_device(device) {}
    ~ComputeMaterial() { if (m_ctx) vcrt_destroy(m_ctx); }                    // Material.cpp:21-29
    ComputeMaterial(const ComputeMaterial&) = delete;
    ComputeMaterial& operator=(const ComputeMaterial&) = delete;

    // Material.h:29-43
    void addStorageImage(const std::shared_ptr<Image>& image, VkShaderStageFlags f) { m_storageImageDescriptors.push_back({image, f}); }
    void addUniformBufferBundle(const std::shared_ptr<BufferBundle>& b, VkShaderStageFlags f) { m_uniformBufferBundleDescriptors.push_back({b, f}); }
    void addStorageBufferBundle(const std::shared_ptr<BufferBundle>& b, VkShaderStageFlags f) { m_storageBufferBundleDescriptors.push_back({b, f}); }
    const std::vector<Descriptor<BufferBundle>>& getUniformBufferBundles() const { return m_uniformBufferBundleDescriptors; }
    const std::vector<Descriptor<BufferBundle>>& getStorageBufferBundles() const { return m_storageBufferBundleDescriptors; }
    const std::vector<Descriptor<Image>>& getStorageImages() const { return m_storageImageDescriptors; }

    // ComputeMaterial.cpp:15-26: descriptor layout + pipeline + pool + sets  ->  context + kernel selection + uploads
    void init() {
        if (m_initialized) return;
        if (m_uniformBufferBundleDescriptors.size() != 1 || m_storageImageDescriptors.size() != 2 || m_storageBufferBundleDescriptors.size() != 5)
            throw std::runtime_error("failed to create compute pipeline layout: expected 1 uniform buffer, 2 storage images, 5 storage buffers");
        if (vcrt_create(m_device, &m_ctx) != VCRT_OK) throw std::runtime_error(vcrt_last_error(nullptr));
        check(vcrt_set_shader(m_ctx, m_computeShaderPath.c_str()));
        auto& target = *m_storageImageDescriptors[0].data;
        auto& accum = *m_storageImageDescriptors[1].data;
        if (target.width != accum.width || target.height != accum.height) throw std::runtime_error("failed to create descriptor sets: image sizes differ");
        check(vcrt_set_image_size(m_ctx, target.width, target.height));
        target.owner = this; target.slot = 0;
        accum.owner = this; accum.slot = 1;
        for (size_t i = 0; i < 5; ++i) {   // bindings 3..7 in insertion order (Material.cpp:258-311)
            const Buffer& b = *m_storageBufferBundleDescriptors[i].data->buffers[0];
            check(vcrt_set_buffer(m_ctx, 3 + (int)i, b.size ? b.data.data() : nullptr, b.size));
        }
        m_initialized = true;
    }

    // ComputeMaterial.cpp:63-68: the descriptor set of `currentFrame` points at buffers[currentFrame] of the uniform bundle
    void bind(VkCommandBuffer&, size_t currentFrame) {
        if (!m_initialized) throw std::runtime_error("failed to bind compute material: init() has not run");
        const Buffer& b = *m_uniformBufferBundleDescriptors[0].data->buffers.at(currentFrame);
        if (b.size != sizeof(vcrt_ubo)) throw std::runtime_error("failed to bind uniform buffer: expected 32 bytes");
        vcrt_ubo u;
        std::memcpy(&u, b.data.data(), sizeof u);
        check(vcrt_set_ubo(m_ctx, &u));
    }

    // Frames in flight (main.cpp:68 MAX_FRAMES_IN_FLIGHT; fences :298-316; vkWaitForFences :325; vkDeviceWaitIdle :419)
    void framesBegin(uint32_t framesInFlight) { requireInit(); check(vcrt_frames_begin(m_ctx, framesInFlight)); }
    void frameWait(uint32_t slot) { check(vcrt_frame_wait(m_ctx, slot)); }
    void framesEnd() { check(vcrt_frames_end(m_ctx)); }

    vcrt_ctx* context() { return m_ctx; }
    void requireInit() const { if (!m_initialized) throw std::runtime_error("failed to use compute material: init() has not run"); }
    void check(int rc) const { if (rc != VCRT_OK) throw std::runtime_error(vcrt_last_error(m_ctx)); }

private:
    std::string m_computeShaderPath;
    int m_device;
    vcrt_ctx* m_ctx = nullptr;
    bool m_initialized = false;
    std::vector<Descriptor<BufferBundle>> m_uniformBufferBundleDescriptors, m_storageBufferBundleDescriptors;
    std::vector<Descriptor<Image>> m_storageImageDescriptors;
};

inline std::vector<uint8_t> Image::read() const {
    if (!owner) throw std::runtime_error("failed to read image: not bound to an initialised ComputeMaterial");
    std::vector<uint8_t> out((size_t)width * height * 4);
    owner->check(slot == 0 ? vcrt_read_target_rgba8(owner->context(), out.data(), out.size()) : vcrt_read_accum_rgba8(owner->context(), out.data(), out.size()));
    return out;
}

// ComputeModel.h:12-22
class ComputeModel {
public:
    explicit ComputeModel(std::shared_ptr<ComputeMaterial> material) : m_material(std::move(material)) { m_material->init(); }   // ComputeModel.cpp:11-14
    std::shared_ptr<ComputeMaterial> getMaterial() { return m_material; }
    // bind + vkCmdDispatch(x, y, z) (ComputeModel.cpp:21-25) + the target -> accumulation copy of main.cpp:253-261
    void computeCommand(VkCommandBuffer& commandBuffer, size_t currentFrame, size_t x, size_t y, size_t z) {
        m_material->bind(commandBuffer, currentFrame);
        m_material->check(vcrt_dispatch(m_material->context(), (uint32_t)x, (uint32_t)y, (uint32_t)z));
    }
    // computeCommand as a frame in flight (between framesBegin and framesEnd): rendered on the next slot's stream, presented into
    // `presented` (W*H*4 bytes, ideally from vcrt_alloc_host; nullptr = no read-back); returns the slot whose fence guards it
    uint32_t frameCommand(VkCommandBuffer& commandBuffer, size_t currentFrame, size_t x, size_t y, size_t z, void* presented, size_t bytes) {
        m_material->bind(commandBuffer, currentFrame);
        uint32_t slot = 0;
        m_material->check(vcrt_frame_dispatch(m_material->context(), (uint32_t)x, (uint32_t)y, (uint32_t)z, presented, bytes, &slot));
        return slot;
    }
    // the same path with run-time parameters (sample loop, depth, RNG / accumulation mode, sharding)
    void renderCommand(VkCommandBuffer& commandBuffer, size_t currentFrame, const vcrt_render_params& params) {
        m_material->bind(commandBuffer, currentFrame);
        m_material->check(vcrt_render(m_material->context(), &params));
    }

private:
    std::shared_ptr<ComputeMaterial> m_material;
};

}  // namespace mcvkp
