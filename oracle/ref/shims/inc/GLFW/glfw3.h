// TEST INFRASTRUCTURE ONLY -- stand-in for <GLFW/glfw3.h> (+ Vulkan) so that the reference's
// unmodified src/scene/mesh.{h,cpp} compile headless.  Only the handful of Vulkan types that
// mesh.h/mesh.cpp name (vertex-input descriptions) are declared; none is used by the scene dump.
#pragma once
#include <cstdint>
#include <cstddef>
typedef enum { VK_VERTEX_INPUT_RATE_VERTEX = 0 } VkVertexInputRate;
typedef enum { VK_FORMAT_R32G32_SFLOAT = 103, VK_FORMAT_R32G32B32_SFLOAT = 106 } VkFormat;
typedef struct { uint32_t binding; uint32_t stride; VkVertexInputRate inputRate; } VkVertexInputBindingDescription;
typedef struct { uint32_t location; uint32_t binding; VkFormat format; uint32_t offset; } VkVertexInputAttributeDescription;
