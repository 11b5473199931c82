// vcrt_devbuild.h -- host-visible entry of the on-device record build (vcrt_devbuild.cu; algorithm: vcrt_devbuild.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <string>

namespace vcrt {
namespace devbuild {

struct Alloc {   // where the two outputs go (device memory owned by the context); return nullptr on failure
    std::function<void*(size_t bytes)> ftris, q4nodes;
};

struct Result {
    uint32_t nslots = 0;          // triangle slots (64 B each) written to the ftris allocation
    uint32_t nwide = 0;           // 4-wide nodes (64 B each) written to the q4nodes allocation; the root is node 0
    int32_t root4 = 0;
    float qorg[3] = {0, 0, 0}, qext[3] = {0, 0, 0};
    uint32_t depth = 0;           // depth of the binary PLOC tree
    uint32_t bound_depth = 0;     // deepest leaf of the bound bvh[]
    uint32_t stack4 = 0;          // traversal-stack entries the 4-wide tree can ask for
    uint32_t ploc_rounds = 0, wide_levels = 0;
    double ms_ranks = 0, ms_sort = 0, ms_ploc = 0, ms_total = 0;   // wall time since the start of the build at the end of steps A, B, C, D
};

// Builds the fast traversal's records from the bound buffers where they lie in device memory.  Returns 0 on success, 1 when
// this scene is not for the device builder (`why` says so: the caller falls back to the host builder, which also produces the
// precise error for malformed trees), -1 on a CUDA error.  Synchronises `stream` a few dozen times (level / round counts).
int run(const void* d_bvh, uint32_t nbvh, const void* d_tris, uint32_t ntris, float max_quantum, uint32_t max_stack, cudaStream_t stream, const Alloc& alloc, Result& out,
        std::string& why);

}  // namespace devbuild
}  // namespace vcrt
