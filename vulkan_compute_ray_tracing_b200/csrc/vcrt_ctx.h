// vcrt_ctx.h -- the context behind the C ABI's opaque vcrt_ctx (private to csrc/: vcrt_api.cu, vcrt_group.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/vcrt.h"
#include "vcrt_launch.h"

struct DevBuf {
    void* ptr = nullptr;
    size_t bytes = 0;      // bytes in use
    size_t capacity = 0;
};

// One frame in flight (vcrt_frame_submit): its own stream, sample buffer, rgba8 image and work counter, so that the render kernels of
// consecutive frames overlap; `folded` orders the folds of consecutive frames, `done` is the fence the host waits on (main.cpp:325).
struct FrameSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t folded = nullptr, done = nullptr;
    DevBuf image, sample;
    unsigned long long* counter = nullptr;
    bool busy = false;
};

struct vcrt_ctx {
    int device = 0;
    FrameSlot frame_slot[VCRT_MAX_FRAMES_IN_FLIGHT];
    int frames_n = 0;                     // frames in flight between vcrt_frames_begin and vcrt_frames_end (0: the synchronous interface is in use)
    uint32_t frame_next = 0;              // slot of the next vcrt_frame_submit
    cudaEvent_t last_folded = nullptr;    // `folded` of the most recently submitted frame
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    std::string error;
    int shader = VCRT_SHADER_FULL;
    DevBuf ssbo[8];                       // bindings 3..7 in the reference's layouts
    std::vector<uint8_t> host_tris, host_bvh;  // host copies for the host-side record build, fetched from the device when that build runs
    bool host_tris_valid = false, host_bvh_valid = false;
    int fast_build = 0;                   // option "fast_build": 0 "auto" (on the device when the scene allows, else on the host) | 1 "host" | 2 "device"
    bool auto_device = false;             // what "auto" picks when both builders apply: the host's binned-SAH tree costs 7 % fewer node visits per ray on the
                                          // C3 scene (11.7 vs 12.5) than the device's PLOC tree, which is built 30-50x faster
    bool built_on_device = false;         // where the current records came from
    bool have_binary = false;             // binary node records (fnodes / qnodes) exist: the megakernel and the q15 / f32 formats walk those
    double fast_build_ms = 0.0;           // wall time of the last record build
    std::string fast_build_stages;        // device builds: where that time went
    bool fast_dirty = true;
    uint32_t continue_threshold = 20;   // option "continue_threshold" (33 - min(leaf, shade) leaves the schedule unchanged)
    uint32_t leaf_threshold = 6, shade_threshold = 8;   // persistent-kernel phase thresholds (options "leaf_threshold", "shade_threshold")
    bool fast_sah = true;                 // option "fast_bvh": "sah" | "sah_plain" (rebuild the topology) | "topology" (keep the bound tree's)
    bool fast_reinsert = true;            // "sah": the rebuilt tree is optimised by reinsertion (vcrt_repack.cpp); "sah_plain": left as built
    int fast_nodes = 0;                   // option "fast_nodes": 0 "auto" (4-wide quantised when the scene extent allows) | 1 "q15" (binary) | 2 "f32" | 3 "q15x4"
    bool quantized = false, wide = false;
    int32_t froot4 = (int32_t)0x80000000;
    uint32_t nf4nodes = 0;
    DevBuf q4nodes;
    float qorg[3] = {0, 0, 0}, qext[3] = {0, 0, 0};
    DevBuf qnodes;
    uint32_t fast_depth = 0, bound_depth = 0;
    int dispatch_trav = 0;                // option "dispatch_traversal": 0 "auto" | 1 "reference" | 2 "fast"
    bool dispatch_fast = false;           // what the last vcrt_dispatch walked
    bool fast_ok = false;
    std::string fast_err;
    DevBuf fnodes, ftris;
    DevBuf wf_q0, wf_q1, wf_hit, wf_color, wf_counts;   // wavefront queues (all pipelines' sets, back to back)
    uint32_t wf_capacity = 0;              // paths per queue set
    int wf_sets = 0;                       // queue sets the buffers are currently carved into
    int wf_pipes_used = 0;                 // info "wf_pipelines": what the last wavefront render ran as
    int wf_streams = 0;                    // option "wf_streams": 0 "auto" | 1..VCRT_MAX_PIPES pipelines of a wavefront render
    cudaStream_t pipe_stream[VCRT_MAX_PIPES] = {nullptr, nullptr, nullptr, nullptr};   // [0] unused (the render stream)
    cudaEvent_t fork_ev = nullptr, join_ev[VCRT_MAX_PIPES] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> ev_pool;      // timing events of finished renders, reused
    uint32_t wf_batch = 256u << 20;        // option "wf_batch_paths": paths per wavefront batch (queue memory = 120 B per path, allocated for what a call needs).
                                           // Every trace launch ends in a ~110 us tail (the longest rays): C3 at 64 spp, 32 Mi / 64 Mi / one batch: 5395 / 5620 / 5763 Mrays/s
    int32_t froot = (int32_t)0x80000000;
    uint32_t nfnodes = 0;
    uint32_t W = 0, H = 0;
    DevBuf target, accum8, accumf, aov, present;
    vcrt_ubo ubo;
    unsigned long long* d_counters = nullptr;   // [0] queries [1] nodes [2] tris [3] work counter [4] traversals [5] bounce-0 queries (vcrt_kernels.inl: flush_stats)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    double kernel_ms = 0.0;
    uint64_t launches = 0;
    vcrt::TraceTimer trace_timer;
    int trace_timing = 0;                 // option "trace_timing": 0 "auto" (multi-sample renders only) | 1 "on" | 2 "off"
    double trace_ms = 0.0;
    uint64_t trace_launches = 0;
};


// shared helpers (vcrt_api.cu)
int vcrt_fail(vcrt_ctx* c, int code, const std::string& msg);
int vcrt_cuda_fail(vcrt_ctx* c, cudaError_t e, const char* what);
int vcrt_ensure(vcrt_ctx* c, DevBuf& b, size_t bytes, const char* what);
