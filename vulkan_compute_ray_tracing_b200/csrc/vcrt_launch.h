// vcrt_launch.h -- host-visible launchers of the render kernels (one per traversal mode / translation unit).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>
#include "vcrt_path.cuh"

#include "vcrt_tunables.h"

namespace vcrt {

// One set of wavefront queues (vcrt_wavefront.cuh); owned by the context.
struct WfQueues {
    float4* q[2];                 // 3 float4 per ray, two queues (ping-pong)
    uint2* hit;                   // per ray of the current queue (bounce 0: per work item): {t bits, winning slot or -1}
    float4* sample_color;         // per path of the batch: final colour (w unused)
    unsigned int* counts;         // [0],[1]: queue sizes  [2]: trace fetch counter
    uint32_t capacity;            // paths per batch
};

// Pipelines of one wavefront render: independent batches (disjoint pixels) run side by side, each on its own stream with its
// own queue set.  stream[0] is the render stream; the others fork from it and join it again through the events.
#define VCRT_MAX_PIPES 4
struct WfPipes {
    int n;
    cudaStream_t stream[VCRT_MAX_PIPES];
    WfQueues q[VCRT_MAX_PIPES];
    cudaEvent_t fork, join[VCRT_MAX_PIPES];
    cudaEvent_t before_accumulate;   // frames in flight: the accumulate kernel waits for this event (the previous frame's fold); nullptr = none
};

// CUDA events around every launch of the dominant kernel (wf_trace_kernel), recorded on the launching stream, so that the
// roofline's kernel duration is measured live in the timed region.  Events are pooled: no creation cost in steady state.
struct TraceTimer {
    struct Rec { cudaEvent_t first, second; bool primary; };
    std::vector<cudaEvent_t> pool;
    std::vector<Rec> pending;
    double primary_ms = 0.0;      // part of the harvested time that went into bounce-0 launches
    bool enabled = true;          // option "trace_timing": off = no events around the trace launches
    cudaError_t begin(cudaStream_t stream, cudaEvent_t* t0, cudaEvent_t* t1) {
        if (!enabled) return cudaSuccess;
        cudaEvent_t* ev[2] = {t0, t1};
        for (cudaEvent_t* e : ev) {
            if (!pool.empty()) { *e = pool.back(); pool.pop_back(); continue; }
            cudaError_t rc = cudaEventCreate(e);
            if (rc != cudaSuccess) return rc;
        }
        return cudaEventRecord(*t0, stream);
    }
    cudaError_t end(cudaStream_t stream, cudaEvent_t t0, cudaEvent_t t1, bool primary) {
        if (!enabled) return cudaSuccess;
        pending.push_back({t0, t1, primary});
        return cudaEventRecord(t1, stream);
    }
    // after a stream synchronise: total milliseconds and count of the pending launches; events go back to the pool
    void drain(double* ms, uint64_t* n) {
        static const bool dump = getenv("VCRT_TRACE_DUMP") != nullptr;   // development aid: every trace launch's duration on stderr
        for (auto& p : pending) {
            float f = 0.0f;
            if (cudaEventElapsedTime(&f, p.first, p.second) == cudaSuccess) { *ms += f; ++*n; if (p.primary) primary_ms += f; if (dump) fprintf(stderr, "trace launch %s %.3f ms\n", p.primary ? "primary" : "bounce", f); }
            pool.push_back(p.first); pool.push_back(p.second);
        }
        pending.clear();
    }
    // the same for the launches that have already finished, without a synchronise
    void harvest(double* ms, uint64_t* n) {
        size_t done = 0;
        while (done < pending.size() && cudaEventQuery(pending[done].second) == cudaSuccess) {
            float f = 0.0f;
            if (cudaEventElapsedTime(&f, pending[done].first, pending[done].second) == cudaSuccess) { *ms += f; ++*n; if (pending[done].primary) primary_ms += f; }
            pool.push_back(pending[done].first); pool.push_back(pending[done].second);
            ++done;
        }
        pending.erase(pending.begin(), pending.begin() + (long)done);
    }
    void destroy() {
        for (auto& p : pending) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
        for (cudaEvent_t e : pool) cudaEventDestroy(e);
        pending.clear(); pool.clear();
    }
};
// Wavefront pipeline of the fast traversal (vcrt_wavefront.cuh); queues are owned by the context.
cudaError_t launch_render_wavefront(const KernelArgs& a, int shader, int rng, int trig, bool count, const WfPipes& pipes, uint32_t* launches, TraceTimer* timer);
cudaError_t launch_render_reference(const KernelArgs& a, int shader, int rng, int trig, bool count, cudaStream_t stream);
cudaError_t launch_render_fast(const KernelArgs& a, int shader, int rng, int trig, bool count, cudaStream_t stream);
cudaError_t launch_render_brute(const KernelArgs& a, int shader, int rng, int trig, bool count, cudaStream_t stream);
cudaError_t launch_tiles(void* image, void* packed, int elem_bytes, bool pack, uint32_t W, uint32_t H, uint32_t rank, uint32_t count, uint32_t owned, cudaStream_t stream);
cudaError_t launch_unpack_all_tiles(void* image, const void* gathered, int elem_bytes, uint32_t W, uint32_t H, uint32_t world, uint32_t tiles_per_rank, cudaStream_t stream);
cudaError_t launch_post_process(const uchar4* tex, uchar4* out, uint32_t w, uint32_t h, float mix, float sigma, float kSigma, float threshold, float inv_gamma,
                                cudaStream_t stream);
cudaError_t launch_resolve(const float4* accumf, uchar4* target, uint32_t npix, float inv_total, float inv_gamma, cudaStream_t stream);
cudaError_t launch_frame_fold(const KernelArgs& a, const float4* sample, uchar4* image, float inv_total, float inv_gamma, cudaStream_t stream);
}  // namespace vcrt
