// vcrt_kernels.inl -- __global__ entry points and their dispatch table for ONE traversal mode.
// Included by vcrt_kernels_{ref,fast,brute}.cu with VCRT_TU_TRAV / VCRT_TU_NAME defined, so that the
// template instantiations of the three traversal modes compile in parallel.
#include <cuda_runtime.h>

#include "vcrt_path.cuh"
#include "vcrt_launch.h"

namespace vcrt {
__device__ __forceinline__ void flush_stats(const KernelArgs& a, const TraceStats& st);
}
#if VCRT_TU_TRAV == 1
#include "vcrt_persistent.cuh"
#include "vcrt_wavefront.cuh"
#endif

namespace vcrt {

__device__ __forceinline__ void flush_stats(const KernelArgs& a, const TraceStats& st) {
    unsigned rays = __reduce_add_sync(0xffffffffu, st.rays);
    unsigned nodes = __reduce_add_sync(0xffffffffu, st.nodes);
    unsigned tris = __reduce_add_sync(0xffffffffu, st.tris);
    if ((threadIdx.x & 31u) == 0u) {
        if (rays) atomicAdd(a.counters + 0, (unsigned long long)rays);
        if (nodes) atomicAdd(a.counters + 1, (unsigned long long)nodes);
        if (tris) atomicAdd(a.counters + 2, (unsigned long long)tris);
    }
}

// One thread per pixel (all of its samples), 32 consecutive threads = one 8x4 pixel block.
// This is the reference-shaped kernel: the parity anchor for every faster variant.
template <int SHADER, int TRAV, int RNG_MODE, int TRIG, bool COUNT>
__global__ void __launch_bounds__(VCRT_BLOCK) render_static_kernel(const __grid_constant__ KernelArgs a) {
    const uint32_t item = blockIdx.x * VCRT_BLOCK + threadIdx.x;
    TraceStats st = {0u, 0u, 0u};
    uint32_t x, y;
    if (item < a.owned_tiles * 1024u && item_to_pixel(a, item, x, y))
        render_pixel<SHADER, TRAV, RNG_MODE, TRIG, COUNT>(a, x, y, st);
    flush_stats(a, st);
}

template <int SHADER, int RNG_MODE, int TRIG, bool COUNT>
static cudaError_t launch_one(const KernelArgs& a, cudaStream_t stream) {
    const uint32_t items = a.owned_tiles * 1024u;
    if (items == 0u) return cudaSuccess;
#if VCRT_TU_TRAV == 1
    if (!(a.flags & VCRT_FLAG_STATIC_KERNEL)) {   // VCRT_FLAG_MEGAKERNEL
        // persistent warps: one resident wave, grid = SM count x resident blocks per SM
        static int grid = 0;
        if (grid == 0) {
            int dev = 0, sms = 0, per_sm = 0;
            cudaError_t e;
            if ((e = cudaGetDevice(&dev)) != cudaSuccess || (e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess ||
                (e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_persistent_kernel<SHADER, RNG_MODE, TRIG, COUNT>, VCRT_PBLOCK, 0)) != cudaSuccess)
                return e;
            grid = sms * (per_sm > 0 ? per_sm : 1);
        }
        const uint32_t needed = (items + VCRT_PBLOCK - 1) / VCRT_PBLOCK;
        render_persistent_kernel<SHADER, RNG_MODE, TRIG, COUNT><<<needed < (uint32_t)grid ? needed : (uint32_t)grid, VCRT_PBLOCK, 0, stream>>>(a);
        return cudaGetLastError();
    }
#endif
    const uint32_t blocks = (items + VCRT_BLOCK - 1) / VCRT_BLOCK;
    render_static_kernel<SHADER, VCRT_TU_TRAV, RNG_MODE, TRIG, COUNT><<<blocks, VCRT_BLOCK, 0, stream>>>(a);
    return cudaGetLastError();
}

template <int SHADER, int RNG_MODE, int TRIG>
static cudaError_t launch_count(const KernelArgs& a, bool count, cudaStream_t stream) {
#if VCRT_TU_TRAV == 1  /* VCRT_TRAVERSAL_FAST (an enum, invisible to the preprocessor) */
    if (count) return launch_one<SHADER, RNG_MODE, TRIG, true>(a, stream);
    return launch_one<SHADER, RNG_MODE, TRIG, false>(a, stream);
#else
    (void)count;
    return launch_one<SHADER, RNG_MODE, TRIG, true>(a, stream);
#endif
}

template <int SHADER, int RNG_MODE>
static cudaError_t launch_trig(const KernelArgs& a, int trig, bool count, cudaStream_t stream) {
    if (trig == VCRT_TRIG_PORTABLE) return launch_count<SHADER, RNG_MODE, VCRT_TRIG_PORTABLE>(a, count, stream);
    return launch_count<SHADER, RNG_MODE, VCRT_TRIG_LIBM>(a, count, stream);
}

template <int SHADER>
static cudaError_t launch_rng(const KernelArgs& a, int rng, int trig, bool count, cudaStream_t stream) {
    if (rng == VCRT_RNG_PHILOX) return launch_trig<SHADER, VCRT_RNG_PHILOX>(a, trig, count, stream);
    return launch_trig<SHADER, VCRT_RNG_PCG_REF>(a, trig, count, stream);
}

cudaError_t VCRT_TU_NAME(const KernelArgs& a, int shader, int rng, int trig, bool count, cudaStream_t stream) {
    if (shader == VCRT_SHADER_SIMPLE) return launch_rng<VCRT_SHADER_SIMPLE>(a, rng, trig, count, stream);
    return launch_rng<VCRT_SHADER_FULL>(a, rng, trig, count, stream);
}

#if VCRT_TU_TRAV == 1
// ---------------------------------------------------------------------------------------------- wavefront pipeline
__global__ void wf_reset_kernel(unsigned int* counts, int which) {
    counts[which] = 0u;   // the queue about to be filled
    counts[2] = 0u;       // the trace fetch counter
}

template <int SHADER, int RNG_MODE, int TRIG>
static cudaError_t wf_run(const KernelArgs& a, bool count, cudaStream_t stream, const WfQueues& w, uint32_t* launches, TraceTimer* timer) {
    // the trace variants: [COUNT][QN][PRIMARY]
    typedef void (*TraceFn)(const KernelArgs, const WfQueues, const WfBatch);
    static const TraceFn trace_fn[2][3][2] = {
        {{wf_trace_kernel<false, 0, false>, wf_trace_kernel<false, 0, true>}, {wf_trace_kernel<false, 1, false>, wf_trace_kernel<false, 1, true>}, {wf_trace_kernel<false, 2, false>, wf_trace_kernel<false, 2, true>}},
        {{wf_trace_kernel<true, 0, false>, wf_trace_kernel<true, 0, true>}, {wf_trace_kernel<true, 1, false>, wf_trace_kernel<true, 1, true>}, {wf_trace_kernel<true, 2, false>, wf_trace_kernel<true, 2, true>}}};
    static int trace_grid[2][3][2] = {{{0, 0}, {0, 0}, {0, 0}}, {{0, 0}, {0, 0}, {0, 0}}}, sms = 0;
    if (sms == 0) {
        int dev = 0, per_sm = 0;
        cudaError_t e;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess || (e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        for (int c = 0; c < 2; ++c)
            for (int q = 0; q < 3; ++q)
                for (int p = 0; p < 2; ++p) {
                    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_fn[c][q][p], VCRT_PBLOCK, 0)) != cudaSuccess) return e;
                    trace_grid[c][q][p] = sms * (per_sm > 0 ? per_sm : 1);   // persistent: one resident wave
                }
    }
    const int ci = count ? 1 : 0, qi = a.scene.q4nodes ? 2 : a.scene.qnodes ? 1 : 0;
    const uint32_t items = a.owned_tiles * 1024u;
    const uint32_t per_batch = w.capacity / a.sample_count;   // capacity >= sample_count is guaranteed by the caller
    const uint32_t shade_grid = (uint32_t)sms * VCRT_SHADE_GRID;
    for (uint32_t item0 = 0; item0 < items; item0 += per_batch) {
        WfBatch b;
        b.item0 = item0;
        b.nitems = items - item0 < per_batch ? items - item0 : per_batch;
        b.npaths = b.nitems * a.sample_count;
        b.cur = 0u; b.bounce = 0u;
        cudaError_t e;
        for (uint32_t bounce = 0; bounce < a.env.max_bounces; ++bounce) {
            b.bounce = bounce;
            const int pi = bounce == 0u ? 1 : 0;   // bounce 0 reads no queue: rays are generated from the path id
            wf_reset_kernel<<<1, 1, 0, stream>>>(w.counts, (int)(b.cur ^ 1u));
            cudaEvent_t t0 = nullptr, t1 = nullptr;
            if (timer && (e = timer->begin(stream, &t0, &t1)) != cudaSuccess) return e;
            trace_fn[ci][qi][pi]<<<trace_grid[ci][qi][pi], VCRT_PBLOCK, 0, stream>>>(a, w, b);
            if (timer && (e = timer->end(stream, t0, t1)) != cudaSuccess) return e;
            if (pi) wf_shade_kernel<SHADER, RNG_MODE, TRIG, true><<<shade_grid, 256, 0, stream>>>(a, w, b);
            else wf_shade_kernel<SHADER, RNG_MODE, TRIG, false><<<shade_grid, 256, 0, stream>>>(a, w, b);
            *launches += 3;
            b.cur ^= 1u;
        }
        wf_accumulate_kernel<<<(b.nitems + 255u) / 256u, 256, 0, stream>>>(a, w, b);
        ++*launches;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <int SHADER, int RNG_MODE>
static cudaError_t wf_trig(const KernelArgs& a, int trig, bool count, cudaStream_t stream, const WfQueues& w, uint32_t* launches, TraceTimer* timer) {
    if (trig == VCRT_TRIG_PORTABLE) return wf_run<SHADER, RNG_MODE, VCRT_TRIG_PORTABLE>(a, count, stream, w, launches, timer);
    return wf_run<SHADER, RNG_MODE, VCRT_TRIG_LIBM>(a, count, stream, w, launches, timer);
}

template <int SHADER>
static cudaError_t wf_rng(const KernelArgs& a, int rng, int trig, bool count, cudaStream_t stream, const WfQueues& w, uint32_t* launches, TraceTimer* timer) {
    if (rng == VCRT_RNG_PHILOX) return wf_trig<SHADER, VCRT_RNG_PHILOX>(a, trig, count, stream, w, launches, timer);
    return wf_trig<SHADER, VCRT_RNG_PCG_REF>(a, trig, count, stream, w, launches, timer);
}

cudaError_t launch_render_wavefront(const KernelArgs& a, int shader, int rng, int trig, bool count, cudaStream_t stream,
                                    float4* q0, float4* q1, uint2* hit, float4* sample_color, unsigned int* counts, uint32_t capacity, uint32_t* launches,
                                    TraceTimer* timer) {
    if (a.owned_tiles == 0u) return cudaSuccess;
    WfQueues w;
    w.q[0] = q0; w.q[1] = q1; w.hit = hit; w.sample_color = sample_color; w.counts = counts; w.capacity = capacity;
    if (shader == VCRT_SHADER_SIMPLE) return wf_rng<VCRT_SHADER_SIMPLE>(a, rng, trig, count, stream, w, launches, timer);
    return wf_rng<VCRT_SHADER_FULL>(a, rng, trig, count, stream, w, launches, timer);
}
#endif

}  // namespace vcrt
