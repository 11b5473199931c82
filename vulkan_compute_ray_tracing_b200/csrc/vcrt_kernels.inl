// vcrt_kernels.inl -- __global__ entry points and their dispatch table for ONE traversal mode.
// Included by vcrt_kernels_{ref,fast,brute}.cu with VCRT_TU_TRAV / VCRT_TU_NAME defined, so that the
// template instantiations of the three traversal modes compile in parallel.
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <utility>

#include "vcrt_path.cuh"
#include "vcrt_launch.h"

namespace vcrt {
__device__ __forceinline__ void flush_stats(const KernelArgs& a, const TraceStats& st, uint32_t queries_per_ray = 1u, bool all_primary = false);
}
#if VCRT_TU_TRAV == 1
#include "vcrt_persistent.cuh"
#include "vcrt_wavefront.cuh"
#endif

namespace vcrt {

// counters: [0] closest-hit queries answered  [1] node records fetched  [2] triangle records fetched  [3] work counter of the
// persistent kernels  [4] traversals run  [5] bounce-0 queries.  A traced ray answers `queries_per_ray` queries (the wavefront
// pipeline traces bounce 0 once per pixel for all of its samples); `all_primary`: every ray of this launch is a primary ray.
__device__ __forceinline__ void flush_stats(const KernelArgs& a, const TraceStats& st, uint32_t queries_per_ray, bool all_primary) {
    unsigned rays = __reduce_add_sync(0xffffffffu, st.rays);
    unsigned nodes = __reduce_add_sync(0xffffffffu, st.nodes);
    unsigned tris = __reduce_add_sync(0xffffffffu, st.tris);
    unsigned prim = __reduce_add_sync(0xffffffffu, all_primary ? st.rays : st.prim);
    if ((threadIdx.x & 31u) == 0u) {
        if (rays) { atomicAdd(a.counters + 0, (unsigned long long)rays * queries_per_ray); atomicAdd(a.counters + 4, (unsigned long long)rays); }
        if (nodes) atomicAdd(a.counters + 1, (unsigned long long)nodes);
        if (tris) atomicAdd(a.counters + 2, (unsigned long long)tris);
        if (prim) atomicAdd(a.counters + 5, (unsigned long long)prim * queries_per_ray);
    }
}

// One thread per pixel (all of its samples), 32 consecutive threads = one 8x4 pixel block.
// This is the reference-shaped kernel: the parity anchor for every faster variant.
template <int SHADER, int TRAV, int RNG_MODE, int TRIG, bool COUNT>
__global__ void __launch_bounds__(VCRT_BLOCK) render_static_kernel(const __grid_constant__ KernelArgs a) {
    const uint32_t item = blockIdx.x * VCRT_BLOCK + threadIdx.x;
    TraceStats st = {0u, 0u, 0u, 0u};
    uint32_t x, y;
    if (item < a.owned_tiles * 1024u && item_to_pixel(a, item, x, y))
        render_pixel<SHADER, TRAV, RNG_MODE, TRIG, COUNT>(a, x, y, st);
    flush_stats(a, st);
}

// Grid of a persistent kernel = one resident wave on the CURRENT device (SM count x resident blocks per SM), cached per
// (device, kernel) behind a mutex: contexts on different devices -- and host threads driving different contexts -- are
// independent (include/vcrt.h), so neither a process-wide static nor an unguarded one will do.  *sms_out: the SM count.
static cudaError_t persistent_grid(const void* fn, int block, int* grid, int* sms_out) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, std::pair<int, int>> cache;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({dev, fn});
    if (it == cache.end()) {
        int sms = 0, per_sm = 0;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess ||
            (e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block, 0)) != cudaSuccess)
            return e;
        it = cache.emplace(std::make_pair(dev, fn), std::make_pair(sms * (per_sm > 0 ? per_sm : 1), sms)).first;
    }
    *grid = it->second.first;
    if (sms_out) *sms_out = it->second.second;
    return cudaSuccess;
}

template <int SHADER, int RNG_MODE, int TRIG, bool COUNT>
static cudaError_t launch_one(const KernelArgs& a, cudaStream_t stream) {
    const uint32_t items = a.owned_tiles * 1024u;
    if (items == 0u) return cudaSuccess;
#if VCRT_TU_TRAV == 1
    if (!(a.flags & VCRT_FLAG_STATIC_KERNEL)) {   // VCRT_FLAG_MEGAKERNEL
        // persistent warps: one resident wave, grid = SM count x resident blocks per SM
        int grid = 0;
        cudaError_t e = persistent_grid((const void*)render_persistent_kernel<SHADER, RNG_MODE, TRIG, COUNT>, VCRT_PBLOCK, &grid, nullptr);
        if (e != cudaSuccess) return e;
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

template <int SHADER, int RNG_MODE, int TRIG>
static cudaError_t wf_run(const KernelArgs& a, bool count, const WfPipes& pipes, uint32_t* launches, TraceTimer* timer) {
    // the trace variants: [COUNT][QN][PRIMARY]
    typedef void (*TraceFn)(const KernelArgs, const WfQueues, const WfBatch);
    static const TraceFn trace_fn[2][3][2] = {
        {{wf_trace_kernel<false, 0, false>, wf_trace_kernel<false, 0, true>}, {wf_trace_kernel<false, 1, false>, wf_trace_kernel<false, 1, true>}, {wf_trace_kernel<false, 2, false>, wf_trace_kernel<false, 2, true>}},
        {{wf_trace_kernel<true, 0, false>, wf_trace_kernel<true, 0, true>}, {wf_trace_kernel<true, 1, false>, wf_trace_kernel<true, 1, true>}, {wf_trace_kernel<true, 2, false>, wf_trace_kernel<true, 2, true>}}};
    const int ci = count ? 1 : 0, qi = a.scene.q4nodes ? 2 : a.scene.qnodes ? 1 : 0;
    int trace_grid[2] = {0, 0}, sms = 0;   // [PRIMARY]; persistent: one resident wave on this device
    for (int p = 0; p < 2; ++p) {
        cudaError_t e = persistent_grid((const void*)trace_fn[ci][qi][p], VCRT_PBLOCK, &trace_grid[p], &sms);
        if (e != cudaSuccess) return e;
    }
    const uint32_t items = a.owned_tiles * 1024u;
    const uint32_t per_batch = pipes.q[0].capacity / a.sample_count;   // capacity >= sample_count is guaranteed by the caller
    const uint32_t nbatches = (items + per_batch - 1) / per_batch;
    const uint32_t shade_grid = (uint32_t)sms * VCRT_SHADE_GRID;
    // Batches are independent (disjoint pixels): with several pipelines they run side by side on their own streams and queue
    // sets, so that the tail of one trace launch (its longest rays) overlaps the other pipelines' work.  Pipeline 0 is the
    // render stream; the others fork from it and join it again (events), so callers see one stream-ordered render.
    const int use = (int)(nbatches < (uint32_t)pipes.n ? nbatches : (uint32_t)pipes.n);
    cudaError_t e;
    if (use > 1) {
        if ((e = cudaEventRecord(pipes.fork, pipes.stream[0])) != cudaSuccess) return e;
        for (int s = 1; s < use; ++s)
            if ((e = cudaStreamWaitEvent(pipes.stream[s], pipes.fork, 0)) != cudaSuccess) return e;
    }
    uint32_t bi = 0;
    for (uint32_t item0 = 0; item0 < items; item0 += per_batch, ++bi) {
        const int si = use > 1 ? (int)(bi % (uint32_t)use) : 0;
        const cudaStream_t stream = pipes.stream[si];
        const WfQueues& w = pipes.q[si];
        WfBatch b;
        b.item0 = item0;
        b.nitems = items - item0 < per_batch ? items - item0 : per_batch;
        b.npaths = b.nitems * a.sample_count;
        b.cur = 0u; b.bounce = 0u;
        // queue sizes and the trace kernel's fetch counter start at zero; from here on the trace kernel clears the queue its shade
        // kernel fills and the shade kernel clears the fetch counter: one memset per batch instead of a reset launch per bounce
        if ((e = cudaMemsetAsync(w.counts, 0, 3 * sizeof(unsigned int), stream)) != cudaSuccess) return e;
        for (uint32_t bounce = 0; bounce < a.env.max_bounces; ++bounce) {
            b.bounce = bounce;
            const int pi = bounce == 0u ? 1 : 0;   // bounce 0 reads no queue: one primary ray per pixel, generated from the item id
            cudaEvent_t t0 = nullptr, t1 = nullptr;
            if (timer && (e = timer->begin(stream, &t0, &t1)) != cudaSuccess) return e;
            const uint32_t rays_max = pi ? b.nitems : b.npaths;   // no more blocks than there can be rays (small frames)
            const uint32_t need = (rays_max + VCRT_PBLOCK - 1) / VCRT_PBLOCK;
            trace_fn[ci][qi][pi]<<<need < (uint32_t)trace_grid[pi] ? need : (uint32_t)trace_grid[pi], VCRT_PBLOCK, 0, stream>>>(a, w, b);
            if (timer && (e = timer->end(stream, t0, t1, pi != 0)) != cudaSuccess) return e;
            if (pi) wf_shade_kernel<SHADER, RNG_MODE, TRIG, true><<<shade_grid, 256, 0, stream>>>(a, w, b);
            else wf_shade_kernel<SHADER, RNG_MODE, TRIG, false><<<shade_grid, 256, 0, stream>>>(a, w, b);
            *launches += 2;
            b.cur ^= 1u;
        }
        if (pipes.before_accumulate && bi == 0 && (e = cudaStreamWaitEvent(stream, pipes.before_accumulate, 0)) != cudaSuccess) return e;   // frames in flight: fold in frame order
        wf_accumulate_kernel<<<(b.nitems + 255u) / 256u, 256, 0, stream>>>(a, w, b);
        ++*launches;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (use > 1)
        for (int s = 1; s < use; ++s) {
            if ((e = cudaEventRecord(pipes.join[s], pipes.stream[s])) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(pipes.stream[0], pipes.join[s], 0)) != cudaSuccess) return e;
        }
    return cudaSuccess;
}

template <int SHADER, int RNG_MODE>
static cudaError_t wf_trig(const KernelArgs& a, int trig, bool count, const WfPipes& pipes, uint32_t* launches, TraceTimer* timer) {
    if (trig == VCRT_TRIG_PORTABLE) return wf_run<SHADER, RNG_MODE, VCRT_TRIG_PORTABLE>(a, count, pipes, launches, timer);
    return wf_run<SHADER, RNG_MODE, VCRT_TRIG_LIBM>(a, count, pipes, launches, timer);
}

template <int SHADER>
static cudaError_t wf_rng(const KernelArgs& a, int rng, int trig, bool count, const WfPipes& pipes, uint32_t* launches, TraceTimer* timer) {
    if (rng == VCRT_RNG_PHILOX) return wf_trig<SHADER, VCRT_RNG_PHILOX>(a, trig, count, pipes, launches, timer);
    return wf_trig<SHADER, VCRT_RNG_PCG_REF>(a, trig, count, pipes, launches, timer);
}

cudaError_t launch_render_wavefront(const KernelArgs& a, int shader, int rng, int trig, bool count, const WfPipes& pipes, uint32_t* launches, TraceTimer* timer) {
    if (a.owned_tiles == 0u || pipes.n < 1) return cudaSuccess;
    if (shader == VCRT_SHADER_SIMPLE) return wf_rng<VCRT_SHADER_SIMPLE>(a, rng, trig, count, pipes, launches, timer);
    return wf_rng<VCRT_SHADER_FULL>(a, rng, trig, count, pipes, launches, timer);
}
#endif

}  // namespace vcrt
