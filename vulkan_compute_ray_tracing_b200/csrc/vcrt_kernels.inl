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
    if (!(a.flags & VCRT_FLAG_STATIC_KERNEL)) {
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

}  // namespace vcrt
