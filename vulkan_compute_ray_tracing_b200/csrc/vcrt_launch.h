// vcrt_launch.h -- host-visible launchers of the render kernels (one per traversal mode / translation unit).
#pragma once
#include <cuda_runtime.h>
#include "vcrt_path.cuh"

#define VCRT_BLOCK 128
#ifndef VCRT_PBLOCK
#define VCRT_PBLOCK 128   /* persistent kernel: threads per block */
#endif
#ifndef VCRT_PMINB
#define VCRT_PMINB 1     /* persistent kernel: min resident blocks per SM (register cap) */
#endif

namespace vcrt {
struct WfQueues;
// Wavefront pipeline of the fast traversal (vcrt_wavefront.cuh); queues are owned by the context.
cudaError_t launch_render_wavefront(const KernelArgs& a, int shader, int rng, int trig, bool count, cudaStream_t stream,
                                    float4* q0, float4* q1, uint2* hit, float4* sample_color, unsigned int* counts, uint32_t capacity, uint32_t* launches);
cudaError_t launch_render_reference(const KernelArgs& a, int shader, int rng, int trig, bool count, cudaStream_t stream);
cudaError_t launch_render_fast(const KernelArgs& a, int shader, int rng, int trig, bool count, cudaStream_t stream);
cudaError_t launch_render_brute(const KernelArgs& a, int shader, int rng, int trig, bool count, cudaStream_t stream);
cudaError_t launch_resolve(const float4* accumf, uchar4* target, uint32_t npix, float inv_total, float inv_gamma, cudaStream_t stream);
}  // namespace vcrt
