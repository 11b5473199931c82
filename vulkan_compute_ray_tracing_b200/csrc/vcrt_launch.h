// vcrt_launch.h -- host-visible launchers of the render kernels (one per traversal mode / translation unit).
#pragma once
#include <cuda_runtime.h>
#include "vcrt_path.cuh"

#define VCRT_BLOCK 128
#define VCRT_PBLOCK 128   /* persistent kernel */

namespace vcrt {
cudaError_t launch_render_reference(const KernelArgs& a, int shader, int rng, int trig, bool count, cudaStream_t stream);
cudaError_t launch_render_fast(const KernelArgs& a, int shader, int rng, int trig, bool count, cudaStream_t stream);
cudaError_t launch_render_brute(const KernelArgs& a, int shader, int rng, int trig, bool count, cudaStream_t stream);
cudaError_t launch_resolve(const float4* accumf, uchar4* target, uint32_t npix, float inv_total, float inv_gamma, cudaStream_t stream);
}  // namespace vcrt
