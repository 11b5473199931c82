// vcrt_post.cu -- resolve of the f32 accumulation buffer into the rgba8 target.
// gamma 2.2 = the reference's post-process fragment shader (post-process-shader.frag:62-70).
#include "vcrt_launch.h"

namespace vcrt {

__global__ void resolve_kernel(const float4* __restrict__ accumf, uchar4* __restrict__ target, uint32_t npix, float inv_total, float inv_gamma) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 a = accumf[i];
    float c[4] = {a.x * inv_total, a.y * inv_total, a.z * inv_total, a.w * inv_total};
    uint8_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float v = c[k];
        v = !(v == v) ? 0.0f : (v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v));
        if (inv_gamma > 0.0f && k < 3) v = powf(v, inv_gamma);
        o[k] = (uint8_t)rintf(v * 255.0f);
    }
    target[i] = make_uchar4(o[0], o[1], o[2], o[3]);
}

cudaError_t launch_resolve(const float4* accumf, uchar4* target, uint32_t npix, float inv_total, float inv_gamma, cudaStream_t stream) {
    if (npix == 0) return cudaSuccess;
    resolve_kernel<<<(npix + 255) / 256, 256, 0, stream>>>(accumf, target, npix, inv_total, inv_gamma);
    return cudaGetLastError();
}

}  // namespace vcrt
