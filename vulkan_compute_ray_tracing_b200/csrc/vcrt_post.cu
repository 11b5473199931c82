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

// Tile-sharded frames (multi-GPU): pixels of the 32x32 tiles owned by (rank, count) <-> a packed buffer holding those
// tiles back to back, 1024 pixels each in row-major order inside the tile (pixels outside the image are zero).  The
// packed buffers of all ranks have equal size after padding, so one all-gather moves them (sharding.py).
template <typename T, bool PACK>
__global__ void tiles_kernel(T* __restrict__ image, T* __restrict__ packed, uint32_t W, uint32_t H, uint32_t tilesX, uint32_t rank, uint32_t count, uint32_t owned) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= owned * 1024u) return;
    const uint32_t j = i >> 10, inner = i & 1023u;
    const uint32_t tile = j * count + rank;
    const uint32_t ty = tile / tilesX, tx = tile - ty * tilesX;
    const uint32_t x = tx * 32u + (inner & 31u), y = ty * 32u + (inner >> 5);
    const bool in = x < W && y < H;
    if (PACK) {
        T v;
        memset(&v, 0, sizeof v);
        if (in) v = image[(size_t)y * W + x];
        packed[i] = v;
    } else if (in) {
        image[(size_t)y * W + x] = packed[i];
    }
}

cudaError_t launch_tiles(void* image, void* packed, int elem_bytes, bool pack, uint32_t W, uint32_t H, uint32_t rank, uint32_t count, uint32_t owned, cudaStream_t stream) {
    if (owned == 0) return cudaSuccess;
    const uint32_t tilesX = (W + 31) / 32, blocks = (owned * 1024u + 255u) / 256u;
    if (elem_bytes == 4) {
        if (pack) tiles_kernel<uchar4, true><<<blocks, 256, 0, stream>>>((uchar4*)image, (uchar4*)packed, W, H, tilesX, rank, count, owned);
        else tiles_kernel<uchar4, false><<<blocks, 256, 0, stream>>>((uchar4*)image, (uchar4*)packed, W, H, tilesX, rank, count, owned);
    } else {
        if (pack) tiles_kernel<float4, true><<<blocks, 256, 0, stream>>>((float4*)image, (float4*)packed, W, H, tilesX, rank, count, owned);
        else tiles_kernel<float4, false><<<blocks, 256, 0, stream>>>((float4*)image, (float4*)packed, W, H, tilesX, rank, count, owned);
    }
    return cudaGetLastError();
}

cudaError_t launch_resolve(const float4* accumf, uchar4* target, uint32_t npix, float inv_total, float inv_gamma, cudaStream_t stream) {
    if (npix == 0) return cudaSuccess;
    resolve_kernel<<<(npix + 255) / 256, 256, 0, stream>>>(accumf, target, npix, inv_total, inv_gamma);
    return cudaGetLastError();
}

}  // namespace vcrt
