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

// All peers' tiles at once (vcrt_group_render): `gathered` = the packed buffers of ranks 0..world-1 back to back, each
// tiles_per_rank * 1024 elements (padded); one thread per image pixel fetches its value from the owner's buffer.
template <typename T>
__global__ void unpack_all_tiles_kernel(T* __restrict__ image, const T* __restrict__ gathered, uint32_t W, uint32_t H, uint32_t tilesX, uint32_t world, uint32_t tiles_per_rank) {
    const uint32_t x = blockIdx.x * 32u + (threadIdx.x & 31u), y = blockIdx.y * 8u + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const uint32_t tile = (y >> 5) * tilesX + (x >> 5);
    const uint32_t rank = tile % world, j = tile / world;
    image[(size_t)y * W + x] = gathered[((size_t)rank * tiles_per_rank + j) * 1024u + ((y & 31u) << 5) + (x & 31u)];
}

cudaError_t launch_unpack_all_tiles(void* image, const void* gathered, int elem_bytes, uint32_t W, uint32_t H, uint32_t world, uint32_t tiles_per_rank, cudaStream_t stream) {
    if (W == 0 || H == 0 || world == 0) return cudaSuccess;
    const dim3 grid((W + 31) / 32, (H + 7) / 8);
    const uint32_t tilesX = (W + 31) / 32;
    if (elem_bytes == 4) unpack_all_tiles_kernel<uchar4><<<grid, 256, 0, stream>>>((uchar4*)image, (const uchar4*)gathered, W, H, tilesX, world, tiles_per_rank);
    else unpack_all_tiles_kernel<float4><<<grid, 256, 0, stream>>>((float4*)image, (const float4*)gathered, W, H, tilesX, world, tiles_per_rank);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- post-process pass
// post-process-shader.frag:26-70 on the rgba8 target: smartDeNoise (a bilateral filter over a disc of radius
// round(kSigma*sigma); the shader ships with it commented out of main, :64) blended with the plain texel by `mix`, then
// pow(rgb, 1/gamma), alpha 1.  Sampling follows the reference's sampler (Image.cpp:353-364: normalised coordinates, LINEAR,
// REPEAT) as the Vulkan specification evaluates it: texel coordinate u*W - 0.5, floor + fraction, weights at 8 bits of
// sub-texel precision, four taps in the order of the spec's formula -- the same fp32 operations as the oracle, which is
// pinned to the fragment shader's own text (oracle/_ref).  The fragment of pixel (px, py) has fragTexCoord = pixel centre /
// image size.  One thread per pixel; the taps of a 16x16 block overlap almost completely, so the rgba8 reads are L1 hits.
__device__ __forceinline__ float4 post_texel(const uchar4* __restrict__ tex, int w, int h, int x, int y) {
    x %= w; if (x < 0) x += w;
    y %= h; if (y < 0) y += h;
    const uchar4 p = tex[(size_t)y * w + x];
    return make_float4((float)p.x / 255.0f, (float)p.y / 255.0f, (float)p.z / 255.0f, (float)p.w / 255.0f);
}

__device__ __forceinline__ float4 post_texture(const uchar4* __restrict__ tex, int w, int h, float uvx, float uvy) {
    const float u = uvx * (float)w - 0.5f, v = uvy * (float)h - 0.5f;
    const float fu = floorf(u), fv = floorf(v);
    const float a = rintf((u - fu) * 256.0f) / 256.0f, b = rintf((v - fv) * 256.0f) / 256.0f;
    const int i0 = (int)fu, j0 = (int)fv;
    const float4 t00 = post_texel(tex, w, h, i0, j0), t10 = post_texel(tex, w, h, i0 + 1, j0), t01 = post_texel(tex, w, h, i0, j0 + 1),
                 t11 = post_texel(tex, w, h, i0 + 1, j0 + 1);
    const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
    return make_float4(((w00 * t00.x + w10 * t10.x) + w01 * t01.x) + w11 * t11.x, ((w00 * t00.y + w10 * t10.y) + w01 * t01.y) + w11 * t11.y,
                       ((w00 * t00.z + w10 * t10.z) + w01 * t01.z) + w11 * t11.z, ((w00 * t00.w + w10 * t10.w) + w01 * t01.w) + w11 * t11.w);
}

__global__ void __launch_bounds__(256) post_process_kernel(const uchar4* __restrict__ tex, uchar4* __restrict__ out, int w, int h, float mix, float sigma,
                                                           float kSigma, float threshold, float inv_gamma) {
    const int px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (px >= w || py >= h) return;
    const float uvx = ((float)px + 0.5f) / (float)w, uvy = ((float)py + 0.5f) / (float)h;
    const float4 centr = post_texture(tex, w, h, uvx, uvy);
    float col[3] = {centr.x, centr.y, centr.z};
    if (mix != 0.0f) {
        const float radius = roundf(kSigma * sigma), radQ = radius * radius;
        const float invSigmaQx2 = 0.5f / (sigma * sigma), invSigmaQx2PI = 0.31830988618379067153776752674503f * invSigmaQx2;
        const float invThresholdSqx2 = 0.5f / (threshold * threshold), invThresholdSqrt2PI = 0.39894228040143267793994605993439f / threshold;
        float zBuff = 0.0f, ax = 0.0f, ay = 0.0f, az = 0.0f;
        for (float x = -radius; x <= radius; x += 1.0f) {
            const float pt = sqrtf(radQ - x * x);
            for (float y = -pt; y <= pt; y += 1.0f) {
                const float blurFactor = expf(-(x * x + y * y) * invSigmaQx2) * invSigmaQx2PI;
                const float4 walk = post_texture(tex, w, h, uvx + x / (float)w, uvy + y / (float)h);
                const float dx = walk.x - centr.x, dy = walk.y - centr.y, dz = walk.z - centr.z, dw = walk.w - centr.w;
                const float deltaFactor = expf(-((dx * dx + dy * dy) + (dz * dz + dw * dw)) * invThresholdSqx2) * invThresholdSqrt2PI * blurFactor;
                zBuff += deltaFactor;
                ax += deltaFactor * walk.x; ay += deltaFactor * walk.y; az += deltaFactor * walk.z;
            }
        }
        col[0] = mix * (ax / zBuff) + (1.0f - mix) * centr.x;
        col[1] = mix * (ay / zBuff) + (1.0f - mix) * centr.y;
        col[2] = mix * (az / zBuff) + (1.0f - mix) * centr.z;
    }
    uint8_t o[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = inv_gamma > 0.0f ? powf(col[k], inv_gamma) : col[k];
        v = !(v == v) ? 0.0f : (v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v));
        o[k] = (uint8_t)rintf(v * 255.0f);
    }
    out[(size_t)py * w + px] = make_uchar4(o[0], o[1], o[2], 255);
}

cudaError_t launch_post_process(const uchar4* tex, uchar4* out, uint32_t w, uint32_t h, float mix, float sigma, float kSigma, float threshold, float inv_gamma,
                                cudaStream_t stream) {
    if (w == 0 || h == 0) return cudaSuccess;
    post_process_kernel<<<dim3((w + 15) / 16, (h + 15) / 16), 256, 0, stream>>>(tex, out, (int)w, (int)h, mix, sigma, kSigma, threshold, inv_gamma);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- frames in flight
// The fold of one frame in flight (vcrt_frame_submit; the reference keeps MAX_FRAMES_IN_FLIGHT = 2 frames going, main.cpp:68,
// :298-316, :325, :394): the render kernels of consecutive frames overlap on their own streams and leave the frame's one sample
// per pixel in `sample`; this kernel, ordered behind the previous frame's fold by an event, applies it to the accumulation
// exactly as the render kernels do in the synchronous path (f32: acc += c, w += 1 -- ray-trace-compute.comp:375-379 for the
// rgba8 running mean), resolves the pixel (resolve_kernel's arithmetic) and writes it to the target and to the frame's own
// rgba8 image, which is what travels to the host while the next frames render.  sample == nullptr: the accumulation is already
// up to date (wavefront pipeline: its accumulate kernel ran behind the same event); only resolve / copy.
__global__ void __launch_bounds__(256) frame_fold_kernel(const __grid_constant__ KernelArgs a, const float4* __restrict__ sample, uchar4* __restrict__ image,
                                                         float inv_total, float inv_gamma) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= a.W * a.H) return;
    const uint32_t y = i / a.W, x = i - y * a.W;
    const uint32_t tile = (y >> 5) * a.tilesX + (x >> 5);
    const bool covered = sample != nullptr && x < a.covW && y < a.covH && tile % a.tile_count == a.tile_rank;
    uchar4 px;
    if (a.accum_mode == VCRT_ACCUM_F32) {
        float4 acc = a.accumf[i];
        if (covered) {
            const float4 c = sample[i];
            if (a.flags & VCRT_FLAG_INTERNAL_RESTART) acc = make_float4(0, 0, 0, 0);
            acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += 1.0f;
            a.accumf[i] = acc;
        }
        const float cc[4] = {acc.x * inv_total, acc.y * inv_total, acc.z * inv_total, acc.w * inv_total};
        uint8_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v = cc[k];
            v = !(v == v) ? 0.0f : (v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v));
            if (inv_gamma > 0.0f && k < 3) v = powf(v, inv_gamma);
            o[k] = (uint8_t)rintf(v * 255.0f);
        }
        px = make_uchar4(o[0], o[1], o[2], o[3]);
        a.target[i] = px;
    } else if (covered) {
        const float4 c = sample[i];
        px = a.accum8[i];
        running_mean_rgba8(px, make_float3(c.x, c.y, c.z), a.sample_begin);
        a.target[i] = px;
        a.accum8[i] = px;
    } else {
        px = a.target[i];
    }
    image[i] = px;
}

cudaError_t launch_frame_fold(const KernelArgs& a, const float4* sample, uchar4* image, float inv_total, float inv_gamma, cudaStream_t stream) {
    const uint32_t npix = a.W * a.H;
    if (npix == 0) return cudaSuccess;
    frame_fold_kernel<<<(npix + 255) / 256, 256, 0, stream>>>(a, sample, image, inv_total, inv_gamma);
    return cudaGetLastError();
}

cudaError_t launch_resolve(const float4* accumf, uchar4* target, uint32_t npix, float inv_total, float inv_gamma, cudaStream_t stream) {
    if (npix == 0) return cudaSuccess;
    resolve_kernel<<<(npix + 255) / 256, 256, 0, stream>>>(accumf, target, npix, inv_total, inv_gamma);
    return cudaGetLastError();
}

}  // namespace vcrt
