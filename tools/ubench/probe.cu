// probe.cu -- in-process peak probes for bench.py's roofline (MEASUREMENT INFRASTRUCTURE, not part of the product library).
//
// The trace kernel's memory side is a stream of dependent, divergent 32-byte gathers (one node half / triangle half per lane
// and visit).  What bounds such a stream is not a bandwidth figure from a copy test but the rate at which L1TEX can pull
// independent 32-byte sectors out of the level that holds the working set.  These probes measure exactly that, on the GPU
// and in the process that runs the benchmark, with the kernel's own access shape:
//   vcrt_probe_gather   every lane of every warp chases its own pointer chain through a table of 32-byte records with
//                       ld.global.nc.v8.b32 (LDG.E.256, the trace kernel's load): table << L1 -> L1-hit ceiling,
//                       table ~ 32 MB -> L2-resident / L1-missing ceiling (C3's regime), table >> L2 -> DRAM gather ceiling (C4's)
//   vcrt_probe_stream   coalesced read of a buffer (L2-resident or HBM-sized): the sequential bandwidth of that level
// Returned figures are best-of-`reps` after one warm-up launch, timed with CUDA events on the launch stream.
#include <cuda_runtime.h>
#include <cstdint>

struct __align__(32) Rec { unsigned a[8]; };

__device__ __forceinline__ Rec ldg256(const Rec* p) {
    Rec r;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.a[0]), "=r"(r.a[1]), "=r"(r.a[2]), "=r"(r.a[3]), "=r"(r.a[4]), "=r"(r.a[5]), "=r"(r.a[6]), "=r"(r.a[7]) : "l"(p));
    return r;
}

__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__global__ void fill_kernel(Rec* rec, size_t n, unsigned mask) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        Rec r;
        const unsigned h = mix32((unsigned)i * 2654435761u + 12345u);
#pragma unroll
        for (int w = 0; w < 8; ++w) r.a[w] = 0x3f800000u + (unsigned)i + w;
        r.a[4] = h & mask;          // next record of the chain
        rec[i] = r;
    }
}

// the same chase over 64-byte records read as two adjacent 256-bit loads (the trace kernel's node / triangle fetch): table of n64 records.
// CHAINS independent chains per lane: with one chain a lane has one record in flight; the trace kernel has more than that in flight
// per lane on average (two halves of a node, a postponed triangle, the stack), so the ceiling of a DRAM-resident table is the
// largest rate over 1, 2 and 4 chains.
template <int CHAINS>
__global__ void __launch_bounds__(128) chase64_kernel(const Rec* __restrict__ rec, unsigned mask64, int steps, unsigned* out) {
    unsigned ix[CHAINS], acc = 0u;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) ix[c] = mix32((blockIdx.x * blockDim.x + threadIdx.x) * CHAINS + c) & mask64;
    for (int s = 0; s < steps; ++s) {
        Rec a[CHAINS], b[CHAINS];
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) { a[c] = ldg256(rec + 2 * (size_t)ix[c]); b[c] = ldg256(rec + 2 * (size_t)ix[c] + 1); }
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            acc ^= a[c].a[0] ^ a[c].a[1] ^ a[c].a[2] ^ a[c].a[3] ^ a[c].a[5] ^ a[c].a[6] ^ a[c].a[7] ^ b[c].a[0] ^ b[c].a[3] ^ b[c].a[7];
            ix[c] = (a[c].a[4] ^ (b[c].a[4] >> 1) ^ (unsigned)c) & mask64;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int CHAINS>
__global__ void __launch_bounds__(128) chase_kernel(const Rec* __restrict__ rec, unsigned mask, int steps, unsigned* out) {
    unsigned ia = mix32(blockIdx.x * blockDim.x + threadIdx.x) & mask, ib = mix32(ia + 977u) & mask, acc = 0u;
    for (int s = 0; s < steps; ++s) {
        const Rec a = ldg256(rec + ia);
        Rec b;
        if (CHAINS == 2) b = ldg256(rec + ib);
        acc ^= a.a[0] ^ a.a[1] ^ a.a[2] ^ a.a[3] ^ a.a[5] ^ a.a[6] ^ a.a[7];
        ia = a.a[4];
        if (CHAINS == 2) { acc ^= b.a[0] ^ b.a[3] ^ b.a[7]; ib = (b.a[4] * 3u + 1u) & mask; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(256) stream_kernel(const uint4* __restrict__ p, size_t n16, unsigned* out) {
    unsigned acc = 0u;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(p + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) out[0] = acc;   // keeps the loads alive
}

extern "C" int vcrt_probe_gather(int device, int records_log2, int steps, int chains, int reps, double* g_per_s) {
    if (!g_per_s || records_log2 < 4 || records_log2 > 28 || steps < 1 || (chains != 1 && chains != 2)) return -1;
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const size_t n = (size_t)1 << records_log2;
    const unsigned mask = (unsigned)(n - 1);
    const int blocks = sms * 10 * 6, threads = 128;     // 10 resident blocks of 128 per SM (the trace kernel's shape), 6 waves
    Rec* d = nullptr; unsigned* out = nullptr;
    if (cudaMalloc(&d, n * sizeof(Rec)) != cudaSuccess || cudaMalloc(&out, (size_t)blocks * threads * 4) != cudaSuccess) { cudaFree(d); return -3; }
    fill_kernel<<<sms * 8, 256>>>(d, n, mask);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep <= reps; ++rep) {
        cudaEventRecord(e0);
        if (chains == 2) chase_kernel<2><<<blocks, threads>>>(d, mask, steps, out);
        else chase_kernel<1><<<blocks, threads>>>(d, mask, steps, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d); cudaFree(out);
    if (err != cudaSuccess) return -4;
    *g_per_s = (double)blocks * threads * steps * chains / (best * 1e-3) / 1e9;
    return 0;
}

// 64-byte records (two adjacent sectors per gather): returns G records/s; bytes/s = 64 x that
extern "C" int vcrt_probe_gather64(int device, int records_log2, int steps, int chains, int reps, double* g_per_s) {
    if (!g_per_s || records_log2 < 4 || records_log2 > 27 || steps < 1 || (chains != 1 && chains != 2 && chains != 4)) return -1;
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const size_t n = (size_t)1 << records_log2;
    const int blocks = sms * 10 * 6, threads = 128;
    Rec* d = nullptr; unsigned* out = nullptr;
    if (cudaMalloc(&d, n * 2 * sizeof(Rec)) != cudaSuccess || cudaMalloc(&out, (size_t)blocks * threads * 4) != cudaSuccess) { cudaFree(d); return -3; }
    fill_kernel<<<sms * 8, 256>>>(d, n * 2, 0xffffffffu);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep <= reps; ++rep) {
        cudaEventRecord(e0);
        if (chains == 4) chase64_kernel<4><<<blocks, threads>>>(d, (unsigned)(n - 1), steps, out);
        else if (chains == 2) chase64_kernel<2><<<blocks, threads>>>(d, (unsigned)(n - 1), steps, out);
        else chase64_kernel<1><<<blocks, threads>>>(d, (unsigned)(n - 1), steps, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d); cudaFree(out);
    if (err != cudaSuccess) return -4;
    *g_per_s = (double)blocks * threads * steps * chains / (best * 1e-3) / 1e9;
    return 0;
}

// 128-byte records (one full L2 line, four sectors per gather): the most a random record fetch can bring in per DRAM transaction.
// Two adjacent 64-byte records of the trace kernel (sibling nodes, neighbouring leaves' triangles) share such a line, so the
// kernel's DRAM rate lies between the 64-byte and the 128-byte gather ceilings.
template <int CHAINS>
__global__ void __launch_bounds__(128) chase128_kernel(const Rec* __restrict__ rec, unsigned mask128, int steps, unsigned* out) {
    unsigned ix[CHAINS], acc = 0u;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) ix[c] = mix32((blockIdx.x * blockDim.x + threadIdx.x) * CHAINS + c) & mask128;
    for (int s = 0; s < steps; ++s) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            const Rec* p = rec + 4 * (size_t)ix[c];
            const Rec a = ldg256(p), b = ldg256(p + 1), d = ldg256(p + 2), e = ldg256(p + 3);
            acc ^= a.a[0] ^ a.a[7] ^ b.a[0] ^ b.a[3] ^ d.a[1] ^ d.a[6] ^ e.a[2] ^ e.a[5];
            ix[c] = (a.a[4] ^ (b.a[4] >> 1) ^ (d.a[4] >> 2) ^ (e.a[4] >> 3) ^ (unsigned)c) & mask128;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

extern "C" int vcrt_probe_gather128(int device, int records_log2, int steps, int chains, int reps, double* g_per_s) {
    if (!g_per_s || records_log2 < 4 || records_log2 > 26 || steps < 1 || (chains != 1 && chains != 2)) return -1;
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const size_t n = (size_t)1 << records_log2;
    const int blocks = sms * 10 * 6, threads = 128;
    Rec* d = nullptr; unsigned* out = nullptr;
    if (cudaMalloc(&d, n * 4 * sizeof(Rec)) != cudaSuccess || cudaMalloc(&out, (size_t)blocks * threads * 4) != cudaSuccess) { cudaFree(d); return -3; }
    fill_kernel<<<sms * 8, 256>>>(d, n * 4, 0xffffffffu);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep <= reps; ++rep) {
        cudaEventRecord(e0);
        if (chains == 2) chase128_kernel<2><<<blocks, threads>>>(d, (unsigned)(n - 1), steps, out);
        else chase128_kernel<1><<<blocks, threads>>>(d, (unsigned)(n - 1), steps, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d); cudaFree(out);
    if (err != cudaSuccess) return -4;
    *g_per_s = (double)blocks * threads * steps * chains / (best * 1e-3) / 1e9;
    return 0;
}

extern "C" int vcrt_probe_stream(int device, size_t bytes, int passes, int reps, double* gb_per_s) {
    if (!gb_per_s || bytes < 4096 || passes < 1) return -1;
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    uint4* d = nullptr; unsigned* out = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess || cudaMalloc(&out, 4) != cudaSuccess) { cudaFree(d); return -3; }
    cudaMemset(d, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep <= reps; ++rep) {
        cudaEventRecord(e0);
        for (int p = 0; p < passes; ++p) stream_kernel<<<sms * 8, 256>>>(d, bytes / 16, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d); cudaFree(out);
    if (err != cudaSuccess) return -4;
    *gb_per_s = (double)bytes * passes / (best * 1e-3) / 1e9;
    return 0;
}
