// gather.cu -- microbenchmark behind the node-layout decision of the fast traversal:
// dependent random gathers of BVH-node-sized records, one chain per lane, as the trace kernel issues them.
//   A  64 B record, 4 x LDG.128      B  64 B record, 2 x LDG.256
//   C  32 B record, 2 x LDG.128      D  32 B record, 1 x LDG.256
// usage: gather [records_log2=20] [steps=64]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct __align__(32) f8 { float a[8]; };
__device__ __forceinline__ f8 ldg256(const void* p) {
    f8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a[0]), "=f"(r.a[1]), "=f"(r.a[2]), "=f"(r.a[3]), "=f"(r.a[4]), "=f"(r.a[5]), "=f"(r.a[6]), "=f"(r.a[7]) : "l"(p));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(128) chase(const float4* __restrict__ rec, unsigned mask, int steps, float* out) {
    unsigned idx = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u & mask;
    float acc = 0.f;
    for (int s = 0; s < steps; ++s) {
        if (MODE == 0) {
            const float4* p = rec + 4 * (size_t)idx;
            float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
            acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.y + d.z + d.w;
            idx = __float_as_uint(d.x) & mask;
        } else if (MODE == 1) {
            const float4* p = rec + 4 * (size_t)idx;
            f8 a = ldg256(p), b = ldg256(p + 2);
            for (int i = 0; i < 8; ++i) acc += a.a[i];
            for (int i = 0; i < 7; ++i) acc += b.a[i + 1] * (i ? 1.f : 0.f);
            idx = __float_as_uint(b.a[4]) & mask;
        } else if (MODE == 2) {
            const float4* p = rec + 2 * (size_t)idx;
            float4 a = __ldg(p), b = __ldg(p + 1);
            acc += a.x + a.y + a.z + a.w + b.y + b.z + b.w;
            idx = __float_as_uint(b.x) & mask;
        } else {
            const float4* p = rec + 2 * (size_t)idx;
            f8 a = ldg256(p);
            for (int i = 0; i < 4; ++i) acc += a.a[i];
            acc += a.a[5] + a.a[6] + a.a[7];
            idx = __float_as_uint(a.a[4]) & mask;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main(int argc, char** argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 20, steps = argc > 2 ? atoi(argv[2]) : 64;
    const size_t n = (size_t)1 << lg;
    const unsigned mask = (unsigned)(n - 1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 9 * 8, threads = 128;
    float* out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    for (int mode = 0; mode < 4; ++mode) {
        const int words = mode < 2 ? 16 : 8;           // floats per record
        const int nextw = mode < 2 ? 12 : 4;           // word holding the next index
        std::vector<unsigned> h(n * words);
        unsigned s = 12345u;
        for (size_t i = 0; i < n; ++i) {
            for (int w = 0; w < words; ++w) h[i * words + w] = 0x3f800000u;
            s = s * 1664525u + 1013904223u;
            h[i * words + nextw] = (s >> 4) & mask;
        }
        float4* d; cudaMalloc(&d, h.size() * 4);
        cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) chase<0><<<blocks, threads>>>(d, mask, steps, out);
            if (mode == 1) chase<1><<<blocks, threads>>>(d, mask, steps, out);
            if (mode == 2) chase<2><<<blocks, threads>>>(d, mask, steps, out);
            if (mode == 3) chase<3><<<blocks, threads>>>(d, mask, steps, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        const double gathers = (double)blocks * threads * steps;
        printf("records 2^%d (%zu MB) mode %c: %.3f ms  %.1f Ggather/s  %.0f GB/s useful  %s\n", lg, n * words * 4 >> 20, "ABCD"[mode], best,
               gathers / best / 1e6, gathers * words * 4 / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
        cudaFree(d);
    }
    return 0;
}
