// gather_tex.cu -- does the texture path add gather capacity next to the LSU path?  (the trace kernel
// sits at ~80 % of the L1 LSU data pipe.)  Dependent random gathers of 32-byte records, one or two chains per lane:
//   L   one chain, 1 x LDG.256                     (the trace kernel's node fetch)
//   T   one chain, 2 x tex1Dfetch<uint4>           (same bytes through the TEX pipe)
//   LL  two independent chains, both LDG.256
//   LT  two independent chains, one LDG.256 + one TEX
//   TT  two independent chains, both TEX
// usage: gather_tex [records_log2=20] [steps=64]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct __align__(32) u8w { unsigned a[8]; };
__device__ __forceinline__ u8w ldg256(const void* p) {
    u8w r;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.a[0]), "=r"(r.a[1]), "=r"(r.a[2]), "=r"(r.a[3]), "=r"(r.a[4]), "=r"(r.a[5]), "=r"(r.a[6]), "=r"(r.a[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ u8w tex256(cudaTextureObject_t t, unsigned idx) {
    const uint4 lo = tex1Dfetch<uint4>(t, 2 * (int)idx), hi = tex1Dfetch<uint4>(t, 2 * (int)idx + 1);
    u8w r;
    r.a[0] = lo.x; r.a[1] = lo.y; r.a[2] = lo.z; r.a[3] = lo.w; r.a[4] = hi.x; r.a[5] = hi.y; r.a[6] = hi.z; r.a[7] = hi.w;
    return r;
}
__device__ __forceinline__ unsigned fold(const u8w& r) { return r.a[0] ^ r.a[1] ^ r.a[2] ^ r.a[3] ^ r.a[5] ^ r.a[6] ^ r.a[7]; }

// MODE: bit 0 = chain A uses TEX, bit 1 = second chain present, bit 2 = chain B uses TEX
template <int MODE>
__global__ void __launch_bounds__(128) chase(const u8w* __restrict__ rec, cudaTextureObject_t tex, unsigned mask, int steps, unsigned* out) {
    unsigned ia = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u & mask, ib = (ia * 40503u + 977u) & mask, acc = 0u;
    for (int s = 0; s < steps; ++s) {
        u8w a = (MODE & 1) ? tex256(tex, ia) : ldg256(rec + ia);
        u8w b;
        if (MODE & 2) b = (MODE & 4) ? tex256(tex, ib) : ldg256(rec + ib);
        acc ^= fold(a); ia = a.a[4] & mask;
        if (MODE & 2) { acc ^= fold(b); ib = (b.a[4] * 3u + 1u) & mask; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main(int argc, char** argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 20, steps = argc > 2 ? atoi(argv[2]) : 64;
    const size_t n = (size_t)1 << lg;
    const unsigned mask = (unsigned)(n - 1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 9 * 8, threads = 128;
    unsigned* out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    std::vector<unsigned> h(n * 8);
    unsigned s = 12345u;
    for (size_t i = 0; i < n; ++i) {
        for (int w = 0; w < 8; ++w) h[i * 8 + w] = 0x3f800000u + (unsigned)i;
        s = s * 1664525u + 1013904223u;
        h[i * 8 + 4] = (s >> 4) & mask;
    }
    u8w* d; cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = d;
    rd.res.linear.desc = cudaCreateChannelDesc<uint4>();
    rd.res.linear.sizeInBytes = h.size() * 4;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    cudaError_t te = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    if (te != cudaSuccess) { printf("texture object: %s\n", cudaGetErrorString(te)); return 1; }
    const int modes[5] = {0, 1, 2, 2 | 4, 1 | 2 | 4};
    const char* names[5] = {"L ", "T ", "LL", "LT", "TT"};
    for (int m = 0; m < 5; ++m) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            switch (modes[m]) {
                case 0: chase<0><<<blocks, threads>>>(d, tex, mask, steps, out); break;
                case 1: chase<1><<<blocks, threads>>>(d, tex, mask, steps, out); break;
                case 2: chase<2><<<blocks, threads>>>(d, tex, mask, steps, out); break;
                case 6: chase<6><<<blocks, threads>>>(d, tex, mask, steps, out); break;
                default: chase<7><<<blocks, threads>>>(d, tex, mask, steps, out); break;
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        const double gathers = (double)blocks * threads * steps * ((modes[m] & 2) ? 2 : 1);
        printf("records 2^%d (%zu MB) mode %s: %.3f ms  %.1f G gathers/s (32 B each)  %s\n", lg, n * 32 >> 20, names[m], best, gathers / best / 1e6,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
