// vcrt_devbuild.cu -- kernels and driver of the on-device record build (algorithm and per-element bodies: vcrt_devbuild.cuh).
#include <cub/cub.cuh>

#include <chrono>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "vcrt_devbuild.h"
#include "vcrt_devbuild.cuh"

namespace vcrt {
namespace devbuild {

namespace {

constexpr int kBlock = 256;
inline unsigned blocks_for(size_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

__global__ void k_link(View v) { const uint32_t i = blockIdx.x * kBlock + threadIdx.x; if (i < v.nbvh) link_children(v, i); }
__global__ void k_count(View v) { const uint32_t i = blockIdx.x * kBlock + threadIdx.x; if (i < v.nbvh) count_up(v, i); }
__global__ void k_rank(View v, uint32_t nleaves) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= v.nbvh || !is_leaf(v.bvh[i])) return;
    uint32_t rank, depth;
    if (!leaf_rank(v, i, rank, depth)) return;
    if (rank < nleaves) v.leaf_node[rank] = i; else atom_max(v.status + ST_ERROR, ERR_SHARED);
    atom_max(v.status + ST_MAXDEPTH, depth);
}
__global__ void k_status_init(uint32_t* status) {
    const int i = threadIdx.x;
    if (i < ST_WORDS) status[i] = (i >= ST_CENTROID && i < ST_CENTROID + 3) ? 0xffffffffu : 0u;
}
__global__ void k_slots(View v, uint32_t n, float4* tris64, float4* lo, float4* hi) {
    const uint32_t s = blockIdx.x * kBlock + threadIdx.x;
    if (s < n) make_slot(v, s, tris64, lo, hi);
}
__global__ void k_morton(const float4* lo, const float4* hi, const uint32_t* status, uint32_t n, uint64_t* keys, uint32_t* vals) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    float clo[3], chi[3];
    for (int k = 0; k < 3; ++k) { clo[k] = ord2f(status[ST_CENTROID + k]); chi[k] = ord2f(status[ST_CENTROID + 3 + k]); }
    keys[i] = morton_of(lo[i], hi[i], clo, chi);
    vals[i] = i;
}
__global__ void k_gather(const float4* lo, const float4* hi, const uint32_t* order, uint32_t n, float4* out_lo, float4* out_hi) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i < n) { out_lo[i] = lo[order[i]]; out_hi[i] = hi[order[i]]; }
}
__global__ void k_nearest(const float4* lo, const float4* hi, uint32_t m, uint32_t* nn) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i < m) nn[i] = nearest(lo, hi, m, i);
}
__global__ void k_flags(const uint32_t* nn, uint32_t m, uint64_t* flags) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i < m) flags[i] = merge_flags(nn, i);
}
__global__ void k_merge(const float4* lo, const float4* hi, const uint32_t* nn, const uint64_t* scan, uint32_t m, uint32_t node_base, float4* out_lo, float4* out_hi,
                        float* nodes, uint32_t* totals) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= m) return;
    merge_write(lo, hi, nn, scan, i, node_base, out_lo, out_hi, nodes);
    if (i == m - 1u) {
        const uint64_t t = scan[i] + merge_flags(nn, i);
        totals[0] = (uint32_t)(t & 0xffffffffull);   // clusters of the next round
        totals[1] = (uint32_t)(t >> 32);             // nodes created in this round
    }
}
// The same three kernels with the cluster count and the node base read from device memory (st = {clusters, nodes created so far}), so
// that several PLOC rounds can be queued without a host round trip in between; `upper` bounds the count from above (the value the host
// saw last).  A round that finds a single cluster left only carries the state (and the root's box) over.
__global__ void k_nearest_d(const float4* lo, const float4* hi, const uint32_t* st, uint32_t* nn) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x, m = st[0];
    if (m > 1u && i < m) nn[i] = nearest(lo, hi, m, i);
}
__global__ void k_flags_d(const uint32_t* nn, const uint32_t* st, uint32_t upper, uint64_t* flags) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x, m = st[0];
    if (i < upper) flags[i] = (m > 1u && i < m) ? merge_flags(nn, i) : 0ull;
}
__global__ void k_merge_d(const float4* lo, const float4* hi, const uint32_t* nn, const uint64_t* scan, const uint32_t* st, float4* out_lo, float4* out_hi,
                          float* nodes, uint32_t* st_out) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x, m = st[0], node_base = st[1];
    if (m <= 1u) {
        if (i == 0u) { out_lo[0] = lo[0]; out_hi[0] = hi[0]; st_out[0] = m; st_out[1] = node_base; }
        return;
    }
    if (i >= m) return;
    merge_write(lo, hi, nn, scan, i, node_base, out_lo, out_hi, nodes);
    if (i == m - 1u) {
        const uint64_t t = scan[i] + merge_flags(nn, i);
        st_out[0] = (uint32_t)(t & 0xffffffffull);               // clusters of the next round
        st_out[1] = node_base + (uint32_t)(t >> 32);             // nodes created so far
    }
}
__global__ void k_state_init(uint32_t* st, uint32_t m) { st[0] = m; st[1] = 0u; }
__global__ void k_wide_count(const float* nodes, const WideItem* items, uint32_t n, uint32_t* inner) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i < n) inner[i] = wide_inner_count(nodes, items[i]);
}
__global__ void k_wide_write(const float* nodes, QFrame q, const WideItem* items, const uint32_t* inner, const uint32_t* offs, uint32_t n, uint32_t level_base, WideItem* next,
                             uint32_t* q4nodes, uint32_t* status, uint32_t* totals) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    wide_write(nodes, q, items[i], level_base + i, level_base + n + offs[i], offs[i], next, q4nodes, status);
    if (i == n - 1u) totals[0] = offs[i] + inner[i];
}
__global__ void k_first_item(WideItem* items, int32_t root) { items[0].node2 = root; items[0].stack_above = 0u; }

struct Temp {   // stream-ordered scratch, freed when the build ends (successfully or not)
    cudaStream_t stream;
    std::vector<void*> ptrs;
    explicit Temp(cudaStream_t s) : stream(s) {}
    ~Temp() { for (void* p : ptrs) cudaFreeAsync(p, stream); }
    template <typename T>
    cudaError_t get(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, (count ? count : 1) * sizeof(T), stream);
        if (e == cudaSuccess) { ptrs.push_back(p); *out = (T*)p; }
        return e;
    }
};

}  // namespace

#define DB_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { why = std::string("CUDA: ") + cudaGetErrorString(e_) + " (" #call ")"; return -1; } } while (0)

int run(const void* d_bvh, uint32_t nbvh, const void* d_tris, uint32_t ntris, float max_quantum, uint32_t max_stack, cudaStream_t stream, const Alloc& alloc, Result& out,
        std::string& why) {
    why.clear();
    if (nbvh < 3) { why = "fewer than two leaves"; return 1; }
    {   // scratch comes from the device's stream-ordered pool: keep what it has handed out across builds instead of returning it
        // to the driver at every synchronise (the default), or each build pays for a few hundred MB of fresh allocations
        int dev = 0;
        cudaMemPool_t pool = nullptr;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&t_begin]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    Temp tmp(stream);
    View v;
    v.bvh = (const vcrt_bvh_node*)d_bvh; v.nbvh = nbvh; v.tris = (const vcrt_triangle*)d_tris; v.ntris = ntris;
    DB_CU(tmp.get(&v.parent, nbvh)); DB_CU(tmp.get(&v.count, nbvh)); DB_CU(tmp.get(&v.arrive, nbvh)); DB_CU(tmp.get(&v.status, ST_WORDS));
    uint32_t* totals = nullptr;
    DB_CU(tmp.get(&totals, 2));
    DB_CU(cudaMemsetAsync(v.parent, 0xff, (size_t)nbvh * 4, stream));
    DB_CU(cudaMemsetAsync(v.count, 0, (size_t)nbvh * 4, stream));
    DB_CU(cudaMemsetAsync(v.arrive, 0, (size_t)nbvh * 4, stream));
    k_status_init<<<1, 32, 0, stream>>>(v.status);
    // ---- A: tie ranks
    k_link<<<blocks_for(nbvh), kBlock, 0, stream>>>(v);
    k_count<<<blocks_for(nbvh), kBlock, 0, stream>>>(v);
    uint32_t h_status[ST_WORDS], nleaves = 0;
    DB_CU(cudaMemcpyAsync(h_status, v.status, sizeof h_status, cudaMemcpyDeviceToHost, stream));
    DB_CU(cudaMemcpyAsync(&nleaves, v.count, 4, cudaMemcpyDeviceToHost, stream));     // leaves below node 0
    DB_CU(cudaStreamSynchronize(stream));
    if (h_status[ST_ERROR]) { why = "the bound tree is not a plain tree (shared subtree, cycle, leaf with children or depth > 4096): flags " + std::to_string(h_status[ST_ERROR]); return 1; }
    if (nleaves < 2) { why = "fewer than two leaves"; return 1; }
    out.ms_ranks = since();
    DB_CU(tmp.get(&v.leaf_node, nleaves));
    DB_CU(cudaMemsetAsync(v.leaf_node, 0xff, (size_t)nleaves * 4, stream));
    k_rank<<<blocks_for(nbvh), kBlock, 0, stream>>>(v, nleaves);
    // ---- B: triangle records, leaf boxes, Morton order
    float4* ftris = (float4*)alloc.ftris((size_t)nleaves * 64);
    if (!ftris) { why = "allocation of the triangle records failed"; return -1; }
    float4 *lo_a, *hi_a, *lo_b, *hi_b;
    DB_CU(tmp.get(&lo_a, nleaves)); DB_CU(tmp.get(&hi_a, nleaves)); DB_CU(tmp.get(&lo_b, nleaves)); DB_CU(tmp.get(&hi_b, nleaves));
    k_slots<<<blocks_for(nleaves), kBlock, 0, stream>>>(v, nleaves, ftris, lo_a, hi_a);
    uint64_t *keys_a, *keys_b, *flags, *scan;
    uint32_t *vals_a, *vals_b, *nn;
    DB_CU(tmp.get(&keys_a, nleaves)); DB_CU(tmp.get(&keys_b, nleaves)); DB_CU(tmp.get(&vals_a, nleaves)); DB_CU(tmp.get(&vals_b, nleaves));
    DB_CU(tmp.get(&flags, nleaves)); DB_CU(tmp.get(&scan, nleaves)); DB_CU(tmp.get(&nn, nleaves));
    k_morton<<<blocks_for(nleaves), kBlock, 0, stream>>>(lo_a, hi_a, v.status, nleaves, keys_a, vals_a);
    size_t sort_bytes = 0, scan_bytes = 0, scan32_bytes = 0;
    DB_CU(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys_a, keys_b, vals_a, vals_b, (int)nleaves, 0, 63, stream));
    DB_CU(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flags, scan, (int)nleaves, stream));
    DB_CU(cub::DeviceScan::ExclusiveSum(nullptr, scan32_bytes, vals_a, vals_b, (int)nleaves, stream));
    size_t cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    if (scan32_bytes > cub_bytes) cub_bytes = scan32_bytes;
    uint8_t* cub_tmp;
    DB_CU(tmp.get(&cub_tmp, cub_bytes));
    DB_CU(cub::DeviceRadixSort::SortPairs(cub_tmp, sort_bytes, keys_a, keys_b, vals_a, vals_b, (int)nleaves, 0, 63, stream));
    k_gather<<<blocks_for(nleaves), kBlock, 0, stream>>>(lo_a, hi_a, vals_b, nleaves, lo_b, hi_b);
    // ---- C: PLOC
    DB_CU(cudaStreamSynchronize(stream));
    out.ms_sort = since();
    float* nodes;                                  // binary nodes, 16 floats each, in creation order (the root is the last one)
    DB_CU(tmp.get(&nodes, (size_t)(nleaves - 1) * 16));
    float4 *cur_lo = lo_b, *cur_hi = hi_b, *nxt_lo = lo_a, *nxt_hi = hi_a;
    // Rounds are queued eight at a time: the kernels read the cluster count from device memory (k_*_d), the host looks at it once per
    // chunk (a round trip per round was 2.5 of the 4 ms of a 1 M-triangle build: 67 rounds).
    uint32_t m = nleaves, node_base = 0, rounds = 0;
    uint32_t* st;                                    // two {clusters, nodes so far} slots, ping-pong
    DB_CU(tmp.get(&st, 4));
    k_state_init<<<1, 1, 0, stream>>>(st, nleaves);
    int cur_st = 0;
    while (m > 1) {
        const uint32_t upper = m;
        const int chunk = 8;
        for (int r = 0; r < chunk; ++r) {
            k_nearest_d<<<blocks_for(upper), kBlock, 0, stream>>>(cur_lo, cur_hi, st + 2 * cur_st, nn);
            k_flags_d<<<blocks_for(upper), kBlock, 0, stream>>>(nn, st + 2 * cur_st, upper, flags);
            DB_CU(cub::DeviceScan::ExclusiveSum(cub_tmp, scan_bytes, flags, scan, (int)upper, stream));
            k_merge_d<<<blocks_for(upper), kBlock, 0, stream>>>(cur_lo, cur_hi, nn, scan, st + 2 * cur_st, nxt_lo, nxt_hi, nodes, st + 2 * (cur_st ^ 1));
            cur_st ^= 1;
            std::swap(cur_lo, nxt_lo); std::swap(cur_hi, nxt_hi);
        }
        uint32_t h_st[2];
        DB_CU(cudaMemcpyAsync(h_st, st + 2 * cur_st, 8, cudaMemcpyDeviceToHost, stream));
        DB_CU(cudaStreamSynchronize(stream));
        if (h_st[0] >= m || h_st[0] + h_st[1] != nleaves) { why = "PLOC made no progress (non-finite boxes?)"; return 1; }
        m = h_st[0]; node_base = h_st[1];
        rounds += chunk;
        if (rounds > 4096) { why = "PLOC did not converge"; return 1; }
    }
    if (node_base != nleaves - 1) { why = "PLOC produced an inconsistent node count"; return 1; }
    out.ms_ploc = since();
    float4 root_lo, root_hi;
    DB_CU(cudaMemcpyAsync(&root_lo, cur_lo, 16, cudaMemcpyDeviceToHost, stream));
    DB_CU(cudaMemcpyAsync(&root_hi, cur_hi, 16, cudaMemcpyDeviceToHost, stream));
    DB_CU(cudaMemcpyAsync(h_status, v.status, sizeof h_status, cudaMemcpyDeviceToHost, stream));
    DB_CU(cudaStreamSynchronize(stream));
    if (h_status[ST_ERROR]) { why = "the bound tree or its triangles cannot be handled on the device: flags " + std::to_string(h_status[ST_ERROR]); return 1; }
    const int32_t root2 = (int32_t)f2u(root_lo.w);
    // ---- quantisation frame (the host builder's rule, quantize_fast_bvh): q stays inside [1, 32766], frame rounded to the floats the kernels use
    QFrame q;
    const double blo[3] = {root_lo.x, root_lo.y, root_lo.z}, bhi[3] = {root_hi.x, root_hi.y, root_hi.z};
    for (int a = 0; a < 3; ++a) {
        if (!(blo[a] <= bhi[a]) || !std::isfinite(blo[a]) || !std::isfinite(bhi[a])) { why = "scene bounds are not finite"; return 1; }
        double ext = bhi[a] - blo[a];
        if (!(ext > 0.0)) ext = 1e-3;
        const double quantum = ext / 32764.0, base = blo[a] - quantum, E = 32768.0 * quantum;
        if (quantum > (double)max_quantum) { why = "scene extent too large for 15-bit bounds"; return 1; }
        q.org[a] = (float)(base - E); q.ext[a] = (float)E;
        if (std::fabs((double)q.org[a] - (base - E)) > 0.01 * quantum || std::fabs((double)q.ext[a] - E) > 1e-6 * E) { why = "quantisation frame does not survive rounding to fp32"; return 1; }
    }
    // ---- D: 4-wide collapse, level by level (breadth-first numbering: the children of a level follow it)
    uint32_t* q4 = (uint32_t*)alloc.q4nodes((size_t)(nleaves - 1) * 64);
    if (!q4) { why = "allocation of the 4-wide nodes failed"; return -1; }
    WideItem *items_a, *items_b;
    uint32_t *inner = vals_a, *offs = vals_b;          // reuse: both hold at least nleaves words
    DB_CU(tmp.get(&items_a, nleaves)); DB_CU(tmp.get(&items_b, nleaves));
    k_first_item<<<1, 1, 0, stream>>>(items_a, root2);
    uint32_t cnt = 1, level_base = 0, levels = 0;
    while (cnt > 0) {
        k_wide_count<<<blocks_for(cnt), kBlock, 0, stream>>>(nodes, items_a, cnt, inner);
        DB_CU(cub::DeviceScan::ExclusiveSum(cub_tmp, scan32_bytes, inner, offs, (int)cnt, stream));
        k_wide_write<<<blocks_for(cnt), kBlock, 0, stream>>>(nodes, q, items_a, inner, offs, cnt, level_base, items_b, q4, v.status, totals);
        uint32_t next = 0;
        DB_CU(cudaMemcpyAsync(&next, totals, 4, cudaMemcpyDeviceToHost, stream));
        DB_CU(cudaStreamSynchronize(stream));
        level_base += cnt;
        cnt = next;
        std::swap(items_a, items_b);
        if (level_base + cnt > nleaves - 1 || ++levels > 4096) { why = "4-wide collapse produced an inconsistent node count"; return 1; }
    }
    DB_CU(cudaMemcpyAsync(h_status, v.status, sizeof h_status, cudaMemcpyDeviceToHost, stream));
    DB_CU(cudaStreamSynchronize(stream));
    DB_CU(cudaGetLastError());
    out.nslots = nleaves; out.nwide = level_base; out.root4 = 0;
    std::memcpy(out.qorg, q.org, sizeof out.qorg); std::memcpy(out.qext, q.ext, sizeof out.qext);
    out.depth = f2u(root_hi.w); out.bound_depth = h_status[ST_MAXDEPTH];
    out.stack4 = h_status[ST_STACK] + 2;          // + the sentinel slot and the register-held top's spill slot
    out.ploc_rounds = rounds; out.wide_levels = levels;
    out.ms_total = since();
    if (out.stack4 > max_stack) { why = "the 4-wide tree could ask for " + std::to_string(out.stack4) + " stack entries"; return 1; }
    return 0;
}

}  // namespace devbuild
}  // namespace vcrt
