// vcrt_wavefront.cuh -- wavefront formulation of the fast path (GPU only): trace and shade are separate kernels that
// hand dense ray queues to each other through HBM; primary rays are never stored (bounce 0 derives them from the path id).
//
// Why: in the megakernel (vcrt_persistent.cuh) shading runs inside the traversal warps with a handful of lanes active
// (ncu: ~30 % of issued instructions at 2-7 active lanes), and lanes that finished a ray idle until enough of them
// want shading.  Here
//   trace  is a persistent-warps kernel whose lanes refill themselves from the ray queue the moment they finish
//          (store 8 bytes, fetch an index, load 48 bytes -- cheap enough to do at low lane counts), so the hot
//          inner-node loop keeps its lanes busy and carries no shading registers;
//   shade  runs one thread per traced ray: all lanes active, divergence only by material;
//   queues carry 48 B per ray: {o.xyz, path id} {d.xyz, rng word} {throughput.xyz, -}.
// A batch is a range of work items (pixels) times ALL samples of the call; consecutive path ids are the samples of
// one pixel.  Every finished path writes its colour to sample_color[path id]; the accumulate kernel then folds the
// samples of a pixel in sample order, so f32 sums and rgba8 histories are bit-identical to the other kernels.
// Bounce 0 is traced ONCE PER PIXEL: the shader's primary ray does not depend on the sample (no sub-pixel jitter,
// ray-trace-compute.comp:371), so the samples of a pixel share one closest-hit answer; the shade kernel fans it out to the
// pixel's paths (hit[item], item = path / sample_count).  Results are unchanged; at 64 spp it removes 98 % of the primary
// traversals (52 % of all closest-hit queries of a C3 step).
#pragma once

#include "vcrt_path.cuh"
#include "vcrt_launch.h"   // WfQueues

namespace vcrt {

struct WfBatch {
    uint32_t item0, nitems;       // work items (pixels) of this batch
    uint32_t npaths;              // nitems * sample_count
    uint32_t cur;                 // queue holding the rays of this bounce
    uint32_t bounce;
};

__device__ __forceinline__ uint32_t wf_append_slot(unsigned int* counter, bool want) {
    // warp-aggregated queue append: one atomic per warp
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0u) return 0u;
    const int leader = __ffs(m) - 1;
    const unsigned lane = threadIdx.x & 31u;
    unsigned base = 0u;
    if ((int)lane == leader) base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}
// The same in two halves, so that the atomic's round trip to L2 (the longest single wait of the shade kernel: 22 % of its
// stall samples sat on the shuffle that broadcasts the result, ncu r01_v10) overlaps the arithmetic between them.
__device__ __forceinline__ unsigned wf_append_begin(unsigned int* counter, bool want, unsigned& m) {
    m = __ballot_sync(0xffffffffu, want);
    unsigned base = 0u;
    if (m != 0u && (int)(threadIdx.x & 31u) == __ffs(m) - 1) base = atomicAdd(counter, (unsigned)__popc(m));
    return base;   // valid in the leader lane only, and only once the atomic has returned
}
__device__ __forceinline__ uint32_t wf_append_end(unsigned base, unsigned m) {
    if (m == 0u) return 0u;
    const unsigned lane = threadIdx.x & 31u;
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    return base + __popc(m & ((1u << lane) - 1u));
}

// Path id of a batch -> its pixel and sample (consecutive ids = the samples of one pixel).  False for pixels outside the
// covered extent (ragged edge tiles).
__device__ __forceinline__ bool wf_path_pixel(const KernelArgs& a, const WfBatch& b, uint32_t path, uint32_t& x, uint32_t& y, uint32_t& k) {
    const uint32_t it = path / a.sample_count;
    k = path - it * a.sample_count;
    return item_to_pixel(a, b.item0 + it, x, y);
}

// The first ray of a path (main(), ray-trace-compute.comp:352-373, :317-319) and the PCG seed (random.glsl:19).  In
// bounce 0 the trace and shade kernels compute it from the path id instead of reading it from a queue: a queue of
// primary rays would cost 48 B written + 80 B read per path for values that are a few flops away.
__device__ __forceinline__ void wf_primary(const KernelArgs& a, uint32_t x, uint32_t y, uint32_t k, Ray& r, uint32_t& rng) {
    const Ray pr = primary_ray(a.cam, x, y);
    r.o = pr.o;
    r.d = normalize(pr.d);
    rng = (600u * x + y) * (a.sample_begin + k + 1u);
}

// ---- trace: persistent warps, lanes refill from the queue; result = {t, slot} per ray
// Lane state: t.node (inner node >= 0 | leaf code < 0 | EMPTY), one postponed leaf, the stack with its sentinel.  The
// loop body is written for predication (trav_inner_step_lean): ncu on the branchy version showed ~50 of ~110 warp
// instructions per iteration spent on control flow around the 57-instruction box test (profiles/r01_v4_*).
// PRIMARY (bounce 0): the "queue" is implicit -- ray i is the primary ray of path i of the batch.
template <bool COUNT, int QN, bool PRIMARY>
__global__ void __launch_bounds__(VCRT_PBLOCK, VCRT_PMINB) wf_trace_kernel(const __grid_constant__ KernelArgs a, const WfQueues w, const WfBatch b) {
    const unsigned FULL = 0xffffffffu;
    const int32_t EMPTY = VCRT_FAST_EMPTY;
    const SceneView& s = a.scene;
    const uint32_t count = PRIMARY ? b.nitems : w.counts[b.cur];   // PRIMARY: one ray per work item (pixel), shared by its samples
    const float4* __restrict__ rays = w.q[b.cur];
    // (queued rays are counted by the shade kernel that consumes this launch's hits: nothing but traversal state lives in this kernel's registers)
    const int leaf_t = (int)a.leaf_threshold, refill_t = (int)a.shade_threshold, cont_t = (int)a.continue_threshold;

    uint32_t idx = 0xffffffffu;          // ray in flight (0xffffffff: none)
    bool done = false;                   // the queue is exhausted for this lane
#if VCRT_SMEMRAY
    // The ray's origin and direction are needed again only by the triangle test (3.6 times per ray on C3): they live in
    // shared memory, [component][thread] (conflict-free), instead of six registers across the inner-node loop.
    __shared__ float s_ray[6][VCRT_PBLOCK];
#endif
    Ray cur; cur.o = cur.d = f3(0, 0, 0);
    TravState t;
    t.idir = t.ood = f3(0, 0, 0); t.closest = VCRT_T_MAX; t.best = -1; t.node = EMPTY; t.sp = 0;
    t.selx = t.sely = t.selz = VCRT_Q15_SEL_LO;
    int32_t pending = EMPTY;             // one postponed leaf
    int32_t tos = EMPTY;                 // top of the traversal stack (vcrt_fast.cuh: trav_inner_step_lean)
    int32_t stack[VCRT_FAST_STACK - VCRT_SSTACK];   // accessed through predicated ld/st.local only (vcrt_fast.cuh: StackRef)
    StackRef sr;
    sr.lbase = (uint32_t)__cvta_generic_to_local(stack);
#if VCRT_SSTACK
    __shared__ int32_t s_stack[VCRT_SSTACK][VCRT_PBLOCK];
    sr.sbase = (uint32_t)__cvta_generic_to_shared(&s_stack[0][threadIdx.x]);
#else
    sr.sbase = 0u;
#endif
    sr.store_if(true, 0, EMPTY);         // sentinel: popping an exhausted stack yields EMPTY
    TraceStats st = {0u, 0u, 0u, 0u};

    for (;;) {
        // ---- refill: lanes whose ray is finished store the result and take the next ray
        if (!done && t.node == EMPTY && pending == EMPTY) {
            if (idx != 0xffffffffu) stream_st(w.hit + idx, make_uint2(f2u(t.closest), (uint32_t)t.best));
            idx = atomicAdd(w.counts + 2, 1u);   // ptxas aggregates this per warp (REDUX + one ATOMG)
            if (idx < count) {
                bool valid = true;
                if (PRIMARY) {
                    uint32_t x, y, rng;
                    valid = item_to_pixel(a, b.item0 + idx, x, y);
                    wf_primary(a, x, y, 0u, cur, rng);
                    if (valid) st.rays++;   // flushed as sample_count queries, one traversal (see the end of the kernel)
                } else {
                    const float4 o = stream_ld(rays + 3 * (size_t)idx), d = stream_ld(rays + 3 * (size_t)idx + 1);
                    cur.o = xyz(o); cur.d = xyz(d);
                }
                trav_begin<QN>(t, s, cur);
#if VCRT_SMEMRAY
                s_ray[0][threadIdx.x] = cur.o.x; s_ray[1][threadIdx.x] = cur.o.y; s_ray[2][threadIdx.x] = cur.o.z;
                s_ray[3][threadIdx.x] = cur.d.x; s_ray[4][threadIdx.x] = cur.d.y; s_ray[5][threadIdx.x] = cur.d.z;
#endif
                t.sp = 1;
                tos = EMPTY;
                if (!valid) { t.node = EMPTY; idx = 0xffffffffu; }   // a path outside the covered extent: nothing to trace or record
            } else {
                done = true;
                idx = 0xffffffffu;
            }
        }
#if VCRT_TAIL_SPLIT
        // the queue has run dry: the rays still in flight are finished by the tail loop below -- once fewer than VCRT_TAIL_ENTER lanes are
        // still walking (with most lanes busy the phase schedule of this loop is the better one, and there is nobody to share with)
        if (__any_sync(FULL, done) && __popc(__ballot_sync(FULL, t.node != EMPTY || pending != EMPTY)) < VCRT_TAIL_ENTER) break;
#else
        if (__all_sync(FULL, done)) break;
#endif

        for (;;) {
            // ---- VCRT_VISITS inner-node visits for every lane that has one; a lane that arrives at a leaf postpones it
            //      (one leaf) and keeps traversing
#pragma unroll
            for (int v = 0; v < (QN == 2 ? VCRT_VISITS4 : VCRT_VISITS); ++v) {
                if (QN == 2) {
                    if (t.node >= 0) {   // 4-wide node: both halves are fetched together, four children decided per round trip
                        if (COUNT) st.nodes++;
                        const Words8* p = s.q4nodes + 2 * (size_t)t.node;
                        const Words8 na = ldg8(p), nb = ldg8(p + 1);
                        int32_t c[4];
                        trav_test4(t, na, nb, c);
                        trav_descend4(t, s, c, pending, tos, sr);
                    }
                } else if (t.node >= 0) {
                    if (COUNT) st.nodes++;
                    float lN, rN;
                    bool hl, hr;
                    int32_t cl, cr;
                    trav_test_children<QN>(t, s, lN, rN, hl, hr, cl, cr);
                    trav_descend(t, lN, rN, hl, hr, cl, cr, pending, tos, sr);
                }
                {   // a leaf that came off the stack while nothing is postponed: postpone it, go on with the next entry.  Skipped
                    // (one vote) when no lane needs it: the step reads `tos`, which is usually still in flight from the visit's pop.
                    const bool park = (uint32_t)t.node > 0x80000000u && pending == EMPTY;
                    if (__any_sync(FULL, park)) {
                        pending = park ? t.node : pending;
                        trav_pop_if(park, t, tos, sr);
                    }
                }
            }
            // ---- phase switch.  Common case first, one vote: enough lanes still have an inner node -> go on.  Only below
            //      that does the warp look at who is blocked on a second leaf (nb) and who wants a new ray (nf).
            const bool inner = t.node >= 0;
            const unsigned mi = __ballot_sync(FULL, inner);
            if (__popc(mi) >= cont_t) continue;
            const int nb = __popc(__ballot_sync(FULL, !inner && pending != EMPTY));   // cannot go on without the leaf phase
            const int nf = __popc(__ballot_sync(FULL, !done && t.node == EMPTY && pending == EMPTY));
            if (mi != 0u && nb < leaf_t && nf < refill_t) continue;
            if (nb != 0 && (nb >= leaf_t || mi == 0u)) {
                // ---- leaf phase: every lane with a postponed leaf tests it
                if (pending != EMPTY) {
                    if (COUNT) st.tris++;
#if VCRT_SMEMRAY
                    Ray lr;
                    lr.o = f3(s_ray[0][threadIdx.x], s_ray[1][threadIdx.x], s_ray[2][threadIdx.x]);
                    lr.d = f3(s_ray[3][threadIdx.x], s_ray[4][threadIdx.x], s_ray[5][threadIdx.x]);
                    trav_leaf_test(t, s, lr, pending);
#else
                    trav_leaf_test(t, s, cur, pending);
#endif
                    pending = EMPTY;
                }
                {   // the leaf the lane was blocked on becomes the postponed one
                    const bool park = (uint32_t)t.node > 0x80000000u;
                    pending = park ? t.node : pending;
                    trav_pop_if(park, t, tos, sr);
                }
                continue;
            }
            break;   // enough lanes want a new ray (nf >= refill_t), or nothing but finished lanes is left
        }
    }
#if VCRT_TAIL_SPLIT
    // ---- tail: the queue is dry, every lane either still walks its last ray or idles.  A launch ends when its LONGEST ray ends, and
    // a ray that needs a few hundred visits is walked by one lane, one dependent round trip after the other, while the rest of
    // the GPU waits (~110 us per launch on C3: 2 % of a 64-spp step, 15 % of an 8-spp one, most of a 1-spp frame).  Here the idle
    // lanes of a warp take subtrees off the stacks of its busy lanes: a donor hands over the top of its stack (an inner node it has
    // not entered yet) together with its ray and its current closest hit; the helper walks that subtree with the donor's
    // arithmetic, and its result is merged into the donor's by the traversal's own rule (smaller t, then smaller slot), so the
    // record that is finally stored is the one a single lane would have found.  Helpers donate in turn; a lane's result is final
    // once its own walk is over and every helper it handed work to has reported back (`out`).
    {
        __shared__ uint32_t s_owner[VCRT_PBLOCK];
        const unsigned lane = threadIdx.x & 31u;
        const uint32_t wbase = threadIdx.x & ~31u;
        bool helper = false;
        uint32_t out = 0u;   // helpers that have not reported back yet
        for (;;) {
            bool working = t.node != EMPTY || pending != EMPTY;
            // (1) a lane whose own ray is finished for good stores it
            if (idx != 0xffffffffu && !working && out == 0u) {
                stream_st(w.hit + idx, make_uint2(f2u(t.closest), (uint32_t)t.best));
                idx = 0xffffffffu;
            }
            // (2) helpers report to their owners: those that are through for good (the owner then waits for one helper less), and --
            //     VCRT_TAIL_EXCHANGE -- all the others too, so that an owner culls with what its helpers have found so far and, the other
            //     way round, a helper with what its owner knows by now
#if VCRT_TAIL_EXCHANGE
            const unsigned fin = __ballot_sync(FULL, helper && !working && out == 0u);
            unsigned fh = __ballot_sync(FULL, helper);
#else
            unsigned fh = __ballot_sync(FULL, helper && !working && out == 0u);
            const unsigned fin = fh;
#endif
            const bool any_helper = fh != 0u;
            while (fh) {
                const int h = __ffs(fh) - 1;
                fh &= fh - 1u;
                const int o = (int)s_owner[wbase + (uint32_t)h];
                const float hc = __shfl_sync(FULL, t.closest, h);
                const int32_t hb = __shfl_sync(FULL, t.best, h);
                const bool done_h = ((fin >> h) & 1u) != 0u;
                if ((int)lane == o) {
                    if (hb >= 0 && (hc < t.closest || (hc == t.closest && hb < t.best))) { t.closest = hc; t.best = hb; }
                    if (done_h) out--;
                }
                if ((int)lane == h && done_h) helper = false;
            }
#if VCRT_TAIL_EXCHANGE
            if (any_helper) {
                const int o = helper ? (int)s_owner[threadIdx.x] : (int)lane;
                const float oc = __shfl_sync(FULL, t.closest, o);
                const int32_t ob = __shfl_sync(FULL, t.best, o);
                if (helper && ob >= 0 && (oc < t.closest || (oc == t.closest && ob < t.best))) { t.closest = oc; t.best = ob; }
            }
#endif
            // (3) idle lanes take the top-of-stack subtree of busy lanes
            const unsigned idle = __ballot_sync(FULL, idx == 0xffffffffu && !helper && !working);
            const unsigned donors = __ballot_sync(FULL, working && tos >= 0);
            if (!__any_sync(FULL, working || idx != 0xffffffffu || helper)) break;   // nobody holds a ray or works for one
#pragma unroll 1
            for (int pass = 0; pass < VCRT_TAIL_PASSES; ++pass) {
            const unsigned idle_now = pass == 0 ? idle : __ballot_sync(FULL, idx == 0xffffffffu && !helper && t.node == EMPTY && pending == EMPTY);
            const unsigned donors_now = pass == 0 ? donors : __ballot_sync(FULL, (t.node != EMPTY || pending != EMPTY) && tos >= 0);
            const int n = min(__popc(idle_now), __popc(donors_now));
            if (n == 0) break;
            {
                __syncwarp();   // the donors' rays in shared memory were written by other lanes of this warp
                const unsigned below = (1u << lane) - 1u;
                const int ri = __popc(idle_now & below), rd = __popc(donors_now & below);
                const bool recv = ((idle_now >> lane) & 1u) && ri < n, give = ((donors_now >> lane) & 1u) && rd < n;
                const int src = recv ? (int)__fns(donors_now, 0u, ri + 1) : (int)lane;
                const int32_t in_node = __shfl_sync(FULL, tos, src);
                const float in_c = __shfl_sync(FULL, t.closest, src);
                const int32_t in_b = __shfl_sync(FULL, t.best, src);
                const float ix = __shfl_sync(FULL, t.idir.x, src), iy = __shfl_sync(FULL, t.idir.y, src), iz = __shfl_sync(FULL, t.idir.z, src);
                const float ox = __shfl_sync(FULL, t.ood.x, src), oy = __shfl_sync(FULL, t.ood.y, src), oz = __shfl_sync(FULL, t.ood.z, src);
                const uint32_t sx = __shfl_sync(FULL, t.selx, src), sy = __shfl_sync(FULL, t.sely, src), sz = __shfl_sync(FULL, t.selz, src);
                if (give) {   // the donor's next stack entry becomes its top; it now waits for one more report
                    t.sp -= 1;
                    sr.load_if(true, t.sp, tos);
                    out++;
                }
                if (recv) {
                    t.node = in_node; t.closest = in_c; t.best = in_b;
                    t.idir = f3(ix, iy, iz); t.ood = f3(ox, oy, oz);
                    t.selx = sx; t.sely = sy; t.selz = sz;
                    t.sp = 1; tos = EMPTY; pending = EMPTY;
                    helper = true;
                    s_owner[threadIdx.x] = (uint32_t)src;
#if VCRT_SMEMRAY
#pragma unroll
                    for (int k = 0; k < 6; ++k) s_ray[k][threadIdx.x] = s_ray[k][wbase + (uint32_t)src];
#endif
                }
#if !VCRT_SMEMRAY
                cur.o = f3(__shfl_sync(FULL, cur.o.x, src), __shfl_sync(FULL, cur.o.y, src), __shfl_sync(FULL, cur.o.z, src));
                cur.d = f3(__shfl_sync(FULL, cur.d.x, src), __shfl_sync(FULL, cur.d.y, src), __shfl_sync(FULL, cur.d.z, src));
#endif
                __syncwarp();
            }
            }
            // (4) a few rounds of: visit, postpone, test -- no votes: with a few lanes left what counts is each lane's own latency
#pragma unroll 1
            for (int round = 0; round < VCRT_TAIL_ROUNDS; ++round) {
                if (t.node >= 0) {
                    if (COUNT) st.nodes++;
                    if (QN == 2) {
                        const Words8* p = s.q4nodes + 2 * (size_t)t.node;
                        const Words8 na = ldg8(p), nb = ldg8(p + 1);
                        int32_t c[4];
                        trav_test4(t, na, nb, c);
#if VCRT_TAIL_PREFETCH
                        // In the tail a lane's speed is the latency of its dependent round trips (an L2 hit per visit and per triangle
                        // test), and the L1TEX pipe that a divergent prefetch costs in the main loop is idle: the records of every child
                        // the ray enters are pulled into L1 now, one visit's arithmetic ahead of their use (or several, for the stack).
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (c[i] != EMPTY) {
                                const char* rec = c[i] >= 0 ? (const char*)(s.q4nodes + 2 * (size_t)c[i]) : (const char*)(s.ftris + 4 * (size_t)(~c[i]));
                                asm volatile("prefetch.global.L1 [%0];" ::"l"(rec));
                                asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + 32));
                            }
#endif
                        trav_descend4(t, s, c, pending, tos, sr);
                    } else {
                        float lN, rN;
                        bool hl, hr;
                        int32_t cl, cr;
                        trav_test_children<QN>(t, s, lN, rN, hl, hr, cl, cr);
                        trav_descend(t, lN, rN, hl, hr, cl, cr, pending, tos, sr);
                    }
                }
                {
                    const bool park = (uint32_t)t.node > 0x80000000u && pending == EMPTY;
                    pending = park ? t.node : pending;
                    trav_pop_if(park, t, tos, sr);
                }
                if (pending != EMPTY) {
                    if (COUNT) st.tris++;
#if VCRT_SMEMRAY
                    Ray lr;
                    lr.o = f3(s_ray[0][threadIdx.x], s_ray[1][threadIdx.x], s_ray[2][threadIdx.x]);
                    lr.d = f3(s_ray[3][threadIdx.x], s_ray[4][threadIdx.x], s_ray[5][threadIdx.x]);
                    trav_leaf_test(t, s, lr, pending);
#else
                    trav_leaf_test(t, s, cur, pending);
#endif
                    pending = EMPTY;
                    const bool park = (uint32_t)t.node > 0x80000000u;
                    pending = park ? t.node : pending;
                    trav_pop_if(park, t, tos, sr);
                }
            }
        }
    }
#endif
    // the queue the shade kernel behind this launch appends to starts empty (this kernel does not touch it; the fetch counter this
    // kernel uses is cleared by the shade kernel in turn: no separate reset launch per bounce)
    if (blockIdx.x == 0 && threadIdx.x == 0) w.counts[b.cur ^ 1u] = 0u;
    // st.rays is non-zero only for PRIMARY (queued rays are counted once per launch above): every traced pixel answers the
    // bounce-0 closest-hit query of all its samples
    flush_stats(a, st, PRIMARY ? a.sample_count : 1u, PRIMARY);
}

// ---- shade: one thread per traced ray (ray_color body, ray-trace-compute.comp:321-340)
template <int SHADER, int RNG_MODE, int TRIG, bool PRIMARY>
__global__ void __launch_bounds__(256, VCRT_SHADE_MINB) wf_shade_kernel(const __grid_constant__ KernelArgs a, const WfQueues w, const WfBatch b) {
    const SceneView& s = a.scene;
    const uint32_t count = PRIMARY ? b.npaths : w.counts[b.cur];
    const float4* __restrict__ rays = w.q[b.cur];
    float4* __restrict__ next = w.q[b.cur ^ 1u];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        w.counts[2] = 0u;   // the trace kernel's fetch counter, for the next bounce (the trace launch that used it is over)
        if (!PRIMARY && count) {   // the rays the preceding trace launch walked: one query = one traversal each
            atomicAdd(a.counters + 0, (unsigned long long)count);
            atomicAdd(a.counters + 4, (unsigned long long)count);
        }
    }
    for (uint32_t base = blockIdx.x * 256u; base < count; base += gridDim.x * 256u) {   // base is warp-uniform
        const uint32_t i = base + threadIdx.x;
        bool cont = false, hit = false;
        float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0;
        uint32_t x = 0, y = 0, k = 0, path = 0, rng = 0, pix = 0;
        Ray cur; cur.o = cur.d = f3(0, 0, 0);
        float3 thr = f3(0, 0, 0);
        Hit rec;
        const uint32_t item = PRIMARY ? i / a.sample_count : 0u;   // bounce 0: the pixel whose (single) primary hit this path shares
        const bool live = i < count && (!PRIMARY || wf_path_pixel(a, b, i, x, y, k));
        if (live) {
            if (PRIMARY) {
                path = i;
                wf_primary(a, x, y, k, cur, rng);
                thr = f3(1.0f, 1.0f, 1.0f);
                q0 = make_float4(0, 0, 0, u2f(path));
            } else {
                q0 = stream_ld(rays + 3 * (size_t)i); q1 = stream_ld(rays + 3 * (size_t)i + 1); q2 = stream_ld(rays + 3 * (size_t)i + 2);
                path = f2u(q0.w); rng = f2u(q1.w);
                cur.o = xyz(q0); cur.d = xyz(q1);
                thr = xyz(q2);
                wf_path_pixel(a, b, path, x, y, k);
            }
            const uint2 h = PRIMARY ? w.hit[item] : stream_ld(w.hit + i);   // PRIMARY: read by every sample of the pixel, keep it cached
            TravState t;
            t.closest = u2f(h.x); t.best = (int32_t)h.y;
            hit = trav_finish(t, s, cur, rec);
            pix = y * a.W + x;
            if (b.bounce == 0 && k == 0 && (a.flags & VCRT_FLAG_WRITE_AOV)) {
                vcrt_aov o;
                if (hit) { o.triangle = rec.triangle; o.material = (int32_t)rec.materialIndex; o.t = rec.t; o.backFace = (uint32_t)rec.backFaceInt; }
                else { o.triangle = -1; o.material = -1; o.t = 0.0f; o.backFace = 0u; }
                a.aov[pix] = o;
            }
            if (hit) {   // whether the path goes on depends on the material type alone (scatter returns type == LIGHT)
                uint32_t type; float3 unused;
                load_mat(s, rec.materialIndex, type, unused);
                cont = type != VCRT_MAT_LIGHT && (b.bounce + 1u) < a.env.max_bounces;
            }
        }
        // reserve the slots of the next queue now; the answer is needed only for the stores at the end
        unsigned am;
        const unsigned abase = wf_append_begin(w.counts + (b.cur ^ 1u), cont, am);
        if (live) {
            if (hit) {
                Rng g;
                g.pcg = rng;
                g.key0 = pix; g.key1 = a.philox_seed; g.ctr0 = a.sample_begin + k;
                rng_begin_bounce<RNG_MODE>(g, b.bounce);
                float3 albedo;
                Ray nx;
                scatter<SHADER, RNG_MODE, TRIG>(s, a.env, cur, rec, albedo, nx, g);
                thr = mul(thr, albedo);
                q0 = make_float4(nx.o.x, nx.o.y, nx.o.z, q0.w);
                q1 = make_float4(nx.d.x, nx.d.y, nx.d.z, u2f(g.pcg));
                q2 = make_float4(thr.x, thr.y, thr.z, 0.0f);
            } else {
                thr = scale(thr, 0.0f);
            }
            if (!cont) stream_st(w.sample_color + path, make_float4(thr.x, thr.y, thr.z, 1.0f));
        }
        const uint32_t slot = wf_append_end(abase, am);
        if (cont) {
            float4* q = next + 3 * (size_t)slot;
            stream_st(q, q0); stream_st(q + 1, q1); stream_st(q + 2, q2);
        }
    }
}

// ---- accumulate: one thread per work item, samples folded in order (ray-trace-compute.comp:375-379)
__global__ void __launch_bounds__(256) wf_accumulate_kernel(const __grid_constant__ KernelArgs a, const WfQueues w, const WfBatch b) {
    const uint32_t it = blockIdx.x * 256u + threadIdx.x;
    if (it >= b.nitems) return;
    uint32_t x, y;
    if (!item_to_pixel(a, b.item0 + it, x, y)) return;
    const uint32_t pix = y * a.W + x;
    const float4* c = w.sample_color + (size_t)it * a.sample_count;
    const bool restart = (a.flags & VCRT_FLAG_INTERNAL_RESTART) != 0u;   // frames in flight: the accumulation starts over with this frame
    if (a.accum_mode == VCRT_ACCUM_F32) {
        float4 acc = restart ? make_float4(0, 0, 0, 0) : a.accumf[pix];
        for (uint32_t k = 0; k < a.sample_count; ++k) { const float4 v = stream_ld(c + k); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += 1.0f; }
        a.accumf[pix] = acc;
    } else {
        uchar4 px = a.accum8[pix];
        for (uint32_t k = 0; k < a.sample_count; ++k) { const float4 v = stream_ld(c + k); running_mean_rgba8(px, f3(v.x, v.y, v.z), a.sample_begin + k); }
        a.target[pix] = px;
        a.accum8[pix] = px;
    }
}

}  // namespace vcrt
