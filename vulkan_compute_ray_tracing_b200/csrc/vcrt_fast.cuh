// vcrt_fast.cuh -- fast closest-hit traversal over records repacked from the bound bvh[] / triangles[].
//
// Same result as hit_bvh (ray-trace-compute.comp:263-311), by construction:
//   * the accept/reject arithmetic of every triangle is tri_test -- the reference's operation order, no FMA;
//   * the reference keeps the FIRST leaf in its fixed right-child-first DFS order among equal-t hits (strict
//     `t < closest_so_far`); triangles are stored in exactly that order, so the slot index is the tie rank and
//     `t == closest && slot < best` reproduces the rule under any visiting order;
//   * boxes only cull.  The reference pads every box by 1e-4 (Bvh.h:16,94-98), orders of magnitude more than
//     the rounding differences between its divide-based slab test and the fma slab test here, so both visit
//     every leaf that holds an acceptable hit.  Culling with `tNear > closest` (strict) keeps equal-t leaves.
//
// Layout (built by vcrt_repack.cpp from the reference arrays, all records 16-byte aligned, 128-bit loads):
//   inner node, 64 B:  {L.min.x L.max.x L.min.y L.max.y} {R.min.x R.max.x R.min.y R.max.y}
//                      {L.min.z L.max.z R.min.z R.max.z} {childL childR - -}
//                      child >= 0: inner node index;  child < 0: leaf, triangle slot = ~child;  empty: box (+inf,-inf)
//   triangle, 64 B:    {v0.xyz, original index} {a.xyz, materialIndex} {b.xyz, -} {n.xyz, -}   in reference DFS leaf order;
//                      a = v0 - v1, b = v2 - v0, n = cross(b, a): the ray-independent part of triIntersect, precomputed with
//                      the shader's own fp32 operations (vcrt_repack.h: precompute_triangles); two 256-bit loads per test
//   quantised inner node, 32 B (QN = 1; used when the scene extent allows, vcrt_repack.h): the same twelve bounds as
//                      15-bit fixed point in a scene-wide frame, rounded outwards, + the two child codes; one 256-bit load.
//                      Boxes only grow, so the visited set is a superset of the float nodes' and results are unchanged.
#pragma once

#include "vcrt_core.cuh"
#include "vcrt_tunables.h"

namespace vcrt {

#ifndef VCRT_FAST_STACK
#define VCRT_FAST_STACK 96   /* traversal-stack entries per ray (local memory, touched only as deep as a ray goes).  The 4-wide tree's exact worst case is computed
                                at build time: 30-40 for the host's SAH trees, 73 for the device's (deeper) PLOC tree of the 10 M-triangle scene */
#endif
#define VCRT_FAST_EMPTY ((int32_t)0x80000000)

VCRT_HD float fmin_(float a, float b) { return fminf(a, b); }
VCRT_HD float fmax_(float a, float b) { return fmaxf(a, b); }

// Per-ray traversal state, advanced one node at a time so that the persistent kernel (vcrt_persistent.cuh) can
// interleave the rays of a warp; hit_bvh_fast below runs the same steps to completion for one ray.
struct TravState {
    float3 idir, ood;       // 1/d (guarded) and o/d for the fma slab test; quantised nodes: 2*qext/d and (o - qorg)/d
    uint32_t selx, sely, selz;   // quantised nodes: PRMT selector of the NEAR plane per axis (far = sel ^ 0x0220)
    float closest;
    int32_t best;           // winning triangle slot or -1
    int32_t node;           // >= 0 inner node to visit, < 0 leaf (~slot), VCRT_FAST_EMPTY: nothing left
    int sp;
};

template <int QN>
VCRT_HD void trav_begin(TravState& t, const SceneView& s, const Ray& r) {
    const float tiny = 1e-30f;   // guard so that 0 * inf never appears
    t.idir = f3(1.0f / (fabsf(r.d.x) > tiny ? r.d.x : copysignf(tiny, r.d.x)),
                1.0f / (fabsf(r.d.y) > tiny ? r.d.y : copysignf(tiny, r.d.y)),
                1.0f / (fabsf(r.d.z) > tiny ? r.d.z : copysignf(tiny, r.d.z)));
    if (QN) {
        // t(bound) = (qorg + 2m*qext - o) / d = m * (2*qext/d) - (o - qorg)/d
        t.ood = f3((r.o.x - s.qorg.x) * t.idir.x, (r.o.y - s.qorg.y) * t.idir.y, (r.o.z - s.qorg.z) * t.idir.z);
        t.idir = f3(2.0f * s.qext.x * t.idir.x, 2.0f * s.qext.y * t.idir.y, 2.0f * s.qext.z * t.idir.z);
        // the plane entered first along an axis is the min plane when the ray travels in +axis, else the max plane
        t.selx = t.idir.x >= 0.0f ? VCRT_Q15_SEL_LO : VCRT_Q15_SEL_HI;
        t.sely = t.idir.y >= 0.0f ? VCRT_Q15_SEL_LO : VCRT_Q15_SEL_HI;
        t.selz = t.idir.z >= 0.0f ? VCRT_Q15_SEL_LO : VCRT_Q15_SEL_HI;
    } else {
        t.ood = f3(r.o.x * t.idir.x, r.o.y * t.idir.y, r.o.z * t.idir.z);
    }
    t.closest = VCRT_T_MAX;
    t.best = -1;
    t.node = QN == 2 ? s.froot4 : s.froot;
    t.sp = 0;
}

// The two children of a quantised 32-byte node (already in registers) against the ray: near/far planes are picked by the
// direction sign at decode time, so the slab test needs no per-axis min/max.
VCRT_HD void trav_test_children_q(const TravState& t, const Words8& n, float& lN, float& rN, bool& hl, bool& hr, int32_t& cl, int32_t& cr) {
    const uint32_t fx = t.selx ^ 0x0220u, fy = t.sely ^ 0x0220u, fz = t.selz ^ 0x0220u;
    const float lnx = fmaf(q15_sel(n.w[0], t.selx), t.idir.x, -t.ood.x), lfx = fmaf(q15_sel(n.w[0], fx), t.idir.x, -t.ood.x);
    const float lny = fmaf(q15_sel(n.w[1], t.sely), t.idir.y, -t.ood.y), lfy = fmaf(q15_sel(n.w[1], fy), t.idir.y, -t.ood.y);
    const float lnz = fmaf(q15_sel(n.w[2], t.selz), t.idir.z, -t.ood.z), lfz = fmaf(q15_sel(n.w[2], fz), t.idir.z, -t.ood.z);
    const float rnx = fmaf(q15_sel(n.w[3], t.selx), t.idir.x, -t.ood.x), rfx = fmaf(q15_sel(n.w[3], fx), t.idir.x, -t.ood.x);
    const float rny = fmaf(q15_sel(n.w[4], t.sely), t.idir.y, -t.ood.y), rfy = fmaf(q15_sel(n.w[4], fy), t.idir.y, -t.ood.y);
    const float rnz = fmaf(q15_sel(n.w[5], t.selz), t.idir.z, -t.ood.z), rfz = fmaf(q15_sel(n.w[5], fz), t.idir.z, -t.ood.z);
    cl = (int32_t)n.w[6]; cr = (int32_t)n.w[7];
    lN = fmax_(fmax_(lnx, lny), fmax_(lnz, 0.0f));
    rN = fmax_(fmax_(rnx, rny), fmax_(rnz, 0.0f));
    const float lF = fmin_(fmin_(lfx, lfy), lfz) * 1.0000004f;
    const float rF = fmin_(fmin_(rfx, rfy), rfz) * 1.0000004f;
    hl = lN <= fmin_(lF, t.closest);
    hr = rN <= fmin_(rF, t.closest);
}
// Both children of inner node t.node against the ray: entry distances (clamped to 0), hit verdicts and child codes.
template <int QN>
VCRT_HD void trav_test_children(const TravState& t, const SceneView& s, float& lN, float& rN, bool& hl, bool& hr, int32_t& cl, int32_t& cr) {
    if (QN) {
        trav_test_children_q(t, ldg8(s.qnodes + t.node), lN, rN, hl, hr, cl, cr);
        return;
    }
    float lx0, lx1, ly0, ly1, lz0, lz1, rx0, rx1, ry0, ry1, rz0, rz1;
    {
        const float4* p = s.fnodes + 4 * (size_t)t.node;
        const float4 n0 = ldg4(p), n1 = ldg4(p + 1), n2 = ldg4(p + 2), n3 = ldg4(p + 3);
        lx0 = fmaf(n0.x, t.idir.x, -t.ood.x); lx1 = fmaf(n0.y, t.idir.x, -t.ood.x);
        ly0 = fmaf(n0.z, t.idir.y, -t.ood.y); ly1 = fmaf(n0.w, t.idir.y, -t.ood.y);
        lz0 = fmaf(n2.x, t.idir.z, -t.ood.z); lz1 = fmaf(n2.y, t.idir.z, -t.ood.z);
        rx0 = fmaf(n1.x, t.idir.x, -t.ood.x); rx1 = fmaf(n1.y, t.idir.x, -t.ood.x);
        ry0 = fmaf(n1.z, t.idir.y, -t.ood.y); ry1 = fmaf(n1.w, t.idir.y, -t.ood.y);
        rz0 = fmaf(n2.z, t.idir.z, -t.ood.z); rz1 = fmaf(n2.w, t.idir.z, -t.ood.z);
        cl = (int32_t)f2u(n3.x); cr = (int32_t)f2u(n3.y);
    }
    lN = fmax_(fmax_(fmin_(lx0, lx1), fmin_(ly0, ly1)), fmax_(fmin_(lz0, lz1), 0.0f));
    const float lF = fmin_(fmin_(fmax_(lx0, lx1), fmax_(ly0, ly1)), fmax_(lz0, lz1)) * 1.0000004f;
    rN = fmax_(fmax_(fmin_(rx0, rx1), fmin_(ry0, ry1)), fmax_(fmin_(rz0, rz1), 0.0f));
    const float rF = fmin_(fmin_(fmax_(rx0, rx1), fmax_(ry0, ry1)), fmax_(rz0, rz1)) * 1.0000004f;
    // an absent child is stored as the box (+inf, -inf), which the min/max slab test above would report as entered at 0 and
    // left at +inf: the verdict is gated on the child code (the quantised formats need no gate: their empty box is empty)
    hl = cl != VCRT_FAST_EMPTY && lN <= fmin_(lF, t.closest);
    hr = cr != VCRT_FAST_EMPTY && rN <= fmin_(rF, t.closest);
}

// ---- 4-wide nodes (vcrt_repack.h: build_wide_bvh): a node is two halves in the quantised binary format.
// Tests the four children and returns them ordered by entry distance: c[0] the nearest hit child ... ; children that are
// missed (or absent) come last with code VCRT_FAST_EMPTY.  Any order is correct (ties are decided by slot rank); nearest
// first makes t-culling effective.
VCRT_HD void trav_test4(const TravState& t, const Words8& a, const Words8& b, int32_t c[4]) {
    float d[4];
    bool h[4];
    trav_test_children_q(t, a, d[0], d[1], h[0], h[1], c[0], c[1]);
    trav_test_children_q(t, b, d[2], d[3], h[2], h[3], c[2], c[3]);
    const float inf = u2f(0x7f800000u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { d[i] = h[i] ? d[i] : inf; c[i] = h[i] ? c[i] : VCRT_FAST_EMPTY; }
    // 5-comparator network; strict compare, so absent children (inf) never move in front of anything
#define VCRT_CSWAP(i, j) { const bool sw = d[j] < d[i]; const float dl = sw ? d[j] : d[i], dh = sw ? d[i] : d[j]; const int32_t cl_ = sw ? c[j] : c[i], ch_ = sw ? c[i] : c[j]; d[i] = dl; d[j] = dh; c[i] = cl_; c[j] = ch_; }
    VCRT_CSWAP(0, 1) VCRT_CSWAP(2, 3) VCRT_CSWAP(0, 2) VCRT_CSWAP(1, 3) VCRT_CSWAP(1, 2)
#undef VCRT_CSWAP
}

// Visit 4-wide inner node t.node (generic form: plain stack; the static kernels and the host emulation).
VCRT_HD void trav_inner_step4(TravState& t, const SceneView& s, int32_t* stack) {
    const Words8* p = s.q4nodes + 2 * (size_t)t.node;
    const Words8 a = ldg8(p), b = ldg8(p + 1);
    int32_t c[4];
    trav_test4(t, a, b, c);
    if (c[0] == VCRT_FAST_EMPTY) { t.node = t.sp ? stack[--t.sp] : VCRT_FAST_EMPTY; return; }
    t.node = c[0];
    for (int i = 3; i >= 1; --i)
        if (c[i] != VCRT_FAST_EMPTY && t.sp < VCRT_FAST_STACK) stack[t.sp++] = c[i];
}

// Visit inner node t.node: test both children, descend into the nearer hit child, push the farther one.
template <int QN>
VCRT_HD void trav_inner_step(TravState& t, const SceneView& s, int32_t* stack) {
    float lN, rN;
    bool hl, hr;
    int32_t cl, cr;
    trav_test_children<QN>(t, s, lN, rN, hl, hr, cl, cr);
    if (hl && hr) {
        const bool left_first = lN <= rN;
        t.node = left_first ? cl : cr;
        if (t.sp < VCRT_FAST_STACK) stack[t.sp++] = left_first ? cr : cl;
    } else if (hl) {
        t.node = cl;
    } else if (hr) {
        t.node = cr;
    } else {
        t.node = t.sp ? stack[--t.sp] : VCRT_FAST_EMPTY;
    }
}

// The same visit for the wavefront trace kernel (device only), written for predication instead of branches, with the top
// of the stack held in a register (`tos`): a pop takes its node from the register and issues the load of the next entry,
// whose latency is then off the critical path.  stack[0] holds the sentinel VCRT_FAST_EMPTY and t.sp starts at 1 with
// tos = EMPTY, so an exhausted stack yields EMPTY without a test; depth never exceeds the stack (vcrt_repack.cpp rejects
// deeper trees), so "push" needs no bound check.
// The stack accesses are predicated ld.local / st.local written in PTX with the load's destination tied to `tos`: left to
// itself the compiler loads the popped entry into a temporary and moves it into the tos register right away, and that
// move waits out the whole load (ncu r01_v9: 16 % of the trace kernel's stall samples sat on those moves).  `sbase` is
// the local-window byte address of stack[0] (__cvta_generic_to_local).
#ifdef __CUDACC__
__device__ __forceinline__ void stk_load_if(bool p, int32_t& dst, uint32_t addr) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.local.b32 %0, [%1];\n\t}" : "+r"(dst) : "r"(addr), "r"((uint32_t)p) : "memory");
}
__device__ __forceinline__ void stk_store_if(bool p, uint32_t addr, int32_t v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.local.b32 [%0], %1;\n\t}" ::"r"(addr), "r"(v), "r"((uint32_t)p) : "memory");
}
__device__ __forceinline__ void stk_lds_if(bool p, int32_t& dst, uint32_t addr) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.shared.b32 %0, [%1];\n\t}" : "+r"(dst) : "r"(addr), "r"((uint32_t)p) : "memory");
}
__device__ __forceinline__ void stk_sts_if(bool p, uint32_t addr, int32_t v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.shared.b32 [%0], %1;\n\t}" ::"r"(addr), "r"(v), "r"((uint32_t)p) : "memory");
}
// Where the entries below the register-held top live.  The first VCRT_SSTACK slots of every lane are in shared memory,
// laid out [slot][thread]: whatever the lanes' depths, a warp's access touches 32 different banks = one wavefront, where
// the same access to local memory costs one L1 sector per distinct depth among the lanes (ncu r01_v9: 3.2 sectors per
// request, and stack traffic was 37 % of all L1 sectors of the kernel).  Deeper slots fall back to local memory.
struct StackRef {
    uint32_t lbase;   // local-window byte address of the local part
    uint32_t sbase;   // shared-window byte address of this thread's slot 0
    __device__ __forceinline__ void store_if(bool p, int sp, int32_t v) const {
#if VCRT_SSTACK
        const bool sh = sp < VCRT_SSTACK;
        stk_sts_if(p && sh, sbase + (uint32_t)sp * (4u * VCRT_PBLOCK), v);
        stk_store_if(p && !sh, lbase + 4u * (uint32_t)(sp - VCRT_SSTACK), v);
#else
        stk_store_if(p, lbase + 4u * (uint32_t)sp, v);
#endif
    }
    __device__ __forceinline__ void load_if(bool p, int sp, int32_t& dst) const {
#if VCRT_SSTACK
        const bool sh = sp < VCRT_SSTACK;
        stk_lds_if(p && sh, dst, sbase + (uint32_t)sp * (4u * VCRT_PBLOCK));
        stk_load_if(p && !sh, dst, lbase + 4u * (uint32_t)(sp - VCRT_SSTACK));
#else
        stk_load_if(p, dst, lbase + 4u * (uint32_t)sp);
#endif
    }
};
// node = tos; tos = stack[--sp], for the lanes where p holds
__device__ __forceinline__ void trav_pop_if(bool p, TravState& t, int32_t& tos, const StackRef& sr) {
    t.node = p ? tos : t.node;
    t.sp -= (int)p;
    sr.load_if(p, t.sp, tos);
}
// The stack side of a visit, given the verdicts on the two children.  The near child, when it is a leaf and nothing is
// parked yet, goes straight into `pending` and the lane goes on with the far child (or with the stack): no push + pop
// round trip through local memory for it.
__device__ __forceinline__ void trav_descend(TravState& t, float lN, float rN, bool hl, bool hr, int32_t cl, int32_t cr, int32_t& pending, int32_t& tos, const StackRef& sr) {
    const bool left_first = hl & (!hr | (lN <= rN));   // bitwise on purpose: one predicate LUT instead of materialised booleans
    const int32_t near_c = left_first ? cl : cr, far_c = left_first ? cr : cl;
    const bool both = hl && hr, none = !(hl || hr);
    const bool park = !none && near_c < 0 && pending == VCRT_FAST_EMPTY;
    pending = park ? near_c : pending;
    const bool push = both && !park, pop = none || (park && !both);
    sr.store_if(push, t.sp, tos);   // push: the old top goes to memory, the far child becomes the top
    t.sp += (int)push;
    t.node = pop ? tos : (park ? far_c : near_c);
    tos = push ? far_c : tos;
    t.sp -= (int)pop;                                        // pop: the top becomes the node, the next entry is loaded into tos
    sr.load_if(pop, t.sp, tos);
}
// The stack side of a 4-wide visit.  c[] as returned by trav_test4.  The nearest child, when it is a leaf and nothing is
// parked yet, goes into `pending`; of the remaining k children the nearest becomes the node, the others go onto the stack
// farthest first (the old top moves to memory, the second nearest becomes the new top).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void trav_descend4(TravState& t, const SceneView& s, int32_t c[4], int32_t& pending, int32_t& tos, const StackRef& sr) {
    const bool park = c[0] != VCRT_FAST_EMPTY && c[0] < 0 && pending == VCRT_FAST_EMPTY;
#if VCRT_PF_L2
    // Working sets beyond L2 (the 10 M-triangle scene): records that will be needed a few iterations from now -- the leaf just
    // parked and the child that becomes the top of the stack -- are pulled into L2 now, so that their DRAM latency overlaps the
    // visits in between instead of stalling the lane's warp when its turn comes (no register, no L1 allocation).
    if (park) prefetch_l2(s.ftris + 4 * (size_t)(~c[0]));
#endif
    pending = park ? c[0] : pending;
    c[0] = park ? c[1] : c[0]; c[1] = park ? c[2] : c[1]; c[2] = park ? c[3] : c[2]; c[3] = park ? VCRT_FAST_EMPTY : c[3];
    const bool k1 = c[0] != VCRT_FAST_EMPTY, k2 = c[1] != VCRT_FAST_EMPTY, k3 = c[2] != VCRT_FAST_EMPTY, k4 = c[3] != VCRT_FAST_EMPTY;
#if VCRT_PF_L2
    if (k2) prefetch_l2(c[1] >= 0 ? (const void*)(s.q4nodes + 2 * (size_t)c[1]) : (const void*)(s.ftris + 4 * (size_t)(~c[1])));
#endif
    sr.store_if(k2, t.sp, tos);
    sr.store_if(k3, t.sp + 1, k4 ? c[3] : c[2]);
    sr.store_if(k4, t.sp + 2, c[2]);
    t.sp += (int)k2 + (int)k3 + (int)k4;
    t.node = k1 ? c[0] : tos;
    tos = k2 ? c[1] : tos;
    t.sp -= (int)!k1;                                        // nothing left here: pop
    sr.load_if(!k1, t.sp, tos);
}
#endif

// Test the triangle of leaf code `leaf` (= ~slot) with the reference's arithmetic and tie rule.
VCRT_HD void trav_leaf_test(TravState& t, const SceneView& s, const Ray& r, int32_t leaf) {
    const int32_t slot = ~leaf;
    const Words8* p = (const Words8*)(s.ftris + 4 * (size_t)slot);
    const Words8 lo = ldg8(p), hi = ldg8(p + 1);
    const float3 v0 = f3(u2f(lo.w[0]), u2f(lo.w[1]), u2f(lo.w[2])), ea = f3(u2f(lo.w[4]), u2f(lo.w[5]), u2f(lo.w[6]));
    const float3 eb = f3(u2f(hi.w[0]), u2f(hi.w[1]), u2f(hi.w[2])), n = f3(u2f(hi.w[4]), u2f(hi.w[5]), u2f(hi.w[6]));
    float tt;
    if (tri_test_pre(v0, ea, eb, n, r, tt) && tt > VCRT_T_MIN && (tt < t.closest || (tt == t.closest && slot < t.best))) {
        t.closest = tt;
        t.best = slot;
    }
}

VCRT_HD bool trav_finish(const TravState& t, const SceneView& s, const Ray& r, Hit& rec) {
    if (t.best < 0) return false;
    const float4* p = s.ftris + 4 * (size_t)t.best;
    const float4 a = ldg4(p), b = ldg4(p + 1), d = ldg4(p + 3);
    finish_triangle_hit_pre(xyz(d), f2u(b.w), (int)f2u(a.w), r, t.closest, rec);
    return true;
}

template <bool COUNT, int QN>
VCRT_HD bool hit_bvh_fast_q(const SceneView& s, const Ray& r, Hit& rec, TraceStats& st) {
    TravState t;
    trav_begin<QN>(t, s, r);
    int32_t stack[VCRT_FAST_STACK];
    while (t.node != VCRT_FAST_EMPTY) {
        if (t.node >= 0) {
            if (COUNT) st.nodes++;
            if (QN == 2) trav_inner_step4(t, s, stack);
            else trav_inner_step<QN>(t, s, stack);
        } else {
            if (COUNT) st.tris++;
            trav_leaf_test(t, s, r, t.node);
            t.node = t.sp ? stack[--t.sp] : VCRT_FAST_EMPTY;
        }
    }
    return trav_finish(t, s, r, rec);
}

template <bool COUNT>
VCRT_HD bool hit_bvh_fast(const SceneView& s, const Ray& r, Hit& rec, TraceStats& st) {
    if (s.q4nodes) return hit_bvh_fast_q<COUNT, 2>(s, r, rec, st);
    return s.qnodes ? hit_bvh_fast_q<COUNT, 1>(s, r, rec, st) : hit_bvh_fast_q<COUNT, 0>(s, r, rec, st);
}

}  // namespace vcrt
