// vcrt_repack.h -- reference bvh[]/triangles[] -> records of the fast traversal (layout: vcrt_fast.cuh).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/vcrt.h"

namespace vcrt {

struct FastBvh {
    std::vector<float> nodes;      // 16 floats per inner node
    std::vector<float> tris;       // 12 floats per triangle slot: {v0, original index} {v1, materialIndex} {v2, -}
    std::vector<float> tris64;     // 16 floats per slot, what the kernels read (precompute_triangles): {v0, index} {a, material} {b, -} {n, -}
    std::vector<uint32_t> qnodes;  // 8 words per inner node (quantised form of `nodes`, see quantize_fast_bvh); empty = not quantised
    std::vector<uint32_t> q4nodes; // 16 words per 4-wide node (build_wide_bvh); empty = not built
    int32_t root4 = (int32_t)0x80000000;   // root of the 4-wide tree: node index, leaf code or "empty", like `root`
    uint32_t stack4 = 0;           // deepest traversal stack the 4-wide tree can ask for (entries)
    float qorg[3] = {0, 0, 0};     // decode frame: coordinate = qorg + (2m) * qext, m = 0.5 * (1 + q / 32768) in [0.5, 1)
    float qext[3] = {0, 0, 0};
    int32_t root = (int32_t)0x80000000;
    uint32_t depth = 0;            // deepest leaf (root = 0)
    uint32_t bound_depth = 0;      // deepest leaf of the bound bvh[] itself (what the reference's 16-entry stack has to cope with)
    uint32_t num_nodes() const { return (uint32_t)(nodes.size() / 16); }
    uint32_t num_slots() const { return (uint32_t)(tris.size() / 12); }
};

// Walks the tree in the reference's visiting order (right child first, ray-trace-compute.comp:301-306) so that
// triangle slots are numbered by the reference's tie rank.  Returns false (err set) for trees the fast traversal
// does not represent: cycles / shared subtrees, leaves that also have children.  The depth of the tree that will be walked
// (the bound topology, or the SAH rebuild's) is checked separately by check_fast_depth.
bool build_fast_bvh(const vcrt_bvh_node* bvh, uint32_t nbvh, const vcrt_triangle* tris, uint32_t ntris, FastBvh& out, std::string& err);
// False (err set) when fb.depth exceeds what the binary traversal stack holds.  Call it on the tree the kernels will walk:
// a deep or degenerate bound tree is fine as long as the SAH rebuild (depth <= 24 + log2 n) replaces its topology.
bool check_fast_depth(const FastBvh& fb, std::string& err);

// Same records, but over a topology built here: binned-SAH top-down build over the leaves collected above (their
// triangles keep their slots = the reference's tie ranks, so results do not change; only the number of nodes a ray
// visits does).  The reference's builder splits at the median of a random axis (Bvh.h:160,175), which costs several
// times more node visits per ray than a surface-area-heuristic tree.  Parallel over subtrees (OpenMP tasks).
bool rebuild_fast_bvh_sah(FastBvh& fb, std::string& err);
// Insertion-based optimisation of the tree rebuild_fast_bvh_sah made (Bittner et al. 2013): `passes` times, the `fraction` of the inner
// nodes with the largest surface area is taken out and their children are put back where the tree's total surface area grows least.
// Leaves keep their slots: results do not change.  A result too deep for the traversal stack is discarded (the tree stays as built).
bool optimize_fast_bvh_reinsert(FastBvh& fb, int passes, float fraction, std::string& err);

// 32-byte form of the inner nodes: the twelve child-box bounds as 15-bit fixed point in one scene-wide frame, rounded
// outwards (boxes only grow, so culling stays conservative and results do not change), followed by the two child codes:
//   {Lx0|Lx1<<16, Ly0|Ly1<<16, Lz0|Lz1<<16, Rx0|Rx1<<16, Ry0|Ry1<<16, Rz0|Rz1<<16, childL, childR}
// One 256-bit load per visit instead of four 128-bit ones (the trace kernel is bound by L1 data-pipe wavefronts,
// DESIGN.md section 6).  Done only when the quantum (scene extent / 32766) is at most `max_quantum` on every axis --
// the reference pads every leaf box by 1e-4 (Bvh.h:16), so quanta of that order cost a few per cent more triangle tests;
// beyond that the 64-byte float nodes are kept.  Returns whether the quantised form was produced.
bool quantize_fast_bvh(FastBvh& fb, float max_quantum);

// 4-wide form of the quantised tree, for the wavefront trace kernel: every node holds up to four children, found by
// repeatedly opening the child with the largest surface area of a binary node's two children (leaves stay leaves, one
// triangle each, slots unchanged).  A visit then decides four boxes after ONE dependent fetch, and a ray needs about half
// as many fetch round trips as in the binary tree -- the trace kernel's time goes into waiting for the slowest lane of
// each round trip.  Record, 64 bytes = two halves in the 32-byte binary format above:
//   {c0.x, c0.y, c0.z, c1.x, c1.y, c1.z, code0, code1} {c2.x, c2.y, c2.z, c3.x, c3.y, c3.z, code2, code3}
// same frame (qorg/qext), same outward rounding, absent children = the empty box with the "empty" code.
// Requires quantize_fast_bvh to have succeeded.  Returns false (and leaves q4nodes empty) when the tree could ask for
// more than `max_stack` stack entries.
bool build_wide_bvh(FastBvh& fb, uint32_t max_stack);

// 64-byte triangle records for the kernels: v0 and the ray-independent part of triIntersect (ray-trace-compute.comp:157-173):
// a = v0 - v1, b = v2 - v0, n = cross(b, a), evaluated here with exactly the fp32 operations the shader performs (no
// contraction), so the values -- and every hit decision -- are bit-identical to computing them per ray.  Two 256-bit loads
// per test instead of three 128-bit ones, and 15 fewer instructions.
void precompute_triangles(FastBvh& fb);

// OpenMP threads used by the functions above (process-wide).
void set_repack_threads(int n);

}  // namespace vcrt
