#include "vcrt_repack.h"

#include <cmath>
#include <cstring>
#include <limits>

namespace vcrt {

namespace {
struct Todo { int32_t ref; int32_t parent; int side; uint32_t depth; };
inline float bits(int32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
inline float ubits(uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
}  // namespace

bool build_fast_bvh(const vcrt_bvh_node* bvh, uint32_t nbvh, const vcrt_triangle* tris, uint32_t ntris, FastBvh& out, std::string& err) {
    out.nodes.clear(); out.tris.clear(); out.root = (int32_t)0x80000000; out.depth = 0;
    if (nbvh == 0) return true;
    const float inf = std::numeric_limits<float>::infinity();
    std::vector<uint8_t> seen(nbvh, 0);
    std::vector<Todo> stack;
    out.nodes.reserve((size_t)nbvh / 2 * 16 + 16);
    out.tris.reserve((size_t)(nbvh / 2 + 1) * 12);
    stack.push_back({0, -1, 0, 0});
    while (!stack.empty()) {
        Todo td = stack.back();
        stack.pop_back();
        int32_t code = (int32_t)0x80000000;
        const vcrt_bvh_node* nd = nullptr;
        if (td.ref >= 0 && (uint32_t)td.ref < nbvh) {
            if (seen[td.ref]) { err = "bvh: node " + std::to_string(td.ref) + " is reachable twice (cycle or shared subtree)"; return false; }
            seen[td.ref] = 1;
            nd = &bvh[td.ref];
            const bool has_children = nd->leftNodeIndex != -1 || nd->rightNodeIndex != -1;
            if (nd->objectIndex != -1) {
                if (has_children) { err = "bvh: node " + std::to_string(td.ref) + " holds a triangle and children; only the reference traversal supports that"; return false; }
                const uint32_t slot = out.num_slots();
                vcrt_triangle t;
                std::memset(&t, 0, sizeof t);
                if ((uint32_t)nd->objectIndex < ntris) t = tris[nd->objectIndex];  // out-of-range -> zero triangle (never hit), as robust reads give
                const float rec[12] = {t.v0[0], t.v0[1], t.v0[2], ubits((uint32_t)nd->objectIndex), t.v1[0], t.v1[1], t.v1[2], ubits(t.materialIndex),
                                       t.v2[0], t.v2[1], t.v2[2], 0.0f};
                out.tris.insert(out.tris.end(), rec, rec + 12);
                code = ~(int32_t)slot;
                if (td.depth > out.depth) out.depth = td.depth;
            } else if (has_children) {
                const uint32_t idx = out.num_nodes();
                float rec[16];
                // both children empty until they report back
                rec[0] = rec[2] = rec[4] = rec[6] = rec[8] = rec[10] = inf;
                rec[1] = rec[3] = rec[5] = rec[7] = rec[9] = rec[11] = -inf;
                rec[12] = rec[13] = bits((int32_t)0x80000000);
                rec[14] = rec[15] = 0.0f;
                out.nodes.insert(out.nodes.end(), rec, rec + 16);
                code = (int32_t)idx;
                // the reference pushes left then right and pops right first
                stack.push_back({nd->leftNodeIndex, (int32_t)idx, 0, td.depth + 1});
                stack.push_back({nd->rightNodeIndex, (int32_t)idx, 1, td.depth + 1});
            }
            // a node with neither a triangle nor children contributes nothing
        }
        if (td.parent < 0) {
            out.root = code;
        } else if (code != (int32_t)0x80000000) {
            float* p = &out.nodes[(size_t)td.parent * 16];
            if (td.side == 0) {
                p[0] = nd->min[0]; p[1] = nd->max[0]; p[2] = nd->min[1]; p[3] = nd->max[1]; p[8] = nd->min[2]; p[9] = nd->max[2];
                p[12] = bits(code);
            } else {
                p[4] = nd->min[0]; p[5] = nd->max[0]; p[6] = nd->min[1]; p[7] = nd->max[1]; p[10] = nd->min[2]; p[11] = nd->max[2];
                p[13] = bits(code);
            }
        }
    }
    if (out.depth + 2 > 48) { err = "bvh: depth " + std::to_string(out.depth) + " exceeds the fast traversal stack (46)"; return false; }
    return true;
}

}  // namespace vcrt
