#include "vcrt_repack.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <omp.h>

namespace vcrt {

namespace {
struct Todo { int32_t ref; int32_t parent; int side; uint32_t depth; };
inline float bits(int32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
inline float ubits(uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
}  // namespace

bool build_fast_bvh(const vcrt_bvh_node* bvh, uint32_t nbvh, const vcrt_triangle* tris, uint32_t ntris, FastBvh& out, std::string& err) {
    out.nodes.clear(); out.tris.clear(); out.root = (int32_t)0x80000000; out.depth = 0; out.bound_depth = 0;
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
    out.bound_depth = out.depth;
    return true;
}

bool check_fast_depth(const FastBvh& fb, std::string& err) {
    if (fb.depth + 3 > 48) { err = "bvh: depth " + std::to_string(fb.depth) + " exceeds the fast traversal stack (45)"; return false; }
    return true;
}

// ------------------------------------------------------------------------------------------ SAH rebuild
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { lo[0] = lo[1] = lo[2] = std::numeric_limits<float>::infinity(); hi[0] = hi[1] = hi[2] = -std::numeric_limits<float>::infinity(); }
    void grow(const Box& b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    void grow(const float* p) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    float half_area() const { const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2]; return dx * dy + dy * dz + dz * dx; }
};

struct Prim { Box box; float c[3]; int32_t slot; };

struct SahBuilder {
    std::vector<Prim> prims;
    std::vector<float> nodes;     // 16 floats per inner node, preallocated (n - 1)
    std::atomic<uint32_t> max_depth{0};

    // Builds the subtree over prims[b,e) (e - b >= 2) into node index b (nodes b .. e-2 belong to this subtree: a binary
    // tree over m leaves has m-1 inner nodes, so the left subtree [b,mid) takes b+1 .. and the right one mid .. ; the
    // layout is therefore independent of the task schedule).  Returns the subtree bounds through `out`.
    void build(uint32_t b, uint32_t e, uint32_t node, uint32_t depth, Box& out) {
        const uint32_t n = e - b;
        Box bounds, cb;
        bounds.reset(); cb.reset();
        for (uint32_t i = b; i < e; ++i) { bounds.grow(prims[i].box); cb.grow(prims[i].c); }
        out = bounds;
        uint32_t mid = b + n / 2;
        bool split_found = false;
        if (n > 2 && depth < 24) {   // past depth 24 only median splits: depth <= 24 + log2(n) stays inside the traversal stack
            constexpr int NB = 16;
            float best_cost = std::numeric_limits<float>::infinity();
            int best_axis = -1, best_bin = -1;
            for (int a = 0; a < 3; ++a) {
                const float ext = cb.hi[a] - cb.lo[a];
                if (!(ext > 0.0f)) continue;
                const float k = NB * (1.0f - 1e-6f) / ext;
                Box bb[NB]; uint32_t cnt[NB];
                for (int i = 0; i < NB; ++i) { bb[i].reset(); cnt[i] = 0; }
                for (uint32_t i = b; i < e; ++i) {
                    int bin = (int)((prims[i].c[a] - cb.lo[a]) * k);
                    bin = bin < 0 ? 0 : (bin >= NB ? NB - 1 : bin);
                    bb[bin].grow(prims[i].box); cnt[bin]++;
                }
                float right_area[NB]; uint32_t right_cnt[NB];
                Box acc; acc.reset(); uint32_t c = 0;
                for (int i = NB - 1; i > 0; --i) { acc.grow(bb[i]); c += cnt[i]; right_area[i] = acc.half_area(); right_cnt[i] = c; }
                acc.reset(); c = 0;
                for (int i = 0; i < NB - 1; ++i) {
                    acc.grow(bb[i]); c += cnt[i];
                    if (c == 0 || right_cnt[i + 1] == 0) continue;
                    const float cost = acc.half_area() * (float)c + right_area[i + 1] * (float)right_cnt[i + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = i; }
                }
            }
            if (best_axis >= 0) {
                const int a = best_axis;
                const float k = NB * (1.0f - 1e-6f) / (cb.hi[a] - cb.lo[a]);
                auto it = std::partition(prims.begin() + b, prims.begin() + e, [&](const Prim& p) {
                    int bin = (int)((p.c[a] - cb.lo[a]) * k);
                    bin = bin < 0 ? 0 : (bin >= NB ? NB - 1 : bin);
                    return bin <= best_bin;
                });
                mid = (uint32_t)(it - prims.begin());
                split_found = mid > b && mid < e;
            }
        }
        if (!split_found) {   // degenerate centroids, tiny ranges or the depth cap: median on the widest centroid axis
            mid = b + n / 2;
            int a = 0;
            if (cb.hi[1] - cb.lo[1] > cb.hi[a] - cb.lo[a]) a = 1;
            if (cb.hi[2] - cb.lo[2] > cb.hi[a] - cb.lo[a]) a = 2;
            std::nth_element(prims.begin() + b, prims.begin() + mid, prims.begin() + e, [a](const Prim& x, const Prim& y) { return x.c[a] < y.c[a]; });
        }
        Box lb, rb;
        int32_t lcode, rcode;
        const uint32_t nl = mid - b, nr = e - mid;
        // node numbering: this node = `node`; left subtree's inner nodes follow immediately, then the right subtree's
        const uint32_t lnode = node + 1, rnode = node + 1 + (nl - 1);
        auto child = [&](uint32_t cb_, uint32_t ce_, uint32_t cnode, Box& bx, int32_t& code) {
            if (ce_ - cb_ == 1) { bx = prims[cb_].box; code = ~prims[cb_].slot; uint32_t d = depth + 1, m = max_depth.load(); while (d > m && !max_depth.compare_exchange_weak(m, d)) {} }
            else { build(cb_, ce_, cnode, depth + 1, bx); code = (int32_t)cnode; }
        };
        if (n > 8192) {
#pragma omp task shared(lb, lcode) firstprivate(b, mid, lnode)
            child(b, mid, lnode, lb, lcode);
#pragma omp task shared(rb, rcode) firstprivate(mid, e, rnode)
            child(mid, e, rnode, rb, rcode);
#pragma omp taskwait
        } else {
            child(b, mid, lnode, lb, lcode);
            child(mid, e, rnode, rb, rcode);
        }
        (void)nr;
        float* p = &nodes[(size_t)node * 16];
        p[0] = lb.lo[0]; p[1] = lb.hi[0]; p[2] = lb.lo[1]; p[3] = lb.hi[1]; p[8] = lb.lo[2]; p[9] = lb.hi[2];
        p[4] = rb.lo[0]; p[5] = rb.hi[0]; p[6] = rb.lo[1]; p[7] = rb.hi[1]; p[10] = rb.lo[2]; p[11] = rb.hi[2];
        p[12] = bits(lcode); p[13] = bits(rcode); p[14] = p[15] = 0.0f;
    }
};

}  // namespace

bool rebuild_fast_bvh_sah(FastBvh& fb, std::string& err) {
    const uint32_t n = fb.num_slots();
    if (n < 2) return true;   // empty scene or a single leaf: nothing to restructure
    SahBuilder sb;
    sb.prims.resize(n);
    const float eps = 0.0001f;   // the reference's leaf padding (Bvh.h:16)
    for (uint32_t i = 0; i < n; ++i) {
        const float* t = &fb.tris[(size_t)i * 12];
        Prim& p = sb.prims[i];
        for (int a = 0; a < 3; ++a) {
            const float lo = std::min(std::min(t[a], t[4 + a]), t[8 + a]), hi = std::max(std::max(t[a], t[4 + a]), t[8 + a]);
            p.box.lo[a] = lo - eps; p.box.hi[a] = hi + eps;
            p.c[a] = 0.5f * (lo + hi);
        }
        p.slot = (int32_t)i;
    }
    sb.nodes.assign((size_t)(n - 1) * 16, 0.0f);
    Box root;
#pragma omp parallel
#pragma omp single
    sb.build(0, n, 0, 0, root);
    if (sb.max_depth.load() + 3 > 48) { err = "sah rebuild: depth " + std::to_string(sb.max_depth.load()) + " exceeds the fast traversal stack"; return false; }
    fb.nodes.swap(sb.nodes);
    fb.root = 0;
    fb.depth = sb.max_depth.load();
    return true;
}

// ------------------------------------------------------------------------------------------ insertion-based optimisation
// After Bittner, Hapala, Havran: "Fast insertion-based optimization of bounding volume hierarchies" (2013).  The top-down build fixes
// the upper levels before it knows what lies below them; here inner nodes are taken out again, largest surface area first, and their
// two children are put back where they increase the tree's total surface area least (branch-and-bound search from the root).  Leaves
// keep their slots, so results cannot change; only the number of boxes a ray enters does.  Sequential (the tree is mutated in place).
namespace {
struct Reinserter {
    uint32_t nleaf = 0, ninner = 0;
    std::vector<Box> box;                 // node ids: inner 0 .. ninner-1, leaf ninner + slot
    std::vector<int32_t> parent;
    std::vector<int32_t> kid;             // 2 per inner node
    int32_t root = 0;
    static float area(const Box& b) { return b.half_area(); }
    static Box merge(const Box& a, const Box& b) { Box r = a; r.grow(b); return r; }
    bool inner(int32_t v) const { return (uint32_t)v < ninner; }
    void refit(int32_t v) {
        while (v >= 0) {
            const Box nb = merge(box[kid[2 * v]], box[kid[2 * v + 1]]);
            if (std::memcmp(&nb, &box[v], sizeof nb) == 0) break;
            box[v] = nb;
            v = parent[v];
        }
    }
    struct Cand { float ind; int32_t node; bool operator<(const Cand& o) const { return ind > o.ind; } };
    std::vector<Cand> heap;
    int32_t find_place(const Box& b) {
        const float ab = area(b);
        float best_cost = std::numeric_limits<float>::infinity();
        int32_t best = root;
        heap.clear();
        heap.push_back({0.0f, root});
        int budget = 4096;   // degenerate scenes (thousands of coincident boxes) would otherwise search the whole tree for every node
        while (!heap.empty() && budget-- > 0) {
            std::pop_heap(heap.begin(), heap.end());
            const Cand c = heap.back();
            heap.pop_back();
            if (c.ind + ab >= best_cost) break;
            const float direct = area(merge(box[c.node], b));
            const float total = c.ind + direct;
            if (total < best_cost) { best_cost = total; best = c.node; }
            const float ci = total - area(box[c.node]);
            if (inner(c.node) && ci + ab < best_cost) {
                heap.push_back({ci, kid[2 * c.node]}); std::push_heap(heap.begin(), heap.end());
                heap.push_back({ci, kid[2 * c.node + 1]}); std::push_heap(heap.begin(), heap.end());
            }
        }
        return best;
    }
    void insert(int32_t v, int32_t at, int32_t fresh) {   // `fresh` (a free inner node) becomes the parent of `at` and `v`
        const int32_t g = parent[at];
        kid[2 * fresh] = at; kid[2 * fresh + 1] = v;
        parent[at] = fresh; parent[v] = fresh; parent[fresh] = g;
        box[fresh] = merge(box[at], box[v]);
        if (g < 0) root = fresh;
        else { kid[2 * g + (kid[2 * g] == at ? 0 : 1)] = fresh; refit(g); }
    }
    bool reinsert(int32_t n) {
        const int32_t p = parent[n];
        if (!inner(n) || p < 0 || parent[p] < 0) return false;   // the root and its children stay
        const int32_t g = parent[p], s = kid[2 * p] == n ? kid[2 * p + 1] : kid[2 * p];
        int32_t l = kid[2 * n], r = kid[2 * n + 1];
        kid[2 * g + (kid[2 * g] == p ? 0 : 1)] = s;
        parent[s] = g;
        refit(g);
        if (area(box[l]) < area(box[r])) std::swap(l, r);
        insert(l, find_place(box[l]), n);
        insert(r, find_place(box[r]), p);
        return true;
    }
};
}  // namespace

bool optimize_fast_bvh_reinsert(FastBvh& fb, int passes, float fraction, std::string& err) {
    const uint32_t n = fb.num_slots();
    if (n < 8 || fb.root != 0 || fb.num_nodes() != n - 1) return true;   // nothing worth doing / not a tree built by rebuild_fast_bvh_sah
    static const int kMin[2][3] = {{0, 2, 8}, {4, 6, 10}};
    Reinserter t;
    t.nleaf = n; t.ninner = n - 1;
    t.box.resize(2 * (size_t)n - 1); t.parent.assign(2 * (size_t)n - 1, -1); t.kid.resize(2 * (size_t)(n - 1));
    for (uint32_t i = 0; i < n - 1; ++i) {
        const float* p = &fb.nodes[(size_t)i * 16];
        for (int c = 0; c < 2; ++c) {
            int32_t code; std::memcpy(&code, &p[12 + c], 4);
            if (code == (int32_t)0x80000000) return true;              // not a full binary tree: leave it
            const int32_t id = code >= 0 ? code : (int32_t)(t.ninner + (uint32_t)(~code));
            if ((uint32_t)id >= t.box.size() || t.parent[id] >= 0) { err = "reinsertion: malformed tree"; return false; }
            t.kid[2 * i + c] = id; t.parent[id] = (int32_t)i;
            for (int a = 0; a < 3; ++a) { t.box[id].lo[a] = p[kMin[c][a]]; t.box[id].hi[a] = p[kMin[c][a] + 1]; }
        }
    }
    t.parent[0] = -1;
    t.box[0] = Reinserter::merge(t.box[t.kid[0]], t.box[t.kid[1]]);
    t.root = 0;
    std::vector<int32_t> order(t.ninner);
    const uint32_t per_pass = std::max(1u, (uint32_t)((double)t.ninner * fraction));
    for (int pass = 0; pass < passes; ++pass) {
        for (uint32_t i = 0; i < t.ninner; ++i) order[i] = (int32_t)i;
        std::partial_sort(order.begin(), order.begin() + std::min<size_t>(per_pass, order.size()), order.end(),
                          [&](int32_t a, int32_t b) { return Reinserter::area(t.box[a]) > Reinserter::area(t.box[b]); });
        for (uint32_t i = 0; i < per_pass && i < t.ninner; ++i) t.reinsert(order[i]);
    }
    // depth of the new tree; a tree the traversal stack could not hold is not taken
    std::vector<float> nodes((size_t)(n - 1) * 16, 0.0f);
    std::vector<int32_t> newid(t.ninner, -1);
    struct Item { int32_t v; uint32_t depth; };
    std::vector<Item> st;
    uint32_t next = 0, max_depth = 0;
    st.push_back({t.root, 0u});
    newid[t.root] = (int32_t)next++;
    while (!st.empty()) {
        const Item it = st.back();
        st.pop_back();
        float* p = &nodes[(size_t)newid[it.v] * 16];
        for (int c = 0; c < 2; ++c) {
            const int32_t k = t.kid[2 * it.v + c];
            int32_t code;
            if (t.inner(k)) { newid[k] = (int32_t)next++; code = newid[k]; st.push_back({k, it.depth + 1}); }
            else { code = ~(int32_t)((uint32_t)k - t.ninner); if (it.depth + 1 > max_depth) max_depth = it.depth + 1; }
            for (int a = 0; a < 3; ++a) { p[kMin[c][a]] = t.box[k].lo[a]; p[kMin[c][a] + 1] = t.box[k].hi[a]; }
            std::memcpy(&p[12 + c], &code, 4);
        }
    }
    if (next != t.ninner) { err = "reinsertion: lost nodes"; return false; }
    if (max_depth + 3 > 48) return true;   // too deep for the binary traversal stack: keep the tree as built
    fb.nodes.swap(nodes);
    fb.root = 0;
    fb.depth = max_depth;
    return true;
}

void precompute_triangles(FastBvh& fb) {
    const uint32_t n = fb.num_slots();
    fb.tris64.assign((size_t)n * 16, 0.0f);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const float* t = &fb.tris[(size_t)i * 12];
        float* o = &fb.tris64[(size_t)i * 16];
        // volatile: each operation is rounded to fp32 on its own, whatever the host compiler's contraction settings
        volatile float a[3], b[3], m[6];
        for (int k = 0; k < 3; ++k) { a[k] = t[k] - t[4 + k]; b[k] = t[8 + k] - t[k]; }
        // cross(b, a) = (b.y*a.z - a.y*b.z, b.z*a.x - a.z*b.x, b.x*a.y - a.x*b.y)   (glm::cross operation order)
        m[0] = b[1] * a[2]; m[1] = a[1] * b[2]; m[2] = b[2] * a[0]; m[3] = a[2] * b[0]; m[4] = b[0] * a[1]; m[5] = a[0] * b[1];
        o[0] = t[0]; o[1] = t[1]; o[2] = t[2]; o[3] = t[3];
        o[4] = a[0]; o[5] = a[1]; o[6] = a[2]; o[7] = t[7];
        o[8] = b[0]; o[9] = b[1]; o[10] = b[2];
        o[12] = m[0] - m[1]; o[13] = m[2] - m[3]; o[14] = m[4] - m[5];
    }
}

void set_repack_threads(int n) { omp_set_num_threads(n < 1 ? 1 : n); }

// ------------------------------------------------------------------------------------------ quantisation
namespace {
// One axis of one box in the frame (qorg, qext): {lo | hi << 16}, rounded outwards; the empty box (min > max) for mn > mx.
inline uint32_t quant_bounds(const FastBvh& fb, int a, double mn, double mx) {
    uint32_t qlo = 32767u, qhi = 0u;         // empty child: min > max on every axis, never entered
    if (mn <= mx) {
        // decoded value = qorg + (1 + q/32768) * qext, evaluated with the floats the kernel uses
        const double org = fb.qorg[a], ext = fb.qext[a];
        double l = std::floor((mn - org - ext) / ext * 32768.0), h = std::ceil((mx - org - ext) / ext * 32768.0);
        while (l > 0.0 && org + (1.0 + l / 32768.0) * ext > mn) l -= 1.0;
        while (h < 32767.0 && org + (1.0 + h / 32768.0) * ext < mx) h += 1.0;
        qlo = (uint32_t)std::min(std::max(l, 0.0), 32767.0);
        qhi = (uint32_t)std::min(std::max(h, 0.0), 32767.0);
    }
    return qlo | (qhi << 16);
}
}  // namespace
bool quantize_fast_bvh(FastBvh& fb, float max_quantum) {
    fb.qnodes.clear();
    const uint32_t n = fb.num_nodes();
    if (n == 0) return false;
    const double inf = std::numeric_limits<double>::infinity();
    double lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    // word index of {min, max} per child and axis in the 64-byte record (layout: vcrt_fast.cuh)
    static const int kMin[2][3] = {{0, 2, 8}, {4, 6, 10}};
    for (uint32_t i = 0; i < n; ++i) {
        const float* p = &fb.nodes[(size_t)i * 16];
        for (int c = 0; c < 2; ++c)
            for (int a = 0; a < 3; ++a) {
                const double mn = p[kMin[c][a]], mx = p[kMin[c][a] + 1];
                if (!(mn <= mx)) continue;               // empty child (+inf, -inf)
                if (!std::isfinite(mn) || !std::isfinite(mx)) return false;
                lo[a] = std::min(lo[a], mn); hi[a] = std::max(hi[a], mx);
            }
    }
    double quantum[3], base[3];
    for (int a = 0; a < 3; ++a) {
        if (!(lo[a] <= hi[a])) return false;             // no finite box at all
        double ext = hi[a] - lo[a];
        if (!(ext > 0.0)) ext = 1e-3;                    // flat scene on this axis
        quantum[a] = ext / 32764.0;                      // q stays inside [1, 32766]
        base[a] = lo[a] - quantum[a];
        if (quantum[a] > (double)max_quantum) return false;
    }
    for (int a = 0; a < 3; ++a) {
        const double E = 32768.0 * quantum[a];
        fb.qorg[a] = (float)(base[a] - E);
        fb.qext[a] = (float)E;
        // the kernel decodes with these rounded floats; make sure rounding did not move the frame by a visible amount
        if (std::fabs((double)fb.qorg[a] - (base[a] - E)) > 0.01 * quantum[a] || std::fabs((double)fb.qext[a] - E) > 1e-6 * E) return false;
    }
    fb.qnodes.resize((size_t)n * 8);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const float* p = &fb.nodes[(size_t)i * 16];
        uint32_t* q = &fb.qnodes[(size_t)i * 8];
        for (int c = 0; c < 2; ++c)
            for (int a = 0; a < 3; ++a) {
                const double mn = p[kMin[c][a]], mx = p[kMin[c][a] + 1];
                const uint32_t word = quant_bounds(fb, a, mn, mx);
                q[c * 3 + a] = word;
            }
        std::memcpy(&q[6], &p[12], 4);
        std::memcpy(&q[7], &p[13], 4);
    }
    return true;
}


// ------------------------------------------------------------------------------------------ 4-wide tree
bool build_wide_bvh(FastBvh& fb, uint32_t max_stack) {
    fb.q4nodes.clear();
    fb.root4 = fb.root;
    fb.stack4 = 0;
    if (fb.qnodes.empty() || fb.root < 0) return false;   // not quantised, or the tree is a single leaf / empty: the binary form serves
    static const int kMin[2][3] = {{0, 2, 8}, {4, 6, 10}};
    const int32_t EMPTY = (int32_t)0x80000000;
    struct Child { float lo[3], hi[3]; int32_t code; };
    auto children_of = [&](int32_t node, Child* out) {   // the (up to two) non-empty children of binary node `node`
        const float* p = &fb.nodes[(size_t)node * 16];
        int n = 0;
        for (int c = 0; c < 2; ++c) {
            int32_t code; std::memcpy(&code, &p[12 + c], 4);
            if (code == EMPTY) continue;
            Child& ch = out[n++];
            for (int a = 0; a < 3; ++a) { ch.lo[a] = p[kMin[c][a]]; ch.hi[a] = p[kMin[c][a] + 1]; }
            ch.code = code;
        }
        return n;
    };
    auto area = [](const Child& c) { const float dx = c.hi[0] - c.lo[0], dy = c.hi[1] - c.lo[1], dz = c.hi[2] - c.lo[2]; return dx * dy + dy * dz + dz * dx; };
    struct Task { int32_t node2; uint32_t node4; uint32_t stack_above; };   // stack_above: entries already on the stack when this node is visited
    std::vector<Task> todo;
    std::vector<uint32_t>& out = fb.q4nodes;
    out.reserve((size_t)fb.num_nodes() / 2 * 16 + 16);
    out.resize(16);
    todo.push_back({fb.root, 0u, 0u});
    uint32_t need = 0;
    while (!todo.empty()) {
        const Task tk = todo.back();
        todo.pop_back();
        Child ch[4];
        int n = children_of(tk.node2, ch);
        while (n < 4) {   // open the inner child with the largest surface area
            int best = -1; float ba = -1.0f;
            for (int i = 0; i < n; ++i) if (ch[i].code >= 0 && area(ch[i]) > ba) { ba = area(ch[i]); best = i; }
            if (best < 0) break;
            Child sub[2];
            const int m = children_of(ch[best].code, sub);
            if (m == 0) { ch[best] = ch[--n]; continue; }      // cannot happen for trees built here (inner nodes have children)
            ch[best] = sub[0];
            if (m == 2) ch[n++] = sub[1];
        }
        // a visit leaves at most n - 1 entries behind and descends into one child
        const uint32_t below = tk.stack_above + (n > 0 ? (uint32_t)(n - 1) : 0u);
        if (below > need) need = below;
        uint32_t w[16];
        for (int i = 0; i < 4; ++i) {
            uint32_t* half = w + (i / 2) * 8;
            const int k = i & 1;
            int32_t code = EMPTY;
            for (int a = 0; a < 3; ++a) half[k * 3 + a] = 32767u;   // the empty box
            if (i < n) {
                for (int a = 0; a < 3; ++a) half[k * 3 + a] = quant_bounds(fb, a, ch[i].lo[a], ch[i].hi[a]);
                code = ch[i].code;
                if (code >= 0) {
                    const uint32_t idx = (uint32_t)(out.size() / 16);
                    out.resize(out.size() + 16);
                    todo.push_back({code, idx, below});
                    code = (int32_t)idx;
                }
            }
            std::memcpy(&half[6 + k], &code, 4);
        }
        std::memcpy(&out[(size_t)tk.node4 * 16], w, sizeof w);
    }
    fb.stack4 = need + 2;   // + the sentinel slot and the register-held top's spill slot
    if (fb.stack4 > max_stack) { fb.q4nodes.clear(); return false; }
    fb.root4 = 0;
    return true;
}

}  // namespace vcrt
