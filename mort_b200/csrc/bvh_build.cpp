// bvh_build.cpp — host binned-SAH builder producing the 4-wide BVH of device_types.h.
//
// Replaces the reference's median-split / bubble-sort builder (objects.cuh:528-661, <= 1024 nodes) — the
// reference only ever builds it for scenes 1 and 10 and scans everything else linearly (world.cuh:122-168).
// Pipeline: 16-bin SAH on centroid bounds per axis -> binary tree with type-homogeneous leaves of at most
// MORT_MAX_LEAF primitives -> collapse to 4-wide by repeatedly opening the child with the largest surface
// area -> breadth-first layout (root = node 0, top levels contiguous so a prefix can be staged in shared
// memory).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <queue>
#include <thread>

#include "flatten.hpp"

namespace mort {
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; a++) { lo[a] = std::numeric_limits<float>::infinity(); hi[a] = -lo[a]; } }
    void grow(const float* l, const float* h) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], l[a]); hi[a] = std::max(hi[a], h[a]); } }
    void grow(const Box& b) { grow(b.lo, b.hi); }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0;
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }
};

struct Node2 { Box box; int left = -1, right = -1, first = 0, count = 0, type = 0; };

constexpr int kBins = 16;
static float kTrav = 1.0f;                 // cost of one more (binary) node relative to one sphere test; MORT_KTRAV overrides (experiments)
inline float prim_cost(int type) { return type == MORT_OBJ_QUAD ? 1.3f : 1.0f; }

// Subtrees over disjoint index ranges are independent, so the top `par_levels` levels hand their left child to another
// thread and splice the two node arrays afterwards.  The tree (and therefore the 4-wide layout, which is rebuilt
// breadth-first from the child links) does not depend on the thread count.
constexpr int kParallelMinPrims = 1 << 14;

struct Builder {
    const std::vector<BuildPrim>& prims;
    std::vector<int> idx;
    std::vector<Node2> nodes;
    int max_leaf = MORT_MAX_LEAF;
    explicit Builder(const std::vector<BuildPrim>& p) : prims(p) {
        if (const char* e = getenv("MORT_MAX_LEAF")) { int v = atoi(e); if (v >= 1 && v <= 8) max_leaf = v; }   // experiments only
        if (const char* e = getenv("MORT_KTRAV")) { float v = (float)atof(e); if (v > 0) kTrav = v; }
        idx.resize(p.size());
        for (size_t i = 0; i < p.size(); i++) idx[i] = (int)i;
        nodes.reserve(p.size() * 2 + 1);
    }
    float centroid(int i, int a) const { return 0.5f * (prims[i].lo[a] + prims[i].hi[a]); }

    static void splice(std::vector<Node2>& dst, const std::vector<Node2>& src) {
        const int off = (int)dst.size();
        for (Node2 n : src) { if (n.left >= 0) n.left += off; if (n.right >= 0) n.right += off; dst.push_back(n); }
    }

    // builds the subtree over idx[b, e) into `nodes` (appending); returns its root's index in `nodes`
    int build(std::vector<Node2>& nodes, int b, int e, int par_levels) {
        int me = (int)nodes.size();
        nodes.emplace_back();
        Box box, cbox; box.reset(); cbox.reset();
        bool homogeneous = true; float cost_sum = 0;
        for (int i = b; i < e; i++) {
            const BuildPrim& p = prims[idx[i]];
            box.grow(p.lo, p.hi);
            float c[3] = {centroid(idx[i], 0), centroid(idx[i], 1), centroid(idx[i], 2)};
            cbox.grow(c, c);
            homogeneous = homogeneous && p.type == prims[idx[b]].type;
            cost_sum += prim_cost(p.type);
        }
        nodes[me].box = box;
        int n = e - b;
        bool can_leaf = n <= max_leaf && homogeneous;

        // best binned SAH split
        float best_cost = std::numeric_limits<float>::infinity(); int best_axis = -1, best_bin = -1;
        float parent_area = std::max(box.area(), 1e-30f);
        if (n >= 2) {
            for (int a = 0; a < 3; a++) {
                float ext = cbox.hi[a] - cbox.lo[a];
                if (!(ext > 0)) continue;
                Box bb[kBins]; int cnt[kBins]; float cst[kBins];
                for (int k = 0; k < kBins; k++) { bb[k].reset(); cnt[k] = 0; cst[k] = 0; }
                float scale = kBins / ext;
                for (int i = b; i < e; i++) {
                    int k = std::min(kBins - 1, std::max(0, (int)((centroid(idx[i], a) - cbox.lo[a]) * scale)));
                    bb[k].grow(prims[idx[i]].lo, prims[idx[i]].hi); cnt[k]++; cst[k] += prim_cost(prims[idx[i]].type);
                }
                float right_area[kBins], right_cost[kBins]; Box r; r.reset(); float rc = 0;
                for (int k = kBins - 1; k > 0; k--) { r.grow(bb[k]); rc += cst[k]; right_area[k] = r.area(); right_cost[k] = rc; }
                Box l; l.reset(); float lc = 0; int ln = 0;
                for (int k = 0; k < kBins - 1; k++) {
                    l.grow(bb[k]); lc += cst[k]; ln += cnt[k];
                    if (ln == 0 || ln == n) continue;
                    float c = kTrav + (l.area() * lc + right_area[k + 1] * right_cost[k + 1]) / parent_area;
                    if (c < best_cost) { best_cost = c; best_axis = a; best_bin = k; }
                }
            }
        }
        if (can_leaf && (best_axis < 0 || cost_sum <= best_cost)) {
            nodes[me].first = b; nodes[me].count = n; nodes[me].type = prims[idx[b]].type;
            return me;
        }
        int mid;
        if (best_axis >= 0) {
            float ext = cbox.hi[best_axis] - cbox.lo[best_axis], scale = kBins / ext, lo = cbox.lo[best_axis];
            int a = best_axis, bin = best_bin;
            auto it = std::partition(idx.begin() + b, idx.begin() + e, [&](int i) {
                int k = std::min(kBins - 1, std::max(0, (int)((centroid(i, a) - lo) * scale)));
                return k <= bin;
            });
            mid = (int)(it - idx.begin());
        } else if (!homogeneous) {
            int t0 = prims[idx[b]].type;
            auto it = std::partition(idx.begin() + b, idx.begin() + e, [&](int i) { return prims[i].type == t0; });
            mid = (int)(it - idx.begin());
        } else {
            mid = b + n / 2;      // coincident centroids: any balanced split
        }
        if (mid <= b || mid >= e) mid = b + n / 2;
        if (par_levels > 0 && n >= kParallelMinPrims) {
            std::vector<Node2> L, R;
            auto left = std::async(std::launch::async, [&] { return build(L, b, mid, par_levels - 1); });
            const int r = build(R, mid, e, par_levels - 1);
            const int l = left.get();
            const int off_l = (int)nodes.size(); splice(nodes, L);
            const int off_r = (int)nodes.size(); splice(nodes, R);
            nodes[me].left = off_l + l; nodes[me].right = off_r + r;
            return me;
        }
        int l = build(nodes, b, mid, 0);
        int r = build(nodes, mid, e, 0);
        nodes[me].left = l; nodes[me].right = r;
        return me;
    }
};

}  // namespace

void build_bvh4(const std::vector<BuildPrim>& prims, std::vector<Bvh4Node>& out, std::vector<int>& order_out, BuildStats& stats) {
    auto t0 = std::chrono::steady_clock::now();
    out.clear(); order_out.clear();
    const float inf = std::numeric_limits<float>::infinity();
    auto clear_node = [&](Bvh4Node& n) {
        for (int k = 0; k < 4; k++) {
            n.lox[k] = n.loy[k] = n.loz[k] = inf; n.hix[k] = n.hiy[k] = n.hiz[k] = -inf;
            n.child[k] = MORT_CHILD_EMPTY; n.spare[k] = 0;
        }
    };
    if (prims.empty()) { Bvh4Node n; clear_node(n); out.push_back(n); stats.n_nodes = 1; return; }

    Builder B(prims);
    // 2^levels concurrent subtrees at most; MORT_BUILD_THREADS=1 forces the serial build
    unsigned hw = std::thread::hardware_concurrency(); if (hw == 0) hw = 1;
    if (const char* e = getenv("MORT_BUILD_THREADS")) { int v = atoi(e); if (v >= 1) hw = (unsigned)v; }
    int levels = 0; while ((1u << levels) < hw && levels < 6) levels++;
    int root = B.build(B.nodes, 0, (int)prims.size(), levels);
    order_out = B.idx;
    stats.n_bvh2_nodes = (int)B.nodes.size();

    struct Item { int n2, n4, depth; };
    std::queue<Item> q;
    out.emplace_back(); clear_node(out[0]);
    q.push(Item{root, 0, 1});
    double sah = 0; float root_area = std::max(B.nodes[root].box.area(), 1e-30f);
    int max_depth4 = 1, leaf_slots = 0;
    while (!q.empty()) {
        Item it = q.front(); q.pop();
        max_depth4 = std::max(max_depth4, it.depth);
        int ch[4]; int nc = 0;
        const Node2& n2 = B.nodes[it.n2];
        if (n2.count > 0) ch[nc++] = it.n2;                 // the whole scene is one leaf
        else { ch[nc++] = n2.left; ch[nc++] = n2.right; }
        while (nc < 4) {
            int pick = -1; float best = -1;
            for (int k = 0; k < nc; k++) if (B.nodes[ch[k]].count == 0 && B.nodes[ch[k]].box.area() > best) { best = B.nodes[ch[k]].box.area(); pick = k; }
            if (pick < 0) break;
            int c = ch[pick];
            ch[pick] = B.nodes[c].left; ch[nc++] = B.nodes[c].right;
        }
        sah += kTrav * B.nodes[it.n2].box.area() / root_area;
        for (int k = 0; k < nc; k++) {
            const Node2& c = B.nodes[ch[k]];
            Bvh4Node& dst = out[it.n4];
            dst.lox[k] = c.box.lo[0]; dst.loy[k] = c.box.lo[1]; dst.loz[k] = c.box.lo[2];
            dst.hix[k] = c.box.hi[0]; dst.hiy[k] = c.box.hi[1]; dst.hiz[k] = c.box.hi[2];
            if (c.count > 0) {
                uint32_t w = MORT_LEAF_BIT | (c.type == MORT_OBJ_QUAD ? MORT_LEAF_QUAD_BIT : 0u) | ((uint32_t)(c.count - 1) << 27) | (uint32_t)c.first;
                out[it.n4].child[k] = w;
                leaf_slots++;
                sah += c.count * prim_cost(c.type) * c.box.area() / root_area;
            } else {
                int n4 = (int)out.size();
                out.emplace_back(); clear_node(out.back());
                out[it.n4].child[k] = (uint32_t)n4;
                q.push(Item{ch[k], n4, it.depth + 1});
            }
        }
    }
    stats.n_nodes = (int)out.size(); stats.max_depth = max_depth4; stats.n_leaf_slots = leaf_slots; stats.sah_cost = sah;
    stats.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace mort
