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
#include <string>
#include <limits>
#include <memory>
#include <queue>
#include <thread>

#include <atomic>

#include "bvh_sah.hpp"
#include "flatten.hpp"

namespace mort {
namespace {

inline float node_area(const Node2& n) { return sah_area(n.lo, n.hi); }

// Subtrees over disjoint index ranges are independent, so the top `par_levels` levels hand their left child to another
// thread.  Nodes come out of one preallocated array through an atomic cursor: their numbering depends on the thread timing,
// the tree does not (bvh_sah.hpp), and the 4-wide layout is rebuilt breadth-first from the child links.
constexpr int kParallelMinPrims = 1 << 14;

struct AtomicAlloc {
    std::atomic<int>* next;
    int pair() { return next->fetch_add(2); }
};

struct Builder {
    const std::vector<BuildPrim>& prims;
    std::vector<int> idx;
    std::unique_ptr<Node2[]> nodes;      // 2n + 1, uninitialised: every node is written before it is read
    std::atomic<int> next{1};
    SahParams P;
    Builder(const std::vector<BuildPrim>& p, const SahParams& par) : prims(p), P(par) {
        idx.resize(p.size());
        for (size_t i = 0; i < p.size(); i++) idx[i] = (int)i;
        nodes.reset(new Node2[p.size() * 2 + 1]);
    }
    // one node of the top levels: the same statistics, sweep and decision as sah_build_subtree, with the two children built concurrently
    void build(int me, int b, int e, int par_levels) {
        if (par_levels <= 0 || e - b < kParallelMinPrims) {
            AtomicAlloc A{&next};
            sah_build_subtree(prims.data(), idx.data(), nodes.get(), me, b, e, P, A);
            return;
        }
        SahNodeStats S; sah_stats_clear(S);
        for (int i = b; i < e; i++) sah_stats_add(S, prims[idx[i]], idx[i]);
        Node2 N; sah_node_set_box(N, S);
        float best_cost = INFINITY; int best_axis = -1, best_bin = -1, best_left = 0;
        const float parent_area = std::max(sah_area(S.lo, S.hi), 1e-30f);
        for (int a = 0; a < 3; a++) {
            if (!sah_axis_open(S, a)) continue;
            SahBins B; sah_bins_clear(B);
            const float scale = sah_axis_scale(S, a), lo = S.clo[a];
            for (int i = b; i < e; i++) {
                const BuildPrim& p = prims[idx[i]];
                const int k = sah_bin_of(sah_centroid(p, a), lo, scale);
                sah_grow(B.lo[k], B.hi[k], p.lo, p.hi); B.cnt[k]++; B.cst[k] += sah_prim_cost(p.type);
            }
            sah_sweep_axis(a, B, S.n, parent_area, P.k_trav, best_cost, best_axis, best_bin, best_left);
        }
        int n_ref_left = 0;
        if (best_axis < 0 && sah_homogeneous(S)) { const int pv = sah_ref_pivot(S); for (int i = b; i < e; i++) n_ref_left += idx[i] < pv ? 1 : 0; }
        const SahSplit d = sah_decide(S, P, best_cost, best_axis, best_bin, best_left, n_ref_left);      // never a leaf: n >= kParallelMinPrims > max_leaf
        auto it = std::partition(idx.begin() + b, idx.begin() + e, [&](int i) { return sah_goes_left(d, prims[i], i); });
        const int mid = (int)(it - idx.begin());
        const int c = next.fetch_add(2);
        N.left = c; N.right = c + 1;
        nodes[me] = N;
        auto left = std::async(std::launch::async, [&] { build(c, b, mid, par_levels - 1); });
        build(c + 1, mid, e, par_levels - 1);
        left.get();
    }
};

}  // namespace

void build_bvh4(const std::vector<BuildPrim>& prims, std::vector<Bvh4Node>& out, std::vector<int>& order_out, BuildStats& stats, const BuildOptions& opt) {
    auto t0 = std::chrono::steady_clock::now();
    out.clear(); order_out.clear(); stats.level_first.clear();
    if (prims.empty()) { Bvh4Node n; bvh4_clear_node(n); out.push_back(n); stats.n_nodes = 1; stats.level_first = {0, 1}; return; }

    const SahParams P = {opt.max_leaf >= 1 && opt.max_leaf <= 8 ? opt.max_leaf : MORT_MAX_LEAF, opt.k_trav > 0 ? opt.k_trav : 1.0f};
    Builder B(prims, P);
    // 2^levels concurrent subtrees at most; threads = 1 forces the serial build
    unsigned hw = opt.threads > 0 ? (unsigned)opt.threads : std::thread::hardware_concurrency(); if (hw == 0) hw = 1;
    int levels = 0; while ((1u << levels) < hw && levels < 6) levels++;
    const int root = 0;
    B.build(root, 0, (int)prims.size(), levels);
    order_out = B.idx;
    stats.n_bvh2_nodes = B.next.load();
    const Node2* N2 = B.nodes.get();

    // collapse, breadth-first: a 4-wide node's index is its position in the queue order, so levels are contiguous
    struct Item { int n2, n4, depth; };
    std::queue<Item> q;
    out.emplace_back(); bvh4_clear_node(out[0]);
    q.push(Item{root, 0, 1});
    double sah = 0; const float root_area = std::max(node_area(N2[root]), 1e-30f);
    int max_depth4 = 1, leaf_slots = 0;
    stats.level_first.push_back(0);
    while (!q.empty()) {
        Item it = q.front(); q.pop();
        if (it.depth > max_depth4) { max_depth4 = it.depth; stats.level_first.push_back(it.n4); }
        int ch[4];
        const int nc = bvh4_open_children(N2, it.n2, ch);
        sah += (double)bvh4_node_sah(N2[it.n2], P.k_trav, root_area);
        for (int k = 0; k < nc; k++) {
            const Node2& c = N2[ch[k]];
            bvh4_set_child_box(out[it.n4], k, c);
            if (c.count > 0) {
                out[it.n4].child[k] = bvh4_leaf_word(c);
                leaf_slots++;
                sah += (double)bvh4_leaf_sah(c, root_area);
            } else {
                int n4 = (int)out.size();
                out.emplace_back(); bvh4_clear_node(out.back());
                out[it.n4].child[k] = (uint32_t)n4;
                q.push(Item{ch[k], n4, it.depth + 1});
            }
        }
    }
    stats.level_first.push_back((int)out.size());
    stats.n_nodes = (int)out.size(); stats.max_depth = max_depth4; stats.n_leaf_slots = leaf_slots; stats.sah_cost = sah;
    stats.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace mort
