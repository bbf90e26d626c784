// gpu_build_driver.hpp — the host loop of the GPU tree builder, written over an executor X that owns the memory and runs
// the bodies of gpu_build_core.cuh: gpu_build.cu passes CUDA kernels on a stream, tests/hostsim/buildsim.cpp passes serial
// loops (no GPU in the development container), so the level logic below is exercised by both.
#pragma once
#include <string>
#include <vector>

#include "gpu_build_core.cuh"

namespace mort {
namespace gb {

// X provides:  void* alloc(size_t) (256-byte aligned, called once; nullptr on failure) · upload(dst, src, bytes) · download(dst, src, bytes) (both ordered with the
// launches, download returns after the data is there) · zero(dst, bytes) · k_init(c, slot_value) · k_clear(c, n) · k_stats(c, cur) ·
// k_bin(c, cur) · k_split(c, cur, n) · k_partition(c, cur) · k_small(c, cur, n) · k_collapse_count(c, cur, n) · k_scan(c, n) ·
// k_collapse_emit(c, cur, base, n) · bool ok(std::string*)
template <class X>
bool build_run(X& x, const std::vector<BuildPrim>& prims, std::vector<Bvh4Node>& nodes_out, std::vector<int>& order_out, BuildStats& stats,
               const BuildOptions& opt, int k_small, std::string* err) {
    auto fail = [&](const std::string& m) { if (err) *err = m; return false; };
    const int N = (int)prims.size();
    nodes_out.clear(); order_out.clear(); stats.level_first.clear();
    if (N == 0) { Bvh4Node n; bvh4_clear_node(n); nodes_out.push_back(n); stats.n_nodes = 1; stats.level_first = {0, 1}; return true; }
    Ctx c; memset(&c, 0, sizeof(c));
    c.N = N;
    c.P.max_leaf = opt.max_leaf >= 1 && opt.max_leaf <= 8 ? opt.max_leaf : MORT_MAX_LEAF;
    c.P.k_trav = opt.k_trav > 0 ? opt.k_trav : 1.0f;
    c.k_small = k_small < c.P.max_leaf ? c.P.max_leaf : k_small;
    const size_t max_active = (size_t)N / (size_t)(c.k_small + 1) + 2;

    // one workspace, carved: first pass measures, second pass hands out
    char* ws = nullptr; size_t off = 0;
    BuildPrim* d_prims = nullptr;
    for (int pass = 0; pass < 2; pass++) {
        off = 0;
        auto take = [&](size_t bytes) { void* p = ws ? ws + off : nullptr; off += (bytes + 255) / 256 * 256; return p; };
        d_prims = (BuildPrim*)take((size_t)N * sizeof(BuildPrim));
        for (int k = 0; k < 2; k++) {
            c.idx[k] = (int*)take((size_t)N * 4); c.slot[k] = (int*)take((size_t)N * 4);
            c.active[k] = (Active*)take(max_active * sizeof(Active));
            c.lvl[k] = (int*)take((size_t)(N + 1) * 4);
        }
        c.nodes = (Node2*)take(((size_t)N * 2 + 1) * sizeof(Node2));
        c.stats = (GStats*)take(max_active * sizeof(GStats));
        c.bins = (uint32_t*)take(max_active * kBinWords * 4);
        c.split = (SplitRec*)take(max_active * sizeof(SplitRec));
        c.small = (Active*)take((size_t)(N + 1) * sizeof(Active));
        c.cnt = (Counters*)take(sizeof(Counters));
        c.kids = (int*)take((size_t)(N + 1) * 16); c.icount = (int*)take((size_t)(N + 1) * 4); c.ioff = (int*)take((size_t)(N + 1) * 4);
        c.out = (Bvh4Node*)take((size_t)(N + 1) * sizeof(Bvh4Node));
        if (pass == 0) { ws = (char*)x.alloc(off); if (!ws) return fail("out of device memory for the build workspace (" + std::to_string(off >> 20) + " MiB)"); }
    }
    c.prims = d_prims;
    stats.gpu_workspace_bytes = off;

    x.upload(d_prims, prims.data(), (size_t)N * sizeof(BuildPrim));
    Counters h; memset(&h, 0, sizeof(h)); h.n_nodes = 1;
    int n_active = 0;
    const Active root = {0, 0, N};
    if (N > c.k_small) { x.upload(c.active[0], &root, sizeof(root)); n_active = 1; }
    else { x.upload(c.small, &root, sizeof(root)); h.n_small = 1; }
    x.upload(c.cnt, &h, sizeof(h));
    x.k_init(c, n_active ? 0 : -1);

    // ---- large nodes, level by level ----
    int cur = 0, levels = 0;
    const int zero = 0;
    while (n_active > 0) {
        x.k_clear(c, n_active);
        x.k_stats(c, cur);
        x.k_bin(c, cur);
        x.k_split(c, cur, n_active);
        x.k_partition(c, cur);
        x.download(&h, c.cnt, sizeof(h));
        if ((size_t)h.n_next > max_active || h.n_nodes > 2 * N + 1 || h.n_small > N + 1) return fail("builder invariant broken (more nodes than primitives allow)");
        n_active = h.n_next;
        x.upload(&c.cnt->n_next, &zero, 4);
        cur ^= 1;
        if (++levels > 8192) return fail("tree deeper than 8192 levels");
    }
    // ---- small subtrees: one thread each ----
    if (h.n_small > 0) x.k_small(c, cur, h.n_small);

    // ---- collapse to 4-wide, breadth-first ----
    Node2 rootn; x.download(&rootn, c.nodes, sizeof(rootn));
    c.root_area = sah_max(sah_area(rootn.lo, rootn.hi), 1e-30f);
    x.upload(c.lvl[0], &zero, 4);
    int base = 0, n = 1, lc = 0, depth = 0;
    stats.level_first.push_back(0);
    while (n > 0) {
        x.k_collapse_count(c, lc, n);
        x.k_scan(c, n);
        x.k_collapse_emit(c, lc, base, n);
        x.download(&h, c.cnt, sizeof(h));
        base += n; n = h.collapse_total; lc ^= 1; depth++;
        stats.level_first.push_back(base);
        if (base + n > N + 1) return fail("builder invariant broken (more 4-wide nodes than primitives)");
    }
    nodes_out.resize((size_t)base); order_out.resize((size_t)N);
    x.download(nodes_out.data(), c.out, (size_t)base * sizeof(Bvh4Node));
    x.download(order_out.data(), c.idx[cur], (size_t)N * 4);
    stats.n_nodes = base; stats.n_bvh2_nodes = h.n_nodes; stats.max_depth = depth; stats.n_leaf_slots = h.leaf_slots;
    stats.sah_cost = (double)h.sah_fx / 4294967296.0;
    stats.gpu_levels = levels; stats.gpu_small_subtrees = h.n_small;
    return x.ok(err);
}

}  // namespace gb
}  // namespace mort
