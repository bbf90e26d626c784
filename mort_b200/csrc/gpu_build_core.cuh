// gpu_build_core.cuh — per-thread bodies of the GPU tree builder (gpu_build.cu wraps them in kernels).
//
// The builder makes the SAME tree as the host builder (bvh_build.cpp): both use the split rule of bvh_sah.hpp, which is a
// function of the set of primitives below a node.  Work decomposition:
//   large nodes (more than `k_small` primitives), level by level, one thread per primitive position:
//       clear -> stats (box, centroid box, cost, type count, index range: min / max / add atomics on order-preserving integer
//       keys) -> bin (16 bins x 3 axes: box keys, count, cost) -> split (one thread per node: the host's sweep over the decoded
//       bins, children allocated, small children handed to the subtree list) -> partition (every position takes a slot of its
//       child's range from an atomic cursor; the order inside a child is arbitrary, which the rule does not see)
//   small subtrees: one thread each runs the sequential builder of bvh_sah.hpp in place (the very code the host runs)
//   collapse to 4-wide, level by level in breadth-first order: count internal children -> exclusive scan -> emit
// The bodies also compile as plain C++: tests/hostsim/buildsim.cpp runs them serially (positions in a shuffled order, to
// stand in for the arbitrary partition order) against the host builder on a machine without a GPU.
#pragma once
#include "bvh_sah.hpp"

namespace mort {
namespace gb {

struct Active { int node, b, e; };
struct GStats { uint32_t lo[3], hi[3], clo[3], chi[3]; int cost, n_spheres, min_ref, max_ref, n_ref_left, pad[3]; };   // 20 words
constexpr int kBinWords = 3 * kSahBins * 8;          // per slot: [axis][bin]{lo3, hi3, cnt, cst}
struct SplitRec { SahSplit d; int b, slot_l, slot_r, cur_l, cur_r; };
struct Counters { int n_nodes, n_next, n_small, collapse_total, leaf_slots, pad[3]; unsigned long long sah_fx; };

struct Ctx {
    int N, k_small;
    SahParams P;
    const BuildPrim* prims;
    int* idx[2]; int* slot[2];                       // ping-pong: primitive at a position, the active-list slot of its node (-1: settled)
    Node2* nodes;
    Active* active[2];
    GStats* stats; uint32_t* bins; SplitRec* split;
    Active* small;
    Counters* cnt;
    // collapse
    int* lvl[2]; int* kids; int* icount; int* ioff;
    Bvh4Node* out;
    float root_area;
};

constexpr uint32_t kKeyPosInf = 0xFF800000u, kKeyNegInf = 0x007FFFFFu;
SAH_HD uint32_t f_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
SAH_HD float bits_f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
// order-preserving map float -> uint32 (so min / max of floats are integer atomics)
SAH_HD uint32_t key_of(float f) { const uint32_t u = f_bits(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
SAH_HD float float_of(uint32_t k) { return bits_f((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k); }

#if defined(__CUDA_ARCH__)
SAH_HD void a_min(uint32_t* p, uint32_t v) { atomicMin(p, v); }
SAH_HD void a_max(uint32_t* p, uint32_t v) { atomicMax(p, v); }
SAH_HD void a_imin(int* p, int v) { atomicMin(p, v); }
SAH_HD void a_imax(int* p, int v) { atomicMax(p, v); }
SAH_HD int a_add(int* p, int v) { return atomicAdd(p, v); }
SAH_HD void a_add64(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }
#else
SAH_HD void a_min(uint32_t* p, uint32_t v) { if (v < *p) *p = v; }
SAH_HD void a_max(uint32_t* p, uint32_t v) { if (v > *p) *p = v; }
SAH_HD void a_imin(int* p, int v) { if (v < *p) *p = v; }
SAH_HD void a_imax(int* p, int v) { if (v > *p) *p = v; }
SAH_HD int a_add(int* p, int v) { const int o = *p; *p = o + v; return o; }
SAH_HD void a_add64(unsigned long long* p, unsigned long long v) { *p += v; }
#endif

SAH_HD void stats_clear(GStats& G) {
    for (int a = 0; a < 3; a++) { G.lo[a] = G.clo[a] = kKeyPosInf; G.hi[a] = G.chi[a] = kKeyNegInf; }
    G.cost = G.n_spheres = 0; G.min_ref = 0x7FFFFFFF; G.max_ref = -1; G.n_ref_left = 0; G.pad[0] = G.pad[1] = G.pad[2] = 0;
}
SAH_HD void stats_decode(const GStats& G, int n, SahNodeStats& S) {
    for (int a = 0; a < 3; a++) { S.lo[a] = float_of(G.lo[a]); S.hi[a] = float_of(G.hi[a]); S.clo[a] = float_of(G.clo[a]); S.chi[a] = float_of(G.chi[a]); }
    S.n = n; S.cost = G.cost; S.n_spheres = G.n_spheres; S.min_ref = G.min_ref; S.max_ref = G.max_ref;
}
SAH_HD uint32_t bin_identity(int w) { const int f = w & 7; return f < 3 ? kKeyPosInf : (f < 6 ? kKeyNegInf : 0u); }

// ---- level kernels ------------------------------------------------------------------------------------------------------
// one thread per active slot: fresh statistics and bins
SAH_HD void body_clear(const Ctx& c, int s) {
    stats_clear(c.stats[s]);
    uint32_t* B = c.bins + (size_t)s * kBinWords;
    for (int w = 0; w < kBinWords; w++) B[w] = bin_identity(w);
}
// one thread per position
SAH_HD void body_stats(const Ctx& c, int cur, int pos) {
    const int s = c.slot[cur][pos];
    if (s < 0) return;
    const int ref = c.idx[cur][pos];
    const BuildPrim p = c.prims[ref];
    GStats& G = c.stats[s];
    for (int a = 0; a < 3; a++) {
        a_min(&G.lo[a], key_of(p.lo[a])); a_max(&G.hi[a], key_of(p.hi[a]));
        const uint32_t ck = key_of(sah_centroid(p, a));
        a_min(&G.clo[a], ck); a_max(&G.chi[a], ck);
    }
    a_add(&G.cost, sah_prim_cost(p.type));
    if (p.type != MORT_OBJ_QUAD) a_add(&G.n_spheres, 1);
    a_imin(&G.min_ref, ref); a_imax(&G.max_ref, ref);
}
// one thread per position; `B` = the bins this thread adds to (its slot's, or a block-private copy of them)
SAH_HD void body_bin(const Ctx& c, int cur, int pos, int s, uint32_t* B) {
    const int ref = c.idx[cur][pos];
    const BuildPrim p = c.prims[ref];
    const GStats& G = c.stats[s];
    bool any_open = false;
    for (int a = 0; a < 3; a++) {
        const float lo = float_of(G.clo[a]), hi = float_of(G.chi[a]);
        if (!((hi - lo) > 0)) continue;
        any_open = true;
        const float scale = (float)kSahBins / (hi - lo);
        const int k = sah_bin_of(sah_centroid(p, a), lo, scale);
        uint32_t* b = B + (a * kSahBins + k) * 8;
        for (int x = 0; x < 3; x++) { a_min(b + x, key_of(p.lo[x])); a_max(b + 3 + x, key_of(p.hi[x])); }
        a_add(reinterpret_cast<int*>(b + 6), 1); a_add(reinterpret_cast<int*>(b + 7), sah_prim_cost(p.type));
    }
    if (!any_open) {                                   // all centroids coincide: the index split needs its left count
        const int pivot = G.min_ref + (G.max_ref - G.min_ref + 1) / 2;
        if (ref < pivot) a_add(&c.stats[s].n_ref_left, 1);
    }
}
// one thread per active slot: the host's decision from the decoded bins; children allocated and routed
SAH_HD void body_split(const Ctx& c, int cur, int s) {
    const Active A = c.active[cur][s];
    SahNodeStats S; stats_decode(c.stats[s], A.e - A.b, S);
    float best_cost = INFINITY; int best_axis = -1, best_bin = -1, best_left = 0;
    const float parent_area = sah_max(sah_area(S.lo, S.hi), 1e-30f);
    const uint32_t* Bw = c.bins + (size_t)s * kBinWords;
    for (int a = 0; a < 3; a++) {
        if (!sah_axis_open(S, a)) continue;
        SahBins B;
        for (int k = 0; k < kSahBins; k++) {
            const uint32_t* b = Bw + (a * kSahBins + k) * 8;
            for (int x = 0; x < 3; x++) { B.lo[k][x] = float_of(b[x]); B.hi[k][x] = float_of(b[3 + x]); }
            B.cnt[k] = (int)b[6]; B.cst[k] = (int)b[7];
        }
        sah_sweep_axis(a, B, S.n, parent_area, c.P.k_trav, best_cost, best_axis, best_bin, best_left);
    }
    const SahSplit d = sah_decide(S, c.P, best_cost, best_axis, best_bin, best_left, c.stats[s].n_ref_left);   // never a leaf: n > k_small >= max_leaf
    Node2 N; sah_node_set_box(N, S);
    const int ch = a_add(&c.cnt->n_nodes, 2);
    N.left = ch; N.right = ch + 1;
    c.nodes[A.node] = N;
    SplitRec R; R.d = d; R.b = A.b; R.cur_l = R.cur_r = 0;
    const int mid = A.b + d.n_left;
    for (int side = 0; side < 2; side++) {
        const Active C = {ch + side, side == 0 ? A.b : mid, side == 0 ? mid : A.e};
        int sl = -1;
        if (C.e - C.b > c.k_small) { sl = a_add(&c.cnt->n_next, 1); c.active[cur ^ 1][sl] = C; }
        else c.small[a_add(&c.cnt->n_small, 1)] = C;
        if (side == 0) R.slot_l = sl; else R.slot_r = sl;
    }
    c.split[s] = R;
}
// one thread per position: move to the child's range.  `k` = this position's rank among the primitives that went the same
// way (from the node's cursor; gpu_build.cu reserves a warp's ranks with one atomic when the warp lies inside one node).
SAH_HD bool partition_side(const Ctx& c, int cur, int pos, int s) {
    const int ref = c.idx[cur][pos];
    return sah_goes_left(c.split[s].d, c.prims[ref], ref);
}
SAH_HD void partition_place(const Ctx& c, int cur, int pos, int s, bool left, int k) {
    const SplitRec& R = c.split[s];
    const int np = left ? R.b + k : R.b + R.d.n_left + k;
    c.idx[cur ^ 1][np] = c.idx[cur][pos];
    c.slot[cur ^ 1][np] = left ? R.slot_l : R.slot_r;
}
SAH_HD void body_partition(const Ctx& c, int cur, int pos) {
    const int s = c.slot[cur][pos];
    if (s < 0) { c.idx[cur ^ 1][pos] = c.idx[cur][pos]; c.slot[cur ^ 1][pos] = -1; return; }
    const bool left = partition_side(c, cur, pos, s);
    const int k = a_add(left ? &c.split[s].cur_l : &c.split[s].cur_r, 1);
    partition_place(c, cur, pos, s, left, k);
}
// one thread per small subtree
struct CounterAlloc { int* next; SAH_HD int pair() { return a_add(next, 2); } };
SAH_HD void body_small(const Ctx& c, int cur, int j) {
    const Active A = c.small[j];
    CounterAlloc al = {&c.cnt->n_nodes};
    sah_build_subtree(c.prims, c.idx[cur], c.nodes, A.node, A.b, A.e, c.P, al);
}

// ---- collapse: level L of the 4-wide tree = nodes [base, base + n); lvl[cur][i] = the binary node the i-th of them stands for ----
SAH_HD void body_collapse_count(const Ctx& c, int cur, int i) {
    int ch[4];
    const int nc = bvh4_open_children(c.nodes, c.lvl[cur][i], ch);
    int internal = 0;
    for (int k = 0; k < 4; k++) { c.kids[4 * i + k] = k < nc ? ch[k] : -1; if (k < nc && c.nodes[ch[k]].count == 0) internal++; }
    c.icount[i] = internal;
}
SAH_HD unsigned long long sah_fx(float v) { return (unsigned long long)((double)v * 4294967296.0); }
SAH_HD void body_collapse_emit(const Ctx& c, int cur, int i, int base, int n) {
    Bvh4Node node; bvh4_clear_node(node);
    int r = 0, leaves = 0;
    unsigned long long fx = sah_fx(bvh4_node_sah(c.nodes[c.lvl[cur][i]], c.P.k_trav, c.root_area));
    for (int k = 0; k < 4; k++) {
        const int kid = c.kids[4 * i + k];
        if (kid < 0) break;
        const Node2 C = c.nodes[kid];
        bvh4_set_child_box(node, k, C);
        if (C.count > 0) { node.child[k] = bvh4_leaf_word(C); leaves++; fx += sah_fx(bvh4_leaf_sah(C, c.root_area)); }
        else { const int j = c.ioff[i] + r; r++; node.child[k] = (uint32_t)(base + n + j); c.lvl[cur ^ 1][j] = kid; }
    }
    c.out[base + i] = node;
    if (leaves) a_add(&c.cnt->leaf_slots, leaves);
    a_add64(&c.cnt->sah_fx, fx);
}

}  // namespace gb
}  // namespace mort
