// bvh_sah.hpp — the binned-SAH split rule and the sequential subtree builder, shared by the host builder (bvh_build.cpp)
// and the GPU builder (gpu_build.cu: level-synchronous kernels for the large nodes, one thread per small subtree).
//
// Every decision is a function of the SET of primitives below a node, never of their order in the index array:
//   * bin boxes are min / max unions, bin counts and costs are integers (a sphere test costs 10, a quad test 13 tenths),
//   * a node whose centroids all coincide is split by primitive type (spheres left) or, when homogeneous, by primitive index
//     around the middle of the node's index range,
//   * the primitives of a leaf are finally sorted by index.
// So the host build, the GPU build with its atomically ordered partitions, and any thread count give the same tree, node box
// for node box — the host builder is the CPU statement the GPU one is tested against (tests/test_gpu_build.py), and "SAH cost
// within x % of the host builder" is an equality.
#pragma once
#include <math.h>
#include <stdint.h>

#include "flatten.hpp"

#if defined(__CUDACC__)
#define SAH_HD __host__ __device__ inline
#else
#define SAH_HD inline
#endif

namespace mort {

constexpr int kSahBins = 16;
struct SahParams { int max_leaf; float k_trav; };   // k_trav: cost of one more (binary) node relative to 10 tenths = one sphere test

struct Node2 {                                      // binary tree the 4-wide collapse starts from
    float lo[3], hi[3];
    int left, right;                                // children (internal) or -1
    int first, count, type;                         // leaf: range in the index array + primitive type; count == 0: internal
    int pad;
};
struct SahBins { float lo[kSahBins][3], hi[kSahBins][3]; int cnt[kSahBins], cst[kSahBins]; };   // one axis
enum { SAH_LEAF = 0, SAH_SPLIT_BIN = 1, SAH_SPLIT_TYPE = 2, SAH_SPLIT_REF = 3 };
struct SahSplit { int mode, axis, bin, pivot; float lo, scale; int n_left; };

SAH_HD int sah_prim_cost(int type) { return type == MORT_OBJ_QUAD ? 13 : 10; }
SAH_HD float sah_cost_f(int tenths) { return (float)tenths * 0.1f; }
SAH_HD float sah_area(const float* lo, const float* hi) {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0.f;
    return 2.f * (dx * dy + dy * dz + dz * dx);
}
SAH_HD float sah_centroid(const BuildPrim& p, int a) { return 0.5f * (p.lo[a] + p.hi[a]); }
SAH_HD int sah_bin_of(float c, float lo, float scale) {
    int k = (int)((c - lo) * scale);
    return k < 0 ? 0 : (k > kSahBins - 1 ? kSahBins - 1 : k);
}
SAH_HD void sah_bins_clear(SahBins& B) {
    for (int k = 0; k < kSahBins; k++) {
        for (int a = 0; a < 3; a++) { B.lo[k][a] = INFINITY; B.hi[k][a] = -INFINITY; }
        B.cnt[k] = 0; B.cst[k] = 0;
    }
}
// plain comparisons (boxes never hold NaNs: flatten_scene rejects them): one min / max instruction on either side
SAH_HD float sah_min(float a, float b) { return b < a ? b : a; }
SAH_HD float sah_max(float a, float b) { return b > a ? b : a; }
SAH_HD void sah_grow(float* lo, float* hi, const float* l, const float* h) {
    for (int a = 0; a < 3; a++) { lo[a] = sah_min(lo[a], l[a]); hi[a] = sah_max(hi[a], h[a]); }
}

// best split plane of one axis; updates (best_cost, best_axis, best_bin, n_left) on a strict improvement, axes in order 0,1,2
SAH_HD void sah_sweep_axis(int axis, const SahBins& B, int n, float parent_area, float k_trav, float& best_cost, int& best_axis, int& best_bin, int& best_left) {
    float right_area[kSahBins], right_cost[kSahBins];
    float rlo[3] = {INFINITY, INFINITY, INFINITY}, rhi[3] = {-INFINITY, -INFINITY, -INFINITY};
    int rc = 0;
    for (int k = kSahBins - 1; k > 0; k--) { sah_grow(rlo, rhi, B.lo[k], B.hi[k]); rc += B.cst[k]; right_area[k] = sah_area(rlo, rhi); right_cost[k] = sah_cost_f(rc); }
    float llo[3] = {INFINITY, INFINITY, INFINITY}, lhi[3] = {-INFINITY, -INFINITY, -INFINITY};
    int lc = 0, ln = 0;
    for (int k = 0; k < kSahBins - 1; k++) {
        sah_grow(llo, lhi, B.lo[k], B.hi[k]); lc += B.cst[k]; ln += B.cnt[k];
        if (ln == 0 || ln == n) continue;
        const float c = k_trav + (sah_area(llo, lhi) * sah_cost_f(lc) + right_area[k + 1] * right_cost[k + 1]) / parent_area;
        if (c < best_cost) { best_cost = c; best_axis = axis; best_bin = k; best_left = ln; }
    }
}

// what every builder knows about a node before it looks at bins
struct SahNodeStats {
    float lo[3], hi[3], clo[3], chi[3];             // box, centroid box
    int n, cost, n_spheres, min_ref, max_ref;       // cost in tenths
};
SAH_HD void sah_stats_clear(SahNodeStats& S) {
    for (int a = 0; a < 3; a++) { S.lo[a] = S.clo[a] = INFINITY; S.hi[a] = S.chi[a] = -INFINITY; }
    S.n = S.cost = S.n_spheres = 0; S.min_ref = 0x7FFFFFFF; S.max_ref = -1;
}
SAH_HD void sah_stats_add(SahNodeStats& S, const BuildPrim& p, int ref) {
    sah_grow(S.lo, S.hi, p.lo, p.hi);
    float c[3] = {sah_centroid(p, 0), sah_centroid(p, 1), sah_centroid(p, 2)};
    sah_grow(S.clo, S.chi, c, c);
    S.n++; S.cost += sah_prim_cost(p.type); S.n_spheres += p.type == MORT_OBJ_QUAD ? 0 : 1;
    S.min_ref = ref < S.min_ref ? ref : S.min_ref; S.max_ref = ref > S.max_ref ? ref : S.max_ref;
}
SAH_HD bool sah_homogeneous(const SahNodeStats& S) { return S.n_spheres == 0 || S.n_spheres == S.n; }
SAH_HD bool sah_axis_open(const SahNodeStats& S, int a) { return (S.chi[a] - S.clo[a]) > 0; }
SAH_HD float sah_axis_scale(const SahNodeStats& S, int a) { return (float)kSahBins / (S.chi[a] - S.clo[a]); }
SAH_HD int sah_ref_pivot(const SahNodeStats& S) { return S.min_ref + (S.max_ref - S.min_ref + 1) / 2; }      // min < pivot <= max when min < max

// the decision, once the best plane over the open axes is known (best_axis < 0: none).  n_ref_left: primitives with
// index < sah_ref_pivot (only read for SAH_SPLIT_REF).
SAH_HD SahSplit sah_decide(const SahNodeStats& S, const SahParams& P, float best_cost, int best_axis, int best_bin, int best_left, int n_ref_left) {
    SahSplit d; d.mode = SAH_LEAF; d.axis = best_axis; d.bin = best_bin; d.pivot = 0; d.lo = 0.f; d.scale = 0.f; d.n_left = 0;
    const bool can_leaf = S.n <= P.max_leaf && sah_homogeneous(S);
    if (can_leaf && (best_axis < 0 || sah_cost_f(S.cost) <= best_cost)) return d;
    if (best_axis >= 0) { d.mode = SAH_SPLIT_BIN; d.lo = S.clo[best_axis]; d.scale = sah_axis_scale(S, best_axis); d.n_left = best_left; }
    else if (!sah_homogeneous(S)) { d.mode = SAH_SPLIT_TYPE; d.n_left = S.n_spheres; }
    else if (S.n >= 2 && S.min_ref < S.max_ref) { d.mode = SAH_SPLIT_REF; d.pivot = sah_ref_pivot(S); d.n_left = n_ref_left; }
    // else: one primitive (or duplicates of one index — cannot happen) that may not be a leaf: impossible with max_leaf >= 1
    return d;
}
SAH_HD bool sah_goes_left(const SahSplit& d, const BuildPrim& p, int ref) {
    if (d.mode == SAH_SPLIT_BIN) return sah_bin_of(sah_centroid(p, d.axis), d.lo, d.scale) <= d.bin;
    if (d.mode == SAH_SPLIT_TYPE) return p.type != MORT_OBJ_QUAD;
    return ref < d.pivot;
}

SAH_HD void sah_node_set_box(Node2& N, const SahNodeStats& S) {
    for (int a = 0; a < 3; a++) { N.lo[a] = S.lo[a]; N.hi[a] = S.hi[a]; }
    N.left = N.right = -1; N.first = 0; N.count = 0; N.type = 0; N.pad = 0;
}

// Sequential builder of the subtree of `root` over idx[b, e).  Alloc::pair() returns the index of the first of two fresh,
// adjacent nodes.  Continues with the smaller child and stacks the larger one, so 32 stack entries hold any tree.
template <class Alloc>
SAH_HD void sah_build_subtree(const BuildPrim* prims, int* idx, Node2* nodes, int root, int b, int e, const SahParams& P, Alloc& alloc) {
    struct Item { int node, b, e; };
    Item stack[32]; int sp = 0;
    Item cur = {root, b, e};
    for (;;) {
        SahNodeStats S; sah_stats_clear(S);
        for (int i = cur.b; i < cur.e; i++) sah_stats_add(S, prims[idx[i]], idx[i]);
        Node2 N; sah_node_set_box(N, S);
        float best_cost = INFINITY; int best_axis = -1, best_bin = -1, best_left = 0;
        const float parent_area = sah_max(sah_area(S.lo, S.hi), 1e-30f);
        if (S.n >= 2)
            for (int a = 0; a < 3; a++) {
                if (!sah_axis_open(S, a)) continue;
                SahBins B; sah_bins_clear(B);
                const float scale = sah_axis_scale(S, a), lo = S.clo[a];
                for (int i = cur.b; i < cur.e; i++) {
                    const BuildPrim& p = prims[idx[i]];
                    const int k = sah_bin_of(sah_centroid(p, a), lo, scale);
                    sah_grow(B.lo[k], B.hi[k], p.lo, p.hi); B.cnt[k]++; B.cst[k] += sah_prim_cost(p.type);
                }
                sah_sweep_axis(a, B, S.n, parent_area, P.k_trav, best_cost, best_axis, best_bin, best_left);
            }
        int n_ref_left = 0;
        if (best_axis < 0 && sah_homogeneous(S)) { const int pv = sah_ref_pivot(S); for (int i = cur.b; i < cur.e; i++) n_ref_left += idx[i] < pv ? 1 : 0; }
        const SahSplit d = sah_decide(S, P, best_cost, best_axis, best_bin, best_left, n_ref_left);
        bool descend = false;
        if (d.mode == SAH_LEAF) {
            N.first = cur.b; N.count = S.n; N.type = prims[idx[cur.b]].type;
            // canonical leaf: primitives by index
            for (int i = cur.b + 1; i < cur.e; i++) { const int v = idx[i]; int j = i - 1; while (j >= cur.b && idx[j] > v) { idx[j + 1] = idx[j]; j--; } idx[j + 1] = v; }
            nodes[cur.node] = N;
        } else {
            int i = cur.b, j = cur.e - 1;
            for (;;) {
                while (i <= j && sah_goes_left(d, prims[idx[i]], idx[i])) i++;
                while (i <= j && !sah_goes_left(d, prims[idx[j]], idx[j])) j--;
                if (i >= j) break;
                const int t = idx[i]; idx[i] = idx[j]; idx[j] = t; i++; j--;
            }
            const int mid = i;                                   // == cur.b + d.n_left
            const int c = alloc.pair();
            N.left = c; N.right = c + 1;
            nodes[cur.node] = N;
            const Item L = {c, cur.b, mid}, R = {c + 1, mid, cur.e};
            const bool left_small = (mid - cur.b) <= (cur.e - mid);
            stack[sp++] = left_small ? R : L;
            cur = left_small ? L : R;
            descend = true;
        }
        if (!descend) { if (sp == 0) break; cur = stack[--sp]; }
    }
}

// ---- collapse of the binary tree to 4-wide nodes (shared by the host queue loop and the GPU level kernels) ----------------
SAH_HD void bvh4_clear_node(Bvh4Node& n) {
    for (int k = 0; k < 4; k++) {
        n.lox[k] = n.loy[k] = n.loz[k] = INFINITY; n.hix[k] = n.hiy[k] = n.hiz[k] = -INFINITY;
        n.child[k] = MORT_CHILD_EMPTY; n.spare[k] = 0;
    }
}
// children of the 4-wide node that stands for binary node n2: its two children, then twice the internal child with the largest
// surface (the first one on ties) is replaced by its own two children.  A binary LEAF as the root stands alone.
SAH_HD int bvh4_open_children(const Node2* nodes, int n2, int ch[4]) {
    int nc = 0;
    if (nodes[n2].count > 0) ch[nc++] = n2;
    else { ch[nc++] = nodes[n2].left; ch[nc++] = nodes[n2].right; }
    while (nc < 4) {
        int pick = -1; float best = -1.f;
        for (int k = 0; k < nc; k++) {
            const Node2& c = nodes[ch[k]];
            if (c.count == 0) { const float a = sah_area(c.lo, c.hi); if (a > best) { best = a; pick = k; } }
        }
        if (pick < 0) break;
        const int c = ch[pick];
        ch[pick] = nodes[c].left; ch[nc++] = nodes[c].right;
    }
    return nc;
}
SAH_HD void bvh4_set_child_box(Bvh4Node& dst, int k, const Node2& c) {
    dst.lox[k] = c.lo[0]; dst.loy[k] = c.lo[1]; dst.loz[k] = c.lo[2];
    dst.hix[k] = c.hi[0]; dst.hiy[k] = c.hi[1]; dst.hiz[k] = c.hi[2];
}
SAH_HD uint32_t bvh4_leaf_word(const Node2& c) {
    return MORT_LEAF_BIT | (c.type == MORT_OBJ_QUAD ? MORT_LEAF_QUAD_BIT : 0u) | ((uint32_t)(c.count - 1) << 27) | (uint32_t)c.first;
}
SAH_HD float bvh4_node_sah(const Node2& n, float k_trav, float root_area) { return k_trav * sah_area(n.lo, n.hi) / root_area; }
SAH_HD float bvh4_leaf_sah(const Node2& c, float root_area) { return (float)c.count * sah_cost_f(sah_prim_cost(c.type)) * sah_area(c.lo, c.hi) / root_area; }

}  // namespace mort
