// refit_core.cuh — per-thread bodies of the bottom-up bounds pass over a committed 4-wide BVH (refit.cu wraps them in
// kernels; tests/hostsim/buildsim.cpp runs them serially against the boxes the builder produced).
//
// Two uses:
//   refit           after spheres of a committed scene moved (mort_update_sphere): the topology stays, every primitive box is
//                   recomputed from the DEVICE records and every node box bottom-up, one kernel per tree level
//   motion bounds   the same pass with two time keys: boxes at ray time 0 and at ray time 1 instead of their union (the
//                   reference's moving-sphere box, objects.cuh:46-55); traversal interpolates them at the ray's time
// Primitive boxes follow flatten.cpp (leaf_world_box + the 2e-6 padding) operation for operation, so a refit of an unchanged
// scene reproduces the builder's node boxes bit for bit.
#pragma once
#include <math.h>
#include <stdint.h>

#include "bvh_sah.hpp"
#include "device_types.h"

namespace mort {
namespace rf {

struct Box { float lo[3], hi[3]; };
enum { TIME_UNION = 0, TIME_0 = 1, TIME_1 = 2 };

struct Ctx {
    Bvh4Node* nodes;            // union boxes, or the time-0 boxes when `node_t1` is set
    Bvh4Node* node_t1;          // motion bounds: time-1 boxes during the pass, then turned into (time 1 - time 0) by body_delta
    const SphereGeom* spheres; const QuadRec* quads; const Instance* instances;
    int n_spheres, n_quads;
    Box* sphere_box[2]; Box* quad_box[2];     // [0]: union or time 0, [1]: time 1 (motion only)
    float pad;
    unsigned* extent_key;       // max |coordinate| over the raw boxes, as float bits (non-negative floats order like unsigned ints)
};

SAH_HD void to_world(const Instance* instances, int inst, float p[3]) {
    if (inst < 0) return;
    const Instance& I = instances[inst];
    for (int k = I.nops - 1; k >= 0; k--) {
        int kind;
#if defined(__CUDA_ARCH__)
        kind = __float_as_int(I.a[k][3]);
#else
        memcpy(&kind, &I.a[k][3], 4);
#endif
        if (kind == INST_OP_TRANSLATE) { for (int a = 0; a < 3; a++) p[a] += I.a[k][a]; }
        else {
            const float sn = I.a[k][0], cs = I.a[k][1];
            const float x = cs * p[0] + sn * p[2], z = -sn * p[0] + cs * p[2];
            p[0] = x; p[2] = z;
        }
    }
}
SAH_HD void grow_point(Box& b, const float* p, float r) {
    for (int a = 0; a < 3; a++) { b.lo[a] = fminf(b.lo[a], p[a] - r); b.hi[a] = fmaxf(b.hi[a], p[a] + r); }
}
SAH_HD void box_clear(Box& b) { for (int a = 0; a < 3; a++) { b.lo[a] = INFINITY; b.hi[a] = -INFINITY; } }
SAH_HD void box_union(Box& b, const Box& o) { for (int a = 0; a < 3; a++) { b.lo[a] = sah_min(b.lo[a], o.lo[a]); b.hi[a] = sah_max(b.hi[a], o.hi[a]); } }
SAH_HD float box_extent(const Box& b) { float m = 0.f; for (int a = 0; a < 3; a++) { m = fmaxf(m, fabsf(b.lo[a])); m = fmaxf(m, fabsf(b.hi[a])); } return m; }
SAH_HD void box_pad(Box& b, float pad) { for (int a = 0; a < 3; a++) { b.lo[a] -= pad + 2e-6f * fabsf(b.lo[a]); b.hi[a] += pad + 2e-6f * fabsf(b.hi[a]); } }

// raw (unpadded) world box of one record at a time key (flatten.cpp: leaf_world_box)
SAH_HD Box sphere_raw_box(const Ctx& c, int i, int time) {
    const SphereGeom g = c.spheres[i];
    float c0[3] = {g.cx, g.cy, g.cz}, c1[3] = {g.cx + g.vx, g.cy + g.vy, g.cz + g.vz};
    const bool moves = g.vx != 0.f || g.vy != 0.f || g.vz != 0.f;
    Box b; box_clear(b);
    to_world(c.instances, g.inst, c0);
    if (moves) to_world(c.instances, g.inst, c1);
    if (time != TIME_1 || !moves) grow_point(b, c0, fabsf(g.r));
    if (moves && time != TIME_0) grow_point(b, c1, fabsf(g.r));
    return b;
}
SAH_HD Box quad_raw_box(const Ctx& c, int i) {
    const QuadRec q = c.quads[i];
    Box b; box_clear(b);
    for (int ci = 0; ci < 2; ci++) for (int cj = 0; cj < 2; cj++) {
        float p[3] = {q.Qx + ci * q.ux + cj * q.vx, q.Qy + ci * q.uy + cj * q.vy, q.Qz + ci * q.uz + cj * q.vz};
        to_world(c.instances, q.inst, p); grow_point(b, p, 0.f);
    }
    return b;
}
#if defined(__CUDA_ARCH__)
SAH_HD void extent_max(unsigned* p, float v) { atomicMax(p, __float_as_uint(v)); }
#else
SAH_HD void extent_max(unsigned* p, float v) { unsigned u; memcpy(&u, &v, 4); if (u > *p) *p = u; }
#endif

// pass 1, one thread per record (spheres first, then quads): raw boxes + the scene extent
SAH_HD void body_raw(const Ctx& c, int i, bool motion) {
    float ext;
    if (i < c.n_spheres) {
        c.sphere_box[0][i] = sphere_raw_box(c, i, motion ? TIME_0 : TIME_UNION);
        ext = box_extent(c.sphere_box[0][i]);
        if (motion) { c.sphere_box[1][i] = sphere_raw_box(c, i, TIME_1); ext = fmaxf(ext, box_extent(c.sphere_box[1][i])); }
    } else {
        const int q = i - c.n_spheres;
        c.quad_box[0][q] = quad_raw_box(c, q);
        if (motion) c.quad_box[1][q] = c.quad_box[0][q];
        ext = box_extent(c.quad_box[0][q]);
    }
    extent_max(c.extent_key, ext);
}
// pass 2, one thread per record: padding (c.pad = 2e-6 * max(extent, |camera centre|), flatten.cpp)
SAH_HD void body_pad(const Ctx& c, int i, bool motion) {
    for (int t = 0; t < (motion ? 2 : 1); t++) {
        if (i < c.n_spheres) box_pad(c.sphere_box[t][i], c.pad); else box_pad(c.quad_box[t][i - c.n_spheres], c.pad);
    }
}
// pass 3, deepest level first, one thread per (node, child slot): leaf = union of its records' boxes, internal = union of the
// child node's four boxes (already written: it lies in a deeper level)
SAH_HD void child_box(const Ctx& c, const Bvh4Node* arr, int t, uint32_t w, Box& b) {
    box_clear(b);
    if (w & MORT_LEAF_BIT) {
        const int first = (int)(w & 0x07FFFFFFu), cnt = (int)((w >> 27) & 7u) + 1;
        const Box* src = (w & MORT_LEAF_QUAD_BIT) ? c.quad_box[t] : c.sphere_box[t];
        for (int k = 0; k < cnt; k++) box_union(b, src[first + k]);
    } else {
        const Bvh4Node& n = arr[w];
        for (int k = 0; k < 4; k++) {
            if (n.child[k] == MORT_CHILD_EMPTY) continue;
            const Box cb = {{n.lox[k], n.loy[k], n.loz[k]}, {n.hix[k], n.hiy[k], n.hiz[k]}};
            box_union(b, cb);
        }
    }
}
SAH_HD void body_level(const Ctx& c, int node, int k, bool motion) {
    Bvh4Node& n = c.nodes[node];
    const uint32_t w = n.child[k];
    if (motion) { Bvh4Node& m = c.node_t1[node]; m.child[k] = w; m.spare[k] = 0; }
    if (w == MORT_CHILD_EMPTY) {
        if (motion) { Bvh4Node& m = c.node_t1[node]; m.lox[k] = m.loy[k] = m.loz[k] = INFINITY; m.hix[k] = m.hiy[k] = m.hiz[k] = -INFINITY; }
        return;
    }
    Box b; child_box(c, c.nodes, 0, w, b);
    n.lox[k] = b.lo[0]; n.loy[k] = b.lo[1]; n.loz[k] = b.lo[2]; n.hix[k] = b.hi[0]; n.hiy[k] = b.hi[1]; n.hiz[k] = b.hi[2];
    if (motion) {
        Bvh4Node& m = c.node_t1[node];
        child_box(c, c.node_t1, 1, w, b);
        m.lox[k] = b.lo[0]; m.loy[k] = b.lo[1]; m.loz[k] = b.lo[2]; m.hix[k] = b.hi[0]; m.hiy[k] = b.hi[1]; m.hiz[k] = b.hi[2];
    }
}
// pass 4 (motion), one thread per (node, child slot): time-1 box -> difference to the time-0 box; empty slots get 0 so that the
// interpolated box stays the inverted (+inf, -inf) box no ray can enter
SAH_HD void body_delta(const Ctx& c, int node, int k) {
    const Bvh4Node& n = c.nodes[node]; Bvh4Node& m = c.node_t1[node];
    if (n.child[k] == MORT_CHILD_EMPTY) { m.lox[k] = m.loy[k] = m.loz[k] = m.hix[k] = m.hiy[k] = m.hiz[k] = 0.f; return; }
    m.lox[k] -= n.lox[k]; m.loy[k] -= n.loy[k]; m.loz[k] -= n.loz[k]; m.hix[k] -= n.hix[k]; m.hiy[k] -= n.hiy[k]; m.hiz[k] -= n.hiz[k];
}

}  // namespace rf
}  // namespace mort
