// trace_body.cuh — one ray of the parity hook mort_trace (what oracle/ref_harness.cu records from the reference): closest hit through
// the tree or by brute force, the full hit record, and the boundary probes of every medium.  Shared by render.cu's trace_kernel and
// the motion-box unit (motion.cu), which compiles the traversal with interpolated node boxes.
#pragma once
#include "rt_core.cuh"

namespace mort {

__device__ __forceinline__ void trace_one(const DeviceScene& sc, const float* __restrict__ rays, int i, mhit_record* __restrict__ out,
                                          mhit_medium_probe* __restrict__ probes, int brute_force, const int32_t* __restrict__ mat_offsets) {
    const float* q = rays + 7 * (size_t)i;
    Ray r; r.o = mk3(q[0], q[1], q[2]); r.d = mk3(q[3], q[4], q[5]); r.tm = q[6];
    Hit h;
    bool any = brute_force ? closest_hit_brute(sc, r, 0.001f, INFINITY, h) : closest_hit<false>(sc, nullptr, 0, r, 0.001f, INFINITY, h);
    mhit_record o;
    o.hit = any ? 1 : 0; o.t = 0.f; o.leaf_type = o.leaf_idx = o.top_type = o.top_idx = -1;
    o.mat_type = o.mat_idx = o.front_face = o.flags = 0;
    o.p[0] = o.p[1] = o.p[2] = o.normal[0] = o.normal[1] = o.normal[2] = o.u = o.v = 0.f;
    if (any) {
        Record rec; resolve_hit(sc, r, h, rec);
        if (rec.sphere_uv) sphere_uv(rec.outward, rec.u, rec.v);
        o.t = rec.t; o.leaf_type = rec.leaf_type; o.leaf_idx = rec.leaf_idx; o.front_face = rec.front_face ? 1 : 0;
        uint32_t ri = h.prim & 0x07FFFFFFu;
        int top = (h.prim & MORT_LEAF_QUAD_BIT) ? sc.quads[ri].pad[0] : sc.sphere_info[ri].pad;
        o.top_type = top >> 24; o.top_idx = top & 0xFFFFFF;
        if (rec.mat_gid >= 0) {
            int type = sc.materials[rec.mat_gid].type;
            o.mat_type = type; o.mat_idx = rec.mat_gid - mat_offsets[type];
        } else { o.mat_type = -1; o.mat_idx = -1; }
        o.p[0] = rec.p.x; o.p[1] = rec.p.y; o.p[2] = rec.p.z;
        o.normal[0] = rec.normal.x; o.normal[1] = rec.normal.y; o.normal[2] = rec.normal.z;
        o.u = rec.u; o.v = rec.v;
    }
    out[i] = o;
    if (probes)
        for (int m = 0, j = 0; m < sc.n_media; m++) {
            if (!sc.media[m].top_level) continue;
            mhit_medium_probe p; p.hit1 = p.hit2 = 0; p.t1 = p.t2 = 0.f;
            float t1, t2;
            if (boundary_probe(sc, sc.media[m], r, -INFINITY, INFINITY, t1)) {
                p.hit1 = 1; p.t1 = t1;
                if (boundary_probe(sc, sc.media[m], r, (float)((double)t1 + 0.0001), INFINITY, t2)) { p.hit2 = 1; p.t2 = t2; }
            }
            probes[(size_t)i * sc.n_media_top + j] = p; j++;
        }
}

}  // namespace mort
