// flatten.cpp — Scene (reference-shaped handles) -> FlatScene (device layout) — see flatten.hpp.
#include "flatten.hpp"

#include <cmath>
#include <cstring>
#include <algorithm>
#include <map>

#define MORT_LINEAR_MAX_LEAVES 40      // == MORT_LINEAR_MAX in rt_core.cuh

namespace mort {
namespace {

// plain comparisons instead of fminf / fmaxf (libm calls without -ffast-math): a NaN coordinate falls through them into the
// box and is refused by the check after leaf_world_box
inline float fmin_(float a, float b) { return b < a ? b : a; }
inline float fmax_(float a, float b) { return b > a ? b : a; }

struct Chain { int n = 0; int kind[MORT_INSTANCE_OPS]; int idx[MORT_INSTANCE_OPS]; bool overflow = false;
    bool operator<(const Chain& o) const {
        if (n != o.n) return n < o.n;
        for (int i = 0; i < n; i++) { if (kind[i] != o.kind[i]) return kind[i] < o.kind[i]; if (idx[i] != o.idx[i]) return idx[i] < o.idx[i]; }
        return false;
    } };

struct Flattener {
    const Scene& s; FlatScene& out; std::string err;
    std::map<Chain, int> inst_ids;
    std::vector<Chain> inst_chain;               // instance id -> its chain (what leaf_world_box needs; one entry per distinct chain, not per leaf)
    Flattener(const Scene& sc, FlatScene& o) : s(sc), out(o) {}

    int instance_of(const Chain& c) {
        if (c.n == 0) return -1;
        auto it = inst_ids.find(c);
        if (it != inst_ids.end()) return it->second;
        Instance I; memset(&I, 0, sizeof(I));
        I.nops = c.n;
        for (int k = 0; k < c.n; k++) {
            int32_t kind;
            if (c.kind[k] == MORT_OBJ_TRANSLATE) {
                kind = INST_OP_TRANSLATE;
                for (int a = 0; a < 3; a++) I.a[k][a] = s.translates[c.idx[k]].offset[a];
            } else {
                kind = INST_OP_ROTATE_Y;
                I.a[k][0] = s.rotates[c.idx[k]].sin_theta; I.a[k][1] = s.rotates[c.idx[k]].cos_theta;
            }
            memcpy(&I.a[k][3], &kind, 4);
        }
        int id = (int)out.instances.size();
        out.instances.push_back(I); inst_ids[c] = id; inst_chain.push_back(c);
        return id;
    }

    // hitDispatch recursion (objects.cuh:858-887) unrolled into a list of leaves in visit order
    template <class F, class G> void collect(int type, int idx, const Chain& c0, int depth, F&& emit, G&& emit_medium) {
        const Chain& c = c0;
        if (depth > 16) { err = "object nesting deeper than 16 levels (cycle?)"; return; }
        switch (type) {
            case MORT_OBJ_SPHERE:
                if (idx < 0 || idx >= (int)s.spheres.size()) { err = "sphere handle out of range"; return; }
                emit(type, idx, c); break;
            case MORT_OBJ_QUAD:
                if (idx < 0 || idx >= (int)s.quads.size()) { err = "quad handle out of range"; return; }
                emit(type, idx, c); break;
            case MORT_OBJ_TRANSLATE:
                if (idx < 0 || idx >= (int)s.translates.size()) { err = "translate handle out of range"; return; }
                if (c.n >= MORT_INSTANCE_OPS) { err = "more than 7 nested translate/rotate_y wrappers are not supported"; return; }
                { Chain c2 = c; c2.kind[c2.n] = MORT_OBJ_TRANSLATE; c2.idx[c2.n++] = idx;
                  collect(s.translates[idx].obj_type, s.translates[idx].obj_idx, c2, depth + 1, emit, emit_medium); } break;
            case MORT_OBJ_ROTATE_Y:
                if (idx < 0 || idx >= (int)s.rotates.size()) { err = "rotate_y handle out of range"; return; }
                if (c.n >= MORT_INSTANCE_OPS) { err = "more than 7 nested translate/rotate_y wrappers are not supported"; return; }
                { Chain c2 = c; c2.kind[c2.n] = MORT_OBJ_ROTATE_Y; c2.idx[c2.n++] = idx;
                  collect(s.rotates[idx].obj_type, s.rotates[idx].obj_idx, c2, depth + 1, emit, emit_medium); } break;
            case MORT_OBJ_HITTABLE_LIST:
                if (idx < 0 || idx >= (int)s.lists.size()) { err = "list handle out of range"; return; }
                for (const Handle& h : s.lists[idx].items) collect(h.type, h.idx, c, depth + 1, emit, emit_medium);
                break;
            case MORT_OBJ_CONSTANT_MEDIUM:                  // hitDispatch reaches a medium through wrappers and lists (objects.cuh:875-877)
                if (idx < 0 || idx >= (int)s.media.size()) { err = "medium handle out of range"; return; }
                emit_medium(idx, c); break;
            default: break;       // hitDispatch has no case for a BVH or unknown tags: never hit
        }
    }

    // object space -> world space through an instance chain (outermost op first)
    void to_world(const Chain& c, float p[3]) const {
        for (int k = c.n - 1; k >= 0; k--) {
            if (c.kind[k] == MORT_OBJ_TRANSLATE) { for (int a = 0; a < 3; a++) p[a] += s.translates[c.idx[k]].offset[a]; }
            else {
                float sn = s.rotates[c.idx[k]].sin_theta, cs = s.rotates[c.idx[k]].cos_theta;
                float x = cs * p[0] + sn * p[2], z = -sn * p[0] + cs * p[2];
                p[0] = x; p[2] = z;
            }
        }
    }
    void leaf_world_box(int type, int idx, const Chain& c, float lo[3], float hi[3]) const {
        for (int a = 0; a < 3; a++) { lo[a] = INFINITY; hi[a] = -INFINITY; }
        auto grow = [&](const float* p, float r) { for (int a = 0; a < 3; a++) { lo[a] = fmin_(lo[a], p[a] - r); hi[a] = fmax_(hi[a], p[a] + r); } };
        if (type == MORT_OBJ_SPHERE) {
            const mscn_sphere& sp = s.spheres[idx];
            float c0[3] = {sp.center[0], sp.center[1], sp.center[2]};
            float c1[3] = {sp.center[0] + sp.center_vec[0], sp.center[1] + sp.center_vec[1], sp.center[2] + sp.center_vec[2]};
            to_world(c, c0); grow(c0, fabsf(sp.radius));
            if (sp.moves) { to_world(c, c1); grow(c1, fabsf(sp.radius)); }
        } else {
            const mscn_quad& q = s.quads[idx];
            for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) {
                float p[3] = {q.Q[0] + i * q.u[0] + j * q.v[0], q.Q[1] + i * q.u[1] + j * q.v[1], q.Q[2] + i * q.u[2] + j * q.v[2]};
                to_world(c, p); grow(p, 0.f);
            }
        }
    }
};

int mat_gid(const Scene& s, int type, int idx) {
    int off[6] = {0, 0, (int)s.lambertians.size(), 0, 0, 0};
    off[3] = off[2] + (int)s.metals.size(); off[4] = off[3] + (int)s.dielectrics.size(); off[5] = off[4] + (int)s.lights.size();
    int cnt[6] = {0, (int)s.lambertians.size(), (int)s.metals.size(), (int)s.dielectrics.size(), (int)s.lights.size(), (int)s.isotropics.size()};
    if (type < 1 || type > 5 || idx < 0 || idx >= cnt[type]) return -1;   // scatterDispatch falls through: absorbs, no emission
    return off[type] + idx;
}
int tex_gid(const Scene& s, int type, int idx) {
    int cnt[5] = {0, (int)s.solids.size(), (int)s.checkers.size(), (int)s.images.size(), (int)s.noises.size()};
    if (type < 1 || type > 4 || idx < 0 || idx >= cnt[type]) return -1;   // valueDispatch's magenta error texture
    int off = 0;
    for (int t = 1; t < type; t++) off += cnt[t];
    return off + idx;
}

void fill_quad_rows(const mscn_quad& q, float a[5][4]) {
    a[0][0] = q.normal[0]; a[0][1] = q.normal[1]; a[0][2] = q.normal[2]; a[0][3] = q.D;
    for (int k = 0; k < 3; k++) { a[1][k] = q.Q[k]; a[2][k] = q.u[k]; a[3][k] = q.v[k]; a[4][k] = q.w[k]; }
    a[1][3] = a[2][3] = a[3][3] = a[4][3] = 0;
}

}  // namespace

void camera_params(const Camera& c, CameraParams& o) {
    memset(&o, 0, sizeof(o));
    auto p3 = [](float* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
    p3(o.center, c.center); p3(o.pixel00, c.pixel00_loc); p3(o.du, c.pixel_delta_u); p3(o.dv, c.pixel_delta_v);
    p3(o.defocus_u, c.defocus_disk_u); p3(o.defocus_v, c.defocus_disk_v); p3(o.background, c.background);
    o.defocus_angle = c.defocus_angle; o.recip_sqrt_spp = c.recip_sqrt_spp; o.pixel_samples_scale = c.pixel_samples_scale;
    o.width = c.image_width; o.height = c.image_height; o.sqrt_spp = c.sqrt_spp; o.bounce_limit = c.bounce_limit;
}

bool flatten_scene(const Scene& s, FlatScene& out, std::string* err, const BuildOptions& opt, BvhBuildFn build, void* build_user) {
    out = FlatScene();
    Flattener F(s, out);
    auto fail = [&](const std::string& m) { if (err) *err = m; return false; };

    // ---- materials / textures: one running index each ----
    for (const auto& m : s.lambertians) { Material d; memset(&d, 0, sizeof(d)); d.type = MORT_MAT_LAMBERTIAN; d.tex_gid = tex_gid(s, m.tex_type, m.tex_idx); out.materials.push_back(d); }
    for (const auto& m : s.metals) { Material d; memset(&d, 0, sizeof(d)); d.type = MORT_MAT_METAL; d.tex_gid = -1; d.ax = m.albedo[0]; d.ay = m.albedo[1]; d.az = m.albedo[2]; d.p0 = m.fuzz; out.materials.push_back(d); }
    for (const auto& m : s.dielectrics) { Material d; memset(&d, 0, sizeof(d)); d.type = MORT_MAT_DIELECTRIC; d.tex_gid = -1; d.ax = m.albedo[0]; d.ay = m.albedo[1]; d.az = m.albedo[2]; d.p0 = m.ior; d.p1 = m.inv_ior; out.materials.push_back(d); }
    for (const auto& m : s.lights) { Material d; memset(&d, 0, sizeof(d)); d.type = MORT_MAT_DIFFUSE_LIGHT; d.tex_gid = tex_gid(s, m.tex_type, m.tex_idx); out.materials.push_back(d); }
    for (const auto& m : s.isotropics) { Material d; memset(&d, 0, sizeof(d)); d.type = MORT_MAT_ISOTROPIC; d.tex_gid = tex_gid(s, m.tex_type, m.tex_idx); out.materials.push_back(d); }
    for (const auto& t : s.solids) { Texture d; memset(&d, 0, sizeof(d)); d.type = MORT_TEX_SOLID; d.c0 = t.color[0]; d.c1 = t.color[1]; d.c2 = t.color[2]; out.textures.push_back(d); }
    for (const auto& t : s.checkers) { Texture d; memset(&d, 0, sizeof(d)); d.type = MORT_TEX_CHECKER; d.c0 = t.inv_scale; d.even_gid = tex_gid(s, t.even_type, t.even_idx); d.odd_gid = tex_gid(s, t.odd_type, t.odd_idx); out.textures.push_back(d); }
    for (size_t i = 0; i < s.images.size(); i++) { Texture d; memset(&d, 0, sizeof(d)); d.type = MORT_TEX_IMAGE; d.even_gid = (int)i; out.textures.push_back(d); }
    for (size_t i = 0; i < s.noises.size(); i++) {
        Texture d; memset(&d, 0, sizeof(d)); d.type = MORT_TEX_NOISE; d.even_gid = (int)i; out.textures.push_back(d);
        NoiseTables n; memset(&n, 0, sizeof(n));
        for (int k = 0; k < 256; k++) {
            for (int a = 0; a < 3; a++) n.ranvec[k][a] = s.noises[i].ranvec[k][a];
            n.perm_x[k] = (uint8_t)s.noises[i].perm_x[k]; n.perm_y[k] = (uint8_t)s.noises[i].perm_y[k]; n.perm_z[k] = (uint8_t)s.noises[i].perm_z[k];
        }
        n.scale = s.noises[i].scale;
        out.noises.push_back(n);
    }

    out.sphere_pinned.assign(s.spheres.size(), 0);
    // ---- visible leaves in world::hit order (world.cuh:110-168) ----
    int top_type = 0, top_idx = 0;
    // gate = intersection of the reference-BVH node boxes above a leaf (everything for leaves outside a bvh): the product
    // ignores those boxes, which is only right while they contain what hangs below them (checked after the world boxes
    // are known)
    struct Gate { float lo[3], hi[3]; };
    Gate open_gate; for (int a = 0; a < 3; a++) { open_gate.lo[a] = -INFINITY; open_gate.hi[a] = INFINITY; }
    Gate cur_gate = open_gate;
    std::vector<Gate> gates;
    { const size_t guess = s.spheres.size() + s.quads.size(); out.leaves.reserve(guess); gates.reserve(guess); }
    auto emit = [&](int type, int idx, const Chain& c) {
        LeafRef L; L.type = type; L.idx = idx; L.inst = F.instance_of(c); L.order = (int)out.leaves.size(); L.top_type = top_type; L.top_idx = top_idx;
        out.leaves.push_back(L); gates.push_back(cur_gate);
    };
    // media in the order world::hit reaches them: the top-level loop (world.cuh:154-160) and, through wrappers and lists, hitDispatch
    // (objects.cuh:875-877).  `pos` = leaves visited before the medium: what it is clipped against.
    struct MediumVisit { int medium, pos; Chain chain; bool top; };
    std::vector<MediumVisit> visits;
    bool under_bvh = false;
    auto emit_m = [&](int m, const Chain& c) {
        if (under_bvh) { F.err = "a constant_medium below a bvh is not supported (the reference culls it with its own node boxes)"; return; }
        visits.push_back(MediumVisit{m, (int)out.leaves.size(), c, false});
    };
    Chain none;
    for (size_t b = 0; b < s.bvhs.size(); b++) {
        if (s.bvhs[b].skip || s.bvhs[b].nodes.empty()) continue;
        top_type = MORT_OBJ_BVH; top_idx = (int)b;
        // depth-first, left before right: the order bvh::hit (objects.cuh:664-723) tests the leaves in.
        // A node whose box has no thickness along an axis (an axis-aligned quad alone in a leaf: the reference does not pad
        // boxes) can never be entered — aabb::hit rejects t_max <= t_min (aabb.cuh:38-59) — so everything below it is
        // invisible in the reference and is left out here.  Every other box contains its objects, and culling by it
        // changes nothing (tests/test_host_scene.py checks the shipped BVH scenes against the reference's hits).
        struct Visit { int node; bool dead; Gate gate; };
        std::vector<Visit> stack; stack.push_back(Visit{0, false, open_gate});
        while (!stack.empty()) {
            const int n = stack.back().node; bool dead = stack.back().dead; Gate g = stack.back().gate; stack.pop_back();
            if (n < 0 || n >= (int)s.bvhs[b].nodes.size()) return fail("bvh node index out of range");
            const mscn_bvh_node& nd = s.bvhs[b].nodes[n];
            for (int a = 0; a < 3; a++) {
                dead = dead || !(nd.bbox[2 * a + 1] > nd.bbox[2 * a]);
                g.lo[a] = fmax_(g.lo[a], nd.bbox[2 * a]); g.hi[a] = fmin_(g.hi[a], nd.bbox[2 * a + 1]);
            }
            cur_gate = g;
            if (nd.is_internal) { stack.push_back(Visit{nd.right_idx, dead, g}); stack.push_back(Visit{nd.left_idx, dead, g}); }
            else if (!dead) {
                under_bvh = true;
                F.collect(nd.left_type, nd.left_idx, none, 0, emit, emit_m);
                if (nd.right_type != nd.left_type || nd.right_idx != nd.left_idx) F.collect(nd.right_type, nd.right_idx, none, 0, emit, emit_m);
                under_bvh = false;
            }
        }
    }
    cur_gate = open_gate;
    if (!s.bvh_mode) {                       // world.cuh:118-120: with a BVH in the world nothing else is visible
        for (size_t i = 0; i < s.spheres.size(); i++) if (!s.spheres[i].skip) { top_type = MORT_OBJ_SPHERE; top_idx = (int)i; F.collect(MORT_OBJ_SPHERE, (int)i, none, 0, emit, emit_m); }
        for (size_t i = 0; i < s.quads.size(); i++) if (!s.quads[i].skip) { top_type = MORT_OBJ_QUAD; top_idx = (int)i; F.collect(MORT_OBJ_QUAD, (int)i, none, 0, emit, emit_m); }
        for (size_t i = 0; i < s.translates.size(); i++) if (!s.translates[i].skip) { top_type = MORT_OBJ_TRANSLATE; top_idx = (int)i; F.collect(MORT_OBJ_TRANSLATE, (int)i, none, 0, emit, emit_m); }
        for (size_t i = 0; i < s.rotates.size(); i++) if (!s.rotates[i].skip) { top_type = MORT_OBJ_ROTATE_Y; top_idx = (int)i; F.collect(MORT_OBJ_ROTATE_Y, (int)i, none, 0, emit, emit_m); }
        for (size_t m = 0; m < s.media.size(); m++) if (!s.media[m].skip) visits.push_back(MediumVisit{(int)m, (int)out.leaves.size(), none, true});
        for (size_t i = 0; i < s.lists.size(); i++) if (!s.lists[i].skip) { top_type = MORT_OBJ_HITTABLE_LIST; top_idx = (int)i; F.collect(MORT_OBJ_HITTABLE_LIST, (int)i, none, 0, emit, emit_m); }
    }
    if (!F.err.empty()) return fail(F.err);
    out.post_media_order = visits.empty() ? (int)out.leaves.size() : visits[0].pos;

    // ---- media records, in visit order; each carries the window of leaves visited between it and the next medium ----
    bool general = false;
    for (size_t v = 0; v < visits.size(); v++) {
        const MediumVisit& V = visits[v];
        const int m = V.medium;
        general = general || V.pos != visits[0].pos || V.chain.n > 0 || !V.top;
        Medium M; memset(&M, 0, sizeof(M));
        M.neg_inv_density = s.media[m].neg_inv_density; M.mat_gid = mat_gid(s, s.media[m].mat_type, s.media[m].mat_idx);
        M.first = (int)out.boundary.size(); M.obj_idx = m;
        M.inst = F.instance_of(V.chain); M.top_level = V.top ? 1 : 0;
        M.after_lo = V.pos; M.after_hi = v + 1 < visits.size() ? visits[v + 1].pos : 0x7FFFFFFF;
        auto emit_b = [&](int type, int idx, const Chain& c) {
            BoundaryPrim B; memset(&B, 0, sizeof(B));
            B.type = type; B.inst = F.instance_of(c);
            if (type == MORT_OBJ_SPHERE) {
                const mscn_sphere& sp = s.spheres[idx];
                out.sphere_pinned[idx] = 1;
                B.a[0][0] = sp.center[0]; B.a[0][1] = sp.center[1]; B.a[0][2] = sp.center[2]; B.a[0][3] = sp.radius;
                if (sp.moves) { B.a[1][0] = sp.center_vec[0]; B.a[1][1] = sp.center_vec[1]; B.a[1][2] = sp.center_vec[2]; }
            } else fill_quad_rows(s.quads[idx], B.a);
            out.boundary.push_back(B);
        };
        auto no_medium = [&](int, const Chain&) { F.err = "a constant_medium whose boundary contains another constant_medium is not supported"; };
        F.collect(s.media[m].obj_type, s.media[m].obj_idx, V.chain, 0, emit_b, no_medium);      // the boundary sits below the medium's own wrappers
        if (!F.err.empty()) return fail(F.err);
        M.count = (int)out.boundary.size() - M.first;
        out.media.push_back(M);
    }
    // 0: one traversal; 1: the shipped form (all media between the surfaces and the top-level lists: a second pass over the lists);
    // 2: media anywhere in the visit order (nested in wrappers / lists): one windowed pass per medium (rt_core.cuh: media_stages)
    out.two_pass = out.media.empty() ? 0 : (general ? 2 : (out.post_media_order < (int)out.leaves.size() ? 1 : 0));
    out.empty = out.leaves.empty() ? 1 : 0;

    // ---- light handle (camera.cuh:118-133, objects.cuh:947-979) ----
    auto light_prim = [&](int type, int idx) {
        LightPrim L; memset(&L, 0, sizeof(L)); L.kind = LIGHT_INVALID;
        if (type == MORT_OBJ_SPHERE && idx >= 0 && idx < (int)s.spheres.size()) {
            L.kind = LIGHT_SPHERE; const mscn_sphere& sp = s.spheres[idx]; out.sphere_pinned[idx] = 1;
            L.a[0][0] = sp.center[0]; L.a[0][1] = sp.center[1]; L.a[0][2] = sp.center[2]; L.a[0][3] = sp.radius;
        } else if (type == MORT_OBJ_QUAD && idx >= 0 && idx < (int)s.quads.size()) {
            L.kind = LIGHT_QUAD; fill_quad_rows(s.quads[idx], L.a); L.area = s.quads[idx].area; L.D = s.quads[idx].D;
        }
        return L;
    };
    const Camera& cam = s.cam;
    if (cam.light_obj_type == -1) out.light_kind = LIGHT_NONE;
    else if (cam.light_obj_type == MORT_OBJ_SPHERE || cam.light_obj_type == MORT_OBJ_QUAD) {
        LightPrim L = light_prim(cam.light_obj_type, cam.light_obj_idx);
        out.light_kind = L.kind; out.lights.push_back(L);
    } else if (cam.light_obj_type == MORT_OBJ_HITTABLE_LIST && cam.light_obj_idx >= 0 && cam.light_obj_idx < (int)s.lists.size()) {
        out.light_kind = LIGHT_LIST;
        for (const Handle& h : s.lists[cam.light_obj_idx].items) {
            if (h.type == MORT_OBJ_HITTABLE_LIST) return fail("nested lists as light handles are not supported");
            out.lights.push_back(light_prim(h.type, h.idx));
        }
        if (out.lights.empty()) return fail("light list is empty");
    } else out.light_kind = LIGHT_INVALID;   // e.g. scene 7's (4,0): pdf 0, direction (1,0,0)

    // ---- boxes, padding, SAH build ----
    std::vector<BuildPrim> prims(out.leaves.size());
    float M = 0;
    for (int a = 0; a < 3; a++) M = fmaxf(M, fabsf(cam.center[a]));
    for (size_t i = 0; i < out.leaves.size(); i++) {
        BuildPrim& p = prims[i];
        F.leaf_world_box(out.leaves[i].type, out.leaves[i].idx, out.leaves[i].inst >= 0 ? F.inst_chain[out.leaves[i].inst] : none, p.lo, p.hi);
        p.type = out.leaves[i].type; p.ref = (int)i;
        if (p.type == MORT_OBJ_SPHERE && s.spheres[out.leaves[i].idx].moves) out.n_moving++;
        for (int a = 0; a < 3; a++) {
            if (!(p.lo[a] <= p.hi[a])) return fail("a primitive has a NaN coordinate (no box bounds it)");
            M = fmax_(M, fabsf(p.lo[a])); M = fmax_(M, fabsf(p.hi[a]));
        }
        // The reference's BVH boxes stop containing an object when its bubble sort physically swapped something a wrapper
        // points at, or a list grew after it was wrapped (objects.cuh:630-661, 463-469): the reference then culls that object
        // view-dependently.  That is not reproducible without running its BVH; refuse instead of rendering something else.
        for (int a = 0; a < 3 && !s.edited; a++) {          // (an edited scene, mort_update_sphere, has left the reference's build behind)
            const float tol = 1e-4f * fmax_(1.0f, fmax_(fabsf(p.lo[a]), fabsf(p.hi[a])));
            if (p.lo[a] < gates[i].lo[a] - tol || p.hi[a] > gates[i].hi[a] + tol)
                return fail("a bvh node box of the reference's build does not contain an object below it (a wrapper's target was moved by the "
                            "build's in-place sort, or a list grew after it was wrapped): the reference culls it view-dependently; not reproducible");
        }
    }
    // The slab test runs in float with one FMA per plane; its error is a few ulp of the largest coordinate
    // in play.  Boxes are padded by 2e-6 * M so the BVH can never cull a hit the exact primitive test accepts
    // (checked against brute force in tests/test_gpu_trace.py).
    float pad = 2e-6f * M;
    for (auto& p : prims) for (int a = 0; a < 3; a++) { p.lo[a] -= pad + 2e-6f * fabsf(p.lo[a]); p.hi[a] += pad + 2e-6f * fabsf(p.hi[a]); }
    out.stats.pad = pad; out.stats.scene_extent = M; out.stats.n_leaves = (int)out.leaves.size();

    std::vector<int> order;
    if ((int)prims.size() <= MORT_LINEAR_MAX_LEAVES) {
        // a handful of primitives: no tree.  Records sorted by (instance, visit order) so consecutive records
        // share the object-space ray; the kernels scan them in lockstep.
        out.linear = 1;
        order.resize(prims.size());
        for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
        // within an instance group: largest surface first — a ray most likely hits something big, and once it has a
        // hit the remaining (smaller) candidates mostly fail the cheap t-range test before their full test
        auto area_of = [&](int i) {
            const LeafRef& L = out.leaves[i];
            if (L.type == MORT_OBJ_SPHERE) { float r = s.spheres[L.idx].radius; return 12.566371f * r * r; }
            return s.quads[L.idx].area;
        };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            if (out.leaves[a].inst != out.leaves[b].inst) return out.leaves[a].inst < out.leaves[b].inst;
            float aa = area_of(a), ab = area_of(b);
            if (aa != ab) return aa > ab;
            return out.leaves[a].order < out.leaves[b].order;
        });
        Bvh4Node n; memset(&n, 0, sizeof(n));
        for (int k = 0; k < 4; k++) { n.lox[k] = n.loy[k] = n.loz[k] = INFINITY; n.hix[k] = n.hiy[k] = n.hiz[k] = -INFINITY; n.child[k] = MORT_CHILD_EMPTY; }
        out.nodes.assign(1, n);
        out.stats.n_nodes = 1; out.stats.max_depth = 0; out.stats.level_first = {0, 1};
    } else {
        std::string berr;
        if (build) { if (!build(build_user, prims, out.nodes, order, out.stats, opt, &berr)) return fail("tree build: " + berr); }
        else build_bvh4(prims, out.nodes, order, out.stats, opt);
    }

    // ---- emit primitive records in leaf order; rewrite leaf child words to per-type record indices ----
    std::vector<int> rec_index(order.size());
    { size_t ns = 0; for (const LeafRef& L : out.leaves) ns += L.type == MORT_OBJ_SPHERE ? 1 : 0; out.spheres.reserve(ns); out.sphere_info.reserve(ns); out.quads.reserve(out.leaves.size() - ns); }
    for (size_t i = 0; i < order.size(); i++) {
        const LeafRef& L = out.leaves[order[i]];
        if (L.type == MORT_OBJ_SPHERE) {
            const mscn_sphere& sp = s.spheres[L.idx];
            SphereGeom g; g.cx = sp.center[0]; g.cy = sp.center[1]; g.cz = sp.center[2]; g.r = sp.radius;
            g.vx = sp.moves ? sp.center_vec[0] : 0.f; g.vy = sp.moves ? sp.center_vec[1] : 0.f; g.vz = sp.moves ? sp.center_vec[2] : 0.f;
            g.inst = L.inst;
            PrimInfo pi; pi.mat_gid = mat_gid(s, sp.mat_type, sp.mat_idx); pi.obj_idx = L.idx; pi.order = L.order; pi.pad = (L.top_type << 24) | (L.top_idx & 0xFFFFFF);
            rec_index[i] = (int)out.spheres.size();
            out.spheres.push_back(g); out.sphere_info.push_back(pi);
        } else {
            const mscn_quad& q = s.quads[L.idx];
            QuadRec r; memset(&r, 0, sizeof(r));
            r.nx = q.normal[0]; r.ny = q.normal[1]; r.nz = q.normal[2]; r.D = q.D;
            r.Qx = q.Q[0]; r.Qy = q.Q[1]; r.Qz = q.Q[2]; r.inst = L.inst;
            r.ux = q.u[0]; r.uy = q.u[1]; r.uz = q.u[2]; r.mat_gid = mat_gid(s, q.mat_type, q.mat_idx);
            r.vx = q.v[0]; r.vy = q.v[1]; r.vz = q.v[2]; r.order = L.order;
            r.wx = q.w[0]; r.wy = q.w[1]; r.wz = q.w[2]; r.obj_idx = L.idx;
            r.area = q.area; r.pad[0] = (L.top_type << 24) | (L.top_idx & 0xFFFFFF);
            rec_index[i] = (int)out.quads.size();
            out.quads.push_back(r);
        }
    }
    for (Bvh4Node& n : out.nodes)
        for (int k = 0; k < 4; k++) {
            uint32_t w = n.child[k];
            if (w == MORT_CHILD_EMPTY || !(w & MORT_LEAF_BIT)) continue;
            uint32_t first = w & 0x07FFFFFFu;
            n.child[k] = (w & ~0x07FFFFFFu) | (uint32_t)rec_index[first];
        }
    // ---- shade class per record (and per medium): the block wavefront's queue for a hit, decided by one byte ----
    auto tex_cold = [&](int gid) {
        int stack[64]; stack[0] = gid;
        for (int guard = 0, stack_n = 1; stack_n > 0 && guard < 64; guard++) {
            const int t = stack[--stack_n];
            if (t < 0 || t >= (int)out.textures.size()) continue;
            const Texture& T = out.textures[t];
            if (T.type == MORT_TEX_IMAGE || T.type == MORT_TEX_NOISE) return true;
            if (T.type == MORT_TEX_CHECKER && stack_n + 2 <= 64) { stack[stack_n++] = T.even_gid; stack[stack_n++] = T.odd_gid; }
        }
        return false;
    };
    auto shade_class = [&](int gid) -> uint8_t {
        if (gid < 0 || gid >= (int)out.materials.size()) return SHADE_TERMINAL;
        const Material& M = out.materials[gid];
        if (M.type == MORT_MAT_LAMBERTIAN || M.type == MORT_MAT_ISOTROPIC) return tex_cold(M.tex_gid) ? SHADE_DIFFUSE_COLD : SHADE_DIFFUSE;
        if (M.type == MORT_MAT_METAL) return SHADE_METAL;
        if (M.type == MORT_MAT_DIELECTRIC) return SHADE_DIELECTRIC;
        return SHADE_TERMINAL;                                   // diffuse_light, unknown tags
    };
    std::vector<uint8_t> cls_of(out.materials.size());
    for (size_t g = 0; g < out.materials.size(); g++) cls_of[g] = shade_class((int)g);
    auto cls = [&](int gid) -> uint8_t { return gid < 0 || gid >= (int)cls_of.size() ? (uint8_t)SHADE_TERMINAL : cls_of[gid]; };
    out.sphere_cls.resize(out.sphere_info.size()); out.quad_cls.resize(out.quads.size());
    for (size_t i = 0; i < out.sphere_info.size(); i++) out.sphere_cls[i] = cls(out.sphere_info[i].mat_gid);
    for (size_t i = 0; i < out.quads.size(); i++) out.quad_cls[i] = cls(out.quads[i].mat_gid);
    for (Medium& M : out.media) M.cls = shade_class(M.mat_gid);
    camera_params(s.cam, out.cam);
    return true;
}

}  // namespace mort
