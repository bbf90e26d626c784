// tests/hostsim/buildsim.cpp — DEVELOPMENT AID, test-only.  Runs the GPU tree builder's per-thread bodies (gpu_build_core.cuh)
// and its level loop (gpu_build_driver.hpp) serially on the CPU — positions in a shuffled order, standing in for the arbitrary
// order atomics resolve in — and compares the result with the host builder (bvh_build.cpp): same 4-wide nodes, same leaf order.
// Not linked into libmort_b200.so.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "flatten.hpp"
#include "gpu_build_driver.hpp"
#include "refit_core.cuh"
#include "scene.hpp"

using namespace mort;

struct SerialExec {
    std::vector<void*> blocks; std::mt19937 rng{12345};
    std::vector<int> perm;
    ~SerialExec() { for (void* p : blocks) free(p); }
    void* alloc(size_t b) { void* p = malloc(b); blocks.push_back(p); return p; }
    void upload(void* d, const void* s, size_t b) { memcpy(d, s, b); }
    void download(void* d, const void* s, size_t b) { memcpy(d, s, b); }
    const std::vector<int>& order(int n) { perm.resize(n); for (int i = 0; i < n; i++) perm[i] = i; std::shuffle(perm.begin(), perm.end(), rng); return perm; }
    void k_init(const gb::Ctx& c, int sv) { for (int i = 0; i < c.N; i++) { c.idx[0][i] = i; c.slot[0][i] = sv; } }
    void k_clear(const gb::Ctx& c, int n) { for (int s = 0; s < n; s++) gb::body_clear(c, s); }
    void k_stats(const gb::Ctx& c, int cur) { for (int i : order(c.N)) gb::body_stats(c, cur, i); }
    void k_bin(const gb::Ctx& c, int cur) { for (int i : order(c.N)) { const int s = c.slot[cur][i]; if (s >= 0) gb::body_bin(c, cur, i, s, c.bins + (size_t)s * gb::kBinWords); } }
    void k_split(const gb::Ctx& c, int cur, int n) { for (int s : order(n)) gb::body_split(c, cur, s); }
    void k_partition(const gb::Ctx& c, int cur) { for (int i : order(c.N)) gb::body_partition(c, cur, i); }
    void k_small(const gb::Ctx& c, int cur, int n) { for (int j : order(n)) gb::body_small(c, cur, j); }
    void k_collapse_count(const gb::Ctx& c, int cur, int n) { for (int i = 0; i < n; i++) gb::body_collapse_count(c, cur, i); }
    void k_scan(const gb::Ctx& c, int n) { int t = 0; for (int i = 0; i < n; i++) { c.ioff[i] = t; t += c.icount[i]; } c.cnt->collapse_total = t; }
    void k_collapse_emit(const gb::Ctx& c, int cur, int base, int n) { for (int i : order(n)) gb::body_collapse_emit(c, cur, i, base, n); }
    bool ok(std::string*) { return true; }
};

static int g_bad = 0, g_ksmall = 64;
static bool emu_build(void*, const std::vector<BuildPrim>& prims, std::vector<Bvh4Node>& nodes, std::vector<int>& order, BuildStats& stats, const BuildOptions& opt, std::string* err) {
    SerialExec x;
    if (!gb::build_run(x, prims, nodes, order, stats, opt, g_ksmall, err)) return false;
    std::vector<Bvh4Node> hn; std::vector<int> ho; BuildStats hs;
    build_bvh4(prims, hn, ho, hs, opt);
    bool same = hn.size() == nodes.size() && ho == order && hs.level_first == stats.level_first;
    size_t diff = 0;
    if (hn.size() == nodes.size())
        for (size_t i = 0; i < hn.size(); i++) {
            const float* a = hn[i].lox; const float* b = nodes[i].lox;
            for (int k = 0; k < 24; k++) if (!(a[k] == b[k])) diff++;
            for (int k = 0; k < 4; k++) if (hn[i].child[k] != nodes[i].child[k]) diff++;
        }
    fprintf(stderr, "prims %zu k_small %d: emulated gpu build %d nodes (bvh2 %d, levels %d, small subtrees %d, depth %d, leaf slots %d, sah %.6f) | host %d nodes (bvh2 %d, depth %d, leaf slots %d, sah %.6f) | order %s, word/box differences %zu\n",
            prims.size(), g_ksmall, stats.n_nodes, stats.n_bvh2_nodes, stats.gpu_levels, stats.gpu_small_subtrees, stats.max_depth, stats.n_leaf_slots, stats.sah_cost,
            hs.n_nodes, hs.n_bvh2_nodes, hs.max_depth, hs.n_leaf_slots, hs.sah_cost, ho == order ? "equal" : "DIFFERENT", diff);
    if (!same || diff || stats.n_leaf_slots != hs.n_leaf_slots || stats.n_bvh2_nodes != hs.n_bvh2_nodes || fabs(stats.sah_cost - hs.sah_cost) > 1e-3 * hs.sah_cost) g_bad++;
    return true;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "buildsim <scene 1-10 | field:G | text:file | dup:N> <assets> [k_small]\n"); return 1; }
    if (argc > 3) g_ksmall = atoi(argv[3]);
    Scene s; std::string a1 = argv[1], assets = argv[2];
    if (a1.rfind("field:", 0) == 0) { if (!build_sphere_field(s, atoi(a1.c_str() + 6), 69420, 0)) { fprintf(stderr, "field: %s\n", s.error.c_str()); return 2; } }
    else if (a1.rfind("text:", 0) == 0) { HostRng g(1); std::string err; if (!load_scene_text(s, g, a1.substr(5), assets, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 2; } }
    else if (a1.rfind("rand:", 0) == 0) {
        // random boxes: clustered centres, several scales, both types, exact duplicates, flat and point-like boxes
        int n = 0, seed = 0; sscanf(a1.c_str() + 5, "%d:%d", &n, &seed);
        std::mt19937 g((unsigned)seed); std::uniform_real_distribution<float> U(0.f, 1.f);
        std::vector<BuildPrim> prims;
        const int n_clusters = 1 + (int)(U(g) * 12);
        std::vector<float> cc(3 * n_clusters); for (float& v : cc) v = (U(g) - 0.5f) * 200.f;
        for (int i = 0; i < n; i++) {
            BuildPrim p; const int c = (int)(U(g) * n_clusters) % n_clusters; const float spread = U(g) < 0.2f ? 60.f : 4.f;
            const float r[3] = {U(g) < 0.1f ? 0.f : U(g) * (U(g) < 0.05f ? 30.f : 0.8f), U(g) * 0.8f, U(g) < 0.1f ? 0.f : U(g) * 0.8f};
            for (int a = 0; a < 3; a++) { const float m = cc[3 * c + a] + (U(g) - 0.5f) * spread; p.lo[a] = m - r[a]; p.hi[a] = m + r[a]; }
            p.type = U(g) < 0.4f ? MORT_OBJ_QUAD : MORT_OBJ_SPHERE; p.ref = i;
            if (i > 0 && U(g) < 0.05f) { const BuildPrim& q = prims[(size_t)(U(g) * i) % i]; memcpy(p.lo, q.lo, 12); memcpy(p.hi, q.hi, 12); if (U(g) < 0.5f) p.type = q.type; }
            prims.push_back(p);
        }
        std::vector<Bvh4Node> nodes; std::vector<int> order; BuildStats st; std::string err;
        if (!emu_build(nullptr, prims, nodes, order, st, BuildOptions(), &err)) { fprintf(stderr, "%s\n", err.c_str()); return 3; }
        return g_bad ? 4 : 0;
    }
    else if (a1.rfind("dup:", 0) == 0) {
        // degenerate input: N concentric spheres + N coincident quads (every centroid equal) + a few ordinary ones
        const int n = atoi(a1.c_str() + 4);
        std::vector<BuildPrim> prims;
        for (int i = 0; i < 2 * n; i++) { BuildPrim p; const float r = 1.f + (i % n) * 0.25f; for (int a = 0; a < 3; a++) { p.lo[a] = -r; p.hi[a] = r; } p.type = i < n ? MORT_OBJ_SPHERE : MORT_OBJ_QUAD; p.ref = i; prims.push_back(p); }
        for (int i = 0; i < 7; i++) { BuildPrim p; for (int a = 0; a < 3; a++) { p.lo[a] = 3.f * i + a; p.hi[a] = p.lo[a] + 1.f; } p.type = MORT_OBJ_SPHERE; p.ref = (int)prims.size(); prims.push_back(p); }
        std::vector<Bvh4Node> nodes; std::vector<int> order; BuildStats st; std::string err;
        if (!emu_build(nullptr, prims, nodes, order, st, BuildOptions(), &err)) { fprintf(stderr, "%s\n", err.c_str()); return 3; }
        return g_bad ? 4 : 0;
    }
    else if (!build_reference_scene(s, atoi(argv[1]), assets)) { fprintf(stderr, "scene: %s\n", s.error.c_str()); return 2; }
    FlatScene f; std::string err;
    if (!flatten_scene(s, f, &err, BuildOptions(), emu_build, nullptr)) { fprintf(stderr, "flatten: %s\n", err.c_str()); return 2; }
    if (f.linear) { fprintf(stderr, "linear-scan scene: no tree\n"); return g_bad ? 4 : 0; }
    // ---- the bottom-up bounds pass (refit_core.cuh) run serially: (1) on the unchanged scene it must reproduce the builder's
    // boxes bit for bit, (2) the motion boxes interpolated at any time must contain the primitives at that time and nest ----
    {
        rf::Ctx c; memset(&c, 0, sizeof(c));
        std::vector<Bvh4Node> nodes = f.nodes, t1(f.nodes.size());
        for (Bvh4Node& n : nodes) for (int k = 0; k < 4; k++) if (n.child[k] != MORT_CHILD_EMPTY) n.lox[k] = n.loy[k] = n.loz[k] = n.hix[k] = n.hiy[k] = n.hiz[k] = 12345.f;
        std::vector<rf::Box> sb[2], qb[2];
        for (int t = 0; t < 2; t++) { sb[t].resize(f.spheres.size()); qb[t].resize(f.quads.size()); c.sphere_box[t] = sb[t].data(); c.quad_box[t] = qb[t].data(); }
        unsigned ext = 0; c.extent_key = &ext;
        c.nodes = nodes.data(); c.node_t1 = nullptr; c.spheres = f.spheres.data(); c.quads = f.quads.data(); c.instances = f.instances.data();
        c.n_spheres = (int)f.spheres.size(); c.n_quads = (int)f.quads.size();
        const int n = c.n_spheres + c.n_quads;
        auto run = [&](bool motion) {
            ext = 0;
            for (int i = 0; i < n; i++) rf::body_raw(c, i, motion);
            float M; memcpy(&M, &ext, 4);
            M = fmaxf(M, fmaxf(fabsf(s.cam.center.x), fmaxf(fabsf(s.cam.center.y), fabsf(s.cam.center.z))));
            c.pad = 2e-6f * M;
            for (int i = 0; i < n; i++) rf::body_pad(c, i, motion);
            for (int L = (int)f.stats.level_first.size() - 2; L >= 0; L--)
                for (int node = f.stats.level_first[L]; node < f.stats.level_first[L + 1]; node++) for (int k = 0; k < 4; k++) rf::body_level(c, node, k, motion);
        };
        run(false);
        size_t diff = 0;
        for (size_t i = 0; i < nodes.size(); i++) { const float* a = nodes[i].lox; const float* b = f.nodes[i].lox; for (int k = 0; k < 24; k++) if (!(a[k] == b[k])) diff++; }
        fprintf(stderr, "refit of the unchanged scene: %zu box words differ from the builder's (pad %g vs %g)\n", diff, c.pad, f.stats.pad);
        if (diff) g_bad++;
        // (1b) after an edit (what mort_update_sphere + mort_refit do): every 7th sphere record moves and changes size; the refitted
        // boxes must contain their primitives and nest.  Done on copies: the motion check below wants the unchanged scene.
        {
            std::vector<SphereGeom> moved = f.spheres; std::vector<Bvh4Node> en = f.nodes;
            std::mt19937 g(7); std::uniform_real_distribution<float> U(-1.f, 1.f);
            for (size_t i = 0; i < moved.size(); i += 7) { moved[i].cx += 3.f * U(g); moved[i].cy += fabsf(U(g)); moved[i].cz += 3.f * U(g); moved[i].r = 0.1f + 0.5f * fabsf(U(g)); moved[i].vy = i % 14 ? 0.f : 0.4f; }
            rf::Ctx e = c; e.spheres = moved.data(); e.nodes = en.data(); e.node_t1 = nullptr;
            ext = 0;
            for (int i = 0; i < n; i++) rf::body_raw(e, i, false);
            float M; memcpy(&M, &ext, 4);
            M = fmaxf(M, fmaxf(fabsf(s.cam.center.x), fmaxf(fabsf(s.cam.center.y), fabsf(s.cam.center.z))));
            e.pad = 2e-6f * M;
            for (int i = 0; i < n; i++) rf::body_pad(e, i, false);
            for (int L = (int)f.stats.level_first.size() - 2; L >= 0; L--)
                for (int node = f.stats.level_first[L]; node < f.stats.level_first[L + 1]; node++) for (int k = 0; k < 4; k++) rf::body_level(e, node, k, false);
            size_t bad_edit = 0, chk = 0;
            for (size_t i = 0; i < en.size(); i++) for (int k = 0; k < 4; k++) {
                const uint32_t w = en[i].child[k];
                if (w == MORT_CHILD_EMPTY) continue;
                const float lo[3] = {en[i].lox[k], en[i].loy[k], en[i].loz[k]}, hi[3] = {en[i].hix[k], en[i].hiy[k], en[i].hiz[k]};
                if (w & MORT_LEAF_BIT) {
                    const int first = (int)(w & 0x07FFFFFFu), cnt = (int)((w >> 27) & 7u) + 1;
                    for (int j = 0; j < cnt; j++) {
                        const rf::Box b = (w & MORT_LEAF_QUAD_BIT) ? rf::quad_raw_box(e, first + j) : rf::sphere_raw_box(e, first + j, rf::TIME_UNION);
                        for (int a = 0; a < 3; a++) { chk++; if (b.lo[a] < lo[a] || b.hi[a] > hi[a]) bad_edit++; }
                    }
                } else for (int j = 0; j < 4; j++) {
                    if (en[w].child[j] == MORT_CHILD_EMPTY) continue;
                    const float clo[3] = {en[w].lox[j], en[w].loy[j], en[w].loz[j]}, chi[3] = {en[w].hix[j], en[w].hiy[j], en[w].hiz[j]};
                    for (int a = 0; a < 3; a++) { chk++; if (clo[a] < lo[a] || chi[a] > hi[a]) bad_edit++; }
                }
            }
            fprintf(stderr, "refit after moving %zu spheres: %zu containment checks, %zu edit-violations\n", (moved.size() + 6) / 7, chk, bad_edit);
            if (bad_edit) g_bad++;
            ext = 0;
        }
        c.node_t1 = t1.data();
        run(true);
        std::vector<Bvh4Node> t1abs = t1;
        for (size_t i = 0; i < nodes.size(); i++) for (int k = 0; k < 4; k++) rf::body_delta(c, (int)i, k);
        size_t viol = 0, checks = 0; int moving = 0;
        for (const SphereGeom& g : f.spheres) if (g.vx != 0.f || g.vy != 0.f || g.vz != 0.f) moving++;
        const float times[5] = {0.f, 0.25f, 0.5f, 0.8125f, 0.99999994f};
        for (size_t i = 0; i < nodes.size(); i++) for (int k = 0; k < 4; k++) {
            const uint32_t w = nodes[i].child[k];
            if (w == MORT_CHILD_EMPTY) { if (t1[i].lox[k] != 0.f || t1[i].hiz[k] != 0.f) viol++; continue; }
            for (float tm : times) {
                float lo[3] = {fmaf(tm, t1[i].lox[k], nodes[i].lox[k]), fmaf(tm, t1[i].loy[k], nodes[i].loy[k]), fmaf(tm, t1[i].loz[k], nodes[i].loz[k])};
                float hi[3] = {fmaf(tm, t1[i].hix[k], nodes[i].hix[k]), fmaf(tm, t1[i].hiy[k], nodes[i].hiy[k]), fmaf(tm, t1[i].hiz[k], nodes[i].hiz[k])};
                // inside the union box the builder made (up to rounding), ...
                const float ulo[3] = {f.nodes[i].lox[k], f.nodes[i].loy[k], f.nodes[i].loz[k]}, uhi[3] = {f.nodes[i].hix[k], f.nodes[i].hiy[k], f.nodes[i].hiz[k]};
                for (int a = 0; a < 3; a++) { checks++; if (lo[a] < ulo[a] - 1e-5f * fabsf(ulo[a]) - 1e-6f || hi[a] > uhi[a] + 1e-5f * fabsf(uhi[a]) + 1e-6f) viol++; }
                if (w & MORT_LEAF_BIT) {          // ... and around every primitive at that time (unpadded box, so the pad is the margin)
                    const int first = (int)(w & 0x07FFFFFFu), cnt = (int)((w >> 27) & 7u) + 1;
                    for (int j = 0; j < cnt; j++) {
                        rf::Box b;
                        if (w & MORT_LEAF_QUAD_BIT) b = rf::quad_raw_box(c, first + j);
                        else {
                            const SphereGeom g = f.spheres[first + j];
                            float p[3] = {g.cx + tm * g.vx, g.cy + tm * g.vy, g.cz + tm * g.vz};
                            rf::to_world(f.instances.data(), g.inst, p);
                            rf::box_clear(b); rf::grow_point(b, p, fabsf(g.r));
                        }
                        for (int a = 0; a < 3; a++) { checks++; if (b.lo[a] < lo[a] || b.hi[a] > hi[a]) viol++; }
                    }
                } else {
                    for (int j = 0; j < 4; j++) {
                        if (nodes[w].child[j] == MORT_CHILD_EMPTY) continue;
                        const float clo[3] = {fmaf(tm, t1[w].lox[j], nodes[w].lox[j]), fmaf(tm, t1[w].loy[j], nodes[w].loy[j]), fmaf(tm, t1[w].loz[j], nodes[w].loz[j])};
                        const float chi[3] = {fmaf(tm, t1[w].hix[j], nodes[w].hix[j]), fmaf(tm, t1[w].hiy[j], nodes[w].hiy[j]), fmaf(tm, t1[w].hiz[j], nodes[w].hiz[j])};
                        for (int a = 0; a < 3; a++) { checks++; if (clo[a] < lo[a] - 4e-7f * fabsf(lo[a]) || chi[a] > hi[a] + 4e-7f * fabsf(hi[a])) viol++; }
                    }
                }
            }
        }
        // how much tighter: surface of the boxes at time 0.5 against the union boxes, summed over all children
        double au = 0, am = 0, rsum = 0; size_t rn = 0;
        for (size_t i = 0; i < nodes.size(); i++) for (int k = 0; k < 4; k++) {
            if (nodes[i].child[k] == MORT_CHILD_EMPTY) continue;
            const float ulo[3] = {f.nodes[i].lox[k], f.nodes[i].loy[k], f.nodes[i].loz[k]}, uhi[3] = {f.nodes[i].hix[k], f.nodes[i].hiy[k], f.nodes[i].hiz[k]};
            const float lo[3] = {nodes[i].lox[k] + 0.5f * t1[i].lox[k], nodes[i].loy[k] + 0.5f * t1[i].loy[k], nodes[i].loz[k] + 0.5f * t1[i].loz[k]};
            const float hi[3] = {nodes[i].hix[k] + 0.5f * t1[i].hix[k], nodes[i].hiy[k] + 0.5f * t1[i].hiy[k], nodes[i].hiz[k] + 0.5f * t1[i].hiz[k]};
            au += sah_area(ulo, uhi); am += sah_area(lo, hi);
            if (sah_area(ulo, uhi) > 0) { rsum += sah_area(lo, hi) / sah_area(ulo, uhi); rn++; }
        }
        fprintf(stderr, "motion boxes: %d moving spheres, %zu containment checks, %zu violations; child-box surface at t = 0.5 is %.1f %% of the union boxes' in total, %.1f %% on average per child\n", moving, checks, viol, 100.0 * am / au, 100.0 * rsum / (rn ? rn : 1));
        if (viol) g_bad++;
    }
    return g_bad ? 4 : 0;
}
