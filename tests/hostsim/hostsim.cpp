// tests/hostsim/hostsim.cpp — DEVELOPMENT AID, test-only.  Compiles mort_b200/csrc/rt_core.cuh as plain
// C++ and runs the product's flattening + BVH traversal + shading functions serially on the CPU so they
// can be diffed against the oracle on a machine without a GPU.  It is not linked into libmort_b200.so and
// no product entry point can reach it; the product itself has no CPU path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define MORT_HOST_COUNTERS
#include "flatten.hpp"
#include "rt_core.cuh"
#include "scene.hpp"

using namespace mort;

static DeviceScene make_device_scene(const Scene& s, const FlatScene& f, std::vector<ImageDesc>& imgs) {
    DeviceScene d; memset(&d, 0, sizeof(d));
    d.nodes = f.nodes.data(); d.n_nodes = (int)f.nodes.size();
    d.spheres = f.spheres.data(); d.sphere_info = f.sphere_info.data(); d.n_spheres = (int)f.spheres.size();
    d.quads = f.quads.data(); d.n_quads = (int)f.quads.size();
    d.sphere_cls = f.sphere_cls.data(); d.quad_cls = f.quad_cls.data();
    d.instances = f.instances.data(); d.n_instances = (int)f.instances.size();
    d.materials = f.materials.data(); d.n_materials = (int)f.materials.size();
    d.textures = f.textures.data(); d.n_textures = (int)f.textures.size();
    imgs.clear();
    for (const ImageRec& im : s.images) { ImageDesc I; I.texels = im.rgb.empty() ? nullptr : im.rgb.data(); I.width = im.width; I.height = im.height; I.cols = im.width * 3; imgs.push_back(I); }
    d.images = imgs.data(); d.n_images = (int)imgs.size();
    d.noises = f.noises.data(); d.n_noises = (int)f.noises.size();
    d.media = f.media.data(); d.n_media = (int)f.media.size();
    d.n_media_top = 0; for (const Medium& M : f.media) d.n_media_top += M.top_level ? 1 : 0;
    d.boundary = f.boundary.data(); d.n_boundary = (int)f.boundary.size();
    d.lights = f.lights.data(); d.n_lights = (int)f.lights.size(); d.light_kind = f.light_kind;
    d.post_media_order = f.post_media_order; d.two_pass = f.two_pass; d.empty = f.empty; d.linear = f.linear;
    return d;
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "hostsim <scene 1-10 | file.mscn> <assets> trace in.mhit out.mhit [brute]\n        ... render W SPP DEPTH SEED out.mimg\n"); return 1; }
    Scene s; std::string assets = argv[2];
    std::string a1 = argv[1];
    if (a1.rfind("text:", 0) == 0) {
        HostRng g(1); std::string err;
        if (!load_scene_text(s, g, a1.substr(5), assets, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
    } else if (a1.rfind("field:", 0) == 0) {
        if (!build_sphere_field(s, atoi(a1.c_str() + 6), 69420, 0)) { fprintf(stderr, "field: %s\n", s.error.c_str()); return 2; }
    } else if (a1.size() > 5 && a1.substr(a1.size() - 5) == ".mscn") {
        std::string err; if (!s.load(a1, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
        ImageRec im; if (!s.images.empty() && load_ppm(assets + "/earthmap.ppm", im)) { s.images[0].rgb = im.rgb; }
    } else if (!build_reference_scene(s, atoi(argv[1]), assets)) { fprintf(stderr, "scene: %s\n", s.error.c_str()); return 2; }
    std::string mode = argv[3];
    if (const char* e = getenv("MORT_EDIT_SPHERES")) {
        // what mort_update_sphere does to the host scene (Scene::update_sphere), for N pseudo-random spheres: new centres, every
        // second one moving, new radii — far outside the boxes the reference's own bvh holds for them
        const int n = atoi(e); uint64_t st = 12345;
        auto u = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (float)((st >> 40) & 0xFFFFFF) / 16777216.f; };
        for (int k = 0; k < n && !s.spheres.empty(); k++) {
            const int idx = (int)(u() * s.spheres.size()) % (int)s.spheres.size();
            const V3 c0((u() - 0.5f) * 20.f, u() * 3.f, (u() - 0.5f) * 20.f), c1 = c0 + V3(0, u(), 0);
            s.update_sphere(idx, c0, (k & 1) ? &c1 : nullptr, 0.1f + u() * 0.6f);
        }
    }
    if (mode == "dump") { return s.dump(argv[4]) ? 0 : 3; }
    if (mode == "dumptext") { std::string err; if (!dump_scene_text(s, argv[4], &err)) { fprintf(stderr, "%s\n", err.c_str()); return 3; } return 0; }
    if (mode == "rand") { HostRng g(1); int n = atoi(argv[4]); for (int i = 0; i < n; i++) fprintf(stderr, "%d\n", g.next()); return 0; }
    if (mode == "render") {
        int W = atoi(argv[4]), spp = atoi(argv[5]), depth = atoi(argv[6]);
        if (W > 0) s.cam.image_width = W;
        if (spp > 0) s.cam.samples_per_pixel = spp;
        if (depth > 0) s.cam.bounce_limit = depth;
        s.cam.initialize();
    }
    FlatScene f; std::string err;
    BuildOptions bo;                                   // (a test harness may read its environment; the product does not)
    if (const char* e = getenv("MORT_BUILD_THREADS")) bo.threads = atoi(e);
    if (!flatten_scene(s, f, &err, bo)) { fprintf(stderr, "flatten: %s\n", err.c_str()); return 2; }
    std::vector<ImageDesc> imgs;
    DeviceScene d = make_device_scene(s, f, imgs);
    fprintf(stderr, "leaves %d nodes %d (bvh2 %d) depth %d sah %.2f pad %g build %.2f ms two_pass %d media %d lights %d kind %d\n", f.stats.n_leaves, f.stats.n_nodes,
            f.stats.n_bvh2_nodes, f.stats.max_depth, f.stats.sah_cost, f.stats.pad, f.stats.build_ms, f.two_pass, d.n_media, d.n_lights, d.light_kind);
    if (mode == "checkbvh") {
        {   // FNV-1a over the node array and the primitive records: identifies the tree and the leaf order
            uint64_t h = 1469598103934665603ull;
            auto mix = [&h](const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } };
            if (!f.nodes.empty()) mix(f.nodes.data(), f.nodes.size() * sizeof(f.nodes[0]));
            if (!f.spheres.empty()) mix(f.spheres.data(), f.spheres.size() * sizeof(f.spheres[0]));
            if (!f.quads.empty()) mix(f.quads.data(), f.quads.size() * sizeof(f.quads[0]));
            fprintf(stderr, "tree hash %016llx\n", (unsigned long long)h);
        }
        // structural invariants of the 4-wide BVH: each primitive record referenced exactly once, leaves
        // type-homogeneous by construction of the child word, un-instanced primitives inside their leaf box,
        // children boxes inside the box their parent holds for them
        std::vector<int> seen_s(f.spheres.size(), 0), seen_q(f.quads.size(), 0);
        int bad = 0; size_t visited = 0;
        struct It { uint32_t node; float lo[3], hi[3]; bool has; };
        std::vector<It> st; It root; root.node = 0; root.has = false; st.push_back(root);
        while (!st.empty()) {
            It it = st.back(); st.pop_back(); visited++;
            const Bvh4Node& n = f.nodes[it.node];
            for (int k = 0; k < 4; k++) {
                uint32_t w = n.child[k];
                if (w == MORT_CHILD_EMPTY) continue;
                float lo[3] = {n.lox[k], n.loy[k], n.loz[k]}, hi[3] = {n.hix[k], n.hiy[k], n.hiz[k]};
                if (it.has) for (int a = 0; a < 3; a++) if (lo[a] < it.lo[a] - 1e-3f || hi[a] > it.hi[a] + 1e-3f) bad++;
                if (w & MORT_LEAF_BIT) {
                    uint32_t first = w & 0x07FFFFFFu; int cnt = (int)((w >> 27) & 7u) + 1;
                    for (int i = 0; i < cnt; i++) {
                        if (w & MORT_LEAF_QUAD_BIT) {
                            if (first + i >= f.quads.size()) { bad++; continue; }
                            seen_q[first + i]++;
                            const QuadRec& q = f.quads[first + i];
                            if (q.inst < 0) for (int c = 0; c < 4; c++) {
                                float p[3] = {q.Qx + (c & 1) * q.ux + (c >> 1) * q.vx, q.Qy + (c & 1) * q.uy + (c >> 1) * q.vy, q.Qz + (c & 1) * q.uz + (c >> 1) * q.vz};
                                for (int a = 0; a < 3; a++) if (p[a] < lo[a] || p[a] > hi[a]) bad++;
                            }
                        } else {
                            if (first + i >= f.spheres.size()) { bad++; continue; }
                            seen_s[first + i]++;
                            const SphereGeom& g = f.spheres[first + i];
                            if (g.inst < 0) {
                                float c[3] = {g.cx, g.cy, g.cz}, v[3] = {g.vx, g.vy, g.vz};
                                for (int a = 0; a < 3; a++) {
                                    float mn = fminf(c[a], c[a] + v[a]) - fabsf(g.r), mx = fmaxf(c[a], c[a] + v[a]) + fabsf(g.r);
                                    if (mn < lo[a] || mx > hi[a]) bad++;
                                }
                            }
                        }
                    }
                } else {
                    if (w >= f.nodes.size()) { bad++; continue; }
                    It c; c.node = w; c.has = true; memcpy(c.lo, lo, 12); memcpy(c.hi, hi, 12); st.push_back(c);
                }
            }
        }
        for (int v : seen_s) if (v != 1) bad++;
        for (int v : seen_q) if (v != 1) bad++;
        if (visited != f.nodes.size()) bad++;
        fprintf(stderr, bad ? "bvh BAD: %d violations\n" : "bvh ok (%d)\n", bad);
        return bad ? 4 : 0;
    }
    if (mode == "trace") {
        FILE* fi = fopen(argv[4], "rb"); if (!fi) return 3;
        uint32_t hd[4]; if (fread(hd, 4, 4, fi) != 4) return 3;
        int n = (int)hd[1];
        std::vector<float> rays((size_t)n * 7); if (fread(rays.data(), 4, rays.size(), fi) != rays.size()) return 3; fclose(fi);
        bool brute = argc > 6 && !strcmp(argv[6], "brute");
        std::vector<mhit_record> out(n);
        std::vector<mhit_medium_probe> probes((size_t)n * d.n_media_top);      // like trace_kernel (render.cu): both boundary probes of every kept medium
        for (int i = 0; i < n; i++) {
            const float* q = &rays[7 * (size_t)i];
            Ray r; r.o = mk3(q[0], q[1], q[2]); r.d = mk3(q[3], q[4], q[5]); r.tm = q[6];
            for (int m = 0, j = 0; m < d.n_media; m++) {
                if (!d.media[m].top_level) continue;
                mhit_medium_probe pr; pr.hit1 = pr.hit2 = 0; pr.t1 = pr.t2 = 0.f;
                float t1, t2;
                if (boundary_probe(d, d.media[m], r, -INFINITY, INFINITY, t1)) {
                    pr.hit1 = 1; pr.t1 = t1;
                    if (boundary_probe(d, d.media[m], r, (float)((double)t1 + 0.0001), INFINITY, t2)) { pr.hit2 = 1; pr.t2 = t2; }
                }
                probes[(size_t)i * d.n_media_top + j] = pr; j++;
            }
            Hit h; bool any = brute ? closest_hit_brute(d, r, 0.001f, INFINITY, h) : closest_hit<false>(d, nullptr, 0, r, 0.001f, INFINITY, h);
            mhit_record o; memset(&o, 0, sizeof(o)); o.hit = any; o.leaf_type = o.leaf_idx = o.top_type = o.top_idx = -1;
            if (any) {
                Record rec; resolve_hit(d, r, h, rec);
                if (rec.sphere_uv) sphere_uv(rec.outward, rec.u, rec.v);
                o.t = rec.t; o.leaf_type = rec.leaf_type; o.leaf_idx = rec.leaf_idx; o.front_face = rec.front_face;
                int gid = rec.mat_gid; o.mat_type = gid < 0 ? -1 : f.materials[gid].type;
                int off = 0; for (int g = 0; g < gid; g++) if (f.materials[g].type != o.mat_type) off = g + 1;
                o.mat_idx = gid - off;
                o.p[0] = rec.p.x; o.p[1] = rec.p.y; o.p[2] = rec.p.z; o.normal[0] = rec.normal.x; o.normal[1] = rec.normal.y; o.normal[2] = rec.normal.z;
                o.u = rec.u; o.v = rec.v;
            }
            out[i] = o;
        }
        FILE* fo = fopen(argv[5], "wb"); uint32_t oh[4] = {MHIT_MAGIC, (uint32_t)n, (uint32_t)d.n_media_top, (uint32_t)f.stats.n_leaves};
        fwrite(oh, 4, 4, fo); fwrite(rays.data(), 4, rays.size(), fo); fwrite(out.data(), sizeof(mhit_record), n, fo);
        if (!probes.empty()) fwrite(probes.data(), sizeof(mhit_medium_probe), probes.size(), fo);
        fclose(fo);
    } else if (mode == "render") {
        uint32_t seed = (uint32_t)strtoul(argv[7], 0, 10);
        const CameraParams& cam = f.cam;
        std::vector<float> hdr((size_t)cam.width * cam.height * 4, 0.f);
        unsigned long long segs = 0;
        for (int pix = 0; pix < cam.width * cam.height; pix++) {
            float sx = 0, sy = 0, sz = 0; int nan_n = 0;
            for (int sj = 0; sj < cam.sqrt_spp; sj++) for (int si = 0; si < cam.sqrt_spp; si++) {
                Path P; Rng g; f3 col;
                path_start(cam, seed, 0, pix, si, sj, P, g);
                static int dbgpix = getenv("HOSTSIM_DEBUG_PIXEL") ? atoi(getenv("HOSTSIM_DEBUG_PIXEL")) : -1;
                if (pix == dbgpix) {
                    Path Q = P; Rng h = g; f3 c2;
                    for (int seg = 0;; seg++) {
                        if (path_exhausted(cam, Q, c2) || ray_is_nan(Q.ray)) break;
                        R4 sb = rng_block(h); SegHit sh; segment_trace<false>(d, nullptr, 0, Q.ray, h, sh);
                        if (sh.h.prim == MORT_PRIM_NONE) { fprintf(stderr, "HOSTSIM px %d smp %d seg %d: miss\n", pix, sj * cam.sqrt_spp + si, seg); break; }
                        Record rec; segment_record(d, Q.ray, sh, rec);
                        fprintf(stderr, "HOSTSIM px %d smp %d seg %d: hit type %d idx %d t %.9g matgid %d ff %d p %.9g %.9g %.9g d %.9g %.9g %.9g\n", pix, sj * cam.sqrt_spp + si, seg,
                                rec.leaf_type, rec.leaf_idx, rec.t, rec.mat_gid, (int)rec.front_face, rec.p.x, rec.p.y, rec.p.z, Q.ray.d.x, Q.ray.d.y, Q.ray.d.z);
                        int stt = segment_shade<CLASS_ANY>(d, cam, sh, Q, h, sb, c2);
                        fprintf(stderr, "HOSTSIM   thr %.9g %.9g %.9g\n", Q.thr.x, Q.thr.y, Q.thr.z);
                        if (stt == SEG_DONE) { fprintf(stderr, "HOSTSIM   color %.9g %.9g %.9g\n", c2.x, c2.y, c2.z); break; }
                    }
                }
                for (;;) { bool tr; int st = path_segment<false>(d, cam, nullptr, 0, P, g, col, tr); segs += tr; if (st == SEG_DONE) break; }
                if (isnan3(col)) nan_n++; else { sx += col.x; sy += col.y; sz += col.z; }
            }
            hdr[4 * (size_t)pix] = sx; hdr[4 * (size_t)pix + 1] = sy; hdr[4 * (size_t)pix + 2] = sz; hdr[4 * (size_t)pix + 3] = (float)nan_n;
        }
        fprintf(stderr, "segments %llu\n", segs);
        if (g_host_counters.queries)
            fprintf(stderr, "tree queries %llu node_steps/query %.3f leaf_visits/query %.3f prim_tests/query %.3f\n", g_host_counters.queries,
                    (double)g_host_counters.node_steps / g_host_counters.queries, (double)g_host_counters.leaf_visits / g_host_counters.queries,
                    (double)g_host_counters.prim_tests / g_host_counters.queries);
        if (g_host_counters.pushes) {
            fprintf(stderr, "stack pushes/query %.3f; share of pushes at depth >= k:", (double)g_host_counters.pushes / g_host_counters.queries);
            unsigned long long above = g_host_counters.pushes;
            for (int k = 0; k < 16; k++) { fprintf(stderr, " %d:%.3f", k, (double)above / g_host_counters.pushes); above -= g_host_counters.push_at[k]; }
            fprintf(stderr, "\n");
        }
        FILE* fo = fopen(argv[8], "wb"); uint32_t hd[5] = {0x474D494Du, (uint32_t)cam.width, (uint32_t)cam.height, 4, 1};
        fwrite(hd, 4, 5, fo); fwrite(hdr.data(), 4, hdr.size(), fo); fclose(fo);
    }
    return 0;
}
