// capi.cu — the C ABI of include/mort_b200.h: context, scene building, commit (flatten + SAH build +
// upload), render / trace launches, statistics.  No CPU rendering path exists: every entry that produces
// pixels or hits requires the CUDA device the context was created on.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "flatten.hpp"
#include "gpu_build.hpp"
#include "mort_b200.h"
#include "refit.hpp"
#include "render.hpp"
#include "scene.hpp"

using namespace mort;


#include "ctx.hpp"

static Handle H(mort_handle h) { return Handle{h.type, h.idx}; }
static void put(mort_handle* out, Handle h) { if (out) { out->type = h.type; out->idx = h.idx; } }
static V3 v3(const float* p) { return V3(p[0], p[1], p[2]); }
static void invalidate(mort_ctx* ctx) { ctx->committed = false; }
// The tree's boxes are padded for slab-test rounding by 2e-6 * M, M = the largest coordinate in play INCLUDING the camera
// centre at commit time (flatten.cpp).  A camera moved outside that radius keeps the geometry but needs wider pads: rebuild.
static bool camera_outgrew_pad(const mort_ctx* ctx) {
    const Camera& c = ctx->scene.cam;
    const float m = std::max(std::fabs(c.center.x), std::max(std::fabs(c.center.y), std::fabs(c.center.z)));
    return m > ctx->flat.stats.scene_extent;
}

// The reference sets cudaLimitStackSize itself (8192 B, mort.cu:703).  The megakernel's frame is < 1 KB, so the default
// limit would do — but re-sizing the context's local-memory pool is measurably worth it: inside a process whose CUDA context
// was created by a host framework (torch), the Cornell frame takes 242.7 ms with the pool as found and 234.5 ms after the
// limit was changed to 2048 B (profiles/r01_process_probe.jsonl; bench.py 1516 -> 1566 Msamples/s, profiles/r01_stack_limit_ab.jsonl); the spilled registers
// of 888 resident blocks live there.
// A larger limit set by the host is never lowered.
static void raise_stack_limit() {
    size_t cur = 0;
    if (cudaDeviceGetLimit(&cur, cudaLimitStackSize) != cudaSuccess) { cudaGetLastError(); return; }
    if (cur < 2048 && cudaDeviceSetLimit(cudaLimitStackSize, 2048) != cudaSuccess) cudaGetLastError();
}

extern "C" {

int mort_create(int cuda_device, mort_ctx** out) {
    if (!out) return MORT_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || cuda_device < 0 || cuda_device >= n) return MORT_ERR_CUDA;   // no CPU fallback
    if (cudaSetDevice(cuda_device) != cudaSuccess) return MORT_ERR_CUDA;
    mort_ctx* ctx = new mort_ctx();
    ctx->device = cuda_device;
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    memset(&ctx->dscene, 0, sizeof(ctx->dscene));
    if (cudaGetDeviceProperties(&ctx->prop, cuda_device) != cudaSuccess || cudaStreamCreate(&ctx->own_stream) != cudaSuccess ||     // a BLOCKING stream: ordered after work the host queued on the legacy default stream
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaMalloc(&ctx->d_counters, 4 * sizeof(unsigned long long)) != cudaSuccess || cudaMalloc(&ctx->d_work, sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc(&ctx->d_mat_offsets, 8 * sizeof(int32_t)) != cudaSuccess || cudaMalloc(&ctx->d_work64, sizeof(unsigned long long)) != cudaSuccess) {
        delete ctx; return MORT_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    raise_stack_limit();                                   // (when to do it was an experiment: profiles/r01_stack_limit_ab.jsonl)
    *out = ctx;
    return MORT_OK;
}

int mort_destroy(mort_ctx* ctx) {
    CTX_CHECK(ctx);
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->comm) mort_comm_detach(ctx);
    ctx->arena.release();
    wavefront_free(ctx->wave);
    cudaFree(ctx->d_counters); cudaFree(ctx->d_work); cudaFree(ctx->d_mat_offsets); cudaFree(ctx->d_accum); cudaFree(ctx->d_rgba); cudaFree(ctx->d_prog); cudaFree(ctx->d_pool_exact); cudaFree(ctx->d_work64); cudaFree(ctx->d_trace);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return MORT_OK;
}

const char* mort_last_error(const mort_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int mort_set_stream(mort_ctx* ctx, void* s) { CTX_CHECK(ctx); ctx->stream = s ? reinterpret_cast<cudaStream_t>(s) : ctx->own_stream; return MORT_OK; }

// ---- scenes ---------------------------------------------------------------------------------------------
int mort_build_scene(mort_ctx* ctx, int scene_id, const char* asset_dir) {
    CTX_CHECK(ctx); invalidate(ctx);
    if (!build_reference_scene(ctx->scene, scene_id, asset_dir ? asset_dir : ".")) return fail(ctx, MORT_ERR_SCENE, ctx->scene.error);
    return MORT_OK;
}
int mort_build_sphere_field(mort_ctx* ctx, int G, uint64_t seed, int camera_kind) {
    CTX_CHECK(ctx); invalidate(ctx);
    if (!build_sphere_field(ctx->scene, G, seed, camera_kind)) return fail(ctx, MORT_ERR_SCENE, ctx->scene.error);
    return MORT_OK;
}
int mort_load_scene(mort_ctx* ctx, const char* path, const char* asset_dir) {
    CTX_CHECK(ctx && path); invalidate(ctx);
    std::string e;
    if (!ctx->scene.load(path, &e)) return fail(ctx, MORT_ERR_IO, e);
    if (!ctx->scene.images.empty()) {
        ImageRec im;
        if (!asset_dir || !load_ppm(std::string(asset_dir) + "/earthmap.ppm", im)) return fail(ctx, MORT_ERR_IO, "scene names an image texture but earthmap.ppm was not found");
        for (ImageRec& r : ctx->scene.images) if (r.width == im.width && r.height == im.height) r.rgb = im.rgb;
    }
    return MORT_OK;
}
int mort_load_scene_text(mort_ctx* ctx, const char* path, const char* asset_dir) {
    CTX_CHECK(ctx && path); invalidate(ctx);
    std::string e;
    ctx->rng.reseed(1);                                   // a scene file starts from the unseeded host stream, like a fresh process
    if (!load_scene_text(ctx->scene, ctx->rng, path, asset_dir ? asset_dir : ".", &e)) return fail(ctx, MORT_ERR_SCENE, e);
    return MORT_OK;
}
int mort_dump_scene_text(mort_ctx* ctx, const char* path) {
    CTX_CHECK(ctx && path);
    std::string e;
    return dump_scene_text(ctx->scene, path, &e) ? MORT_OK : fail(ctx, MORT_ERR_IO, e);
}
int mort_dump_scene(mort_ctx* ctx, const char* path) { CTX_CHECK(ctx && path); return ctx->scene.dump(path) ? MORT_OK : fail(ctx, MORT_ERR_IO, std::string("cannot write ") + path); }
int mort_clear_scene(mort_ctx* ctx) { CTX_CHECK(ctx); invalidate(ctx); ctx->scene.clear(); ctx->rng.reseed(1); return MORT_OK; }

int mort_add_solid(mort_ctx* ctx, float r, float g, float b, mort_handle* out) { CTX_CHECK(ctx); invalidate(ctx); put(out, ctx->scene.add_solid(V3(r, g, b))); return MORT_OK; }
int mort_add_checker(mort_ctx* ctx, float scale, mort_handle even, mort_handle odd, mort_handle* out) { CTX_CHECK(ctx); invalidate(ctx); put(out, ctx->scene.add_checker(scale, H(even), H(odd))); return MORT_OK; }
int mort_add_image(mort_ctx* ctx, const uint8_t* rgb, int w, int h, mort_handle* out) { CTX_CHECK(ctx); invalidate(ctx); put(out, ctx->scene.add_image(rgb, w, h)); return MORT_OK; }
int mort_add_noise(mort_ctx* ctx, float scale, mort_handle* out) { CTX_CHECK(ctx); invalidate(ctx); put(out, ctx->scene.add_noise(scale, ctx->rng)); return MORT_OK; }
int mort_add_lambertian(mort_ctx* ctx, mort_handle tex, mort_handle* out) { CTX_CHECK(ctx); invalidate(ctx); put(out, ctx->scene.add_lambertian(H(tex))); return MORT_OK; }
int mort_add_metal(mort_ctx* ctx, float r, float g, float b, float fuzz, mort_handle* out) { CTX_CHECK(ctx); invalidate(ctx); put(out, ctx->scene.add_metal(V3(r, g, b), fuzz)); return MORT_OK; }
int mort_add_dielectric(mort_ctx* ctx, float ior, mort_handle* out) { CTX_CHECK(ctx); invalidate(ctx); put(out, ctx->scene.add_dielectric(ior)); return MORT_OK; }
int mort_add_diffuse_light(mort_ctx* ctx, mort_handle tex, mort_handle* out) { CTX_CHECK(ctx); invalidate(ctx); put(out, ctx->scene.add_diffuse_light(H(tex))); return MORT_OK; }
int mort_add_isotropic(mort_ctx* ctx, mort_handle tex, mort_handle* out) { CTX_CHECK(ctx); invalidate(ctx); put(out, ctx->scene.add_isotropic(H(tex))); return MORT_OK; }
int mort_add_sphere(mort_ctx* ctx, const float c[3], float r, mort_handle mat, int skip, mort_handle* out) { CTX_CHECK(ctx && c); invalidate(ctx); put(out, ctx->scene.add_sphere(v3(c), r, H(mat), skip != 0)); return MORT_OK; }
int mort_add_moving_sphere(mort_ctx* ctx, const float c0[3], const float c1[3], float r, mort_handle mat, int skip, mort_handle* out) { CTX_CHECK(ctx && c0 && c1); invalidate(ctx); put(out, ctx->scene.add_moving_sphere(v3(c0), v3(c1), r, H(mat), skip != 0)); return MORT_OK; }
int mort_add_quad(mort_ctx* ctx, const float Q[3], const float u[3], const float v[3], mort_handle mat, int skip, mort_handle* out) { CTX_CHECK(ctx && Q && u && v); invalidate(ctx); put(out, ctx->scene.add_quad(v3(Q), v3(u), v3(v), H(mat), skip != 0)); return MORT_OK; }
int mort_add_translate(mort_ctx* ctx, mort_handle obj, const float off[3], int skip, mort_handle* out) { CTX_CHECK(ctx && off); invalidate(ctx); put(out, ctx->scene.add_translate(H(obj), v3(off), skip != 0)); return MORT_OK; }
int mort_add_rotate_y(mort_ctx* ctx, mort_handle obj, float deg, int skip, mort_handle* out) { CTX_CHECK(ctx); invalidate(ctx); put(out, ctx->scene.add_rotate_y(H(obj), deg, skip != 0)); return MORT_OK; }
int mort_add_constant_medium(mort_ctx* ctx, mort_handle b, float density, mort_handle mat, int skip, mort_handle* out) { CTX_CHECK(ctx); invalidate(ctx); put(out, ctx->scene.add_constant_medium(H(b), density, H(mat), skip != 0)); return MORT_OK; }
int mort_add_list(mort_ctx* ctx, int skip, mort_handle* out) { CTX_CHECK(ctx); invalidate(ctx); put(out, ctx->scene.add_list(skip != 0)); return MORT_OK; }
int mort_list_add(mort_ctx* ctx, mort_handle list, mort_handle obj) { CTX_CHECK(ctx); invalidate(ctx); return ctx->scene.list_add(H(list), H(obj)) == 0 ? MORT_OK : fail(ctx, MORT_ERR_ARG, "mort_list_add: not a list handle"); }
int mort_add_bvh(mort_ctx* ctx, mort_handle list, int skip, mort_handle* out) { CTX_CHECK(ctx); invalidate(ctx); put(out, ctx->scene.add_bvh(H(list), skip != 0)); return MORT_OK; }
int mort_add_box(mort_ctx* ctx, const float a[3], const float b[3], mort_handle mat) { CTX_CHECK(ctx && a && b); invalidate(ctx); ctx->scene.box(v3(a), v3(b), H(mat)); return MORT_OK; }
int mort_add_rotated_box(mort_ctx* ctx, const float size[3], const float tr[3], float deg, mort_handle mat, mort_handle* out) { CTX_CHECK(ctx && size && tr); invalidate(ctx); put(out, ctx->scene.rotated_box(v3(size), v3(tr), deg, H(mat))); return MORT_OK; }
int mort_host_rand(mort_ctx* ctx) { return ctx ? ctx->rng.next() : 0; }

// ---- camera ---------------------------------------------------------------------------------------------
int mort_get_camera(mort_ctx* ctx, mort_camera_desc* o) {
    CTX_CHECK(ctx && o);
    const Camera& c = ctx->scene.cam;
    o->aspect_ratio = c.aspect_ratio; o->image_width = c.image_width; o->samples_per_pixel = c.samples_per_pixel; o->bounce_limit = c.bounce_limit; o->vfov = c.vfov;
    o->background[0] = c.background.x; o->background[1] = c.background.y; o->background[2] = c.background.z;
    o->lookfrom[0] = c.lookfrom.x; o->lookfrom[1] = c.lookfrom.y; o->lookfrom[2] = c.lookfrom.z;
    o->lookat[0] = c.lookat.x; o->lookat[1] = c.lookat.y; o->lookat[2] = c.lookat.z;
    o->vup[0] = c.vup.x; o->vup[1] = c.vup.y; o->vup[2] = c.vup.z;
    o->defocus_angle = c.defocus_angle; o->focus_dist = c.focus_dist; o->light_obj_type = c.light_obj_type; o->light_obj_idx = c.light_obj_idx;
    return MORT_OK;
}
int mort_set_camera(mort_ctx* ctx, const mort_camera_desc* d) {
    CTX_CHECK(ctx && d);
    if (d->image_width < 1 || !(d->aspect_ratio > 0) || d->samples_per_pixel < 1 || d->bounce_limit < 0) return fail(ctx, MORT_ERR_ARG, "mort_set_camera: bad width / aspect / spp / depth");
    Camera& c = ctx->scene.cam;
    c.aspect_ratio = d->aspect_ratio; c.image_width = d->image_width; c.samples_per_pixel = d->samples_per_pixel; c.bounce_limit = d->bounce_limit; c.vfov = d->vfov;
    c.background = v3(d->background); c.lookfrom = v3(d->lookfrom); c.lookat = v3(d->lookat); c.vup = v3(d->vup);
    c.defocus_angle = d->defocus_angle; c.focus_dist = d->focus_dist;
    // a camera move keeps the committed geometry (the reference re-runs only cam.initialize() per frame, mort.cu:90);
    // a different light handle changes the device light table, so it needs a new commit
    if (c.light_obj_type != d->light_obj_type || c.light_obj_idx != d->light_obj_idx) invalidate(ctx);
    c.light_obj_type = d->light_obj_type; c.light_obj_idx = d->light_obj_idx;
    c.initialize();
    if (ctx->committed) {
        if (camera_outgrew_pad(ctx)) return mort_commit(ctx);
        camera_params(c, ctx->flat.cam);
    }
    return MORT_OK;
}
int mort_override_camera(mort_ctx* ctx, int w, float aspect, int spp, int depth) {
    CTX_CHECK(ctx);
    Camera& c = ctx->scene.cam;
    if (w > 0) c.image_width = w;
    if (aspect > 0) c.aspect_ratio = aspect;
    if (spp > 0) c.samples_per_pixel = spp;
    if (depth > 0) c.bounce_limit = depth;
    c.initialize();
    if (ctx->committed) camera_params(c, ctx->flat.cam);       // geometry untouched: no rebuild needed
    return MORT_OK;
}
int mort_get_camera_record(mort_ctx* ctx, mscn_camera* out) { CTX_CHECK(ctx && out); ctx->scene.cam.to_record(*out); return MORT_OK; }

// ---- commit ---------------------------------------------------------------------------------------------
// tree builder handed to flatten_scene: the GPU builder (gpu_build.cu) for large scenes, the host builder otherwise
static bool commit_builder(void* user, const std::vector<BuildPrim>& prims, std::vector<Bvh4Node>& nodes, std::vector<int>& order, BuildStats& stats,
                           const BuildOptions& opt, std::string* err) {
    mort_ctx* ctx = static_cast<mort_ctx*>(user);
    const int b = ctx->build_opts.builder;
    const bool gpu = b == MORT_BUILD_GPU || (b == MORT_BUILD_AUTO && prims.size() >= 16384);
    if (!gpu) { build_bvh4(prims, nodes, order, stats, opt); return true; }
    if (gpu_build_bvh4(prims, nodes, order, stats, opt, ctx->stream, ctx->build_opts.gpu_flags, err)) return true;
    if (b == MORT_BUILD_GPU) return false;                // asked for explicitly: report
    cudaGetLastError();                                   // AUTO: e.g. no room for the workspace next to a host framework's pool — same tree from the host builder
    if (err) err->clear();
    build_bvh4(prims, nodes, order, stats, opt);
    return true;
}
static BuildOptions build_options(const mort_ctx* ctx) {
    BuildOptions o;
    const mort_build_opts& b = ctx->build_opts;
    if (b.max_leaf >= 1 && b.max_leaf <= MORT_MAX_LEAF) o.max_leaf = b.max_leaf;
    if (b.k_trav > 0) o.k_trav = b.k_trav;
    o.threads = b.host_threads > 0 ? b.host_threads : 0;
    o.gpu_small = b.gpu_small > 0 ? b.gpu_small : 0;
    return o;
}
// bottom-up bounds pass on the device records (refit.cu).  union_pass: the tree every kernel reads (needed after edits; a fresh build
// already holds these boxes).  motion: the time-0 / time-1 copy that motion.cu's kernels interpolate.
static int run_refit(mort_ctx* ctx, bool union_pass, bool motion) {
    const FlatScene& f = ctx->flat;
    const size_t ns = f.spheres.size(), nq = f.quads.size(), nb = f.nodes.size() * sizeof(Bvh4Node);
    for (int t = 0; t < (motion ? 2 : 1); t++) {
        if (!ctx->d_sphere_box[t] && ns) CU(ctx->arena.alloc(ns * 24, &ctx->d_sphere_box[t]));
        if (!ctx->d_quad_box[t] && nq) CU(ctx->arena.alloc(nq * 24, &ctx->d_quad_box[t]));
    }
    if (!ctx->d_extent) { void* p = nullptr; CU(ctx->arena.alloc(256, &p)); ctx->d_extent = static_cast<unsigned*>(p); }
    if (motion && !ctx->d_node_t1) {
        void* p = nullptr; CU(ctx->arena.alloc(nb, &p)); ctx->d_node_t0 = static_cast<Bvh4Node*>(p);
        CU(ctx->arena.alloc(nb, &p)); ctx->d_node_t1 = static_cast<Bvh4Node*>(p);
        CU(cudaMemcpyAsync(ctx->d_node_t0, ctx->dscene.nodes, nb, cudaMemcpyDeviceToDevice, ctx->stream));      // child words + the inverted boxes of empty slots
    }
    RefitArgs a;
    a.n_nodes = (int)f.nodes.size();
    a.spheres = ctx->dscene.spheres; a.quads = ctx->dscene.quads; a.instances = ctx->dscene.instances; a.n_spheres = (int)ns; a.n_quads = (int)nq;
    for (int t = 0; t < 2; t++) { a.sphere_box[t] = ctx->d_sphere_box[t]; a.quad_box[t] = ctx->d_quad_box[t]; }
    a.extent_key = ctx->d_extent; a.level_first = f.stats.level_first;
    const Camera& c = ctx->scene.cam;
    a.cam_center[0] = c.center.x; a.cam_center[1] = c.center.y; a.cam_center[2] = c.center.z;
    float pad = 0.f, extent = 0.f;
    if (union_pass) {
        a.nodes = const_cast<Bvh4Node*>(ctx->dscene.nodes); a.node_t1 = nullptr; a.motion = false;
        CU(refit_run(a, ctx->stream, &pad, &extent));
        ctx->flat.stats.pad = pad; ctx->flat.stats.scene_extent = extent;
    }
    if (motion) {
        a.nodes = ctx->d_node_t0; a.node_t1 = ctx->d_node_t1; a.motion = true;
        CU(refit_run(a, ctx->stream, &pad, &extent));
        ctx->motion_active = true; ctx->motion_nodes = (int)f.nodes.size();
    }
    return MORT_OK;
}

int mort_set_build_opts(mort_ctx* ctx, const mort_build_opts* o) {
    CTX_CHECK(ctx);
    mort_build_opts d; memset(&d, 0, sizeof(d));
    if (o) d = *o;
    if (d.builder < MORT_BUILD_AUTO || d.builder > MORT_BUILD_GPU || d.max_leaf < 0 || d.max_leaf > MORT_MAX_LEAF || d.k_trav < 0 || d.gpu_small < 0)
        return fail(ctx, MORT_ERR_ARG, "mort_set_build_opts: builder 0..2, max_leaf 0..4, k_trav >= 0, gpu_small >= 0");
    ctx->build_opts = d; invalidate(ctx);
    return MORT_OK;
}
int mort_get_build_info(mort_ctx* ctx, mort_build_info* out) {
    CTX_CHECK(ctx && out);
    memset(out, 0, sizeof(*out));
    const BuildStats& b = ctx->flat.stats;
    out->built_on_gpu = b.built_on_gpu; out->gpu_levels = b.gpu_levels; out->gpu_small_subtrees = b.gpu_small_subtrees; out->bvh2_nodes = b.n_bvh2_nodes;
    out->motion_nodes = ctx->motion_active ? ctx->motion_nodes : 0; out->refits = ctx->refits;
    out->flatten_ms = ctx->flatten_ms; out->build_ms = b.build_ms; out->gpu_stream_ms = b.gpu_kernel_ms; out->refit_ms = ctx->refit_ms;
    out->gpu_workspace_bytes = b.gpu_workspace_bytes;
    return MORT_OK;
}

int mort_commit(mort_ctx* ctx) {
    CTX_CHECK(ctx);
    invalidate(ctx);                                      // a failed re-commit must not leave a "committed" context over freed memory
    CU(cudaSetDevice(ctx->device));
    std::string e;
    const auto tf = std::chrono::steady_clock::now();
    if (!flatten_scene(ctx->scene, ctx->flat, &e, build_options(ctx), commit_builder, ctx)) return fail(ctx, MORT_ERR_SCENE, e);
    ctx->flatten_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tf).count() - ctx->flat.stats.build_ms;
    // the traversal stack holds 48 entries and a 4-wide node pushes at most 3: refuse (loudly) a tree it could overflow on
    if (!ctx->flat.linear && ctx->flat.stats.max_depth * 3 > 48)
        return fail(ctx, MORT_ERR_SCENE, "BVH deeper than 16 levels: the traversal stack (48 entries) could overflow; scene rejected");
    if (ctx->flat.spheres.size() >= (1u << 27) || ctx->flat.quads.size() >= (1u << 27))
        return fail(ctx, MORT_ERR_SCENE, "more than 2^27 primitives of one kind are not addressable by a leaf word");
    auto t0 = std::chrono::steady_clock::now();
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->stream != ctx->own_stream) CU(cudaStreamSynchronize(ctx->own_stream));
    ctx->arena.release();
    memset(&ctx->dscene, 0, sizeof(ctx->dscene));
    for (int t = 0; t < 2; t++) ctx->d_sphere_box[t] = ctx->d_quad_box[t] = nullptr;      // arena-owned: gone
    ctx->d_node_t0 = ctx->d_node_t1 = nullptr; ctx->d_extent = nullptr; ctx->motion_active = false; ctx->motion_nodes = 0;
    ctx->sphere_dirty.assign(ctx->scene.spheres.size(), 0); ctx->n_dirty = 0; ctx->refits = 0; ctx->refit_ms = 0;
    DeviceScene d; memset(&d, 0, sizeof(d));
    const FlatScene& f = ctx->flat;
    CU(ctx->arena.upload(f.nodes, &d.nodes)); d.n_nodes = (int)f.nodes.size();
    CU(ctx->arena.upload(f.spheres, &d.spheres)); CU(ctx->arena.upload(f.sphere_info, &d.sphere_info)); d.n_spheres = (int)f.spheres.size();
    CU(ctx->arena.upload(f.quads, &d.quads)); d.n_quads = (int)f.quads.size();
    CU(ctx->arena.upload(f.sphere_cls, &d.sphere_cls)); CU(ctx->arena.upload(f.quad_cls, &d.quad_cls));
    CU(ctx->arena.upload(f.instances, &d.instances)); d.n_instances = (int)f.instances.size();
    CU(ctx->arena.upload(f.materials, &d.materials)); d.n_materials = (int)f.materials.size();
    CU(ctx->arena.upload(f.textures, &d.textures)); d.n_textures = (int)f.textures.size();
    std::vector<ImageDesc> imgs;
    for (const ImageRec& im : ctx->scene.images) {
        ImageDesc I; I.texels = nullptr; I.width = im.width; I.height = im.height; I.cols = im.width * 3;
        if (!im.rgb.empty()) CU(ctx->arena.upload(im.rgb, &I.texels));
        imgs.push_back(I);
    }
    CU(ctx->arena.upload(imgs, &d.images)); d.n_images = (int)imgs.size();
    CU(ctx->arena.upload(f.noises, &d.noises)); d.n_noises = (int)f.noises.size();
    CU(ctx->arena.upload(f.media, &d.media)); d.n_media = (int)f.media.size();
    d.n_media_top = 0; for (const Medium& M : f.media) d.n_media_top += M.top_level ? 1 : 0;
    CU(ctx->arena.upload(f.boundary, &d.boundary)); d.n_boundary = (int)f.boundary.size();
    CU(ctx->arena.upload(f.lights, &d.lights)); d.n_lights = (int)f.lights.size(); d.light_kind = f.light_kind;
    d.post_media_order = f.post_media_order; d.two_pass = f.two_pass; d.empty = f.empty; d.linear = f.linear;
    ctx->dscene = d;
    const Scene& s = ctx->scene;
    int32_t off[8] = {0, 0, (int32_t)s.lambertians.size(), 0, 0, 0, 0, 0};
    off[3] = off[2] + (int32_t)s.metals.size(); off[4] = off[3] + (int32_t)s.dielectrics.size(); off[5] = off[4] + (int32_t)s.lights.size();
    CU(cudaMemcpy(ctx->d_mat_offsets, off, sizeof(off), cudaMemcpyHostToDevice));
    ctx->upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    // fingerprint of everything the kernels read (FNV-1a 64): ties a checkpoint to its scene and camera
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } };
    auto mixv = [&mix](const auto& v) { uint64_t n = v.size(); mix(&n, sizeof(n)); if (n) mix(v.data(), n * sizeof(v[0])); };
    mixv(f.nodes); mixv(f.spheres); mixv(f.sphere_info); mixv(f.quads); mixv(f.instances); mixv(f.materials); mixv(f.textures);
    mixv(f.noises); mixv(f.media); mixv(f.boundary); mixv(f.lights);
    for (const ImageRec& im : ctx->scene.images) { int32_t wh[2] = {im.width, im.height}; mix(wh, sizeof(wh)); mixv(im.rgb); }
    int32_t fl[5] = {f.light_kind, f.post_media_order, f.two_pass, f.empty, f.linear}; mix(fl, sizeof(fl));
    ctx->geometry_hash = h;                              // the camera is mixed in on demand: it may move without a new commit
    // motion-aware node boxes: one bottom-up pass over the uploaded tree (the topology was built from the union boxes)
    if (ctx->build_opts.motion_bounds && !f.linear && f.n_moving > 0) { const int rc = run_refit(ctx, false, true); if (rc != MORT_OK) return rc; }
    ctx->committed = true;
    return MORT_OK;
}

// ---- dynamic scenes: edit + refit -------------------------------------------------------------------------------------------
int mort_update_sphere(mort_ctx* ctx, mort_handle sphere, const float c0[3], const float c1[3], float radius) {
    CTX_CHECK(ctx && c0);
    if (sphere.type != MORT_OBJ_SPHERE || sphere.idx < 0 || sphere.idx >= (int)ctx->scene.spheres.size()) return fail(ctx, MORT_ERR_ARG, "mort_update_sphere: not a sphere handle");
    if (!(radius == radius) || !(c0[0] == c0[0] && c0[1] == c0[1] && c0[2] == c0[2]) || (c1 && !(c1[0] == c1[0] && c1[1] == c1[1] && c1[2] == c1[2])))
        return fail(ctx, MORT_ERR_ARG, "mort_update_sphere: NaN");
    const V3 a = v3(c0); V3 b; if (c1) b = v3(c1);
    ctx->scene.update_sphere(sphere.idx, a, c1 ? &b : nullptr, radius);
    // a sphere copied into a light or a medium-boundary record, or a scene that was not committed: only a new commit will do
    if (!ctx->committed || (size_t)sphere.idx >= ctx->flat.sphere_pinned.size() || ctx->flat.sphere_pinned[sphere.idx] || ctx->sphere_dirty.size() != ctx->scene.spheres.size()) { invalidate(ctx); return MORT_OK; }
    if (!ctx->sphere_dirty[sphere.idx]) { ctx->sphere_dirty[sphere.idx] = 1; ctx->n_dirty++; }
    return MORT_OK;
}
int mort_refit(mort_ctx* ctx) {
    CTX_CHECK(ctx);
    if (!ctx->committed) return fail(ctx, MORT_ERR_STATE, "mort_refit: no committed scene (a changed light / medium-boundary sphere or a structural change needs mort_commit)");
    CU(cudaSetDevice(ctx->device));
    const auto t0 = std::chrono::steady_clock::now();
    FlatScene& f = ctx->flat;
    if (ctx->n_dirty > 0) {
        // patch the device records of the edited spheres (a sphere slot may appear in several records: instanced lists)
        size_t lo = f.spheres.size(), hi = 0;
        uint64_t h = ctx->geometry_hash;
        auto mix = [&h](const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } };
        for (size_t r = 0; r < f.spheres.size(); r++) {
            const int slot = f.sphere_info[r].obj_idx;
            if (slot < 0 || (size_t)slot >= ctx->sphere_dirty.size() || !ctx->sphere_dirty[slot]) continue;
            const mscn_sphere& sp = ctx->scene.spheres[slot];
            SphereGeom& g = f.spheres[r];
            g.cx = sp.center[0]; g.cy = sp.center[1]; g.cz = sp.center[2]; g.r = sp.radius;
            g.vx = sp.moves ? sp.center_vec[0] : 0.f; g.vy = sp.moves ? sp.center_vec[1] : 0.f; g.vz = sp.moves ? sp.center_vec[2] : 0.f;
            lo = std::min(lo, r); hi = std::max(hi, r + 1);
            mix(&r, sizeof(r)); mix(&g, sizeof(g));
        }
        if (hi > lo) CU(cudaMemcpyAsync(const_cast<SphereGeom*>(ctx->dscene.spheres) + lo, f.spheres.data() + lo, (hi - lo) * sizeof(SphereGeom), cudaMemcpyHostToDevice, ctx->stream));
        ctx->geometry_hash = h;
        std::fill(ctx->sphere_dirty.begin(), ctx->sphere_dirty.end(), 0); ctx->n_dirty = 0;
    }
    if (!f.linear) { const int rc = run_refit(ctx, true, ctx->motion_active); if (rc != MORT_OK) return rc; }
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->refits++;
    ctx->refit_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return MORT_OK;
}

// ---- render ---------------------------------------------------------------------------------------------
void mort_default_render_opts(mort_render_opts* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->seed = 69420; o->frame = 0; o->mode = MORT_MODE_POOL; o->sample_mod = 1; o->sample_rem = 0; o->stage_nodes = 0;
}

static int ensure_accum(mort_ctx* ctx, size_t npix) {
    if (ctx->accum_pixels < npix) {
        cudaFree(ctx->d_accum); ctx->d_accum = nullptr; ctx->accum_pixels = 0;
        CU(cudaMalloc(&ctx->d_accum, npix * sizeof(float4))); ctx->accum_pixels = npix;
    }
    if (ctx->rgba_pixels < npix) {
        cudaFree(ctx->d_rgba); ctx->d_rgba = nullptr; ctx->rgba_pixels = 0;
        CU(cudaMalloc(&ctx->d_rgba, npix * 4)); ctx->rgba_pixels = npix;
    }
    return MORT_OK;
}

int mort_render_device(mort_ctx* ctx, const mort_render_opts* opts_in, void* d_accum) {
    CTX_CHECK(ctx && d_accum);
    if (!ctx->committed) return fail(ctx, MORT_ERR_STATE, "mort_render: scene not committed (call mort_commit)");
    mort_render_opts o; if (opts_in) o = *opts_in; else mort_default_render_opts(&o);
    CU(cudaSetDevice(ctx->device));
    // a scene with media nested in wrappers / lists is rendered by the one kernel that carries that visit order (stages.cu)
    const bool general_media = ctx->flat.two_pass == 2;
    if (general_media) { o.mode = MORT_MODE_POOL; o.threads_per_block = 512; o.blocks_per_sm = 2; o.pool_paths = 1024; o.pool_refill = -1; }
    if (o.mode == MORT_MODE_AUTO) {
        // Per scene class, from the two schedulers' measured rates on the ten shipped scenes (BASELINE.md sections 3 and 4; they render the same
        // exact frame, so this is a speed choice only).  The megakernel wins where a path is a few cheap segments in a linear-scan scene — no
        // light sampling, no noise texture (scenes 2, 3, 5: 5623 / 24 499 / 7214 against 4205 / 10 260 / 4814 Msamples/s) — and on linear-scan
        // scenes with media (scene 7: 1239 against 1143); the block wavefront wins everywhere else (scenes 1, 4, 6, 8, 9, the sphere fields).
        const FlatScene& fs = ctx->flat;
        const bool light_free = fs.light_kind == LIGHT_NONE && fs.noises.empty();
        o.mode = (o.n_frames <= 1 && fs.linear && (light_free || !fs.media.empty())) ? MORT_MODE_MEGAKERNEL : MORT_MODE_POOL;
        if (o.mode == MORT_MODE_MEGAKERNEL) { o.threads_per_block = 0; o.blocks_per_sm = 0; }      // pool shapes do not apply
    }
    const CameraParams& cam = ctx->flat.cam;
    if (o.sample_mod < 1 || o.sample_rem < 0 || o.sample_rem >= o.sample_mod) return fail(ctx, MORT_ERR_ARG, "mort_render: bad sample split");
    FrameParams p; memset(&p, 0, sizeof(p));
    p.sc = ctx->dscene; p.cam = cam; p.seed = o.seed; p.frame = o.frame; p.sj_mod = o.sample_mod; p.sj_rem = o.sample_rem;
    p.n_rows = cam.sqrt_spp > o.sample_rem ? (cam.sqrt_spp - o.sample_rem + o.sample_mod - 1) / o.sample_mod : 0;
    p.n_subset = p.n_rows * cam.sqrt_spp;
    p.n_pixels = cam.width * cam.height;
    p.tile_mod = o.tile_mod > 1 ? o.tile_mod : 1; p.tile_rem = o.tile_mod > 1 ? o.tile_rem : 0;
    const int rows = o.tile_rows > 0 ? o.tile_rows : 8;      // band height of the tile split
    p.band_px = rows * cam.width;
    if (p.tile_mod > 1) {
        if (o.tile_rem < 0 || o.tile_rem >= o.tile_mod) return fail(ctx, MORT_ERR_ARG, "mort_render: bad tile split");
        if (o.mode == MORT_MODE_WAVEFRONT) return fail(ctx, MORT_ERR_ARG, "mort_render: tile split needs the megakernel or the block wavefront");
        int n = 0;                                           // pixels in this rank's bands
        for (int b = p.tile_rem; b * rows < cam.height; b += p.tile_mod) n += std::min(rows, cam.height - b * rows) * cam.width;
        p.n_pixels = n;
    }
    // pixels per warp task: enough samples per task (~2048) to amortise the end-of-task tail, at most 16 pixels
    const int task_samples = 2048;                                    // (swept 256..8192 in profiles/r01_task_sweep.jsonl)
    int PT = p.n_subset > 0 ? (task_samples + p.n_subset - 1) / p.n_subset : 1;
    PT = std::max(1, std::min(16, PT));
    p.lanes_per_pixel = PT;
    // guided tail: tasks shrink to >= 256 samples (8 per lane) in the last round of the frame
    p.min_task_px = p.n_subset > 0 ? std::max(1, std::min(PT, (256 + p.n_subset - 1) / p.n_subset)) : PT;
    if (o.exact_accum && o.mode == MORT_MODE_WAVEFRONT) return fail(ctx, MORT_ERR_ARG, "mort_render: exact_accum needs the megakernel or the block wavefront");
    if (o.accumulate && !o.exact_accum) return fail(ctx, MORT_ERR_ARG, "mort_render: accumulate needs exact_accum (float sums are not order-independent)");
    p.accumulate = o.accumulate ? 1 : 0;
    p.accum = o.exact_accum ? nullptr : reinterpret_cast<float4*>(d_accum);
    p.accum_exact = o.exact_accum ? reinterpret_cast<unsigned long long*>(d_accum) : nullptr;
    p.counters = ctx->d_counters; p.work_counter = ctx->d_work;

    const int threads = o.threads_per_block > 0 ? (o.threads_per_block + 31) / 32 * 32 : 128;
    if (threads > 128 && o.mode == MORT_MODE_MEGAKERNEL) return fail(ctx, MORT_ERR_ARG, "mort_render: the megakernel takes at most 128 threads per block");
    // staging is opt-in (stage_nodes > 0): measured slower than L1-resident LDG.128 nodes (profiles/r01: scene 8 154 vs 276
    // Msamples/s), so <= 0 means none.  The megakernel keeps 6 KB of static shared memory of its own.
    int n_staged = o.stage_nodes;
    const int max_stage = (int)((ctx->prop.sharedMemPerBlockOptin > 6144 ? ctx->prop.sharedMemPerBlockOptin - 6144 : 0) / sizeof(Bvh4Node));
    if (n_staged < 0 || ctx->flat.linear || o.mode != MORT_MODE_MEGAKERNEL) n_staged = 0;     // nothing to stage for a linear-scan scene
    n_staged = std::min(n_staged, (int)ctx->flat.nodes.size());          // "more than the tree has" = the whole tree
    if (n_staged > max_stage) return fail(ctx, MORT_ERR_ARG, "mort_render: stage_nodes exceeds the shared memory of a block (" + std::to_string(max_stage) + " nodes at most)");
    p.n_staged = n_staged;

    CU(cudaMemsetAsync(ctx->d_counters, 0, 4 * sizeof(unsigned long long), ctx->stream));
    CU(cudaMemsetAsync(ctx->d_work, 0, sizeof(unsigned int), ctx->stream));
    uint64_t launches = 0;
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    if (o.n_frames > 1 && o.mode != MORT_MODE_POOL) return fail(ctx, MORT_ERR_ARG, "mort_render: n_frames > 1 is a block-wavefront feature");
    if (o.mode == MORT_MODE_MEGAKERNEL) {
        int occ = 0, regs = 0;
        // measured (profiles/r01_v4_variants.jsonl): 6 blocks/SM (80 regs) wins for the lockstep linear scan, 8 (64 regs) for the
        // larger BVH scenes; the spread between 6 and 8 is <= 3 % everywhere
        const int min_blocks = o.blocks_per_sm > 0 ? o.blocks_per_sm : (ctx->flat.linear ? 6 : 8);   // selects the register-capped variant
        CU(mega_query(threads, n_staged, min_blocks, ctx->flat.linear != 0, &occ, &regs));
        if (occ < 1) return fail(ctx, MORT_ERR_CUDA, "megakernel does not fit on an SM with this configuration");
        int bps = o.blocks_per_sm > 0 ? std::min(o.blocks_per_sm, occ) : occ;
        LaunchShape sh; sh.threads = threads; sh.blocks = bps * ctx->prop.multiProcessorCount; sh.smem_bytes = n_staged * (int)sizeof(Bvh4Node);
        CU(mega_launch(p, sh, min_blocks, ctx->stream));
        launches = 1;
        ctx->stats.threads_per_block = threads; ctx->stats.blocks_per_sm = bps; ctx->stats.regs_per_thread = regs; ctx->stats.staged_nodes = n_staged;
    } else if (o.mode == MORT_MODE_POOL) {
        // block wavefront (pool.cu): samples are added into an exact frame with integer reductions; a float4 request
        // goes through a context-owned exact frame and one resolve pass
        const size_t npix_full = (size_t)cam.width * cam.height;
        // a batch of frames (opts.n_frames > 1): frame keys frame .. frame + n - 1 rendered by ONE launch into n consecutive frames of
        // d_accum — the reference's frame loop (mort.cu:93-120) without a kernel tail and a launch gap per frame
        const int n_frames = o.n_frames > 1 ? o.n_frames : 1;
        if (n_frames > 1 && (n_frames > 256 || npix_full > (1u << 24) || p.tile_mod > 1 || o.accumulate))
            return fail(ctx, MORT_ERR_ARG, "mort_render: a batch takes at most 256 frames of at most 2^24 pixels, whole frames, no accumulate");
        unsigned long long* target = reinterpret_cast<unsigned long long*>(d_accum);
        if (!o.exact_accum) {
            if (ctx->pool_exact_pixels < npix_full * n_frames) {
                cudaFree(ctx->d_pool_exact); ctx->d_pool_exact = nullptr; ctx->pool_exact_pixels = 0;
                CU(cudaMalloc(&ctx->d_pool_exact, npix_full * n_frames * 4 * sizeof(unsigned long long))); ctx->pool_exact_pixels = npix_full * n_frames;
            }
            target = ctx->d_pool_exact;
        }
        p.accum = nullptr; p.accum_exact = target;
        p.frame_samples = (unsigned long long)p.n_pixels * (unsigned long long)p.n_subset;
        p.total_samples = p.frame_samples * (unsigned long long)n_frames; p.frame_words = (unsigned long long)npix_full * 4ull;
        p.pix_mask = n_frames > 1 ? 0x00FFFFFFu : 0xFFFFFFFFu; p.frame_shift = n_frames > 1 ? 24 : 0;
        p.work64 = ctx->d_work64;
        // Block shape, pool size and trace-phase form per scene class, from same-box A/B runs (profiles/r02_pool_ab.md):
        //   linear-scan scenes (<= 40 leaves)        2 blocks x 512 threads, 1024 paths each, generic kernel
        //   tree scenes without media                2 blocks x 512 threads, 1024 paths each, fixed 32-ray trace chunks
        //   tree scenes with media                   1 block x 640 threads (96 registers), 2048 paths, lane refill + classify pass
        //                                            that starts while the slowest rays are still in the tree
        const bool tree = !ctx->flat.linear, media = !ctx->flat.media.empty();
        PoolShape ps; ps.tree = tree ? 1 : 0;
        ps.threads = o.threads_per_block > 0 ? o.threads_per_block : (tree && media ? 640 : 512);
        ps.min_blocks = o.blocks_per_sm > 0 ? o.blocks_per_sm : (ps.threads >= 640 ? 1 : 2);
        ps.pool_paths = o.pool_paths > 0 ? o.pool_paths : (ps.min_blocks >= 2 ? 1024 : 2048);
        if (ps.pool_paths < 32 || ps.pool_paths > 32736) return fail(ctx, MORT_ERR_ARG, "mort_render: pool_paths must be in [32, 32736]");
        ps.pool_paths = (ps.pool_paths + 31) / 32 * 32;
        p.pool_paths = ps.pool_paths;
        p.pool_refill = o.pool_refill < 0 ? 0 : (o.pool_refill == 0 ? (tree && media ? 16 : 0) : std::min(31, o.pool_refill));
        p.pool_overlap = (o.pool_flags & 1) ? 0 : ((o.pool_flags & 2) ? 1 : (media && ps.min_blocks == 1 ? 1 : 0));
        int occ = 0, regs = 0, smem = 0;
        const bool motion = ctx->motion_active && tree && !general_media;      // kernels of motion.cu over the time-0 boxes + their change to time 1
        if (motion) { p.sc.nodes = ctx->d_node_t0; p.sc.node_dt = ctx->d_node_t1; }
        CU(general_media ? pool_query_stages(ps, &occ, &regs, &smem) : motion ? pool_query_motion(ps, &occ, &regs, &smem) : pool_query(ps, &occ, &regs, &smem));
        if (occ < 1) return fail(ctx, MORT_ERR_ARG, "mort_render: a pool of " + std::to_string(ps.pool_paths) + " paths (" + std::to_string(smem) + " B) does not fit in a block's shared memory");
        const int bps = std::min(occ, ps.min_blocks);
        CU(cudaMemsetAsync(ctx->d_work64, 0, sizeof(unsigned long long), ctx->stream));
        if (!o.accumulate || !o.exact_accum) CU(zero_exact_launch(target, p.n_pixels * n_frames, p.band_px, p.tile_mod, p.tile_rem, ctx->stream));
        const int grid = bps * ctx->prop.multiProcessorCount;
        CU(general_media ? pool_launch_stages(p, ps, grid, ctx->stream) : motion ? pool_launch_motion(p, ps, grid, ctx->stream) : pool_launch(p, ps, grid, ctx->stream));
        launches = 2;
        if (!o.exact_accum) { CU(resolve_exact_tiles_launch(target, p.n_pixels * n_frames, p.band_px, p.tile_mod, p.tile_rem, reinterpret_cast<float4*>(d_accum), ctx->stream)); launches = 3; }
        ctx->stats.threads_per_block = ps.threads; ctx->stats.blocks_per_sm = bps; ctx->stats.regs_per_thread = regs; ctx->stats.staged_nodes = 0;
    } else if (o.mode == MORT_MODE_WAVEFRONT) {
        int n_paths = std::max(o.wavefront_paths > 0 ? o.wavefront_paths : 1 << 21, p.n_pixels);   // at least one slot per pixel
        if (!ctx->wave || ctx->wave_paths != n_paths) {
            wavefront_free(ctx->wave); ctx->wave = nullptr;
            CU(wavefront_alloc(&ctx->wave, n_paths)); ctx->wave_paths = n_paths;
        }
        CU(wavefront_render(p, ctx->wave, n_paths, ctx->prop.multiProcessorCount, ctx->stream, &launches));
        ctx->stats.threads_per_block = 128; ctx->stats.blocks_per_sm = 8; ctx->stats.staged_nodes = 0; ctx->stats.regs_per_thread = 0;
    } else return fail(ctx, MORT_ERR_ARG, "mort_render: unknown mode");
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    unsigned long long cnt[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(cnt, ctx->d_counters, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (cnt[2] != 0) return fail(ctx, MORT_ERR_CUDA, "mort_render: the block wavefront's scheduler watchdog tripped (a warp waited for work for seconds): frame incomplete");
    float ms = 0; CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.reserved0 = (int32_t)sizeof(FrameParams);      // bytes of kernel parameters that go host->device per frame
    ctx->stats.last_render_ms = ms; ctx->stats.last_segments = cnt[0]; ctx->stats.last_samples = cnt[1]; ctx->stats.last_kernel_launches = launches;
    return MORT_OK;
}

int mort_resolve_exact_device(mort_ctx* ctx, const void* d_exact, void* d_accum) {
    CTX_CHECK(ctx && d_exact && d_accum);
    if (!ctx->committed) return fail(ctx, MORT_ERR_STATE, "scene not committed");
    const CameraParams& cam = ctx->flat.cam;
    CU(resolve_exact_launch(reinterpret_cast<const unsigned long long*>(d_exact), cam.width * cam.height, reinterpret_cast<float4*>(d_accum), ctx->stream));
    ctx->stats.last_kernel_launches += 1;
    return MORT_OK;
}

// ---- progressive accumulation + checkpoint / resume -------------------------------------------------------
int mort_accumulate_exact_device(mort_ctx* ctx, void* d_sum, const void* d_frame) {
    CTX_CHECK(ctx && d_sum && d_frame);
    if (!ctx->committed) return fail(ctx, MORT_ERR_STATE, "scene not committed");
    const CameraParams& cam = ctx->flat.cam;
    CU(accumulate_exact_launch(reinterpret_cast<unsigned long long*>(d_sum), reinterpret_cast<const unsigned long long*>(d_frame), cam.width * cam.height, ctx->stream));
    ctx->stats.last_kernel_launches += 1;
    return MORT_OK;
}

int mort_scene_fingerprint(mort_ctx* ctx, uint64_t* out) {
    CTX_CHECK(ctx && out);
    if (!ctx->committed) return fail(ctx, MORT_ERR_STATE, "scene not committed");
    *out = fingerprint(ctx);
    return MORT_OK;
}

namespace {
struct CheckpointHeader {            // 64 bytes, little endian
    char magic[4]; uint32_t version; int32_t width, height, samples_per_frame; uint32_t frames_done, seed, pad0;
    uint64_t fingerprint, payload_bytes; uint64_t reserved[2];
};
static_assert(sizeof(CheckpointHeader) == 64, "checkpoint header layout");
}  // namespace

int mort_save_checkpoint(mort_ctx* ctx, const char* path, const void* d_sum, uint32_t seed, uint32_t frames_done) {
    CTX_CHECK(ctx && path && d_sum);
    if (!ctx->committed) return fail(ctx, MORT_ERR_STATE, "scene not committed");
    const CameraParams& cam = ctx->flat.cam;
    const size_t bytes = (size_t)cam.width * cam.height * 4 * sizeof(unsigned long long);
    std::vector<unsigned long long> host((size_t)cam.width * cam.height * 4);
    CU(cudaMemcpyAsync(host.data(), d_sum, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CheckpointHeader hd; memset(&hd, 0, sizeof(hd));
    memcpy(hd.magic, "MCKP", 4); hd.version = 1; hd.width = cam.width; hd.height = cam.height; hd.samples_per_frame = cam.sqrt_spp * cam.sqrt_spp;
    hd.frames_done = frames_done; hd.seed = seed; hd.fingerprint = fingerprint(ctx); hd.payload_bytes = bytes;
    // write to a sibling file and rename: a crash mid-write never destroys the previous checkpoint
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(ctx, MORT_ERR_IO, "cannot write " + tmp);
    bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1 && fwrite(host.data(), 1, bytes, f) == bytes;
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return fail(ctx, MORT_ERR_IO, std::string("cannot write ") + path); }
    return MORT_OK;
}

int mort_load_checkpoint(mort_ctx* ctx, const char* path, void* d_sum, uint32_t* seed, uint32_t* frames_done) {
    CTX_CHECK(ctx && path && d_sum);
    if (!ctx->committed) return fail(ctx, MORT_ERR_STATE, "scene not committed");
    const CameraParams& cam = ctx->flat.cam;
    const size_t bytes = (size_t)cam.width * cam.height * 4 * sizeof(unsigned long long);
    FILE* f = fopen(path, "rb");
    if (!f) return fail(ctx, MORT_ERR_IO, std::string("cannot read ") + path);
    CheckpointHeader hd;
    if (fread(&hd, sizeof(hd), 1, f) != 1 || memcmp(hd.magic, "MCKP", 4) != 0 || hd.version != 1) { fclose(f); return fail(ctx, MORT_ERR_IO, std::string(path) + ": not a mort checkpoint"); }
    if (hd.width != cam.width || hd.height != cam.height || hd.samples_per_frame != cam.sqrt_spp * cam.sqrt_spp || hd.fingerprint != fingerprint(ctx) || hd.payload_bytes != bytes) {
        fclose(f);
        return fail(ctx, MORT_ERR_STATE, std::string(path) + ": checkpoint belongs to a different scene, camera or frame size");
    }
    std::vector<unsigned long long> host((size_t)cam.width * cam.height * 4);
    const bool ok = fread(host.data(), 1, bytes, f) == bytes;
    fclose(f);
    if (!ok) return fail(ctx, MORT_ERR_IO, std::string(path) + ": truncated checkpoint");
    CU(cudaMemcpyAsync(d_sum, host.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (seed) *seed = hd.seed;
    if (frames_done) *frames_done = hd.frames_done;
    return MORT_OK;
}

int mort_reset_progressive(mort_ctx* ctx) { CTX_CHECK(ctx); ctx->prog_frames = 0; ctx->prog_fingerprint = 0; return MORT_OK; }

int mort_render_progressive(mort_ctx* ctx, const mort_render_opts* opts, int n_frames, const char* checkpoint_path, int resume,
                            uint8_t* rgba8_out, float* accum_out, uint32_t* frames_total) {
    CTX_CHECK(ctx);
    if (!ctx->committed) return fail(ctx, MORT_ERR_STATE, "mort_render_progressive: scene not committed (call mort_commit)");
    if (n_frames < 0) return fail(ctx, MORT_ERR_ARG, "mort_render_progressive: n_frames < 0");
    CU(cudaSetDevice(ctx->device));
    const CameraParams& cam = ctx->flat.cam;
    const size_t npix = (size_t)cam.width * cam.height;
    int rc = ensure_accum(ctx, npix);
    if (rc != MORT_OK) return rc;
    mort_render_opts o; if (opts) o = *opts; else mort_default_render_opts(&o);
    if (o.mode == MORT_MODE_WAVEFRONT) return fail(ctx, MORT_ERR_ARG, "mort_render_progressive: needs exact sums (megakernel or block wavefront)");
    // the running image restarts when the scene, the camera, the frame size or the seed changed
    const uint64_t fp = fingerprint(ctx);
    if (ctx->prog_pixels != npix || ctx->prog_fingerprint != fp || ctx->prog_seed != o.seed || !ctx->d_prog) {
        if (ctx->prog_pixels != npix || !ctx->d_prog) {
            cudaFree(ctx->d_prog); ctx->d_prog = nullptr; ctx->prog_pixels = 0;
            CU(cudaMalloc(&ctx->d_prog, npix * 4 * sizeof(unsigned long long))); ctx->prog_pixels = npix;
        }
        ctx->prog_fingerprint = fp; ctx->prog_seed = o.seed; ctx->prog_frames = 0;
    }
    if (checkpoint_path && resume) {
        FILE* probe = fopen(checkpoint_path, "rb");
        if (probe) {
            fclose(probe);
            uint32_t seed = 0, done = 0;
            rc = mort_load_checkpoint(ctx, checkpoint_path, ctx->d_prog, &seed, &done);
            if (rc != MORT_OK) return rc;
            if (seed != o.seed) return fail(ctx, MORT_ERR_STATE, std::string(checkpoint_path) + ": checkpoint was rendered with a different seed");
            ctx->prog_frames = done;
        }
    }
    double ms = 0; uint64_t segs = 0, smps = 0, launches = 0;
    for (int f = 0; f < n_frames; f++) {
        o.exact_accum = 1; o.accumulate = ctx->prog_frames > 0 ? 1 : 0; o.frame = ctx->prog_frames; o.n_frames = 0;
        rc = mort_render_device(ctx, &o, ctx->d_prog);
        if (rc != MORT_OK) return rc;
        ctx->prog_frames++;
        ms += ctx->stats.last_render_ms; segs += ctx->stats.last_segments; smps += ctx->stats.last_samples; launches += ctx->stats.last_kernel_launches;
    }
    if (ctx->prog_frames == 0) return fail(ctx, MORT_ERR_STATE, "mort_render_progressive: no frame rendered or resumed yet");
    if (checkpoint_path && (n_frames > 0 || !resume)) {
        rc = mort_save_checkpoint(ctx, checkpoint_path, ctx->d_prog, o.seed, ctx->prog_frames);
        if (rc != MORT_OK) return rc;
    }
    rc = mort_resolve_exact_device(ctx, ctx->d_prog, ctx->d_accum);
    if (rc != MORT_OK) return rc;
    if (rgba8_out) {
        const long long spp_total = (long long)cam.sqrt_spp * cam.sqrt_spp * ctx->prog_frames;
        if (spp_total > 0x7FFFFFFF) return fail(ctx, MORT_ERR_ARG, "mort_render_progressive: more than 2^31 samples per pixel");
        rc = mort_tonemap_device(ctx, ctx->d_accum, (int)spp_total, ctx->d_rgba);
        if (rc != MORT_OK) return rc;
        CU(cudaMemcpyAsync(rgba8_out, ctx->d_rgba, npix * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (accum_out) CU(cudaMemcpyAsync(accum_out, ctx->d_accum, npix * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (n_frames > 0) { ctx->stats.last_render_ms = ms; ctx->stats.last_segments = segs; ctx->stats.last_samples = smps; ctx->stats.last_kernel_launches = launches + 2; }
    if (frames_total) *frames_total = ctx->prog_frames;
    return MORT_OK;
}

int mort_tonemap_device(mort_ctx* ctx, const void* d_accum, int spp_total, void* d_rgba8) {
    CTX_CHECK(ctx && d_accum && d_rgba8);
    if (!ctx->committed) return fail(ctx, MORT_ERR_STATE, "scene not committed");
    if (spp_total < 1) return fail(ctx, MORT_ERR_ARG, "mort_tonemap_device: samples_per_pixel_total < 1");
    const CameraParams& cam = ctx->flat.cam;
    float scale = (float)(1.0 / (double)spp_total);                       // camera.cuh:52 for the full sample set
    CU(tonemap_launch(reinterpret_cast<const float4*>(d_accum), cam.width * cam.height, scale, reinterpret_cast<uint8_t*>(d_rgba8), ctx->stream));
    ctx->stats.last_kernel_launches += 1;
    return MORT_OK;
}

int mort_render(mort_ctx* ctx, const mort_render_opts* opts, uint8_t* rgba8_out, float* accum_out) {
    CTX_CHECK(ctx);
    if (!ctx->committed) return fail(ctx, MORT_ERR_STATE, "mort_render: scene not committed (call mort_commit)");
    const CameraParams& cam = ctx->flat.cam;
    size_t npix = (size_t)cam.width * cam.height;
    int rc = ensure_accum(ctx, npix);
    if (rc != MORT_OK) return rc;
    mort_render_opts oh; if (opts) oh = *opts; else mort_default_render_opts(&oh);
    oh.exact_accum = 0; oh.accumulate = 0; oh.n_frames = 0;   // the host-buffer call returns ONE float4 image
    // A host-buffer frame is a whole frame: a tile split would return the other ranks' bands uninitialised (partial frames
    // are combined on the device: mort_render_device + mort_group_* / the caller's collective).  A sample split is
    // allowed; its 8-bit frame is the mean over the samples this call rendered.
    if (oh.tile_mod > 1) return fail(ctx, MORT_ERR_ARG, "mort_render: a tile split has no host-buffer form (use mort_render_device or mort_group_render)");
    rc = mort_render_device(ctx, &oh, ctx->d_accum);
    if (rc != MORT_OK) return rc;
    if (rgba8_out) {
        const int sm = oh.sample_mod > 1 ? oh.sample_mod : 1;
        const int rows = cam.sqrt_spp > oh.sample_rem ? (cam.sqrt_spp - oh.sample_rem + sm - 1) / sm : 0;
        rc = mort_tonemap_device(ctx, ctx->d_accum, std::max(1, rows * cam.sqrt_spp), ctx->d_rgba);
        if (rc != MORT_OK) return rc;
        CU(cudaMemcpyAsync(rgba8_out, ctx->d_rgba, npix * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (accum_out) CU(cudaMemcpyAsync(accum_out, ctx->d_accum, npix * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MORT_OK;
}

// ---- parity hook ----------------------------------------------------------------------------------------
int mort_trace(mort_ctx* ctx, const float* rays7, int n, mhit_record* out, mhit_medium_probe* probes, int flags) {
    CTX_CHECK(ctx && (n == 0 || (rays7 && out)));
    if (!ctx->committed) return fail(ctx, MORT_ERR_STATE, "mort_trace: scene not committed");
    if (n == 0) return MORT_OK;
    CU(cudaSetDevice(ctx->device));
    // one grow-only device buffer for the rays, the records and the probes (the parity tests call this thousands of times)
    const int nm = ctx->dscene.n_media_top;
    const size_t b_rays = ((size_t)n * 28 + 255) / 256 * 256, b_out = ((size_t)n * sizeof(mhit_record) + 255) / 256 * 256;
    const size_t b_pr = (probes && nm) ? (size_t)n * nm * sizeof(mhit_medium_probe) : 0;
    if (ctx->trace_bytes < b_rays + b_out + b_pr) {
        cudaFree(ctx->d_trace); ctx->d_trace = nullptr; ctx->trace_bytes = 0;
        CU(cudaMalloc(&ctx->d_trace, b_rays + b_out + b_pr)); ctx->trace_bytes = b_rays + b_out + b_pr;
    }
    char* base = static_cast<char*>(ctx->d_trace);
    float* d_rays = reinterpret_cast<float*>(base); mhit_record* d_out = reinterpret_cast<mhit_record*>(base + b_rays);
    mhit_medium_probe* d_pr = b_pr ? reinterpret_cast<mhit_medium_probe*>(base + b_rays + b_out) : nullptr;
    cudaError_t e = cudaMemcpyAsync(d_rays, rays7, (size_t)n * 28, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        if (ctx->motion_active && !(flags & MORT_TRACE_BRUTE_FORCE)) {      // through the interpolated boxes, like the renders of this scene
            DeviceScene ms = ctx->dscene; ms.nodes = ctx->d_node_t0; ms.node_dt = ctx->d_node_t1;
            e = trace_launch_motion(ms, d_rays, n, d_out, d_pr, 0, ctx->d_mat_offsets, ctx->stream);
        } else e = trace_launch(ctx->dscene, d_rays, n, d_out, d_pr, (flags & MORT_TRACE_BRUTE_FORCE) ? 1 : 0, ctx->d_mat_offsets, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(mhit_record), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && d_pr) e = cudaMemcpyAsync(probes, d_pr, (size_t)n * nm * sizeof(mhit_medium_probe), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(ctx, MORT_ERR_CUDA, std::string("mort_trace: ") + cudaGetErrorString(e));
    return MORT_OK;
}

// ---- stats ----------------------------------------------------------------------------------------------
int mort_get_stats(mort_ctx* ctx, mort_stats* out) {
    CTX_CHECK(ctx && out);
    mort_stats& s = ctx->stats;
    const Camera& c = ctx->scene.cam;
    s.width = c.image_width; s.height = c.image_height; s.sqrt_spp = c.sqrt_spp; s.bounce_limit = c.bounce_limit;
    const FlatScene& f = ctx->flat;
    s.n_leaves = f.stats.n_leaves; s.n_spheres = (int)f.spheres.size(); s.n_quads = (int)f.quads.size(); s.n_nodes = f.stats.n_nodes; s.bvh_depth = f.stats.max_depth;
    s.n_media = ctx->dscene.n_media_top; s.n_instances = (int)f.instances.size(); s.n_materials = (int)f.materials.size(); s.n_textures = (int)f.textures.size();
    s.sah_cost = f.stats.sah_cost; s.build_ms = f.stats.build_ms; s.upload_ms = ctx->upload_ms;
    s.sm_count = ctx->prop.multiProcessorCount; s.device_bytes = ctx->arena.bytes;
    *out = s;
    return MORT_OK;
}

}  // extern "C"
