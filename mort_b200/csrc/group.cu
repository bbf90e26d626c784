// group.cu — multi-GPU rendering behind the C ABI: the frame is sharded over GPUs by samples or by
// screen tiles, every GPU renders its share independently into an EXACT partial frame (accum.cuh), and the partial
// frames are combined ONCE per frame with one NCCL sum-reduce of 64-bit integers over NVLink.  Integer sums are
// associative, so the N-GPU frame equals the single-GPU frame bit for bit, whatever NCCL's reduction order.
//
// Two ways to drive it, same collective:
//   mort_comm_*   one PROCESS per GPU (torchrun, MPI, ...): rank 0 makes a 128-byte id (mort_comm_unique_id), the
//                 launcher hands it to every rank, each rank attaches its context (mort_comm_attach) and calls
//                 mort_comm_reduce_exact on its partial frame
//   mort_group_*  one process drives every GPU of the box (the `mort --gpus N` CLI): a context and a host thread per
//                 device, ncclCommInitAll, mort_group_render = render shares + reduce + resolve + tone map + copy
// NCCL is bound at run time with dlopen("libnccl.so.2") — whichever copy the process already has (a host framework's)
// or the system one — so libmort_b200.so neither pins an NCCL build nor needs one for single-GPU use.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ctx.hpp"

namespace {

struct NcclApi {
    void* lib = nullptr; std::string err;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl() {
    static NcclApi api; static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
        if (!api.lib) { api.err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "?"); return; }
        auto sym = [&](const char* n) -> void* { void* p = dlsym(api.lib, n); if (!p && api.err.empty()) api.err = std::string("NCCL lacks ") + n; return p; };
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion"); api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank"); api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy"); api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart"); api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    return &api;
}
#define NC(ctx, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(ctx, MORT_ERR_CUDA, std::string(#call) + ": " + nccl()->GetErrorString(r_)); } while (0)

}  // namespace

struct mort_group {
    int n = 0;
    std::vector<mort_ctx*> ctx;
    std::vector<ncclComm_t> comm;
    std::vector<unsigned long long*> d_exact; std::vector<size_t> exact_pixels;
    std::string err;
    mort_group_stats stats;
};

extern "C" {

// ---- one process per GPU -----------------------------------------------------------------------------------
int mort_comm_unique_id(void* id128) {
    if (!id128) return MORT_ERR_ARG;
    NcclApi* N = nccl();
    if (!N->err.empty()) return MORT_ERR_CUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (N->GetUniqueId(&id) != ncclSuccess) return MORT_ERR_CUDA;
    memcpy(id128, &id, sizeof(id));
    return MORT_OK;
}

int mort_comm_attach(mort_ctx* ctx, const void* id128, int world, int rank) {
    CTX_CHECK(ctx && id128);
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, MORT_ERR_ARG, "mort_comm_attach: bad world / rank");
    NcclApi* N = nccl();
    if (!N->err.empty()) return fail(ctx, MORT_ERR_CUDA, N->err);
    if (ctx->comm) { N->CommDestroy((ncclComm_t)ctx->comm); ctx->comm = nullptr; }
    CU(cudaSetDevice(ctx->device));
    ncclUniqueId id; memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    NC(ctx, N->CommInitRank(&c, world, id, rank));
    ctx->comm = c; ctx->comm_world = world; ctx->comm_rank = rank;
    return MORT_OK;
}

int mort_comm_detach(mort_ctx* ctx) {
    CTX_CHECK(ctx);
    if (ctx->comm) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); nccl()->CommDestroy((ncclComm_t)ctx->comm); ctx->comm = nullptr; }
    ctx->comm_world = 1; ctx->comm_rank = 0;
    return MORT_OK;
}

int mort_comm_reduce_exact(mort_ctx* ctx, void* d_exact, int root) {
    CTX_CHECK(ctx && d_exact);
    if (!ctx->committed) return fail(ctx, MORT_ERR_STATE, "mort_comm_reduce_exact: scene not committed");
    if (!ctx->comm) return ctx->comm_world == 1 ? MORT_OK : fail(ctx, MORT_ERR_STATE, "mort_comm_reduce_exact: no communicator attached");
    if (root < 0 || root >= ctx->comm_world) return fail(ctx, MORT_ERR_ARG, "mort_comm_reduce_exact: bad root");
    const size_t words = (size_t)ctx->flat.cam.width * ctx->flat.cam.height * 4;
    CU(cudaSetDevice(ctx->device));
    NC(ctx, nccl()->Reduce(d_exact, d_exact, words, ncclUint64, ncclSum, root, (ncclComm_t)ctx->comm, ctx->stream));
    ctx->stats.last_kernel_launches += 1;
    return MORT_OK;
}

// ---- one process, every GPU of the box ---------------------------------------------------------------------
int mort_group_create(int n_devices, const int* devices, mort_group** out) {
    if (!out || n_devices < 1) return MORT_ERR_ARG;
    *out = nullptr;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) return MORT_ERR_CUDA;
    std::vector<int> devs(n_devices);
    for (int i = 0; i < n_devices; i++) { devs[i] = devices ? devices[i] : i; if (devs[i] < 0 || devs[i] >= have) return MORT_ERR_ARG; }
    mort_group* g = new mort_group();
    g->n = n_devices; g->ctx.assign(n_devices, nullptr); g->comm.assign(n_devices, nullptr);
    g->d_exact.assign(n_devices, nullptr); g->exact_pixels.assign(n_devices, 0);
    memset(&g->stats, 0, sizeof(g->stats));
    for (int i = 0; i < n_devices; i++)
        if (mort_create(devs[i], &g->ctx[i]) != MORT_OK) { mort_group_destroy(g); return MORT_ERR_CUDA; }
    if (n_devices > 1) {
        NcclApi* N = nccl();
        if (!N->err.empty() || N->CommInitAll(g->comm.data(), n_devices, devs.data()) != ncclSuccess) { mort_group_destroy(g); return MORT_ERR_CUDA; }
    }
    *out = g;
    return MORT_OK;
}

int mort_group_destroy(mort_group* g) {
    if (!g) return MORT_ERR_ARG;
    for (int i = 0; i < g->n; i++) {
        if (g->ctx[i]) { cudaSetDevice(g->ctx[i]->device); cudaDeviceSynchronize(); }
        if (g->comm[i]) nccl()->CommDestroy(g->comm[i]);
        if (g->d_exact[i]) cudaFree(g->d_exact[i]);
        if (g->ctx[i]) mort_destroy(g->ctx[i]);
    }
    delete g;
    return MORT_OK;
}

int mort_group_size(const mort_group* g) { return g ? g->n : 0; }
mort_ctx* mort_group_ctx(mort_group* g, int rank) { return (g && rank >= 0 && rank < g->n) ? g->ctx[rank] : nullptr; }
const char* mort_group_last_error(const mort_group* g) { return g ? g->err.c_str() : "null group"; }

int mort_group_render(mort_group* g, const mort_render_opts* opts_in, int split, uint8_t* rgba8_out, float* accum_out) {
    if (!g) return MORT_ERR_ARG;
    auto gfail = [&](int code, const std::string& m) { g->err = m; return code; };
    if (split != MORT_SPLIT_SAMPLE && split != MORT_SPLIT_TILE) return gfail(MORT_ERR_ARG, "mort_group_render: unknown split");
    mort_render_opts base; if (opts_in) base = *opts_in; else mort_default_render_opts(&base);
    if (base.mode == MORT_MODE_WAVEFRONT) return gfail(MORT_ERR_ARG, "mort_group_render: partial frames must be exact (megakernel or block wavefront)");
    const int n = g->n;
    for (int i = 0; i < n; i++) if (!g->ctx[i]->committed) return gfail(MORT_ERR_STATE, "mort_group_render: rank " + std::to_string(i) + " has no committed scene");
    const CameraParams cam = g->ctx[0]->flat.cam;
    const uint64_t fp0 = fingerprint(g->ctx[0]);
    for (int i = 1; i < n; i++) if (fingerprint(g->ctx[i]) != fp0) return gfail(MORT_ERR_STATE, "mort_group_render: the ranks hold different scenes or cameras");
    const size_t npix = (size_t)cam.width * cam.height;

    std::vector<int> rc(n, MORT_OK);
    std::vector<double> ms(n, 0.0); std::vector<uint64_t> segs(n, 0), smps(n, 0);
    cudaEvent_t c0 = nullptr, c1 = nullptr;
    // each GPU renders its share on its own host thread (a context is single-threaded; the launches are asynchronous but
    // mort_render_device waits for its counters)
    auto work = [&](int i) {
        mort_ctx* ctx = g->ctx[i];
        if (cudaSetDevice(ctx->device) != cudaSuccess) { rc[i] = MORT_ERR_CUDA; return; }
        if (g->exact_pixels[i] < npix) {
            cudaFree(g->d_exact[i]); g->d_exact[i] = nullptr; g->exact_pixels[i] = 0;
            if (cudaMalloc(&g->d_exact[i], npix * 32) != cudaSuccess) { rc[i] = fail(ctx, MORT_ERR_CUDA, "mort_group_render: cudaMalloc of the partial frame failed"); return; }
            g->exact_pixels[i] = npix;
        }
        mort_render_opts o = base;
        o.exact_accum = 1; o.accumulate = 0; o.n_frames = 0;
        if (split == MORT_SPLIT_SAMPLE) { o.sample_mod = n; o.sample_rem = i; o.tile_mod = 0; o.tile_rem = 0; }
        else { o.sample_mod = 1; o.sample_rem = 0; o.tile_mod = n; o.tile_rem = i; if (o.tile_rows <= 0) o.tile_rows = 2; }   // thin bands: every rank sees every part of the frame
        // a tile split leaves the other ranks' bands untouched: they must be zero for the sum
        if (split == MORT_SPLIT_TILE && n > 1 && cudaMemsetAsync(g->d_exact[i], 0, npix * 32, ctx->stream) != cudaSuccess) { rc[i] = MORT_ERR_CUDA; return; }
        rc[i] = mort_render_device(ctx, &o, g->d_exact[i]);
        ms[i] = ctx->stats.last_render_ms; segs[i] = ctx->stats.last_segments; smps[i] = ctx->stats.last_samples;
    };
    if (n == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < n; i++) th.emplace_back(work, i);
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < n; i++) if (rc[i] != MORT_OK) return gfail(rc[i], "rank " + std::to_string(i) + ": " + g->ctx[i]->err);

    mort_ctx* root = g->ctx[0];
    cudaSetDevice(root->device);
    float coll_ms = 0.f;
    if (n > 1) {
        // the one exchange step of the frame
        NcclApi* N = nccl();
        cudaEventCreate(&c0); cudaEventCreate(&c1);
        cudaEventRecord(c0, root->stream);
        ncclResult_t r = N->GroupStart();
        for (int i = 0; i < n && r == ncclSuccess; i++)
            r = N->Reduce(g->d_exact[i], g->d_exact[i], npix * 4, ncclUint64, ncclSum, 0, g->comm[i], g->ctx[i]->stream);
        if (r == ncclSuccess) r = N->GroupEnd();
        if (r != ncclSuccess) { cudaEventDestroy(c0); cudaEventDestroy(c1); return gfail(MORT_ERR_CUDA, std::string("ncclReduce: ") + N->GetErrorString(r)); }
        cudaSetDevice(root->device);
        cudaEventRecord(c1, root->stream);
    }
    // resolve + tone map + copy back on rank 0
    mort_ctx* ctx = root;
    if (ctx->accum_pixels < npix) { cudaFree(ctx->d_accum); ctx->d_accum = nullptr; ctx->accum_pixels = 0; CU(cudaMalloc(&ctx->d_accum, npix * sizeof(float4))); ctx->accum_pixels = npix; }
    if (ctx->rgba_pixels < npix) { cudaFree(ctx->d_rgba); ctx->d_rgba = nullptr; ctx->rgba_pixels = 0; CU(cudaMalloc(&ctx->d_rgba, npix * 4)); ctx->rgba_pixels = npix; }
    int rcr = mort_resolve_exact_device(ctx, g->d_exact[0], ctx->d_accum);
    if (rcr != MORT_OK) return gfail(rcr, ctx->err);
    if (rgba8_out) {
        rcr = mort_tonemap_device(ctx, ctx->d_accum, cam.sqrt_spp * cam.sqrt_spp, ctx->d_rgba);
        if (rcr != MORT_OK) return gfail(rcr, ctx->err);
        CU(cudaMemcpyAsync(rgba8_out, ctx->d_rgba, npix * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (accum_out) CU(cudaMemcpyAsync(accum_out, ctx->d_accum, npix * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 1; i < n; i++) { cudaSetDevice(g->ctx[i]->device); cudaStreamSynchronize(g->ctx[i]->stream); }
    cudaSetDevice(root->device);
    if (c0) { cudaEventElapsedTime(&coll_ms, c0, c1); cudaEventDestroy(c0); cudaEventDestroy(c1); }
    mort_group_stats& s = g->stats;
    s.n_gpus = n; s.split = split; s.collective_ms = coll_ms; s.collective_bytes = n > 1 ? npix * 32 : 0;
    s.kernel_ms_max = *std::max_element(ms.begin(), ms.end()); s.kernel_ms_min = *std::min_element(ms.begin(), ms.end());
    s.segments = 0; s.samples = 0;
    for (int i = 0; i < n; i++) { s.segments += segs[i]; s.samples += smps[i]; }
    return MORT_OK;
}

int mort_group_get_stats(mort_group* g, mort_group_stats* out) {
    if (!g || !out) return MORT_ERR_ARG;
    *out = g->stats;
    return MORT_OK;
}

}  // extern "C"
