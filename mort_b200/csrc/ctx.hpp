// ctx.hpp — the context behind the C ABI's opaque mort_ctx (shared by capi.cu and group.cu; not part of the public interface).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "flatten.hpp"
#include "mort_b200.h"
#include "render.hpp"
#include "scene.hpp"

namespace mort {

struct DeviceArena {                 // every device allocation of one committed scene
    std::vector<void*> ptrs; size_t bytes = 0;
    template <class T> cudaError_t upload(const std::vector<T>& v, const T** out) {
        *out = nullptr;
        if (v.empty()) return cudaSuccess;
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, v.size() * sizeof(T));
        if (e != cudaSuccess) return e;
        ptrs.push_back(p); bytes += v.size() * sizeof(T);
        e = cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
        *out = reinterpret_cast<const T*>(p);
        return e;
    }
    cudaError_t alloc(size_t n_bytes, void** out) {
        *out = nullptr;
        if (n_bytes == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(out, n_bytes);
        if (e != cudaSuccess) { *out = nullptr; return e; }
        ptrs.push_back(*out); bytes += n_bytes;
        return cudaSuccess;
    }
    void release() { for (void* p : ptrs) cudaFree(p); ptrs.clear(); bytes = 0; }
};

}  // namespace mort

using namespace mort;

struct mort_ctx {
    int device = 0;
    cudaDeviceProp prop;
    Scene scene;
    HostRng rng;
    FlatScene flat;
    bool committed = false;
    DeviceArena arena;
    DeviceScene dscene;
    int32_t* d_mat_offsets = nullptr;
    unsigned long long* d_counters = nullptr;     // [0] segments [1] samples
    unsigned int* d_work = nullptr;
    float4* d_accum = nullptr; size_t accum_pixels = 0;
    uint8_t* d_rgba = nullptr; size_t rgba_pixels = 0;
    WavefrontBuffers* wave = nullptr; int wave_paths = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    mort_stats stats;
    double upload_ms = 0;
    uint64_t geometry_hash = 0;
    void* d_trace = nullptr; size_t trace_bytes = 0;              // mort_trace: rays + records + probes (grow-only)
    void* comm = nullptr; int comm_world = 1, comm_rank = 0;      // NCCL communicator of a multi-process job (group.cu)
    unsigned long long* d_prog = nullptr; size_t prog_pixels = 0;   // progressive exact image (mort_render_progressive)
    unsigned long long* d_pool_exact = nullptr; size_t pool_exact_pixels = 0;   // block wavefront: exact frame behind a float4 request
    unsigned long long* d_work64 = nullptr;     // block wavefront: head of the sample index space
    uint64_t prog_fingerprint = 0; uint32_t prog_frames = 0, prog_seed = 0;
    // tree build / refit / motion bounds
    mort_build_opts build_opts = {};
    double flatten_ms = 0, refit_ms = 0;
    int refits = 0, motion_nodes = 0;
    bool motion_active = false;                                   // d_node_t0 / d_node_t1 hold the tree's boxes at time 0 and their change to time 1
    Bvh4Node* d_node_t0 = nullptr;
    void* d_sphere_box[2] = {nullptr, nullptr}; void* d_quad_box[2] = {nullptr, nullptr};   // refit work buffers (arena-owned, allocated on first use)
    Bvh4Node* d_node_t1 = nullptr; unsigned* d_extent = nullptr;
    std::vector<uint8_t> sphere_dirty; int n_dirty = 0;           // mort_update_sphere since the last refit / commit
};

#define CTX_CHECK(c) do { if (!(c)) return MORT_ERR_ARG; } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return MORT_ERR_CUDA; } } while (0)

static inline int fail(mort_ctx* ctx, int code, const std::string& m) { ctx->err = m; return code; }
static inline uint64_t fingerprint(const mort_ctx* ctx) {      // committed geometry + current camera (FNV-1a 64)
    uint64_t h = ctx->geometry_hash;
    const uint8_t* b = reinterpret_cast<const uint8_t*>(&ctx->flat.cam);
    for (size_t i = 0; i < sizeof(ctx->flat.cam); i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
