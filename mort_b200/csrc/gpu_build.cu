// gpu_build.cu — binned-SAH build of the 4-wide BVH on the GPU (the reference's builder is a host
// bubble sort capped at 1024 nodes, objects.cuh:528-661; the host SAH builder here takes ~0.4-1 s for the 1 M-sphere field of
// BASELINE config 4, about one frame).  Kernels = the bodies of gpu_build_core.cuh; the level loop is gpu_build_driver.hpp.
// The tree equals the host builder's node for node (bvh_sah.hpp), so nothing downstream can tell which one ran.
//
// Hardware notes: the per-primitive kernels are atomic-bound at the top levels (10^6 primitives, one node), so
//   k_stats      a warp that lies inside one node reduces its 16 values with redux.sync and issues 16 atomics instead of 512
//   k_bin        a block that lies inside one node bins into a private copy in shared memory (native shared atomics) and
//                flushes at most 384 global atomics
//   k_partition  a warp inside one node reserves its ranks with two atomics (ballot + popc)
// Deeper levels have many small nodes, little contention, and take the plain per-thread path.  Every kernel is a streaming
// pass over <= 40 B per primitive: the build is latency- / launch-bound (about 5 launches and one 48-byte read-back per level).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "gpu_build.hpp"
#include "gpu_build_driver.hpp"

namespace mort {
namespace {

using namespace gb;
constexpr int kT = 256;
constexpr unsigned kFull = 0xffffffffu;

__global__ void __launch_bounds__(kT) k_init(Ctx c, int sv) {
    const int i = blockIdx.x * kT + threadIdx.x;
    if (i < c.N) { c.idx[0][i] = i; c.slot[0][i] = sv; }
}
__global__ void __launch_bounds__(kT) k_clear(Ctx c, int n) {
    const size_t total = (size_t)n * kBinWords;
    for (size_t w = (size_t)blockIdx.x * kT + threadIdx.x; w < total; w += (size_t)gridDim.x * kT) c.bins[w] = bin_identity((int)(w & 7));
    for (int s = blockIdx.x * kT + threadIdx.x; s < n; s += gridDim.x * kT) stats_clear(c.stats[s]);
}
__global__ void __launch_bounds__(kT) k_stats(Ctx c, int cur, int aggregate) {
    const int pos = blockIdx.x * kT + threadIdx.x;
    const int s = pos < c.N ? c.slot[cur][pos] : -1;
    const int s0 = __shfl_sync(kFull, s, 0);
    const bool uni = aggregate && __all_sync(kFull, s == s0) && s0 >= 0;
    if (!uni) { if (s >= 0) body_stats(c, cur, pos); return; }
    const int ref = c.idx[cur][pos];
    const BuildPrim p = c.prims[ref];
    uint32_t lo[3], hi[3], clo[3], chi[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        lo[a] = __reduce_min_sync(kFull, key_of(p.lo[a])); hi[a] = __reduce_max_sync(kFull, key_of(p.hi[a]));
        const uint32_t ck = key_of(sah_centroid(p, a));
        clo[a] = __reduce_min_sync(kFull, ck); chi[a] = __reduce_max_sync(kFull, ck);
    }
    const int cost = (int)__reduce_add_sync(kFull, (unsigned)sah_prim_cost(p.type));
    const int nsph = (int)__reduce_add_sync(kFull, p.type != MORT_OBJ_QUAD ? 1u : 0u);
    const int rmin = (int)__reduce_min_sync(kFull, (unsigned)ref), rmax = (int)__reduce_max_sync(kFull, (unsigned)ref);
    if ((threadIdx.x & 31) == 0) {
        GStats& G = c.stats[s0];
#pragma unroll
        for (int a = 0; a < 3; a++) { atomicMin(&G.lo[a], lo[a]); atomicMax(&G.hi[a], hi[a]); atomicMin(&G.clo[a], clo[a]); atomicMax(&G.chi[a], chi[a]); }
        atomicAdd(&G.cost, cost); if (nsph) atomicAdd(&G.n_spheres, nsph);
        atomicMin(&G.min_ref, rmin); atomicMax(&G.max_ref, rmax);
    }
}
__global__ void __launch_bounds__(kT) k_bin(Ctx c, int cur, int privatize) {
    __shared__ uint32_t sb[kBinWords];
    __shared__ int s_first;
    const int pos = blockIdx.x * kT + threadIdx.x;
    const int s = pos < c.N ? c.slot[cur][pos] : -1;
    if (threadIdx.x == 0) s_first = s;
    __syncthreads();
    const int s0 = s_first;
    const int same = __syncthreads_and(s == s0 && s0 >= 0);
    if (privatize && same) {
        for (int w = threadIdx.x; w < kBinWords; w += kT) sb[w] = bin_identity(w);
        __syncthreads();
        body_bin(c, cur, pos, s0, sb);
        __syncthreads();
        uint32_t* B = c.bins + (size_t)s0 * kBinWords;
        for (int w = threadIdx.x; w < kBinWords; w += kT) {
            const uint32_t v = sb[w];
            if (v == bin_identity(w)) continue;
            const int f = w & 7;
            if (f < 3) atomicMin(B + w, v); else if (f < 6) atomicMax(B + w, v); else atomicAdd(B + w, v);
        }
    } else if (s >= 0) body_bin(c, cur, pos, s, c.bins + (size_t)s * kBinWords);
}
__global__ void __launch_bounds__(64) k_split(Ctx c, int cur, int n) {
    const int s = blockIdx.x * 64 + threadIdx.x;
    if (s < n) body_split(c, cur, s);
}
__global__ void __launch_bounds__(kT) k_partition(Ctx c, int cur, int aggregate) {
    const int pos = blockIdx.x * kT + threadIdx.x;
    const bool in = pos < c.N;
    const int s = in ? c.slot[cur][pos] : -1;
    const int s0 = __shfl_sync(kFull, s, 0);
    const bool uni = aggregate && __all_sync(kFull, s == s0) && s0 >= 0;
    if (!uni) { if (in) body_partition(c, cur, pos); return; }
    const int lane = threadIdx.x & 31;
    const bool left = partition_side(c, cur, pos, s0);
    const unsigned ml = __ballot_sync(kFull, left);
    const int nl = __popc(ml);
    int bl = 0, br = 0;
    if (lane == 0) {
        if (nl) bl = atomicAdd(&c.split[s0].cur_l, nl);
        if (nl < 32) br = atomicAdd(&c.split[s0].cur_r, 32 - nl);
    }
    bl = __shfl_sync(kFull, bl, 0); br = __shfl_sync(kFull, br, 0);
    const unsigned lt = (1u << lane) - 1u;
    partition_place(c, cur, pos, s0, left, left ? bl + __popc(ml & lt) : br + __popc(~ml & lt));
}
__global__ void __launch_bounds__(64) k_small(Ctx c, int cur, int n) {
    const int j = blockIdx.x * 64 + threadIdx.x;
    if (j < n) body_small(c, cur, j);
}
__global__ void __launch_bounds__(kT) k_collapse_count(Ctx c, int cur, int n) {
    const int i = blockIdx.x * kT + threadIdx.x;
    if (i < n) body_collapse_count(c, cur, i);
}
// exclusive scan of icount[0, n) by ONE block of 1024 threads, 1024 values per round with a running carry; total -> collapse_total
__global__ void __launch_bounds__(1024) k_scan(Ctx c, int n) {
    __shared__ int warp_sum[32];
    __shared__ int tile_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + (int)threadIdx.x;
        const int v = i < n ? c.icount[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(kFull, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            const int w = warp_sum[lane];
            int xs = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(kFull, xs, o); if (lane >= o) xs += y; }
            warp_sum[lane] = xs - w;
            if (lane == 31) tile_total = xs;
        }
        __syncthreads();
        if (i < n) c.ioff[i] = carry + warp_sum[warp] + x - v;
        carry += tile_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) c.cnt->collapse_total = carry;
}
__global__ void __launch_bounds__(kT) k_collapse_emit(Ctx c, int cur, int base, int n) {
    const int i = blockIdx.x * kT + threadIdx.x;
    if (i < n) body_collapse_emit(c, cur, i, base, n);
}

struct CudaExec {
    cudaStream_t st; int flags;
    void* ws = nullptr; cudaError_t e = cudaSuccess; const char* where = "";
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void note(cudaError_t r, const char* w) { if (e == cudaSuccess && r != cudaSuccess) { e = r; where = w; } }
    int grid(int n) const { return (n + kT - 1) / kT; }
    void* alloc(size_t b) { note(cudaMalloc(&ws, b), "cudaMalloc(workspace)"); return e == cudaSuccess ? ws : nullptr; }
    void upload(void* d, const void* s, size_t b) { note(cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, st), "upload"); }
    void download(void* d, const void* s, size_t b) { note(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, st), "download"); note(cudaStreamSynchronize(st), "sync"); }
    void k_init(const Ctx& c, int sv) { mort::k_init<<<grid(c.N), kT, 0, st>>>(c, sv); }
    void k_clear(const Ctx& c, int n) { const size_t w = (size_t)n * kBinWords; mort::k_clear<<<(int)std::min<size_t>((w + kT - 1) / kT, 4096), kT, 0, st>>>(c, n); }
    void k_stats(const Ctx& c, int cur) { mort::k_stats<<<grid(c.N), kT, 0, st>>>(c, cur, (flags & 1) ? 0 : 1); }
    void k_bin(const Ctx& c, int cur) { mort::k_bin<<<grid(c.N), kT, 0, st>>>(c, cur, (flags & 1) ? 0 : 1); }
    void k_split(const Ctx& c, int cur, int n) { mort::k_split<<<(n + 63) / 64, 64, 0, st>>>(c, cur, n); }
    void k_partition(const Ctx& c, int cur) { mort::k_partition<<<grid(c.N), kT, 0, st>>>(c, cur, (flags & 1) ? 0 : 1); }
    void k_small(const Ctx& c, int cur, int n) { mort::k_small<<<(n + 63) / 64, 64, 0, st>>>(c, cur, n); }
    void k_collapse_count(const Ctx& c, int cur, int n) { mort::k_collapse_count<<<grid(n), kT, 0, st>>>(c, cur, n); }
    void k_scan(const Ctx& c, int n) { mort::k_scan<<<1, 1024, 0, st>>>(c, n); }
    void k_collapse_emit(const Ctx& c, int cur, int base, int n) { mort::k_collapse_emit<<<grid(n), kT, 0, st>>>(c, cur, base, n); }
    bool ok(std::string* err) {
        note(cudaGetLastError(), "kernel launch");
        if (e != cudaSuccess && err) *err = std::string(where) + ": " + cudaGetErrorString(e);
        return e == cudaSuccess;
    }
};

}  // namespace

bool gpu_build_bvh4(const std::vector<BuildPrim>& prims, std::vector<Bvh4Node>& nodes, std::vector<int>& order, BuildStats& stats,
                    const BuildOptions& opt, void* cuda_stream, int flags, std::string* err) {
    const auto t0 = std::chrono::steady_clock::now();
    CudaExec x; x.st = (cudaStream_t)cuda_stream; x.flags = flags;
    cudaEventCreate(&x.ev0); cudaEventCreate(&x.ev1);
    cudaEventRecord(x.ev0, x.st);
    const int k_small = opt.gpu_small > 0 ? opt.gpu_small : 64;
    bool good = build_run(x, prims, nodes, order, stats, opt, k_small, err);
    if (good) good = x.ok(err); else x.ok(nullptr);
    cudaEventRecord(x.ev1, x.st);
    cudaStreamSynchronize(x.st);
    float ms = 0.f; cudaEventElapsedTime(&ms, x.ev0, x.ev1);
    cudaEventDestroy(x.ev0); cudaEventDestroy(x.ev1);
    if (x.ws) cudaFree(x.ws);
    if (!good && err && err->empty()) *err = "GPU tree build failed";
    stats.built_on_gpu = good ? 1 : 0;
    stats.gpu_kernel_ms = ms;                    // stream time from the first upload to the last read-back (includes the per-level read-backs)
    stats.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return good;
}

}  // namespace mort
