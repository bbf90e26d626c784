// render.cu — megakernel, parity-trace and tone-map kernels for sm_100a.
//
// Megakernel = the reference's renderKernel (mort.cu:44-47 -> Camera::render, camera.cuh:178-208) rebuilt for
// B200: instead of one thread walking all sqrt_spp^2 samples of a pixel serially, a persistent WARP owns a
// pixel group at a time (work fetched from an atomic counter), its lanes take the pixel's samples in strides
// (lane l: samples l, l+G, ...), every lane regenerates its next camera path the moment the previous one
// ends (so the warp's lanes stay busy although path lengths differ), and the lane sums are combined with a
// fixed-order xor-shuffle tree — no atomics on the image, bit-reproducible frames.  All per-ray work lives
// in rt_core.cuh.
#include <cuda_runtime.h>

#include <cstdlib>

#include "accum.cuh"
#include "render.hpp"
#include "rt_core.cuh"
#include "trace_body.cuh"

namespace mort {

#define MEGA_PMAX 16

// 64-bit add into shared memory as two native 32-bit atomics.  atomicAdd(unsigned long long*) on shared memory compiles to a
// compare-and-swap spin loop (ATOMS.CAST.SPIN.64; 5 % of the stall samples on scene 1).  The low words wrap exactly
// floor(sum / 2^32) times whatever the order, so carrying each wrap into the high word keeps the sum exact mod 2^64.
// (Out of line it costs 28 % of Cornell's throughput — the call forces the caller's live state through the ABI — so it is
// force-inlined; profiles/r01_ab_bisect.jsonl, variant v1.)
__device__ __forceinline__ void smem_add64(unsigned long long* p, unsigned long long v) {
    if (v == 0ull) return;
    unsigned* w = reinterpret_cast<unsigned*>(p);
    const unsigned lo = (unsigned)v;
    const unsigned old = atomicAdd(w, lo);
    const unsigned hi = (unsigned)(v >> 32) + ((old + lo) < lo ? 1u : 0u);
    if (hi) atomicAdd(w + 1, hi);
}

// kMinBlocks = occupancy target handed to ptxas (register cap 65536 / (128 * kMinBlocks)): 4 -> 128 regs,
// 6 -> 80, 8 -> 64.  Which one wins is a measurement (profiles/), selectable through mort_render_opts.blocks_per_sm.
// kLinear: -1 = the scene's linear flag decides at run time (the shipped configuration), 1 / 0 = specialised builds
template <bool kStaged, int kMinBlocks, int kLinear>
__global__ void __launch_bounds__(128, kMinBlocks) mega_kernel(const __grid_constant__ FrameParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned long long part[4][MEGA_PMAX][4];                    // per warp: per task pixel: r, g, b, flags
    __shared__ unsigned long long lacc[4][4][32];                           // per warp: r, g, b, flags of every lane's current pixel
                                                                            // (shared memory instead of 8 live registers per lane; lane-private, conflict-free)
    const Bvh4Node* staged = reinterpret_cast<const Bvh4Node*>(smem_raw);
    if (kStaged) {
        float4* dst = reinterpret_cast<float4*>(smem_raw);
        const float4* src = reinterpret_cast<const float4*>(P.sc.nodes);
        const int nvec = P.n_staged * (int)(sizeof(Bvh4Node) / sizeof(float4));
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = __ldg(src + i);
        __syncthreads();
    }
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int PT = P.lanes_per_pixel;                     // pixels per warp task (1..MEGA_PMAX)
    const int sqrt_spp = P.cam.sqrt_spp, n_subset = P.n_subset;
    const int band_px = P.band_px;
    unsigned n_seg = 0, n_smp = 0;                        // per task, flushed to the 64-bit global counters at task end

    // A warp task = PT consecutive pixels = PT * n_subset samples in pixel-major order.  Lanes pull the next
    // sample the moment their path ends (ballot + prefix count: deterministic, no atomics), so the warp only
    // idles in the last few iterations of a task instead of after every lane's private quota.
    // Guided hand-out: once fewer than one full task per warp is left, tasks shrink with the remaining work (down to
    // P.min_task_px pixels) so the last round of the frame is spread over every warp instead of leaving most of
    // them idle behind the few that drew a late full-size task.  The peek is racy on purpose: any chunk size is
    // valid, and the exact accumulation makes the frame independent of how pixels were grouped into tasks.
    for (;;) {
        int base = 0, npx = 0;
        if (lane == 0) {
            int chunk = PT;
            if (P.min_task_px < PT) {
                const int n_warps = (int)gridDim.x * (int)(blockDim.x >> 5);
                const int left = P.n_pixels - (int)*reinterpret_cast<volatile unsigned int*>(P.work_counter);
                if (left < n_warps * PT) chunk = max(P.min_task_px, min(PT, left / n_warps));
            }
            base = (int)atomicAdd(P.work_counter, (unsigned)chunk);
            npx = min(chunk, P.n_pixels - base);
        }
        base = __shfl_sync(full, base, 0); npx = __shfl_sync(full, npx, 0);
        if (npx <= 0) break;
        const int items = npx * n_subset;
        if (lane < npx * 4) part[warp][lane >> 2][lane & 3] = 0ull;
        if (lane + 32 < npx * 4) part[warp][(lane + 32) >> 2][(lane + 32) & 3] = 0ull;
        __syncwarp();
        int next = 0;                                     // warp-uniform: next unassigned sample of the task
        int cur_p = -1;                                   // task pixel this lane currently accumulates for (in lacc)
        lacc[warp][0][lane] = 0ull; lacc[warp][1][lane] = 0ull; lacc[warp][2][lane] = 0ull; lacc[warp][3][lane] = 0ull;
        bool alive = false;
        Path path; Rng g;
        path.depth = 0; path.thr = mk3(1, 1, 1); path.ray.o = path.ray.d = mk3(0, 0, 0); path.ray.tm = 0.f;
        rng_init(g, 0, 0, 0, 0);
        for (;;) {
            const unsigned need = __ballot_sync(full, !alive);
            if (need != 0u && next < items) {
                const int idx = next + __popc(need & lt_mask);
                next += __popc(need);
                if (!alive && idx < items) {
                    const int p = idx / n_subset, k = idx - p * n_subset;
                    if (p != cur_p) {
                        if (cur_p >= 0) {                 // pixel switch: hand the finished partial sums over
#pragma unroll
                            for (int c = 0; c < 4; c++) { smem_add64(&part[warp][cur_p][c], lacc[warp][c][lane]); lacc[warp][c][lane] = 0ull; }
                        }
                        cur_p = p;
                    }
                    const int row = k / sqrt_spp;
                    path_start(P.cam, P.seed, P.frame, tile_to_global(base + p, band_px, P.tile_mod, P.tile_rem), k - row * sqrt_spp, P.sj_rem + row * P.sj_mod, path, g);
                    alive = true; n_smp++;
                }
            }
            if (__ballot_sync(full, alive) == 0u) break;
            if (alive) {
                f3 col; bool traced;
                const int st = path_segment<kStaged, kLinear>(P.sc, P.cam, staged, P.n_staged, path, g, col, traced);
                n_seg += traced ? 1u : 0u;
                if (st == SEG_DONE) {
                    unsigned long long af = 0ull;
                    if (isnan3(col)) af = 1ull;
                    else {
                        long long dr = 0, dg = 0, db = 0;
                        fx_add(dr, af, col.x, 20); fx_add(dg, af, col.y, 34); fx_add(db, af, col.z, 48);
                        lacc[warp][0][lane] += (unsigned long long)dr; lacc[warp][1][lane] += (unsigned long long)dg; lacc[warp][2][lane] += (unsigned long long)db;
                    }
                    if (af) lacc[warp][3][lane] += af;
                    alive = false;
                }
            }
        }
        if (cur_p >= 0) {
#pragma unroll
            for (int c = 0; c < 4; c++) smem_add64(&part[warp][cur_p][c], lacc[warp][c][lane]);
        }
        __syncwarp();
        if (P.accum_exact) {                                  // 4 words per pixel, coalesced over the task's pixels
            // every pixel belongs to exactly one task of a launch, so the progressive "+=" needs no atomics
#pragma unroll
            for (int h = 0; h < 64; h += 32) {
                const int l = lane + h;
                if (l < npx * 4) {
                    unsigned long long* dst = P.accum_exact + (size_t)tile_to_global(base + (l >> 2), band_px, P.tile_mod, P.tile_rem) * 4 + (l & 3);
                    unsigned long long v = part[warp][l >> 2][l & 3];
                    if (P.accumulate) v += *dst;
                    *dst = v;
                }
            }
        } else if (lane < npx)
            P.accum[tile_to_global(base + lane, band_px, P.tile_mod, P.tile_rem)] = fx_resolve((long long)part[warp][lane][0], (long long)part[warp][lane][1], (long long)part[warp][lane][2], part[warp][lane][3]);
        __syncwarp();
        for (int off = 16; off > 0; off >>= 1) { n_seg += __shfl_xor_sync(full, n_seg, off); n_smp += __shfl_xor_sync(full, n_smp, off); }
        if (lane == 0) { atomicAdd(P.counters, (unsigned long long)n_seg); atomicAdd(P.counters + 1, (unsigned long long)n_smp); }
        n_seg = 0; n_smp = 0;
    }
}

typedef void (*MegaFn)(const FrameParams);
// Specialising the kernel for linear-scan vs tree scenes (kLinear = 1 / 0: the other traversal is not compiled in) was
// measured and rejected: ptxas allocates the 80-register linear kernel worse without the tree code in it (268 / 428 B of
// spill stores / loads instead of 212 / 372) and Cornell drops from 1559 to 1458 Msamples/s, scenes 1 and 8 do not move
// (profiles/r01_ab_specialise.jsonl).  The template parameter stays for experiments; every variant decides at run time.
static MegaFn mega_variant(bool staged, int min_blocks, bool linear) {
    (void)linear;
    if (staged) return min_blocks >= 7 ? (MegaFn)mega_kernel<true, 8, -1> : (min_blocks >= 5 ? (MegaFn)mega_kernel<true, 6, -1> : (MegaFn)mega_kernel<true, 4, -1>);
    return min_blocks >= 7 ? (MegaFn)mega_kernel<false, 8, -1> : (min_blocks >= 5 ? (MegaFn)mega_kernel<false, 6, -1> : (MegaFn)mega_kernel<false, 4, -1>);
}

cudaError_t mega_query(int threads, int n_staged, int min_blocks, bool linear, int* max_blocks_per_sm, int* regs) {
    MegaFn fn = mega_variant(n_staged > 0, min_blocks, linear);
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, (const void*)fn);
    if (e != cudaSuccess) return e;
    if (regs) *regs = fa.numRegs;
    size_t smem = (size_t)n_staged * sizeof(Bvh4Node);
    if (n_staged > 0) {
        e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks_per_sm, (const void*)fn, threads, smem);
}

cudaError_t mega_launch(const FrameParams& p, const LaunchShape& shape, int min_blocks, cudaStream_t st) {
    MegaFn fn = mega_variant(p.n_staged > 0, min_blocks, p.sc.linear != 0);
    if (p.n_staged > 0) {
        cudaError_t e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, shape.smem_bytes);
        if (e != cudaSuccess) return e;
    }
    void* args[] = {(void*)&p};
    return cudaLaunchKernel((const void*)fn, dim3(shape.blocks), dim3(shape.threads), args, (size_t)shape.smem_bytes, st);
}

// ------------------------------------------------------------------------------------------------------
// mort_trace: one thread per ray; closest hit with media disabled + the two boundary probes per medium
// (what oracle/ref_harness.cu records from the reference).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) trace_kernel(const __grid_constant__ DeviceScene sc, const float* __restrict__ rays, int n,
                                                    mhit_record* __restrict__ out, mhit_medium_probe* __restrict__ probes,
                                                    int brute_force, const int32_t* __restrict__ mat_offsets) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) trace_one(sc, rays, i, out, probes, brute_force, mat_offsets);
}

cudaError_t trace_launch(const DeviceScene& sc, const float* d_rays, int n, mhit_record* d_out, mhit_medium_probe* d_probes,
                         int brute_force, const int32_t* d_mat_offsets, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    trace_kernel<<<(n + 127) / 128, 128, 0, st>>>(sc, d_rays, n, d_out, d_probes, brute_force, d_mat_offsets);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) resolve_exact_kernel(const unsigned long long* __restrict__ ex, int n, float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ulonglong2 a = reinterpret_cast<const ulonglong2*>(ex)[2 * (size_t)i], b = reinterpret_cast<const ulonglong2*>(ex)[2 * (size_t)i + 1];
    out[i] = fx_resolve((long long)a.x, (long long)a.y, (long long)b.x, b.y);
}
// progressive accumulation / merging partial frames: sum += frame, word by word (the sums are integers: any order, any grouping)
__global__ void __launch_bounds__(256) accumulate_exact_kernel(ulonglong2* __restrict__ sum, const ulonglong2* __restrict__ frame, size_t n2) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    ulonglong2 a = sum[i]; const ulonglong2 b = frame[i];
    a.x += b.x; a.y += b.y;
    sum[i] = a;
}
cudaError_t accumulate_exact_launch(unsigned long long* d_sum, const unsigned long long* d_frame, int n_pixels, cudaStream_t st) {
    if (n_pixels <= 0) return cudaSuccess;
    const size_t n2 = (size_t)n_pixels * 2;
    accumulate_exact_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(reinterpret_cast<ulonglong2*>(d_sum), reinterpret_cast<const ulonglong2*>(d_frame), n2);
    return cudaGetLastError();
}
cudaError_t resolve_exact_launch(const unsigned long long* d_exact, int n_pixels, float4* d_accum, cudaStream_t st) {
    if (n_pixels <= 0) return cudaSuccess;
    resolve_exact_kernel<<<(n_pixels + 255) / 256, 256, 0, st>>>(d_exact, n_pixels, d_accum);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------
// tone pipeline of camera.cuh:194-207
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tonemap_kernel(const float4* __restrict__ accum, int n, float scale, uchar4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = accum[i];
    uint8_t px[4];
    tonemap_pixel(a.x, a.y, a.z, scale, px);
    out[i] = make_uchar4(px[0], px[1], px[2], px[3]);
}
cudaError_t tonemap_launch(const float4* d_accum, int n_pixels, float scale, uint8_t* d_rgba8, cudaStream_t st) {
    if (n_pixels <= 0) return cudaSuccess;
    tonemap_kernel<<<(n_pixels + 255) / 256, 256, 0, st>>>(d_accum, n_pixels, scale, reinterpret_cast<uchar4*>(d_rgba8));
    return cudaGetLastError();
}

}  // namespace mort
