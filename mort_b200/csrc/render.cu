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

#include "render.hpp"
#include "rt_core.cuh"

namespace mort {

// ------------------------------------------------------------------------------------------------------
template <bool kStaged>
__global__ void __launch_bounds__(128) mega_kernel(const __grid_constant__ FrameParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Bvh4Node* staged = reinterpret_cast<const Bvh4Node*>(smem_raw);
    if (kStaged) {
        float4* dst = reinterpret_cast<float4*>(smem_raw);
        const float4* src = reinterpret_cast<const float4*>(P.sc.nodes);
        const int nvec = P.n_staged * (int)(sizeof(Bvh4Node) / sizeof(float4));
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = __ldg(src + i);
        __syncthreads();
    }
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int G = P.lanes_per_pixel, ppw = 32 / G;
    const int lp = lane / G, ls = lane - lp * G;
    const int sqrt_spp = P.cam.sqrt_spp;
    unsigned long long n_seg = 0, n_smp = 0;

    for (;;) {
        int base = 0;
        if (lane == 0) base = (int)atomicAdd(P.work_counter, (unsigned)ppw);
        base = __shfl_sync(full, base, 0);
        if (base >= P.n_pixels) break;
        const int pixel = base + lp;
        const bool valid = pixel < P.n_pixels;
        int k = ls;                                    // index into this call's sample subset
        float ax = 0.f, ay = 0.f, az = 0.f, an = 0.f;
        bool alive = false;
        Path path; Rng g;
        path.depth = 0; path.thr = mk3(1, 1, 1); path.ray.o = path.ray.d = mk3(0, 0, 0); path.ray.tm = 0.f;
        rng_init(g, 0, 0, 0, 0);
        for (;;) {
            if (!alive && valid && k < P.n_subset) {
                const int row = k / sqrt_spp;
                const int s_i = k - row * sqrt_spp, s_j = P.sj_rem + row * P.sj_mod;
                path_start(P.cam, P.seed, P.frame, pixel, s_i, s_j, path, g);
                alive = true; k += G; n_smp++;
            }
            if (__ballot_sync(full, alive) == 0u) break;
            if (alive) {
                f3 col; bool traced;
                const int st = path_segment<kStaged>(P.sc, P.cam, staged, P.n_staged, path, g, col, traced);
                n_seg += traced ? 1u : 0u;
                if (st == SEG_DONE) {
                    ax += col.x; ay += col.y; az += col.z;            // IEEE: a NaN sample poisons the channel (camera.cuh:190-198)
                    an += isnan3(col) ? 1.f : 0.f;
                    alive = false;
                }
            }
        }
        for (int off = G >> 1; off > 0; off >>= 1) {
            ax += __shfl_xor_sync(full, ax, off); ay += __shfl_xor_sync(full, ay, off);
            az += __shfl_xor_sync(full, az, off); an += __shfl_xor_sync(full, an, off);
        }
        if (ls == 0 && valid) P.accum[pixel] = make_float4(ax, ay, az, an);
    }
    for (int off = 16; off > 0; off >>= 1) { n_seg += __shfl_xor_sync(full, n_seg, off); n_smp += __shfl_xor_sync(full, n_smp, off); }
    if (lane == 0) { atomicAdd(P.counters, n_seg); atomicAdd(P.counters + 1, n_smp); }
}

cudaError_t mega_query(int threads, int n_staged, int* max_blocks_per_sm, int* regs) {
    cudaFuncAttributes fa;
    cudaError_t e = n_staged > 0 ? cudaFuncGetAttributes(&fa, mega_kernel<true>) : cudaFuncGetAttributes(&fa, mega_kernel<false>);
    if (e != cudaSuccess) return e;
    if (regs) *regs = fa.numRegs;
    size_t smem = (size_t)n_staged * sizeof(Bvh4Node);
    if (n_staged > 0) {
        e = cudaFuncSetAttribute(mega_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks_per_sm, mega_kernel<true>, threads, smem);
    }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks_per_sm, mega_kernel<false>, threads, 0);
}

cudaError_t mega_launch(const FrameParams& p, const LaunchShape& shape, cudaStream_t st) {
    if (p.n_staged > 0) {
        cudaError_t e = cudaFuncSetAttribute(mega_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, shape.smem_bytes);
        if (e != cudaSuccess) return e;
        mega_kernel<true><<<shape.blocks, shape.threads, shape.smem_bytes, st>>>(p);
    } else {
        mega_kernel<false><<<shape.blocks, shape.threads, 0, st>>>(p);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------
// mort_trace: one thread per ray; closest hit with media disabled + the two boundary probes per medium
// (what oracle/ref_harness.cu records from the reference).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) trace_kernel(const __grid_constant__ DeviceScene sc, const float* __restrict__ rays, int n,
                                                    mhit_record* __restrict__ out, mhit_medium_probe* __restrict__ probes,
                                                    int brute_force, const int32_t* __restrict__ mat_offsets) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
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
        for (int m = 0; m < sc.n_media; m++) {
            mhit_medium_probe p; p.hit1 = p.hit2 = 0; p.t1 = p.t2 = 0.f;
            float t1, t2;
            if (boundary_probe(sc, sc.media[m], r, -INFINITY, INFINITY, t1)) {
                p.hit1 = 1; p.t1 = t1;
                if (boundary_probe(sc, sc.media[m], r, (float)((double)t1 + 0.0001), INFINITY, t2)) { p.hit2 = 1; p.t2 = t2; }
            }
            probes[(size_t)i * sc.n_media + m] = p;
        }
}

cudaError_t trace_launch(const DeviceScene& sc, const float* d_rays, int n, mhit_record* d_out, mhit_medium_probe* d_probes,
                         int brute_force, const int32_t* d_mat_offsets, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    trace_kernel<<<(n + 127) / 128, 128, 0, st>>>(sc, d_rays, n, d_out, d_probes, brute_force, d_mat_offsets);
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
