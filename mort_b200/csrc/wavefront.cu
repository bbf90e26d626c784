// wavefront.cu — wavefront form of the integrator for sm_100a (the alternative to render.cu's megakernel).
//
// Path state lives in HBM as coalesced SoA arrays of 16-byte vectors indexed by SLOT; slot = pixel * S + j owns
// the samples j, j+S, j+2S, ... of its pixel and a private accumulator, so — exactly like the megakernel's
// lanes — no atomics ever touch the image and frames are bit-reproducible.  One wave =
//     wf_extend            closest hit + media for every active slot, then a material-class split:
//                          slots are appended to per-class queues with warp-ballot / popc aggregation
//     wf_shade<class> x4   terminal (miss / emitter / absorbed), diffuse (texture + ONB + pdf + light
//                          sampling), metal, dielectric — each over its own compacted queue; finished samples
//                          are accumulated and the slot REGENERATES its next camera path in place; survivors
//                          are compacted into the next wave's active queue
// Every kernel is a persistent grid whose warps pull 32-item chunks from an atomic work counter.
// The per-ray code is rt_core.cuh, shared with the megakernel: same Philox stream, same arithmetic, same
// estimator — the two variants render the same image up to the order of the per-pixel float sums.
#include <cuda_runtime.h>

#include "render.hpp"
#include "rt_core.cuh"

namespace mort {

enum { C_ACTIVE0 = 0, C_ACTIVE1 = 1, C_CLASS0 = 2, /* 2..5 */ C_HEAD_EXTEND = 6, C_HEAD_SHADE0 = 7, /* 7..10 */ C_HEAD_START = 11, C_N = 16 };

struct WavefrontBuffers {
    int n = 0;
    float4 *ray_o = nullptr, *ray_d = nullptr, *thr = nullptr, *hit = nullptr, *acc = nullptr;
    uint4* rng = nullptr;                       // Philox block counter, -, next subset sample index, current sample id
    uint32_t* q_active[2] = {nullptr, nullptr};
    uint32_t* q_class[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t* counters = nullptr;               // C_N words
    uint32_t* h_counters = nullptr;             // pinned
};

size_t wavefront_bytes(int n) { return (size_t)n * (5 * 16 + 16 + 6 * 4) + C_N * 4; }

cudaError_t wavefront_alloc(WavefrontBuffers** out, int n) {
    WavefrontBuffers* b = new WavefrontBuffers();
    b->n = n;
    cudaError_t e = cudaSuccess;
    auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A((void**)&b->ray_o, (size_t)n * 16); A((void**)&b->ray_d, (size_t)n * 16); A((void**)&b->thr, (size_t)n * 16);
    A((void**)&b->hit, (size_t)n * 16); A((void**)&b->acc, (size_t)n * 16); A((void**)&b->rng, (size_t)n * 16);
    for (int k = 0; k < 2; k++) A((void**)&b->q_active[k], (size_t)n * 4);
    for (int k = 0; k < 4; k++) A((void**)&b->q_class[k], (size_t)n * 4);
    A((void**)&b->counters, C_N * 4);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&b->h_counters, C_N * 4);
    if (e != cudaSuccess) { wavefront_free(b); *out = nullptr; return e; }
    *out = b;
    return cudaSuccess;
}
void wavefront_free(WavefrontBuffers* b) {
    if (!b) return;
    cudaFree(b->ray_o); cudaFree(b->ray_d); cudaFree(b->thr); cudaFree(b->hit); cudaFree(b->acc); cudaFree(b->rng);
    for (int k = 0; k < 2; k++) cudaFree(b->q_active[k]);
    for (int k = 0; k < 4; k++) cudaFree(b->q_class[k]);
    cudaFree(b->counters);
    if (b->h_counters) cudaFreeHost(b->h_counters);
    delete b;
}

struct WfParams {
    FrameParams f;
    int S, n_slots;
    float4 *ray_o, *ray_d, *thr, *hit, *acc; uint4* rng;
    uint32_t *q_in, *q_out; uint32_t* q_class[4];
    uint32_t* counters; int c_in, c_out;
};

// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void queue_push(uint32_t* q, uint32_t* count, bool pred, uint32_t value) {
    const unsigned full = 0xffffffffu;
    unsigned m = __ballot_sync(full, pred);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(full, base, leader);
    if (pred) q[base + __popc(m & ((1u << lane) - 1u))] = value;
}

// persistent warps: 32-item chunks from an atomic head; `body(item, valid)` runs warp-convergent
template <class F>
__device__ __forceinline__ void for_each_item(const uint32_t* q, uint32_t count, uint32_t* head, F body) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(head, 32u);
        base = __shfl_sync(full, base, 0);
        if (base >= count) break;
        uint32_t i = base + lane;
        bool valid = i < count;
        body(valid ? (q ? q[i] : i) : 0u, valid);
    }
}

struct SlotState { Path path; Rng g; uint32_t k_next; };

__device__ __forceinline__ void rng_restore(Rng& g, const FrameParams& f, uint32_t pixel, uint4 r) {
    g.k0 = f.seed; g.k1 = f.frame; g.pixel = pixel; g.sample = r.w; g.block = r.x;             // block-aligned stream: no buffered words
}
__device__ __forceinline__ void load_slot(const WfParams& P, uint32_t slot, SlotState& s) {
    float4 o = P.ray_o[slot], d = P.ray_d[slot], t = P.thr[slot]; uint4 r = P.rng[slot];
    s.path.ray.o = mk3(o.x, o.y, o.z); s.path.ray.tm = o.w; s.path.ray.d = mk3(d.x, d.y, d.z);
    s.path.thr = mk3(t.x, t.y, t.z); s.path.depth = __float_as_int(t.w);
    s.k_next = r.z;
    rng_restore(s.g, P.f, slot / (uint32_t)P.S, r);
}
__device__ __forceinline__ void store_slot(const WfParams& P, uint32_t slot, const SlotState& s) {
    P.ray_o[slot] = make_float4(s.path.ray.o.x, s.path.ray.o.y, s.path.ray.o.z, s.path.ray.tm);
    P.ray_d[slot] = make_float4(s.path.ray.d.x, s.path.ray.d.y, s.path.ray.d.z, 0.f);
    P.thr[slot] = make_float4(s.path.thr.x, s.path.thr.y, s.path.thr.z, __int_as_float(s.path.depth));
    P.rng[slot] = make_uint4(s.g.block, 0u, s.k_next, s.g.sample);
}

// Brings a slot to its next traceable segment: finishes samples that are already decided (bounce limit,
// NaN ray), regenerating camera paths until one needs tracing or the slot's sample quota is used up.
// Returns true when the slot stays active.  `color_done` = a finished sample to account first.
__device__ __forceinline__ bool settle_slot(const WfParams& P, uint32_t slot, SlotState& s, bool have_path, bool done, f3 color,
                                            float4& acc, unsigned long long& n_smp) {
    const CameraParams& cam = P.f.cam;
    const int pixel = (int)(slot / (uint32_t)P.S);
    for (;;) {
        if (have_path && !done) {
            if (path_exhausted(cam, s.path, color)) done = true;
            else if (ray_is_nan(s.path.ray)) { color = mk3(NAN, NAN, NAN); done = true; }
            else return true;
        }
        if (have_path && done) {
            acc.x += color.x; acc.y += color.y; acc.z += color.z; acc.w += isnan3(color) ? 1.f : 0.f;
        }
        if ((int)s.k_next >= P.f.n_subset) return false;
        const int k = (int)s.k_next, row = k / cam.sqrt_spp;
        path_start(cam, P.f.seed, P.f.frame, pixel, k - row * cam.sqrt_spp, P.f.sj_rem + row * P.f.sj_mod, s.path, s.g);
        s.k_next += (uint32_t)P.S; n_smp++;
        have_path = true; done = false;
    }
}

__global__ void __launch_bounds__(128) wf_reset(uint32_t* c, int c_out) {
    int i = threadIdx.x;
    if (i == c_out || (i >= C_CLASS0 && i < C_N)) c[i] = 0u;
}

__global__ void __launch_bounds__(128) wf_start(const __grid_constant__ WfParams P) {
    unsigned long long n_smp = 0;
    for_each_item(nullptr, (uint32_t)P.n_slots, P.counters + C_HEAD_START, [&](uint32_t slot, bool valid) {
        bool active = false;
        if (valid) {
            SlotState s; s.k_next = slot % (uint32_t)P.S;
            s.path.depth = 0; s.path.thr = mk3(1, 1, 1); s.path.ray.o = s.path.ray.d = mk3(0, 0, 0); s.path.ray.tm = 0.f;
            rng_init(s.g, P.f.seed, P.f.frame, 0, 0);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            active = settle_slot(P, slot, s, false, false, mk3(0, 0, 0), acc, n_smp);
            P.acc[slot] = acc;
            if (active) store_slot(P, slot, s);
        }
        queue_push(P.q_out, P.counters + P.c_out, active, slot);
    });
    for (int off = 16; off > 0; off >>= 1) n_smp += __shfl_xor_sync(0xffffffffu, n_smp, off);
    if ((threadIdx.x & 31) == 0 && n_smp) atomicAdd(P.f.counters + 1, n_smp);
}

__global__ void __launch_bounds__(128) wf_extend(const __grid_constant__ WfParams P) {
    unsigned long long n_seg = 0;
    const uint32_t count = P.counters[P.c_in];
    for_each_item(P.q_in, count, P.counters + C_HEAD_EXTEND, [&](uint32_t slot, bool valid) {
        int cls = -1;
        if (valid) {
            float4 o = P.ray_o[slot], d = P.ray_d[slot];
            Ray r; r.o = mk3(o.x, o.y, o.z); r.tm = o.w; r.d = mk3(d.x, d.y, d.z);
            // canonical stream: the bounce's stage block comes first (wf_shade regenerates it from block - 1 ... see
            // rng.y), then the media draws of this segment
            Rng g; uint4 rs = P.rng[slot]; rng_restore(g, P.f, slot / (uint32_t)P.S, rs);
            const uint32_t stage_block = g.block; g.block++;
            SegHit sh;
            segment_trace<false>(P.f.sc, nullptr, 0, r, g, sh);
            n_seg++;
            P.rng[slot] = make_uint4(g.block, stage_block, rs.z, rs.w);
            P.hit[slot] = make_float4(sh.h.t, __uint_as_float(sh.h.prim), sh.h.a, sh.h.b);
            cls = material_class(P.f.sc, seghit_material(P.f.sc, sh));
        }
#pragma unroll
        for (int c = 0; c < 4; c++) queue_push(P.q_class[c], P.counters + C_CLASS0 + c, cls == c, slot);
    });
    for (int off = 16; off > 0; off >>= 1) n_seg += __shfl_xor_sync(0xffffffffu, n_seg, off);
    if ((threadIdx.x & 31) == 0 && n_seg) atomicAdd(P.f.counters, n_seg);
}

template <int kClass>
__global__ void __launch_bounds__(128) wf_shade(const __grid_constant__ WfParams P) {
    unsigned long long n_smp = 0;
    const uint32_t count = P.counters[C_CLASS0 + kClass];
    for_each_item(P.q_class[kClass], count, P.counters + C_HEAD_SHADE0 + kClass, [&](uint32_t slot, bool valid) {
        bool active = false;
        if (valid) {
            SlotState s; load_slot(P, slot, s);
            float4 hv = P.hit[slot];
            SegHit sh; sh.h.t = hv.x; sh.h.prim = __float_as_uint(hv.y); sh.h.a = hv.z; sh.h.b = hv.w;
            f3 color = mk3(0, 0, 0);
            Rng sgen = s.g; sgen.block = P.rng[slot].y;                       // the stage block reserved by wf_extend
            const R4 sb = kClass == CLASS_TERMINAL || kClass == CLASS_METAL ? R4{0.f, 0.f, 0.f, 0.f} : rng_block(sgen);
            const int st = segment_shade<kClass>(P.f.sc, P.f.cam, sh, s.path, s.g, sb, color);
            float4 acc = P.acc[slot];
            active = settle_slot(P, slot, s, true, st == SEG_DONE, color, acc, n_smp);
            P.acc[slot] = acc;
            if (active) store_slot(P, slot, s);
        }
        queue_push(P.q_out, P.counters + P.c_out, active, slot);
    });
    for (int off = 16; off > 0; off >>= 1) n_smp += __shfl_xor_sync(0xffffffffu, n_smp, off);
    if ((threadIdx.x & 31) == 0 && n_smp) atomicAdd(P.f.counters + 1, n_smp);
}

// per-pixel fixed-order sum of the S slot accumulators
__global__ void __launch_bounds__(256) wf_finalize(const float4* __restrict__ acc, int S, int n_pixels, float4* __restrict__ out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pixels) return;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < S; j++) { float4 a = acc[(size_t)p * S + j]; t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w; }
    out[p] = t;
}

cudaError_t wavefront_render(const FrameParams& f, WavefrontBuffers* b, int n_paths, int sm_count, cudaStream_t st, uint64_t* launches) {
    if (!b || b->n < 1) return cudaErrorInvalidValue;
    uint64_t nl = 0;
    int S = b->n / (f.n_pixels > 0 ? f.n_pixels : 1);
    if (S > f.n_subset) S = f.n_subset;
    if (S > 64) S = 64;
    if (S < 1) return cudaErrorInvalidValue;            // the caller sizes the buffers to at least one slot per pixel
    WfParams P; memset(&P, 0, sizeof(P));
    P.f = f; P.S = S; P.n_slots = f.n_pixels * S;
    P.ray_o = b->ray_o; P.ray_d = b->ray_d; P.thr = b->thr; P.hit = b->hit; P.acc = b->acc; P.rng = b->rng;
    for (int c = 0; c < 4; c++) P.q_class[c] = b->q_class[c];
    P.counters = b->counters;
    const int threads = 128, blocks = sm_count * 8;
    cudaError_t e = cudaMemsetAsync(b->counters, 0, C_N * 4, st);
    if (e != cudaSuccess) return e;
    int cur = 0;
    P.q_in = nullptr; P.q_out = b->q_active[cur]; P.c_in = C_ACTIVE1; P.c_out = C_ACTIVE0;
    wf_start<<<blocks, threads, 0, st>>>(P); nl++;
    for (int wave = 0;; wave++) {
        if ((wave & 3) == 0) {                            // termination check every 4 waves (empty queues make the kernels no-ops)
            e = cudaMemcpyAsync(b->h_counters, b->counters, C_N * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
            if (b->h_counters[cur == 0 ? C_ACTIVE0 : C_ACTIVE1] == 0u) break;
        }
        P.q_in = b->q_active[cur]; P.q_out = b->q_active[cur ^ 1];
        P.c_in = cur == 0 ? C_ACTIVE0 : C_ACTIVE1; P.c_out = cur == 0 ? C_ACTIVE1 : C_ACTIVE0;
        wf_reset<<<1, 32, 0, st>>>(b->counters, P.c_out);
        wf_extend<<<blocks, threads, 0, st>>>(P);
        wf_shade<CLASS_TERMINAL><<<blocks, threads, 0, st>>>(P);
        wf_shade<CLASS_DIFFUSE><<<blocks, threads, 0, st>>>(P);
        wf_shade<CLASS_METAL><<<blocks, threads, 0, st>>>(P);
        wf_shade<CLASS_DIELECTRIC><<<blocks, threads, 0, st>>>(P);
        nl += 6;
        cur ^= 1;
        if (wave > 4000000) return cudaErrorLaunchTimeout;
    }
    wf_finalize<<<(f.n_pixels + 255) / 256, 256, 0, st>>>(b->acc, S, f.n_pixels, f.accum); nl++;
    if (launches) *launches = nl;
    (void)n_paths;
    return cudaGetLastError();
}

}  // namespace mort
