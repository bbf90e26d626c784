// pool.cu — the "block wavefront": ONE persistent kernel in which every thread block runs a wavefront over a pool of
// paths that lives in its shared memory (B200: 227 KB per SM).
//
// Why (profiles/r01_scene8_fetch_bound.md, r01_source_level.md): in the warp-task megakernel (render.cu) a lane carries its
// whole path through traversal AND shading, so (1) the warps of an SM sit in every stage of the integrator at once and
// the ~100 KB of SASS they walk through thrash the 32 KB instruction cache (scene 8: 10.6 fetch stalls per issued
// instruction), (2) the traversal registers and the shading registers are live together (64-register cap -> 15.6 GB of
// spill write-back per launch), and (3) a warp's 32 lanes shade 32 unrelated materials (10 of 32 lanes live in shading).
// Here the path state (19 words per path: ray, throughput, depth, Philox counters, hit) is an SoA pool in shared memory and
// the block alternates between two phases separated by __syncthreads():
//   TRACE   warps pull 32-path chunks of the live list from a shared counter: closest hit + media (rt_core.cuh), the hit
//           goes back to the pool and the path's slot is appended to its material-class list (ballot / popc aggregation)
//   SHADE   warps pull 32-slot chunks of ONE class (terminal / diffuse / metal / dielectric): record, texture, ONB, pdf
//           and light sampling for 32 hits of the same kind; a finished sample is added to the exact frame with
//           64-bit integer reductions (accum.cuh: order-independent, so the frame equals the megakernel's bit for bit) and the
//           lane at once claims the next camera sample of the frame from a global counter and regenerates in place
// so all warps of a block execute the same few KB of code at the same time, each phase only keeps its own state in
// registers, and shading is coherent by construction.  No state ever goes to HBM: the only global traffic is the scene,
// the sample counter and the reductions into the frame.
//
// Same per-ray code, Philox stream and arithmetic as the megakernel: the two schedulers render bit-identical exact frames
// (tests/test_gpu_pool.py).
#include <cuda_runtime.h>

#include "accum.cuh"
#include "render.hpp"
#include "rt_core.cuh"
#include "trace_body.cuh"

namespace mort {

namespace {

constexpr int kPoolWords = 19;                 // 32-bit words of state per path
constexpr int kPoolLists = 3 + SHADE_CLASSES;  // 16-bit slot lists per path: 2 live lists (ping-pong), one list per shade class, the traced list

// The pool is addressed as word k of path s = smem[k * np + s] straight off the extern array, so that every access
// compiles to LDS / STS (pointers kept in a struct decayed to generic LD / ST).
extern __shared__ __align__(16) unsigned char pool_raw[];
enum { W_OX = 0, W_OY, W_OZ, W_DX, W_DY, W_DZ, W_TM,        // ray
       W_TX, W_TY, W_TZ, W_DEPTH,                          // throughput, bounces so far
       W_PIX, W_SMP, W_BLK, W_STAGE,                       // Philox counter words: pixel, sample, next block, this bounce's stage block
       W_HT, W_HPRIM, W_HA, W_HB };                        // hit
static_assert(W_HB + 1 == kPoolWords, "pool layout");
enum { L_LIVE0 = 0, L_LIVE1 = 1, L_CLS0 = 2, L_DONE = 2 + SHADE_CLASSES };   // 16-bit slot lists: 2 live lists (ping-pong), the class lists, paths whose traversal is finished
struct Pool {
    int np;
    __device__ __forceinline__ float& f(int k, unsigned s) const { return reinterpret_cast<float*>(pool_raw)[k * np + (int)s]; }
    __device__ __forceinline__ uint32_t& u(int k, unsigned s) const { return reinterpret_cast<uint32_t*>(pool_raw)[k * np + (int)s]; }
    __device__ __forceinline__ uint16_t* list(int l) const { return reinterpret_cast<uint16_t*>(pool_raw) + (2 * kPoolWords + l) * np; }
};

struct PoolCtl {
    unsigned t_n[2], t_head[2];                // live list: entries, chunk head
    unsigned c_n[2][SHADE_CLASSES], c_head[2];             // class lists: entries, chunk head
    unsigned k_head[2];                        // classify pass: chunk head
    unsigned d_n[2];                           // traced list: entries reserved
};

// warp-aggregated append to a shared-memory list (every lane of the warp must call it)
__device__ __forceinline__ void list_push(uint16_t* list, unsigned* count, bool pred, unsigned value) {
    const unsigned full = 0xffffffffu;
    const unsigned m = __ballot_sync(full, pred);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(full, base, leader);
    if (pred) list[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)value;
}

// A finished sample goes straight into the exact frame.  Black samples (most of them in the scenes with a black
// background) add nothing and cost nothing.
__device__ __forceinline__ void add_sample(const FrameParams& P, const Rng& g, f3 col) {
    unsigned long long* dst = P.accum_exact + (size_t)g.pixel * 4 + (size_t)(g.k1 - P.frame) * P.frame_words;      // frame of a batch: its own exact frame
    if (isnan3(col)) { atomicAdd(dst + 3, 1ull); return; }
    long long dr = 0, dg = 0, db = 0; unsigned long long fl = 0ull;
    fx_add(dr, fl, col.x, 20); fx_add(dg, fl, col.y, 34); fx_add(db, fl, col.z, 48);
    if (dr) atomicAdd(dst, (unsigned long long)dr);
    if (dg) atomicAdd(dst + 1, (unsigned long long)dg);
    if (db) atomicAdd(dst + 2, (unsigned long long)db);
    if (fl) atomicAdd(dst + 3, fl);
}

struct Lane { Path path; Rng g; };
// Philox identity of a path <-> its three pool words.  A launch may render a BATCH of frames (the reference's frame loop,
// mort.cu:93-120, as one launch): the frame-in-batch index rides in the upper bits of the pixel word.
__device__ __forceinline__ void rng_load(const FrameParams& P, const Pool& S, unsigned slot, Rng& g) {
    const uint32_t w = S.u(W_PIX, slot);
    g.k0 = P.seed; g.k1 = P.frame + (P.frame_shift ? (w >> P.frame_shift) : 0u); g.pixel = w & P.pix_mask;
    g.sample = S.u(W_SMP, slot); g.block = S.u(W_BLK, slot);
}

// Brings a lane to a path whose next segment must be traced: a path that reached the bounce limit or whose ray went NaN is
// finished here (camera.cuh:161-163; render.cu's path_segment does the same checks before its closest-hit query), and a
// lane without a path claims the frame's next camera sample.  Warp-convergent (ballots inside).  Returns "lane holds a live path".
// (force-inlined on purpose: one out-of-line copy was measured 6-15 % slower on every scene, profiles/r02_pool_ab.md)
__device__ __forceinline__ bool settle(const FrameParams& P, Lane& L, bool check, bool need_new, unsigned& n_smp) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    bool live = false;
    for (;;) {
        if (check) {
            f3 col;
            if (path_exhausted(P.cam, L.path, col)) { add_sample(P, L.g, col); need_new = true; }
            else if (ray_is_nan(L.path.ray)) { add_sample(P, L.g, mk3(NAN, NAN, NAN)); need_new = true; }
            else live = true;
            check = false;
        }
        const unsigned m = __ballot_sync(full, need_new);
        if (m == 0u) break;
        const int leader = __ffs(m) - 1;
        unsigned long long base = 0ull;
        if (lane == leader) base = atomicAdd(P.work64, (unsigned long long)__popc(m));
        base = __shfl_sync(full, base, leader);
        if (need_new) {
            need_new = false;
            const unsigned long long gidx = base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
            if (gidx < P.total_samples) {
                const unsigned long long fb = P.frame_shift ? gidx / P.frame_samples : 0ull;      // frame of the batch
                const unsigned long long fi = gidx - fb * P.frame_samples;
                const unsigned long long pl = fi / (unsigned long long)P.n_subset;
                const int k = (int)(fi - pl * (unsigned long long)P.n_subset);
                const int row = k / P.cam.sqrt_spp;
                const int pixel = tile_to_global((int)pl, P.band_px, P.tile_mod, P.tile_rem);
                path_start(P.cam, P.seed, P.frame + (uint32_t)fb, pixel, k - row * P.cam.sqrt_spp, P.sj_rem + row * P.sj_mod, L.path, L.g);
                n_smp++;
                check = true;
            }
        }
    }
    return live;
}

__device__ __forceinline__ void store_path(const FrameParams& P, const Pool& S, unsigned s, const Lane& L) {
    S.f(W_OX, s) = L.path.ray.o.x; S.f(W_OY, s) = L.path.ray.o.y; S.f(W_OZ, s) = L.path.ray.o.z;
    S.f(W_DX, s) = L.path.ray.d.x; S.f(W_DY, s) = L.path.ray.d.y; S.f(W_DZ, s) = L.path.ray.d.z; S.f(W_TM, s) = L.path.ray.tm;
    S.f(W_TX, s) = L.path.thr.x; S.f(W_TY, s) = L.path.thr.y; S.f(W_TZ, s) = L.path.thr.z; S.u(W_DEPTH, s) = (uint32_t)L.path.depth;
    S.u(W_PIX, s) = L.g.pixel | (P.frame_shift ? (L.g.k1 - P.frame) << P.frame_shift : 0u); S.u(W_SMP, s) = L.g.sample; S.u(W_BLK, s) = L.g.block;
}
__device__ __forceinline__ void load_ray(const Pool& S, unsigned s, Ray& r) {
    r.o = mk3(S.f(W_OX, s), S.f(W_OY, s), S.f(W_OZ, s)); r.d = mk3(S.f(W_DX, s), S.f(W_DY, s), S.f(W_DZ, s)); r.tm = S.f(W_TM, s);
}

// one 32-slot chunk of one material class
template <int kClass>
__device__ __forceinline__ bool shade_chunk(const FrameParams& P, const Pool& S, bool valid, unsigned slot, unsigned& n_smp) {
    Lane L;
    bool done = false;
    if (valid) {
        load_ray(S, slot, L.path.ray);
        L.path.thr = mk3(S.f(W_TX, slot), S.f(W_TY, slot), S.f(W_TZ, slot)); L.path.depth = (int)S.u(W_DEPTH, slot);
        rng_load(P, S, slot, L.g);
        SegHit sh; sh.h.t = S.f(W_HT, slot); sh.h.prim = S.u(W_HPRIM, slot); sh.h.a = S.f(W_HA, slot); sh.h.b = S.f(W_HB, slot);
        R4 sb = {0.f, 0.f, 0.f, 0.f};
        if (kClass == CLASS_DIFFUSE || kClass == CLASS_DIFFUSE_COLD || kClass == CLASS_DIELECTRIC) {            // the bounce's stage block, reserved by the trace phase
            Rng sg = L.g; sg.block = S.u(W_STAGE, slot);
            sb = rng_block(sg);
        }
        f3 color = mk3(0, 0, 0);
        const int st = segment_shade<kClass>(P.sc, P.cam, sh, L.path, L.g, sb, color);
        if (st == SEG_DONE) { add_sample(P, L.g, color); done = true; }
    } else {
        L.path.ray.o = L.path.ray.d = L.path.thr = mk3(0, 0, 0); L.path.ray.tm = 0.f; L.path.depth = 0;
        rng_init(L.g, P.seed, P.frame, 0, 0);
    }
    const bool live = settle(P, L, valid && !done, valid && done, n_smp);
    if (live) store_path(P, S, slot, L);
    return live;
}

// closest hit + media of one lane's path; returns its material class
template <int kLinear>
__device__ __forceinline__ int trace_lane(const FrameParams& P, const Pool& S, unsigned slot, unsigned& n_seg) {
    Ray r; load_ray(S, slot, r);
    Rng g; rng_load(P, S, slot, g);
    const uint32_t stage = g.block; g.block++;        // canonical stream: the stage block precedes the segment's media draws
    SegHit sh;
    segment_trace<false, kLinear>(P.sc, nullptr, 0, r, g, sh);
    n_seg++;
    S.u(W_BLK, slot) = g.block; S.u(W_STAGE, slot) = stage;
    S.f(W_HT, slot) = sh.h.t; S.u(W_HPRIM, slot) = sh.h.prim; S.f(W_HA, slot) = sh.h.a; S.f(W_HB, slot) = sh.h.b;
    return seghit_class(P.sc, sh);
}

// TRACE phase for tree scenes with lane refill (Aila & Laine's persistent while-while, inside one block).  With fixed
// 32-ray chunks a lane whose ray has left the tree idles until the slowest ray of its chunk is done: ncu of the phased
// kernel on scene 8 shows the traversal code — 60 % of all warp instructions — running with 12.6 of 32 lanes
// (profiles/r02_pool_kernel.md).  Here the phase is split in two:
//   trace_refill   pure traversal.  As soon as `want` lanes of a warp are idle they store their hits (4 words) and take
//                  new rays from the block's live list while the other lanes keep their traversal state in registers.
//                  The refill is deliberately tiny (no RNG, no media, no material lookup, no list pushes): it runs with
//                  only the idle lanes active, so every instruction in it is paid at low utilisation.
//   classify       everything else of the segment (media, the rare second pass, Philox block bookkeeping, material-class
//                  split) for 32 paths at a time after a barrier: uniform work, all lanes busy.
__device__ __forceinline__ void trace_refill(const FrameParams& P, const Pool& S, PoolCtl& ctl, int par, unsigned nt, int want, bool overlap) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const float tmin = 0.001f;
    StackEntry stack[MORT_STACK];
    Trav T; T.idx = T.idy = T.idz = T.oix = T.oiy = T.oiz = 0.f; T.nxo = 0; T.nyo = 4; T.nzo = 8; T.cur = MORT_CHILD_EMPTY; T.sp = 0; MORT_SET_TM(T, 0.f);
    Ray r; r.o = r.d = mk3(0, 0, 0); r.tm = 0.f;
    Hit best; best.t = INFINITY; best.prim = MORT_PRIM_NONE; best.a = best.b = 0.f;
    unsigned slot = 0; bool have = false;
    bool exhausted = false;                               // warp-uniform: the live list has no unclaimed entry left
    const uint16_t* live = S.list(L_LIVE0 + par);
    for (;;) {
        const bool fin = have && T.cur == MORT_CHILD_EMPTY;
        const unsigned busy = __ballot_sync(full, have && !fin);
        if (busy == 0u || (!exhausted && 32 - __popc(busy) >= want)) {
            const unsigned fm = __ballot_sync(full, fin);
            if (!overlap) {
                if (fin) { S.f(W_HT, slot) = best.t; S.u(W_HPRIM, slot) = best.prim; S.f(W_HA, slot) = best.a; S.f(W_HB, slot) = best.b; have = false; }
            } else if (fm != 0u) {
                // hits go to the pool, the slots to the traced list: the classify pass starts on them while other warps still trace
                if (fin) { S.f(W_HT, slot) = best.t; S.u(W_HPRIM, slot) = best.prim; S.f(W_HA, slot) = best.a; S.f(W_HB, slot) = best.b; }
                __threadfence_block();
                const int fl = __ffs(fm) - 1;
                unsigned db = 0;
                if (lane == fl) db = atomicAdd(&ctl.d_n[par], (unsigned)__popc(fm));
                db = __shfl_sync(full, db, fl);
                if (fin) { reinterpret_cast<volatile uint16_t*>(S.list(L_DONE))[db + (unsigned)__popc(fm & lt)] = (uint16_t)slot; have = false; }
            }
            if (!exhausted) {
                const unsigned need = ~busy;                            // every lane that is not traversing
                const int leader = __ffs(need) - 1;
                unsigned base = 0;
                if (lane == leader) base = atomicAdd(&ctl.t_head[par], (unsigned)__popc(need));
                base = __shfl_sync(full, base, leader);
                const unsigned i = base + (unsigned)__popc(need & lt);
                if (!have && i < nt) {
                    slot = live[i];
                    load_ray(S, slot, r);
                    best.t = INFINITY; best.prim = MORT_PRIM_NONE; best.a = best.b = 0.f;
                    trav_begin(T, r);
                    have = true;
                }
                exhausted = base + (unsigned)__popc(need) >= nt;
            }
            if (__ballot_sync(full, have) == 0u) break;
        }
        for (;;) {                                        // traversal burst: until enough lanes have run dry
#if defined(MORT_POSTPONE)
            uint32_t pend = MORT_CHILD_EMPTY;
            while (!(T.cur & MORT_LEAF_BIT)) trav_node<false>(P.sc, nullptr, 0, T, stack, tmin, best.t);
            if (T.cur != MORT_CHILD_EMPTY) {
                pend = T.cur; T.cur = trav_pop(T, stack, best.t);
                while (!(T.cur & MORT_LEAF_BIT)) trav_node<false>(P.sc, nullptr, 0, T, stack, tmin, best.t);
            }
#pragma unroll 1
            for (int k = 0; k < 2; k++) {
                const uint32_t L = k == 0 ? pend : T.cur;
                if (L != MORT_CHILD_EMPTY) leaf_intersect(P.sc, L, r, tmin, best, 0, 0x7FFFFFFF);
            }
            if (T.cur != MORT_CHILD_EMPTY) T.cur = trav_pop(T, stack, best.t);
#else
            while (!(T.cur & MORT_LEAF_BIT)) trav_node<false>(P.sc, nullptr, 0, T, stack, tmin, best.t);
            if (T.cur != MORT_CHILD_EMPTY) trav_leaf(P.sc, T, stack, r, tmin, best, 0, 0x7FFFFFFF);
#endif
            const unsigned act = __ballot_sync(full, T.cur != MORT_CHILD_EMPTY);
            if (act == 0u || (!exhausted && 32 - __popc(act) >= want)) break;
        }
    }
}
// the rest of the segment for a path whose closest surface hit is in the pool: -> material class
__device__ __forceinline__ int classify_lane(const FrameParams& P, const Pool& S, unsigned slot, unsigned& n_seg) {
    Rng g; rng_load(P, S, slot, g);
    const uint32_t stage = g.block; g.block++;
    Hit h; h.t = S.f(W_HT, slot); h.prim = S.u(W_HPRIM, slot); h.a = S.f(W_HA, slot); h.b = S.f(W_HB, slot);
    SegHit sh;
    if (P.sc.n_media > 0) {
        Ray r; load_ray(S, slot, r);
        segment_finish(P.sc, r, g, h.prim != MORT_PRIM_NONE, h, sh);
        S.f(W_HT, slot) = sh.h.t; S.u(W_HPRIM, slot) = sh.h.prim; S.f(W_HA, slot) = sh.h.a; S.f(W_HB, slot) = sh.h.b;
    } else sh.h = h;
    n_seg++;
    S.u(W_BLK, slot) = g.block; S.u(W_STAGE, slot) = stage;
    return seghit_class(P.sc, sh);
}

// kTree: built for tree scenes only (no lockstep linear scan in it: scene 1 +6 %, 1 M-sphere field +7 % over the generic kernel);
// !kTree: the GENERIC kernel (run-time flag, both closest-hit codes compiled in), used for linear-scan scenes — a kernel with the
// linear scan alone is allocated worse by ptxas and loses 10 % on Cornell (profiles/r02_pool_ab.md; round 1 saw the same).
template <int NT, int MINB, bool kTree>
__global__ void __launch_bounds__(NT, MINB) pool_kernel(const __grid_constant__ FrameParams P) {
    __shared__ PoolCtl ctl;
    const int NP = P.pool_paths;
    Pool S; S.np = NP;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned n_seg = 0, n_smp = 0;

    if (threadIdx.x < (int)(sizeof(PoolCtl) / 4)) reinterpret_cast<unsigned*>(&ctl)[threadIdx.x] = 0u;
    for (int i = threadIdx.x; i < NP; i += NT) S.list(L_DONE)[i] = (uint16_t)0xFFFFu;
    __syncthreads();
    // initial fill: every slot is free and claims a camera sample
    for (int s0 = 0; s0 < NP; s0 += NT) {
        const int s = s0 + (int)threadIdx.x;
        Lane L;
        L.path.ray.o = L.path.ray.d = L.path.thr = mk3(0, 0, 0); L.path.ray.tm = 0.f; L.path.depth = 0;
        rng_init(L.g, P.seed, P.frame, 0, 0);
        const bool live = settle(P, L, false, s < NP, n_smp);
        if (live) store_path(P, S, (unsigned)s, L);
        list_push(S.list(L_LIVE0), &ctl.t_n[0], live, (unsigned)s);
    }
    __syncthreads();

    for (int round = 0;; round++) {
        const int par = round & 1, nxt = par ^ 1;
        // the other parity's counters are idle during this trace phase (last read in the previous round, next written
        // in this round's shade phase / the next round's trace phase): reset them here
        if (threadIdx.x == 0) {
            ctl.t_n[nxt] = 0u; ctl.t_head[nxt] = 0u; ctl.c_head[nxt] = 0u; ctl.k_head[nxt] = 0u; ctl.d_n[nxt] = 0u;
            for (int c = 0; c < SHADE_CLASSES; c++) ctl.c_n[nxt][c] = 0u;
        }
        const unsigned nt = ctl.t_n[par];
        if (nt == 0u) break;                                    // block-uniform: no live path and no sample left to claim

        // ---------------- TRACE: closest hit + media for every live path, then the material-class split ----------------
        if (P.pool_refill > 0 && (kTree || !P.sc.linear) && !P.sc.two_pass && !P.sc.empty) {
            const bool overlap = P.pool_overlap != 0;
            trace_refill(P, S, ctl, par, nt, P.pool_refill, overlap);
            if (!overlap) __syncthreads();
            // classify: media + class split, 32 traced paths per warp.  No barrier: a warp that has run out of rays starts on the
            // paths already traced (in completion order) while the slowest rays of the phase are still in the tree; an entry that
            // has not been published yet is waited for (every one of the nt live paths ends up in the list).
            for (;;) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&ctl.k_head[par], 32u);
                base = __shfl_sync(full, base, 0);
                if (base >= nt) break;
                const unsigned i = base + (unsigned)lane;
                int cls = -1; unsigned slot = 0;
                if (i < nt) {
                    if (overlap) {
                        volatile uint16_t* e = reinterpret_cast<volatile uint16_t*>(S.list(L_DONE)) + i;
                        unsigned v;
                        while ((v = *e) == 0xFFFFu) __nanosleep(20);
                        slot = v;
                    } else slot = S.list(L_LIVE0 + par)[i];
                }
                if (overlap) __threadfence_block();
                if (i < nt) cls = classify_lane(P, S, slot, n_seg);
#pragma unroll
                for (int c = 0; c < SHADE_CLASSES; c++) list_push(S.list(L_CLS0 + c), &ctl.c_n[par][c], cls == c, slot);
            }
        } else
        for (;;) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&ctl.t_head[par], 32u);
            base = __shfl_sync(full, base, 0);
            if (base >= nt) break;
            const unsigned i = base + (unsigned)lane;
            const bool valid = i < nt;
            int cls = -1; unsigned slot = 0;
            if (valid) { slot = S.list(L_LIVE0 + par)[i]; cls = trace_lane<kTree ? 0 : -1>(P, S, slot, n_seg); }
#pragma unroll
            for (int c = 0; c < SHADE_CLASSES; c++) list_push(S.list(L_CLS0 + c), &ctl.c_n[par][c], cls == c, slot);
        }
        __syncthreads();

        // ---------------- SHADE: one material class per 32-slot chunk; finished lanes regenerate in place ----------------
        if (P.pool_overlap != 0 && P.pool_refill > 0 && !P.sc.linear && !P.sc.two_pass && !P.sc.empty)
            for (int i = threadIdx.x; i < NP; i += NT) S.list(L_DONE)[i] = (uint16_t)0xFFFFu;      // consumed by classify; unpublished again for the next round
        // chunk j of the phase -> (class, offset); heavy classes first so the phase's tail is made of cheap chunks
        const int order[SHADE_CLASSES] = {CLASS_DIFFUSE_COLD, CLASS_DIFFUSE, CLASS_DIELECTRIC, CLASS_METAL, CLASS_TERMINAL};
        unsigned cnt[SHADE_CLASSES], first[SHADE_CLASSES + 1];
        first[0] = 0u;
#pragma unroll
        for (int k = 0; k < SHADE_CLASSES; k++) { cnt[k] = ctl.c_n[par][order[k]]; first[k + 1] = first[k] + ((cnt[k] + 31u) >> 5); }
        for (;;) {
            unsigned j = 0;
            if (lane == 0) j = atomicAdd(&ctl.c_head[par], 1u);
            j = __shfl_sync(full, j, 0);
            if (j >= first[SHADE_CLASSES]) break;
            int k = 0; unsigned fk = 0u, ck = cnt[0]; int ok = order[0];
#pragma unroll
            for (int q = 1; q < SHADE_CLASSES; q++) if (j >= first[q]) { k = q; fk = first[q]; ck = cnt[q]; ok = order[q]; }      // constant indices: registers
            const unsigned off = (j - fk) * 32u + (unsigned)lane;
            const bool v = off < ck;
            const unsigned sl = v ? S.list(L_CLS0 + ok)[off] : 0u;
            bool live;
            if (k == 0) live = shade_chunk<CLASS_DIFFUSE_COLD>(P, S, v, sl, n_smp);
            else if (k == 1) live = shade_chunk<CLASS_DIFFUSE>(P, S, v, sl, n_smp);
            else if (k == 2) live = shade_chunk<CLASS_DIELECTRIC>(P, S, v, sl, n_smp);
            else if (k == 3) live = shade_chunk<CLASS_METAL>(P, S, v, sl, n_smp);
            else live = shade_chunk<CLASS_TERMINAL>(P, S, v, sl, n_smp);
            list_push(S.list(L_LIVE0 + nxt), &ctl.t_n[nxt], live, sl);
        }
        __syncthreads();
    }
    for (int off = 16; off > 0; off >>= 1) { n_seg += __shfl_xor_sync(full, n_seg, off); n_smp += __shfl_xor_sync(full, n_smp, off); }
    if (lane == 0) { atomicAdd(P.counters, (unsigned long long)n_seg); atomicAdd(P.counters + 1, (unsigned long long)n_smp); }
}


typedef void (*PoolFn)(const FrameParams);
template <bool kTree>
PoolFn pool_variant_of(const PoolShape& s) {
    if (s.threads >= 1024) return (PoolFn)pool_kernel<1024, 1, kTree>;
    if (s.threads >= 768) return (PoolFn)pool_kernel<768, 1, kTree>;
    if (s.threads >= 640) return (PoolFn)pool_kernel<640, 1, kTree>;
    if (s.threads >= 512) return s.min_blocks >= 2 ? (PoolFn)pool_kernel<512, 2, kTree> : (PoolFn)pool_kernel<512, 1, kTree>;
    if (s.threads >= 384) return s.min_blocks >= 2 ? (PoolFn)pool_kernel<384, 2, kTree> : (PoolFn)pool_kernel<384, 1, kTree>;
    return s.min_blocks >= 3 ? (PoolFn)pool_kernel<256, 3, kTree> : (PoolFn)pool_kernel<256, 2, kTree>;
}
#if defined(MORT_MOTION_BOUNDS)
PoolFn pool_variant(const PoolShape& s) { return pool_variant_of<true>(s); }         // motion boxes live in a tree
#elif defined(MORT_GENERAL_MEDIA)
PoolFn pool_variant(const PoolShape&) { return (PoolFn)pool_kernel<512, 2, false>; } // one shape: the generic kernel (run-time linear / tree flag)
#else
PoolFn pool_variant(const PoolShape& s) { return s.tree ? pool_variant_of<true>(s) : pool_variant_of<false>(s); }
#endif
#if defined(MORT_GENERAL_MEDIA)
int pool_threads(const PoolShape&) { return 512; }
#else
int pool_threads(const PoolShape& s) { return s.threads >= 1024 ? 1024 : s.threads >= 768 ? 768 : s.threads >= 640 ? 640 : s.threads >= 512 ? 512 : s.threads >= 384 ? 384 : 256; }
#endif
int pool_smem(const PoolShape& s) { return s.pool_paths * (kPoolWords * 4 + kPoolLists * 2); }

}  // namespace

#if defined(MORT_MOTION_BOUNDS)
#define POOL_EXPORT(name) name##_motion
#elif defined(MORT_GENERAL_MEDIA)
#define POOL_EXPORT(name) name##_stages
#else
#define POOL_EXPORT(name) name
#endif
cudaError_t POOL_EXPORT(pool_query)(const PoolShape& want, int* blocks_per_sm, int* regs, int* smem_bytes) {
    PoolFn fn = pool_variant(want);
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, (const void*)fn);
    if (e != cudaSuccess) return e;
    if (regs) *regs = fa.numRegs;
    const int smem = pool_smem(want);
    if (smem_bytes) *smem_bytes = smem;
    e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, (const void*)fn, pool_threads(want), (size_t)smem);
}

cudaError_t POOL_EXPORT(pool_launch)(const FrameParams& p, const PoolShape& shape, int blocks, cudaStream_t st) {
    PoolFn fn = pool_variant(shape);
    const int smem = pool_smem(shape);
    cudaError_t e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    void* args[] = {(void*)&p};
    return cudaLaunchKernel((const void*)fn, dim3(blocks), dim3(pool_threads(shape)), args, (size_t)smem, st);
}

#if defined(MORT_MOTION_BOUNDS)
// parity hook for scenes committed with motion boxes: the same record as render.cu's trace_kernel, through THIS unit's traversal
__global__ void __launch_bounds__(128) trace_motion_kernel(const __grid_constant__ DeviceScene sc, const float* __restrict__ rays, int n,
                                                           mhit_record* __restrict__ out, mhit_medium_probe* __restrict__ probes,
                                                           int brute_force, const int32_t* __restrict__ mat_offsets) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) trace_one(sc, rays, i, out, probes, brute_force, mat_offsets);
}
cudaError_t trace_launch_motion(const DeviceScene& sc, const float* d_rays, int n, mhit_record* d_out, mhit_medium_probe* d_probes,
                                int brute_force, const int32_t* d_mat_offsets, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    trace_motion_kernel<<<(n + 127) / 128, 128, 0, st>>>(sc, d_rays, n, d_out, d_probes, brute_force, d_mat_offsets);
    return cudaGetLastError();
}
#elif defined(MORT_GENERAL_MEDIA)
#else
// ------------------------------------------------------------------------------------------------------
// zero / resolve the exact frame over the pixels one call renders (all of them, or a rank's 8-row bands)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) zero_exact_kernel(ulonglong2* __restrict__ ex, int n_local, int band_px, int mod, int rem) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    const size_t g = (size_t)tile_to_global(i, band_px, mod, rem);
    ex[2 * g] = make_ulonglong2(0ull, 0ull); ex[2 * g + 1] = make_ulonglong2(0ull, 0ull);
}
__global__ void __launch_bounds__(256) resolve_exact_tiles_kernel(const ulonglong2* __restrict__ ex, int n_local, int band_px, int mod, int rem, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    const size_t g = (size_t)tile_to_global(i, band_px, mod, rem);
    const ulonglong2 a = ex[2 * g], b = ex[2 * g + 1];
    out[g] = fx_resolve((long long)a.x, (long long)a.y, (long long)b.x, b.y);
}
cudaError_t zero_exact_launch(unsigned long long* d_exact, int n_local, int band_px, int tile_mod, int tile_rem, cudaStream_t st) {
    if (n_local <= 0) return cudaSuccess;
    if (tile_mod <= 1) return cudaMemsetAsync(d_exact, 0, (size_t)n_local * 32, st);
    zero_exact_kernel<<<(n_local + 255) / 256, 256, 0, st>>>(reinterpret_cast<ulonglong2*>(d_exact), n_local, band_px, tile_mod, tile_rem);
    return cudaGetLastError();
}
cudaError_t resolve_exact_tiles_launch(const unsigned long long* d_exact, int n_local, int band_px, int tile_mod, int tile_rem, float4* d_accum, cudaStream_t st) {
    if (n_local <= 0) return cudaSuccess;
    resolve_exact_tiles_kernel<<<(n_local + 255) / 256, 256, 0, st>>>(reinterpret_cast<const ulonglong2*>(d_exact), n_local, band_px, tile_mod > 1 ? tile_mod : 1, tile_rem, d_accum);
    return cudaGetLastError();
}
#endif

}  // namespace mort
