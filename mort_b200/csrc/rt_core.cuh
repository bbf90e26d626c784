// rt_core.cuh — per-ray device code of mort-b200: Philox stream, exact primitive tests, 4-wide BVH
// traversal, media, record reconstruction, materials / textures / PDFs and one path segment.
//
// Everything here is a MORT_HD inline function over plain structs: the CUDA kernels in render.cu are thin
// schedulers (persistent warps, queues) around them.  The same header also compiles as ordinary C++
// (tests/hostsim) — a development aid that lets traversal and shading be debugged against the oracle on a
// machine without a GPU; it is not part of, nor reachable from, libmort_b200.so.
//
// Parity-critical arithmetic (anything that decides WHICH primitive a ray hits and at what t) is written
// with explicit single-rounding intrinsics in the contraction pattern the reference's device code compiles
// to (SURVEY.md §7 "hard parts"), so nvcc can neither fuse nor reorder it:
//   sphere::hit objects.cuh:60-88 · quad::hit objects.cuh:190-215 · translate / rotate_y objects.cuh:268-366
// Shading arithmetic uses ordinary operators.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "device_types.h"
#include "mort_scene_format.h"

#if defined(__CUDACC__)
#define MORT_HD __host__ __device__ __forceinline__
#define MORT_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define MORT_HD inline
#define MORT_HD_NOINLINE static inline
#endif

namespace mort {

// ---------------------------------------------------------------------------------------------------
// exact single-rounding ops + vector loads
// ---------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
MORT_HD float xfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
MORT_HD float xmul(float a, float b) { return __fmul_rn(a, b); }
MORT_HD float xadd(float a, float b) { return __fadd_rn(a, b); }
MORT_HD float xsub(float a, float b) { return __fsub_rn(a, b); }
MORT_HD float xdiv(float a, float b) { return __fdiv_rn(a, b); }
MORT_HD float xsqrt(float a) { return __fsqrt_rn(a); }
struct F4 { float x, y, z, w; };
MORT_HD F4 ld4(const void* p) { float4 v = __ldg(reinterpret_cast<const float4*>(p)); F4 r = {v.x, v.y, v.z, v.w}; return r; }
MORT_HD uint32_t ldu8(const uint8_t* p) { return __ldg(p); }
MORT_HD int f2i_bits(float f) { return __float_as_int(f); }
MORT_HD float i2f_bits(int i) { return __int_as_float(i); }
#else
MORT_HD float xfma(float a, float b, float c) { return fmaf(a, b, c); }    // host build uses -ffp-contract=off
MORT_HD float xmul(float a, float b) { return a * b; }
MORT_HD float xadd(float a, float b) { return a + b; }
MORT_HD float xsub(float a, float b) { return a - b; }
MORT_HD float xdiv(float a, float b) { return a / b; }
MORT_HD float xsqrt(float a) { return sqrtf(a); }
struct F4 { float x, y, z, w; };
MORT_HD F4 ld4(const void* p) { return *reinterpret_cast<const F4*>(p); }
MORT_HD uint32_t ldu8(const uint8_t* p) { return *p; }
MORT_HD int f2i_bits(float f) { int i; memcpy(&i, &f, 4); return i; }
MORT_HD float i2f_bits(int i) { float f; memcpy(&f, &i, 4); return f; }
#endif

struct f3 { float x, y, z; };
MORT_HD f3 mk3(float x, float y, float z) { f3 r = {x, y, z}; return r; }
MORT_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
MORT_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
MORT_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
MORT_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
MORT_HD f3 operator*(float t, f3 a) { return mk3(t * a.x, t * a.y, t * a.z); }
// Device code is compiled with -fmad=false (no implicit contraction: the arithmetic of a path must not depend on the kernel its
// code was inlined into), so the fused forms that matter for speed in SHADING are written out here.  fmaf is exact-by-definition on
// both the device and the host build, whatever the surrounding code.
MORT_HD float dot3(f3 a, f3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
MORT_HD f3 madd3(float t, f3 a, f3 b) { return mk3(fmaf(t, a.x, b.x), fmaf(t, a.y, b.y), fmaf(t, a.z, b.z)); }      // t * a + b
MORT_HD float len2(f3 a) { return dot3(a, a); }
MORT_HD f3 cross3(f3 u, f3 v) { return mk3(fmaf(u.y, v.z, -(u.z * v.y)), fmaf(u.z, v.x, -(u.x * v.z)), fmaf(u.x, v.y, -(u.y * v.x))); }
MORT_HD f3 unit3(f3 a);                                                           // vec3.cuh:133-136, defined after rsqrt_fast
MORT_HD bool isnan3(f3 a) { return a.x != a.x || a.y != a.y || a.z != a.z; }
// exact forms (reference contraction: e2*f2 fused last, first product fused onto the plain middle one)
MORT_HD float xdot(f3 a, f3 b) { return xfma(a.z, b.z, xfma(a.x, b.x, xmul(a.y, b.y))); }
MORT_HD f3 xsub3(f3 a, f3 b) { return mk3(xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)); }
MORT_HD f3 xadd3(f3 a, f3 b) { return mk3(xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z)); }
MORT_HD f3 xat(f3 o, f3 d, float t) { return mk3(xfma(t, d.x, o.x), xfma(t, d.y, o.y), xfma(t, d.z, o.z)); }
MORT_HD f3 xcross(f3 u, f3 v) {
    return mk3(xfma(u.y, v.z, -xmul(u.z, v.y)), xfma(u.z, v.x, -xmul(u.x, v.z)), xfma(u.x, v.y, -xmul(u.y, v.x)));
}

struct Ray { f3 o, d; float tm; };

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10, canonical stream: key = (seed, frame), counter = (pixel, sample, block, 0);
// uniforms u = (x >> 8) * 2^-24 consumed in program order (identical to oracle/mort_oracle.c).
// ---------------------------------------------------------------------------------------------------
struct Rng { uint32_t k0, k1, pixel, sample, block; };
struct U4 { uint32_t x, y, z, w; };
struct R4 { float x, y, z, w; };

#if !defined(MORT_PHILOX_UNROLL)
#define MORT_PHILOX_UNROLL 2
#endif
constexpr int kPhiloxUnroll = MORT_PHILOX_UNROLL;
// One Philox block.  Deliberately NOT inlined: a path draws at many sites, and inlined copies of the 10-round
// block (28 copies, ~2.8 k SASS instructions in the first version) pushed the megakernel out of the
// instruction cache (profiles/r01_mega_cornell_v1.md: 72 % of stall samples were `no_inst`).  The single
// out-of-line copy is unrolled by 2.  Same-box A/B of the unroll factor (profiles/r01_ab_final.jsonl, Msamples/s on
// Cornell / scene 1 / scene 8 / 1 M field): factor 10: 1585 / 2623 / 354 / 885, factor 2: 1578 / 2588 / 371 / 905 —
// the fully unrolled block costs the scenes whose hot code already overflows the instruction cache.
MORT_HD_NOINLINE U4 philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t k0, uint32_t k1) {
    uint32_t c3 = 0u;
#pragma unroll (kPhiloxUnroll)
    for (int r = 0; r < 10; r++) {
#if defined(__CUDA_ARCH__)
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    U4 o = {c0, c1, c2, c3};
    return o;
}
MORT_HD void rng_init(Rng& g, uint32_t seed, uint32_t frame, uint32_t pixel, uint32_t sample) {
    g.k0 = seed; g.k1 = frame; g.pixel = pixel; g.sample = sample; g.block = 0;
}
// The canonical stream is consumed a block at a time: every stage that draws (camera, one medium free-flight,
// one scatter stage, one rejection attempt) opens a fresh block and uses its words in order — the oracle's
// rng_align() points.  No buffered state, no refill branches: the generator is 3 live registers.
MORT_HD R4 rng_block(Rng& g) {
    U4 o = philox_block(g.pixel, g.sample, g.block, g.k0, g.k1);
    g.block++;
    const float s = 1.0f / 16777216.0f;
    R4 r = {(float)(o.x >> 8) * s, (float)(o.y >> 8) * s, (float)(o.z >> 8) * s, (float)(o.w >> 8) * s};
    return r;
}
MORT_HD int rnd_int_from(float u, int lo, int hi) {                                             // rng.cuh:30-42
    float r = 1.0f - u;                                  // curand_uniform is (0,1] = 1 - [0,1)
    r = (float)((double)r * ((double)(hi - lo) + 0.999999));   // `random *= (max - min + 0.999999)`: a double product rounded to float
    r += (float)lo;
    return (int)truncf(r);
}

MORT_HD float rsqrt_fast(float x);
// fast single-precision helpers for SHADING arithmetic only (never for hit decisions)
#if defined(__CUDA_ARCH__)
MORT_HD float rsqrt_fast(float x) { return rsqrtf(x); }
MORT_HD float div_fast(float a, float b) { return __fdividef(a, b); }
MORT_HD float rcp_fast(float a) { return __frcp_rn(a); }
#else
MORT_HD float rsqrt_fast(float x) { return 1.0f / sqrtf(x); }
MORT_HD float div_fast(float a, float b) { return a / b; }
MORT_HD float rcp_fast(float a) { return 1.0f / a; }
#endif
MORT_HD f3 unit3(f3 a) { float r = rsqrt_fast(len2(a)); return r * a; }

// ---------------------------------------------------------------------------------------------------
// instance transforms (objects.cuh:268-278, 334-366)
// ---------------------------------------------------------------------------------------------------
MORT_HD void ray_to_object(const Instance* insts, int inst, f3& o, f3& d) {
    if (inst < 0) return;
    const Instance* I = insts + inst;
    const int nops = I->nops;
#pragma unroll 1
    for (int k = 0; k < nops; k++) {
        F4 a = ld4(&I->a[k][0]);                          // {x,y,z | sin,cos,-} + op kind in the 4th word
        const int kind = f2i_bits(a.w);
        if (kind == INST_OP_TRANSLATE) { o = xsub3(o, mk3(a.x, a.y, a.z)); }
        else {
            float s = a.x, c = a.y;
            float ox = xfma(c, o.x, -xmul(s, o.z)), oz = xfma(s, o.x, xmul(c, o.z));
            float dx = xfma(c, d.x, -xmul(s, d.z)), dz = xfma(s, d.x, xmul(c, d.z));
            o.x = ox; o.z = oz; d.x = dx; d.z = dz;
        }
    }
}
MORT_HD void record_to_world(const Instance* insts, int inst, f3& p, f3& n) {
    if (inst < 0) return;
    const Instance* I = insts + inst;
    const int nops = I->nops;
#pragma unroll 1
    for (int k = nops - 1; k >= 0; k--) {
        F4 a = ld4(&I->a[k][0]);
        const int kind = f2i_bits(a.w);
        if (kind == INST_OP_TRANSLATE) { p = xadd3(p, mk3(a.x, a.y, a.z)); }
        else {
            float s = a.x, c = a.y;
            float px = xfma(c, p.x, xmul(s, p.z)), pz = xfma(c, p.z, -xmul(s, p.x));   // -s*x + c*z compiles as c*z - s*x
            float nx = xfma(c, n.x, xmul(s, n.z)), nz = xfma(c, n.z, -xmul(s, n.x));
            p.x = px; p.z = pz; n.x = nx; n.z = nz;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// exact primitive tests.  Both return the reference's t for the interval test the reference applies
// (sphere: reject root < tmin || tmax < root; quad: reject t < tmin || t > tmax) or a negative "miss" flag.
// ---------------------------------------------------------------------------------------------------
MORT_HD bool sphere_test(f3 c, float r, f3 vel, f3 o, f3 d, float tm, float tmin, float tmax, float& t_out) {
    f3 cen = mk3(xfma(tm, vel.x, c.x), xfma(tm, vel.y, c.y), xfma(tm, vel.z, c.z));
    f3 oc = xsub3(o, cen);
    float a = xdot(d, d);
    float half_b = xdot(oc, d);
    float cc = xfma(-r, r, xdot(oc, oc));
    float disc = xfma(half_b, half_b, -xmul(a, cc));
    if (disc < 0) return false;
    float sq = xsqrt(disc);
    float root = xdiv(xsub(-half_b, sq), a);
    if (root < tmin || tmax < root) {
        root = xdiv(xadd(-half_b, sq), a);
        if (root < tmin || tmax < root) return false;
    }
    t_out = root;
    return true;
}
MORT_HD bool quad_test(F4 nD, const float* rec /* QuadRec rows 1..4 */, f3 o, f3 d, float tmin, float tmax, float& t_out, float& alpha, float& beta) {
    f3 n = mk3(nD.x, nD.y, nD.z);
    float denom = xdot(n, d);
    if (fabsf(denom) <= 1e-8f) return false;          // == ((double)fabsf(denom) < 1e-8): 1e-8f is the largest float below 1e-8
    const float num = xsub(nD.w, xdot(n, o));
    // a ray leaving a surface has its origin on that surface's plane: num ~ 0, and the IEEE division would take its
    // slow path.  |num| < tmin/2 * |denom| implies |t| < tmin, i.e. the reference rejects it (t < t_min) as well.
    // (Deciding every out-of-range plane with an approximate quotient first was measured and rejected: Cornell
    // 1584 -> 1499 Msamples/s, profiles/r01_ab_variants3.jsonl — the lanes that pass pay for both divisions.)
    if (tmin > 0.f && fabsf(num) < 0.5f * tmin * fabsf(denom)) return false;
    float t = xdiv(num, denom);
    if (t < tmin || t > tmax) return false;
    F4 Q = ld4(rec);
    f3 P = xat(o, d, t);
    f3 hp = xsub3(P, mk3(Q.x, Q.y, Q.z));
    F4 U = ld4(rec + 4), Vv = ld4(rec + 8), W = ld4(rec + 12);
    f3 w = mk3(W.x, W.y, W.z);
    float al = xdot(w, xcross(hp, mk3(Vv.x, Vv.y, Vv.z)));
    float be = xdot(w, xcross(mk3(U.x, U.y, U.z), hp));
    if ((al < 0) || (al > 1) || (be < 0) || (be > 1)) return false;
    t_out = t; alpha = al; beta = be;
    return true;
}

// ---------------------------------------------------------------------------------------------------
// closest hit over the 4-wide BVH
// ---------------------------------------------------------------------------------------------------
#define MORT_STACK 48
#if !defined(__CUDA_ARCH__) && defined(MORT_HOST_COUNTERS)      // tests/hostsim only: traversal work per query, to compare tree builders on the CPU
struct HostCounters { unsigned long long queries, node_steps, leaf_visits, prim_tests, pushes, push_at[MORT_STACK]; };
static HostCounters g_host_counters = {};
#define MORT_COUNT(field, n) (g_host_counters.field += (n))
#else
#define MORT_COUNT(field, n) ((void)0)
#endif
struct alignas(8) StackEntry { uint32_t child; float t; };
#define MORT_PRIM_NONE 0xFFFFFFFFu

struct Hit { float t; uint32_t prim; float a, b; };      // prim = leaf-word type bits | record index ; (a,b) = quad alpha/beta

MORT_HD int prim_order(const DeviceScene& sc, uint32_t prim) {
    uint32_t i = prim & 0x07FFFFFFu;
    if (prim & MORT_LEAF_QUAD_BIT) { F4 v = ld4(reinterpret_cast<const float*>(sc.quads + i) + 12); return f2i_bits(v.w); }
    F4 v = ld4(sc.sphere_info + i); return f2i_bits(v.z);
}

// Candidate acceptance = the reference's sequential rule restated order-independently: nearest t wins; on
// bit-equal t the primitive the reference visits LATER wins (App. A-Q3).  order_lo/order_hi restrict the
// candidates to a visit-order window (used only when media and top-level lists coexist).
MORT_HD void consider(const DeviceScene& sc, Hit& best, float t, uint32_t prim, float a, float b, int order_lo, int order_hi) {
    if (order_lo > 0 || order_hi < 0x7FFFFFFF) { int o = prim_order(sc, prim); if (o < order_lo || o >= order_hi) return; }
    if (t == best.t && best.prim != MORT_PRIM_NONE) { if (prim_order(sc, prim) < prim_order(sc, best.prim)) return; }
    best.t = t; best.prim = prim; best.a = a; best.b = b;
}

MORT_HD void leaf_intersect(const DeviceScene& sc, uint32_t w, const Ray& r, float tmin, Hit& best, int order_lo, int order_hi) {
    uint32_t first = w & 0x07FFFFFFu; int count = (int)((w >> 27) & 7u) + 1;
    int cur_inst = -1; f3 o = r.o, d = r.d;               // neighbours in a leaf usually share their instance: transform the ray once
    if (w & MORT_LEAF_QUAD_BIT) {
#pragma unroll 1
        for (int i = 0; i < count; i++) {
            const float* q = reinterpret_cast<const float*>(sc.quads + first + i);
            F4 nD = ld4(q); F4 Qi = ld4(q + 4);
            const int inst = f2i_bits(Qi.w);
            if (inst != cur_inst) { o = r.o; d = r.d; ray_to_object(sc.instances, inst, o, d); cur_inst = inst; }
            float t, al, be;
            if (quad_test(nD, q + 4, o, d, tmin, best.t, t, al, be))
                consider(sc, best, t, MORT_LEAF_BIT | MORT_LEAF_QUAD_BIT | (first + i), al, be, order_lo, order_hi);
        }
    } else {
#pragma unroll 1
        for (int i = 0; i < count; i++) {
            const float* s = reinterpret_cast<const float*>(sc.spheres + first + i);
            F4 c = ld4(s), v = ld4(s + 4);
            const int inst = f2i_bits(v.w);
            if (inst != cur_inst) { o = r.o; d = r.d; ray_to_object(sc.instances, inst, o, d); cur_inst = inst; }
            float t;
            if (sphere_test(mk3(c.x, c.y, c.z), c.w, mk3(v.x, v.y, v.z), o, d, r.tm, tmin, best.t, t))
                consider(sc, best, t, MORT_LEAF_BIT | (first + i), 0.f, 0.f, order_lo, order_hi);
        }
    }
}

// Small scenes (<= MORT_LINEAR_MAX leaves; records sorted by instance): every lane walks the same record
// sequence, so the warp stays in lockstep and the loads are warp-uniform broadcasts — for a dozen primitives
// this beats any tree (profiles/r01_mega_cornell_v2.md: in the BVH loops only half of the live lanes did work).
#define MORT_LINEAR_MAX 40
MORT_HD void closest_hit_linear(const DeviceScene& sc, const Ray& r, float tmin, Hit& best, int order_lo, int order_hi) {
    int cur_inst = -1; f3 o = r.o, d = r.d;
#pragma unroll 1
    for (int i = 0; i < sc.n_spheres; i++) {
        const float* s = reinterpret_cast<const float*>(sc.spheres + i);
        F4 c = ld4(s), v = ld4(s + 4);
        const int inst = f2i_bits(v.w);
        if (inst != cur_inst) { o = r.o; d = r.d; ray_to_object(sc.instances, inst, o, d); cur_inst = inst; }
        float t;
        if (sphere_test(mk3(c.x, c.y, c.z), c.w, mk3(v.x, v.y, v.z), o, d, r.tm, tmin, best.t, t))
            consider(sc, best, t, MORT_LEAF_BIT | (uint32_t)i, 0.f, 0.f, order_lo, order_hi);
    }
#pragma unroll 1
    for (int i = 0; i < sc.n_quads; i++) {
        const float* q = reinterpret_cast<const float*>(sc.quads + i);
        F4 nD = ld4(q), Qi = ld4(q + 4);
        const int inst = f2i_bits(Qi.w);
        if (inst != cur_inst) { o = r.o; d = r.d; ray_to_object(sc.instances, inst, o, d); cur_inst = inst; }
        float t, al, be;
        if (quad_test(nD, q + 4, o, d, tmin, best.t, t, al, be))
            consider(sc, best, t, MORT_LEAF_BIT | MORT_LEAF_QUAD_BIT | (uint32_t)i, al, be, order_lo, order_hi);
    }
}

MORT_HD float safe_rcp_dir(float d) {
    // axis-parallel rays: keep the slab arithmetic finite (inf * 0 would make NaN); 1e-20 tilts the ray by
    // far less than the box padding
    float a = fabsf(d) < 1e-20f ? (d < 0 ? -1e-20f : 1e-20f) : d;
    return 1.0f / a;
}

// Traversal of the 4-wide BVH as a resumable state machine: per-ray constants + current node + stack depth.  closest_hit()
// below runs it to completion for one ray; the block wavefront's trace phase (pool.cu) runs it in bursts and hands lanes
// whose ray has finished a new ray while their neighbours are still traversing.
#if defined(MORT_STACK_TOP)
#define MORT_SET_TOP(T, e) ((T).top = (e))
#else
#define MORT_SET_TOP(T, e) ((void)0)
#endif
#if defined(MORT_MOTION_BOUNDS)
#define MORT_SET_TM(T, v) ((T).tm = (v))
#else
#define MORT_SET_TM(T, v) ((void)0)
#endif
struct Trav {
    float idx, idy, idz, oix, oiy, oiz;                 // 1 / d and o / d per axis: a slab plane is one FMA
    int nxo, nyo, nzo;                                   // float offsets of the ray's near-plane rows inside a Bvh4Node (far = 12/20/28 - near)
    uint32_t cur;                                        // node index | leaf word | MORT_CHILD_EMPTY = finished
    int sp;
#if defined(MORT_MOTION_BOUNDS)
    float tm;                                            // ray time
#endif
    StackEntry top;                                      // copy of stack[sp - 1]: a pop hands it out at once and starts loading the entry below
};
MORT_HD void trav_begin(Trav& T, const Ray& r) {
    T.idx = safe_rcp_dir(r.d.x); T.idy = safe_rcp_dir(r.d.y); T.idz = safe_rcp_dir(r.d.z);
    T.oix = r.o.x * T.idx; T.oiy = r.o.y * T.idy; T.oiz = r.o.z * T.idz;
    // per-ray octant: which of the node's lo/hi planes is the entry ("near") plane on each axis.  Picking the
    // rows by address replaces 12 of the 18 min/max per child (float offsets into Bvh4Node: lo rows at 0/4/8, hi at 12/16/20).
    T.nxo = T.idx < 0.f ? 12 : 0; T.nyo = T.idy < 0.f ? 16 : 4; T.nzo = T.idz < 0.f ? 20 : 8;
    T.cur = 0; T.sp = 0; MORT_SET_TM(T, r.tm); T.top.child = MORT_CHILD_EMPTY; T.top.t = 0.f;       // root
}
// entries whose entry distance is beyond the current best cannot contain a closer-or-equal hit
MORT_HD uint32_t trav_pop(Trav& T, const StackEntry* stack, float best_t) {
    // The top entry lives in registers (stack[] in local memory is its backing store), so the usual pop does not wait for
    // a local-memory load: the load it issues (the new top) is only needed by the NEXT pop (ncu, scene 8: the pop loop was
    // 3.8 % of the instructions but 9.1 % of the stall samples).
    uint32_t next = MORT_CHILD_EMPTY;
#if defined(MORT_STACK_TOP)
    while (T.sp > 0 && next == MORT_CHILD_EMPTY) {
        const StackEntry e = T.top;
        T.sp--;
        if (T.sp > 0) T.top = stack[T.sp - 1];
        if (e.t <= best_t) next = e.child;
    }
#else
    while (T.sp > 0 && next == MORT_CHILD_EMPTY) { T.sp--; const StackEntry e = stack[T.sp]; if (e.t <= best_t) next = e.child; }
#endif
    return next;
}
// one internal node: slab-test its 4 children, continue with the nearest hit child, push the others
// kStaged: the first n_staged nodes (breadth-first prefix = top levels) are read from a shared-memory copy,
// the rest from global memory, through generic 128-bit loads; otherwise every node is a read-only LDG.128.
template <bool kStaged>
MORT_HD void trav_node(const DeviceScene& sc, const Bvh4Node* staged, int n_staged, Trav& T, StackEntry* stack, float tmin, float best_t) {
    MORT_COUNT(node_steps, 1);
    const uint32_t cur = T.cur;
    const int nxo = T.nxo, nyo = T.nyo, nzo = T.nzo, fxo = 12 - nxo, fyo = 20 - nyo, fzo = 28 - nzo;
    const float idx = T.idx, idy = T.idy, idz = T.idz, oix = T.oix, oiy = T.oiy, oiz = T.oiz;
    F4 nx, ny, nz, fx, fy, fz, chf;
#if defined(__CUDA_ARCH__)
    if (kStaged) {
        // generic 128-bit loads: the pointer may be shared or global
        const float* n = reinterpret_cast<const float*>(((int)cur < n_staged ? staged : sc.nodes) + cur);
        float4 a0 = *reinterpret_cast<const float4*>(n + nxo), a1 = *reinterpret_cast<const float4*>(n + nyo), a2 = *reinterpret_cast<const float4*>(n + nzo);
        float4 a3 = *reinterpret_cast<const float4*>(n + fxo), a4 = *reinterpret_cast<const float4*>(n + fyo), a5 = *reinterpret_cast<const float4*>(n + fzo);
        float4 a6 = *reinterpret_cast<const float4*>(n + 24);
        nx = F4{a0.x, a0.y, a0.z, a0.w}; ny = F4{a1.x, a1.y, a1.z, a1.w}; nz = F4{a2.x, a2.y, a2.z, a2.w};
        fx = F4{a3.x, a3.y, a3.z, a3.w}; fy = F4{a4.x, a4.y, a4.z, a4.w}; fz = F4{a5.x, a5.y, a5.z, a5.w}; chf = F4{a6.x, a6.y, a6.z, a6.w};
    } else
#endif
    {
        const float* n = reinterpret_cast<const float*>(sc.nodes + cur);
        nx = ld4(n + nxo); ny = ld4(n + nyo); nz = ld4(n + nzo); fx = ld4(n + fxo); fy = ld4(n + fyo); fz = ld4(n + fzo); chf = ld4(n + 24);
    }
    (void)staged; (void)n_staged;
    // Motion-aware bounds (mort_build_opts.motion_bounds; the reference bounds a moving sphere by the union of its end boxes,
    // objects.cuh:46-55): sc.nodes then holds every child box at time 0 and sc.node_dt the difference to the box at time 1, and the
    // box a ray is tested against is the one at ITS time.  Centres move linearly and radii are constant, so lerp(box0, box1, t)
    // contains every primitive below at time t; parents hold unions of their children's end boxes, which contain the union of
    // the lerps.  One FMA per plane.  Compiled only into the kernels of motion.cu (a run-time branch here cost the 64-register
    // kernels 260 B of extra spills, ptxas -v), which are launched only for scenes committed with motion boxes.
#if defined(MORT_MOTION_BOUNDS)
    {
        const float* m = reinterpret_cast<const float*>(sc.node_dt + cur);
        const float tm = T.tm;
        F4 d;
        d = ld4(m + nxo); nx.x = fmaf(tm, d.x, nx.x); nx.y = fmaf(tm, d.y, nx.y); nx.z = fmaf(tm, d.z, nx.z); nx.w = fmaf(tm, d.w, nx.w);
        d = ld4(m + nyo); ny.x = fmaf(tm, d.x, ny.x); ny.y = fmaf(tm, d.y, ny.y); ny.z = fmaf(tm, d.z, ny.z); ny.w = fmaf(tm, d.w, ny.w);
        d = ld4(m + nzo); nz.x = fmaf(tm, d.x, nz.x); nz.y = fmaf(tm, d.y, nz.y); nz.z = fmaf(tm, d.z, nz.z); nz.w = fmaf(tm, d.w, nz.w);
        d = ld4(m + fxo); fx.x = fmaf(tm, d.x, fx.x); fx.y = fmaf(tm, d.y, fx.y); fx.z = fmaf(tm, d.z, fx.z); fx.w = fmaf(tm, d.w, fx.w);
        d = ld4(m + fyo); fy.x = fmaf(tm, d.x, fy.x); fy.y = fmaf(tm, d.y, fy.y); fy.z = fmaf(tm, d.z, fy.z); fy.w = fmaf(tm, d.w, fy.w);
        d = ld4(m + fzo); fz.x = fmaf(tm, d.x, fz.x); fz.y = fmaf(tm, d.y, fz.y); fz.z = fmaf(tm, d.z, fz.z); fz.w = fmaf(tm, d.w, fz.w);
    }
#endif
    float tn[4]; uint32_t cw[4] = {(uint32_t)f2i_bits(chf.x), (uint32_t)f2i_bits(chf.y), (uint32_t)f2i_bits(chf.z), (uint32_t)f2i_bits(chf.w)};
    const float nxa[4] = {nx.x, nx.y, nx.z, nx.w}, nya[4] = {ny.x, ny.y, ny.z, ny.w}, nza[4] = {nz.x, nz.y, nz.z, nz.w};
    const float fxa[4] = {fx.x, fx.y, fx.z, fx.w}, fya[4] = {fy.x, fy.y, fy.z, fy.w}, fza[4] = {fz.x, fz.y, fz.z, fz.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float tnear = fmaxf(fmaxf(fmaf(nxa[k], idx, -oix), fmaf(nya[k], idy, -oiy)), fmaxf(fmaf(nza[k], idz, -oiz), tmin));
        float tfar = fminf(fminf(fmaf(fxa[k], idx, -oix), fmaf(fya[k], idy, -oiy)), fminf(fmaf(fza[k], idz, -oiz), best_t));
        bool h = tnear <= tfar;                         // empty slots carry an inverted box (lo = +inf, hi = -inf): never hit
        tn[k] = h ? tnear : INFINITY;
        cw[k] = h ? cw[k] : MORT_CHILD_EMPTY;
    }
#define MORT_CSWAP(i, j) { bool sw = tn[j] < tn[i]; float ta = sw ? tn[j] : tn[i], tb = sw ? tn[i] : tn[j]; uint32_t ca = sw ? cw[j] : cw[i], cb = sw ? cw[i] : cw[j]; tn[i] = ta; tn[j] = tb; cw[i] = ca; cw[j] = cb; }
    // Only the nearest hit is brought to slot 0 (3 comparators); the other hits are pushed in slot order and
    // culled by their entry distance when popped.  Measured against the full 5-comparator sort
    // (profiles/r01_ab_variants2.jsonl): scene 1 +3.6 %, 1 M-sphere field +4.5 %, scene 8 +1 % — most node
    // visits hit at most two children, where the two orders coincide.
    MORT_CSWAP(0, 1) MORT_CSWAP(2, 3) MORT_CSWAP(0, 2)
#undef MORT_CSWAP
    // push far-to-near, continue with the nearest; nothing hit -> pop
    int sp = T.sp;
#pragma unroll
    for (int k = 3; k >= 1; k--)
        if (cw[k] != MORT_CHILD_EMPTY) { MORT_COUNT(pushes, 1); MORT_COUNT(push_at[sp], 1); StackEntry e; e.child = cw[k]; e.t = tn[k]; stack[sp] = e; MORT_SET_TOP(T, e); sp++; }   // depth * 3 <= MORT_STACK is checked at commit
    T.sp = sp;
    uint32_t next = cw[0];
    if (next == MORT_CHILD_EMPTY) next = trav_pop(T, stack, best_t);
    T.cur = next;
}
// one leaf: exact primitive tests, then the next stack entry
MORT_HD void trav_leaf(const DeviceScene& sc, Trav& T, const StackEntry* stack, const Ray& r, float tmin, Hit& best, int order_lo, int order_hi) {
    MORT_COUNT(leaf_visits, 1); MORT_COUNT(prim_tests, ((T.cur >> 27) & 7u) + 1);
    leaf_intersect(sc, T.cur, r, tmin, best, order_lo, order_hi);
    T.cur = trav_pop(T, stack, best.t);
}

// kLinear: -1 = sc.linear decides at run time, 1 / 0 = the caller's kernel is specialised for linear-scan / tree scenes
// (the other traversal is not even compiled into it).
template <bool kStaged, int kLinear = -1>
MORT_HD bool closest_hit(const DeviceScene& sc, const Bvh4Node* staged, int n_staged, const Ray& r, float tmin, float tmax, Hit& best,
                         int order_lo = 0, int order_hi = 0x7FFFFFFF) {
    best.t = tmax; best.prim = MORT_PRIM_NONE; best.a = best.b = 0.f;
    if (sc.empty) return false;
    if (kLinear == 1 || (kLinear < 0 && sc.linear)) { closest_hit_linear(sc, r, tmin, best, order_lo, order_hi); return best.prim != MORT_PRIM_NONE; }
    Trav T; trav_begin(T, r);
    StackEntry stack[MORT_STACK];                        // one 64-bit local store / load per push / pop
    // "while-while" (Aila & Laine): an inner loop that only descends internal nodes, then one leaf step.  The
    // loops are structured (no continue / break across them) so the compiler's convergence barriers sit at
    // the loop exits: the warp's lanes test nodes together and intersect leaves together instead of drifting
    // apart for the whole traversal (lane utilisation 6.6/32 with the former single loop).
    MORT_COUNT(queries, 1);
#if defined(MORT_POSTPONE)
    // experiment (Aila & Laine's "speculative traversal"): the first leaf a lane reaches is postponed while it descends to a second
    // one, so the lanes of a warp leave the node loop less often
    while (T.cur != MORT_CHILD_EMPTY) {
        uint32_t pend = MORT_CHILD_EMPTY;
        while (!(T.cur & MORT_LEAF_BIT)) trav_node<kStaged>(sc, staged, n_staged, T, stack, tmin, best.t);
        if (T.cur != MORT_CHILD_EMPTY) {
            pend = T.cur; T.cur = trav_pop(T, stack, best.t);
            while (!(T.cur & MORT_LEAF_BIT)) trav_node<kStaged>(sc, staged, n_staged, T, stack, tmin, best.t);
        }
#pragma unroll 1
        for (int k = 0; k < 2; k++) {
            const uint32_t L = k == 0 ? pend : T.cur;
            if (L != MORT_CHILD_EMPTY) leaf_intersect(sc, L, r, tmin, best, order_lo, order_hi);
        }
        if (T.cur != MORT_CHILD_EMPTY) T.cur = trav_pop(T, stack, best.t);
    }
#else
    while (T.cur != MORT_CHILD_EMPTY) {
        while (!(T.cur & MORT_LEAF_BIT)) trav_node<kStaged>(sc, staged, n_staged, T, stack, tmin, best.t);
        if (T.cur != MORT_CHILD_EMPTY) trav_leaf(sc, T, stack, r, tmin, best, order_lo, order_hi);
    }
#endif
    return best.prim != MORT_PRIM_NONE;
}

// brute force over every leaf record (validation only: mort_trace(..., MORT_TRACE_BRUTE_FORCE))
MORT_HD bool closest_hit_brute(const DeviceScene& sc, const Ray& r, float tmin, float tmax, Hit& best) {
    best.t = tmax; best.prim = MORT_PRIM_NONE; best.a = best.b = 0.f;
    for (int i = 0; i < sc.n_spheres; i++) leaf_intersect(sc, MORT_LEAF_BIT | (uint32_t)i, r, tmin, best, 0, 0x7FFFFFFF);
    for (int i = 0; i < sc.n_quads; i++) leaf_intersect(sc, MORT_LEAF_BIT | MORT_LEAF_QUAD_BIT | (uint32_t)i, r, tmin, best, 0, 0x7FFFFFFF);
    return best.prim != MORT_PRIM_NONE;
}

// ---------------------------------------------------------------------------------------------------
// hit record (hit_record.cuh) reconstructed once for the winning primitive
// ---------------------------------------------------------------------------------------------------
struct Record {
    f3 p, normal; float t, u, v; int mat_gid; bool front_face;
    int leaf_type, leaf_idx;      // reference (type, slot) of the primitive
    bool sphere_uv;               // u,v still to be derived from the outward normal (only image textures need them)
    f3 outward;                   // sphere outward normal in object space (for compute_uv)
};

MORT_HD void sphere_uv(f3 p, float& u, float& v) {                                   // objects.cuh:101-108
    float theta = acosf(-p.y);
    float phi = (float)((double)atan2f(-p.z, p.x) + 3.141592565);
    u = (float)((double)phi / (2.0 * 3.141592565));
    v = (float)((double)theta / 3.141592565);
}

MORT_HD void resolve_hit(const DeviceScene& sc, const Ray& r, const Hit& h, Record& rec) {
    uint32_t i = h.prim & 0x07FFFFFFu;
    f3 o = r.o, d = r.d;
    rec.t = h.t;
    if (h.prim & MORT_LEAF_QUAD_BIT) {
        const float* q = reinterpret_cast<const float*>(sc.quads + i);
        F4 nD = ld4(q), Qi = ld4(q + 4), U = ld4(q + 8), Vv = ld4(q + 12), W = ld4(q + 16);
        int inst = f2i_bits(Qi.w);
        ray_to_object(sc.instances, inst, o, d);
        f3 n = mk3(nD.x, nD.y, nD.z);
        rec.p = xat(o, d, h.t);
        rec.front_face = xdot(d, n) < 0;
        rec.normal = rec.front_face ? n : -n;
        rec.u = h.a; rec.v = h.b; rec.sphere_uv = false;
        rec.mat_gid = f2i_bits(U.w); rec.leaf_type = MORT_OBJ_QUAD; rec.leaf_idx = f2i_bits(W.w);
        (void)Vv;
        record_to_world(sc.instances, inst, rec.p, rec.normal);
    } else {
        const float* s = reinterpret_cast<const float*>(sc.spheres + i);
        F4 c = ld4(s), v = ld4(s + 4), info = ld4(sc.sphere_info + i);
        int inst = f2i_bits(v.w);
        ray_to_object(sc.instances, inst, o, d);
        f3 cen = mk3(xfma(r.tm, v.x, c.x), xfma(r.tm, v.y, c.y), xfma(r.tm, v.z, c.z));
        rec.p = xat(o, d, h.t);
        float rr = xdiv(1.0f, c.w);
        f3 pc = xsub3(rec.p, cen);
        f3 outward = mk3(xmul(rr, pc.x), xmul(rr, pc.y), xmul(rr, pc.z));
        rec.front_face = xdot(d, outward) < 0;
        rec.normal = rec.front_face ? outward : -outward;
        rec.outward = outward; rec.sphere_uv = true; rec.u = rec.v = 0.f;
        rec.mat_gid = f2i_bits(info.x); rec.leaf_type = MORT_OBJ_SPHERE; rec.leaf_idx = f2i_bits(info.y);
        record_to_world(sc.instances, inst, rec.p, rec.normal);
    }
}

// ---------------------------------------------------------------------------------------------------
// constant_medium (objects.cuh:396-434): boundary probes by linear scan in the reference's order
// ---------------------------------------------------------------------------------------------------
// out of line: medium_hit calls it twice, and each copy carries both exact primitive tests
MORT_HD_NOINLINE bool boundary_probe(const DeviceScene& sc, const Medium& m, Ray r, float tmin, float tmax, float& t_hit) {
    bool any = false; float closest = tmax;
    for (int k = 0; k < m.count; k++) {
        const BoundaryPrim* B = sc.boundary + m.first + k;
        F4 hd = ld4(B);
        f3 o = r.o, d = r.d;
        ray_to_object(sc.instances, f2i_bits(hd.y), o, d);
        float t;
        if (f2i_bits(hd.x) == MORT_OBJ_SPHERE) {
            F4 c = ld4(&B->a[0][0]), v = ld4(&B->a[1][0]);
            if (sphere_test(mk3(c.x, c.y, c.z), c.w, mk3(v.x, v.y, v.z), o, d, r.tm, tmin, closest, t)) { any = true; closest = t; }
        } else {
            float al, be;
            if (quad_test(ld4(&B->a[0][0]), &B->a[1][0], o, d, tmin, closest, t, al, be)) { any = true; closest = t; }
        }
    }
    t_hit = closest;
    return any;
}

// returns true when the ray scatters inside medium m before `closest`
MORT_HD bool medium_hit(const DeviceScene& sc, const Medium& m, const Ray& r, float tmin, float tmax, Rng& g, float& t_out) {
    float t1, t2;
    if (!boundary_probe(sc, m, r, -INFINITY, INFINITY, t1)) return false;
    if (!boundary_probe(sc, m, r, (float)((double)t1 + 0.0001), INFINITY, t2)) return false;
    if (t1 < tmin) t1 = tmin;
    if (t2 > tmax) t2 = tmax;
    if (t1 >= t2) return false;
    if (t1 < 0) t1 = 0;
#if defined(MORT_GENERAL_MEDIA)
    f3 mo = r.o, md = r.d;
    ray_to_object(sc.instances, m.inst, mo, md);                                  // the medium sees the ray its wrappers hand down (objects.cuh:268-366)
    float ray_length = xsqrt(xdot(md, md));
#else
    float ray_length = xsqrt(xdot(r.d, r.d));
#endif
    float inside = (t2 - t1) * ray_length;
    double hit_distance = m.neg_inv_density * (double)logf(rng_block(g).x);       // aligned draw (rng_align in the oracle)
    if (hit_distance > (double)inside) return false;
    t_out = (float)((double)t1 + hit_distance / (double)ray_length);
    return true;
}

MORT_HD_NOINLINE Hit closest_hit_second_pass(const DeviceScene& sc, Ray r, float tmin, float tmax) {
    Hit h; closest_hit<false>(sc, nullptr, 0, r, tmin, tmax, h, sc.post_media_order, 0x7FFFFFFF);
    return h;
}

// all media of the scene in array order, each clipped to the closest event so far (world.cuh:154-160)
struct MediaOut { Rng g; int med; float t; };
MORT_HD_NOINLINE MediaOut media_scan(const DeviceScene& sc, Ray r, float tmin, float closest, Rng g) {
    MediaOut o; o.med = -1; o.t = 0.f;
    for (int m = 0; m < sc.n_media; m++) {
        float t;
        if (medium_hit(sc, sc.media[m], r, tmin, closest, g, t)) { o.med = m; o.t = t; closest = t; }
    }
    o.g = g;
    return o;
}

#if defined(MORT_GENERAL_MEDIA)
// Compiled only into the kernels of stages.cu (and the test-only host build): the kernels every shipped scene runs do not carry
// this code (with it inlined next to the shipped form, ptxas spilled 268 instead of 132 bytes in the config-3 kernel).
// media anywhere in the visit order (a constant_medium inside a translate / rotate_y / list, objects.cuh:875-877): every medium is
// clipped against what world::hit has found BEFORE it reaches the medium, and the leaves it visits afterwards only win with t <= the
// medium's event — one windowed closest-hit pass per medium that has leaves behind it.  Cold (no shipped scene), out of line.
struct StagesOut { Rng g; Hit h; int med; float t; bool any; };
MORT_HD_NOINLINE StagesOut media_stages(const DeviceScene& sc, Ray r, float tmin, bool any, Hit h, Rng g) {
    StagesOut o; o.med = -1; o.t = 0.f;
    float closest = any ? h.t : INFINITY;
    for (int m = 0; m < sc.n_media; m++) {
        const Medium& M = sc.media[m];
        float t;
        if (medium_hit(sc, M, r, tmin, closest, g, t)) { o.med = m; o.t = t; closest = t; }
        if (M.after_hi > M.after_lo) {
            Hit h2; closest_hit<false>(sc, nullptr, 0, r, tmin, closest, h2, M.after_lo, M.after_hi);
            if (h2.prim != MORT_PRIM_NONE) { h = h2; any = true; o.med = -1; closest = h2.t; }
        }
    }
    o.g = g; o.h = h; o.any = any;
    return o;
}
#endif

// world::hit (world.cuh:104-171) in two stages so the megakernel and the wavefront kernels share it:
//   segment_trace  : surfaces, then media clipped to the closest surface, then top-level lists -> SegHit
//   segment_record : the hit_record of the winner (surface primitive or medium event)
#define MORT_PRIM_MEDIUM 0xFFFFFFFEu
struct SegHit { Hit h; };                         // h.prim == MORT_PRIM_NONE: miss; == MORT_PRIM_MEDIUM: medium event, h.a = medium index bits

// everything of world::hit after the surfaces: media clipped to the closest surface, then (rare) the post-media leaves
MORT_HD void segment_finish(const DeviceScene& sc, const Ray& r, Rng& g, bool any, Hit h, SegHit& out) {
    const float tmin = 0.001f;
    float closest = any ? h.t : INFINITY;
    int med = -1; float tmed = 0.f;
#if defined(MORT_GENERAL_MEDIA)
    if (sc.two_pass == 2) {
        const StagesOut so = media_stages(sc, r, tmin, any, h, g);
        g = so.g; h = so.h; any = so.any; med = so.med; tmed = so.t;
    } else
#endif
    if (sc.n_media > 0) {                            // cold for most scenes: kept out of line
        MediaOut mo = media_scan(sc, r, tmin, closest, g);
        g = mo.g; med = mo.med; tmed = mo.t;
        if (med >= 0) closest = tmed;
    }
    if (sc.two_pass == 1) {
        Hit h2 = closest_hit_second_pass(sc, r, tmin, closest);
        if (h2.prim != MORT_PRIM_NONE) { h = h2; any = true; med = -1; }
    }
    if (med >= 0) { h.t = tmed; h.prim = MORT_PRIM_MEDIUM; h.a = i2f_bits(med); h.b = 0.f; }
    else if (!any) { h.prim = MORT_PRIM_NONE; }
    out.h = h;
}
template <bool kStaged, int kLinear = -1>
MORT_HD void segment_trace(const DeviceScene& sc, const Bvh4Node* staged, int n_staged, const Ray& r, Rng& g, SegHit& out) {
    Hit h;
    // one traversal; only when media and top-level lists coexist (no shipped scene) is the visit-order window
    // narrower than everything and a second, out-of-line pass needed
    const bool any = closest_hit<kStaged, kLinear>(sc, staged, n_staged, r, 0.001f, INFINITY, h, 0, sc.two_pass ? sc.post_media_order : 0x7FFFFFFF);
    segment_finish(sc, r, g, any, h, out);
}
MORT_HD int seghit_material(const DeviceScene& sc, const SegHit& sh) {     // material gid of the winner (-1: none)
    if (sh.h.prim == MORT_PRIM_NONE) return -1;
    if (sh.h.prim == MORT_PRIM_MEDIUM) return sc.media[f2i_bits(sh.h.a)].mat_gid;
    uint32_t i = sh.h.prim & 0x07FFFFFFu;
    if (sh.h.prim & MORT_LEAF_QUAD_BIT) { F4 v = ld4(reinterpret_cast<const float*>(sc.quads + i) + 8); return f2i_bits(v.w); }
    F4 v = ld4(sc.sphere_info + i); return f2i_bits(v.x);
}
MORT_HD void segment_record(const DeviceScene& sc, const Ray& r, const SegHit& sh, Record& rec) {
    if (sh.h.prim == MORT_PRIM_MEDIUM) {                                 // objects.cuh:425-431
        int med = f2i_bits(sh.h.a);
#if defined(MORT_GENERAL_MEDIA)
        f3 mo = r.o, md = r.d;
        ray_to_object(sc.instances, sc.media[med].inst, mo, md);         // the record is made in the medium's frame and handed up through its wrappers
        rec.t = sh.h.t; rec.p = xat(mo, md, sh.h.t); rec.normal = mk3(1, 0, 0); rec.front_face = true;
        record_to_world(sc.instances, sc.media[med].inst, rec.p, rec.normal);
#else
        rec.t = sh.h.t; rec.p = xat(r.o, r.d, sh.h.t); rec.normal = mk3(1, 0, 0); rec.front_face = true;
#endif
        rec.mat_gid = sc.media[med].mat_gid; rec.u = rec.v = 0.f; rec.sphere_uv = false;
        rec.leaf_type = MORT_OBJ_CONSTANT_MEDIUM; rec.leaf_idx = sc.media[med].obj_idx;
        return;
    }
    resolve_hit(sc, r, sh.h, rec);
}

// ---------------------------------------------------------------------------------------------------
// textures (textures.cuh)
// ---------------------------------------------------------------------------------------------------
// The octave loop stays rolled: fully unrolled, the 7 octaves x 8 lattice corners were 1282 SASS instructions (20 KB)
// streaming through the instruction cache on every marble hit of scene 8 — a kernel that is instruction-fetch bound there
// (profiles/r01_scene8_fetch_bound.md; scene 8 397 -> 437 Msamples/s with everything rolled).  The 8 corners are unrolled
// again (their index bits fold into the code): with them rolled too the noise-only scene 4 lost a third of its speed.
// The corner weights are selects (di ? uu : 1 - uu), bit-identical to the reference's di*uu + (1-di)*(1-uu) for finite uu.
MORT_HD_NOINLINE float perlin_turb(const NoiseTables* N, f3 p) {                      // textures.cuh:174-196, 232-265
    double accum = 0.0, weight = 1.0;
#pragma unroll 1
    for (int oct = 0; oct < 7; oct++) {
        float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
        float u = p.x - fx, v = p.y - fy, w = p.z - fz;
        u = u * u * (3 - 2 * u); v = v * v * (3 - 2 * v); w = w * w * (3 - 2 * w);
        int i = (int)fx, j = (int)fy, k = (int)fz;
        double du = u, dv = v, dw = w;
        double uu = du * du * (3 - 2 * du), vv = dv * dv * (3 - 2 * dv), ww = dw * dw * (3 - 2 * dw);
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int di = c >> 2, dj = (c >> 1) & 1, dk = c & 1;
            int idx = (int)(ldu8(N->perm_x + ((i + di) & 255)) ^ ldu8(N->perm_y + ((j + dj) & 255)) ^ ldu8(N->perm_z + ((k + dk) & 255)));
            F4 cv = ld4(&N->ranvec[idx][0]);
            f3 wv = mk3((float)(du - di), (float)(dv - dj), (float)(dw - dk));
            float dt = xfma(cv.z, wv.z, xfma(cv.x, wv.x, xmul(cv.y, wv.y)));
            acc += (di ? uu : 1 - uu) * (dj ? vv : 1 - vv) * (dk ? ww : 1 - ww) * (double)dt;
        }
        accum += weight * (double)(float)acc;
        weight *= 0.5;
        p = 2.0f * p;
    }
    return fabsf((float)accum);
}

// image / noise / error textures: rare in the shipped scenes' hot loops, transcendental-heavy -> out of line
MORT_HD_NOINLINE f3 texture_cold(const DeviceScene& sc, int gid, int need_sphere_uv, f3 outward, float u, float v, f3 p) {
    if (need_sphere_uv) sphere_uv(outward, u, v);
    if (gid >= 0) {
        const Texture* T = sc.textures + gid;
        F4 a = ld4(T), b = ld4(reinterpret_cast<const float*>(T) + 4);
        int type = f2i_bits(a.x);
        if (type == MORT_TEX_IMAGE) {                                                 // textures.cuh:129-146
            const ImageDesc im = sc.images[f2i_bits(b.x)];
            if (im.height <= 0 || im.texels == nullptr) return mk3(0, 1, 1);
            float uc = u < 0 ? 0 : (u > 1 ? 1 : u);
            float vc = v < 0 ? 0 : (v > 1 ? 1 : v);
            float vf = (float)(1.0 - (double)vc);
            int i = (int)(uc * im.width), j = (int)(vf * im.height);
            if (j > im.height - 1) j = im.height - 1;
            if (j < 0) j = 0;
            int x0 = i * 3; int cmax = im.cols - 1;
            int xr = x0 < 0 ? 0 : (x0 > cmax ? cmax : x0), xg = x0 + 1 > cmax ? cmax : (x0 + 1 < 0 ? 0 : x0 + 1), xb = x0 + 2 > cmax ? cmax : (x0 + 2 < 0 ? 0 : x0 + 2);
            const uint8_t* row = im.texels + (size_t)j * im.cols;
            float sc255 = (float)(1.0 / 255.0);
            return mk3(sc255 * (float)ldu8(row + xr), sc255 * (float)ldu8(row + xg), sc255 * (float)ldu8(row + xb));
        }
        if (type == MORT_TEX_NOISE) {                                                 // textures.cuh:198-202
            const NoiseTables* N = sc.noises + f2i_bits(b.x);
            f3 s = N->scale * p;
            float f = (float)(1.0 + sin((double)s.z + 10.0 * (double)perlin_turb(N, s)));
            return mk3(0.5f * f, 0.5f * f, 0.5f * f);
        }
    }
    float e = (float)(((int)floorf(u * 1000.0f) % 2) == ((int)floorf(v * 1000.0f) % 2));   // textures.cuh:347-348
    return mk3(e, 0.f, e);
}

MORT_HD f3 texture_value(const DeviceScene& sc, int gid, Record& rec) {               // textures.cuh:327-349
    for (int guard = 0; guard < 8 && gid >= 0; guard++) {
        const Texture* T = sc.textures + gid;
        F4 a = ld4(T);
        int type = f2i_bits(a.x);
        if (type == MORT_TEX_SOLID) return mk3(a.y, a.z, a.w);
        if (type != MORT_TEX_CHECKER) break;
        F4 b = ld4(reinterpret_cast<const float*>(T) + 4);                            // textures.cuh:52-60
        float inv = a.y;
        int xi = (int)floorf(inv * rec.p.x), yi = (int)floorf(inv * rec.p.y), zi = (int)floorf(inv * rec.p.z);
        bool even = (xi + yi + zi) % 2 == 0;
        gid = even ? f2i_bits(b.x) : f2i_bits(b.y);
    }
    return texture_cold(sc, gid, rec.sphere_uv ? 1 : 0, rec.outward, rec.u, rec.v, rec.p);
}

// ---------------------------------------------------------------------------------------------------
// sampling helpers (vec3.cuh:148-212, onb.cuh:41-50) and light sampling (objects.cuh:110-145, 217-235, 488-504)
// ---------------------------------------------------------------------------------------------------
struct Onb { f3 u, v, w; };
MORT_HD void onb_from_w(Onb& b, f3 w) {
    f3 uw = unit3(w);
    f3 a = (fabsf(uw.x) > 0.9f) ? mk3(0, 1, 0) : mk3(1, 0, 0);
    f3 v = unit3(cross3(uw, a));
    b.u = cross3(uw, v); b.v = v; b.w = uw;
}
MORT_HD f3 onb_local(const Onb& b, f3 a) { return madd3(a.z, b.w, madd3(a.y, b.v, a.x * b.u)); }
MORT_HD f3 random_unit_vector(Rng& g) {
    for (;;) {
        R4 b = rng_block(g);                               // one aligned block per rejection attempt
        float x = b.x * 2.0f - 1.0f, y = b.y * 2.0f - 1.0f, z = b.z * 2.0f - 1.0f;
        f3 p = mk3(x, y, z);
        if (len2(p) >= 1) continue;
        return unit3(p);
    }
}
// Shading-side arithmetic is single precision: where the reference silently promotes to double through a
// literal (2 * 3.1415926 * r1, x / 3.1415926, sqrt(1.0 - c*c) ...) the float result differs by at most an ulp,
// far below Monte-Carlo noise, and FP64 sequences + libm slow paths are what the kernel cannot afford.
MORT_HD void sincos_fast(float x, float& s, float& c) {
#if defined(__CUDA_ARCH__)
    __sincosf(x, &s, &c);
#else
    s = sinf(x); c = cosf(x);
#endif
}
MORT_HD f3 random_cosine_direction(float r1, float r2) {
    float phi = 6.2831852f * r1;                       // 2 * 3.1415926 (vec3.cuh:185)
    float sr = sqrtf(r2), sn, cs;
    sincos_fast(phi, sn, cs);
    return mk3(cs * sr, sn * sr, sqrtf(1 - r2));
}
MORT_HD f3 reflect3(f3 v, f3 n) { return madd3(-(2 * dot3(v, n)), n, v); }
MORT_HD f3 refract3(f3 uv, f3 n, float eta) {
    float cos_theta = fminf(dot3(-uv, n), 1.0f);
    f3 perp = eta * madd3(cos_theta, n, uv);
    return madd3(-sqrtf(fabsf(1.0f - len2(perp))), n, perp);
}
MORT_HD float reflectance(float cosine, float ref_idx) {
    float r0 = div_fast(1 - ref_idx, 1 + ref_idx); r0 = r0 * r0;
    float x = 1 - cosine, x2 = x * x;
    return r0 + (1 - r0) * (x2 * x2 * x);
}

MORT_HD float light_prim_pdf(const LightPrim* L, f3 origin, f3 dir) {
    F4 hd = ld4(L);
    int kind = f2i_bits(hd.x);
    if (kind == LIGHT_SPHERE) {                                                       // objects.cuh:110-122
        F4 c = ld4(&L->a[0][0]);
        float t;
        if (!sphere_test(mk3(c.x, c.y, c.z), c.w, mk3(0, 0, 0), origin, dir, 0.f, 0.001f, HUGE_VALF, t)) return 0.f;
        float cos_theta_max = sqrtf(1 - div_fast(c.w * c.w, len2(mk3(c.x, c.y, c.z) - origin)));
        float solid_angle = 6.2831852f * (1 - cos_theta_max);
        return 1.0f / solid_angle;
    }
    if (kind == LIGHT_QUAD) {                                                         // objects.cuh:217-229
        float t, al, be;
        F4 nD = ld4(&L->a[0][0]);
        if (!quad_test(nD, &L->a[1][0], origin, dir, 0.001f, HUGE_VALF, t, al, be)) return 0.f;
        float d2 = t * t * len2(dir);
        f3 n = mk3(nD.x, nD.y, nD.z);                  // |dot| is the same for the flipped normal
        float cosine = fabsf(dot3(dir, n) * rsqrt_fast(len2(dir)));
        return d2 / (cosine * hd.y);              // IEEE: cosine = 0 must give inf
    }
    return 0.f;
}
MORT_HD f3 light_prim_random(const LightPrim* L, f3 origin, float r1, float r2) {
    F4 hd = ld4(L);
    int kind = f2i_bits(hd.x);
    if (kind == LIGHT_SPHERE) {                                                       // objects.cuh:124-145
        F4 c = ld4(&L->a[0][0]);
        f3 direction = mk3(c.x, c.y, c.z) - origin;
        float d2 = len2(direction);
        Onb uvw; onb_from_w(uvw, direction);
        float z = 1 + r2 * (sqrtf(1 - div_fast(c.w * c.w, d2)) - 1);
        float phi = 6.283184f * r1;                     // 2 * 3.141592 (objects.cuh:140)
        float s = sqrtf(1 - z * z), sn, cs;
        sincos_fast(phi, sn, cs);
        return onb_local(uvw, mk3(cs * s, sn * s, z));
    }
    if (kind == LIGHT_QUAD) {                                                         // objects.cuh:231-235
        F4 Q = ld4(&L->a[1][0]), U = ld4(&L->a[2][0]), Vv = ld4(&L->a[3][0]);
        f3 p = madd3(r2, mk3(Vv.x, Vv.y, Vv.z), madd3(r1, mk3(U.x, U.y, U.z), mk3(Q.x, Q.y, Q.z)));
        return p - origin;
    }
    return mk3(1, 0, 0);
}
MORT_HD float light_pdf_value(const DeviceScene& sc, f3 origin, f3 dir) {             // objects.cuh:947-962
    if (sc.light_kind == LIGHT_SPHERE || sc.light_kind == LIGHT_QUAD) return light_prim_pdf(sc.lights, origin, dir);
    if (sc.light_kind == LIGHT_LIST) {
        float weight = 1.0f / (float)sc.n_lights, sum = 0.f;
        for (int i = 0; i < sc.n_lights; i++) sum += weight * light_prim_pdf(sc.lights + i, origin, dir);
        return sum;
    }
    return 0.f;
}
// b = the scatter stage's block; b.x was the mixture coin, so the light draws are b.y, b.z, b.w in order
MORT_HD f3 light_random(const DeviceScene& sc, f3 origin, R4 b) {                     // objects.cuh:964-979
    if (sc.light_kind == LIGHT_SPHERE || sc.light_kind == LIGHT_QUAD) return light_prim_random(sc.lights, origin, b.y, b.z);
    if (sc.light_kind == LIGHT_LIST) { int k = rnd_int_from(b.y, 0, sc.n_lights - 1); return light_prim_random(sc.lights + k, origin, b.z, b.w); }
    return mk3(1, 0, 0);
}

// ---------------------------------------------------------------------------------------------------
// camera (camera.cuh:210-242) and one path segment (camera.cuh:96-159, forward form)
// ---------------------------------------------------------------------------------------------------
MORT_HD void camera_ray(const CameraParams& c, int x, int y, int s_i, int s_j, Rng& g, Ray& r) {
    // camera.cuh:237-241 + 213 evaluate these in double; every step is a single rounding of an exactly
    // representable double result, so the float forms below give the same bits
    R4 b = rng_block(g);                                  // jitter x, jitter y, time (no lens) in one block
    float ox = xsub(xmul(xadd((float)s_i, b.x), c.recip_sqrt_spp), 0.5f);
    float oy = xsub(xmul(xadd((float)s_j, b.y), c.recip_sqrt_spp), 0.5f);
    float tm = b.z;
    float tu = xadd((float)x, ox), tv = xadd((float)y, oy);
    f3 ps = mk3(xfma(tv, c.dv[0], xfma(tu, c.du[0], c.pixel00[0])), xfma(tv, c.dv[1], xfma(tu, c.du[1], c.pixel00[1])),
                xfma(tv, c.dv[2], xfma(tu, c.du[2], c.pixel00[2])));
    f3 origin = mk3(c.center[0], c.center[1], c.center[2]);
    if (!(c.defocus_angle <= 0)) {
        float dx, dy;
        for (;;) { R4 d = rng_block(g); dx = d.x * 2.0f - 1.0f; dy = d.y * 2.0f - 1.0f; tm = d.z; if (dx * dx + dy * dy < 1) break; }
        origin = mk3(xfma(dy, c.defocus_v[0], xfma(dx, c.defocus_u[0], c.center[0])), xfma(dy, c.defocus_v[1], xfma(dx, c.defocus_u[1], c.center[1])),
                     xfma(dy, c.defocus_v[2], xfma(dx, c.defocus_u[2], c.center[2])));
    }
    r.o = origin; r.d = xsub3(ps, origin); r.tm = tm;
}

// Path state.  The reference stores (A, E, spdf, pdf) per bounce and unwinds L = E + A*spdf*L/pdf backwards
// (camera.cuh:137-173).  Only non-scattering hits emit, so E_i = 0 on every stored bounce and the unwind is
// exactly  L = (prod_i A_i*spdf_i/pdf_i) * terminal ; the product is carried forward in `thr`.  IEEE
// inf/NaN propagate through the product the same way they do through the unwind (0*inf, 0/0 -> NaN).
struct Path { Ray ray; f3 thr; int depth; };

enum { SEG_CONTINUE = 0, SEG_DONE = 1 };

// Material classes = the wavefront's shade queues.
enum { CLASS_TERMINAL = SHADE_TERMINAL, CLASS_DIFFUSE = SHADE_DIFFUSE, CLASS_METAL = SHADE_METAL, CLASS_DIELECTRIC = SHADE_DIELECTRIC,
       CLASS_DIFFUSE_COLD = SHADE_DIFFUSE_COLD, CLASS_ANY = SHADE_CLASSES };
// the shading queue of a segment's winner, from the byte tables built at commit (flatten.cpp): one dependent load
MORT_HD int seghit_class(const DeviceScene& sc, const SegHit& sh) {
    if (sh.h.prim == MORT_PRIM_NONE) return CLASS_TERMINAL;
    if (sh.h.prim == MORT_PRIM_MEDIUM) return sc.media[f2i_bits(sh.h.a)].cls;
    const uint32_t i = sh.h.prim & 0x07FFFFFFu;
    return (int)ldu8(((sh.h.prim & MORT_LEAF_QUAD_BIT) ? sc.quad_cls : sc.sphere_cls) + i);
}
MORT_HD int material_class(const DeviceScene& sc, int mat_gid) {
    if (mat_gid < 0) return CLASS_TERMINAL;
    int type = f2i_bits(ld4(sc.materials + mat_gid).x);
    if (type == MORT_MAT_LAMBERTIAN || type == MORT_MAT_ISOTROPIC) return CLASS_DIFFUSE;
    if (type == MORT_MAT_METAL) return CLASS_METAL;
    if (type == MORT_MAT_DIELECTRIC) return CLASS_DIELECTRIC;
    return CLASS_TERMINAL;                                     // diffuse_light, unknown tags
}

// Shades one hit (or miss).  On SEG_DONE `color` is the finished sample (may be NaN/inf, like the reference's).
// kClass prunes the material switch for the per-material wavefront kernels; CLASS_ANY keeps all of it.
template <int kClass>
MORT_HD int segment_shade(const DeviceScene& sc, const CameraParams& cam, const SegHit& sh, Path& P, Rng& g, const R4& sb, f3& color) {
    if (sh.h.prim == MORT_PRIM_NONE) {
        color = P.thr * mk3(cam.background[0], cam.background[1], cam.background[2]);   // camera.cuh:154-158
        return SEG_DONE;
    }
    Record rec;
    segment_record(sc, P.ray, sh, rec);
    if (rec.mat_gid < 0) { color = P.thr * mk3(0, 0, 0); return SEG_DONE; }             // dispatchers' default: no scatter, no emission
    const Material* M = sc.materials + rec.mat_gid;
    F4 m0 = ld4(M), m1 = ld4(reinterpret_cast<const float*>(M) + 4);
    int type = f2i_bits(m0.x), tex = f2i_bits(m0.y);
    if ((kClass == CLASS_ANY || kClass == CLASS_TERMINAL) && type == MORT_MAT_DIFFUSE_LIGHT) {   // materials.cuh:151-163
        f3 e = rec.front_face ? texture_value(sc, tex, rec) : mk3(0, 0, 0);
        color = P.thr * e;
        return SEG_DONE;
    }
    if ((kClass == CLASS_ANY || kClass == CLASS_METAL) && type == MORT_MAT_METAL) {      // materials.cuh:73-84
        f3 refl = reflect3(P.ray.d, rec.normal);
        float fuzz = m1.y;
        refl = madd3(fuzz, random_unit_vector(g), unit3(refl));
        P.thr = P.thr * mk3(m0.z, m0.w, m1.x);
        P.ray.o = rec.p; P.ray.d = refl; P.depth++;
        return SEG_CONTINUE;
    }
    if ((kClass == CLASS_ANY || kClass == CLASS_DIELECTRIC) && type == MORT_MAT_DIELECTRIC) {   // materials.cuh:107-130
        float ratio = rec.front_face ? m1.z : m1.y;
        f3 ud = unit3(P.ray.d);
        float cos_theta = fminf(dot3(-ud, rec.normal), 1.0f);
        float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
        bool cant_refract = (ratio * sin_theta) > 1.0f;
        f3 dir;
        if (cant_refract || reflectance(cos_theta, ratio) > sb.x) dir = reflect3(ud, rec.normal);
        else dir = refract3(ud, rec.normal, ratio);
        P.ray.o = rec.p; P.ray.d = dir; P.depth++;       // attenuation (1,1,1)
        return SEG_CONTINUE;
    }
    if (!(kClass == CLASS_ANY || kClass == CLASS_DIFFUSE || kClass == CLASS_DIFFUSE_COLD) || (type != MORT_MAT_LAMBERTIAN && type != MORT_MAT_ISOTROPIC)) {
        color = P.thr * mk3(0, 0, 0); return SEG_DONE;
    }

    // lambertian / isotropic: importance-sampled bounce (camera.cuh:115-140, pdf.cuh)
    f3 atten = texture_value(sc, tex, rec);
    const bool cosine = type == MORT_MAT_LAMBERTIAN;
    Onb uvw;
    if (cosine) onb_from_w(uvw, rec.normal);
    f3 dir; float pdf;
    const float inv4pi = (float)(1 / (4 * 3.1415926));
    const float inv_pi_a = (float)(1 / 3.1415926), inv_pi_b = (float)(1 / 3.141592565);   // pdf.cuh:48, materials.cuh:54
    if (sc.light_kind == LIGHT_NONE) {
        if (cosine) dir = onb_local(uvw, random_cosine_direction(sb.x, sb.y));
        else dir = random_unit_vector(g);
        pdf = cosine ? fmaxf(0.f, dot3(unit3(dir), uvw.w) * inv_pi_a) : inv4pi;
    } else {
        // sb = the bounce's stage block: mixture coin, then the chosen branch's draws
        if (sb.x < 0.5f) dir = light_random(sc, rec.p, sb);
        else if (cosine) dir = onb_local(uvw, random_cosine_direction(sb.y, sb.z));
        else dir = random_unit_vector(g);
        float pm = cosine ? fmaxf(0.f, dot3(unit3(dir), uvw.w) * inv_pi_a) : inv4pi;
        pdf = 0.5f * light_pdf_value(sc, rec.p, dir) + 0.5f * pm;
    }
    float spdf;
    if (cosine) { float ct = dot3(rec.normal, unit3(dir)); spdf = (ct < 0) ? 0.f : ct * inv_pi_b; }   // materials.cuh:51-55
    else spdf = inv4pi;
    // (attenuation * scattering_pdf * L) / pdf  with vec/scalar = (1/pdf) * vec  (camera.cuh:172, vec3.cuh:109-112)
    float rp = 1.0f / pdf;                                  // IEEE: pdf = 0 must give inf (NaN semantics of camera.cuh:172)
    P.thr = rp * ((spdf * atten) * P.thr);
    P.ray.o = rec.p; P.ray.d = dir; P.depth++;
    return SEG_CONTINUE;
}

// A finished path contributes thr * terminal; a path that reaches the bounce limit contributes thr * 0
// (camera.cuh:161-163; IEEE keeps inf * 0 = NaN exactly like the reference's unwind).
MORT_HD bool path_exhausted(const CameraParams& cam, const Path& P, f3& color) {
    if (P.depth < cam.bounce_limit) return false;
    color = P.thr * mk3(0, 0, 0);
    return true;
}
MORT_HD bool ray_is_nan(const Ray& r) { return isnan3(r.d) || isnan3(r.o); }

// Megakernel form: one whole segment.  `traced` is set when a closest-hit query was issued (the unit of Mrays/s).
template <bool kStaged, int kLinear = -1>
MORT_HD int path_segment(const DeviceScene& sc, const CameraParams& cam, const Bvh4Node* staged, int n_staged, Path& P, Rng& g, f3& color, bool& traced) {
    traced = false;
    if (path_exhausted(cam, P, color)) return SEG_DONE;
    if (ray_is_nan(P.ray)) { color = mk3(NAN, NAN, NAN); return SEG_DONE; }
    traced = true;
    // the bounce's stage block is generated here, by every live lane together, whether or not the shading
    // below ends up drawing from it (one convergent Philox call instead of one per divergent material branch)
    const R4 sb = rng_block(g);
    SegHit sh;
    segment_trace<kStaged, kLinear>(sc, staged, n_staged, P.ray, g, sh);
    return segment_shade<CLASS_ANY>(sc, cam, sh, P, g, sb, color);
}

MORT_HD void path_start(const CameraParams& cam, uint32_t seed, uint32_t frame, int pixel, int s_i, int s_j, Path& P, Rng& g) {
    int x = pixel % cam.width, y = pixel / cam.width;
    rng_init(g, seed, frame, (uint32_t)pixel, (uint32_t)(s_j * cam.sqrt_spp + s_i));
    camera_ray(cam, x, y, s_i, s_j, g, P.ray);
    P.thr = mk3(1, 1, 1); P.depth = 0;
}

// Camera::render's tail (camera.cuh:194-207): mean, NaN flush, gamma 2, 8-bit quantisation.
MORT_HD void tonemap_pixel(float sx, float sy, float sz, float scale, uint8_t out[4]) {
    float c[3] = {sx * scale, sy * scale, sz * scale};
    for (int k = 0; k < 3; k++) {
        float v = c[k];
        if (v != v) v = 0.0f;
        v = sqrtf(v);
        v = v < 0.0f ? 0.0f : (v > 0.999f ? 0.999f : v);
        out[k] = (uint8_t)(int)(256 * v);
    }
    out[3] = 255;
}

}  // namespace mort
