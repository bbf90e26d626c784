/* mort_oracle.c — CPU restatement (plain C) of the reference renderer's hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see mort_oracle.h): the checker, never the thing measured or shipped.
 * Parity status: PINNED against outputs of the reference itself (tests/golden/, tests/test_oracle_golden.py).
 *
 * It follows the reference function by function (citations are /root/reference/<file>:<lines>) and keeps
 * the reference's structure on purpose — linear scans in world::hit order, true recursion through
 * hitDispatch, the binary BVH with its double-precision slab test, per-bounce arrays unwound backwards —
 * because its job is to be obviously the same algorithm, not to be fast.
 *
 * Two deliberate substitutions, both dictated by BASELINE.json's north_star:
 *   * RNG: cuRAND XORWOW per-pixel state (rng.cuh) -> counter-based Philox4x32-10, key = (seed, frame),
 *     counter = (pixel, sample, block, 0), uniforms u = (x >> 8) * 2^-24 in [0,1) consumed in the reference's
 *     program order, with ALIGNMENT POINTS (rng_align: the rest of the current block is discarded, the next
 *     draw opens a new block) before every medium free-flight draw and before every attempt of a rejection
 *     loop; the (at most four) non-loop draws of a bounce's scatter stage — mixture coin, light pick, r1, r2,
 *     or the dielectric's reflectance draw — come from ONE block that is generated at the top of the bounce,
 *     before the closest-hit query, whether or not the bounce ends up using it (so all lanes of a warp can
 *     generate it together).  Every draw is therefore "word k of a known block", which is how the product
 *     indexes them statically.  The product uses the identical stream, so
 *     product-vs-oracle images agree far below Monte-Carlo noise; oracle-vs-reference agreement is statistical.
 *   * Primitive ids: hit_record (hit_record.cuh:10-17) has none; this restatement carries (type, slot) of
 *     the sphere / quad that produced the record.
 *
 * Floating point: the reference's device code is compiled with FMA contraction on.  Where a rounding can
 * change a primary-hit decision (sphere / quad / instance transforms) the contraction nvcc 12.9 produces
 * for sm_100a is written out with fmaf() and the file is built with -ffp-contract=off; the pattern is
 * "a*b + c*d + e*f -> fma(e,f, fma(a,b, c*d))", "a*b - c*d -> fma(a,b, -(c*d))", "x - a*b -> fma(-a,b,x)".
 * It is validated bit-for-bit against the reference's t values in tests/golden/.
 */
#define _GNU_SOURCE
#include "mort_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* vec3 with the device contraction patterns (vec3.cuh)                                             */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { float x, y, z; } v3;
static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(float t, v3 v) { return V(t * v.x, t * v.y, t * v.z); }
static inline v3 vdiv(v3 v, float t) { return vscale(1.0f / t, v); }                  /* vec3.cuh:109-112 */
static inline float vdot(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.x, b.x, a.y * b.y)); }   /* vec3.cuh:114-119 */
static inline float vlen2(v3 a) { return vdot(a, a); }                                /* vec3.cuh:52-55 */
static inline float vlen(v3 a) { return sqrtf(vlen2(a)); }
static inline v3 vunit(v3 a) { return vdiv(a, vlen(a)); }                             /* vec3.cuh:133-136 */
static inline v3 vcross(v3 u, v3 v) {                                                 /* vec3.cuh:121-126 */
    return V(fmaf(u.y, v.z, -(u.z * v.y)), fmaf(u.z, v.x, -(u.x * v.z)), fmaf(u.x, v.y, -(u.y * v.x)));
}
static inline v3 vfma(float t, v3 d, v3 o) { return V(fmaf(t, d.x, o.x), fmaf(t, d.y, o.y), fmaf(t, d.z, o.z)); } /* o + t*d */
static inline v3 V3P(const float* p) { return V(p[0], p[1], p[2]); }
static inline int visnan(v3 a) { return a.x != a.x || a.y != a.y || a.z != a.z; }

typedef struct { v3 o, d; float tm; } ray_t;                                          /* ray.cuh */
static inline v3 ray_at(const ray_t* r, float t) { return vfma(t, r->d, r->o); }      /* ray.cuh:16-19 */

typedef struct {                                                                      /* hit_record.cuh:10-17 */
    v3 p, normal; int mat_idx, mat_type; float t, u, v; int front_face;
    int leaf_type, leaf_idx;                       /* oracle-only: which sphere / quad */
} hitrec;
static inline void set_face_normal(hitrec* rec, const ray_t* r, v3 outward) {        /* hit_record.cuh:19-22 */
    rec->front_face = vdot(r->d, outward) < 0;
    rec->normal = rec->front_face ? outward : vneg(outward);
}

/* ------------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011; KATs in SURVEY.md App. D) and the canonical per-sample stream   */
/* ------------------------------------------------------------------------------------------------ */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
typedef struct { uint32_t key[2], ctr[4], buf[4]; int have; } rng_t;
static void rng_init(rng_t* g, uint32_t seed, uint32_t frame, uint32_t pixel, uint32_t sample) {
    g->key[0] = seed; g->key[1] = frame; g->ctr[0] = pixel; g->ctr[1] = sample; g->ctr[2] = 0; g->ctr[3] = 0; g->have = 0;
}
static float random_float(rng_t* g) {                /* replaces rng.cuh:17-22; same range [0,1) */
    if (g->have == 0) { oracle_philox4x32_10(g->ctr, g->key, g->buf); g->ctr[2]++; g->have = 4; }
    uint32_t x = g->buf[4 - g->have]; g->have--;
    return (float)(x >> 8) * (1.0f / 16777216.0f);
}
static void rng_align(rng_t* g) { g->have = 0; }    /* canonical stream: the next draw is word 0 of a fresh block */
static float random_float_range(rng_t* g, float lo, float hi) { return random_float(g) * (hi - lo) + lo; }  /* rng.cuh:25-28 */
static int random_int(rng_t* g, int lo, int hi) {    /* rng.cuh:30-42: curand_uniform is (0,1] = 1 - [0,1) */
    float r = 1.0f - random_float(g);
    r = (float)((double)r * (hi - lo + 0.999999));
    r += (float)lo;
    return (int)truncf(r);
}
void oracle_stream_uniforms(uint32_t seed, uint32_t frame, uint32_t pixel, uint32_t sample, int n, float* out) {
    rng_t g; rng_init(&g, seed, frame, pixel, sample);
    for (int i = 0; i < n; i++) out[i] = random_float(&g);
}

/* ------------------------------------------------------------------------------------------------ */
/* scene                                                                                            */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { int skip, num; const int32_t* items; } list_t;
typedef struct { int skip, n; const mscn_bvh_node* nodes; } bvh_t;
struct oracle_scene {
    unsigned char* blob;
    mscn_header h;
    const mscn_sphere* sph; const mscn_quad* quad; const mscn_translate* tr; const mscn_rotate_y* rot; const mscn_medium* med;
    list_t* lists; bvh_t* bvhs;
    const mscn_lambertian* lam; const mscn_metal* met; const mscn_dielectric* die; const mscn_diffuse_light* lig; const mscn_isotropic* iso;
    const mscn_solid* sol; const mscn_checker* chk; const mscn_image* img; const mscn_noise* noi;
    mscn_camera cam;
    uint8_t* rgb; int rgb_w, rgb_h;
};

oracle_scene* oracle_scene_load(const char* path, const uint8_t* rgb, int rgb_w, int rgb_h) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    oracle_scene* s = (oracle_scene*)calloc(1, sizeof(*s));
    s->blob = (unsigned char*)malloc(sz > 0 ? sz : 1);
    if (fread(s->blob, 1, sz, f) != (size_t)sz) { fclose(f); free(s->blob); free(s); return NULL; }
    fclose(f);
    const unsigned char* p = s->blob; const unsigned char* end = s->blob + sz;
    memcpy(&s->h, p, sizeof(s->h)); p += sizeof(s->h);
    if (s->h.magic != MSCN_MAGIC || s->h.version != MSCN_VERSION) { free(s->blob); free(s); return NULL; }
#define TAKE(field, type, n) do { s->field = (const type*)p; p += sizeof(type) * (size_t)(n); } while (0)
    TAKE(sph, mscn_sphere, s->h.n_sphere); TAKE(quad, mscn_quad, s->h.n_quad); TAKE(tr, mscn_translate, s->h.n_translate);
    TAKE(rot, mscn_rotate_y, s->h.n_rotate_y); TAKE(med, mscn_medium, s->h.n_medium);
    s->lists = (list_t*)calloc(s->h.n_list + 1, sizeof(list_t));
    for (int i = 0; i < s->h.n_list; i++) {
        const int32_t* q = (const int32_t*)p; s->lists[i].skip = q[0]; s->lists[i].num = q[1]; s->lists[i].items = q + 2;
        p += 8 + 8 * (size_t)q[1];
    }
    s->bvhs = (bvh_t*)calloc(s->h.n_bvh + 1, sizeof(bvh_t));
    for (int i = 0; i < s->h.n_bvh; i++) {
        const int32_t* q = (const int32_t*)p; s->bvhs[i].skip = q[0]; s->bvhs[i].n = q[1];
        s->bvhs[i].nodes = (const mscn_bvh_node*)(q + 2); p += 8 + sizeof(mscn_bvh_node) * (size_t)q[1];
    }
    TAKE(lam, mscn_lambertian, s->h.n_lambertian); TAKE(met, mscn_metal, s->h.n_metal); TAKE(die, mscn_dielectric, s->h.n_dielectric);
    TAKE(lig, mscn_diffuse_light, s->h.n_diffuse_light); TAKE(iso, mscn_isotropic, s->h.n_isotropic);
    TAKE(sol, mscn_solid, s->h.n_solid); TAKE(chk, mscn_checker, s->h.n_checker); TAKE(img, mscn_image, s->h.n_image);
    TAKE(noi, mscn_noise, s->h.n_noise);
#undef TAKE
    if (p + sizeof(mscn_camera) > end) { oracle_scene_free(s); return NULL; }
    memcpy(&s->cam, p, sizeof(mscn_camera));
    if (rgb && rgb_w > 0 && rgb_h > 0) {
        s->rgb = (uint8_t*)malloc((size_t)rgb_w * rgb_h * 3); memcpy(s->rgb, rgb, (size_t)rgb_w * rgb_h * 3);
        s->rgb_w = rgb_w; s->rgb_h = rgb_h;
    }
    return s;
}
void oracle_scene_free(oracle_scene* s) { if (!s) return; free(s->lists); free(s->bvhs); free(s->rgb); free(s->blob); free(s); }
void oracle_scene_camera(const oracle_scene* s, mscn_camera* out) { *out = s->cam; }
int oracle_scene_counts(const oracle_scene* s, mscn_header* out) { *out = s->h; return 0; }

/* Camera::initialize, camera.cuh:47-84 (host code: no contraction, mixed float/double as written there) */
static v3 h_scale(float t, v3 v) { return V(t * v.x, t * v.y, t * v.z); }
static float h_len(v3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
static v3 h_unit(v3 v) { return h_scale(1 / h_len(v), v); }
static v3 h_cross(v3 u, v3 v) { return V(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x); }
static float h_deg2rad(float d) { const float pi = 3.1415926535897932385f; return (float)(d * pi / 180.0); }
static void P3(float* d, v3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
static void camera_initialize(mscn_camera* c) {
    c->image_height = (int)(c->image_width / c->aspect_ratio);
    if (c->image_height < 1) c->image_height = 1;
    c->sqrt_spp = (int)sqrt((double)c->samples_per_pixel);
    c->pixel_samples_scale = (float)(1.0 / (c->sqrt_spp * c->sqrt_spp));
    c->recip_sqrt_spp = (float)(1.0 / c->sqrt_spp);
    v3 lookfrom = V3P(c->lookfrom), lookat = V3P(c->lookat), vup = V3P(c->vup);
    v3 center = lookfrom;
    float theta = h_deg2rad((float)c->vfov);
    float h = tanf(theta / 2);
    float viewport_height = 2 * h * c->focus_dist;
    double viewport_width = viewport_height * ((double)c->image_width / c->image_height);
    v3 w = h_unit(vsub(lookfrom, lookat)), u = h_unit(h_cross(vup, w)), v = h_cross(w, u);
    v3 viewport_u = h_scale((float)viewport_width, u), viewport_v = h_scale(viewport_height, vneg(v));
    v3 du = h_scale(1 / (float)c->image_width, viewport_u);
    v3 dv = h_scale(1 / (float)c->image_height, vneg(viewport_v));
    v3 ul = vadd(vsub(vsub(center, h_scale(c->focus_dist, w)), h_scale(1 / 2.0f, viewport_u)), h_scale(1 / 2.0f, viewport_v));
    v3 p00 = vadd(ul, h_scale(0.5f, vadd(du, dv)));
    float defocus_radius = c->focus_dist * tanf(h_deg2rad(c->defocus_angle / 2));
    P3(c->center, center); P3(c->pixel00_loc, p00); P3(c->pixel_delta_u, du); P3(c->pixel_delta_v, dv);
    P3(c->v, v); P3(c->u, u); P3(c->w, w);
    P3(c->defocus_disk_u, h_scale(defocus_radius, u)); P3(c->defocus_disk_v, h_scale(defocus_radius, v));
}
void oracle_camera_override(oracle_scene* s, int width, float aspect, int spp, int depth) {
    if (width > 0) s->cam.image_width = width;
    if (aspect > 0) s->cam.aspect_ratio = aspect;
    if (spp > 0) s->cam.samples_per_pixel = spp;
    if (depth > 0) s->cam.bounce_limit = depth;
    camera_initialize(&s->cam);
}

/* ------------------------------------------------------------------------------------------------ */
/* hittables (objects.cuh)                                                                          */
/* ------------------------------------------------------------------------------------------------ */
static int hit_dispatch(const oracle_scene* S, int type, int idx, const ray_t* r, float tmin, float tmax, hitrec* rec, rng_t* g);

static inline v3 sphere_center(const mscn_sphere* s, float time) {                    /* objects.cuh:90-97 */
    if (!s->moves) return V3P(s->center);
    return vfma(time, V3P(s->center_vec), V3P(s->center));
}
static void sphere_uv(v3 p, float* u, float* v) {                                     /* objects.cuh:101-108 */
    float theta = acosf(-p.y);
    float phi = (float)(atan2f(-p.z, p.x) + 3.141592565);
    *u = (float)(phi / (2.0 * 3.141592565));
    *v = (float)(theta / 3.141592565);
}
static int sphere_hit(const mscn_sphere* s, int idx, const ray_t* r, float tmin, float tmax, hitrec* rec) {   /* objects.cuh:60-88 */
    v3 cen = sphere_center(s, r->tm);
    v3 oc = vsub(r->o, cen);
    float a = vlen2(r->d);
    float half_b = vdot(oc, r->d);
    float c = fmaf(-s->radius, s->radius, vlen2(oc));
    float disc = fmaf(half_b, half_b, -(a * c));
    if (disc < 0) return 0;
    float sqrtd = sqrtf(disc);
    float root = (-half_b - sqrtd) / a;
    if (root < tmin || tmax < root) {
        root = (-half_b + sqrtd) / a;
        if (root < tmin || tmax < root) return 0;
    }
    rec->t = root;
    rec->p = ray_at(r, rec->t);
    v3 outward = vdiv(vsub(rec->p, cen), s->radius);
    set_face_normal(rec, r, outward);
    sphere_uv(outward, &rec->u, &rec->v);
    rec->mat_type = s->mat_type; rec->mat_idx = s->mat_idx;
    rec->leaf_type = MORT_OBJ_SPHERE; rec->leaf_idx = idx;
    return 1;
}
static int quad_hit(const mscn_quad* q, int idx, const ray_t* r, float tmin, float tmax, hitrec* rec) {        /* objects.cuh:190-215 */
    v3 n = V3P(q->normal);
    float denom = vdot(n, r->d);
    if ((double)fabsf(denom) < 1e-8) return 0;
    float t = (q->D - vdot(n, r->o)) / denom;
    if (t < tmin || t > tmax) return 0;
    v3 isect = ray_at(r, t);
    v3 hp = vsub(isect, V3P(q->Q));
    v3 w = V3P(q->w);
    float alpha = vdot(w, vcross(hp, V3P(q->v)));
    float beta = vdot(w, vcross(V3P(q->u), hp));
    if ((alpha < 0) || (alpha > 1) || (beta < 0) || (beta > 1)) return 0;
    rec->t = t; rec->p = isect; rec->mat_type = q->mat_type; rec->mat_idx = q->mat_idx;
    rec->u = alpha; rec->v = beta;
    set_face_normal(rec, r, n);
    rec->leaf_type = MORT_OBJ_QUAD; rec->leaf_idx = idx;
    return 1;
}
static int translate_hit(const oracle_scene* S, const mscn_translate* t, const ray_t* r, float tmin, float tmax, hitrec* rec, rng_t* g) { /* objects.cuh:268-278 */
    ray_t rr = { vsub(r->o, V3P(t->offset)), r->d, r->tm };
    if (!hit_dispatch(S, t->obj_type, t->obj_idx, &rr, tmin, tmax, rec, g)) return 0;
    rec->p = vadd(rec->p, V3P(t->offset));
    return 1;
}
static int rotate_hit(const oracle_scene* S, const mscn_rotate_y* t, const ray_t* r, float tmin, float tmax, hitrec* rec, rng_t* g) {   /* objects.cuh:334-366 */
    float c = t->cos_theta, s = t->sin_theta;
    ray_t rr = *r;
    rr.o.x = fmaf(c, r->o.x, -(s * r->o.z));
    rr.o.z = fmaf(s, r->o.x, c * r->o.z);
    rr.d.x = fmaf(c, r->d.x, -(s * r->d.z));
    rr.d.z = fmaf(s, r->d.x, c * r->d.z);
    if (!hit_dispatch(S, t->obj_type, t->obj_idx, &rr, tmin, tmax, rec, g)) return 0;
    v3 p = rec->p, n = rec->normal;
    p.x = fmaf(c, rec->p.x, s * rec->p.z);
    p.z = fmaf(c, rec->p.z, -(s * rec->p.x));      /* (-s*x + c*z) is canonicalised to c*z - s*x before contraction */
    n.x = fmaf(c, rec->normal.x, s * rec->normal.z);
    n.z = fmaf(c, rec->normal.z, -(s * rec->normal.x));
    rec->p = p; rec->normal = n;
    return 1;
}
static int medium_probe(const oracle_scene* S, const mscn_medium* m, const ray_t* r, hitrec* rec1, hitrec* rec2, rng_t* g, int* h1) { /* objects.cuh:398-406 */
    *h1 = 0;
    if (!hit_dispatch(S, m->obj_type, m->obj_idx, r, -INFINITY, INFINITY, rec1, g)) return 0;
    *h1 = 1;
    if (!hit_dispatch(S, m->obj_type, m->obj_idx, r, (float)(rec1->t + 0.0001), INFINITY, rec2, g)) return 0;
    return 1;
}
static int medium_hit(const oracle_scene* S, const mscn_medium* m, const ray_t* r, float tmin, float tmax, hitrec* rec, rng_t* g) {    /* objects.cuh:396-434 */
    hitrec rec1, rec2; int h1;
    if (!g) return 0;                              /* oracle_trace: surfaces only (the product's mort_trace has no medium loop either) */
    if (!medium_probe(S, m, r, &rec1, &rec2, g, &h1)) return 0;
    if (rec1.t < tmin) rec1.t = tmin;
    if (rec2.t > tmax) rec2.t = tmax;
    if (rec1.t >= rec2.t) return 0;
    if (rec1.t < 0) rec1.t = 0;
    float ray_length = vlen(r->d);
    float distance_inside = (rec2.t - rec1.t) * ray_length;
    rng_align(g);
    double hit_distance = m->neg_inv_density * (double)logf(random_float(g));
    if (hit_distance > (double)distance_inside) return 0;
    rec->t = (float)(rec1.t + hit_distance / ray_length);
    rec->p = ray_at(r, rec->t);
    rec->normal = V(1, 0, 0); rec->front_face = 1;
    rec->mat_type = m->mat_type; rec->mat_idx = m->mat_idx;
    rec->leaf_type = MORT_OBJ_CONSTANT_MEDIUM; rec->leaf_idx = (int)(m - S->med);
    return 1;
}
static int list_hit(const oracle_scene* S, const list_t* l, const ray_t* r, float tmin, float tmax, hitrec* rec, rng_t* g) {          /* objects.cuh:471-486 */
    hitrec tmp; int any = 0; float closest = tmax;
    tmp.u = tmp.v = 0;
    for (int i = 0; i < l->num; i++)
        if (hit_dispatch(S, l->items[2 * i], l->items[2 * i + 1], r, tmin, closest, &tmp, g)) { any = 1; closest = tmp.t; *rec = tmp; }
    return any;
}
static int hit_dispatch(const oracle_scene* S, int type, int idx, const ray_t* r, float tmin, float tmax, hitrec* rec, rng_t* g) {    /* objects.cuh:858-887 */
    switch (type) {
        case MORT_OBJ_SPHERE: return sphere_hit(&S->sph[idx], idx, r, tmin, tmax, rec);
        case MORT_OBJ_QUAD: return quad_hit(&S->quad[idx], idx, r, tmin, tmax, rec);
        case MORT_OBJ_TRANSLATE: return translate_hit(S, &S->tr[idx], r, tmin, tmax, rec, g);
        case MORT_OBJ_ROTATE_Y: return rotate_hit(S, &S->rot[idx], r, tmin, tmax, rec, g);
        case MORT_OBJ_CONSTANT_MEDIUM: return medium_hit(S, &S->med[idx], r, tmin, tmax, rec, g);
        case MORT_OBJ_HITTABLE_LIST: return list_hit(S, &S->lists[idx], r, tmin, tmax, rec, g);
    }
    return 0;
}
/* aabb::hit, aabb.cuh:37-59: the subtraction is float, the multiply double, t_min/t_max are float */
static int aabb_hit(const float* bb, const ray_t* r, float tmin, float tmax) {
    const float o[3] = { r->o.x, r->o.y, r->o.z }, d[3] = { r->d.x, r->d.y, r->d.z };
    for (int a = 0; a < 3; a++) {
        double invD = 1.0 / d[a];
        float orig = o[a];
        double t0 = (bb[2 * a] - orig) * invD;
        double t1 = (bb[2 * a + 1] - orig) * invD;
        if (invD < 0) { double aux = t0; t0 = t1; t1 = aux; }
        if (t0 > tmin) tmin = (float)t0;
        if (t1 < tmax) tmax = (float)t1;
        if (tmax <= tmin) return 0;
    }
    return 1;
}
/* bvh::hit, objects.cuh:664-723: the explicit 20-frame stack there emulates exactly this recursion */
static int bvh_node_hit(const oracle_scene* S, const bvh_t* B, int node, const ray_t* r, float tmin, float tmax, hitrec* rec, rng_t* g) {
    const mscn_bvh_node* n = &B->nodes[node];
    if (!aabb_hit(n->bbox, r, tmin, tmax)) return 0;
    if (!n->is_internal) {
        int hl = hit_dispatch(S, n->left_type, n->left_idx, r, tmin, tmax, rec, g);
        int hr = hit_dispatch(S, n->right_type, n->right_idx, r, tmin, hl ? rec->t : tmax, rec, g);
        return hl || hr;
    }
    int hl = bvh_node_hit(S, B, n->left_idx, r, tmin, tmax, rec, g);
    int hr = bvh_node_hit(S, B, n->right_idx, r, tmin, hl ? rec->t : tmax, rec, g);
    return hl || hr;
}

/* world::hit, world.cuh:104-171.  top = which top-level object produced the record. */
static int world_hit(const oracle_scene* S, const ray_t* r, float tmin, float tmax, hitrec* rec, rng_t* g, int with_media, int* top_type, int* top_idx) {
    hitrec tmp; int any = 0; float closest = tmax;
    tmp.u = tmp.v = 0;
#define ACCEPT(ty, ix) do { any = 1; closest = tmp.t; *rec = tmp; if (top_type) { *top_type = (ty); *top_idx = (ix); } } while (0)
    for (int i = 0; i < S->h.n_bvh; i++)
        if (!S->bvhs[i].skip && S->bvhs[i].n > 0 && bvh_node_hit(S, &S->bvhs[i], 0, r, tmin, closest, &tmp, g)) ACCEPT(MORT_OBJ_BVH, i);
    if (S->h.bvh_mode) return any;
    for (int i = 0; i < S->h.n_sphere; i++)
        if (!S->sph[i].skip && sphere_hit(&S->sph[i], i, r, tmin, closest, &tmp)) ACCEPT(MORT_OBJ_SPHERE, i);
    for (int i = 0; i < S->h.n_quad; i++)
        if (!S->quad[i].skip && quad_hit(&S->quad[i], i, r, tmin, closest, &tmp)) ACCEPT(MORT_OBJ_QUAD, i);
    for (int i = 0; i < S->h.n_translate; i++)
        if (!S->tr[i].skip && translate_hit(S, &S->tr[i], r, tmin, closest, &tmp, g)) ACCEPT(MORT_OBJ_TRANSLATE, i);
    for (int i = 0; i < S->h.n_rotate_y; i++)
        if (!S->rot[i].skip && rotate_hit(S, &S->rot[i], r, tmin, closest, &tmp, g)) ACCEPT(MORT_OBJ_ROTATE_Y, i);
    if (with_media)
        for (int i = 0; i < S->h.n_medium; i++)
            if (!S->med[i].skip && medium_hit(S, &S->med[i], r, tmin, closest, &tmp, g)) ACCEPT(MORT_OBJ_CONSTANT_MEDIUM, i);
    for (int i = 0; i < S->h.n_list; i++)
        if (!S->lists[i].skip && list_hit(S, &S->lists[i], r, tmin, closest, &tmp, g)) ACCEPT(MORT_OBJ_HITTABLE_LIST, i);
#undef ACCEPT
    return any;
}

int oracle_trace(const oracle_scene* S, const float* rays7, int n, mhit_record* out, mhit_medium_probe* probes) {
    rng_t g; rng_init(&g, 0, 0, 0, 0);
    for (int i = 0; i < n; i++) {
        const float* q = rays7 + 7 * (size_t)i;
        ray_t r = { V(q[0], q[1], q[2]), V(q[3], q[4], q[5]), q[6] };
        hitrec rec; memset(&rec, 0, sizeof(rec));
        int tt = -1, ti = -1;
        int h = world_hit(S, &r, 0.001f, INFINITY, &rec, NULL, 0, &tt, &ti);      /* no RNG: media nested in wrappers / lists stay out as well */
        mhit_record o; memset(&o, 0, sizeof(o));
        o.hit = h; o.leaf_type = o.leaf_idx = o.top_type = o.top_idx = -1;
        if (h) {
            o.t = rec.t; o.leaf_type = rec.leaf_type; o.leaf_idx = rec.leaf_idx; o.top_type = tt; o.top_idx = ti;
            o.mat_type = rec.mat_type; o.mat_idx = rec.mat_idx; o.front_face = rec.front_face;
            P3(o.p, rec.p); P3(o.normal, rec.normal); o.u = rec.u; o.v = rec.v;
        }
        out[i] = o;
        if (probes)
            for (int m = 0; m < S->h.n_medium; m++) {
                hitrec r1, r2; int h1 = 0; mhit_medium_probe pr = { 0, 0, 0.f, 0.f };
                int h2 = medium_probe(S, &S->med[m], &r, &r1, &r2, &g, &h1);
                if (h1) { pr.hit1 = 1; pr.t1 = r1.t; }
                if (h2) { pr.hit2 = 1; pr.t2 = r2.t; }
                probes[(size_t)i * S->h.n_medium + m] = pr;
            }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* textures (textures.cuh)                                                                          */
/* ------------------------------------------------------------------------------------------------ */
static float perlin_interp(v3 c[2][2][2], double u, double v, double w) {             /* textures.cuh:232-250 */
    double uu = u * u * (3 - 2 * u), vv = v * v * (3 - 2 * v), ww = w * w * (3 - 2 * w), accum = 0.0;
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
        v3 wv = V((float)(u - i), (float)(v - j), (float)(w - k));
        accum += (i * uu + (1 - i) * (1 - uu)) * (j * vv + (1 - j) * (1 - vv)) * (k * ww + (1 - k) * (1 - ww)) * vdot(c[i][j][k], wv);
    }
    return (float)accum;
}
static float perlin_noise(const mscn_noise* n, v3 p) {                                /* textures.cuh:174-196 */
    float u = p.x - floorf(p.x), v = p.y - floorf(p.y), w = p.z - floorf(p.z);
    u = u * u * (3 - 2 * u); v = v * v * (3 - 2 * v); w = w * w * (3 - 2 * w);
    int i = (int)floorf(p.x), j = (int)floorf(p.y), k = (int)floorf(p.z);
    v3 c[2][2][2];
    for (int di = 0; di < 2; di++) for (int dj = 0; dj < 2; dj++) for (int dk = 0; dk < 2; dk++) {
        int idx = n->perm_x[(i + di) & 255] ^ n->perm_y[(j + dj) & 255] ^ n->perm_z[(k + dk) & 255];
        c[di][dj][dk] = V3P(n->ranvec[idx]);
    }
    return perlin_interp(c, u, v, w);
}
static float perlin_turb(const mscn_noise* n, v3 p) {                                 /* textures.cuh:252-265, depth 7 */
    double accum = 0.0, weight = 1.0; v3 tp = p;
    for (int i = 0; i < 7; i++) { accum += weight * perlin_noise(n, tp); weight *= 0.5; tp = vscale(2, tp); }
    return fabsf((float)accum);
}
static v3 texture_value(const oracle_scene* S, int type, int idx, float u, float v, v3 p) {   /* textures.cuh:327-349 */
    switch (type) {
        case MORT_TEX_SOLID: return V3P(S->sol[idx].color);
        case MORT_TEX_CHECKER: {                                                      /* textures.cuh:52-60 */
            const mscn_checker* c = &S->chk[idx];
            int xi = (int)floorf(c->inv_scale * p.x), yi = (int)floorf(c->inv_scale * p.y), zi = (int)floorf(c->inv_scale * p.z);
            int even = (xi + yi + zi) % 2 == 0;
            return even ? texture_value(S, c->even_type, c->even_idx, u, v, p) : texture_value(S, c->odd_type, c->odd_idx, u, v, p);
        }
        case MORT_TEX_IMAGE: {                                                        /* textures.cuh:129-146 */
            const mscn_image* im = &S->img[idx];
            if (im->height <= 0 || !S->rgb) return V(0, 1, 1);
            u = u < 0 ? 0 : (u > 1 ? 1 : u);
            float vc = v < 0 ? 0 : (v > 1 ? 1 : v);
            v = (float)(1.0 - vc);
            int i = (int)(u * im->width), j = (int)(v * im->height);
            /* tex2D on an unnormalised, point-sampled pitch-2D byte texture clamps to the edge */
            int cols = im->width * 3; if (j > im->height - 1) j = im->height - 1; if (j < 0) j = 0;
            int ch[3];
            for (int k = 0; k < 3; k++) { int x = i * 3 + k; if (x > cols - 1) x = cols - 1; if (x < 0) x = 0; ch[k] = S->rgb[(size_t)j * cols + x]; }
            float sc = (float)(1.0 / 255.0);
            return V(sc * ch[0], sc * ch[1], sc * ch[2]);
        }
        case MORT_TEX_NOISE: {                                                        /* textures.cuh:198-202 */
            const mscn_noise* n = &S->noi[idx];
            v3 s = vscale(n->scale, p);
            float f = (float)(1 + sin(s.z + 10.0 * perlin_turb(n, s)));
            return vscale(f, V(0.5f, 0.5f, 0.5f));
        }
    }
    float e = (float)(((int)floorf(u * 1000.0f) % 2) == ((int)floorf(v * 1000.0f) % 2));   /* textures.cuh:347-348 */
    return V(e, 0, e);
}

/* ------------------------------------------------------------------------------------------------ */
/* sampling helpers (vec3.cuh:148-212, onb.cuh:41-50)                                               */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { v3 u, v, w; } onb_t;
static void onb_from_w(onb_t* b, v3 w) {
    v3 uw = vunit(w);
    v3 a = (fabsf(uw.x) > 0.9f) ? V(0, 1, 0) : V(1, 0, 0);
    v3 v = vunit(vcross(uw, a));
    b->u = vcross(uw, v); b->v = v; b->w = uw;
}
static v3 onb_local(const onb_t* b, v3 a) {       /* a.x*u + a.y*v + a.z*w with the device contraction */
    return V(fmaf(a.z, b->w.x, fmaf(a.x, b->u.x, a.y * b->v.x)), fmaf(a.z, b->w.y, fmaf(a.x, b->u.y, a.y * b->v.y)),
             fmaf(a.z, b->w.z, fmaf(a.x, b->u.z, a.y * b->v.z)));
}
static v3 random_in_unit_sphere(rng_t* g) {
    for (;;) {
        rng_align(g);
        float x = random_float_range(g, -1, 1), y = random_float_range(g, -1, 1), z = random_float_range(g, -1, 1);
        v3 p = V(x, y, z);
        if (vlen2(p) >= 1) continue;
        return p;
    }
}
static v3 random_unit_vector(rng_t* g) { return vunit(random_in_unit_sphere(g)); }
static v3 random_in_unit_disk(rng_t* g) {
    for (;;) {
        rng_align(g);
        float x = random_float_range(g, -1, 1), y = random_float_range(g, -1, 1);
        v3 p = V(x, y, 0);
        if (vlen2(p) < 1) return p;
    }
}
static v3 random_cosine_direction(rng_t* g) {                                         /* vec3.cuh:180-191 */
    float r1 = random_float(g), r2 = random_float(g);
    float phi = (float)(2 * 3.1415926 * r1);
    float x = cosf(phi) * sqrtf(r2), y = sinf(phi) * sqrtf(r2), z = sqrtf(1 - r2);
    return V(x, y, z);
}
static v3 reflect(v3 v, v3 n) { return vsub(v, vscale(2 * vdot(v, n), n)); }           /* vec3.cuh:193-196 */
static v3 refract(v3 uv, v3 n, float eta) {                                           /* vec3.cuh:198-204 */
    float cos_theta = (float)fmin((double)vdot(vneg(uv), n), 1.0);
    v3 perp = vscale(eta, vadd(uv, vscale(cos_theta, n)));
    v3 par = vscale((float)(-sqrt(fabs(1.0 - vlen2(perp)))), n);
    return vadd(perp, par);
}
static float reflectance(float cosine, float ref_idx) {                               /* vec3.cuh:206-212 */
    float r0 = (1 - ref_idx) / (1 + ref_idx); r0 = r0 * r0;
    float x = 1 - cosine, x2 = x * x;
    return r0 + (1 - r0) * (x2 * x2 * x);
}

/* ------------------------------------------------------------------------------------------------ */
/* light sampling (objects.cuh:110-145, 217-235, 488-504, 947-979)                                  */
/* ------------------------------------------------------------------------------------------------ */
static float pdf_value_dispatch(const oracle_scene* S, int type, int idx, v3 origin, v3 dir);
static v3 random_dispatch(const oracle_scene* S, int type, int idx, v3 origin, rng_t* g);

static float sphere_pdf_value(const mscn_sphere* s, int idx, v3 origin, v3 dir) {
    hitrec rec; ray_t r = { origin, dir, 0 };
    if (!sphere_hit(s, idx, &r, 0.001f, HUGE_VALF, &rec)) return 0;
    float cos_theta_max = sqrtf(1 - s->radius * s->radius / vlen2(vsub(V3P(s->center), origin)));
    float solid_angle = (float)(2 * 3.1415926 * (1 - cos_theta_max));
    return (float)(1.0 / solid_angle);
}
static v3 sphere_random(const mscn_sphere* s, v3 origin, rng_t* g) {
    v3 direction = vsub(V3P(s->center), origin);
    float d2 = vlen2(direction);
    onb_t uvw; onb_from_w(&uvw, direction);
    float r1 = random_float(g), r2 = random_float(g);
    float z = 1 + r2 * (sqrtf(1 - s->radius * s->radius / d2) - 1);
    float phi = (float)(2 * 3.141592 * r1);
    float x = cosf(phi) * sqrtf(1 - z * z), y = sinf(phi) * sqrtf(1 - z * z);
    return onb_local(&uvw, V(x, y, z));
}
static float quad_pdf_value(const mscn_quad* q, int idx, v3 origin, v3 dir) {
    hitrec rec; ray_t r = { origin, dir, 0 };
    if (!quad_hit(q, idx, &r, 0.001f, HUGE_VALF, &rec)) return 0;
    float d2 = rec.t * rec.t * vlen2(dir);
    float cosine = fabsf(vdot(dir, rec.normal) / vlen(dir));
    return d2 / (cosine * q->area);
}
static v3 quad_random(const mscn_quad* q, v3 origin, rng_t* g) {
    float r1 = random_float(g), r2 = random_float(g);
    v3 p = vadd(vadd(V3P(q->Q), vscale(r1, V3P(q->u))), vscale(r2, V3P(q->v)));
    return vsub(p, origin);
}
static float pdf_value_dispatch(const oracle_scene* S, int type, int idx, v3 origin, v3 dir) {
    switch (type) {
        case MORT_OBJ_SPHERE: return sphere_pdf_value(&S->sph[idx], idx, origin, dir);
        case MORT_OBJ_QUAD: return quad_pdf_value(&S->quad[idx], idx, origin, dir);
        case MORT_OBJ_HITTABLE_LIST: {
            const list_t* l = &S->lists[idx];
            float weight = (float)(1.0 / (float)l->num), sum = 0.0f;
            for (int i = 0; i < l->num; i++) sum += weight * pdf_value_dispatch(S, l->items[2 * i], l->items[2 * i + 1], origin, dir);
            return sum;
        }
    }
    return 0.0f;
}
static v3 random_dispatch(const oracle_scene* S, int type, int idx, v3 origin, rng_t* g) {
    switch (type) {
        case MORT_OBJ_SPHERE: return sphere_random(&S->sph[idx], origin, g);
        case MORT_OBJ_QUAD: return quad_random(&S->quad[idx], origin, g);
        case MORT_OBJ_HITTABLE_LIST: {
            const list_t* l = &S->lists[idx];
            int k = random_int(g, 0, l->num - 1);
            return random_dispatch(S, l->items[2 * k], l->items[2 * k + 1], origin, g);
        }
    }
    return V(1, 0, 0);
}

/* ------------------------------------------------------------------------------------------------ */
/* materials (materials.cuh) and pdfs (pdf.cuh)                                                     */
/* ------------------------------------------------------------------------------------------------ */
enum { PDF_NONE = 0, PDF_COSINE = 1, PDF_SPHERE = 2 };
typedef struct { v3 attenuation; int pdf_kind; onb_t uvw; int skip_pdf; ray_t skip_ray; } scatter_rec;

static int scatter_dispatch(const oracle_scene* S, const ray_t* r_in, const hitrec* rec, scatter_rec* sr, rng_t* g, rng_t* sg) {   /* materials.cuh:272-296 */
    switch (rec->mat_type) {
        case MORT_MAT_LAMBERTIAN: {                                                   /* materials.cuh:38-44 */
            const mscn_lambertian* m = &S->lam[rec->mat_idx];
            sr->attenuation = texture_value(S, m->tex_type, m->tex_idx, rec->u, rec->v, rec->p);
            sr->pdf_kind = PDF_COSINE; onb_from_w(&sr->uvw, rec->normal); sr->skip_pdf = 0;
            return 1;
        }
        case MORT_MAT_METAL: {                                                        /* materials.cuh:73-84 */
            const mscn_metal* m = &S->met[rec->mat_idx];
            v3 refl = reflect(r_in->d, rec->normal);
            refl = vadd(vunit(refl), vscale(m->fuzz, random_unit_vector(g)));
            sr->attenuation = V3P(m->albedo); sr->pdf_kind = PDF_NONE; sr->skip_pdf = 1;
            sr->skip_ray.o = rec->p; sr->skip_ray.d = refl; sr->skip_ray.tm = r_in->tm;
            return 1;
        }
        case MORT_MAT_DIELECTRIC: {                                                   /* materials.cuh:107-130 */
            const mscn_dielectric* m = &S->die[rec->mat_idx];
            sr->attenuation = V(1, 1, 1); sr->pdf_kind = PDF_NONE; sr->skip_pdf = 1;
            float ratio = rec->front_face ? m->inv_ior : m->ior;
            v3 ud = vunit(r_in->d);
            float cos_theta = (float)fmin((double)vdot(vneg(ud), rec->normal), 1.0);
            float sin_theta = (float)sqrt(1.0 - cos_theta * cos_theta);
            int cant_refract = (double)(ratio * sin_theta) > 1.0;
            v3 dir;
            if (cant_refract || reflectance(cos_theta, ratio) > random_float(sg)) dir = reflect(ud, rec->normal);
            else dir = refract(ud, rec->normal, ratio);
            sr->skip_ray.o = rec->p; sr->skip_ray.d = dir; sr->skip_ray.tm = r_in->tm;
            return 1;
        }
        case MORT_MAT_DIFFUSE_LIGHT: return 0;                                        /* materials.cuh:151-154 */
        case MORT_MAT_ISOTROPIC: {                                                    /* materials.cuh:182-188 */
            const mscn_isotropic* m = &S->iso[rec->mat_idx];
            sr->attenuation = texture_value(S, m->tex_type, m->tex_idx, rec->u, rec->v, rec->p);
            sr->pdf_kind = PDF_SPHERE; sr->skip_pdf = 0;
            return 1;
        }
    }
    return 0;
}
static v3 emit_dispatch(const oracle_scene* S, const hitrec* rec) {                  /* materials.cuh:298-322, 156-163 */
    if (rec->mat_type == MORT_MAT_DIFFUSE_LIGHT) {
        if (!rec->front_face) return V(0, 0, 0);
        const mscn_diffuse_light* m = &S->lig[rec->mat_idx];
        return texture_value(S, m->tex_type, m->tex_idx, rec->u, rec->v, rec->p);
    }
    return V(0, 0, 0);
}
static float scatter_pdf_dispatch(const hitrec* rec, const ray_t* scattered) {       /* materials.cuh:324-349, 51-55, 195-198 */
    if (rec->mat_type == MORT_MAT_LAMBERTIAN) {
        float cos_theta = vdot(rec->normal, vunit(scattered->d));
        return (float)((cos_theta < 0) ? 0 : cos_theta / 3.141592565);
    }
    if (rec->mat_type == MORT_MAT_ISOTROPIC) return (float)(1 / (4 * 3.1415926));
    return 0;
}
static float mat_pdf_value(const scatter_rec* sr, v3 dir) {                           /* pdf.cuh:30-32, 46-49 */
    if (sr->pdf_kind == PDF_COSINE) { float c = vdot(vunit(dir), sr->uvw.w); return fmaxf(0, (float)(c / 3.1415926)); }
    return (float)(1 / (4 * 3.1415926));
}
static v3 mat_pdf_generate(const scatter_rec* sr, rng_t* g, rng_t* sg) {              /* pdf.cuh:35-37, 52-54 */
    if (sr->pdf_kind == PDF_COSINE) return onb_local(&sr->uvw, random_cosine_direction(sg));
    return random_unit_vector(g);                       /* rejection loop: own aligned blocks */
}

/* Camera::ray_color, camera.cuh:86-176 (forward pass into per-bounce arrays, then the unwind) */
#define ORACLE_MAX_DEPTH 1024
static int g_debug_pixel = -2;
static v3 ray_color(const oracle_scene* S, const ray_t* r, rng_t* g, uint64_t* segments) {
    if (g_debug_pixel == -2) { const char* e = getenv("ORACLE_DEBUG_PIXEL"); g_debug_pixel = e ? atoi(e) : -1; }
    const int dbg = g_debug_pixel >= 0 && (int)g->ctr[0] == g_debug_pixel;
    const mscn_camera* cam = &S->cam;
    int limit = cam->bounce_limit > ORACLE_MAX_DEPTH ? ORACLE_MAX_DEPTH : cam->bounce_limit;
    v3 att[ORACLE_MAX_DEPTH], em[ORACLE_MAX_DEPTH]; float spdf[ORACLE_MAX_DEPTH], pdfv[ORACLE_MAX_DEPTH];
    hitrec rec; memset(&rec, 0, sizeof(rec));
    int iter = 0; ray_t cur = *r; v3 final = V(0, 0, 0);
    while (iter < limit) {
        (*segments)++;
        /* canonical stream: this bounce's stage block, generated before the closest-hit query */
        rng_t sg = *g; rng_align(&sg); (void)random_float(&sg); sg.have = 4;      /* sg.buf = block, nothing consumed yet */
        g->ctr[2] = sg.ctr[2]; rng_align(g);                                       /* the main stream continues after it */
        if (world_hit(S, &cur, 0.001f, INFINITY, &rec, g, 1, NULL, NULL)) {
            if (dbg) fprintf(stderr, "ORACLE px %u smp %u seg %d: hit type %d idx %d t %.9g mat %d/%d ff %d p %.9g %.9g %.9g d %.9g %.9g %.9g\n", g->ctr[0], g->ctr[1], iter,
                             rec.leaf_type, rec.leaf_idx, rec.t, rec.mat_type, rec.mat_idx, rec.front_face, rec.p.x, rec.p.y, rec.p.z, cur.d.x, cur.d.y, cur.d.z);
            ray_t scattered; v3 emission = emit_dispatch(S, &rec);
            float pdf, scattering_pdf; scatter_rec sr;
            if (scatter_dispatch(S, &cur, &rec, &sr, g, &sg)) {
                if (sr.skip_pdf) {
                    cur = sr.skip_ray; att[iter] = sr.attenuation; em[iter] = V(0, 0, 0); spdf[iter] = 1.0f; pdfv[iter] = 1.0f;
                    iter++; continue;
                }
                if (cam->light_obj_type == -1) {
                    scattered.o = rec.p; scattered.d = mat_pdf_generate(&sr, g, &sg); scattered.tm = r->tm;
                    pdf = mat_pdf_value(&sr, scattered.d);
                } else {                                                             /* pdf.cuh:85-103 */
                    v3 dir;
                    if (random_float(&sg) < 0.5f) dir = random_dispatch(S, cam->light_obj_type, cam->light_obj_idx, rec.p, &sg);
                    else dir = mat_pdf_generate(&sr, g, &sg);
                    scattered.o = rec.p; scattered.d = dir; scattered.tm = r->tm;
                    pdf = (float)(0.5 * pdf_value_dispatch(S, cam->light_obj_type, cam->light_obj_idx, rec.p, dir) + 0.5 * mat_pdf_value(&sr, dir));
                }
                scattering_pdf = scatter_pdf_dispatch(&rec, &scattered);
                cur = scattered; att[iter] = sr.attenuation; em[iter] = emission; spdf[iter] = scattering_pdf; pdfv[iter] = pdf;
                iter++;
            } else { final = emission; break; }
        } else { final = V3P(cam->background); break; }
    }
    if (iter == limit) final = V(0, 0, 0);
    while (iter > 0) {
        iter--;
        if (dbg) fprintf(stderr, "ORACLE   unwind %d: att %.9g %.9g %.9g spdf %.9g pdf %.9g\n", iter, att[iter].x, att[iter].y, att[iter].z, spdf[iter], pdfv[iter]);
        v3 num = vmul(vscale(spdf[iter], att[iter]), final);      /* attenuation * scattering_pdf * finalValue */
        final = vadd(em[iter], vdiv(num, pdfv[iter]));
    }
    if (dbg) fprintf(stderr, "ORACLE   color %.9g %.9g %.9g\n", final.x, final.y, final.z);
    return final;
}

/* Camera::get_ray + sample_square_stratified + defocus_disk_sample, camera.cuh:210-242 */
static void get_ray(const mscn_camera* c, int x, int y, int s_i, int s_j, rng_t* g, ray_t* out) {
    double px = ((s_i + random_float(g)) * c->recip_sqrt_spp) - 0.5;
    double py = ((s_j + random_float(g)) * c->recip_sqrt_spp) - 0.5;
    float ox = (float)px, oy = (float)py;
    float tu = (float)((double)x + ox), tv = (float)((double)y + oy);
    v3 ps = vfma(tv, V3P(c->pixel_delta_v), vfma(tu, V3P(c->pixel_delta_u), V3P(c->pixel00_loc)));
    v3 origin;
    if (c->defocus_angle <= 0) origin = V3P(c->center);
    else { v3 p = random_in_unit_disk(g); origin = vfma(p.y, V3P(c->defocus_disk_v), vfma(p.x, V3P(c->defocus_disk_u), V3P(c->center))); }
    out->o = origin; out->d = vsub(ps, origin); out->tm = random_float(g);
}
void oracle_camera_ray(const oracle_scene* s, uint32_t seed, uint32_t frame, int x, int y, int s_i, int s_j, float* ray7) {
    rng_t g; rng_init(&g, seed, frame, (uint32_t)(x + y * s->cam.image_width), (uint32_t)(s_j * s->cam.sqrt_spp + s_i));
    ray_t r; get_ray(&s->cam, x, y, s_i, s_j, &g, &r);
    ray7[0] = r.o.x; ray7[1] = r.o.y; ray7[2] = r.o.z; ray7[3] = r.d.x; ray7[4] = r.d.y; ray7[5] = r.d.z; ray7[6] = r.tm;
}

/* Camera::render, camera.cuh:178-208 */
typedef struct { const oracle_scene* S; uint32_t seed, frame; int sj_mod, sj_rem, tid, nthreads; float* hdr; uint8_t* rgba8; uint64_t segments, samples; } job_t;
static float clampf(float x, float lo, float hi) { if (x < lo) return lo; if (x > hi) return hi; return x; }
static void* render_rows(void* arg) {
    job_t* J = (job_t*)arg; const oracle_scene* S = J->S; const mscn_camera* c = &S->cam;
    int W = c->image_width, H = c->image_height;
    for (int y = J->tid; y < H; y += J->nthreads)
        for (int x = 0; x < W; x++) {
            int offset = x + y * W;
            v3 all = V(0, 0, 0), fin = V(0, 0, 0); int nan_n = 0;
            for (int s_j = 0; s_j < c->sqrt_spp; s_j++) {
                if (s_j % J->sj_mod != J->sj_rem) continue;
                for (int s_i = 0; s_i < c->sqrt_spp; s_i++) {
                    rng_t g; rng_init(&g, J->seed, J->frame, (uint32_t)offset, (uint32_t)(s_j * c->sqrt_spp + s_i));
                    ray_t r; get_ray(c, x, y, s_i, s_j, &g, &r);
                    v3 col = ray_color(S, &r, &g, &J->segments);
                    J->samples++;
                    all = vadd(all, col);
                    if (visnan(col)) nan_n++; else fin = vadd(fin, col);
                }
            }
            if (J->hdr) { float* o = J->hdr + 4 * (size_t)offset; o[0] = fin.x; o[1] = fin.y; o[2] = fin.z; o[3] = (float)nan_n; }
            if (J->rgba8) {
                v3 pc = vscale(c->pixel_samples_scale, all);
                if (pc.x != pc.x) pc.x = 0.0f; if (pc.y != pc.y) pc.y = 0.0f; if (pc.z != pc.z) pc.z = 0.0f;
                pc.x = sqrtf(pc.x); pc.y = sqrtf(pc.y); pc.z = sqrtf(pc.z);           /* utils.h:41-43 */
                uint8_t* o = J->rgba8 + 4 * (size_t)offset;
                o[0] = (uint8_t)(int)(256 * clampf(pc.x, 0.0f, 0.999f)); o[1] = (uint8_t)(int)(256 * clampf(pc.y, 0.0f, 0.999f));
                o[2] = (uint8_t)(int)(256 * clampf(pc.z, 0.0f, 0.999f)); o[3] = 255;
            }
        }
    return NULL;
}
int oracle_render(const oracle_scene* S, uint32_t seed, uint32_t frame, int sj_mod, int sj_rem, int n_threads,
                  float* hdr, uint8_t* rgba8, uint64_t counters[2]) {
    if (sj_mod < 1) sj_mod = 1;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    job_t* jobs = (job_t*)calloc(n_threads, sizeof(job_t)); pthread_t* th = (pthread_t*)calloc(n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) {
        job_t j = { S, seed, frame, sj_mod, sj_rem, t, n_threads, hdr, rgba8, 0, 0 }; jobs[t] = j;
        if (n_threads > 1) pthread_create(&th[t], NULL, render_rows, &jobs[t]);
    }
    if (n_threads == 1) render_rows(&jobs[0]);
    uint64_t seg = 0, smp = 0;
    for (int t = 0; t < n_threads; t++) { if (n_threads > 1) pthread_join(th[t], NULL); seg += jobs[t].segments; smp += jobs[t].samples; }
    if (counters) { counters[0] = seg; counters[1] = smp; }
    free(jobs); free(th);
    return 0;
}
