// scene.cpp — host scene model: builder methods, camera set-up, the reference's median-split BVH
// ordering (needed only because it defines primitive ids), dump/load.
#include "scene.hpp"

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>

namespace mort {

float length(V3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }

// ------------------------------------------------------------------------------------------------
// HostRng: glibc random_r TYPE_3 (degree 31, separation 3), as used by rand().
// ------------------------------------------------------------------------------------------------
void HostRng::reseed(uint32_t seed) {
    if (seed == 0) seed = 1;
    int32_t r[344 + 34];
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
        int64_t v = (16807LL * r[i - 1]) % 2147483647LL;
        if (v < 0) v += 2147483647LL;
        r[i] = (int32_t)v;
    }
    for (int i = 31; i < 34; i++) r[i] = r[i - 31];
    for (int i = 34; i < 344; i++) r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
    // keep the last 34 values as a ring; next output index is 344
    for (int i = 0; i < 34; i++) r_[i] = r[344 - 34 + i];
    f_ = 0; b_ = 0; draws_ = 0;
}

int HostRng::next() {
    draws_++;
    // ring of 34: position p holds o[k-34+p]; o[k] = o[k-31] + o[k-3]
    int32_t v = (int32_t)((uint32_t)r_[(f_ + 3) % 34] + (uint32_t)r_[(f_ + 31) % 34]);
    r_[f_] = v;
    f_ = (f_ + 1) % 34;
    return (int)((uint32_t)v >> 1);
}

float HostRng::random_float() { return (float)(next() / (2147483647 + 1.0)); }
float HostRng::random_float(float lo, float hi) { return lo + (hi - lo) * random_float(); }

// ------------------------------------------------------------------------------------------------
// Camera::initialize — camera.cuh:47-84, same promotions (float / double) expression by expression.
// ------------------------------------------------------------------------------------------------
static float degrees_to_radians(float degrees) {       // utils.h:21-27: float pi, double 180.0
    const float pi = 3.1415926535897932385f;
    return (float)(degrees * pi / 180.0);
}

void Camera::initialize() {
    image_height = static_cast<int>(image_width / aspect_ratio);
    image_height = (image_height < 1) ? 1 : image_height;

    sqrt_spp = int(std::sqrt((double)samples_per_pixel));
    pixel_samples_scale = (float)(1.0 / (sqrt_spp * sqrt_spp));
    recip_sqrt_spp = (float)(1.0 / sqrt_spp);

    center = lookfrom;

    float theta = degrees_to_radians((float)vfov);
    float h = tanf(theta / 2);
    float viewport_height = 2 * h * focus_dist;
    double viewport_width = viewport_height * (static_cast<double>(image_width) / image_height);

    w = unit(lookfrom - lookat);
    u = unit(cross(vup, w));
    v = cross(w, u);

    V3 viewport_u = (float)viewport_width * u;
    V3 viewport_v = viewport_height * -v;

    pixel_delta_u = viewport_u / (float)image_width;
    pixel_delta_v = -viewport_v / (float)image_height;

    V3 viewport_upper_left = center - (focus_dist * w) - viewport_u / 2 + viewport_v / 2;
    pixel00_loc = viewport_upper_left + 0.5f * (pixel_delta_u + pixel_delta_v);

    float defocus_radius = focus_dist * tanf(degrees_to_radians(defocus_angle / 2));
    defocus_disk_u = defocus_radius * u;     // u * r == r * u (vec3.cuh:104-107)
    defocus_disk_v = defocus_radius * v;
}

static void put3(float* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
static V3 get3(const float* d) { return V3(d[0], d[1], d[2]); }

void Camera::to_record(mscn_camera& c) const {
    memset(&c, 0, sizeof(c));
    c.aspect_ratio = aspect_ratio; c.image_width = image_width; c.image_height = image_height;
    c.samples_per_pixel = samples_per_pixel; c.pixel_samples_scale = pixel_samples_scale;
    c.sqrt_spp = sqrt_spp; c.recip_sqrt_spp = recip_sqrt_spp; c.bounce_limit = bounce_limit; c.vfov = vfov;
    put3(c.background, background);
    c.light_obj_type = light_obj_type; c.light_obj_idx = light_obj_type == -1 ? 0 : light_obj_idx;
    put3(c.center, center); put3(c.pixel00_loc, pixel00_loc);
    put3(c.pixel_delta_u, pixel_delta_u); put3(c.pixel_delta_v, pixel_delta_v);
    put3(c.lookfrom, lookfrom); put3(c.lookat, lookat); put3(c.vup, vup);
    put3(c.v, v); put3(c.u, u); put3(c.w, w);
    c.defocus_angle = defocus_angle; c.focus_dist = focus_dist;
    put3(c.defocus_disk_u, defocus_disk_u); put3(c.defocus_disk_v, defocus_disk_v);
}

// ------------------------------------------------------------------------------------------------
// boxes (aabb.cuh:17-29, interval.cuh:12-15)
// ------------------------------------------------------------------------------------------------
static void box_from_points(float* b, V3 a, V3 c) {
    b[0] = fminf(a.x, c.x); b[1] = fmaxf(a.x, c.x);
    b[2] = fminf(a.y, c.y); b[3] = fmaxf(a.y, c.y);
    b[4] = fminf(a.z, c.z); b[5] = fmaxf(a.z, c.z);
}
static void box_union(float* out, const float* a, const float* b) {
    for (int k = 0; k < 3; k++) {
        float lo = a[2 * k] < b[2 * k] ? a[2 * k] : b[2 * k];
        float hi = a[2 * k + 1] > b[2 * k + 1] ? a[2 * k + 1] : b[2 * k + 1];
        out[2 * k] = lo; out[2 * k + 1] = hi;
    }
}

uint32_t fnv1a32(const uint8_t* p, size_t n) {
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 16777619u; }
    return h;
}

// ------------------------------------------------------------------------------------------------
// journal (scene text statements, scene_text.cpp)
// ------------------------------------------------------------------------------------------------
std::string handle_token(Handle h, char category) {
    static const char* obj[] = {"?", "sph", "qd", "tr", "ry", "med", "list", "bvh"};
    static const char* mat[] = {"?", "lam", "met", "die", "lgt", "iso"};
    static const char* tex[] = {"?", "sol", "chk", "img", "noi"};
    const char* const* tab = category == 'o' ? obj : (category == 'm' ? mat : tex);
    const int n = category == 'o' ? 8 : (category == 'm' ? 6 : 5);
    if (h.type < 1 || h.type >= n || h.idx < 0) return "none";
    return std::string(tab[h.type]) + std::to_string(h.idx);
}
// shortest decimal that reads back as the same float (std::to_chars): round-trips like "%.9g", 10x cheaper — a 10^6-sphere field journals 7 floats per sphere
static std::string fs(float v) { char b[48]; const auto r = std::to_chars(b, b + sizeof(b), v); return std::string(b, r.ptr); }
static std::string fs3(V3 v) { return fs(v.x) + " " + fs(v.y) + " " + fs(v.z); }
static void jlog(Scene& s, const std::string& line) { if (s.journal_mute == 0) s.journal.push_back(line); }
// one buffer per statement for the primitives that come by the million (same text as the concatenations above would give)
struct JLine {
    std::string s;
    explicit JLine(const char* head) { s.reserve(112); s = head; }
    JLine& f(float v) { char b[48]; const auto r = std::to_chars(b, b + sizeof(b), v); s.push_back(' '); s.append(b, r.ptr); return *this; }
    JLine& v(V3 p) { return f(p.x).f(p.y).f(p.z); }
    JLine& h(Handle hd, char cat) { s.push_back(' '); s += handle_token(hd, cat); return *this; }
    JLine& hidden(bool skip) { if (skip) s += " hidden"; return *this; }
};
static void jlog(Scene& s, JLine& l) { if (s.journal_mute == 0) s.journal.push_back(std::move(l.s)); }
static const char* hid(bool skip) { return skip ? " hidden" : ""; }

// ------------------------------------------------------------------------------------------------
// textures / materials
// ------------------------------------------------------------------------------------------------
Handle Scene::add_solid(V3 c) {
    jlog(*this, "solid " + fs3(c));
    mscn_solid s; put3(s.color, c); solids.push_back(s);
    return Handle{MORT_TEX_SOLID, (int)solids.size() - 1};
}
Handle Scene::add_checker(float scale, Handle even, Handle odd) {
    jlog(*this, "checker " + fs(scale) + " " + handle_token(even, 't') + " " + handle_token(odd, 't'));
    mscn_checker c; c.inv_scale = (float)(1.0 / scale);                       // textures.cuh:43
    c.even_type = even.type; c.even_idx = even.idx; c.odd_type = odd.type; c.odd_idx = odd.idx;
    checkers.push_back(c);
    return Handle{MORT_TEX_CHECKER, (int)checkers.size() - 1};
}
Handle Scene::add_image(const uint8_t* rgb, int width, int height, const char* source) {
    ImageRec im; im.width = width; im.height = height;
    if (source) im.source = source;
    jlog(*this, source ? std::string("image ") + source : "image @" + std::to_string(images.size()));
    if (rgb && width > 0 && height > 0) {
        im.rgb.assign(rgb, rgb + (size_t)width * height * 3);
        im.fnv1a = fnv1a32(im.rgb.data(), im.rgb.size());
    }
    images.push_back(std::move(im));
    return Handle{MORT_TEX_IMAGE, (int)images.size() - 1};
}
Handle Scene::add_noise(float scale, HostRng& rng) {
    // textures.cuh:164-172 then 216-230.  Constructor arguments are evaluated right to left by the
    // reference's host compiler: z is drawn first.
    mscn_noise n; memset(&n, 0, sizeof(n));
    n.scale = scale;
    jlog(*this, "noise " + fs(scale) + " at " + std::to_string(rng.draws()));
    for (int i = 0; i < MORT_PERLIN_POINTS; i++) {
        float z = rng.random_float(-1, 1), y = rng.random_float(-1, 1), x = rng.random_float(-1, 1);
        put3(n.ranvec[i], unit(V3(x, y, z)));
    }
    int32_t* perms[3] = {n.perm_x, n.perm_y, n.perm_z};
    for (int a = 0; a < 3; a++) {
        int32_t* p = perms[a];
        for (int i = 0; i < MORT_PERLIN_POINTS; i++) p[i] = i;
        for (int i = MORT_PERLIN_POINTS - 1; i > 0; i--) {
            int target = (int)rng.random_float(0.0f, (float)i);
            int32_t tmp = p[i]; p[i] = p[target]; p[target] = tmp;
        }
    }
    noises.push_back(n);
    return Handle{MORT_TEX_NOISE, (int)noises.size() - 1};
}
Handle Scene::add_noise_tables(const mscn_noise& n) {
    journal_complete = false;                            // explicit tables have no text form
    noises.push_back(n);
    return Handle{MORT_TEX_NOISE, (int)noises.size() - 1};
}

Handle Scene::add_lambertian(Handle tex) {
    jlog(*this, "lambertian " + handle_token(tex, 't'));
    lambertians.push_back(mscn_lambertian{tex.type, tex.idx});
    return Handle{MORT_MAT_LAMBERTIAN, (int)lambertians.size() - 1};
}
Handle Scene::add_metal(V3 albedo, float fuzz) {
    jlog(*this, "metal " + fs3(albedo) + " " + fs(fuzz));
    mscn_metal m; put3(m.albedo, albedo); m.fuzz = fuzz; metals.push_back(m);
    return Handle{MORT_MAT_METAL, (int)metals.size() - 1};
}
Handle Scene::add_dielectric(float ior) {
    jlog(*this, "dielectric " + fs(ior));
    mscn_dielectric d; d.ior = ior; d.inv_ior = (float)(1.0 / ior); d.albedo[0] = d.albedo[1] = d.albedo[2] = 1.0f;
    dielectrics.push_back(d);
    return Handle{MORT_MAT_DIELECTRIC, (int)dielectrics.size() - 1};
}
Handle Scene::add_diffuse_light(Handle tex) {
    jlog(*this, "light " + handle_token(tex, 't'));
    lights.push_back(mscn_diffuse_light{tex.type, tex.idx});
    return Handle{MORT_MAT_DIFFUSE_LIGHT, (int)lights.size() - 1};
}
Handle Scene::add_isotropic(Handle tex) {
    jlog(*this, "isotropic " + handle_token(tex, 't'));
    isotropics.push_back(mscn_isotropic{tex.type, tex.idx});
    return Handle{MORT_MAT_ISOTROPIC, (int)isotropics.size() - 1};
}

// ------------------------------------------------------------------------------------------------
// hittables
// ------------------------------------------------------------------------------------------------
Handle Scene::add_sphere(V3 c, float r, Handle mat, bool skip) {
    if (journal_mute == 0) { JLine l("sphere"); jlog(*this, l.v(c).f(r).h(mat, 'm').hidden(skip)); }
    mscn_sphere s; memset(&s, 0, sizeof(s));
    put3(s.center, c); s.radius = r; s.moves = 0; s.mat_type = mat.type; s.mat_idx = mat.idx; s.skip = skip ? 1 : 0;
    V3 rv(r, r, r);
    box_from_points(s.bbox, c - rv, c + rv);                                  // objects.cuh:39-42
    spheres.push_back(s);
    return Handle{MORT_OBJ_SPHERE, (int)spheres.size() - 1};
}
Handle Scene::add_moving_sphere(V3 c1, V3 c2, float r, Handle mat, bool skip) {
    if (journal_mute == 0) { JLine l("moving_sphere"); jlog(*this, l.v(c1).v(c2).f(r).h(mat, 'm').hidden(skip)); }
    mscn_sphere s; memset(&s, 0, sizeof(s));
    put3(s.center, c1); s.radius = r; s.moves = 1; put3(s.center_vec, c2 - c1);
    s.mat_type = mat.type; s.mat_idx = mat.idx; s.skip = skip ? 1 : 0;
    V3 rv(r, r, r); float b1[6], b2[6];
    box_from_points(b1, c1 - rv, c1 + rv); box_from_points(b2, c2 - rv, c2 + rv);
    box_union(s.bbox, b1, b2);                                                // objects.cuh:52-54
    spheres.push_back(s);
    return Handle{MORT_OBJ_SPHERE, (int)spheres.size() - 1};
}
bool Scene::update_sphere(int idx, V3 c1, const V3* c2, float r) {
    if (idx < 0 || idx >= (int)spheres.size()) return false;
    mscn_sphere& s = spheres[idx];
    put3(s.center, c1); s.radius = r;
    V3 rv(r, r, r); float b1[6];
    box_from_points(b1, c1 - rv, c1 + rv);
    if (c2) { float b2[6]; s.moves = 1; put3(s.center_vec, *c2 - c1); box_from_points(b2, *c2 - rv, *c2 + rv); box_union(s.bbox, b1, b2); }
    else { s.moves = 0; put3(s.center_vec, V3(0, 0, 0)); memcpy(s.bbox, b1, sizeof(b1)); }
    edited = true; journal_complete = false;            // the journal no longer describes the arrays
    return true;
}
Handle Scene::add_quad(V3 Q, V3 u, V3 v, Handle mat, bool skip) {
    jlog(*this, "quad " + fs3(Q) + " " + fs3(u) + " " + fs3(v) + " " + handle_token(mat, 'm') + hid(skip));
    mscn_quad q; memset(&q, 0, sizeof(q));
    V3 n = cross(u, v);                                                       // objects.cuh:173-184
    V3 normal = unit(n);
    put3(q.Q, Q); put3(q.u, u); put3(q.v, v); put3(q.normal, normal);
    q.D = dot(normal, Q);
    put3(q.w, n / dot(n, n));
    q.area = length(n);
    q.mat_type = mat.type; q.mat_idx = mat.idx; q.skip = skip ? 1 : 0;
    float d1[6], d2[6];
    box_from_points(d1, Q, Q + u + v); box_from_points(d2, Q + u, Q + v);
    box_union(q.bbox, d1, d2);
    quads.push_back(q);
    return Handle{MORT_OBJ_QUAD, (int)quads.size() - 1};
}
Handle Scene::add_translate(Handle obj, V3 offset, bool skip) {
    jlog(*this, "translate " + handle_token(obj, 'o') + " " + fs3(offset) + hid(skip));
    mscn_translate t; t.obj_type = obj.type; t.obj_idx = obj.idx; put3(t.offset, offset); t.skip = skip ? 1 : 0;
    translates.push_back(t);
    Box6 bb; float c[6];                                                     // objects.cuh:264: child's box + displacement
    if (!bbox_of(obj, c)) memset(c, 0, sizeof(c));
    for (int a = 0; a < 3; a++) { bb.b[2 * a] = c[2 * a] + offset[a]; bb.b[2 * a + 1] = c[2 * a + 1] + offset[a]; }
    translate_bbox.push_back(bb);
    return Handle{MORT_OBJ_TRANSLATE, (int)translates.size() - 1};
}
Handle Scene::add_rotate_y(Handle obj, float theta_deg, bool skip) {
    jlog(*this, "rotate_y " + handle_token(obj, 'o') + " " + fs(theta_deg) + hid(skip));
    mscn_rotate_y r; r.obj_type = obj.type; r.obj_idx = obj.idx; r.skip = skip ? 1 : 0;
    float radians = (float)(theta_deg * 3.1415926535897932385 / 180.0);      // objects.cuh:297-299
    r.sin_theta = sinf(radians); r.cos_theta = cosf(radians);
    rotates.push_back(r);
    Box6 bb; float c[6];                                                     // objects.cuh:304-329: box of the 8 rotated corners
    if (!bbox_of(obj, c)) memset(c, 0, sizeof(c));
    float pmin[3] = {HUGE_VALF, HUGE_VALF, HUGE_VALF}, pmax[3] = {-HUGE_VALF, -HUGE_VALF, -HUGE_VALF};
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
        float x = i * c[1] + (1 - i) * c[0], y = j * c[3] + (1 - j) * c[2], z = k * c[5] + (1 - k) * c[4];
        float nx = r.cos_theta * x + r.sin_theta * z, nz = -r.sin_theta * x + r.cos_theta * z;
        float t[3] = {nx, y, nz};
        for (int a = 0; a < 3; a++) { pmin[a] = fminf(pmin[a], t[a]); pmax[a] = fmaxf(pmax[a], t[a]); }
    }
    box_from_points(bb.b, V3(pmin[0], pmin[1], pmin[2]), V3(pmax[0], pmax[1], pmax[2]));
    rotate_bbox.push_back(bb);
    return Handle{MORT_OBJ_ROTATE_Y, (int)rotates.size() - 1};
}
Handle Scene::add_constant_medium(Handle boundary, float density, Handle mat, bool skip) {
    jlog(*this, "medium " + handle_token(boundary, 'o') + " " + fs(density) + " " + handle_token(mat, 'm') + hid(skip));
    mscn_medium m; memset(&m, 0, sizeof(m));
    m.obj_type = boundary.type; m.obj_idx = boundary.idx;
    m.neg_inv_density = -(1.0 / density);                                     // objects.cuh:387 (double)
    m.mat_type = mat.type; m.mat_idx = mat.idx; m.skip = skip ? 1 : 0;
    media.push_back(m);
    Box6 bb;                                                                 // objects.cuh:392: the boundary's box
    if (!bbox_of(boundary, bb.b)) memset(bb.b, 0, sizeof(bb.b));
    medium_bbox.push_back(bb);
    return Handle{MORT_OBJ_CONSTANT_MEDIUM, (int)media.size() - 1};
}
Handle Scene::add_list(bool skip) {
    jlog(*this, std::string("list") + hid(skip));
    ListRec l; l.skip = skip ? 1 : 0; lists.push_back(l);
    return Handle{MORT_OBJ_HITTABLE_LIST, (int)lists.size() - 1};
}
int Scene::list_add(Handle list, Handle obj) {
    if (list.type != MORT_OBJ_HITTABLE_LIST || list.idx < 0 || list.idx >= (int)lists.size()) return -1;
    ListRec& L = lists[list.idx];
    float b[6];                                                              // objects.cuh:463-469: first item's box, then unions
    if (!bbox_of(obj, b)) memset(b, 0, sizeof(b));
    if (L.items.empty()) memcpy(L.bbox, b, sizeof(b));
    else { float u[6]; box_union(u, L.bbox, b); memcpy(L.bbox, u, sizeof(u)); }
    L.has_bbox = true;
    L.items.push_back(obj);
    jlog(*this, "add " + handle_token(list, 'o') + " " + handle_token(obj, 'o'));
    return 0;
}

// host_getBboxInfo (objects.cuh:918-945); any other handle type falls off the reference's switch (garbage there, zeros here)
bool Scene::bbox_of(Handle h, float out[6]) const {
    auto in = [&](size_t n) { return h.idx >= 0 && (size_t)h.idx < n; };
    if (h.type == MORT_OBJ_SPHERE && in(spheres.size())) { memcpy(out, spheres[h.idx].bbox, 24); return true; }
    if (h.type == MORT_OBJ_QUAD && in(quads.size())) { memcpy(out, quads[h.idx].bbox, 24); return true; }
    if (h.type == MORT_OBJ_TRANSLATE && in(translate_bbox.size())) { memcpy(out, translate_bbox[h.idx].b, 24); return true; }
    if (h.type == MORT_OBJ_ROTATE_Y && in(rotate_bbox.size())) { memcpy(out, rotate_bbox[h.idx].b, 24); return true; }
    if (h.type == MORT_OBJ_CONSTANT_MEDIUM && in(medium_bbox.size())) { memcpy(out, medium_bbox[h.idx].b, 24); return true; }
    if (h.type == MORT_OBJ_HITTABLE_LIST && in(lists.size())) { memcpy(out, lists[h.idx].bbox, 24); return true; }
    return false;
}

// The reference's BVH constructor (objects.cuh:528-611) is restated here for ONE reason: its bubble sort
// physically swaps same-type primitives in the world arrays (objects.cuh:630-661, 815-845), so it defines
// which array slot — i.e. which primitive id — every sphere ends up in.  The node arrays it produces are
// kept for the scene dump; rendering uses the SAH wide BVH of bvh_build.cpp instead.
Handle Scene::add_bvh(Handle list, bool skip) {
    jlog(*this, "bvh " + handle_token(list, 'o') + hid(skip));
    BvhRec B; B.skip = skip ? 1 : 0; B.list_idx = list.idx;
    if (list.type != MORT_OBJ_HITTABLE_LIST || list.idx < 0 || list.idx >= (int)lists.size()) {
        error = "add_bvh: not a list"; bvhs.push_back(B); bvh_mode = true;
        return Handle{MORT_OBJ_BVH, (int)bvhs.size() - 1};
    }
    std::vector<Handle>& it = lists[list.idx].items;
    auto box = [&](int k, float* b) {
        if (!bbox_of(it[k], b)) { for (int a = 0; a < 6; a++) b[a] = 0; }
    };
    auto cmp = [&](int a, int b, int axis) {                                  // objects.cuh:981-1000
        float ba[6], bb[6]; box(a, ba); box(b, bb);
        if (ba[2 * axis] < bb[2 * axis]) return -1;
        if (ba[2 * axis] > bb[2 * axis]) return 1;
        return 0;
    };
    auto swap_objs = [&](Handle a, Handle b) {
        if (a.type == MORT_OBJ_SPHERE) std::swap(spheres[a.idx], spheres[b.idx]);
        else if (a.type == MORT_OBJ_QUAD) std::swap(quads[a.idx], quads[b.idx]);
        else if (a.type == MORT_OBJ_TRANSLATE) { std::swap(translates[a.idx], translates[b.idx]); std::swap(translate_bbox[a.idx], translate_bbox[b.idx]); }
        else if (a.type == MORT_OBJ_ROTATE_Y) { std::swap(rotates[a.idx], rotates[b.idx]); std::swap(rotate_bbox[a.idx], rotate_bbox[b.idx]); }
        else if (a.type == MORT_OBJ_CONSTANT_MEDIUM) { std::swap(media[a.idx], media[b.idx]); std::swap(medium_bbox[a.idx], medium_bbox[b.idx]); }
    };
    struct Span { int b, e; };
    std::vector<Span> spans; spans.push_back(Span{0, (int)it.size()});
    size_t cur = 0;
    while (cur < spans.size()) {
        int s0 = spans[cur].b, s1 = spans[cur].e;
        mscn_bvh_node nd; memset(&nd, 0, sizeof(nd));
        const float inf = std::numeric_limits<float>::infinity();
        float bx[6] = {inf, -inf, inf, -inf, inf, -inf};                       // aabb::empty
        for (int i = s0; i < s1; i++) { float b[6]; box(i, b); float u[6]; box_union(u, bx, b); memcpy(bx, u, 24); }
        memcpy(nd.bbox, bx, 24);
        float sx = bx[1] - bx[0], sy = bx[3] - bx[2], sz = bx[5] - bx[4];
        int axis = (sx > sy) ? (sx > sz ? 0 : 2) : (sy > sz ? 1 : 2);         // aabb.cuh:62-67
        int span = s1 - s0;
        if (span == 1) {
            nd.left_type = nd.right_type = it[s0].type; nd.left_idx = nd.right_idx = it[s0].idx; nd.is_internal = 0;
        } else if (span == 2) {
            int a = s0, b = s0 + 1;
            if (cmp(s0, s0 + 1, axis) > 0) { a = s0 + 1; b = s0; }
            nd.left_type = it[a].type; nd.left_idx = it[a].idx; nd.right_type = it[b].type; nd.right_idx = it[b].idx;
            nd.is_internal = 0;
        } else if (span > 2) {
            for (int i = 0; i < span - 1; i++) {                               // bubble sort, objects.cuh:631-661
                bool swapped = false;
                for (int j = s0; j < s1 - i - 1; j++) {
                    if (cmp(j, j + 1, axis) == 1) {
                        if (it[j].type == it[j + 1].type) swap_objs(it[j], it[j + 1]);
                        else std::swap(it[j], it[j + 1]);
                        swapped = true;
                    }
                }
                if (!swapped) break;
            }
            int mid = s0 + (span / 2 + (span % 2 != 0));
            nd.left_type = MORT_OBJ_BVH; nd.left_idx = (int)spans.size(); spans.push_back(Span{s0, mid});
            nd.right_type = MORT_OBJ_BVH; nd.right_idx = (int)spans.size(); spans.push_back(Span{mid, s1});
            nd.is_internal = 1;
        }
        B.nodes.push_back(nd);
        cur++;
    }
    bvhs.push_back(B);
    bvh_mode = true;                                                            // world.cuh:51-54
    return Handle{MORT_OBJ_BVH, (int)bvhs.size() - 1};
}

// ------------------------------------------------------------------------------------------------
// box helpers (utils.h:51-126)
// ------------------------------------------------------------------------------------------------
void Scene::box(V3 a, V3 b, Handle mat) {
    V3 mn(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z));
    V3 mx(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z));
    V3 dx(mx.x - mn.x, 0, 0), dy(0, mx.y - mn.y, 0), dz(0, 0, mx.z - mn.z);
    add_quad(V3(mn.x, mn.y, mx.z), dx, dy, mat);     // front
    add_quad(V3(mx.x, mn.y, mx.z), -dz, dy, mat);    // right
    add_quad(V3(mx.x, mn.y, mn.z), -dx, dy, mat);    // back
    add_quad(V3(mn.x, mn.y, mn.z), dz, dy, mat);     // left
    add_quad(V3(mn.x, mx.y, mx.z), dx, -dz, mat);    // top
    add_quad(V3(mn.x, mn.y, mn.z), dx, dz, mat);     // bottom
}

static Handle six_sided(Scene& s, V3 size, Handle mat) {
    V3 dx(size.x, 0, 0), dy(0, size.y, 0), dz(0, 0, size.z);
    Handle q[6];
    q[0] = s.add_quad(V3(0, 0, size.z), dx, dy, mat, true);
    q[1] = s.add_quad(V3(size.x, 0, size.z), -dz, dy, mat, true);
    q[2] = s.add_quad(V3(size.x, 0, 0), -dx, dy, mat, true);
    q[3] = s.add_quad(V3(0, 0, 0), dz, dy, mat, true);
    q[4] = s.add_quad(V3(0, size.y, size.z), dx, -dz, mat, true);
    q[5] = s.add_quad(V3(0, 0, 0), dx, dz, mat, true);
    Handle l = s.add_list(true);
    for (int i = 0; i < 6; i++) s.list_add(l, q[i]);
    return l;
}
Handle Scene::rotated_box(V3 size, V3 translation, float theta, Handle mat) {
    Handle l = six_sided(*this, size, mat);
    Handle rot = add_rotate_y(l, theta, true);
    return add_translate(rot, translation, false);
}
Handle Scene::rotated_smoke_box(V3 size, V3 translation, float theta, float density, Handle mat) {
    Handle l = six_sided(*this, size, mat);
    Handle rot = add_rotate_y(l, theta, true);
    Handle tr = add_translate(rot, translation, true);
    return add_constant_medium(tr, density, mat, false);
}

void Scene::clear() { *this = Scene(); }

// ------------------------------------------------------------------------------------------------
// dump / load (include/mort_scene_format.h)
// ------------------------------------------------------------------------------------------------
template <class T> static void putv(FILE* f, const std::vector<T>& v) { if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f); }

bool Scene::dump(const std::string& path) const {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    mscn_header h; memset(&h, 0, sizeof(h));
    h.magic = MSCN_MAGIC; h.version = MSCN_VERSION;
    h.n_sphere = (int)spheres.size(); h.n_quad = (int)quads.size(); h.n_translate = (int)translates.size();
    h.n_rotate_y = (int)rotates.size(); h.n_medium = (int)media.size(); h.n_list = (int)lists.size(); h.n_bvh = (int)bvhs.size();
    h.n_lambertian = (int)lambertians.size(); h.n_metal = (int)metals.size(); h.n_dielectric = (int)dielectrics.size();
    h.n_diffuse_light = (int)lights.size(); h.n_isotropic = (int)isotropics.size();
    h.n_solid = (int)solids.size(); h.n_checker = (int)checkers.size(); h.n_image = (int)images.size(); h.n_noise = (int)noises.size();
    h.bvh_mode = bvh_mode ? 1 : 0;
    fwrite(&h, sizeof(h), 1, f);
    putv(f, spheres); putv(f, quads); putv(f, translates); putv(f, rotates); putv(f, media);
    for (const ListRec& l : lists) {
        int32_t hd[2] = {l.skip, (int32_t)l.items.size()}; fwrite(hd, 4, 2, f);
        for (const Handle& it : l.items) { int32_t p[2] = {it.type, it.idx}; fwrite(p, 4, 2, f); }
    }
    for (const BvhRec& b : bvhs) {
        int32_t hd[2] = {b.skip, (int32_t)b.nodes.size()}; fwrite(hd, 4, 2, f);
        putv(f, b.nodes);
    }
    putv(f, lambertians); putv(f, metals); putv(f, dielectrics); putv(f, lights); putv(f, isotropics);
    putv(f, solids); putv(f, checkers);
    for (const ImageRec& im : images) { mscn_image r{im.width, im.height, im.fnv1a}; fwrite(&r, sizeof(r), 1, f); }
    putv(f, noises);
    mscn_camera c; cam.to_record(c); fwrite(&c, sizeof(c), 1, f);
    fclose(f);
    return true;
}

template <class T> static bool getv(FILE* f, std::vector<T>& v, int n) {
    if (n < 0) return false;
    v.resize(n);
    return n == 0 || fread(v.data(), sizeof(T), n, f) == (size_t)n;
}

bool Scene::load(const std::string& path, std::string* err) {
    clear();
    journal_complete = false;                            // arrays come from the dump, not from builder calls
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { if (err) *err = "cannot open " + path; return false; }
    mscn_header h;
    bool ok = fread(&h, sizeof(h), 1, f) == 1 && h.magic == MSCN_MAGIC && h.version == MSCN_VERSION;
    ok = ok && getv(f, spheres, h.n_sphere) && getv(f, quads, h.n_quad) && getv(f, translates, h.n_translate) &&
         getv(f, rotates, h.n_rotate_y) && getv(f, media, h.n_medium);
    for (int i = 0; ok && i < h.n_list; i++) {
        int32_t hd[2]; ok = fread(hd, 4, 2, f) == 2 && hd[1] >= 0;
        ListRec l; l.skip = hd[0];
        for (int k = 0; ok && k < hd[1]; k++) { int32_t p[2]; ok = fread(p, 4, 2, f) == 2; l.items.push_back(Handle{p[0], p[1]}); }
        lists.push_back(l);
    }
    for (int i = 0; ok && i < h.n_bvh; i++) {
        int32_t hd[2]; ok = fread(hd, 4, 2, f) == 2;
        BvhRec b; b.skip = hd[0]; ok = ok && getv(f, b.nodes, hd[1]);
        bvhs.push_back(b);
    }
    ok = ok && getv(f, lambertians, h.n_lambertian) && getv(f, metals, h.n_metal) && getv(f, dielectrics, h.n_dielectric) &&
         getv(f, lights, h.n_diffuse_light) && getv(f, isotropics, h.n_isotropic) && getv(f, solids, h.n_solid) &&
         getv(f, checkers, h.n_checker);
    for (int i = 0; ok && i < h.n_image; i++) {
        mscn_image r; ok = fread(&r, sizeof(r), 1, f) == 1;
        ImageRec im; im.width = r.width; im.height = r.height; im.fnv1a = r.fnv1a; images.push_back(im);
    }
    ok = ok && getv(f, noises, h.n_noise);
    mscn_camera c;
    ok = ok && fread(&c, sizeof(c), 1, f) == 1;
    fclose(f);
    if (!ok) { if (err) *err = "malformed scene file " + path; clear(); return false; }
    Box6 zero; memset(&zero, 0, sizeof(zero));           // the dump does not carry the wrappers' boxes (only a later add_bvh would read them)
    translate_bbox.assign(translates.size(), zero); rotate_bbox.assign(rotates.size(), zero); medium_bbox.assign(media.size(), zero);
    bvh_mode = h.bvh_mode != 0;
    cam.aspect_ratio = c.aspect_ratio; cam.image_width = c.image_width; cam.image_height = c.image_height;
    cam.samples_per_pixel = c.samples_per_pixel; cam.pixel_samples_scale = c.pixel_samples_scale;
    cam.sqrt_spp = c.sqrt_spp; cam.recip_sqrt_spp = c.recip_sqrt_spp; cam.bounce_limit = c.bounce_limit; cam.vfov = c.vfov;
    cam.background = get3(c.background); cam.light_obj_type = c.light_obj_type; cam.light_obj_idx = c.light_obj_idx;
    cam.center = get3(c.center); cam.pixel00_loc = get3(c.pixel00_loc);
    cam.pixel_delta_u = get3(c.pixel_delta_u); cam.pixel_delta_v = get3(c.pixel_delta_v);
    cam.lookfrom = get3(c.lookfrom); cam.lookat = get3(c.lookat); cam.vup = get3(c.vup);
    cam.v = get3(c.v); cam.u = get3(c.u); cam.w = get3(c.w);
    cam.defocus_angle = c.defocus_angle; cam.focus_dist = c.focus_dist;
    cam.defocus_disk_u = get3(c.defocus_disk_u); cam.defocus_disk_v = get3(c.defocus_disk_v);
    return true;
}

bool load_ppm(const std::string& path, ImageRec& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    int w = 0, h = 0, mx = 0; char magic[3] = {0, 0, 0};
    bool ok = fscanf(f, "%2s %d %d %d", magic, &w, &h, &mx) == 4 && magic[0] == 'P' && magic[1] == '6' && mx == 255 && w > 0 && h > 0;
    if (ok) {
        fgetc(f);
        out.rgb.resize((size_t)w * h * 3);
        ok = fread(out.rgb.data(), 1, out.rgb.size(), f) == out.rgb.size();
    }
    fclose(f);
    if (!ok) return false;
    out.width = w; out.height = h; out.fnv1a = fnv1a32(out.rgb.data(), out.rgb.size());
    return true;
}

}  // namespace mort
