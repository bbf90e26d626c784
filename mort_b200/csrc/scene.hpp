// scene.hpp — host-side scene model of mort-b200.
//
// Mirrors the reference's scene-builder surface (value handles = (type tag, array slot); one growable
// array per hittable / material / texture kind; a Camera that is a bag of public fields) so scene code
// written against the reference maps 1:1 onto mort_add_* calls:
//   world::add overloads          /root/reference/world.cuh:27-90
//   object ctors                  /root/reference/objects.cuh:38,46,170,258,296,384,459-469,529
//   material / texture ctors      /root/reference/materials.cuh:36,71,104,149,180; textures.cuh:20,42,79,164
//   Camera fields + initialize()  /root/reference/camera.cuh:12-84
// Unlike the reference there are no fixed capacities (objects.cuh:451,521,746-764): everything is a
// std::vector.  Records reuse the on-disk structs of include/mort_scene_format.h so a scene can be
// dumped and compared bit-for-bit with the reference's.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "mort_scene_format.h"

namespace mort {

struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
// All host vector math is plain single-precision, one rounding per operation (no contraction): the
// reference's host code is built for baseline x86-64, which has no FMA.
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator*(float t, V3 v) { return V3(t * v.x, t * v.y, t * v.z); }
inline V3 operator/(V3 v, float t) { return (1 / t) * v; }          // vec3.cuh:109-112: multiply by reciprocal
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
float length(V3 v);
inline V3 unit(V3 v) { return v / length(v); }

// glibc's rand() (TYPE_3 additive feedback, r[i] = r[i-3] + r[i-31], unseeded => seed 1), restated so
// that the scene layouts do not depend on the libc in use.  The reference never calls srand()
// (SURVEY.md App. A-Q12); its scene layout on Linux is this stream.
class HostRng {
public:
    explicit HostRng(uint32_t seed = 1) { reseed(seed); }
    void reseed(uint32_t seed);
    int next();                                         // rand()
    float random_float();                               // rng.cuh:44-47
    float random_float(float lo, float hi);             // rng.cuh:49-53
    uint64_t draws() const { return draws_; }           // rand() calls since the last reseed (scene text files pin noise tables to it)
    void skip(uint64_t n) { while (n--) next(); }
private:
    int32_t r_[34];
    int f_, b_;
    uint64_t draws_ = 0;
};

struct Handle { int type; int idx; };

struct Camera {                                        // camera.cuh:12-45 (defaults included)
    float aspect_ratio = 1.0f;
    int image_width = 1500;
    int image_height = 0;
    int samples_per_pixel = 50;
    float pixel_samples_scale = 0;
    int sqrt_spp = 0;
    float recip_sqrt_spp = 0;
    int bounce_limit = 10;
    int vfov = 90;
    V3 background = V3(0.70f, 0.80f, 1.00f);
    int light_obj_type = -1, light_obj_idx = 0;
    V3 center, pixel00_loc, pixel_delta_u, pixel_delta_v;
    V3 lookfrom = V3(0, 0, 1), lookat = V3(0, 0, 0), vup = V3(0, 1, 0);
    V3 v, u, w;
    float defocus_angle = 0, focus_dist = 10;
    V3 defocus_disk_u, defocus_disk_v;
    void initialize();                                  // camera.cuh:47-84
    void to_record(mscn_camera& c) const;
};

struct ListRec { int skip = 0; std::vector<Handle> items; float bbox[6] = {0, 0, 0, 0, 0, 0}; bool has_bbox = false; };
struct BvhRec { int skip = 0; int list_idx = -1; std::vector<mscn_bvh_node> nodes; };
struct ImageRec { int width = 0, height = 0; std::vector<uint8_t> rgb; uint32_t fnv1a = 0; std::string source; /* file name inside the asset dir, if any */ };

class Scene {
public:
    // ---- textures (textures.cuh) ----
    Handle add_solid(V3 c);
    Handle add_checker(float scale, Handle even, Handle odd);
    Handle add_image(const uint8_t* rgb, int width, int height, const char* source = nullptr);   // RGB8, rows top-down (stb order)
    Handle add_noise(float scale, HostRng& rng);                     // draws 3*256 + 3*255 host randoms
    Handle add_noise_tables(const mscn_noise& n);                    // explicit tables (scene files / tests)
    // ---- materials (materials.cuh) ----
    Handle add_lambertian(Handle tex);
    Handle add_metal(V3 albedo, float fuzz);
    Handle add_dielectric(float ior);
    Handle add_diffuse_light(Handle tex);
    Handle add_isotropic(Handle tex);
    // ---- hittables (objects.cuh) ----
    Handle add_sphere(V3 c, float r, Handle mat, bool skip = false);
    Handle add_moving_sphere(V3 c1, V3 c2, float r, Handle mat, bool skip = false);
    bool update_sphere(int idx, V3 c1, const V3* c2, float r);          // dynamic scenes: new centre(s) / radius for an existing sphere (marks the scene edited)
    Handle add_quad(V3 Q, V3 u, V3 v, Handle mat, bool skip = false);
    Handle add_translate(Handle obj, V3 offset, bool skip = false);
    Handle add_rotate_y(Handle obj, float theta_deg, bool skip = false);
    Handle add_constant_medium(Handle boundary, float density, Handle mat, bool skip = false);
    Handle add_list(bool skip);
    int list_add(Handle list, Handle obj);
    Handle add_bvh(Handle list, bool skip);             // reference's median-split build incl. its physical sort
    // helpers of utils.h:51-126
    void box(V3 a, V3 b, Handle mat);
    Handle rotated_box(V3 size, V3 translation, float theta, Handle mat);
    Handle rotated_smoke_box(V3 size, V3 translation, float theta, float density, Handle mat);

    void clear();
    bool dump(const std::string& path) const;           // include/mort_scene_format.h
    bool load(const std::string& path, std::string* err);  // inverse of dump (images resolved by the caller)

    bool bbox_of(Handle h, float out[6]) const;         // reference's host_getBboxInfo (objects.cuh:918-945)

    std::vector<mscn_sphere> spheres;
    std::vector<mscn_quad> quads;
    std::vector<mscn_translate> translates;
    std::vector<mscn_rotate_y> rotates;
    std::vector<mscn_medium> media;
    // bounding boxes the reference computes in the wrappers' constructors (objects.cuh:258-266, 296-330, 384-393); they are not part
    // of the scene dump and only matter to the reference's own BVH build (node boxes, sort order) when a wrapper is put in a bvh
    struct Box6 { float b[6]; };
    std::vector<Box6> translate_bbox, rotate_bbox, medium_bbox;
    std::vector<ListRec> lists;
    std::vector<BvhRec> bvhs;
    std::vector<mscn_lambertian> lambertians;
    std::vector<mscn_metal> metals;
    std::vector<mscn_dielectric> dielectrics;
    std::vector<mscn_diffuse_light> lights;
    std::vector<mscn_isotropic> isotropics;
    std::vector<mscn_solid> solids;
    std::vector<mscn_checker> checkers;
    std::vector<ImageRec> images;
    std::vector<mscn_noise> noises;
    bool bvh_mode = false;
    bool edited = false;                                 // a primitive changed after construction (mort_update_sphere): the reference-BVH boxes are stale
    Camera cam;
    std::string error;
    // Journal of the builder calls in the order they were made, one scene-text statement each (scene_text.cpp).
    // Replaying it rebuilds the same arrays slot for slot; scenes loaded from a binary dump have none.
    std::vector<std::string> journal;
    bool journal_complete = true;
    int journal_mute = 0;                                // > 0 inside helpers that log themselves
};

// The ten shipped scenes (mort.cu:129-631, dispatch mort.cu:649-689).  `asset_dir` holds earthmap.ppm.
bool build_reference_scene(Scene& s, int scene_id, const std::string& asset_dir);
// BASELINE.json config 4: scene-1 recipe generalised to cells [-G,G)^2 with an own seeded generator.
// camera_kind 0 = book view (13,2,3) vfov 20; 1 = aerial.
bool build_sphere_field(Scene& s, int G, uint64_t seed, int camera_kind);
bool load_ppm(const std::string& path, ImageRec& out);
// scene_text.cpp — line-oriented scene description, one mort_add_* call per statement (grammar in the file header)
bool load_scene_text(Scene& s, HostRng& rng, const std::string& path, const std::string& asset_dir, std::string* err);
bool dump_scene_text(const Scene& s, const std::string& path, std::string* err);
// canonical handle spelling: kind prefix + array slot ("sph12", "lam0", "sol3"); category 'o' object, 'm' material, 't' texture
std::string handle_token(Handle h, char category);
uint32_t fnv1a32(const uint8_t* p, size_t n);

}  // namespace mort
