// device_types.h — the HBM layout of a committed scene (plain structs shared by the host flattener,
// the SAH builder and the CUDA kernels).  Everything a kernel touches per ray is 16-byte vectorisable and
// 32-byte aligned:
//
//   Bvh4Node   128 B   4 child boxes in SoA (6 x float4) + 4 child words + 4 spare words; read as 8 x LDG.128
//   SphereGeom  32 B   {cx,cy,cz,r} {vx,vy,vz,inst}          hot: intersected during traversal
//   PrimInfo    16 B   {mat_gid, obj_idx, order, 0}           cold: read once for the winning primitive / on ties
//   QuadRec     96 B   {n,D} {Q,inst} {u,mat_gid} {v,order} {w,obj_idx} {area,top,-,-}; first 16 B decide most misses
//   Instance   128 B   up to 7 ops (translate / rotate_y), outermost first
//   Material    32 B   Texture 32 B
//
// Child word of a Bvh4Node:  0xFFFFFFFF empty | internal: node index (bit31 = 0)
//                            | leaf: bit31=1, bit30 = type (0 sphere, 1 quad), bits 27..29 = count-1, bits 0..26 = first record
#pragma once
#include <stdint.h>

namespace mort {

#define MORT_CHILD_EMPTY 0xFFFFFFFFu
#define MORT_LEAF_BIT 0x80000000u
#define MORT_LEAF_QUAD_BIT 0x40000000u
#define MORT_MAX_LEAF 4

struct alignas(128) Bvh4Node {
    float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
    uint32_t child[4];
    uint32_t spare[4];
};

struct alignas(32) SphereGeom { float cx, cy, cz, r; float vx, vy, vz; int32_t inst; };
struct alignas(16) PrimInfo { int32_t mat_gid, obj_idx, order, pad; };   // pad = (top-level object type << 24) | slot, for mort_trace
struct alignas(32) QuadRec {
    float nx, ny, nz, D;
    float Qx, Qy, Qz; int32_t inst;
    float ux, uy, uz; int32_t mat_gid;
    float vx, vy, vz; int32_t order;
    float wx, wy, wz; int32_t obj_idx;
    float area; int32_t pad[3];             // pad[0] = (top-level object type << 24) | slot
};

enum { INST_OP_TRANSLATE = 1, INST_OP_ROTATE_Y = 2 };
// Shading queues of the block wavefront (pool.cu).  COLD = a diffuse material whose texture tree contains a noise or image
// texture (Perlin turbulence in double precision / texel fetches): kept apart so those warps run converged instead of 4 lanes at a time.
enum { SHADE_TERMINAL = 0, SHADE_DIFFUSE = 1, SHADE_METAL = 2, SHADE_DIELECTRIC = 3, SHADE_DIFFUSE_COLD = 4, SHADE_CLASSES = 5 };
#define MORT_INSTANCE_OPS 7                 // 16 + 7 * 16 = 128 B; only the first nops rows are ever read
struct alignas(32) Instance {
    int32_t nops; int32_t pad[3];           // up to MORT_INSTANCE_OPS nested wrappers, outermost first
    float a[MORT_INSTANCE_OPS][4];                          // translate: {x,y,z, kind} ; rotate_y: {sin, cos, 0, kind}  (kind as int bits)
};

struct alignas(32) Material {               // gid = running index over lambertian|metal|dielectric|diffuse_light|isotropic
    int32_t type; int32_t tex_gid;
    float ax, ay, az;                       // metal albedo
    float p0, p1;                           // metal: fuzz ; dielectric: ior, inv_ior
    int32_t pad;
};
struct alignas(32) Texture {                // gid = running index over solid|checker|image|noise
    int32_t type;
    float c0, c1, c2;                       // solid: rgb ; checker: inv_scale
    int32_t even_gid, odd_gid;              // checker ; image: image slot in even_gid ; noise: noise slot in even_gid
    int32_t pad[2];
};
struct ImageDesc { const uint8_t* texels; int32_t width, height, cols; };   // cols = bytes per row
struct alignas(16) NoiseTables { float ranvec[256][4]; uint8_t perm_x[256], perm_y[256], perm_z[256]; float scale; float pad[3]; };

// One light-list entry (objects.cuh pdf_value/random): a sphere, a quad, or something the reference's
// dispatchers do not handle (pdf 0, direction (1,0,0)).
enum { LIGHT_NONE = 0, LIGHT_SPHERE = 1, LIGHT_QUAD = 2, LIGHT_LIST = 3, LIGHT_INVALID = 4 };
struct alignas(32) LightPrim {
    int32_t kind; float area; float D; int32_t pad;
    float a[5][4];                          // sphere: a[0] = {cx,cy,cz,r} ; quad: a[0]=n a[1]=Q a[2]=u a[3]=v a[4]=w
};

// A constant_medium: boundary primitives in the reference's visit order (a flattened list).
struct alignas(32) Medium {                  // one VISIT of a constant_medium in world::hit's order (a medium reachable twice has two records)
    double neg_inv_density;
    int32_t mat_gid; int32_t first, count;  // range in boundary[]
    int32_t obj_idx; int32_t cls;            // reference slot; shade class of the phase material
    int32_t inst;                            // wrappers above the medium (-1: none): the frame its ray length and its record are taken in
    int32_t after_lo, after_hi;              // visit-order window of the leaves between this medium and the next one (two_pass == 2)
    int32_t top_level;                       // reached by world::hit's own medium loop (mort_trace reports boundary probes for these)
    int32_t pad[5];
};
struct alignas(32) BoundaryPrim {           // same geometry as SphereGeom / QuadRec, one union-sized record
    int32_t type; int32_t inst; int32_t pad[2];
    float a[5][4];                          // sphere: a[0] = {cx,cy,cz,r}, a[1] = {vx,vy,vz,0} ; quad: n+D, Q, u, v, w
};

struct CameraParams {                        // camera.cuh fields a kernel needs
    float center[3], pixel00[3], du[3], dv[3], defocus_u[3], defocus_v[3], background[3];
    float defocus_angle, recip_sqrt_spp, pixel_samples_scale;
    int32_t width, height, sqrt_spp, bounce_limit;
};

struct DeviceScene {
    const Bvh4Node* nodes; int32_t n_nodes;
    const Bvh4Node* node_dt;                 // motion-aware bounds: box at time 1 minus box at time 0 per child (nodes then holds the time-0 boxes); else null
    const SphereGeom* spheres; const PrimInfo* sphere_info; int32_t n_spheres;
    const QuadRec* quads; int32_t n_quads;
    const uint8_t* sphere_cls; const uint8_t* quad_cls;   // shade class (SHADE_*) of each record's material: one byte decides a hit's shading queue
    const Instance* instances; int32_t n_instances;
    const Material* materials; int32_t n_materials;
    const Texture* textures; int32_t n_textures;
    const ImageDesc* images; int32_t n_images;
    const NoiseTables* noises; int32_t n_noises;
    const Medium* media; int32_t n_media;    // visits in world::hit's order
    int32_t n_media_top;                     // of which reached by world::hit's own medium loop (the ones mort_trace probes)
    const BoundaryPrim* boundary; int32_t n_boundary;
    const LightPrim* lights; int32_t n_lights; int32_t light_kind;
    int32_t post_media_order;                // leaves with order >= this are visited after the media (top-level lists)
    int32_t two_pass;                        // 1: media, then post-media leaves (a second pass); 2: media anywhere in the visit order (windowed passes)
    int32_t empty;                           // no visible primitive at all
    int32_t linear;                          // few leaves: records sorted by instance, scanned linearly (no tree)
};

}  // namespace mort
