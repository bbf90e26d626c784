/* mort_scene_format.h — on-disk layouts shared by the reference harness (oracle/ref_harness.cu),
 * the CPU oracle (oracle/mort_oracle.c), the product library (mort_dump_scene) and the tests.
 *
 * These are DATA layouts only (no algorithm): a field-by-field, little-endian flattening of the
 * reference's host-side scene arrays as they stand right before world::toDevice()
 * (/root/reference/world.cuh:98-102), i.e. after every scene function has run and after the
 * reference's BVH build has physically permuted the sphere array (/root/reference/objects.cuh:630-661).
 * Primitive ids everywhere in this repo are (type tag, array slot) in exactly these arrays.
 *
 * Type tags are the reference's: objects /root/reference/objects.cuh:13-19,
 * materials /root/reference/materials.cuh:14-18, textures /root/reference/textures.cuh:10-13.
 */
#ifndef MORT_SCENE_FORMAT_H
#define MORT_SCENE_FORMAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSCN_MAGIC 0x4E43534Du /* "MSCN" */
#define MSCN_VERSION 1u

enum { MORT_OBJ_SPHERE = 1, MORT_OBJ_QUAD = 2, MORT_OBJ_TRANSLATE = 3, MORT_OBJ_ROTATE_Y = 4,
       MORT_OBJ_CONSTANT_MEDIUM = 5, MORT_OBJ_HITTABLE_LIST = 6, MORT_OBJ_BVH = 7 };
enum { MORT_MAT_LAMBERTIAN = 1, MORT_MAT_METAL = 2, MORT_MAT_DIELECTRIC = 3,
       MORT_MAT_DIFFUSE_LIGHT = 4, MORT_MAT_ISOTROPIC = 5 };
enum { MORT_TEX_SOLID = 1, MORT_TEX_CHECKER = 2, MORT_TEX_IMAGE = 3, MORT_TEX_NOISE = 4 };

#define MORT_PERLIN_POINTS 256 /* /root/reference/textures.cuh:158 */

typedef struct {
    uint32_t magic, version;
    int32_t n_sphere, n_quad, n_translate, n_rotate_y, n_medium, n_list, n_bvh;
    int32_t n_lambertian, n_metal, n_dielectric, n_diffuse_light, n_isotropic;
    int32_t n_solid, n_checker, n_image, n_noise;
    int32_t bvh_mode;
    int32_t reserved[7];
} mscn_header; /* 26 words */

/* bbox order everywhere: xmin,xmax,ymin,ymax,zmin,zmax (interval-per-axis like aabb.cuh:8) */
typedef struct { float center[3]; float radius; int32_t moves; float center_vec[3];
                 int32_t mat_type, mat_idx, skip; float bbox[6]; } mscn_sphere;        /* 17 words */
typedef struct { float Q[3], u[3], v[3], normal[3], w[3]; float D, area;
                 int32_t mat_type, mat_idx, skip; float bbox[6]; } mscn_quad;          /* 26 words */
typedef struct { int32_t obj_type, obj_idx; float offset[3]; int32_t skip; } mscn_translate;
typedef struct { int32_t obj_type, obj_idx; float sin_theta, cos_theta; int32_t skip; } mscn_rotate_y;
typedef struct { int32_t obj_type, obj_idx; double neg_inv_density;
                 int32_t mat_type, mat_idx, skip, pad; } mscn_medium;
/* list:  int32 skip, int32 num, then num x {int32 type, int32 idx}                       */
/* bvh:   int32 skip, int32 n_nodes, then n_nodes x mscn_bvh_node                          */
typedef struct { int32_t left_type, left_idx, right_type, right_idx, is_internal;
                 float bbox[6]; } mscn_bvh_node;                                        /* 11 words */

typedef struct { int32_t tex_type, tex_idx; } mscn_lambertian;
typedef struct { float albedo[3]; float fuzz; } mscn_metal;
typedef struct { float ior, inv_ior; float albedo[3]; } mscn_dielectric;
typedef struct { int32_t tex_type, tex_idx; } mscn_diffuse_light;
typedef struct { int32_t tex_type, tex_idx; } mscn_isotropic;

typedef struct { float color[3]; } mscn_solid;
typedef struct { float inv_scale; int32_t even_type, even_idx, odd_type, odd_idx; } mscn_checker;
/* image pixels are not embedded: RGB8 rows top-down, looked up by (width,height,fnv1a32 of bytes) */
typedef struct { int32_t width, height; uint32_t fnv1a; } mscn_image;
typedef struct { float scale; float ranvec[MORT_PERLIN_POINTS][3];
                 int32_t perm_x[MORT_PERLIN_POINTS], perm_y[MORT_PERLIN_POINTS],
                         perm_z[MORT_PERLIN_POINTS]; } mscn_noise;

/* Every field of the reference's Camera after initialize() (/root/reference/camera.cuh:12-84),
 * minus the four device scratch pointers. */
typedef struct {
    float aspect_ratio; int32_t image_width, image_height, samples_per_pixel;
    float pixel_samples_scale; int32_t sqrt_spp; float recip_sqrt_spp;
    int32_t bounce_limit, vfov; float background[3];
    int32_t light_obj_type, light_obj_idx;
    float center[3], pixel00_loc[3], pixel_delta_u[3], pixel_delta_v[3];
    float lookfrom[3], lookat[3], vup[3], v[3], u[3], w[3];
    float defocus_angle, focus_dist; float defocus_disk_u[3], defocus_disk_v[3];
} mscn_camera; /* 52 words */

/* ---- primary-hit parity records (written by the harness' --trace, by mort_trace tests) ---- */
#define MHIT_MAGIC 0x5449484Du /* "MHIT" */
typedef struct {
    int32_t hit;            /* 0/1: world::hit with the medium loop disabled                  */
    float   t;              /* rec.t                                                          */
    int32_t leaf_type, leaf_idx;   /* sphere/quad (type, slot) that produced it, -1 if miss   */
    int32_t top_type, top_idx;     /* top-level object world::hit reached it through          */
    int32_t mat_type, mat_idx, front_face;
    int32_t flags;          /* bit0: leaf matched by nearest t, not bit-equal t; bit1: >1 leaf at equal t */
    float   p[3], normal[3], u, v;
} mhit_record; /* 18 words */
typedef struct { int32_t hit1, hit2; float t1, t2; } mhit_medium_probe;

#ifdef __cplusplus
}
#endif
#endif
