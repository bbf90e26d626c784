// scenes.cpp — the ten shipped scenes, restated against the Scene builder, plus the synthetic
// sphere field of BASELINE.json config 4.
//
// Geometry, materials, camera and light handles follow /root/reference/mort.cu:129-631 (dispatch
// mort.cu:649-689).  Two things that are NOT visible in that source but decide the bits of the layout
// are reproduced deliberately and pinned by tests/test_host_scene.py against the reference's own dumps:
//   * the host random stream is glibc rand() unseeded (HostRng), and
//   * draws that appear inside one argument list are consumed right-to-left,
//     so every draw below is bound to a named temporary in the order the reference consumes it.
// Mixed float/double arithmetic (e.g. `a + 0.9 * random_float()` is evaluated in double and then
// narrowed by the point3 constructor) is kept expression by expression.
#include <cmath>

#include "scene.hpp"

namespace mort {

static V3 rand_vec(HostRng& g) {                        // vec3::random(), vec3.cuh:63-65: z, y, x
    float z = g.random_float(), y = g.random_float(), x = g.random_float();
    return V3(x, y, z);
}
static V3 rand_vec(HostRng& g, float lo, float hi) {    // vec3::random(min,max), vec3.cuh:67-69
    float z = g.random_float(lo, hi), y = g.random_float(lo, hi), x = g.random_float(lo, hi);
    return V3(x, y, z);
}

static void book_camera(Camera& cam) {                  // shared by scenes 1, 2, 4, 10
    cam.vfov = 20;
    cam.lookfrom = V3(13, 2, 3);
    cam.lookat = V3(0, 0, 0);
    cam.vup = V3(0, 1, 0);
    cam.defocus_angle = 0;
}

static void random_spheres(Scene& s, HostRng& g) {      // mort.cu:129-226
    Handle spheres = s.add_list(true);
    Handle c1 = s.add_solid(V3(.2f, .3f, .1f));
    Handle c2 = s.add_solid(V3(.9f, .9f, .9f));
    Handle checker = s.add_checker(0.32f, c1, c2);
    Handle ground_mat = s.add_lambertian(checker);
    s.list_add(spheres, s.add_sphere(V3(0, -1000, 0), 1000, ground_mat, true));

    for (int a = -11; a < 11; a++) {
        for (int b = -11; b < 11; b++) {
            float choose_mat = g.random_float();
            float rz = g.random_float(), rx = g.random_float();
            V3 center((float)(a + 0.9 * rx), 0.2f, (float)(b + 0.9 * rz));
            if ((double)length(center - V3(4, 0.2f, 0)) > 0.9) {
                if (choose_mat < 0.8) {
                    V3 q = rand_vec(g), p = rand_vec(g);     // right operand first; the product commutes anyway
                    V3 albedo = p * q;
                    V3 center2 = center + V3(0, g.random_float(0.0f, 0.5f), 0);
                    Handle mat = s.add_lambertian(s.add_solid(albedo));
                    s.list_add(spheres, s.add_moving_sphere(center, center2, 0.2f, mat, true));
                } else if (choose_mat < 0.95) {
                    V3 albedo = rand_vec(g, 0.5f, 1);
                    float fuzz = g.random_float(0.0f, 0.5f);
                    s.list_add(spheres, s.add_sphere(center, 0.2f, s.add_metal(albedo, fuzz), true));
                } else {
                    s.list_add(spheres, s.add_sphere(center, 0.2f, s.add_dielectric(1.5f), true));
                }
            }
        }
    }
    s.list_add(spheres, s.add_sphere(V3(0, 1, 0), 1.0f, s.add_dielectric(1.5f), true));
    s.list_add(spheres, s.add_sphere(V3(-4, 1, 0), 1.0f, s.add_lambertian(s.add_solid(V3(0.4f, 0.2f, 0.1f))), true));
    s.list_add(spheres, s.add_sphere(V3(4, 1, 0), 1.0f, s.add_metal(V3(0.7f, 0.6f, 0.5f), 0.0f), true));
    s.add_bvh(spheres, false);

    Camera& cam = s.cam;
    cam.aspect_ratio = (float)(16.0 / 9.0); cam.image_width = 1200; cam.samples_per_pixel = 100; cam.bounce_limit = 20;
    cam.light_obj_type = -1;
    book_camera(cam);
    cam.focus_dist = 10.0f;
}

static void two_spheres(Scene& s) {                     // mort.cu:228-253
    Handle c1 = s.add_solid(V3(.2f, .3f, .1f));
    Handle c2 = s.add_solid(V3(.9f, .9f, .9f));
    Handle mat = s.add_lambertian(s.add_checker(0.32f, c1, c2));
    s.add_sphere(V3(0, -10, 0), 10, mat);
    s.add_sphere(V3(0, 10, 0), 10, mat);
    Camera& cam = s.cam;
    cam.aspect_ratio = (float)(16.0 / 9.0); cam.image_width = 1200; cam.samples_per_pixel = 20; cam.bounce_limit = 50;
    cam.light_obj_type = -1;
    book_camera(cam);
}

static void out_of_order_spheres(Scene& s, HostRng& g, int n) {   // mort.cu:255-290
    Handle spheres = s.add_list(true);
    for (int i = 0; i < n; i++) {
        V3 q = rand_vec(g), p = rand_vec(g);
        V3 albedo = p * q;
        float c = (float)(n - i);
        Handle mat = s.add_lambertian(s.add_solid(albedo));
        s.list_add(spheres, s.add_sphere(V3(c, c, c), 0.2f, mat, true));
    }
    s.add_bvh(spheres, false);
    Camera& cam = s.cam;
    cam.aspect_ratio = (float)(16.0 / 9.0); cam.image_width = 1200; cam.samples_per_pixel = 1; cam.bounce_limit = 5;
    cam.light_obj_type = -1;
    book_camera(cam);
    cam.focus_dist = 10.0f;
}

static bool earth(Scene& s, const std::string& assets) {           // mort.cu:292-313
    ImageRec im;
    if (!load_ppm(assets + "/earthmap.ppm", im)) { s.error = "cannot load " + assets + "/earthmap.ppm"; return false; }
    Handle tex = s.add_image(im.rgb.data(), im.width, im.height, "earthmap.ppm");
    s.add_sphere(V3(0, 0, 0), 2, s.add_lambertian(tex));
    Camera& cam = s.cam;
    cam.aspect_ratio = (float)(16.0 / 9.0); cam.image_width = 1200; cam.samples_per_pixel = 100; cam.bounce_limit = 50;
    cam.light_obj_type = -1;
    cam.vfov = 20; cam.lookfrom = V3(0, 0, 12); cam.lookat = V3(0, 0, 0); cam.vup = V3(0, 1, 0); cam.defocus_angle = 0;
    return true;
}

static void two_perlin_spheres(Scene& s, HostRng& g) {             // mort.cu:315-338
    Handle mat = s.add_lambertian(s.add_noise(4.0f, g));            // reference reads an unset idx here; slot 0 (A-Q9)
    s.add_sphere(V3(0, -1000, 0), 1000, mat);
    s.add_sphere(V3(0, 2, 0), 2, mat);
    Camera& cam = s.cam;
    cam.aspect_ratio = (float)(16.0 / 9.0); cam.image_width = 1200; cam.samples_per_pixel = 5; cam.bounce_limit = 10;
    cam.light_obj_type = -1;
    book_camera(cam);
}

static void quads(Scene& s) {                                       // mort.cu:340-390
    Handle red = s.add_solid(V3(1.0f, 0.2f, 0.2f)), green = s.add_solid(V3(0.2f, 1.0f, 0.2f));
    Handle blue = s.add_solid(V3(0.2f, 0.2f, 1.0f)), orange = s.add_solid(V3(1.0f, 0.5f, 0.0f));
    Handle teal = s.add_solid(V3(0.2f, 0.8f, 0.8f));
    Handle left = s.add_lambertian(red), back = s.add_lambertian(green), right = s.add_lambertian(blue);
    Handle upper = s.add_lambertian(orange), lower = s.add_lambertian(teal);
    s.add_quad(V3(-3, -2, 5), V3(0, 0, -4), V3(0, 4, 0), left);
    s.add_quad(V3(-2, -2, 0), V3(4, 0, 0), V3(0, 4, 0), back);
    s.add_quad(V3(3, -2, 1), V3(0, 0, 4), V3(0, 4, 0), right);
    s.add_quad(V3(-2, 3, 1), V3(4, 0, 0), V3(0, 0, 4), upper);
    s.add_quad(V3(-2, -3, 5), V3(4, 0, 0), V3(0, 0, -4), lower);
    Camera& cam = s.cam;
    cam.aspect_ratio = 1.0f; cam.image_width = 400; cam.samples_per_pixel = 100; cam.bounce_limit = 50;
    cam.light_obj_type = -1;
    cam.vfov = 20; cam.lookfrom = V3(0, 0, 9); cam.lookat = V3(0, 0, 0); cam.vup = V3(0, 1, 0); cam.defocus_angle = 0;
}

static void cornell_camera(Camera& cam) {
    cam.aspect_ratio = 1.0f; cam.background = V3(0, 0, 0);
    cam.vfov = 40; cam.lookfrom = V3(278, 278, -800); cam.lookat = V3(278, 278, 0); cam.vup = V3(0, 1, 0);
    cam.defocus_angle = 0;
}

static void cornell_box(Scene& s) {                                 // mort.cu:392-448
    Handle red = s.add_solid(V3(.65f, .05f, .05f)), white = s.add_solid(V3(.73f, .73f, .73f));
    Handle green = s.add_solid(V3(.12f, .45f, .15f)), light = s.add_solid(V3(15, 15, 10));
    Handle red_wall = s.add_lambertian(red), white_wall = s.add_lambertian(white), green_wall = s.add_lambertian(green);
    Handle lamp = s.add_diffuse_light(light);
    Handle glass = s.add_dielectric(1.5f);

    Handle lights = s.add_list(false);
    s.list_add(lights, s.add_quad(V3(343, 554, 332), V3(-130, 0, 0), V3(0, 0, -105), lamp, true));
    s.list_add(lights, s.add_sphere(V3(190, 90, 190), 90, glass, true));

    s.add_quad(V3(555, 0, 0), V3(0, 555, 0), V3(0, 0, 555), green_wall);
    s.add_quad(V3(0, 0, 0), V3(0, 555, 0), V3(0, 0, 555), red_wall);
    s.add_quad(V3(0, 0, 0), V3(555, 0, 0), V3(0, 0, 555), white_wall);
    s.add_quad(V3(555, 555, 555), V3(-555, 0, 0), V3(0, 0, -555), white_wall);
    s.add_quad(V3(0, 0, 555), V3(555, 0, 0), V3(0, 555, 0), white_wall);
    s.rotated_box(V3(165, 330, 165), V3(265, 0, 295), 15, white_wall);

    Camera& cam = s.cam;
    cam.image_width = 600; cam.samples_per_pixel = 1000; cam.bounce_limit = 50;
    cornell_camera(cam);
    cam.light_obj_type = lights.type; cam.light_obj_idx = lights.idx;
}

static void cornell_smoke(Scene& s) {                               // mort.cu:450-504
    Handle red = s.add_solid(V3(.65f, .05f, .05f)), white = s.add_solid(V3(.73f, .73f, .73f));
    Handle green = s.add_solid(V3(.12f, .45f, .15f)), light = s.add_solid(V3(15, 15, 10));
    Handle black_c = s.add_solid(V3(0, 0, 0)), white_c = s.add_solid(V3(1, 1, 1));
    Handle red_wall = s.add_lambertian(red), white_wall = s.add_lambertian(white), green_wall = s.add_lambertian(green);
    Handle lamp = s.add_diffuse_light(light);
    Handle black_smoke = s.add_lambertian(black_c), white_smoke = s.add_lambertian(white_c);

    s.add_quad(V3(555, 0, 0), V3(0, 555, 0), V3(0, 0, 555), green_wall);
    s.add_quad(V3(0, 0, 0), V3(0, 555, 0), V3(0, 0, 555), red_wall);
    s.add_quad(V3(343, 554, 332), V3(-130, 0, 0), V3(0, 0, -105), lamp);
    s.add_quad(V3(0, 0, 0), V3(555, 0, 0), V3(0, 0, 555), white_wall);
    s.add_quad(V3(555, 555, 555), V3(-555, 0, 0), V3(0, 0, -555), white_wall);
    s.add_quad(V3(0, 0, 555), V3(555, 0, 0), V3(0, 555, 0), white_wall);
    s.rotated_smoke_box(V3(165, 330, 165), V3(265, 0, 295), 15, 0.01f, black_smoke);
    s.rotated_smoke_box(V3(165, 165, 165), V3(130, 0, 65), -18, 0.01f, white_smoke);

    Camera& cam = s.cam;
    cam.image_width = 800; cam.samples_per_pixel = 2000; cam.bounce_limit = 50;
    cornell_camera(cam);
    // mort.cu:495-496 stores the lamp MATERIAL's handle here — (4,0), which the object dispatchers read
    // as rotate_y #0.  Kept bug-compatible: parity is against the reference as it is.
    cam.light_obj_type = lamp.type; cam.light_obj_idx = lamp.idx;
}

static bool final_scene(Scene& s, HostRng& g, const std::string& assets, int image_width, int spp, int max_depth) {  // mort.cu:506-631
    Handle ground_mat = s.add_lambertian(s.add_solid(V3(0.48f, 0.83f, 0.53f)));
    const int boxes_per_side = 20;
    for (int i = 0; i < boxes_per_side; i++)
        for (int j = 0; j < boxes_per_side; j++) {
            double w = 100.0, x0 = -1000.0 + i * w, z0 = -1000.0 + j * w, y0 = 0.0, x1 = x0 + w, z1 = z0 + w;
            float y1 = g.random_float(1, 101);
            s.box(V3((float)x0, (float)y0, (float)z0), V3((float)x1, y1, (float)z1), ground_mat);
        }

    Handle light_mat = s.add_diffuse_light(s.add_solid(V3(7, 7, 7)));
    Handle light = s.add_quad(V3(123, 554, 147), V3(300, 0, 0), V3(0, 0, 265), light_mat);

    V3 center1(400, 400, 200), center2 = center1 + V3(30, 0, 0);
    s.add_moving_sphere(center1, center2, 50, s.add_lambertian(s.add_solid(V3(0.7f, 0.3f, 0.1f))));

    Handle glass = s.add_dielectric(1.5f);
    s.add_sphere(V3(260, 150, 45), 50, glass);
    s.add_sphere(V3(0, 150, 145), 50, s.add_metal(V3(0.8f, 0.8f, 0.9f), 1.0f));

    // "subsurface" sphere: glass boundary that is also visible, filled with a lambertian-scattering medium (A-Q7)
    Handle sub_mat = s.add_lambertian(s.add_solid(V3(0.2f, 0.4f, 0.9f)));
    Handle sub = s.add_sphere(V3(360, 150, 145), 70, glass);
    s.add_constant_medium(sub, 0.2f, sub_mat);

    Handle fog_mat = s.add_lambertian(s.add_solid(V3(1, 1, 1)));
    Handle fog_boundary = s.add_sphere(V3(0, 0, 0), 5000, glass);
    s.add_constant_medium(fog_boundary, 0.0001f, fog_mat);

    ImageRec im;
    if (!load_ppm(assets + "/earthmap.ppm", im)) { s.error = "cannot load " + assets + "/earthmap.ppm"; return false; }
    s.add_sphere(V3(400, 200, 400), 100, s.add_lambertian(s.add_image(im.rgb.data(), im.width, im.height, "earthmap.ppm")));

    s.add_sphere(V3(220, 280, 300), 80, s.add_lambertian(s.add_noise(0.1f, g)));

    Handle cluster_color = s.add_solid(V3(.73f, .73f, .73f));
    Handle cluster_mat = s.add_lambertian(cluster_color);
    Handle cluster = s.add_list(true);
    for (int j = 0; j < 1000; j++) s.list_add(cluster, s.add_sphere(rand_vec(g, 0, 165), 10, cluster_mat, true));
    s.add_translate(s.add_rotate_y(cluster, 15, true), V3(-100, 270, 395));

    Camera& cam = s.cam;
    cam.aspect_ratio = 1.0f; cam.image_width = image_width; cam.samples_per_pixel = spp; cam.bounce_limit = max_depth;
    cam.background = V3(0, 0, 0);
    cam.light_obj_type = light.type; cam.light_obj_idx = light.idx;
    cam.vfov = 40; cam.lookfrom = V3(478, 278, -600); cam.lookat = V3(278, 278, 0); cam.vup = V3(0, 1, 0); cam.defocus_angle = 0;
    return true;
}

bool build_reference_scene(Scene& s, int scene_id, const std::string& asset_dir) {
    s.clear();
    HostRng g(1);
    bool ok = true;
    switch (scene_id) {
        case 1: random_spheres(s, g); break;
        case 2: two_spheres(s); break;
        case 3: ok = earth(s, asset_dir); break;
        case 4: two_perlin_spheres(s, g); break;
        case 5: quads(s); break;
        case 6: cornell_box(s); break;
        case 7: cornell_smoke(s); break;
        case 8: ok = final_scene(s, g, asset_dir, 800, 1000, 40); break;
        case 9: ok = final_scene(s, g, asset_dir, 400, 250, 4); break;
        case 10: out_of_order_spheres(s, g, 35); break;
        default: break;     // mort.cu:644: the range check is vacuous; any other index renders the empty world
    }
    s.cam.initialize();
    return ok;
}

// ------------------------------------------------------------------------------------------------
// Synthetic field: the scene-1 recipe over cells [-G,G)^2 (G = 500 -> 10^6 candidate spheres), with its
// own seeded SplitMix64 stream (not rand()).  No reference counterpart: the reference's fixed capacities
// (objects.cuh:451,521,746) cannot hold it.  No reference-style BVH either (its O(n^2) sort is hopeless at
// this size); all spheres are plain visible top-level objects for the SAH builder.
// ------------------------------------------------------------------------------------------------
static inline uint64_t splitmix64(uint64_t& st) {
    uint64_t z = (st += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline float sm_u01(uint64_t& st) { return (float)(splitmix64(st) >> 40) * (1.0f / 16777216.0f); }

bool build_sphere_field(Scene& s, int G, uint64_t seed, int camera_kind) {
    s.clear();
    if (G < 1) { s.error = "sphere field: G must be >= 1"; return false; }
    uint64_t st = seed;
    Handle c1 = s.add_solid(V3(.2f, .3f, .1f)), c2 = s.add_solid(V3(.9f, .9f, .9f));
    // ground: a flat checkered quad (a radius-1000 sphere as in scene 1 would drop 134 units by x = 500).  It sits
    // half a millimetre below y = 0: the 3-D checker takes floor(y / 0.32), and on a plane exactly at y = 0 the
    // sign of the rounding noise of the hit point would pick the colour.
    float L = (float)G + 50.0f;
    s.add_quad(V3(-L, -0.0005f, -L), V3(0, 0, 2 * L), V3(2 * L, 0, 0), s.add_lambertian(s.add_checker(0.32f, c1, c2)));
    Handle glass = s.add_dielectric(1.5f);
    // a palette of materials instead of one per sphere: 10^6 spheres share 4096 lambertians / 1024 metals
    std::vector<Handle> lam, met;
    for (int i = 0; i < 4096; i++) {
        V3 a(sm_u01(st) * sm_u01(st), sm_u01(st) * sm_u01(st), sm_u01(st) * sm_u01(st));
        lam.push_back(s.add_lambertian(s.add_solid(a)));
    }
    for (int i = 0; i < 1024; i++) {
        V3 a(0.5f + 0.5f * sm_u01(st), 0.5f + 0.5f * sm_u01(st), 0.5f + 0.5f * sm_u01(st));
        met.push_back(s.add_metal(a, 0.5f * sm_u01(st)));
    }
    for (int a = -G; a < G; a++)
        for (int b = -G; b < G; b++) {
            float choose = sm_u01(st);
            V3 center((float)(a + 0.9 * sm_u01(st)), 0.2f, (float)(b + 0.9 * sm_u01(st)));
            uint32_t pick = (uint32_t)(splitmix64(st) >> 32);
            if ((double)length(center - V3(4, 0.2f, 0)) > 0.9) {
                if (choose < 0.8f) {
                    V3 c2v = center + V3(0, 0.5f * sm_u01(st), 0);
                    s.add_moving_sphere(center, c2v, 0.2f, lam[pick % lam.size()]);
                } else if (choose < 0.95f) {
                    s.add_sphere(center, 0.2f, met[pick % met.size()]);
                } else {
                    s.add_sphere(center, 0.2f, glass);
                }
            }
        }
    s.add_sphere(V3(0, 1, 0), 1.0f, glass);
    s.add_sphere(V3(-4, 1, 0), 1.0f, s.add_lambertian(s.add_solid(V3(0.4f, 0.2f, 0.1f))));
    s.add_sphere(V3(4, 1, 0), 1.0f, s.add_metal(V3(0.7f, 0.6f, 0.5f), 0.0f));

    Camera& cam = s.cam;
    cam.aspect_ratio = (float)(16.0 / 9.0); cam.image_width = 1920; cam.samples_per_pixel = 256; cam.bounce_limit = 50;
    cam.light_obj_type = -1;
    cam.vup = V3(0, 1, 0); cam.defocus_angle = 0; cam.focus_dist = 10.0f;
    if (camera_kind == 0) { cam.vfov = 20; cam.lookfrom = V3(13, 2, 3); cam.lookat = V3(0, 0, 0); }
    else { cam.vfov = 40; cam.lookfrom = V3(0, 0.6f * G, 1.2f * G); cam.lookat = V3(0, 0, 0); }
    cam.initialize();
    return true;
}

}  // namespace mort
