// flatten.hpp — host side of mort_commit(): turns the reference-shaped Scene (handles, lists, instance
// wrappers, media, the reference's own BVH objects) into the flat HBM layout of device_types.h and builds
// the binned-SAH 4-wide BVH over every visible leaf primitive.
#pragma once
#include <string>
#include <vector>

#include "device_types.h"
#include "scene.hpp"

namespace mort {

// One visible sphere / quad as world::hit (world.cuh:104-171) reaches it.  `order` is its position in the
// reference's visit sequence: on bit-equal t the reference keeps the LATER candidate (objects.cuh:73-76,197),
// so "largest order wins" reproduces its tie rule under any traversal order.
struct LeafRef { int type, idx, inst, order, top_type, top_idx; };

struct BuildStats {
    int n_leaves = 0, n_nodes = 0, n_bvh2_nodes = 0, max_depth = 0, n_leaf_slots = 0;
    double sah_cost = 0, build_ms = 0;
    float pad = 0, scene_extent = 0;
};

struct FlatScene {
    std::vector<Bvh4Node> nodes;
    std::vector<SphereGeom> spheres; std::vector<PrimInfo> sphere_info;
    std::vector<QuadRec> quads;
    std::vector<uint8_t> sphere_cls, quad_cls;   // SHADE_* per record
    std::vector<Instance> instances;
    std::vector<Material> materials;
    std::vector<Texture> textures;
    std::vector<NoiseTables> noises;
    std::vector<Medium> media;
    std::vector<BoundaryPrim> boundary;
    std::vector<LightPrim> lights;
    std::vector<LeafRef> leaves;               // visit order (kept for tests / stats)
    int light_kind = LIGHT_NONE;
    int post_media_order = 0, two_pass = 0, empty = 1, linear = 0;
    CameraParams cam;
    BuildStats stats;
};

bool flatten_scene(const Scene& s, FlatScene& out, std::string* err);
void camera_params(const Camera& c, CameraParams& out);

// bvh_build.cpp: binned SAH over boxes -> BVH2 -> collapse to 4-wide, breadth-first layout.
struct BuildPrim { float lo[3], hi[3]; int type; int ref; };   // ref = index into FlatScene::leaves
struct Bvh4Leaf { int first, count, type; };                    // range in `order_out`
void build_bvh4(const std::vector<BuildPrim>& prims, std::vector<Bvh4Node>& nodes, std::vector<int>& order_out,
                BuildStats& stats);

}  // namespace mort
